python -m pytest tests -m gpu -x -q -k "randomised" 2>&1 | tail -8
