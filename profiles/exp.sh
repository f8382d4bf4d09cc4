python -m pytest tests -m gpu -x -q -k "pruning" 2>&1 | tail -5
