python __graft_entry__.py smoke 2>&1 | tail -3
python profiles/time_invproc.py 256 2>&1 | grep cledb_invproc
bash profiles/quick.sh r01h
