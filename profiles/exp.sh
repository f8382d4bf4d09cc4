python profiles/t_e2e.py
CLEDB_B200_NO_CHUNK=1 python profiles/t_e2e.py
