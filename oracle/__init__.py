"""CPU oracle of the CLEDB hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke()
and the cpu_baseline / --impl reference legs of bench.py, never by the product package."""
