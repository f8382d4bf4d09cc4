"""Minimal stand-in so the reference's CLEDB_PREPINV imports in a container without astropy.

Test infrastructure only (used by oracle/run_reference.py).  Only the Cryo-NIRSP header branch of the
reference touches astropy; nothing on the hot path does.
"""
