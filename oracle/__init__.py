"""ORACLE — test infrastructure only (see oracle/selective_scan_ref.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; sigma_b200/ never does.
"""
