"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

Nothing under openmc_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
