"""oracle/ -- CPU restatement of the reference algorithm.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import from here, and only as the checker (never as the thing measured or shipped).
"""
