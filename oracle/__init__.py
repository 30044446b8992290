"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the batched-IOD hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (outfit_b200) never does.
"""
