"""CPU oracle for the SVDSS hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (svdss_b200/) never does.  PARITY UNPINNED: see oracle/sfs_oracle.c header.
"""
from .binding import *  # noqa: F401,F403
