"""CPU oracle for the VQ bottleneck hot path (TEST INFRASTRUCTURE, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product path (vq_seg_b200/) never does.
"""
from .vq_oracle import *  # noqa: F401,F403
