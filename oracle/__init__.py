"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/cv_ransac_oracle.c for the full header).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package; the
product package (code-reproduction-ransac_b200/) never does.
"""
from .cvoracle import *  # noqa: F401,F403
