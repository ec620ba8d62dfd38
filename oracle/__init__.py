"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/gsm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (gammaspikemodel_b200/) never does.
"""
from .gsm_oracle import *  # noqa: F401,F403
