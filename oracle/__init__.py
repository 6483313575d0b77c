"""CPU oracle for the OH path — TEST INFRASTRUCTURE ONLY (see oracle/qc_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package.  PARITY UNPINNED: the reference has no golden vectors and its arithmetic
lives in the absent libxgboost 1.6.0; anchoring is by known-answer boosters (tests/).
"""
