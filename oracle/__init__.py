"""CPU oracle for the SkellySim pair-kernel hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under skellysim_b200/ does.  See oracle/oracle.c for the
reference file:line each routine follows and for the parity-pinning status.
"""
from .oracle import *  # noqa: F401,F403
