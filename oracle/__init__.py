"""CPU oracle (test infrastructure only -- see oracle/oracle.py header)."""
from .oracle import *  # noqa: F401,F403
