"""
miosqp_b200 -- Blackwell-native batched QP-relaxation engine behind miOSQP's API.

    from miosqp_b200 import MIOSQP, solve_many          # reference API (solver.py) + multi-instance driver
    from miosqp_b200 import engine                       # ctypes binding of the C ABI (include/bqp.h)
"""
from .constants import *          # noqa: F401,F403
from .problem_data import Data, add_bounds          # noqa: F401
from .results import Results      # noqa: F401
from .miqp import MIOSQP, solve_many, setup_many                # noqa: F401
