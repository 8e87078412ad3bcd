"""MIQP status strings, byte-identical to /root/reference/miosqp/constants.py:2-7 (including the 'Unolved' typo)."""
MI_UNSOLVED = 'Unolved'
MI_SOLVED = 'Solved'
MI_PRIMAL_INFEASIBLE = 'Primal Infeasible'
MI_DUAL_INFEASIBLE = 'Dual Infeasible'
MI_MAX_ITER_FEASIBLE = 'Max-iter feasible'
MI_MAX_ITER_UNSOLVED = 'Max-iter unsolved'
