#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
BQP_LIB_SUFFIX=_smalldbg BQP_BUILD_DEFS="-DBQP_SMALL_DEBUG" timeout 300 python - <<'PY' 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/s45_small_check_timers.log
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
sys.argv = ["x"]
# the config-3 problem, 8 leaves, 200 iterations with a check every 25 (convergence off)
import scipy.sparse as spa
from miosqp_b200 import engine, problems, power_converter as pc
drive = pc.Drive(); system = pc.System(drive, 300, 5.5)
prog = pc.MpcProgram(system, 10, pc.TailCost(system, 0.95, "delta_550"))
q, l, u = prog.vectors(drive.initial_state())
A = spa.vstack([prog.A, spa.identity(60, format="csc")[prog.i_idx, :]]).tocsc()
l = np.append(l, prog.i_l); u = np.append(u, prog.i_u)
e = engine.BatchedQP().setup(prog.P, q, A, l, u, i_idx=np.asarray(prog.i_idx), eps_abs=1e-12, eps_rel=1e-12, eps_prim_inf=1e-12, eps_dual_inf=1e-12, max_iter=200, check_termination=25)
ls, us = problems.branched_nodes(l, u, 60, 8, np.random.default_rng(0))
for _ in range(2):
    r = e.solve_batch(ls, us, np.zeros((8, 60)), np.zeros((8, 150)))
print(engine.last_timing())
PY
