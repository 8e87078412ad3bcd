"""
BASELINE config 3 workload: finite-control-set MPC of a three-level inverter driving an induction motor, posed as
an MIQP with horizon N and re-solved every sampling instant through MIOSQP.setup / update_vectors / set_x0 / solve.

The reference builds this workload in /root/reference/examples/power_converter/ (model `power_converter.py:48-167`,
steady state `:206-260`, per-unit parameters `:263-355`, MIQP matrices `quadratic_program.py:11-136`, MPC step
`power_converter.py:421-508`, closed loop `:589-649`, switching statistics `utils.py:81-98`).  Here the same
mathematics is restated compactly: the prediction matrices come from one running power of A instead of nested
`matrix_power` loops, the closed loop is a plain function over a solver factory, and everything that is
presentation in the reference (plots, THD, Gurobi arm, LaTeX tables) is left out.  The matrices are checked against
the reference's own builder in tests/test_power_converter_workload.py (fixture: tests/golden/power_converter_model.npz).

State (12): stator currents (2), rotor fluxes (2), rotating current reference (2), previous switch position (3),
two low-pass filter states of the switching effort, constant 1 (switching-frequency target).
Input (6): switch positions u in {-1,0,1}^3 and auxiliary |du| bounds (3); all 6N inputs are integer variables.
"""
import os

import numpy as np
import scipy.linalg as sla
import scipy.sparse as spa

# settings of the reference's MPC step, power_converter.py:451-466
MPC_SETTINGS = {'eps_int_feas': 1e-02, 'max_iter_bb': 2000, 'tree_explor_rule': 1, 'branching_rule': 0,
                'verbose': False, 'print_interval': 1}
MPC_QP_SETTINGS = {'eps_abs': 1e-03, 'eps_rel': 1e-03, 'eps_prim_inf': 1e-04, 'verbose': False}

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class Drive(object):
    """Per-unit machine and inverter constants (power_converter.py:263-355)."""
    Rs, Rr = 0.0108, 0.0091                 # stator / rotor resistance
    Xls, Xlr, Xm = 0.1493, 0.1104, 2.3489   # leakage and mutual reactances
    omegar = 0.9911                         # rotor speed
    Vdc = 1.930                             # dc-link voltage
    kT = 1.2361                             # torque constant
    D = 0.6266                              # Xs Xr - Xm^2 (rounded, as the reference uses it)

    def __init__(self, Ts=25.0e-06, freq=50., k1=0.8e03, k2=0.8e03, torque=1.):
        self.Ts, self.freq, self.k1, self.k2, self.torque = Ts, freq, k1, k2, torque
        self.Tspu = Ts * 2 * np.pi * freq
        self.steps_per_period = int(1. / freq / Ts)
        self.Xs = self.Xls + self.Xm
        self.Xr = self.Xlr + self.Xm
        self.taus = (self.Xr * self.D) / (self.Rs * self.Xr ** 2 + self.Rr * self.Xm ** 2)
        self.taur = self.Xr / self.Rr
        s3 = np.sqrt(3.) / 2.
        self.clarke = 2. / 3. * np.array([[1., -.5, -.5], [0., s3, -s3]])     # abc -> alpha/beta
        self.clarke_inv = np.array([[1., 0.], [-.5, s3], [-.5, -s3]])

    def steady_state(self, torque, psi_s=1.):
        """[i_s(2), psi_r(2)] with the stator flux of magnitude psi_s on the alpha axis (power_converter.py:206-260)."""
        psi_rb = -torque / psi_s * self.D / self.Xm / self.kT
        disc = np.sqrt(self.Xm ** 2 * psi_s ** 2 - 4. * self.Xs ** 2 * psi_rb ** 2)
        psi_ra = (self.Xm * psi_s + disc) / (2. * self.Xs)
        flux = np.array([psi_s, 0., psi_ra, psi_rb])
        to_current = 1. / self.D * np.array([[self.Xr, 0., -self.Xm, 0.], [0., self.Xr, 0., -self.Xm]])
        return np.append(to_current.dot(flux), flux[2:])

    def initial_state(self):
        xs = self.steady_state(self.torque)
        return np.concatenate((xs, xs[:2], np.zeros(3), np.ones(3)))


class System(object):
    """Discrete-time extended model x+ = A x + B u, y = C x (power_converter.py:48-167)."""

    def __init__(self, drive, fsw_des=300, delta=5.5):
        d = drive
        c = d.Xm / (d.taur * d.D); w = d.omegar * d.Xm / d.D
        F = np.array([[-1. / d.taus, 0., c, w],
                      [0., -1. / d.taus, -w, c],
                      [d.Xm / d.taur, 0., -1. / d.taur, -d.omegar],
                      [0., d.Xm / d.taur, d.omegar, -1. / d.taur]])
        Gc = d.Xr / d.D * d.Vdc / 2. * np.vstack((np.eye(2), np.zeros((2, 2)))).dot(d.clarke)
        A_phys = sla.expm(F * d.Tspu)
        B_phys = -np.linalg.inv(F).dot(np.eye(4) - A_phys).dot(Gc)             # exact zero-order hold
        cs, sn = np.cos(d.Tspu), np.sin(d.Tspu)
        A_ref = np.array([[cs, -sn], [sn, cs]])                                 # rotating reference
        a1, a2 = 1. - 1. / d.k1, 1. - 1. / d.k2
        A_filt = np.array([[a1, 0.], [1. - a1, a2]])
        # averaged over the 12 semiconductor switches and normalised by the desired switching frequency
        B_filt = 1. / fsw_des * 1. / 12. * (1 - a1) / d.Ts * np.array([[1., 1., 1.], [0., 0., 0.]])
        self.A = sla.block_diag(A_phys, A_ref, np.zeros((3, 3)), A_filt, np.ones((1, 1)))
        B = np.zeros((12, 6))
        B[0:4, 0:3] = B_phys
        B[6:9, 0:3] = np.eye(3)
        B[9:11, 3:6] = B_filt
        self.B = B
        C = np.zeros((3, 12))
        C[0, 0], C[0, 4] = 1., -1.
        C[1, 1], C[1, 5] = 1., -1.
        C[2, 10], C[2, 11] = -delta, delta
        self.C = C
        self.prev_input = np.hstack((np.zeros((3, 6)), np.eye(3), np.zeros((3, 3))))   # W: state -> u_{k-1}
        self.sel_u = np.hstack((np.eye(3), np.zeros((3, 3))))                          # G: input -> switch position
        self.sel_aux = np.hstack((np.zeros((3, 3)), np.eye(3)))                        # T: input -> auxiliary bound
        self.fsw_des, self.delta = fsw_des, delta


class TailCost(object):
    """ADP tail cost x'P0x + q0'x + r0 (tail_cost.py:7-19); `name` picks one of the reference's precomputed tails."""

    def __init__(self, system, gamma=0.95, name="delta_550"):
        self.gamma = gamma
        if name is None:
            self.P0 = system.C.T.dot(system.C); self.q0 = np.zeros(system.C.shape[1]); self.r0 = 0.
        else:
            z = np.load(os.path.join(_DATA, "tail_%s.npz" % name))
            self.P0, self.q0, self.r0 = z["P0"], z["q0"], float(z["r0"])


class MpcProgram(object):
    """MIQP  min 1/2 u'Pu + q(x)'u  s.t. l <= A u <= u(x),  u integer in [-1,1]  over the stacked inputs of an
    N-step horizon (quadratic_program.py:11-136).  q(x) = 2(q_x x + q_u); the first 6N rows of the upper bound
    follow the state through `SA` (|u_k - u_{k-1}| <= auxiliary input), the last 3N rows box the auxiliary inputs."""

    def __init__(self, system, N, tail):
        A, B, C = system.A, system.B, system.C
        nx, nu = B.shape
        g = tail.gamma
        powers = [np.eye(nx)]
        for _ in range(N):
            powers.append(A.dot(powers[-1]))
        free = np.vstack(powers)                                       # x_k = A^k x_0, k = 0..N
        forced = np.zeros(((N + 1) * nx, N * nu))                      # x_k += sum_{j<k} A^(k-1-j) B u_j
        AB = [p.dot(B) for p in powers[:N]]
        for k in range(1, N + 1):
            for j in range(k):
                forced[k * nx:(k + 1) * nx, j * nu:(j + 1) * nu] = AB[k - 1 - j]
        free_N, forced_N = free[-nx:], forced[-nx:]
        stage = C.T.dot(C)
        H = sla.block_diag(*([g ** k * stage for k in range(N)] + [np.zeros((nx, nx))]))
        gN = g ** N
        self.P = spa.csc_matrix(2. * (forced.T.dot(H).dot(forced) + gN * forced_N.T.dot(tail.P0).dot(forced_N)))
        self.q_x = forced.T.dot(H.T).dot(free) + gN * forced_N.T.dot(tail.P0).dot(free_N)
        self.q_u = gN * forced_N.T.dot(tail.q0)
        W = system.prev_input
        S = np.hstack((np.vstack((np.kron(np.eye(N), W), np.kron(np.eye(N), -W))), np.zeros((6 * N, nx))))
        R = np.vstack((np.kron(np.eye(N), system.sel_u - system.sel_aux), np.kron(np.eye(N), -system.sel_u - system.sel_aux)))
        self.A = spa.csc_matrix(np.vstack((R - S.dot(forced), np.kron(np.eye(N), system.sel_aux))))
        self.SA = S.dot(free)
        self.l = np.append(-np.inf * np.ones(6 * N), -np.ones(3 * N))
        self.u = np.append(np.zeros(6 * N), np.ones(3 * N))
        self.N, self.nu = N, nu
        self.i_idx = np.arange(nu * N)
        self.i_l = -np.ones(nu * N)
        self.i_u = np.ones(nu * N)

    def vectors(self, x):
        """(q, l, u) at state x (power_converter.py:429-434)."""
        u = self.u.copy()
        u[:6 * self.N] = self.SA.dot(x)
        return 2. * (self.q_x.dot(x) + self.q_u), self.l.copy(), u


def on_transitions(u, u_prev):
    """ON transitions of the 12 semiconductor switches for one change of the three switch positions (utils.py:81-98)."""
    t = np.zeros(12)
    slot = {(0., 1.): 0, (-1., 0.): 1, (1., 0.): 2, (0., -1.): 3}
    for ph in range(3):
        k = slot.get((float(u_prev[ph]), float(u[ph])))
        if k is not None:
            t[4 * ph + k] = 1
    return t


class ClosedLoopResult(object):
    pass


def closed_loop(steps, N=10, make_solver=None, drive=None, fsw_des=300, delta=5.5, gamma=0.95, tail="delta_550",
                on_step=None, speculation=0, qp_settings=None, replay=None):
    """First `steps` sampling instants of the reference's closed loop (power_converter.py:589-649): at every instant
    build (q,l,u) from the state, warm start from the shifted previous plan, solve the MIQP, apply the first input.
    `make_solver()` returns an object with the MIOSQP interface (default: miosqp_b200.MIOSQP on the CUDA engine).
    `replay='native'` runs the B&B loop in C++ (csrc/bqp_bnb.cpp) instead of tree.py: same decisions, no interpreter time
    per node.  `qp_settings` overrides entries of the reference's OSQP settings.  `speculation` > 0 lets every launch also solve up to that many nodes ahead of the replay (tree.py `speculate`):
    same answers, far fewer host round trips on this workload's deep, narrow trees.
    Returns the trajectories and the per-step B&B statistics."""
    if make_solver is None:
        from .miqp import MIOSQP as make_solver
    drive = drive or Drive()
    system = System(drive, fsw_des, delta)
    prog = MpcProgram(system, N, TailCost(system, gamma, tail))
    nx, nu = system.B.shape
    X = np.zeros((nx, steps + 1)); U = np.zeros((nu, steps)); Y = np.zeros((3, steps))
    X[:, 0] = drive.initial_state()
    plan = np.zeros(nu * N)
    res = ClosedLoopResult()
    res.obj = np.zeros(steps); res.run_time = np.zeros(steps); res.osqp_solve_time = np.zeros(steps)
    res.osqp_iter_avg = np.zeros(steps); res.nodes = np.zeros(steps, dtype=int); res.admm_iters = np.zeros(steps, dtype=int)
    res.status = []
    solver = None
    for k in range(steps):
        q, l, u = prog.vectors(X[:, k])
        if solver is None:
            solver = make_solver()
            solver.setup(prog.P, q, prog.A, l, u, prog.i_idx, prog.i_l, prog.i_u,
                         dict(MPC_SETTINGS, speculation=speculation, replay=replay), dict(MPC_QP_SETTINGS, **(qp_settings or {})))
        else:
            solver.update_vectors(q, l, u)
        solver.set_x0(plan)
        r = solver.solve()
        # the reference stops in a debugger on anything but MI_SOLVED (power_converter.py:493); a workload run keeps
        # going on the incumbent when the node limit max_iter_bb is hit and counts those steps
        if r.status not in ('Solved', 'Max-iter feasible'):
            raise RuntimeError("MPC step %d: MIQP status %r" % (k, r.status))
        res.status.append(r.status)
        plan = np.asarray(r.x, dtype=float)
        U[:, k] = plan[:nu]
        res.obj[k], res.run_time[k], res.osqp_solve_time[k], res.osqp_iter_avg[k] = r.upper_glob, r.run_time, r.osqp_solve_time, r.osqp_iter_avg
        res.nodes[k] = solver.work.iter_num - 1; res.admm_iters[k] = solver.work.osqp_iter
        if on_step is not None:
            on_step(k, solver, r)
        X[:, k + 1] = system.A.dot(X[:, k]) + system.B.dot(U[:, k])
        Y[:, k] = system.C.dot(X[:, k])
        plan = np.append(plan[nu:], plan[-nu:])                       # shifted plan = next warm start
    res.X, res.U, res.Y, res.program, res.system, res.drive, res.solver = X, U, Y, prog, system, drive, solver
    res.phase_currents = drive.clarke_inv.dot(X[0:2, :steps])
    res.phase_references = drive.clarke_inv.dot(X[4:6, :steps])
    res.torque = drive.kT * (drive.Xm / drive.Xr) * (X[2, :steps] * X[1, :steps] - X[3, :steps] * X[0, :steps])
    sw = np.zeros(12)
    for k in range(1, steps):
        sw += on_transitions(U[:3, k], U[:3, k - 1])
    res.switching_frequency = float(np.mean(sw / (steps * drive.Ts))) if steps > 1 else 0.
    return res
