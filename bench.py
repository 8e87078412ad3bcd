#!/usr/bin/env python
"""
bench.py -- QP-relaxations/sec of the batched ADMM hot path (BASELINE.json metric).

Workload (config "cfg2"): random_miqp n=500 m=1000 |i_idx|=50 density 0.7, 100 instances drawn by the
reference generator (/root/reference/examples/random_miqp/run_example.py:71-83, np.random.seed(1)),
each contributing the 8 leaves of its depth-3 B&B subtree on the three most fractional integer
variables of its root relaxation, warm-started from the root's (x, y) exactly as
Workspace.add_left/add_right create children (/root/reference/miosqp/workspace.py:157-203).
One STEP = one pass of the hot path over that frontier batch: every leaf's full OSQP ADMM solve
(Node.solve, /root/reference/miosqp/node.py:96-143), all leaves concurrently in one kernel launch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--instances I] [--mode both|frontier|bnb]

The same line carries, under "bnb", the reference's OWN use of the path: branch-and-bound TO COMPLETION over the same 100
instances with the reference's settings (/root/reference/examples/random_miqp/run_example.py:98-116; loop solver.py:85-123),
every relaxation on the engine, all trees sharing one rolling engine session (bqp_bnb_solve_rolling): consumed QP relaxations
per second by wall clock (host replay, H2D and D2H inside), next to the same B&B on the CPU oracle (bounded sample).

value   : leaves solved per second, inputs resident in HBM, CUDA-event time of the launches.
e2e     : the same through the public host-buffer call (pack + H2D + launch + D2H inside the timed region).
roofline: algorithmic bytes / kernel time against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference: the CPU oracle (oracle/, a restatement of OSQP; the reference's own
`osqp` dependency is not installable here) on all host threads over a bounded sample of the same leaves.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QP_SETTINGS = dict(eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4)   # run_example.py:113-116
N_VAR, M_CON, P_INT, DENSITY = 500, 1000, 50, 0.7
LEAF_DEPTH = 3


def make_instances(count, seed):
    from miosqp_b200 import problems
    return [problems.extend(pr) for pr in problems.random_miqp(N_VAR, M_CON, P_INT, DENSITY, seed=seed, count=count)]


def make_leaves(inst, x_root, y_root):
    """8 leaves of the depth-3 subtree below the root on its 3 most fractional integer variables."""
    P, q, A, l, u, i_idx = inst
    m = l.shape[0]
    n_int = len(i_idx)
    xi = x_root[i_idx]
    frac = np.abs(xi - np.round(xi))
    order = np.argsort(-frac, kind="stable")[:LEAF_DEPTH]
    L, U = [], []
    for code in range(1 << LEAF_DEPTH):
        ll = l.copy(); uu = u.copy()
        for k, pos in enumerate(order):
            row = m - n_int + pos
            if (code >> k) & 1:
                ll[row] = np.ceil(xi[pos])        # right child, workspace.py:189-190
            else:
                uu[row] = np.floor(xi[pos])       # left child, workspace.py:165-166
        if np.any(ll > uu):                       # (integral root entry) keep the node valid
            continue
        L.append(ll); U.append(uu)
    return L, U


def algorithmic_bytes_per_node_iter(inst, T, check_every=25):
    """8(2n+6m) + (2*12*nnzL + 16 N)/T + checks/25/T."""
    P, q, A, l, u, i_idx = inst
    n, m = A.shape[1], A.shape[0]
    import scipy.sparse as spa
    nnzA = A.nnz
    nnz_triuP = spa.triu(P).nnz
    nnzL = nnzA + n * (n - 1) // 2
    N = n + m
    return 8.0 * (2 * n + 6 * m) + (24.0 * nnzL + 16.0 * N) / T + 12.0 * (2 * nnzA + nnz_triuP) / check_every / T


class ClockSampler(object):
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [v.strip() for v in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_oracle_sample(insts_nodes, threads, repeats=1):
    """insts_nodes: list of (instance tuple, l, u, x0, y0).  Returns (seconds per pass, nodes, node_iters)."""
    from oracle import oracle
    solvers = {}
    S, L, U, X0, Y0 = [], [], [], [], []
    for inst, l, u, x0, y0 in insts_nodes:
        key = id(inst)
        if key not in solvers:
            o = oracle.OSQP()
            o.setup(inst[0], inst[1], inst[2], inst[3], inst[4], **QP_SETTINGS)
            solvers[key] = o
        S.append(solvers[key]); L.append(l); U.append(u); X0.append(x0); Y0.append(y0)
    best = None
    iters = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, _, st, it, _ = oracle.solve_multi(S, L, U, X0, Y0, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        iters = int(np.sum(it))
    return best, len(S), iters


def kernel_source_sha256(rel):
    import hashlib
    try:
        with open(os.path.join(ROOT, rel), "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()
    except OSError:
        return None


def build_report(args, world, workload, B, B_all, iters_all, node_iters, dev_s, dev_s_max, e2e_s_max, wall_max,
                 tm, tm2, iters, status, factor_mb, inst0, clocks, t_setup, root_iters):
    """The JSON line of the bench contract (kept separate from the GPU code so that it is unit-tested on CPU)."""
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    T = 1 << LEAF_DEPTH
    bytes_per_ni = algorithmic_bytes_per_node_iter(inst0, T)
    step_s = dev_s / args.steps
    kernel = {0: "admm_tile_kernel<%d>", 1: "admm_stream_kernel<%d>", 2: "admm_panel_kernel (%d nodes per tile)",
              3: "admm_rows_kernel (%d nodes per tile)"}[int(tm.get("kernel", 1))] % tm["tile_nodes"]
    traffic, traffic_src, traffic_scope = None, None, None
    try:   # dram__bytes_read+write of one captured launch of this kernel (ncu --set full), valid only for the kernel source it
        # was captured from: the capture records the sha256 of the kernel's .cu file and a different source nulls it
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tr = json.load(f)
        if tr["kernel"] == kernel and tr.get("kernel_source_sha256") == kernel_source_sha256(tr.get("kernel_source", "")):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            traffic_src = tr["source"]
            traffic_scope = tr.get("scope")
    except (OSError, KeyError, ValueError):
        pass
    achieved = bytes_per_ni * node_iters / step_s / 1e9
    return {
        "metric": "QP-relaxations/sec", "value": B_all * args.steps / dev_s_max, "unit": "QP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "leaves_per_step": int(B_all), "qp_settings": QP_SETTINGS,
                   "parallelism": "instances sharded across %d GPU(s), no data-path collective" % world,
                   "l2": "inputs larger than L2 (%.0f MB of factors per GPU streamed every ADMM iteration)" % factor_mb,
                   "tile_nodes": tm["tile_nodes"], "threads_per_cta": tm["threads"], "tiles_first_launch": tm["tiles"],
                   "smem_bytes_per_cta": tm["smem_bytes"], "ring_slots": int(tm.get("ring_slots", 0))},
        "admm_node_iters_per_s": iters_all * args.steps / dev_s_max,
        "admm_iters_per_leaf": node_iters / float(B),
        "admm_iters_max": int(np.max(iters)), "admm_iters_median": float(np.median(iters)),
        "status_counts": {str(k): int(v) for k, v in zip(*np.unique(status, return_counts=True))},
        "e2e": {"value": B_all * args.steps / e2e_s_max, "unit": "QP/s",
                "h2d_bytes_per_step": int(tm2["h2d_bytes"]), "d2h_bytes_per_step": int(tm2["d2h_bytes"])},
        "gpu_launches": args.steps * int(tm["launches"]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "traffic": traffic, "traffic_source": traffic_src, "traffic_scope": traffic_scope, "kernel": kernel,
                     "algorithmic_bytes_per_node_iter": bytes_per_ni, "node_iters_per_step": node_iters,
                     "launches_per_step": int(tm["launches"]), "step_kernel_ms": 1e3 * step_s,
                     "note": "one step = %d launches of the same kernel (rounds of ADMM iterations, re-tiled in between); "
                             "achieved and streamed bytes are per step; traffic is the DRAM traffic of the launch named in traffic_scope" % int(tm["launches"]),
                     "streamed_bytes_per_step": int(tm["stream_bytes"]),
                     "streamed_gbs": tm["stream_bytes"] / step_s / 1e9},
        "setup_s": t_setup, "root_iters": root_iters, "wall_s_resident_loop": wall_max,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=100)
    ap.add_argument("--tile-nodes", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sort", action="store_true")
    ap.add_argument("--parallel-setup", action="store_true", help="(default now) bqp_setup_many: factorise the instances on all host threads")
    ap.add_argument("--serial-setup", action="store_true", help="one bqp_setup per instance")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "mpc", "cfg4"],
                    help="cfg2 (default): BASELINE config 2, the headline; mpc: BASELINE config 3, the power-converter MPC closed loop "
                         "(horizon 10, --mpc-steps sampling instants, warm-started), ms per MPC step and consumed QP/s, CPU oracle beside it; "
                         "cfg4: BASELINE config 4, one n=2000 m=4000 |i_idx|=200 MIQP (5 %% dense), B&B cut at --cfg4-nodes nodes")
    ap.add_argument("--cfg4-nodes", type=int, default=40)
    ap.add_argument("--mpc-steps", type=int, default=1000)
    ap.add_argument("--mpc-cpu-steps", type=int, default=40)
    ap.add_argument("--adaptive-rho-interval", type=int, default=0,
                    help="> 0: the whole run (GPU and CPU arms) under osqp's adaptive_rho with this fixed interval instead of the "
                         "fixed-rho contract")
    ap.add_argument("--no-adaptive-extra", action="store_true",
                    help="skip the 'adaptive_rho' object of the default line (a second, shorter run of this script under adaptive rho)")
    ap.add_argument("--mode", default="both", choices=["both", "frontier", "bnb"],
                    help="frontier: the batched-frontier step (value, e2e, roofline); bnb: B&B to completion; both (default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = host_cores()
    workload = "random_miqp n=%d m=%d |i_idx|=%d density=%.1f, %d instances/GPU x %d leaves (depth-%d subtree, warm-started from the root)" % (
        N_VAR, M_CON, P_INT, DENSITY, args.instances, 1 << LEAF_DEPTH, LEAF_DEPTH)
    if args.adaptive_rho_interval > 0:
        # osqp's adaptive_rho (what the reference gets from osqp's defaults, workspace.py:63-68) with a FIXED interval -- the
        # reproducible variant; GPU and CPU arms alike
        QP_SETTINGS.update(adaptive_rho=True, adaptive_rho_interval=args.adaptive_rho_interval)
        workload += ", osqp adaptive_rho with a fixed interval of %d iterations" % args.adaptive_rho_interval

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, cores, workload)
    if args.workload == "mpc":
        return mpc_workload(args, cores) if rank == 0 else 0
    if args.workload == "cfg4":
        return cfg4_workload(args, cores) if rank == 0 else 0

    import torch
    import torch.distributed as dist
    from miosqp_b200 import engine
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- setup (untimed): instances, device-resident factors, root relaxations, leaf batch
    t_setup = time.perf_counter()
    from miosqp_b200 import problems as _problems
    raw = _problems.random_miqp(N_VAR, M_CON, P_INT, DENSITY, seed=1 + rank, count=args.instances)
    insts = [_problems.extend(pr) for pr in raw]
    if not args.serial_setup:   # host halves of the 100 factorisations on all host threads (untimed either way)
        qps = engine.setup_many(insts, device=local_rank, **QP_SETTINGS)
    else:
        qps = [engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, device=local_rank, **QP_SETTINGS)
               for (P, q, A, l, u, i_idx) in insts]
    n, m = N_VAR, M_CON + P_INT
    xs, ys, sc = engine.solve_multi(qps, [i[3] for i in insts], [i[4] for i in insts],
                                    [np.zeros(n)] * len(insts), [np.zeros(m)] * len(insts))
    root_iters = int(np.sum(sc.iters))
    Q, L, U, X0, Y0, owner = [], [], [], [], [], []
    # longest-first submission: a leaf's ADMM iteration count correlates with its parent's, and tiles are scheduled in
    # submission order, so the instances whose root needed most iterations go first (shorter tail when tiles > SMs)
    order = np.argsort(-sc.iters, kind="stable") if not args.no_sort else np.arange(len(insts))
    for k in order:
        inst = insts[k]
        x_root = np.nan_to_num(xs[k]); y_root = np.nan_to_num(ys[k])
        ll, uu = make_leaves(inst, x_root, y_root)
        for a, b in zip(ll, uu):
            Q.append(qps[k]); L.append(a); U.append(b); X0.append(x_root); Y0.append(y_root); owner.append(k)
    B = len(Q)
    engine.set_tuning(args.tile_nodes, args.threads)
    t_setup = time.perf_counter() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident measurement: inputs stay in HBM, one launch per step
    rb = engine.ResidentBatch(Q, L, U, X0, Y0)
    for _ in range(args.warmup):
        rb.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    kernel_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rb.run()                                   # blocking; CUDA events on the engine's stream
        kernel_ms.append(engine.last_timing()["kernel_ms"])
    barrier()
    wall_resident = time.perf_counter() - t0
    clocks = sampler.stop()
    xs, ys, sc = rb.download()
    tm = engine.last_timing()
    dev_s = float(np.sum(kernel_ms)) / 1e3
    node_iters = int(tm["node_iters"])

    # ---- end to end through the public host-buffer API
    for _ in range(min(args.warmup, 2)):
        engine.solve_multi(Q, L, U, X0, Y0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        xs2, ys2, sc2 = engine.solve_multi(Q, L, U, X0, Y0)
    barrier()
    e2e_s = time.perf_counter() - t0
    tm2 = engine.last_timing()

    # ---- max over ranks
    times = torch.tensor([dev_s, e2e_s, wall_resident], dtype=torch.float64, device="cuda")
    counts = torch.tensor([B, node_iters], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    dev_s_max, e2e_s_max, wall_max = [float(v) for v in times.tolist()]
    B_all, iters_all = [float(v) for v in counts.tolist()]

    if rank == 0:
        out = build_report(args, world, workload, B, B_all, iters_all, node_iters, dev_s, dev_s_max, e2e_s_max, wall_max,
                           tm, tm2, np.asarray(sc.iters), np.asarray(sc.status),
                           sum(qp.dims()["factor_bytes"] for qp in qps) / 1e6, insts[0], clocks, t_setup, root_iters)
        if not args.no_cpu_baseline:
            # every leaf of `cores` instances spread uniformly over the (longest-first) submission order
            pos = sorted(set(int(v) for v in np.linspace(0, len(order) - 1, min(len(order), cores)).round()))
            chosen = set(int(order[k]) for k in pos)
            sample = [(insts[owner[b]], L[b], U[b], X0[b], Y0[b]) for b in range(B) if owner[b] in chosen]
            secs, nodes, its = run_oracle_sample(sample, cores)
            out["cpu_baseline"] = {"value": nodes / secs, "unit": "QP/s", "cores": cores, "kind": "port",
                                   "sample": "all %d leaves of %d instances spread uniformly over the step's (longest-first) batch, CPU "
                                             "oracle (oracle/osqp_oracle.c; parity with PyPI osqp unpinned), %d threads, %.1f s" % (
                                                 nodes, len(chosen), cores, secs),
                                   "admm_node_iters_per_s": its / secs}
            try:    # the reference's own execution model: one node at a time on one core (SURVEY 8d asks for both figures)
                s1, n1, i1 = run_oracle_sample(sample[::max(1, len(sample) // 8)][:8], 1)
                out["cpu_baseline"]["single_core"] = {"value": n1 / s1, "unit": "QP/s", "cores": 1,
                                                      "sample": "%d of those leaves, one thread, %.1f s" % (n1, s1),
                                                      "admm_node_iters_per_s": i1 / s1}
            except Exception as e:      # never lose the bench line over an extra figure
                out["cpu_baseline"]["single_core"] = {"error": repr(e)}
    bnb = None
    if args.mode in ("both", "bnb"):
        rb = None                      # the resident frontier batch is not needed any more
        bnb = bnb_section(args, raw, qps, rank, world, cores, torch, dist)
    if rank == 0:
        if bnb is not None:
            out["bnb"] = bnb
        if world == 1 and args.adaptive_rho_interval == 0 and not args.no_adaptive_extra and args.workload == "cfg2":
            out["adaptive_rho"] = adaptive_extra(args)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def adaptive_extra(args):
    """The same workload under osqp's adaptive_rho (fixed interval 50): a second, shorter run of this script, so that the default
    line shows both contracts.  The headline `value` stays on the fixed-rho contract the reference arm is timed on."""
    cmd = [sys.executable, os.path.abspath(__file__), "--adaptive-rho-interval", "50", "--steps", "3", "--warmup", "3",
           "--instances", str(args.instances), "--mode", args.mode]
    if args.no_cpu_baseline:
        cmd.append("--no-cpu-baseline")
    try:
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        line = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")][-1]
        d = json.loads(line)
        keep = {k: d.get(k) for k in ("value", "unit", "ms_per_step", "admm_node_iters_per_s", "admm_iters_per_leaf", "admm_iters_max",
                                      "status_counts", "e2e", "gpu_launches", "cpu_baseline")}
        keep["workload"] = d["config"]["workload"]
        keep["kernel"] = d["roofline"]["kernel"]
        keep["launches_per_step"] = d["roofline"]["launches_per_step"]
        keep["streamed_gbs"] = d["roofline"]["streamed_gbs"]
        if "bnb" in d:
            keep["bnb"] = {k: d["bnb"].get(k) for k in ("value", "unit", "rolling", "cpu", "decisions_identical_to_reference_golden")}
        return keep
    except Exception as e:          # never lose the bench line over the extra figure
        return {"error": repr(e)}


BNB_WORKLOAD = ("random_miqp n=%d m=%d |i_idx|=%d density=%.1f, %%d instances/GPU, branch-and-bound TO COMPLETION with the reference's "
                "settings (run_example.py:98-116), every relaxation on the engine" % (N_VAR, M_CON, P_INT, DENSITY))


def oracle_many_fn(oracles, threads):
    """bqp_solve_many_fn (include/bqp.h) backed by the CPU oracle: the native B&B replay then runs the reference's loop on the
    host cores -- the CPU arm of the B&B numbers.  Only this function touches oracle/ in the B&B section."""
    from oracle import oracle
    from miosqp_b200 import engine

    def fn(ctx, B, owner, l, u, x0, y0, x, y, status, iters):
        try:
            os_ = [oracles[int(owner[b])] for b in range(B)]
            arr = lambda pp, b, size: np.ctypeslib.as_array(pp[b], shape=(size,))
            Ls = [np.array(arr(l, b, os_[b].m)) for b in range(B)]; Us = [np.array(arr(u, b, os_[b].m)) for b in range(B)]
            X0 = [np.array(arr(x0, b, os_[b].n)) for b in range(B)]; Y0 = [np.array(arr(y0, b, os_[b].m)) for b in range(B)]
            xs, ys, st, it, _ = oracle.solve_multi(os_, Ls, Us, X0, Y0, threads=threads)
            for b in range(B):
                arr(x, b, os_[b].n)[:] = xs[b]; arr(y, b, os_[b].m)[:] = ys[b]; status[b] = int(st[b]); iters[b] = int(it[b])
            return 0
        except Exception:       # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return -1
    return engine.SOLVE_MANY_FN(fn)


def bnb_cpu(raw, count, threads, max_iter_bb=None):
    """The same B&B on the CPU oracle for the first `count` instances: one tree per host thread, each running the reference's
    sequential loop (native replay + oracle solves, one relaxation at a time) -- trees are independent, so this is the
    embarrassingly parallel CPU execution with no thread waiting for another.  Returns (seconds, consumed nodes, iters, outs)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    from miosqp_b200 import engine, problems
    from miosqp_b200.problem_data import Data
    datas, fns = [], []
    for pr in raw[:count]:
        d = Data(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'])
        o = oracle.OSQP(); o.setup(d.P, d.q, d.A, d.l, d.u, **QP_SETTINGS)
        datas.append(d); fns.append(oracle_many_fn([o], 1))
    st = dict(problems.RANDOM_MIQP_SETTINGS)
    if max_iter_bb:
        st['max_iter_bb'] = max_iter_bb

    def one(k):     # ctypes releases the GIL inside the native replay and inside the oracle's solve
        return engine.bnb_solve_many([None], [datas[k]], [st], [QP_SETTINGS['eps_abs']], [None], [np.inf], many_fn=fns[k])[0]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        outs = list(ex.map(one, range(count)))
    secs = time.perf_counter() - t0
    return secs, sum(r["iter_num"] - 1 for _, r, _ in outs), sum(int(r["osqp_iter"]) for _, r, _ in outs), outs


def bnb_section(args, raw, qps, rank, world, cores, torch, dist):
    """B&B to completion over this rank's instances on the engine (rolling session), max over ranks; CPU arm on rank 0."""
    import miosqp_b200
    from miosqp_b200 import engine, problems, miqp
    solvers = miqp.wrap_many(raw, qps, dict(problems.RANDOM_MIQP_SETTINGS, replay='native'), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    runs = {}
    for driver in ("rolling", "lockstep"):
        for s in solvers:
            s.work.reset(); s.work.first_run = 0
            s.work.batches = s.work.batched_nodes = s.work.spec_nodes = s.work.spec_hits = 0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = miosqp_b200.solve_many(solvers, rolling=(driver == "rolling"))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        works = [s.work for s in solvers]
        tm = engine.last_timing()
        consumed = sum(w.iter_num - 1 for w in works)
        iters = sum(int(w.osqp_iter) for w in works)
        solved = sum(int(w.batched_nodes) for w in works)
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        c = torch.tensor([consumed, iters, solved], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
        wall_max = float(t.item()); consumed_all, iters_all, solved_all = [float(v) for v in c.tolist()]
        runs[driver] = {"wall_s": wall_max, "qp_consumed": int(consumed_all), "qp_solved": int(solved_all),
                        "qp_per_s": consumed_all / wall_max, "admm_node_iters_per_s": iters_all / wall_max,
                        "status": {st: [r.status for r in res].count(st) for st in sorted(set(r.status for r in res))}}
        if driver == "rolling":
            runs[driver].update({"launches": int(tm["launches"]), "kernel_ms": tm["kernel_ms"],
                                 "mean_tiles_per_launch": tm["tile_iters"] / max(1.0, 100.0 * tm["launches"]),
                                 "mean_live_nodes_per_tile_iteration": iters / max(1.0, float(tm["tile_iters"])),
                                 "h2d_bytes": int(tm["h2d_bytes"]), "d2h_bytes": int(tm["d2h_bytes"])})
            sig = [(r.status, float(r.upper_glob), w.iter_num, int(w.osqp_iter), list(map(tuple, w.decisions))) for r, w in zip(res, works)]
        else:
            # the rolling session runs sparse rounds on clusters of 4 (other summation order): decisions, node and iteration
            # counts must be identical, incumbents agree to 1e-9
            sig2 = [(r.status, float(r.upper_glob), w.iter_num, int(w.osqp_iter), list(map(tuple, w.decisions))) for r, w in zip(res, works)]
            runs[driver]["same_decisions_and_counts_as_rolling"] = bool(
                [(a[0], a[2], a[3], a[4]) for a in sig] == [(a[0], a[2], a[3], a[4]) for a in sig2])
            runs[driver]["max_rel_incumbent_difference_to_rolling"] = float(max(abs(a[1] - b[1]) / (1 + abs(b[1])) for a, b in zip(sig, sig2)))
    if rank != 0:
        return None
    gold = None
    gname = "bnb_cfg2_adaptive%d.json" % QP_SETTINGS["adaptive_rho_interval"] if QP_SETTINGS.get("adaptive_rho") else "bnb_cfg2.json"
    gp = os.path.join(ROOT, "tests", "golden", gname)
    if os.path.exists(gp):      # rank 0 draws seed 1: its first instances are the golden ones (unmodified reference on the oracle)
        g = json.load(open(gp))
        gold = True
        for i in range(min(len(sig), len(g))):
            r = g["cfg2_inst%d" % i]["result"]
            gold = gold and (sig[i][4] == [tuple(d) for d in r["decisions"]] and sig[i][0] == r["status"] and sig[i][2] == r["iter_num"]
                             and sig[i][3] == r["osqp_iter"] and abs(sig[i][1] - r["upper_glob"]) <= 1e-9 * (1 + abs(r["upper_glob"])))
    ro = runs["rolling"]
    out = {"workload": BNB_WORKLOAD % args.instances, "driver": "rolling engine session shared by all trees (bqp_bnb_solve_rolling), native replay",
           "metric": "QP-relaxations/sec (consumed by the B&B, wall clock incl. host replay and all copies)",
           "value": ro["qp_per_s"], "unit": "QP/s",
           "e2e": {"value": ro["qp_per_s"], "unit": "QP/s", "h2d_bytes_per_run": ro.get("h2d_bytes"), "d2h_bytes_per_run": ro.get("d2h_bytes")},
           "rolling": ro, "lockstep": runs["lockstep"],
           "decisions_identical_to_reference_golden": gold,
           "golden": "tests/golden/%s: first 3 instances, UNMODIFIED reference package on the CPU oracle (make_bnb_golden.py --cfg2)" % gname}
    if not args.no_cpu_baseline:
        try:
            cnt = max(1, min(args.instances, cores))
            secs, nodes, its, _ = bnb_cpu(raw, cnt, cores, max_iter_bb=int(os.environ.get("BENCH_BNB_CPU_NODES", "12")))
            out["cpu"] = {"value": nodes / secs, "unit": "QP/s", "cores": cores, "kind": "port",
                          "sample": "first %d instances, B&B cut at %s nodes each (the reference's max_iter_bb), one tree per host thread running "
                                    "the reference's sequential loop, %d threads, CPU oracle, %.1f s" % (cnt, os.environ.get("BENCH_BNB_CPU_NODES", "12"), cores, secs),
                          "admm_node_iters_per_s": its / secs,
                          # the cut keeps the top of every tree, whose relaxations need more iterations than the average node of a
                          # full run: QP/s at the full run's mean iteration count per node, the figure to compare `value` with
                          "qp_per_s_at_full_run_mean_iters": (its / secs) / (ro["admm_node_iters_per_s"] / ro["qp_per_s"])}
        except Exception as e:      # never lose the bench line over the CPU arm
            out["cpu"] = {"error": repr(e)}
    return out


def mpc_workload(args, cores):
    """BASELINE config 3: examples/power_converter closed loop (power_converter.py:589-649), horizon N = 10, first
    --mpc-steps sampling instants, every MIQP warm-started from the shifted previous plan (set_x0) and solved by the native
    B&B replay with look-ahead on the engine.  CPU arm: the same loop, same replay, relaxations on the CPU oracle one at a
    time (the reference's execution model), first --mpc-cpu-steps instants (the start-up transient: the heaviest steps)."""
    import torch
    from miosqp_b200 import power_converter as pc, engine
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU path")
    pc.closed_loop(2, N=10, speculation=0, replay='native').solver.work.solver.free()        # CUDA context, first-launch costs
    out = {"metric": "ms per MPC step (closed loop, horizon 10)", "unit": "ms", "higher_is_better": False, "n_gpus": 1, "dtype": "f64",
           "data": "synthetic", "config": {"workload": "power_converter MPC N=10, %d sampling instants, warm-started (BASELINE config 3)" % args.mpc_steps}}
    runs = {}
    for K in (128, 32):
        t0 = time.perf_counter()
        r = pc.closed_loop(args.mpc_steps if K == 128 else min(args.mpc_steps, args.mpc_cpu_steps), N=10, speculation=K, replay='native')
        wall = time.perf_counter() - t0
        w = r.solver.work
        steps = len(r.status)
        runs[K] = {"steps": steps, "ms_per_mpc_step": 1e3 * wall / steps, "qp_consumed": int(r.nodes.sum()), "qp_solved": int(w.batched_nodes),
                   "qp_per_s_consumed": float(r.nodes.sum()) / wall, "launches_per_step": w.batches / float(steps),
                   "admm_iters": int(r.admm_iters.sum()), "node_limit_steps": sum(s != 'Solved' for s in r.status),
                   "mean_torque": float(r.torque.mean()), "switching_frequency_hz": r.switching_frequency}
        if K == 128:
            Ugpu = r.U.copy(); nodes_gpu = r.nodes.copy()
        w.solver.free()
    out["value"] = runs[128]["ms_per_mpc_step"]
    out["lookahead_128"] = runs[128]; out["lookahead_32_first_steps"] = runs[32]
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["FAKE_ENGINE_THREADS"] = "1"
        import fake_engine                      # oracle-backed stand-in for the engine (test infrastructure): the CPU arm
        real = (engine.BatchedQP, engine.solve_multi)
        engine.BatchedQP, engine.solve_multi = fake_engine.FakeBatchedQP, fake_engine.solve_multi
        try:
            S = min(args.mpc_steps, args.mpc_cpu_steps)
            t0 = time.perf_counter()
            rc = pc.closed_loop(S, N=10, speculation=0, replay='native')
            wall = time.perf_counter() - t0
            t0 = time.perf_counter()
            rg = None
        finally:
            engine.BatchedQP, engine.solve_multi = real
        same = bool(np.array_equal(rc.U, Ugpu[:, :S]) and np.array_equal(rc.nodes, nodes_gpu[:S]))
        out["cpu_baseline"] = {"value": 1e3 * wall / S, "unit": "ms per MPC step", "cores": 1, "kind": "port",
                               "sample": "first %d sampling instants, native replay + CPU oracle, one relaxation at a time (the reference's "
                                         "execution model), %.1f s" % (S, wall),
                               "qp_per_s_consumed": float(rc.nodes.sum()) / wall, "qp_consumed": int(rc.nodes.sum()),
                               "same_inputs_and_node_counts_as_gpu": same}
        out["gpu_over_cpu_on_the_same_steps"] = out["cpu_baseline"]["value"] / runs[32]["ms_per_mpc_step"]
    print(json.dumps(out))
    return 0


def cfg4_workload(args, cores):
    """BASELINE config 4 on ONE GPU: random_miqp n=2000 m=4000 |i_idx|=200 density 0.05, the reference's B&B cut at
    --cfg4-nodes nodes (its own max_iter_bb), native replay.  The tree offers one or two leaves per step = ONE tile, solved by
    the whole-GPU kernel (bqp_grid.cu: every SM on that tile; BQP_GRID=0 falls back to the one-CTA streamed kernel): the figure
    that matters is the time per ADMM iteration of that tile, reported with the bytes one iteration moves (L2-resident) and the
    SURVEY 8d algorithmic bytes.  CPU arm: the same B&B on the oracle."""
    import torch
    import scipy.sparse as spa
    from miosqp_b200 import engine, problems, miqp
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU path")
    n, m, p, dens = 2000, 4000, 200, 0.05
    raw = problems.random_miqp(n, m, p, dens, seed=1)
    st = dict(problems.RANDOM_MIQP_SETTINGS, replay='native', max_iter_bb=args.cfg4_nodes)
    t0 = time.perf_counter()
    solver = miqp.setup_many(raw, st, dict(problems.RANDOM_MIQP_QP_SETTINGS))[0]
    t_setup = time.perf_counter() - t0
    runs = []
    for rep in range(2):        # first run warms the context up
        solver.work.reset(); solver.work.first_run = 0
        solver.work.batches = solver.work.batched_nodes = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = solver.solve()
        torch.cuda.synchronize()
        runs.append(time.perf_counter() - t0)
    w = solver.work
    wall = runs[-1]
    nodes, iters = w.iter_num - 1, int(w.osqp_iter)
    dims = w.solver.dims()
    P, q, A, l, u, i_idx = problems.extend(raw[0])
    nnzA, nnz_triuP = A.nnz, spa.triu(P).nnz
    nnzL = nnzA + n * (n - 1) // 2                      # constraints-first elimination: A' A fills the trailing block
    N = n + A.shape[0]
    T = 2
    alg = 8.0 * (2 * n + 6 * A.shape[0]) + (24.0 * nnzL + 16.0 * N) / T + 12.0 * (2 * nnzA + nnz_triuP) / 25 / T
    # every step holds the two children: the step's tile runs max(iters of the two) iterations; approximate the tile-iterations by
    # half the node-iterations (siblings need similar counts) -- reported as such
    tile_iters = iters / 2.0
    tm = engine.last_timing()
    grid = int(tm.get("kernel", 1)) == 4
    if grid:
        tile_iters = float(iters)          # the native replay consumes one leaf per launch on this dive: a tile-iteration per node-iteration
    kernel_ms_last = float(tm.get("kernel_ms", 0.0))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"metric": "QP-relaxations/sec", "value": nodes / wall, "unit": "QP/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f64",
           "data": "synthetic", "vs_baseline": None,
           "config": {"workload": "random_miqp n=2000 m=4000 |i_idx|=200 density=0.05 (BASELINE config 4), one tree, B&B cut at %d nodes, "
                                  "native replay, 1 GPU" % args.cfg4_nodes, "qp_settings": QP_SETTINGS},
           "status": r.status, "qp_consumed": nodes, "admm_iters": iters, "launches": int(w.batches), "wall_s": wall, "setup_s": t_setup,
           "us_per_admm_iteration_of_the_tile": 1e6 * wall / max(1.0, tile_iters),
           "last_launch_kernel_ms": kernel_ms_last, "threads_per_cta": int(tm.get("threads", 0)), "ctas": int(tm.get("tiles", 0)),
           "e2e": {"value": nodes / wall, "unit": "QP/s", "note": "host buffers, replay and copies inside (the B&B loop has no device-resident variant)"},
           "roofline": {"bound": "hbm", "kernel": "admm_grid_kernel (one tile on every SM, 512 threads per CTA)" if grid else "admm_stream_kernel (one CTA per tile, two leaves)",
                        "unit": "GB/s", "peak": peak,
                        "algorithmic_bytes_per_node_iter": alg, "achieved": alg * iters / wall / 1e9, "frac": alg * iters / wall / 1e9 / peak,
                        "streamed_factor_bytes_per_tile_iteration": int(dims["factor_bytes"]),
                        "streamed_gbs_into_one_sm": dims["factor_bytes"] * tile_iters / wall / 1e9, "traffic": None,
                        "note": ("the tile's matrices (explicit reduced inverse 32.5 MB, CSR A and A' 9.6 MB) stay in L2 (126 MB): no HBM "
                                 "traffic to speak of, the kernel is bound by three grid-wide barriers and L2 round trips per iteration "
                                 "(23.7 us measured with convergence off, profiles/r02_summary.md F); the wall time also holds 39 launches "
                                 "with host replay, H2D and D2H") if grid else
                                ("one tile = one CTA = one SM streams the 28.8 MB factor + A twice per iteration out of L2 (126 MB): the "
                                 "kernel is bound by what one SM can pull, not by HBM; 147 SMs idle")}}
    if not args.no_cpu_baseline:
        try:
            secs, cn, cits, outs = bnb_cpu(raw, 1, 1, max_iter_bb=args.cfg4_nodes)
            out["cpu_baseline"] = {"value": cn / secs, "unit": "QP/s", "cores": 1, "kind": "port",
                                   "sample": "the same B&B (cut at %d nodes) on the CPU oracle, one relaxation at a time, %.1f s" % (args.cfg4_nodes, secs),
                                   "admm_node_iters_per_s": cits / secs, "same_node_and_iteration_counts_as_gpu": bool(cn == nodes and cits == iters)}
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(out))
    return 0


def reference_arm(args, cores, workload):
    """The reference's CPU path: its OSQP dependency is not installable here, so this times the CPU oracle
    (a restatement of the same algorithm) on every host thread, on a bounded sample of the same leaves."""
    from oracle import oracle
    S_inst = max(1, min(args.instances, cores // 2))   # instances in the sample
    if os.environ.get("BENCH_REFERENCE_MAX_INSTANCES"):   # tests: smaller sample
        S_inst = max(1, min(S_inst, int(os.environ["BENCH_REFERENCE_MAX_INSTANCES"])))
    insts = make_instances(S_inst, seed=1)
    n, m = N_VAR, M_CON + P_INT
    solvers = []
    for inst in insts:
        o = oracle.OSQP(); o.setup(inst[0], inst[1], inst[2], inst[3], inst[4], **QP_SETTINGS)
        solvers.append(o)
    xr, yr, st, it, _ = oracle.solve_multi(solvers, [i[3] for i in insts], [i[4] for i in insts],
                                           [np.zeros(n)] * S_inst, [np.zeros(m)] * S_inst, threads=cores)
    S, L, U, X0, Y0 = [], [], [], [], []
    for k, inst in enumerate(insts):
        x_root = np.nan_to_num(xr[k]); y_root = np.nan_to_num(yr[k])
        ll, uu = make_leaves(inst, x_root, y_root)
        for a, b in zip(ll, uu):
            S.append(solvers[k]); L.append(a); U.append(b); X0.append(x_root); Y0.append(y_root)
    for _ in range(min(args.warmup, 1)):
        oracle.solve_multi(S, L, U, X0, Y0, threads=cores)
    t0 = time.perf_counter()
    iters = 0
    for _ in range(args.steps):
        _, _, st, it, _ = oracle.solve_multi(S, L, U, X0, Y0, threads=cores)
        iters += int(np.sum(it))
    secs = time.perf_counter() - t0
    value = len(S) * args.steps / secs
    sample = "%d leaves (first %d instances x 8) per step, CPU oracle, %d threads" % (len(S), S_inst, cores)
    out = {"impl": "reference", "metric": "QP-relaxations/sec", "value": value, "unit": "QP/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload, "sample": sample, "qp_settings": QP_SETTINGS},
           "admm_node_iters_per_s": iters / secs,
           "cpu_baseline": {"value": value, "unit": "QP/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
