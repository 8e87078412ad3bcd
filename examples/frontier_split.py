#!/usr/bin/env python
"""
BASELINE config 4: ONE large MIQP whose B&B frontier is split across the GPUs of a node.  Every rank holds
the factor and replays the same tree; each B&B step the batch of unsolved leaves -- plus, with --speculation K, up to K
look-ahead nodes, which is what gives a single instance more than two nodes per step to spread -- is dealt round-robin
to the ranks, solved on the local GPU, exchanged with one all-gather, and the incumbent agreed with one
all-reduce(MIN).  Rank 0 prints one JSON line.

    python examples/frontier_split.py --vars 500 --rows 1000 --ints 50 --density 0.7 --speculation 64          # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        examples/frontier_split.py --vars 2000 --rows 4000 --ints 200 --density 0.05 --speculation 512           # config 4
"""
import argparse
import json
import os
import time

import _common  # noqa: F401


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--vars", dest="n", type=int, default=2000, help="n (torchrun claims every prefix of --n...)")
    ap.add_argument("--rows", dest="m", type=int, default=4000, help="m")
    ap.add_argument("--ints", dest="p", type=int, default=200, help="|i_idx|")
    ap.add_argument("--density", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--speculation", type=int, default=64)
    ap.add_argument("--max-nodes", type=int, default=1000, help="max_iter_bb of the reference's settings")
    ap.add_argument("--dist-backend", default="nccl", choices=("nccl", "gloo"),
                    help="torch.distributed backend for the node-result exchange (gloo: host-side exchange, used by the CPU tests)")
    args = ap.parse_args(argv)

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    ctx = None
    if world > 1:
        import torch
        import torch.distributed as dist
        if args.dist_backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            ctx = (rank, world, None, torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
            ctx = (rank, world, None, torch.device("cpu"))
    if ctx is not None:
        from miosqp_b200 import sharding
        ctx = sharding.DistCtx.of(ctx)
    import miosqp_b200
    from miosqp_b200 import problems

    pr = problems.random_miqp(args.n, args.m, args.p, args.density, seed=args.seed)[0]      # same draw on every rank
    s = miosqp_b200.MIOSQP()
    t0 = time.perf_counter()
    s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS, max_iter_bb=args.max_nodes, speculation=args.speculation),
            dict(problems.RANDOM_MIQP_QP_SETTINGS, device=local))
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = s.solve(dist_ctx=ctx)
    wall = time.perf_counter() - t0
    w = s.work
    if rank == 0:
        print(json.dumps({"workload": "random_miqp n=%d m=%d p=%d density=%g, frontier split over %d GPU(s)" % (args.n, args.m, args.p, args.density, world),
                          "status": r.status, "upper_glob": float(r.upper_glob), "nodes": w.iter_num - 1, "admm_iters": int(w.osqp_iter),
                          "launches": w.batches, "solved_nodes": w.batched_nodes, "speculation": args.speculation,
                          "spec_hit_rate": w.spec_hits / float(max(1, w.spec_nodes)), "solve_s": wall, "setup_s": t_setup,
                          "qp_per_s_consumed": (w.iter_num - 1) / wall, "qp_per_s_solved": w.batched_nodes / wall,
                          "exchange": None if ctx is None else {
                              "collective": "one all_gather_into_tensor of packed [status, iters, seconds, x, y] rows + one all-reduce(MIN) of the incumbent per B&B step",
                              "backend": args.dist_backend, "steps": ctx.exchanges, "seconds_total": ctx.exchange_s,
                              "ms_per_step": 1e3 * ctx.exchange_s / max(1, ctx.exchanges), "bytes_gathered_total": ctx.exchange_bytes}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
