#!/usr/bin/env python
"""C5 across GPUs: one process per GPU (torchrun), database sharded contiguously by record index,
same query batch everywhere, local exact top-k on each shard, ONE NCCL all_gather of Q*k*(8+8) bytes
per rank, merge with the same key.  Strong scaling: the database size is fixed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_search_dist.py [--records 10000000] [--queries 64] [--k 10] [--steps 10]
Rank 0 prints one JSON line; with --check the merged result is compared with the oracle on a sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

G.build()
import rna_sequence_diff_patch_b200 as R  # noqa: E402
from rna_sequence_diff_patch_b200.dist_search import shard_bounds, slice_packed  # noqa: E402
from rna_sequence_diff_patch_b200.engine import topk_merge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=64)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # every rank builds the same synthetic database (seeded) and keeps only its shard
    rng = np.random.default_rng(20260005)
    lens = rng.integers(24, 32, size=args.records)
    off = np.zeros(args.records + 1, np.int64); np.cumsum(lens, out=off[1:])
    codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
    codes[rng.random(codes.shape[0]) < 1e-3] = 14
    qs, qo = [], [0]
    for r in rng.integers(0, args.records, size=args.queries):
        s = codes[off[r]:off[r + 1]].copy()
        hit = rng.random(s.shape[0]) < 0.1
        s[hit] = rng.integers(0, 4, size=int(hit.sum()), dtype=np.uint8)
        qs.append(s); qo.append(qo[-1] + s.shape[0])
    Q = R.pack((np.concatenate(qs), np.array(qo, np.int64)), bits=4)
    lo, hi = shard_bounds(lens, world)[rank]
    shard = R.pack((codes[off[lo]:off[hi]], (off[lo:hi + 1] - off[lo]).copy()), bits=4)
    shard.symmask |= 1 << 14
    costs = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).default_costs()
    eng = R.Engine(local); eng.set_costs(costs)
    eng.db_load(shard, global_index_base=lo)

    qd = {k: torch.from_numpy(v).to(dev) for k, v in dict(w=Q.words, s=Q.start, l=Q.len).items()}
    top_i = torch.empty((args.queries, args.k), dtype=torch.int64, device=dev)
    top_s = torch.empty((args.queries, args.k), dtype=torch.float64, device=dev)
    gi = torch.empty((world, args.queries, args.k), dtype=torch.int64, device=dev)
    gs = torch.empty((world, args.queries, args.k), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        eng.db_search_topk_dev(qd["w"].data_ptr(), qd["s"].data_ptr(), qd["l"].data_ptr(), args.queries, Q.max_len, 4,
                               Q.symmask, args.k, top_i.data_ptr(), top_s.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(gi, top_i)
            dist.all_gather_into_tensor(gs, top_s)
        else:
            gi[0].copy_(top_i); gs[0].copy_(top_s)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record(); e1.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    mi, msc = topk_merge(gi.cpu().numpy(), gs.cpu().numpy())
    cells = float(np.array([len(q) for q in qs], np.float64).sum() * lens.astype(np.float64).sum())
    if rank == 0:
        out = {"config": "C5-dist", "n_gpus": world, "records": args.records, "queries": args.queries, "k": args.k,
               "ms_per_query_batch": ms / args.steps, "gcups": cells * args.steps / (ms * 1e-3) * 1e-9,
               "wall_ms_per_batch": wall / args.steps * 1e3, "collective": "all_gather_into_tensor x2 (Q*k*16 B per rank)",
               "mode": eng.last_mode, "top1_head": msc[:3, 0].tolist()}
        if args.check:
            from oracle import oracle as O
            n_chk = min(args.records, 200000)
            idx, sc = O.search_topk(O.decode(qs[0]), codes[:off[n_chk]], off[:n_chk + 1].copy(), costs, args.k)
            if args.records == n_chk:
                out["check"] = bool(np.array_equal(idx, mi[0]) and np.array_equal(sc, msc[0]))
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
