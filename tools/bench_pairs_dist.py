#!/usr/bin/env python
"""C2 pair list across GPUs, strong scaling (first row): ONE list of pairs shared by all ranks is cut
into contiguous ranges balanced by the number of cells, every rank scores its range through the host C-ABI call
(pinned host buffers, H2D + D2H inside), and the 8 B/pair results are gathered in pair order with one NCCL
all_gather.  Rank 0 prints one JSON line; --check compares a sample with the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_pairs_dist.py [--pairs 1000000] [--steps 10]"""
import argparse
import ctypes as C
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
from rna_sequence_diff_patch_b200 import _lib  # noqa: E402
from rna_sequence_diff_patch_b200.dist_pairs import pair_shard_bounds  # noqa: E402
from rna_sequence_diff_patch_b200.dist_search import slice_packed  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
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
    ca, oa, cb, ob = bench.gen_pairs(args.pairs, bench.SEEDS["c2"], 4)           # the same list on every rank
    A = R.pack((ca, oa)); B = R.pack((cb, ob))
    cells = float((np.diff(oa) * np.diff(ob)).sum())
    bounds = pair_shard_bounds(A.len, B.len, world)
    lo, hi = bounds[rank]
    a, b = slice_packed(A, lo, hi), slice_packed(B, lo, hi)
    costs = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs()
    eng = R.Engine(local); eng.set_costs(costs)
    lib = R.load_library()
    hp = [torch.from_numpy(x).pin_memory() for x in (a.words, a.start, a.len, b.words, b.start, b.len)]
    width = max(h - l for l, h in bounds)
    out_host = torch.zeros(width, dtype=torch.float64).pin_memory()
    out_dev = torch.zeros(width, dtype=torch.float64, device=dev)
    gathered = torch.zeros(world * width, dtype=torch.float64, device=dev)
    ptr = lambda t, ty: C.cast(t.data_ptr(), C.POINTER(ty))
    mode = C.c_int()

    def step():
        _lib.check(lib.rsd_distance_batch(eng.ctx, ptr(hp[0], C.c_uint32), ptr(hp[1], C.c_int64), ptr(hp[2], C.c_int32), a.words.shape[0],
                                          ptr(hp[3], C.c_uint32), ptr(hp[4], C.c_int64), ptr(hp[5], C.c_int32), b.words.shape[0],
                                          hi - lo, A.max_len, B.max_len, A.bits, 0xF, 0, ptr(out_host, C.c_double), C.byref(mode)))
        if world > 1:
            out_dev.copy_(out_host, non_blocking=True)
            dist.all_gather_into_tensor(gathered, out_dev)
        else:
            gathered.copy_(out_host, non_blocking=True)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    res = gathered.cpu().numpy().reshape(world, width)
    full = np.concatenate([res[r, :h - l] for r, (l, h) in enumerate(bounds)])
    if rank == 0:
        out = {"config": "C2-pairs-dist", "n_gpus": world, "pairs": args.pairs, "ms_per_list": ms / args.steps,
               "gcups_e2e": cells * args.steps / (ms * 1e-3) * 1e-9, "scaling": "strong",
               "shard_cells_max_over_mean": float(max(float((A.len[l:h].astype(np.float64) * B.len[l:h]).sum()) for l, h in bounds) / (cells / world)),
               "collective": "one all_gather_into_tensor of 8 B/pair", "mode": mode.value}
        if args.check:
            from oracle import oracle as O
            sub = np.random.default_rng(1).choice(args.pairs, size=2000, replace=False)
            sa = [O.decode(ca[oa[k]:oa[k + 1]]) for k in sub]; sb = [O.decode(cb[ob[k]:ob[k + 1]]) for k in sub]
            xa, xo = O.concat(sa); ya, yo = O.concat(sb)
            out["check"] = bool(np.array_equal(full[sub], O.distance_batch(xa, xo, ya, yo, costs)))
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
