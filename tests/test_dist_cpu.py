"""The N>1 path of the sharded search on CPU: world_size-2 gloo, local top-k from the oracle
(no GPU here), one all_gather, merge with rsd_topk_merge == the unsharded top-k."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import json
    import torch.distributed as dist
    from oracle import oracle as O
    from rna_sequence_diff_patch_b200.dist_search import shard_bounds, gather_merge
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = json.load(open(os.path.join(ROOT, "rna-sequence-diff-patch_b200", "dropin", "costs.json")))
    rng = np.random.default_rng(5)
    n = 3001
    lens = rng.integers(24, 32, size=n)
    off = np.zeros(n + 1, np.int64); np.cumsum(lens, out=off[1:])
    codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
    codes[::7] = codes[0]                                           # plenty of ties
    queries = [O.decode(codes[off[r]:off[r + 1]]) for r in (5, 77, 2999)]
    k = 9
    lo, hi = shard_bounds(lens, world)[rank]
    li = np.zeros((len(queries), k), np.int64); ls = np.zeros((len(queries), k), np.float64)
    for q, query in enumerate(queries):
        sub_off = (off[lo:hi + 1] - off[lo]).copy()
        i, s = O.search_topk(query, codes[off[lo]:off[hi]].copy(), sub_off, costs, min(k, hi - lo))
        li[q, :len(i)] = i + lo; ls[q, :len(s)] = s
        li[q, len(i):] = -1
    mi, ms = gather_merge(li, ls)
    if rank == 0:
        wi = np.stack([O.search_topk(query, codes, off, costs, k)[0] for query in queries])
        ws = np.stack([O.search_topk(query, codes, off, costs, k)[1] for query in queries])
        np.save(os.path.join(tmp, "ok.npy"), np.array([np.array_equal(mi, wi) and np.array_equal(ms, ws)]))
    dist.destroy_process_group()


def test_gloo_world2_gather_merge(tmp_path):
    import torch.multiprocessing as mp
    import __graft_entry__ as G
    G.build()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(os.path.join(str(tmp_path), "ok.npy"))[0]


def test_shard_bounds_cover_and_balance():
    sys.path.insert(0, ROOT)
    from rna_sequence_diff_patch_b200.dist_search import shard_bounds
    rng = np.random.default_rng(1)
    lens = rng.integers(24, 32, size=10007)
    for world in (1, 2, 4, 8):
        b = shard_bounds(lens, world)
        assert b[0][0] == 0 and b[-1][1] == 10007
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        sums = [int(lens[lo:hi].sum()) for lo, hi in b]
        assert max(sums) - min(sums) <= 64
    assert shard_bounds(np.zeros(0, np.int64), 4) == [(0, 0)] * 4
