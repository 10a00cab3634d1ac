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
    costs = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).default_costs()
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


class _OracleEngine:
    """Stands in for Engine in the CPU-only world_size-2 test of the pair-list sharding (no GPU here)."""

    def __init__(self, costs):
        self.costs = costs

    def distance_batch(self, A, B, force_mode=0):
        from oracle import oracle as O
        import rna_sequence_diff_patch_b200 as R
        from rna_sequence_diff_patch_b200.encoding import unpack
        ca, oa = unpack(A); cb, ob = unpack(B)
        return O.distance_batch(ca, oa, cb, ob, self.costs)


def _pairs_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import json
    import torch.distributed as dist
    from oracle import oracle as O
    import rna_sequence_diff_patch_b200 as R
    from rna_sequence_diff_patch_b200.dist_pairs import ShardedPairs
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs()
    rng = np.random.default_rng(11)
    n = 501
    la = rng.integers(0, 60, size=n); lb = rng.integers(0, 60, size=n)
    oa = np.zeros(n + 1, np.int64); np.cumsum(la, out=oa[1:]); ob = np.zeros(n + 1, np.int64); np.cumsum(lb, out=ob[1:])
    ca = rng.integers(0, 4, size=int(oa[-1]), dtype=np.uint8); cb = rng.integers(0, 4, size=int(ob[-1]), dtype=np.uint8)
    A = R.pack((ca, oa)); B = R.pack((cb, ob))
    got = ShardedPairs(_OracleEngine(costs), rank, world).distance_batch(A, B)
    want = O.distance_batch(ca, oa, cb, ob, costs)
    np.save(os.path.join(tmp, f"ok{rank}.npy"), np.array([np.array_equal(got, want)]))
    dist.destroy_process_group()


def test_gloo_world2_sharded_pairs(tmp_path):
    import torch.multiprocessing as mp
    import __graft_entry__ as G
    G.build()
    port = 30100 + (os.getpid() % 500)
    mp.spawn(_pairs_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(np.load(os.path.join(str(tmp_path), f"ok{r}.npy"))[0] for r in range(2))


def test_pair_shard_bounds_balance_cells():
    sys.path.insert(0, ROOT)
    from rna_sequence_diff_patch_b200.dist_pairs import pair_shard_bounds
    rng = np.random.default_rng(2)
    la = rng.integers(100, 301, size=20011); lb = rng.integers(100, 301, size=20011)
    la[:5000] = 300; lb[:5000] = 300                                 # heavy head: equal counts would be unbalanced
    cells = la.astype(np.int64) * lb
    for world in (1, 2, 3, 8):
        b = pair_shard_bounds(la, lb, world)
        assert b[0][0] == 0 and b[-1][1] == 20011 and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        sums = [int(cells[lo:hi].sum()) for lo, hi in b]
        assert max(sums) - min(sums) <= 2 * 300 * 300
    assert pair_shard_bounds(np.zeros(0, np.int32), np.zeros(0, np.int32), 3) == [(0, 0)] * 3
