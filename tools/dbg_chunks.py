"""Sweep the host-path chunk count / growth ratio of rsd_distance_batch on the C2 workload (debug tool)."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import _lib
import bench
ca, oa, cb, ob = bench.gen_pairs(1_000_000, 20260002, 4)
A = R.pack((ca, oa)); B = R.pack((cb, ob))
eng = R.Engine(0); eng.set_costs(__import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs())
lib = R.load_library()
arrs = [A.words, A.start, A.len, B.words, B.start, B.len]
tp = [torch.from_numpy(a).pin_memory() for a in arrs]; outp = torch.zeros(A.n, dtype=torch.float64).pin_memory()
ptr = lambda a, t: C.cast(a, C.POINTER(t))
def run(reps=12):
    mode = C.c_int(); ts = []
    for r in range(reps + 3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(lib.rsd_distance_batch(eng.ctx, ptr(tp[0].data_ptr(), C.c_uint32), ptr(tp[1].data_ptr(), C.c_int64), ptr(tp[2].data_ptr(), C.c_int32), A.words.shape[0],
                   ptr(tp[3].data_ptr(), C.c_uint32), ptr(tp[4].data_ptr(), C.c_int64), ptr(tp[5].data_ptr(), C.c_int32), B.words.shape[0], A.n, A.max_len, B.max_len, A.bits, 15, 0,
                   ptr(outp.data_ptr(), C.c_double), C.byref(mode)))
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts[3:]) * 1e3), float(np.min(ts[3:]) * 1e3)
for n, r in [(1, 1), (4, 2), (4, 1.5), (5, 1.5), (6, 1.4), (6, 1.5), (7, 1.3), (7, 1.4), (7, 1.5), (8, 1.3), (8, 1.4), (8, 1.5), (8, 1.6), (6, 1.7), (5, 2)]:
    os.environ["RSD_CHUNKS"] = str(n); os.environ["RSD_CHUNK_RATIO"] = str(r)
    print("chunks", n, "ratio", r, "median/min ms", run(), flush=True)
if len(sys.argv) > 1:
    os.environ["RSD_CHUNKS"], os.environ["RSD_CHUNK_RATIO"] = sys.argv[1], sys.argv[2]
    os.environ["RSD_TRACE"] = "1"; eng.set_timing(True); run(1)
