import ctypes as C, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import _lib
import bench
ca, oa, cb, ob = bench.gen_pairs(1_000_000, 20260002, 4)
A = R.pack((ca, oa)); B = R.pack((cb, ob))
eng = R.Engine(0); eng.set_costs(__import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs())
lib = R.load_library()
def run(bufs, out, label, reps=5):
    ptr = lambda a, t: C.cast(a, C.POINTER(t))
    mode = C.c_int()
    ts = []
    for r in range(reps + 2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(lib.rsd_distance_batch(eng.ctx, ptr(bufs[0], C.c_uint32), ptr(bufs[1], C.c_int64), ptr(bufs[2], C.c_int32), A.words.shape[0],
                   ptr(bufs[3], C.c_uint32), ptr(bufs[4], C.c_int64), ptr(bufs[5], C.c_int32), B.words.shape[0], A.n, A.max_len, B.max_len, A.bits, 15, 0, ptr(out, C.c_double), C.byref(mode)))
        ts.append(time.perf_counter() - t0)
    print(label, "ms:", [round(t * 1e3, 2) for t in ts], "kernel_ms", eng.last_kernel_ms(), flush=True)
arrs = [A.words, A.start, A.len, B.words, B.start, B.len]
# pageable
out = np.zeros(A.n); run([a.ctypes.data for a in arrs], out.ctypes.data, "pageable")
# torch pinned
tp = [torch.from_numpy(a).pin_memory() for a in arrs]; outp = torch.zeros(A.n, dtype=torch.float64).pin_memory()
print("is_pinned", [t.is_pinned() for t in tp])
run([t.data_ptr() for t in tp], outp.data_ptr(), "torch-pinned")
# rsd_host_alloc
hb = []
for a in arrs:
    p = C.c_void_p(); _lib.check(lib.rsd_host_alloc(C.byref(p), a.nbytes)); C.memmove(p.value, a.ctypes.data, a.nbytes); hb.append(p.value)
po = C.c_void_p(); _lib.check(lib.rsd_host_alloc(C.byref(po), A.n * 8))
run(hb, po.value, "rsd_host_alloc")
# torch H2D speed of the same pinned tensors
dev = torch.device("cuda", 0)
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = [t.to(dev, non_blocking=True) for t in tp]; torch.cuda.synchronize()
    print("torch H2D all ms", round((time.perf_counter() - t0) * 1e3, 2), "bytes", sum(t.numel() * t.element_size() for t in tp))
eng.set_timing(True)
run([t.data_ptr() for t in tp], outp.data_ptr(), "torch-pinned+timing")
