"""Host-path diagnostics on the C2 workload: H2D rate of the box, then RSD_TRACE timelines of the three input forms
of the host distance call (explicit start[], canonical start == NULL, raw codes)."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import _lib
import bench
n = int(os.environ.get("PAIRS", 1_000_000))
ca, oa, cb, ob = bench.gen_pairs(n, 20260002, 4)
A = R.pack((ca, oa)); B = R.pack((cb, ob))
eng = R.Engine(0); eng.set_costs(__import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs())
eng.set_timing(True)
lib = R.load_library()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hp = dict(aw=pin(A.words), as_=pin(A.start), al=pin(A.len), bw=pin(B.words), bs=pin(B.start), bl=pin(B.len), ca=pin(ca), cb=pin(cb))
out = torch.zeros(n, dtype=torch.float64).pin_memory()
p = lambda t, ty: C.cast(t.data_ptr(), C.POINTER(ty))
dev = torch.device("cuda", 0)
for name in ("aw", "ca"):
    for r in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d = hp[name].to(dev, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"H2D {name}: {hp[name].numel() * hp[name].element_size() / dt * 1e-9:.1f} GB/s ({dt * 1e3:.2f} ms)", flush=True)
mode = C.c_int()
def explicit():
    _lib.check(lib.rsd_distance_batch(eng.ctx, p(hp["aw"], C.c_uint32), p(hp["as_"], C.c_int64), p(hp["al"], C.c_int32), A.words.shape[0],
               p(hp["bw"], C.c_uint32), p(hp["bs"], C.c_int64), p(hp["bl"], C.c_int32), B.words.shape[0], n, A.max_len, B.max_len, A.bits, 15, 0, p(out, C.c_double), C.byref(mode)))
def canonical():
    _lib.check(lib.rsd_distance_batch(eng.ctx, p(hp["aw"], C.c_uint32), None, p(hp["al"], C.c_int32), A.words.shape[0],
               p(hp["bw"], C.c_uint32), None, p(hp["bl"], C.c_int32), B.words.shape[0], n, A.max_len, B.max_len, A.bits, 15, 0, p(out, C.c_double), C.byref(mode)))
def codes():
    _lib.check(lib.rsd_distance_batch_codes(eng.ctx, p(hp["ca"], C.c_uint8), p(hp["al"], C.c_int32), p(hp["cb"], C.c_uint8), p(hp["bl"], C.c_int32),
               n, A.max_len, B.max_len, A.bits, 15, 0, p(out, C.c_double), C.byref(mode)))
for label, fn in (("explicit start", explicit), ("canonical", canonical), ("raw codes", codes)):
    ts = []
    for r in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(label, "ms:", [round(t * 1e3, 2) for t in ts], "kernel_ms", round(eng.last_kernel_ms(), 3), flush=True)
    os.environ["RSD_TRACE"] = "1"; fn(); del os.environ["RSD_TRACE"]
