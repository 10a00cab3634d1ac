// rsd_api.cu — the C ABI of librsd.so (include/rsd.h): context, cost classification, packing,
// buffer management and kernel dispatch.  No torch types, no CPU fallback: every compute entry
// point needs a CUDA device and says so when there is none.
#include "../../include/rsd.h"
#include "rsd_ctx.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "k_dist.cuh"
#include "k_matrix.cuh"
#include "k_ubench.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int rsd_fail(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}
extern "C" const char *rsd_last_error(void) { return g_err; }
extern "C" int rsd_abi_version(void) { return RSD_ABI_VERSION; }

extern "C" int rsd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int rsd_create(int device, rsd_ctx **out) {
    if (!out) return rsd_fail(RSD_EINVAL, "rsd_create: out is NULL");
    rsd_ctx *c = new (std::nothrow) rsd_ctx();
    if (!c) return rsd_fail(RSD_ENOMEM, "rsd_create: out of host memory");
    c->device = device;
    *out = c;
    return RSD_OK;
}

int rsd_ctx::ensure_device() {
    if (inited) {
        if (pid != getpid())
            return rsd_fail(RSD_EINVAL, "rsd: context used after fork(); create a new context in the child process");
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return rsd_fail(RSD_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return RSD_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return rsd_fail(RSD_ENODEV, "rsd: no CUDA device available (%s); librsd has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return rsd_fail(RSD_EINVAL, "rsd: device %d out of range (0..%d)", device, n - 1);
    RSD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RSD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return rsd_fail(RSD_ENODEV, "rsd: device %d is sm_%d%d; this library is built for sm_100a only", device,
                        prop.major, prop.minor);
    sm_count = prop.multiProcessorCount;
    RSD_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    RSD_CUDA(cudaEventCreate(&ev0));
    RSD_CUDA(cudaEventCreate(&ev1));
    RSD_CUDA(cudaMalloc(&d_ic, sizeof(IntCosts)));
    RSD_CUDA(cudaMalloc(&d_fc, sizeof(F64Costs)));
    pid = getpid();
    inited = true;
    return RSD_OK;
}

extern "C" int rsd_destroy(rsd_ctx *c) {
    if (!c) return RSD_OK;
    if (c->inited && c->pid == getpid()) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        c->free_all();
        cudaFree(c->d_ic); cudaFree(c->d_fc);
        cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
        cudaStreamDestroy(c->stream);
    }
    delete c;
    return RSD_OK;
}

extern "C" int rsd_host_alloc(void **out, int64_t bytes) {
    if (!out || bytes < 0) return rsd_fail(RSD_EINVAL, "rsd_host_alloc: bad arguments");
    cudaError_t e = cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { cudaGetLastError(); return rsd_fail(RSD_ENOMEM, "cudaMallocHost(%lld): %s", (long long)bytes, cudaGetErrorString(e)); }
    return RSD_OK;
}
extern "C" int rsd_host_free(void *p) {
    if (p) cudaFreeHost(p);
    return RSD_OK;
}

extern "C" int64_t rsd_launch_count(rsd_ctx *c) { return c ? c->launches : 0; }
extern "C" int rsd_set_timing(rsd_ctx *c, int on) { if (!c) return rsd_fail(RSD_EINVAL, "ctx is NULL"); c->timing = on != 0; return RSD_OK; }
extern "C" double rsd_last_kernel_ms(rsd_ctx *c) {
    if (!c || !c->inited || !c->timed) return 0.0;
    float ms = 0.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return 0.0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return 0.0;
    return (double)ms;
}

// ------------------------------------------------------------------------------------------------
// costs + numeric-mode classifier
// ------------------------------------------------------------------------------------------------
extern "C" int rsd_set_costs(rsd_ctx *c, double ins, double del, const double *sub) {
    if (!c || !sub) return rsd_fail(RSD_EINVAL, "rsd_set_costs: NULL argument");
    auto bad = [](double v) { return !(v >= 0.0) || std::isinf(v); };
    if (bad(ins) || bad(del)) return rsd_fail(RSD_ECOSTS, "rsd_set_costs: insert/delete cost must be finite and >= 0");
    for (int a = 0; a < 15; ++a)
        for (int b = 0; b < 15; ++b)
            if (a != b && bad(sub[a * 15 + b]))
                return rsd_fail(RSD_ECOSTS, "rsd_set_costs: update cost [%d][%d] must be finite and >= 0", a, b);
    c->ins = ins; c->del = del;
    for (int a = 0; a < 15; ++a)
        for (int b = 0; b < 15; ++b) c->sub[a][b] = a == b ? 0.0 : sub[a * 15 + b];
    c->have_costs = true;
    return RSD_OK;
}

int rsd_ctx::classify(uint32_t symmask, int64_t max_m, int64_t max_n, int bits, int force_mode, ModeInfo &mi) const {
    if (!have_costs) return rsd_fail(RSD_EINVAL, "rsd: rsd_set_costs has not been called");
    memset(&mi, 0, sizeof mi);
    mi.fc.ins = ins; mi.fc.del = del;
    for (int a = 0; a < 15; ++a)
        for (int b = 0; b < 15; ++b) mi.fc.sub[a][b] = a == b ? 0.0 : sub[a][b];
    // reachable costs: ins, del, sub[a][b] for a != b both present
    double reach[2 + 15 * 15]; int nr = 0;
    reach[nr++] = ins; reach[nr++] = del;
    for (int a = 0; a < 15; ++a) if (symmask >> a & 1)
        for (int b = 0; b < 15; ++b) if (a != b && (symmask >> b & 1)) reach[nr++] = sub[a][b];
    int k = -1;
    for (int kk = 0; kk <= 16 && k < 0; ++kk) {
        bool ok = true;
        for (int r = 0; r < nr && ok; ++r) {
            double s = std::ldexp(reach[r], kk);
            ok = s == std::floor(s) && s < 1073741824.0;
        }
        if (ok) k = kk;
    }
    bool i32_ok = false, i16_ok = false, fast_ok = false;
    if (k >= 0) {
        IntCosts &ic = mi.ic;
        ic.scale_log2 = k;
        ic.ins = (int32_t)std::ldexp(ins, k); ic.del = (int32_t)std::ldexp(del, k);
        int64_t maxabsw = (int64_t)ic.ins + ic.del;
        bool w8 = true;
        for (int a = 0; a < 16; ++a)
            for (int b = 0; b < 16; ++b) {
                int64_t w;
                if (a == b) w = -((int64_t)ic.ins + ic.del);
                else if (a < 15 && b < 15 && (symmask >> a & 1) && (symmask >> b & 1))
                    w = (int64_t)std::ldexp(sub[a][b], k) - ic.ins - ic.del;
                else w = 0;                                   // unreachable entry
                if (std::llabs(w) > maxabsw) maxabsw = std::llabs(w);
                if (a < 4 && b < 4 && (w < -128 || w > 127)) w8 = false;
                ic.w[a][b] = (int32_t)std::max<int64_t>(std::min<int64_t>(w, INT32_MAX), INT32_MIN);
            }
        // |H'| <= i*del + j*ins on the extended (strip-padded) matrix
        const int64_t pad_n = max_n + 64;
        const double bound = (double)max_m * ic.del + (double)pad_n * ic.ins + (double)maxabsw;
        i32_ok = bound < 2147483000.0;
        i16_ok = bound < 32760.0;
        fast_ok = i16_ok && w8 && bits == 2 && (symmask & ~0xFu) == 0;
        if (fast_ok)
            for (int a = 0; a < 4; ++a) {
                uint32_t r = 0;
                for (int b = 0; b < 4; ++b) r |= (uint32_t)(uint8_t)(int8_t)ic.w[a][b] << (8 * b);
                ic.rowtab4[a] = r;
            }
    }
    int mode;
    if (force_mode == 0) mode = fast_ok ? RSD_MODE_I16X2 : (i32_ok ? RSD_MODE_I32 : RSD_MODE_F64);
    else if (force_mode == RSD_MODE_I16X2) {
        if (!fast_ok) return rsd_fail(RSD_EINVAL, "rsd: I16X2 mode not exact/applicable for these costs, symbols or lengths");
        mode = force_mode;
    } else if (force_mode == RSD_MODE_I32) {
        if (!i32_ok) return rsd_fail(RSD_EINVAL, "rsd: I32 mode not exact for these costs (not dyadic) or lengths");
        mode = force_mode;
    } else if (force_mode == RSD_MODE_F64) mode = force_mode;
    else return rsd_fail(RSD_EINVAL, "rsd: unknown force_mode %d", force_mode);
    mi.mode = mode;
    mi.k = k < 0 ? 0 : k;
    return RSD_OK;
}

extern "C" int rsd_classify(rsd_ctx *c, uint32_t symmask, int64_t max_m, int64_t max_n, int force_mode,
                            int *mode_out, int *scale_log2_out) {
    if (!c) return rsd_fail(RSD_EINVAL, "ctx is NULL");
    ModeInfo mi;
    int bits = (symmask & ~0xFu) == 0 ? 2 : 4;
    int rc = c->classify(symmask, max_m, max_n, bits, force_mode, mi);
    if (rc) return rc;
    if (mode_out) *mode_out = mi.mode;
    if (scale_log2_out) *scale_log2_out = mi.k;
    return RSD_OK;
}

int rsd_ctx::upload_costs(const ModeInfo &mi, cudaStream_t st) {
    // staged through pinned-free small copies; cudaMemcpyAsync from pageable memory copies the
    // source before returning, so the stack ModeInfo is safe.
    RSD_CUDA(cudaMemcpyAsync(d_ic, &mi.ic, sizeof(IntCosts), cudaMemcpyHostToDevice, st));
    RSD_CUDA(cudaMemcpyAsync(d_fc, &mi.fc, sizeof(F64Costs), cudaMemcpyHostToDevice, st));
    return RSD_OK;
}

// ------------------------------------------------------------------------------------------------
// packing (ingest)
// ------------------------------------------------------------------------------------------------
extern "C" int64_t rsd_pack_words(const int32_t *len, int64_t n, int bits) {
    if (!len || (bits != 2 && bits != 4)) return -1;
    const int per = 32 / bits;
    int64_t w = 0;
    for (int64_t i = 0; i < n; ++i) w += (len[i] + per - 1) / per;
    return w + 4;
}

extern "C" int rsd_pack(const uint8_t *codes, const int64_t *off, int64_t n, int bits,
                        uint32_t *words, int64_t *start, int32_t *len, uint32_t *symmask_inout) {
    if (!off || !words || !start || !len || (n > 0 && !codes)) return rsd_fail(RSD_EINVAL, "rsd_pack: NULL argument");
    if (bits != 2 && bits != 4) return rsd_fail(RSD_EINVAL, "rsd_pack: bits must be 2 or 4");
    const int per = 32 / bits;
    const uint32_t lim = 1u << bits;
    int64_t w = 0;
    uint32_t mask = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t L = off[i + 1] - off[i];
        if (L < 0 || L > INT32_MAX) return rsd_fail(RSD_EINVAL, "rsd_pack: bad offsets at sequence %lld", (long long)i);
        start[i] = w; len[i] = (int32_t)L;
        const uint8_t *s = codes + off[i];
        for (int64_t k = 0; k < L; k += per) {
            uint32_t word = 0;
            const int lim_k = (int)std::min<int64_t>(per, L - k);
            for (int c = 0; c < lim_k; ++c) {
                const uint32_t v = s[k + c];
                if (v >= lim) return rsd_fail(RSD_EINVAL, "rsd_pack: code %u at sequence %lld does not fit %d bits", v, (long long)i, bits);
                mask |= 1u << v;
                word |= v << (bits * c);
            }
            words[w++] = word;
        }
    }
    for (int p = 0; p < 4; ++p) words[w++] = 0;
    if (symmask_inout) *symmask_inout |= mask;
    return RSD_OK;
}

// ------------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------------
int rsd_ctx::make_plan(const int32_t *d_alen, const int32_t *d_blen, int64_t n_pairs, int C, int allow_twin,
                       double *d_out, cudaStream_t st, PlanView &pv) {
    RSD_OK_OR_RETURN(plan_pair_bin.ensure(sizeof(int) * (size_t)n_pairs));
    RSD_OK_OR_RETURN(plan_bins.ensure(sizeof(int) * (size_t)(4 * (RSD_NB + 1) + 8)));
    RSD_OK_OR_RETURN(plan_groups.ensure(sizeof(int2) * (size_t)n_pairs));
    int *bins = (int *)plan_bins.p;
    pv.pair_bin = (int *)plan_pair_bin.p;
    pv.bin_cnt = bins;
    pv.bin_cursor = bins + (RSD_NB + 1);
    pv.bin_group_off = bins + 2 * (RSD_NB + 1);
    pv.bin_warp_off = bins + 3 * (RSD_NB + 1);
    pv.totals = bins + 4 * (RSD_NB + 1);
    pv.work_counter = pv.totals + 4;
    pv.groups = (int2 *)plan_groups.p;
    pv.C = C; pv.allow_twin = allow_twin;
    RSD_CUDA(cudaMemsetAsync(bins, 0, sizeof(int) * (size_t)(4 * (RSD_NB + 1) + 8), st));
    RSD_CUDA(cudaMemsetAsync(pv.groups, 0xFF, sizeof(int2) * (size_t)n_pairs, st));
    const int T = 256;
    const unsigned G = (unsigned)((n_pairs + T - 1) / T);
    k_plan_count<<<G, T, 0, st>>>(d_alen, d_blen, n_pairs, pv, ins, del, d_out);
    k_plan_scan<<<1, 1024, 0, st>>>(pv);
    k_plan_fill<<<G, T, 0, st>>>(n_pairs, pv);
    launches += 3;
    RSD_CUDA(cudaGetLastError());
    return RSD_OK;
}

// ------------------------------------------------------------------------------------------------
// distance batch
// ------------------------------------------------------------------------------------------------
template <typename K>
static int persistent_grid(K kernel, int threads, int sm_count, int &blocks) {
    int per_sm = 0;
    RSD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
    if (per_sm < 1) per_sm = 1;
    blocks = per_sm * sm_count;
    return RSD_OK;
}

int rsd_ctx::distance_dev(const uint32_t *a_words, const int64_t *a_start, const int32_t *a_len,
                          const uint32_t *b_words, const int64_t *b_start, const int32_t *b_len,
                          int64_t n_pairs, int64_t max_m, int64_t max_n, int bits, uint32_t symmask,
                          int force_mode, double *d_out, int *mode_out, cudaStream_t st) {
    if (n_pairs < 0 || n_pairs > INT32_MAX) return rsd_fail(RSD_ERANGE, "rsd: n_pairs out of range");
    if (bits != 2 && bits != 4) return rsd_fail(RSD_EINVAL, "rsd: bits must be 2 or 4");
    if (bits == 2 && (symmask & ~0xFu)) return rsd_fail(RSD_EINVAL, "rsd: 2-bit packing with symbols outside ACGU");
    ModeInfo mi;
    RSD_OK_OR_RETURN(classify(symmask, max_m, max_n, bits, force_mode, mi));
    if (mode_out) *mode_out = mi.mode;
    timed = false;
    if (n_pairs == 0) return RSD_OK;
    RSD_OK_OR_RETURN(upload_costs(mi, st));
    SeqView A{a_words, a_start, a_len}, B{b_words, b_start, b_len};
    PlanView pv;
    constexpr int THREADS = 128;
    const int wpb = THREADS / 32;
    int blocks = 0;
    if (mi.mode == RSD_MODE_I16X2) {
        constexpr int C = 32;
        RSD_OK_OR_RETURN(make_plan(a_len, b_len, n_pairs, C, 1, d_out, st, pv));
        RSD_OK_OR_RETURN(persistent_grid(k_dist_twin16<C>, THREADS, sm_count, blocks));
        const int stride = max_n > 32 * C ? (int)max_m : 0;
        RSD_OK_OR_RETURN(scratch.ensure(sizeof(uint32_t) * (size_t)stride * blocks * wpb + 16));
        if (timing) RSD_CUDA(cudaEventRecord(ev0, st));
        k_dist_twin16<C><<<blocks, THREADS, 0, st>>>(pv, A, B, mi.ic, d_out, (uint32_t *)scratch.p, stride);
    } else if (mi.mode == RSD_MODE_I32) {
        constexpr int C = 32;
        RSD_OK_OR_RETURN(make_plan(a_len, b_len, n_pairs, C, 0, d_out, st, pv));
        const int stride = max_n > 32 * C ? (int)max_m : 0;
        if (bits == 2) {
            RSD_OK_OR_RETURN(persistent_grid(k_dist_gen<int, 2, C>, THREADS, sm_count, blocks));
            RSD_OK_OR_RETURN(scratch.ensure(sizeof(int) * (size_t)stride * blocks * wpb + 16));
            if (timing) RSD_CUDA(cudaEventRecord(ev0, st));
            k_dist_gen<int, 2, C><<<blocks, THREADS, 0, st>>>(pv, A, B, d_ic, d_fc, d_out, (int *)scratch.p, stride);
        } else {
            RSD_OK_OR_RETURN(persistent_grid(k_dist_gen<int, 4, C>, THREADS, sm_count, blocks));
            RSD_OK_OR_RETURN(scratch.ensure(sizeof(int) * (size_t)stride * blocks * wpb + 16));
            if (timing) RSD_CUDA(cudaEventRecord(ev0, st));
            k_dist_gen<int, 4, C><<<blocks, THREADS, 0, st>>>(pv, A, B, d_ic, d_fc, d_out, (int *)scratch.p, stride);
        }
    } else {
        constexpr int C = 16;
        RSD_OK_OR_RETURN(make_plan(a_len, b_len, n_pairs, C, 0, d_out, st, pv));
        const int stride = max_n > 32 * C ? (int)max_m : 0;
        if (bits == 2) {
            RSD_OK_OR_RETURN(persistent_grid(k_dist_gen<double, 2, C>, THREADS, sm_count, blocks));
            RSD_OK_OR_RETURN(scratch.ensure(sizeof(double) * (size_t)stride * blocks * wpb + 16));
            if (timing) RSD_CUDA(cudaEventRecord(ev0, st));
            k_dist_gen<double, 2, C><<<blocks, THREADS, 0, st>>>(pv, A, B, d_ic, d_fc, d_out, (double *)scratch.p, stride);
        } else {
            RSD_OK_OR_RETURN(persistent_grid(k_dist_gen<double, 4, C>, THREADS, sm_count, blocks));
            RSD_OK_OR_RETURN(scratch.ensure(sizeof(double) * (size_t)stride * blocks * wpb + 16));
            if (timing) RSD_CUDA(cudaEventRecord(ev0, st));
            k_dist_gen<double, 4, C><<<blocks, THREADS, 0, st>>>(pv, A, B, d_ic, d_fc, d_out, (double *)scratch.p, stride);
        }
    }
    if (timing) { RSD_CUDA(cudaEventRecord(ev1, st)); timed = true; }
    launches += 1;
    RSD_CUDA(cudaGetLastError());
    return RSD_OK;
}

extern "C" int rsd_distance_batch_dev(rsd_ctx *c,
                                      const uint32_t *a_words, const int64_t *a_start, const int32_t *a_len,
                                      const uint32_t *b_words, const int64_t *b_start, const int32_t *b_len,
                                      int64_t n_pairs, int64_t max_m, int64_t max_n, int bits, uint32_t symmask,
                                      int force_mode, double *out, int *mode_out, void *stream) {
    if (!c) return rsd_fail(RSD_EINVAL, "ctx is NULL");
    RSD_OK_OR_RETURN(c->ensure_device());
    return c->distance_dev(a_words, a_start, a_len, b_words, b_start, b_len, n_pairs, max_m, max_n, bits, symmask,
                           force_mode, out, mode_out, (cudaStream_t)stream);
}

static int64_t max_len(const int32_t *len, int64_t n) {
    int32_t m = 0;
    for (int64_t i = 0; i < n; ++i) m = std::max(m, len[i]);
    return m;
}

int rsd_ctx::upload_seqs(SeqBufs &sb, const uint32_t *words, const int64_t *start, const int32_t *len, int64_t n,
                         int64_t n_words, cudaStream_t st) {
    RSD_OK_OR_RETURN(sb.words.ensure(sizeof(uint32_t) * (size_t)(n_words + 8)));
    RSD_OK_OR_RETURN(sb.start.ensure(sizeof(int64_t) * (size_t)std::max<int64_t>(n, 1)));
    RSD_OK_OR_RETURN(sb.len.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1)));
    RSD_CUDA(cudaMemcpyAsync(sb.words.p, words, sizeof(uint32_t) * (size_t)n_words, cudaMemcpyHostToDevice, st));
    RSD_CUDA(cudaMemsetAsync((uint32_t *)sb.words.p + n_words, 0, sizeof(uint32_t) * 8, st));
    RSD_CUDA(cudaMemcpyAsync(sb.start.p, start, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    RSD_CUDA(cudaMemcpyAsync(sb.len.p, len, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    return RSD_OK;
}

extern "C" int rsd_distance_batch(rsd_ctx *c,
                                  const uint32_t *a_words, const int64_t *a_start, const int32_t *a_len, int64_t a_nwords,
                                  const uint32_t *b_words, const int64_t *b_start, const int32_t *b_len, int64_t b_nwords,
                                  int64_t n_pairs, int bits, uint32_t symmask, int force_mode,
                                  double *out, int *mode_out) {
    if (!c) return rsd_fail(RSD_EINVAL, "ctx is NULL");
    if (n_pairs < 0) return rsd_fail(RSD_EINVAL, "rsd_distance_batch: n_pairs < 0");
    if (n_pairs > 0 && (!a_words || !a_start || !a_len || !b_words || !b_start || !b_len || !out))
        return rsd_fail(RSD_EINVAL, "rsd_distance_batch: NULL buffer");
    RSD_OK_OR_RETURN(c->ensure_device());
    if (n_pairs == 0) return RSD_OK;
    cudaStream_t st = c->stream;
    RSD_OK_OR_RETURN(c->upload_seqs(c->bufA, a_words, a_start, a_len, n_pairs, a_nwords, st));
    RSD_OK_OR_RETURN(c->upload_seqs(c->bufB, b_words, b_start, b_len, n_pairs, b_nwords, st));
    RSD_OK_OR_RETURN(c->out_f64.ensure(sizeof(double) * (size_t)n_pairs));
    const int64_t max_m = max_len(a_len, n_pairs), max_n = max_len(b_len, n_pairs);
    RSD_OK_OR_RETURN(c->distance_dev((const uint32_t *)c->bufA.words.p, (const int64_t *)c->bufA.start.p,
                                     (const int32_t *)c->bufA.len.p, (const uint32_t *)c->bufB.words.p,
                                     (const int64_t *)c->bufB.start.p, (const int32_t *)c->bufB.len.p, n_pairs, max_m,
                                     max_n, bits, symmask, force_mode, (double *)c->out_f64.p, mode_out, st));
    RSD_CUDA(cudaMemcpyAsync(out, c->out_f64.p, sizeof(double) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
    RSD_CUDA(cudaStreamSynchronize(st));
    return RSD_OK;
}

// ------------------------------------------------------------------------------------------------
// one pair, whole matrix
// ------------------------------------------------------------------------------------------------
extern "C" int rsd_matrix(rsd_ctx *c, const uint8_t *a, int32_t m, const uint8_t *b, int32_t n,
                          double *values, uint8_t *mask) {
    if (!c) return rsd_fail(RSD_EINVAL, "ctx is NULL");
    if (m < 0 || n < 0 || !values || !mask || (m > 0 && !a) || (n > 0 && !b))
        return rsd_fail(RSD_EINVAL, "rsd_matrix: bad arguments");
    if ((int64_t)(m + 1) * (n + 1) > ((int64_t)1 << 31)) return rsd_fail(RSD_ERANGE, "rsd_matrix: matrix too large; use the batch entry points");
    RSD_OK_OR_RETURN(c->ensure_device());
    ModeInfo mi;
    RSD_OK_OR_RETURN(c->classify(0x7FFF, m, n, 4, RSD_MODE_F64, mi));
    cudaStream_t st = c->stream;
    RSD_OK_OR_RETURN(c->upload_costs(mi, st));
    const size_t cells = (size_t)(m + 1) * (n + 1);
    RSD_OK_OR_RETURN(c->mat_vals.ensure(sizeof(double) * cells));
    RSD_OK_OR_RETURN(c->mat_mask.ensure(cells));
    RSD_OK_OR_RETURN(c->mat_ab.ensure((size_t)m + n + 2));
    uint8_t *da = (uint8_t *)c->mat_ab.p, *db = da + m + 1;
    if (m) RSD_CUDA(cudaMemcpyAsync(da, a, m, cudaMemcpyHostToDevice, st));
    if (n) RSD_CUDA(cudaMemcpyAsync(db, b, n, cudaMemcpyHostToDevice, st));
    k_matrix_f64<<<1, 1024, 0, st>>>(da, m, db, n, c->d_fc, (double *)c->mat_vals.p, (uint8_t *)c->mat_mask.p);
    c->launches += 1;
    RSD_CUDA(cudaGetLastError());
    RSD_CUDA(cudaMemcpyAsync(values, c->mat_vals.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, st));
    RSD_CUDA(cudaMemcpyAsync(mask, c->mat_mask.p, cells, cudaMemcpyDeviceToHost, st));
    RSD_CUDA(cudaStreamSynchronize(st));
    return RSD_OK;
}

// ------------------------------------------------------------------------------------------------
// issue-rate microbenchmark
// ------------------------------------------------------------------------------------------------
extern "C" int rsd_ubench(rsd_ctx *c, int which, double *ops_per_s) {
    if (!c || !ops_per_s) return rsd_fail(RSD_EINVAL, "rsd_ubench: NULL argument");
    RSD_OK_OR_RETURN(c->ensure_device());
    cudaStream_t st = c->stream;
    RSD_OK_OR_RETURN(c->scratch.ensure(64));
    const int threads = 256, blocks = c->sm_count * 8, iters = 2000;
    const uint32_t y = 0x00010003u, z = 0x00070002u;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        RSD_CUDA(cudaEventRecord(c->ev0, st));
        switch (which) {
            case 0: k_ubench_u32<0><<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            case 1: k_ubench_u32<1><<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            case 2: k_ubench_u32<2><<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            case 3: k_ubench_u32<3><<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            case 4: k_ubench_f64<<<blocks, threads, 0, st>>>(iters, 1.5, (double *)c->scratch.p); break;
            case 5: k_ubench_u32<5><<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            case 6: k_ubench_u32<6><<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            case 7: k_ubench_mix<<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            case 8: k_ubench_u32<8><<<blocks, threads, 0, st>>>(iters, y, z, (uint32_t *)c->scratch.p); break;
            default: return rsd_fail(RSD_EINVAL, "rsd_ubench: unknown kind %d", which);
        }
        RSD_CUDA(cudaEventRecord(c->ev1, st));
        RSD_CUDA(cudaEventSynchronize(c->ev1));
        float ms = 0.f;
        RSD_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0) best = std::min(best, ms);
        c->launches += 1;
    }
    const double per_thread = (double)iters * RSD_UB_CHAINS * RSD_UB_REPS * (which == 7 ? 3.0 : 1.0);
    *ops_per_s = per_thread * threads * blocks / (best * 1e-3);
    return RSD_OK;
}

// ------------------------------------------------------------------------------------------------
// top-k merge after the gather (host; O(G*k) per query).  Key: score descending, global index
// ascending == the stable descending sort of performance.py:12-15 over the whole collection.
// ------------------------------------------------------------------------------------------------
extern "C" int rsd_topk_merge(const int64_t *idx, const double *score, int n_shards, int64_t n_queries, int k,
                              int64_t *out_idx, double *out_score) {
    if (!idx || !score || !out_idx || !out_score || n_shards < 1 || k < 1 || n_queries < 0)
        return rsd_fail(RSD_EINVAL, "rsd_topk_merge: bad arguments");
    std::vector<int> cur((size_t)n_shards);
    for (int64_t q = 0; q < n_queries; ++q) {
        std::fill(cur.begin(), cur.end(), 0);
        for (int r = 0; r < k; ++r) {
            int best = -1; int64_t bi = -1; double bs = 0;
            for (int g = 0; g < n_shards; ++g) {
                if (cur[g] >= k) continue;
                const size_t o = ((size_t)g * n_queries + q) * k + cur[g];
                const int64_t ci = idx[o];
                if (ci < 0) continue;                       // shard exhausted
                const double cs = score[o];
                if (best < 0 || cs > bs || (cs == bs && ci < bi)) { best = g; bi = ci; bs = cs; }
            }
            out_idx[q * k + r] = bi;
            out_score[q * k + r] = best < 0 ? 0.0 : bs;
            if (best >= 0) ++cur[best];
        }
    }
    return RSD_OK;
}

#include "rsd_stubs.cuh"
extern "C" int rsd_script_batch(rsd_ctx *, const uint32_t *, const int64_t *, const int32_t *, int64_t,
                                const uint32_t *, const int64_t *, const int32_t *, int64_t, int64_t, int, uint32_t, int,
                                int64_t, uint8_t *, int32_t *, int32_t *, int32_t *, double *, int *) { RSD_NOT_YET("rsd_script_batch"); }
extern "C" int rsd_patch_batch(rsd_ctx *, const uint8_t *, const int32_t *, const int32_t *, const int32_t *, int64_t,
                               const uint32_t *, const int64_t *, const int32_t *, int64_t,
                               const uint32_t *, const int64_t *, const int32_t *, int64_t,
                               const uint32_t *, const int64_t *, const int32_t *, int64_t,
                               int64_t, int, int64_t, uint8_t *, int32_t *, int32_t *) { RSD_NOT_YET("rsd_patch_batch"); }
extern "C" int rsd_script_patch_check_batch(rsd_ctx *, const uint32_t *, const int64_t *, const int32_t *, int64_t,
                                const uint32_t *, const int64_t *, const int32_t *, int64_t, int64_t, int, uint32_t, int,
                                int64_t, uint8_t *, int32_t *, int32_t *, int32_t *, double *, uint8_t *, int *) { RSD_NOT_YET("rsd_script_patch_check_batch"); }
extern "C" int rsd_db_load(rsd_ctx *, const uint32_t *, const int64_t *, const int32_t *, int64_t, int64_t, int, uint32_t, int64_t) { RSD_NOT_YET("rsd_db_load"); }
extern "C" int rsd_db_free(rsd_ctx *) { RSD_NOT_YET("rsd_db_free"); }
extern "C" int rsd_db_search_topk(rsd_ctx *, const uint32_t *, const int64_t *, const int32_t *, int64_t, int64_t, int, uint32_t,
                                  int, int, int64_t *, double *, double *, int *) { RSD_NOT_YET("rsd_db_search_topk"); }
extern "C" int rsd_db_search_topk_dev(rsd_ctx *, const uint32_t *, const int64_t *, const int32_t *, int64_t, int64_t, int,
                                      uint32_t, int, int, int64_t *, double *, int *, void *) { RSD_NOT_YET("rsd_db_search_topk_dev"); }
extern "C" int rsd_long_pair(rsd_ctx *, const uint8_t *, int64_t, const uint8_t *, int64_t, int, int, int64_t,
                             uint8_t *, int32_t *, int32_t *, int64_t *, double *, int *) { RSD_NOT_YET("rsd_long_pair"); }
