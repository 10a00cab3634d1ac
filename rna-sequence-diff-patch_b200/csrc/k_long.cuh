// k_long.cuh — one long pair (>= ~10 kb per side; BASELINE config 4: 50 kb x 50 kb with traceback).
//
// Block-tiled diagonal wavefront: the matrix is cut into column panels of 32*C columns; panel w is
// owned by warp w (one warp per CTA so panels spread over all SMs) and, inside the panel, lane s
// owns C columns in registers — the same systolic row pipeline as the batched kernels.  Panels run
// concurrently as a second-level pipeline: the right-most column of panel w streams through a
// global boundary array to panel w+1, which follows ~48 rows behind (progress counter with
// release/acquire fences).  All CTAs are co-resident (cooperative launch), so the waits cannot
// deadlock.  Directions are 2 bit/cell, accumulated 16 rows per register and written with 128-bit
// stores; the traceback walks them from (m,n).
//
// Keys are (cost, steps) folded into one int64 (cost << S | steps) in the H' form of k_script.cuh
// — int32 would overflow at 50 kb — or (fp64 cost, int steps) when the costs are not dyadic.
// Replaces (reference): wagnerFisher + create_paths(dp)[0] + generate_es for a pair the reference
// cannot hold in memory (~500 B per cell).
#pragma once
#include "k_script.cuh"

struct LongArgs {
    const uint8_t *a; int m;           // 1 byte / symbol codes on the device
    const uint8_t *b; int n;
    int n_panels, n_pad;               // n_pad = n_panels * 32 * C
    uint32_t *dirs;                    // [ceil(m/16)][n_pad]
    void *bound;                       // [n_panels][m] keys (int64 or double)
    int *bound_steps;                  // [n_panels][m] (fp64 mode)
    int *progress;                     // [n_panels] rows published
    double *dist;
    int S;
};

template <bool F64, int C>
__global__ void __launch_bounds__(32)
k_long_fwd(LongArgs la, const IntCosts *__restrict__ icp, const F64Costs *__restrict__ fcp) {
    using T = typename std::conditional<F64, double, long long>::type;
    __shared__ T s_w[256];
    for (int k = threadIdx.x; k < 256; k += 32) {
        if constexpr (F64) s_w[k] = fcp->sub[k >> 4][k & 15];
        else s_w[k] = ((long long)icp->w[k >> 4][k & 15] << la.S) - 1;            // (w << S) - 1
    }
    __syncwarp();
    const int lane = threadIdx.x;
    const int w = blockIdx.x;                       // panel
    const int m = la.m, n = la.n;
    const int col0 = (w * 32 + lane) * C;
    const bool strip_on = col0 < n;
    T c_ins = 0, c_del = 0;
    if constexpr (F64) { c_ins = fcp->ins; c_del = fcp->del; }

    int bc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) bc[c] = (col0 + c < n) ? la.b[col0 + c] : 0;
    T H[C]; int HS[F64 ? C : 1]; uint32_t acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        acc[c] = 0u;
        if constexpr (F64) { H[c] = __dmul_rn((double)(col0 + c + 1), c_ins); HS[c] = col0 + c + 1; }
        else H[c] = 0;
    }
    T last = 0, prev_recv = 0; int last_s = 0, prev_recv_s = 0;
    T *bin = w > 0 ? (T *)la.bound + (size_t)(w - 1) * m : nullptr;
    T *bout = (T *)la.bound + (size_t)w * m;
    int *bin_s = (F64 && w > 0) ? la.bound_steps + (size_t)(w - 1) * m : nullptr;
    int *bout_s = F64 ? la.bound_steps + (size_t)w * m : nullptr;
    volatile int *prog_in = w > 0 ? la.progress + (w - 1) : nullptr;
    const bool publish = (w + 1 < la.n_panels);
    uint32_t *dcol = la.dirs + col0;
    const int steps = m + 31;
    int avail = w > 0 ? 0 : m;                      // rows of the left boundary known to be published

#pragma unroll 1
    for (int t = 0; t < steps; ++t) {
        // lane 0 consumes boundary row t: make sure the producer panel has published it
        if (w > 0 && t < m && t >= avail) {
            if (lane == 0) {
                int p;
                do { p = *prog_in; } while (p <= t);
                avail = p;
            }
            avail = __shfl_sync(RSD_FULL, avail, 0);
            __threadfence();
        }
        T recv = __shfl_up_sync(RSD_FULL, last, 1);
        int recv_s = 0;
        if constexpr (F64) recv_s = __shfl_up_sync(RSD_FULL, last_s, 1);
        const int i = t - lane;
        const bool row_on = strip_on && (unsigned)i < (unsigned)m;
        if (lane == 0) {
            if (w > 0) {
                if (row_on) { recv = __ldcg(bin + i); if constexpr (F64) recv_s = __ldcg(bin_s + i); }
            } else if constexpr (F64) { recv = __dmul_rn((double)(i + 1), c_del); recv_s = i + 1; }
            else recv = 0;
        }
        if (row_on) {
            const int rowbase = (int)la.a[i] << 4;
            T left = recv, diag = prev_recv; int left_s = recv_s, diag_s = prev_recv_s;
            if constexpr (F64) {
                if (i == 0) { diag = (col0 == 0) ? 0.0 : __dmul_rn((double)col0, c_ins); diag_s = col0; }
                else if (col0 == 0) { diag = __dmul_rn((double)i, c_del); diag_s = i; }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const T wv = s_w[rowbase + bc[c]];
                uint32_t code;
                if constexpr (F64) {
                    const double c0 = __dadd_rn(left, c_ins), c1 = __dadd_rn(H[c], c_del), c2 = __dadd_rn(diag, wv);
                    const int s0 = left_s + 1, s1 = HS[c] + 1, s2 = diag_s + 1;
                    const double v = fmin(fmin(c0, c1), c2);
                    int bs = (c0 == v) ? s0 : 0x7fffffff; code = 0u;
                    if (c1 == v && s1 < bs) { bs = s1; code = 1u; }
                    if (c2 == v && s2 < bs) { bs = s2; code = 2u; }
                    diag = H[c]; diag_s = HS[c]; H[c] = v; HS[c] = bs; left = v; left_s = bs;
                } else {
                    const long long x = diag + wv;
                    const bool p_del = H[c] <= x;                 // DEL before UPD
                    const long long t2 = p_del ? H[c] : x;
                    const bool p_ins = left <= t2;                // INS first
                    diag = H[c];
                    H[c] = p_ins ? left : t2;
                    left = H[c];
                    code = p_ins ? 0u : (p_del ? 1u : 2u);
                }
                acc[c] = __funnelshift_r(acc[c], code, 2);
            }
            last = left; prev_recv = recv;
            if constexpr (F64) { last_s = left_s; prev_recv_s = recv_s; }
            if ((i & 15) == 15 || i == m - 1) {
                const int sh = 2 * (15 - (i & 15));
                uint4 *dst = reinterpret_cast<uint4 *>(dcol + (size_t)(i >> 4) * la.n_pad);
#pragma unroll
                for (int c = 0; c < C; c += 4)
                    dst[c >> 2] = make_uint4(acc[c] >> sh, acc[c + 1] >> sh, acc[c + 2] >> sh, acc[c + 3] >> sh);
            }
        }
        // lane 31 finished row i31 = t - 31: publish it for the next panel, 16 rows at a time
        if (publish && lane == 31) {
            const int i31 = t - 31;
            if (i31 >= 0 && i31 < m) {
                bout[i31] = last; if constexpr (F64) bout_s[i31] = last_s;
                if ((i31 & 15) == 15 || i31 == m - 1) { __threadfence(); *(volatile int *)(la.progress + w) = i31 + 1; }
            }
        }
    }
    // the cell (m, n) lives in panel (n-1)/(32C), lane ((n-1)/C)%32, column (n-1)%C
    if (strip_on && col0 <= n - 1 && n - 1 < col0 + C) {
        const int cl = (n - 1) - col0;
        T res = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) if (c == cl) res = H[c];
        if constexpr (F64) la.dist[0] = res;
        else {
            const long long key = res + (long long)m * (((long long)icp->del << la.S) + 1) + (long long)n * (((long long)icp->ins << la.S) + 1);
            la.dist[0] = (double)(key >> la.S) / (double)(1 << icp->scale_log2);
        }
    }
}

// one thread: walk the direction words; ops written sink->origin from the end of tmp[0 .. m+n)
__global__ void k_long_traceback(int m, int n, const uint32_t *__restrict__ dirs, int n_pad,
                                 uint8_t *__restrict__ tmp, int32_t *__restrict__ n_ops) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t pos = (int64_t)m + n;
    int i = m, j = n;
    while (i > 0 && j > 0) {
        const uint32_t wv = dirs[(size_t)((i - 1) >> 4) * n_pad + (j - 1)];
        const uint32_t code = (wv >> (2 * ((i - 1) & 15))) & 3u;
        tmp[--pos] = (uint8_t)code;
        if (code == 0u) --j; else if (code == 1u) --i; else { --i; --j; }
    }
    while (j > 0) { tmp[--pos] = 0; --j; }
    while (i > 0) { tmp[--pos] = 1; --i; }
    n_ops[0] = (int32_t)((int64_t)m + n - pos);
}

// packed script of the single pair: op + entered cell via two block scans (same as k_finalize, unpacked input)
__global__ void __launch_bounds__(1024) k_long_emit(const uint8_t *__restrict__ tmp, int m, int n, const int32_t *__restrict__ n_ops,
                                                    uint8_t *__restrict__ op, int32_t *__restrict__ oi, int32_t *__restrict__ oj) {
    __shared__ int s_ai[1024], s_bj[1024];
    __shared__ int carry_i, carry_j;
    const int tid = threadIdx.x;
    const int k_ops = n_ops[0];
    const uint8_t *src = tmp + ((int64_t)m + n - k_ops);
    if (tid == 0) { carry_i = 0; carry_j = 0; }
    __syncthreads();
    for (int base = 0; base < k_ops; base += 1024) {
        const int k = base + tid;
        const int o = k < k_ops ? src[k] : 3;
        s_ai[tid] = (o == 1 || o == 2); s_bj[tid] = (o == 0 || o == 2);
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int xi = tid >= off ? s_ai[tid - off] : 0, xj = tid >= off ? s_bj[tid - off] : 0;
            __syncthreads();
            s_ai[tid] += xi; s_bj[tid] += xj;
            __syncthreads();
        }
        const int vi = carry_i + s_ai[tid], vj = carry_j + s_bj[tid];
        if (k < k_ops) { op[k] = (uint8_t)o; if (oi) oi[k] = vi; if (oj) oj[k] = vj; }
        __syncthreads();
        if (tid == 1023) { carry_i = vi; carry_j = vj; }
        __syncthreads();
    }
}
