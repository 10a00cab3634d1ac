// k_dist.cuh — batched distance-only kernels (BASELINE config 2; also the scorer under wf_score).
//
// Systolic warp design: a pair's matrix is cut into strips of C columns; lane s of a group keeps
// row i of strip s in C registers and works on row (t - s) at step t, so the only inter-lane
// traffic is ONE __shfl_up per row (the strip's right-most cell).  Rows of the source string are
// streamed from the packed words through L1; the destination codes of a strip are turned into
// per-column selector registers once.  No shared-memory matrix, no HBM traffic besides the
// packed inputs (2 or 4 bit / symbol) and 8 B / pair of output.
//
// Replaces (reference): the double loop of wagnerFisher SED:185-222 + min_cost SED:92-128,
// read-out IR:439.
#pragma once
#include "k_plan.cuh"

struct SeqView {
    const uint32_t *words;
    const int64_t *start;
    const int32_t *len;
};

// =============================================================================================
// Fast path: 2-bit codes (ACGU), scaled-int16 H' values, TWO pairs per register (hi/lo halves).
// Max form: N = -H' >= 0 and v = -w = ins + del - sub clamped to >= 0 (a substitution dearer than
// delete+insert never wins, so the clamp leaves every value unchanged):
//     N[i][j] = max(N[i][j-1], N[i-1][j], N[i-1][j-1] + v(a_i, b_j)),   D = m*del + n*ins - N.
// Per packed cell (= 2 matrix cells): PRMT (v lookup for both pairs, alu pipe), one plain 32-bit add
// (halves are non-negative and < 2^16, so no carry crosses; issued as IMAD on the fma pipe) and
// VIMNMX3.U16x2 (alu pipe): 2 alu + 1 fma instructions per two cells.
// =============================================================================================
template <int C>
__global__ void __launch_bounds__(128)
k_dist_twin16(PlanView pv, SeqView A, SeqView B, IntCosts ic, double *__restrict__ out,
              uint32_t *__restrict__ scratch, int scratch_stride, uint32_t one) {
    static_assert(C % 16 == 0, "strip width must be a multiple of the 2-bit word");
    __shared__ __align__(16) uint32_t s_tab[4];
    if (threadIdx.x < 4) s_tab[threadIdx.x] = ic.rowtab4[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n_warps = pv.totals[1];
    uint32_t *scr = scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * scratch_stride;
    const double inv_scale = 1.0 / (double)(1 << ic.scale_log2);

    for (;;) {
        int W = 0;
        if (lane == 0) W = atomicAdd(pv.work_counter, 1);
        W = __shfl_sync(RSD_FULL, W, 0);
        if (W >= n_warps) break;
        const WarpTask tk = plan_decode(pv, W, lane);
        int m = 0, nA = 0, nB = 0;
        const uint32_t *awA = A.words, *awB = A.words, *bwA = B.words, *bwB = B.words;
        if (tk.on) {
            m = A.len[tk.pA]; nA = B.len[tk.pA]; nB = B.len[tk.pB];
            awA = A.words + A.start[tk.pA]; awB = A.words + A.start[tk.pB];
            bwA = B.words + B.start[tk.pA]; bwB = B.words + B.start[tk.pB];
        }
        const int nmax = max(nA, nB);
        const int ns = (nmax + C - 1) / C;
        const int npass = tk.multi ? __shfl_sync(RSD_FULL, (ns + 31) >> 5, 0) : 1;
        int resA = 0, resB = 0;

        for (int pass = 0; pass < npass; ++pass) {
            const int s = pass * 32 + tk.s0;
            const bool strip_on = tk.on && s < ns;
            const int col0 = s * C;
            uint32_t sel[C];
#pragma unroll
            for (int k = 0; k < C / 16; ++k) {
                uint32_t xa = strip_on ? __ldg(bwA + (col0 >> 4) + k) : 0u;
                uint32_t xb = strip_on ? __ldg(bwB + (col0 >> 4) + k) : 0u;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    uint32_t ca = (xa >> (2 * c)) & 3u, cb = (xb >> (2 * c)) & 3u;
                    // nibbles: [byte ca of ra][sign of it][byte cb of rb][sign of it]
                    sel[k * 16 + c] = 0xC480u + ca * 0x11u + cb * 0x1100u;
                }
            }
            uint32_t H[C];
#pragma unroll
            for (int c = 0; c < C; ++c) H[c] = 0u;
            uint32_t last = 0u, prev_recv = 0u, curA = 0u, curB = 0u;
            const int steps = warp_max(strip_on ? m + tk.s0 : 0);

            if (!tk.multi) {
                // ---- single pass (n <= 32*C): the hot loop.  Rows are consumed in blocks of 16 steps so
                // the source codes of a block sit in one register per pair: an unaligned 16-code window
                // (funnel shift of two packed words), rotated so the current code is at bits [3:2] and can
                // be OR-ed into the shared-memory address of the 4-entry row table.
                const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_tab);
                const int mrow = strip_on ? m : 0;
                const bool lead = tk.s0 == 0;
                int i = -tk.s0;
#pragma unroll 1
                for (int t0 = 0; t0 < steps; t0 += 16) {
                    {
                        const int wi = i >> 4, bit = (i & 15) * 2;            // floor division also for i < 0
                        const int w0 = max(wi, 0), w1 = max(wi + 1, 0);
                        const uint32_t a0 = __ldg(awA + w0), a1 = __ldg(awA + w1);
                        const uint32_t b0 = __ldg(awB + w0), b1 = __ldg(awB + w1);
                        curA = __funnelshift_l(__funnelshift_r(a0, a1, bit), __funnelshift_r(a0, a1, bit), 2);
                        curB = __funnelshift_l(__funnelshift_r(b0, b1, bit), __funnelshift_r(b0, b1, bit), 2);
                    }
                    const int tend = min(16, steps - t0);
#pragma unroll 1
                    for (int k = 0; k < tend; ++k, ++i) {
                        uint32_t recv = __shfl_up_sync(RSD_FULL, last, 1);
                        if (lead) recv = 0u;
                        uint32_t ra, rb;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ra) : "r"(sbase | (curA & 0xCu)));
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rb) : "r"(sbase | (curB & 0xCu)));
                        curA = __funnelshift_r(curA, curA, 2);
                        curB = __funnelshift_r(curB, curB, 2);
                        if ((unsigned)i < (unsigned)mrow) {
                            uint32_t left = recv, diag = prev_recv;
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const uint32_t v = prmt(ra, rb, sel[c]);
                                const uint32_t x = add_fma(v, diag, one);
                                diag = H[c];
                                H[c] = max3u16x2(x, H[c], left);
                                left = H[c];
                            }
                            last = left; prev_recv = recv;
                        }
                    }
                }
            } else {
            const bool wr_scr = tk.s0 == 31 && pass + 1 < npass;
#pragma unroll 1
            for (int t = 0; t < steps; ++t) {
                uint32_t recv = __shfl_up_sync(RSD_FULL, last, 1);
                const int i = t - tk.s0;
                const bool row_on = strip_on && (unsigned)i < (unsigned)m;
                if (tk.s0 == 0) recv = (pass > 0 && row_on) ? scr[i] : 0u;
                if (row_on) {
                    if ((i & 15) == 0) { curA = __ldg(awA + (i >> 4)); curB = __ldg(awB + (i >> 4)); }
                    const uint32_t ra = s_tab[curA & 3u]; curA >>= 2;
                    const uint32_t rb = s_tab[curB & 3u]; curB >>= 2;
                    uint32_t left = recv, diag = prev_recv;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const uint32_t v = prmt(ra, rb, sel[c]);
                        const uint32_t x = add_fma(v, diag, one);
                        diag = H[c];
                        H[c] = max3u16x2(x, H[c], left);
                        left = H[c];
                    }
                    last = left; prev_recv = recv;
                    if (wr_scr) scr[i] = last;
                }
            }
            }
            if (strip_on) {
                if (s == (nA - 1) / C) {
                    const int cl = (nA - 1) - s * C;
                    uint32_t v = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) if (c == cl) v = H[c];
                    resA = (int)(v & 0xffffu);
                }
                if (s == (nB - 1) / C) {
                    const int cl = (nB - 1) - s * C;
                    uint32_t v = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) if (c == cl) v = H[c];
                    resB = (int)(v >> 16);
                }
            }
            __syncwarp();
        }
        if (tk.on) {
            const int sA = tk.multi ? ((nA - 1) / C) & 31 : (nA - 1) / C;
            if (tk.s0 == sA)
                out[tk.pA] = (double)(m * ic.del + nA * ic.ins - resA) * inv_scale;
            const int sB = tk.multi ? ((nB - 1) / C) & 31 : (nB - 1) / C;
            if (tk.hasB && tk.s0 == sB)
                out[tk.pB] = (double)(m * ic.del + nB * ic.ins - resB) * inv_scale;
        }
    }
}

// =============================================================================================
// General path: any symbols (2- or 4-bit packing), T = int32 (scaled, H' transform, table lookup
// of w from shared memory) or T = double (reference operation order, SED:95-109; borders are
// products SED:159,177; no FMA contraction — every add/mul is an explicit _rn intrinsic).
// =============================================================================================
template <typename T> struct GenTab;
template <> struct GenTab<int> { const IntCosts *c; };
template <> struct GenTab<double> { const F64Costs *c; };

template <typename T, int BITS, int C>
__global__ void __launch_bounds__(128)
k_dist_gen(PlanView pv, SeqView A, SeqView B, const IntCosts *__restrict__ icp,
           const F64Costs *__restrict__ fcp, double *__restrict__ out,
           T *__restrict__ scratch, int scratch_stride) {
    constexpr bool F64 = sizeof(T) == 8;
    constexpr int PER = 32 / BITS;               // symbols per packed word
    static_assert(C % PER == 0, "strip width must be a multiple of the packed word");
    __shared__ T s_w[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        if constexpr (F64) s_w[k] = fcp->sub[k >> 4][k & 15];
        else s_w[k] = icp->w[k >> 4][k & 15];
    }
    __syncthreads();
    T c_ins, c_del;
    double inv_scale = 1.0;
    int i_ins = 0, i_del = 0;
    if constexpr (F64) { c_ins = fcp->ins; c_del = fcp->del; }
    else { c_ins = 0; c_del = 0; i_ins = icp->ins; i_del = icp->del; inv_scale = 1.0 / (double)(1 << icp->scale_log2); }
    const int lane = threadIdx.x & 31;
    const int n_warps = pv.totals[1];
    T *scr = scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * scratch_stride;

    for (;;) {
        int W = 0;
        if (lane == 0) W = atomicAdd(pv.work_counter, 1);
        W = __shfl_sync(RSD_FULL, W, 0);
        if (W >= n_warps) break;
        const WarpTask tk = plan_decode(pv, W, lane);
        int m = 0, n = 0;
        const uint32_t *aw = A.words, *bw = B.words;
        if (tk.on) {
            m = A.len[tk.pA]; n = B.len[tk.pA];
            aw = A.words + A.start[tk.pA]; bw = B.words + B.start[tk.pA];
        }
        const int ns = (n + C - 1) / C;
        const int npass = tk.multi ? __shfl_sync(RSD_FULL, (ns + 31) >> 5, 0) : 1;
        T res = 0;

        for (int pass = 0; pass < npass; ++pass) {
            const int s = pass * 32 + tk.s0;
            const bool strip_on = tk.on && s < ns;
            const int col0 = s * C;
            int bc[C];                            // destination code of each column
#pragma unroll
            for (int k = 0; k < C / PER; ++k) {
                uint32_t x = strip_on ? __ldg(bw + col0 / PER + k) : 0u;
#pragma unroll
                for (int c = 0; c < PER; ++c) bc[k * PER + c] = (x >> (BITS * c)) & ((1u << BITS) - 1u);
            }
            {   // keep each column's shared-memory byte address; the row offset is added per row (one add per lookup)
                const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_w);
#pragma unroll
                for (int c = 0; c < C; ++c) bc[c] = (int)(sbase + (uint32_t)bc[c] * (uint32_t)sizeof(T));
            }
            T H[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if constexpr (F64) H[c] = __dmul_rn((double)(col0 + c + 1), c_ins);   // SED:159
                else H[c] = 0;
            }
            T last = 0, prev_recv = 0;
            uint32_t cur = 0u;
            const int steps = warp_max(strip_on ? m + tk.s0 : 0);
            const bool wr_scr = tk.multi && tk.s0 == 31 && pass + 1 < npass;

#pragma unroll 1
            for (int t = 0; t < steps; ++t) {
                T recv = __shfl_up_sync(RSD_FULL, last, 1);
                const int i = t - tk.s0;
                const bool row_on = strip_on && (unsigned)i < (unsigned)m;
                if (tk.s0 == 0) {
                    if (pass > 0) recv = row_on ? scr[i] : (T)0;
                    else if constexpr (F64) recv = __dmul_rn((double)(i + 1), c_del);     // SED:177
                    else recv = 0;
                }
                if (row_on) {
                    if (i % PER == 0) cur = __ldg(aw + i / PER);
                    const uint32_t rowoff = ((cur & ((1u << BITS) - 1u)) << 4) * (uint32_t)sizeof(T); cur >>= BITS;
                    T left = recv, diag = prev_recv;
                    if constexpr (F64) {
                        if (i == 0) diag = (s == 0) ? 0.0 : __dmul_rn((double)col0, c_ins);   // row-0 border
                        else if (s == 0) diag = __dmul_rn((double)i, c_del);
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        T w;
                        if constexpr (F64) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(w) : "r"((uint32_t)bc[c] + rowoff));
                        else asm volatile("ld.shared.s32 %0, [%1];" : "=r"(w) : "r"((uint32_t)bc[c] + rowoff));
                        if constexpr (F64) {
                            const double c0 = __dadd_rn(left, c_ins);      // SED:95
                            const double c1 = __dadd_rn(H[c], c_del);      // SED:97
                            const double c2 = __dadd_rn(diag, w);          // SED:99
                            diag = H[c];
                            H[c] = dmin2(dmin2(c0, c1), c2);                 // SED:106-107
                        } else {
                            const int t2 = addmin32(diag, w, H[c]);
                            diag = H[c];
                            H[c] = min(t2, left);
                        }
                        left = H[c];
                    }
                    last = left; prev_recv = recv;
                    if (wr_scr) scr[i] = last;
                }
            }
            if (strip_on && s == (n - 1) / C) {
                const int cl = (n - 1) - s * C;
#pragma unroll
                for (int c = 0; c < C; ++c) if (c == cl) res = H[c];
            }
            __syncwarp();
        }
        if (tk.on) {
            const int sl = tk.multi ? ((n - 1) / C) & 31 : (n - 1) / C;
            if (tk.s0 == sl) {
                if constexpr (F64) out[tk.pA] = res;
                else out[tk.pA] = (double)(res + m * i_del + n * i_ins) * inv_scale;
            }
        }
    }
}
