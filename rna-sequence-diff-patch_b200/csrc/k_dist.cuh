// k_dist.cuh — batched distance-only kernels (BASELINE config 2; also the scorer under wf_score).
//
// Systolic warp design: a pair's matrix is cut into strips of C columns; a lane keeps one row of
// its strip in C registers and the lanes of a group run one row apart, so the only inter-lane
// traffic is ONE __shfl_up per row (the strip's right-most cell).  Rows of the source string are
// streamed from the packed words through L1; the destination codes of a strip are turned into
// per-column selector registers once.  No shared-memory matrix, no HBM traffic besides the
// packed inputs (2 or 4 bit / symbol) and 8 B / pair of output.  Work arrives as warp tasks laid
// out on a tape of 32*P lane slots (k_plan.cuh).
//
// Replaces (reference): the double loop of wagnerFisher SED:185-222 + min_cost SED:92-128,
// read-out IR:439.
#pragma once
#include <type_traits>
#include "k_plan.cuh"

struct SeqView {
    const uint32_t *words;
    const int64_t *start;
    const int32_t *len;
};

// =============================================================================================
// Fast path: 2-bit codes (ACGU), scaled-int16 values, TWO pairs per register (hi/lo halves).
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
    __shared__ __align__(32) uint32_t s_tab[8];          // [0..3] v[a][.], [4..7] the transposed table (swapped pairs)
    if (threadIdx.x < 8) s_tab[threadIdx.x] = ic.rowtab4[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n_tasks = pv.totals[1];
    // two boundary columns per warp (ping-pong between passes)
    uint32_t *scr = scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 * scratch_stride;
    const double inv_scale = 1.0 / (double)(1 << ic.scale_log2);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_tab);

    for (;;) {
        int W = 0;
        if (lane == 0) W = atomicAdd(pv.work_counter, 1);
        W = __shfl_sync(RSD_FULL, W, 0);
        if (W >= n_tasks) break;
        const TapeTask tt = plan_decode(pv, W);
        int ns = tt.ns, P = tt.P;
        if (ns == 0) {                               // class of long pairs: one group, shape from the pair
            const int p0 = pv.groups[tt.gfirst].x;
            const int m0 = A.len[p0], n0 = B.len[p0];
            ns = ((plan_swap(m0, n0, C) ? m0 : n0) + C - 1) / C; P = (ns + 31) >> 5;
        }

        for (int pass = 0; pass < P; ++pass) {
            const LaneSlot ls = tape_slot(pv, tt, ns, pass, lane);
            // rows / columns after the orientation choice of the plan (plan_swap): the twins share m
            int m = 0, nA = 0, nB = 0, baseA = 0, baseB = 0;
            const uint32_t *awA = A.words, *awB = A.words, *bwA = B.words, *bwB = B.words;
            uint32_t sbaseA = sbase, sbaseB = sbase;
            if (ls.on) {
                const int mA = A.len[ls.pA], mB = A.len[ls.pB];
                nA = B.len[ls.pA]; nB = B.len[ls.pB];
                baseA = mA * ic.del + nA * ic.ins; baseB = mB * ic.del + nB * ic.ins;      // D = base - N either way
                awA = A.words + A.start[ls.pA]; awB = A.words + A.start[ls.pB];
                bwA = B.words + B.start[ls.pA]; bwB = B.words + B.start[ls.pB];
                m = mA;
                if (plan_swap(mA, nA, C)) { const uint32_t *t = awA; awA = bwA; bwA = t; m = nA; nA = mA; sbaseA = sbase + 16u; }
                if (plan_swap(mB, nB, C)) { const uint32_t *t = awB; awB = bwB; bwB = t; nB = mB; sbaseB = sbase + 16u; }
            }
            const int col0 = ls.s * C;
            uint32_t sel[C];
#pragma unroll
            for (int k = 0; k < C / 16; ++k) {
                uint32_t xa = ls.on ? __ldg(bwA + (col0 >> 4) + k) : 0u;
                uint32_t xb = ls.on ? __ldg(bwB + (col0 >> 4) + k) : 0u;
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
            const int steps = warp_max(ls.on ? m + ls.sk : 0);
            const bool has_in = __shfl_sync(RSD_FULL, (int)ls.from_scratch, 0) != 0;
            const bool has_out = __shfl_sync(RSD_FULL, (int)ls.to_scratch, 31) != 0;
            const uint32_t *scr_in = scr + ((pass & 1) ? scratch_stride : 0);
            uint32_t *scr_out = scr + ((pass & 1) ? 0 : scratch_stride);

            // Rows are consumed in blocks of 16 steps so the source codes of a block sit in one register
            // per pair: an unaligned 16-code window (funnel shift of two packed words), rotated so the
            // current code is at bits [3:2] and can be OR-ed into the shared-memory address of the
            // 4-entry row table.  Two copies of the loop: passes without a boundary hand-off (the common
            // case) run the lean one.
            const int mrow = ls.on ? m : 0;
            auto run_rows = [&](auto handoff_tag) {
                constexpr bool HANDOFF = decltype(handoff_tag)::value;
                int i = -ls.sk;
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
#pragma unroll 2
                    for (int k = 0; k < tend; ++k, ++i) {
                        uint32_t recv = __shfl_up_sync(RSD_FULL, last, 1);
                        if (ls.lead) recv = 0u;
                        const bool row_on = (unsigned)i < (unsigned)mrow;
                        if constexpr (HANDOFF) { if (ls.from_scratch && row_on) recv = scr_in[i]; }
                        uint32_t ra, rb;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ra) : "r"(sbaseA | (curA & 0xCu)));
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rb) : "r"(sbaseB | (curB & 0xCu)));
                        curA = __funnelshift_r(curA, curA, 2);
                        curB = __funnelshift_r(curB, curB, 2);
                        if (row_on) {
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
                            if constexpr (HANDOFF) { if (ls.to_scratch) scr_out[i] = last; }
                        }
                    }
                }
            };
            if (has_in || has_out) run_rows(std::true_type{}); else run_rows(std::false_type{});
            if (ls.on) {
                if (ls.s == (nA - 1) / C) {
                    const int cl = (nA - 1) - ls.s * C;
                    uint32_t v = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) if (c == cl) v = H[c];
                    out[ls.pA] = (double)(baseA - (int)(v & 0xffffu)) * inv_scale;
                }
                if (ls.hasB && ls.s == (nB - 1) / C) {
                    const int cl = (nB - 1) - ls.s * C;
                    uint32_t v = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) if (c == cl) v = H[c];
                    out[ls.pB] = (double)(baseB - (int)(v >> 16)) * inv_scale;
                }
            }
            __syncwarp();
        }
    }
}

// =============================================================================================
// General path: any symbols (2- or 4-bit packing), T = int32 (scaled, H' transform, table lookup
// of w from shared memory) or T = double (reference operation order, SED:95-109; borders are
// products SED:159,177; no FMA contraction — every add/mul is an explicit _rn intrinsic).
// =============================================================================================
template <typename T, int BITS, int C>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? 5 : 4))
k_dist_gen(PlanView pv, SeqView A, SeqView B, const IntCosts *__restrict__ icp,
           const F64Costs *__restrict__ fcp, double *__restrict__ out,
           T *__restrict__ scratch, int scratch_stride) {
    constexpr bool F64 = sizeof(T) == 8;
    constexpr int PER = 32 / BITS;               // symbols per packed word
    static_assert(C % PER == 0, "strip width must be a multiple of the packed word");
    // The 16x16 table is replicated once per bank (group): entry e of replica r lives at word-pair /
    // word e*REP + r, so lane l always reads bank (pair) l % REP and the per-cell lookups of a warp are
    // free of bank conflicts whatever symbols the lanes hold (one table would serialise 3-4 ways).
    constexpr int REP = 128 / (int)sizeof(T);    // 16 replicas of 8-byte entries, 32 of 4-byte entries
    __shared__ T s_w[256 * REP];
    for (int k = threadIdx.x; k < 256 * REP; k += blockDim.x) {
        const int e = k / REP;
        if constexpr (F64) s_w[k] = fcp->sub[e >> 4][e & 15];
        else s_w[k] = icp->w[e >> 4][e & 15];
    }
    __syncthreads();
    T c_ins, c_del;
    double inv_scale = 1.0;
    int i_ins = 0, i_del = 0;
    if constexpr (F64) { c_ins = fcp->ins; c_del = fcp->del; }
    else { c_ins = 0; c_del = 0; i_ins = icp->ins; i_del = icp->del; inv_scale = 1.0 / (double)(1 << icp->scale_log2); }
    const int lane = threadIdx.x & 31;
    const int n_tasks = pv.totals[1];
    T *scr = scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 * scratch_stride;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_w);

    for (;;) {
        int W = 0;
        if (lane == 0) W = atomicAdd(pv.work_counter, 1);
        W = __shfl_sync(RSD_FULL, W, 0);
        if (W >= n_tasks) break;
        const TapeTask tt = plan_decode(pv, W);
        int ns = tt.ns, P = tt.P;
        if (ns == 0) {
            const int p0 = pv.groups[tt.gfirst].x;
            ns = (B.len[p0] + C - 1) / C; P = (ns + 31) >> 5;
        }

        for (int pass = 0; pass < P; ++pass) {
            const LaneSlot ls = tape_slot(pv, tt, ns, pass, lane);
            int m = 0, n = 0;
            const uint32_t *aw = A.words, *bw = B.words;
            if (ls.on) {
                m = A.len[ls.pA]; n = B.len[ls.pA];
                aw = A.words + A.start[ls.pA]; bw = B.words + B.start[ls.pA];
            }
            const int s = ls.s;
            const int col0 = s * C;
            int bc[C];                            // shared-memory byte address of each column's table column
#pragma unroll
            for (int k = 0; k < C / PER; ++k) {
                uint32_t x = ls.on ? __ldg(bw + col0 / PER + k) : 0u;
#pragma unroll
                for (int c = 0; c < PER; ++c)
                    bc[k * PER + c] = (int)(sbase + (((x >> (BITS * c)) & ((1u << BITS) - 1u)) * REP + (lane % REP)) * (uint32_t)sizeof(T));
            }
            T H[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if constexpr (F64) H[c] = __dmul_rn((double)(col0 + c + 1), c_ins);   // SED:159
                else H[c] = 0;
            }
            T last = 0, prev_recv = 0;
            uint32_t cur = 0u;
            const int steps = warp_max(ls.on ? m + ls.sk : 0);
            const bool has_in = __shfl_sync(RSD_FULL, (int)ls.from_scratch, 0) != 0;
            const bool has_out = __shfl_sync(RSD_FULL, (int)ls.to_scratch, 31) != 0;
            const T *scr_in = scr + ((pass & 1) ? scratch_stride : 0);
            T *scr_out = scr + ((pass & 1) ? 0 : scratch_stride);

            auto run_rows = [&](auto handoff_tag) {          // two copies: passes without a hand-off run the lean one
            constexpr bool HANDOFF = decltype(handoff_tag)::value;
#pragma unroll (F64 ? 1 : 2)
            for (int t = 0; t < steps; ++t) {
                T recv = __shfl_up_sync(RSD_FULL, last, 1);
                const int i = t - ls.sk;
                const bool row_on = ls.on && (unsigned)i < (unsigned)m;
                if (ls.lead) {
                    if constexpr (F64) recv = __dmul_rn((double)(i + 1), c_del);          // SED:177
                    else recv = 0;
                }
                if constexpr (HANDOFF) { if (ls.from_scratch && row_on) recv = scr_in[i]; }
                if (row_on) {
                    if (i % PER == 0) cur = __ldg(aw + i / PER);
                    const uint32_t rowoff = ((cur & ((1u << BITS) - 1u)) << 4) * (uint32_t)(REP * sizeof(T)); cur >>= BITS;
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
                            // SED:106-107 min(c0, c1, c2): the value of a min does not depend on the association, so
                            // the two candidates that do not involve the left neighbour are combined first and only
                            // DADD + one compare-select sit on the row's dependency chain
                            H[c] = dmin2(c0, dmin2(c1, c2));
                        } else {
                            const int t2 = addmin32(diag, w, H[c]);
                            diag = H[c];
                            H[c] = min(t2, left);
                        }
                        left = H[c];
                    }
                    last = left; prev_recv = recv;
                    if constexpr (HANDOFF) { if (ls.to_scratch) scr_out[i] = last; }
                }
            }
            };
            if (has_in || has_out) run_rows(std::true_type{}); else run_rows(std::false_type{});
            if (ls.on && s == (n - 1) / C) {
                const int cl = (n - 1) - s * C;
                T res = 0;
#pragma unroll
                for (int c = 0; c < C; ++c) if (c == cl) res = H[c];
                if constexpr (F64) out[ls.pA] = res;
                else out[ls.pA] = (double)(res + m * i_del + n * i_ins) * inv_scale;
            }
            __syncwarp();
        }
    }
}
