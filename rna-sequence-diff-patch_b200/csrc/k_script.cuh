// k_script.cuh — forward pass with 2-bit direction codes, traceback, and the finalize kernel that
// turns the traced path into the packed edit script, applies it (prefix-sum scatter patch) and
// checks the round trip on the device (BASELINE config 3).
//
// Replaces (reference): wagnerFisher SED:133-224 + create_paths(dp)[0] SED:228-271 +
// generate_es SED:274-334 + patching SED:380-457 for generated scripts.
//
// Canonical rule (verified against the reference): among the predecessors that tie on
// cost, take the one with the fewest edges from the origin, then the first of INS, DEL, UPD.
// Integer mode folds both keys into one word, key = cost * 2^S + steps (steps < 2^S), and runs the
// same H' recurrence as the distance kernels on keys:
//     key'[i][j] = min(key'[i][j-1], key'[i-1][j], key'[i-1][j-1] + ((w << S) - 1))
// where the "- 1" is the diagonal edge saving one step against INS+DEL.  Ties are resolved INS,
// DEL, UPD by the order of the two min instructions (VIMNMX with predicate = __vibmin_s32).
//
// Direction storage: dirs[pair][rb][col] is one u32 holding rows 16*rb .. 16*rb+15 of column col,
// two bits per row pushed in from the right (row r sits at bits 2*(15 - (r & 15)) + {1,0});
// col < n_pad = strips * C.  Bit 1 = "not INS" (the left candidate lost), bit 0 = "UPD rather than DEL"
// (the diagonal candidate beat the upper one): 0x -> INS, 10 -> DEL, 11 -> UPD.  In the integer kernel
// both bits are sign bits of differences the recurrence already has (t2 - left, t2 - up), shifted in
// with one funnel shift each — no compare, no select.
#pragma once
#include <type_traits>
#include "k_dist.cuh"

// direction word -> op code (0 INS, 1 DEL, 2 UPD) of matrix row `row` (0-based interior row)
__device__ __forceinline__ uint32_t dir_decode(uint32_t word, int row) {
    const uint32_t bits = (word >> (2 * (15 - (row & 15)))) & 3u;
    return bits < 2u ? 0u : bits - 1u;
}

struct ScriptView {
    uint32_t *dirs;            // chunk-local direction words
    const int64_t *dir_off;    // [n_pairs] word offset of each pair inside dirs
    double *dist;              // [n_pairs]
};

template <bool F64, int BITS, int C>
__global__ void __launch_bounds__(128)
k_script_fwd(PlanView pv, SeqView A, SeqView B, const IntCosts *__restrict__ icp,
             const F64Costs *__restrict__ fcp, ScriptView sv, int S,
             void *__restrict__ scratch_v, int scratch_stride, int minus_one) {
    using T = typename std::conditional<F64, double, int>::type;
    constexpr int PER = 32 / BITS;
    static_assert(C % PER == 0 && C % 4 == 0, "strip width");
    // table replicated per bank (group) so the per-cell lookups never conflict (see k_dist_gen)
    constexpr int REP = 128 / (int)sizeof(T);
    __shared__ T s_w[256 * REP];
    for (int k = threadIdx.x; k < 256 * REP; k += blockDim.x) {
        const int e = k / REP;
        if constexpr (F64) s_w[k] = fcp->sub[e >> 4][e & 15];
        else s_w[k] = (int)(((unsigned)icp->w[e >> 4][e & 15] << S) - 1u);     // (w << S) - 1
    }
    __syncthreads();
    T c_ins = 0, c_del = 0;
    int i_ins = 0, i_del = 0;
    double inv_scale = 1.0;
    if constexpr (F64) { c_ins = fcp->ins; c_del = fcp->del; }
    else { i_ins = icp->ins; i_del = icp->del; inv_scale = 1.0 / (double)(1 << icp->scale_log2); }
    const int lane = threadIdx.x & 31;
    const int n_tasks = pv.totals[1];
    // per warp: two boundary columns (ping-pong between passes); fp64 mode keeps the step counts in a
    // second slab behind all the cost columns
    const size_t wslot = (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 * scratch_stride;
    T *scr = (T *)scratch_v + wslot;
    int *scr_steps = nullptr;
    if constexpr (F64)
        scr_steps = (int *)((T *)scratch_v + (size_t)gridDim.x * (blockDim.x >> 5) * 2 * scratch_stride) + wslot;
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
        const int n_pad = ns * C;

        for (int pass = 0; pass < P; ++pass) {
            const LaneSlot ls = tape_slot(pv, tt, ns, pass, lane);
            int m = 0, n = 0;
            const uint32_t *aw = A.words, *bw = B.words;
            uint32_t *dbase = sv.dirs;
            if (ls.on) {
                m = A.len[ls.pA]; n = B.len[ls.pA];
                aw = A.words + A.start[ls.pA]; bw = B.words + B.start[ls.pA];
                dbase = sv.dirs + sv.dir_off[ls.pA];
            }
            const int s = ls.s;
            const int col0 = s * C;
            int bc[C];
#pragma unroll
            for (int k = 0; k < C / PER; ++k) {
                uint32_t x = ls.on ? __ldg(bw + col0 / PER + k) : 0u;
#pragma unroll
                for (int c = 0; c < PER; ++c) {
                    const uint32_t code = (x >> (BITS * c)) & ((1u << BITS) - 1u);
                    bc[k * PER + c] = (int)(sbase + (code * REP + (lane % REP)) * (uint32_t)sizeof(T));   // shared byte address; row offset added per row
                }
            }
            T H[C];
            int HS[F64 ? C : 1];
            uint32_t acc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                acc[c] = 0u;
                if constexpr (F64) { H[c] = __dmul_rn((double)(col0 + c + 1), c_ins); HS[c] = col0 + c + 1; }
                else H[c] = 0;
            }
            T last = 0, prev_recv = 0;
            int last_s = 0, prev_recv_s = 0;
            uint32_t cur = 0u;
            const int steps = warp_max(ls.on ? m + ls.sk : 0);
            const bool has_in = __shfl_sync(RSD_FULL, (int)ls.from_scratch, 0) != 0;
            const bool has_out = __shfl_sync(RSD_FULL, (int)ls.to_scratch, 31) != 0;
            const int off_in = (pass & 1) ? scratch_stride : 0, off_out = (pass & 1) ? 0 : scratch_stride;
            uint32_t *dcol = dbase + col0;

            auto run_rows = [&](auto handoff_tag) {          // two copies: passes without a hand-off run the lean one
            constexpr bool HANDOFF = decltype(handoff_tag)::value;
#pragma unroll 1
            for (int t = 0; t < steps; ++t) {
                T recv = __shfl_up_sync(RSD_FULL, last, 1);
                int recv_s = 0;
                if constexpr (F64) recv_s = __shfl_up_sync(RSD_FULL, last_s, 1);
                const int i = t - ls.sk;
                const bool row_on = ls.on && (unsigned)i < (unsigned)m;
                if (ls.lead) {
                    if constexpr (F64) { recv = __dmul_rn((double)(i + 1), c_del); recv_s = i + 1; }
                    else recv = 0;
                }
                if constexpr (HANDOFF) {
                    if (ls.from_scratch && row_on) {
                        recv = scr[off_in + i];
                        if constexpr (F64) recv_s = scr_steps[off_in + i];
                    }
                }
                if (row_on) {
                    if (i % PER == 0) cur = __ldg(aw + i / PER);
                    const uint32_t rowoff = ((cur & ((1u << BITS) - 1u)) << 4) * (uint32_t)(REP * sizeof(T)); cur >>= BITS;
                    T left = recv, diag = prev_recv;
                    [[maybe_unused]] int left_s = recv_s, diag_s = prev_recv_s;
                    if constexpr (F64) {
                        if (i == 0) { diag = (s == 0) ? 0.0 : __dmul_rn((double)col0, c_ins); diag_s = col0; }
                        else if (s == 0) { diag = __dmul_rn((double)i, c_del); diag_s = i; }
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        T w;
                        if constexpr (F64) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(w) : "r"((uint32_t)bc[c] + rowoff));
                        else asm volatile("ld.shared.s32 %0, [%1];" : "=r"(w) : "r"((uint32_t)bc[c] + rowoff));
                        if constexpr (F64) {
                            uint32_t code;
                            const double c0 = __dadd_rn(left, c_ins);
                            const double c1 = __dadd_rn(H[c], c_del);
                            const double c2 = __dadd_rn(diag, w);
                            const int s0 = left_s + 1, s1 = HS[c] + 1, s2 = diag_s + 1;
                            const double v = dmin2(c0, dmin2(c1, c2));         // same value; only one compare-select after left + ins
                            int bs = (c0 == v) ? s0 : 0x7fffffff; code = 0u;
                            if (c1 == v && s1 < bs) { bs = s1; code = 1u; }
                            if (c2 == v && s2 < bs) { bs = s2; code = 2u; }
                            diag = H[c]; diag_s = HS[c];
                            H[c] = v; HS[c] = bs;
                            left = v; left_s = bs;
                            acc[c] = (acc[c] << 2) | (code == 0u ? 0u : code + 1u);      // 0 -> 00, 1 -> 10, 2 -> 11
                        } else {
                            const int up = H[c];
                            const int t2 = addmin32(diag, w, up);              // min(diag + w, up): ties keep DEL
                            const int d_upd = sub_fma(t2, up, minus_one);      // < 0  <=>  diagonal strictly better than up
                            const int d_ins = sub_fma(t2, left, minus_one);    // < 0  <=>  left loses (ties keep INS)
                            diag = up;
                            H[c] = min(t2, left);
                            left = H[c];
                            acc[c] = __funnelshift_l((uint32_t)d_ins, acc[c], 1);   // sign bits shifted in: "not INS",
                            acc[c] = __funnelshift_l((uint32_t)d_upd, acc[c], 1);   // then "UPD rather than DEL"
                        }
                    }
                    last = left; prev_recv = recv;
                    if constexpr (F64) { last_s = left_s; prev_recv_s = recv_s; }
                    if constexpr (HANDOFF) {
                        if (ls.to_scratch) { scr[off_out + i] = last; if constexpr (F64) scr_steps[off_out + i] = last_s; }
                    }
                    if ((i & 15) == 15 || i == m - 1) {
                        const int sh = 2 * (15 - (i & 15));                  // partial last block: align as if 16 rows
                        uint4 *dst = reinterpret_cast<uint4 *>(dcol + (size_t)(i >> 4) * n_pad);
#pragma unroll
                        for (int c = 0; c < C; c += 4)
                            dst[c >> 2] = make_uint4(acc[c] << sh, acc[c + 1] << sh, acc[c + 2] << sh, acc[c + 3] << sh);
                    }
                }
            }
            };
            if (has_in || has_out) run_rows(std::true_type{}); else run_rows(std::false_type{});
            if (ls.on && s == (n - 1) / C) {
                const int cl = (n - 1) - s * C;
                T res = 0;
#pragma unroll
                for (int c = 0; c < C; ++c) if (c == cl) res = H[c];
                if constexpr (F64) sv.dist[ls.pA] = res;
                else {
                    // undo the H' transform on the key, then split cost / steps
                    const long long key = (long long)res + (long long)m * (((long long)i_del << S) + 1)
                                          + (long long)n * (((long long)i_ins << S) + 1);
                    sv.dist[ls.pA] = (double)(key >> S) * inv_scale;
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Traceback: one thread per pair walks the direction words from (m,n) to the origin and writes the
// ops sink->origin into tmp[pair slot] from the END of the slot; n_ops[p] = count.
// ------------------------------------------------------------------------------------------------
__global__ void k_traceback(const int32_t *__restrict__ a_len, const int32_t *__restrict__ b_len, int64_t n_pairs,
                            const uint32_t *__restrict__ dirs, const int64_t *__restrict__ dir_off, int C,
                            uint8_t *__restrict__ tmp, int64_t max_ops, int32_t *__restrict__ n_ops) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int m = a_len[p], n = b_len[p];
    const int n_pad = ((n + C - 1) / C) * C;
    const uint32_t *d = dirs + dir_off[p];
    uint8_t *slot = tmp + p * max_ops;
    int64_t pos = (int64_t)m + n;          // slot[pos-1] is the last op (enters (m,n))
    int i = m, j = n;
    while (i > 0 && j > 0) {
        const uint32_t w = d[(size_t)((i - 1) >> 4) * n_pad + (j - 1)];
        const uint32_t code = dir_decode(w, i - 1);
        slot[--pos] = (uint8_t)code;
        if (code == 0u) --j; else if (code == 1u) --i; else { --i; --j; }
    }
    while (j > 0) { slot[--pos] = 0; --j; }        // row 0: inserts (SED:146-164)
    while (i > 0) { slot[--pos] = 1; --i; }        // column 0: deletes (SED:167-182)
    n_ops[p] = (int32_t)((int64_t)m + n - pos);
}

// ------------------------------------------------------------------------------------------------
// Finalize: one CTA per pair.  Reads the reversed-in-place ops, block-scans the two 0/1 streams
//   ai[k] = [op != INS] (consumes a source symbol), bj[k] = [op != DEL] (produces a destination symbol)
// which gives, per op, the matrix cell it enters (oi = incl. scan of ai, oj = incl. scan of bj) — the
// reference's source.index / destination.index are oi-1 / oj-1 (SED:302-323) — and the output
// position of every produced symbol: the closed-form patch  out = dest chars of non-delete ops ++
// x[len(src):]  (patching SED:380-457 on generated scripts).
// Also evaluates the reference's error code (SED:389-399) against x and, for the round-trip mode,
// compares the patched string with B.
// ------------------------------------------------------------------------------------------------
struct FinalizeArgs {
    const uint8_t *tmp; int64_t max_ops; const int32_t *n_ops;          // ops per pair slot
    int end_aligned;                                                      // 1: ops end at slot offset m+n (traceback), 0: start at 0
    SeqView A, B, X; int bits;                                            // X == A for the round-trip check
    uint8_t *op; int32_t *oi; int32_t *oj;                                // optional packed script outputs (may be NULL)
    int64_t out_stride;                                                   // stride of op/oi/oj slots
    uint8_t *patched; int64_t max_out; int32_t *out_len; int32_t *err;    // optional patch outputs (may be NULL)
    uint8_t *ok;                                                          // optional: 1 iff err == 0 and patched == B
};

__global__ void __launch_bounds__(256) k_finalize(FinalizeArgs fa, int64_t n_pairs) {
    const int64_t p = blockIdx.x;
    if (p >= n_pairs) return;
    __shared__ int s_ai[256], s_bj[256];
    __shared__ int s_carry_i, s_carry_j, s_bad_src;
    const int tid = threadIdx.x;
    const int m = fa.A.len[p], n = fa.B.len[p];
    const int k_ops = fa.n_ops[p];
    const uint8_t *src = fa.tmp + p * fa.max_ops + (fa.end_aligned ? ((int64_t)m + n - k_ops) : 0);
    const int64_t a0 = fa.A.start[p], b0 = fa.B.start[p];
    const bool do_patch = fa.patched != nullptr || fa.ok != nullptr || fa.err != nullptr;
    const int xlen = do_patch ? fa.X.len[p] : 0;
    const int64_t x0 = do_patch ? fa.X.start[p] : 0;
    if (tid == 0) { s_carry_i = 0; s_carry_j = 0; s_bad_src = 0; }
    __syncthreads();
    for (int base = 0; base < k_ops; base += 256) {
        const int k = base + tid;
        const int o = k < k_ops ? src[k] : 3;
        const int ai = (o == 1 || o == 2), bj = (o == 0 || o == 2);
        s_ai[tid] = ai; s_bj[tid] = bj;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            int xi = tid >= off ? s_ai[tid - off] : 0, xj = tid >= off ? s_bj[tid - off] : 0;
            __syncthreads();
            s_ai[tid] += xi; s_bj[tid] += xj;
            __syncthreads();
        }
        const int ci = s_carry_i, cj = s_carry_j;
        const int oi = ci + s_ai[tid], oj = cj + s_bj[tid];       // inclusive: the cell entered
        if (k < k_ops) {
            if (fa.op) fa.op[p * fa.out_stride + k] = (uint8_t)o;
            if (fa.oi) fa.oi[p * fa.out_stride + k] = oi;
            if (fa.oj) fa.oj[p * fa.out_stride + k] = oj;
            if (do_patch) {
                if (ai) {       // source symbol consumed: compare with x (error code, SED:391)
                    // reference wraps index -1 to the last char (SED:302); oi >= 1 here so no wrap
                    const uint32_t sc = pk_get(fa.A.words, a0, oi - 1, fa.bits);
                    if (oi - 1 >= xlen || pk_get(fa.X.words, x0, oi - 1, fa.bits) != sc) s_bad_src = 1;
                }
                if (bj) {       // destination symbol produced at output position oj-1
                    const uint32_t dc = pk_get(fa.B.words, b0, oj - 1, fa.bits);
                    if (fa.patched) fa.patched[p * fa.max_out + (oj - 1)] = (uint8_t)dc;
                }
            }
        }
        __syncthreads();
        if (tid == 255) { s_carry_i = oi; s_carry_j = oj; }
        __syncthreads();
    }
    if (!do_patch) return;
    const int srclen = s_carry_i;      // == m for a complete script
    const int produced = s_carry_j;    // == n
    // tail of x beyond the consumed source (patching keeps it: SED:393 error code 1 path)
    for (int k = srclen + tid; k < xlen; k += 256)
        if (fa.patched) fa.patched[p * fa.max_out + produced + (k - srclen)] = (uint8_t)pk_get(fa.X.words, x0, k, fa.bits);
    __syncthreads();
    if (tid == 0) {
        int code;
        if (!s_bad_src && xlen == srclen) code = 0;
        else if (xlen >= srclen) code = 1;
        else code = -1;
        if (fa.err) fa.err[p] = code;
        if (fa.out_len) fa.out_len[p] = code < 0 ? 0 : produced + (xlen - srclen);
        // round trip: the produced symbols are B's symbols by construction iff produced == n and no tail
        if (fa.ok) fa.ok[p] = (uint8_t)(code == 0 && produced == n && srclen == m && xlen == m);
    }
}
