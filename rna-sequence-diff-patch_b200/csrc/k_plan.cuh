// k_plan.cuh — device-side work planning for the systolic warp kernels.
//
// A pair (m source symbols, n destination symbols) needs ns = ceil(n / C) lanes ("strips" of C
// matrix columns held in registers).  Pairs are binned by (ns, m) with a counting sort so that
// every warp is filled with groups of the same shape:
//   * ns <= 32 : floor(32/ns) groups per warp, one pass;
//   * ns  > 32 : one group per warp, ceil(ns/32) passes (class 33);
//   * "twin" bins put two pairs of identical (ns, m) into one group — the int16x2 kernels carry
//     one pair in each half of every register.
// Bins are numbered heaviest first so the persistent kernels hand out the long tasks first.
#pragma once
#include "rsd_common.cuh"

#define RSD_MQ_MAX 2048                   // m is binned exactly below MQ-1, clamped above; MQ <= this
#define RSD_NB_MAX (34 * RSD_MQ_MAX)      // classes ns = 0..33  (0 unused)

struct PlanView {
    int *pair_bin;        // [n_pairs] bin of each pair, -1 = trivial (m == 0 or n == 0)
    int *bin_cnt;         // [NB]
    int *bin_cursor;      // [NB]
    int *bin_group_off;   // [NB + 1] exclusive scan of groups per bin
    int *bin_warp_off;    // [NB + 1] exclusive scan of warps per bin
    int2 *groups;         // [n_pairs] {pair A, pair B or -1}
    int *totals;          // {n_groups, n_warps}
    int *work_counter;    // persistent-kernel ticket
    int C;                // columns per lane
    int allow_twin;
    int MQ;               // row bins per class for this call: min(max_m + 2, RSD_MQ_MAX)
    int NB;               // 34 * MQ
};

__device__ __forceinline__ int plan_bin(int m, int n, int C, int MQ) {
    int ns = (n + C - 1) / C;
    int nsq = ns > 32 ? 33 : ns;
    int mq = m < MQ - 1 ? m : MQ - 1;
    return (33 - nsq) * MQ + (MQ - 1 - mq);
}
__device__ __forceinline__ int bin_nsq(int bin, int MQ) { return 33 - bin / MQ; }
__device__ __forceinline__ int bin_mq(int bin, int MQ) { return MQ - 1 - bin % MQ; }
__device__ __forceinline__ bool bin_twin(int bin, int allow_twin, int MQ) {
    return allow_twin && bin_nsq(bin, MQ) <= 32 && bin_mq(bin, MQ) < MQ - 1;
}

// trivial pairs are answered here: D = n*ins (m == 0) or m*del (n == 0) — one fp64 multiply,
// like the reference's border rows (SED:159,177).
__global__ void k_plan_count(const int32_t *__restrict__ a_len, const int32_t *__restrict__ b_len,
                             int64_t n_pairs, PlanView pv, double ins, double del, double *out) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int m = a_len[p], n = b_len[p];
    if (m == 0 || n == 0) {
        pv.pair_bin[p] = -1;
        if (out) out[p] = m == 0 ? __dmul_rn((double)n, ins) : __dmul_rn((double)m, del);
        return;
    }
    int bin = plan_bin(m, n, pv.C, pv.MQ);
    pv.pair_bin[p] = bin;
    atomicAdd(&pv.bin_cnt[bin], 1);
}

// one block of 1024 threads walks the bins in coalesced chunks of 1024; (groups, warps) are scanned
// together as one 64-bit value with warp shuffles (three barriers per chunk)
__global__ void __launch_bounds__(1024) k_plan_scan(PlanView pv) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    unsigned long long carry = 0ull;                     // same value in every thread
    for (int base = 0; base < pv.NB; base += 1024) {
        const int b = base + t;
        unsigned long long v = 0ull;
        if (b < pv.NB) {
            const int c = pv.bin_cnt[b];
            if (c) {
                const int groups = bin_twin(b, pv.allow_twin, pv.MQ) ? (c + 1) >> 1 : c;
                const int nsq = bin_nsq(b, pv.MQ);
                const int gpw = nsq > 32 ? 1 : 32 / nsq;
                v = ((unsigned long long)groups << 32) | (unsigned)((groups + gpw - 1) / gpw);
            }
        }
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(RSD_FULL, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long x = s_warp[lane];
            unsigned long long xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(RSD_FULL, xi, o);
                if (lane >= o) xi += y;
            }
            s_warp[lane] = xi - x;                       // exclusive offset of each warp
            if (lane == 31) s_total = xi;
        }
        __syncthreads();
        const unsigned long long excl = carry + s_warp[wid] + (inc - v);
        if (b < pv.NB) {
            pv.bin_group_off[b] = (int)(excl >> 32); pv.bin_warp_off[b] = (int)(excl & 0xffffffffull);
            pv.bin_cursor[b] = 0;
            const int c = pv.bin_cnt[b];
            // a twin bin with an odd count leaves its last group without a partner
            if ((c & 1) && bin_twin(b, pv.allow_twin, pv.MQ)) pv.groups[(int)(excl >> 32) + (c >> 1)].y = -1;
        }
        carry += s_total;
        __syncthreads();
    }
    if (t == 0) {
        pv.bin_group_off[pv.NB] = (int)(carry >> 32); pv.bin_warp_off[pv.NB] = (int)(carry & 0xffffffffull);
        pv.totals[0] = (int)(carry >> 32); pv.totals[1] = (int)(carry & 0xffffffffull);
        *pv.work_counter = 0;
    }
}

__global__ void k_plan_fill(int64_t n_pairs, PlanView pv) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int bin = pv.pair_bin[p];
    if (bin < 0) return;
    int r = atomicAdd(&pv.bin_cursor[bin], 1);
    int *g = reinterpret_cast<int *>(pv.groups);
    if (bin_twin(bin, pv.allow_twin, pv.MQ)) g[2 * (pv.bin_group_off[bin] + (r >> 1)) + (r & 1)] = (int)p;
    else pv.groups[pv.bin_group_off[bin] + r] = make_int2((int)p, -1);
    pv.bin_cnt[bin] = 0;        // leave the counters zeroed for the next plan (nobody reads them after the scan)
}

// ---- what one warp task looks like, decoded by every lane ------------------------------------
struct WarpTask {
    int pA, pB;        // pair indices (pB == pA when the group has no twin); -1 when the lane idles
    bool hasB;
    int s0;            // strip index of this lane inside its group for pass 0 (== skew in rows)
    bool multi;        // class 33: one group on the whole warp, several passes
    bool on;           // lane belongs to a live group
};

__device__ __forceinline__ WarpTask plan_decode(const PlanView &pv, int W, int lane) {
    int lo = 0, hi = pv.NB;                        // largest b with bin_warp_off[b] <= W
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(&pv.bin_warp_off[mid]) <= W) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int nsq = bin_nsq(b, pv.MQ);
    WarpTask t;
    t.multi = nsq > 32;
    const int gpw = t.multi ? 1 : 32 / nsq;
    const int gbase = __ldg(&pv.bin_group_off[b]);
    const int gcount = __ldg(&pv.bin_group_off[b + 1]) - gbase;
    const int wl = W - __ldg(&pv.bin_warp_off[b]);
    const int g = t.multi ? 0 : lane / nsq;
    t.s0 = t.multi ? lane : lane - g * nsq;
    const int gi = wl * gpw + g;
    t.on = g < gpw && gi < gcount;
    t.pA = t.pB = -1; t.hasB = false;
    if (t.on) {
        int2 gp = pv.groups[gbase + gi];
        t.pA = gp.x; t.hasB = gp.y >= 0; t.pB = t.hasB ? gp.y : gp.x;
    }
    return t;
}
