// k_ingest.cuh — device side of ingest: word offsets of a canonically packed batch from its lengths, and packing
// of raw symbol codes (1 byte per symbol) into the 2 / 4 bit words the kernels read.
//
// Canonical layout = what rsd_pack writes: sequence p starts at word sum_{q<p} nwords(len[q]).  The host call only
// sums the lengths of blocks of RSD_SCAN_BLOCK sequences (a vectorised pass it needs anyway for its chunk
// boundaries); the scan inside a block runs here, so start[] (8 bytes per sequence) never crosses PCIe.
#pragma once
#include "rsd_common.cuh"

#define RSD_SCAN_BLOCK 4096               // sequences per block of the start[] scan (4 per thread, 1024 threads)

// start[p] = base[p / 4096] + (exclusive scan of nwords(len[.]) inside the block); grid = number of blocks.
// With sym_start != nullptr the symbol offsets (exclusive scan of len) are produced the same way from sym_base.
__global__ void __launch_bounds__(1024)
k_starts_from_len(const int32_t *__restrict__ len, int64_t n, int sh, const int64_t *__restrict__ base,
                  int64_t *__restrict__ start, const int64_t *__restrict__ sym_base, int64_t *__restrict__ sym_start) {
    __shared__ unsigned long long s_warp[32];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * RSD_SCAN_BLOCK + 4 * t;
    const int add = (1 << sh) - 1;
    int L[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) L[q] = p0 + q < n ? len[p0 + q] : 0;
    // words in the low half, symbols in the high half of one 64-bit scan value (a block holds < 2^32 of either)
    unsigned long long v = 0ull;
#pragma unroll
    for (int q = 0; q < 4; ++q) v += (unsigned long long)((L[q] + add) >> sh) | ((unsigned long long)(unsigned)L[q] << 32);
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long x = __shfl_up_sync(RSD_FULL, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long x = __shfl_up_sync(RSD_FULL, wi, o); if (lane >= o) wi += x; }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    unsigned long long ex = s_warp[wid] + inc - v;
    int64_t wo = base[blockIdx.x] + (int64_t)(ex & 0xffffffffull);
    int64_t so = sym_start ? sym_base[blockIdx.x] + (int64_t)(ex >> 32) : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (p0 + q < n) { start[p0 + q] = wo; if (sym_start) sym_start[p0 + q] = so; }
        wo += (L[q] + add) >> sh; so += L[q];
    }
}

// Pack raw codes into words: one thread per output word.  The owner of a word is the last sequence whose start is
// <= the word index (sequences of length 0 share their start with their successor, which owns the word): lane 0
// finds the owner of the warp's first word by binary search over start[] (L2-resident), the other lanes search the
// short window behind it.  The pass is bandwidth-bound on the code bytes (1 byte per symbol in, BITS bits out).
// bad[0] receives 1 + the index of a sequence holding a code that does not fit BITS; mask[0] the symbols seen.
__device__ __forceinline__ int64_t last_start_le(const int64_t *__restrict__ start, int64_t lo, int64_t hi, int64_t w) {
    while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if (start[mid] <= w) lo = mid; else hi = mid - 1; }
    return lo;
}

template <int BITS>
__global__ void __launch_bounds__(256)
k_pack_codes(const uint8_t *__restrict__ codes, const int64_t *__restrict__ sym_start, const int64_t *__restrict__ start,
             const int32_t *__restrict__ len, int64_t n, int64_t w_begin, int64_t w_end, int64_t sym_origin,
             uint32_t *__restrict__ words, unsigned long long *__restrict__ bad, uint32_t *__restrict__ mask) {
    constexpr int PER = 32 / BITS;
    const int lane = threadIdx.x & 31;
    const int64_t w = w_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t seen = 0u;
    int64_t p_first = 0;
    if (lane == 0 && w < w_end) p_first = last_start_le(start, 0, n - 1, w);
    p_first = __shfl_sync(RSD_FULL, p_first, 0);
    if (w < w_end) {
        int64_t hi = min(p_first + 64, n - 1);
        if (start[hi] <= w) hi = n - 1;                         // a run of empty sequences: search the rest
        const int64_t p = last_start_le(start, p_first, hi, w);
        const int64_t k0 = (w - start[p]) * PER;
        const int L = len[p];
        const uint8_t *src = codes + (sym_start[p] - sym_origin) + k0;
        const int cnt = (int)min((int64_t)PER, (int64_t)L - k0);
        uint32_t word = 0u;
        if (cnt == PER && ((uintptr_t)src & 7) == 0) {
            const uint2 *s8 = reinterpret_cast<const uint2 *>(src);
#pragma unroll
            for (int h = 0; h < PER / 8; ++h) {
                const uint2 x = s8[h];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t a = (x.x >> (8 * c)) & 255u, b2 = (x.y >> (8 * c)) & 255u;
                    if (a >> BITS) atomicMax(bad, (unsigned long long)p + 1ull); else seen |= 1u << a;
                    if (b2 >> BITS) atomicMax(bad, (unsigned long long)p + 1ull); else seen |= 1u << b2;
                    word |= (a & ((1u << BITS) - 1u)) << (BITS * (8 * h + c));
                    word |= (b2 & ((1u << BITS) - 1u)) << (BITS * (8 * h + 4 + c));
                }
            }
        } else {
            for (int c = 0; c < cnt; ++c) {
                const uint32_t a = src[c];
                if (a >> BITS) atomicMax(bad, (unsigned long long)p + 1ull); else seen |= 1u << a;
                word |= (a & ((1u << BITS) - 1u)) << (BITS * c);
            }
        }
        words[w] = word;
    }
    if (mask) {                                                  // optional: which symbols occur
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) seen |= __shfl_xor_sync(RSD_FULL, seen, o);
        if (lane == 0 && seen) atomicOr(mask, seen);
    }
}
