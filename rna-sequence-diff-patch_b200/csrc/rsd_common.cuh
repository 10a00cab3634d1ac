// rsd_common.cuh — shared device helpers for the sm_100a Wagner–Fischer kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RSD_WARP 32
#define RSD_FULL 0xffffffffu

// ------------------------------------------------------------------------------------------
// Cost tables as the kernels see them (built on the host by the classifier, rsd_api.cu).
// H' transform used by the integer kernels:  H'[i][j] = D[i][j] - i*del - j*ins, so
//   H'[i][j] = min(H'[i][j-1], H'[i-1][j], H'[i-1][j-1] + w(a_i,b_j)),  w = sub - ins - del,
// both borders are 0, and D[m][n] = H'[m][n] + m*del + n*ins.  Integer adds are exact, so this is
// the same number the reference's left+ins / up+del / diag+sub (SED:95-99) produces.
// ------------------------------------------------------------------------------------------
struct IntCosts {
    int32_t ins, del;        // scaled by 2^k
    int32_t w[16][16];       // w[a][b] = sub(a,b) - ins - del (scaled); w[a][a] = -(ins+del)
    int32_t scale_log2;      // k
    uint32_t rowtab4[8];     // 2-bit fast path: byte b of rowtab4[a] = v[a][b] = max(0, -w[a][b]) (<= 127);
                             // rowtab4[4 + a] holds the transposed table v[b][a] for pairs processed with swapped roles
};
struct F64Costs {
    double ins, del;
    double sub[16][16];      // sub[a][a] = 0.0 (SED:79-81)
};

// ------------------------------------------------------------------------------------------
// packed code access
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pk_get(const uint32_t *__restrict__ w, int64_t start, int k, int bits) {
    if (bits == 4) return (w[start + (k >> 3)] >> ((k & 7) * 4)) & 15u;
    return (w[start + (k >> 4)] >> ((k & 15) * 2)) & 3u;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// min(a + b, c) per signed halfword — one VIADDMNMX.S16x2 on sm_90+/sm_100a
__device__ __forceinline__ uint32_t addmin16x2(uint32_t a, uint32_t b, uint32_t c) {
    return __viaddmin_s16x2(a, b, c);
}
__device__ __forceinline__ uint32_t min16x2(uint32_t a, uint32_t b) {
    return __vimin3_s16x2(a, b, b);
}
// packed add on the fma pipe: a * one + b with `one` an opaque kernel argument equal to 1, so ptxas
// keeps an IMAD instead of folding it into an alu-pipe IADD3.  Halves must not carry into each other.
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}
// a - b as b * (-1) + a on the fma pipe (`minus_one` is an opaque kernel argument equal to -1)
__device__ __forceinline__ int sub_fma(int a, int b, int minus_one) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(minus_one), "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t max3u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ int addmin32(int a, int b, int c) { return __viaddmin_s32(a, b, c); }

// min of two non-NaN doubles as compare + select (fmin() also pays for NaN quieting: ~7 instructions)
__device__ __forceinline__ double dmin2(double a, double b) { return a < b ? a : b; }

__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(RSD_FULL, v, o));
    return v;
}
