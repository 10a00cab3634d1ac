// k_ubench.cuh — issue-rate microbenchmarks: the denominators of the cell-update roofline
// (MEASURED_PEAKS.json has no INT32 entry; SURVEY 8d asks for a measured one).
// Every thread runs 8 independent dependency chains of one instruction kind.
#pragma once
#include "rsd_common.cuh"

#define RSD_UB_CHAINS 8
#define RSD_UB_REPS 16

template <int WHICH>
__device__ __forceinline__ uint32_t ub_op(uint32_t x, uint32_t y, uint32_t z) {
    if constexpr (WHICH == 0) {            // IADD3: three-input add cannot be merged further
        uint32_t r; asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(r) : "r"(x), "r"(y), "r"(z));
        return r;
    } else if constexpr (WHICH == 1) {     // VIADDMNMX (s32)
        return (uint32_t)__viaddmin_s32((int)x, (int)y, (int)z);
    } else if constexpr (WHICH == 2) {     // VIADDMNMX.S16x2
        return __viaddmin_s16x2(x, y, z);
    } else if constexpr (WHICH == 3) {     // PRMT
        return prmt(x, y, z);
    } else if constexpr (WHICH == 5) {     // IMAD
        return x * y + z;
    } else if constexpr (WHICH == 6) {     // VIMNMX3
        return (uint32_t)__vimin3_s32((int)x, (int)y, (int)z);
    } else if constexpr (WHICH == 8) {     // VIMNMX.S16x2 two-input form used by the kernels
        return __vimin3_s16x2(x, y, y);
    } else if constexpr (WHICH == 9) {     // IMAD.HI.U32
        uint32_t r; asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z));
        return r;
    } else if constexpr (WHICH == 10) {    // IMAD.WIDE.U32 (64-bit product, low half fed back)
        unsigned long long r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x), "r"(y));
        return (uint32_t)r ^ (uint32_t)(r >> 32);
    } else if constexpr (WHICH == 11) {    // SHF (funnel shift, as used for the direction bits)
        return __funnelshift_l(y, x, 1);
    } else {
        return x;
    }
}

template <int WHICH>
__global__ void __launch_bounds__(256) k_ubench_u32(int iters, uint32_t y, uint32_t z, uint32_t *sink) {
    uint32_t x[RSD_UB_CHAINS];
#pragma unroll
    for (int u = 0; u < RSD_UB_CHAINS; ++u) x[u] = threadIdx.x * 2654435761u + u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < RSD_UB_REPS; ++r)
#pragma unroll
            for (int u = 0; u < RSD_UB_CHAINS; ++u) x[u] = ub_op<WHICH>(x[u], y, z);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int u = 0; u < RSD_UB_CHAINS; ++u) acc ^= x[u];
    if (acc == 0x12345678u) sink[0] = acc;
}

// WHICH == 7: the instruction mix of the int16x2 cell (PRMT + VIADDMNMX.S16x2 + VIMNMX.S16x2),
// counted as 3 ops per "cell" so the number is comparable with the single-kind rates.
__global__ void __launch_bounds__(256) k_ubench_mix(int iters, uint32_t y, uint32_t z, uint32_t *sink) {
    uint32_t x[RSD_UB_CHAINS];
#pragma unroll
    for (int u = 0; u < RSD_UB_CHAINS; ++u) x[u] = threadIdx.x * 2654435761u + u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < RSD_UB_REPS; ++r)
#pragma unroll
            for (int u = 0; u < RSD_UB_CHAINS; ++u) {
                uint32_t w = prmt(y, z, x[u]);
                uint32_t t = __viaddmin_s16x2(x[u], w, y);
                x[u] = __vimin3_s16x2(t, z, z);
            }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int u = 0; u < RSD_UB_CHAINS; ++u) acc ^= x[u];
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void __launch_bounds__(256) k_ubench_f64(int iters, double y, double *sink) {
    double x[RSD_UB_CHAINS];
#pragma unroll
    for (int u = 0; u < RSD_UB_CHAINS; ++u) x[u] = (double)(threadIdx.x + u);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < RSD_UB_REPS; ++r)
#pragma unroll
            for (int u = 0; u < RSD_UB_CHAINS; ++u) x[u] = __dadd_rn(x[u], y);
    }
    double acc = 0;
#pragma unroll
    for (int u = 0; u < RSD_UB_CHAINS; ++u) acc += x[u];
    if (acc == 1.2345) sink[0] = acc;
}
