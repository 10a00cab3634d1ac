// k_matrix.cuh — one pair, the whole (m+1) x (n+1) matrix of fp64 values plus the 3-bit tie mask
// (bit0 INS, bit1 DEL, bit2 UPD): everything the reference's dp object exposes to its callers
// (SED:133-224; gui.py:372-379 reads every value, create_paths SED:228-271 walks the edges).
// Always fp64 in the reference's operation order, so it is exact for every cost table.
// GUI-sized problems (tens to a few thousand symbols): one CTA walks the anti-diagonals.
#pragma once
#include "rsd_common.cuh"

__global__ void __launch_bounds__(1024)
k_matrix_f64(const uint8_t *__restrict__ a, int m, const uint8_t *__restrict__ b, int n,
             const F64Costs *__restrict__ fc, double *__restrict__ D, uint8_t *__restrict__ mask) {
    const int W = n + 1;
    const double ins = fc->ins, del = fc->del;
    for (int j = threadIdx.x; j <= n; j += blockDim.x) {
        D[j] = __dmul_rn((double)j, ins);                      // SED:159
        mask[j] = j ? 1 : 0;
    }
    for (int i = 1 + threadIdx.x; i <= m; i += blockDim.x) {
        D[(size_t)i * W] = __dmul_rn((double)i, del);          // SED:177
        mask[(size_t)i * W] = 2;
    }
    __syncthreads();
    // anti-diagonal d = i + j, cells with 1 <= i <= m, 1 <= j <= n
    for (int d = 2; d <= m + n; ++d) {
        const int ilo = max(1, d - n), ihi = min(m, d - 1);
        for (int i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
            const int j = d - i;
            const uint8_t ca = a[i - 1], cb = b[j - 1];
            const double sub = ca == cb ? 0.0 : fc->sub[ca][cb];          // SED:79-87
            const double c0 = __dadd_rn(D[(size_t)i * W + j - 1], ins);   // SED:95
            const double c1 = __dadd_rn(D[(size_t)(i - 1) * W + j], del); // SED:97
            const double c2 = __dadd_rn(D[(size_t)(i - 1) * W + j - 1], sub); // SED:99
            const double v = dmin2(dmin2(c0, c1), c2);
            D[(size_t)i * W + j] = v;
            mask[(size_t)i * W + j] = (uint8_t)((c0 == v) | ((c1 == v) << 1) | ((c2 == v) << 2)); // SED:109
        }
        __syncthreads();
    }
}
