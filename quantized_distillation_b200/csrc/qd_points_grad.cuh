// qd_points_grad.cuh -- gradient of the loss w.r.t. the K centroids
// (nonUniformQuantization_variable.backward, quant_functions.py:471-506):
//
//     grad_points[k] = sum_{i : idx_i = k} fl32(g_i * alpha_row(i))
//
// The reference makes K masked passes over the tensor; here it is one pass,
// 5 B/elt (float32 g + uint8 idx).  Work item = up to 1024 consecutive elements
// of one row (so alpha is a per-item scalar), one warp per item, float64
// accumulators in registers for 8 centroids at a time.  Reduction order is
// fixed (lane tree -> warps in order -> CTAs in order), so the result is
// deterministic and data-parallel replicas stay bit-identical without any
// communication.
#pragma once
#include "qd_common.cuh"

namespace qd {

constexpr int kPgThreads = 256;
constexpr int kPgWarps = kPgThreads / 32;
constexpr int kPgItem = 1024;
constexpr int kPgGroup = 8;  // centroids accumulated per sweep

template <typename IdxT>
__global__ void __launch_bounds__(kPgThreads) points_grad_partial(const float* __restrict__ g,
                                                                 const IdxT* __restrict__ idx,
                                                                 const float* __restrict__ alpha, int K, Geometry geo,
                                                                 double* __restrict__ partial /*[gridDim.x][K]*/) {
    extern __shared__ double s_acc[];  // [kPgWarps][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kPgWarps * K; i += kPgThreads) s_acc[i] = 0.0;
    __syncthreads();

    const int64_t items_per_row = (geo.row_len + kPgItem - 1) / kPgItem;
    const int64_t items = geo.rows * items_per_row;
    const int64_t stride = (int64_t)gridDim.x * kPgWarps;
    for (int kg = 0; kg < K; kg += kPgGroup) {
        double acc[kPgGroup];
#pragma unroll
        for (int k = 0; k < kPgGroup; ++k) acc[k] = 0.0;
        for (int64_t item = (int64_t)blockIdx.x * kPgWarps + warp; item < items; item += stride) {
            const int64_t row = item / items_per_row, sub = item % items_per_row;
            const int64_t start = row * geo.row_len + sub * kPgItem;
            const int64_t row_end = min((row + 1) * geo.row_len, geo.n);
            const int len = (int)min((int64_t)kPgItem, row_end - start);
            if (len <= 0) continue;
            const float a = alpha[row];
            const float* gp = g + start;
            const IdxT* ip = idx + start;
            const bool vec = sizeof(IdxT) == 1 && ((reinterpret_cast<uintptr_t>(gp) & 15) == 0) &&
                             ((reinterpret_cast<uintptr_t>(ip) & 3) == 0);
            const int vlen = vec ? (len & ~3) : 0;
            for (int e = lane * 4; e < vlen; e += 128) {
                float4 t = *reinterpret_cast<const float4*>(gp + e);
                uint32_t w = *reinterpret_cast<const uint32_t*>(ip + e);
                float pv[4] = {__fmul_rn(t.x, a), __fmul_rn(t.y, a), __fmul_rn(t.z, a), __fmul_rn(t.w, a)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int id = (int)((w >> (8 * j)) & 0xffu) - kg;
#pragma unroll
                    for (int k = 0; k < kPgGroup; ++k) acc[k] += (id == k) ? (double)pv[j] : 0.0;
                }
            }
            for (int e = vlen + lane; e < len; e += 32) {
                const float pv = __fmul_rn(gp[e], a);  // in-place multiply of the reference (:495)
                const int id = (int)ip[e] - kg;
#pragma unroll
                for (int k = 0; k < kPgGroup; ++k) acc[k] += (id == k) ? (double)pv : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < kPgGroup; ++k) {
            double s = warp_sum(acc[k]);
            if (lane == 0 && kg + k < K) s_acc[warp * K + kg + k] = s;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += kPgThreads) {
        double s = 0.0;
        for (int w = 0; w < kPgWarps; ++w) s += s_acc[w * K + k];
        partial[(int64_t)blockIdx.x * K + k] = s;
    }
}

__global__ void points_grad_final(const double* __restrict__ partial, int nblocks, int K, float* __restrict__ out) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * K + k];
        out[k] = (float)s;
    }
}

}  // namespace qd
