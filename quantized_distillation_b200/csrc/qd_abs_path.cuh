// qd_abs_path.cuh -- row a10, an EXTENSION with NO parity target ("parity: unpinned").
//
// The reference offers 'absmax' / 'absnorm' scaling (quant_functions.py:109-127, 144-146; selectable through
// quantizationFunctionToUse='uniformAbsMaxScaling', conv_forward_model.py:206-208) but the code cannot execute:
// `tensor.max(p=2)` is not a torch call and `self.norm_scaling = norm_scaling.view` stores a bound method.  What the
// lines evidently intend, once those two slips are repaired, is implemented here:
//     sign = sign(x); v = |x|;  norm_b = max_j v_j  (absmax)  or  sqrt(sum_j v_j^2)  (absnorm, padded tail included),
//     norm_b < 1e-10 -> 1;  x_hat = v / norm_b;  inverse: y * norm_b * sign (+ mean)
// and the uniform op on top of it: level = rint(x_hat * S), q = ((level / S) * norm_b) * sign.  One rounding per
// reference op as everywhere else; the L2 norm is accumulated in float64 and rounded once (torch's float32
// reduction order is unspecified, one more reason this row stays unpinned).
// One kernel for every row length: a GROUP of threads (a warp for rows <= 1024, a 256-thread CTA beyond) reads its
// row twice (the second time from L1/L2).
#pragma once
#include "qd_block_path.cuh"

namespace qd {

enum AbsMode : int { ABS_SCALE = 0, ABS_UNIFORM = 1 };

struct AbsParams {
    const float* x;
    float* out;        // x_hat (padded layout) for ABS_SCALE, q for ABS_UNIFORM
    float* sign;       // ABS_SCALE: sign tensor, padded layout (the reference keeps torch.sign(tensor))
    uint8_t* idx8;     // ABS_UNIFORM: optional integer levels
    float* norm;       // per-row scale
    const float* mean;
    float max_element;
    Geometry geo;
    float S;
    int kind;          // QD_SCALE_ABSMAX / QD_SCALE_ABSNORM
};

__device__ __forceinline__ float sign_of(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : v); }  // torch.sign: 0 -> 0, NaN -> NaN

template <int MODE, int GROUP>
__global__ void __launch_bounds__(256) abs_rows_kernel(const __grid_constant__ AbsParams P) {
    __shared__ double s_scratch[8];
    const int tid = (GROUP == 32) ? (threadIdx.x & 31) : threadIdx.x;
    constexpr int kGroups = 256 / GROUP;
    const bool pre = (P.mean != nullptr) || (P.max_element > 0.f);
    const float mean = P.mean ? *P.mean : 0.f;
    for (int64_t row = (int64_t)blockIdx.x * kGroups + (GROUP == 32 ? (threadIdx.x >> 5) : 0); row < P.geo.rows;
         row += (int64_t)gridDim.x * kGroups) {
        const int64_t base = row * P.geo.row_len;
        const int64_t len = min(P.geo.row_len, P.geo.n - base);
        const float* src = P.x + base;
        // ---- pass 1: norm of the row -----------------------------------------------------------
        float mx = 0.f;
        double ss = 0.0;
        bool nan = false;
        for (int64_t e = tid; e < len; e += GROUP) {
            float v = src[e];
            if (pre) v = pre_op(v, mean, P.max_element);
            const float a = fabsf(v);
            nan = nan || (a != a);
            mx = fmaxf(mx, a);
            ss += (double)a * (double)a;
        }
        if (P.kind == QD_SCALE_ABSNORM && len < P.geo.row_len && P.geo.rows > 1 && tid == 0) {
            // the padded tail repeats the last element of the tensor (help_functions.py:80-86) and counts in the norm
            float v = P.x[P.geo.n - 1];
            if (pre) v = pre_op(v, mean, P.max_element);
            ss += (double)(P.geo.row_len - len) * (double)v * (double)v;
        }
        mx = grp_minmax<GROUP, false>(nan ? __int_as_float(0x7fc00000) : mx, reinterpret_cast<float*>(s_scratch));
        if (P.kind == QD_SCALE_ABSNORM) ss = grp_sum<GROUP>(ss, s_scratch);
        float norm = (P.kind == QD_SCALE_ABSMAX) ? mx : (float)sqrt(ss);
        if (norm < kTolDiffZero) norm = 1.0f;                                   // :118-123
        if (P.norm != nullptr && tid == 0) P.norm[row] = norm;
        // ---- pass 2 ----------------------------------------------------------------------------
        const int64_t plen = (MODE == ABS_SCALE) ? P.geo.row_len : len;          // x_hat / sign are written in the padded layout
        float lastv = 0.f;
        if (MODE == ABS_SCALE && len < plen) {
            lastv = P.x[P.geo.n - 1];
            if (pre) lastv = pre_op(lastv, mean, P.max_element);
        }
        for (int64_t e = tid; e < plen; e += GROUP) {
            float v = (e < len) ? src[e] : lastv;
            if (pre && e < len) v = pre_op(v, mean, P.max_element);
            const float sg = sign_of(v);
            const float xh = __fdiv_rn(fabsf(v), norm);                         // abs_, div_ (:110, :127)
            if constexpr (MODE == ABS_SCALE) {
                P.out[base + e] = xh;
                if (P.sign != nullptr) P.sign[base + e] = sg;
            } else {
                const float lvl = rintf(__fmul_rn(xh, P.S));                    // mul_, round_ (:189-190)
                float q = __fmul_rn(__fmul_rn(__fdiv_rn(lvl, P.S), norm), sg);  // div_ (:191), mul_(norm), mul_(sign) (:145-146)
                if (pre) q = __fadd_rn(q, mean);                                // add_(mean) (:148)
                P.out[base + e] = q;
                if (P.idx8 != nullptr) P.idx8[base + e] = (uint8_t)(int)lvl;
            }
        }
        if constexpr (GROUP != 32) __syncthreads();
    }
}

// y * norm * sign (+ mean), padding dropped (quant_functions.py:144-150)
__global__ void __launch_bounds__(256) abs_inv_scale_kernel(const float* __restrict__ y, const float* __restrict__ sign,
                                                            const float* __restrict__ norm, const float* __restrict__ mean,
                                                            float* __restrict__ out, Geometry geo) {
    const float m = mean ? *mean : 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < geo.n; i += stride) {
        const int64_t row = (geo.rows == 1) ? 0 : i / geo.row_len;
        float v = __fmul_rn(__fmul_rn(y[i], norm[row]), sign[i]);
        if (mean) v = __fadd_rn(v, m);
        out[i] = v;
    }
}

}  // namespace qd
