// qd_plan.cuh -- one launch over every parameter tensor of a model (SURVEY.md
// section 8 f1).  The training loop of the reference quantizes the model one
// tensor at a time (cnn_models/conv_forward_model.py:236-247: 22-60 tensors,
// ~12 stock launches each); most of those tensors are 10-500 elements and are
// pure launch latency.  A plan flattens all rows of all tensors into one row
// space; a warp maps its global row to (tensor, local row) with a binary search
// over the per-tensor row prefix held in shared memory and then runs exactly
// the single-tensor warp path on it.
#pragma once
#include "qd_warp_path.cuh"

namespace qd {

struct PlanEntry {
    const float* src;
    float* dst;
    float* save;        // optional: full-precision copy of src written by the forward launch (shadow buffer)
    int64_t n;
    int64_t row_start;  // first global row of this tensor
    int64_t rows;
    int64_t row_len;
    float S;            // levels - 1
    float rS;           // RN(1/S)
    float lim;          // 0.5 - S*2^-20
    int vec;            // 16-byte aligned rows
};

constexpr int kPlanSmemEntries = 256;

// Gradient pointers of the backward launch travel BY VALUE in the kernel parameters (up to 256
// tensors = 2 KB): no host->device table copy per step, and the launch can be captured in a
// CUDA graph (a captured memcpy from a temporary host array could not).
constexpr int kPlanGradsByValue = 256;
struct GradTable {
    float* g[kPlanGradsByValue];
};

// forward of one plan row; with a shadow pointer the row is also written, untouched, to the
// master copy (save + quantize in one pass: read 4, write 4 + 4 bytes per element)
template <int R, bool VEC, bool FULL>
__device__ __forceinline__ void plan_forward_row(const Params& P, float* save, int64_t row, int lane) {
    float v[4 * R], gv[4 * R];
    Centroids cen{nullptr, nullptr, 0};
    LaneTable<OP_UNIFORM, BWD_OFF> rt;
    warp_load_row<OP_UNIFORM, BWD_OFF, R, VEC, FULL>(P, row, lane, v, gv);
    if (save != nullptr) {
        const int64_t base = row * P.geo.row_len;
        const int len = FULL ? R * 128 : (int)min(P.geo.row_len, P.geo.n - base);
        store_row<R, VEC, FULL>(save + base, len, lane, v);
    }
    warp_compute_row<OP_UNIFORM, BWD_OFF, R, VEC, FULL>(P, cen, rt, row, lane, v, gv);
}

template <int BWD, int R>
__global__ void __launch_bounds__(kWarpCtaThreads) plan_rows_kernel(const PlanEntry* __restrict__ entries, int count,
                                                                   int64_t total_rows, float* const* __restrict__ grads,
                                                                   int with_save, const __grid_constant__ GradTable gtab) {
    __shared__ int64_t s_start[kPlanSmemEntries];
    const bool in_smem = count <= kPlanSmemEntries;
    if (in_smem) {
        for (int i = threadIdx.x; i < count; i += blockDim.x) s_start[i] = entries[i].row_start;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * kWarpsPerCta;
    Centroids cen{nullptr, nullptr, 0};
    LaneTable<OP_UNIFORM, BWD_OFF> rt;
    for (int64_t grow = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); grow < total_rows; grow += stride) {
        int lo = 0, hi = count - 1;  // largest t with row_start[t] <= grow
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            int64_t s = in_smem ? s_start[mid] : entries[mid].row_start;
            if (s <= grow) lo = mid; else hi = mid - 1;
        }
        const PlanEntry en = entries[lo];
        Params P;
        P.x = en.src;
        P.q = (BWD == BWD_OFF) ? en.dst : nullptr;
        float* gptr = nullptr;
        if constexpr (BWD != BWD_OFF) gptr = (grads != nullptr) ? grads[lo] : gtab.g[lo];
        P.g = gptr;
        P.gout = gptr;
        P.xhat = nullptr; P.idx8 = nullptr; P.idx64 = nullptr;
        P.alpha = nullptr; P.beta = nullptr; P.argmin = nullptr; P.argmax = nullptr;
        P.mean = nullptr; P.max_element = 0.f; P.points = nullptr; P.num_points = 0; P.rule = 0;
        P.geo.n = en.n; P.geo.row_len = en.row_len; P.geo.rows = en.rows;
        P.S = en.S; P.rS = en.rS; P.half_minus_band = en.lim; P.stochastic = 0; P.seed = 0; P.offset = 0;
        const int64_t row = grow - en.row_start;
        bool vec = en.vec != 0;
        if constexpr (BWD != BWD_OFF) vec = vec && ((reinterpret_cast<uintptr_t>(P.g) & 15) == 0);
        const bool full = (en.row_len == R * 128) && ((row + 1) * en.row_len <= en.n);
        if constexpr (BWD == BWD_OFF) {
            float* save = with_save ? en.save : nullptr;
            if (vec && save != nullptr) vec = (reinterpret_cast<uintptr_t>(save) & 15) == 0;
            if (vec) {
                if (full) plan_forward_row<R, true, true>(P, save, row, lane);
                else plan_forward_row<R, true, false>(P, save, row, lane);
            } else {
                plan_forward_row<R, false, false>(P, save, row, lane);
            }
        } else if (vec) {
            if (full) warp_process_row<OP_UNIFORM, BWD, R, true, true>(P, cen, rt, row, lane);
            else warp_process_row<OP_UNIFORM, BWD, R, true, false>(P, cen, rt, row, lane);
        } else {
            warp_process_row<OP_UNIFORM, BWD, R, false, false>(P, cen, rt, row, lane);
        }
    }
}

}  // namespace qd
