// qd_select.cuh -- next-row f3: the two reductions of the differentiable-quantization SETUP as kernels.
//
// (1) Exact order statistics of a tensor (the 2K values np.percentile(x_hat, linspace(0,100,K)) reads,
//     help_functions.py:140-154) WITHOUT sorting it:
//        pass 1  histogram of the value bin  b(v) = clamp(floor(v * 2048), 0, 2047)  (monotone in v; x_hat
//                lives in [0, 1]) -- shared-memory privatised, one read of the tensor;
//        plan    one CTA: prefix sums, the bin and the residual rank of every requested rank, one slot
//                per distinct bin, slot offsets;
//        pass 2  the elements of the selected bins are appended (as order-preserving 32-bit keys) to
//                their slot's buffer -- second read of the tensor, output typically < 1 % of it;
//        final   one CTA per rank: most-significant-digit radix select (4 x 8 bits) of the residual
//                rank inside the slot buffer.
//     8 bytes per element instead of a full radix sort (~32+), exact for any input (the bin function
//     only has to be monotone; ties and the order inside a bin are resolved on the exact keys).
// (2) L2 norms of many tensors in two launches, float64 partials in a fixed order (the gradient norms
//     of the bit allocation, conv_forward_model.py:424-448).
#pragma once
#include "qd_rowops.cuh"

namespace qd {

constexpr int kSelBins = 2048;
constexpr int kSelMaxRanks = 512;
constexpr int kSelThreads = 512;

struct SelectState {                 // lives in the workspace, after the histogram
    int num_slots;
    int rank_slot[kSelMaxRanks];     // slot of rank r
    int64_t rank_rest[kSelMaxRanks]; // residual rank inside the slot
    int64_t slot_offset[kSelMaxRanks + 1];
    int slot_cursor[kSelMaxRanks];
    short bin_slot[kSelBins];        // -1: bin not selected
};

__device__ __forceinline__ int select_bin(float v) {
    if (v != v) return kSelBins - 1;  // NaN sorts last, like numpy
    const float t = v * (float)kSelBins;
    int b = (t >= (float)kSelBins) ? kSelBins - 1 : (int)t;   // (int) of a huge / inf value is avoided
    return b < 0 ? 0 : b;
}
// NaN above +inf in key order too
__device__ __forceinline__ uint32_t select_key(float v) { return (v != v) ? 0xffffffffu : float_key(v); }

__global__ void __launch_bounds__(kSelThreads) select_hist_kernel(const float* __restrict__ v, int64_t n,
                                                                 unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_h[kSelBins];
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) s_h[i] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = (reinterpret_cast<uintptr_t>(v) & 15) == 0;
    const int64_t nv = vec ? (n >> 2) : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 t = ld_stream4(v + 4 * i);
        atomicAdd(&s_h[select_bin(t.x)], 1u);
        atomicAdd(&s_h[select_bin(t.y)], 1u);
        atomicAdd(&s_h[select_bin(t.z)], 1u);
        atomicAdd(&s_h[select_bin(t.w)], 1u);
    }
    for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&s_h[select_bin(v[i])], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
        if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);   // integer counts: order does not matter
}

__global__ void __launch_bounds__(kSelThreads) select_plan_kernel(const unsigned long long* __restrict__ hist,
                                                                 const int64_t* __restrict__ ranks, int num_ranks,
                                                                 SelectState* __restrict__ st) {
    __shared__ long long s_prefix[kSelBins + 1];
    __shared__ int s_first[kSelBins];   // first rank (in request order) that landed in the bin, or -1
    if (threadIdx.x == 0) {             // 2048 additions: a serial prefix is a few microseconds and trivially exact
        long long acc = 0;
        for (int b = 0; b < kSelBins; ++b) { s_prefix[b] = acc; acc += (long long)hist[b]; }
        s_prefix[kSelBins] = acc;
    }
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) { s_first[b] = -1; st->bin_slot[b] = -1; }
    __syncthreads();
    __shared__ int s_rank_bin[kSelMaxRanks];
    for (int r = threadIdx.x; r < num_ranks; r += blockDim.x) {
        long long k = ranks[r];
        k = k < 0 ? 0 : (k >= s_prefix[kSelBins] ? s_prefix[kSelBins] - 1 : k);
        int lo = 0, hi = kSelBins - 1;  // largest b with prefix[b] <= k
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_prefix[mid] <= k) lo = mid; else hi = mid - 1;
        }
        s_rank_bin[r] = lo;
        st->rank_rest[r] = k - s_prefix[lo];
    }
    __syncthreads();
    if (threadIdx.x == 0) {             // slots in order of first use: deterministic layout
        int slots = 0;
        long long off = 0;
        for (int r = 0; r < num_ranks; ++r) {
            const int b = s_rank_bin[r];
            if (s_first[b] < 0) {
                s_first[b] = slots;
                st->bin_slot[b] = (short)slots;
                st->slot_offset[slots] = off;
                st->slot_cursor[slots] = 0;
                off += s_prefix[b + 1] - s_prefix[b];
                ++slots;
            }
            st->rank_slot[r] = s_first[b];
        }
        st->slot_offset[slots] = off;
        st->num_slots = slots;
    }
}

__global__ void __launch_bounds__(kSelThreads) select_compact_kernel(const float* __restrict__ v, int64_t n,
                                                                    SelectState* __restrict__ st, uint32_t* __restrict__ buf) {
    __shared__ short s_slot[kSelBins];
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) s_slot[i] = st->bin_slot[i];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float t = v[i];
        const int s = s_slot[select_bin(t)];
        if (s >= 0) {
            const int pos = atomicAdd(&st->slot_cursor[s], 1);   // position inside the slot is arbitrary: selection is by value
            buf[st->slot_offset[s] + pos] = select_key(t);
        }
    }
}

// one CTA per requested rank: MSD radix select of the residual rank among the slot's keys
__global__ void __launch_bounds__(kSelThreads) select_final_kernel(const SelectState* __restrict__ st,
                                                                  const uint32_t* __restrict__ buf, float* __restrict__ out) {
    __shared__ unsigned int s_h[256];
    __shared__ uint32_t s_prefix;
    __shared__ long long s_rest;
    const int r = blockIdx.x;
    const int slot = st->rank_slot[r];
    const uint32_t* keys = buf + st->slot_offset[slot];
    const long long m = st->slot_offset[slot + 1] - st->slot_offset[slot];
    if (threadIdx.x == 0) { s_prefix = 0u; s_rest = st->rank_rest[r]; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_h[i] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t mask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
        for (long long i = threadIdx.x; i < m; i += blockDim.x) {
            const uint32_t k = keys[i];
            if ((k & mask) == prefix) atomicAdd(&s_h[(k >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long rest = s_rest;
            int d = 0;
            for (; d < 255; ++d) {
                if (rest < (long long)s_h[d]) break;
                rest -= (long long)s_h[d];
            }
            s_prefix = prefix | ((uint32_t)d << shift);
            s_rest = rest;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[r] = (s_prefix == 0xffffffffu) ? __int_as_float(0x7fc00000) : key_float(s_prefix);
}

// ---------------------------------------------------------------- multi-tensor L2 norm
constexpr int kNormChunk = 16384;
struct NormEntry {
    const float* ptr;
    int64_t n;
    int64_t chunk_start;
    int64_t chunks;
};

__global__ void __launch_bounds__(256) multi_norm_partial(const NormEntry* __restrict__ entries, int count, int64_t total_chunks,
                                                         double* __restrict__ partial) {
    __shared__ double s_w[8];
    for (int64_t c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        int lo = 0, hi = count - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (entries[mid].chunk_start <= c) lo = mid; else hi = mid - 1;
        }
        const NormEntry en = entries[lo];
        const int64_t start = (c - en.chunk_start) * kNormChunk;
        const int len = (int)min((int64_t)kNormChunk, en.n - start);
        double acc = 0.0;
        for (int e = threadIdx.x; e < len; e += 256) {
            const float t = en.ptr[start + e];
            acc += (double)t * (double)t;
        }
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += s_w[w];
            partial[c] = s;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) multi_norm_final(const NormEntry* __restrict__ entries, int count,
                                                       const double* __restrict__ partial, float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const NormEntry en = entries[t];
    double s = 0.0;
    for (int64_t c = 0; c < en.chunks; ++c) s += partial[en.chunk_start + c];
    out[t] = (float)sqrt(s);
}

}  // namespace qd
