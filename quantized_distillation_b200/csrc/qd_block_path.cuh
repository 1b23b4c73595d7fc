// qd_block_path.cuh -- rows of 1025 .. QD_MAX_STAGED_BUCKET elements: ONE CTA
// PER ROW, the row is staged in shared memory so that it still crosses HBM once.
//
// Staging uses the TMA bulk-copy engine (cp.async.bulk.shared::cluster.global
// with mbarrier complete_tx; SASS UBLKCP): one elected thread enqueues the row
// in 32 KB chunks, each chunk signalling its own mbarrier, and the 512 threads
// start the min/max reduction of chunk c while chunks c+1.. are still in
// flight.  Rows whose global address is not 16-byte aligned fall back to a
// cooperative ld.global -> st.shared copy.  Shared memory is sized to the row
// (dynamic), so short rows give several resident CTAs per SM and the store
// phase of one row overlaps the load phase of another.
#pragma once
#include "qd_rowops.cuh"

namespace qd {

constexpr int kBlockCtaThreads = 512;
constexpr int kStageChunk = 8192;  // floats per TMA bulk copy (32 KB)
constexpr int kMaxStageChunks = (QD_MAX_STAGED_BUCKET + kStageChunk - 1) / kStageChunk;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// CTA-wide reductions through a small shared scratch (16 warps)
template <bool IS_MIN>
__device__ __forceinline__ float cta_minmax(float v, float* scratch) {
    v = IS_MIN ? warp_min(v) : warp_max(v);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float r = scratch[lane < (blockDim.x >> 5) ? lane : 0];
    return IS_MIN ? warp_min(r) : warp_max(r);
}
__device__ __forceinline__ int cta_min_int(int v, int* scratch) {
    v = warp_min_int(v);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    int r = scratch[lane < (blockDim.x >> 5) ? lane : 0];
    return warp_min_int(r);
}
__device__ __forceinline__ double cta_sum(double v, double* scratch) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = (lane < (blockDim.x >> 5)) ? scratch[lane] : 0.0;
    return warp_sum(r);  // fixed tree: deterministic
}

template <int OP, int BWD>
__global__ void __launch_bounds__(kBlockCtaThreads) block_rows_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(128) float s_row[];
    __shared__ __align__(8) uint64_t s_bar[kMaxStageChunks];
    __shared__ float s_k[OP == OP_NONUNIFORM ? 256 : 1];
    __shared__ float s_m[OP == OP_NONUNIFORM ? 256 : 1];
    __shared__ double s_scratch[kBlockCtaThreads / 32];

    Centroids cen{s_k, s_m, P.num_points};
    if constexpr (OP == OP_NONUNIFORM) centroid_setup(s_k, s_m, P.points, P.num_points);
    if (threadIdx.x == 0) {
        for (int c = 0; c < kMaxStageChunks; ++c) mbar_init(&s_bar[c], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int tid = threadIdx.x;
    const bool pre = (P.mean != nullptr) || (P.max_element > 0.f);
    const float mean = P.mean ? *P.mean : 0.f;
    uint32_t phase_bits = 0;  // bit c = parity the next wait on s_bar[c] must see

    for (int64_t row = blockIdx.x; row < P.geo.rows; row += gridDim.x) {
        const int64_t base = row * P.geo.row_len;
        const int len = (int)min(P.geo.row_len, P.geo.n - base);
        const float* src = P.x + base;
        const bool tma_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const int bulk_len = tma_ok ? (len & ~3) : 0;  // multiple of 16 bytes
        const int nchunks = (bulk_len + kStageChunk - 1) / kStageChunk;

        // ---- stage the row -------------------------------------------------
        if (tid == 0 && nchunks > 0) {
            // generic-proxy reads of the previous row are complete (barrier at loop end);
            // order them before the async-proxy writes that follow
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int c = 0; c < nchunks; ++c) {
                const int off = c * kStageChunk;
                const uint32_t bytes = (uint32_t)min(kStageChunk, bulk_len - off) * 4u;
                mbar_expect_tx(&s_bar[c], bytes);
                tma_bulk_g2s(s_row + off, src + off, bytes, &s_bar[c]);
            }
        }
        for (int e = bulk_len + tid; e < len; e += kBlockCtaThreads) s_row[e] = ld_stream1(src + e);

        // ---- min / max, chunk by chunk as the copies land ---------------------
        float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&s_bar[c], (phase_bits >> c) & 1u);
            phase_bits ^= (1u << c);
            const int off = c * kStageChunk;
            const int cl = min(kStageChunk, bulk_len - off);
            for (int e = tid * 4; e < cl; e += kBlockCtaThreads * 4) {
                float4 t = *reinterpret_cast<const float4*>(s_row + off + e);
                if (pre) {
                    t.x = pre_op(t.x, mean, P.max_element); t.y = pre_op(t.y, mean, P.max_element);
                    t.z = pre_op(t.z, mean, P.max_element); t.w = pre_op(t.w, mean, P.max_element);
                    *reinterpret_cast<float4*>(s_row + off + e) = t;
                }
                mn = min_nan(min_nan(mn, t.x), min_nan(t.y, min_nan(t.z, t.w)));
                mx = max_nan(max_nan(mx, t.x), max_nan(t.y, max_nan(t.z, t.w)));
            }
        }
        __syncthreads();  // scalar-staged tail visible
        for (int e = bulk_len + tid; e < len; e += kBlockCtaThreads) {
            float t = s_row[e];
            if (pre) { t = pre_op(t, mean, P.max_element); s_row[e] = t; }
            mn = min_nan(mn, t);
            mx = max_nan(mx, t);
        }
        mn = cta_minmax<true>(mn, reinterpret_cast<float*>(s_scratch));
        mx = cta_minmax<false>(mx, reinterpret_cast<float*>(s_scratch));
        RowState rs;
        rs.mean = mean;
        rs.beta = mn;
        rs.alpha = make_alpha(mn, mx);
        if (P.alpha != nullptr && tid == 0) { P.alpha[row] = rs.alpha; P.beta[row] = rs.beta; }
        if (P.argmin != nullptr) {
            int imin = 0x7fffffff, imax = 0x7fffffff;
            for (int e = tid; e < len; e += kBlockCtaThreads) {
                float t = s_row[e];
                if (t == mn) imin = min(imin, e);
                if (t == mx) imax = min(imax, e);
            }
            imin = cta_min_int(imin, reinterpret_cast<int*>(s_scratch));
            imax = cta_min_int(imax, reinterpret_cast<int*>(s_scratch));
            if (tid == 0) {
                P.argmin[row] = (imin == 0x7fffffff) ? 0 : imin;
                P.argmax[row] = (imax == 0x7fffffff) ? 0 : imax;
            }
        }

        // ---- element-wise pass from shared memory ----------------------------
        if constexpr (OP == OP_SCALE) {
            const int plen = (int)P.geo.row_len;
            const float last = to_unit(s_row[len - 1], rs.beta, rs.alpha);
            float* dst = P.xhat + base;
            for (int e = tid; e < plen; e += kBlockCtaThreads)
                st_stream1(dst + e, e < len ? to_unit(s_row[e], rs.beta, rs.alpha) : last);
        } else if constexpr (OP == OP_UNIFORM) {
            float rb = 0.f;
            int imin2 = 0, imax2 = 0;
            const UniformFast uf = make_uniform_fast(rs.alpha, P.S);
            if constexpr (BWD == BWD_MINMAX) {
                float qmn = __int_as_float(0x7f800000), qmx = __int_as_float(0xff800000);
                for (int e = tid; e < len; e += kBlockCtaThreads) {
                    float lvl;
                    float qv = uniform_quantize(s_row[e], rs, P.S, lvl);
                    qmn = min_nan(qmn, qv);
                    qmx = max_nan(qmx, qv);
                }
                qmn = cta_minmax<true>(qmn, reinterpret_cast<float*>(s_scratch));
                qmx = cta_minmax<false>(qmx, reinterpret_cast<float*>(s_scratch));
                rs.beta2 = qmn;
                rs.alpha2 = make_alpha(qmn, qmx);
                double acc = 0.0;
                imin2 = 0x7fffffff; imax2 = 0x7fffffff;
                for (int e = tid; e < len; e += kBlockCtaThreads) {
                    float lvl;
                    float xv = s_row[e];
                    float qv = uniform_quantize(xv, rs, P.S, lvl);
                    if (qv == qmn) imin2 = min(imin2, e);
                    if (qv == qmx) imax2 = min(imax2, e);
                    acc += (double)minmax_term(xv, qv, P.g[base + e], rs);
                }
                imin2 = cta_min_int(imin2, reinterpret_cast<int*>(s_scratch));
                imax2 = cta_min_int(imax2, reinterpret_cast<int*>(s_scratch));
                rb = (float)cta_sum(acc, s_scratch);
            }
            for (int e = tid; e < len; e += kBlockCtaThreads) {
                float lvl, qv;
                const float xv = s_row[e];
                if (P.stochastic) {
                    Philox rng(P.seed);
                    const int64_t ge = base + e;
                    uint4 rnd = rng(P.offset + (uint64_t)(ge >> 2));
                    uint32_t w = (ge & 3) == 0 ? rnd.x : (ge & 3) == 1 ? rnd.y : (ge & 3) == 2 ? rnd.z : rnd.w;
                    qv = uniform_quantize_stochastic(xv, rs, P.S, u01(w), lvl);
                } else {
                    qv = uniform_quantize_auto(xv, rs, uf, P.S, P.rS, P.half_minus_band, lvl);
                }
                if constexpr (BWD != BWD_OFF) {
                    float gv = P.g[base + e];
                    if constexpr (BWD == BWD_TRUNC) gv = (fabsf(xv) > 1.0f) ? 0.f : gv;
                    if (BWD == BWD_MINMAX && imin2 != imax2) {
                        if (e == imax2) gv = __fadd_rn(gv, rb);
                        if (e == imin2) gv = __fadd_rn(gv, -rb);
                    }
                    st_stream1(P.gout + base + e, gv);
                }
                if (P.q != nullptr) st_stream1(P.q + base + e, pre ? __fadd_rn(qv, mean) : qv);
                if (P.idx8 != nullptr) P.idx8[base + e] = (uint8_t)(int)lvl;
            }
        } else if constexpr (OP == OP_NONUNIFORM) {
            for (int e = tid; e < len; e += kBlockCtaThreads) {
                float xh = to_unit(s_row[e], rs.beta, rs.alpha);
                int id = centroid_index(cen, xh, P.rule);
                float qv = from_unit(cen.k[id], rs.alpha, rs.beta);
                if (P.q != nullptr) st_stream1(P.q + base + e, pre ? __fadd_rn(qv, mean) : qv);
                if (P.idx8 != nullptr) P.idx8[base + e] = (uint8_t)id;
                if (P.idx64 != nullptr) P.idx64[base + e] = id;
            }
        }
        __syncthreads();  // everyone is done with s_row before the next row is staged
    }
}

}  // namespace qd
