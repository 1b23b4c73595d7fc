// qd_host.cu -- host-buffer entry points: what a caller holding CPU tensors
// gets (the reference's functions accept CPU tensors; here the arithmetic still
// runs on the GPU).  The tensor is cut into row-aligned chunks and each chunk
// travels H2D -> fused kernel -> D2H on one of three streams, so the two PCIe
// directions and the kernel overlap; end-to-end time is bounded by the slower
// PCIe direction, not by their sum.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "qd_b200.h"

// message of the calling thread (qd_last_error), defined in qd_api.cu
extern "C" int qd_internal_fail(int code, const char* fmt, ...);

#define QDH_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return qd_internal_fail(QD_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

namespace {

// pipeline depth and chunk size: three slots of 16 MiB per buffer sit on the plateau of the tuning
// sweep of round 1 (tools/e2e_tune.py: ~44 GB/s per PCIe direction from 2 slots x 4 MiB upwards)
constexpr int kMaxSlots = 8;
constexpr int kSlots = 3;
constexpr int64_t kChunkElems = 4 << 20;

struct Slot {
    cudaStream_t stream = nullptr;
    float *x = nullptr, *g = nullptr, *q = nullptr, *gout = nullptr;
    void* ws = nullptr;
    size_t ws_bytes = 0;
};

struct HostCtx {
    bool ready = false;
    Slot slot[kMaxSlots];
    float *big_x = nullptr, *big_g = nullptr, *big_q = nullptr, *big_gout = nullptr;  // bucket=None path
    int64_t big_cap = 0;
    void* big_ws = nullptr;
    size_t big_ws_bytes = 0;
};

HostCtx g_ctx[64];
std::mutex g_mu[64];  // one lock per device: callers on different GPUs never wait for each other

// the caller's current device is put back when the call returns, whatever the path
struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ensure_ctx(int device, HostCtx** out) {
    HostCtx& c = g_ctx[device];
    if (!c.ready) {
        for (int i = 0; i < kSlots; ++i) {
            Slot& s = c.slot[i];
            QDH_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            const size_t bytes = (size_t)kChunkElems * sizeof(float);
            QDH_CUDA(cudaMalloc(&s.x, bytes));
            QDH_CUDA(cudaMalloc(&s.g, bytes));
            QDH_CUDA(cudaMalloc(&s.q, bytes));
            QDH_CUDA(cudaMalloc(&s.gout, bytes));
            s.ws_bytes = qd_workspace_bytes(kChunkElems, 0);
            QDH_CUDA(cudaMalloc(&s.ws, s.ws_bytes));
        }
        c.ready = true;
    }
    *out = &c;
    return QD_OK;
}

int run_host(const float* hx, const float* hg, float* hq, float* hgout, int64_t n, int64_t bucket, int levels, int mode,
             int device) {
    if (hx == nullptr || hq == nullptr || n <= 0 || bucket < 0)
        return qd_internal_fail(QD_ERR_INVALID_ARG, "host entry point: NULL buffer, n <= 0 or bucket < 0 (n=%lld bucket=%lld)", (long long)n, (long long)bucket);
    const bool bwd = hg != nullptr;
    if (bwd && hgout == nullptr) return qd_internal_fail(QD_ERR_INVALID_ARG, "gout_host is NULL");
    if (device < 0 || device >= 64) return qd_internal_fail(QD_ERR_INVALID_ARG, "device ordinal %d out of range", device);
    DeviceGuard guard;
    QDH_CUDA(cudaGetDevice(&guard.prev));
    QDH_CUDA(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_mu[device]);
    HostCtx* c;
    int rc = ensure_ctx(device, &c);
    if (rc) return rc;

    int64_t rows, row_len, padded;
    rc = qd_bucket_geometry(n, bucket, &rows, &row_len, &padded);
    if (rc) return rc;

    if (row_len > kChunkElems) {
        // one row spans more than a chunk (bucket None on a large tensor): no row-aligned
        // cut exists, so stage the whole tensor; copies still run at PCIe rate.
        if (c->big_cap < n) {
            cudaFree(c->big_x); cudaFree(c->big_g); cudaFree(c->big_q); cudaFree(c->big_gout); cudaFree(c->big_ws);
            const size_t bytes = (size_t)n * sizeof(float);
            c->big_ws_bytes = qd_workspace_bytes(n, bucket);
            c->big_cap = 0;
            c->big_x = c->big_g = c->big_q = c->big_gout = nullptr;
            c->big_ws = nullptr;
            QDH_CUDA(cudaMalloc(&c->big_x, bytes));
            QDH_CUDA(cudaMalloc(&c->big_g, bytes));
            QDH_CUDA(cudaMalloc(&c->big_q, bytes));
            QDH_CUDA(cudaMalloc(&c->big_gout, bytes));
            QDH_CUDA(cudaMalloc(&c->big_ws, c->big_ws_bytes));
            c->big_cap = n;
        }
        cudaStream_t s = c->slot[0].stream;
        const size_t bytes = (size_t)n * sizeof(float);
        QDH_CUDA(cudaMemcpyAsync(c->big_x, hx, bytes, cudaMemcpyHostToDevice, s));
        if (bwd) QDH_CUDA(cudaMemcpyAsync(c->big_g, hg, bytes, cudaMemcpyHostToDevice, s));
        rc = bwd ? qd_uniform_fwd_bwd(c->big_x, c->big_g, c->big_q, c->big_gout, n, bucket, levels, mode, c->big_ws,
                                      c->big_ws_bytes, s)
                 : qd_uniform_fwd(c->big_x, c->big_q, nullptr, nullptr, nullptr, nullptr, nullptr, n, bucket, levels,
                                  nullptr, 0.f, 0, 0, 0, c->big_ws, c->big_ws_bytes, s);
        if (rc) return rc;
        QDH_CUDA(cudaMemcpyAsync(hq, c->big_q, bytes, cudaMemcpyDeviceToHost, s));
        if (bwd) QDH_CUDA(cudaMemcpyAsync(hgout, c->big_gout, bytes, cudaMemcpyDeviceToHost, s));
        QDH_CUDA(cudaStreamSynchronize(s));
        return QD_OK;
    }

    // row-aligned chunks; the kernel sees each chunk as an independent tensor with the
    // same bucket size, which is exact because rows never straddle a chunk boundary and
    // only the last chunk holds the (short) tail row.
    const int64_t rows_per_chunk = kChunkElems / row_len;
    const int64_t chunk = rows_per_chunk * row_len;
    int k = 0;
    for (int64_t off = 0; off < n; off += chunk, ++k) {
        Slot& s = c->slot[k % kSlots];
        const int64_t len = (n - off < chunk) ? (n - off) : chunk;
        const size_t bytes = (size_t)len * sizeof(float);
        // a chunk shorter than the bucket must still be bucketed like the tail of the
        // full tensor: with rows >= 2 overall the tail row is "padded", never a short
        // single row -- both cases give the same min/max, so passing bucket is exact.
        QDH_CUDA(cudaMemcpyAsync(s.x, hx + off, bytes, cudaMemcpyHostToDevice, s.stream));
        if (bwd) QDH_CUDA(cudaMemcpyAsync(s.g, hg + off, bytes, cudaMemcpyHostToDevice, s.stream));
        rc = bwd ? qd_uniform_fwd_bwd(s.x, s.g, s.q, s.gout, len, bucket, levels, mode, s.ws, s.ws_bytes, s.stream)
                 : qd_uniform_fwd(s.x, s.q, nullptr, nullptr, nullptr, nullptr, nullptr, len, bucket, levels, nullptr,
                                  0.f, 0, 0, 0, s.ws, s.ws_bytes, s.stream);
        if (rc) return rc;
        QDH_CUDA(cudaMemcpyAsync(hq + off, s.q, bytes, cudaMemcpyDeviceToHost, s.stream));
        if (bwd) QDH_CUDA(cudaMemcpyAsync(hgout + off, s.gout, bytes, cudaMemcpyDeviceToHost, s.stream));
    }
    for (int i = 0; i < kSlots; ++i) QDH_CUDA(cudaStreamSynchronize(c->slot[i].stream));
    QDH_CUDA(cudaGetLastError());
    return QD_OK;
}

}  // namespace

extern "C" int qd_uniform_fwd_host(const float* x_host, float* q_host, int64_t n, int64_t bucket, int levels, int device) {
    return run_host(x_host, nullptr, q_host, nullptr, n, bucket, levels, 0, device);
}

extern "C" int qd_uniform_fwd_bwd_host(const float* x_host, const float* g_host, float* q_host, float* gout_host,
                                       int64_t n, int64_t bucket, int levels, int mode, int device) {
    if (g_host == nullptr) return qd_internal_fail(QD_ERR_INVALID_ARG, "g_host is NULL");
    return run_host(x_host, g_host, q_host, gout_host, n, bucket, levels, mode, device);
}
