// host_timeline.cu -- measurement only: the chunked H2D -> fused kernel -> D2H pipeline of qd_host.cu replayed with a CUDA
// event after every operation, so that the gaps of the two copy engines can be read off (no nsys in this image).
//   nvcc -O2 -o /tmp/host_timeline tools/host_timeline.cu -Iinclude -Lquantized_distillation_b200 -lqd_b200 \
//        -Xlinker -rpath=$PWD/quantized_distillation_b200 && /tmp/host_timeline <chunk MiB> <slots> <schedule>
// schedule: 0 = uniform chunks, 1 = tail ramp (.., 1/2, 1/4, 1/8, 1/8 chunk), 2 = head and tail ramp
// streams:  0 = one stream per slot (H2D, kernel, D2H of a chunk in order on it), 1 = one stream per engine (all H2D on one,
//           all kernels on a second, all D2H on a third; events carry the chunk's dependencies and the buffer reuse)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "qd_b200.h"

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const int chunk_mib = argc > 1 ? atoi(argv[1]) : 16;
    const int slots = argc > 2 ? atoi(argv[2]) : 3;
    const int schedule = argc > 3 ? atoi(argv[3]) : 0;
    const int per_engine = argc > 4 ? atoi(argv[4]) : 0;
    const int64_t n = 1 << 26, bucket = 256;
    const int64_t chunk = (int64_t)chunk_mib << 18;
    float *hx, *hg, *hq, *hgo;
    CK(cudaMallocHost(&hx, n * 4)); CK(cudaMallocHost(&hg, n * 4)); CK(cudaMallocHost(&hq, n * 4)); CK(cudaMallocHost(&hgo, n * 4));
    for (int64_t i = 0; i < n; ++i) { hx[i] = (float)((i * 2654435761u) % 1000) * 1e-4f - 0.05f; hg[i] = 1.0f; hq[i] = 0; hgo[i] = 0; }
    struct Slot { cudaStream_t s; float *x, *g, *q, *go; };
    std::vector<Slot> sl(slots);
    for (auto& s : sl) {
        CK(cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking));
        CK(cudaMalloc(&s.x, chunk * 4)); CK(cudaMalloc(&s.g, chunk * 4)); CK(cudaMalloc(&s.q, chunk * 4)); CK(cudaMalloc(&s.go, chunk * 4));
    }
    std::vector<int64_t> lens;
    {
        std::vector<int64_t> ramp = {chunk / 8, chunk / 8, chunk / 4, chunk / 2};
        int64_t left = n;
        if (schedule == 2) for (int64_t r : ramp) { lens.push_back(r); left -= r; }
        int64_t tail = 0;
        if (schedule >= 1) for (int64_t r : ramp) tail += r;
        for (left -= tail; left > 0; left -= chunk) lens.push_back(left < chunk ? left : chunk);
        if (schedule >= 1) for (int i = 3; i >= 0; --i) lens.push_back(ramp[i]);
    }
    const int C = (int)lens.size();
    std::vector<cudaEvent_t> ev(C * 3 + 1);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    cudaStream_t s_in, s_k, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(ev[C * 3], per_engine ? s_in : sl[0].s));
        if (per_engine) { CK(cudaStreamWaitEvent(s_k, ev[C * 3], 0)); CK(cudaStreamWaitEvent(s_out, ev[C * 3], 0)); }
        int64_t off = 0;
        for (int k = 0; k < C; ++k) {
            Slot& s = sl[k % slots];
            const int64_t len = lens[k];
            cudaStream_t a = per_engine ? s_in : s.s, b = per_engine ? s_k : s.s, c = per_engine ? s_out : s.s;
            if (per_engine && k >= slots) CK(cudaStreamWaitEvent(a, ev[(k - slots) * 3 + 1], 0));   // inputs of the slot consumed
            CK(cudaMemcpyAsync(s.x, hx + off, len * 4, cudaMemcpyHostToDevice, a));
            CK(cudaMemcpyAsync(s.g, hg + off, len * 4, cudaMemcpyHostToDevice, a));
            CK(cudaEventRecord(ev[k * 3], a));
            if (per_engine) {
                CK(cudaStreamWaitEvent(b, ev[k * 3], 0));
                if (k >= slots) CK(cudaStreamWaitEvent(b, ev[(k - slots) * 3 + 2], 0));               // outputs of the slot copied out
            }
            if (qd_uniform_fwd_bwd(s.x, s.g, s.q, s.go, len, bucket, 16, QD_BWD_MINMAX, nullptr, 0, (qd_stream_t)b)) { printf("%s\n", qd_last_error()); return 1; }
            CK(cudaEventRecord(ev[k * 3 + 1], b));
            if (per_engine) CK(cudaStreamWaitEvent(c, ev[k * 3 + 1], 0));
            CK(cudaMemcpyAsync(hq + off, s.q, len * 4, cudaMemcpyDeviceToHost, c));
            CK(cudaMemcpyAsync(hgo + off, s.go, len * 4, cudaMemcpyDeviceToHost, c));
            CK(cudaEventRecord(ev[k * 3 + 2], c));
            off += len;
        }
        CK(cudaDeviceSynchronize());
    }
    printf("chunk %d MiB, %d slots, schedule %d, %s: %d chunks\n chunk MiB | h2d done | kernel done | d2h done   (ms since start; last repetition)\n", chunk_mib, slots, schedule, per_engine ? "one stream per engine" : "one stream per slot", C);
    float last = 0;
    for (int k = 0; k < C; ++k) {
        float a, b, c;
        CK(cudaEventElapsedTime(&a, ev[C * 3], ev[k * 3]));
        CK(cudaEventElapsedTime(&b, ev[C * 3], ev[k * 3 + 1]));
        CK(cudaEventElapsedTime(&c, ev[C * 3], ev[k * 3 + 2]));
        printf(" %3d %5.1f | %7.3f | %7.3f | %7.3f\n", k, lens[k] / 262144.0, a, b, c);
        last = c;
    }
    printf("total %.3f ms = %.2f GB/s\n", last, n * 16 / (last * 1e-3) / 1e9);
    return 0;
}
