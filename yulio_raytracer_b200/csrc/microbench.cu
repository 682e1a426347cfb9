// microbench.cu — the roofline denominators SURVEY §8d asks the builder to MEASURE on the box the benchmark runs on:
//   kind 0  FP32 issue: dependent-free FFMA chains on every SM                         -> TFLOP/s (2 flop per FFMA lane)
//   kind 1  read bandwidth of a working set of `bytes` re-read with 16-byte ld.global.cg (L1 bypassed): below the L2 size this is
//           the L2 -> SM read bandwidth that bounds an L2-resident BVH walk, far above it the HBM read bandwidth     -> GB/s
//   kind 2  warp-instruction issue rate: independent integer ops, one warp instruction per scheduler per clock        -> G warp-inst/s
// bench.py calls these once per run (yrtxMicrobench) and prints them inside `roofline`; none of it is on the render path.
#include "device_impl.hpp"

namespace yrt {

__global__ void __launch_bounds__(256) k_mb_fma(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;                           // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) k_mb_issue(unsigned* out, int iters, unsigned a) {
    unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x0 = (x0 ^ a) + x1; x1 = (x1 ^ a) + x2; x2 = (x2 ^ a) + x3; x3 = (x3 ^ a) + x4;
            x4 = (x4 ^ a) + x5; x5 = (x5 ^ a) + x6; x6 = (x6 ^ a) + x7; x7 = (x7 ^ a) + x0;
        }
    }
    const unsigned s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if (s == 0x12345u) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mb_read(const uint4* __restrict__ buf, size_t n16, int reps, unsigned* out) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
#pragma unroll 4
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345u) out[0] = acc;
}

double microbench(yrt_device* dev, int kind, size_t bytes) {
    cudaStream_t st = dev->stream;
    cudaEvent_t a, b; YRT_CK(cudaEventCreate(&a)); YRT_CK(cudaEventCreate(&b));
    DevBuf<unsigned> out; out.alloc(4);
    double result = 0.0; float ms = 0.f;
    if (kind == 0 || kind == 2) {
        const int blocks = dev->numSMs * 8, iters = 4096;
        for (int pass = 0; pass < 3; pass++) {               // pass 0 warms up; best of the other two
            YRT_CK(cudaEventRecord(a, st));
            if (kind == 0) k_mb_fma<<<blocks, 256, 0, st>>>((float*)out.p, iters, 1.0000001f, 1e-9f);
            else k_mb_issue<<<blocks, 256, 0, st>>>(out.p, iters, 0x9E3779B9u);
            YRT_CK(cudaEventRecord(b, st)); YRT_CK(cudaEventSynchronize(b)); YRT_CK(cudaEventElapsedTime(&ms, a, b));
            const double laneOps = (double)blocks * 256.0 * iters * 64.0;
            const double v = kind == 0 ? laneOps * 2.0 / (ms * 1e-3) / 1e12                 // TFLOP/s
                                       : laneOps * 2.0 / 32.0 / (ms * 1e-3) / 1e9;           // G warp-instructions/s (LOP3 + IADD per step)
            if (pass > 0 && v > result) result = v;
        }
    } else if (kind == 1) {
        if (bytes < (1u << 20)) bytes = 1u << 20;
        DevBuf<uint4> buf; buf.alloc(bytes / 16);
        YRT_CK(cudaMemsetAsync(buf.p, 1, bytes / 16 * 16, st));
        const size_t total = (size_t)8 << 30;                // ~8 GB of reads per timed pass
        const int reps = (int)std::max<size_t>(1, total / bytes);
        for (int pass = 0; pass < 3; pass++) {
            YRT_CK(cudaEventRecord(a, st));
            k_mb_read<<<dev->numSMs * 8, 256, 0, st>>>(buf.p, bytes / 16, reps, out.p);
            YRT_CK(cudaEventRecord(b, st)); YRT_CK(cudaEventSynchronize(b)); YRT_CK(cudaEventElapsedTime(&ms, a, b));
            const double v = (double)(bytes / 16 * 16) * reps / (ms * 1e-3) / 1e9;
            if (pass > 0 && v > result) result = v;
        }
    } else { cudaEventDestroy(a); cudaEventDestroy(b); throw std::runtime_error("device_cuda: unknown microbenchmark kind"); }
    YRT_CK(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b);
    return result;
}

}  // namespace yrt
