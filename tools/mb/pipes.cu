// Development microbenchmark: which pipe executes PRMT / IDP.4A / IMAD.HI / FMNMX / FFMA on sm_100a, measured as warp instructions per clock per SM
// for each instruction alone and for mixtures (a mixture that runs faster than the sum of its parts uses two pipes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_pipes tools/mb/pipes.cu && ./mb_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE> __global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, uint32_t magic, uint32_t mul256) {
    uint32_t a[8]; float f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = seed * (threadIdx.x + 1 + j); f[j] = (float)(threadIdx.x + j); }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 8) asm volatile("prmt.b32 %0, %0, %1, 0x7651;" : "+r"(a[j]) : "r"(magic));
            if (MODE == 1 || MODE == 4) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(0x00000100u), "r"(magic));
            if (MODE == 2 || MODE == 5) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(mul256), "r"(magic));
            if (MODE == 3 || MODE == 6 || MODE == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(1.5f));
            if (MODE == 6 || MODE == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(1.0001f), "f"(0.5f));
            if (MODE == 9) { asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a[j]) : "r"(magic)); }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r += a[j] + __float_as_uint(f[j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int instPerSlot, uint32_t* out) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(out, 3u, 0x4B000000u, 256u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, 3u, 0x4B000000u, 256u); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double warpInst = (double)blocks * 8 * ITER * 8 * instPerSlot;
    printf("%-28s %8.3f ms  %7.1f G warp-inst/s  = %5.2f inst/clk/SM at the nominal %d MHz\n", name, best, warpInst / best / 1e6, warpInst / (best * 1e-3) / sms / (khz * 1e3), khz / 1000);
}
int main() {
    uint32_t* out; cudaMalloc(&out, 148 * 16 * 256 * 4);
    run<0>("PRMT", 1, out); run<1>("IDP.4A", 1, out); run<2>("IMAD.HI", 1, out); run<3>("FMNMX", 1, out); run<7>("FFMA", 1, out); run<9>("SHF", 1, out);
    run<4>("PRMT + IDP.4A", 2, out); run<5>("PRMT + IMAD.HI", 2, out); run<6>("FMNMX + FFMA", 2, out); run<8>("PRMT + FMNMX", 2, out);
    return cudaGetLastError() != cudaSuccess;
}
