// Microbenchmark: per-sub-partition throughput of the instruction mix of the attention softmax (MUFU.EX2, FMNMX, F2FP, FFMA2)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_mix alu_mix.cu && ./alu_mix
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t *>(&h); }

// bit 0: 32 ex2, bit 1: 32 fminf, bit 2: 16 F2FP packs, bit 3: 32 fmaf
__global__ void __launch_bounds__(512, 1) k(long long *out, int iters, int nwarps, int mask, float s, float b)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = lane * 0.01f + q;
    uint32_t sink = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            float x[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                x[q] = v[q];
                if (mask & 8) x[q] = fmaf(x[q], s, b);
                if (mask & 2) x[q] = fminf(x[q], 120.f);
                if (mask & 1) x[q] = ex2(x[q]);
            }
            if (mask & 4) {
#pragma unroll
                for (int q = 0; q < 32; q += 2) sink ^= pack(x[q], x[q + 1]);
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) sink ^= __float_as_uint(x[q]);
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(__float_as_uint(v[q]) ^ (sink & 1));
        }
        t1 = clock64();
    }
    if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = sink; }
}

int main()
{
    long long *d, h[32];
    cudaMalloc(&d, sizeof(h));
    const int iters = 2000;
    const int masks[] = {0, 1, 2, 4, 8, 3, 5, 9, 15};
    const char *names[] = {"baseline (xor only)", "32 ex2", "32 fmin", "16 f2fp", "32 fma", "ex2+fmin", "ex2+f2fp", "ex2+fma", "ex2+fmin+f2fp+fma"};
    for (int m = 0; m < 9; ++m)
        for (int nw : {4, 8, 16}) {
            k<<<1, 512>>>(d, iters, nw, masks[m], 0.18f, -3.f);
            cudaDeviceSynchronize();
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w * 2] > mx ? h[w * 2] : mx;
            printf("%-22s warps/SMSP %d: %7.1f clk per 32-element chunk per warp, %6.1f clk per chunk per sub-partition\n", names[m], nw / 4,
                   (double)mx / iters, (double)mx / iters / (nw / 4));
        }
    return 0;
}
