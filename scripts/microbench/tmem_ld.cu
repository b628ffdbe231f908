// Microbenchmark: tensor-memory -> register read rate of tcgen05.ld shapes on sm_100a (cycles per load, bytes per clock and SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(SHAPE, taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 " \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), \
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), \
          "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory")

// mode 0: 32x32b.x32 (4 KB / warp), 1: 16x256b.x8 (32 regs: 16 lanes x 256 B = 4 KB), 2: 16x128b.x16, 3: 32x32b.x16 twice
// fma_per_ld: independent FMA work issued between load and wait (tests overlap with math of the same warp)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(long long *out, int iters, int nwarps, int fma_per_ld, int two_in_flight)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc0 = lane, acc1 = lane + 1, acc2 = lane + 2, acc3 = lane + 3;
    uint32_t sink = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        uint32_t v[32], w[32];
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t col = (uint32_t)((i * 32) & 255) + (uint32_t)((warp >> 2) & 1) * 256u;
            if (MODE == 0) LD32("32x32b.x32", base + col, v);
            else if (MODE == 1) LD32("16x256b.x8", base + col, v);     // half the lanes (16) of the quarter
            else if (MODE == 2) LD32("16x128b.x16", base + col, v);
            if (two_in_flight) {
                if (MODE == 0) LD32("32x32b.x32", base + ((col + 32) & 511), w);
                else if (MODE == 1) LD32("16x256b.x8", base + ((col + 32) & 511), w);
                else LD32("16x128b.x16", base + ((col + 32) & 511), w);
            }
            for (int f = 0; f < fma_per_ld; ++f) {
                acc0 = fmaf(acc0, 1.0001f, 0.5f); acc1 = fmaf(acc1, 1.0001f, 0.5f);
                acc2 = fmaf(acc2, 1.0001f, 0.5f); acc3 = fmaf(acc3, 1.0001f, 0.5f);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 32; ++q) sink ^= v[q];
            if (two_in_flight)
#pragma unroll
                for (int q = 0; q < 32; ++q) sink ^= w[q];
        }
        t1 = clock64();
    }
    if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = (long long)sink + (long long)(acc0 + acc1 + acc2 + acc3); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int MODE>
void run(const char *name, int nwarps, int fma, int two)
{
    long long *d, h[32];
    cudaMalloc(&d, sizeof(h));
    const int iters = 2000;
    k<MODE><<<1, 512>>>(d, iters, nwarps, fma, two);
    k<MODE><<<1, 512>>>(d, iters, nwarps, fma, two);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < nwarps; ++w) mx = h[w * 2] > mx ? h[w * 2] : mx;
    const double loads = (double)iters * (two ? 2 : 1);
    printf("%-12s warps %2d fma/ld %4d in-flight %d: %7.1f clk per load per warp, %6.1f B/clk/SM  (%s)\n", name, nwarps, fma * 4, two ? 2 : 1,
           mx / loads, 4096.0 * loads * nwarps / mx, cudaGetErrorString(e));
    cudaFree(d);
}

int main()
{
    for (int nw : {1, 4, 8, 16}) run<0>("32x32b.x32", nw, 0, 0);
    for (int nw : {4, 8}) run<0>("32x32b.x32", nw, 0, 1);
    for (int f : {16, 32, 64, 128}) run<0>("32x32b.x32", 4, f, 0);
    for (int f : {32, 64}) run<0>("32x32b.x32", 8, f, 0);
    for (int nw : {1, 4, 8}) run<1>("16x256b.x8", nw, 0, 0);
    for (int nw : {1, 4, 8}) run<2>("16x128b.x16", nw, 0, 0);
    return 0;
}
