// Microbenchmark: tcgen05.st cost, alone and interleaved with tcgen05.ld as in the attention softmax (read 32 columns of S,
// write 16 packed columns of P).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_st tmem_st.cu && ./tmem_st
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// mode 0: st only; 1: ld + wait + st (data dependent: st of values derived from the load); 2: as 1 plus 32 MUFU.EX2 per load;
// 3: 32 MUFU only (no tensor memory); 4: ld + wait + 32 MUFU (no st); 5: as 2 with wait::st every iteration
__global__ void __launch_bounds__(544, 1) k(long long *out, int iters, int nwarps, int mode)
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
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 1) * 256u;
    uint32_t sink = lane;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        uint32_t v[32], pk[16];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = lane * 33 + q;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t c = (uint32_t)(i % 7);
            if (mode == 1 || mode == 2 || mode == 4 || mode == 5) {
                ld32(base + c * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            if (mode == 2 || mode == 3 || mode == 4 || mode == 5) {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(ex2(__uint_as_float(v[q]) * 1e-30f));
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) pk[q] = v[2 * q] ^ (v[2 * q + 1] << 16);
            if (mode == 0 || mode == 1 || mode == 2 || mode == 5) st16(base + c * 16, pk);
            else
#pragma unroll
                for (int q = 0; q < 16; ++q) sink ^= pk[q];
            if (mode == 5) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (mode == 0)
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] += 1;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        t1 = clock64();
#pragma unroll
        for (int q = 0; q < 32; ++q) sink ^= v[q];
    }
    if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = sink; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main()
{
    const char *names[] = {"st16 only", "ld32+st16", "ld32+32 ex2+st16", "32 ex2 only", "ld32+32 ex2", "ld32+ex2+st16+wait::st"};
    long long *d, h[40];
    cudaMalloc(&d, sizeof(h));
    const int iters = 2000;
    for (int mode = 0; mode < 6; ++mode)
        for (int nw : {1, 4, 8, 16}) {
            k<<<1, 544>>>(d, iters, nw, mode);
            k<<<1, 544>>>(d, iters, nw, mode);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w * 2] > mx ? h[w * 2] : mx;
            printf("%-26s warps %2d: %7.1f clk per iteration (slowest warp)  (%s)\n", names[mode], nw, (double)mx / iters, cudaGetErrorString(e));
        }
    return 0;
}
