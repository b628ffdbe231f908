// attention.cu -- softmax(q k^T / sqrt(64)) v for the CLIP ViT blocks (nn.MultiheadAttention inside openai-CLIP's
// ResidualAttentionBlock [3P]; LoRA variant models/lora.py:165-303 computes the same core with merged weights).
//
// Input is the packed in_proj output qkv bf16 [n_img*L, 3d] (q | k | v; head h at columns h*64..h*64+63), output
// bf16 [n_img*L, d].  L is 50 / 197 / 257 tokens, head_dim is 64 for every CLIP ViT.
// One CTA per (image, head): K and V of the whole sequence staged once in shared memory (cp.async), each of the
// 8 warps owns 16-query row tiles and runs an online-softmax loop over 64-key chunks on bf16 mma.sync tiles.
// (4 % of the encoder FLOPs; the tcgen05 budget goes to the GEMMs -- see DESIGN.md.)
#include "common.cuh"

namespace {

constexpr int HD = 64;        // head dim
constexpr int LDS = 72;       // smem row stride in elements (144 B: conflict-free ldmatrix)
constexpr int WARPS = 8;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *smem)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__global__ void __launch_bounds__(WARPS * 32) attention_kernel(const __nv_bfloat16 *__restrict__ qkv,
                                                               __nv_bfloat16 *__restrict__ out, int L, int heads)
{
    const int img = blockIdx.y, h = blockIdx.x;
    const int d = heads * HD;
    const int ld = 3 * d;
    const int Lp = (L + 63) & ~63;                 // keys padded to whole 64-key chunks
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16 *sK = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *sV = sK + (size_t)Lp * LDS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const __nv_bfloat16 *base = qkv + (size_t)img * L * ld + h * HD;

    // stage K and V (rows >= L are zero so that 0-probability keys contribute exact zeros)
    for (int i = tid; i < Lp * 8; i += WARPS * 32) {
        const int row = i >> 3, c = (i & 7) * 8;
        if (row < L) {
            cp_async16(sK + row * LDS + c, base + (size_t)row * ld + d + c);
            cp_async16(sV + row * LDS + c, base + (size_t)row * ld + 2 * d + c);
        } else {
            *reinterpret_cast<uint4 *>(sK + row * LDS + c) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(sV + row * LDS + c) = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();

    const float sl2 = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
    const int n_rt = (L + 15) / 16;
    for (int rt = warp; rt < n_rt; rt += WARPS) {
        const int r0 = rt * 16 + g, r1 = r0 + 8;
        // Q fragments straight from global: 4 k-steps of 16 dims
        uint32_t qa[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int c = kk * 16 + 2 * tig;
            qa[kk][0] = r0 < L ? *reinterpret_cast<const uint32_t *>(base + (size_t)r0 * ld + c) : 0u;
            qa[kk][1] = r1 < L ? *reinterpret_cast<const uint32_t *>(base + (size_t)r1 * ld + c) : 0u;
            qa[kk][2] = r0 < L ? *reinterpret_cast<const uint32_t *>(base + (size_t)r0 * ld + c + 8) : 0u;
            qa[kk][3] = r1 < L ? *reinterpret_cast<const uint32_t *>(base + (size_t)r1 * ld + c + 8) : 0u;
        }
        float o[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

        for (int kc = 0; kc < Lp; kc += 64) {
            // S = Q K^T for 64 keys: 8 n-tiles of 8 keys
            float s[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
                for (int kk2 = 0; kk2 < 2; ++kk2) {
                    uint32_t b[4];
                    ldmatrix_x4(b, sK + (size_t)(kc + j * 8 + (lane & 7)) * LDS + kk2 * 32 + (lane >> 3) * 8);
                    mma_bf16(s[j], qa[kk2 * 2], b[0], b[1]);
                    mma_bf16(s[j], qa[kk2 * 2 + 1], b[2], b[3]);
                }
            }
            // mask padded keys, chunk row max
            float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = kc + j * 8 + 2 * tig;
                if (key >= L) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
                if (key + 1 >= L) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
                cm0 = fmaxf(cm0, fmaxf(s[j][0], s[j][1]));
                cm1 = fmaxf(cm1, fmaxf(s[j][2], s[j][3]));
            }
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
            const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);   // finite: every chunk holds >= 1 real key
            const float a0 = exp2f((m0 - nm0) * sl2), a1 = exp2f((m1 - nm1) * sl2);
            m0 = nm0; m1 = nm1;
            float rs0 = 0.f, rs1 = 0.f;
            uint32_t pa[4][4];   // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p0 = exp2f((s[j][0] - m0) * sl2), p1 = exp2f((s[j][1] - m0) * sl2);
                const float p2 = exp2f((s[j][2] - m1) * sl2), p3 = exp2f((s[j][3] - m1) * sl2);
                rs0 += p0 + p1; rs1 += p2 + p3;
                const int kk = j >> 1;
                if ((j & 1) == 0) { pa[kk][0] = pack_bf16(p0, p1); pa[kk][1] = pack_bf16(p2, p3); }
                else              { pa[kk][2] = pack_bf16(p0, p1); pa[kk][3] = pack_bf16(p2, p3); }
            }
            l0 = l0 * a0 + rs0; l1 = l1 * a1 + rs1;
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1; }
            // O += P V : k = keys (4 steps of 16), n = dims (8 tiles of 8)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) {
                    uint32_t b[4];
                    ldmatrix_x4_trans(b, sV + (size_t)(kc + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + j2 * 16 + (lane >> 4) * 8);
                    mma_bf16(o[j2 * 2], pa[kk], b[0], b[1]);
                    mma_bf16(o[j2 * 2 + 1], pa[kk], b[2], b[3]);
                }
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
        __nv_bfloat16 *ob = out + (size_t)img * L * d + h * HD;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 8 + 2 * tig;
            if (r0 < L) *reinterpret_cast<uint32_t *>(ob + (size_t)r0 * d + c) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
            if (r1 < L) *reinterpret_cast<uint32_t *>(ob + (size_t)r1 * d + c) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
        }
    }
}

}  // namespace

extern "C" int ec_attention(const void *qkv, void *out, int n_img, int L, int heads, void *stream)
{
    EC_REQUIRE(qkv && out && n_img > 0 && L > 0 && heads > 0, "ec_attention: bad arguments");
    EC_REQUIRE(L <= 1024, "ec_attention: L=%d exceeds the shared-memory K/V staging limit", L);
    EC_REQUIRE(n_img <= 65535, "ec_attention: n_img=%d exceeds grid.y", n_img);
    const int Lp = (L + 63) & ~63;
    const size_t smem = (size_t)2 * Lp * LDS * sizeof(__nv_bfloat16);
    EC_REQUIRE(smem <= 220 * 1024, "ec_attention: L=%d needs %zu bytes of shared memory", L, smem);
    static size_t attr = 0;
    if (smem > attr) {
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    attention_kernel<<<dim3(heads, n_img), WARPS * 32, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)qkv,
                                                                                     (__nv_bfloat16 *)out, L, heads);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}
