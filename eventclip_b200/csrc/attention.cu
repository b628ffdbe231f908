// attention.cu -- softmax(q k^T / sqrt(64)) v for the CLIP ViT blocks (nn.MultiheadAttention inside openai-CLIP's
// ResidualAttentionBlock [3P]; LoRA variant models/lora.py:165-303 computes the same core with merged weights).
//
// Input is the packed in_proj output qkv bf16 [n_img*L, 3d] (q | k | v; head h at columns h*64..h*64+63), output
// bf16 [n_img*L, d].  L is 50 / 197 / 257 tokens, head_dim is 64 for every CLIP ViT.
// One CTA per (image, head): Q, K and V of the whole sequence staged once in shared memory (cp.async), each of the
// 7 warps owns 16-query row tiles and runs an online-softmax loop over 64-key chunks on bf16 mma.sync tiles.
// (4 % of the encoder FLOPs; the tcgen05 budget goes to the GEMMs -- see DESIGN.md.)
#include <cstdlib>
#include <string>

#include "common.cuh"

namespace {

constexpr int HD = 64;        // head dim
constexpr int LDS = 72;       // smem row stride in elements (144 B: conflict-free ldmatrix)
constexpr int WARPS = 7;      // 13 row tiles of 16 queries (L = 197) in two rounds

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
__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// One 64-key chunk of one 16-query tile.  NT = number of 8-key n-tiles that hold real keys (8, or fewer in the tail);
// MASK = the chunk contains keys >= L.
template <int NT, bool MASK>
__device__ __forceinline__ void chunk(const __nv_bfloat16 *sK, const __nv_bfloat16 *sV, int kc, int L, int lane, int tig,
                                      const uint32_t (&qa)[4][4], float (&o)[8][4], float &m0, float &m1, float &l0,
                                      float &l1, float sl2)
{
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int kk2 = 0; kk2 < 2; ++kk2) {
            uint32_t b[4];
            ldmatrix_x4(b, sK + (size_t)(kc + j * 8 + (lane & 7)) * LDS + kk2 * 32 + (lane >> 3) * 8);
            mma_bf16(s[j], qa[kk2 * 2], b[0], b[1]);
            mma_bf16(s[j], qa[kk2 * 2 + 1], b[2], b[3]);
        }
    }
    float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        if (MASK) {
            const int key = kc + j * 8 + 2 * tig;
            if (key >= L) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (key + 1 >= L) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
        }
        cm0 = fmaxf(cm0, fmaxf(s[j][0], s[j][1]));
        cm1 = fmaxf(cm1, fmaxf(s[j][2], s[j][3]));
    }
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
    const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);   // finite: every chunk holds >= 1 real key
    const float a0 = fast_exp2((m0 - nm0) * sl2), a1 = fast_exp2((m1 - nm1) * sl2);
    m0 = nm0; m1 = nm1;
    const float ms0 = nm0 * sl2, ms1 = nm1 * sl2;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pa[NT / 2][4];   // P as A fragments: NT/2 k-steps of 16 keys
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const float p0 = fast_exp2(fmaf(s[j][0], sl2, -ms0)), p1 = fast_exp2(fmaf(s[j][1], sl2, -ms0));
        const float p2 = fast_exp2(fmaf(s[j][2], sl2, -ms1)), p3 = fast_exp2(fmaf(s[j][3], sl2, -ms1));
        rs0 += p0 + p1; rs1 += p2 + p3;
        if ((j & 1) == 0) { pa[j >> 1][0] = pack_bf16(p0, p1); pa[j >> 1][1] = pack_bf16(p2, p3); }
        else              { pa[j >> 1][2] = pack_bf16(p0, p1); pa[j >> 1][3] = pack_bf16(p2, p3); }
    }
    l0 = l0 * a0 + rs0; l1 = l1 * a1 + rs1;
    if (__any_sync(0xffffffffu, a0 != 1.f || a1 != 1.f)) {   // the running max moved: rescale the accumulator
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1; }
    }
    // O += P V : k = keys (NT/2 steps of 16), n = dims (8 tiles of 8)
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, sV + (size_t)(kc + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + j2 * 16 + (lane >> 4) * 8);
            mma_bf16(o[j2 * 2], pa[kk], b[0], b[1]);
            mma_bf16(o[j2 * 2 + 1], pa[kk], b[2], b[3]);
        }
    }
}

__global__ void __launch_bounds__(WARPS * 32, 2) attention_kernel(const __nv_bfloat16 *__restrict__ qkv,
                                                               __nv_bfloat16 *__restrict__ out, int L, int heads)
{
    const int img = blockIdx.y, h = blockIdx.x;
    const int d = heads * HD;
    const int ld = 3 * d;
    const int Lp = (L + 15) & ~15;                 // rows staged: whole 16-row / 16-key steps
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *sK = sQ + (size_t)Lp * LDS;
    __nv_bfloat16 *sV = sK + (size_t)Lp * LDS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const __nv_bfloat16 *base = qkv + (size_t)img * L * ld + h * HD;

    // stage Q, K and V (rows >= L are zero so that 0-probability keys contribute exact zeros)
    for (int i = tid; i < Lp * 8; i += WARPS * 32) {
        const int row = i >> 3, c = (i & 7) * 8;
        if (row < L) {
            cp_async16(sQ + row * LDS + c, base + (size_t)row * ld + c);
            cp_async16(sK + row * LDS + c, base + (size_t)row * ld + d + c);
            cp_async16(sV + row * LDS + c, base + (size_t)row * ld + 2 * d + c);
        } else {
            *reinterpret_cast<uint4 *>(sQ + row * LDS + c) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(sK + row * LDS + c) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(sV + row * LDS + c) = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();

    const float sl2 = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
    const int n_rt = Lp >> 4;
    const int full_end = L & ~63;                     // keys [0, full_end) form whole 64-key chunks
    const int tail_nt = ((L - full_end + 15) >> 4) << 1;   // n-tiles of the tail chunk: 0, 2, 4, 6, 8
    for (int rt = warp; rt < n_rt; rt += WARPS) {
        const int r0 = rt * 16 + g, r1 = r0 + 8;
        uint32_t qa[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            ldmatrix_x4(qa[kk], sQ + (size_t)(rt * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8);
        float o[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

        for (int kc = 0; kc < full_end; kc += 64) chunk<8, false>(sK, sV, kc, L, lane, tig, qa, o, m0, m1, l0, l1, sl2);
        if (tail_nt == 2) chunk<2, true>(sK, sV, full_end, L, lane, tig, qa, o, m0, m1, l0, l1, sl2);
        else if (tail_nt == 4) chunk<4, true>(sK, sV, full_end, L, lane, tig, qa, o, m0, m1, l0, l1, sl2);
        else if (tail_nt == 6) chunk<6, true>(sK, sV, full_end, L, lane, tig, qa, o, m0, m1, l0, l1, sl2);
        else if (tail_nt == 8) chunk<8, true>(sK, sV, full_end, L, lane, tig, qa, o, m0, m1, l0, l1, sl2);

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
    return ec_attention_ex(qkv, out, n_img, L, heads, 0, stream);
}

extern "C" int ec_attention_fwd_lse(const void *qkv, void *out, float *lse, int n_seq, int L, int heads, int causal, void *stream)
{
    EC_REQUIRE(qkv && out && lse && n_seq > 0 && L > 0 && heads > 0, "ec_attention_fwd_lse: bad arguments");
    const int rc = ec::attention_tc(qkv, out, n_seq, L, heads, causal, (cudaStream_t)stream, lse);
    if (rc == EC_ERR_UNSUPPORTED) ec::set_error("ec_attention_fwd_lse: L=%d is outside the tensor-memory kernels' range (<= 384)", L);
    return rc;
}

extern "C" int ec_attention_ex(const void *qkv, void *out, int n_img, int L, int heads, int causal, void *stream)
{
    EC_REQUIRE(qkv && out && n_img > 0 && L > 0 && heads > 0, "ec_attention: bad arguments");
    EC_REQUIRE(L <= 1024, "ec_attention: L=%d exceeds the shared-memory K/V staging limit", L);
    EC_REQUIRE(n_img <= 65535, "ec_attention: n_img=%d exceeds grid.y", n_img);
    // tensor-memory kernels for L <= 384 (every CLIP ViT at 224 px); EC_ATTN=mma forces the mma.sync kernel below
    static const bool force_mma = getenv("EC_ATTN") && std::string(getenv("EC_ATTN")) == "mma";
    if (!force_mma) {
        const int rc = ec::attention_tc(qkv, out, n_img, L, heads, causal, (cudaStream_t)stream);
        if (rc != EC_ERR_UNSUPPORTED) return rc;
    }
    EC_REQUIRE(!causal, "ec_attention_ex: the causal mask and fp16 operands are only built into the tensor-memory kernels (L <= 384)");
    const int Lp = (L + 15) & ~15;
    const size_t smem = (size_t)3 * Lp * LDS * sizeof(__nv_bfloat16);
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
