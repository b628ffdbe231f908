// attention_bwd.cu -- backward of softmax(q k^T / 8) v for the fine-tune step (the
// reference obtains it from autograd through nn.MultiheadAttention / models/lora.py:165-303).
//
// qkv bf16 [n_img*L, 3d], forward output o and its gradient d_o bf16 [n_img*L, d]  ->  dqkv bf16 [n_img*L, 3d].
// One CTA per (image, head); Q, K, V, dO of the whole sequence live in shared memory, nothing is atomically accumulated:
//   phase 0  D_i = <dO_i, O_i>
//   phase A  warp per 16-query tile: log-sum-exp of the row (recomputed), then dS = P * (dP - D) / 8 and dQ = dS K
//   phase B  warp per 16-key tile:  S^T and dP^T recomputed with the keys as rows, dV = P^T dO, dK = dS^T Q
// bf16 mma.sync m16n8k16 tiles; the training batch is small (config 5: 64 images), so this kernel favours simplicity.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int LDS = 72;       // smem row stride in elements (144 B: conflict-free ldmatrix)
constexpr int WARPS = 8;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void ldsm(uint32_t (&r)[4], const void *smem)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void ldsm_t(uint32_t (&r)[4], const void *smem)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// A fragments (16 rows x 64 dims) of the tile starting at row r0 of a [rows][LDS] smem matrix
__device__ __forceinline__ void load_a(uint32_t (&a)[4][4], const __nv_bfloat16 *s, int r0, int lane)
{
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldsm(a[kk], s + (size_t)(r0 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8);
}
// C[16 x 64] = A[16 x 64 dims] . M[c0.. c0+63][dims]^T   (rows of M are the output columns: K-major B operand)
__device__ __forceinline__ void mm_nt(float (&c)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16 *m, int c0, int lane)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
#pragma unroll
        for (int kk2 = 0; kk2 < 2; ++kk2) {
            uint32_t b[4];
            ldsm(b, m + (size_t)(c0 + j * 8 + (lane & 7)) * LDS + kk2 * 32 + (lane >> 3) * 8);
            mma(c[j], a[kk2 * 2], b[0], b[1]);
            mma(c[j], a[kk2 * 2 + 1], b[2], b[3]);
        }
    }
}
// acc[16 x 64 dims] += P[16 x 64 (k index)] . M[k0.. k0+63][dims]   (rows of M are the reduction index: ldmatrix.trans)
__device__ __forceinline__ void mm_nn_acc(float (&acc)[8][4], const uint32_t (&pa)[4][4], const __nv_bfloat16 *m, int k0, int lane)
{
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
            uint32_t b[4];
            ldsm_t(b, m + (size_t)(k0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + j2 * 16 + (lane >> 4) * 8);
            mma(acc[j2 * 2], pa[kk], b[0], b[1]);
            mma(acc[j2 * 2 + 1], pa[kk], b[2], b[3]);
        }
    }
}

__global__ void __launch_bounds__(WARPS * 32, 1)
attention_bwd_kernel(const __nv_bfloat16 *__restrict__ qkv, const __nv_bfloat16 *__restrict__ o,
                     const __nv_bfloat16 *__restrict__ d_o, __nv_bfloat16 *__restrict__ dqkv, int L, int heads)
{
    const int img = blockIdx.y, h = blockIdx.x;
    const int d = heads * HD, ld = 3 * d;
    const int Lp = (L + 63) & ~63;                  // rows staged: whole 64-row chunks (zero padded)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *sK = sQ + (size_t)Lp * LDS, *sV = sK + (size_t)Lp * LDS, *sdO = sV + (size_t)Lp * LDS;
    float *sD = reinterpret_cast<float *>(sdO + (size_t)Lp * LDS);    // D_i
    float *sLse = sD + Lp;                                             // log2-domain log-sum-exp of row i (+inf for padding)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    const __nv_bfloat16 *base = qkv + (size_t)img * L * ld + h * HD;
    const __nv_bfloat16 *obase = o + (size_t)img * L * d + h * HD, *dobase = d_o + (size_t)img * L * d + h * HD;
    __nv_bfloat16 *dbase = dqkv + (size_t)img * L * ld + h * HD;

    for (int i = tid; i < Lp * 8; i += WARPS * 32) {
        const int row = i >> 3, c = (i & 7) * 8;
        if (row < L) {
            cp_async16(sQ + row * LDS + c, base + (size_t)row * ld + c);
            cp_async16(sK + row * LDS + c, base + (size_t)row * ld + d + c);
            cp_async16(sV + row * LDS + c, base + (size_t)row * ld + 2 * d + c);
            cp_async16(sdO + row * LDS + c, dobase + (size_t)row * d + c);
        } else {
            const uint4 z = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(sQ + row * LDS + c) = z; *reinterpret_cast<uint4 *>(sK + row * LDS + c) = z;
            *reinterpret_cast<uint4 *>(sV + row * LDS + c) = z; *reinterpret_cast<uint4 *>(sdO + row * LDS + c) = z;
        }
    }
    asm volatile("cp.async.commit_group;");
    // phase 0: D_i = <dO_i, O_i> (fp32), one row per thread
    for (int r = tid; r < Lp; r += WARPS * 32) {
        float acc = 0.f;
        if (r < L) {
            const uint4 *po = reinterpret_cast<const uint4 *>(obase + (size_t)r * d), *pd = reinterpret_cast<const uint4 *>(dobase + (size_t)r * d);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 a = po[c], b = pd[c];
                const __nv_bfloat162 *ah = reinterpret_cast<const __nv_bfloat162 *>(&a), *bh = reinterpret_cast<const __nv_bfloat162 *>(&b);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(bh[e]);
                    acc += x.x * y.x + x.y * y.y;
                }
            }
        }
        sD[r] = acc;
        sLse[r] = INFINITY;
    }
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();

    const float sl2 = 0.125f * 1.4426950408889634f;
    const int n_t = (L + 15) >> 4;

    // ---- phase A: rows = queries ----
    for (int rt = warp; rt < n_t; rt += WARPS) {
        const int r0 = rt * 16 + g, r1 = r0 + 8;
        uint32_t qa[4][4], da[4][4];
        load_a(qa, sQ, rt * 16, lane);
        load_a(da, sdO, rt * 16, lane);
        // log-sum-exp of the two rows this thread holds
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        for (int kc = 0; kc < Lp; kc += 64) {
            float s[8][4];
            mm_nt(s, qa, sK, kc, lane);
            float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = kc + j * 8 + 2 * tig;
                if (key >= L) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
                if (key + 1 >= L) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
                c0 = fmaxf(c0, fmaxf(s[j][0], s[j][1])); c1 = fmaxf(c1, fmaxf(s[j][2], s[j][3]));
            }
            c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1)); c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
            c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
            const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a0 += exp2f((s[j][0] - n0) * sl2) + exp2f((s[j][1] - n0) * sl2);
                a1 += exp2f((s[j][2] - n1) * sl2) + exp2f((s[j][3] - n1) * sl2);
            }
            l0 = l0 * exp2f((m0 - n0) * sl2) + a0; l1 = l1 * exp2f((m1 - n1) * sl2) + a1;
            m0 = n0; m1 = n1;
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float lse0 = m0 * sl2 + log2f(l0), lse1 = m1 * sl2 + log2f(l1);
        if (tig == 0) { if (r0 < L) sLse[r0] = lse0; if (r1 < L) sLse[r1] = lse1; }
        const float D0 = sD[r0], D1 = sD[r1];
        float dq[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
        for (int kc = 0; kc < Lp; kc += 64) {
            float s[8][4], dp[8][4];
            mm_nt(s, qa, sK, kc, lane);
            mm_nt(dp, da, sV, kc, lane);
            uint32_t ds[4][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = kc + j * 8 + 2 * tig;
                const float p0 = key < L ? exp2f(s[j][0] * sl2 - lse0) : 0.f, p1 = key + 1 < L ? exp2f(s[j][1] * sl2 - lse0) : 0.f;
                const float p2 = key < L ? exp2f(s[j][2] * sl2 - lse1) : 0.f, p3 = key + 1 < L ? exp2f(s[j][3] * sl2 - lse1) : 0.f;
                const float e0 = p0 * (dp[j][0] - D0) * 0.125f, e1 = p1 * (dp[j][1] - D0) * 0.125f;
                const float e2 = p2 * (dp[j][2] - D1) * 0.125f, e3 = p3 * (dp[j][3] - D1) * 0.125f;
                if ((j & 1) == 0) { ds[j >> 1][0] = pack(e0, e1); ds[j >> 1][1] = pack(e2, e3); }
                else              { ds[j >> 1][2] = pack(e0, e1); ds[j >> 1][3] = pack(e2, e3); }
            }
            mm_nn_acc(dq, ds, sK, kc, lane);      // dQ += dS K
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 8 + 2 * tig;
            if (r0 < L) *reinterpret_cast<uint32_t *>(dbase + (size_t)r0 * ld + c) = pack(dq[j][0], dq[j][1]);
            if (r1 < L) *reinterpret_cast<uint32_t *>(dbase + (size_t)r1 * ld + c) = pack(dq[j][2], dq[j][3]);
        }
    }
    __syncthreads();      // every row's log-sum-exp is in shared memory

    // ---- phase B: rows = keys, columns = queries ----
    for (int kt = warp; kt < n_t; kt += WARPS) {
        const int r0 = kt * 16 + g, r1 = r0 + 8;
        uint32_t ka[4][4], va[4][4];
        load_a(ka, sK, kt * 16, lane);
        load_a(va, sV, kt * 16, lane);
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f; dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f; }
        for (int qc = 0; qc < Lp; qc += 64) {
            float st[8][4], dpt[8][4];
            mm_nt(st, ka, sQ, qc, lane);          // S^T  = K_j Q^T
            mm_nt(dpt, va, sdO, qc, lane);        // dP^T = V_j dO^T
            uint32_t pt[4][4], dst[4][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = qc + j * 8 + 2 * tig;
                const float lq0 = sLse[q], lq1 = sLse[q + 1], Dq0 = sD[q], Dq1 = sD[q + 1];   // +inf lse for padded queries -> p = 0
                const float p0 = exp2f(st[j][0] * sl2 - lq0), p1 = exp2f(st[j][1] * sl2 - lq1);
                const float p2 = exp2f(st[j][2] * sl2 - lq0), p3 = exp2f(st[j][3] * sl2 - lq1);
                const float e0 = p0 * (dpt[j][0] - Dq0) * 0.125f, e1 = p1 * (dpt[j][1] - Dq1) * 0.125f;
                const float e2 = p2 * (dpt[j][2] - Dq0) * 0.125f, e3 = p3 * (dpt[j][3] - Dq1) * 0.125f;
                if ((j & 1) == 0) { pt[j >> 1][0] = pack(p0, p1); pt[j >> 1][1] = pack(p2, p3); dst[j >> 1][0] = pack(e0, e1); dst[j >> 1][1] = pack(e2, e3); }
                else              { pt[j >> 1][2] = pack(p0, p1); pt[j >> 1][3] = pack(p2, p3); dst[j >> 1][2] = pack(e0, e1); dst[j >> 1][3] = pack(e2, e3); }
            }
            mm_nn_acc(dv, pt, sdO, qc, lane);     // dV += P^T dO
            mm_nn_acc(dk, dst, sQ, qc, lane);     // dK += dS^T Q
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 8 + 2 * tig;
            if (r0 < L) {
                *reinterpret_cast<uint32_t *>(dbase + (size_t)r0 * ld + d + c) = pack(dk[j][0], dk[j][1]);
                *reinterpret_cast<uint32_t *>(dbase + (size_t)r0 * ld + 2 * d + c) = pack(dv[j][0], dv[j][1]);
            }
            if (r1 < L) {
                *reinterpret_cast<uint32_t *>(dbase + (size_t)r1 * ld + d + c) = pack(dk[j][2], dk[j][3]);
                *reinterpret_cast<uint32_t *>(dbase + (size_t)r1 * ld + 2 * d + c) = pack(dv[j][2], dv[j][3]);
            }
        }
    }
}

}  // namespace

extern "C" int ec_attention_bwd(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int n_img, int L,
                                int heads, void *stream)
{
    EC_REQUIRE(qkv && o && d_o && dqkv && n_img > 0 && L > 0 && heads > 0, "ec_attention_bwd: bad arguments");
    static const bool force_mma = getenv("EC_ATTN_BWD") && !strcmp(getenv("EC_ATTN_BWD"), "mma");
    if (lse && L <= 256 && !force_mma) {       // tensor-memory kernel; it needs the forward's log-sum-exp
        const int rc = ec::attention_bwd_tc(qkv, o, d_o, lse, dqkv, n_img, L, heads, (cudaStream_t)stream);
        if (rc != EC_ERR_UNSUPPORTED) return rc;
    }
    EC_REQUIRE(n_img <= 65535, "ec_attention_bwd: n_img=%d exceeds grid.y", n_img);
    const int Lp = (L + 63) & ~63;
    const size_t smem = (size_t)4 * Lp * LDS * sizeof(__nv_bfloat16) + (size_t)2 * Lp * sizeof(float);
    EC_REQUIRE(smem <= 220 * 1024, "ec_attention_bwd: L=%d needs %zu bytes of shared memory", L, smem);
    static size_t attr[64] = {0};
    int dev_id = 0;
    EC_CUDA_CHECK(cudaGetDevice(&dev_id));
    if (dev_id < 64 && smem > attr[dev_id]) {
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[dev_id] = smem;
    }
    attention_bwd_kernel<<<dim3(heads, n_img), WARPS * 32, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)qkv, (const __nv_bfloat16 *)o, (const __nv_bfloat16 *)d_o, (__nv_bfloat16 *)dqkv, L, heads);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}
