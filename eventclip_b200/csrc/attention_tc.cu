// attention_tc.cu -- softmax(q k^T / sqrt(64)) v on the 5th-generation tensor cores (tcgen05 + TMEM), L <= 384.
//
// Same contract as attention.cu (nn.MultiheadAttention core of openai-CLIP's ResidualAttentionBlock [3P]; LoRA variant
// models/lora.py:165-303): qkv bf16 [n_img*L, 3d] -> out bf16 [n_img*L, d], head_dim 64.
//
// Persistent CTAs (one per SM) walk over (image, head) units; shared memory holds two units (next one prefetched),
// tensor memory both 128-query tiles of the current unit (2 x 256 columns):
//   warp 0      issues everything asynchronous: six TMA box loads (Q, K, V; rows >= L are zero-filled by the 3-D tensor
//               map), then per 128-query tile  S = Q K^T  (tcgen05.mma, A and B from shared memory, N = ceil16(L))
//               and  O = P V  (A = P read from TENSOR MEMORY, B = V as an MN-major shared-memory operand)
//   warps 1-8   one thread per query row (one warpgroup per 128-query tile): a single pass over S in TMEM -- exponentials
//               relative to the maximum of the row's first 32 scores, P written back as packed bf16 over the first
//               half of S's own columns (tcgen05.st); the softmax denominators come from a companion MMA (P . ones);
//               O / denominator -> global.
// S never leaves the SM and P never touches shared memory.
//
// Kernels of this file (dispatch in ec::attention_tc): attention_tc4_kernel (L <= 64: four chains in flight), attention_tc2_kernel
// (64 < L <= 256: two slots, two threads per query row), attention_tc2x_kernel (L = 257: the same with the class key and the last
// query row folded in), attention_tc_big_kernel (other 256 < L <= 384), attention_tc_kernel (round 1, EC_ATTN_V=1), and the
// tensor-memory backward attention_bwd_tc_kernel.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int TILE_BYTES = 128 * 128;           // 128 rows x 64 bf16, SWIZZLE_128B
constexpr int NTHREADS = 288;                   // warp 0: TMA + MMA issue; warps 1-4 / 5-8: one thread per query row of tile 0 / 1
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TILE_COLS = 256;             // tensor-memory columns per 128-query tile
constexpr int STAGE_BYTES = 6 * TILE_BYTES;     // Q, K, V of one (image, head): 2 tiles each
constexpr uint32_t O_COL = 128;                 // O accumulator columns [128, 192)
constexpr uint32_t SUM_COL = 192;               // row sums = P . ones, columns [192, 208)
constexpr int ONES_BYTES = 2048;                // 16 rows x 64 bf16 of 1.0: B operand of the row-sum MMA (layout-proof)

struct AttnParams {
    __nv_bfloat16 *out;
    const __nv_bfloat16 *qkv;      // the packed input (attention_tc2x_kernel reads key / value row 0 and query row 256 directly)
    float *lse;           // optional [n_img, heads, L]: log2-domain log-sum-exp of every row (kept for the backward pass)
    int L, heads, d, n_img, causal;
    int f16;              // q, k, v, P and the output are fp16 instead of bf16 (the inference forward's fp16-operand mode)
    long long *dbg;       // EC_ATTN_DBG: clock64 stamps of CTA 0 (profiling only)
    int exact;            // row reference = the exact row maximum (one extra pass over S) instead of the maximum of the first 32 scores
    int stagger;          // two-tile units: issue the tiles' MMAs half a period apart (EC_ATTN_STAGGER=0 restores side by side)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// two fp32 values -> one packed 16-bit pair (low half = a), bf16 or fp16, round to nearest even
__device__ __forceinline__ uint32_t pack16x2(float a, float b, int f16)
{
    if (f16) {
        const __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<const uint32_t *>(&h);
    }
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    uint64_t t0 = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && (spins & 0x3ff) == 0x3ff) {      // a protocol bug traps after ~2 s instead of hanging the GPU
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t.reg .b32 rx;\n\t"
        "elect.sync rx|P1, 0xffffffff;\n\t"
        "@P1 mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}

// shared-memory operand descriptors, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tmem_ld1(uint32_t taddr)
{
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(r);
}
// two probabilities per SFU op: exp2 of a packed bf16 pair, result already in the packed bf16 form P is stored in
__device__ __forceinline__ uint32_t exp2_bf16x2(float lo, float hi)
{
    uint32_t x, y;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(hi), "f"(lo));
    asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32 pairs (Blackwell FFMA2 / FADD2: two lanes of fp32 math per issued instruction)
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk2(float a, float b)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ f2 mk2u(uint32_t a, uint32_t b)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void un2(f2 x, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
// Softmax of one 128-query tile (thread = query row) followed by the O epilogue.
// causal != 0: key j contributes to query `row` only if j <= row (text tower).
template <uint32_t OCOL, uint32_t SUMCOL, int NF>      // NF: tensor-memory loads in flight in the exact-maximum pass
__device__ __forceinline__ void softmax_tile(uint32_t lane_base, int L, int nch, bool live, int row, int lane, uint64_t *bar_p,
                                             uint64_t *bar_o, uint32_t o_parity, uint64_t *bar_oe, __nv_bfloat16 *orow,
                                             int causal, float *lse_row, int f16, int exact, long long *dbg = nullptr)
{
    if (dbg && lane == 0) dbg[0] = clock64();                  // S landed
    float ms_keep = 0.f;
    const int klim = causal ? min(L, row + 1) : L;     // keys [0, klim) are visible to this row
    const float sl2 = 0.125f * 1.4426950408889634f;
                // Single pass over S (tensor-memory reads are the scarce resource: ~64 B/clk per SM).  Softmax is invariant to
                // the constant subtracted before the exponential, so the reference point is the maximum of the FIRST 32 scores
                // of the row instead of the exact row maximum; exponents are evaluated in fp32 (range 2^+-126, clamped at
                // +120) and the denominators come from the tensor core (P . ones), so the result is the same softmax.
                // P chunk c lands on columns [16c, 16c+16), which this thread has already read (chunks are taken in order).
                if (live) {
                    uint32_t va[32], vb[32];
                    tmem_ld32_issue(lane_base, va);
                    tmem_ld_wait();
                    if (nch > 1) tmem_ld32_issue(lane_base + 32u, vb);
                    float m = -INFINITY;
    #pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < klim) m = fmaxf(m, __uint_as_float(va[j]));
                    if (exact) {                       // the other chunks too, NF loads in flight (vb is one more: the waits cover it)
                        for (int c = 1; c < nch; c += NF) {
                            uint32_t vc[NF][32];
    #pragma unroll
                            for (int u = 0; u < NF; ++u)
                                if (c + u < nch) tmem_ld32_issue(lane_base + (uint32_t)((c + u) * 32), vc[u]);
                            tmem_ld_wait();
    #pragma unroll
                            for (int u = 0; u < NF; ++u)
                                if (c + u < nch) {
                                    if ((c + u + 1) * 32 <= klim) {
                                        float m0 = __uint_as_float(vc[u][0]), m1 = __uint_as_float(vc[u][1]);
    #pragma unroll
                                        for (int j = 2; j < 32; j += 2) {
                                            m0 = fmaxf(m0, __uint_as_float(vc[u][j]));
                                            m1 = fmaxf(m1, __uint_as_float(vc[u][j + 1]));
                                        }
                                        m = fmaxf(m, fmaxf(m0, m1));
                                    } else {
    #pragma unroll
                                        for (int j = 0; j < 32; ++j)
                                            if ((c + u) * 32 + j < klim) m = fmaxf(m, __uint_as_float(vc[u][j]));
                                    }
                                }
                        }
                    }
                    const float ms = m * sl2;
                    ms_keep = ms;
                    // chunk c of the row: 32 scores -> 16 packed probabilities.  `full` chunks (every key visible) take the
                    // straight-line path without the visibility selects.
                    const f2 sl2x = mk2(sl2, sl2), msx = mk2(-ms, -ms);
                    auto emit = [&](const uint32_t (&v)[32], int c) {
                        uint32_t pk[16];
                        if ((c + 1) * 32 <= klim) {
    #pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                float x0, x1;
                                un2(fma2(mk2u(v[j], v[j + 1]), sl2x, msx), x0, x1);
                                x0 = fminf(x0, 120.f);
                                x1 = fminf(x1, 120.f);
                                pk[j >> 1] = pack16x2(fast_exp2(x0), fast_exp2(x1), f16);
                            }
                        } else {
    #pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                float x0 = fminf(fmaf(__uint_as_float(v[j]), sl2, -ms), 120.f);
                                float x1 = fminf(fmaf(__uint_as_float(v[j + 1]), sl2, -ms), 120.f);
                                if (c * 32 + j >= klim) x0 = -INFINITY;
                                if (c * 32 + j + 1 >= klim) x1 = -INFINITY;
                                pk[j >> 1] = pack16x2(fast_exp2(x0), fast_exp2(x1), f16);
                            }
                        }
                        tmem_st16(lane_base + (uint32_t)(c * 16), pk);
                    };
                    for (int c = 0; c < nch; c += 2) {
                        if (c > 0) {
                            tmem_ld_wait();
                            if (c + 1 < nch) tmem_ld32_issue(lane_base + (uint32_t)((c + 1) * 32), vb);
                        }
                        emit(va, c);
                        if (c + 1 < nch) {
                            tmem_ld_wait();
                            if (c + 2 < nch) tmem_ld32_issue(lane_base + (uint32_t)((c + 2) * 32), va);
                            emit(vb, c + 1);
                        }
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p);
                if (dbg && lane == 0) dbg[1] = clock64();      // P written
                // epilogue: O / rowsum -> bf16 -> global
                mbar_wait(bar_o, o_parity);
                tc_fence_after();
                if (dbg && lane == 0) dbg[2] = clock64();      // O landed
                if (!live) {      // nothing to store: keep the barrier protocol and move on
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_oe);
                    return;
                }
                const float rsum = tmem_ld1(lane_base + SUMCOL);
                const float inv = 1.f / rsum;
                if (lse_row && row < L) *lse_row = ms_keep + log2f(rsum);     // p_ij = exp2(s_ij * sl2 - lse)
    #pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + OCOL + (uint32_t)(c * 32), v);
                    if (row < L) {
    #pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 o;
                            o.x = pack16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv, f16);
                            o.y = pack16x2(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv, f16);
                            o.z = pack16x2(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv, f16);
                            o.w = pack16x2(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv, f16);
                            *reinterpret_cast<uint4 *>(orow + c * 32 + j) = o;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_oe);
                if (dbg && lane == 0) dbg[3] = clock64();      // epilogue done
}

// Persistent: one CTA per SM walks over (image, head) units.  Shared memory holds two units (the next one is fetched
// while the current one is computed); tensor memory holds both 128-query tiles of a unit, each served by its own
// softmax warpgroup, so the two tiles of a unit run side by side.
__global__ void __launch_bounds__(NTHREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnParams p)
{
    const bool stagger = p.stagger != 0;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    // stage s: Q (2 tiles) | K (2 tiles, K-major B of S = Q K^T) | V (2 tiles, MN-major B of O = P V)
    unsigned char *sOnes = smem + 2 * STAGE_BYTES;    // all-ones operand: O's companion MMA yields the softmax denominators
    __shared__ __align__(8) uint64_t bar_qk[2], bar_v[2], bar_free[2];       // per smem stage
    __shared__ __align__(8) uint64_t bar_s[2], bar_p[2], bar_o[2], bar_oe[2];   // per query tile / warpgroup
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.L, d = p.d, heads = p.heads;
    const int KP = (L + 15) & ~15;                    // keys rounded to a whole MMA K step
    const int MT = (L + 127) >> 7;                    // 128-query tiles (1 or 2)
    const int n_units = p.n_img * heads;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_qk[i], 1); mbar_init(&bar_v[i], 1); mbar_init(&bar_free[i], 1);
            mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 4); mbar_init(&bar_o[i], 1); mbar_init(&bar_oe[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS) reinterpret_cast<uint32_t *>(sOnes)[i] = p.f16 ? 0x3c003c00u : 0x3f803f80u;   // 1.0 x2 (fp16 / bf16)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== TMA + MMA issuer =====================
        auto load_unit = [&](int unit, int stage) {      // elected lane only
            const int img = unit / heads, h = unit % heads;
            unsigned char *sQ = smem + stage * STAGE_BYTES, *sK = sQ + 2 * TILE_BYTES, *sV = sQ + 4 * TILE_BYTES;
            // Q and K gate the first MMA; V is only needed once P exists, so it gets its own barrier
            mbar_expect_tx(&bar_qk[stage], (uint32_t)(2 * MT * TILE_BYTES));
            for (int b = 0; b < MT; ++b) {
                tma_load_3d(sK + b * TILE_BYTES, &map_qkv, &bar_qk[stage], d + h * HD, b * 128, img);
                tma_load_3d(sQ + b * TILE_BYTES, &map_qkv, &bar_qk[stage], h * HD, b * 128, img);
            }
            mbar_expect_tx(&bar_v[stage], (uint32_t)(MT * TILE_BYTES));
            for (int b = 0; b < MT; ++b) tma_load_3d(sV + b * TILE_BYTES, &map_qkv, &bar_v[stage], 2 * d + h * HD, b * 128, img);
        };
        // S: D fp32, A/B bf16 (bits 7, 10) or fp16 K-major, M = 128, N = KP.   O: B is MN-major (bit 16), N = 64.   sums: N = 16.
        const uint32_t FMT16 = p.f16 ? 0u : ((1u << 7) | (1u << 10));
        const uint32_t idesc_s = (1u << 4) | FMT16 | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_o = (1u << 4) | FMT16 | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_1 = (1u << 4) | FMT16 | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t odesc = make_desc(smem_u32(sOnes));
        if ((int)blockIdx.x < n_units && elect_one()) load_unit(blockIdx.x, 0);
        __syncwarp();
        int dbg_i = 0, dbg_pv = 0;      // unit index of the S issue / (unit * 2 + tile) of the PV issue, for the stamps
        auto issue_s = [&](int stage, int t) {
            unsigned char *sQ = smem + stage * STAGE_BYTES, *sK = sQ + 2 * TILE_BYTES;
            const uint64_t kdesc = make_desc(smem_u32(sK)), qdesc = make_desc(smem_u32(sQ + t * TILE_BYTES));
            const uint32_t tb = tmem_base + (uint32_t)(t * TILE_COLS);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0);
                umma_commit(&bar_s[t]);
                if (p.dbg && blockIdx.x == 0 && dbg_i < 16) p.dbg[(dbg_i * 2 + t) * 2] = clock64();
            }
            __syncwarp();
        };
        auto issue_pv = [&](int stage, int t, bool release_stage) {
            unsigned char *sV = smem + stage * STAGE_BYTES + 4 * TILE_BYTES;
            const uint64_t vdesc = make_desc(smem_u32(sV));
            const uint32_t tb = tmem_base + (uint32_t)(t * TILE_COLS);
            if (elect_one()) {
                for (int j = 0; j < KP / 16; ++j) {   // 16 keys per step: 8 packed columns of P, 16 rows (2048 B) of V
                    umma_ts(tb + O_COL, tb + (uint32_t)(8 * j), vdesc + (uint64_t)(128 * j), idesc_o, j != 0);
                    umma_ts(tb + SUM_COL, tb + (uint32_t)(8 * j), odesc, idesc_1, j != 0);   // += P . 1
                }
                umma_commit(&bar_o[t]);
                if (release_stage) umma_commit(&bar_free[stage]);   // every MMA that reads this stage has retired
                if (p.dbg && blockIdx.x == 0 && dbg_pv < 32) { p.dbg[(dbg_pv >> 1) * 4 + (dbg_pv & 1) * 2 + 1] = clock64(); }
            }
            __syncwarp();
        };
        if (MT == 2 && stagger) {
            // Two 128-query tiles = two dependent chains  S -> softmax -> P V -> epilogue -> S (next unit).  Issued side by side
            // they run in phase: both softmax warpgroups compete for the SFU and the tensor-memory read port at the same time and
            // both wait at the same time.  Issue order  S0(i), PV1(i-1), S1(i), PV0(i)  keeps the chains half a period apart: the
            // softmax of one tile runs while the other tile is in its MMA / epilogue / barrier phases.
            int i = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
                const int stage = i & 1;
                const uint32_t sph = (uint32_t)(i >> 1) & 1, uph = (uint32_t)i & 1;
                const int next = unit + gridDim.x;
                mbar_wait(&bar_qk[stage], sph);
                if (i >= 1) mbar_wait(&bar_oe[0], uph ^ 1);      // tile 0's tensor memory of unit i-1 fully consumed
                tc_fence_after();
                dbg_i = i;
                issue_s(stage, 0);
                if (i >= 1) {
                    mbar_wait(&bar_p[1], uph ^ 1);               // P of tile 1, unit i-1 (its V was awaited for tile 0)
                    tc_fence_after();
                    dbg_pv = (i - 1) * 2 + 1;
                    issue_pv(stage ^ 1, 1, true);
                }
                if (next < n_units) {
                    if (i >= 1) mbar_wait(&bar_free[stage ^ 1], (uint32_t)((i - 1) >> 1) & 1);
                    if (elect_one()) load_unit(next, stage ^ 1);
                    __syncwarp();
                }
                if (i >= 1) { mbar_wait(&bar_oe[1], uph ^ 1); tc_fence_after(); }
                issue_s(stage, 1);
                mbar_wait(&bar_v[stage], sph);
                mbar_wait(&bar_p[0], uph);
                tc_fence_after();
                dbg_pv = i * 2;
                issue_pv(stage, 0, false);
            }
            if (i >= 1) {                                        // tile 1 of the last unit
                mbar_wait(&bar_p[1], (uint32_t)(i - 1) & 1);
                tc_fence_after();
                dbg_pv = (i - 1) * 2 + 1;
                issue_pv((i - 1) & 1, 1, true);
            }
        } else {
            int i = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
                const int stage = i & 1;
                const uint32_t sph = (uint32_t)(i >> 1) & 1, uph = (uint32_t)i & 1;
                const int next = unit + gridDim.x;
                if (next < n_units) {
                    // the other stage was last read by the MMAs of unit i-1
                    if (i >= 1) mbar_wait(&bar_free[stage ^ 1], (uint32_t)((i - 1) >> 1) & 1);
                    if (elect_one()) load_unit(next, stage ^ 1);
                    __syncwarp();
                }
                mbar_wait(&bar_qk[stage], sph);
                tc_fence_after();
                for (int t = 0; t < MT; ++t) {
                    if (i >= 1) { mbar_wait(&bar_oe[t], uph ^ 1); tc_fence_after(); }   // tile t's TMEM of unit i-1 fully consumed
                    dbg_i = i;
                    issue_s(stage, t);
                }
                mbar_wait(&bar_v[stage], sph);
                for (int t = 0; t < MT; ++t) {
                    mbar_wait(&bar_p[t], uph);               // P is in tensor memory
                    tc_fence_after();
                    dbg_pv = i * 2 + t;
                    issue_pv(stage, t, t == MT - 1);
                }
            }
        }
    } else if ((warp - 1) / 4 < MT) {
        // ===================== softmax + epilogue: thread = query row, one warpgroup per 128-query tile =====================
        const int t = (warp - 1) >> 2;                // tile served by this warpgroup
        const int quarter = warp & 3;                 // TMEM lane quarter of this warp (hardware: lanes 32*(warp%4)..+31)
        const uint32_t lane_base = tmem_base + (uint32_t)(t * TILE_COLS) + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f;
        const int nch = (KP + 31) >> 5;               // 32-column chunks of S
        const int row = t * 128 + quarter * 32 + lane;
        const bool live = t * 128 + quarter * 32 < L; // warp-uniform: this warp owns at least one real query row
        int i = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
            const uint32_t uph = (uint32_t)i & 1;
            const int img = unit / heads, h = unit % heads;
            mbar_wait(&bar_s[t], uph);
            tc_fence_after();
            softmax_tile<O_COL, SUM_COL, 1>(lane_base, L, nch, live, row, lane, &bar_p[t], &bar_o[t], uph, &bar_oe[t],
                                         p.out + ((size_t)img * L + row) * d + h * HD, p.causal,
                                         p.lse ? p.lse + (size_t)unit * L + row : nullptr, p.f16, p.exact,
                                         (p.dbg && blockIdx.x == 0 && quarter == 0 && i < 16) ? p.dbg + 64 + (i * 2 + t) * 4 : nullptr);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ================================================================================================================
// v2 (L <= 256): two threads per query row and two independent chains in flight.
//
// What the round-2 timeline (clock64 stamps per phase, profiles/r02_attention_timeline.txt) showed about v1: a tcgen05.ld
// of 32 columns holds its warp for ~256 cycles (tensor memory -> registers moves 16 B/clk per SM sub-partition) and does not
// overlap that warp's own math, so one thread per row spends 7 x (256 load + ~360 math/SFU) = 4300 cycles in the softmax of
// a 197-key row and 2000 more in the O epilogue, with the tensor pipe idle meanwhile.  Here
//   * every 128-row tile is served by EIGHT warps: the two warps that share a tensor-memory lane quarter (same sub-partition)
//     take the even / odd 32-column chunks of S, so one warp's load overlaps the other's exponentials; the row reference
//     (maximum of the first 32 scores) goes from the even warp to the odd one through shared memory and a 64-thread named
//     barrier, which also orders the P stores (P chunk c overwrites S columns of chunk c/2) against the partner's loads;
//   * the two tensor-memory slots are two chains  S -> softmax -> P V -> epilogue  issued half a period apart
//     (S(j), PV(j-1), S(j+1), ...): for L <= 128 the slots hold DIFFERENT (image, head) units, four of which are resident
//     in shared memory, so ViT-B/32 (L = 50) and the text tower (L = 77) no longer leave half the CTA idle.
// ================================================================================================================
constexpr int NTHREADS2 = 544;                  // warp 0: TMA + MMA issue; warps 1-8 / 9-16: slot 0 / 1 (x even / odd chunks)

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int MT, int F16>      // MT: 128-query tiles per unit; F16: fp16 (else bf16) operands and output
__global__ void __launch_bounds__(NTHREADS2, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *sOnes = smem + 2 * STAGE_BYTES;    // all-ones operand: O's companion MMA yields the softmax denominators
    __shared__ __align__(8) uint64_t bar_qk[4], bar_v[4], bar_free[4];          // per smem stage
    __shared__ __align__(8) uint64_t bar_s[2], bar_p[2], bar_o[2], bar_oe[2];   // per tensor-memory slot
    __shared__ float s_ref[2][128];                                              // row reference, even -> odd warp of a pair
    __shared__ float s_ref1[2][128];                                             // exact mode: the odd warp's partial row maximum
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.L, d = p.d, heads = p.heads;
    const int KP = (L + 15) & ~15;                    // keys rounded to a whole MMA K step
    const int NST = MT == 2 ? 2 : 4;                  // units resident in shared memory
    const int stage_bytes = 3 * MT * TILE_BYTES;
    const int n_units = p.n_img * heads;
    const int n_my = (int)blockIdx.x < n_units ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_items = n_my * MT;                    // item j = (unit j / MT, tile j % MT) runs in slot j & 1

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_qk[i], 1); mbar_init(&bar_v[i], 1); mbar_init(&bar_free[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 8); mbar_init(&bar_o[i], 1); mbar_init(&bar_oe[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS2) reinterpret_cast<uint32_t *>(sOnes)[i] = F16 ? 0x3c003c00u : 0x3f803f80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== TMA + MMA issuer =====================
        auto load_unit = [&](int u) {                 // elected lane only; u = this CTA's u-th unit
            const int unit = (int)blockIdx.x + u * (int)gridDim.x, stage = u % NST;
            const int img = unit / heads, h = unit % heads;
            unsigned char *sQ = smem + stage * stage_bytes, *sK = sQ + MT * TILE_BYTES, *sV = sQ + 2 * MT * TILE_BYTES;
            mbar_expect_tx(&bar_qk[stage], (uint32_t)(2 * MT * TILE_BYTES));
            for (int b = 0; b < MT; ++b) {
                tma_load_3d(sK + b * TILE_BYTES, &map_qkv, &bar_qk[stage], d + h * HD, b * 128, img);
                tma_load_3d(sQ + b * TILE_BYTES, &map_qkv, &bar_qk[stage], h * HD, b * 128, img);
            }
            mbar_expect_tx(&bar_v[stage], (uint32_t)(MT * TILE_BYTES));
            for (int b = 0; b < MT; ++b) tma_load_3d(sV + b * TILE_BYTES, &map_qkv, &bar_v[stage], 2 * d + h * HD, b * 128, img);
        };
        const uint32_t FMT16 = F16 ? 0u : ((1u << 7) | (1u << 10));
        const uint32_t idesc_s = (1u << 4) | FMT16 | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_o = (1u << 4) | FMT16 | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_1 = (1u << 4) | FMT16 | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t odesc = make_desc(smem_u32(sOnes));
        if (elect_one())
            for (int u = 0; u < NST && u < n_my; ++u) load_unit(u);
        __syncwarp();
        auto issue_s = [&](int j) {
            const int u = j / MT, t = j - u * MT, stage = u % NST, slot = j & 1;
            mbar_wait(&bar_qk[stage], (uint32_t)(u / NST) & 1);
            if (j >= 2) mbar_wait(&bar_oe[slot], (uint32_t)((j - 2) >> 1) & 1);      // the slot's previous item is fully consumed
            tc_fence_after();
            unsigned char *sQ = smem + stage * stage_bytes, *sK = sQ + MT * TILE_BYTES;
            const uint64_t kdesc = make_desc(smem_u32(sK)), qdesc = make_desc(smem_u32(sQ + t * TILE_BYTES));
            const uint32_t tb = tmem_base + (uint32_t)(slot * TILE_COLS);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0);
                umma_commit(&bar_s[slot]);
                if (p.dbg && blockIdx.x == 0 && j < 32) p.dbg[j * 2] = clock64();
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int u = j / MT, t = j - u * MT, stage = u % NST, slot = j & 1;
            mbar_wait(&bar_v[stage], (uint32_t)(u / NST) & 1);
            mbar_wait(&bar_p[slot], (uint32_t)(j >> 1) & 1);                         // P is in tensor memory
            tc_fence_after();
            unsigned char *sV = smem + stage * stage_bytes + 2 * MT * TILE_BYTES;
            const uint64_t vdesc = make_desc(smem_u32(sV));
            const uint32_t tb = tmem_base + (uint32_t)(slot * TILE_COLS);
            const bool last = t == MT - 1;
            const int nk = KP / 16;
            if (elect_one()) {
                for (int k = 0; k < nk; ++k) {        // 16 keys per step: 8 packed columns of P, 16 rows (2048 B) of V
                    umma_ts(tb + O_COL, tb + (uint32_t)(8 * k), vdesc + (uint64_t)(128 * k), idesc_o, k != 0);
                    umma_ts(tb + SUM_COL, tb + (uint32_t)(8 * k), odesc, idesc_1, k != 0);   // += P . 1
                }
                umma_commit(&bar_o[slot]);
                if (last) umma_commit(&bar_free[stage]);      // every MMA that reads this stage has retired
                if (p.dbg && blockIdx.x == 0 && j < 32) p.dbg[j * 2 + 1] = clock64();
            }
            __syncwarp();
            if (last && u + NST < n_my) {             // the stage is free once those MMAs have retired: fetch the unit NST ahead
                mbar_wait(&bar_free[stage], (uint32_t)(u / NST) & 1);
                if (elect_one()) load_unit(u + NST);
                __syncwarp();
            }
        };
        for (int j = 0; j < n_items; ++j) {
            issue_s(j);
            if (j >= 1) issue_pv(j - 1);
        }
        if (n_items >= 1) issue_pv(n_items - 1);
    } else {
        // ===================== softmax + epilogue: two threads per query row =====================
        const int slot = (warp - 1) >> 3;             // tensor-memory slot served by this warp
        const int half = ((warp - 1) >> 2) & 1;       // 0: even 32-column chunks (and the row reference), 1: odd chunks
        const int quarter = warp & 3;                 // TMEM lane quarter (hardware: a warp reaches lanes 32 * (warp % 4) .. + 31)
        const int pair_id = 1 + slot * 4 + quarter;   // named barrier of the two warps that share the rows
        const uint32_t lane_base = tmem_base + (uint32_t)(slot * TILE_COLS) + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f;
        const int nch = (KP + 31) >> 5;               // 32-column chunks of S
        const int n_steps = max(2, (nch - half + 1) >> 1);
        float *ref = &s_ref[slot][quarter * 32 + lane];
        constexpr int f16 = F16;
        int k = 0;
        for (int j = slot; j < n_items; j += 2, ++k) {
            const int u = j / MT, t = j - u * MT;
            const int unit = (int)blockIdx.x + u * (int)gridDim.x;
            const int img = unit / heads, h = unit % heads;
            const int row = t * 128 + quarter * 32 + lane;
            const bool live = t * 128 + quarter * 32 < L;     // pair-uniform: these warps own at least one real query row
            const uint32_t ph = (uint32_t)k & 1;
            long long *dbg = (p.dbg && blockIdx.x == 0 && quarter == 0 && half == 0 && lane == 0 && j < 32) ? p.dbg + 64 + j * 4 : nullptr;
            mbar_wait(&bar_s[slot], ph);
            tc_fence_after();
            if (dbg) dbg[0] = clock64();
            float ms = 0.f;
            if (live) {
                const int klim = p.causal ? min(L, row + 1) : L;     // keys [0, klim) are visible to this row
                if (p.exact) {
                    // exact row maximum: one extra pass over this thread's chunks, partial maxima exchanged through shared memory.
                    // With it no exponent is positive, so the 16-bit P cannot overflow whatever the scores are (fp16 operands:
                    // a key 11 nats above the first 32 would otherwise give inf).
                    float m = -INFINITY;
                    for (int c = half; c < nch; c += 2) {
                        uint32_t v[32];
                        tmem_ld32_issue(lane_base + (uint32_t)(c * 32), v);
                        tmem_ld_wait();
                        if ((c + 1) * 32 <= klim) {
                            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
                            for (int q = 2; q < 32; q += 2) {
                                m0 = fmaxf(m0, __uint_as_float(v[q]));
                                m1 = fmaxf(m1, __uint_as_float(v[q + 1]));
                            }
                            m = fmaxf(m, fmaxf(m0, m1));
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; ++q)
                                if (c * 32 + q < klim) m = fmaxf(m, __uint_as_float(v[q]));
                        }
                    }
                    float *mine = half ? &s_ref1[slot][quarter * 32 + lane] : ref;
                    float *other = half ? ref : &s_ref1[slot][quarter * 32 + lane];
                    *mine = m;
                    pair_sync(pair_id);
                    ms = fmaxf(m, *other) * sl2;
                }
                for (int st = 0; st < n_steps; ++st) {
                    const int c = 2 * st + half;
                    const bool has = c < nch;
                    uint32_t v[32];
                    if (has) {
                        tmem_ld32_issue(lane_base + (uint32_t)(c * 32), v);
                        tmem_ld_wait();
                    }
                    if (st == 0) {
                        if (!p.exact && half == 0) {
                            float m = -INFINITY;
#pragma unroll
                            for (int q = 0; q < 32; ++q)
                                if (q < klim) m = fmaxf(m, __uint_as_float(v[q]));
                            ms = m * sl2;
                            *ref = ms;
                        }
                        pair_sync(pair_id);           // reference visible; chunks 0 and 1 are in registers
                        if (!p.exact && half == 1) ms = *ref;
                    } else if (st == 1)
                        pair_sync(pair_id);           // chunks 2 and 3 are in registers: P chunks 5 and 6 may overwrite them
                    if (has) {
                        uint32_t pk[16];
                        const f2 sl2x = mk2(sl2, sl2), msx = mk2(-ms, -ms);
                        if ((c + 1) * 32 <= klim) {
#pragma unroll
                            for (int q = 0; q < 32; q += 2) {
                                float x0, x1;
                                un2(fma2(mk2u(v[q], v[q + 1]), sl2x, msx), x0, x1);
                                x0 = fminf(x0, 120.f);
                                x1 = fminf(x1, 120.f);
                                pk[q >> 1] = pack16x2(fast_exp2(x0), fast_exp2(x1), f16);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; q += 2) {
                                float x0 = fminf(fmaf(__uint_as_float(v[q]), sl2, -ms), 120.f);
                                float x1 = fminf(fmaf(__uint_as_float(v[q + 1]), sl2, -ms), 120.f);
                                if (c * 32 + q >= klim) x0 = -INFINITY;
                                if (c * 32 + q + 1 >= klim) x1 = -INFINITY;
                                pk[q >> 1] = pack16x2(fast_exp2(x0), fast_exp2(x1), f16);
                            }
                        }
                        tmem_st16(lane_base + (uint32_t)(c * 16), pk);
                    }
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[slot]);
            if (dbg) dbg[1] = clock64();
            // epilogue: O / rowsum -> 16 bits -> global; this warp takes 32 of the 64 head columns
            mbar_wait(&bar_o[slot], ph);
            tc_fence_after();
            if (dbg) dbg[2] = clock64();
            if (live) {
                uint32_t v[32];
                tmem_ld32_issue(lane_base + O_COL + (uint32_t)(half * 32), v);
                const float rsum = tmem_ld1(lane_base + SUM_COL);      // (waits for both loads)
                const float inv = 1.f / rsum;
                if (row < L) {
                    if (p.lse && half == 0) p.lse[(size_t)unit * L + row] = ms + log2f(rsum);     // p_ij = exp2(s_ij * sl2 - lse)
                    __nv_bfloat16 *orow = p.out + ((size_t)img * L + row) * d + h * HD + half * 32;
#pragma unroll
                    for (int q = 0; q < 32; q += 8) {
                        uint4 o;
                        o.x = pack16x2(__uint_as_float(v[q]) * inv, __uint_as_float(v[q + 1]) * inv, f16);
                        o.y = pack16x2(__uint_as_float(v[q + 2]) * inv, __uint_as_float(v[q + 3]) * inv, f16);
                        o.z = pack16x2(__uint_as_float(v[q + 4]) * inv, __uint_as_float(v[q + 5]) * inv, f16);
                        o.w = pack16x2(__uint_as_float(v[q + 6]) * inv, __uint_as_float(v[q + 7]) * inv, f16);
                        *reinterpret_cast<uint4 *>(orow + q) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_oe[slot]);
            if (dbg) dbg[3] = clock64();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ================================================================================================================
// ViT-L/14 (L = 257 = class token + 16 x 16 patches) on the two-slot kernel.
//
// 257 keys need 272 score columns and two slots of them do not fit tensor memory, which is why attention_tc_big_kernel runs ONE
// chain at a time.  Here the tensor cores see exactly 256 keys (keys 1 .. 256: K and V boxes start at row 1) and 256 queries
// (rows 0 .. 255), i.e. attention_tc2_kernel's geometry with two 256-column slots, and the two leftovers are folded in:
//   * key 0: every softmax thread computes its half of q_row . k_0 (q from the swizzled tile in shared memory, k_0 from global),
//     the two halves meet in the exchange of the exact-maximum pass; p_0 = exp2(s_0 - max) joins the denominator and p_0 v_0 is
//     added to the O row in the epilogue (fp32);
//   * query row 256: warps 17-19 compute it with mma.sync m16n8k16 from the K / V tiles in shared memory (the row is row 0 of the A
//     fragments), a third of the keys each with its own maximum; warp 17 combines the three partial rows.
// Always the exact row maximum.  Non-causal only.
// ================================================================================================================
constexpr int NTHREADS2X = 640;                 // warp 0: TMA + MMA issue; warps 1-16: softmax; warps 17-19: query row 256 (mma.sync)

__device__ __forceinline__ void unpack16x2(uint32_t w, int f16, float &a, float &b)
{
    if (f16) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w));
        a = f.x; b = f.y;
    } else {
        a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
    }
}

template <int F16>
__device__ __forceinline__ void mma16816(float (&dd)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(dd[0]), "+f"(dd[1]), "+f"(dd[2]), "+f"(dd[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(dd[0]), "+f"(dd[1]), "+f"(dd[2]), "+f"(dd[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int F16>
__global__ void __launch_bounds__(NTHREADS2X, 1)
attention_tc2x_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *sOnes = smem + 2 * STAGE_BYTES;    // all-ones operand: O's companion MMA yields the softmax denominators
    __shared__ __align__(8) uint64_t bar_qk[4], bar_v[4], bar_free[4];          // per smem stage
    __shared__ __align__(8) uint64_t bar_s[2], bar_p[2], bar_o[2], bar_oe[2];   // per tensor-memory slot
    __shared__ float s_ref[2][128];                                              // row reference, even -> odd warp of a pair
    __shared__ float s_ref1[2][128];                                             // the odd warp's partial row maximum
    __shared__ float s_dot[2][2][128];                                           // partial q . k_0 of the two warps of a pair
    __shared__ float s_tail[272];                                                // scores / probabilities of query row 256 (warps 17-19)
    __shared__ float s_tpart[3][66];                                             // their partial rows, maxima and sums
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.L, d = p.d, heads = p.heads;
    constexpr int MT = 2;                             // query rows 0 .. 255; row 256 belongs to warp 17
    const int KP = 256;                               // keys 1 .. 256 go through the tensor cores; key 0 is folded in by the softmax threads
    const int NST = 2;                                // units resident in shared memory
    const int stage_bytes = 3 * MT * TILE_BYTES;
    const int n_units = p.n_img * heads;
    const int n_my = (int)blockIdx.x < n_units ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_items = n_my * MT;                    // item j = (unit j / MT, tile j % MT) runs in slot j & 1

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_qk[i], 1); mbar_init(&bar_v[i], 1); mbar_init(&bar_free[i], 4); }   // bar_free: MMA commit + warps 17-19
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 8); mbar_init(&bar_o[i], 1); mbar_init(&bar_oe[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS2X) reinterpret_cast<uint32_t *>(sOnes)[i] = F16 ? 0x3c003c00u : 0x3f803f80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== TMA + MMA issuer =====================
        auto load_unit = [&](int u) {                 // elected lane only; u = this CTA's u-th unit
            const int unit = (int)blockIdx.x + u * (int)gridDim.x, stage = u % NST;
            const int img = unit / heads, h = unit % heads;
            unsigned char *sQ = smem + stage * stage_bytes, *sK = sQ + MT * TILE_BYTES, *sV = sQ + 2 * MT * TILE_BYTES;
            mbar_expect_tx(&bar_qk[stage], (uint32_t)(2 * MT * TILE_BYTES));
            for (int b = 0; b < MT; ++b) {
                tma_load_3d(sK + b * TILE_BYTES, &map_qkv, &bar_qk[stage], d + h * HD, 1 + b * 128, img);      // keys 1 .. 256
                tma_load_3d(sQ + b * TILE_BYTES, &map_qkv, &bar_qk[stage], h * HD, b * 128, img);
            }
            mbar_expect_tx(&bar_v[stage], (uint32_t)(MT * TILE_BYTES));
            for (int b = 0; b < MT; ++b) tma_load_3d(sV + b * TILE_BYTES, &map_qkv, &bar_v[stage], 2 * d + h * HD, 1 + b * 128, img);
        };
        const uint32_t FMT16 = F16 ? 0u : ((1u << 7) | (1u << 10));
        const uint32_t idesc_s = (1u << 4) | FMT16 | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_o = (1u << 4) | FMT16 | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_1 = (1u << 4) | FMT16 | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t odesc = make_desc(smem_u32(sOnes));
        if (elect_one())
            for (int u = 0; u < NST && u < n_my; ++u) load_unit(u);
        __syncwarp();
        auto issue_s = [&](int j) {
            const int u = j / MT, t = j - u * MT, stage = u % NST, slot = j & 1;
            mbar_wait(&bar_qk[stage], (uint32_t)(u / NST) & 1);
            if (j >= 2) mbar_wait(&bar_oe[slot], (uint32_t)((j - 2) >> 1) & 1);      // the slot's previous item is fully consumed
            tc_fence_after();
            unsigned char *sQ = smem + stage * stage_bytes, *sK = sQ + MT * TILE_BYTES;
            const uint64_t kdesc = make_desc(smem_u32(sK)), qdesc = make_desc(smem_u32(sQ + t * TILE_BYTES));
            const uint32_t tb = tmem_base + (uint32_t)(slot * TILE_COLS);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0);
                umma_commit(&bar_s[slot]);
                if (p.dbg && blockIdx.x == 0 && j < 32) p.dbg[j * 2] = clock64();
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int u = j / MT, t = j - u * MT, stage = u % NST, slot = j & 1;
            mbar_wait(&bar_v[stage], (uint32_t)(u / NST) & 1);
            mbar_wait(&bar_p[slot], (uint32_t)(j >> 1) & 1);                         // P is in tensor memory
            tc_fence_after();
            unsigned char *sV = smem + stage * stage_bytes + 2 * MT * TILE_BYTES;
            const uint64_t vdesc = make_desc(smem_u32(sV));
            const uint32_t tb = tmem_base + (uint32_t)(slot * TILE_COLS);
            const bool last = t == MT - 1;
            const int nk = KP / 16;
            if (elect_one()) {
                for (int k = 0; k < nk; ++k) {        // 16 keys per step: 8 packed columns of P, 16 rows (2048 B) of V
                    umma_ts(tb + O_COL, tb + (uint32_t)(8 * k), vdesc + (uint64_t)(128 * k), idesc_o, k != 0);
                    umma_ts(tb + SUM_COL, tb + (uint32_t)(8 * k), odesc, idesc_1, k != 0);   // += P . 1
                }
                umma_commit(&bar_o[slot]);
                if (last) umma_commit(&bar_free[stage]);      // every MMA that reads this stage has retired
                if (p.dbg && blockIdx.x == 0 && j < 32) p.dbg[j * 2 + 1] = clock64();
            }
            __syncwarp();
            if (last && u + NST < n_my) {             // the stage is free once those MMAs have retired: fetch the unit NST ahead
                mbar_wait(&bar_free[stage], (uint32_t)(u / NST) & 1);
                if (elect_one()) load_unit(u + NST);
                __syncwarp();
            }
        };
        for (int j = 0; j < n_items; ++j) {
            issue_s(j);
            if (j >= 1) issue_pv(j - 1);
        }
        if (n_items >= 1) issue_pv(n_items - 1);
    } else if (warp <= 16) {
        // ===================== softmax + epilogue: two threads per query row =====================
        const int slot = (warp - 1) >> 3;             // tensor-memory slot served by this warp
        const int half = ((warp - 1) >> 2) & 1;       // 0: even 32-column chunks (and the row reference), 1: odd chunks
        const int quarter = warp & 3;                 // TMEM lane quarter (hardware: a warp reaches lanes 32 * (warp % 4) .. + 31)
        const int pair_id = 1 + slot * 4 + quarter;   // named barrier of the two warps that share the rows
        const uint32_t lane_base = tmem_base + (uint32_t)(slot * TILE_COLS) + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f;
        const int nch = (KP + 31) >> 5;               // 32-column chunks of S
        const int n_steps = max(2, (nch - half + 1) >> 1);
        float *ref = &s_ref[slot][quarter * 32 + lane];
        constexpr int f16 = F16;
        int k = 0;
        for (int j = slot; j < n_items; j += 2, ++k) {
            const int u = j / MT, t = j - u * MT;
            const int unit = (int)blockIdx.x + u * (int)gridDim.x;
            const int img = unit / heads, h = unit % heads;
            const int row = t * 128 + quarter * 32 + lane;
            const bool live = t * 128 + quarter * 32 < L;     // pair-uniform: these warps own at least one real query row
            const uint32_t ph = (uint32_t)k & 1;
            long long *dbg = (p.dbg && blockIdx.x == 0 && quarter == 0 && half == 0 && lane == 0 && j < 32) ? p.dbg + 64 + j * 4 : nullptr;
            // key 0 (the class token's key): this thread's half of q_row . k_0, q from the swizzled tile in shared memory, k_0 from global
            const int stage = u % NST, rt = quarter * 32 + lane;
            const __nv_bfloat16 *kv0 = p.qkv + (size_t)img * L * 3 * d + d + h * HD;      // k_0 of this head; v_0 lies d elements further
            mbar_wait(&bar_qk[stage], (uint32_t)(u / NST) & 1);                           // the TMA writes of Q are visible to this thread
            float dp = 0.f;
            {
                const unsigned char *qrow = smem + stage * stage_bytes + t * TILE_BYTES + rt * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 qa = *reinterpret_cast<const uint4 *>(qrow + (((4 * half + c) ^ (rt & 7)) << 4));
                    const uint4 kb = __ldg(reinterpret_cast<const uint4 *>(kv0) + 4 * half + c);
                    const uint32_t qa4[4] = {qa.x, qa.y, qa.z, qa.w}, kb4[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float q0, q1, k0, k1;
                        unpack16x2(qa4[e], f16, q0, q1);
                        unpack16x2(kb4[e], f16, k0, k1);
                        dp = fmaf(q0, k0, fmaf(q1, k1, dp));
                    }
                }
            }
            mbar_wait(&bar_s[slot], ph);
            tc_fence_after();
            if (dbg) dbg[0] = clock64();
            float ms = 0.f, p0 = 0.f;
            if (live) {
                const int klim = 256;                                // every tensor-core key (1 .. 256) is visible
                {
                    // exact row maximum: one extra pass over this thread's chunks, partial maxima exchanged through shared memory.
                    // With it no exponent is positive, so the 16-bit P cannot overflow whatever the scores are (fp16 operands:
                    // a key 11 nats above the first 32 would otherwise give inf).
                    float m = -INFINITY;
                    for (int c = half; c < nch; c += 2) {
                        uint32_t v[32];
                        tmem_ld32_issue(lane_base + (uint32_t)(c * 32), v);
                        tmem_ld_wait();
                        if ((c + 1) * 32 <= klim) {
                            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
                            for (int q = 2; q < 32; q += 2) {
                                m0 = fmaxf(m0, __uint_as_float(v[q]));
                                m1 = fmaxf(m1, __uint_as_float(v[q + 1]));
                            }
                            m = fmaxf(m, fmaxf(m0, m1));
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; ++q)
                                if (c * 32 + q < klim) m = fmaxf(m, __uint_as_float(v[q]));
                        }
                    }
                    float *mine = half ? &s_ref1[slot][quarter * 32 + lane] : ref;
                    float *other = half ? ref : &s_ref1[slot][quarter * 32 + lane];
                    *mine = m;
                    s_dot[half][slot][rt] = dp;
                    pair_sync(pair_id);
                    const float s0 = dp + s_dot[half ^ 1][slot][rt];         // q_row . k_0
                    ms = fmaxf(fmaxf(m, *other), s0) * sl2;
                    p0 = fast_exp2(fmaf(s0, sl2, -ms));
                }
                for (int st = 0; st < n_steps; ++st) {
                    const int c = 2 * st + half;
                    const bool has = c < nch;
                    uint32_t v[32];
                    if (has) {
                        tmem_ld32_issue(lane_base + (uint32_t)(c * 32), v);
                        tmem_ld_wait();
                    }
                    if (st == 0) {
                        pair_sync(pair_id);           // chunks 0 and 1 are in registers
                    } else if (st == 1)
                        pair_sync(pair_id);           // chunks 2 and 3 are in registers: P chunks 5 and 6 may overwrite them
                    if (has) {
                        uint32_t pk[16];
                        const f2 sl2x = mk2(sl2, sl2), msx = mk2(-ms, -ms);
                        if ((c + 1) * 32 <= klim) {
#pragma unroll
                            for (int q = 0; q < 32; q += 2) {
                                float x0, x1;
                                un2(fma2(mk2u(v[q], v[q + 1]), sl2x, msx), x0, x1);
                                x0 = fminf(x0, 120.f);
                                x1 = fminf(x1, 120.f);
                                pk[q >> 1] = pack16x2(fast_exp2(x0), fast_exp2(x1), f16);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; q += 2) {
                                float x0 = fminf(fmaf(__uint_as_float(v[q]), sl2, -ms), 120.f);
                                float x1 = fminf(fmaf(__uint_as_float(v[q + 1]), sl2, -ms), 120.f);
                                if (c * 32 + q >= klim) x0 = -INFINITY;
                                if (c * 32 + q + 1 >= klim) x1 = -INFINITY;
                                pk[q >> 1] = pack16x2(fast_exp2(x0), fast_exp2(x1), f16);
                            }
                        }
                        tmem_st16(lane_base + (uint32_t)(c * 16), pk);
                    }
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[slot]);
            if (dbg) dbg[1] = clock64();
            // epilogue: O / rowsum -> 16 bits -> global; this warp takes 32 of the 64 head columns
            mbar_wait(&bar_o[slot], ph);
            tc_fence_after();
            if (dbg) dbg[2] = clock64();
            if (live) {
                uint32_t v[32];
                tmem_ld32_issue(lane_base + O_COL + (uint32_t)(half * 32), v);
                const float rsum = tmem_ld1(lane_base + SUM_COL) + p0;      // (waits for both loads); + key 0
                const float inv = 1.f / rsum;
                {
                    if (p.lse && half == 0) p.lse[(size_t)unit * L + row] = ms + log2f(rsum);     // p_ij = exp2(s_ij * sl2 - lse)
                    __nv_bfloat16 *orow = p.out + ((size_t)img * L + row) * d + h * HD + half * 32;
                    const uint4 *v0p = reinterpret_cast<const uint4 *>(kv0 + d) + 4 * half;       // v_0, this warp's 32 head columns
#pragma unroll
                    for (int q = 0; q < 32; q += 8) {
                        const uint4 vv = __ldg(v0p + (q >> 3));
                        float a0, a1, a2, a3, a4, a5, a6, a7;
                        unpack16x2(vv.x, f16, a0, a1); unpack16x2(vv.y, f16, a2, a3);
                        unpack16x2(vv.z, f16, a4, a5); unpack16x2(vv.w, f16, a6, a7);
                        uint4 o;
                        o.x = pack16x2(fmaf(p0, a0, __uint_as_float(v[q])) * inv, fmaf(p0, a1, __uint_as_float(v[q + 1])) * inv, f16);
                        o.y = pack16x2(fmaf(p0, a2, __uint_as_float(v[q + 2])) * inv, fmaf(p0, a3, __uint_as_float(v[q + 3])) * inv, f16);
                        o.z = pack16x2(fmaf(p0, a4, __uint_as_float(v[q + 4])) * inv, fmaf(p0, a5, __uint_as_float(v[q + 5])) * inv, f16);
                        o.w = pack16x2(fmaf(p0, a6, __uint_as_float(v[q + 6])) * inv, fmaf(p0, a7, __uint_as_float(v[q + 7])) * inv, f16);
                        *reinterpret_cast<uint4 *>(orow + q) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_oe[slot]);
            if (dbg) dbg[3] = clock64();
        }
    } else {
        // ===================== warps 17-19: query row 256 with mma.sync (m16n8k16, the row is row 0 of the A fragments) =====================
        // scores: A = q (16 x 16 per k step, rows 1-15 zero), B = K^T straight from the swizzled tile (two 32-bit loads per step);
        // P V: A = p rounded to 16 bits, B = V through ldmatrix.trans; key 0 (k_0, v_0 from global) is added with plain FMAs by warp 17.
        // Each warp takes a third of the keys (whole 16-key steps) with its own maximum; warp 17 combines the three partial rows.
        const float sl2 = 0.125f * 1.4426950408889634f;
        constexpr int f16 = F16;
        const int row = 256, g = lane >> 2, tq = lane & 3, w = warp - 17;
        const int ks0 = w == 0 ? 0 : (w == 1 ? 6 : 11), ks1 = w == 0 ? 6 : (w == 1 ? 11 : 16);      // 16-key steps [ks0, ks1) of the tiles
        const int jlo = w == 0 ? 0 : 1 + 16 * ks0, jhi = 1 + 16 * ks1;                              // s_tail indices (keys) of this warp
        for (int u = 0; u < n_my; ++u) {
            const int unit = (int)blockIdx.x + u * (int)gridDim.x, stage = u % NST;
            const int img = unit / heads, h = unit % heads;
            const uint32_t sph = (uint32_t)(u / NST) & 1;
            const __nv_bfloat16 *base = p.qkv + (size_t)img * L * 3 * d + h * HD;
            const uint32_t *qw = reinterpret_cast<const uint32_t *>(base + (size_t)row * 3 * d);     // 32 words = 64 dims of q_256
            uint32_t qa[4][2];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qa[ks][0] = g == 0 ? __ldg(qw + 8 * ks + tq) : 0u;
                qa[ks][1] = g == 0 ? __ldg(qw + 8 * ks + 4 + tq) : 0u;
            }
            if (w == 0) {                                     // key 0: every lane two dims
                float q0, q1, k0, k1;
                unpack16x2(__ldg(qw + lane), f16, q0, q1);
                unpack16x2(__ldg(reinterpret_cast<const uint32_t *>(base + d) + lane), f16, k0, k1);
                float s0 = fmaf(q0, k0, q1 * k1);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                if (lane == 0) s_tail[0] = s0 * sl2;
            }
            mbar_wait(&bar_qk[stage], sph);
            const uint32_t sK_s = smem_u32(smem + stage * stage_bytes + MT * TILE_BYTES);
            const uint32_t sV_s = smem_u32(smem + stage * stage_bytes + 2 * MT * TILE_BYTES);
            for (int nt = 2 * ks0; nt < 2 * ks1; nt += 2) {   // two independent 8-key tiles per pass: tile rows 8 nt + g, 8 nt + 8 + g
                const int kka = 8 * nt + g, kkb = kka + 8;
                const uint32_t ra = sK_s + (uint32_t)((kka >> 7) * TILE_BYTES + (kka & 127) * 128 + 4 * tq);
                const uint32_t rb = sK_s + (uint32_t)((kkb >> 7) * TILE_BYTES + (kkb & 127) * 128 + 4 * tq);
                float da[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t a0, a1, b0, b1;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a0) : "r"(ra + (uint32_t)(((2 * ks) ^ (kka & 7)) << 4)));
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a1) : "r"(ra + (uint32_t)(((2 * ks + 1) ^ (kka & 7)) << 4)));
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b0) : "r"(rb + (uint32_t)(((2 * ks) ^ (kkb & 7)) << 4)));
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b1) : "r"(rb + (uint32_t)(((2 * ks + 1) ^ (kkb & 7)) << 4)));
                    mma16816<F16>(da, qa[ks][0], 0u, qa[ks][1], 0u, a0, a1);
                    mma16816<F16>(db, qa[ks][0], 0u, qa[ks][1], 0u, b0, b1);
                }
                if (g == 0) {
                    s_tail[1 + 8 * nt + 2 * tq] = da[0] * sl2; s_tail[2 + 8 * nt + 2 * tq] = da[1] * sl2;
                    s_tail[9 + 8 * nt + 2 * tq] = db[0] * sl2; s_tail[10 + 8 * nt + 2 * tq] = db[1] * sl2;
                }
            }
            __syncwarp();
            float m = -INFINITY;
            for (int j = jlo + lane; j < jhi; j += 32) m = fmaxf(m, s_tail[j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float sum = 0.f;
            for (int j = jlo + lane; j < jhi; j += 32) {
                const float e = fast_exp2(s_tail[j] - m);
                s_tail[j] = e;
                sum += e;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            __syncwarp();
            mbar_wait(&bar_v[stage], sph);
            float oacc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f; }
            for (int ks = ks0; ks < ks1; ++ks) {              // sixteen keys per step: tile rows 16 ks .. + 15 = keys 1 + 16 ks ..
                uint32_t pa0 = 0u, pa2 = 0u;
                if (g == 0) {
                    pa0 = pack16x2(s_tail[1 + 16 * ks + 2 * tq], s_tail[2 + 16 * ks + 2 * tq], f16);
                    pa2 = pack16x2(s_tail[9 + 16 * ks + 2 * tq], s_tail[10 + 16 * ks + 2 * tq], f16);
                }
                const int vr = 16 * ks + (lane & 15);         // the row this lane addresses for ldmatrix (lanes 16-31 repeat 0-15)
                const uint32_t vrow = sV_s + (uint32_t)((vr >> 7) * TILE_BYTES + (vr & 127) * 128);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    uint32_t b0, b1;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                                 : "=r"(b0), "=r"(b1) : "r"(vrow + (uint32_t)((nt ^ (vr & 7)) << 4)));
                    mma16816<F16>(oacc[nt], pa0, 0u, pa2, 0u, b0, b1);
                }
            }
            // partial row (dims 8 nt + 2 tq, + 1 in the lanes with g == 0), maximum and sum of this warp's keys
            if (g == 0) {
                float p0 = 0.f;
                const uint32_t *v0w = reinterpret_cast<const uint32_t *>(base + 2 * d);              // v_0: 32 words
                if (w == 0) p0 = s_tail[0];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    float a = 0.f, b2 = 0.f;
                    if (w == 0) unpack16x2(__ldg(v0w + 4 * nt + tq), f16, a, b2);
                    s_tpart[w][8 * nt + 2 * tq] = fmaf(p0, a, oacc[nt][0]);
                    s_tpart[w][8 * nt + 2 * tq + 1] = fmaf(p0, b2, oacc[nt][1]);
                }
                if (tq == 0) { s_tpart[w][64] = m; s_tpart[w][65] = sum; }
            }
            asm volatile("bar.sync 9, 96;" ::: "memory");
            if (w == 0) {
                const float m0 = s_tpart[0][64], m1 = s_tpart[1][64], m2 = s_tpart[2][64];
                const float mm = fmaxf(m0, fmaxf(m1, m2));
                const float c0 = fast_exp2(m0 - mm), c1 = fast_exp2(m1 - mm), c2 = fast_exp2(m2 - mm);
                const float tot = s_tpart[0][65] * c0 + s_tpart[1][65] * c1 + s_tpart[2][65] * c2;
                const float inv = 1.f / tot;
                const float x0 = s_tpart[0][2 * lane] * c0 + s_tpart[1][2 * lane] * c1 + s_tpart[2][2 * lane] * c2;
                const float x1 = s_tpart[0][2 * lane + 1] * c0 + s_tpart[1][2 * lane + 1] * c1 + s_tpart[2][2 * lane + 1] * c2;
                __nv_bfloat16 *orow = p.out + ((size_t)img * L + row) * d + h * HD + 2 * lane;
                *reinterpret_cast<uint32_t *>(orow) = pack16x2(x0 * inv, x1 * inv, f16);
                if (p.lse && lane == 0) p.lse[(size_t)unit * L + row] = mm + log2f(tot);
            }
            asm volatile("bar.sync 9, 96;" ::: "memory");        // the partials have been read: the next unit may overwrite them
            if (lane == 0) mbar_arrive(&bar_free[stage]);        // this warp no longer reads the stage
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// ================================================================================================================
// Small sequences (L <= 64: ViT-B/32's 50 tokens): FOUR chains in flight.
//
// A (image, head) unit of 50 x 50 scores is a few hundred cycles of work for every pipe, but its chain S -> softmax -> P V ->
// epilogue costs thousands of cycles of latency (barrier hand-offs, tensor-memory round trips): with the two slots of
// attention_tc2_kernel a unit takes ~6 900 cycles.  Here a slot needs only 128 tensor-memory columns (S [0,64) with P packed over
// its first half, O [64,128); the denominators are summed by the row's own thread), so four units are in flight, each served by
// one warpgroup with ONE thread per query row; Q, K, V travel as 64-row boxes (24 KB per unit, eight units resident in shared
// memory, fetched four units ahead).  The 128-row MMA reads 64 more rows behind Q (this unit's K): finite garbage in accumulator
// rows 64-127, which no thread reads.  Warp 0 issues the S MMAs as far ahead as slots free up, warp 17 the TMA loads and the P V MMAs.
// ================================================================================================================
constexpr int S4_TILE = 64 * 128;               // 64 rows x 64 16-bit columns, SWIZZLE_128B
constexpr int S4_UNIT = 3 * S4_TILE;            // Q, K, V of one unit
constexpr int S4_STAGES = 8;
constexpr uint32_t S4_SLOT_W = 128, S4_O_COL = 64;
constexpr int NTHREADS4 = 576;                  // warp 0: S issue; warps 1-16: four softmax warpgroups; warp 17: TMA + P V issue

template <int F16>
__global__ void __launch_bounds__(NTHREADS4, 1)
attention_tc4_kernel(const __grid_constant__ CUtensorMap map_qkv64, const AttnParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_ld[S4_STAGES], bar_free[S4_STAGES];      // per smem stage
    __shared__ __align__(8) uint64_t bar_s[4], bar_p[4], bar_o[4], bar_oe[4];    // per tensor-memory slot
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.L, d = p.d, heads = p.heads;
    const int KP = (L + 15) & ~15;                    // <= 64
    const int n_units = p.n_img * heads;
    const int n_my = (int)blockIdx.x < n_units ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv64) : "memory");
        for (int i = 0; i < S4_STAGES; ++i) { mbar_init(&bar_ld[i], 1); mbar_init(&bar_free[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 4); mbar_init(&bar_o[i], 1); mbar_init(&bar_oe[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0 || warp == 17) {
        // ===================== warp 0: S issuer; warp 17: TMA + P V issuer (no head-of-line blocking between the two) =====================
        auto load_unit = [&](int u) {                 // elected lane only; u = this CTA's u-th unit
            const int unit = (int)blockIdx.x + u * (int)gridDim.x, stage = u % S4_STAGES;
            const int img = unit / heads, h = unit % heads;
            unsigned char *sQ = smem + stage * S4_UNIT;
            mbar_expect_tx(&bar_ld[stage], (uint32_t)S4_UNIT);
            tma_load_3d(sQ + S4_TILE, &map_qkv64, &bar_ld[stage], d + h * HD, 0, img);
            tma_load_3d(sQ, &map_qkv64, &bar_ld[stage], h * HD, 0, img);
            tma_load_3d(sQ + 2 * S4_TILE, &map_qkv64, &bar_ld[stage], 2 * d + h * HD, 0, img);
        };
        const uint32_t FMT16 = F16 ? 0u : ((1u << 7) | (1u << 10));
        const uint32_t idesc_s = (1u << 4) | FMT16 | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_o = (1u << 4) | FMT16 | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (warp == 17) {
            if (elect_one())
                for (int u = 0; u < S4_STAGES && u < n_my; ++u) load_unit(u);
            __syncwarp();
        }
        auto issue_s = [&](int j) {
            const int stage = j % S4_STAGES, slot = j & 3;
            mbar_wait(&bar_ld[stage], (uint32_t)(j / S4_STAGES) & 1);
            if (j >= 4) mbar_wait(&bar_oe[slot], (uint32_t)((j >> 2) - 1) & 1);      // the slot's previous unit is fully consumed
            tc_fence_after();
            unsigned char *sQ = smem + stage * S4_UNIT;
            const uint64_t qdesc = make_desc(smem_u32(sQ)), kdesc = make_desc(smem_u32(sQ + S4_TILE));
            const uint32_t tb = tmem_base + (uint32_t)slot * S4_SLOT_W;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0);
                umma_commit(&bar_s[slot]);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int stage = j % S4_STAGES, slot = j & 3;
            mbar_wait(&bar_p[slot], (uint32_t)(j >> 2) & 1);                         // P is in tensor memory
            tc_fence_after();
            const uint64_t vdesc = make_desc(smem_u32(smem + stage * S4_UNIT + 2 * S4_TILE));
            const uint32_t tb = tmem_base + (uint32_t)slot * S4_SLOT_W;
            const int nk = KP / 16;
            if (elect_one()) {
                for (int k = 0; k < nk; ++k)          // 16 keys per step: 8 packed columns of P, 16 rows (2048 B) of V
                    umma_ts(tb + S4_O_COL, tb + (uint32_t)(8 * k), vdesc + (uint64_t)(128 * k), idesc_o, k != 0);
                umma_commit(&bar_o[slot]);
                umma_commit(&bar_free[stage]);        // every MMA that reads this stage has retired
            }
            __syncwarp();
        };
        if (warp == 0) {
            for (int j = 0; j < n_my; ++j) issue_s(j);            // as far ahead as slots free up
        } else {
            for (int j = 0; j < n_my; ++j) {
                if (j >= 2 && j - 2 + S4_STAGES < n_my) {     // the P V of unit j - 2 retired long ago: refill its stage (no wait in practice)
                    mbar_wait(&bar_free[(j - 2) % S4_STAGES], (uint32_t)((j - 2) / S4_STAGES) & 1);
                    if (elect_one()) load_unit(j - 2 + S4_STAGES);
                    __syncwarp();
                }
                issue_pv(j);
            }
        }
    } else {
        // ===================== softmax + epilogue: one warpgroup per slot, one thread per query row =====================
        const int slot = (warp - 1) >> 2;
        const int quarter = warp & 3;                 // TMEM lane quarter (hardware: a warp reaches lanes 32 * (warp % 4) .. + 31)
        const uint32_t lane_base = tmem_base + (uint32_t)slot * S4_SLOT_W + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f;
        const int row = quarter * 32 + lane;
        const bool live = quarter * 32 < L;           // warp-uniform
        const int klim = p.causal ? min(L, row + 1) : L;
        const int nch = (KP + 31) >> 5;               // 1 or 2
        constexpr int f16 = F16;
        int k = 0;
        for (int j = slot; j < n_my; j += 4, ++k) {
            const int unit = (int)blockIdx.x + j * (int)gridDim.x;
            const int img = unit / heads, h = unit % heads;
            const uint32_t ph = (uint32_t)k & 1;
            mbar_wait(&bar_s[slot], ph);
            tc_fence_after();
            float ms = 0.f, rsum = 0.f;
            if (live) {
                uint32_t va[32];
                tmem_ld32_issue(lane_base, va);
                // row reference: exact maximum over the visible keys (the second chunk is loaded again for its exponentials:
                // 96 registers do not hold two chunks next to the packed probabilities)
                float m = -INFINITY;
                if (nch > 1) {
                    uint32_t vb[32];
                    tmem_ld32_issue(lane_base + 32u, vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (32 + q < klim) m = fmaxf(m, __uint_as_float(vb[q]));
                } else
                    tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    if (q < klim) m = fmaxf(m, __uint_as_float(va[q]));
                ms = m * sl2;
                f2 acc2 = mk2(0.f, 0.f);
                auto emit = [&](const uint32_t (&v)[32], int c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int q = 0; q < 32; q += 2) {
                        float x0 = fmaf(__uint_as_float(v[q]), sl2, -ms), x1 = fmaf(__uint_as_float(v[q + 1]), sl2, -ms);
                        if (c * 32 + q >= klim) x0 = -INFINITY;
                        if (c * 32 + q + 1 >= klim) x1 = -INFINITY;
                        const float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
                        acc2 = add2(acc2, mk2(e0, e1));
                        pk[q >> 1] = pack16x2(e0, e1, f16);
                    }
                    tmem_st16(lane_base + (uint32_t)(c * 16), pk);
                };
                emit(va, 0);
                if (nch > 1) {
                    tmem_ld32_issue(lane_base + 32u, va);
                    tmem_ld_wait();
                    emit(va, 1);
                }
                float a0, a1;
                un2(acc2, a0, a1);
                rsum = a0 + a1;
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[slot]);
            // epilogue: O / rowsum -> 16 bits -> global
            mbar_wait(&bar_o[slot], ph);
            tc_fence_after();
            if (live) {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(lane_base + S4_O_COL, v0);
                tmem_ld32_issue(lane_base + S4_O_COL + 32u, v1);
                tmem_ld_wait();
                const float inv = 1.f / rsum;
                if (row < L) {
                    if (p.lse) p.lse[(size_t)unit * L + row] = ms + log2f(rsum);     // p_ij = exp2(s_ij * sl2 - lse)
                    __nv_bfloat16 *orow = p.out + ((size_t)img * L + row) * d + h * HD;
#pragma unroll
                    for (int q = 0; q < 32; q += 8) {
                        uint4 o;
                        o.x = pack16x2(__uint_as_float(v0[q]) * inv, __uint_as_float(v0[q + 1]) * inv, f16);
                        o.y = pack16x2(__uint_as_float(v0[q + 2]) * inv, __uint_as_float(v0[q + 3]) * inv, f16);
                        o.z = pack16x2(__uint_as_float(v0[q + 4]) * inv, __uint_as_float(v0[q + 5]) * inv, f16);
                        o.w = pack16x2(__uint_as_float(v0[q + 6]) * inv, __uint_as_float(v0[q + 7]) * inv, f16);
                        *reinterpret_cast<uint4 *>(orow + q) = o;
                        o.x = pack16x2(__uint_as_float(v1[q]) * inv, __uint_as_float(v1[q + 1]) * inv, f16);
                        o.y = pack16x2(__uint_as_float(v1[q + 2]) * inv, __uint_as_float(v1[q + 3]) * inv, f16);
                        o.z = pack16x2(__uint_as_float(v1[q + 4]) * inv, __uint_as_float(v1[q + 5]) * inv, f16);
                        o.w = pack16x2(__uint_as_float(v1[q + 6]) * inv, __uint_as_float(v1[q + 7]) * inv, f16);
                        *reinterpret_cast<uint4 *>(orow + 32 + q) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_oe[slot]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// Variant for 256 < L <= 384 (ViT-L/14: 257 tokens).  The S row no longer fits twice in tensor memory, so the 128-query
// tiles of a unit are processed one after the other by a single softmax warpgroup:
//   TMEM  S [0,384)  P [0,192)  O [384,448)  sums [448,464);   S = Q K^T is issued as N = 256 plus N = KP-256 MMAs.
// Shared memory holds one unit (3 tiles each of Q, K, V); Q/K of the next unit are fetched as soon as the last S MMA
// of the current one has retired, V after the last P V MMA.
constexpr uint32_t BIG_O_COL = 384, BIG_SUM_COL = 448;
constexpr int BIG_NTHREADS = 192;                // warp 0: TMA + MMA issue; warps 1-4: softmax; warp 5: rows >= 256 when there are at most BIG_TAIL_MAX
constexpr int BIG_TILES = 3;
constexpr int BIG_TAIL_MAX = 4;

__global__ void __launch_bounds__(BIG_NTHREADS, 1)
attention_tc_big_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *sQ = smem, *sK = smem + BIG_TILES * TILE_BYTES, *sV = smem + 2 * BIG_TILES * TILE_BYTES;
    unsigned char *sOnes = smem + 3 * BIG_TILES * TILE_BYTES;
    __shared__ __align__(8) uint64_t bar_qk, bar_v, bar_qk_free, bar_v_free, bar_s, bar_p, bar_o, bar_oe;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_tail[BIG_TAIL_MAX][BIG_TILES * 128];     // scores / probabilities of the tail rows (warp 5)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.L, d = p.d, heads = p.heads;
    const int KP = (L + 15) & ~15;
    const int MT_all = (L + 127) >> 7;                // tiles of Q, K, V in shared memory: 3 for L = 257
    // A unit of L = 257 is two full 128-query tiles plus ONE row (the class token's 256 patches + itself).  As a third tile that row
    // would cost a whole S -> softmax -> P V -> epilogue chain (a third of the kernel's time); instead warp 5 computes the rows
    // >= 256 with plain FMAs from the operands in shared memory while the tensor cores work on the two full tiles.
    const int n_tail = (L > 256 && L - 256 <= BIG_TAIL_MAX) ? L - 256 : 0;
    const int MT = n_tail ? 2 : MT_all;               // 128-query tiles that go through the tensor cores
    const int n_units = p.n_img * heads;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
        mbar_init(&bar_qk, 1); mbar_init(&bar_v, 1);
        mbar_init(&bar_qk_free, n_tail ? 2 : 1); mbar_init(&bar_v_free, n_tail ? 2 : 1);      // + the tail warp's arrival
        mbar_init(&bar_s, 1); mbar_init(&bar_p, 4); mbar_init(&bar_o, 1); mbar_init(&bar_oe, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += BIG_NTHREADS) reinterpret_cast<uint32_t *>(sOnes)[i] = p.f16 ? 0x3c003c00u : 0x3f803f80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        auto load_qk = [&](int unit) {
            const int img = unit / heads, h = unit % heads;
            mbar_expect_tx(&bar_qk, (uint32_t)(2 * MT_all * TILE_BYTES));
            for (int b = 0; b < MT_all; ++b) {
                tma_load_3d(sK + b * TILE_BYTES, &map_qkv, &bar_qk, d + h * HD, b * 128, img);
                tma_load_3d(sQ + b * TILE_BYTES, &map_qkv, &bar_qk, h * HD, b * 128, img);
            }
        };
        auto load_v = [&](int unit) {
            const int img = unit / heads, h = unit % heads;
            mbar_expect_tx(&bar_v, (uint32_t)(MT_all * TILE_BYTES));
            for (int b = 0; b < MT_all; ++b) tma_load_3d(sV + b * TILE_BYTES, &map_qkv, &bar_v, 2 * d + h * HD, b * 128, img);
        };
        const int n1 = KP < 256 ? KP : 256, n2 = KP - n1;          // S is issued in two column blocks
        const uint32_t FMT16 = p.f16 ? 0u : ((1u << 7) | (1u << 10));
        const uint32_t idesc_s1 = (1u << 4) | FMT16 | ((uint32_t)(n1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_s2 = (1u << 4) | FMT16 | ((uint32_t)(n2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_o = (1u << 4) | FMT16 | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_1 = (1u << 4) | FMT16 | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t odesc = make_desc(smem_u32(sOnes));
        const uint64_t kdesc = make_desc(smem_u32(sK)), k2desc = make_desc(smem_u32(sK + 2 * TILE_BYTES));
        const uint64_t vdesc = make_desc(smem_u32(sV));
        if ((int)blockIdx.x < n_units && elect_one()) { load_qk(blockIdx.x); load_v(blockIdx.x); }
        __syncwarp();
        int i = 0;
        uint32_t n = 0;                                   // tiles issued so far (parity source of the per-tile barriers)
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
            const uint32_t uph = (uint32_t)i & 1;
            const int next = unit + gridDim.x;
            mbar_wait(&bar_qk, uph);
            tc_fence_after();
            for (int t = 0; t < MT; ++t, ++n) {
                if (n >= 1) { mbar_wait(&bar_oe, (n - 1) & 1); tc_fence_after(); }   // previous tile's TMEM fully consumed
                const uint64_t qdesc = make_desc(smem_u32(sQ + t * TILE_BYTES));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_ss(tmem_base, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s1, k != 0);
                        if (n2 > 0) umma_ss(tmem_base + 256, qdesc + (uint64_t)(2 * k), k2desc + (uint64_t)(2 * k), idesc_s2, k != 0);
                    }
                    umma_commit(&bar_s);
                    if (t == MT - 1) umma_commit(&bar_qk_free);
                }
                __syncwarp();
                if (t == MT - 1 && next < n_units) {          // Q and K of this unit are dead: fetch the next unit's
                    mbar_wait(&bar_qk_free, uph);
                    if (elect_one()) load_qk(next);
                    __syncwarp();
                }
                mbar_wait(&bar_p, n & 1);                    // P is in tensor memory
                if (t == 0) mbar_wait(&bar_v, uph);
                tc_fence_after();
                if (elect_one()) {
                    for (int j = 0; j < KP / 16; ++j) {
                        umma_ts(tmem_base + BIG_O_COL, tmem_base + (uint32_t)(8 * j), vdesc + (uint64_t)(128 * j), idesc_o, j != 0);
                        umma_ts(tmem_base + BIG_SUM_COL, tmem_base + (uint32_t)(8 * j), odesc, idesc_1, j != 0);
                    }
                    umma_commit(&bar_o);
                    if (t == MT - 1) umma_commit(&bar_v_free);
                }
                __syncwarp();
            }
            if (next < n_units) {
                mbar_wait(&bar_v_free, uph);
                if (elect_one()) load_v(next);
                __syncwarp();
            }
        }
    } else if (warp <= 4) {
        const int quarter = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int nch = (KP + 31) >> 5;
        uint32_t n = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int img = unit / heads, h = unit % heads;
            for (int t = 0; t < MT; ++t, ++n) {
                const int row = t * 128 + quarter * 32 + lane;
                const bool live = t * 128 + quarter * 32 < L;
                mbar_wait(&bar_s, n & 1);
                tc_fence_after();
                softmax_tile<BIG_O_COL, BIG_SUM_COL, 3>(lane_base, L, nch, live, row, lane, &bar_p, &bar_o, n & 1, &bar_oe,
                                                     p.out + ((size_t)img * L + row) * d + h * HD, p.causal,
                                                     p.lse ? p.lse + (size_t)unit * L + row : nullptr, p.f16, p.exact);
            }
        }
    } else if (n_tail) {
        // ===================== tail rows (>= 256) on the FMA pipe, one warp =====================
        // element (r, col) of a 128 x 64 SWIZZLE_128B tile: r * 128 + (((col >> 3) ^ (r & 7)) << 4) + (col & 7) * 2 bytes
        const float sl2 = 0.125f * 1.4426950408889634f;
        const int f16 = p.f16;
        auto unpack = [&](uint32_t w, float &a, float &b) {
            if (f16) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w));
                a = f.x; b = f.y;
            } else {
                a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
            }
        };
        int i = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
            const uint32_t uph = (uint32_t)i & 1;
            const int img = unit / heads, h = unit % heads;
            mbar_wait(&bar_qk, uph);
            for (int tr = 0; tr < n_tail; ++tr) {
                const int row = 256 + tr;
                const int klim = p.causal ? min(L, row + 1) : L;
                float q[64];
                const unsigned char *qrow = sQ + 2 * TILE_BYTES + tr * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(qrow + ((c ^ (tr & 7)) << 4));
                    unpack(w.x, q[8 * c], q[8 * c + 1]); unpack(w.y, q[8 * c + 2], q[8 * c + 3]);
                    unpack(w.z, q[8 * c + 4], q[8 * c + 5]); unpack(w.w, q[8 * c + 6], q[8 * c + 7]);
                }
                for (int j = lane; j < BIG_TILES * 128; j += 32) {
                    float sc = -INFINITY;
                    if (j < klim) {
                        const unsigned char *krow = sK + (j >> 7) * TILE_BYTES + (j & 127) * 128;
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 w = *reinterpret_cast<const uint4 *>(krow + ((c ^ (j & 7)) << 4));
                            float k0, k1;
                            unpack(w.x, k0, k1); a0 = fmaf(q[8 * c], k0, a0); a1 = fmaf(q[8 * c + 1], k1, a1);
                            unpack(w.y, k0, k1); a0 = fmaf(q[8 * c + 2], k0, a0); a1 = fmaf(q[8 * c + 3], k1, a1);
                            unpack(w.z, k0, k1); a0 = fmaf(q[8 * c + 4], k0, a0); a1 = fmaf(q[8 * c + 5], k1, a1);
                            unpack(w.w, k0, k1); a0 = fmaf(q[8 * c + 6], k0, a0); a1 = fmaf(q[8 * c + 7], k1, a1);
                        }
                        sc = (a0 + a1) * sl2;
                    }
                    s_tail[tr][j] = sc;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_qk_free);        // Q and K of this unit are no longer read by this warp
            mbar_wait(&bar_v, uph);
            for (int tr = 0; tr < n_tail; ++tr) {
                const int row = 256 + tr;
                const int klim = p.causal ? min(L, row + 1) : L;
                float m = -INFINITY;
                for (int j = lane; j < BIG_TILES * 128; j += 32) m = fmaxf(m, s_tail[tr][j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float sum = 0.f;
                for (int j = lane; j < BIG_TILES * 128; j += 32) {
                    const float e = fast_exp2(s_tail[tr][j] - m);          // exp2(-inf) = 0 for the keys this row does not see
                    s_tail[tr][j] = e;
                    sum += e;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                __syncwarp();
                // this lane's two head columns 2 * lane, 2 * lane + 1: chunk lane >> 2, byte (lane & 3) * 4 inside it
                float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                const int cb = lane >> 2, inb = (lane & 3) * 4;
                int j = 0;
                for (; j + 1 < klim; j += 2) {
                    const unsigned char *v0 = sV + (j >> 7) * TILE_BYTES + (j & 127) * 128 + ((cb ^ (j & 7)) << 4) + inb;
                    const unsigned char *v1 = sV + ((j + 1) >> 7) * TILE_BYTES + ((j + 1) & 127) * 128 + ((cb ^ ((j + 1) & 7)) << 4) + inb;
                    float a, b, c2, d2;
                    unpack(*reinterpret_cast<const uint32_t *>(v0), a, b);
                    unpack(*reinterpret_cast<const uint32_t *>(v1), c2, d2);
                    const float p0 = s_tail[tr][j], p1 = s_tail[tr][j + 1];
                    o0 = fmaf(p0, a, o0); o1 = fmaf(p0, b, o1);
                    o2 = fmaf(p1, c2, o2); o3 = fmaf(p1, d2, o3);
                }
                if (j < klim) {
                    const unsigned char *v0 = sV + (j >> 7) * TILE_BYTES + (j & 127) * 128 + ((cb ^ (j & 7)) << 4) + inb;
                    float a, b;
                    unpack(*reinterpret_cast<const uint32_t *>(v0), a, b);
                    const float p0 = s_tail[tr][j];
                    o0 = fmaf(p0, a, o0); o1 = fmaf(p0, b, o1);
                }
                const float inv = 1.f / sum;
                __nv_bfloat16 *orow = p.out + ((size_t)img * L + row) * d + h * HD + 2 * lane;
                *reinterpret_cast<uint32_t *>(orow) = pack16x2((o0 + o2) * inv, (o1 + o3) * inv, f16);
                if (p.lse && lane == 0) p.lse[(size_t)unit * L + row] = m + log2f(sum);
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&bar_v_free);         // V of this unit is no longer read by this warp
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ================================================================================================================
// Backward on the tensor cores (L <= 256).  One persistent CTA per SM walks over (image, head) units.
//   shared memory  Q, K, V, dO as two [128 x 64] bf16 tiles each (TMA, rows >= L zero-filled), plus P and dS of the
//                  current (128 queries x 128 keys) block as bf16 operand tiles written by the softmax warps
//   tensor memory  S [0,128)  dP [128,256)  dQ (both query tiles) [256,384)  dK [384,448)  dV [448,512)
//   warp 0  TMA;  warp 1  MMA issue;  warps 2-5  one thread per row: P = exp2(S*sl2 - lse), dS = P*(dP - D)/8 -> smem,
//           and the epilogues (dK, dV per key tile; dQ per unit) -> global.
// Per (key tile j, query tile t):  S = Q_t K_j^T, dP = dO_t V_j^T  (K-major operands)
//                                  dV_j += P^T dO_t, dK_j += dS^T Q_t  (A = the P / dS tile read MN-major, B MN-major)
//                                  dQ_t += dS K_j                      (A = the dS tile read K-major, B MN-major)
// The same P / dS tile serves as a K-major and as an MN-major operand: only the descriptor changes.
constexpr int BWD_THREADS = 192;
constexpr uint32_t B_S = 0, B_DP = 128, B_DQ = 256, B_DK = 384, B_DV = 448;
constexpr int BWD_SMEM = 8 * TILE_BYTES + 2 * 2 * TILE_BYTES;      // inputs + P + dS (two 64-key sub-tiles each)

struct AttnBwdParams {
    __nv_bfloat16 *dqkv;
    const __nv_bfloat16 *o;
    const float *lse;
    int L, heads, d, n_img;
};

// MN-major operand: 8-row groups 1024 bytes apart along K (SBO), 64-wide atoms along M/N `lbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, const AttnBwdParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *sQ = smem, *sK = smem + 2 * TILE_BYTES, *sV = smem + 4 * TILE_BYTES, *sdO = smem + 6 * TILE_BYTES;
    unsigned char *sP = smem + 8 * TILE_BYTES, *sdS = smem + 10 * TILE_BYTES;
    __shared__ __align__(8) uint64_t bar_load, bar_done, bar_s, bar_p, bar_free, bar_kv, bar_kv_free, bar_q, bar_q_free;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.L, d = p.d, heads = p.heads;
    const int MT = (L + 127) >> 7;
    const int n_units = p.n_img * heads;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
        mbar_init(&bar_load, 1); mbar_init(&bar_done, 1); mbar_init(&bar_s, 1); mbar_init(&bar_p, 4); mbar_init(&bar_free, 1);
        mbar_init(&bar_kv, 1); mbar_init(&bar_kv_free, 4); mbar_init(&bar_q, 1); mbar_init(&bar_q_free, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int i = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
            const int img = unit / heads, h = unit % heads;
            if (i > 0) mbar_wait(&bar_done, (uint32_t)(i - 1) & 1);      // every MMA of the previous unit has retired
            if (elect_one()) {
                mbar_expect_tx(&bar_load, (uint32_t)(4 * MT * TILE_BYTES));
                for (int t = 0; t < MT; ++t) {
                    tma_load_3d(sQ + t * TILE_BYTES, &map_qkv, &bar_load, h * HD, t * 128, img);
                    tma_load_3d(sK + t * TILE_BYTES, &map_qkv, &bar_load, d + h * HD, t * 128, img);
                    tma_load_3d(sV + t * TILE_BYTES, &map_qkv, &bar_load, 2 * d + h * HD, t * 128, img);
                    tma_load_3d(sdO + t * TILE_BYTES, &map_do, &bar_load, h * HD, t * 128, img);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // instruction descriptors: D fp32, A/B bf16, M = 128
        const uint32_t idesc_s0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);     // N is filled in per key tile
        const uint32_t idesc_kv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_q = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t p_mn = make_desc_mn(smem_u32(sP), TILE_BYTES), ds_mn = make_desc_mn(smem_u32(sdS), TILE_BYTES);
        const uint64_t ds_k0 = make_desc(smem_u32(sdS)), ds_k1 = make_desc(smem_u32(sdS + TILE_BYTES));
        uint32_t it = 0, kvc = 0;
        int i = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
            mbar_wait(&bar_load, (uint32_t)i & 1);
            tc_fence_after();
            for (int j = 0; j < MT; ++j) {
                const uint64_t k_k = make_desc(smem_u32(sK + j * TILE_BYTES)), v_k = make_desc(smem_u32(sV + j * TILE_BYTES));
                const uint64_t k_mn = make_desc_mn(smem_u32(sK + j * TILE_BYTES), TILE_BYTES);
                const int kj = (min(128, L - 128 * j) + 15) >> 4;       // 16-key steps that hold real keys in this tile
                const uint32_t idesc_s = idesc_s0 | ((uint32_t)(kj * 2) << 17);      // N = 16 * kj
                for (int t = 0; t < MT; ++t, ++it) {
                    const int qt = (min(128, L - 128 * t) + 15) >> 4;   // 16-query steps that hold real queries
                    const uint64_t q_k = make_desc(smem_u32(sQ + t * TILE_BYTES)), do_k = make_desc(smem_u32(sdO + t * TILE_BYTES));
                    const uint64_t q_mn = make_desc_mn(smem_u32(sQ + t * TILE_BYTES), TILE_BYTES);
                    const uint64_t do_mn = make_desc_mn(smem_u32(sdO + t * TILE_BYTES), TILE_BYTES);
                    // S and dP may overwrite tensor memory as soon as the softmax warps have consumed the previous block
                    // (they signalled bar_p for it, waited on below in the previous iteration)
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + B_S, q_k + (uint64_t)(2 * k), k_k + (uint64_t)(2 * k), idesc_s, k != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + B_DP, do_k + (uint64_t)(2 * k), v_k + (uint64_t)(2 * k), idesc_s, k != 0);
                        umma_commit(&bar_s);
                    }
                    __syncwarp();
                    mbar_wait(&bar_p, it & 1);                                   // P and dS are in shared memory
                    if (t == 0 && kvc > 0) mbar_wait(&bar_kv_free, (kvc - 1) & 1);   // dK / dV of the previous key tile were read out
                    if (j == 0 && i > 0) mbar_wait(&bar_q_free, (uint32_t)(i - 1) & 1);   // dQ of the previous unit was read out
                    tc_fence_after();
                    if (elect_one()) {
                        // rows / columns of P and dS beyond the real queries / keys are never written: their steps are skipped
                        for (int k = 0; k < qt; ++k)     // 16 queries per step = 2048 bytes of the query-row-major tiles
                            umma_ss(tmem_base + B_DV, p_mn + (uint64_t)(128 * k), do_mn + (uint64_t)(128 * k), idesc_kv, (t | k) != 0);
                        for (int k = 0; k < qt; ++k)
                            umma_ss(tmem_base + B_DK, ds_mn + (uint64_t)(128 * k), q_mn + (uint64_t)(128 * k), idesc_kv, (t | k) != 0);
                        for (int k = 0; k < kj; ++k)     // 16 keys per step: 32 bytes inside a 64-key sub-tile of dS, 2048 bytes of K
                            umma_ss(tmem_base + B_DQ + (uint32_t)(64 * t), (k < 4 ? ds_k0 : ds_k1) + (uint64_t)(2 * (k & 3)),
                                    k_mn + (uint64_t)(128 * k), idesc_q, (j | k) != 0);
                        umma_commit(&bar_free);
                        if (t == MT - 1) umma_commit(&bar_kv);
                        if (t == MT - 1 && j == MT - 1) { umma_commit(&bar_q); umma_commit(&bar_done); }
                    }
                    __syncwarp();
                }
                ++kvc;
            }
        }
    } else {
        // ===================== softmax + epilogues: thread = row of a 128-row tile =====================
        const int quarter = warp & 3;                 // tensor-memory lane quarter this warp may access
        const int rt = quarter * 32 + lane;           // row inside a tile
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f;
        const uint32_t sw = (uint32_t)(rt & 7);
        uint32_t it = 0, kvc = 0;
        int i = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++i) {
            const int img = unit / heads, h = unit % heads;
            mbar_wait(&bar_load, (uint32_t)i & 1);
            // D = <dO, O> and the log-sum-exp of this thread's row in each query tile
            float Dv[2] = {0.f, 0.f}, lse[2] = {INFINITY, INFINITY};
            for (int t = 0; t < MT; ++t) {
                const int row = t * 128 + rt;
                if (row < L) {
                    const uint4 *po = reinterpret_cast<const uint4 *>(p.o + ((size_t)img * L + row) * d + h * HD);
                    const uint32_t srow = smem_u32(sdO + t * TILE_BYTES) + (uint32_t)rt * 128;
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 a = po[c], b = lds128u(srow + (((uint32_t)c ^ sw) << 4));
                        const __nv_bfloat162 *ah = reinterpret_cast<const __nv_bfloat162 *>(&a), *bh = reinterpret_cast<const __nv_bfloat162 *>(&b);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(bh[e]);
                            acc += x.x * y.x + x.y * y.y;
                        }
                    }
                    Dv[t] = acc;
                    lse[t] = p.lse[(size_t)unit * L + row];
                }
            }
            for (int j = 0; j < MT; ++j) {
                for (int t = 0; t < MT; ++t, ++it) {
                    mbar_wait(&bar_s, it & 1);
                    if (it > 0) mbar_wait(&bar_free, (it - 1) & 1);      // the MMAs that read the previous P / dS have retired
                    tc_fence_after();
                    const float lr = lse[t], Dr = Dv[t];
                    const int ncol = ((min(128, L - 128 * j) + 15) >> 4) * 16;       // S / dP columns the MMAs produced
                    const int nrow = ((min(128, L - 128 * t) + 15) >> 4) * 16;       // P / dS rows the MMAs will read
                    const int nc = quarter * 32 < nrow ? (ncol + 31) >> 5 : 0;       // warp-uniform: nothing to do for padding rows
#pragma unroll 1
                    for (int c = 0; c < nc; ++c) {
                        uint32_t vs[32], vp[32];
                        tmem_ld32_issue(lane_base + B_S + (uint32_t)(c * 32), vs);
                        tmem_ld32(lane_base + B_DP + (uint32_t)(c * 32), vp);        // waits for both loads
                        uint32_t pk[16], dk[16];
                        const int key0 = j * 128 + c * 32;
#pragma unroll
                        for (int e = 0; e < 32; e += 2) {
                            float p0 = fast_exp2(fmaf(__uint_as_float(vs[e]), sl2, -lr));
                            float p1 = fast_exp2(fmaf(__uint_as_float(vs[e + 1]), sl2, -lr));
                            float s0 = p0 * (__uint_as_float(vp[e]) - Dr) * 0.125f;
                            float s1 = p1 * (__uint_as_float(vp[e + 1]) - Dr) * 0.125f;
                            if (key0 + e >= L) { p0 = 0.f; s0 = 0.f; }          // select, not multiply: the columns may be stale
                            if (key0 + e + 1 >= L) { p1 = 0.f; s1 = 0.f; }
                            __nv_bfloat162 hp = __floats2bfloat162_rn(p0, p1), hs = __floats2bfloat162_rn(s0, s1);
                            pk[e >> 1] = *reinterpret_cast<uint32_t *>(&hp);
                            dk[e >> 1] = *reinterpret_cast<uint32_t *>(&hs);
                        }
                        // columns [32c, 32c+32) of this row: 64-key sub-tile c/2, 16-byte chunks (c%2)*4 .. +3, SWIZZLE_128B
                        const uint32_t off = (uint32_t)(c >> 1) * TILE_BYTES + (uint32_t)rt * 128;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t ch = ((uint32_t)((c & 1) * 4 + q) ^ sw) << 4;
                            sts128(smem_u32(sP) + off + ch, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                            sts128(smem_u32(sdS) + off + ch, dk[4 * q], dk[4 * q + 1], dk[4 * q + 2], dk[4 * q + 3]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_p);
                }
                // dK_j, dV_j: rows = keys of tile j
                mbar_wait(&bar_kv, kvc & 1);
                tc_fence_after();
                {
                    const int key = j * 128 + rt;
                    __nv_bfloat16 *ok = p.dqkv + ((size_t)img * L + key) * 3 * d + d + h * HD;
#pragma unroll
                    for (int m = 0; m < 2; ++m) {                // m = 0: dK, m = 1: dV
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t v[32];
                            tmem_ld32(lane_base + (m ? B_DV : B_DK) + (uint32_t)(c * 32), v);
                            if (key < L) {
#pragma unroll
                                for (int e = 0; e < 32; e += 8) {
                                    uint4 o;
                                    __nv_bfloat162 hh;
                                    hh = __floats2bfloat162_rn(__uint_as_float(v[e]), __uint_as_float(v[e + 1])); o.x = *reinterpret_cast<uint32_t *>(&hh);
                                    hh = __floats2bfloat162_rn(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])); o.y = *reinterpret_cast<uint32_t *>(&hh);
                                    hh = __floats2bfloat162_rn(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])); o.z = *reinterpret_cast<uint32_t *>(&hh);
                                    hh = __floats2bfloat162_rn(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])); o.w = *reinterpret_cast<uint32_t *>(&hh);
                                    *reinterpret_cast<uint4 *>(ok + m * d + c * 32 + e) = o;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_kv_free);
                ++kvc;
            }
            // dQ of both query tiles
            mbar_wait(&bar_q, (uint32_t)i & 1);
            tc_fence_after();
            for (int t = 0; t < MT; ++t) {
                const int row = t * 128 + rt;
                __nv_bfloat16 *oq = p.dqkv + ((size_t)img * L + row) * 3 * d + h * HD;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + B_DQ + (uint32_t)(64 * t + c * 32), v);
                    if (row < L) {
#pragma unroll
                        for (int e = 0; e < 32; e += 8) {
                            uint4 o;
                            __nv_bfloat162 hh;
                            hh = __floats2bfloat162_rn(__uint_as_float(v[e]), __uint_as_float(v[e + 1])); o.x = *reinterpret_cast<uint32_t *>(&hh);
                            hh = __floats2bfloat162_rn(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])); o.y = *reinterpret_cast<uint32_t *>(&hh);
                            hh = __floats2bfloat162_rn(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])); o.z = *reinterpret_cast<uint32_t *>(&hh);
                            hh = __floats2bfloat162_rn(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])); o.w = *reinterpret_cast<uint32_t *>(&hh);
                            *reinterpret_cast<uint4 *>(oq + c * 32 + e) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_q_free);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

}  // namespace

namespace ec {

// Returns EC_OK when the tcgen05 kernel was launched, EC_ERR_UNSUPPORTED when the shape is outside its range.
int attention_tc(const void *qkv, void *out, int n_img, int L, int heads, int causal, cudaStream_t stream, float *lse)
{
    if (L > 384) return EC_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return EC_ERR_CUDA; }
    const int d = heads * HD;
    CUtensorMap map;
    cuuint64_t gdim[3] = {(cuuint64_t)3 * d, (cuuint64_t)L, (cuuint64_t)n_img};
    cuuint64_t gstr[2] = {(cuuint64_t)3 * d * 2, (cuuint64_t)L * 3 * d * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(qkv), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention) failed with CUresult %d", (int)r); return EC_ERR_CUDA; }
    const size_t smem = 2 * STAGE_BYTES + ONES_BYTES + 1024;
    const size_t smem_big = 3 * BIG_TILES * TILE_BYTES + ONES_BYTES + 1024;
    static bool attr_set[64] = {false};
    int dev_id = 0;
    EC_CUDA_CHECK(cudaGetDevice(&dev_id));
    if (dev_id < 64 && !attr_set[dev_id]) {
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc2_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc2_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc2_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc2_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
        attr_set[dev_id] = true;
    }
    // A/B hooks: EC_ATTN_V=1 selects the round-1 kernel (one thread per row), EC_ATTN_STAGGER=0 issues its two tiles side by side
    static const int stagger = [] { const char *e = getenv("EC_ATTN_STAGGER"); return e ? atoi(e) : 1; }();
    static const int version = [] { const char *e = getenv("EC_ATTN_V"); return e ? atoi(e) : 2; }();
    AttnParams p;
    p.out = (__nv_bfloat16 *)out; p.qkv = (const __nv_bfloat16 *)qkv; p.lse = lse; p.L = L; p.heads = heads; p.d = d; p.n_img = n_img;
    p.causal = causal & 1; p.f16 = (causal >> 1) & 1;       // EC_ATTN_CAUSAL | EC_ATTN_F16
    p.stagger = stagger;
    // EC_ATTN_EXACT: 1 = exact row maximum as the softmax reference, 0 = maximum of the first 32 scores (single pass).  Default: exact
    // with fp16 operands (P is stored as fp16: an exponent above 2^16 would be inf), single pass with bf16 (range 2^127, clamped).
    static const int exact_env = [] { const char *e = getenv("EC_ATTN_EXACT"); return e ? atoi(e) : -1; }();
    p.exact = exact_env >= 0 ? exact_env : p.f16;
    p.dbg = nullptr;
    static const int dbg_on = [] { const char *e = getenv("EC_ATTN_DBG"); return e ? atoi(e) : 0; }();
    static long long *dbg_buf = nullptr;
    if (dbg_on) {
        if (!dbg_buf) EC_CUDA_CHECK(cudaMalloc(&dbg_buf, 256 * sizeof(long long)));
        EC_CUDA_CHECK(cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), stream));
        p.dbg = dbg_buf;
    }
    const int units = n_img * heads;
    const int grid = units < sm_count() ? units : sm_count();
    static const int small_on = [] { const char *e = getenv("EC_ATTN_SMALL"); return e ? atoi(e) : 1; }();
    static const int x257 = [] { const char *e = getenv("EC_ATTN_257"); return e ? atoi(e) : 1; }();
    if (L == 257 && !p.causal && x257 && version == 2) {
        static bool attrx[64] = {false};
        if (dev_id < 64 && !attrx[dev_id]) {
            EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc2x_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc2x_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attrx[dev_id] = true;
        }
        if (p.f16) attention_tc2x_kernel<1><<<grid, NTHREADS2X, smem, stream>>>(map, p);
        else attention_tc2x_kernel<0><<<grid, NTHREADS2X, smem, stream>>>(map, p);
    } else if (L > 256) attention_tc_big_kernel<<<grid, BIG_NTHREADS, smem_big, stream>>>(map, p);
    else if (version == 2 && L <= 64 && small_on) {
        // four-chain kernel for small sequences: 64-row boxes
        CUtensorMap map64;
        cuuint32_t box64[3] = {64, 64, 1};
        CUresult r64 = enc(&map64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(qkv), gdim, gstr, box64, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r64 != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention, 64-row boxes) failed with CUresult %d", (int)r64); return EC_ERR_CUDA; }
        const size_t smem4 = (size_t)S4_STAGES * S4_UNIT + 1024;
        static bool attr4[64] = {false};
        if (dev_id < 64 && !attr4[dev_id]) {
            EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            EC_CUDA_CHECK(cudaFuncSetAttribute(attention_tc4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            attr4[dev_id] = true;
        }
        if (p.f16) attention_tc4_kernel<1><<<grid, NTHREADS4, smem4, stream>>>(map64, p);
        else attention_tc4_kernel<0><<<grid, NTHREADS4, smem4, stream>>>(map64, p);
    } else if (version == 2) {
        const bool two = L > 128;
        auto k = two ? (p.f16 ? attention_tc2_kernel<2, 1> : attention_tc2_kernel<2, 0>) : (p.f16 ? attention_tc2_kernel<1, 1> : attention_tc2_kernel<1, 0>);
        k<<<grid, NTHREADS2, smem, stream>>>(map, p);
    } else
        attention_tc_kernel<<<grid, NTHREADS, smem, stream>>>(map, p);
    if (dbg_on && L <= 256 && !(L <= 64 && small_on)) {      // profiling only: per-unit timeline of CTA 0 in SM clocks, relative to the first S issue
        long long h[256];
        EC_CUDA_CHECK(cudaStreamSynchronize(stream));
        EC_CUDA_CHECK(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
        const long long t0 = h[0];
        for (int u = 0; u < 12; ++u)
            for (int t = 0; t < 2; ++t) {
                const long long *m = h + (u * 2 + t) * 2, *w = h + 64 + (u * 2 + t) * 4;
                fprintf(stderr, "unit %2d tile %d: S issue %7lld | S landed %7lld | P written %7lld | PV issue %7lld | O landed %7lld | epilogue done %7lld\n",
                        u, t, m[0] - t0, w[0] - t0, w[1] - t0, m[1] - t0, w[2] - t0, w[3] - t0);
            }
    }
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

int attention_bwd_tc(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int n_img, int L, int heads,
                     cudaStream_t stream)
{
    if (L > 256 || !lse) return EC_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return EC_ERR_CUDA; }
    const int d = heads * HD;
    CUtensorMap mq, md;
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    {
        cuuint64_t gdim[3] = {(cuuint64_t)3 * d, (cuuint64_t)L, (cuuint64_t)n_img};
        cuuint64_t gstr[2] = {(cuuint64_t)3 * d * 2, (cuuint64_t)L * 3 * d * 2};
        CUresult r = enc(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(qkv), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention bwd qkv) failed with CUresult %d", (int)r); return EC_ERR_CUDA; }
    }
    {
        cuuint64_t gdim[3] = {(cuuint64_t)d, (cuuint64_t)L, (cuuint64_t)n_img};
        cuuint64_t gstr[2] = {(cuuint64_t)d * 2, (cuuint64_t)L * d * 2};
        CUresult r = enc(&md, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(d_o), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention bwd dO) failed with CUresult %d", (int)r); return EC_ERR_CUDA; }
    }
    const size_t smem = BWD_SMEM + 1024;
    static bool attr_set[64] = {false};
    int dev_id = 0;
    EC_CUDA_CHECK(cudaGetDevice(&dev_id));
    if (dev_id < 64 && !attr_set[dev_id]) {
        EC_CUDA_CHECK(cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev_id] = true;
    }
    AttnBwdParams p;
    p.dqkv = (__nv_bfloat16 *)dqkv; p.o = (const __nv_bfloat16 *)o; p.lse = lse; p.L = L; p.heads = heads; p.d = d; p.n_img = n_img;
    const int units = n_img * heads;
    const int grid = units < sm_count() ? units : sm_count();
    attention_bwd_tc_kernel<<<grid, BWD_THREADS, smem, stream>>>(mq, md, p);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

}  // namespace ec
