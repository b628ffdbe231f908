// head.cu -- classifier heads: text-cosine logits + per-sample aggregation, and the fp32 few-shot adapter pieces.
//
// Reference semantics (paths relative to the reference root):
//   zero-shot head      models/clip_cls.py:144-154   logits = scale * feats @ text.T (feats NOT normalised)
//   few-shot / FT head  models/clip_cls.py:326-342, models/clip_cls_ft.py:232-248  (L2-normalise, mask, logits)
//   _aggregate_logits   models/clip_cls.py:104-121   sum | mean over valid | max with -1e6 on invalid
//   _aggregate_probs    models/clip_cls.py:123-129   softmax per view, masked mean
//   top-1 / top-5       test.py:66-81
//   adapter             models/adapter.py:82-105 (nn.TransformerEncoderLayer, norm_first, ReLU FFN, key padding mask)
// All of it is fp32 SIMT: < 0.1 % of the FLOPs and it decides top-1, so it is kept in the reference's precision.
#include "common.cuh"

namespace {

constexpr int HEAD_THREADS = 256;

// block-wide (value, index) arg-max; ties resolve to the lowest index (torch.argmax convention)
__device__ void block_argmax(float v, int idx, float *sv, int *si, float &best, int &besti)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    __syncthreads();
    if (lane == 0) { sv[warp] = v; si[warp] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = sv[0];
        int bi = si[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (sv[w] > b || (sv[w] == b && si[w] < bi)) { b = sv[w]; bi = si[w]; }
        sv[0] = b; si[0] = bi;
    }
    __syncthreads();
    best = sv[0]; besti = si[0];
}

__device__ float block_reduce(float v, float *sv, bool is_max)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? ec::warp_max(v) : ec::warp_sum(v);
    __syncthreads();
    if (lane == 0) sv[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = sv[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) b = is_max ? fmaxf(b, sv[w]) : b + sv[w];
        sv[0] = b;
    }
    __syncthreads();
    return sv[0];
}

// logits of class row k for all T views of sample b: <scale * f_t, text_k>, TT views per pass over the text row
template <int TT, int MODE>
__device__ __forceinline__ void head_dots(const float *sf, const float *__restrict__ tx, const float *vmask, int T, int C, int n_cls,
                                          int b, int k, int lane, float *out_full, float *sl)
{
    for (int t0 = 0; t0 < T; t0 += TT) {
        float acc[TT];
#pragma unroll
        for (int j = 0; j < TT; ++j) acc[j] = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float w = __ldg(tx + c);
#pragma unroll
            for (int j = 0; j < TT; ++j)
                if (t0 + j < T) acc[j] = fmaf(sf[(size_t)(t0 + j) * C + c], w, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < TT; ++j)
            if (t0 + j < T) {
                const int t = t0 + j;
                const float v = ec::warp_sum(acc[j]);
                if (lane == 0) {
                    if (MODE == 1) out_full[((size_t)b * T + t) * n_cls + k] = vmask[t] != 0.f ? v : 0.f;
                    else sl[(size_t)t * n_cls + k] = vmask[t] != 0.f ? v : 0.f;
                }
            }
    }
}

// One CTA per sample.  smem: feats [T][C] | logits [T][n_cls] | agg [n_cls] | probs [n_cls]
// MODE 0: everything in one launch.  With many classes (N-ImageNet: 1000) one CTA per sample walks 125 text rows per warp with
// little memory-level parallelism (ncu: 1.1 ms for 64 samples), so the work is split: MODE 1, grid (B, S): logits of a slice of
// the classes -> out_full;  MODE 2, grid B: reads out_full back and does the aggregation, the probabilities and top-5.
template <int MODE>
__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const float *__restrict__ feats, const uint8_t *__restrict__ valid,
                                                            const float *__restrict__ text, int T, int C, int n_cls,
                                                            float scale, int normalize, int agg, float *out_full,
                                                            float *out_logits, float *out_probs, int32_t *out_top)
{
    extern __shared__ __align__(16) float sm[];
    float *sf = sm;
    float *sl = sf + (size_t)T * C;
    float *sagg = sl + (size_t)T * n_cls;
    float *sprob = sagg + n_cls;
    __shared__ float sv[32];
    __shared__ int si[32];
    __shared__ float vmask[16];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;

    if (tid < T) vmask[tid] = valid[(size_t)b * T + tid] ? 1.f : 0.f;
    __syncthreads();
    float nvalid = 0.f;
    for (int t = 0; t < T; ++t) nvalid += vmask[t];

    // features of the T views: optional L2 normalisation (F.normalize eps 1e-12), mask, then * scale
    if (MODE != 2)
    for (int t = warp; t < T; t += nwarp) {
        const float *f = feats + ((size_t)b * T + t) * C;
        float mul = 0.f;
        if (vmask[t] != 0.f) {
            mul = 1.f;
            if (normalize) {
                float ss = 0.f;
                for (int c = lane; c < C; c += 32) { const float x = f[c]; ss += x * x; }
                ss = ec::warp_sum(ss);
                mul = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            }
        }
        for (int c = lane; c < C; c += 32) sf[(size_t)t * C + c] = vmask[t] != 0.f ? (f[c] * mul) * scale : 0.f;
    }
    __syncthreads();

    // logits[t][k] = <scale * f_t, text_k>; one warp per class row, all views at once
    const int k_per = MODE == 1 ? (n_cls + (int)gridDim.y - 1) / (int)gridDim.y : n_cls;
    const int k_lo = MODE == 1 ? (int)blockIdx.y * k_per : 0, k_hi = min(n_cls, k_lo + k_per);
    if (MODE != 2)
    for (int k = k_lo + warp; k < k_hi; k += nwarp) {
        const float *tx = text + (size_t)k * C;
        // TT views per pass over the text row (register tile): the old fixed 16-view tile issued 16 predicated LDS + FMA pairs per
        // column whatever T was -- 8 x the useful work at T = 2, 3 x at T = 5
        if (T == 1) head_dots<1, MODE>(sf, tx, vmask, T, C, n_cls, b, k, lane, out_full, sl);
        else if (T == 2) head_dots<2, MODE>(sf, tx, vmask, T, C, n_cls, b, k, lane, out_full, sl);
        else head_dots<4, MODE>(sf, tx, vmask, T, C, n_cls, b, k, lane, out_full, sl);
    }
    if (MODE == 1) return;
    if (MODE == 2)
        for (int i = tid; i < T * n_cls; i += blockDim.x) sl[i] = out_full[(size_t)b * T * n_cls + i];
    __syncthreads();

    // aggregated logits
    for (int k = tid; k < n_cls; k += blockDim.x) {
        float a;
        if (agg == EC_AGG_MAX) {
            a = -INFINITY;
            for (int t = 0; t < T; ++t) a = fmaxf(a, sl[(size_t)t * n_cls + k] - (1.f - vmask[t]) * 1e6f);
        } else {
            a = 0.f;
            for (int t = 0; t < T; ++t) a += sl[(size_t)t * n_cls + k];
            if (agg == EC_AGG_MEAN) a = a / nvalid;
        }
        sagg[k] = a;
        sprob[k] = 0.f;
        if (out_logits) out_logits[(size_t)b * n_cls + k] = a;
        if (MODE == 0 && out_full)
            for (int t = 0; t < T; ++t) out_full[((size_t)b * T + t) * n_cls + k] = sl[(size_t)t * n_cls + k];
    }
    __syncthreads();

    // probs = mean over valid views of softmax(logits_t)
    for (int t = 0; t < T; ++t) {
        if (vmask[t] == 0.f) continue;   // uniform across the block
        float m = -INFINITY;
        for (int k = tid; k < n_cls; k += blockDim.x) m = fmaxf(m, sl[(size_t)t * n_cls + k]);
        m = block_reduce(m, sv, true);
        float s = 0.f;
        for (int k = tid; k < n_cls; k += blockDim.x) s += expf(sl[(size_t)t * n_cls + k] - m);
        s = block_reduce(s, sv, false);
        for (int k = tid; k < n_cls; k += blockDim.x) sprob[k] += expf(sl[(size_t)t * n_cls + k] - m) / s;
    }
    __syncthreads();
    for (int k = tid; k < n_cls; k += blockDim.x) {
        const float pr = sprob[k] / nvalid;
        sprob[k] = pr;
        if (out_probs) out_probs[(size_t)b * n_cls + k] = pr;
    }
    __syncthreads();

    // top-5 of logits and probs (repeated arg-max with exclusion)
    if (out_top) {
        for (int which = 0; which < 2; ++which) {
            float *src = which == 0 ? sagg : sprob;
            for (int r = 0; r < 5; ++r) {
                float v = -INFINITY;
                int idx = 0x7fffffff;
                for (int k = tid; k < n_cls; k += blockDim.x)
                    if (src[k] > v) { v = src[k]; idx = k; }
                float bv;
                int bi;
                block_argmax(v, idx, sv, si, bv, bi);
                if (tid == 0) {
                    out_top[((size_t)b * 2 + which) * 5 + r] = (r < n_cls && bi != 0x7fffffff) ? bi : -1;
                    if (bi != 0x7fffffff) src[bi] = -INFINITY;
                }
                __syncthreads();
            }
        }
    }
}

// ---- fp32 GEMM for a few hundred rows (the few-shot adapter: 128 x 512 activations): lane = output column, eight rows per warp ----
// out = act(A W^T + bias) (+ res).  A CTA of four warps owns 32 rows x 32 columns and walks K in chunks of 128: the chunk of A
// ([32][128], read back as broadcast 16-byte words) and of W ([32 columns][128], pitch 132 floats: the 16-byte reads of the eight lanes
// of a quarter warp fall on 32 distinct banks) are staged in shared memory with coalesced 16-byte loads; the K sum of an output is one
// sequential chain in a single thread -- no shuffles, a fixed summation order (deterministic), 8 independent chains per lane.
// The lane-split kernel below reduces every output with two 5-step shuffle trees (36 us for 128 x 1536 x 512).
constexpr int RT_KC = 128, RT_WP = RT_KC + 4;
__global__ void __launch_bounds__(128) sgemm_rows_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                         const float *__restrict__ bias, const float *__restrict__ res, int M,
                                                         int N, int K, int act, float *__restrict__ out)
{
    __shared__ __align__(16) float sa[32][RT_KC];
    __shared__ __align__(16) float sw[32][RT_WP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int kc = 0; kc < K; kc += RT_KC) {
        const int kn = min(RT_KC, K - kc);                 // a multiple of 4
        const int q4 = kn >> 2;
        for (int i = tid; i < 32 * q4; i += 128) {
            const int row = i / q4, c4 = i - row * q4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vw = va;
            if (m0 + row < M) va = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + row) * K + kc + 4 * c4);
            if (n0 + row < N) vw = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(n0 + row) * K + kc + 4 * c4));
            *reinterpret_cast<float4 *>(&sa[row][4 * c4]) = va;
            *reinterpret_cast<float4 *>(&sw[row][4 * c4]) = vw;
        }
        __syncthreads();
#pragma unroll 4
        for (int k4 = 0; k4 < q4; ++k4) {
            const float4 w = *reinterpret_cast<const float4 *>(&sw[lane][4 * k4]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 a = *reinterpret_cast<const float4 *>(&sa[8 * warp + r][4 * k4]);
                acc[r] = fmaf(a.w, w.w, fmaf(a.z, w.z, fmaf(a.y, w.y, fmaf(a.x, w.x, acc[r]))));
            }
        }
        __syncthreads();
    }
    const int n = n0 + lane;
    if (n < N) {
        const float bv = bias ? bias[n] : 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int m = m0 + 8 * warp + r;
            if (m < M) {
                float v = acc[r] + bv;
                if (act == 1) v = fmaxf(v, 0.f);
                if (res) v += res[(size_t)m * N + n];
                out[(size_t)m * N + n] = v;
            }
        }
    }
}

// ---- fp32 SGEMM for the adapter: out = act(A W^T + bias) (+ res); 64x64 tile, 16-deep, 256 threads, 4x4 per thread ----
__global__ void __launch_bounds__(256) sgemm_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                    const float *__restrict__ bias, const float *__restrict__ res, int M,
                                                    int N, int K, int act, float *__restrict__ out)
{
    __shared__ float sa[16][64 + 4];
    __shared__ float sw[16][64 + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = tid; i < 64 * 16; i += 256) {
            const int r = i >> 4, c = i & 15;
            sa[c][r] = (m0 + r < M && k0 + c < K) ? A[(size_t)(m0 + r) * K + k0 + c] : 0.f;
            sw[c][r] = (n0 + r < N && k0 + c < K) ? W[(size_t)(n0 + r) * K + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sa[c][ty * 4 + i]; w[i] = sw[c][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) {
                float v = acc[i][j] + (bias ? bias[n] : 0.f);
                if (act == 1) v = fmaxf(v, 0.f);
                if (res) v += res[(size_t)m * N + n];
                out[(size_t)m * N + n] = v;
            }
        }
}

// Skinny variant for the adapter's shapes (M = B*T <= a few hundred rows, K a multiple of 32 up to 1024): the 64 x 64 tile kernel
// above runs 8-32 CTAs with a load -> barrier -> 16 k-steps -> barrier loop and is latency-bound (ncu: 65 us per launch, 10 launches
// per few-shot step).  Here a CTA owns NC output columns, keeps their W rows in shared memory, and its warps stream the rows of A
// two at a time through registers; each dot product is reduced over the lanes in a fixed order (deterministic).
template <int NC>
__global__ void __launch_bounds__(256) sgemm_skinny_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                           const float *__restrict__ bias, const float *__restrict__ res, int M,
                                                           int N, int K, int act, float *__restrict__ out)
{
    extern __shared__ __align__(16) float sw[];        // [NC][K]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * NC;
    const int KJ = K >> 5;                             // values per lane (K % 32 == 0, K <= 1024)
    for (int i = tid; i < NC * K; i += 256) {
        const int c = i / K, k = i - c * K;
        sw[i] = n0 + c < N ? W[(size_t)(n0 + c) * K + k] : 0.f;
    }
    __syncthreads();
    for (int m = 2 * warp; m < M; m += 16) {
        const bool two = m + 1 < M;
        float a0[32], a1[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            a0[j] = 0.f; a1[j] = 0.f;
            if (j < KJ) {
                a0[j] = A[(size_t)m * K + lane + 32 * j];
                if (two) a1[j] = A[(size_t)(m + 1) * K + lane + 32 * j];
            }
        }
        float r0 = 0.f, r1 = 0.f;                      // lane c keeps column n0 + c
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < KJ) {
                    const float w = sw[c * K + lane + 32 * j];
                    s0 = fmaf(a0[j], w, s0);
                    s1 = fmaf(a1[j], w, s1);
                }
            s0 = ec::warp_sum(s0);
            s1 = ec::warp_sum(s1);
            if (lane == c) { r0 = s0; r1 = s1; }
        }
        if (lane < NC && n0 + lane < N) {
            const int n = n0 + lane;
            float v = r0 + (bias ? bias[n] : 0.f);
            if (act == 1) v = fmaxf(v, 0.f);
            if (res) v += res[(size_t)m * N + n];
            out[(size_t)m * N + n] = v;
            if (two) {
                float u = r1 + (bias ? bias[n] : 0.f);
                if (act == 1) u = fmaxf(u, 0.f);
                if (res) u += res[(size_t)(m + 1) * N + n];
                out[(size_t)(m + 1) * N + n] = u;
            }
        }
    }
}

// One CTA per sample, one warp per head; T <= 16 views.
__global__ void adapter_attention_kernel(const float *__restrict__ qkv, const uint8_t *__restrict__ valid, int T, int D,
                                         int heads, float *__restrict__ out)
{
    const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (h >= heads) return;
    const int hd = D / heads;
    const float sc = rsqrtf((float)hd);
    const float *base = qkv + (size_t)b * T * 3 * D + h * hd;
    for (int t = 0; t < T; ++t) {
        float s[16];
        float m = -INFINITY;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            s[u] = -INFINITY;
            if (u < T) {
                float acc = 0.f;
                for (int c = lane; c < hd; c += 32) acc = fmaf(base[(size_t)t * 3 * D + c], base[(size_t)u * 3 * D + D + c], acc);
                acc = ec::warp_sum(acc) * sc;
                if (valid[(size_t)b * T + u]) { s[u] = acc; m = fmaxf(m, acc); }
            }
        }
        float den = 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            s[u] = (u < T && s[u] != -INFINITY) ? expf(s[u] - m) : 0.f;
            den += s[u];
        }
        for (int c = lane; c < hd; c += 32) {
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (u < T) acc = fmaf(s[u], base[(size_t)u * 3 * D + 2 * D + c], acc);
            out[((size_t)b * T + t) * D + h * hd + c] = acc / den;
        }
    }
}

// Backward of adapter_attention_kernel: one CTA per sample, one warp per head, T <= 16 views.  The probabilities are recomputed
// (T x T per head); dV = P^T dO, dP = dO V^T, dS = P (dP - rowsum(P dP)), dQ = dS K / sqrt(hd), dK = dS^T Q / sqrt(hd).
// Masked keys have P = 0, so they receive no dK / dV; every query row (valid or padded) is differentiated like the reference's
// nn.TransformerEncoder does.
__global__ void adapter_attention_bwd_kernel(const float *__restrict__ qkv, const uint8_t *__restrict__ valid,
                                             const float *__restrict__ d_out, int T, int D, int heads, float *__restrict__ d_qkv)
{
    extern __shared__ float s_att[];             // [heads][2][16][16]: probabilities and dS of every head
    const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (h >= heads) return;
    const int hd = D / heads;
    const float sc = rsqrtf((float)hd);
    const float *base = qkv + (size_t)b * T * 3 * D + h * hd;
    const float *dob = d_out + (size_t)b * T * D + h * hd;
    float *dqb = d_qkv + (size_t)b * T * 3 * D + h * hd;
    float (*P)[16] = reinterpret_cast<float (*)[16]>(s_att + (size_t)h * 512);
    float (*dS)[16] = reinterpret_cast<float (*)[16]>(s_att + (size_t)h * 512 + 256);
    // probabilities and dP
    for (int t = 0; t < T; ++t) {
        float s[16], dp[16];
        float m = -INFINITY;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            s[u] = -INFINITY; dp[u] = 0.f;
            if (u < T) {
                float acc = 0.f, acc2 = 0.f;
                for (int c = lane; c < hd; c += 32) {
                    acc = fmaf(base[(size_t)t * 3 * D + c], base[(size_t)u * 3 * D + D + c], acc);
                    acc2 = fmaf(dob[(size_t)t * D + c], base[(size_t)u * 3 * D + 2 * D + c], acc2);
                }
                acc = ec::warp_sum(acc) * sc;
                dp[u] = ec::warp_sum(acc2);
                if (valid[(size_t)b * T + u]) { s[u] = acc; m = fmaxf(m, acc); }
            }
        }
        float den = 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            s[u] = (u < T && s[u] != -INFINITY) ? expf(s[u] - m) : 0.f;
            den += s[u];
        }
        float dot = 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) { s[u] /= den; dot = fmaf(s[u], dp[u], dot); }
        if (lane < 16) {
            // every lane holds the full row; lane u keeps column u
            float pu = 0.f, du = 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (u == lane) { pu = s[u]; du = s[u] * (dp[u] - dot); }
            P[t][lane] = pu;
            dS[t][lane] = du;
        }
    }
    __syncwarp();
    for (int c = lane; c < hd; c += 32) {
        for (int t = 0; t < T; ++t) {          // dQ[t] = sc * sum_u dS[t][u] K[u]
            float acc = 0.f;
            for (int u = 0; u < T; ++u) acc = fmaf(dS[t][u], base[(size_t)u * 3 * D + D + c], acc);
            dqb[(size_t)t * 3 * D + c] = acc * sc;
        }
        for (int u = 0; u < T; ++u) {          // dK[u] = sc * sum_t dS[t][u] Q[t];  dV[u] = sum_t P[t][u] dO[t]
            float ak = 0.f, av = 0.f;
            for (int t = 0; t < T; ++t) {
                ak = fmaf(dS[t][u], base[(size_t)t * 3 * D + c], ak);
                av = fmaf(P[t][u], dob[(size_t)t * D + c], av);
            }
            dqb[(size_t)u * 3 * D + D + c] = ak * sc;
            dqb[(size_t)u * 3 * D + 2 * D + c] = av;
        }
    }
}

__global__ void relu_bwd_kernel(const float *__restrict__ y, const float *__restrict__ dy, float *__restrict__ dx, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

}  // namespace

extern "C" int ec_adapter_attention_bwd(const float *qkv, const uint8_t *valid, const float *d_out, int B, int T, int D, int heads,
                                        float *d_qkv, void *stream)
{
    EC_REQUIRE(qkv && valid && d_out && d_qkv, "ec_adapter_attention_bwd: null pointer");
    EC_REQUIRE(B > 0 && T > 0 && T <= 16 && heads > 0 && heads <= 32 && D % heads == 0,
               "ec_adapter_attention_bwd: bad shape B=%d T=%d D=%d heads=%d", B, T, D, heads);
    EC_REQUIRE(heads <= 16, "ec_adapter_attention_bwd: at most 16 heads (got %d)", heads);
    adapter_attention_bwd_kernel<<<B, heads * 32, (size_t)heads * 2048, (cudaStream_t)stream>>>(qkv, valid, d_out, T, D, heads, d_qkv);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

/* dx = dy where y > 0 else 0: backward of the ReLU fused into ec_gemm_f32 (act = 1); y is the activation's OUTPUT */
extern "C" int ec_relu_bwd(const float *y, const float *dy, float *dx, int64_t n, void *stream)
{
    EC_REQUIRE(y && dy && dx && n >= 0, "ec_relu_bwd: bad arguments");
    if (n == 0) return EC_OK;
    relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_head(const float *feats, const uint8_t *valid, const float *text, int B, int T, int C, int n_cls,
                       float scale, int normalize, int agg, float *out_full, float *out_logits, float *out_probs,
                       int32_t *out_top, void *stream)
{
    EC_REQUIRE(feats && valid && text, "ec_head: null pointer");
    EC_REQUIRE(B > 0 && T > 0 && T <= 16 && C > 0 && n_cls > 0, "ec_head: bad shape B=%d T=%d C=%d n_cls=%d (T <= 16)", B, T, C,
               n_cls);
    EC_REQUIRE(agg >= EC_AGG_SUM && agg <= EC_AGG_MAX, "ec_head: bad aggregation %d", agg);
    const size_t smem = ((size_t)T * C + (size_t)T * n_cls + 2 * (size_t)n_cls) * sizeof(float);
    EC_REQUIRE(smem <= 220 * 1024, "ec_head: T*C + T*n_cls too large for shared memory (%zu bytes)", smem);
    // one CTA per sample is latency- and issue-bound as soon as a sample has a few thousand (view, class) pairs (C1: 32 samples x 5 views x
    // 101 classes took 166 us on 32 SMs): spread the logits over the GPU unless the batch alone fills it
    const bool split = out_full != nullptr && n_cls >= 32;
    static size_t attr[3] = {0, 0, 0};
    auto grant = [&](int mode, const void *fn) -> int {
        if (smem > 48 * 1024 && smem > attr[mode]) {
            EC_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr[mode] = smem;
        }
        return EC_OK;
    };
    if (split) {
        int rc = grant(1, (const void *)head_kernel<1>);
        if (rc != EC_OK) return rc;
        rc = grant(2, (const void *)head_kernel<2>);
        if (rc != EC_OK) return rc;
        int S = (4 * ec::sm_count() + B - 1) / B;                    // about four CTAs per SM in the logits launch
        S = S < 1 ? 1 : (S > n_cls / 8 ? n_cls / 8 : S);               // at least one class row per warp of a slice
        head_kernel<1><<<dim3(B, S), HEAD_THREADS, smem, (cudaStream_t)stream>>>(feats, valid, text, T, C, n_cls, scale, normalize, agg,
                                                                                out_full, out_logits, out_probs, out_top);
        head_kernel<2><<<B, HEAD_THREADS, smem, (cudaStream_t)stream>>>(feats, valid, text, T, C, n_cls, scale, normalize, agg, out_full,
                                                                        out_logits, out_probs, out_top);
    } else {
        const int rc = grant(0, (const void *)head_kernel<0>);
        if (rc != EC_OK) return rc;
        head_kernel<0><<<B, HEAD_THREADS, smem, (cudaStream_t)stream>>>(feats, valid, text, T, C, n_cls, scale, normalize, agg, out_full,
                                                                        out_logits, out_probs, out_top);
    }
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_gemm_f32(const float *A, const float *W, const float *bias, const float *res, int M, int N, int K, int act,
                           float *out, void *stream)
{
    EC_REQUIRE(A && W && out && M > 0 && N > 0 && K > 0, "ec_gemm_f32: bad arguments");
    EC_REQUIRE(act == 0 || act == 1, "ec_gemm_f32: bad activation %d", act);
    static const int rows_on = [] { const char *e = getenv("EC_SGEMM_ROWS"); return e ? atoi(e) : 1; }();
    if (rows_on && M >= 16 && M <= 4096 && K % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
        sgemm_rows_kernel<<<dim3((N + 31) / 32, (M + 31) / 32), 128, 0, (cudaStream_t)stream>>>(A, W, bias, res, M, N, K, act, out);
    } else if (M <= 1024 && K % 32 == 0 && K <= 1024) {
        constexpr int NC = 8;
        const size_t smem = (size_t)NC * K * sizeof(float);          // <= 32 KB
        sgemm_skinny_kernel<NC><<<(N + NC - 1) / NC, 256, smem, (cudaStream_t)stream>>>(A, W, bias, res, M, N, K, act, out);
    } else {
        sgemm_kernel<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, (cudaStream_t)stream>>>(A, W, bias, res, M, N, K, act, out);
    }
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_adapter_attention(const float *qkv, const uint8_t *valid, int B, int T, int D, int heads, float *out,
                                    void *stream)
{
    EC_REQUIRE(qkv && valid && out, "ec_adapter_attention: null pointer");
    EC_REQUIRE(B > 0 && T > 0 && T <= 16 && heads > 0 && heads <= 32 && D % heads == 0,
               "ec_adapter_attention: bad shape B=%d T=%d D=%d heads=%d", B, T, D, heads);
    adapter_attention_kernel<<<B, heads * 32, 0, (cudaStream_t)stream>>>(qkv, valid, T, D, heads, out);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}
