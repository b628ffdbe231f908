// elementwise.cu -- memory-bound glue kernels of the encoder / heads: LayerNorm, class-token rows, im2col,
// dtype conversion, LoRA weight merge, adapter residual blend.
//
// Reference semantics (paths relative to the reference root / openai-CLIP `clip/model.py` [3P]):
//   LayerNorm         fp32 statistics, eps 1e-5 (CLIP's LayerNorm subclass upcasts to fp32)
//   class token       x = cat([class_embedding, patches]) + positional_embedding
//   LoRA merge        models/lora.py:138-149 (q/k/v) and 49-52 (out_proj): W + up @ down, no alpha/r scaling
//   residual blend    models/adapter.py:22-25: in*r + new*(1-r)
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// One warp per row; two-pass (mean, then centred variance) on register-cached values.  TIn = float or __half (the fp16
// residual stream); outputs: bf16 (GEMM operand), fp32, and / or fp16 (ln_pre starting an fp16 residual stream).
__device__ __forceinline__ float4 ld4(const float *p, int j) { return reinterpret_cast<const float4 *>(p)[j]; }
__device__ __forceinline__ float4 ld4(const __half *p, int j)
{
    const uint2 u = reinterpret_cast<const uint2 *>(p)[j];
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

template <bool CACHE, typename TIn>
__global__ void __launch_bounds__(256) layernorm_kernel(const TIn *__restrict__ x, int64_t row_stride,
                                                        const float *__restrict__ gamma, const float *__restrict__ beta,
                                                        int M, int d, __nv_bfloat16 *__restrict__ out_bf16,
                                                        float *__restrict__ out_f32, __half *__restrict__ out_f16,
                                                        __half *__restrict__ out_lo = nullptr)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const TIn *xr = x + (size_t)row * row_stride;
    const int n4 = d >> 2;
    constexpr int MAXV = 8;   // d <= 1024 when CACHE
    float4 v[MAXV];
    float s = 0.f;
    if (CACHE) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = lane + i * 32;
            if (j < n4) { v[i] = ld4(xr, j); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
        }
    } else {
        for (int j = lane; j < n4; j += 32) { const float4 t = ld4(xr, j); s += (t.x + t.y) + (t.z + t.w); }
    }
    const float mean = ec::warp_sum(s) / (float)d;
    float q = 0.f;
    if (CACHE) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = lane + i * 32;
            if (j < n4) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                q += (a * a + b * b) + (c * c + e * e);
            }
        }
    } else {
        for (int j = lane; j < n4; j += 32) {
            const float4 t = ld4(xr, j);
            const float a = t.x - mean, b = t.y - mean, c = t.z - mean, e = t.w - mean;
            q += (a * a + b * b) + (c * c + e * e);
        }
    }
    const float rstd = rsqrtf(ec::warp_sum(q) / (float)d + 1e-5f);
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma);
    const float4 *b4 = reinterpret_cast<const float4 *>(beta);
    auto emit = [&](int j, const float4 &t) {
        const float4 g = g4[j], b = b4[j];
        float4 o;
        o.x = (t.x - mean) * rstd * g.x + b.x;
        o.y = (t.y - mean) * rstd * g.y + b.y;
        o.z = (t.z - mean) * rstd * g.z + b.z;
        o.w = (t.w - mean) * rstd * g.w + b.w;
        if (out_f32) reinterpret_cast<float4 *>(out_f32 + (size_t)row * d)[j] = o;
        if (out_bf16) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t *>(&h0);
            u.y = *reinterpret_cast<uint32_t *>(&h1);
            reinterpret_cast<uint2 *>(out_bf16 + (size_t)row * d)[j] = u;
        }
        if (out_f16) {
            __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t *>(&h0);
            u.y = *reinterpret_cast<uint32_t *>(&h1);
            reinterpret_cast<uint2 *>(out_f16 + (size_t)row * d)[j] = u;
            if (out_lo) {      // residue of the fp16 rounding: the (hi, lo) residual stream of ec_gemm_bf16_stats2
                const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
                __half2 l0 = __floats2half2_rn(o.x - b0.x, o.y - b0.y), l1 = __floats2half2_rn(o.z - b1.x, o.w - b1.y);
                uint2 ul;
                ul.x = *reinterpret_cast<uint32_t *>(&l0);
                ul.y = *reinterpret_cast<uint32_t *>(&l1);
                reinterpret_cast<uint2 *>(out_lo + (size_t)row * d)[j] = ul;
            }
        }
    };
    if (CACHE) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = lane + i * 32;
            if (j < n4) emit(j, v[i]);
        }
    } else {
        for (int j = lane; j < n4; j += 32) emit(j, ld4(xr, j));
    }
}

// fp16 rows with 16-byte accesses: a lane owns 8 consecutive elements per step (d % 8 == 0, d <= 1024, 16-byte aligned rows).
// A warp normalises TWO rows at a time and keeps them packed in registers: all loads of both rows are issued before the
// first reduction, which doubles the bytes in flight per SM (the kernel is latency-, not bandwidth-limited otherwise).
__global__ void __launch_bounds__(256) layernorm_h8_kernel(const __half *__restrict__ x, int64_t row_stride,
                                                           const float *__restrict__ gamma, const float *__restrict__ beta,
                                                           int M, int d, __nv_bfloat16 *__restrict__ out_bf16,
                                                           float *__restrict__ out_f32, __half *__restrict__ out_f16)
{
    constexpr int R = 2, MAXV = 4;
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
    const int lane = threadIdx.x & 31;
    if (row0 >= M) return;
    const int n8 = d >> 3;
    uint4 raw[R][MAXV];
    auto unpack = [](const uint4 &u, float (&f)[8]) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2 *>(&w[e]));
            f[2 * e] = t.x; f[2 * e + 1] = t.y;
        }
    };
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint4 *xr = reinterpret_cast<const uint4 *>(x + (size_t)min(row0 + r, M - 1) * row_stride);
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (lane + i * 32 < n8) raw[r][i] = xr[lane + i * 32];
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            if (lane + i * 32 < n8) {
                float f[8];
                unpack(raw[r][i], f);
                s += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
            }
        }
        mean[r] = s;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = ec::warp_sum(mean[r]) / (float)d;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            if (lane + i * 32 < n8) {
                float f[8];
                unpack(raw[r][i], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) { const float a = f[e] - mean[r]; q += a * a; }
            }
        }
        rstd[r] = q;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(ec::warp_sum(rstd[r]) / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int j = lane + i * 32;
        if (j >= n8) continue;
        const float4 g0 = reinterpret_cast<const float4 *>(gamma)[2 * j], g1 = reinterpret_cast<const float4 *>(gamma)[2 * j + 1];
        const float4 b0 = reinterpret_cast<const float4 *>(beta)[2 * j], b1 = reinterpret_cast<const float4 *>(beta)[2 * j + 1];
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = row0 + r;
            if (row >= M) continue;
            float f[8], o[8];
            unpack(raw[r][i], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (f[e] - mean[r]) * rstd[r] * g[e] + b[e];
            if (out_f32) {
                float4 *of = reinterpret_cast<float4 *>(out_f32 + (size_t)row * d) + 2 * j;
                of[0] = make_float4(o[0], o[1], o[2], o[3]);
                of[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
            if (out_bf16) {
                uint4 u;
                __nv_bfloat162 h;
                h = __floats2bfloat162_rn(o[0], o[1]); u.x = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2bfloat162_rn(o[2], o[3]); u.y = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2bfloat162_rn(o[4], o[5]); u.z = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2bfloat162_rn(o[6], o[7]); u.w = *reinterpret_cast<uint32_t *>(&h);
                reinterpret_cast<uint4 *>(out_bf16 + (size_t)row * d)[j] = u;
            }
            if (out_f16) {
                uint4 u;
                __half2 h;
                h = __floats2half2_rn(o[0], o[1]); u.x = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2half2_rn(o[2], o[3]); u.y = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2half2_rn(o[4], o[5]); u.z = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2half2_rn(o[6], o[7]); u.w = *reinterpret_cast<uint32_t *>(&h);
                reinterpret_cast<uint4 *>(out_f16 + (size_t)row * d)[j] = u;
            }
        }
    }
}

template <typename TIn>
int launch_ln_t(const TIn *x, int64_t row_stride, const float *gamma, const float *beta, int M, int d, void *out_bf16,
                float *out_f32, void *out_f16, cudaStream_t stream)
{
    EC_REQUIRE(x && gamma && beta && (out_bf16 || out_f32 || out_f16), "layernorm: null pointer");
    EC_REQUIRE(M > 0 && d > 0 && d % 4 == 0 && row_stride % 4 == 0, "layernorm: d and row stride must be multiples of 4");
    const int rows_per_block = 8;
    const unsigned grid = (unsigned)((M + rows_per_block - 1) / rows_per_block);
    if (sizeof(TIn) == 2 && d <= 1024 && d % 8 == 0 && row_stride % 8 == 0 && ((uintptr_t)x & 15) == 0 &&
        (((uintptr_t)out_bf16 | (uintptr_t)out_f32 | (uintptr_t)out_f16) & 15) == 0) {
        layernorm_h8_kernel<<<(grid + 1) / 2, 256, 0, stream>>>((const __half *)x, row_stride, gamma, beta, M, d, (__nv_bfloat16 *)out_bf16,
                                                      out_f32, (__half *)out_f16);
        EC_CUDA_CHECK(cudaGetLastError());
        return EC_OK;
    }
    if (d <= 1024)
        layernorm_kernel<true, TIn><<<grid, 256, 0, stream>>>(x, row_stride, gamma, beta, M, d, (__nv_bfloat16 *)out_bf16, out_f32,
                                                              (__half *)out_f16);
    else
        layernorm_kernel<false, TIn><<<grid, 256, 0, stream>>>(x, row_stride, gamma, beta, M, d, (__nv_bfloat16 *)out_bf16, out_f32,
                                                               (__half *)out_f16);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

int launch_ln(const float *x, int64_t row_stride, const float *gamma, const float *beta, int M, int d, void *out_bf16,
              float *out_f32, cudaStream_t stream)
{
    return launch_ln_t<float>(x, row_stride, gamma, beta, M, d, out_bf16, out_f32, nullptr, stream);
}

__global__ void cls_rows_kernel(float *x, const float *cls, const float *pos, int n_img, int L, int d)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_img * d) return;
    const int img = (int)(i / d), c = (int)(i % d);
    x[(size_t)img * L * d + c] = cls[c] + pos[c];
}

__global__ void f32_to_bf16_kernel(const float *src, __nv_bfloat16 *dst, int64_t n)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 t = *reinterpret_cast<const float4 *>(src + i);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(t.x, t.y), h1 = __floats2bfloat162_rn(t.z, t.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t *>(&h0);
        u.y = *reinterpret_cast<uint32_t *>(&h1);
        *reinterpret_cast<uint2 *>(dst + i) = u;
    } else {
        for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
    }
}

// NCHW [n,3,224,224] -> rows [n*G*G, ldk], column order (c, dy, dx) = conv1.weight.flatten(1) order.
template <typename T>
__global__ void im2col_kernel(const T *img, int n_img, int P, int G, int ldk, __nv_bfloat16 *out, int out_f16)
{
    const int K = 3 * P * P;
    const int64_t total = (int64_t)n_img * G * G * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % K);
        const int64_t row = i / K;
        const int img_i = (int)(row / (G * G)), pr = (int)(row % (G * G));
        const int py = pr / G, px = pr % G;
        const int c = col / (P * P), dy = (col / P) % P, dx = col % P;
        const float v = (float)img[(((size_t)img_i * 3 + c) * 224 + py * P + dy) * 224 + px * P + dx];
        if (out_f16) reinterpret_cast<__half *>(out)[row * ldk + col] = __float2half_rn(v);
        else out[row * ldk + col] = __float2bfloat16(v);
    }
}

__global__ void lora_merge_kernel(const float *W, const float *up, const float *down, int rows, int d, int r,
                                  __nv_bfloat16 *Wm)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * d) return;
    const int row = (int)(i / d), col = (int)(i % d);
    float acc = 0.f;
    if (up && down)
        for (int k = 0; k < r; ++k) acc = fmaf(up[(size_t)row * r + k], down[(size_t)k * d + col], acc);
    Wm[i] = __float2bfloat16(W[i] + acc);
}

__global__ void blend_kernel(const float *a, const float *b, float r, float q, float *out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(__fmul_rn(a[i], r), __fmul_rn(b[i], q));
}

// dst[s] = idx[s] >= 0 ? src[idx[s]] : 0   (the zeros + masked assignment of models/clip_cls.py:320-321)
__global__ void gather_rows_kernel(const float *src, const int32_t *idx, float *dst, int n_rows, int C)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_rows * C) return;
    const int r = (int)(i / C), c = (int)(i % C);
    const int sr = idx[r];
    dst[i] = sr >= 0 ? src[(size_t)sr * C + c] : 0.f;
}

// F.normalize(x, p=2, dim=-1): x / max(||x||, 1e-12); one warp per row
__global__ void l2norm_rows_kernel(const float *x, float *out, int M, int C)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = x[(size_t)row * C + c]; ss += v * v; }
    const float inv = 1.f / fmaxf(sqrtf(ec::warp_sum(ss)), 1e-12f);
    for (int c = lane; c < C; c += 32) out[(size_t)row * C + c] = x[(size_t)row * C + c] * inv;
}

// center_events (datasets/utils.py:38-57): one CTA per sample, min/max reduction then the shift, float32 like numpy
__global__ void __launch_bounds__(1024) center_events_kernel(float4 *ev, const int64_t *offsets, int H, int W)
{
    const int64_t lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
    __shared__ float red[5][32];
    __shared__ float sh[3];
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY, tmin = INFINITY;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float4 e = ev[i];
        xmin = fminf(xmin, e.x); xmax = fmaxf(xmax, e.x);
        ymin = fminf(ymin, e.y); ymax = fmaxf(ymax, e.y);
        tmin = fminf(tmin, e.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = xmin; red[1][w] = xmax; red[2][w] = ymin; red[3][w] = ymax; red[4][w] = tmin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            red[0][0] = fminf(red[0][0], red[0][k]); red[1][0] = fmaxf(red[1][0], red[1][k]);
            red[2][0] = fminf(red[2][0], red[2][k]); red[3][0] = fmaxf(red[3][0], red[3][k]);
            red[4][0] = fminf(red[4][0], red[4][k]);
        }
        // ((max + min + 1.) - W) // 2.  in float32
        sh[0] = floorf(__fdiv_rn(__fsub_rn(__fadd_rn(__fadd_rn(red[1][0], red[0][0]), 1.0f), (float)W), 2.0f));
        sh[1] = floorf(__fdiv_rn(__fsub_rn(__fadd_rn(__fadd_rn(red[3][0], red[2][0]), 1.0f), (float)H), 2.0f));
        sh[2] = red[4][0];
    }
    __syncthreads();
    const float sx = sh[0], sy = sh[1], st = sh[2];
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float4 e = ev[i];
        e.x = __fsub_rn(e.x, sx); e.y = __fsub_rn(e.y, sy); e.z = __fsub_rn(e.z, st);
        ev[i] = e;
    }
}

// h-flip / t-flip with p = 1 (datasets/utils.py:18-35); one CTA per sample
__global__ void __launch_bounds__(1024) flip_events_kernel(const float4 *src, float4 *dst, const int64_t *offsets, int W,
                                                           int hflip, int tflip)
{
    const int64_t lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
    if (hi <= lo) return;
    const float t_last = src[hi - 1].z;     // after the reversal this is events[0, 2]
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float4 e = src[tflip ? (hi - 1 - (i - lo)) : i];
        if (hflip) e.x = __fsub_rn((float)(W - 1), e.x);
        if (tflip) { e.z = __fsub_rn(t_last, e.z); e.w = -e.w; }
        dst[i] = e;
    }
}

// x[n*L + l, :] = table[tokens[n, l], :] + pos[l, :]   (CLIP text tower input, openai-CLIP encode_text [3P])
__global__ void embed_tokens_kernel(const float *table, const int32_t *tokens, const float *pos, float *out, int n_rows, int L,
                                    int d, int vocab)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_rows * d) return;
    const int r = (int)(i / d), c = (int)(i % d);
    int tok = tokens[r];
    tok = min(max(tok, 0), vocab - 1);
    out[i] = table[(size_t)tok * d + c] + pos[(size_t)(r % L) * d + c];
}

}  // namespace

extern "C" int ec_embed_tokens(const float *table, const int32_t *tokens, const float *pos, float *out, int n_seq, int L,
                               int d, int vocab, void *stream)
{
    EC_REQUIRE(table && tokens && pos && out && n_seq > 0 && L > 0 && d > 0 && vocab > 0, "ec_embed_tokens: bad arguments");
    const int64_t n = (int64_t)n_seq * L * d;
    embed_tokens_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(table, tokens, pos, out, n_seq * L, L, d,
                                                                                       vocab);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_center_events(float *events, const int64_t *offsets, int B, int H, int W, void *stream)
{
    EC_REQUIRE(events && offsets && B >= 0 && H > 0 && W > 0, "ec_center_events: bad arguments");
    EC_REQUIRE(((uintptr_t)events & 15) == 0, "ec_center_events: events must be 16-byte aligned");
    if (B == 0) return EC_OK;
    center_events_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(events), offsets, H, W);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_flip_events(const float *src, float *dst, const int64_t *offsets, int B, int W, int hflip, int tflip,
                              void *stream)
{
    EC_REQUIRE(src && dst && offsets && src != dst && B >= 0 && W > 0, "ec_flip_events: bad arguments (out of place only)");
    EC_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "ec_flip_events: events must be 16-byte aligned");
    if (B == 0) return EC_OK;
    flip_events_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst),
                                                              offsets, W, hflip, tflip);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_gather_rows(const float *src, const int32_t *idx, float *dst, int n_rows, int C, void *stream)
{
    EC_REQUIRE(src && idx && dst && n_rows > 0 && C > 0, "ec_gather_rows: bad arguments");
    const int64_t n = (int64_t)n_rows * C;
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n_rows, C);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_l2norm_rows(const float *x, float *out, int M, int C, void *stream)
{
    EC_REQUIRE(x && out && M > 0 && C > 0, "ec_l2norm_rows: bad arguments");
    l2norm_rows_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, out, M, C);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_layernorm(const float *x, int64_t row_stride_in, const float *gamma, const float *beta, int M, int d,
                            void *out_bf16, float *out_f32, void *stream)
{
    return launch_ln(x, row_stride_in, gamma, beta, M, d, out_bf16, out_f32, (cudaStream_t)stream);
}

extern "C" int ec_layernorm_f16x2(const float *x, int64_t row_stride, const float *gamma, const float *beta, int M, int d, void *out_hi,
                                  void *out_lo, void *stream)
{
    EC_REQUIRE(x && gamma && beta && out_hi && out_lo, "ec_layernorm_f16x2: null pointer");
    EC_REQUIRE(M > 0 && d > 0 && d % 4 == 0 && d <= 1024 && row_stride % 4 == 0, "ec_layernorm_f16x2: d must be a multiple of 4, at most 1024");
    const unsigned grid = (unsigned)((M + 7) / 8);
    layernorm_kernel<true, float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, row_stride, gamma, beta, M, d, nullptr, nullptr, (__half *)out_hi,
                                                                           (__half *)out_lo);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_layernorm_ex(const void *x, int x_is_f16, int64_t row_stride_in, const float *gamma, const float *beta, int M,
                               int d, void *out_bf16, float *out_f32, void *out_f16, void *stream)
{
    if (x_is_f16)
        return launch_ln_t<__half>((const __half *)x, row_stride_in, gamma, beta, M, d, out_bf16, out_f32, out_f16, (cudaStream_t)stream);
    return launch_ln_t<float>((const float *)x, row_stride_in, gamma, beta, M, d, out_bf16, out_f32, out_f16, (cudaStream_t)stream);
}

// (sum, sum of squares) of every fp16 row into part 0 of a [M, n_parts] float2 table (other parts zeroed): the statistics the
// LayerNorm-folded GEMM (ec_gemm_ln) reads when the rows were not written by a statistics-producing epilogue
__global__ void __launch_bounds__(256) row_stats_f16_kernel(const __half *__restrict__ x, int64_t row_stride, int M, int d,
                                                            float2 *__restrict__ stats, int n_parts)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const uint4 *xr = reinterpret_cast<const uint4 *>(x + (size_t)row * row_stride);
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < (d >> 3); j += 32) {
        const uint4 u = xr[j];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[e]));
            s1 += f.x + f.y;
            s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
        }
    }
    s1 = ec::warp_sum(s1); s2 = ec::warp_sum(s2);
    if (lane < n_parts) stats[(size_t)row * n_parts + lane] = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
}

extern "C" int ec_row_stats_f16(const void *x, int64_t row_stride, int M, int d, float *stats, int n_parts, void *stream)
{
    EC_REQUIRE(x && stats && M > 0 && d > 0 && d % 8 == 0 && row_stride % 8 == 0 && n_parts >= 1 && n_parts <= 32 &&
               ((uintptr_t)x & 15) == 0 && ((uintptr_t)stats & 7) == 0, "ec_row_stats_f16: bad arguments");
    row_stats_f16_kernel<<<(unsigned)((M + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const __half *)x, row_stride, M, d,
                                                                                    (float2 *)stats, n_parts);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_layernorm_f32(const float *x, const float *gamma, const float *beta, int M, int d, float *out,
                                void *stream)
{
    return launch_ln(x, d, gamma, beta, M, d, nullptr, out, (cudaStream_t)stream);
}

extern "C" int ec_cls_rows(float *x, const float *class_embedding, const float *pos, int n_img, int L, int d,
                           void *stream)
{
    EC_REQUIRE(x && class_embedding && pos && n_img > 0 && L > 0 && d > 0, "ec_cls_rows: bad arguments");
    const int64_t n = (int64_t)n_img * d;
    cls_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, class_embedding, pos, n_img, L, d);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_f32_to_bf16(const float *src, void *dst, int64_t n, void *stream)
{
    EC_REQUIRE(src && dst && n >= 0, "ec_f32_to_bf16: bad arguments");
    EC_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "ec_f32_to_bf16: misaligned pointers");
    if (n == 0) return EC_OK;
    const int64_t thr = (n + 3) / 4;
    f32_to_bf16_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16 *)dst, n);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_im2col(const void *img, int in_is_bf16, int n_img, int patch, int ldk, void *out, void *stream)
{
    EC_REQUIRE(img && out && n_img > 0, "ec_im2col: bad arguments");
    EC_REQUIRE(patch > 0 && 224 % patch == 0 && ldk >= 3 * patch * patch, "ec_im2col: bad patch %d / ldk %d", patch, ldk);
    const int G = 224 / patch;
    const int64_t total = (int64_t)n_img * G * G * 3 * patch * patch;
    const unsigned grid = (unsigned)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    const int out_f16 = (in_is_bf16 >> 1) & 1;
    if (in_is_bf16 & 1)
        im2col_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)img, n_img, patch, G, ldk,
                                                                              (__nv_bfloat16 *)out, out_f16);
    else
        im2col_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)img, n_img, patch, G, ldk,
                                                                      (__nv_bfloat16 *)out, out_f16);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_lora_merge(const float *W, const float *up, const float *down, int rows, int d, int r, void *Wm_bf16,
                             void *stream)
{
    EC_REQUIRE(W && Wm_bf16 && rows > 0 && d > 0, "ec_lora_merge: bad arguments");
    EC_REQUIRE((up == nullptr) == (down == nullptr), "ec_lora_merge: up and down must both be given or both be NULL");
    const int64_t n = (int64_t)rows * d;
    lora_merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, up, down, rows, d, r,
                                                                                     (__nv_bfloat16 *)Wm_bf16);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_blend(const float *a, const float *b, double r, float *out, int64_t n, void *stream)
{
    EC_REQUIRE(a && b && out && n >= 0, "ec_blend: bad arguments");
    if (n == 0) return EC_OK;
    // python evaluates (1. - residual) in double before torch narrows both scalars to float32
    blend_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, (float)r, (float)(1.0 - r), out, n);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}
