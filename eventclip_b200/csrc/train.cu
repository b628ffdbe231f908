// train.cu -- elementwise / row kernels of the fine-tune step's backward pass (SURVEY.md section 8 row A12:
// models/clip_cls_ft.py:214-269 trains through model.visual; the reference relies on autograd for all of this).
//   LayerNorm backward, QuickGELU forward/backward on the saved pre-activation, bf16 transpose (weight-gradient GEMMs
//   reduce over the token dimension), Adam (method.py:150-191 uses torch.optim.Adam with two learning rates).
#include "common.cuh"

namespace {

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma ; one warp per row; dx is ADDED to `acc` when given.
// VPL > 0: the row (x and g) stays in registers, VPL float4 per lane (d == 128 * VPL): every global load is issued before
// the first reduction.  VPL == 0: generic three-pass version.  dx_bf16 (optional): bf16 copy of dx, the A operand of the
// next data-gradient GEMM.
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float *__restrict__ x, int64_t x_stride,
                                                            const float *__restrict__ dy, const float *__restrict__ gamma,
                                                            const float *__restrict__ acc, int64_t acc_stride, int M, int d,
                                                            float *__restrict__ dx, int64_t dx_stride, __nv_bfloat16 *__restrict__ dxb)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *xr = x + (size_t)row * x_stride, *dr = dy + (size_t)row * d;
    if (VPL > 0) {
        float4 xv[VPL > 0 ? VPL : 1], gv[VPL > 0 ? VPL : 1], av[VPL > 0 ? VPL : 1];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            xv[i] = *reinterpret_cast<const float4 *>(xr + c);
            gv[i] = *reinterpret_cast<const float4 *>(dr + c);
            av[i] = acc ? *reinterpret_cast<const float4 *>(acc + (size_t)row * acc_stride + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
        const float mean = ec::warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
            q += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
        }
        const float rstd = rsqrtf(ec::warp_sum(q) / (float)d + 1e-5f);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 gm = *reinterpret_cast<const float4 *>(gamma + (i * 32 + lane) * 4);
            gv[i].x *= gm.x; gv[i].y *= gm.y; gv[i].z *= gm.z; gv[i].w *= gm.w;
            xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;          // xhat
            sg += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
            sgx += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
        }
        sg = ec::warp_sum(sg) / (float)d; sgx = ec::warp_sum(sgx) / (float)d;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            float4 o;
            o.x = rstd * (gv[i].x - sg - xv[i].x * sgx) + av[i].x;
            o.y = rstd * (gv[i].y - sg - xv[i].y * sgx) + av[i].y;
            o.z = rstd * (gv[i].z - sg - xv[i].z * sgx) + av[i].z;
            o.w = rstd * (gv[i].w - sg - xv[i].w * sgx) + av[i].w;
            *reinterpret_cast<float4 *>(dx + (size_t)row * dx_stride + c) = o;
            if (dxb) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
                uint2 pk; pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                *reinterpret_cast<uint2 *>(dxb + (size_t)row * d + c) = pk;
            }
        }
        return;
    }
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = ec::warp_sum(s) / (float)d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; q += t * t; }
    const float rstd = rsqrtf(ec::warp_sum(q) / (float)d + 1e-5f);
    float sg = 0.f, sgx = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float g = dr[c] * gamma[c], xh = (xr[c] - mean) * rstd;
        sg += g; sgx += g * xh;
    }
    sg = ec::warp_sum(sg) / (float)d; sgx = ec::warp_sum(sgx) / (float)d;
    for (int c = lane; c < d; c += 32) {
        const float g = dr[c] * gamma[c], xh = (xr[c] - mean) * rstd;
        float v = rstd * (g - sg - xh * sgx);
        if (acc) v += acc[(size_t)row * acc_stride + c];
        dx[(size_t)row * dx_stride + c] = v;
        if (dxb) dxb[(size_t)row * d + c] = __float2bfloat16(v);
    }
}

// 8 bf16 per thread (one 16-byte access); n8 = n / 8 full vectors, the tail is handled by the last thread scalar-wise
__device__ __forceinline__ float qg(float x) { return x / (1.f + __expf(-1.702f * x)); }
// d/da [a * sigmoid(1.702 a)] = s + 1.702 a s (1 - s)
__device__ __forceinline__ float qg_grad(float x)
{
    const float sg = 1.f / (1.f + __expf(-1.702f * x));
    return sg + 1.702f * x * sg * (1.f - sg);
}

__global__ void __launch_bounds__(256) quickgelu_kernel(const __nv_bfloat16 *__restrict__ a, __nv_bfloat16 *__restrict__ h, int64_t n)
{
    const int64_t n8 = n >> 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v = reinterpret_cast<const uint4 *>(a)[i];
        __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = __bfloat1622float2(p[e]);
            p[e] = __floats2bfloat162_rn(qg(x.x), qg(x.y));
        }
        reinterpret_cast<uint4 *>(h)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = n8 << 3; i < n; ++i) h[i] = __float2bfloat16(qg(__bfloat162float(a[i])));
}

__global__ void __launch_bounds__(256) quickgelu_bwd_kernel(const __nv_bfloat16 *__restrict__ a, const __nv_bfloat16 *__restrict__ dh,
                                                            __nv_bfloat16 *__restrict__ da, int64_t n)
{
    const int64_t n8 = n >> 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 va = reinterpret_cast<const uint4 *>(a)[i];
        uint4 vd = reinterpret_cast<const uint4 *>(dh)[i];
        const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&va);
        __nv_bfloat162 *pd = reinterpret_cast<__nv_bfloat162 *>(&vd);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = __bfloat1622float2(pa[e]), g = __bfloat1622float2(pd[e]);
            pd[e] = __floats2bfloat162_rn(g.x * qg_grad(x.x), g.y * qg_grad(x.y));
        }
        reinterpret_cast<uint4 *>(da)[i] = vd;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = n8 << 3; i < n; ++i)
            da[i] = __float2bfloat16(__bfloat162float(dh[i]) * qg_grad(__bfloat162float(a[i])));
}

// LoRA factor gradients from the gradient of the merged weight (W_eff = W + up . down, models/lora.py:138-149, 49-52):
//   d_up[rows, r] = dW . down^T        d_down[r, d] = up^T . dW
// dW holds n_mat matrices of `rows` rows stacked (q | k | v of in_proj, or the single out_proj); blockIdx.y picks one.
struct LoraGradArgs {
    const float *up[4], *down[4];
    float *d_up[4], *d_down[4];
};

// One launch, both products: blockIdx.x < ceil(rows/16) computes a 16-row tile of d_up (all ranks, 16 at a time), the other
// blocks a 16-column tile of d_down.  256 threads = 16 x 16 outputs; the reduction index is staged through shared memory
// 128 deep, so a CTA issues 6 rounds of wide loads for d = rows = 768 instead of a dependent load per multiply-add.
constexpr int LG_KC = 128;
__global__ void __launch_bounds__(256) lora_grad_kernel(const float *__restrict__ dW, int64_t ld, int rows, int d, int r, LoraGradArgs g)
{
    __shared__ float sa[16][LG_KC + 1];      // d_up: dW[row][k]      d_down: up^T[j][k]   (k = reduction index)
    __shared__ float sb[16][LG_KC + 1];      // d_up: down[j][k]      d_down: dW^T[c][k]
    const int z = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    if (!g.d_up[z]) return;
    const float *w = dW + (size_t)z * rows * ld, *up = g.up[z], *dn = g.down[z];
    const int up_tiles = (rows + 15) / 16;
    if ((int)blockIdx.x < up_tiles) {
        // d_up[row0 + ty][j0 + tx] = sum_c dW[row][c] * down[j][c]
        const int row0 = blockIdx.x * 16;
        for (int j0 = 0; j0 < r; j0 += 16) {
            float acc = 0.f;
            for (int k0 = 0; k0 < d; k0 += LG_KC) {
                for (int e = tid; e < 16 * LG_KC; e += 256) {
                    const int rr = e / LG_KC, k = e - rr * LG_KC;
                    sa[rr][k] = (row0 + rr < rows && k0 + k < d) ? w[(size_t)(row0 + rr) * ld + k0 + k] : 0.f;
                    sb[rr][k] = (j0 + rr < r && k0 + k < d) ? dn[(size_t)(j0 + rr) * d + k0 + k] : 0.f;
                }
                __syncthreads();
#pragma unroll 16
                for (int k = 0; k < LG_KC; ++k) acc += sa[ty][k] * sb[tx][k];
                __syncthreads();
            }
            if (row0 + ty < rows && j0 + tx < r) g.d_up[z][(size_t)(row0 + ty) * r + j0 + tx] = acc;
        }
    } else {
        // d_down[j0 + ty][c0 + tx] = sum_i up[i][j] * dW[i][c]
        const int c0 = ((int)blockIdx.x - up_tiles) * 16;
        for (int j0 = 0; j0 < r; j0 += 16) {
            float acc = 0.f;
            for (int k0 = 0; k0 < rows; k0 += LG_KC) {
                for (int e = tid; e < 16 * LG_KC; e += 256) {
                    const int k = e >> 4, q = e & 15;          // 16 consecutive floats of one row per half-warp
                    sa[q][k] = (k0 + k < rows && j0 + q < r) ? up[(size_t)(k0 + k) * r + j0 + q] : 0.f;
                    sb[q][k] = (k0 + k < rows && c0 + q < d) ? w[(size_t)(k0 + k) * ld + c0 + q] : 0.f;
                }
                __syncthreads();
#pragma unroll 16
                for (int k = 0; k < LG_KC; ++k) acc += sa[ty][k] * sb[tx][k];
                __syncthreads();
            }
            if (j0 + ty < r && c0 + tx < d) g.d_down[z][(size_t)(j0 + ty) * d + c0 + tx] = acc;
        }
    }
}

// out[c, r] = in[r, c]; 64x64 tiles through shared memory, 4-byte (bf16x2) global accesses on both sides when the
// strides allow it (even ld_in / ld_out), 128 B per warp row
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16 *__restrict__ in, __nv_bfloat16 *__restrict__ out,
                                                             int rows, int cols, int64_t ld_in, int64_t ld_out, int vec)
{
    __shared__ __nv_bfloat16 t[64][66];
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const __nv_bfloat16 zero = __float2bfloat16(0.f);
    for (int i = ty; i < 64; i += 8) {
        const int r = r0 + i, c = c0 + 2 * tx;
        __nv_bfloat16 a = zero, b = zero;
        if (r < rows) {
            if (vec && c + 1 < cols) {
                const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(in + (size_t)r * ld_in + c);
                a = v.x; b = v.y;
            } else {
                if (c < cols) a = in[(size_t)r * ld_in + c];
                if (c + 1 < cols) b = in[(size_t)r * ld_in + c + 1];
            }
        }
        t[i][2 * tx] = a; t[i][2 * tx + 1] = b;
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i, r = r0 + 2 * tx;
        if (c >= cols) continue;
        const __nv_bfloat16 a = t[2 * tx][i], b = t[2 * tx + 1][i];
        if (vec && r + 1 < rows) {
            __nv_bfloat162 v; v.x = a; v.y = b;
            *reinterpret_cast<__nv_bfloat162 *>(out + (size_t)c * ld_out + r) = v;
        } else {
            if (r < rows) out[(size_t)c * ld_out + r] = a;
            if (r + 1 < rows) out[(size_t)c * ld_out + r + 1] = b;
        }
    }
}

__global__ void adam_kernel(float *p, const float *g, float *m, float *v, int64_t n, float lr, float b1, float b2, float eps,
                            float wd, float bc1, float bc2)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] + wd * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);   // torch.optim.Adam: denom = sqrt(v_hat) + eps
}


// out[m,n] (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]; 32x32 tiles (LoRA factor gradients, head backward: small shapes)
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float *__restrict__ A, int64_t sam, int64_t sak,
                                                            const float *__restrict__ B, int64_t sbk, int64_t sbn, int M, int N,
                                                            int K, float alpha, float *__restrict__ out, int64_t ldo, int accumulate)
{
    __shared__ float sa[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int i = ty; i < 32; i += 8) {
            // pick the thread->element mapping that is contiguous in memory for each operand
            if (sak == 1) { const int m = m0 + i, k = k0 + tx; sa[i][tx] = (m < M && k < K) ? A[m * sam + k] : 0.f; }
            else          { const int m = m0 + tx, k = k0 + i; sa[tx][i] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f; }
            if (sbn == 1) { const int k = k0 + i, n = n0 + tx; sb[i][tx] = (k < K && n < N) ? B[k * sbk + n] : 0.f; }
            else          { const int k = k0 + tx, n = n0 + i; sb[tx][i] = (k < K && n < N) ? B[k * sbk + n * sbn] : 0.f; }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float b = sb[k][tx];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] += sa[ty + 8 * r][k] * b;
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + ty + 8 * r, n = n0 + tx;
        if (m < M && n < N) {
            float *o = out + (size_t)m * ldo + n;
            *o = accumulate ? *o + alpha * acc[r] : alpha * acc[r];
        }
    }
}

// Backward of F.normalize(x, dim=-1) (eps 1e-12) followed by the valid mask: dx = (dy - y <y, dy>) / max(|x|, eps), y = x / max(|x|, eps)
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                         const uint8_t *__restrict__ mask, int M, int C, float *__restrict__ dx)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *xr = x + (size_t)row * C, *dr = dy + (size_t)row * C;
    float *o = dx + (size_t)row * C;
    if (mask && !mask[row]) { for (int c = lane; c < C; c += 32) o[c] = 0.f; return; }
    float ss = 0.f, sd = 0.f;
    for (int c = lane; c < C; c += 32) { ss += xr[c] * xr[c]; sd += xr[c] * dr[c]; }
    ss = ec::warp_sum(ss); sd = ec::warp_sum(sd);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    const float dot = sd * inv;                       // <y, dy>
    for (int c = lane; c < C; c += 32) o[c] = (dr[c] - xr[c] * inv * dot) * inv;
}

// Loss head of the fine-tune step (clip_cls_ft.py:191-212 aggregate, 258-269 F.cross_entropy on the aggregated logits):
// one warp per sample; writes the per-sample loss, the gradient of the MEAN loss w.r.t. full_logits, and (one CTA only,
// after a grid-wide count) nothing else -- the mean is taken by mean_kernel below so the result is order-deterministic.
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float *__restrict__ full, const uint8_t *__restrict__ valid,
                                                     const int32_t *__restrict__ labels, int B, int T, int K, int agg,
                                                     float *__restrict__ loss_b, float *__restrict__ dfull)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    int nv = 0;
    for (int t = 0; t < T; ++t) nv += valid[b * T + t] ? 1 : 0;
    const float w = agg == EC_AGG_MEAN ? 1.f / (float)nv : 1.f;     // nv == 0 gives inf/nan exactly like the reference's 0/0
    const float *fb = full + (size_t)b * T * K;
    const int y = labels[b];
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) {
        float z = 0.f;
        for (int t = 0; t < T; ++t) z += fb[(size_t)t * K + k];
        mx = fmaxf(mx, z * w);
    }
    mx = ec::warp_max(mx);
    float se = 0.f, zy = 0.f;
    for (int k = lane; k < K; k += 32) {
        float z = 0.f;
        for (int t = 0; t < T; ++t) z += fb[(size_t)t * K + k];
        z *= w;
        se += __expf(z - mx);
        if (k == y) zy = z;
    }
    se = ec::warp_sum(se); zy = ec::warp_sum(zy);
    const float lse = mx + __logf(se);
    if (lane == 0) loss_b[b] = lse - zy;
    const float invB = 1.f / (float)B;
    for (int k = lane; k < K; k += 32) {
        float z = 0.f;
        for (int t = 0; t < T; ++t) z += fb[(size_t)t * K + k];
        z *= w;
        const float g = (__expf(z - lse) - (k == y ? 1.f : 0.f)) * invB * w;
        // every VALID view slot receives the gradient of the sum.  The reference sums over all T rows, but its padded rows are
        // features multiplied by the valid mask (clip_cls.py:330-333): zero rows whose logits carry no gradient to anything.
        // Writing zeros there keeps d_text exact when the padded slots hold non-zero adapter outputs (few-shot training).
        for (int t = 0; t < T; ++t) dfull[((size_t)b * T + t) * K + k] = valid[b * T + t] ? g : 0.f;
    }
}

// Probability loss of the reference (use_probs_loss, models/clip_cls_ft.py:265-267 / clip_cls.py:173-175): probs = mean over the
// valid views of softmax(full_logits[b,t,:]) (clip_cls.py:123-129), loss = -log(probs[label] + 1e-6).  One warp per sample.
// d loss / d z[t,k] = -s_t[y] (delta_ky - s_t[k]) / ((p_y + 1e-6) n_valid) for valid views, 0 for padded ones.
constexpr int PROBS_MAX_T = 16;
__global__ void __launch_bounds__(256) probs_bwd_kernel(const float *__restrict__ full, const uint8_t *__restrict__ valid,
                                                        const int32_t *__restrict__ labels, int B, int T, int K,
                                                        float *__restrict__ loss_b, float *__restrict__ dfull)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float *fb = full + (size_t)b * T * K;
    const int y = labels[b];
    float lse[PROBS_MAX_T], sy[PROBS_MAX_T];
    int nv = 0;
    float py = 0.f;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        lse[t] = 0.f; sy[t] = 0.f;
        if (!valid[b * T + t]) continue;
        ++nv;
        const float *z = fb + (size_t)t * K;
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[k]);
        mx = ec::warp_max(mx);
        float se = 0.f;
        for (int k = lane; k < K; k += 32) se += expf(z[k] - mx);
        se = ec::warp_sum(se);
        lse[t] = mx + logf(se);
        sy[t] = expf(z[y] - lse[t]);
        py += sy[t];
    }
    py /= (float)nv;                                   // nv == 0 gives nan exactly like the reference's 0/0
    if (lane == 0) loss_b[b] = -logf(py + 1e-6f);
    const float coef = -1.f / ((py + 1e-6f) * (float)nv * (float)B);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        float *g = dfull + ((size_t)b * T + t) * K;
        if (!valid[b * T + t]) {
            for (int k = lane; k < K; k += 32) g[k] = 0.f;
            continue;
        }
        const float *z = fb + (size_t)t * K;
        const float c = coef * sy[t];
        for (int k = lane; k < K; k += 32) g[k] = c * ((k == y ? 1.f : 0.f) - expf(z[k] - lse[t]));
    }
}

__global__ void mean_kernel(const float *v, int n, float *out)
{
    __shared__ float s[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += v[i];
    a = ec::warp_sum(a);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
        a = ec::warp_sum(a);
        if (threadIdx.x == 0) out[0] = a / (float)n;
    }
}


// ---- column reductions over the token dimension (bias / LayerNorm affine / embedding gradients), two deterministic stages:
//      partial sums of row slabs, then a fixed-order fold ----
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(128) colsum_part_kernel(const T *__restrict__ a, int M, int N, int64_t ld, float *__restrict__ part)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= N) return;
    const int per = (M + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(M, r0 + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += to_f(a[(size_t)r * ld + c]); s1 += to_f(a[(size_t)(r + 1) * ld + c]);
        s2 += to_f(a[(size_t)(r + 2) * ld + c]); s3 += to_f(a[(size_t)(r + 3) * ld + c]);
    }
    for (; r < r1; ++r) s0 += to_f(a[(size_t)r * ld + c]);
    part[(size_t)blockIdx.y * N + c] = (s0 + s1) + (s2 + s3);
}

// out[c] = sum_p part[p, c] for n_out stacked outputs of width N each (part is [n_out][P][N])
__global__ void __launch_bounds__(128) fold_kernel(const float *__restrict__ part, int P, int N, float *const o0, float *const o1)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= N) return;
    float *const out = blockIdx.y == 0 ? o0 : o1;
    const float *p = part + (size_t)blockIdx.y * P * N;
    float s = 0.f;
    for (int q = 0; q < P; ++q) s += p[(size_t)q * N + c];
    out[c] = s;
}

// LayerNorm affine gradients: dgamma[c] = sum_rows dy * xhat, dbeta[c] = sum_rows dy.  Row statistics are recomputed per
// row by a warp (the row is read once into registers for d <= 1024), then the 8 rows of a CTA are folded through shared
// memory; slabs of rows -> part[2][P][d].
__global__ void __launch_bounds__(256) ln_param_part_kernel(const float *__restrict__ x, int64_t x_stride, const float *__restrict__ dy,
                                                            int M, int d, float *__restrict__ part, int P)
{
    extern __shared__ float sm[];                 // [2][d] accumulators of this CTA
    float *sg = sm, *sbv = sm + d;
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) sm[c] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (M + P - 1) / P, r0 = blockIdx.x * per, r1 = min(M, r0 + per);
    // rows are taken by the 8 warps in turn; the shared accumulators are updated one warp at a time (fixed order)
    for (int rb = r0; rb < r1; rb += 8) {
        const int row = rb + warp;
        float mean = 0.f, rstd = 0.f;
        if (row < r1) {
            const float *xr = x + (size_t)row * x_stride;
            float s = 0.f;
            for (int c = lane; c < d; c += 32) s += xr[c];
            mean = ec::warp_sum(s) / (float)d;
            float q = 0.f;
            for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; q += t * t; }
            rstd = rsqrtf(ec::warp_sum(q) / (float)d + 1e-5f);
        }
        for (int w = 0; w < 8; ++w) {
            if (w == warp && row < r1) {
                const float *xr = x + (size_t)row * x_stride, *dr = dy + (size_t)row * d;
                for (int c = lane; c < d; c += 32) {
                    const float g = dr[c];
                    sg[c] += g * (xr[c] - mean) * rstd;
                    sbv[c] += g;
                }
            }
            __syncthreads();
        }
    }
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        part[(size_t)blockIdx.x * d + c] = sg[c];
        part[(size_t)(P + blockIdx.x) * d + c] = sbv[c];
    }
}

// dst[r, :] = src[map(r), :] where map skips the class-token row of every image: r -> (r / G2) * (G2 + 1) + 1 + r % G2;
// fp32 -> bf16 (the patch-token rows of d_x0 as the A operand of conv1's weight gradient)
__global__ void __launch_bounds__(256) patch_rows_bf16_kernel(const float *__restrict__ src, int n_rows, int G2, int d,
                                                              __nv_bfloat16 *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 elements
    const int d4 = d >> 2;
    if (i >= (int64_t)n_rows * d4) return;
    const int r = (int)(i / d4), c = (int)(i - (int64_t)r * d4) * 4;
    const int img = r / G2, tok = r - img * G2;
    const float4 v = *reinterpret_cast<const float4 *>(src + ((size_t)img * (G2 + 1) + 1 + tok) * d + c);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t *>(&a); o.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(dst + (size_t)r * d + c) = o;
}

}  // namespace

extern "C" int ec_layernorm_bwd(const float *x, int64_t x_stride, const float *dy, const float *gamma, const float *acc,
                                int64_t acc_stride, int M, int d, float *dx, int64_t dx_stride, void *dx_bf16, void *stream)
{
    EC_REQUIRE(x && dy && gamma && dx && M > 0 && d > 0, "ec_layernorm_bwd: bad arguments");
    const dim3 grid((M + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16 *dxb = (__nv_bfloat16 *)dx_bf16;
    const bool vec = d % 128 == 0 && d <= 1024 && x_stride % 4 == 0 && dx_stride % 4 == 0 && (!acc || acc_stride % 4 == 0) &&
                     (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)gamma | (uintptr_t)dx | (uintptr_t)acc) & 15) == 0 &&
                     ((uintptr_t)dxb & 7) == 0;
#define EC_LNB(V) layernorm_bwd_kernel<V><<<grid, 256, 0, st>>>(x, x_stride, dy, gamma, acc, acc_stride, M, d, dx, dx_stride, dxb)
    if (vec) {
        switch (d / 128) {
            case 1: EC_LNB(1); break;
            case 2: EC_LNB(2); break;
            case 3: EC_LNB(3); break;
            case 4: EC_LNB(4); break;
            case 5: EC_LNB(5); break;
            case 6: EC_LNB(6); break;
            case 7: EC_LNB(7); break;
            default: EC_LNB(8); break;
        }
    } else {
        EC_LNB(0);
    }
#undef EC_LNB
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_quickgelu(const void *a, void *h, int64_t n, void *stream)
{
    EC_REQUIRE(a && h && n >= 0 && (((uintptr_t)a | (uintptr_t)h) & 15) == 0, "ec_quickgelu: bad arguments (16-byte aligned pointers)");
    if (n == 0) return EC_OK;
    quickgelu_kernel<<<(unsigned)((n / 8 + 255) / 256 + 1 < 148 * 16 ? (n / 8 + 255) / 256 + 1 : 148 * 16), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)a, (__nv_bfloat16 *)h, n);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_quickgelu_bwd(const void *a, const void *dh, void *da, int64_t n, void *stream)
{
    EC_REQUIRE(a && dh && da && n >= 0 && (((uintptr_t)a | (uintptr_t)dh | (uintptr_t)da) & 15) == 0,
               "ec_quickgelu_bwd: bad arguments (16-byte aligned pointers)");
    if (n == 0) return EC_OK;
    quickgelu_bwd_kernel<<<(unsigned)((n / 8 + 255) / 256 + 1 < 148 * 16 ? (n / 8 + 255) / 256 + 1 : 148 * 16), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)a,
                                                                                        (const __nv_bfloat16 *)dh,
                                                                                        (__nv_bfloat16 *)da, n);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_transpose_bf16(const void *in, void *out, int rows, int cols, int64_t ld_in, int64_t ld_out, void *stream)
{
    EC_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "ec_transpose_bf16: bad arguments");
    const int vec = (ld_in % 2 == 0) && (ld_out % 2 == 0) && (((uintptr_t)in | (uintptr_t)out) & 3) == 0;
    transpose_bf16_kernel<<<dim3((cols + 63) / 64, (rows + 63) / 64), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)in, (__nv_bfloat16 *)out, rows, cols, ld_in, ld_out, vec);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_adam(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, void *stream)
{
    EC_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "ec_adam: bad arguments");
    if (n == 0) return EC_OK;
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                               eps, weight_decay, bc1, bc2);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_gemm_f32_strided(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn, int M, int N,
                                   int K, float alpha, float *out, int64_t ldo, int accumulate, void *stream)
{
    EC_REQUIRE(A && B && out && M > 0 && N > 0 && K > 0 && ldo >= N, "ec_gemm_f32_strided: bad arguments");
    sgemm_strided_kernel<<<dim3((N + 31) / 32, (M + 31) / 32), 256, 0, (cudaStream_t)stream>>>(A, sam, sak, B, sbk, sbn, M, N, K, alpha, out,
                                                                                               ldo, accumulate);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_l2norm_rows_bwd(const float *x, const float *dy, const uint8_t *mask, int M, int C, float *dx, void *stream)
{
    EC_REQUIRE(x && dy && dx && M > 0 && C > 0, "ec_l2norm_rows_bwd: bad arguments");
    l2norm_bwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, dy, mask, M, C, dx);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_ce_loss_bwd(const float *full_logits, const uint8_t *valid, const int32_t *labels, int B, int T, int n_cls, int agg,
                              float *loss_per_sample, float *loss_mean, float *d_full, void *stream)
{
    EC_REQUIRE(full_logits && valid && labels && loss_per_sample && loss_mean && d_full && B > 0 && T > 0 && n_cls > 0,
               "ec_ce_loss_bwd: bad arguments");
    EC_REQUIRE(agg == EC_AGG_SUM || agg == EC_AGG_MEAN, "ec_ce_loss_bwd: training supports agg sum / mean (got %d)", agg);
    ce_bwd_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(full_logits, valid, labels, B, T, n_cls, agg, loss_per_sample, d_full);
    mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(loss_per_sample, B, loss_mean);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_probs_loss_bwd(const float *full_logits, const uint8_t *valid, const int32_t *labels, int B, int T, int n_cls,
                                 float *loss_per_sample, float *loss_mean, float *d_full, void *stream)
{
    EC_REQUIRE(full_logits && valid && labels && loss_per_sample && loss_mean && d_full && B > 0 && T > 0 && n_cls > 0,
               "ec_probs_loss_bwd: bad arguments");
    EC_REQUIRE(T <= PROBS_MAX_T, "ec_probs_loss_bwd: at most %d views per sample (got %d)", PROBS_MAX_T, T);
    probs_bwd_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(full_logits, valid, labels, B, T, n_cls, loss_per_sample, d_full);
    mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(loss_per_sample, B, loss_mean);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_lora_grad(const float *dW, int64_t ld, int n_mat, int rows, int d, int r, const float *const *up,
                            const float *const *down, float *const *d_up, float *const *d_down, void *stream)
{
    EC_REQUIRE(dW && up && down && d_up && d_down && n_mat >= 1 && n_mat <= 4 && rows > 0 && d > 0 && r > 0 && ld >= d,
               "ec_lora_grad: bad arguments");
    LoraGradArgs g;
    for (int z = 0; z < 4; ++z) {
        const bool on = z < n_mat && up[z] && down[z];
        g.up[z] = on ? up[z] : nullptr; g.down[z] = on ? down[z] : nullptr;
        g.d_up[z] = on ? d_up[z] : nullptr; g.d_down[z] = on ? d_down[z] : nullptr;
        if (on) EC_REQUIRE(d_up[z] && d_down[z], "ec_lora_grad: matrix %d has factors but no gradient buffers", z);
    }
    lora_grad_kernel<<<dim3((rows + 15) / 16 + (d + 15) / 16, n_mat), 256, 0, (cudaStream_t)stream>>>(dW, ld, rows, d, r, g);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_colsum(const void *a, int is_bf16, int M, int N, int64_t ld, float *scratch, int n_part, float *out, void *stream)
{
    EC_REQUIRE(a && scratch && out && M > 0 && N > 0 && ld >= N && n_part >= 1 && n_part <= 4096, "ec_colsum: bad arguments");
    const dim3 grid((N + 127) / 128, n_part);
    if (is_bf16) colsum_part_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)a, M, N, ld, scratch);
    else colsum_part_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float *)a, M, N, ld, scratch);
    fold_kernel<<<dim3((N + 127) / 128, 1), 128, 0, (cudaStream_t)stream>>>(scratch, n_part, N, out, out);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_layernorm_param_grad(const float *x, int64_t x_stride, const float *dy, int M, int d, float *scratch, int n_part,
                                       float *dgamma, float *dbeta, void *stream)
{
    EC_REQUIRE(x && dy && scratch && dgamma && dbeta && M > 0 && d > 0 && n_part >= 1 && n_part <= 4096 && d <= 4096,
               "ec_layernorm_param_grad: bad arguments");
    ln_param_part_kernel<<<n_part, 256, 2 * d * sizeof(float), (cudaStream_t)stream>>>(x, x_stride, dy, M, d, scratch, n_part);
    fold_kernel<<<dim3((d + 127) / 128, 2), 128, 0, (cudaStream_t)stream>>>(scratch, n_part, d, dgamma, dbeta);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

extern "C" int ec_patch_rows_bf16(const float *src, int n_img, int G2, int d, void *dst, void *stream)
{
    EC_REQUIRE(src && dst && n_img > 0 && G2 > 0 && d > 0 && d % 4 == 0, "ec_patch_rows_bf16: bad arguments");
    const int64_t n = (int64_t)n_img * G2 * (d / 4);
    patch_rows_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n_img * G2, G2, d, (__nv_bfloat16 *)dst);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}
