// event2img.cu -- fused event stream -> CLIP input tensor, one launch, no intermediate in HBM.
//
// Replaces the reference's per-sample CPU pipeline (paths relative to the reference root):
//   parse_events            datasets/vis.py:44-52     (float -> int32 truncation)
//   make_event_histogram    datasets/vis.py:6-41      (counts, hot-pixel removal, /max, gray, white blend, round)
//   colour map              datasets/vis.py:95-101    (grayscale: both polarities map to 127)
//   CLIP preprocess [3P]    datasets/event2img.py:119-122 (Pillow bicubic Resize(224), CenterCrop(224),
//                           ToTensor, Normalize)
// Chunk boundaries (vis.py:55-72) and view selection (event2img.py:80-92) are planned on the host by
// ec_plan_frames() and arrive as the ec_frame table.
//
// One thread-block cluster per frame.  CTA `rank` of a cluster owns a band of RB sensor rows:
//   P1  scan the frame's events (16-byte loads, L1 bypass), shared-memory atomics into packed bins
//       (pos count in bits 0-15, neg count in bits 16-31 of one word per pixel)
//   P2  exact integer statistics (n, S1, S2) -> cluster reduction over DSMEM -> hot-pixel cut `keep`
//   P3  max of the surviving bins -> cluster reduction
//   P4  per-pixel gray value in IEEE fp64 (bit-exact with numpy's float64 path), in place
//   P5  Pillow horizontal 8-bit bicubic pass for the 224 cropped columns -> band rows in smem
//   P6  vertical pass reading neighbour bands through DSMEM, 256-entry normalise LUT, output store
#include <cooperative_groups.h>

#include <map>
#include <mutex>
#include <vector>
#include <cmath>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int OUT = 224;               // CLIP input resolution
constexpr int PREC = 22;               // Pillow PRECISION_BITS = 32 - 8 - 2
constexpr int GLUT_N = 32;             // per-frame gray LUT covers pos,neg < 32

struct E2IParams {
    const float4 *events;
    const ec_frame *frames;
    int H, W, RB, CS;
    int flags, out_fmt, patch, ldk, G;
    void *out;
    int32_t *dbg_counts;
    uint8_t *dbg_gray;
    uint8_t *dbg_u8;
    int32_t *status;
    const int32_t *hx;   // [224][2+KH]: lo, cnt, taps  (cropped output columns)
    const int32_t *vy;   // [224][2+KV]: lo, cnt, taps  (cropped output rows)
    const float *nlut;   // [3][256] normalise LUT
    int KH, KV;
};

struct Part {            // per-CTA partial statistics exchanged over DSMEM
    unsigned long long s1, s2, nnz, nacc;
    unsigned mx;
    unsigned bad;
};

__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// c is removed  <=>  c > mean + t*std  <=>  c*n > S1 and (c*n - S1)^2 > t^2 (n*S2 - S1^2)   (exact integers)
__device__ bool is_hot(long long c, unsigned long long n, unsigned long long S1, unsigned __int128 var_n2, int t)
{
    __int128 d = (__int128)c * (__int128)n - (__int128)S1;
    if (d <= 0) return false;
    return (unsigned __int128)(d * d) > (unsigned __int128)(t * t) * var_n2;
}

// largest count that survives the hot-pixel rule (vis.py:16-23, thresh = 10)
__device__ unsigned compute_keep(unsigned long long n, unsigned long long S1, unsigned long long S2, int t)
{
    if (t <= 0 || n == 0) return 0xffffffffu;
    unsigned __int128 var = (unsigned __int128)n * S2 - (unsigned __int128)S1 * S1;
    double varf = (double)(unsigned long long)(var >> 64) * 18446744073709551616.0 + (double)(unsigned long long)var;
    double thr = ((double)S1 + (double)t * sqrt(varf)) / (double)n;
    long long c = (long long)floor(thr);
    if (c < 0) c = 0;
    for (int it = 0; it < 64 && c > 0 && is_hot(c, n, S1, var, t); ++it) --c;
    for (int it = 0; it < 64 && !is_hot(c + 1, n, S1, var, t); ++it) ++c;
    return c > 0xfffffffell ? 0xfffffffeu : (unsigned)c;
}

// vis.py:27-39 evaluated as numpy >= 2 does (float64); `hist @ cmap` is a 2-term BLAS dot = one fma.
__device__ __forceinline__ unsigned gray_px(unsigned pos, unsigned neg, unsigned mx, bool mask)
{
    if (mx == 0) return 0;   // 0/0 in the reference (undefined there; NaN -> uint8 gives 0 on x86)
    const double m = (double)mx;
    const double gp = __ddiv_rn((double)pos, m);
    const double gn = __ddiv_rn((double)neg, m);
    double img = __fma_rn(gn, 127.0, __dmul_rn(gp, 127.0));
    if (mask) {
        double w = __dadd_rn(gp, gn);
        w = fmin(fmax(w, 0.0), 1.0);
        img = __dadd_rn(__dmul_rn(img, w), __dmul_rn(255.0, __dsub_rn(1.0, w)));
    }
    return (unsigned)__double2int_rn(img);   // np.round: half to even
}

__device__ __forceinline__ unsigned clip8(int v)
{
    v >>= PREC;
    return (unsigned)min(max(v, 0), 255);
}

__device__ __forceinline__ void store_pair(const E2IParams &p, int slot, int yo, int x, unsigned v0, unsigned v1,
                                           const float *nl)
{
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float f0 = nl[c * 256 + v0], f1 = nl[c * 256 + v1];
        if (p.out_fmt == EC_OUT_F32_NCHW) {
            float *o = (float *)p.out + (((size_t)slot * 3 + c) * OUT + yo) * OUT + x;
            *reinterpret_cast<float2 *>(o) = make_float2(f0, f1);
        } else if (p.out_fmt == EC_OUT_BF16_NCHW) {
            __nv_bfloat16 *o = (__nv_bfloat16 *)p.out + (((size_t)slot * 3 + c) * OUT + yo) * OUT + x;
            *reinterpret_cast<__nv_bfloat162 *>(o) = __floats2bfloat162_rn(f0, f1);
        } else {
            const int P = p.patch;
            const size_t row = (size_t)slot * p.G * p.G + (size_t)(yo / P) * p.G + x / P;
            const int col = c * P * P + (yo % P) * P + (x % P);
            __nv_bfloat16 *o = (__nv_bfloat16 *)p.out + row * p.ldk + col;
            *reinterpret_cast<__nv_bfloat162 *>(o) = __floats2bfloat162_rn(f0, f1);
        }
    }
}

__global__ void __launch_bounds__(1024, 1) event2img_kernel(const E2IParams p)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nwarps = NT >> 5;
    const int CS = p.CS;
    const int rank = blockIdx.x % CS;
    const int fid = blockIdx.x / CS;
    const ec_frame fr = p.frames[fid];
    const int H = p.H, W = p.W, RB = p.RB;
    const int y0 = rank * RB;
    const int rows = max(0, min(H, y0 + RB) - y0);
    const int nband = rows * W;
    const long long band_lo = (long long)y0 * W;
    const long long HW = (long long)H * W;
    const bool mask = (p.flags & EC_FLAG_BACKGROUND_MASK) != 0;
    const bool cnz = (p.flags & EC_FLAG_COUNT_NON_ZERO) != 0;
    const int yo_begin = (OUT * rank) / CS, yo_end = (OUT * (rank + 1)) / CS;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);            // RB*W packed bins, later gray bytes
    uint8_t *hrow = smem_raw + (size_t)RB * W * 4;                       // RB*224 horizontally resampled rows
    __shared__ Part part;
    __shared__ unsigned long long red[32][4];
    __shared__ unsigned redm[32];
    __shared__ unsigned s_keep, s_mx;
    __shared__ uint8_t glut[GLUT_N * GLUT_N];
    __shared__ float nlut[768];

    for (int i = tid; i < 768; i += NT) nlut[i] = p.nlut[i];

    // ---- padding frame: the reference pads missing views with zeros (event2img.py:89-91) ----
    if (fr.ev_count <= 0) {
        __syncthreads();
        const int slot = fr.out_slot;
        if (p.out_fmt == EC_OUT_BF16_PATCH) {
            const int P = p.patch, G = p.G;
            const int cols = 3 * P * P;
            for (int r = rank; r < G * G; r += CS) {
                __nv_bfloat16 *o = (__nv_bfloat16 *)p.out + ((size_t)slot * G * G + r) * p.ldk;
                for (int c = tid; c < cols; c += NT) o[c] = __float2bfloat16(0.f);
            }
        } else {
            const size_t n = (size_t)3 * OUT * OUT;
            const size_t b = n * rank / CS, e = n * (rank + 1) / CS;
            if (p.out_fmt == EC_OUT_F32_NCHW) {
                float *o = (float *)p.out + (size_t)slot * n;
                for (size_t i = b + tid; i < e; i += NT) o[i] = 0.f;
            } else {
                __nv_bfloat16 *o = (__nv_bfloat16 *)p.out + (size_t)slot * n;
                for (size_t i = b + tid; i < e; i += NT) o[i] = __float2bfloat16(0.f);
            }
        }
        return;
    }

    // ---- P0: clear the band's bins ----
    for (int i = tid; i < nband; i += NT) hist[i] = 0u;
    __syncthreads();

    // ---- P1: scan events, accumulate the ones that fall into this band ----
    unsigned long long nacc = 0;
    unsigned bad = 0;
    {
        const float4 *ev = p.events + fr.ev_start;
        const int n = fr.ev_count;
        for (int base = 0; base < n; base += NT * 4) {
            float4 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * NT + tid;
                if (i < n) e[u] = ld_stream(ev + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * NT + tid;
                if (i < n) {
                    const int x = __float2int_rz(e[u].x), y = __float2int_rz(e[u].y);
                    const int pol = __float2int_rz(e[u].w);
                    if (pol != 0) {
                        const long long idx = (long long)x + (long long)y * W;   // flat index as np.bincount sees it
                        if (idx < 0 || idx >= HW) {
                            bad = 1;
                        } else {
                            const long long l = idx - band_lo;
                            if (l >= 0 && l < nband) {
                                atomicAdd(&hist[(int)l], pol > 0 ? 1u : 65536u);
                                ++nacc;
                            }
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- P2: exact statistics of the band ----
    {
        unsigned long long s1 = 0, s2 = 0, nnz = 0;
        int32_t *dc = p.dbg_counts ? p.dbg_counts + ((size_t)fid * HW + band_lo) * 2 : nullptr;
        for (int i = tid; i < nband; i += NT) {
            const uint32_t w = hist[i];
            const uint32_t a = w & 0xffffu, b = w >> 16;
            s1 += a + b;
            s2 += (unsigned long long)(a * a) + (unsigned long long)(b * b);
            nnz += (a > 0) + (b > 0);
            if (dc) { dc[2 * i] = (int32_t)a; dc[2 * i + 1] = (int32_t)b; }
        }
        s1 = warp_sum_u64(s1); s2 = warp_sum_u64(s2); nnz = warp_sum_u64(nnz); nacc = warp_sum_u64(nacc);
        bad = __any_sync(0xffffffffu, bad) ? 1u : 0u;
        if (lane == 0) { red[wid][0] = s1; red[wid][1] = s2; red[wid][2] = nnz; red[wid][3] = nacc; redm[wid] = bad; }
        __syncthreads();
        if (tid == 0) {
            Part t = {0, 0, 0, 0, 0, 0};
            for (int w = 0; w < nwarps; ++w) {
                t.s1 += red[w][0]; t.s2 += red[w][1]; t.nnz += red[w][2]; t.nacc += red[w][3]; t.bad |= redm[w];
            }
            part = t;
        }
    }
    cluster.sync();
    if (tid == 0) {
        unsigned long long S1 = 0, S2 = 0, NNZ = 0, NACC = 0;
        unsigned anybad = 0;
        for (int r = 0; r < CS; ++r) {
            const Part *q = cluster.map_shared_rank(&part, r);
            S1 += q->s1; S2 += q->s2; NNZ += q->nnz; NACC += q->nacc; anybad |= q->bad;
        }
        const unsigned long long n = cnz ? NNZ : (unsigned long long)HW * 2ull;
        s_keep = compute_keep(n, S1, S2, 10);
        if (rank == 0) {
            unsigned st = 0;
            if (anybad) st |= EC_STATUS_BAD_COORD;
            if (S1 != NACC) st |= EC_STATUS_COUNT_OVERFLOW;   // a 16-bit field wrapped
            if (st) atomicOr(p.status, (int)st);
        }
    }
    __syncthreads();
    const unsigned keep = s_keep;

    // ---- P3: max of the surviving bins ----
    {
        unsigned mx = 0;
        for (int i = tid; i < nband; i += NT) {
            const uint32_t w = hist[i];
            const uint32_t a = w & 0xffffu, b = w >> 16;
            if (a <= keep) mx = max(mx, a);
            if (b <= keep) mx = max(mx, b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) redm[wid] = mx;
        __syncthreads();
        if (tid == 0) {
            unsigned m = 0;
            for (int w = 0; w < nwarps; ++w) m = max(m, redm[w]);
            part.mx = m;
        }
    }
    cluster.sync();
    if (tid == 0) {
        unsigned m = 0;
        for (int r = 0; r < CS; ++r) m = max(m, cluster.map_shared_rank(&part, r)->mx);
        s_mx = m;
    }
    __syncthreads();
    const unsigned mx = s_mx;

    // ---- P4: gray value per pixel, in place (small counts through a per-frame LUT) ----
    for (int i = tid; i < GLUT_N * GLUT_N; i += NT) glut[i] = (uint8_t)gray_px(i % GLUT_N, i / GLUT_N, mx, mask);
    __syncthreads();
    {
        uint8_t *dg = p.dbg_gray ? p.dbg_gray + (size_t)fid * HW + band_lo : nullptr;
        for (int i = tid; i < nband; i += NT) {
            const uint32_t w = hist[i];
            uint32_t a = w & 0xffffu, b = w >> 16;
            if (a > keep) a = 0;
            if (b > keep) b = 0;
            const unsigned g = (a < GLUT_N && b < GLUT_N) ? glut[b * GLUT_N + a] : gray_px(a, b, mx, mask);
            hist[i] = g;
            if (dg) dg[i] = (uint8_t)g;
        }
    }
    __syncthreads();

    // ---- P5: horizontal Pillow pass (only the 224 columns that survive the centre crop) ----
    {
        const int stride = 2 + p.KH;
        uint32_t *hrow32 = reinterpret_cast<uint32_t *>(hrow);
        const int items = rows * (OUT / 4);
        for (int it = tid; it < items; it += NT) {
            const int y = it / (OUT / 4), xg = it % (OUT / 4);
            const uint32_t *src = hist + y * W;
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int32_t *tab = p.hx + (xg * 4 + j) * stride;
                const int lo = __ldg(tab), cnt = __ldg(tab + 1);
                int ss = 1 << (PREC - 1);
                for (int t = 0; t < cnt; ++t) ss += (int)(src[lo + t] & 0xffu) * __ldg(tab + 2 + t);
                packed |= clip8(ss) << (8 * j);
            }
            hrow32[y * (OUT / 4) + xg] = packed;
        }
    }
    cluster.sync();

    // ---- P6: vertical pass (+DSMEM reads of neighbour bands), normalise, store ----
    {
        const int stride = 2 + p.KV;
        const int items = (yo_end - yo_begin) * (OUT / 2);
        uint8_t *du = p.dbg_u8 ? p.dbg_u8 + (size_t)fid * OUT * OUT : nullptr;
        for (int it = tid; it < items; it += NT) {
            const int yo = yo_begin + it / (OUT / 2), x = (it % (OUT / 2)) * 2;
            const int32_t *tab = p.vy + yo * stride;
            const int lo = __ldg(tab), cnt = __ldg(tab + 1);
            int s0 = 1 << (PREC - 1), s1 = s0;
            for (int t = 0; t < cnt; ++t) {
                const int ys = lo + t;
                const int owner = ys / RB;
                const uint8_t *hr = (owner == rank) ? hrow : cluster.map_shared_rank(hrow, owner);
                const unsigned two = *reinterpret_cast<const uint16_t *>(hr + (ys - owner * RB) * OUT + x);
                const int k = __ldg(tab + 2 + t);
                s0 += (int)(two & 0xffu) * k;
                s1 += (int)(two >> 8) * k;
            }
            const unsigned v0 = clip8(s0), v1 = clip8(s1);
            if (du) { du[yo * OUT + x] = (uint8_t)v0; du[yo * OUT + x + 1] = (uint8_t)v1; }
            store_pair(p, fr.out_slot, yo, x, v0, v1, nlut);
        }
    }
    cluster.sync();   // peers may still be reading this CTA's hrow
}

// ------------------------------------------------------------------------------------------------
// host side: Pillow coefficient tables (ImagingResample precompute_coeffs + normalize_coeffs_8bpc)
// ------------------------------------------------------------------------------------------------
double bicubic(double x)
{
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// table rows for output positions [first, first+224) of an in -> out resample; returns taps per row
int build_axis(int in, int out, int first, std::vector<int32_t> &tab)
{
    const double scale = (double)in / (double)out;
    const double fscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * fscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    tab.assign((size_t)OUT * (2 + ksize), 0);
    std::vector<double> k(ksize);
    for (int i = 0; i < OUT; ++i) {
        int32_t *row = tab.data() + (size_t)i * (2 + ksize);
        if (in == out) {   // Pillow skips a pass that does not change the size: identity tap
            row[0] = first + i; row[1] = 1; row[2] = 1 << PREC;
            continue;
        }
        const int xx = first + i;
        const double center = (xx + 0.5) * scale;
        const double ss = 1.0 / fscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in) xmax = in;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        row[0] = xmin; row[1] = xmax;
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            const double v = k[x] * (double)(1 << PREC);
            row[2 + x] = v < 0 ? (int32_t)(-0.5 + v) : (int32_t)(0.5 + v);
        }
    }
    return ksize;
}

struct Tables {
    int32_t *dev = nullptr;   // hx | vy | nlut(float bits)
    int KH = 0, KV = 0;
    size_t off_vy = 0, off_lut = 0;
};

std::mutex g_mu;
std::map<std::tuple<int, int, int>, Tables> g_tables;

int python_round_half_even(double v) { return (int)std::nearbyint(v); }

int get_tables(int H, int W, cudaStream_t stream, Tables &out)
{
    int dev = 0;
    EC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, H, W);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) { out = it->second; return EC_OK; }
    // torchvision Resize(224): shorter side -> 224, longer side -> int(224 * long / short)
    int Ho, Wo;
    if (W <= H) { Wo = OUT; Ho = (int)(224.0 * H / W); } else { Ho = OUT; Wo = (int)(224.0 * W / H); }
    // CenterCrop(224): offsets int(round((size - 224) / 2.0))
    const int top = python_round_half_even((Ho - OUT) / 2.0), left = python_round_half_even((Wo - OUT) / 2.0);
    std::vector<int32_t> hx, vy;
    Tables t;
    t.KH = build_axis(W, Wo, left, hx);
    t.KV = build_axis(H, Ho, top, vy);
    std::vector<int32_t> all(hx);
    t.off_vy = all.size();
    all.insert(all.end(), vy.begin(), vy.end());
    t.off_lut = all.size();
    // ToTensor (/255 in float32) + Normalize ((x - mean) / std in float32), method.py:17-18 constants
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
    const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) {
            volatile float f = (float)v / 255.0f;
            volatile float d = f - mean[c];
            volatile float r = d / stdv[c];
            float rf = r;
            int32_t bits;
            memcpy(&bits, &rf, 4);
            all.push_back(bits);
        }
    EC_CUDA_CHECK(cudaMalloc(&t.dev, all.size() * sizeof(int32_t)));
    EC_CUDA_CHECK(cudaMemcpyAsync(t.dev, all.data(), all.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    EC_CUDA_CHECK(cudaStreamSynchronize(stream));   // one-time table upload; `all` is a stack-lifetime buffer
    g_tables[key] = t;
    out = t;
    return EC_OK;
}

int geometry(int H, int W, int &CS, int &RB, int &NT, size_t &smem)
{
    const size_t budget = 220 * 1024;   // 227 KB per CTA minus static shared memory
    const size_t per_row = (size_t)W * 4 + OUT;
    const int rb_max = (int)(budget / per_row);
    if (rb_max < 1) return EC_ERR_UNSUPPORTED;
    CS = 1;
    while ((H + CS - 1) / CS > rb_max) CS *= 2;
    if (CS > 8) return EC_ERR_UNSUPPORTED;   // portable cluster limit
    RB = (H + CS - 1) / CS;
    smem = (size_t)RB * per_row;
    smem = (smem + 15) & ~(size_t)15;
    NT = ((size_t)RB * W <= 16384) ? 512 : 1024;
    return EC_OK;
}

}  // namespace

extern "C" int ec_event2img_geometry(int H, int W, int *cluster_size, int *threads, int *smem_bytes)
{
    int CS, RB, NT;
    size_t smem;
    if (H <= 0 || W <= 0 || geometry(H, W, CS, RB, NT, smem) != EC_OK) {
        ec::set_error("ec_event2img_geometry: sensor %dx%d unsupported", H, W);
        return EC_ERR_UNSUPPORTED;
    }
    if (cluster_size) *cluster_size = CS;
    if (threads) *threads = NT;
    if (smem_bytes) *smem_bytes = (int)smem;
    return EC_OK;
}

extern "C" int ec_event2img(const float *events, const ec_frame *frames, int n_frames, int H, int W, int flags,
                            int out_fmt, int patch, int ldk, void *out, int32_t *dbg_counts, uint8_t *dbg_gray,
                            uint8_t *dbg_u8, int32_t *status, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    EC_REQUIRE(n_frames >= 0, "ec_event2img: n_frames < 0");
    if (n_frames == 0) return EC_OK;
    EC_REQUIRE(events && frames && out && status, "ec_event2img: null pointer");
    EC_REQUIRE(((uintptr_t)events & 15) == 0, "ec_event2img: events must be 16-byte aligned");
    EC_REQUIRE(H > 0 && W > 0 && H >= 8 && W >= 8, "ec_event2img: bad sensor shape %dx%d", H, W);
    EC_REQUIRE(out_fmt >= EC_OUT_F32_NCHW && out_fmt <= EC_OUT_BF16_PATCH, "ec_event2img: bad out_fmt %d", out_fmt);
    int G = 0;
    if (out_fmt == EC_OUT_BF16_PATCH) {
        EC_REQUIRE(patch > 0 && patch % 2 == 0 && OUT % patch == 0, "ec_event2img: patch %d must be even and divide 224", patch);
        EC_REQUIRE(ldk >= 3 * patch * patch && ldk % 2 == 0, "ec_event2img: ldk %d too small / odd", ldk);
        G = OUT / patch;
    }
    int CS, RB, NT;
    size_t smem;
    if (geometry(H, W, CS, RB, NT, smem) != EC_OK) {
        ec::set_error("ec_event2img: sensor %dx%d does not fit an 8-CTA cluster", H, W);
        return EC_ERR_UNSUPPORTED;
    }
    Tables tb;
    int rc = get_tables(H, W, stream, tb);
    if (rc != EC_OK) return rc;

    E2IParams p;
    p.events = reinterpret_cast<const float4 *>(events);
    p.frames = frames;
    p.H = H; p.W = W; p.RB = RB; p.CS = CS;
    p.flags = flags; p.out_fmt = out_fmt; p.patch = patch; p.ldk = ldk; p.G = G;
    p.out = out; p.dbg_counts = dbg_counts; p.dbg_gray = dbg_gray; p.dbg_u8 = dbg_u8; p.status = status;
    p.hx = tb.dev; p.vy = tb.dev + tb.off_vy; p.nlut = reinterpret_cast<const float *>(tb.dev + tb.off_lut);
    p.KH = tb.KH; p.KV = tb.KV;

    EC_CUDA_CHECK(cudaFuncSetAttribute(event2img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_frames * CS);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    EC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, event2img_kernel, p));
    return EC_OK;
}

// ------------------------------------------------------------------------------------------------
// host planning: chunk boundaries + view selection
// ------------------------------------------------------------------------------------------------
extern "C" int ec_plan_frames(const int64_t *offsets, int B, int64_t N, int T, const int32_t *sel, int compact,
                              ec_frame *frames, int cap, uint8_t *valid, int32_t *chunks, int *n_frames,
                              int *n_valid)
{
    EC_REQUIRE(offsets && frames && valid && n_frames && n_valid, "ec_plan_frames: null pointer");
    EC_REQUIRE(B >= 0 && N > 0 && T > 0, "ec_plan_frames: bad B/N/T");
    int nf = 0, nv = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t E = offsets[b + 1] - offsets[b];
        // the reference never sees an empty stream (datasets/caltech.py:181-182 resamples) and would fail on t[0]
        EC_REQUIRE(E > 0, "ec_plan_frames: sample %d has no events", b);
        // split_event_count (vis.py:55-72): K chunks, chunk k = [s_k, s_k + len_k)
        int64_t K, full, tail_start = -1;
        if (E < N) { K = 1; full = 0; }
        else {
            const int64_t m = (E + N - 1) / N;     // len(arange(0, E, N))
            full = m - 1;
            const int64_t last = (m - 1) * N;
            const bool tail = (double)(E - last) > (double)N * 0.5;
            K = full + (tail ? 1 : 0);
            if (tail) tail_start = E - N;
        }
        if (chunks) chunks[b] = (int32_t)K;
        const int nvb = (int)(K > T ? T : K);
        for (int t = 0; t < T; ++t) {
            const bool v = t < nvb;
            valid[(size_t)b * T + t] = v ? 1 : 0;
            if (!v && compact) continue;
            if (nf >= cap) { ec::set_error("ec_plan_frames: frame table capacity %d exceeded", cap); return EC_ERR_CAPACITY; }
            ec_frame f;
            f.out_slot = compact ? nv : b * T + t;
            if (v) {
                int64_t k = t;
                if (K > T && sel) k = sel[(size_t)b * T + t];
                EC_REQUIRE(k >= 0 && k < K, "ec_plan_frames: sample %d slot %d selects chunk %lld of %lld", b, t,
                           (long long)k, (long long)K);
                int64_t s, len;
                if (E < N) { s = 0; len = E; }
                else if (k < full) { s = k * N; len = N; }
                else { s = tail_start; len = N; }
                EC_REQUIRE(len <= INT32_MAX, "ec_plan_frames: chunk too long");
                f.ev_start = offsets[b] + s;
                f.ev_count = (int32_t)len;
                ++nv;
            } else {
                f.ev_start = 0;
                f.ev_count = 0;
            }
            frames[nf++] = f;
        }
    }
    *n_frames = nf;
    *n_valid = nv;
    return EC_OK;
}
