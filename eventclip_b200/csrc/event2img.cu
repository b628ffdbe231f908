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
#include <cuda_fp16.h>

#include <map>
#include <tuple>
#include <mutex>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int OUT = 224;               // CLIP input resolution
constexpr int PREC = 22;               // Pillow PRECISION_BITS = 32 - 8 - 2
constexpr int GLUT_N = 32;             // per-frame gray LUT covers pos,neg < 32

struct E2IParams {
    const float4 *events;
    const uint32_t *events_c;   // compact wire format (row F2): one word per event, see ec_pack_events; NULL = float4 events
    const ec_frame *frames;
    int n_frames;
    int H, W, RB, CS;
    int flags, out_fmt, patch, ldk, G;
    void *out;
    int32_t *dbg_counts;
    uint8_t *dbg_gray;
    uint8_t *dbg_u8;
    int32_t *status;
    unsigned long long band_magic; // ceil(2^40 / (RB*W)): flat index -> owning CTA by multiply-shift
    const int32_t *hx;   // [224][2+KH]: lo, cnt, taps  (cropped output columns)
    const int32_t *vy;   // [224][2+KV]: lo, cnt, taps  (cropped output rows)
    const float *nlut;   // [3][256] normalise LUT
    int KH, KV;
    // tensor-core resample (single-CTA frames, W % 4 == 0): int8 A fragments of the tap tables split into three signed
    // base-256 digits, [14 tiles][KS][3 digits][32 lanes][4 regs]; ws = first source column / row of each 16-output tile
    const int32_t *fragH, *fragV, *wsH, *wsV;
    int KSH, KSV;
    int gray_off;   // tensor-core kernel: byte offset of the gray plane in dynamic shared memory
    // bf16 outputs: bf16(fma(v, na[c], nb[c])) equals the bf16 rounding of the exact float32 normalise LUT for all 256 bytes
    // (checked on the host when the tables are built); affine = 0 falls back to the LUT
    float na[3], nb[3];
    int affine;
    int f16;        // 16-bit outputs are fp16 instead of bf16 (EC_OUT_F16_PATCH; out_fmt then reads EC_OUT_BF16_PATCH for the layout)
    int gray;       // EC_OUT_GRAY_*_PATCH: ONE plane per patch row -- the resampled byte itself as an exact 16-bit float (P*P columns);
                    // the three normalised channels are affine in it, so conv1 is folded to K = P*P (clip.py::packed_gray)
    // band-exchange cluster kernel (event2img_big_kernel): byte offsets of its regions in dynamic shared memory, pitches of the
    // gray plane / the transposed horizontal result, events per CTA and exchange round, multiply-shift constant of y / RB
    int big_off_b, big_off_ht, big_gp, big_hpb, big_cap;
    unsigned big_rb_magic;
};

struct Part {            // per-CTA partial statistics exchanged over DSMEM
    unsigned long long s2;   // sum of squared counts
    unsigned nnz;            // non-empty bins
    unsigned nacc;           // events accumulated (= sum of counts)
    unsigned mall;           // largest count (before hot-pixel removal)
    unsigned mx;             // largest surviving count (second exchange, only when needed)
    unsigned flags;          // EC_STATUS_* bits seen by this CTA
};

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

__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_max_u32(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
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

// Same result as compute_keep() for the sizes one CTA histograms (n * S2 < 2^56): 64-bit integers and a float estimate;
// the exact integer comparison walks the estimate to the answer, so the estimate's rounding never matters.
__device__ unsigned compute_keep_small(unsigned long long n, unsigned long long S1, unsigned long long S2, int t)
{
    if (t <= 0 || n == 0) return 0xffffffffu;
    if (__umul64hi(n, S2) != 0 || n * S2 >= (1ull << 56) || n >= (1ull << 31) || t != 10) return compute_keep(n, S1, S2, t);
    const unsigned long long T = 100ull * (n * S2 - S1 * S1);                 // t^2 n^2 var < 2^63
    auto hot = [&](long long c) {
        const long long d = c * (long long)n - (long long)S1;
        if (d <= 0) return false;
        if (d > 3037000499ll) return true;                                       // d^2 >= 2^63 > T
        return (unsigned long long)(d * d) > T;
    };
    const float thr = ((float)S1 + sqrtf((float)T)) / (float)n;                 // mean + 10 std
    long long c = (long long)floorf(thr);
    if (c < 0) c = 0;
    if (c > 0x7fffffffll) c = 0x7fffffffll;
    int it = 0;
    for (; it < 64 && c > 0 && hot(c); ++it) --c;
    for (; it < 128 && !hot(c + 1); ++it) ++c;
    if (it >= 128) return compute_keep(n, S1, S2, t);                            // estimate far off: take the exact slow path
    return c > 0xfffffffell ? 0xfffffffeu : (unsigned)c;
}

// vis.py:27-39 evaluated as numpy >= 2 does (float64); `hist @ cmap` is a 2-term BLAS dot = one fma.
__device__ __noinline__ unsigned gray_px(unsigned pos, unsigned neg, unsigned mx, bool mask)
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

// D (s32, 16x8) = A (s8 coefficient digits, 16x32, row) . B (u8 pixels, 32x8, col) + C (same value in all four slots)
__device__ __forceinline__ void imma_s8u8(int (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1, int c)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "r"(c));
}

__device__ __forceinline__ unsigned clip8(int v)
{
    v >>= PREC;
    return (unsigned)min(max(v, 0), 255);
}


// Eight resampled bytes b as eight 16-bit floats holding b / 128 EXACTLY -- the rows of the gray patch format (the 2^-7 keeps the
// folded conv1 weights at the scale of the original ones, inside fp16's normal range).  fp16: 0x4800 | b = 8 + b / 128 (the ulp of
// [8, 16) is 2^-7), minus 8; bf16: the high half of (65536 + b / 128) - 65536 (float ulp 2^-7 there, 8 significant bits).
__device__ __forceinline__ uint4 gray16x8(uint2 v, int f16)
{
    uint32_t w[4];
    if (f16) {
        const uint32_t k = 0x48004800u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t src = j < 2 ? v.x : v.y;
            const uint32_t h = __byte_perm(src, k, (j & 1) ? 0x5352u : 0x5150u);
            const __half2 r = __hsub2(*reinterpret_cast<const __half2 *>(&h), *reinterpret_cast<const __half2 *>(&k));
            w[j] = *reinterpret_cast<const uint32_t *>(&r);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t src = j < 2 ? v.x : v.y;
            const float a = __uint_as_float(__byte_perm(src, 0x47800000u, 0x7650u + 2 * (j & 1))) - 65536.0f;
            const float b = __uint_as_float(__byte_perm(src, 0x47800000u, 0x7651u + 2 * (j & 1))) - 65536.0f;
            w[j] = __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);      // exact: a byte has at most 8 significant bits
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------
// P1 shared by both kernels: event scan with returning shared-memory atomics
// ------------------------------------------------------------------------------------------------
struct ScanAcc {
    unsigned long long s2;   // sum over events of 2c+1 (c = count before the increment) = sum of squared counts
    unsigned nnz, nacc, mall, flags;
};

// One round = 128 consecutive events per warp, 4 per lane (lane-contiguous 512-byte loads at immediate offsets):
// all loads, then the decode, then all atomics, then the statistics from the values the atomics returned.
// An event becomes (flat bin index l, increment): 1 = positive field, 65536 = negative field, 0 = not histogrammed.
struct ScanRound {      // one lane's four events of a round (float4 rows or compact words)
    float4 e[4];
    uint32_t w[4];
};

template <bool COMPACT, bool TAIL>
__device__ __forceinline__ void load_round(const float4 *ev, const uint32_t *evc, int rem, ScanRound &r)
{
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        r.e[u] = make_float4(0.f, 0.f, 0.f, 0.f);      // a missing tail event decodes to "add 0 to bin 0"
        r.w[u] = 0;
        if (!TAIL || u * 32 < rem) {
            if (COMPACT) r.w[u] = ld_stream_u32(evc + u * 32);
            else r.e[u] = ld_stream(ev + u * 32);
        }
    }
}

template <bool COMPACT, bool MULTI, bool TAIL>
__device__ __forceinline__ void scan_round(const ScanRound &r, int W, unsigned uHW_fast, unsigned uHW, long long HW, uint32_t hist_s,
                                           unsigned long long magic, unsigned bandpx, ScanAcc &acc)
{
    constexpr int U = 4;
    const float4 (&ev4)[4] = r.e;
    const uint32_t (&wc)[4] = r.w;
    unsigned l[U], inc[U], old[U];
    int xs[U], ys[U];
    bool ok[U], all_ok = true;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (COMPACT) {
            // word = flat index (30 bits) | polarity code << 30 (0: p == 0, 1: p > 0, 2: p < 0, 3: index rejected at pack time)
            const unsigned pc = wc[u] >> 30;
            l[u] = wc[u] & 0x3fffffffu;
            inc[u] = pc == 1u ? 1u : (pc == 2u ? 65536u : 0u);
            ok[u] = pc != 3u && l[u] < uHW;
        } else {
            // .astype(int) truncates: trunc(p) != 0 <=> |p| >= 1 (vis.py:44-52); a NaN polarity is dropped like p == 0
            xs[u] = __float2int_rz(ev4[u].x); ys[u] = __float2int_rz(ev4[u].y);
            inc[u] = ev4[u].w >= 1.0f ? 1u : (ev4[u].w <= -1.0f ? 65536u : 0u);
            l[u] = (unsigned)(ys[u] * W + xs[u]);      // flat index as np.bincount sees it; exact in 32 bits for x, y, W < 2^15
            ok[u] = (unsigned)(xs[u] | ys[u]) < 32768u && l[u] < uHW_fast;
        }
        all_ok = all_ok && ok[u];
    }
    if (!all_ok) {      // rare: negative / huge coordinates (the flat index may still be in range), rejected words
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ok[u]) continue;
            bool good = false;
            if (!COMPACT) {
                const long long i64 = (long long)xs[u] + (long long)ys[u] * W;
                good = i64 >= 0 && i64 < HW;
                l[u] = (unsigned)i64;
            }
            if (!good) {
                // p == 0 events are never indexed by the reference; a rejected compact word was a p != 0 event
                if (COMPACT ? (wc[u] >> 30) != 0u : inc[u] != 0u) acc.flags |= EC_STATUS_BAD_COORD;
                inc[u] = 0; l[u] = 0;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        uint32_t addr;
        old[u] = 0;
        if (TAIL && inc[u] == 0u) continue;     // full rounds add 0 to bin 0 instead (no divergence); a tail would hammer that bin
        if (MULTI) {
            // route the event to the CTA that owns its sensor row: one distributed-shared-memory atomic
            const unsigned owner = (unsigned)(((unsigned long long)l[u] * magic) >> 40);   // l / (RB*W), exact for l < 2^24
            const uint32_t local = hist_s + 4u * (l[u] - owner * bandpx);
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(local), "r"(owner));
            asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old[u]) : "r"(addr), "r"(inc[u]) : "memory");
        } else {
            addr = hist_s + 4u * l[u];
            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old[u]) : "r"(addr), "r"(inc[u]) : "memory");
        }
    }
    // adding 1 to a field holding c raises sum(c^2) by 2c+1 and the non-empty count by [c == 0]
    const bool all_valid = inc[0] && inc[1] && inc[2] && inc[3];
    if (__all_sync(0xffffffffu, all_valid)) {
        unsigned cs = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned c = __byte_perm(old[u], 0u, inc[u] == 1u ? 0x4410u : 0x4432u);
            cs += c;
            acc.nnz += c == 0u;
            acc.mall = max(acc.mall, c + 1u);          // 65536 here means the 16-bit field wrapped
        }
        acc.s2 += 2u * cs + U;                         // <= 4 * (2*65535 + 1): no 32-bit overflow within one round
        acc.nacc += U;
    } else {
        unsigned s2p = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned c = __byte_perm(old[u], 0u, inc[u] == 1u ? 0x4410u : 0x4432u);
            const unsigned v = inc[u] != 0u;
            s2p += (2u * c + 1u) * v;
            acc.nnz += (c == 0u) & v;
            acc.mall = max(acc.mall, (c + 1u) * v);
            acc.nacc += v;
        }
        acc.s2 += s2p;
    }
}

// CTA `rank` of a CS-CTA cluster scans its 1/CS slice of the frame's events
template <bool COMPACT, bool MULTI>
__device__ __forceinline__ void scan_events(const E2IParams &p, const ec_frame &fr, int rank, int CS, uint32_t *hist, int tid, int NT,
                                            ScanAcc &acc)
{
    const int per = (fr.ev_count + CS - 1) / CS;
    const int e_lo = min(rank * per, fr.ev_count);
    const int n = min(per, fr.ev_count - e_lo);
    const long long HW = (long long)p.H * p.W;
    const unsigned uHW = HW > 0x7fffffffll ? 0x7fffffffu : (unsigned)HW;
    const unsigned uHW_fast = p.W < 32768 ? uHW : 0u;      // wider sensors take the 64-bit index path
    const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist);
    const unsigned bandpx = (unsigned)(p.RB * p.W);
    const int first = (tid >> 5) * 128 + (tid & 31);       // this lane's first event of a round
    const float4 *ev = p.events + (COMPACT ? 0 : fr.ev_start + e_lo + first);
    const uint32_t *evc = p.events_c + (COMPACT ? fr.ev_start + e_lo + first : 0);
    const int step = NT * 4;
    // software pipeline: the loads of round r+1 are in flight while round r is decoded and histogrammed
    ScanRound cur, nxt;
    if (n >= step) {
        load_round<COMPACT, false>(ev, evc, 0, cur);
        int base = step;
        for (; base + step <= n; base += step) {
            ev += step; evc += step;
            load_round<COMPACT, false>(ev, evc, 0, nxt);
            scan_round<COMPACT, MULTI, false>(cur, p.W, uHW_fast, uHW, HW, hist_s, p.band_magic, bandpx, acc);
            cur = nxt;
        }
        if (base < n) {
            ev += step; evc += step;
            load_round<COMPACT, true>(ev, evc, n - base - first, nxt);
            scan_round<COMPACT, MULTI, false>(cur, p.W, uHW_fast, uHW, HW, hist_s, p.band_magic, bandpx, acc);
            scan_round<COMPACT, MULTI, true>(nxt, p.W, uHW_fast, uHW, HW, hist_s, p.band_magic, bandpx, acc);
        } else {
            scan_round<COMPACT, MULTI, false>(cur, p.W, uHW_fast, uHW, HW, hist_s, p.band_magic, bandpx, acc);
        }
    } else if (n > 0) {
        load_round<COMPACT, true>(ev, evc, n - first, cur);
        scan_round<COMPACT, MULTI, true>(cur, p.W, uHW_fast, uHW, HW, hist_s, p.band_magic, bandpx, acc);
    }
}

// KHMAX: compile-time bound on the horizontal taps (5 when upsampling, 11 for 640 -> 298); 0 = dynamic loop.
template <int KHMAX, bool DBG, bool COMPACT>
__global__ void __launch_bounds__(1024, 1) event2img_kernel(const E2IParams p)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nwarps = NT >> 5;
    const int CS = p.CS;
    const int rank = blockIdx.x % CS;
    const int n_clusters = gridDim.x / CS;        // persistent: each cluster walks over frames fid, fid + n_clusters, ...
    const int H = p.H, W = p.W, RB = p.RB;
    const int y0 = rank * RB;
    const int rows = max(0, min(H, y0 + RB) - y0);
    const int nband = rows * W;
    const long long band_lo = (long long)y0 * W;
    const long long HW = (long long)H * W;
    const bool mask = (p.flags & EC_FLAG_BACKGROUND_MASK) != 0;
    const bool cnz = (p.flags & EC_FLAG_COUNT_NON_ZERO) != 0;
    const int yo_begin = (OUT * rank) / CS, yo_end = (OUT * (rank + 1)) / CS;

    // One buffer, three lives: packed bins (RB*W words) -> gray bytes (first RB*W bytes) + resampled rows after them.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);
    uint8_t *gray = smem_raw;
    uint8_t *hrow = smem_raw + (((size_t)RB * W + 15) & ~(size_t)15);   // RB*224 bytes, inside the bins' footprint
    __shared__ Part part;
    __shared__ unsigned long long red64[32];
    __shared__ unsigned red32[32][4];
    __shared__ unsigned s_keep, s_mx, s_mall;
    __shared__ uint8_t glut[GLUT_N * GLUT_N];
    __shared__ float nlut[768];
    __shared__ uint2 nlut3[256];                 // bf16 (c0, c1, c2, 0) of each uint8 value: one 8-byte load per pixel
    __shared__ int vtab[OUT * (2 + 11)];         // vertical taps (lo, cnt, k[]) when KV <= 11
    __shared__ short ypq[OUT], ypr[OUT], xpq[OUT / 2], xpr[OUT / 2];   // patch coordinates of output rows / column pairs

    for (int i = tid; i < 768; i += NT) nlut[i] = p.nlut[i];
    for (int i = tid; i < 256; i += NT) {
        nlut3[i] = make_uint2(pack16x2(p.nlut[i], p.nlut[256 + i], p.f16), pack16x2(p.nlut[512 + i], 0.f, p.f16));
    }
    const bool vsm = p.KV <= 11;
    if (vsm)
        for (int i = tid; i < OUT * (2 + p.KV); i += NT) vtab[i] = p.vy[i];
    const int *vy = vsm ? vtab : p.vy;
    if (p.out_fmt == EC_OUT_BF16_PATCH) {
        const int P = p.patch;
        for (int i = tid; i < OUT; i += NT) { ypq[i] = (short)(i / P); ypr[i] = (short)(i % P); }
        for (int i = tid; i < OUT / 2; i += NT) { xpq[i] = (short)((2 * i) / P); xpr[i] = (short)((2 * i) % P); }
    }

    for (int fid = blockIdx.x / CS; fid < p.n_frames; fid += n_clusters) {
    const ec_frame fr = p.frames[fid];
    // ---- padding frame: the reference pads missing views with zeros (event2img.py:89-91) ----
    if (fr.ev_count <= 0) {
        const int slot = fr.out_slot;
        if (p.out_fmt == EC_OUT_BF16_PATCH) {
            const int P = p.patch, G = p.G;
            const int cols = (p.gray ? 1 : 3) * P * P;
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
        continue;     // uniform across the cluster; the previous frame ended with a cluster barrier
    }

    // ---- P0: clear the band's bins (16-byte stores; the buffer is padded to 16 bytes) ----
    {
        uint4 *h4 = reinterpret_cast<uint4 *>(hist);
        const int n4 = (nband + 3) >> 2;
        for (int i = tid; i < n4; i += NT) h4[i] = make_uint4(0, 0, 0, 0);
    }
    if (CS > 1) cluster.sync(); else __syncthreads();

    // ---- P1: scan events; returning atomics give the statistics for free:
    //      adding 1 to a bin holding c raises sum(c^2) by 2c+1 and the non-empty count by [c == 0].
    //      In a multi-CTA cluster every CTA scans 1/CS of the frame's events and routes each one to the CTA that
    //      owns its sensor row with a distributed-shared-memory atomic (no event is read twice). ----
    unsigned long long s2 = 0;
    unsigned nnz = 0, nacc = 0, mall = 0, flags = 0;
    {
        ScanAcc acc = {0, 0, 0, 0, 0};
        if (CS > 1) scan_events<COMPACT, true>(p, fr, rank, CS, hist, tid, NT, acc);
        else scan_events<COMPACT, false>(p, fr, 0, 1, hist, tid, NT, acc);
        s2 = acc.s2; nnz = acc.nnz; nacc = acc.nacc; mall = acc.mall; flags = acc.flags;
        if (mall > 0xffffu) flags |= EC_STATUS_COUNT_OVERFLOW;     // a 16-bit field wrapped
    }
    // while this frame is reduced, resampled and stored, pull this CTA's slice of the NEXT frame's events into L2
    if (fid + n_clusters < p.n_frames) {
        const ec_frame nx = p.frames[fid + n_clusters];
        if (nx.ev_count > 0) {
            const int per = (nx.ev_count + CS - 1) / CS;
            const int e_lo = min(rank * per, nx.ev_count);
            const int n = min(per, nx.ev_count - e_lo);
            const char *b = COMPACT ? reinterpret_cast<const char *>(p.events_c + nx.ev_start + e_lo)
                                    : reinterpret_cast<const char *>(p.events + nx.ev_start + e_lo);
            const long long bytes = (long long)n * (COMPACT ? 4 : 16);
            for (long long o = (long long)tid * 128; o < bytes; o += (long long)NT * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
        }
    }
    // block reduction of the partials
    s2 = warp_sum_u64(s2); nnz = warp_sum_u32(nnz); nacc = warp_sum_u32(nacc); mall = warp_max_u32(mall);
    flags |= __shfl_xor_sync(0xffffffffu, flags, 16); flags |= __shfl_xor_sync(0xffffffffu, flags, 8);
    flags |= __shfl_xor_sync(0xffffffffu, flags, 4); flags |= __shfl_xor_sync(0xffffffffu, flags, 2);
    flags |= __shfl_xor_sync(0xffffffffu, flags, 1);
    if (lane == 0) { red64[wid] = s2; red32[wid][0] = nnz; red32[wid][1] = nacc; red32[wid][2] = mall; red32[wid][3] = flags; }
    __syncthreads();   // also: all atomics of this CTA have landed
    if (tid == 0) {
        Part t = {0, 0, 0, 0, 0, 0};
        for (int w = 0; w < nwarps; ++w) {
            t.s2 += red64[w]; t.nnz += red32[w][0]; t.nacc += red32[w][1]; t.mall = max(t.mall, red32[w][2]);
            t.flags |= red32[w][3];
        }
        part = t;
    }
    cluster.sync();
    if (tid == 0) {
        unsigned long long S1 = 0, S2 = 0, NNZ = 0;
        unsigned M = 0, fl = 0;
        for (int r = 0; r < CS; ++r) {
            const Part *q = cluster.map_shared_rank(&part, r);
            S1 += q->nacc; S2 += q->s2; NNZ += q->nnz; M = max(M, q->mall); fl |= q->flags;
        }
        const unsigned long long n = cnz ? NNZ : (unsigned long long)HW * 2ull;
        s_keep = compute_keep(n, S1, S2, 10);
        s_mall = M;
        if (rank == 0 && fl) atomicOr(p.status, (int)fl);
    }
    __syncthreads();
    const unsigned keep = s_keep;
    unsigned mx = s_mall;
    const bool hot = mx > keep;

    // ---- P3 (only when some bin exceeds `keep`): max of the surviving bins ----
    if (hot) {   // uniform across the cluster: s_keep / s_mall derive from the same cluster totals
        unsigned m = 0;
        {
            const uint4 *h4 = reinterpret_cast<const uint4 *>(hist);   // words past nband are zero (P0 cleared them)
            const int n4 = (nband + 3) >> 2;
            for (int i = tid; i < n4; i += NT) {
                const uint4 w4 = h4[i];
                if ((w4.x | w4.y | w4.z | w4.w) == 0) continue;
                const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t a = ws[q] & 0xffffu, b = ws[q] >> 16;
                    if (a <= keep) m = max(m, a);
                    if (b <= keep) m = max(m, b);
                }
            }
        }
        m = warp_max_u32(m);
        if (lane == 0) red32[wid][0] = m;
        __syncthreads();
        if (tid == 0) {
            unsigned mm = 0;
            for (int w = 0; w < nwarps; ++w) mm = max(mm, red32[w][0]);
            part.mx = mm;
        }
        cluster.sync();
        if (tid == 0) {
            unsigned mm = 0;
            for (int r = 0; r < CS; ++r) mm = max(mm, cluster.map_shared_rank(&part, r)->mx);
            s_mx = mm;
        }
        __syncthreads();
        mx = s_mx;
    }

    // ---- P4: gray byte per pixel (small counts through a per-frame LUT), compacted in place ----
    // the hot-pixel cut is folded into the LUT: a count above `keep` looks up the entry of a zero count
    {
        const int side = (int)min(s_mall, (unsigned)(GLUT_N - 1)) + 1;      // pairs up to the largest raw count occur
        for (int k = tid; k < side * side; k += NT) {
            const unsigned ln = k / side, lp = k - ln * side;
            glut[ln * GLUT_N + lp] = (uint8_t)gray_px(lp > keep ? 0u : lp, ln > keep ? 0u : ln, mx, mask);
        }
    }
    __syncthreads();
    {
        // Round r packs pixel groups [r*NT, (r+1)*NT): it reads words [4rNT, 4(r+1)NT) and writes words [rNT, (r+1)NT),
        // which only earlier rounds (or this one, before the barrier) have read.
        const int ng = (nband + 3) >> 2;
        const uint4 *h4 = reinterpret_cast<const uint4 *>(hist);
        uint32_t *g32 = reinterpret_cast<uint32_t *>(gray);
        for (int j0 = 0; j0 < ng; j0 += NT) {
            const int j = j0 + tid;
            uint32_t pk = 0;
            if (j < ng) {
                const uint4 w4 = h4[j];
                const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
                const uint32_t any = w4.x | w4.y | w4.z | w4.w;
                if (!DBG && any == 0) {
                    pk = 0x01010101u * glut[0];          // four empty pixels
                } else if (!DBG && (any & 0xffe0ffe0u) == 0) {
                    // all eight counts < 32: four byte lookups (the LUT already applies the cut)
                    const unsigned g0 = glut[((w4.x >> 11) & 0x3e0u) | w4.x & 0x1fu], g1 = glut[((w4.y >> 11) & 0x3e0u) | w4.y & 0x1fu];
                    const unsigned g2 = glut[((w4.z >> 11) & 0x3e0u) | w4.z & 0x1fu], g3 = glut[((w4.w >> 11) & 0x3e0u) | w4.w & 0x1fu];
                    pk = __byte_perm(__byte_perm(g0, g1, 0x0040), __byte_perm(g2, g3, 0x0040), 0x5410);
                } else
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    uint32_t w = ws[q];
                    if (DBG && p.dbg_counts && 4 * j + q < nband) {
                        int32_t *dc = p.dbg_counts + ((size_t)fid * HW + band_lo + 4 * j + q) * 2;
                        dc[0] = (int32_t)(w & 0xffffu); dc[1] = (int32_t)(w >> 16);
                    }
                    if (hot) {   // some bin exceeds the cut (uniform): remove those fields
                        if ((w & 0xffffu) > keep) w &= 0xffff0000u;
                        if ((w >> 16) > keep) w &= 0x0000ffffu;
                    }
                    unsigned g;
                    if ((w & 0xffe0ffe0u) == 0) g = glut[((w >> 11) & 0x3e0u) | (w & 0x1fu)];   // both counts < 32
                    else g = gray_px(w & 0xffffu, w >> 16, mx, mask);
                    pk |= g << (8 * q);
                }
            }
            __syncthreads();
            if (j < ng) {
                g32[j] = pk;
                if (DBG && p.dbg_gray)
                    for (int q = 0; q < 4 && 4 * j + q < nband; ++q)
                        p.dbg_gray[(size_t)fid * HW + band_lo + 4 * j + q] = (uint8_t)(pk >> (8 * q));
            }
        }
    }
    __syncthreads();

    // ---- P5: horizontal Pillow pass, one thread per cropped output column, taps in registers ----
    {
        const int RG = NT / OUT;
        const int x = tid % OUT, rg = tid / OUT;
        if (rg < RG) {
            const int32_t *tab = p.hx + x * (2 + p.KH);
            const int lo = __ldg(tab), cnt = __ldg(tab + 1);
            if (KHMAX > 0) {
                // taps beyond cnt carry weight 0 and read at most KHMAX-1 bytes past the row: still inside the buffer
                int k[KHMAX > 0 ? KHMAX : 1];
#pragma unroll
                for (int t = 0; t < KHMAX; ++t) k[t] = t < cnt ? __ldg(tab + 2 + t) : 0;
                const uint8_t *src = gray + rg * W + lo;
                uint8_t *dst = hrow + rg * OUT + x;
                const int sstep = RG * W, dstep = RG * OUT;
                for (int y = rg; y < rows; y += RG, src += sstep, dst += dstep) {
                    int ss = 1 << (PREC - 1);
#pragma unroll
                    for (int t = 0; t < KHMAX; ++t) ss += (int)src[t] * k[t];
                    *dst = (uint8_t)clip8(ss);
                }
            } else {
                for (int y = rg; y < rows; y += RG) {
                    const uint8_t *src = gray + y * W + lo;
                    int ss = 1 << (PREC - 1);
                    for (int t = 0; t < cnt; ++t) ss += (int)src[t] * __ldg(tab + 2 + t);
                    hrow[y * OUT + x] = (uint8_t)clip8(ss);
                }
            }
        }
    }
    cluster.sync();

    // ---- P6: vertical pass (+DSMEM reads of neighbour bands), normalise, store ----
    {
        const int stride = 2 + p.KV;
        uint8_t *du = (DBG && p.dbg_u8) ? p.dbg_u8 + (size_t)fid * OUT * OUT : nullptr;
        const int slot = fr.out_slot;
        const bool wide = p.out_fmt != EC_OUT_BF16_PATCH || (p.patch % 8) == 0;
        if (wide) {
            constexpr int NG = OUT / 8;
            const int items = (yo_end - yo_begin) * NG;
            for (int it = tid; it < items; it += NT) {
                const int yo = yo_begin + it / NG, xg = it % NG;
                const int *tab = vy + yo * stride;
                const int lo = tab[0], cnt = tab[1];
                int acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = 1 << (PREC - 1);
                if (CS == 1) {
                    const uint8_t *src = hrow + lo * OUT + xg * 8;
                    for (int t = 0; t < cnt; ++t, src += OUT) {
                        const uint2 v = *reinterpret_cast<const uint2 *>(src);
                        const int k = tab[2 + t];
                        acc[0] += (int)__byte_perm(v.x, 0, 0x4440) * k; acc[1] += (int)__byte_perm(v.x, 0, 0x4441) * k;
                        acc[2] += (int)__byte_perm(v.x, 0, 0x4442) * k; acc[3] += (int)__byte_perm(v.x, 0, 0x4443) * k;
                        acc[4] += (int)__byte_perm(v.y, 0, 0x4440) * k; acc[5] += (int)__byte_perm(v.y, 0, 0x4441) * k;
                        acc[6] += (int)__byte_perm(v.y, 0, 0x4442) * k; acc[7] += (int)__byte_perm(v.y, 0, 0x4443) * k;
                    }
                } else {
                    int own = lo / RB, rin = lo - own * RB;
                    const uint8_t *base = (own == rank) ? hrow : cluster.map_shared_rank(hrow, own);
                    for (int t = 0; t < cnt; ++t) {
                        const uint2 v = *reinterpret_cast<const uint2 *>(base + rin * OUT + xg * 8);
                        const int k = tab[2 + t];
                        acc[0] += (int)__byte_perm(v.x, 0, 0x4440) * k; acc[1] += (int)__byte_perm(v.x, 0, 0x4441) * k;
                        acc[2] += (int)__byte_perm(v.x, 0, 0x4442) * k; acc[3] += (int)__byte_perm(v.x, 0, 0x4443) * k;
                        acc[4] += (int)__byte_perm(v.y, 0, 0x4440) * k; acc[5] += (int)__byte_perm(v.y, 0, 0x4441) * k;
                        acc[6] += (int)__byte_perm(v.y, 0, 0x4442) * k; acc[7] += (int)__byte_perm(v.y, 0, 0x4443) * k;
                        if (++rin == RB) { rin = 0; ++own; base = (own == rank) ? hrow : cluster.map_shared_rank(hrow, own); }
                    }
                }
                unsigned v8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v8[j] = clip8(acc[j]);
                const int x = xg * 8;
                if (du) {
                    uint2 pk;
                    pk.x = v8[0] | (v8[1] << 8) | (v8[2] << 16) | (v8[3] << 24);
                    pk.y = v8[4] | (v8[5] << 8) | (v8[6] << 16) | (v8[7] << 24);
                    *reinterpret_cast<uint2 *>(du + yo * OUT + x) = pk;
                }
                if (p.out_fmt == EC_OUT_F32_NCHW) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float *o = (float *)p.out + (((size_t)slot * 3 + c) * OUT + yo) * OUT + x;
                        const float *nl = nlut + c * 256;
                        *reinterpret_cast<float4 *>(o) = make_float4(nl[v8[0]], nl[v8[1]], nl[v8[2]], nl[v8[3]]);
                        *reinterpret_cast<float4 *>(o + 4) = make_float4(nl[v8[4]], nl[v8[5]], nl[v8[6]], nl[v8[7]]);
                    }
                } else if (p.gray) {
                    uint2 pk;
                    pk.x = v8[0] | (v8[1] << 8) | (v8[2] << 16) | (v8[3] << 24);
                    pk.y = v8[4] | (v8[5] << 8) | (v8[6] << 16) | (v8[7] << 24);
                    const int P = p.patch;
                    const size_t obase = ((size_t)slot * p.G * p.G + (size_t)ypq[yo] * p.G + xpq[x >> 1]) * p.ldk + ypr[yo] * P + xpr[x >> 1];
                    *reinterpret_cast<uint4 *>((__nv_bfloat16 *)p.out + obase) = gray16x8(pk, p.f16);
                } else {
                    // bf16: one 8-byte LUT read per pixel yields all three channels
                    uint2 t8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t8[j] = nlut3[v8[j]];
                    uint4 o0, o1, o2;
                    o0.x = __byte_perm(t8[0].x, t8[1].x, 0x5410); o1.x = __byte_perm(t8[0].x, t8[1].x, 0x7632); o2.x = __byte_perm(t8[0].y, t8[1].y, 0x5410);
                    o0.y = __byte_perm(t8[2].x, t8[3].x, 0x5410); o1.y = __byte_perm(t8[2].x, t8[3].x, 0x7632); o2.y = __byte_perm(t8[2].y, t8[3].y, 0x5410);
                    o0.z = __byte_perm(t8[4].x, t8[5].x, 0x5410); o1.z = __byte_perm(t8[4].x, t8[5].x, 0x7632); o2.z = __byte_perm(t8[4].y, t8[5].y, 0x5410);
                    o0.w = __byte_perm(t8[6].x, t8[7].x, 0x5410); o1.w = __byte_perm(t8[6].x, t8[7].x, 0x7632); o2.w = __byte_perm(t8[6].y, t8[7].y, 0x5410);
                    size_t obase, cstride;
                    if (p.out_fmt == EC_OUT_BF16_NCHW) {
                        obase = (((size_t)slot * 3) * OUT + yo) * OUT + x;
                        cstride = (size_t)OUT * OUT;
                    } else {
                        const int P = p.patch;
                        obase = ((size_t)slot * p.G * p.G + (size_t)ypq[yo] * p.G + xpq[x >> 1]) * p.ldk + ypr[yo] * P + xpr[x >> 1];
                        cstride = (size_t)P * P;
                    }
                    __nv_bfloat16 *ob = (__nv_bfloat16 *)p.out + obase;
                    *reinterpret_cast<uint4 *>(ob) = o0;
                    *reinterpret_cast<uint4 *>(ob + cstride) = o1;
                    *reinterpret_cast<uint4 *>(ob + 2 * cstride) = o2;
                }
            }
        } else {
            // patch size not a multiple of 8 (ViT-L/14): column pairs never straddle a patch
            const int P = p.patch;
            const int items = (yo_end - yo_begin) * (OUT / 2);
            for (int it = tid; it < items; it += NT) {
                const int yo = yo_begin + it / (OUT / 2), xp = it % (OUT / 2), x = 2 * xp;
                const int *tab = vy + yo * stride;
                const int lo = tab[0], cnt = tab[1];
                int s0 = 1 << (PREC - 1), s1 = s0;
                int own = lo / RB, rin = lo - own * RB;
                const uint8_t *base = (own == rank) ? hrow : cluster.map_shared_rank(hrow, own);
                for (int t = 0; t < cnt; ++t) {
                    const unsigned two = *reinterpret_cast<const uint16_t *>(base + rin * OUT + x);
                    const int k = tab[2 + t];
                    s0 += (int)(two & 0xffu) * k;
                    s1 += (int)(two >> 8) * k;
                    if (++rin == RB) { rin = 0; ++own; base = (own == rank) ? hrow : cluster.map_shared_rank(hrow, own); }
                }
                const unsigned v0 = clip8(s0), v1 = clip8(s1);
                if (du) { du[yo * OUT + x] = (uint8_t)v0; du[yo * OUT + x + 1] = (uint8_t)v1; }
                const size_t obase = ((size_t)slot * p.G * p.G + (size_t)ypq[yo] * p.G + xpq[xp]) * p.ldk + ypr[yo] * P + xpr[xp];
                __nv_bfloat16 *ob = (__nv_bfloat16 *)p.out + obase;
                if (p.gray) {
                    *reinterpret_cast<unsigned *>(ob) = gray16x8(make_uint2(v0 | (v1 << 8), 0u), p.f16).x;
                    continue;
                }
                const uint2 t0 = nlut3[v0], t1 = nlut3[v1];
                *reinterpret_cast<unsigned *>(ob) = __byte_perm(t0.x, t1.x, 0x5410);
                *reinterpret_cast<unsigned *>(ob + (size_t)P * P) = __byte_perm(t0.x, t1.x, 0x7632);
                *reinterpret_cast<unsigned *>(ob + 2 * (size_t)P * P) = __byte_perm(t0.y, t1.y, 0x5410);
            }
        }
    }
    cluster.sync();   // peers may still be reading this CTA's hrow; also fences the shared state for the next frame
    }   // frame loop
}

// two fp32 fmas in one instruction (sm_100 packed math): d = a * b + c on both halves
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

// Normalise and store eight consecutive resampled bytes (row yo, columns x .. x+7, x % 8 == 0) in all three channels.
__device__ __forceinline__ void emit8(const E2IParams &p, uint2 v, int yo, int x, char *ofr, uint8_t *du, bool wide, int cstride,
                                      const int *rowoff, const int *coloff, const float *nlut, const uint2 *nlut3)
{
    if (du) *reinterpret_cast<uint2 *>(du + yo * OUT + x) = v;
    if (p.gray) {
        const uint4 o = gray16x8(v, p.f16);
        __nv_bfloat16 *orow = (__nv_bfloat16 *)ofr + rowoff[yo];
        if (wide) *reinterpret_cast<uint4 *>(orow + coloff[x >> 3]) = o;
        else {
            const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint32_t *>(orow + coloff[(x >> 1) + j]) = w4[j];
        }
        return;
    }
    if (p.out_fmt == EC_OUT_F32_NCHW) {
        unsigned v8[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { v8[j] = __byte_perm(v.x, 0u, 0x4440u + j); v8[4 + j] = __byte_perm(v.y, 0u, 0x4440u + j); }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float *o = (float *)ofr + c * OUT * OUT + yo * OUT + x;
            const float *nl = nlut + c * 256;
            *reinterpret_cast<float4 *>(o) = make_float4(nl[v8[0]], nl[v8[1]], nl[v8[2]], nl[v8[3]]);
            *reinterpret_cast<float4 *>(o + 4) = make_float4(nl[v8[4]], nl[v8[5]], nl[v8[6]], nl[v8[7]]);
        }
        return;
    }
    uint32_t w[3][4];      // [channel][pixel pair] as packed bf16
    if (p.affine) {
        // byte -> float through the 2^23 trick, one fma per channel, packed bf16 conversion
        float2 f[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            f[j] = make_float2(__uint_as_float(__byte_perm(v.x, 0x4b000000u, 0x7650u + 2 * j)) - 8388608.0f,
                               __uint_as_float(__byte_perm(v.x, 0x4b000000u, 0x7651u + 2 * j)) - 8388608.0f);
            f[2 + j] = make_float2(__uint_as_float(__byte_perm(v.y, 0x4b000000u, 0x7650u + 2 * j)) - 8388608.0f,
                                   __uint_as_float(__byte_perm(v.y, 0x4b000000u, 0x7651u + 2 * j)) - 8388608.0f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 na = make_float2(p.na[c], p.na[c]), nb = make_float2(p.nb[c], p.nb[c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 r = ffma2(f[j], na, nb);
                w[c][j] = pack16x2(r.x, r.y, p.f16);
            }
        }
    } else {
        uint2 t8[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t8[j] = nlut3[__byte_perm(v.x, 0u, 0x4440u + j)];
            t8[4 + j] = nlut3[__byte_perm(v.y, 0u, 0x4440u + j)];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            w[0][j] = __byte_perm(t8[2 * j].x, t8[2 * j + 1].x, 0x5410);
            w[1][j] = __byte_perm(t8[2 * j].x, t8[2 * j + 1].x, 0x7632);
            w[2][j] = __byte_perm(t8[2 * j].y, t8[2 * j + 1].y, 0x5410);
        }
    }
    __nv_bfloat16 *orow = (__nv_bfloat16 *)ofr + rowoff[yo];
    if (wide) {
        __nv_bfloat16 *ob = orow + coloff[x >> 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) *reinterpret_cast<uint4 *>(ob + c * cstride) = make_uint4(w[c][0], w[c][1], w[c][2], w[c][3]);
    } else {
        // patch size not a multiple of 8 (ViT-L/14): column pairs never straddle a patch
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat16 *ob = orow + coloff[(x >> 1) + j];
#pragma unroll
            for (int c = 0; c < 3; ++c) *reinterpret_cast<uint32_t *>(ob + c * cstride) = w[c][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core variant: one CTA per frame (the whole sensor's bins fit one SM), W % 4 == 0, upsampling or mild
// downsampling (every 16-output tile reads a source window of <= 32 pixels).
//
// Both Pillow passes are banded matrix products evaluated with mma.sync.m16n8k32 (s8 x u8 -> s32):
//     hT[xo][y]   = clip8((sum_x kh[xo][x] gray[y][x] + 2^21) >> 22)      A = kh digits, B = gray rows as stored
//     out[yo][xo] = clip8((sum_y kv[yo][y] hT[xo][y]  + 2^21) >> 22)      A = kv digits, B = rows of the transposed hT
// The 22-bit taps are split into three signed base-256 digits; the three s32 partial sums recombine to exactly Pillow's
// integer accumulator.  Shared memory: [bins | later hT][gray plane].  The vertical pass permutes its B columns so that
// each lane ends up with 8 consecutive output pixels and stores them straight from registers (three 16-byte stores).
// ------------------------------------------------------------------------------------------------
template <bool DBG, bool COMPACT>
__global__ void __launch_bounds__(1024, 1) event2img_tc_kernel(const E2IParams p)
{
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nwarps = NT >> 5;
    const int H = p.H, W = p.W;
    const int HW = H * W;
    const int HP = (H + 3) & ~3;                 // row pitch of the transposed intermediate
    const bool mask = (p.flags & EC_FLAG_BACKGROUND_MASK) != 0;
    const bool cnz = (p.flags & EC_FLAG_COUNT_NON_ZERO) != 0;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);
    uint8_t *hT = smem_raw;                      // [224][HP], over the dead bins
    uint8_t *gray = smem_raw + p.gray_off;       // [H][W] (+ slack for the fragment loads of the last row tile)
    __shared__ unsigned long long red64[32];
    __shared__ unsigned red32[32][4];
    __shared__ unsigned s_keep;
    __shared__ uint8_t glut[GLUT_N * GLUT_N];
    __shared__ float nlut[768];
    __shared__ uint2 nlut3[256];                 // bf16 (c0, c1, c2, 0) of each uint8 value: one 8-byte load per pixel
    __shared__ int rowoff[OUT], coloff[OUT / 2];   // element offset of an output row / of a column group (8 px, or 2 px for P % 8 != 0)
    __shared__ int s_ws[2 * (OUT / 16)];            // first source column / row of each 16-output tile

    constexpr int MT = OUT / 16;
    const int parts = nwarps / MT;                  // warps per 16-output tile (>= 1: the launch uses >= 512 threads)
    const int g = lane >> 2, tig = lane & 3;
    const uint4 *fragH = reinterpret_cast<const uint4 *>(p.fragH) + lane;
    const uint4 *fragV = reinterpret_cast<const uint4 *>(p.fragV) + lane;
    const bool wide = p.out_fmt != EC_OUT_BF16_PATCH || (p.patch % 8) == 0;
    const int cstride = p.out_fmt == EC_OUT_BF16_PATCH ? p.patch * p.patch : OUT * OUT;     // channel stride (elements)
    const int frame_elems = p.out_fmt == EC_OUT_BF16_PATCH ? p.G * p.G * p.ldk : 3 * OUT * OUT;

    if (p.out_fmt == EC_OUT_F32_NCHW)
        for (int i = tid; i < 768; i += NT) nlut[i] = p.nlut[i];
    for (int i = tid; i < 256; i += NT) {
        nlut3[i] = make_uint2(pack16x2(p.nlut[i], p.nlut[256 + i], p.f16), pack16x2(p.nlut[512 + i], 0.f, p.f16));
    }
    {
        const int cw = wide ? 8 : 2;                // pixels per column group
        for (int i = tid; i < OUT; i += NT)
            rowoff[i] = p.out_fmt == EC_OUT_BF16_PATCH ? (i / p.patch) * p.G * p.ldk + (i % p.patch) * p.patch : i * OUT;
        for (int i = tid; i < OUT / cw; i += NT)
            coloff[i] = p.out_fmt == EC_OUT_BF16_PATCH ? ((i * cw) / p.patch) * p.ldk + (i * cw) % p.patch : i * cw;
        for (int i = tid; i < 2 * MT; i += NT) s_ws[i] = i < MT ? p.wsH[i] : p.wsV[i - MT];
    }
    __syncthreads();

    // 16-byte groups of the bins that hT overlays, plus the 32 bytes the last rows' source windows of P6 may read beyond hT's end
    // (zero taps, but the spare warps must not clear them concurrently: compute-sanitizer racecheck)
    const int ht16 = ((OUT * HP + 15) >> 4) + 2;
    bool bins_clean = false;                        // bins above hT already zero (cleared by the spare warps of the last frame)
    for (int fid = blockIdx.x; fid < p.n_frames; fid += gridDim.x) {
        const ec_frame fr = p.frames[fid];
        const int slot = fr.out_slot;
        // ---- padding frame: the reference pads missing views with zeros (event2img.py:89-91) ----
        if (fr.ev_count <= 0) {
            if (p.out_fmt == EC_OUT_BF16_PATCH) {
                const int G = p.G, cols = (p.gray ? 1 : 3) * p.patch * p.patch;
                for (int r = 0; r < G * G; ++r) {
                    __nv_bfloat16 *o = (__nv_bfloat16 *)p.out + ((size_t)slot * G * G + r) * p.ldk;
                    for (int c = tid; c < cols; c += NT) o[c] = __float2bfloat16(0.f);
                }
            } else if (p.out_fmt == EC_OUT_F32_NCHW) {
                float *o = (float *)p.out + (size_t)slot * 3 * OUT * OUT;
                for (int i = tid; i < 3 * OUT * OUT; i += NT) o[i] = 0.f;
            } else {
                __nv_bfloat16 *o = (__nv_bfloat16 *)p.out + (size_t)slot * 3 * OUT * OUT;
                for (int i = tid; i < 3 * OUT * OUT; i += NT) o[i] = __float2bfloat16(0.f);
            }
            continue;     // uniform; the previous frame ended with a barrier
        }

        // ---- P0: clear the bins (W % 4 == 0: whole 16-byte groups).  After the first frame only the part the previous
        //      frame's hT dirtied is left: the warps without matrix work cleared the rest during the vertical pass ----
        {
            uint4 *h4 = reinterpret_cast<uint4 *>(hist);
            const int n16 = bins_clean ? min(ht16, HW >> 2) : (HW >> 2);
#pragma unroll 4
            for (int i = tid; i < n16; i += NT) h4[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();

        // ---- P1: event scan ----
        ScanAcc acc = {0, 0, 0, 0, 0};
        scan_events<COMPACT, false>(p, fr, 0, 1, hist, tid, NT, acc);
        // while this frame is reduced, resampled and stored, pull the next frame's events into L2
        if (fid + (int)gridDim.x < p.n_frames) {
            const ec_frame nx = p.frames[fid + gridDim.x];
            if (nx.ev_count > 0) {
                const char *b = COMPACT ? reinterpret_cast<const char *>(p.events_c + nx.ev_start)
                                        : reinterpret_cast<const char *>(p.events + nx.ev_start);
                const int bytes = nx.ev_count * (COMPACT ? 4 : 16);
                for (int o = tid * 128; o < bytes; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
            }
        }
        // ---- P2: exact integer statistics -> hot-pixel cut (one lane of warp 0), while the other warps already fill the
        //      per-frame gray LUT for the common case that no bin exceeds the cut (max = largest count) ----
        unsigned mx;
        {
            unsigned long long s2 = warp_sum_u64(acc.s2);
            unsigned nnz = warp_sum_u32(acc.nnz), nacc = warp_sum_u32(acc.nacc), mall = warp_max_u32(acc.mall);
            unsigned fl = __reduce_or_sync(0xffffffffu, acc.flags);
            if (lane == 0) { red64[wid] = s2; red32[wid][0] = nnz; red32[wid][1] = nacc; red32[wid][2] = mall; red32[wid][3] = fl; }
            __syncthreads();   // also: all atomics have landed
            const bool on = lane < nwarps;
            mx = warp_max_u32(on ? red32[lane][2] : 0u);
            if (wid == 0) {
                s2 = warp_sum_u64(on ? red64[lane] : 0ull);
                nnz = warp_sum_u32(on ? red32[lane][0] : 0u);
                nacc = warp_sum_u32(on ? red32[lane][1] : 0u);
                fl = __reduce_or_sync(0xffffffffu, on ? red32[lane][3] : 0u);
                if (lane == 0) {
                    if (mx > 0xffffu) fl |= EC_STATUS_COUNT_OVERFLOW;      // a 16-bit field wrapped
                    const unsigned long long n = cnz ? (unsigned long long)nnz : (unsigned long long)HW * 2ull;
                    s_keep = compute_keep_small(n, nacc, s2, 10);
                    if (fl) atomicOr(p.status, (int)fl);
                }
            } else {
                // only (pos, neg) pairs up to the largest count are ever looked up.  This fill bets on "no bin exceeds the cut"
                // (max = largest count); frames with large counts usually lose the bet, so it stops at 8 x 8 entries
                const int side = (int)min(mx, 7u) + 1;
                for (int k = tid - 32; k < side * side; k += NT - 32) {
                    const int neg = k / side, pos = k - neg * side;
                    glut[neg * GLUT_N + pos] = (uint8_t)gray_px(pos, neg, mx, mask);
                }
            }
            __syncthreads();
        }
        const unsigned keep = s_keep;
        const bool hot = mx > keep;
        const unsigned mall = mx;         // largest raw count
        // ---- P3 (only when some bin exceeds `keep`): max of the surviving bins ----
        if (hot) {
            unsigned m = 0;
            const uint4 *h4 = reinterpret_cast<const uint4 *>(hist);
            for (int i = tid; i < (HW >> 2); i += NT) {
                const uint4 w4 = h4[i];
                const unsigned m4 = __vmaxu2(__vmaxu2(w4.x, w4.y), __vmaxu2(w4.z, w4.w));     // per-polarity maxima of the four pixels
                if (m4 == 0) continue;
                const unsigned hi = m4 >> 16, lo = m4 & 0xffffu;
                if (max(hi, lo) <= keep) { m = max(m, max(hi, lo)); continue; }               // no field of the group is cut
                const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t a = ws[q] & 0xffffu, b = ws[q] >> 16;
                    if (a <= keep) m = max(m, a);
                    if (b <= keep) m = max(m, b);
                }
            }
            m = warp_max_u32(m);
            if (lane == 0) red32[wid][0] = m;
            __syncthreads();
            mx = warp_max_u32(lane < nwarps ? red32[lane][0] : 0u);
        }
        // ---- gray LUT for the final maximum: entries up to the cut are computed (fp64), the rest of the square that the raw
        //      counts can index repeats them (a count above `keep` reads as 0).  Frames without a cut only extend the 8 x 8
        //      entries filled above ----
        {
            const int side = (int)min(mall, (unsigned)(GLUT_N - 1)) + 1;
            const int kside = (int)min((unsigned)(side - 1), keep) + 1;                       // distinct surviving counts per axis
            if (hot || side > 8) {
                const int skip = hot ? 0 : 8;                                                 // not hot: [0,8) x [0,8) is already there
                for (int k = tid; k < kside * kside; k += NT) {
                    const int ln = k / kside, lp = k - ln * kside;
                    if (ln >= skip || lp >= skip) glut[ln * GLUT_N + lp] = (uint8_t)gray_px(lp, ln, mx, mask);
                }
                __syncthreads();
                if (kside < side) {
                    for (int k = tid; k < side * side; k += NT) {
                        const int ln = k / side, lp = k - ln * side;
                        if (ln >= kside || lp >= kside) glut[ln * GLUT_N + lp] = glut[(ln >= kside ? 0 : ln) * GLUT_N + (lp >= kside ? 0 : lp)];
                    }
                }
                __syncthreads();
            }
        }

        // ---- P4: gray byte per pixel (small counts through the LUT) into its own plane ----
        {
            const uint4 *h4 = reinterpret_cast<const uint4 *>(hist);
            uint32_t *g32 = reinterpret_cast<uint32_t *>(gray);
            const uint32_t bg4 = 0x01010101u * glut[0];          // four empty pixels
#pragma unroll 2
            for (int j = tid; j < (HW >> 2); j += NT) {
                const uint4 w4 = h4[j];
                const uint32_t any = w4.x | w4.y | w4.z | w4.w;
                uint32_t pk = bg4;
                if (!DBG && (any & 0xffe0ffe0u) == 0) {
                    // all eight counts < 32 (the common case): four byte lookups, no per-pixel branches
                    if (any) {
                        const unsigned g0 = glut[((w4.x >> 11) & 0x3e0u) | w4.x & 0x1fu], g1 = glut[((w4.y >> 11) & 0x3e0u) | w4.y & 0x1fu];
                        const unsigned g2 = glut[((w4.z >> 11) & 0x3e0u) | w4.z & 0x1fu], g3 = glut[((w4.w >> 11) & 0x3e0u) | w4.w & 0x1fu];
                        pk = __byte_perm(__byte_perm(g0, g1, 0x0040), __byte_perm(g2, g3, 0x0040), 0x5410);
                    }
                } else {
                    const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
                    pk = 0;
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w = ws[q];
                        if (DBG && p.dbg_counts) {
                            int32_t *dc = p.dbg_counts + ((size_t)fid * HW + 4 * j + q) * 2;
                            dc[0] = (int32_t)(w & 0xffffu); dc[1] = (int32_t)(w >> 16);
                        }
                        if (hot) {   // some bin exceeds the cut (uniform): remove those fields
                            if ((w & 0xffffu) > keep) w &= 0xffff0000u;
                            if ((w >> 16) > keep) w &= 0x0000ffffu;
                        }
                        unsigned g;
                        if ((w & 0xffe0ffe0u) == 0) g = glut[((w >> 11) & 0x3e0u) | (w & 0x1fu)];   // both counts < 32
                        else g = gray_px(w & 0xffffu, w >> 16, mx, mask);
                        pk |= g << (8 * q);
                    }
                }
                g32[j] = pk;
                if (DBG && p.dbg_gray) *reinterpret_cast<uint32_t *>(p.dbg_gray + (size_t)fid * HW + 4 * j) = pk;
            }
        }
        __syncthreads();

        // ---- P5: horizontal pass on the tensor cores, gray [y][x] -> hT [xo][y].  Warp w owns the 16-column tile
        //      mt = w % 14 (its A fragments stay in registers for the whole pass) and 1 / parts of the 8-row tiles ----
        {
            const int n_tiles = (H + 7) >> 3;
            if (wid < MT * parts) {
                const int mt = wid % MT, part = wid / MT;
                const int ws = s_ws[mt];
                uint4 a[3];
#pragma unroll
                for (int dg = 0; dg < 3; ++dg) a[dg] = __ldg(fragH + (mt * 3 + dg) * 32);
                const int nt0 = part * n_tiles / parts, nt1 = (part + 1) * n_tiles / parts;
                const uint8_t *src = gray + (nt0 * 8 + g) * W + ws + tig * 4;
                uint8_t *dst = hT + (mt * 16 + g) * HP + nt0 * 8 + tig * 2;
                int y = nt0 * 8 + tig * 2;
#pragma unroll 2
                for (int nt = nt0; nt < nt1; ++nt, src += 8 * W, dst += 8, y += 8) {
                    const uint32_t b0 = *reinterpret_cast<const uint32_t *>(src);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t *>(src + 16);
                    int c0[4], c1[4], c2[4];
                    imma_s8u8(c0, a[0], b0, b1, 1 << (PREC - 1));      // the rounding term rides in digit 0
                    imma_s8u8(c1, a[1], b0, b1, 0);
                    imma_s8u8(c2, a[2], b0, b1, 0);
                    unsigned px[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) px[e] = (unsigned)__vimin_s32_relu((c0[e] + (c1[e] << 8) + (c2[e] << 16)) >> PREC, 255);
                    if (y < HP) {      // rows past H (garbage in, zero vertical taps) are only kept inside the pitch
                        *reinterpret_cast<uint16_t *>(dst) = (uint16_t)__byte_perm(px[0], px[1], 0x0040);
                        *reinterpret_cast<uint16_t *>(dst + 8 * HP) = (uint16_t)__byte_perm(px[2], px[3], 0x0040);
                    }
                }
            }
        }
        __syncthreads();

        // ---- P6: vertical pass on the tensor cores fused with the normalisation and the output stores.
        //      Column n of tile step t reads hT row 32 cg + 8 (n >> 1) + 2 t + (n & 1), so after four steps a lane holds
        //      eight consecutive output pixels of rows yo and yo + 8: three 16-byte stores per row, no staging ----
        {
            uint8_t *du = (DBG && p.dbg_u8) ? p.dbg_u8 + (size_t)fid * OUT * OUT : nullptr;
            char *ofr = (char *)p.out + (size_t)slot * frame_elems * (p.out_fmt == EC_OUT_F32_NCHW ? 4 : 2);
            if (wid < MT * parts) {
                const int mt = wid % MT, part = wid / MT;
                const int ws = s_ws[MT + mt];
                uint4 a[3];
#pragma unroll
                for (int dg = 0; dg < 3; ++dg) a[dg] = __ldg(fragV + (mt * 3 + dg) * 32);
                const uint8_t *src0 = hT + (8 * (g >> 1) + (g & 1)) * HP + ws + tig * 4;
                const int yo = mt * 16 + g;
                for (int cg = part; cg < OUT / 32; cg += parts) {
                    uint32_t row_a[2], row_b[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t pa[2], pb[2];
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            const uint8_t *src = src0 + (32 * cg + 2 * (2 * h + t)) * HP;
                            const uint32_t b0 = *reinterpret_cast<const uint32_t *>(src);
                            const uint32_t b1 = *reinterpret_cast<const uint32_t *>(src + 16);
                            int c0[4], c1[4], c2[4];
                            imma_s8u8(c0, a[0], b0, b1, 1 << (PREC - 1));
                            imma_s8u8(c1, a[1], b0, b1, 0);
                            imma_s8u8(c2, a[2], b0, b1, 0);
                            unsigned px[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                px[e] = (unsigned)__vimin_s32_relu((c0[e] + (c1[e] << 8) + (c2[e] << 16)) >> PREC, 255);
                            pa[t] = __byte_perm(px[0], px[1], 0x0040);
                            pb[t] = __byte_perm(px[2], px[3], 0x0040);
                        }
                        row_a[h] = __byte_perm(pa[0], pa[1], 0x5410);
                        row_b[h] = __byte_perm(pb[0], pb[1], 0x5410);
                    }
                    const int x = 32 * cg + 8 * tig;
                    emit8(p, make_uint2(row_a[0], row_a[1]), yo, x, ofr, du, wide, cstride, rowoff, coloff, nlut, nlut3);
                    emit8(p, make_uint2(row_b[0], row_b[1]), yo + 8, x, ofr, du, wide, cstride, rowoff, coloff, nlut, nlut3);
                }
            } else {
                // spare warps: the bins above hT are dead since P4 -- clear them for the next frame now
                uint4 *h4 = reinterpret_cast<uint4 *>(hist);
                const int spare = NT - MT * parts * 32;
                for (int i = ht16 + tid - MT * parts * 32; i < (HW >> 2); i += spare) h4[i] = make_uint4(0, 0, 0, 0);
            }
            bins_clean = nwarps > MT * parts;
        }
        __syncthreads();   // hT shares the bins' storage
    }   // frame loop
}


// ------------------------------------------------------------------------------------------------
// Band-exchange cluster kernel: sensors whose bins do not fit one SM (N-ImageNet 480x640: 1.2 MB of packed bins).
//
// A cluster of CS CTAs owns a frame; CTA `rank` owns the bins of a band of RB sensor rows.  Only LOCAL shared-memory atomics
// touch the bins.  Per frame (and per round of at most CS * cap events):
//   S1  every CTA reads 1 / CS of the events once (16-byte loads), decodes them, and counting-sorts them BY OWNER BAND into a
//       word list in its own shared memory: word = bin index inside the owner's band | polarity << 16.  The events of a
//       round live in registers between the counting pass (one shared-memory atomic per event on a per-warp, per-owner
//       counter, which also yields the event's slot) and the placement pass.
//   S2  every CTA PULLS its segment of each peer's list through distributed shared memory (coalesced 4-byte loads, four in
//       flight per lane) and histograms it with returning local atomics; the returned counts give the exact statistics on the
//       fly: sum of squares, largest count, and how many fields reached 1, 2, ... 8 events -- from which the largest
//       surviving count after the hot-pixel cut follows without another pass over the bins whenever the cut is <= 7.
//   P2  one cluster exchange of the partial statistics -> cut, maximum, per-frame gray LUT
//   P4  gray byte per pixel into a padded plane (over the dead word list)
//   P5  horizontal Pillow pass on the tensor cores (two 32-pixel source windows per 16 outputs), band rows -> hT band
//   P6  vertical pass on the tensor cores: (16 output rows x 32 columns) units spread over all warps of the cluster, source
//       rows fetched from the owners' hT bands through distributed shared memory, normalise + store from registers
// Three cluster barriers per frame; the bins are cleared for the next frame by warps that carry no matrix work.
// A 16-bit field that wraps (>= 65536 events of one polarity on one pixel) marks its pixel as suspect: those pixels (at most
// eight per band) are recounted exactly in 32 bits and the statistics are recomputed from the bins -- a rare path that
// keeps the result identical to the reference's int64 histogram (datasets/vis.py:9-14).
// ------------------------------------------------------------------------------------------------
constexpr int BIG_NT = 1024;
constexpr int BIG_KMAX = 10;                        // events per thread and round (kept in registers between the two S1 passes)
constexpr int BIG_CAP = BIG_NT * BIG_KMAX;          // events per CTA and round
constexpr int BIG_SUS = 8;                          // suspect pixels per band

struct BigPart {                 // per-CTA partial statistics exchanged over DSMEM
    unsigned long long s2;       // sum of squared counts
    unsigned ge[8];              // ge[v-1] = fields whose count reached v (v = 1 .. 8)
    unsigned nacc, mall, flags, nsus, mx, pad;
};

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n)      // PTX shl: amounts >= 32 give 0
{
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}
__device__ __forceinline__ uint32_t map_cluster(uint32_t saddr, int cta)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t caddr)
{
    uint32_t r;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(r) : "r"(caddr) : "memory");
    return r;
}
__device__ __forceinline__ void imma_s8u8_acc(int (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// one event of the frame as (flat bin index, polarity 0 none / 1 positive / 2 negative); false = not histogrammed
template <bool COMPACT>
__device__ __forceinline__ bool decode_event(const E2IParams &p, long long e, long long HW, unsigned &l, unsigned &pol, unsigned &flags)
{
    if (COMPACT) {
        const uint32_t w = ld_stream_u32(p.events_c + e);
        const unsigned pc = w >> 30;
        l = w & 0x3fffffffu;
        pol = pc == 1u ? 1u : (pc == 2u ? 2u : 0u);
        if (pc == 3u || (pc != 0u && (long long)l >= HW)) { flags |= EC_STATUS_BAD_COORD; return false; }
        return pol != 0u;
    }
    const float4 ev = ld_stream(p.events + e);
    const int xs = __float2int_rz(ev.x), ys = __float2int_rz(ev.y);
    pol = ev.w >= 1.0f ? 1u : (ev.w <= -1.0f ? 2u : 0u);
    if (pol == 0u) return false;                 // p == 0 (or NaN): never indexed by the reference
    const long long i64 = (long long)xs + (long long)ys * p.W;
    if (i64 < 0 || i64 >= HW) { flags |= EC_STATUS_BAD_COORD; return false; }
    l = (unsigned)i64;
    return true;
}

template <bool DBG, bool COMPACT>
__global__ void __launch_bounds__(BIG_NT, 1) event2img_big_kernel(const E2IParams p)
{
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int NT = BIG_NT, NW = BIG_NT / 32, MT = OUT / 16;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int CS = p.CS;
    const int rank = blockIdx.x % CS;
    const int n_clusters = gridDim.x / CS;
    const int H = p.H, W = p.W, RB = p.RB, GP = p.big_gp, HPB = p.big_hpb;
    const int y0 = rank * RB;
    const int rows = max(0, min(H, y0 + RB) - y0);
    const int nband = rows * W;
    const unsigned bandpx = (unsigned)(RB * W);
    const long long band_lo = (long long)y0 * W;
    const long long HW = (long long)H * W;
    const bool mask = (p.flags & EC_FLAG_BACKGROUND_MASK) != 0;
    const bool cnz = (p.flags & EC_FLAG_COUNT_NON_ZERO) != 0;
    const int W4 = W >> 2;
    const int n16 = (RB * W) >> 2;                      // 16-byte groups of the bins (RB * W is a multiple of 4)
    const unsigned uHW = (unsigned)HW;                  // < 2^24 (big_plan)
    const unsigned owner_magic = (unsigned)p.band_magic;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *bins = reinterpret_cast<uint32_t *>(smem_raw);
    uint8_t *gray = smem_raw + p.big_off_b;             // [RB rounded up to 8][GP], over the dead word list
    uint32_t *sorted = reinterpret_cast<uint32_t *>(smem_raw + p.big_off_b);
    uint8_t *hT = smem_raw + p.big_off_ht;              // [224][HPB]: the band's rows of the horizontal result, transposed
    __shared__ BigPart part;
    __shared__ uint32_t s_seg[16];                      // (first word, words) of this CTA's list for each owner band
    __shared__ uint32_t s_cnt[NW * 8], s_base[NW * 8];
    __shared__ unsigned long long red64[NW];
    __shared__ unsigned red32[NW][12];
    __shared__ unsigned s_keep, s_mx, s_mall, s_need_pass, s_nsus_tot, s_nacc;
    __shared__ unsigned s_sus[BIG_SUS], s_nsus, s_exact[BIG_SUS][2];
    __shared__ uint8_t glut[GLUT_N * GLUT_N];
    __shared__ uint2 nlut3[256];
    __shared__ int rowoff[OUT], coloff[OUT / 2];
    __shared__ int s_ws[2 * MT];

    const uint32_t bins_s = (uint32_t)__cvta_generic_to_shared(bins);
    const uint32_t sorted_s = (uint32_t)__cvta_generic_to_shared(sorted);
    const uint32_t hT_s = (uint32_t)__cvta_generic_to_shared(hT);
    const uint32_t seg_s = (uint32_t)__cvta_generic_to_shared(s_seg);
    const bool wide = p.out_fmt != EC_OUT_BF16_PATCH || (p.patch % 8) == 0;
    const int cstride = p.out_fmt == EC_OUT_BF16_PATCH ? p.patch * p.patch : OUT * OUT;
    const int frame_elems = p.out_fmt == EC_OUT_BF16_PATCH ? p.G * p.G * p.ldk : 3 * OUT * OUT;

    for (int i = tid; i < 256; i += NT) {
        nlut3[i] = make_uint2(pack16x2(p.nlut[i], p.nlut[256 + i], p.f16), pack16x2(p.nlut[512 + i], 0.f, p.f16));
    }
    {
        const int cw = wide ? 8 : 2;
        for (int i = tid; i < OUT; i += NT)
            rowoff[i] = p.out_fmt == EC_OUT_BF16_PATCH ? (i / p.patch) * p.G * p.ldk + (i % p.patch) * p.patch : i * OUT;
        for (int i = tid; i < OUT / cw; i += NT)
            coloff[i] = p.out_fmt == EC_OUT_BF16_PATCH ? ((i * cw) / p.patch) * p.ldk + (i * cw) % p.patch : i * cw;
        for (int i = tid; i < 2 * MT; i += NT) s_ws[i] = i < MT ? p.wsH[i] : p.wsV[i - MT];
        uint4 *b4 = reinterpret_cast<uint4 *>(bins);
        for (int i = tid; i < n16; i += NT) b4[i] = make_uint4(0, 0, 0, 0);
        if (tid < NW * 8) s_cnt[tid] = 0;
    }
    cluster.sync();        // every CTA of the cluster runs before anyone touches a peer's shared memory

    // block reduction of the per-thread statistics into `part`
    auto publish = [&](unsigned long long s2, unsigned (&ge)[8], unsigned mall, unsigned flags) {
        s2 = warp_sum_u64(s2);
#pragma unroll
        for (int v = 0; v < 8; ++v) ge[v] = warp_sum_u32(ge[v]);
        mall = warp_max_u32(mall);
        flags = __reduce_or_sync(0xffffffffu, flags);
        if (lane == 0) {
            red64[wid] = s2;
#pragma unroll
            for (int v = 0; v < 8; ++v) red32[wid][v] = ge[v];
            red32[wid][8] = mall; red32[wid][9] = flags;
        }
        __syncthreads();
        if (tid < 8) { unsigned a = 0; for (int w = 0; w < NW; ++w) a += red32[w][tid]; part.ge[tid] = a; }
        else if (tid == 8) { unsigned a = 0; for (int w = 0; w < NW; ++w) a = max(a, red32[w][8]); part.mall = a; }
        else if (tid == 9) { unsigned a = 0; for (int w = 0; w < NW; ++w) a |= red32[w][9]; part.flags = a; }
        else if (tid == 10) { unsigned long long a = 0; for (int w = 0; w < NW; ++w) a += red64[w]; part.s2 = a; }
        else if (tid == 11) { part.nacc = s_nacc; part.nsus = s_nsus; }
    };
    // cluster totals -> hot-pixel cut, largest surviving count (or the request for a pass over the bins)
    auto derive = [&]() {
        if (wid == 0) {
            // lane r < CS fetches CTA r's partials over DSMEM (all in flight at once), the warp reduces them
            BigPart q;
            q.s2 = 0; q.nacc = q.mall = q.flags = q.nsus = 0;
#pragma unroll
            for (int v = 0; v < 8; ++v) q.ge[v] = 0;
            if (lane < CS) q = *cluster.map_shared_rank(&part, lane);
            const unsigned long long S2 = warp_sum_u64(q.s2);
            const unsigned long long S1 = warp_sum_u64((unsigned long long)q.nacc);
            unsigned long long GE[9];
#pragma unroll
            for (int v = 0; v < 8; ++v) GE[v] = warp_sum_u64((unsigned long long)q.ge[v]);
            GE[8] = 0;
            const unsigned M = warp_max_u32(q.mall), ns = warp_sum_u32(q.nsus);
            const unsigned fl = __reduce_or_sync(0xffffffffu, q.flags);
            if (lane == 0) {
                const unsigned long long n = cnz ? GE[0] : (unsigned long long)HW * 2ull;
                const unsigned keep = compute_keep_small(n, S1, S2, 10);
                unsigned mx = M, need = 0;
                if (M > keep) {
                    if (keep <= 7u) {
                        mx = 0;
                        for (int v = (int)keep; v >= 1; --v)
                            if (GE[v - 1] > GE[v]) { mx = (unsigned)v; break; }      // some field ended at exactly v
                    } else
                        need = 1;
                }
                s_keep = keep; s_mall = M; s_mx = mx; s_need_pass = need; s_nsus_tot = ns;
                if (rank == 0 && fl) atomicOr(p.status, (int)fl);
            }
        }
        __syncthreads();
    };

    ec_frame fr_next = p.frames[min((int)(blockIdx.x / CS), p.n_frames - 1)];
    for (int fid = blockIdx.x / CS; fid < p.n_frames; fid += n_clusters) {
        const ec_frame fr = fr_next;
        fr_next = p.frames[min(fid + n_clusters, p.n_frames - 1)];       // in flight during this frame
        const int slot = fr.out_slot;
        // ---- padding frame: the reference pads missing views with zeros (event2img.py:89-91) ----
        if (fr.ev_count <= 0) {
            if (p.out_fmt == EC_OUT_BF16_PATCH) {
                const int G = p.G, cols = (p.gray ? 1 : 3) * p.patch * p.patch;
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
            continue;     // uniform across the cluster
        }

        unsigned long long acc_s2 = 0;
        unsigned ge32[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned mall = 0, flags = 0;
        if (tid == 0) { s_nsus = 0; s_nacc = 0; }
        const int EC = fr.ev_count;
        const int round_cap = CS * p.big_cap;
        for (int rb = 0; rb < EC; rb += round_cap) {
            const int n_round = min(round_cap, EC - rb);
            const int per = (n_round + CS - 1) / CS;                 // <= big_cap
            const int e_lo = min(rank * per, n_round);
            const int n = min(per, n_round - e_lo);
            // ================= S1: decode this CTA's slice, count by owner, sort into the word list =================
            const int pw = ((((n + 31) >> 5) + 31) >> 5) << 5;       // events per warp: a multiple of 32, <= 32 * KMAX
            const int w_lo = wid * pw;
            const int w_n = max(0, min(pw, n - w_lo));
            const long long e0 = fr.ev_start + rb + e_lo + w_lo + lane;
            uint32_t pk[BIG_KMAX];     // local bin (16) | negative << 16 | owner << 17 | slot within (warp, owner) << 20 | valid << 31
#pragma unroll
            for (int k0 = 0; k0 < BIG_KMAX; k0 += 5) {
                float4 ev[5];
                uint32_t wc[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int i = 32 * (k0 + j) + lane;
                    ev[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    wc[j] = 0;
                    if (i < w_n) {
                        if (COMPACT) wc[j] = ld_stream_u32(p.events_c + e0 + 32 * (k0 + j));
                        else ev[j] = ld_stream(p.events + e0 + 32 * (k0 + j));
                    }
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int k = k0 + j;
                    pk[k] = 0;
                    if (32 * k < w_n) {                               // warp-uniform
                        unsigned l, neg;
                        bool take;
                        if (COMPACT) {
                            const unsigned pc = wc[j] >> 30;
                            l = wc[j] & 0x3fffffffu;
                            neg = pc == 2u;
                            take = (pc == 1u || pc == 2u) && l < uHW;
                            if (pc == 3u || (pc != 0u && !take)) flags |= EC_STATUS_BAD_COORD;
                        } else {
                            // .astype(int) truncates: trunc(p) != 0 <=> |p| >= 1 (vis.py:44-52); a NaN polarity is dropped like p == 0
                            const int xs = __float2int_rz(ev[j].x), ys = __float2int_rz(ev[j].y);
                            const bool pos = ev[j].w >= 1.0f;
                            neg = ev[j].w <= -1.0f;
                            l = (unsigned)(ys * W + xs);              // flat index as np.bincount sees it; exact in 32 bits for x, y, W < 2^15
                            bool ok = (unsigned)(xs | ys) < 32768u && l < uHW;
                            if (!ok) {                                // rare: negative / huge coordinates whose flat index may still be in range
                                const long long i64 = (long long)xs + (long long)ys * W;
                                ok = i64 >= 0 && i64 < HW;
                                l = (unsigned)i64;
                            }
                            take = ok && (pos || neg);
                            if (!ok && (pos || neg)) flags |= EC_STATUS_BAD_COORD;    // p == 0 events are never range-checked
                        }
                        if (take && 32 * k + lane < w_n) {
                            const unsigned owner = __umulhi(l, owner_magic) >> 4;      // l / (RB * W): ceil(2^36 / d), checked on the host
                            const unsigned local = l - owner * bandpx;
                            const unsigned pos_in = atomicAdd(&s_cnt[wid * 8 + owner], 1u);
                            pk[k] = local | (neg << 16) | (owner << 17) | (pos_in << 20) | 0x80000000u;
                        }
                    }
                }
            }
            __syncthreads();
            if (wid == 0) {
                // one warp turns the 32 x 8 counters into list offsets in owner-major, warp-minor order: lane q covers the eight
                // consecutive entries (owner q / 4, producer warps 8 (q % 4) .. + 7); the counters are left zeroed for the next round
                const int o = lane >> 2, w0 = 8 * (lane & 3);
                unsigned v[8], tot = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { v[i] = s_cnt[(w0 + i) * 8 + o]; s_cnt[(w0 + i) * 8 + o] = 0; tot += v[i]; }
                unsigned incl = tot;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                unsigned run = incl - tot;
                const unsigned ostart = __shfl_sync(0xffffffffu, run, lane & ~3);     // first entry of this owner
                const unsigned oend = __shfl_sync(0xffffffffu, incl, lane | 3);
                if ((lane & 3) == 0) { s_seg[2 * o] = ostart; s_seg[2 * o + 1] = oend - ostart; }
#pragma unroll
                for (int i = 0; i < 8; ++i) { s_base[(w0 + i) * 8 + o] = run; run += v[i]; }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BIG_KMAX; ++k)
                if (pk[k] & 0x80000000u)
                    sorted[s_base[wid * 8 + ((pk[k] >> 17) & 7u)] + ((pk[k] >> 20) & 0x7ffu)] = pk[k] & 0x1ffffu;
            cluster.sync();            // every list and segment table of the cluster is complete

            // ================= S2: pull this band's segments, local returning atomics, statistics on the fly =================
            {
                const int wps = NW / CS;                              // warps per source CTA
                const int src = (rank + (wid % CS)) % CS;
                const int sub = wid / CS;
                const uint32_t rseg = map_cluster(seg_s + 8u * (uint32_t)rank, src);
                const unsigned start = ld_cluster_u32(rseg), len = ld_cluster_u32(rseg + 4u);
                if (sub == 0 && lane == 0) atomicAdd(&s_nacc, len);
                const uint32_t rlist = map_cluster(sorted_s, src) + 4u * start;
                const unsigned stride = (unsigned)wps * 32u;
                const unsigned jb = (unsigned)sub * 32u + (unsigned)lane;
                unsigned ge_lo = 0, ge_hi = 0;
                for (unsigned base = 0; base < len; base += 4u * stride) {
                    uint32_t w[4];
                    bool v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned j = base + (unsigned)u * stride + jb;
                        v[u] = j < len;
                        w[u] = 0;
                        if (v[u]) w[u] = ld_cluster_u32(rlist + 4u * j);
                    }
                    const bool all4 = __all_sync(0xffffffffu, v[0] && v[1] && v[2] && v[3]);
                    unsigned old[4], inc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        inc[u] = v[u] ? 1u + (w[u] >> 16) * 65535u : 0u;
                        old[u] = 0;
                        if (all4 || v[u])
                            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old[u]) : "r"(bins_s + 4u * (w[u] & 0xffffu)), "r"(inc[u]) : "memory");
                    }
                    unsigned s2p = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned c = __byte_perm(old[u], 0u, 0x4410u + (w[u] >> 16) * 0x22u);      // the incremented field
                        const unsigned on = all4 ? 1u : (unsigned)v[u];
                        s2p += (2u * c + 1u) * on;
                        mall = max(mall, (c + 1u) * on);
                        const unsigned sh = 8u * c;
                        ge_lo += shl_clamp(on, sh);                   // fields that reach 1 .. 4 events
                        ge_hi += shl_clamp(on, sh - 32u);             // ... 5 .. 8 (a shift amount outside [0, 32) yields 0)
                        if (on && c == 65535u) {                      // the 16-bit field wrapped: recount this pixel exactly later
                            const unsigned q = atomicAdd(&s_nsus, 1u);
                            if (q < BIG_SUS) s_sus[q] = w[u] & 0xffffu;
                        }
                    }
                    acc_s2 += s2p;
                }
#pragma unroll
                for (int v8 = 0; v8 < 4; ++v8) {
                    ge32[v8] += (ge_lo >> (8 * v8)) & 0xffu;
                    ge32[4 + v8] += (ge_hi >> (8 * v8)) & 0xffu;
                }
            }
            if (rb + round_cap < EC) cluster.sync();                  // peers have consumed this round's list before it is rewritten
        }   // rounds

        // ================= P2: statistics -> cut, maximum =================
        publish(acc_s2, ge32, mall, flags);
        cluster.sync();                // partials visible; also: every peer has finished pulling (the word list is dead)
        derive();
        if (s_nsus_tot) {              // cluster-uniform, rare: some 16-bit field wrapped
            if (tid == 0) {
                const unsigned raw = s_nsus;
                unsigned m = 0;
                for (unsigned i = 0; i < min(raw, (unsigned)BIG_SUS); ++i) {
                    bool dup = false;
                    for (unsigned j = 0; j < m; ++j) dup = dup || s_sus[j] == s_sus[i];
                    if (!dup) s_sus[m++] = s_sus[i];
                }
                s_nsus = m;
                if (raw > (unsigned)BIG_SUS) atomicOr(p.status, EC_STATUS_COUNT_OVERFLOW);     // more wrapped pixels than the side table holds
            }
            if (tid < 2 * BIG_SUS) (&s_exact[0][0])[tid] = 0;
            __syncthreads();
            const int ns = (int)s_nsus;
            if (ns > 0) {
                unsigned fl2 = 0;
                for (long long e = tid; e < EC; e += NT) {
                    unsigned l, pol;
                    if (!decode_event<COMPACT>(p, fr.ev_start + e, HW, l, pol, fl2)) continue;
                    const long long ll = (long long)l - band_lo;
                    if (ll < 0 || ll >= (long long)nband) continue;
                    for (int s = 0; s < ns; ++s)
                        if ((unsigned)ll == s_sus[s]) atomicAdd(&s_exact[s][pol - 1u], 1u);
                }
            }
            __syncthreads();
            unsigned long long s2 = 0;
            unsigned ge[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            unsigned ml = 0;
            for (int i = tid; i < nband; i += NT) {
                const uint32_t w = bins[i];
                unsigned a = w & 0xffffu, b = w >> 16;
                for (int s = 0; s < ns; ++s)
                    if ((unsigned)i == s_sus[s]) { a = s_exact[s][0]; b = s_exact[s][1]; }
                s2 += (unsigned long long)a * a + (unsigned long long)b * b;
                ml = max(ml, max(a, b));
#pragma unroll
                for (int v = 0; v < 8; ++v) ge[v] += (a > (unsigned)v) + (b > (unsigned)v);
            }
            __syncthreads();
            publish(s2, ge, ml, 0u);
            cluster.sync();
            derive();
        }
        const unsigned keep = s_keep;
        const unsigned mall_all = s_mall;
        const bool hot = mall_all > keep;
        const int nsus = s_nsus_tot ? (int)s_nsus : 0;
        if (s_need_pass) {             // cluster-uniform: the cut is above 8 and some bin exceeds it -> max of the survivors from the bins
            unsigned m = 0;
            const uint4 *h4 = reinterpret_cast<const uint4 *>(bins);
            for (int i = tid; i < ((nband + 3) >> 2); i += NT) {
                const uint4 w4 = h4[i];
                const unsigned m4 = __vmaxu2(__vmaxu2(w4.x, w4.y), __vmaxu2(w4.z, w4.w));
                if (m4 == 0) continue;
                const unsigned hi = m4 >> 16, lo = m4 & 0xffffu;
                if (nsus == 0 && max(hi, lo) <= keep) { m = max(m, max(hi, lo)); continue; }
                const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    unsigned a = ws[q] & 0xffffu, b = ws[q] >> 16;
                    for (int s = 0; s < nsus; ++s)
                        if ((unsigned)(4 * i + q) == s_sus[s]) { a = s_exact[s][0]; b = s_exact[s][1]; }
                    if (a <= keep) m = max(m, a);
                    if (b <= keep) m = max(m, b);
                }
            }
            m = warp_max_u32(m);
            if (lane == 0) red32[wid][10] = m;
            __syncthreads();
            if (tid == 0) {
                unsigned mm = 0;
                for (int w = 0; w < NW; ++w) mm = max(mm, red32[w][10]);
                part.mx = mm;
            }
            cluster.sync();
            if (tid == 0) {
                unsigned mm = 0;
                for (int r = 0; r < CS; ++r) mm = max(mm, cluster.map_shared_rank(&part, r)->mx);
                s_mx = mm;
            }
            __syncthreads();
        }
        const unsigned mx = s_mx;

        // ================= P4: per-frame gray LUT (the cut folded in), gray byte per pixel =================
        {
            // entries up to the cut are computed (fp64); the rest of the square the raw counts can index repeats them
            const int side = (int)min(mall_all, (unsigned)(GLUT_N - 1)) + 1;
            const int kside = (int)min((unsigned)(side - 1), keep) + 1;
            for (int k = tid; k < kside * kside; k += NT) {
                const int ln = k / kside, lp = k - ln * kside;
                glut[ln * GLUT_N + lp] = (uint8_t)gray_px(lp, ln, mx, mask);
            }
            __syncthreads();
            if (kside < side)
                for (int k = tid; k < side * side; k += NT) {
                    const int ln = k / side, lp = k - ln * side;
                    if (ln >= kside || lp >= kside) glut[ln * GLUT_N + lp] = glut[(ln >= kside ? 0 : ln) * GLUT_N + (lp >= kside ? 0 : lp)];
                }
        }
        __syncthreads();
        {
            const uint32_t bg4 = 0x01010101u * glut[0];
            for (int y = wid; y < rows; y += NW) {
                const uint4 *h4 = reinterpret_cast<const uint4 *>(bins + y * W);
                uint32_t *g32 = reinterpret_cast<uint32_t *>(gray + y * GP);
                for (int c4 = lane; c4 < W4; c4 += 32) {
                    const uint4 w4 = h4[c4];
                    const uint32_t any = w4.x | w4.y | w4.z | w4.w;
                    uint32_t pkb = bg4;
                    if (!DBG && (any & 0xffe0ffe0u) == 0) {
                        if (any) {
                            const unsigned g0 = glut[((w4.x >> 11) & 0x3e0u) | w4.x & 0x1fu], g1 = glut[((w4.y >> 11) & 0x3e0u) | w4.y & 0x1fu];
                            const unsigned g2 = glut[((w4.z >> 11) & 0x3e0u) | w4.z & 0x1fu], g3 = glut[((w4.w >> 11) & 0x3e0u) | w4.w & 0x1fu];
                            pkb = __byte_perm(__byte_perm(g0, g1, 0x0040), __byte_perm(g2, g3, 0x0040), 0x5410);
                        }
                    } else {
                        const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
                        pkb = 0;
#pragma unroll 1
                        for (int q = 0; q < 4; ++q) {
                            uint32_t w = ws[q];
                            if (DBG && p.dbg_counts) {
                                int32_t *dc = p.dbg_counts + ((size_t)fid * HW + band_lo + (size_t)y * W + 4 * c4 + q) * 2;
                                dc[0] = (int32_t)(w & 0xffffu); dc[1] = (int32_t)(w >> 16);
                            }
                            if (hot) {
                                if ((w & 0xffffu) > keep) w &= 0xffff0000u;
                                if ((w >> 16) > keep) w &= 0x0000ffffu;
                            }
                            unsigned gq;
                            if ((w & 0xffe0ffe0u) == 0) gq = glut[((w >> 11) & 0x3e0u) | (w & 0x1fu)];
                            else gq = gray_px(w & 0xffffu, w >> 16, mx, mask);
                            pkb |= gq << (8 * q);
                        }
                    }
                    g32[c4] = pkb;
                    if (DBG && p.dbg_gray) *reinterpret_cast<uint32_t *>(p.dbg_gray + (size_t)fid * HW + band_lo + (size_t)y * W + 4 * c4) = pkb;
                }
            }
        }
        __syncthreads();
        if (nsus > 0 && tid < nsus) {      // wrapped pixels: exact 32-bit counts
            const unsigned i = s_sus[tid];
            const unsigned a = s_exact[tid][0], b = s_exact[tid][1];
            const unsigned gq = gray_px(a > keep ? 0u : a, b > keep ? 0u : b, mx, mask);
            const unsigned yy = i / (unsigned)W, xx = i - yy * (unsigned)W;
            gray[yy * GP + xx] = (uint8_t)gq;
            if (DBG && p.dbg_counts) {
                int32_t *dc = p.dbg_counts + ((size_t)fid * HW + band_lo + i) * 2;
                dc[0] = (int32_t)a; dc[1] = (int32_t)b;
            }
            if (DBG && p.dbg_gray) p.dbg_gray[(size_t)fid * HW + band_lo + i] = (uint8_t)gq;
        }
        if (nsus > 0) __syncthreads();
        // while this frame is resampled and stored, pull this CTA's first slice of the NEXT frame's events into L2
        if (fid + n_clusters < p.n_frames) {
            const ec_frame nx = fr_next;
            if (nx.ev_count > 0) {
                const int n_round = min(round_cap, nx.ev_count);
                const int per = (n_round + CS - 1) / CS;
                const int e_lo = min(rank * per, n_round);
                const int n = min(per, n_round - e_lo);
                const char *b = COMPACT ? reinterpret_cast<const char *>(p.events_c + nx.ev_start + e_lo)
                                        : reinterpret_cast<const char *>(p.events + nx.ev_start + e_lo);
                const long long bytes = (long long)n * (COMPACT ? 4 : 16);
                for (long long o = (long long)tid * 128; o < bytes; o += (long long)NT * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
            }
        }

        // ================= P5: horizontal pass on the tensor cores, gray [y][x] -> hT [xo][y - y0] =================
        {
            constexpr int HW_WARPS = 2 * MT;           // 28 warps: tile mt = wid % 14, half of the 8-row tiles each
            if (wid < HW_WARPS) {
                const int mt = wid % MT, half = wid / MT;
                const int ws = s_ws[mt];
                const uint4 *fragH = reinterpret_cast<const uint4 *>(p.fragH) + lane;
                uint4 a[2][3];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int dg = 0; dg < 3; ++dg)
                        a[ks][dg] = p.KSH > ks ? __ldg(fragH + ((mt * p.KSH + ks) * 3 + dg) * 32) : make_uint4(0, 0, 0, 0);
                const int n_tiles = (rows + 7) >> 3;
                const int nt0 = half * n_tiles / 2, nt1 = (half + 1) * n_tiles / 2;
                for (int nt = nt0; nt < nt1; ++nt) {
                    const uint8_t *src = gray + (nt * 8 + g) * GP + ws + tig * 4;
                    const uint32_t b00 = *reinterpret_cast<const uint32_t *>(src), b01 = *reinterpret_cast<const uint32_t *>(src + 16);
                    const uint32_t b10 = *reinterpret_cast<const uint32_t *>(src + 32), b11 = *reinterpret_cast<const uint32_t *>(src + 48);
                    int c0[4], c1[4], c2[4];
                    imma_s8u8(c0, a[0][0], b00, b01, 1 << (PREC - 1));
                    imma_s8u8(c1, a[0][1], b00, b01, 0);
                    imma_s8u8(c2, a[0][2], b00, b01, 0);
                    imma_s8u8_acc(c0, a[1][0], b10, b11);
                    imma_s8u8_acc(c1, a[1][1], b10, b11);
                    imma_s8u8_acc(c2, a[1][2], b10, b11);
                    unsigned px[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) px[e] = (unsigned)__vimin_s32_relu((c0[e] + (c1[e] << 8) + (c2[e] << 16)) >> PREC, 255);
                    uint8_t *dst = hT + (mt * 16 + g) * HPB + nt * 8 + tig * 2;
                    *reinterpret_cast<uint16_t *>(dst) = (uint16_t)__byte_perm(px[0], px[1], 0x0040);
                    *reinterpret_cast<uint16_t *>(dst + 8 * HPB) = (uint16_t)__byte_perm(px[2], px[3], 0x0040);
                }
            } else {
                // spare warps: the bins are dead since P4 -- clear the first quarter for the next frame
                uint4 *b4 = reinterpret_cast<uint4 *>(bins);
                for (int i = tid - HW_WARPS * 32; i < (n16 >> 2); i += NT - HW_WARPS * 32) b4[i] = make_uint4(0, 0, 0, 0);
            }
        }
        cluster.sync();                // every band of hT is complete

        // ================= P6: vertical pass on the tensor cores (source rows over DSMEM), normalise, store =================
        {
            uint8_t *du = (DBG && p.dbg_u8) ? p.dbg_u8 + (size_t)fid * OUT * OUT : nullptr;
            char *ofr = (char *)p.out + (size_t)slot * frame_elems * (p.out_fmt == EC_OUT_F32_NCHW ? 4 : 2);
            constexpr int UNITS = MT * (OUT / 32);     // (16 output rows) x (32 output columns)
            const int my_units = rank < UNITS ? (UNITS - rank + CS - 1) / CS : 0;     // units u = rank + CS * k, one per warp k
            const int unit_warps = min(my_units, NW);
            if (wid < unit_warps) {
                const uint4 *fragV = reinterpret_cast<const uint4 *>(p.fragV) + lane;
                for (int u = rank + CS * wid; u < UNITS; u += CS * NW) {
                    const int mt = u % MT, cgp = u / MT;
                    const int ws = s_ws[MT + mt];
                    uint4 a[2][3];
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int dg = 0; dg < 3; ++dg)
                            a[ks][dg] = p.KSV > ks ? __ldg(fragV + ((mt * p.KSV + ks) * 3 + dg) * 32) : make_uint4(0, 0, 0, 0);
                    // the lane's four 4-byte source groups of a K step pair: sensor rows ws + 32 ks + 16 h + 4 tig .. + 3
                    uint32_t raddr[4];
                    bool rok[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const unsigned y = (unsigned)(ws + 16 * q + 4 * tig);
                        const unsigned band = (y * p.big_rb_magic) >> 20;             // y / RB
                        rok[q] = band < (unsigned)CS;
                        raddr[q] = rok[q] ? map_cluster(hT_s + (y - band * (unsigned)RB), (int)band) : 0u;
                    }
                    const int yo = mt * 16 + g;
                    uint32_t row_a[2], row_b[2];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t pa[2], pb[2];
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            const uint32_t off = (uint32_t)((32 * cgp + 8 * (g >> 1) + (g & 1) + 2 * (2 * hh + t)) * HPB);
                            uint32_t b[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) b[q] = rok[q] ? ld_cluster_u32(raddr[q] + off) : 0u;
                            int c0[4], c1[4], c2[4];
                            imma_s8u8(c0, a[0][0], b[0], b[1], 1 << (PREC - 1));
                            imma_s8u8(c1, a[0][1], b[0], b[1], 0);
                            imma_s8u8(c2, a[0][2], b[0], b[1], 0);
                            imma_s8u8_acc(c0, a[1][0], b[2], b[3]);
                            imma_s8u8_acc(c1, a[1][1], b[2], b[3]);
                            imma_s8u8_acc(c2, a[1][2], b[2], b[3]);
                            unsigned px[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                px[e] = (unsigned)__vimin_s32_relu((c0[e] + (c1[e] << 8) + (c2[e] << 16)) >> PREC, 255);
                            pa[t] = __byte_perm(px[0], px[1], 0x0040);
                            pb[t] = __byte_perm(px[2], px[3], 0x0040);
                        }
                        row_a[hh] = __byte_perm(pa[0], pa[1], 0x5410);
                        row_b[hh] = __byte_perm(pb[0], pb[1], 0x5410);
                    }
                    const int x = 32 * cgp + 8 * tig;
                    emit8(p, make_uint2(row_a[0], row_a[1]), yo, x, ofr, du, wide, cstride, rowoff, coloff, p.nlut, nlut3);
                    emit8(p, make_uint2(row_b[0], row_b[1]), yo + 8, x, ofr, du, wide, cstride, rowoff, coloff, p.nlut, nlut3);
                }
            } else {
                // warps without a unit clear the rest of the bins
                uint4 *b4 = reinterpret_cast<uint4 *>(bins);
                const int spare = NT - unit_warps * 32;
                for (int i = (n16 >> 2) + tid - unit_warps * 32; i < n16; i += spare) b4[i] = make_uint4(0, 0, 0, 0);
            }
            if (unit_warps == NW) {    // no spare warp (tiny clusters): everybody clears after the units
                __syncthreads();
                uint4 *b4 = reinterpret_cast<uint4 *>(bins);
                for (int i = (n16 >> 2) + tid; i < n16; i += NT) b4[i] = make_uint4(0, 0, 0, 0);
            }
        }
        // no barrier here: the next frame's S1 only writes the word list (this CTA's P5 has read the gray plane), its
        // barriers order the bin clearing before S2, and peers reach the next cluster barrier only after their P6
    }   // frame loop
    cluster.sync();                    // no CTA leaves while a peer may still read its shared memory
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
    int32_t *dev = nullptr;   // hx | vy | nlut(float bits) | fragH | fragV | wsH | wsV
    int KH = 0, KV = 0;
    size_t off_vy = 0, off_lut = 0;
    size_t off_fragH = 0, off_fragV = 0, off_wsH = 0, off_wsV = 0;
    int KSH = 0, KSV = 0;
    float na[3] = {0, 0, 0}, nb[3] = {0, 0, 0};
    int affine = 0;
    float na16[3] = {0, 0, 0}, nb16[3] = {0, 0, 0};      // the same for fp16 outputs
    int affine16 = 0;
};

// signed base-256 digits of a Pillow coefficient (|k| < 2^23): k = d0 + 256 d1 + 65536 d2, each in [-128, 127]
void digits3(int32_t k, int d[3])
{
    d[0] = ((k + 128) & 255) - 128;
    const int32_t k1 = (k - d[0]) >> 8;
    d[1] = ((k1 + 128) & 255) - 128;
    d[2] = (k1 - d[1]) >> 8;
}

// A fragments of mma.sync.m16n8k32 (row-major s8) for every 16-output tile of an axis table (lo, cnt, taps):
// register q of lane l holds A[row = l/4 + 8*(q&1)][k = (l%4)*4 + 16*(q>>1) .. +3], lowest k in the lowest byte.
// Returns the K steps (32 source positions each) a tile needs, or 0 when a coefficient does not fit three digits.
int build_frags(const std::vector<int32_t> &axis, int ksize, std::vector<int32_t> &frag, std::vector<int32_t> &ws)
{
    const int stride = 2 + ksize, tiles = OUT / 16;
    ws.assign(tiles, 0);
    int KS = 1;
    for (int mt = 0; mt < tiles; ++mt) {
        int lo_min = 1 << 30, hi = 0;
        for (int r = 0; r < 16; ++r) {
            const int32_t *row = axis.data() + (size_t)(mt * 16 + r) * stride;
            lo_min = std::min(lo_min, (int)row[0]);
            hi = std::max(hi, (int)(row[0] + row[1]));
        }
        ws[mt] = lo_min & ~3;
        KS = std::max(KS, (hi - ws[mt] + 31) / 32);
    }
    frag.assign((size_t)tiles * KS * 3 * 32 * 4, 0);
    for (int mt = 0; mt < tiles; ++mt)
        for (int ks = 0; ks < KS; ++ks)
            for (int lane = 0; lane < 32; ++lane)
                for (int q = 0; q < 4; ++q) {
                    const int r = lane / 4 + 8 * (q & 1);
                    const int32_t *row = axis.data() + (size_t)(mt * 16 + r) * stride;
                    uint32_t w[3] = {0, 0, 0};
                    for (int b = 0; b < 4; ++b) {
                        const int src = ws[mt] + ks * 32 + (lane % 4) * 4 + 16 * (q >> 1) + b;
                        int32_t k = 0;
                        if (src >= row[0] && src < row[0] + row[1]) k = row[2 + src - row[0]];
                        if (k <= -(1 << 23) || k >= (1 << 23)) return 0;
                        int d[3];
                        digits3(k, d);
                        for (int dg = 0; dg < 3; ++dg) w[dg] |= (uint32_t)(d[dg] & 255) << (8 * b);
                    }
                    for (int dg = 0; dg < 3; ++dg)
                        frag[((((size_t)mt * KS + ks) * 3 + dg) * 32 + lane) * 4 + q] = (int32_t)w[dg];
                }
    return KS;
}

// torchvision Resize(224) + CenterCrop(224) as the two cropped axis tables
void resample_axes(int H, int W, std::vector<int32_t> &hx, std::vector<int32_t> &vy, int &KH, int &KV)
{
    // Resize(224): shorter side -> 224, longer side -> int(224 * long / short)
    int Ho, Wo;
    if (W <= H) { Wo = OUT; Ho = (int)(224.0 * H / W); } else { Ho = OUT; Wo = (int)(224.0 * W / H); }
    // CenterCrop(224): offsets int(round((size - 224) / 2.0)), Python's round-half-even
    const int top = (int)std::nearbyint((Ho - OUT) / 2.0), left = (int)std::nearbyint((Wo - OUT) / 2.0);
    KH = build_axis(W, Wo, left, hx);
    KV = build_axis(H, Ho, top, vy);
}

// Tensor-core kernel: whole sensor in one CTA, 4-byte aligned rows, one 32-wide source window per 16-output tile
// (KSH == KSV == 1).  Returns whether it applies and its shared-memory layout.
bool tc_plan(int H, int W, int CS, int KSH, int KSV, size_t &smem, int &gray_off)
{
    static const bool off = getenv("EC_E2I_TC") && atoi(getenv("EC_E2I_TC")) == 0;    // 0 forces the SIMT kernel
    const int HP = (H + 3) & ~3;
    size_t region_a = (size_t)H * W * 4;                       // bins, later the transposed horizontal result
    if (region_a < (size_t)OUT * HP) region_a = (size_t)OUT * HP;
    region_a = (region_a + 15) & ~(size_t)15;
    const size_t gray_bytes = ((size_t)(H + 8) * W + 32 + 15) & ~(size_t)15;       // last row tile reads up to 7 rows + 31 bytes past the plane
    if (off || CS != 1 || W % 4 != 0 || H < 16 || KSH != 1 || KSV != 1 || region_a + gray_bytes > (size_t)218 * 1024) return false;     // + 8.4 KB of static tables <= 227 KB
    smem = region_a + gray_bytes;
    gray_off = (int)region_a;
    return true;
}

// Band-exchange cluster kernel: the bins do not fit one CTA, rows are 4-byte aligned, every 16-output tile of both axes reads
// a source window of <= 64 pixels (two K steps), a band holds <= 65536 bins.  Returns whether it applies, the cluster size,
// the band height (a multiple of 4 rows, so that a 4-byte group of the transposed intermediate never straddles two bands)
// and the shared-memory layout.  EC_E2I_BIG=0 disables it, EC_E2I_BIG=force also takes sensors that fit one CTA (tests of the
// 2- and 4-CTA paths), EC_E2I_BIG_CAP lowers the events per CTA and round (tests of the multi-round exchange).
struct BigPlan {
    int CS = 0, RB = 0, gp = 0, hpb = 0, off_b = 0, off_ht = 0, cap = 0;
    unsigned rb_magic = 0;
    size_t smem = 0;
};

bool big_plan(int H, int W, int KSH, int KSV, bool fits_one_cta, BigPlan &bp)
{
    const char *mode = getenv("EC_E2I_BIG");
    const bool force = mode && !strcmp(mode, "force");
    if (mode && atoi(mode) == 0 && !force) return false;
    if (fits_one_cta && !force) return false;
    if (W % 4 != 0 || H < 16 || W < 64 || KSH < 1 || KSH > 2 || KSV < 1 || KSV > 2 || (long long)H * W >= (1 << 24)) return false;
    const size_t limit = (size_t)227 * 1024 - 10 * 1024;       // static shared memory of the kernel: 9.5 KB
    for (int CS = 2; CS <= 8; CS *= 2) {
        const int RB = (((H + CS - 1) / CS) + 3) & ~3;
        if ((long long)RB * (CS - 1) >= H) continue;           // the last band would be empty
        if ((long long)RB * W > 65536) continue;
        const int RB8 = (RB + 7) & ~7;
        int gpw = W / 4;
        while (gpw % 8 != 4) ++gpw;                             // pitch = 4 (mod 8) words: the eight rows of a B fragment hit distinct banks
        const int gp = 4 * gpw;
        const int hpb = RB8 + 4;
        const size_t bins = (size_t)RB * W * 4;
        size_t regb = (size_t)RB8 * gp + 128;
        int cap = BIG_CAP;
        if (const char *e = getenv("EC_E2I_BIG_CAP")) { const int c = atoi(e); if (c >= 32 && c < cap) cap = c; }
        if (regb < (size_t)cap * 4) {
            // a small band: the word list does not fit the gray plane's footprint -- shrink the round rather than grow the region
            // beyond what the gray plane needs, unless shared memory is plentiful
            const size_t want = (size_t)cap * 4;
            if (bins + want + (size_t)OUT * hpb + 64 <= limit) regb = want;
            else cap = (int)(regb / 4);
        }
        regb = (regb + 15) & ~(size_t)15;
        const size_t ht = ((size_t)OUT * hpb + 15) & ~(size_t)15;
        if (bins + regb + ht > limit) continue;
        bp.CS = CS; bp.RB = RB; bp.gp = gp; bp.hpb = hpb; bp.cap = cap;
        bp.off_b = (int)bins; bp.off_ht = (int)(bins + regb);
        bp.smem = bins + regb + ht;
        bp.rb_magic = (unsigned)(((1u << 20) + (unsigned)RB - 1) / (unsigned)RB);
        for (unsigned y = 0; y < (unsigned)(CS * RB + 128); ++y)                 // the multiply-shift division is exact where it is used
            if (((y * bp.rb_magic) >> 20) != y / (unsigned)RB) return false;
        const unsigned long long d = (unsigned long long)RB * W, magic = ((1ull << 36) + d - 1) / d;     // owner = (l * magic) >> 36
        if (magic >> 32) return false;
        for (unsigned long long k = 1; k <= (unsigned long long)CS; ++k)
            if ((((k * d - 1) * magic) >> 36) != k - 1 || (((k * d) * magic) >> 36) != k) return false;
        return true;
    }
    return false;
}

std::mutex g_mu;
std::map<std::tuple<int, int, int>, Tables> g_tables;

int get_tables(int H, int W, cudaStream_t stream, Tables &out)
{
    int dev = 0;
    EC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, H, W);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) { out = it->second; return EC_OK; }
    std::vector<int32_t> hx, vy;
    Tables t;
    resample_axes(H, W, hx, vy, t.KH, t.KV);
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
    // one fma per channel that reproduces the 16-bit rounding (bf16, fp16) of the LUT above for every byte value (else: keep the LUT)
    {
        auto bf16_bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return (uint32_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16); };
        auto f16_bits = [](float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); };
        for (int fmt = 0; fmt < 2; ++fmt) {
            int ok_all = 1;
            float *pa = fmt ? t.na16 : t.na, *pb = fmt ? t.nb16 : t.nb;
            for (int c = 0; c < 3 && ok_all; ++c) {
                const float a0 = (float)(1.0 / (255.0 * (double)stdv[c])), b0 = (float)(-(double)mean[c] / (double)stdv[c]);
                bool found = false;
                for (int r = 0; r <= 4 && !found; ++r)          // nudge a / b by a few ulps if the nominal pair misses a rounding boundary
                    for (int da = -r; da <= r && !found; ++da)
                        for (int db = -r; db <= r && !found; ++db) {
                            if (std::max(std::abs(da), std::abs(db)) != r) continue;
                            float a = a0, b = b0;
                            for (int k = 0; k < std::abs(da); ++k) a = std::nextafterf(a, da > 0 ? INFINITY : -INFINITY);
                            for (int k = 0; k < std::abs(db); ++k) b = std::nextafterf(b, db > 0 ? INFINITY : -INFINITY);
                            bool ok = true;
                            for (int v = 0; v < 256 && ok; ++v) {
                                float lut;
                                memcpy(&lut, &all[t.off_lut + (size_t)c * 256 + v], 4);
                                const float y = fmaf((float)v, a, b);
                                ok = fmt ? f16_bits(y) == f16_bits(lut) : bf16_bits(y) == bf16_bits(lut);
                            }
                            if (ok) { pa[c] = a; pb[c] = b; found = true; }
                        }
                if (!found) ok_all = 0;
            }
            if (getenv("EC_E2I_AFFINE") && atoi(getenv("EC_E2I_AFFINE")) == 0) ok_all = 0;     // experiment hook
            (fmt ? t.affine16 : t.affine) = ok_all;
        }
    }
    {
        std::vector<int32_t> fh, fv, wh, wv;
        t.KSH = build_frags(hx, t.KH, fh, wh);
        t.KSV = build_frags(vy, t.KV, fv, wv);
        while (all.size() % 4) all.push_back(0);      // fragments are read as 16-byte vectors
        t.off_fragH = all.size(); all.insert(all.end(), fh.begin(), fh.end());
        t.off_fragV = all.size(); all.insert(all.end(), fv.begin(), fv.end());
        t.off_wsH = all.size(); all.insert(all.end(), wh.begin(), wh.end());
        t.off_wsV = all.size(); all.insert(all.end(), wv.begin(), wv.end());
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
    const size_t budget = 206 * 1024;   // 227 KB per CTA minus ~20 KB of static shared memory (LUTs, tap tables)
    // bins (4 bytes/pixel) are later reused as gray bytes (1 byte/pixel) followed by the 224-byte resampled rows
    const size_t per_row = (size_t)W * 4 > (size_t)W + OUT + 16 ? (size_t)W * 4 : (size_t)W + OUT + 16;
    const int rb_max = (int)(budget / per_row);
    if (rb_max < 1) return EC_ERR_UNSUPPORTED;
    CS = 1;
    while ((H + CS - 1) / CS > rb_max) CS *= 2;
    if (CS > 8) return EC_ERR_UNSUPPORTED;   // portable cluster limit
    RB = (H + CS - 1) / CS;
    smem = (size_t)RB * per_row;
    smem = (smem + 31) & ~(size_t)15;
    NT = ((size_t)RB * W <= 16384) ? 512 : 1024;
    // experiment hooks (profiling only): EC_E2I_CS forces a larger cluster, EC_E2I_NT the block size
    if (const char *e = getenv("EC_E2I_CS")) {
        const int want = atoi(e);
        if (want > CS && want <= 8 && (want & (want - 1)) == 0) {
            CS = want;
            RB = (H + CS - 1) / CS;
            smem = ((size_t)RB * per_row + 31) & ~(size_t)15;
        }
    }
    if (const char *e = getenv("EC_E2I_NT")) {
        const int want = atoi(e);
        if (want == 256 || want == 512 || want == 1024) NT = want;
    }
    return EC_OK;
}

}  // namespace

extern "C" int ec_event2img_geometry(int H, int W, int *cluster_size, int *threads, int *smem_bytes)
{
    int CS, RB, NT;
    size_t smem;
    if (H < 8 || W < 8 || geometry(H, W, CS, RB, NT, smem) != EC_OK) {
        ec::set_error("ec_event2img_geometry: sensor %dx%d unsupported", H, W);
        return EC_ERR_UNSUPPORTED;
    }
    {      // same choice ec_event2img makes: the tensor-core kernel when its tiling applies, the band-exchange kernel for clusters
        std::vector<int32_t> hx, vy, frag, ws;
        int KH, KV, gray_off;
        resample_axes(H, W, hx, vy, KH, KV);
        const int KSH = build_frags(hx, KH, frag, ws), KSV = build_frags(vy, KV, frag, ws);
        size_t tsmem = smem;
        const bool tc = CS == 1 && tc_plan(H, W, CS, KSH, KSV, tsmem, gray_off);
        BigPlan bp;
        if (big_plan(H, W, KSH, KSV, CS == 1, bp)) { CS = bp.CS; NT = BIG_NT; smem = bp.smem; }
        else if (tc) { smem = tsmem; if (NT < 512) NT = 512; }
    }
    if (cluster_size) *cluster_size = CS;
    if (threads) *threads = NT;
    if (smem_bytes) *smem_bytes = (int)smem;
    return EC_OK;
}

// compact wire format of row F2: word = flat pixel index as np.bincount sees it (x + y*W, 30 bits) | polarity code << 30
__global__ void __launch_bounds__(256) pack_events_kernel(const float4 *__restrict__ ev, int64_t n, long long HW, int W,
                                                          uint32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 e = ld_stream(ev + i);
    const int x = __float2int_rz(e.x), y = __float2int_rz(e.y), pol = __float2int_rz(e.w);
    uint32_t w = 0;                                    // p == 0: never histogrammed, never range-checked (vis.py:9-14)
    if (pol != 0) {
        const long long l = (long long)x + (long long)y * W;
        w = (l >= 0 && l < HW) ? ((uint32_t)l | (pol > 0 ? 1u << 30 : 2u << 30)) : (3u << 30);
    }
    out[i] = w;
}

static int event2img_impl(const float *events, const uint32_t *events_c, const ec_frame *frames, int n_frames, int H, int W, int flags,
                          int out_fmt, int patch, int ldk, void *out, int32_t *dbg_counts, uint8_t *dbg_gray, uint8_t *dbg_u8,
                          int32_t *status, void *stream_);

extern "C" int ec_event2img(const float *events, const ec_frame *frames, int n_frames, int H, int W, int flags,
                            int out_fmt, int patch, int ldk, void *out, int32_t *dbg_counts, uint8_t *dbg_gray,
                            uint8_t *dbg_u8, int32_t *status, void *stream_)
{
    if (n_frames > 0) {
        EC_REQUIRE(events, "ec_event2img: null pointer");
        EC_REQUIRE(((uintptr_t)events & 15) == 0, "ec_event2img: events must be 16-byte aligned");
    }
    return event2img_impl(events, nullptr, frames, n_frames, H, W, flags, out_fmt, patch, ldk, out, dbg_counts, dbg_gray, dbg_u8,
                          status, stream_);
}

extern "C" int ec_event2img_compact(const uint32_t *events, const ec_frame *frames, int n_frames, int H, int W, int flags,
                                    int out_fmt, int patch, int ldk, void *out, int32_t *dbg_counts, uint8_t *dbg_gray,
                                    uint8_t *dbg_u8, int32_t *status, void *stream_)
{
    if (n_frames > 0) {
        EC_REQUIRE(events, "ec_event2img_compact: null pointer");
        EC_REQUIRE(((uintptr_t)events & 3) == 0, "ec_event2img_compact: events must be 4-byte aligned");
        EC_REQUIRE((long long)H * W < (1ll << 30), "ec_event2img_compact: sensor too large for the 30-bit index");
    }
    return event2img_impl(nullptr, events, frames, n_frames, H, W, flags, out_fmt, patch, ldk, out, dbg_counts, dbg_gray, dbg_u8,
                          status, stream_);
}

extern "C" int ec_pack_events(const float *events, int64_t n_events, int H, int W, uint32_t *out, void *stream)
{
    EC_REQUIRE(n_events >= 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "ec_pack_events: bad arguments");
    if (n_events == 0) return EC_OK;
    EC_REQUIRE(events && out && ((uintptr_t)events & 15) == 0, "ec_pack_events: null or misaligned pointer");
    pack_events_kernel<<<(unsigned)((n_events + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(events), n_events, (long long)H * W, W, out);
    EC_CUDA_CHECK(cudaGetLastError());
    return EC_OK;
}

static int event2img_impl(const float *events, const uint32_t *events_c, const ec_frame *frames, int n_frames, int H, int W, int flags,
                          int out_fmt, int patch, int ldk, void *out, int32_t *dbg_counts, uint8_t *dbg_gray, uint8_t *dbg_u8,
                          int32_t *status, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    EC_REQUIRE(n_frames >= 0, "ec_event2img: n_frames < 0");
    if (n_frames == 0) return EC_OK;
    EC_REQUIRE(frames && out && status, "ec_event2img: null pointer");
    EC_REQUIRE(H > 0 && W > 0 && H >= 8 && W >= 8, "ec_event2img: bad sensor shape %dx%d", H, W);
    EC_REQUIRE(out_fmt >= EC_OUT_F32_NCHW && out_fmt <= EC_OUT_GRAY_F16_PATCH, "ec_event2img: bad out_fmt %d", out_fmt);
    const int f16 = out_fmt == EC_OUT_F16_PATCH || out_fmt == EC_OUT_GRAY_F16_PATCH;
    const int gray = out_fmt == EC_OUT_GRAY_BF16_PATCH || out_fmt == EC_OUT_GRAY_F16_PATCH;
    if (f16 || gray) out_fmt = EC_OUT_BF16_PATCH;      // same row layout; the kernels pick the 16-bit format / the plane count from E2IParams
    int G = 0;
    if (out_fmt == EC_OUT_BF16_PATCH) {
        EC_REQUIRE(patch > 0 && patch % 2 == 0 && OUT % patch == 0, "ec_event2img: patch %d must be even and divide 224", patch);
        EC_REQUIRE(ldk >= (gray ? 1 : 3) * patch * patch && ldk % 2 == 0, "ec_event2img: ldk %d too small / odd", ldk);
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
    p.events_c = events_c;
    p.frames = frames;
    p.n_frames = n_frames;
    p.H = H; p.W = W; p.RB = RB; p.CS = CS;
    p.flags = flags; p.out_fmt = out_fmt; p.patch = patch; p.ldk = ldk; p.G = G;
    p.out = out; p.dbg_counts = dbg_counts; p.dbg_gray = dbg_gray; p.dbg_u8 = dbg_u8; p.status = status;
    p.hx = tb.dev; p.vy = tb.dev + tb.off_vy; p.nlut = reinterpret_cast<const float *>(tb.dev + tb.off_lut);
    p.KH = tb.KH; p.KV = tb.KV;
    p.fragH = tb.dev + tb.off_fragH; p.fragV = tb.dev + tb.off_fragV; p.wsH = tb.dev + tb.off_wsH; p.wsV = tb.dev + tb.off_wsV;
    p.KSH = tb.KSH; p.KSV = tb.KSV;
    p.gray_off = 0;
    for (int c = 0; c < 3; ++c) { p.na[c] = f16 ? tb.na16[c] : tb.na[c]; p.nb[c] = f16 ? tb.nb16[c] : tb.nb[c]; }
    p.affine = f16 ? tb.affine16 : tb.affine;
    p.f16 = f16;
    p.gray = gray;
    p.band_magic = ((1ull << 40) + (unsigned long long)RB * W - 1) / ((unsigned long long)RB * W);
    p.big_off_b = p.big_off_ht = p.big_gp = p.big_hpb = p.big_cap = 0;
    p.big_rb_magic = 0;

    const bool dbg = dbg_counts || dbg_gray || dbg_u8;
    BigPlan bp;
    const bool big = big_plan(H, W, tb.KSH, tb.KSV, CS == 1, bp);
    const bool tc = !big && tc_plan(H, W, CS, tb.KSH, tb.KSV, smem, p.gray_off);
    if (tc && NT < 512) NT = 512;      // 14 warps carry the matrix passes
    typedef void (*kern_t)(const E2IParams);
    kern_t kern;
    if (big) {
        CS = bp.CS; RB = bp.RB; NT = BIG_NT; smem = bp.smem;
        p.CS = CS; p.RB = RB;
        p.big_off_b = bp.off_b; p.big_off_ht = bp.off_ht; p.big_gp = bp.gp; p.big_hpb = bp.hpb; p.big_cap = bp.cap;
        p.big_rb_magic = bp.rb_magic;
        p.band_magic = ((1ull << 36) + (unsigned long long)RB * W - 1) / ((unsigned long long)RB * W);     // 32-bit, __umulhi(l, magic) >> 4
        kern = events_c ? (dbg ? event2img_big_kernel<true, true> : event2img_big_kernel<false, true>)
                        : (dbg ? event2img_big_kernel<true, false> : event2img_big_kernel<false, false>);
    } else if (tc)
        kern = events_c ? (dbg ? event2img_tc_kernel<true, true> : event2img_tc_kernel<false, true>)
                        : (dbg ? event2img_tc_kernel<true, false> : event2img_tc_kernel<false, false>);
    else
        kern = events_c
        ? (dbg ? (tb.KH <= 5 ? event2img_kernel<5, true, true> : (tb.KH <= 11 ? event2img_kernel<11, true, true> : event2img_kernel<0, true, true>))
               : (tb.KH <= 5 ? event2img_kernel<5, false, true> : (tb.KH <= 11 ? event2img_kernel<11, false, true> : event2img_kernel<0, false, true>)))
        : (dbg ? (tb.KH <= 5 ? event2img_kernel<5, true, false> : (tb.KH <= 11 ? event2img_kernel<11, true, false> : event2img_kernel<0, true, false>))
               : (tb.KH <= 5 ? event2img_kernel<5, false, false> : (tb.KH <= 11 ? event2img_kernel<11, false, false> : event2img_kernel<0, false, false>)));
    {
        static std::mutex mu;
        static std::map<std::pair<int, const void *>, size_t> granted;   // the attribute is per device
        std::lock_guard<std::mutex> lk(mu);
        int dev_id = 0;
        EC_CUDA_CHECK(cudaGetDevice(&dev_id));
        size_t &g = granted[std::make_pair(dev_id, (const void *)kern)];
        if (smem > g) {
            EC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            g = smem;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_frames * CS);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // persistent: as many clusters as the device holds at once; each walks over the frames with that stride
    {
        static std::mutex mu2;
        static std::map<std::tuple<int, const void *, int, size_t, int>, int> resident;
        std::lock_guard<std::mutex> lk(mu2);
        int dev_id = 0;
        EC_CUDA_CHECK(cudaGetDevice(&dev_id));
        int &nc = resident[std::make_tuple(dev_id, (const void *)kern, NT, smem, CS)];
        if (nc == 0) {
            cudaLaunchConfig_t q = cfg;
            q.gridDim = dim3((unsigned)(ec::sm_count() * 8));     // any multiple of the cluster size
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) {
                cudaGetLastError();
                n = ec::sm_count() / CS;
            }
            static const int force = getenv("EC_E2I_CLUSTERS") ? atoi(getenv("EC_E2I_CLUSTERS")) : 0;
            nc = force > 0 ? force : n;
        }
        if (n_frames > nc) cfg.gridDim = dim3((unsigned)nc * CS);
    }
    EC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
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
