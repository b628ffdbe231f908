/*
 * oracle/event2img_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's event -> frame -> CLIP-input arithmetic
 * (SURVEY.md section 8 rows A1-A6).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product path (eventclip_b200/) never does.
 *
 * Parity pin: tests/golden/make_golden.py runs the UNMODIFIED reference
 * (/root/reference/datasets/vis.py + PIL/torchvision preprocess) in the
 * authoring container and commits its outputs; tests/test_oracle_golden.py
 * checks this file against those vectors bit for bit.
 *
 * Reference lines restated (relative to /root/reference):
 *   A1 parse_events            datasets/vis.py:44-52
 *   A2 split_event_count       datasets/vis.py:55-72
 *   A3 make_event_histogram    datasets/vis.py:9-14     (counts)
 *   A4 make_event_histogram    datasets/vis.py:16-41    (hot pixels, gray, blend, round)
 *      colour map              datasets/vis.py:95-101
 *   A5 CLIP preprocess [3P]    datasets/event2img.py:119-122 -> torchvision
 *      Resize(224,BICUBIC)/CenterCrop(224)/ToTensor/Normalize, Pillow 8bpc resample
 *   A6 _subsample_imgs         datasets/event2img.py:80-92
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: the only fused multiply-add in the reference's
 * float64 pipeline is the one inside the BLAS dot product of `hist @ cmap`
 * (vis.py:31), restated explicitly with fma() below.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_COORD -1   /* numpy would raise ValueError (bincount / reshape) */
#define ORC_ERR_CAP -2
#define ORC_ERR_ARG -3

/* ---- A2: chunk boundaries by event index (vis.py:55-72) ------------------ */
/* Returns K and fills idx0/idx1 (capacity cap). */
int orc_split_event_count(int64_t E, int64_t N, int64_t *idx0, int64_t *idx1, int cap)
{
    if (E <= 0 || N <= 0) return ORC_ERR_ARG;
    if (E < N) {                      /* vis.py:60-61 */
        if (cap < 1) return ORC_ERR_CAP;
        idx0[0] = 0; idx1[0] = E;
        return 1;
    }
    /* idx = arange(0, E, N): m starts; pairs (idx[k], idx[k+1]) for k < m-1 */
    int64_t m = (E + N - 1) / N;
    int64_t K = m - 1;
    int64_t last = (m - 1) * N;       /* idx[-1] */
    int tail = (double)(E - last) > (double)N * 0.5;  /* vis.py:67 */
    if (K + tail > cap) return ORC_ERR_CAP;
    for (int64_t k = 0; k < K; ++k) { idx0[k] = k * N; idx1[k] = (k + 1) * N; }
    if (tail) { idx0[K] = E - N; idx1[K] = E; ++K; }
    return (int)K;
}

/* ---- A1 + A3: per-pixel polarity counts (vis.py:44-52, 9-14) ------------- */
/* events: float32 [n,4] rows (x,y,t,p).  counts: int64 [H,W,2] (pos,neg), zeroed here. */
int orc_histogram(const float *ev, int64_t n, int H, int W, int64_t *counts)
{
    int64_t bins = (int64_t)H * W;
    memset(counts, 0, sizeof(int64_t) * bins * 2);
    for (int64_t i = 0; i < n; ++i) {
        int32_t x = (int32_t)ev[4 * i + 0];   /* astype(int32): truncation */
        int32_t y = (int32_t)ev[4 * i + 1];
        int32_t p = (int32_t)ev[4 * i + 3];
        if (p == 0) continue;                 /* neither p>0 nor p<0 */
        int64_t idx = (int64_t)x + (int64_t)y * W;   /* flat index, as bincount sees it */
        if (idx < 0 || idx >= bins) return ORC_ERR_COORD;
        counts[2 * idx + (p > 0 ? 0 : 1)] += 1;
    }
    return ORC_OK;
}

/* ---- A4: hot-pixel zeroing + normalise + colour + blend + round ---------- */
/* The comparison `c > thresh*std + mean` (vis.py:17-23) is evaluated in exact
 * integer arithmetic: with n bins, S1 = sum c, S2 = sum c^2,
 *   c > mean + t*std  <=>  c*n > S1  and  (c*n - S1)^2 > t^2 * (n*S2 - S1^2).
 * numpy evaluates the same inequality in float64; the two agree except when
 * numpy's own rounding error decides an exact tie. */
static int hot(int64_t c, int64_t n, int64_t S1, int64_t S2, int64_t t)
{
    __int128 d = (__int128)c * n - S1;
    if (d <= 0) return 0;
    __int128 var = (__int128)n * S2 - (__int128)S1 * S1;   /* n^2 * variance >= 0 */
    return d * d > (__int128)t * t * var;
}

/* gray value of one pixel; float64 sequence of vis.py:27-39 under numpy>=2
 * (float32 array / int64 scalar -> float64), BLAS k-loop = fma. */
static uint8_t gray_px(int64_t pos, int64_t neg, int64_t mx, int bgmask)
{
    if (mx == 0) return 0;   /* 0/0 = NaN -> uint8 cast yields 0 on x86 (undefined in the reference) */
    double gp = (double)(float)pos / (double)mx;
    double gn = (double)(float)neg / (double)mx;
    double img = fma(gn, 127.0, gp * 127.0);          /* hist @ cmap, both cmap rows = 127 */
    if (bgmask) {
        double w = gp + gn;                           /* hist.sum(-1) */
        if (w < 0.0) w = 0.0;
        if (w > 1.0) w = 1.0;
        double a = img * w;
        double b = 255.0 * (1.0 - w);
        img = a + b;
    }
    return (uint8_t)nearbyint(img);                   /* np.round = half-to-even */
}

/* counts int64 [H,W,2] -> gray uint8 [H,W] (the three channels are identical);
 * zeroed (nullable) uint8 [H,W,2] = 1 where the bin was removed.
 * stats_out (nullable) int64[5] = {n, S1, S2, max_after, n_zeroed}. */
int orc_frame_from_counts(const int64_t *counts, int H, int W, int count_non_zero,
                          int bgmask, int thresh, uint8_t *gray, uint8_t *zeroed,
                          int64_t *stats_out)
{
    int64_t bins2 = (int64_t)H * W * 2;
    int64_t n = 0, S1 = 0, S2 = 0;
    for (int64_t i = 0; i < bins2; ++i) {
        int64_t c = counts[i];
        if (count_non_zero && c == 0) continue;
        n += 1; S1 += c; S2 += c * c;
    }
    int64_t mx = 0, nz = 0;
    for (int64_t i = 0; i < bins2; ++i) {
        int64_t c = counts[i];
        int z = thresh > 0 && n > 0 && hot(c, n, S1, S2, thresh);
        if (zeroed) zeroed[i] = (uint8_t)z;
        if (z) { nz++; continue; }
        if (c > mx) mx = c;
    }
    for (int64_t px = 0; px < (int64_t)H * W; ++px) {
        int64_t pos = counts[2 * px], neg = counts[2 * px + 1];
        if (thresh > 0 && n > 0) {
            if (hot(pos, n, S1, S2, thresh)) pos = 0;
            if (hot(neg, n, S1, S2, thresh)) neg = 0;
        }
        gray[px] = gray_px(pos, neg, mx, bgmask);
    }
    if (stats_out) { stats_out[0] = n; stats_out[1] = S1; stats_out[2] = S2; stats_out[3] = mx; stats_out[4] = nz; }
    return ORC_OK;
}

/* ---- A5: Pillow 8-bit bicubic resample (ImagingResample, 8bpc) ----------- */
#define PRECISION_BITS (32 - 8 - 2)

static double bicubic_filter(double x)
{
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

/* Fills bounds[2*out] = (lo, count) and kk[out*ksize] fixed-point weights. Returns ksize. */
static int precompute_coeffs(int in, int out, int **bounds_p, int32_t **kk_p)
{
    double scale = (double)in / (double)out;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    double support = 2.0 * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    int *bounds = (int *)malloc(sizeof(int) * 2 * out);
    int32_t *kk = (int32_t *)calloc((size_t)out * ksize, sizeof(int32_t));
    double *k = (double *)malloc(sizeof(double) * ksize);
    for (int xx = 0; xx < out; ++xx) {
        double center = (xx + 0.5) * scale;
        double ww = 0.0, ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in) xmax = in;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w; ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            double v = k[x] * (double)(1 << PRECISION_BITS);
            kk[(size_t)xx * ksize + x] = v < 0 ? (int32_t)(-0.5 + v) : (int32_t)(0.5 + v);
        }
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
    }
    free(k);
    *bounds_p = bounds; *kk_p = kk;
    return ksize;
}

static uint8_t clip8(int32_t v)
{
    v >>= PRECISION_BITS;            /* arithmetic shift, as Pillow's lookup index */
    return v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v);
}

/* torchvision Resize(224): shorter side -> 224, longer -> int(224*long/short). */
void orc_resized_shape(int H, int W, int *Ho, int *Wo)
{
    if (W <= H) { *Wo = 224; *Ho = (int)(224.0 * H / W); }
    else        { *Ho = 224; *Wo = (int)(224.0 * W / H); }
}

/* gray u8 [H,W] -> u8 [224,224]: Pillow resize (horizontal pass, u8, vertical pass, u8)
 * then CenterCrop(224).  hpass (nullable): u8 [H, Wo] intermediate for stage checks;
 * full (nullable): u8 [Ho, Wo] resized image before the crop. */
int orc_resize_crop_224(const uint8_t *gray, int H, int W, uint8_t *out, uint8_t *hpass, uint8_t *full)
{
    int Ho, Wo;
    orc_resized_shape(H, W, &Ho, &Wo);
    if (Ho < 224 || Wo < 224) return ORC_ERR_ARG;
    int *bh, *bv; int32_t *kh, *kv;
    uint8_t *tmp = (uint8_t *)malloc((size_t)H * Wo);
    const uint8_t *src = gray;
    int srcW = W;
    if (Wo != W) {   /* Pillow skips a pass whose size does not change */
        int ks = precompute_coeffs(W, Wo, &bh, &kh);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < Wo; ++x) {
                int32_t ss = 1 << (PRECISION_BITS - 1);
                int lo = bh[2 * x], cnt = bh[2 * x + 1];
                for (int j = 0; j < cnt; ++j) ss += (int32_t)gray[(size_t)y * W + lo + j] * kh[(size_t)x * ks + j];
                tmp[(size_t)y * Wo + x] = clip8(ss);
            }
        free(bh); free(kh);
        src = tmp; srcW = Wo;
    }
    if (hpass) memcpy(hpass, src == gray ? gray : tmp, (size_t)H * Wo);
    uint8_t *res = (uint8_t *)malloc((size_t)Ho * Wo);
    if (Ho != H) {
        int ks = precompute_coeffs(H, Ho, &bv, &kv);
        for (int y = 0; y < Ho; ++y) {
            int lo = bv[2 * y], cnt = bv[2 * y + 1];
            for (int x = 0; x < Wo; ++x) {
                int32_t ss = 1 << (PRECISION_BITS - 1);
                for (int j = 0; j < cnt; ++j) ss += (int32_t)src[(size_t)(lo + j) * srcW + x] * kv[(size_t)y * ks + j];
                res[(size_t)y * Wo + x] = clip8(ss);
            }
        }
        free(bv); free(kv);
    } else {
        memcpy(res, src, (size_t)Ho * Wo);
    }
    if (full) memcpy(full, res, (size_t)Ho * Wo);
    /* CenterCrop(224): top = int(round((Ho-224)/2.0)), left likewise (python round = half-even) */
    int top = (int)nearbyint((Ho - 224) / 2.0), left = (int)nearbyint((Wo - 224) / 2.0);
    for (int y = 0; y < 224; ++y) memcpy(out + (size_t)y * 224, res + (size_t)(y + top) * Wo + left, 224);
    free(res); free(tmp);
    return ORC_OK;
}

/* ToTensor + Normalize: float32 ((u/255) - mean_c) / std_c, CLIP constants (method.py:17-18). */
static const float CLIP_MEAN[3] = {0.48145466f, 0.4578275f, 0.40821073f};
static const float CLIP_STD[3] = {0.26862954f, 0.26130258f, 0.27577711f};

void orc_normalize(const uint8_t *u8, float *out /* [3,224,224] */)
{
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 224 * 224; ++i) {
            volatile float v = (float)u8[i] / 255.0f;   /* ToTensor: u8 -> f32 .div(255) */
            volatile float d = v - CLIP_MEAN[c];
            out[(size_t)c * 224 * 224 + i] = d / CLIP_STD[c];
        }
}

/* ---- events2frames (vis.py:75-117): all K chunks -> uint8 [K,H,W] gray ---- */
int orc_events2frames(const float *ev, int64_t E, int H, int W, int64_t N, int count_non_zero,
                      int bgmask, uint8_t *frames, int capK)
{
    int64_t *i0 = (int64_t *)malloc(sizeof(int64_t) * (capK + 1));
    int64_t *i1 = (int64_t *)malloc(sizeof(int64_t) * (capK + 1));
    int K = orc_split_event_count(E, N, i0, i1, capK);
    if (K < 0) { free(i0); free(i1); return K; }
    int64_t *counts = (int64_t *)malloc(sizeof(int64_t) * H * W * 2);
    int rc = ORC_OK;
    for (int k = 0; k < K && rc == ORC_OK; ++k) {
        rc = orc_histogram(ev + 4 * i0[k], i1[k] - i0[k], H, W, counts);
        if (rc == ORC_OK)
            rc = orc_frame_from_counts(counts, H, W, count_non_zero, bgmask, 10,
                                       frames + (size_t)k * H * W, NULL, NULL);
    }
    free(counts); free(i0); free(i1);
    return rc == ORC_OK ? K : rc;
}

/* ---- one sample, events -> img f32 [T,3,224,224] + valid u8 [T] ----------
 * sel (nullable): chunk ids to use when K > T (the reference draws
 * torch.randperm(K)[:T], event2img.py:85 -- the caller supplies that draw).
 * only_selected != 0 skips chunks that are not selected (what the CUDA path does);
 * 0 converts every chunk first, like the reference (event2img.py:118-122). */
int orc_event2img_sample(const float *ev, int64_t E, int H, int W, int64_t N, int T,
                         int count_non_zero, int bgmask, const int64_t *sel,
                         int only_selected, float *img, uint8_t *valid)
{
    int capK = (int)(E / (N > 0 ? N : 1)) + 2;
    int64_t *i0 = (int64_t *)malloc(sizeof(int64_t) * capK);
    int64_t *i1 = (int64_t *)malloc(sizeof(int64_t) * capK);
    int K = orc_split_event_count(E, N, i0, i1, capK);
    if (K < 0) { free(i0); free(i1); return K; }
    if (K > T && !sel) { free(i0); free(i1); return ORC_ERR_ARG; }
    int64_t *counts = (int64_t *)malloc(sizeof(int64_t) * H * W * 2);
    uint8_t *gray = (uint8_t *)malloc((size_t)H * W);
    uint8_t *u8 = (uint8_t *)malloc(224 * 224);
    float *all = NULL;
    int rc = ORC_OK;
    memset(img, 0, sizeof(float) * (size_t)T * 3 * 224 * 224);
    memset(valid, 0, (size_t)T);
    if (!only_selected) {
        all = (float *)malloc(sizeof(float) * (size_t)K * 3 * 224 * 224);
        for (int k = 0; k < K && rc == ORC_OK; ++k) {
            rc = orc_histogram(ev + 4 * i0[k], i1[k] - i0[k], H, W, counts);
            if (rc) break;
            orc_frame_from_counts(counts, H, W, count_non_zero, bgmask, 10, gray, NULL, NULL);
            rc = orc_resize_crop_224(gray, H, W, u8, NULL, NULL);
            orc_normalize(u8, all + (size_t)k * 3 * 224 * 224);
        }
    }
    int nv = K > T ? T : K;
    for (int t = 0; t < nv && rc == ORC_OK; ++t) {
        int64_t k = K > T ? sel[t] : t;
        if (k < 0 || k >= K) { rc = ORC_ERR_ARG; break; }
        float *dst = img + (size_t)t * 3 * 224 * 224;
        if (all) {
            memcpy(dst, all + (size_t)k * 3 * 224 * 224, sizeof(float) * 3 * 224 * 224);
        } else {
            rc = orc_histogram(ev + 4 * i0[k], i1[k] - i0[k], H, W, counts);
            if (rc) break;
            orc_frame_from_counts(counts, H, W, count_non_zero, bgmask, 10, gray, NULL, NULL);
            rc = orc_resize_crop_224(gray, H, W, u8, NULL, NULL);
            orc_normalize(u8, dst);
        }
        valid[t] = 1;
    }
    free(all); free(u8); free(gray); free(counts); free(i0); free(i1);
    return rc == ORC_OK ? K : rc;
}
