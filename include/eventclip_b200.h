/*
 * eventclip_b200.h -- C ABI of the B200-native EventCLIP inference hot path.
 *
 * Every entry point is plain `extern "C"`: raw device/host pointers, sizes and a
 * cudaStream_t passed as void*.  No torch types, no C++ types, no exceptions.
 * Functions return 0 on success or a negative EC_ERR_* code; ec_last_error()
 * returns a thread-local message for the last failure.  All device work is
 * enqueued on the given stream and is asynchronous; the library never
 * synchronises.  The caller (PyTorch on the reference side) owns every buffer.
 *
 * Each function cites the reference interface it replaces (paths relative to
 * the reference repository root).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 */
#ifndef EVENTCLIP_B200_H
#define EVENTCLIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EC_API __attribute__((visibility("default")))
#else
#define EC_API
#endif

#define EC_OK 0
#define EC_ERR_ARG -1         /* bad shape / null pointer / unsupported option   */
#define EC_ERR_CUDA -2        /* CUDA runtime error (see ec_last_error)          */
#define EC_ERR_UNSUPPORTED -3 /* device is not sm_100 or shape exceeds limits    */
#define EC_ERR_CAPACITY -4    /* caller-provided table / workspace too small     */

/* bits of the device status word written by ec_event2img (0 = clean) */
#define EC_STATUS_BAD_COORD 1      /* x + y*W outside [0, H*W): numpy would raise ValueError (vis.py:12-14) */
#define EC_STATUS_COUNT_OVERFLOW 2 /* a per-pixel polarity count exceeded 65535 in the packed histogram     */

/* flags of ec_event2img */
#define EC_FLAG_COUNT_NON_ZERO 1  /* quantize_args['count_non_zero'] (vis.py:18-20)   */
#define EC_FLAG_BACKGROUND_MASK 2 /* quantize_args['background_mask'] (vis.py:34-37)  */

/* output formats of ec_event2img */
#define EC_OUT_F32_NCHW 0  /* float32 [slot,3,224,224] -- the tensor the reference DataLoader yields */
#define EC_OUT_BF16_NCHW 1 /* same layout, bf16 (round-to-nearest-even of the float32 value)      */
#define EC_OUT_BF16_PATCH 2 /* bf16 [slot, G*G, ldk] im2col rows (c,dy,dx) feeding the patch GEMM  */
#define EC_OUT_F16_PATCH 3  /* the same rows in fp16 (round-to-nearest-even of the float32 value): the fp16-operand inference forward */
#define EC_OUT_GRAY_BF16_PATCH 4 /* ONE plane: bf16 [slot, G*G, ldk >= P*P] rows (dy,dx) holding the resampled gray byte / 128 exactly.  The three
                                    normalised channels of a frame are affine in that byte (grayscale frames, datasets/vis.py:94-104 +
                                    CLIP's Normalize), so the caller folds conv1 to K = P*P and a per-output bias (SURVEY 8(d) note) */
#define EC_OUT_GRAY_F16_PATCH 5  /* the same in fp16 */

EC_API const char *ec_last_error(void);
EC_API int ec_version(void);
/* 0 when the current CUDA device is sm_100 (B200); EC_ERR_UNSUPPORTED otherwise. No CPU fallback exists. */
EC_API int ec_device_check(void);

/* One frame of work for ec_event2img: events [ev_start, ev_start+ev_count) of the packed
 * stream are histogrammed and written to output image `out_slot`.  ev_count == 0 writes an
 * all-zero image (the reference's zero padding of missing views, datasets/event2img.py:89-91). */
typedef struct ec_frame {
    int64_t ev_start;
    int32_t ev_count;
    int32_t out_slot;
} ec_frame;

/* HOST function.  Replaces split_event_count (datasets/vis.py:55-72), the view-count rule and
 * _subsample_imgs (datasets/event2img.py:70-72, 80-92) for a batch of B packed samples.
 *   offsets   host int64 [B+1]  event offsets of each sample in the packed stream
 *   N         events per frame;  T = number of view slots per sample (max_imgs)
 *   sel       host int32 [B,T] or NULL: chunk ids to use for samples with K > T (the caller's
 *             torch.randperm(K)[:T] draw, event2img.py:85); NULL selects chunks 0..T-1
 *   compact   0: out_slot = b*T + t and padded slots get an ev_count==0 frame (reference layout
 *             [B,T,...]);  1: only valid views, out_slot = running index (the imgs[valid_mask]
 *             gather of models/clip_cls.py:139)
 *   frames    host ec_frame [cap] out;  valid host uint8 [B,T] out;  chunks host int32 [B] out
 *             (K per sample, nullable);  n_frames / n_valid out. */
EC_API int ec_plan_frames(const int64_t *offsets, int B, int64_t N, int T, const int32_t *sel, int compact,
                   ec_frame *frames, int cap, uint8_t *valid, int32_t *chunks, int *n_frames, int *n_valid);

/* DEVICE.  Fused event stream -> CLIP input.  Replaces, in one launch and with no intermediate
 * tensor in HBM: parse_events + make_event_histogram (datasets/vis.py:44-52, 6-41, colour map
 * 95-101) and the CLIP preprocess applied at datasets/event2img.py:119-122
 * (Resize(224,bicubic) / CenterCrop(224) / ToTensor / Normalize).
 *   events    device float32 [*,4] packed (x,y,t,p) rows, 16-byte aligned
 *   frames    device ec_frame [n_frames]
 *   H, W      sensor shape;  flags EC_FLAG_*;  out_fmt EC_OUT_*
 *   patch, ldk  only for the *_PATCH formats: patch size P and row stride (elements, >= 3*P*P; >= P*P for the gray formats)
 *   out       device output images (format above)
 *   dbg_counts device int32 [n_frames,H,W,2] or NULL   (parity taps: raw counts,
 *   dbg_gray   device uint8 [n_frames,H,W]   or NULL    uint8 frame = any channel of vis.py's output,
 *   dbg_u8     device uint8 [n_frames,224,224] or NULL   resized+cropped uint8) indexed by frame, not slot
 *   status    device int32 [1], OR-ed with EC_STATUS_* bits (caller zeroes it) */
EC_API int ec_event2img(const float *events, const ec_frame *frames, int n_frames, int H, int W, int flags,
                 int out_fmt, int patch, int ldk, void *out, int32_t *dbg_counts, uint8_t *dbg_gray,
                 uint8_t *dbg_u8, int32_t *status, void *stream);

/* Row F2 -- compact event wire format.  One 32-bit word per event: bits [0,30) = the flat pixel index x + y*W exactly as
 * np.bincount receives it at datasets/vis.py:9-14 (coordinates truncated like .astype(int); x >= W aliasing preserved),
 * bits [30,32) = polarity code (0: p == 0, ignored by the histogram; 1: p > 0; 2: p < 0; 3: index outside [0, H*W) ->
 * EC_STATUS_BAD_COORD when consumed).  t is dropped: nothing reads it after the chunking (vis.py:44-52).  4 bytes per
 * event instead of 16 on the wire and in HBM.  ec_pack_events converts float32 [n,4] rows on the device (after
 * ec_center_events / ec_flip_events where those apply); ec_event2img_compact is ec_event2img on packed words, with
 * ec_frame.ev_start / ev_count counting events as before.  Frames are bit-identical to the float path. */
EC_API int ec_pack_events(const float *events, int64_t n_events, int H, int W, uint32_t *out, void *stream);
EC_API int ec_event2img_compact(const uint32_t *events, const ec_frame *frames, int n_frames, int H, int W, int flags,
                                int out_fmt, int patch, int ldk, void *out, int32_t *dbg_counts, uint8_t *dbg_gray,
                                uint8_t *dbg_u8, int32_t *status, void *stream);

/* DEVICE, in place.  center_events (datasets/utils.py:38-57), applied to every sample before the frames are built
 * (datasets/caltech.py:176): per sample t -= min t, x -= ((x_max + x_min + 1) - W) // 2, y likewise (float32 math).
 *   events device float32 [*,4];  offsets device int64 [B+1] */
EC_API int ec_center_events(float *events, const int64_t *offsets, int B, int H, int W, void *stream);

/* DEVICE, out of place.  The deterministic flips of the reference's test-time augmentation
 * (datasets/utils.py:18-35 with p = 1, used at datasets/event2img.py:100-103): hflip x -> W-1-x;
 * tflip reverses the event order of each sample, t -> t_last - t, p -> -p.  src/dst device float32 [*,4]. */
EC_API int ec_flip_events(const float *src, float *dst, const int64_t *offsets, int B, int W, int hflip, int tflip,
                          void *stream);

/* Launch geometry ec_event2img will use for a sensor (for bench/roofline reporting). */
EC_API int ec_event2img_geometry(int H, int W, int *cluster_size, int *threads, int *smem_bytes);

/* ---- CLIP ViT image encoder (openai/CLIP VisionTransformer; call sites models/clip_cls.py:101,
 *      models/clip_cls_ft.py:180) ------------------------------------------------------------ */

/* C[M,N] = epilogue(A[M,K] . W[N,K]^T): bf16 operands, fp32 accumulation in TMEM (tcgen05.mma), TMA loads.
 * The nn.Linear / conv1 / proj contractions of the encoder all go through this entry.
 *   A   device bf16 [M,lda]   W device bf16 [N,ldw]   (lda, ldw in elements, multiples of 8, >= K)
 *   bias device fp32 [N] or NULL
 *   epi  EC_EPI_*;  out: bf16 or fp32 [M,ldo] depending on epi;  res: fp32 [M,ldo] residual or NULL
 *   row_map: for EC_EPI_PATCH, tokens per image G*G (output row = (m/G2)*(G2+1) + 1 + m%G2 and
 *            pos = fp32 [G2+1, N] positional embedding passed through `res`) */
#define EC_EPI_BF16 0        /* out bf16 = acc + bias                                       */
#define EC_EPI_BF16_QGELU 1  /* out bf16 = QuickGELU(acc + bias)  (x * sigmoid(1.702 x))     */
#define EC_EPI_F32_RESADD 2  /* out fp32 = res + acc + bias       (residual stream update)   */
#define EC_EPI_F32 3         /* out fp32 = acc + bias                                        */
#define EC_EPI_PATCH 4       /* out fp32 token rows = acc + pos[1 + m%G2]  (patch embedding) */
#define EC_EPI_F16_RESADD 5  /* out fp16 = res(fp16) + acc + bias: residual stream in the reference's CUDA precision;
                                `out` and `res` point to fp16 [M,ldo] (res is passed through the float* parameter) */
#define EC_EPI_F16X2_RESADD 6 /* ec_gemm_bf16_stats2 only: the fp16 residual stream as a (hi, lo) pair of fp16 planes, see there */
#define EC_EPI_F16_OPERANDS 0x100 /* OR-ed into epi: A and W hold fp16 instead of bf16 (tcgen05.mma kind::f16 takes either, at the same
                                rate, but not a mixed pair) and the 16-bit outputs (EC_EPI_BF16, EC_EPI_BF16_QGELU) are written as
                                fp16 -- the reference's own CUDA inference dtype (clip.load keeps fp16 weights, test.py:26-29),
                                3 more mantissa bits than bf16 at every rounding point of the encoder */
EC_API int ec_gemm_bf16(const void *A, int lda, const void *W, int ldw, const float *bias, int M, int N, int K,
                 int epi, void *out, int ldo, const float *res, int row_map, void *stream);

/* Measurement hook (bench.py's roofline): after ec_gemm_timing(buf, capacity) the i-th following GEMM launch (i < capacity)
 * writes its device-side start / end time (%globaltimer, ns) to buf[2i], buf[2i+1] -- also as a node of a replayed CUDA
 * graph, where events cannot be recorded.  buf: device uint64 [2*capacity], zeroed by the caller; NULL disables. */
EC_API int ec_gemm_timing(uint64_t *buf, int capacity);

/* Split-K variant for weight-gradient shapes (small M x N, long K = tokens): out fp32 [M,ldo] = A[M,K] . W[N,K]^T with the
 * K range cut into `splits` parts that run as separate tiles and are summed in a fixed order (deterministic).
 * workspace: device fp32 [splits * M * N] (unused when splits == 1).  ec_gemm_splitk_choose returns the split count that
 * fills the SMs once for a shape (>= 512 K elements per split). */
EC_API int ec_gemm_splitk_choose(int M, int N, int K);
EC_API int ec_gemm_bf16_splitk(const void *A, int lda, const void *W, int ldw, int M, int N, int K, int splits,
                               float *workspace, float *out, int ldo, void *stream);

/* Weight gradient without transposing the activations: out fp32 [M,ldo] = A^T . B with A bf16 [K, lda >= M] and B bf16
 * [K, ldb >= N] both TOKEN-major (K = tokens is the row index) -- e.g. dW[out,in] = dY^T . X.  The tiles are fetched as
 * 64 x 64 TMA boxes and fed to tcgen05.mma as MN-major operands; K is split as in ec_gemm_bf16_splitk (rows >= K read as
 * zeros, so K needs no padding).  M, N, lda, ldb multiples of 8. */
EC_API int ec_gemm_bf16_tn_splitk(const void *A, int lda, const void *B, int ldb, int M, int N, int K, int splits,
                                  float *workspace, float *out, int ldo, void *stream);

/* LayerNorm over the last dim (eps 1e-5, fp32 statistics): x fp32 [M, ldx] rows -> bf16 [M,d] (out_bf16)
 * and/or fp32 [M,d] (out_f32); either output may be NULL.  row_stride_in lets ln_post read only the
 * class-token rows (stride L*d). */
EC_API int ec_layernorm(const float *x, int64_t row_stride_in, const float *gamma, const float *beta, int M, int d,
                 void *out_bf16, float *out_f32, void *stream);
/* Same LayerNorm for the fp16 residual stream of the inference forward (the reference's CUDA precision, test.py:26-29 loads
 * CLIP in fp16): x is fp16 when x_is_f16 != 0; any of the three outputs may be NULL.  out_f16 is what ln_pre writes to
 * start the fp16 stream. */
EC_API int ec_layernorm_ex(const void *x, int x_is_f16, int64_t row_stride_in, const float *gamma, const float *beta, int M,
                           int d, void *out_bf16, float *out_f32, void *out_f16, void *stream);

/* ---- LayerNorm folded into the GEMMs of the fp16 residual stream (ln_1 -> in_proj, ln_2 -> c_fc of openai/CLIP's
 *      ResidualAttentionBlock [3P], called through models/clip_cls.py:101).  With Wg = fp16(gamma * W) (column k scaled by
 *      gamma_k; fp16 because tcgen05.mma kind::f16 rejects an fp16 x bf16 operand pair), s_j = sum_k Wg[j,k] and
 *      c_j = beta . W[j] + b_j:   LN(x) W^T + b  =  rstd (x Wg^T - mean s) + c.   When every row of Wg is centred
 *      (Wg[j,:] -= mean_k Wg[j,k]; allowed because sum_k (x_k - mean) = 0) s vanishes: pass colsum = NULL and the epilogue is
 *      one fma per element.  The GEMM reads the fp16 residual rows themselves as its A operand; the row statistics arrive as
 *      partial (sum, sum of squares) pairs written by the epilogue that produced the rows. */
/* The same update on a residual stream kept as TWO fp16 planes, x = hi + lo (hi = fp16(x), lo = fp16(x - hi)): the value keeps
 * ~22 mantissa bits across the 24 residual updates of a 12-block tower (an fp16 stream rounds it to 11 bits after every update),
 * while hi stays directly usable as the fp16 A operand of the next LayerNorm-folded GEMM (ec_gemm_ln).  Both planes [M, ldo] are
 * read and written in place; stats_out as in ec_gemm_bf16_stats (statistics of hi + lo before rounding).
 * Replaces the same reference lines as ec_gemm_bf16 with EC_EPI_F32_RESADD (x += out_proj(...) / c_proj(...) of openai-CLIP's
 * ResidualAttentionBlock, called through models/clip_cls.py:101). */
EC_API int ec_gemm_bf16_stats2(const void *A, int lda, const void *W, int ldw, const float *bias, int M, int N, int K, void *x_hi,
                               void *x_lo, int ldo, float *stats_out, int f16_operands, void *stream);
/* LayerNorm of fp32 rows written as that (hi, lo) pair: out_hi = fp16(y), out_lo = fp16(y - out_hi)  (ln_pre of the tower) */
EC_API int ec_layernorm_f16x2(const float *x, int64_t row_stride, const float *gamma, const float *beta, int M, int d, void *out_hi,
                              void *out_lo, void *stream);
/* number of float2 statistics slots per row that ec_gemm_bf16_stats writes for an N-column output */
EC_API int ec_gemm_stats_parts(int N);
/* out fp16 = res(fp16) + A W^T + bias (EC_EPI_F16_RESADD) and stats_out float [M, parts, 2] = per-row partial sums of the values
 * written (parts = ec_gemm_stats_parts(N)) */
EC_API int ec_gemm_bf16_stats(const void *A, int lda, const void *W, int ldw, const float *bias, int M, int N, int K, void *out,
                              int ldo, const void *res, float *stats_out, int f16_operands, void *stream);
/* out bf16 = epi(LayerNorm(X) W^T + b) from X fp16 [M,K] (ldx), Wg fp16 [N,K], colsum s [N] or NULL, cbias c [N], stats float
 * [M, n_parts, 2];  epi = EC_EPI_BF16 or EC_EPI_BF16_QGELU, optionally | EC_EPI_F16_OPERANDS for an fp16 output */
EC_API int ec_gemm_ln(const void *X, int ldx, const void *Wg, int ldw, const float *colsum, const float *cbias,
                      const float *stats, int n_parts, int M, int N, int K, int epi, void *out, int ldo, void *stream);
/* stats[row, 0] = (sum, sum of squares) of the fp16 row, other parts zero: statistics for rows no GEMM epilogue produced */
EC_API int ec_row_stats_f16(const void *x, int64_t row_stride, int M, int d, float *stats, int n_parts, void *stream);

/* Multi-head self-attention core on the packed QKV activations of nn.MultiheadAttention:
 * qkv bf16 [n_img*L, 3*d] (q | k | v, head h at columns h*64), out bf16 [n_img*L, d].
 * softmax(q k^T / sqrt(64)) v per (image, head); head_dim is 64 for every CLIP ViT. */
EC_API int ec_attention(const void *qkv, void *out, int n_img, int L, int heads, void *stream);

/* Same with an optional causal mask (key j visible to query i iff j <= i): the attention of CLIP's text tower
 * (openai-CLIP encode_text [3P], called at models/clip_cls.py:84, models/clip_cls_ft.py:152). */
#define EC_ATTN_CAUSAL 1
#define EC_ATTN_F16 2 /* qkv, the probabilities and out are fp16 instead of bf16 (tensor-memory kernels, L <= 384) */
EC_API int ec_attention_ex(const void *qkv, void *out, int n_seq, int L, int heads, int causal /* EC_ATTN_* bits */, void *stream);

/* Text-tower input: out[n*L + l, :] = token_embedding[tokens[n,l], :] + positional_embedding[l, :]  (fp32).
 * tokens device int32 [n_seq, L]. */
EC_API int ec_embed_tokens(const float *table, const int32_t *tokens, const float *pos, float *out, int n_seq, int L, int d,
                           int vocab, void *stream);

/* Writes the class-token rows of the pre-LN token matrix: x[n*L + 0, :] = class_embedding + pos[0]. */
EC_API int ec_cls_rows(float *x, const float *class_embedding, const float *pos, int n_img, int L, int d, void *stream);

/* float32 -> bf16 conversion (weights, activations). n elements. */
EC_API int ec_f32_to_bf16(const float *src, void *dst, int64_t n, void *stream);

/* NCHW image (fp32 or bf16) -> bf16 im2col rows [n_img*G*G, ldk] for images that did not come from
 * ec_event2img (the reference's data_dict['img'] input, models/clip_cls.py:133). in_is_bf16: 0 fp32, 1 bf16. */
EC_API int ec_im2col(const void *img, int in_is_bf16 /* bit 0: input is bf16; bit 1: write fp16 rows */, int n_img, int patch, int ldk,
                     void *out, void *stream);

/* LoRA merge (models/lora.py:138-149 q/k/v, 49-52 out_proj): Wm[rows,d] bf16 = W[rows,d] + up[rows,r] . down[r,d],
 * fp32 math, one rounding to bf16.  up/down NULL copies W. */
EC_API int ec_lora_merge(const float *W, const float *up, const float *down, int rows, int d, int r, void *Wm_bf16,
                  void *stream);

/* ---- classifier heads ---------------------------------------------------------------------- */
#define EC_AGG_SUM 0
#define EC_AGG_MEAN 1
#define EC_AGG_MAX 2
/* Logit head of ZS/FS/FT classifiers (models/clip_cls.py:144-154, 326-342; models/clip_cls_ft.py:232-248).
 *   feats  fp32 [B*T, C] view features in slot order (zeros for invalid slots are not required)
 *   valid  uint8 [B*T];  text fp32 [n_cls, C] (already L2-normalised)
 *   normalize  0: zero-shot (image features used as is, clip_cls.py:148)  1: L2-normalise each view first
 *   out_full fp32 [B,T,n_cls] (invalid rows = 0), out_logits / out_probs fp32 [B,n_cls],
 *   out_top int32 [B,2,5]: top-5 class ids of logits and of probs (nullable). */
EC_API int ec_head(const float *feats, const uint8_t *valid, const float *text, int B, int T, int C, int n_cls,
            float scale, int normalize, int agg, float *out_full, float *out_logits, float *out_probs,
            int32_t *out_top, void *stream);

/* fp32 SIMT GEMM for the few-shot adapter (models/adapter.py:82-105 runs in fp32):
 * out[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ res);  act: 0 none, 1 ReLU. */
EC_API int ec_gemm_f32(const float *A, const float *W, const float *bias, const float *res, int M, int N, int K,
                int act, float *out, void *stream);

/* Adapter self-attention over the <=16 views of each sample with key-padding mask
 * (nn.TransformerEncoderLayer, src_key_padding_mask=~valid; models/adapter.py:96-97).
 * qkv fp32 [B*T, 3*D], out fp32 [B*T, D]. */
EC_API int ec_adapter_attention(const float *qkv, const uint8_t *valid, int B, int T, int D, int heads, float *out,
                         void *stream);

/* Backward of ec_adapter_attention (training of the few-shot feature adapter, models/adapter.py:82-105 under autograd in the
 * reference): d_qkv fp32 [B*T, 3*D] from qkv, the mask and d_out fp32 [B*T, D].  heads <= 16, T <= 16. */
EC_API int ec_adapter_attention_bwd(const float *qkv, const uint8_t *valid, const float *d_out, int B, int T, int D, int heads,
                                    float *d_qkv, void *stream);
/* dx = dy where y > 0, else 0: backward of the ReLU that ec_gemm_f32 (act = 1) fuses; y is the activation's output. */
EC_API int ec_relu_bwd(const float *y, const float *dy, float *dx, int64_t n, void *stream);

/* out = r*a + (1-r)*b  (Adapter.residual_add, models/adapter.py:22-25); n elements. */
EC_API int ec_blend(const float *a, const float *b, double r, float *out, int64_t n, void *stream);

/* dst[s,:] = idx[s] >= 0 ? src[idx[s],:] : 0 -- places the features of the valid views into the zero-initialised
 * [B*T, C] slot matrix (models/clip_cls.py:320-321).  idx device int32 [n_rows]. */
EC_API int ec_gather_rows(const float *src, const int32_t *idx, float *dst, int n_rows, int C, void *stream);

/* F.normalize(x, p=2, dim=-1) on fp32 rows [M,C] (text features, models/clip_cls.py:85, 295). */
EC_API int ec_l2norm_rows(const float *x, float *out, int M, int C, void *stream);

/* fp32 LayerNorm rows [M,d] -> fp32 (adapter pre-norm). */
EC_API int ec_layernorm_f32(const float *x, const float *gamma, const float *beta, int M, int d, float *out, void *stream);

/* ---- fine-tune step (models/clip_cls_ft.py:214-269; backward the reference leaves to autograd) ------------------ */

/* LayerNorm backward over rows: dx = acc + rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma (eps 1e-5).
 * x fp32 rows with stride x_stride (the LayerNorm input), dy fp32 [M,d]; acc (nullable) fp32 rows added to the result;
 * dx_bf16 (nullable) bf16 [M,d] copy of the result (A operand of the next data-gradient GEMM). */
EC_API int ec_layernorm_bwd(const float *x, int64_t x_stride, const float *dy, const float *gamma, const float *acc,
                            int64_t acc_stride, int M, int d, float *dx, int64_t dx_stride, void *dx_bf16, void *stream);

/* QuickGELU on a saved bf16 pre-activation (training forward keeps it) and its backward; n elements. */
EC_API int ec_quickgelu(const void *a, void *h, int64_t n, void *stream);
EC_API int ec_quickgelu_bwd(const void *a, const void *dh, void *da, int64_t n, void *stream);

/* out[c, r] = in[r, c] for bf16 matrices (weight-gradient GEMMs reduce over the token dimension). */
EC_API int ec_transpose_bf16(const void *in, void *out, int rows, int cols, int64_t ld_in, int64_t ld_out, void *stream);

/* Training forward of the attention core: ec_attention_ex plus lse fp32 [n_seq, heads, L], the log2-domain log-sum-exp of
 * every score row (p_ij = exp2(s_ij * log2(e)/8 - lse_i)), kept for the backward.  L <= 384 (tensor-memory kernels). */
EC_API int ec_attention_fwd_lse(const void *qkv, void *out, float *lse, int n_seq, int L, int heads, int causal, void *stream);

/* Attention backward for the packed QKV layout of ec_attention: dqkv bf16 [n_img*L, 3d] from qkv, the forward
 * output o bf16 [n_img*L, d] and its gradient do bf16 [n_img*L, d]; scores recomputed per tile (flash-style).
 * lse: the forward's ec_attention_fwd_lse output -> tcgen05 kernel (L <= 256: S, dP and the dQ / dK / dV accumulators in
 * tensor memory); NULL (or L > 256, or EC_ATTN_BWD=mma) -> mma.sync kernel that recomputes the row statistics. */
EC_API int ec_attention_bwd(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int n_img, int L,
                            int heads, void *stream);

/* One Adam update (torch.optim.Adam semantics, method.py:150-191): fp32 parameters, bias correction by `step` >= 1. */
EC_API int ec_adam(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, void *stream);

/* out[m,n] (+)= alpha * sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]: fp32 SIMT GEMM with arbitrary operand strides for the small
 * backward products -- LoRA factor gradients dUp = dW . down^T, dDown = up^T . dW (models/lora.py:138-149 under autograd)
 * and the logit head's d_feats / d_text. */
EC_API int ec_gemm_f32_strided(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn, int M, int N,
                               int K, float alpha, float *out, int64_t ldo, int accumulate, void *stream);

/* LoRA factor gradients from the gradient of the merged weight W_eff = W + up . down (models/lora.py:138-149 q/k/v,
 * 49-52 out_proj; the reference gets them from autograd):  d_up[rows,r] = dW . down^T,  d_down[r,d] = up^T . dW.
 * dW fp32 holds n_mat (<= 4) matrices of `rows` rows stacked (q | k | v of in_proj, or out_proj alone), row stride ld.
 * up / down / d_up / d_down: HOST arrays of n_mat device pointers; a NULL up[z] skips matrix z (e.g. no LoRA on k). */
EC_API int ec_lora_grad(const float *dW, int64_t ld, int n_mat, int rows, int d, int r, const float *const *up,
                        const float *const *down, float *const *d_up, float *const *d_down, void *stream);

/* Column sums over the token dimension: out[c] = sum_r a[r, c] for a fp32 or bf16 matrix [M, ld] (bias gradients of the
 * nn.Linear layers, positional / class embedding gradients).  Two deterministic stages; scratch fp32 [n_part * N]. */
EC_API int ec_colsum(const void *a, int is_bf16, int M, int N, int64_t ld, float *scratch, int n_part, float *out, void *stream);

/* LayerNorm affine gradients: dgamma[c] = sum_r dy[r,c] * xhat[r,c], dbeta[c] = sum_r dy[r,c]; x fp32 rows of stride x_stride
 * (the LayerNorm INPUT, statistics recomputed), dy fp32 [M,d]; scratch fp32 [2 * n_part * d]. */
EC_API int ec_layernorm_param_grad(const float *x, int64_t x_stride, const float *dy, int M, int d, float *scratch, int n_part,
                                   float *dgamma, float *dbeta, void *stream);

/* dst bf16 [n_img*G2, d] = the patch-token rows of the fp32 token matrix src [n_img*(G2+1), d] (class-token rows skipped):
 * the A operand of conv1's weight gradient. */
EC_API int ec_patch_rows_bf16(const float *src, int n_img, int G2, int d, void *dst, void *stream);

/* Backward of F.normalize(x, p=2, dim=-1) followed by the valid-mask multiply (clip_cls_ft.py:229-232; text features
 * :163): dx = (dy - y<y,dy>) / max(|x|, 1e-12); rows with mask == 0 get dx = 0.  mask nullable. */
EC_API int ec_l2norm_rows_bwd(const float *x, const float *dy, const uint8_t *mask, int M, int C, float *dx, void *stream);

/* Loss of the fine-tune step and its gradient: logits = aggregate(full_logits) (clip_cls_ft.py:191-202, sum or mean),
 * loss = F.cross_entropy(logits, labels) (:258-264).  full_logits fp32 [B,T,n_cls] with zero rows for invalid views,
 * labels device int32 [B].  Outputs: per-sample loss [B], mean loss [1], d(mean loss)/d(full_logits) [B,T,n_cls]. */
EC_API int ec_ce_loss_bwd(const float *full_logits, const uint8_t *valid, const int32_t *labels, int B, int T, int n_cls, int agg,
                          float *loss_per_sample, float *loss_mean, float *d_full, void *stream);

/* The reference's other loss (loss_dict['use_probs_loss'], models/clip_cls_ft.py:265-267, clip_cls.py:173-175):
 * probs = mean over the valid views of softmax(full_logits[b,t,:]) (clip_cls.py:123-129), loss = F.nll_loss(log(probs + 1e-6)).
 * Same outputs as ec_ce_loss_bwd; padded views receive a zero gradient.  T <= 16. */
EC_API int ec_probs_loss_bwd(const float *full_logits, const uint8_t *valid, const int32_t *labels, int B, int T, int n_cls,
                             float *loss_per_sample, float *loss_mean, float *d_full, void *stream);

#ifdef __cplusplus
}
#endif
#endif
