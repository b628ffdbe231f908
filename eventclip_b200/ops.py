"""Tensor-level wrappers over the C ABI.  PyTorch is plumbing here (device memory, streams); all arithmetic happens
in libeventclip_b200.so.  Every function enqueues on torch's current stream and returns immediately."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

OUT_FMT = {"f32": L.EC_OUT_F32_NCHW, "bf16": L.EC_OUT_BF16_NCHW, "patch": L.EC_OUT_BF16_PATCH, "patch_f16": L.EC_OUT_F16_PATCH,
           "gray": L.EC_OUT_GRAY_BF16_PATCH, "gray_f16": L.EC_OUT_GRAY_F16_PATCH}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _dev(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise L.ECError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if not t.is_contiguous():
        raise L.ECError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise L.ECError(f"{name} must be {dtype}, got {t.dtype}")
    L.require_device(t.device.index)
    return t


# ------------------------------------------------------------------------------------------------ event2img
def plan_frames(offsets, N, T, sel=None, compact=False):
    """Host planning (ec_plan_frames).  offsets: int64 [B+1] (CPU tensor / array).
    Returns (frames uint8 CPU tensor [n_frames,16] (pinned when CUDA is available), valid bool [B,T] CPU,
    chunks int32 [B] CPU, n_valid)."""
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    B = off.shape[0] - 1
    cap = B * T
    pin = torch.cuda.is_available()
    frames = torch.empty((max(cap, 1), 16), dtype=torch.uint8, pin_memory=pin)
    valid = np.zeros((B, T), np.uint8)
    chunks = np.zeros(B, np.int32)
    selp = C.c_void_p(0)
    if sel is not None:
        sel = np.ascontiguousarray(np.asarray(sel, dtype=np.int32))
        if sel.shape != (B, T):
            raise L.ECError(f"sel must have shape {(B, T)}, got {sel.shape}")
        selp = C.c_void_p(sel.ctypes.data)
    nf, nv = C.c_int(0), C.c_int(0)
    rc = L.load().ec_plan_frames(C.c_void_p(off.ctypes.data), B, int(N), int(T), selp, int(bool(compact)),
                                 C.c_void_p(frames.data_ptr()), cap, C.c_void_p(valid.ctypes.data),
                                 C.c_void_p(chunks.ctypes.data), C.byref(nf), C.byref(nv))
    if rc == L.EC_ERR_ARG and "no events" in L.load().ec_last_error().decode():
        raise AssertionError(L.load().ec_last_error().decode())
    L.check(rc, "ec_plan_frames")
    return frames[:nf.value], torch.from_numpy(valid.astype(bool)), torch.from_numpy(chunks), nv.value


def event2img(events, frames, shape, n_slots, count_non_zero=False, background_mask=True, out="f32", patch=0,
              ldk=0, out_tensor=None, debug=False, status=None):
    """Fused frames (ec_event2img).  events: CUDA float32 [E,4], or CUDA int32 [E] words of the compact wire format
    (pack_events, row F2); frames: CUDA uint8 [n_frames,16].
    Returns (images, status int32[1] CUDA tensor, debug dict or None)."""
    compact = events.dtype == torch.int32
    _dev(events, torch.int32 if compact else torch.float32, "events")
    _dev(frames, torch.uint8, "frames")
    if compact:
        if events.dim() != 1:
            raise L.ECError("compact events must be a 1-D int32 tensor of packed words")
    elif events.dim() != 2 or events.shape[1] != 4:
        raise L.ECError("events must be [E, 4] rows of (x, y, t, p)")
    H, W = shape
    n_frames = frames.shape[0]
    fmt = OUT_FMT[out]
    dev = events.device
    if out_tensor is None:
        if fmt == L.EC_OUT_F32_NCHW:
            out_tensor = torch.empty((n_slots, 3, 224, 224), dtype=torch.float32, device=dev)
        elif fmt == L.EC_OUT_BF16_NCHW:
            out_tensor = torch.empty((n_slots, 3, 224, 224), dtype=torch.bfloat16, device=dev)
        else:
            G = 224 // patch
            gray = fmt in (L.EC_OUT_GRAY_BF16_PATCH, L.EC_OUT_GRAY_F16_PATCH)
            ldk = ldk or (1 if gray else 3) * patch * patch
            out_tensor = torch.zeros((n_slots * G * G, ldk), device=dev,
                                     dtype=torch.float16 if fmt in (L.EC_OUT_F16_PATCH, L.EC_OUT_GRAY_F16_PATCH) else torch.bfloat16)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    dbg = None
    dc = dg = du = None
    if debug:
        dc = torch.zeros((n_frames, H, W, 2), dtype=torch.int32, device=dev)
        dg = torch.zeros((n_frames, H, W), dtype=torch.uint8, device=dev)
        du = torch.zeros((n_frames, 224, 224), dtype=torch.uint8, device=dev)
        dbg = dict(counts=dc, gray=dg, u8=du)
    flags = (L.EC_FLAG_COUNT_NON_ZERO if count_non_zero else 0) | (L.EC_FLAG_BACKGROUND_MASK if background_mask else 0)
    with torch.cuda.device(dev):
        fn = L.load().ec_event2img_compact if compact else L.load().ec_event2img
        rc = fn(_ptr(events), _ptr(frames), n_frames, H, W, flags, fmt, int(patch), int(ldk),
                _ptr(out_tensor), _ptr(dc), _ptr(dg), _ptr(du), _ptr(status), _stream())
    L.check(rc, "ec_event2img")
    return out_tensor, status, dbg


def pack_events(events, shape):
    """float32 [E,4] events -> int32 [E] compact words (ec_pack_events): flat pixel index | polarity code << 30."""
    _dev(events, torch.float32, "events")
    H, W = shape
    out = torch.empty(events.shape[0], dtype=torch.int32, device=events.device)
    with torch.cuda.device(events.device):
        L.check(L.load().ec_pack_events(_ptr(events), events.shape[0], H, W, _ptr(out), _stream()), "ec_pack_events")
    return out


def raise_on_status(status):
    """Synchronising check of the device status word; mirrors the reference's exception types."""
    s = int(status.item())
    if s & L.EC_STATUS_BAD_COORD:
        # np.bincount / reshape raise ValueError on such input (datasets/vis.py:12-14)
        raise ValueError("event coordinates outside the sensor: x + y*W must lie in [0, H*W)")
    if s & L.EC_STATUS_COUNT_OVERFLOW:
        raise L.ECError("a per-pixel event count exceeded 65535 in one frame (packed histogram overflow)")


def center_events(events, offsets_dev, shape):
    """In place on the device (ec_center_events); offsets_dev: CUDA int64 [B+1]."""
    _dev(events, torch.float32, "events")
    _dev(offsets_dev, torch.int64, "offsets")
    with torch.cuda.device(events.device):
        L.check(L.load().ec_center_events(_ptr(events), _ptr(offsets_dev), offsets_dev.numel() - 1, shape[0], shape[1],
                                          _stream()), "ec_center_events")
    return events


def flip_events(events, offsets_dev, W, hflip=False, tflip=False):
    """Out of place (ec_flip_events): the reference's h-/t-flip TTA variants of a packed batch."""
    _dev(events, torch.float32, "events")
    _dev(offsets_dev, torch.int64, "offsets")
    out = torch.empty_like(events)
    with torch.cuda.device(events.device):
        L.check(L.load().ec_flip_events(_ptr(events), _ptr(out), _ptr(offsets_dev), offsets_dev.numel() - 1, int(W),
                                        int(bool(hflip)), int(bool(tflip)), _stream()), "ec_flip_events")
    return out


def event2img_geometry(shape):
    cs, nt, sm = C.c_int(), C.c_int(), C.c_int()
    L.check(L.load().ec_event2img_geometry(shape[0], shape[1], C.byref(cs), C.byref(nt), C.byref(sm)),
            "ec_event2img_geometry")
    return dict(cluster=cs.value, threads=nt.value, smem=sm.value)


# ------------------------------------------------------------------------------------------------ encoder pieces
def gemm_bf16(A, W, bias=None, epi="bf16", out=None, res=None, row_map=0, M=None):
    """out = epilogue(A @ W.T).  A [M,K] (row stride may exceed K) and W [N,K] are both bf16 or both fp16; with fp16 operands the
    16-bit outputs ("bf16", "bf16_qgelu" epilogues) are fp16 as well (EC_EPI_F16_OPERANDS)."""
    f16 = A.dtype == torch.float16
    _dev(A, torch.float16 if f16 else torch.bfloat16, "A")
    _dev(W, A.dtype, "W")
    M = A.shape[0] if M is None else M
    N, K = W.shape
    epi_id = {"bf16": L.EC_EPI_BF16, "bf16_qgelu": L.EC_EPI_BF16_QGELU, "f32_resadd": L.EC_EPI_F32_RESADD,
              "f32": L.EC_EPI_F32, "patch": L.EC_EPI_PATCH, "f16_resadd": L.EC_EPI_F16_RESADD}[epi]
    flag = L.EC_EPI_F16_OPERANDS if f16 else 0
    if epi == "f16_resadd":
        if out is None or res is None or out.dtype != torch.float16 or res.dtype != torch.float16:
            raise L.ECError("f16_resadd epilogue needs fp16 `out` and `res` (the fp16 residual stream)")
    if out is None:
        if epi == "patch":
            raise L.ECError("patch epilogue needs a preallocated token matrix")
        out = torch.empty((M, N), dtype=A.dtype if epi_id <= 1 else torch.float32, device=A.device)
    elif epi_id <= 1 and out.dtype != A.dtype:
        raise L.ECError(f"16-bit epilogue writes {A.dtype} (the operands' dtype), got an output of {out.dtype}")
    with torch.cuda.device(A.device):
        rc = L.load().ec_gemm_bf16(_ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(bias), M, N, K, epi_id | flag,
                                   _ptr(out), out.stride(0), _ptr(res), int(row_map), _stream())
    L.check(rc, "ec_gemm_bf16")
    return out


def gemm_stats_parts(N):
    """float2 statistics slots per row written by gemm_bf16_stats for an N-column output."""
    return int(L.load().ec_gemm_stats_parts(int(N)))


def gemm_bf16_stats(A, W, bias, x, stats):
    """x (fp16 [M,N], in place) += A @ W.T + bias, and stats float32 [M, parts, 2] = per-row partial (sum, sum of squares) of
    the new x: what the LayerNorm-folded GEMM that reads x next needs.  A and W: both bf16 or both fp16."""
    f16 = A.dtype == torch.float16
    _dev(A, torch.float16 if f16 else torch.bfloat16, "A")
    _dev(W, A.dtype, "W")
    _dev(x, torch.float16, "x")
    _dev(stats, torch.float32, "stats")
    M, (N, K) = A.shape[0], W.shape
    if stats.numel() < M * gemm_stats_parts(N) * 2:
        raise L.ECError("gemm_bf16_stats: statistics buffer too small")
    with torch.cuda.device(A.device):
        L.check(L.load().ec_gemm_bf16_stats(_ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(bias), M, N, K, _ptr(x), x.stride(0),
                                            _ptr(x), _ptr(stats), int(f16), _stream()), "ec_gemm_bf16_stats")
    return x


def gemm_bf16_stats2(A, W, bias, x2, stats):
    """The same update on a residual stream kept as two fp16 planes x2 [2, M, N] (x = x2[0] + x2[1], hi = fp16(x), lo = fp16(x - hi)),
    both updated in place; stats as in gemm_bf16_stats."""
    f16 = A.dtype == torch.float16
    _dev(A, torch.float16 if f16 else torch.bfloat16, "A")
    _dev(W, A.dtype, "W")
    _dev(x2, torch.float16, "x2")
    _dev(stats, torch.float32, "stats")
    M, (N, K) = A.shape[0], W.shape
    if x2.dim() != 3 or x2.shape[0] != 2 or x2.shape[1] != M or x2.shape[2] != N or not x2.is_contiguous():
        raise L.ECError("gemm_bf16_stats2: x2 must be a contiguous fp16 [2, M, N] tensor")
    if stats.numel() < M * gemm_stats_parts(N) * 2:
        raise L.ECError("gemm_bf16_stats2: statistics buffer too small")
    with torch.cuda.device(A.device):
        L.check(L.load().ec_gemm_bf16_stats2(_ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(bias), M, N, K, _ptr(x2[0]), _ptr(x2[1]),
                                             N, _ptr(stats), int(f16), _stream()), "ec_gemm_bf16_stats2")
    return x2


def layernorm_f16x2(x, gamma, beta, M, d, out2, row_stride=None):
    """LayerNorm of fp32 rows into the (hi, lo) fp16 pair out2 [2, M, d]."""
    _dev(x, torch.float32, "x")
    _dev(out2, torch.float16, "out2")
    with torch.cuda.device(x.device):
        L.check(L.load().ec_layernorm_f16x2(_ptr(x), int(row_stride or d), _ptr(gamma), _ptr(beta), M, d, _ptr(out2[0]), _ptr(out2[1]),
                                            _stream()), "ec_layernorm_f16x2")
    return out2


def gemm_ln(x, Wg, colsum, cbias, stats, n_parts, epi="bf16", out=None, out_dtype=torch.bfloat16):
    """out = epi(LayerNorm(x) @ W.T + b) with the LayerNorm folded into the GEMM: x fp16 [M,K] is the A operand itself,
    Wg = fp16(gamma * W), colsum[j] = sum_k Wg[j,k], cbias[j] = beta . W[j] + b[j], stats = per-row partial sums.
    The output is bf16, or fp16 when `out` / `out_dtype` says so."""
    _dev(x, torch.float16, "x")
    _dev(Wg, torch.float16, "Wg")
    _dev(stats, torch.float32, "stats")
    M, (N, K) = x.shape[0], Wg.shape
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    epi_id = {"bf16": L.EC_EPI_BF16, "bf16_qgelu": L.EC_EPI_BF16_QGELU}[epi]
    if out.dtype == torch.float16:
        epi_id |= L.EC_EPI_F16_OPERANDS
    with torch.cuda.device(x.device):
        L.check(L.load().ec_gemm_ln(_ptr(x), x.stride(0), _ptr(Wg), Wg.stride(0), _ptr(colsum), _ptr(cbias), _ptr(stats), int(n_parts),
                                    M, N, K, epi_id, _ptr(out), out.stride(0), _stream()), "ec_gemm_ln")
    return out


def row_stats_f16(x, stats, n_parts, M=None, row_stride=None):
    """stats[row, 0] = (sum, sum of squares) of the fp16 row; the other parts are zeroed."""
    _dev(x, torch.float16, "x")
    _dev(stats, torch.float32, "stats")
    M = x.shape[0] if M is None else M
    d = x.shape[-1]
    with torch.cuda.device(x.device):
        L.check(L.load().ec_row_stats_f16(_ptr(x), int(row_stride or d), M, d, _ptr(stats), int(n_parts), _stream()), "ec_row_stats_f16")
    return stats


def layernorm(x, gamma, beta, M, d, row_stride=None, out_bf16=None, out_f32=None, out_f16=None):
    """x: fp32, or fp16 (the fp16 residual stream); out_f16 is what ln_pre writes to start that stream."""
    if x.dtype == torch.float16 or out_f16 is not None:
        _dev(x, x.dtype if x.dtype in (torch.float16, torch.float32) else torch.float32, "x")
        with torch.cuda.device(x.device):
            rc = L.load().ec_layernorm_ex(_ptr(x), int(x.dtype == torch.float16), int(row_stride or d), _ptr(gamma), _ptr(beta),
                                          M, d, _ptr(out_bf16), _ptr(out_f32), _ptr(out_f16), _stream())
        L.check(rc, "ec_layernorm_ex")
        return
    _dev(x, torch.float32, "x")
    with torch.cuda.device(x.device):
        rc = L.load().ec_layernorm(_ptr(x), int(row_stride or d), _ptr(gamma), _ptr(beta), M, d, _ptr(out_bf16),
                                   _ptr(out_f32), _stream())
    L.check(rc, "ec_layernorm")


def attention(qkv, out, n_img, Ltok, heads, causal=False):
    """softmax(q k^T / 8) v on packed qkv [n*L, 3d]; qkv and out are both bf16 or both fp16."""
    f16 = qkv.dtype == torch.float16
    _dev(qkv, torch.float16 if f16 else torch.bfloat16, "qkv")
    _dev(out, qkv.dtype, "out")
    flags = (L.EC_ATTN_CAUSAL if causal else 0) | (L.EC_ATTN_F16 if f16 else 0)
    with torch.cuda.device(qkv.device):
        rc = L.load().ec_attention_ex(_ptr(qkv), _ptr(out), n_img, Ltok, heads, flags, _stream())
    L.check(rc, "ec_attention")
    return out


def embed_tokens(table, tokens_i32, pos):
    """x[n*L + l] = table[tokens[n,l]] + pos[l]  (fp32 [n*L, d])."""
    _dev(table, torch.float32, "table")
    _dev(tokens_i32, torch.int32, "tokens")
    n, Lc = tokens_i32.shape
    d = table.shape[1]
    out = torch.empty((n * Lc, d), dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        L.check(L.load().ec_embed_tokens(_ptr(table), _ptr(tokens_i32), _ptr(pos), _ptr(out), n, Lc, d, table.shape[0],
                                         _stream()), "ec_embed_tokens")
    return out


def cls_rows(x, cls, pos, n_img, Ltok, d):
    _dev(x, torch.float32, "x")
    with torch.cuda.device(x.device):
        L.check(L.load().ec_cls_rows(_ptr(x), _ptr(cls), _ptr(pos), n_img, Ltok, d, _stream()), "ec_cls_rows")


def f32_to_bf16(src, dst=None):
    _dev(src, torch.float32, "src")
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        L.check(L.load().ec_f32_to_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "ec_f32_to_bf16")
    return dst


def im2col(img, patch, ldk=None, dtype=torch.bfloat16):
    """img CUDA [n,3,224,224] fp32 or bf16 -> bf16 (or fp16) [n*G*G, ldk]."""
    _dev(img, None, "img")
    if img.dtype not in (torch.float32, torch.bfloat16):
        raise L.ECError(f"img must be float32 or bfloat16, got {img.dtype}")
    n = img.shape[0]
    G = 224 // patch
    K = 3 * patch * patch
    ldk = ldk or K
    out = (torch.zeros if ldk != K else torch.empty)((n * G * G, ldk), dtype=dtype, device=img.device)
    with torch.cuda.device(img.device):
        L.check(L.load().ec_im2col(_ptr(img), int(img.dtype == torch.bfloat16) | (2 if dtype == torch.float16 else 0), n, patch, ldk,
                                   _ptr(out), _stream()), "ec_im2col")
    return out


def lora_merge(W, up, down, out=None):
    """bf16(W + up @ down) in fp32 math (models/lora.py:138-149)."""
    _dev(W, torch.float32, "W")
    rows, d = W.shape
    r = up.shape[1] if up is not None else 0
    if out is None:
        out = torch.empty((rows, d), dtype=torch.bfloat16, device=W.device)
    with torch.cuda.device(W.device):
        L.check(L.load().ec_lora_merge(_ptr(W), _ptr(up), _ptr(down), rows, d, r, _ptr(out), _stream()), "ec_lora_merge")
    return out


# ------------------------------------------------------------------------------------------------ heads
def head(feats, valid_u8, text, B, T, scale, normalize, agg, want_top=True):
    _dev(feats, torch.float32, "feats")
    _dev(text, torch.float32, "text")
    _dev(valid_u8, torch.uint8, "valid")
    C_ = feats.shape[-1]
    n_cls = text.shape[0]
    dev = feats.device
    full = torch.empty((B, T, n_cls), dtype=torch.float32, device=dev)
    logits = torch.empty((B, n_cls), dtype=torch.float32, device=dev)
    probs = torch.empty((B, n_cls), dtype=torch.float32, device=dev)
    top = torch.empty((B, 2, 5), dtype=torch.int32, device=dev) if want_top else None
    with torch.cuda.device(dev):
        rc = L.load().ec_head(_ptr(feats), _ptr(valid_u8), _ptr(text), B, T, C_, n_cls, float(scale), int(normalize),
                              L.EC_AGG[agg], _ptr(full), _ptr(logits), _ptr(probs), _ptr(top), _stream())
    L.check(rc, "ec_head")
    return full, logits, probs, top


def gemm_f32(A, W, bias=None, res=None, act=0):
    _dev(A, torch.float32, "A")
    _dev(W, torch.float32, "W")
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().ec_gemm_f32(_ptr(A), _ptr(W), _ptr(bias), _ptr(res), M, N, K, act, _ptr(out), _stream()),
                "ec_gemm_f32")
    return out


def adapter_attention(qkv, valid_u8, B, T, D, heads):
    _dev(qkv, torch.float32, "qkv")
    out = torch.empty((B * T, D), dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        L.check(L.load().ec_adapter_attention(_ptr(qkv), _ptr(valid_u8), B, T, D, heads, _ptr(out), _stream()),
                "ec_adapter_attention")
    return out


def adapter_attention_bwd(qkv, valid_u8, d_out, B, T, D, heads):
    """d_qkv of ec_adapter_attention (few-shot adapter training)."""
    _dev(qkv, torch.float32, "qkv")
    _dev(d_out, torch.float32, "d_out")
    d_qkv = torch.empty_like(qkv)
    with torch.cuda.device(qkv.device):
        L.check(L.load().ec_adapter_attention_bwd(_ptr(qkv), _ptr(valid_u8), _ptr(d_out), B, T, D, heads, _ptr(d_qkv), _stream()),
                "ec_adapter_attention_bwd")
    return d_qkv


def relu_bwd(y, dy):
    """dy masked by y > 0 (y = output of the ReLU fused into gemm_f32(act=1))."""
    _dev(y, torch.float32, "y")
    _dev(dy, torch.float32, "dy")
    dx = torch.empty_like(y)
    with torch.cuda.device(y.device):
        L.check(L.load().ec_relu_bwd(_ptr(y), _ptr(dy), _ptr(dx), y.numel(), _stream()), "ec_relu_bwd")
    return dx


def blend(a, b, r):
    _dev(a, torch.float32, "a")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        L.check(L.load().ec_blend(_ptr(a), _ptr(b), float(r), _ptr(out), a.numel(), _stream()), "ec_blend")
    return out


def layernorm_f32(x, gamma, beta):
    _dev(x, torch.float32, "x")
    M, d = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.load().ec_layernorm_f32(_ptr(x), _ptr(gamma), _ptr(beta), M, d, _ptr(out), _stream()),
                "ec_layernorm_f32")
    return out


def gather_rows(src, idx_i32, n_rows):
    """dst[s] = src[idx[s]] or zeros where idx[s] < 0."""
    _dev(src, torch.float32, "src")
    _dev(idx_i32, torch.int32, "idx")
    Cdim = src.shape[-1]
    dst = torch.empty((n_rows, Cdim), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        L.check(L.load().ec_gather_rows(_ptr(src), _ptr(idx_i32), _ptr(dst), n_rows, Cdim, _stream()), "ec_gather_rows")
    return dst


def l2norm_rows(x):
    _dev(x, torch.float32, "x")
    M, Cdim = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.load().ec_l2norm_rows(_ptr(x), _ptr(out), M, Cdim, _stream()), "ec_l2norm_rows")
    return out


# ------------------------------------------------------------------------------------------------ fine-tune step
def layernorm_bwd(x, dy, gamma, M, d, acc=None, x_stride=None, dx=None, dx_stride=None, acc_stride=None, dx_bf16=None):
    """dx = (acc +) LayerNorm'(x) . dy; strides let ln_post touch only the class-token rows; dx_bf16: optional bf16 copy."""
    _dev(x, torch.float32, "x")
    _dev(dy, torch.float32, "dy")
    if dx is None:
        dx = torch.empty((M, d), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().ec_layernorm_bwd(_ptr(x), int(x_stride or d), _ptr(dy), _ptr(gamma), _ptr(acc), int(acc_stride or d),
                                          M, d, _ptr(dx), int(dx_stride or d), _ptr(dx_bf16), _stream()), "ec_layernorm_bwd")
    return dx


def quickgelu(a, out=None):
    _dev(a, torch.bfloat16, "a")
    out = torch.empty_like(a) if out is None else out
    with torch.cuda.device(a.device):
        L.check(L.load().ec_quickgelu(_ptr(a), _ptr(out), a.numel(), _stream()), "ec_quickgelu")
    return out


def quickgelu_bwd(a, dh, out=None):
    _dev(a, torch.bfloat16, "a")
    _dev(dh, torch.bfloat16, "dh")
    out = torch.empty_like(a) if out is None else out
    with torch.cuda.device(a.device):
        L.check(L.load().ec_quickgelu_bwd(_ptr(a), _ptr(dh), _ptr(out), a.numel(), _stream()), "ec_quickgelu_bwd")
    return out


def transpose_bf16(x, out=None, pad=8):
    """x bf16 [R, Ccols] -> bf16 [Ccols, Rp] with Rp = R rounded up to `pad` (zero filled): the K-major operand of a
    weight-gradient GEMM."""
    _dev(x, torch.bfloat16, "x")
    R, Cc = x.shape
    Rp = (R + pad - 1) // pad * pad
    if out is None:
        out = torch.zeros((Cc, Rp), dtype=torch.bfloat16, device=x.device) if Rp != R else \
            torch.empty((Cc, Rp), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().ec_transpose_bf16(_ptr(x), _ptr(out), R, Cc, x.stride(0), out.stride(0), _stream()),
                "ec_transpose_bf16")
    return out


def attention_fwd_lse(qkv, out, n_img, Ltok, heads, causal=False):
    """Training forward: attention plus the per-row log-sum-exp [n_img, heads, Ltok] the tcgen05 backward needs."""
    _dev(qkv, torch.bfloat16, "qkv")
    lse = torch.empty((n_img, heads, Ltok), dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        L.check(L.load().ec_attention_fwd_lse(_ptr(qkv), _ptr(out), _ptr(lse), n_img, Ltok, heads, int(bool(causal)), _stream()),
                "ec_attention_fwd_lse")
    return lse


def attention_bwd(qkv, o, d_o, n_img, Ltok, heads, out=None, lse=None):
    _dev(qkv, torch.bfloat16, "qkv")
    _dev(o, torch.bfloat16, "o")
    _dev(d_o, torch.bfloat16, "d_o")
    out = torch.empty_like(qkv) if out is None else out
    with torch.cuda.device(qkv.device):
        L.check(L.load().ec_attention_bwd(_ptr(qkv), _ptr(o), _ptr(d_o), _ptr(lse), _ptr(out), n_img, Ltok, heads, _stream()),
                "ec_attention_bwd")
    return out


def adam(param, grad, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _dev(t, torch.float32, n)
    with torch.cuda.device(param.device):
        L.check(L.load().ec_adam(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr),
                                 float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), _stream()),
                "ec_adam")


def mm_f32(A, B, trans_a=False, trans_b=False, out=None, accumulate=False, alpha=1.0):
    """fp32 matrix product of 2-D views (any strides): out (+)= alpha * op(A) @ op(B)."""
    if trans_a:
        A = A.t()
    if trans_b:
        B = B.t()
    M, K = A.shape
    K2, N = B.shape
    if K != K2:
        raise L.ECError(f"mm_f32: inner dimensions differ ({K} vs {K2})")
    if not (A.is_cuda and B.is_cuda and A.dtype == B.dtype == torch.float32):
        raise L.ECError("mm_f32 needs fp32 CUDA tensors (no CPU fallback exists)")
    L.require_device(A.device.index)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().ec_gemm_f32_strided(_ptr(A), A.stride(0), A.stride(1), _ptr(B), B.stride(0), B.stride(1), M, N, K,
                                             float(alpha), _ptr(out), out.stride(0), int(bool(accumulate)), _stream()),
                "ec_gemm_f32_strided")
    return out


def l2norm_rows_bwd(x, dy, mask_u8=None):
    _dev(x, torch.float32, "x")
    _dev(dy, torch.float32, "dy")
    M, Cdim = x.shape
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.load().ec_l2norm_rows_bwd(_ptr(x), _ptr(dy), _ptr(mask_u8), M, Cdim, _ptr(dx), _stream()),
                "ec_l2norm_rows_bwd")
    return dx


def probs_loss_bwd(full_logits, valid_u8, labels_i32):
    """The reference's probability loss (use_probs_loss, clip_cls_ft.py:265-267): nll of log(mean-over-views softmax + 1e-6).
    Returns (per-sample loss [B], mean loss [1], d mean-loss / d full_logits [B,T,K])."""
    _dev(full_logits, torch.float32, "full_logits")
    _dev(labels_i32, torch.int32, "labels")
    B, T, K = full_logits.shape
    loss_b = torch.empty(B, dtype=torch.float32, device=full_logits.device)
    loss = torch.empty(1, dtype=torch.float32, device=full_logits.device)
    dfull = torch.empty_like(full_logits)
    with torch.cuda.device(full_logits.device):
        L.check(L.load().ec_probs_loss_bwd(_ptr(full_logits), _ptr(valid_u8), _ptr(labels_i32), B, T, K, _ptr(loss_b),
                                           _ptr(loss), _ptr(dfull), _stream()), "ec_probs_loss_bwd")
    return loss_b, loss, dfull


def ce_loss_bwd(full_logits, valid_u8, labels_i32, agg):
    """Returns (per-sample loss [B], mean loss [1], d mean-loss / d full_logits [B,T,K])."""
    _dev(full_logits, torch.float32, "full_logits")
    _dev(labels_i32, torch.int32, "labels")
    B, T, K = full_logits.shape
    loss_b = torch.empty(B, dtype=torch.float32, device=full_logits.device)
    loss = torch.empty(1, dtype=torch.float32, device=full_logits.device)
    dfull = torch.empty_like(full_logits)
    agg_id = {"sum": L.EC_AGG["sum"], "mean": L.EC_AGG["mean"]}.get(agg)
    if agg_id is None:
        raise L.ECError(f"training supports agg_func 'sum' or 'mean', got {agg!r}")
    with torch.cuda.device(full_logits.device):
        L.check(L.load().ec_ce_loss_bwd(_ptr(full_logits), _ptr(valid_u8), _ptr(labels_i32), B, T, K, agg_id, _ptr(loss_b),
                                        _ptr(loss), _ptr(dfull), _stream()), "ec_ce_loss_bwd")
    return loss_b, loss, dfull


def lora_grad(dW, rows, factors, dests=None):
    """factors: list (<= 4) of (up [rows,r], down [r,d]) or None per stacked matrix of dW [n*rows, d]; dests: matching list
    of (d_up, d_down) output tensors or None.  Returns the list of (d_up, d_down) (None where skipped)."""
    _dev(dW, torch.float32, "dW")
    n = len(factors)
    d = dW.shape[1]
    outs, arrs = [], [(C.c_void_p * n)() for _ in range(4)]
    r = 0
    for z, f in enumerate(factors):
        if f is None:
            outs.append(None)
            continue
        up, down = f
        r = up.shape[1]
        dst = dests[z] if dests is not None and dests[z] is not None else (None, None)
        du = dst[0] if dst[0] is not None else torch.empty((rows, r), dtype=torch.float32, device=dW.device)
        dd = dst[1] if dst[1] is not None else torch.empty((r, d), dtype=torch.float32, device=dW.device)
        for t, nme in ((up, "up"), (down, "down"), (du, "d_up"), (dd, "d_down")):
            _dev(t, torch.float32, nme)
        arrs[0][z], arrs[1][z], arrs[2][z], arrs[3][z] = up.data_ptr(), down.data_ptr(), du.data_ptr(), dd.data_ptr()
        outs.append((du, dd))
    if r == 0:
        return outs
    with torch.cuda.device(dW.device):
        L.check(L.load().ec_lora_grad(_ptr(dW), dW.stride(0), n, rows, d, r, arrs[0], arrs[1], arrs[2], arrs[3], _stream()),
                "ec_lora_grad")
    return outs


def gemm_bf16_splitk(A, W, splits=None, out=None, M=None, K=None):
    """fp32 out = A @ W.T for weight-gradient shapes (small M x N, long K): split-K over the SMs, deterministic sum."""
    _dev(A, torch.bfloat16, "A")
    _dev(W, torch.bfloat16, "W")
    M = A.shape[0] if M is None else M
    N = W.shape[0]
    K = W.shape[1] if K is None else K
    lib = L.load()
    if splits is None:
        with torch.cuda.device(A.device):
            splits = lib.ec_gemm_splitk_choose(M, N, K)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    ws = torch.empty((splits, M, N), dtype=torch.float32, device=A.device) if splits > 1 else None
    with torch.cuda.device(A.device):
        L.check(lib.ec_gemm_bf16_splitk(_ptr(A), A.stride(0), _ptr(W), W.stride(0), M, N, K, int(splits), _ptr(ws), _ptr(out),
                                        out.stride(0), _stream()), "ec_gemm_bf16_splitk")
    return out


def colsum(a, out=None, M=None, n_part=None):
    """out[c] = sum_r a[r, c] (fp32 result) for a 2-D fp32 / bf16 tensor (row stride may exceed the width)."""
    if a.dtype not in (torch.float32, torch.bfloat16) or not a.is_cuda or a.stride(1) != 1:
        raise L.ECError("colsum needs a CUDA fp32 / bf16 matrix with unit column stride")
    L.require_device(a.device.index)
    M = a.shape[0] if M is None else M
    N = a.shape[1]
    n_part = n_part or max(1, min(512, M // 64))
    scratch = torch.empty((n_part, N), dtype=torch.float32, device=a.device)
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        L.check(L.load().ec_colsum(_ptr(a), int(a.dtype == torch.bfloat16), M, N, a.stride(0), _ptr(scratch), n_part, _ptr(out),
                                   _stream()), "ec_colsum")
    return out


def layernorm_param_grad(x, dy, M, d, x_stride=None, dgamma=None, dbeta=None, n_part=None):
    _dev(x, torch.float32, "x")
    _dev(dy, torch.float32, "dy")
    n_part = n_part or max(1, min(296, M // 32))
    scratch = torch.empty((2, n_part, d), dtype=torch.float32, device=x.device)
    dgamma = torch.empty(d, dtype=torch.float32, device=x.device) if dgamma is None else dgamma
    dbeta = torch.empty(d, dtype=torch.float32, device=x.device) if dbeta is None else dbeta
    with torch.cuda.device(x.device):
        L.check(L.load().ec_layernorm_param_grad(_ptr(x), int(x_stride or d), _ptr(dy), M, d, _ptr(scratch), n_part, _ptr(dgamma),
                                                 _ptr(dbeta), _stream()), "ec_layernorm_param_grad")
    return dgamma, dbeta


def patch_rows_bf16(x, n_img, G2, d):
    _dev(x, torch.float32, "x")
    out = torch.empty((n_img * G2, d), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().ec_patch_rows_bf16(_ptr(x), n_img, G2, d, _ptr(out), _stream()), "ec_patch_rows_bf16")
    return out


def gemm_bf16_tn(A, B, splits=None, out=None):
    """fp32 out [M,N] = A.T @ B for token-major bf16 matrices A [K,M], B [K,N] (weight gradients dY^T X); no transposes.
    A / B may be column blocks of wider matrices (unit column stride, row stride a multiple of 8)."""
    for t, n in ((A, "A"), (B, "B")):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1):
            raise L.ECError(f"{n} must be a CUDA bf16 matrix with unit column stride (no CPU fallback exists)")
    L.require_device(A.device.index)
    K, M = A.shape
    N = B.shape[1]
    if B.shape[0] != K:
        raise L.ECError(f"gemm_bf16_tn: row counts differ ({K} vs {B.shape[0]})")
    lib = L.load()
    if splits is None:
        with torch.cuda.device(A.device):
            splits = lib.ec_gemm_splitk_choose(M, N, max(K, 64))
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    ws = torch.empty((splits, M, N), dtype=torch.float32, device=A.device) if splits > 1 else None
    with torch.cuda.device(A.device):
        L.check(lib.ec_gemm_bf16_tn_splitk(_ptr(A), A.stride(0), _ptr(B), B.stride(0), M, N, K, int(splits), _ptr(ws), _ptr(out),
                                           out.stride(0), _stream()), "ec_gemm_bf16_tn_splitk")
    return out
