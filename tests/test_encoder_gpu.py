"""GPU parity of the CLIP ViT image encoder pieces (through the C ABI) against plain fp32 PyTorch on the CPU.

Tolerances (bf16 operands, fp32 accumulation, fp32 residual stream):
  GEMM / attention vs an fp32 reference fed the SAME bf16-rounded operands: rel-L2 <= 2e-3 (accumulation order only)
  whole encoder vs the fp32 oracle: rel-L2 <= 2e-2 (SURVEY.md section 8(c) pin 7)
"""
import numpy as np
import pytest
import torch

from eventclip_b200 import clip, ops
from oracle import clip_oracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (256, 512, 768), (1000, 768, 3072), (197 * 3, 2304, 768),
                                   (50, 64, 128), (130, 384, 592), (4096, 1024, 1024), (77, 512, 768)])
def test_gemm_plain(cuda_dev, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = bf(torch.randn(M, K, generator=g))
    W = bf(torch.randn(N, K, generator=g) * K ** -0.5)
    bias = torch.randn(N, generator=g)
    ref = A.float() @ W.float().t() + bias
    out = ops.gemm_bf16(A.to(cuda_dev), W.to(cuda_dev), bias.to(cuda_dev), "f32")
    torch.cuda.synchronize()
    assert rel(out, ref) < 2e-3, rel(out, ref)
    out16 = ops.gemm_bf16(A.to(cuda_dev), W.to(cuda_dev), bias.to(cuda_dev), "bf16")
    assert rel(out16.float(), ref) < 6e-3


def test_gemm_epilogues(cuda_dev):
    g = torch.Generator().manual_seed(3)
    M, N, K = 394, 768, 768
    A, W = bf(torch.randn(M, K, generator=g)), bf(torch.randn(N, K, generator=g) * K ** -0.5)
    bias, res = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    acc = A.float() @ W.float().t() + bias
    Ad, Wd, bd = A.to(cuda_dev), W.to(cuda_dev), bias.to(cuda_dev)
    q = ops.gemm_bf16(Ad, Wd, bd, "bf16_qgelu")
    assert rel(q.float(), acc * torch.sigmoid(1.702 * acc)) < 6e-3
    x = res.to(cuda_dev).clone()
    ops.gemm_bf16(Ad, Wd, bd, "f32_resadd", out=x, res=x)               # in place on the residual stream
    assert rel(x, acc + res) < 2e-3
    nobias = ops.gemm_bf16(Ad, Wd, None, "f32")
    assert rel(nobias, acc - bias) < 2e-3
    # fp16 residual stream: in place on halves (TMA boxes of 32 x 32 halves), fp32 math, one rounding to fp16;
    # shapes that exercise the single-CTA (M < 256) and the CTA-pair kernels and a ragged last row tile
    for Mh in (394, 100, 1000):
        Ah = bf(torch.randn(Mh, K, generator=g))
        resh = torch.randn(Mh, N, generator=g).half()
        xh = resh.to(cuda_dev).clone()
        ops.gemm_bf16(Ah.to(cuda_dev), Wd, bd, "f16_resadd", out=xh, res=xh)
        wanth = Ah.float() @ W.float().t() + bias + resh.float()
        assert xh.dtype == torch.float16 and rel(xh.float(), wanth) < 2e-3
        assert (xh.float().cpu() - wanth).abs().max() < 2.0 ** -9 * wanth.abs().max() + 0.03      # within fp16 rounding + bf16 product noise
    with pytest.raises(Exception):
        ops.gemm_bf16(Ad, Wd, bd, "f16_resadd", out=x, res=x)           # fp32 tensors are rejected
    # patch epilogue: rows land at token 1 + m % G2 of image m // G2 and get the positional embedding
    G2, n_img, d = 49, 6, 256
    P = bf(torch.randn(n_img * G2, 128, generator=g))
    Wp = bf(torch.randn(d, 128, generator=g) * 0.1)
    pos = torch.randn(G2 + 1, d, generator=g)
    tok = torch.full((n_img * (G2 + 1), d), 7.0, device=cuda_dev)
    ops.gemm_bf16(P.to(cuda_dev), Wp.to(cuda_dev), None, "patch", out=tok, res=pos.to(cuda_dev), row_map=G2)
    want = (P.float() @ Wp.float().t()).view(n_img, G2, d) + pos[1:]
    got = tok.view(n_img, G2 + 1, d).cpu()
    assert rel(got[:, 1:], want) < 2e-3 and (got[:, 0] == 7.0).all()


def test_layernorm_folded_into_gemms(cuda_dev):
    """ec_gemm_bf16_stats writes the row statistics of the fp16 stream it produces; ec_gemm_ln reads the stream itself as A and
    applies rstd (acc - mean s_j) + c_j -- against LayerNorm + Linear in fp32 on the same fp16 rows."""
    g = torch.Generator().manual_seed(12)
    for M, d, N2 in ((394, 768, 2304), (1000, 768, 3072), (300, 1024, 1024), (70, 256, 384)):
        A = bf(torch.randn(M, d, generator=g))
        Wo = bf(torch.randn(d, d, generator=g) * d ** -0.5)
        bo = torch.randn(d, generator=g)
        x0 = (torch.randn(M, d, generator=g) * 2 + 0.5).half()
        gamma, beta = torch.rand(d, generator=g) + 0.5, torch.randn(d, generator=g) * 0.1
        W2 = torch.randn(N2, d, generator=g) * d ** -0.5
        b2 = torch.randn(N2, generator=g)
        parts = ops.gemm_stats_parts(d)
        x = x0.to(cuda_dev).clone()
        stats = torch.full((M, parts, 2), 7.0, device=cuda_dev)
        ops.gemm_bf16_stats(A.to(cuda_dev), Wo.to(cuda_dev), bo.to(cuda_dev), x, stats)
        want_x = A.float() @ Wo.float().t() + bo + x0.float()
        assert rel(x.float(), want_x) < 2e-3
        xs = x.float().cpu()                                     # statistics are those of the ROUNDED stream
        st = stats.cpu().sum(1)
        assert (st[:, 0] - xs.sum(1)).abs().max() < 1e-3 * xs.abs().sum(1).max()
        assert (st[:, 1] - (xs * xs).sum(1)).abs().max() < 1e-4 * (xs * xs).sum(1).max()
        # consumer
        Wg = (W2 * gamma[None, :]).half()
        colsum, cbias = Wg.float().sum(1), W2 @ beta + b2
        ref = torch.nn.functional.layer_norm(xs, (d,), gamma, beta, 1e-5) @ W2.t() + b2
        for epi, want in (("bf16", ref), ("bf16_qgelu", ref * torch.sigmoid(1.702 * ref))):
            got = ops.gemm_ln(x, Wg.to(cuda_dev), colsum.to(cuda_dev), cbias.to(cuda_dev), stats, parts, epi)
            assert got.dtype == torch.bfloat16 and rel(got.float(), want) < 8e-3, (M, d, epi, rel(got.float(), want))
        # centred rows of Wg: the mean term drops out, no column sums needed (what the encoder uses)
        Wc = W2 * gamma[None, :]
        Wc = (Wc - Wc.mean(1, keepdim=True)).half()
        gotc = ops.gemm_ln(x, Wc.to(cuda_dev), None, cbias.to(cuda_dev), stats, parts, "bf16")
        assert rel(gotc.float(), ref) < 8e-3, (M, d, rel(gotc.float(), ref))
        # statistics from the stand-alone kernel (rows no epilogue produced) give the same result
        st2 = torch.full((M, parts, 2), 3.0, device=cuda_dev)
        ops.row_stats_f16(x, st2, parts)
        assert (st2[:, 1:] == 0).all() and (st2[:, 0].cpu() - st).abs().max() < 1e-3 * st.abs().max()
        got2 = ops.gemm_ln(x, Wg.to(cuda_dev), colsum.to(cuda_dev), cbias.to(cuda_dev), st2, parts, "bf16")
        assert rel(got2.float(), ref) < 8e-3


def test_layernorm_folding_on_clip_like_statistics(cuda_dev):
    """Pretrained CLIP streams are not Gaussian: a few channels sit at 50-500 x the typical magnitude on every token and rows
    can carry a mean well away from zero.  The folded LayerNorm (var = E[x^2] - mean^2 on un-centred fp16 rows) is checked
    against the fp32 oracle on towers engineered to have (a) two outlier channels at +80 / -60 sigma, (b) a common offset of
    6 sigma on every channel, (c) an offset of 3000 sigma -- a regime where the folding loses most of its bits and the guard
    (VisionTransformer.validate_ln_fold) has to switch it off."""
    arch = "ViT-tiny/16"
    imgs = torch.randn(6, 3, 224, 224, generator=torch.Generator().manual_seed(10))
    for tag, edit, fold_must_hold in (("outlier_channels", lambda v: (v.ln_pre.bias.data.__setitem__(3, 80.0), v.ln_pre.bias.data.__setitem__(77, -60.0)), True),
                                      ("row_offset_6_sigma", lambda v: v.ln_pre.bias.data.add_(6.0), True),
                                      ("row_offset_3000_sigma", lambda v: v.ln_pre.bias.data.add_(3000.0), False),
                                      # (d) a stream beyond fp16's range: the guard also moves the residual stream to float32
                                      ("stream_beyond_fp16_range", lambda v: v.ln_pre.weight.data.mul_(4.0e4), None)):
        oracle = clip_oracle.build_clip(arch, seed=24)
        edit(oracle.visual)
        model = clip.CLIP(arch)
        model.load_state_dict(oracle.state_dict())
        model = model.to(cuda_dev).eval()
        vis = model.visual
        assert vis.fold_ln and vis.residual_dtype == torch.float16
        with torch.no_grad():
            ref = oracle.encode_image(imgs)
            patches = ops.im2col(imgs.to(cuda_dev), vis.patch_size, vis.k_patch, dtype=vis.operand_dtype)
            diff = vis.validate_ln_fold(patches, imgs.shape[0])
            got = model.encode_image(imgs.to(cuda_dev)).cpu()
        from parity_util import record_metric
        record_metric("ln_fold_robustness", case=tag, fold_vs_unfolded=diff, fold_kept=bool(vis.fold_ln), rel_l2=rel(got, ref))
        if fold_must_hold is None:
            assert vis.residual_dtype == torch.float32 and not vis.fold_ln, tag
            assert torch.isfinite(got).all() and rel(got, ref) < 2e-2, (tag, rel(got, ref))
        elif fold_must_hold:
            assert vis.fold_ln and diff < 1e-2, (tag, diff)
            assert rel(got, ref) < 2e-2, (tag, rel(got, ref))
        else:
            # the guard notices and falls back to the LayerNorm kernels; fp16 cannot hold a 3000-sigma offset exactly either
            # (2^-11 * 3000 = 1.5 sigma per element), so only the switch itself and finiteness are asserted
            assert not vis.fold_ln, (tag, diff)
            assert torch.isfinite(got).all()


def test_layernorm_and_helpers(cuda_dev):
    g = torch.Generator().manual_seed(4)
    for M, d in ((37, 768), (5, 1024), (9, 128), (3, 2048)):
        x = torch.randn(M, d, generator=g) * 3 + 1
        w, b = torch.randn(d, generator=g), torch.randn(d, generator=g)
        ref = torch.nn.functional.layer_norm(x, (d,), w, b, 1e-5)
        o32 = torch.empty(M, d, device=cuda_dev)
        o16 = torch.empty(M, d, dtype=torch.bfloat16, device=cuda_dev)
        ops.layernorm(x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev), M, d, out_bf16=o16, out_f32=o32)
        assert (o32.cpu() - ref).abs().max() < 2e-5 * ref.abs().max().clamp_min(1)
        assert torch.equal(o16, o32.to(torch.bfloat16))
        assert (ops.layernorm_f32(x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev)).cpu() - ref).abs().max() < 1e-4
    # fp16 residual stream: fp16 input (statistics in fp32), fp16 output (what ln_pre writes)
    for M, d in ((37, 768), (5, 1024), (3, 2048)):
        x = (torch.randn(M, d, generator=g) * 3 + 1).half()
        w, b = torch.randn(d, generator=g), torch.randn(d, generator=g)
        ref = torch.nn.functional.layer_norm(x.float(), (d,), w, b, 1e-5)
        o32 = torch.empty(M, d, device=cuda_dev)
        ob = torch.empty(M, d, dtype=torch.bfloat16, device=cuda_dev)
        oh = torch.empty(M, d, dtype=torch.float16, device=cuda_dev)
        ops.layernorm(x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev), M, d, out_bf16=ob, out_f32=o32, out_f16=oh)
        assert (o32.cpu() - ref).abs().max() < 2e-5 * ref.abs().max().clamp_min(1)
        assert torch.equal(ob, o32.to(torch.bfloat16)) and torch.equal(oh, o32.half())
        oh2 = torch.empty(M, d, dtype=torch.float16, device=cuda_dev)
        ops.layernorm(x.float().to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev), M, d, out_f16=oh2)       # fp32 in, fp16 out
        assert (oh2.float() - oh.float()).abs().max() <= 2.0 ** -9 * oh.float().abs().max()      # other summation order, fp32 input
    # strided rows (ln_post reads the class token of every image)
    x = torch.randn(4, 5, 128, generator=g)
    w, b = torch.ones(128), torch.zeros(128)
    o = torch.empty(4, 128, device=cuda_dev)
    ops.layernorm(x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev), 4, 128, row_stride=5 * 128, out_f32=o)
    assert (o.cpu() - torch.nn.functional.layer_norm(x[:, 0], (128,))).abs().max() < 1e-5
    s = torch.randn(1001, generator=g)
    assert torch.equal(ops.f32_to_bf16(s.to(cuda_dev)).cpu(), s.to(torch.bfloat16))
    # LoRA merge: bf16(W + up @ down), fp32 math (models/lora.py:138-149)
    W, up, down = torch.randn(96, 64, generator=g), torch.randn(96, 4, generator=g), torch.randn(4, 64, generator=g)
    m = ops.lora_merge(W.to(cuda_dev), up.to(cuda_dev), down.to(cuda_dev)).cpu()
    assert (m.float() - (W + up @ down)).abs().max() <= 2 ** -8 * (W + up @ down).abs().max()
    assert torch.equal(ops.lora_merge(W.to(cuda_dev), None, None).cpu(), W.to(torch.bfloat16))


# the last three cases give every persistent CTA several (image, head) units: the shared-memory ring of the two-slot kernel
# wraps (four resident units for L <= 128, two for L <= 256), the slots alternate, odd and even unit counts per CTA
@pytest.mark.parametrize("L,heads,n_img", [(50, 12, 3), (197, 12, 2), (257, 16, 2), (5, 2, 4), (64, 2, 1), (65, 2, 1),
                                           (50, 12, 131), (128, 4, 333), (197, 12, 57), (129, 2, 260),
                                           # 256 < L: rows >= 256 on the FMA pipe (at most four), else a third tile; many units per CTA
                                           (258, 2, 3), (260, 2, 2), (261, 2, 2), (257, 16, 40), (384, 2, 2)])
def test_attention(cuda_dev, L, heads, n_img):
    g = torch.Generator().manual_seed(L)
    d = heads * 64
    qkv = bf(torch.randn(n_img * L, 3 * d, generator=g))
    out = torch.empty(n_img * L, d, dtype=torch.bfloat16, device=cuda_dev)
    ops.attention(qkv.to(cuda_dev), out, n_img, L, heads)
    q, k, v = [t.view(n_img, L, heads, 64).transpose(1, 2) for t in qkv.float().chunk(3, -1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v).transpose(1, 2).reshape(n_img * L, d)
    assert rel(out.float(), ref) < 8e-3, rel(out.float(), ref)


@pytest.mark.parametrize("L,heads,n_img", [(50, 12, 131), (257, 16, 40), (197, 12, 57), (64, 2, 300)])
def test_attention_fp16_many_units_per_cta(cuda_dev, L, heads, n_img):
    """The fp16-operand launches (what the inference forward uses) with several units per persistent CTA: the eight-stage ring of
    the small-sequence kernel, the two-stage ring + tail-row warps of the L = 257 kernel, the two-slot kernel."""
    g = torch.Generator().manual_seed(1000 + L)
    d = heads * 64
    qkv = torch.randn(n_img * L, 3 * d, generator=g).half()
    out = torch.empty(n_img * L, d, dtype=torch.float16, device=cuda_dev)
    ops.attention(qkv.to(cuda_dev), out, n_img, L, heads)
    q, k, v = [t.view(n_img, L, heads, 64).transpose(1, 2) for t in qkv.float().chunk(3, -1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v).transpose(1, 2).reshape(n_img * L, d)
    got = out.float().cpu()
    assert rel(got, ref) < 2e-3, rel(got, ref)
    # the last query row and the first key of the L = 257 kernel take their own routes: check them on their own
    if L == 257:
        rows = torch.arange(n_img) * L + 256
        assert rel(got[rows], ref[rows]) < 2e-3, rel(got[rows], ref[rows])


@pytest.mark.parametrize("L,heads,n_img,dt", [(197, 2, 3, torch.float16), (197, 2, 3, torch.bfloat16), (50, 2, 5, torch.float16),
                                             (257, 2, 2, torch.float16), (77, 2, 4, torch.float16)])
def test_attention_peaked_rows_beyond_the_first_keys(cuda_dev, L, heads, n_img, dt):
    """Trained towers have attention sinks: one key can sit 20+ nats above every other, anywhere in the row.  The tcgen05 kernels
    take the EXACT row maximum as the softmax reference with fp16 operands (EC_ATTN_EXACT; bf16 keeps the single pass, range
    2^127), so the 16-bit P never overflows: rows whose dominant key lies beyond the first 32 keys (the single-pass reference
    window) and is 24 nats above the rest must come out as that key's value row, finite."""
    g = torch.Generator().manual_seed(L + heads)
    d = heads * 64
    q = torch.ones(n_img, L, heads, 64) + 0.05 * torch.randn(n_img, L, heads, 64, generator=g)
    k = 0.05 * torch.randn(n_img, L, heads, 64, generator=g)
    v = torch.randn(n_img, L, heads, 64, generator=g)
    hot = L - 7                                       # far beyond key 31; with a causal mask it would be visible to the last rows only
    k[:, hot] = 3.0                                   # q . k = 192 -> 24 nats after the 1/8 scale
    k[:, :, 1] = 0.05 * torch.randn(n_img, L, 64, generator=g)      # head 1 stays flat: ordinary rows in the same launch
    qkv = torch.cat([q.reshape(n_img * L, d), k.reshape(n_img * L, d), v.reshape(n_img * L, d)], 1).to(dt)
    out = torch.empty(n_img * L, d, dtype=dt, device=cuda_dev)
    ops.attention(qkv.to(cuda_dev), out, n_img, L, heads)
    qf, kf, vf = [t.view(n_img, L, heads, 64).transpose(1, 2) for t in qkv.double().chunk(3, -1)]
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, -1) @ vf).transpose(1, 2).reshape(n_img * L, d).float()
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    assert rel(got, ref) < 8e-3, rel(got, ref)


@pytest.mark.parametrize("arch,n", [("ViT-tiny/32", 5), ("ViT-tiny/16", 3), ("ViT-B/32", 4)])
def test_encoder_vs_oracle(cuda_dev, arch, n):
    oracle = clip_oracle.build_clip(arch, seed=21)
    model = clip.CLIP(arch)
    model.load_state_dict(oracle.state_dict())
    model = model.to(cuda_dev).eval()
    g = torch.Generator().manual_seed(8)
    imgs = torch.randn(n, 3, 224, 224, generator=g)
    with torch.no_grad():
        ref = oracle.encode_image(imgs)
        got = model.encode_image(imgs.to(cuda_dev))
        got_bf = model.encode_image(imgs.to(cuda_dev).to(torch.bfloat16))
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert rel(got, ref) < 2e-2, rel(got, ref)
    from parity_util import record_metric, rel_l2_centered
    record_metric("encoder_vs_oracle", arch=arch, rel_l2=rel(got, ref), rel_l2_centered=rel_l2_centered(got, ref))
    assert rel_l2_centered(got, ref) < 5e-2, rel_l2_centered(got, ref)      # error against the input-dependent part only
    assert rel(got_bf, ref) < 2e-2
    # weight updates are picked up (version counters) -- optimizer steps / load_state_dict must not see stale packs
    with torch.no_grad():
        model.visual.proj.mul_(2.0)
        got2 = model.encode_image(imgs.to(cuda_dev))
    assert rel(got2, 2 * ref) < 2e-2


@pytest.mark.parametrize("arch", ["ViT-B/16", "ViT-L/14"])
def test_encoder_large_archs_vs_oracle(cuda_dev, arch):
    oracle = clip_oracle.build_clip(arch, seed=22)
    model = clip.CLIP(arch)
    model.load_state_dict(oracle.state_dict())
    model = model.to(cuda_dev).eval()
    imgs = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        ref = oracle.encode_image(imgs)
        got = model.encode_image(imgs.to(cuda_dev))
    assert rel(got, ref) < 2e-2, rel(got, ref)


def test_split_residual_stream_update(cuda_dev):
    """ec_gemm_bf16_stats2: the stream as a (hi, lo) pair of fp16 planes.  After an update hi + lo equals the fp32 result to
    ~2^-21 relative (hi alone only to 2^-11), lo stays below half an ulp of hi, and the statistics are those of the sum; a chain of
    updates drifts far less than the plain fp16 stream."""
    g = torch.Generator().manual_seed(31)
    for M, d, K in ((394, 768, 768), (1000, 768, 3072), (70, 256, 128)):
        A = torch.randn(M, K, generator=g).half()
        Wo = (torch.randn(d, K, generator=g) * K ** -0.5).half()
        bo = torch.randn(d, generator=g)
        x_true = (torch.randn(M, d, generator=g) * 2 + 0.5).double()
        hi0 = x_true.float().half()
        lo0 = (x_true.float() - hi0.float()).half()
        x2 = torch.stack([hi0, lo0]).to(cuda_dev).contiguous()
        x1 = hi0.to(cuda_dev).clone()
        parts = ops.gemm_stats_parts(d)
        stats = torch.full((M, parts, 2), 7.0, device=cuda_dev)
        stats1 = torch.empty_like(stats)
        upd = A.double() @ Wo.double().t() + bo.double()
        want = hi0.double() + lo0.double()
        for it in range(6):
            ops.gemm_bf16_stats2(A.to(cuda_dev), Wo.to(cuda_dev), bo.to(cuda_dev), x2, stats)
            ops.gemm_bf16_stats(A.to(cuda_dev), Wo.to(cuda_dev), bo.to(cuda_dev), x1, stats1)
            want = want + upd
        hi, lo = x2[0].cpu(), x2[1].cpu()
        got = hi.double() + lo.double()
        err2 = float((got - want).norm() / want.norm())
        err1 = float((x1.cpu().double() - want).norm() / want.norm())
        assert err2 < 5e-6 and err1 > 20 * err2, (M, d, K, err1, err2)       # fp32-accumulator noise vs six fp16 roundings
        assert (lo.float().abs() <= hi.float().abs() * 2.0 ** -10 + 1e-7).all()    # lo is a rounding residue: at most half an ulp of hi
        st = stats.cpu().sum(1).double()
        assert (st[:, 0] - got.sum(1)).abs().max() < 1e-3 * got.abs().sum(1).max()
        assert (st[:, 1] - (got * got).sum(1)).abs().max() < 1e-4 * (got * got).sum(1).max()


def test_split_residual_stream_encoder_vs_oracle(cuda_dev):
    """EC_RESIDUAL=fp16x2 / visual.residual_split: the encoder with the (hi, lo) stream is closer to the fp32 oracle than with
    the plain fp16 stream, on the part of the features that differs between images (centred error)."""
    from parity_util import record_metric, rel_l2_centered
    arch = "ViT-B/16"
    oracle = clip_oracle.build_clip(arch, seed=29)
    model = clip.CLIP(arch)
    model.load_state_dict(oracle.state_dict())
    model = model.to(cuda_dev).eval()
    imgs = torch.randn(6, 3, 224, 224, generator=torch.Generator().manual_seed(11))
    with torch.no_grad():
        ref = oracle.encode_image(imgs)
        got16 = model.encode_image(imgs.to(cuda_dev)).cpu()
        model.visual.residual_split = True
        got2 = model.encode_image(imgs.to(cuda_dev)).cpu()
        model.visual.residual_split = False
    e16, e2 = rel_l2_centered(got16, ref), rel_l2_centered(got2, ref)
    record_metric("encoder_split_stream", arch=arch, centred_fp16=e16, centred_fp16x2=e2, plain_fp16x2=rel(got2, ref))
    # with bf16 operands (EC_OPERANDS=bf16) the operand roundings dominate and the stream's format hardly matters
    gain = 0.75 if model.visual.operand_dtype == torch.float16 else 1.05
    assert rel(got2, ref) < 2e-2 and e2 < gain * e16, (e16, e2)


def test_residual_stream_fp16_and_fp32_vs_oracle(cuda_dev):
    """The inference forward keeps the residual stream in fp16 (the reference's CUDA precision) by default and in fp32 on
    request; both stay within the encoder tolerance of the fp32 oracle and close to each other."""
    arch = "ViT-B/16"
    oracle = clip_oracle.build_clip(arch, seed=23)
    model = clip.CLIP(arch)
    model.load_state_dict(oracle.state_dict())
    model = model.to(cuda_dev).eval()
    imgs = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(10))
    with torch.no_grad():
        ref = oracle.encode_image(imgs)
        assert model.visual.residual_dtype == torch.float16
        got16 = model.encode_image(imgs.to(cuda_dev)).cpu()
        model.visual.residual_dtype = torch.float32
        got32 = model.encode_image(imgs.to(cuda_dev)).cpu()
    assert rel(got16, ref) < 2e-2 and rel(got32, ref) < 2e-2, (rel(got16, ref), rel(got32, ref))
    assert rel(got16, got32) < 1e-2 and not torch.equal(got16, got32)


@pytest.mark.parametrize("L,heads,n_seq", [(77, 8, 5), (16, 1, 3), (197, 12, 2), (257, 16, 1), (77, 8, 301), (197, 12, 40)])
def test_causal_attention(cuda_dev, L, heads, n_seq):
    """Text-tower attention: key j is visible to query i iff j <= i (both tensor-memory kernels)."""
    g = torch.Generator().manual_seed(L + 1)
    d = heads * 64
    qkv = bf(torch.randn(n_seq * L, 3 * d, generator=g))
    out = torch.empty(n_seq * L, d, dtype=torch.bfloat16, device=cuda_dev)
    ops.attention(qkv.to(cuda_dev), out, n_seq, L, heads, causal=True)
    q, k, v = [t.view(n_seq, L, heads, 64).transpose(1, 2) for t in qkv.float().chunk(3, -1)]
    s = q @ k.transpose(-1, -2) / 8.0 + torch.full((L, L), float("-inf")).triu_(1)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n_seq * L, d)
    assert rel(out.float(), ref) < 8e-3, rel(out.float(), ref)


@pytest.mark.parametrize("arch,n", [("ViT-tiny/32", 7), ("ViT-B/16", 5)])
def test_text_tower_vs_oracle(cuda_dev, arch, n):
    """encode_text from token ids (models/clip_cls.py:84) against the fp32 oracle."""
    oracle = clip_oracle.build_clip(arch, seed=23, text=True)
    model = clip.CLIP(arch, text=True)
    model.load_state_dict(oracle.state_dict())
    model = model.to(cuda_dev).eval()
    tok = clip_oracle.synth_tokens(n, oracle.context_length, oracle.vocab_size, 4)
    with torch.no_grad():
        ref = oracle.encode_text(tok)
        got = model.encode_text(tok.to(cuda_dev))
    assert got.shape == ref.shape
    assert rel(got, ref) < 2e-2, rel(got, ref)
    with pytest.raises(NotImplementedError):
        clip.CLIP(arch).encode_text(tok)
