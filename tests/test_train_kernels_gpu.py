"""GPU parity of the fine-tune step's backward kernels (through the C ABI) against torch autograd in fp32 on the CPU.

The reference obtains every one of these gradients from autograd (models/clip_cls_ft.py:214-269 trains through
model.visual), so autograd on the same bf16-rounded operands is the oracle.  Tolerances: fp32 kernels rel-L2 <= 1e-5;
bf16-operand kernels (attention backward, QuickGELU) rel-L2 <= 1.5e-2 (P and dS are rounded to bf16 for the second
matmul, outputs are bf16).
"""
import pytest
import torch
import torch.nn.functional as F

from eventclip_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,d,stride", [(37, 768, None), (5, 1024, 3 * 1024), (64, 256, None), (33, 200, None), (9, 1280, None)])
def test_layernorm_bwd(cuda_dev, M, d, stride):
    g = torch.Generator().manual_seed(M)
    rows = stride // d if stride else 1
    xfull = (torch.randn(M * rows, d, generator=g) * 2 + 0.5)
    x = xfull.view(M, rows * d)[:, :d].clone().requires_grad_(True)
    gamma, beta = torch.randn(d, generator=g), torch.randn(d, generator=g)
    dy, acc = torch.randn(M, d, generator=g), torch.randn(M, d, generator=g)
    F.layer_norm(x, (d,), gamma, beta).backward(dy)
    dx = ops.layernorm_bwd(xfull.to(cuda_dev), dy.to(cuda_dev), gamma.to(cuda_dev), M, d, x_stride=stride)
    assert rel(dx, x.grad) < 1e-5
    dxb = torch.empty((M, d), dtype=torch.bfloat16, device=cuda_dev)
    dx2 = ops.layernorm_bwd(xfull.to(cuda_dev), dy.to(cuda_dev), gamma.to(cuda_dev), M, d, acc=acc.to(cuda_dev), x_stride=stride,
                            dx_bf16=dxb)
    assert rel(dx2, x.grad + acc) < 1e-5
    assert torch.equal(dxb, dx2.to(torch.bfloat16))


def test_quickgelu_fwd_bwd(cuda_dev):
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(1000, 96, generator=g) * 3).to(torch.bfloat16)
    dh = torch.randn(1000, 96, generator=g).to(torch.bfloat16)
    af = a.float().requires_grad_(True)
    h = af * torch.sigmoid(1.702 * af)
    h.backward(dh.float())
    out = ops.quickgelu(a.to(cuda_dev))
    assert rel(out.float(), h.detach()) < 4e-3
    da = ops.quickgelu_bwd(a.to(cuda_dev), dh.to(cuda_dev))
    assert rel(da.float(), af.grad) < 4e-3


@pytest.mark.parametrize("R,Cc", [(197 * 3, 768), (50, 64), (1001, 40)])
def test_transpose_bf16(cuda_dev, R, Cc):
    x = torch.randn(R, Cc).to(torch.bfloat16)
    out = ops.transpose_bf16(x.to(cuda_dev)).cpu()
    Rp = (R + 7) // 8 * 8
    assert out.shape == (Cc, Rp)
    assert torch.equal(out[:, :R], x.t())
    assert (out[:, R:] == 0).all()


@pytest.mark.parametrize("n_img,Ltok,heads", [(2, 197, 12), (3, 50, 4), (1, 257, 16), (2, 64, 2), (1, 77, 8), (40, 197, 12),
                                              (1, 128, 1), (2, 129, 3), (1, 256, 2)])
def test_attention_bwd(cuda_dev, n_img, Ltok, heads):
    g = torch.Generator().manual_seed(Ltok)
    d = heads * 64
    qkv = (torch.randn(n_img * Ltok, 3 * d, generator=g) * 1.5).to(torch.bfloat16)
    d_o = torch.randn(n_img * Ltok, d, generator=g).to(torch.bfloat16)
    x = qkv.float().requires_grad_(True)
    q, k, v = [t.reshape(n_img, Ltok, heads, 64).transpose(1, 2) for t in x.split(d, dim=1)]
    p = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(n_img * Ltok, d)
    o.backward(d_o.float())
    dev_qkv = qkv.to(cuda_dev)
    o_dev = torch.empty((n_img * Ltok, d), dtype=torch.bfloat16, device=cuda_dev)
    ops.attention(dev_qkv, o_dev, n_img, Ltok, heads)
    dqkv = ops.attention_bwd(dev_qkv, o_dev, d_o.to(cuda_dev), n_img, Ltok, heads)        # mma.sync kernel (no lse given)
    torch.cuda.synchronize()
    for j, name in enumerate("qkv"):
        r = rel(dqkv[:, j * d:(j + 1) * d].float(), x.grad[:, j * d:(j + 1) * d])
        assert r < 1.5e-2, (name, r)
    if Ltok <= 256:
        # tcgen05 kernel: forward with log-sum-exp export, backward with S / dP / dQ / dK / dV in tensor memory
        o2 = torch.empty_like(o_dev)
        lse = ops.attention_fwd_lse(dev_qkv, o2, n_img, Ltok, heads)
        assert torch.equal(o2, o_dev)
        s = (q @ k.transpose(-1, -2) * 0.125).detach()
        ref_lse = torch.logsumexp(s, dim=-1) * 1.4426950408889634                           # [n_img, heads, L], log2 domain
        assert (lse.cpu() - ref_lse).abs().max().item() < 2e-2
        dq2 = torch.full_like(dqkv, float("nan"))
        ops.attention_bwd(dev_qkv, o_dev, d_o.to(cuda_dev), n_img, Ltok, heads, out=dq2, lse=lse)
        torch.cuda.synchronize()
        assert not torch.isnan(dq2.float()).any()
        for j, name in enumerate("qkv"):
            r = rel(dq2[:, j * d:(j + 1) * d].float(), x.grad[:, j * d:(j + 1) * d])
            assert r < 1.5e-2, ("tc", name, r)


def test_adam_matches_torch(cuda_dev):
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(10007, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-3)
    p = p0.to(cuda_dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn(10007, generator=g)
        ref.grad = grad.clone()
        opt.step()
        ops.adam(p, grad.to(cuda_dev), m, v, lr=2e-3, step=step)
    assert rel(p, ref.detach()) < 1e-6


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_mm_f32(cuda_dev, ta, tb):
    g = torch.Generator().manual_seed(11)
    M, N, K = 45, 70, 133
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    ref = (A.t() if ta else A) @ (B.t() if tb else B)
    out = ops.mm_f32(A.to(cuda_dev), B.to(cuda_dev), ta, tb)
    assert rel(out, ref) < 1e-5
    out2 = ops.mm_f32(A.to(cuda_dev), B.to(cuda_dev), ta, tb, out=out.clone(), accumulate=True)
    assert rel(out2, 2 * ref) < 1e-5


def test_l2norm_rows_bwd(cuda_dev):
    g = torch.Generator().manual_seed(13)
    x = torch.randn(20, 512, generator=g).requires_grad_(True)
    mask = torch.rand(20, generator=g) > 0.3
    dy = torch.randn(20, 512, generator=g)
    (F.normalize(x, p=2, dim=-1) * mask.float()[:, None]).backward(dy)
    dx = ops.l2norm_rows_bwd(x.detach().to(cuda_dev), dy.to(cuda_dev), mask.to(torch.uint8).to(cuda_dev))
    assert rel(dx, x.grad) < 1e-5


@pytest.mark.parametrize("agg", ["sum", "mean"])
def test_ce_loss_bwd(cuda_dev, agg):
    g = torch.Generator().manual_seed(17)
    B, T, K = 9, 3, 101
    valid = torch.rand(B, T, generator=g) > 0.3
    valid[:, 0] = True
    full = (torch.randn(B, T, K, generator=g) * 3 * valid[..., None].float()).requires_grad_(True)
    labels = torch.randint(0, K, (B,), generator=g)
    logits = full.sum(1)
    if agg == "mean":
        logits = logits / valid.float().sum(1, keepdim=True)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    lb, lm, dfull = ops.ce_loss_bwd(full.detach().to(cuda_dev), valid.to(torch.uint8).to(cuda_dev),
                                    labels.to(torch.int32).to(cuda_dev), agg)
    assert abs(lm.item() - loss.item()) < 1e-5 * max(1.0, abs(loss.item()))
    assert rel(lb, F.cross_entropy(logits, labels, reduction="none").detach()) < 1e-5
    # padded views: the reference's rows are features * valid mask, i.e. constants (clip_cls.py:330-333) -> zero gradient here
    assert rel(dfull.cpu()[valid], full.grad[valid]) < 1e-5
    assert (dfull.cpu()[~valid] == 0).all()


def test_probs_loss_bwd(cuda_dev):
    """use_probs_loss (clip_cls_ft.py:265-267): nll of log(mean-over-valid-views softmax + 1e-6), restated with torch and
    differentiated by autograd; padded views (zero logits) get a zero gradient."""
    g = torch.Generator().manual_seed(19)
    for B, T, K in ((9, 3, 101), (4, 10, 1000), (5, 1, 2)):
        valid = torch.rand(B, T, generator=g) > 0.3
        valid[:, 0] = True
        full = (torch.randn(B, T, K, generator=g) * 3 * valid[..., None].float()).requires_grad_(True)
        labels = torch.randint(0, K, (B,), generator=g)
        vm = valid.float()
        probs = (full.softmax(-1) * vm[..., None]).sum(1) / vm.sum(1, keepdim=True)        # clip_cls.py:123-129
        loss = F.nll_loss((probs + 1e-6).log(), labels)
        loss.backward()
        lb, lm, dfull = ops.probs_loss_bwd(full.detach().to(cuda_dev), valid.to(torch.uint8).to(cuda_dev),
                                           labels.to(torch.int32).to(cuda_dev))
        assert abs(lm.item() - loss.item()) < 1e-5 * max(1.0, abs(loss.item()))
        assert rel(lb, F.nll_loss((probs + 1e-6).log(), labels, reduction="none").detach()) < 1e-5
        assert rel(dfull, full.grad) < 1e-4
        assert (dfull.cpu()[~valid] == 0).all()


@pytest.mark.parametrize("rows,d,r,skip", [(128, 128, 4, None), (768, 768, 16, 1), (96, 200, 20, None)])
def test_lora_grad(cuda_dev, rows, d, r, skip):
    """d_up = dW . down^T, d_down = up^T . dW for three stacked matrices (q | k | v), optionally without LoRA on k."""
    g = torch.Generator().manual_seed(rows + r)
    dW = torch.randn(3 * rows, d, generator=g)
    fac = [(torch.randn(rows, r, generator=g), torch.randn(r, d, generator=g)) for _ in range(3)]
    dev_fac = [None if z == skip else (u.to(cuda_dev), dn.to(cuda_dev)) for z, (u, dn) in enumerate(fac)]
    outs = ops.lora_grad(dW.to(cuda_dev), rows, dev_fac)
    for z, (u, dn) in enumerate(fac):
        if z == skip:
            assert outs[z] is None
            continue
        w = dW[z * rows:(z + 1) * rows]
        assert rel(outs[z][0], w @ dn.t()) < 1e-5 and rel(outs[z][1], u.t() @ w) < 1e-5


@pytest.mark.parametrize("M,N,K,splits", [(768, 768, 12608, None), (2304, 768, 3 * 197 + 5, None), (384, 128, 304, 4),
                                          (128, 128, 1000, 16), (768, 768, 12608, 1)])
def test_gemm_splitk(cuda_dev, M, N, K, splits):
    """Weight-gradient shapes: split-K partial tiles + fixed-order reduction equal the single-pass GEMM's result."""
    g = torch.Generator().manual_seed(M + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp).to(torch.bfloat16)
    W = torch.zeros(N, Kp).to(torch.bfloat16)
    A[:, :K] = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W[:, :K] = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16)
    ref = A.float() @ W.float().t()
    out = ops.gemm_bf16_splitk(A.to(cuda_dev), W.to(cuda_dev), splits=splits)
    assert rel(out, ref) < 2e-3
    again = ops.gemm_bf16_splitk(A.to(cuda_dev), W.to(cuda_dev), splits=splits)
    assert torch.equal(out, again)            # deterministic


@pytest.mark.parametrize("M,N,dtype", [(12608, 768, torch.bfloat16), (197, 2304, torch.float32), (6, 50 * 128, torch.float32),
                                       (1000, 40, torch.bfloat16)])
def test_colsum(cuda_dev, M, N, dtype):
    a = torch.randn(M, N).to(dtype)
    out = ops.colsum(a.to(cuda_dev))
    assert rel(out, a.double().sum(0)) < 1e-5
    assert torch.equal(out, ops.colsum(a.to(cuda_dev)))           # deterministic


@pytest.mark.parametrize("M,d,stride", [(1000, 768, None), (37, 128, None), (5, 1024, 3 * 1024)])
def test_layernorm_param_grad(cuda_dev, M, d, stride):
    g = torch.Generator().manual_seed(M + d)
    rows = stride // d if stride else 1
    xfull = torch.randn(M * rows, d, generator=g) * 2 + 0.5
    x = xfull.view(M, rows * d)[:, :d].clone()
    gamma, beta = torch.randn(d, generator=g).requires_grad_(True), torch.randn(d, generator=g).requires_grad_(True)
    dy = torch.randn(M, d, generator=g)
    F.layer_norm(x, (d,), gamma, beta).backward(dy)
    dg, db = ops.layernorm_param_grad(xfull.to(cuda_dev), dy.to(cuda_dev), M, d, x_stride=stride)
    assert rel(dg, gamma.grad) < 1e-5 and rel(db, beta.grad) < 1e-5


def test_patch_rows_bf16(cuda_dev):
    n_img, G2, d = 3, 49, 128
    x = torch.randn(n_img * (G2 + 1), d)
    out = ops.patch_rows_bf16(x.to(cuda_dev), n_img, G2, d).cpu()
    ref = x.view(n_img, G2 + 1, d)[:, 1:].reshape(n_img * G2, d).to(torch.bfloat16)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("K,M,N,splits", [(12608, 768, 768, None), (3 * 197, 2304, 768, None), (300, 384, 128, 1), (1000, 128, 128, 4),
                                          (64, 256, 512, 1), (12608, 3072, 768, None), (777, 768, 3072, 2), (500, 768, 592, None)])
def test_gemm_tn_token_major(cuda_dev, K, M, N, splits):
    """out = A^T B straight from token-major activations (MN-major UMMA operands): equals the transposed-copy route."""
    g = torch.Generator().manual_seed(K + M)
    A = torch.randn(K, M, generator=g).to(torch.bfloat16)
    B = (torch.randn(K, N, generator=g) * K ** -0.5).to(torch.bfloat16)
    ref = A.float().t() @ B.float()
    out = ops.gemm_bf16_tn(A.to(cuda_dev), B.to(cuda_dev), splits=splits)
    torch.cuda.synchronize()
    assert rel(out, ref) < 2e-3, rel(out, ref)
    # strided views (a column block of a wider matrix, e.g. the q / k / v thirds of d_qkv)
    if M % 3 == 0 and (M // 3) % 8 == 0:
        Ad = A.to(cuda_dev)
        part = ops.gemm_bf16_tn(Ad[:, M // 3:2 * M // 3], B.to(cuda_dev), splits=splits)
        assert rel(part, ref[M // 3:2 * M // 3]) < 2e-3
