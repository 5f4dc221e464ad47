"""Kernel-level parity (GPU): every C-ABI entry point against a plain PyTorch fp32 reference of the same op
or against the CPU oracle.  All calls go through libspe_b200.so (ctypes)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from spe_b200 import ops
    return ops


def dev():
    return torch.device("cuda:0")


def bf(x):
    return x.to(torch.bfloat16)


def rel_err(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K_", [(128, 128, 64), (256, 384, 384), (300, 200, 136), (77, 48, 48), (1600, 1152, 384)])
def test_gemm_majors_and_tails(K, a_major, b_major, M, N, K_):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K_ + a_major * 2 + b_major)
    A = torch.randn(M, K_, generator=g)
    B = torch.randn(N, K_, generator=g)
    A16, B16 = bf(A).to(dev()), bf(B).to(dev())
    ref = A16.float() @ B16.float().t()
    pad = lambda n: (n + 7) // 8 * 8
    if a_major == 0:
        a_st = torch.zeros(M, pad(K_), dtype=torch.bfloat16, device=dev()); a_st[:, :K_] = A16; lda = pad(K_)
    else:
        a_st = torch.zeros(K_, pad(M), dtype=torch.bfloat16, device=dev()); a_st[:, :M] = A16.t(); lda = pad(M)
    if b_major == 0:
        b_st = torch.zeros(N, pad(K_), dtype=torch.bfloat16, device=dev()); b_st[:, :K_] = B16; ldb = pad(K_)
    else:
        b_st = torch.zeros(K_, pad(N), dtype=torch.bfloat16, device=dev()); b_st[:, :N] = B16.t(); ldb = pad(N)
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev())
    K.gemm(a_st, b_st, out, M, N, K_, a_major=a_major, lda=lda, b_major=b_major, ldb=ldb, ldc=N)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)


@pytest.mark.parametrize("M,N,K_,mn", [(384, 384, 2400, 1), (304, 200, 136, 1), (1536, 384, 12800, 1), (88, 384, 640, 1), (8, 384, 2400, 1),
                                       (300, 130, 64, 0), (77, 81, 1600, 0), (4, 384, 2400, 0)])
def test_gemm_accumulate_in_place(K, M, N, K_, mn):
    """residual aliasing C = C += A B (wgrad into .grad): TMA reduce-add path, split-K path and the direct (unaligned pitch) path."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K_)
    A = bf(torch.randn(M, K_, generator=g)).to(dev())
    B = bf(torch.randn(N, K_, generator=g)).to(dev())
    c0 = torch.randn(M, N, generator=g).to(dev())
    prod = A.float() @ B.float().t()
    out = c0.clone()
    if mn:      # wgrad layout: A = dy [K_, M], B = x [K_, N], both MN-major
        a_st, b_st, kw = A.t().contiguous(), B.t().contiguous(), dict(a_major=K.MAJOR_MN, lda=M, b_major=K.MAJOR_MN, ldb=N)
    else:
        a_st, b_st, kw = A, B, dict(lda=K_, ldb=K_)
    for reps in (1, 2):
        K.gemm(a_st, b_st, out, M, N, K_, ldc=N, residual=out, ldr=N, **kw)
        torch.cuda.synchronize()
        want = c0 + reps * prod
        assert rel_err(out, want) < 3e-5, (reps, rel_err(out, want))


def test_gemm_cta_pair_mode_subprocess():
    """cta_group::2 (256 x BN CTA-pair tiles) is opt-in (SPE_GEMM_CG2, read once per process): run the GEMM tests under it."""
    import os, subprocess, sys
    env = dict(os.environ, SPE_GEMM_CG2="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_kernels_gpu.py"), "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "gemm_majors_and_tails or gemm_accumulate or gemm_epilogue or linear_fn or ffn_fn"], env=env, cwd=root,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]


def test_gemm_epilogue_features(K):
    g = torch.Generator().manual_seed(3)
    M, N, Kd = 300, 384, 192
    x = bf(torch.randn(M, Kd, generator=g)).to(dev())
    w = bf(torch.randn(N, Kd, generator=g) / math.sqrt(Kd)).to(dev())
    bias = torch.randn(N, generator=g).to(dev())
    gamma = torch.rand(N, generator=g).to(dev())
    res = torch.randn(M, N, generator=g).to(dev())
    pre = x.float() @ w.float().t() * 0.5 + bias
    # gelu + aux_out (pre-activation) + bf16 out
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev())
    aux = torch.empty(M, N, dtype=torch.bfloat16, device=dev())
    K.gemm(x, w, out, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, alpha=0.5, bias=bias, act=K.ACT_GELU, aux_out=aux, ld_aux=N)
    assert rel_err(out, F.gelu(pre)) < 1e-2
    assert rel_err(aux, pre) < 1e-2
    # layerscale + residual, fp32 out
    out32 = torch.empty(M, N, dtype=torch.float32, device=dev())
    K.gemm(x, w, out32, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, alpha=0.5, bias=bias, gamma=gamma, residual=res, ldr=N, aux_out=aux, ld_aux=N)
    assert rel_err(out32, res + gamma * pre) < 1e-5
    # relu, then relu-grad / gelu-grad epilogues
    K.gemm(x, w, out32, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, alpha=0.5, bias=bias, act=K.ACT_RELU)
    assert rel_err(out32, F.relu(pre)) < 1e-5
    auxin = bf(torch.randn(M, N, generator=g)).to(dev())
    K.gemm(x, w, out32, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, alpha=0.5, act=K.ACT_RELU_GRAD, aux_in=auxin, ld_aux=N)
    assert rel_err(out32, (pre - bias) * (auxin.float() > 0)) < 1e-5
    K.gemm(x, w, out32, M, N, Kd, lda=Kd, ldb=Kd, ldc=N, alpha=0.5, act=K.ACT_GELU_GRAD, aux_in=auxin, ld_aux=N)
    a = auxin.float().requires_grad_(True)
    F.gelu(a).sum().backward()
    assert rel_err(out32, (pre - bias) * a.grad) < 1e-4
    # scalar epilogue path: N = 81, fp32 out with odd ld
    w81 = bf(torch.randn(81, Kd, generator=g) / math.sqrt(Kd)).to(dev())
    b81 = torch.randn(81, generator=g).to(dev())
    o81 = torch.empty(M, 81, dtype=torch.float32, device=dev())
    K.gemm(x, w81, o81, M, 81, Kd, lda=Kd, ldb=Kd, ldc=81, bias=b81)
    assert rel_err(o81, x.float() @ w81.float().t() + b81) < 1e-5
    # head-interleave remap
    o2 = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=dev())
    K.gemm(x, w, o2, M, N, Kd, lda=Kd, ldb=Kd, ldc=2 * N, split=48, split_stride=96)
    ref = (x.float() @ w.float().t()).view(M, N // 48, 48)
    assert rel_err(o2.view(M, N // 48, 96)[:, :, :48], ref) < 1e-2
    assert float(o2.view(M, N // 48, 96)[:, :, 48:].abs().max()) == 0.0


def test_gemm_batched_attention_shapes(K):
    """QK^T / PV / their backward layouts: heads packed in the feature dim, d=48 (K tail zero-filled by TMA)."""
    g = torch.Generator().manual_seed(5)
    B, H, Lq, Lk, d = 2, 4, 150, 200, 48
    q = bf(torch.randn(B, Lq, H * d, generator=g)).to(dev())
    k = bf(torch.randn(B, Lk, H * d, generator=g)).to(dev())
    v = bf(torch.randn(B, Lk, H * d, generator=g)).to(dev())
    ld = (Lk + 7) // 8 * 8
    S = torch.zeros(B, H, Lq, ld, dtype=torch.float32, device=dev())
    K._qk_logits(q, k, H, 0.25, S, ld)
    qh = q.float().view(B, Lq, H, d).transpose(1, 2)
    kh = k.float().view(B, Lk, H, d).transpose(1, 2)
    vh = v.float().view(B, Lk, H, d).transpose(1, 2)
    Sref = 0.25 * qh @ kh.transpose(-1, -2)
    assert rel_err(S[..., :Lk], Sref) < 1e-5
    # accumulate a second QK^T into S (conditional cross-attention)
    K._qk_logits(q, k, H, 0.25, S, ld, q2=q, k2=k)
    assert rel_err(S[..., :Lk], 2 * Sref) < 1e-5
    P = torch.zeros(B, H, Lq, ld, dtype=torch.bfloat16, device=dev())
    P[..., :Lk] = bf(torch.softmax(Sref, -1))
    out = torch.empty(B, Lq, H * d, dtype=torch.bfloat16, device=dev())
    K._pv(P, v, H, out, Lq, Lk, ld)
    Oref = (P[..., :Lk].float() @ vh).transpose(1, 2).reshape(B, Lq, H * d)
    assert rel_err(out, Oref) < 1e-2
    dO = bf(torch.randn(B, Lq, H * d, generator=g)).to(dev())
    dP, dV = K._attn_bwd_common(dO, P, v, H, Lq, Lk, ld)
    doh = dO.float().view(B, Lq, H, d).transpose(1, 2)
    assert rel_err(dP[..., :Lk], doh @ vh.transpose(-1, -2)) < 1e-2
    assert rel_err(dV, (P[..., :Lk].float().transpose(-1, -2) @ doh).transpose(1, 2).reshape(B, Lk, H * d)) < 1e-2
    dq, dk = K._dq_dk(P, q, k, H, 0.5, Lq, Lk, ld)
    assert rel_err(dq, (0.5 * P[..., :Lk].float() @ kh).transpose(1, 2).reshape(B, Lq, H * d)) < 1e-2
    assert rel_err(dk, (0.5 * P[..., :Lk].float().transpose(-1, -2) @ qh).transpose(1, 2).reshape(B, Lk, H * d)) < 1e-2


# ------------------------------------------------------------------------------------------------
# autograd Functions vs torch fp32
# ------------------------------------------------------------------------------------------------
def _grads(loss, *ts):
    return torch.autograd.grad(loss, ts, allow_unused=True)


@pytest.mark.parametrize("N_out", [384, 81, 4])
def test_linear_fn(K, N_out):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 100, 192, generator=g).to(dev())
    w = (torch.randn(N_out, 192, generator=g) / 14).to(dev()).requires_grad_(True)
    b = torch.randn(N_out, generator=g).to(dev()).requires_grad_(True)
    x16 = bf(x).requires_grad_(True)
    y = K.linear(x16, w, b, out_f32=True)
    go = torch.randn(y.shape, generator=g).to(dev())
    dx, dw, db = _grads((y * go).sum(), x16, w, b)
    xr = x16.detach().float().requires_grad_(True)
    wr = bf(w.detach()).float().requires_grad_(True)
    yr = F.linear(xr, wr, b)
    dxr, dwr, dbr = _grads((yr * bf(go).float()).sum(), xr, wr, b)
    assert rel_err(y, yr) < 1e-5
    assert rel_err(dx, dxr) < 1e-2 and rel_err(dw, dwr) < 1e-2 and rel_err(db, dbr) < 1e-2


@pytest.mark.parametrize("act,use_gamma", [("gelu", True), ("relu", False)])
def test_ffn_fn(K, act, use_gamma):
    g = torch.Generator().manual_seed(12)
    D, Fh = 192, 768
    x = torch.randn(2, 130, D, generator=g).to(dev())
    res = torch.randn(2, 130, D, generator=g).to(dev()).requires_grad_(True)
    w1 = (torch.randn(Fh, D, generator=g) / math.sqrt(D)).to(dev()).requires_grad_(True)
    b1 = (0.1 * torch.randn(Fh, generator=g)).to(dev()).requires_grad_(True)
    w2 = (torch.randn(D, Fh, generator=g) / math.sqrt(Fh)).to(dev()).requires_grad_(True)
    b2 = (0.1 * torch.randn(D, generator=g)).to(dev()).requires_grad_(True)
    gamma = (torch.rand(D, generator=g) + 0.5).to(dev()).requires_grad_(True) if use_gamma else None
    x16 = bf(x).requires_grad_(True)
    y = K.ffn(x16, w1, b1, w2, b2, res, gamma, act)
    go = torch.randn(y.shape, generator=g).to(dev())
    ins = [x16, w1, b1, w2, b2, res] + ([gamma] if use_gamma else [])
    got = _grads((y * go).sum(), *ins)
    xr = x16.detach().float().requires_grad_(True)
    actf = F.gelu if act == "gelu" else F.relu
    # reference on the bf16-rounded weights (so ReLU masks agree); hidden activation rounded like the kernel's
    w1r, w2r = bf(w1.detach()).float().requires_grad_(True), bf(w2.detach()).float().requires_grad_(True)
    yr = F.linear(actf(F.linear(xr, w1r, b1)), w2r, b2)
    yr = res + (gamma * yr if use_gamma else yr)
    refs = _grads((yr * go).sum(), *([xr, w1r, b1, w2r, b2] + ins[5:]))
    assert rel_err(y, yr) < 2e-2
    for a, b_, n in zip(got, refs, ["dx", "dw1", "db1", "dw2", "db2", "dres", "dgamma"]):
        assert rel_err(a, b_) < 3e-2, (n, rel_err(a, b_))


@pytest.mark.parametrize("D,eps", [(128, 1e-6), (192, 1e-6), (384, 1e-5), (768, 1e-5)])
def test_layernorm_fn(K, D, eps):
    g = torch.Generator().manual_seed(13)
    x = (torch.randn(5, 67, D, generator=g) * 2 + 0.3).to(dev()).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(dev()).requires_grad_(True)
    b = (0.1 * torch.randn(D, generator=g)).to(dev()).requires_grad_(True)
    y32, y16 = K.layernorm(x, w, b, eps, want_f32=True)
    g32 = torch.randn(y32.shape, generator=g).to(dev())
    g16 = bf(torch.randn(y32.shape, generator=g)).to(dev())
    got = _grads((y32 * g32).sum() + (y16.float() * g16.float()).sum(), x, w, b)
    yr = F.layer_norm(x, (D,), w, b, eps)
    ref = _grads((yr * (g32 + g16.float())).sum(), x, w, b)
    assert rel_err(y32, yr) < 1e-5 and rel_err(y16, yr) < 1e-2
    for a, b_ in zip(got, ref):
        assert rel_err(a, b_) < 1e-4, rel_err(a, b_)


@pytest.mark.parametrize("masked,two", [(False, False), (True, False), (False, True)])
def test_attention_fn(K, masked, two):
    g = torch.Generator().manual_seed(14)
    B, H, Lq, Lk, d, dv = 2, 8, 60, 100, 24, 24
    mk = lambda *s: bf(torch.randn(*s, generator=g)).to(dev()).requires_grad_(True)
    q, k, v = mk(B, Lq, H * d), mk(B, Lk, H * d), mk(B, Lk, H * dv)
    q2, k2 = (mk(B, Lq, H * d), mk(B, Lk, H * d)) if two else (None, None)
    mask = None
    if masked:
        mask = torch.zeros(B, Lk, dtype=torch.uint8)
        mask[0, 90:] = 1
        mask[1, 50:70] = 1
        mask = mask.to(dev())
    scale = (2 * d if two else d) ** -0.5
    out, pmean = K.attention(q, k, v, H, scale, mask_u8=mask, q2=q2, k2=k2, want_mean=True)
    go = bf(torch.randn(out.shape, generator=g)).to(dev())
    ins = [q, k, v] + ([q2, k2] if two else [])
    got = _grads((out.float() * go.float()).sum(), *ins)
    fr = [t.detach().float().requires_grad_(True) for t in ins]
    hd = lambda t, L, dd: t.view(B, L, H, dd).transpose(1, 2)
    S = hd(fr[0], Lq, d) @ hd(fr[1], Lk, d).transpose(-1, -2)
    if two:
        S = S + hd(fr[3], Lq, d) @ hd(fr[4], Lk, d).transpose(-1, -2)
    S = S * scale
    if masked:
        S = S.masked_fill(mask.bool()[:, None, None, :], float("-inf"))
    Pm = S.softmax(-1)
    outr = (Pm @ hd(fr[2], Lk, dv)).transpose(1, 2).reshape(B, Lq, H * dv)
    ref = _grads((outr * go.float()).sum(), *fr)
    assert rel_err(out, outr) < 2e-2
    assert rel_err(pmean, Pm.mean(1)) < 1e-2
    for a, b_ in zip(got, ref):
        assert rel_err(a, b_) < 3e-2, rel_err(a, b_)


@pytest.mark.parametrize("masked,two,Lq,Lk,d,dv,packed", [(False, False, 60, 100, 48, 48, False), (True, False, 300, 300, 48, 48, False),
                                                          (True, True, 600, 1600, 48, 48, False), (False, False, 1600, 1600, 48, 48, True),
                                                          (True, True, 130, 257, 32, 16, False), (False, False, 128, 128, 64, 64, False)])
def test_fused_attention_fwd(K, masked, two, Lq, Lk, d, dv, packed):
    """attn_fused.cu through ops.attention (no head-mean requested -> fused forward): output, the saved probabilities P (incl. zero
    padding columns) and the log-sum-exp against an fp32 reference; gradients flow through the (GEMM) backward from the saved P."""
    g = torch.Generator().manual_seed(140 + Lq)
    B, H = 2, 8
    mk = lambda *s: bf(torch.randn(*s, generator=g)).to(dev())
    if packed:       # q | k | v as strided views of one packed projection output (the encoder / backbone layout)
        qkv = mk(B, Lq, 3 * H * d)
        q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].requires_grad_(True) for i in range(3)]
    else:
        q, k, v = mk(B, Lq, H * d).requires_grad_(True), mk(B, Lk, H * d).requires_grad_(True), mk(B, Lk, H * dv).requires_grad_(True)
    q2, k2 = (mk(B, Lq, H * d).requires_grad_(True), mk(B, Lk, H * d).requires_grad_(True)) if two else (None, None)
    mask = None
    if masked:
        mask = torch.zeros(B, Lk, dtype=torch.uint8)
        mask[0, Lk - Lk // 7:] = 1
        mask[1, Lk // 3:Lk // 2] = 1
        mask = mask.to(dev())
    scale = (2 * d if two else d) ** -0.5
    assert K._fused_attention_ok(q, k, v, q2, k2, H)
    # raw kernel: P and lse
    ld = K.rup(Lk, 8)
    P = torch.full((B, H, Lq, ld), 7.0, dtype=torch.bfloat16, device=dev())
    lse = torch.empty((B, H, Lq), dtype=torch.float32, device=dev())
    out0 = torch.empty((B, Lq, H * dv), dtype=torch.bfloat16, device=dev())
    K.fused_attention_fwd(q.detach(), k.detach(), v.detach(), None if q2 is None else q2.detach(), None if k2 is None else k2.detach(), mask, H, scale,
                          out0, P=P, lse=lse)
    out = K.attention(q, k, v, H, scale, mask_u8=mask, q2=q2, k2=k2)
    assert torch.equal(out.detach(), out0)
    go = bf(torch.randn(out.shape, generator=g)).to(dev())
    ins = [q, k, v] + ([q2, k2] if two else [])
    got = _grads((out.float() * go.float()).sum(), *ins)
    fr = [t.detach().float().requires_grad_(True) for t in ins]
    hd = lambda t, L, dd: t.reshape(B, L, H, dd).transpose(1, 2)
    S = hd(fr[0], Lq, d) @ hd(fr[1], Lk, d).transpose(-1, -2)
    if two:
        S = S + hd(fr[3], Lq, d) @ hd(fr[4], Lk, d).transpose(-1, -2)
    S = S * scale
    if masked:
        S = S.masked_fill(mask.bool()[:, None, None, :], float("-inf"))
    Pm = S.softmax(-1)
    outr = (Pm @ hd(fr[2], Lk, dv)).transpose(1, 2).reshape(B, Lq, H * dv)
    ref = _grads((outr * go.float()).sum(), *fr)
    assert rel_err(out, outr) < 2e-2, rel_err(out, outr)
    assert rel_err(P[..., :Lk], Pm) < 1e-2, rel_err(P[..., :Lk], Pm)
    assert float(P[..., Lk:].float().abs().max()) == 0.0 if ld > Lk else True
    lse_ref = torch.logsumexp(S, -1) * 1.4426950408889634
    assert float((lse.cpu() - lse_ref.cpu()).abs().max()) < 2e-3
    for a, b_ in zip(got, ref):
        assert rel_err(a, b_) < 3e-2, rel_err(a, b_)
    # the GEMM-flavoured fused backward (attn_bwd_gemms_kernel with the softmax backward folded in): dP and P from HBM
    dP = torch.zeros((B, H, Lq, ld), dtype=torch.bfloat16, device=dev())
    dP[..., :Lk] = (hd(go.float(), Lq, dv) @ hd(fr[2].detach(), Lk, dv).transpose(-1, -2)).to(torch.bfloat16)
    delta = (go.float() * out0.float()).view(B, Lq, H, dv).sum(-1).permute(0, 2, 1).contiguous()
    dq, dk, dvv = torch.empty_like(q.detach()).contiguous(), torch.empty_like(k.detach()).contiguous(), torch.empty_like(v.detach()).contiguous()
    K.fused_attention_bwd_gemms(dP, P, q.detach(), k.detach(), go, H, scale, dq, dk, dvv, delta=delta)
    for a, b_ in zip((dq, dk, dvv), ref[:3]):
        assert rel_err(a, b_) < 3e-2, rel_err(a, b_)


@pytest.mark.parametrize("H,N,dh", [(2, 35, 64), (4, 196, 48), (8, 130, 48), (8, 1600, 48), (4, 1024, 48), (16, 300, 48), (6, 77, 48), (12, 261, 32)])   # (8, 1600) = cfg2 = the benchmarked shape (fused kernels); 16 = CaiT-M36 (cfg4, talking_h16.cu), 6/12: generic-H kernels
def test_talking_heads_attention_fn(K, H, N, dh):
    g = torch.Generator().manual_seed(15)
    B, D = 2, H * dh
    qkv = bf(torch.randn(B, N, 3 * D, generator=g)).to(dev()).requires_grad_(True)
    Wl = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev()).requires_grad_(True)
    bl = (0.1 * torch.randn(H, generator=g)).to(dev()).requires_grad_(True)
    Ww = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev()).requires_grad_(True)
    bw = (0.01 * torch.randn(H, generator=g)).to(dev()).requires_grad_(True)
    out = K.talking_heads_attention(qkv, Wl, bl, Ww, bw, H)
    go = bf(torch.randn(out.shape, generator=g)).to(dev())
    got = _grads((out.float() * go.float()).sum(), qkv, Wl, bl, Ww, bw)
    qr = qkv.detach().float().requires_grad_(True)
    t = qr.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    S = (t[0] * dh ** -0.5) @ t[1].transpose(-1, -2)
    L = torch.einsum("gh,bhij->bgij", Wl, S) + bl.view(1, H, 1, 1)
    P = L.softmax(-1)
    A = torch.einsum("gh,bhij->bgij", Ww, P) + bw.view(1, H, 1, 1)
    outr = (A @ t[2]).transpose(1, 2).reshape(B, N, D)
    ref = _grads((outr * go.float()).sum(), qr, Wl, bl, Ww, bw)
    assert rel_err(out, outr) < 2e-2
    for a, b_, n in zip(got, ref, ["dqkv", "dWl", "dbl", "dWw", "dbw"]):
        if n == "dbl":      # softmax is shift invariant: the true gradient is 0, both sides are rounding noise
            assert float(a.abs().max()) < 1e-3 * float(ref[1].abs().max()) + 1e-5
            continue
        assert rel_err(a, b_) < 3e-2, (n, rel_err(a, b_))


def test_patch_embed_bicubic_sine(K):
    from oracle import spe_oracle as O
    g = torch.Generator().manual_seed(16)
    B, D, p = 2, 128, 16
    img = torch.randn(B, 3, 48, 64, generator=g).to(dev())
    w = (torch.randn(D, 3, p, p, generator=g) / 27).to(dev()).requires_grad_(True)
    b = (0.1 * torch.randn(D, generator=g)).to(dev()).requires_grad_(True)
    pe = (0.5 * torch.randn(1, 6 * 7, D, generator=g)).to(dev()).requires_grad_(True)
    pos = K.BicubicTokensFn.apply(pe, 6, 7, 3, 4)
    x = K.PatchEmbedFn.apply(img, w, b, pos, p)
    go = torch.randn(x.shape, generator=g).to(dev())
    got = _grads((x * go).sum(), w, b, pe)
    w16 = bf(w.detach()).float().requires_grad_(True)
    xr = F.conv2d(bf(img).float(), w16, b, stride=p).flatten(2).transpose(1, 2)
    per = F.interpolate(pe.transpose(1, 2).reshape(1, D, 6, 7), size=(3, 4), mode="bicubic", align_corners=False).flatten(2).transpose(1, 2)
    xr = xr + per
    ref = _grads((xr * go).sum(), w16, b, pe)
    assert rel_err(pos, per[0]) < 1e-5
    assert rel_err(x, xr) < 1e-4
    assert rel_err(got[0], ref[0]) < 1e-2 and rel_err(got[1], ref[1]) < 1e-2 and rel_err(got[2], ref[2]) < 1e-2
    # sine encodings vs oracle
    mask = torch.zeros(2, 5, 7, dtype=torch.bool)
    mask[1, :, 5:] = True
    mask[1, 4:, :] = True
    pos32, pos16 = K.sine_pos_2d(mask.to(torch.uint8).to(dev()), 128)
    pr = O.sine_pos_2d(mask, 128).flatten(2).transpose(1, 2)
    assert float((pos32.cpu() - pr).abs().max()) < 2e-5
    ref_pts = torch.rand(3, 9, 2, generator=g).to(dev()).requires_grad_(True)
    emb = K.query_sine_embed(ref_pts, 192)
    rr = ref_pts.detach().cpu().requires_grad_(True)
    er = O.query_sine_embed(rr, 192)
    # arguments reach 2*pi*1e0 .. fp32 sin/cos of O(6) inputs: abs tol
    assert float((emb.cpu() - er).abs().max()) < 1e-4
    ge = torch.randn(emb.shape, generator=g)
    (dref,) = _grads((emb * ge.to(dev())).sum(), ref_pts)
    (drr,) = _grads((er * ge).sum(), rr)
    assert rel_err(dref.cpu(), drr) < 1e-4


# ------------------------------------------------------------------------------------------------
# matcher + criterion vs oracle
# ------------------------------------------------------------------------------------------------
def _rand_targets(g, B, C, max_gt, repeat=1, scores=False):
    tg = []
    for _ in range(B):
        n = int(torch.randint(0 if max_gt > 3 else 1, max_gt + 1, (1,), generator=g))
        c = torch.rand(n, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.05
        t = {"labels": torch.randint(1, C, (n,), generator=g).repeat_interleave(repeat),
             "boxes": torch.cat([c, wh], 1).repeat_interleave(repeat, 0)}
        if scores:
            t["scores"] = (torch.rand(n, generator=g) * 0.8 + 0.1).repeat_interleave(repeat)
        tg.append(t)
    return tg


def test_match_cost_and_lsap_vs_oracle():
    from oracle import spe_oracle as O
    from spe_b200 import criterion_ops as CO
    g = torch.Generator().manual_seed(21)
    B, Q, C = 6, 300, 81
    logits = torch.randn(B, Q, C, generator=g)
    boxes = torch.cat([torch.rand(B, Q, 2, generator=g) * 0.8 + 0.1, torch.rand(B, Q, 2, generator=g) * 0.45 + 0.02], -1)
    targets = _rand_targets(g, B, C, 10, repeat=5)
    T = CO.pack_targets(targets, dev())
    cost = CO.match_cost(logits.to(dev()), boxes.to(dev()), T, (2.0, 5.0, 2.0))
    r2g = CO.lsap(cost, T).cpu()
    for b, t in enumerate(targets):
        G = len(t["labels"])
        if G == 0:
            assert (r2g[b] == -1).all()
            continue
        cref = O.match_cost(logits[b], boxes[b], t["labels"], t["boxes"])
        assert float((cost[b, :, :G].cpu() - cref).abs().max()) < 5e-6
        i, j = O.lsap(cref)
        rows = torch.nonzero(r2g[b] >= 0).flatten()
        assert torch.equal(rows, i) and torch.equal(r2g[b][rows].long(), j)


def test_matcher_cfg5_full_size_bit_exact():
    """BASELINE configs[4]: 300 queries x 1000 GT, batch 256, through HungarianMatcher.forward (the reference API).  Every image is
    compared with the oracle (reference cost arithmetic in torch + scipy) -- indices bit-exact -- plus the size-independent
    properties of an assignment: min(Q, G) pairs, no query and no GT used twice, predictions sorted (scipy's transposed branch)."""
    import bench
    from oracle import spe_oracle as O
    from spe_b200.models.matcher import HungarianMatcher
    logits, boxes, targets = bench.cfg5_inputs()
    m = HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2)
    got = m({"pred_logits": logits.to(dev()), "pred_boxes": boxes.to(dev())}, [{k: v.to(dev()) for k, v in t.items()} for t in targets])
    assert len(got) == 256
    ref = O.hungarian_match(logits, boxes, targets)
    for (i, j), (ri, rj) in zip(got, ref):
        assert i.dtype == torch.int64 and j.dtype == torch.int64 and not i.is_cuda
        assert len(i) == 300 and len(set(i.tolist())) == 300 and len(set(j.tolist())) == 300
        assert torch.equal(i, torch.arange(300))
        assert torch.equal(i, ri) and torch.equal(j, rj)


def test_lsap_kernel_bit_exact_on_scipy_vectors(golden_dir):
    from spe_b200 import criterion_ops as CO
    from tests.test_lsap_oracle import regen_lsap_cases
    n = 0
    for c, case in regen_lsap_cases(golden_dir):
        nr, nc = c.shape
        r2c = CO.lsap_raw(c.to(dev()).unsqueeze(0).contiguous(), None)[0].cpu()
        rows = torch.nonzero(r2c >= 0).flatten()
        assert torch.equal(rows, case["rows"]) and torch.equal(r2c[rows].long(), case["cols"]), (case["shape"], case["kind"])
        n += 1
    assert n == 33


@pytest.mark.parametrize("refine,gamma", [(False, 2.0), (True, 0.5)])
def test_criterion_kernels_vs_oracle(refine, gamma):
    from oracle import spe_oracle as O
    from spe_b200 import criterion_ops as CO
    g = torch.Generator().manual_seed(22)
    B, Q, C = 4, 50, 21
    logits = torch.randn(B, Q, C, generator=g).requires_grad_(True)
    boxes = torch.cat([torch.rand(B, Q, 2, generator=g) * 0.8 + 0.1, torch.rand(B, Q, 2, generator=g) * 0.45 + 0.02], -1).requires_grad_(True)
    targets = _rand_targets(g, B, C, 3, repeat=2, scores=refine)
    ld, idx = O.criterion_forward({"pred_logits": logits, "pred_boxes": boxes}, targets, gamma=gamma, refine=refine, return_indices=True)
    gl_ce, = torch.autograd.grad(ld["loss_ce"], logits, retain_graph=True)
    gb_l1, = torch.autograd.grad(ld["loss_bbox"], boxes, retain_graph=True)
    gb_g, = torch.autograd.grad(ld["loss_giou"], boxes)
    T = CO.pack_targets(targets, dev())
    lg, bx = logits.detach().to(dev()).requires_grad_(True), boxes.detach().to(dev()).requires_grad_(True)
    out = CO.set_losses(lg, bx, T, (2.0, 5.0, 2.0), 0.25, gamma, refine)
    for k in ("loss_ce", "loss_bbox", "loss_giou", "class_error", "cardinality_error"):
        assert abs(float(out[k]) - float(ld[k])) < 1e-3 * max(1.0, abs(float(ld[k]))), (k, float(out[k]), float(ld[k]))
    a, = torch.autograd.grad(out["loss_ce"], lg, retain_graph=True)
    b1, = torch.autograd.grad(out["loss_bbox"], bx, retain_graph=True)
    b2, = torch.autograd.grad(out["loss_giou"], bx)
    assert rel_err(a.cpu(), gl_ce) < 1e-3 and rel_err(b1.cpu(), gb_l1) < 1e-3 and rel_err(b2.cpu(), gb_g) < 1e-3
