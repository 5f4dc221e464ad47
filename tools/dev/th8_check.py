"""H = 16 talking-heads kernels (csrc/talking_h16.cu) vs fp32 at a ragged small shape and at the cfg4 token count, both logit storage
formats, + per-kernel device time (library profiler).  SPE_TH16_GENERIC=1 in the environment times the CUDA-core kernels instead."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spe_b200 import _lib, ops as K

dev = torch.device("cuda")
rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def case(H, N, B, s16, check=True):
    K._TH_S16 = s16
    g = torch.Generator().manual_seed(31)
    dh = 48
    D = H * dh
    qkv = torch.randn(B, N, 3 * D, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    Wl = (torch.eye(H) + 0.2 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    bl = (0.1 * torch.randn(H, generator=g)).to(dev).requires_grad_(True)
    Ww = (torch.eye(H) + 0.2 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    bw = (0.01 * torch.randn(H, generator=g)).to(dev).requires_grad_(True)
    go = torch.randn(B, N, D, generator=g).to(torch.bfloat16).to(dev)
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize(); _lib.prof_enable(True)
        out = K.talking_heads_attention(qkv, Wl, bl, Ww, bw, H)
        got = torch.autograd.grad((out.float() * go.float()).sum(), [qkv, Wl, Ww])
    _lib.prof_enable(False)
    fam = _lib.prof_collect()
    msg = "H=%d N=%d B=%d s16=%d: fwd %.3f ms bwd %.3f ms" % (H, N, B, s16, fam["talking_softmax_fwd"][0], fam["talking_softmax_bwd"][0])
    if check:
        qr = qkv.detach().float().requires_grad_(True)
        Wlr, Wwr = Wl.detach().clone().requires_grad_(True), Ww.detach().clone().requires_grad_(True)
        t = qr.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
        S = (t[0] * dh ** -0.5) @ t[1].transpose(-1, -2)
        L = torch.einsum("gh,bhij->bgij", Wlr, S) + bl.detach().view(1, H, 1, 1)
        A = torch.einsum("gh,bhij->bgij", Wwr, L.softmax(-1)) + bw.detach().view(1, H, 1, 1)
        outr = (A @ t[2]).transpose(1, 2).reshape(B, N, D)
        ref = torch.autograd.grad((outr * go.float()).sum(), [qr, Wlr, Wwr])
        errs = [rel(out, outr)] + [rel(a, b) for a, b in zip(got, ref)]
        msg += "  rel err out %.2e dqkv %.2e dWl %.2e dWw %.2e %s" % (*errs, "OK" if max(errs) < 3e-2 and errs[0] < 2e-2 else "FAIL")
    print(msg, flush=True)


for s16 in (True, False):
    case(8, 300, 2, s16)
    case(8, 77, 1, s16)
    case(8, 1600, 8, s16, check=(os.environ.get("TH8_CHECK_BIG", "1") == "1"))
