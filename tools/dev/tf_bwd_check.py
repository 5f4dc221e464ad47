"""Development check of the fused talking-heads backward (delta / dq / dkv kernels) against fp32 torch autograd. GPU only."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, ".")
from spe_b200 import _lib
from spe_b200._lib import lib, check, stream

def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))

def run(B, H, N, dh=48, seed=0):
    g = torch.Generator().manual_seed(seed)
    D = H * dh
    dev = "cuda"
    qkv = (torch.randn(B, N, 3 * D, generator=g)).to(torch.bfloat16).to(dev)
    Wl = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev)
    bl = (0.1 * torch.randn(H, generator=g)).to(dev)
    Ww = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev)
    bw = (0.01 * torch.randn(H, generator=g)).to(dev)
    dO = (torch.randn(B, N, D, generator=g)).to(torch.bfloat16).to(dev)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    out = torch.zeros(B, N, D, dtype=torch.bfloat16, device=dev)
    lse2 = torch.zeros(B, H, N, dtype=torch.float32, device=dev)
    ws = torch.zeros(int(lib().spe_talking_fused_fwd_workspace(B, H, N, dh)), dtype=torch.uint8, device=dev)
    a = _lib.TalkingFusedArgs(B, H, N, dh, q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(), v.stride(1), v.stride(0),
                              Wl.data_ptr(), bl.data_ptr(), Ww.data_ptr(), bw.data_ptr(), dh ** -0.5, out.data_ptr(), out.stride(1), out.stride(0),
                              lse2.data_ptr(), ws.data_ptr(), ws.numel())
    check(lib().spe_talking_fused_fwd(C.byref(a), stream()))
    dqkv = torch.zeros(B, N, 3 * D, dtype=torch.bfloat16, device=dev)
    dWl = torch.zeros(H, H, device=dev)
    dWw = torch.zeros(H, H, device=dev)
    ws2 = torch.zeros(int(lib().spe_talking_fused_bwd_workspace(B, H, N, dh)), dtype=torch.uint8, device=dev)
    ab = _lib.TalkingFusedBwdArgs(B, H, N, dh, q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(), v.stride(1), v.stride(0),
                                  dO.data_ptr(), dO.stride(1), dO.stride(0), Wl.data_ptr(), bl.data_ptr(), Ww.data_ptr(), bw.data_ptr(), dh ** -0.5, lse2.data_ptr(),
                                  dqkv.data_ptr(), dqkv.stride(1), dWl.data_ptr(), dWw.data_ptr(), ws2.data_ptr(), ws2.numel())
    check(lib().spe_talking_fused_bwd(C.byref(ab), stream()))
    torch.cuda.synchronize()
    qr = qkv.float().requires_grad_(True)
    Wlr, Wwr = Wl.clone().requires_grad_(True), Ww.clone().requires_grad_(True)
    t = qr.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    S = (t[0] * dh ** -0.5) @ t[1].transpose(-1, -2)
    L = torch.einsum("gh,bhij->bgij", Wlr, S) + bl.view(1, H, 1, 1)
    P = L.softmax(-1)
    A = torch.einsum("gh,bhij->bgij", Wwr, P) + bw.view(1, H, 1, 1)
    ref = (A @ t[2]).transpose(1, 2).reshape(B, N, D)
    (ref * dO.float()).sum().backward()
    gq = qr.grad
    e = dict(dq=rel(dqkv[:, :, :D], gq[:, :, :D]), dk=rel(dqkv[:, :, D:2 * D], gq[:, :, D:2 * D]), dv=rel(dqkv[:, :, 2 * D:], gq[:, :, 2 * D:]),
             dWl=rel(dWl, Wlr.grad), dWw=rel(dWw, Wwr.grad))
    print("B=%d H=%d N=%d " % (B, H, N), " ".join("%s %.3e" % kv for kv in e.items()), flush=True)
    return max(v if v == v else 1e9 for v in e.values())

if __name__ == "__main__":
    shapes = [(1, 8, 64), (1, 8, 16), (2, 8, 130), (2, 4, 196), (1, 8, 1600), (2, 8, 333)]
    if os.environ.get("TF_ONLY_BIG"):
        shapes = []
    bad = 0
    for (B, H, N) in shapes:
        bad += run(B, H, N) > 3e-2
    if os.environ.get("TF_TIME"):
        B, H, N = 8, 8, 1600
        run(B, H, N)
        _lib.prof_enable(True)
        for _ in range(3):
            run(B, H, N)
        _lib.prof_collect()
        if os.environ.get("SPE_PROF_CSV"):
            import collections
            acc = collections.defaultdict(list)
            for line in open(os.environ["SPE_PROF_CSV"]):
                f = line.strip().split(",")
                acc[f[1]].append(float(f[3]))
            print("TIMES", {k: round(1000 * sum(v) / len(v), 1) for k, v in acc.items()}, "us")
    if int(os.environ.get("SPE_TF_DBG", "0")) & 1024:
        import numpy as np
        buf = (C.c_longlong * (4 * 32 * 8))()
        lib().spe_talking_fused_trace.argtypes = [C.c_void_p]
        lib().spe_talking_fused_trace(buf)
        t = np.array(buf[:]).reshape(4, 32, 8)
        t0 = t[t > 0].min()
        names = {1: "MMA  [top, yfull, sempty, tiles issued, tfull, acc issued]", 2: "POS  [top, sfull, loaded, tempty, math done, arrived]"}
        for role in (1, 2):
            print(names[role])
            for blk in range(4, 12):
                print("   blk %2d " % blk, " ".join("%7d" % (v - t0 if v else -1) for v in t[role, blk, :6]))
    print("FAIL" if bad else "OK")
