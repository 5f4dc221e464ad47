"""Development check of the fused talking-heads forward (stats + main kernels) against fp32 torch. GPU only."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from spe_b200 import _lib
from spe_b200._lib import lib, check, stream

def run(B, H, N, dh=48, seed=0, nchunk=None):
    g = torch.Generator().manual_seed(seed)
    D = H * dh
    dev = "cuda"
    qkv = (torch.randn(B, N, 3 * D, generator=g)).to(torch.bfloat16).to(dev)
    Wl = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev)
    bl = (0.1 * torch.randn(H, generator=g)).to(dev)
    Ww = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev)
    bw = (0.01 * torch.randn(H, generator=g)).to(dev)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    out = torch.zeros(B, N, D, dtype=torch.bfloat16, device=dev)
    lse2 = torch.zeros(B, H, N, dtype=torch.float32, device=dev)
    ws = torch.zeros(int(lib().spe_talking_fused_fwd_workspace(B, H, N, dh)), dtype=torch.uint8, device=dev)
    a = _lib.TalkingFusedArgs(B, H, N, dh, q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(), v.stride(1), v.stride(0),
                              Wl.data_ptr(), bl.data_ptr(), Ww.data_ptr(), bw.data_ptr(), dh ** -0.5, out.data_ptr(), out.stride(1), out.stride(0),
                              lse2.data_ptr(), ws.data_ptr(), ws.numel())
    check(lib().spe_talking_fused_fwd(C.byref(a), stream()))
    torch.cuda.synchronize()
    t = qkv.float().view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    S = (t[0] * dh ** -0.5) @ t[1].transpose(-1, -2)
    L = torch.einsum("gh,bhij->bgij", Wl, S) + bl.view(1, H, 1, 1)
    P = L.softmax(-1)
    A = torch.einsum("gh,bhij->bgij", Ww, P) + bw.view(1, H, 1, 1)
    ref = (A @ t[2]).transpose(1, 2).reshape(B, N, D)
    lse_ref = torch.logsumexp(L, -1) * 1.4426950408889634
    e_lse = float((lse2 - lse_ref).abs().max())
    e_out = float((out.float() - ref).norm() / ref.norm())
    print("B=%d H=%d N=%d  lse max abs err %.3e   out rel err %.3e" % (B, H, N, e_lse, e_out), flush=True)
    return e_lse, e_out

if __name__ == "__main__":
    import os
    shapes = [(1, 8, 64), (1, 8, 16), (2, 8, 130), (2, 4, 196), (1, 8, 1600), (2, 8, 333)]
    if os.environ.get("TF_ONLY_BIG"):
        shapes = []
    bad = 0
    for (B, H, N) in shapes:
        e1, e2 = run(B, H, N)
        bad += (e1 > 5e-2) or (e2 > 2e-2) or e1 != e1 or e2 != e2
    if os.environ.get("TF_TIME"):
        B, H, N = 8, 8, 1600
        import time
        run(B, H, N)
        _lib.prof_enable(True)
        for _ in range(5):
            run(B, H, N)
        r = _lib.prof_collect()
        print({k: v for k, v in r.items() if v[2]})
        if os.environ.get("SPE_PROF_CSV"):
            import collections
            acc = collections.defaultdict(list)
            for line in open(os.environ["SPE_PROF_CSV"]):
                f = line.strip().split(",")
                acc[f[1]].append(float(f[3]))
            print("DBG", os.environ.get("SPE_TF_DBG"), "NCHUNK", os.environ.get("SPE_TF_NCHUNK"), {k: round(1000 * sum(v) / len(v), 1) for k, v in acc.items()}, "us")
    if int(os.environ.get("SPE_TF_DBG", "0")) & 512:
        import numpy as np
        buf = (C.c_longlong * (4 * 32 * 8))()
        lib().spe_talking_fused_trace.argtypes = [C.c_void_p]
        lib().spe_talking_fused_trace(buf)
        t = np.array(buf[:]).reshape(4, 32, 8)
        t0 = t[t > 0].min()
        names = {0: "TMA  [yempty, zempty]", 1: "MMA  [top, yfull, sempty, S issued, afull, zfull, PV issued]", 2: "POS  [top, exp done, token, mix2+st done, next load+mix1 done, next load done]"}
        for role in (0, 1, 2):
            print(names[role])
            for blk in range(4, 14):
                print("   blk %2d " % blk, " ".join("%7d" % (v - t0 if v else -1) for v in t[role, blk, :7]))
    print("FAIL" if bad else "OK")
