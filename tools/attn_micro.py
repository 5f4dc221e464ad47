"""Timing of the fused attention forward (attn_fused.cu) at the detector's shapes (CUDA events, rotating buffers > L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spe_b200 import ops
dev = torch.device("cuda")

def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def case(name, B, H, Lq, Lk, d, dv, two, store_p=True, masked=False):
    q = torch.randn(B, Lq, H * d, device=dev).to(torch.bfloat16); k = torch.randn(B, Lk, H * d, device=dev).to(torch.bfloat16)
    v = torch.randn(B, Lk, H * dv, device=dev).to(torch.bfloat16)
    q2 = torch.randn_like(q) if two else None; k2 = torch.randn_like(k) if two else None
    ld = ops.rup(Lk, 8)
    P = torch.empty(B, H, Lq, ld, device=dev, dtype=torch.bfloat16) if store_p else None
    out = torch.empty(B, Lq, H * dv, device=dev, dtype=torch.bfloat16)
    mask = torch.zeros(B, Lk, dtype=torch.uint8, device=dev) if masked else None
    scale = (d * (2 if two else 1)) ** -0.5
    ms = timeit(lambda: ops.fused_attention_fwd(q, k, v, q2, k2, mask, H, scale, out, P=P))
    flops = 4.0 * B * H * Lq * Lk * (d * (2 if two else 1) + dv) / 2 * 1.0   # QK^T (once) + PV, 2*MAC
    print("%-28s %.3f ms   %.1f TFLOP/s (algorithmic QK^T+PV)   P bytes %.0f MB -> %.0f GB/s" % (
        name, ms, flops / ms / 1e9, (P.numel() * 2 / 1e6 if P is not None else 0), (P.numel() * 2 / ms / 1e6 if P is not None else 0)))

case("encoder 1600x1600 P", 8, 8, 1600, 1600, 48, 48, False, masked=True)
case("encoder 1600x1600 noP", 8, 8, 1600, 1600, 48, 48, False, store_p=False, masked=True)
case("cross 600x1600 two P", 8, 8, 600, 1600, 48, 48, True, masked=True)
case("cross 600x1600 two noP", 8, 8, 600, 1600, 48, 48, True, store_p=False, masked=True)
case("self 300x300 x16 P", 16, 8, 300, 300, 48, 48, False)
case("class 81x1681 P", 8, 8, 81, 1681, 48, 48, False)

# ---- backward GEMMs: fused (one pass over dS, P) vs softmax-backward folded in (from dP) -- graph-captured timing
def timeit_graph(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def bwd_case(name, B, H, Lq, Lk, d):
    q = torch.randn(B, Lq, H * d, device=dev).to(torch.bfloat16); k = torch.randn(B, Lk, H * d, device=dev).to(torch.bfloat16)
    dO = torch.randn(B, Lq, H * d, device=dev).to(torch.bfloat16)
    ld = ops.rup(Lk, 8)
    dS = (torch.randn(B, H, Lq, ld, device=dev) * 0.01).to(torch.bfloat16); P = torch.rand(B, H, Lq, ld, device=dev).to(torch.bfloat16)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(k)
    delta = torch.zeros(B, H, Lq, device=dev)
    t0 = timeit_graph(lambda: ops.fused_attention_bwd_gemms(dS, P, q, k, dO, H, 0.1, dq, dk, dv))
    t1 = timeit_graph(lambda: ops.fused_attention_bwd_gemms(dS, P, q, k, dO, H, 0.1, dq, dk, dv, delta=delta))
    t2 = timeit_graph(lambda: ops.fused_attention_bwd_gemms(dS, None, q, k, None, H, 0.1, dq, dk, None))
    n2 = B * H * Lq * ld * 2 / 1e6
    print("%-24s fused dS+P %.3f ms (%.0f GB/s)   from dP %.3f ms   dQ,dK only %.3f ms (%.0f GB/s)" % (name, t0, 2 * n2 / t0 / 1e3, t1, t2, n2 / t2 / 1e3))

bwd_case("bwd 1600x1600", 8, 8, 1600, 1600, 48)
bwd_case("bwd cross 600x1600", 8, 8, 600, 1600, 48)

# ---- fully fused backward (recompute)
from spe_b200 import _lib
import ctypes as C
def bwd2_case(name, B, H, Lq, Lk, d, two):
    mk = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
    q, k, v, dO = mk(B, Lq, H * d), mk(B, Lk, H * d), mk(B, Lk, H * d), mk(B, Lq, H * d)
    q2, k2 = (mk(B, Lq, H * d), mk(B, Lk, H * d)) if two else (None, None)
    out = torch.empty_like(dO); lse = torch.empty(B, H, Lq, device=dev)
    scale = (d * (2 if two else 1)) ** -0.5
    mask = torch.zeros(B, Lk, dtype=torch.uint8, device=dev)
    ops.fused_attention_fwd(q, k, v, q2, k2, mask, H, scale, out, lse=lse)
    class Ctx: pass
    ctx = Ctx(); ctx.saved_tensors = (q, k, v, q2, k2, lse, out, mask); ctx.H, ctx.scale = H, scale
    t = timeit_graph(lambda: ops._attention_backward_recompute(ctx, dO))
    flops = 2.0 * B * H * Lq * Lk * (d * (2 if two else 1) * 3 + d * 2) * (2 if two else 1)
    print("%-28s fused recompute bwd %.3f ms" % (name, t))
bwd2_case("bwd2 encoder 1600x1600", 8, 8, 1600, 1600, 48, False)
bwd2_case("bwd2 cross 600x1600 two", 8, 8, 600, 1600, 48, True)
bwd2_case("bwd2 self 300x300 x16", 16, 8, 300, 300, 48, False)
