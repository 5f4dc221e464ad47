"""Timing experiments on single GEMM shapes (CUDA events, L2 flushed between reps by rotating buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spe_b200 import ops
dev = torch.device("cuda")

def timeit(fn, reps=20):
    """GPU time per launch: the launches are captured into a CUDA graph (eager launches from Python are host-bound at ~15 us each)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B, H, N, dh = 8, 8, 1600, 48
D = H * dh
q = torch.randn(B, N, D, device=dev).to(torch.bfloat16); k = torch.randn(B, N, D, device=dev).to(torch.bfloat16)
S32 = torch.empty(B, H, N, N, device=dev); S16 = torch.empty(B, H, N, N, device=dev, dtype=torch.bfloat16)
x = torch.randn(B * N, D, device=dev).to(torch.bfloat16)
w1 = torch.randn(4 * D, D, device=dev).to(torch.bfloat16); b1 = torch.zeros(4 * D, device=dev)
h = torch.empty(B * N, 4 * D, device=dev, dtype=torch.bfloat16); a = torch.empty_like(h)
wq = torch.randn(3 * D, D, device=dev).to(torch.bfloat16); oq = torch.empty(B * N, 3 * D, device=dev, dtype=torch.bfloat16)
res = torch.randn(B * N, D, device=dev); w2 = torch.randn(D, 4 * D, device=dev).to(torch.bfloat16); o32 = torch.empty(B * N, D, device=dev)
tag = os.environ.get("SPE_GEMM_DBG", "0")
print("dbg=%s S fp32 : %.3f ms" % (tag, timeit(lambda: ops._qk_logits(q, k, H, 0.1, S32, N))))
print("dbg=%s S bf16 : %.3f ms" % (tag, timeit(lambda: ops._qk_logits(q, k, H, 0.1, S16, N))))
print("dbg=%s qkv    : %.3f ms" % (tag, timeit(lambda: ops.gemm(x, wq, oq, B * N, 3 * D, D, lda=D, ldb=D, ldc=3 * D))))
print("dbg=%s fc1gelu: %.3f ms" % (tag, timeit(lambda: ops.gemm(x, w1, h, B * N, 4 * D, D, lda=D, ldb=D, ldc=4 * D, bias=b1, act=ops.ACT_GELU, aux_out=a, ld_aux=4 * D))))
print("dbg=%s fc2res : %.3f ms" % (tag, timeit(lambda: ops.gemm(h, w2, o32, B * N, D, 4 * D, lda=4 * D, ldb=4 * D, ldc=D, residual=res, ldr=D))))
