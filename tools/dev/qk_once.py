"""A few eager launches of the K = 48 batched QK^T GEMM (fp16 logits out) at the cfg2 backbone shape, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spe_b200 import ops
dev = torch.device("cuda")
B, H, N, dh = 8, 8, 1600, 48
D = H * dh
q = torch.randn(B, N, D, device=dev).to(torch.bfloat16); k = torch.randn(B, N, D, device=dev).to(torch.bfloat16)
S = torch.empty(B, H, N, N, device=dev, dtype=torch.float16)
for _ in range(4):
    ops._qk_logits(q, k, H, 0.1, S, N)
torch.cuda.synchronize()
print("done")
