"""The kernels added in round 2, one launch set each, as a short command for ncu:
  * csrc/talking_fused.cu (SPE_TH_FUSED=1): fused talking-heads forward (stats + main) and the three recomputing backward kernels at the
    cfg2 backbone shape (B=8, H=8, N=1600, dh=48);
  * csrc/talking_h16.cu: the H=16 mix/softmax/mix kernels at the cfg4 shape (B=1, H=16, N=4150), fp16 logits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spe_b200 import ops

torch.manual_seed(0)
dev = torch.device("cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def run(B, H, N, fused):
    dh = 48
    D = H * dh
    ops._TH_FUSED = "1" if fused else "0"
    qkv = (torch.randn(B, N, 3 * D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    Wl = (torch.eye(H, device=dev) + 0.1 * torch.randn(H, H, device=dev)).requires_grad_(True)
    Ww = (torch.eye(H, device=dev) + 0.1 * torch.randn(H, H, device=dev)).requires_grad_(True)
    bl = torch.zeros(H, device=dev, requires_grad=True)
    bw = torch.zeros(H, device=dev, requires_grad=True)
    for _ in range(reps):
        o = ops.talking_heads_attention(qkv, Wl, bl, Ww, bw, H)
        o.float().sum().backward()
    torch.cuda.synchronize()


run(8, 8, 1600, True)
run(1, 16, 4150, False)
print("done")
