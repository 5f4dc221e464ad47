"""One talking-heads attention fwd+bwd at cfg2 shapes (B=8,H=8,N=1600,dh=48) + one LN/FFN block: a short command for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spe_b200 import ops

torch.manual_seed(0)
dev = torch.device("cuda")
B, H, N, dh = 8, 8, 1600, 48
D = H * dh
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
qkv = (torch.randn(B, N, 3 * D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
Wl = (torch.eye(H, device=dev) + 0.1 * torch.randn(H, H, device=dev)).requires_grad_(True)
Ww = (torch.eye(H, device=dev) + 0.1 * torch.randn(H, H, device=dev)).requires_grad_(True)
bl = torch.zeros(H, device=dev, requires_grad=True)
bw = torch.zeros(H, device=dev, requires_grad=True)
x = torch.randn(B, N, D, device=dev).to(torch.bfloat16).requires_grad_(True)
res = torch.randn(B, N, D, device=dev)
w1 = (torch.randn(4 * D, D, device=dev) / 20).requires_grad_(True); b1 = torch.zeros(4 * D, device=dev, requires_grad=True)
w2 = (torch.randn(D, 4 * D, device=dev) / 40).requires_grad_(True); b2 = torch.zeros(D, device=dev, requires_grad=True)
gamma = torch.full((D,), 0.1, device=dev, requires_grad=True)
# encoder self-attention (fused forward, attn_fused.cu) and conditional cross-attention (two QK segments) at cfg2 shapes
eq = (torch.randn(B, N, 2 * D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
ev = (torch.randn(B, N, D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
mask = torch.zeros(B, N, dtype=torch.uint8, device=dev)
cq = [(torch.randn(B, 600, D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True) for _ in range(2)]
ck = [(torch.randn(B, N, D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True) for _ in range(3)]
for _ in range(reps):
    ea = ops.attention(eq[:, :, :D], eq[:, :, D:], ev, H, dh ** -0.5, mask_u8=mask)
    ca = ops.attention(cq[0], ck[0], ck[2], H, (2 * dh) ** -0.5, mask_u8=mask, q2=cq[1], k2=ck[1])
    (ea.float().sum() + ca.float().sum()).backward()
    o = ops.talking_heads_attention(qkv, Wl, bl, Ww, bw, H)
    y = ops.ffn(x, w1, b1, w2, b2, res, gamma, "gelu")
    (o.float().sum() + y.sum()).backward()
torch.cuda.synchronize()
print("done")
