"""Small invocations of every kernel added in round 2, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/dev/sanitize_new.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spe_b200 import ops as K, criterion_ops as CO, pseudo_labels as PL
from spe_b200.dp import FlatGradBuffer
from spe_b200.optim import FlatAdamW, clip_grad_norm_

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)


def th(H, N, B, fused, s16=True):
    K._TH_FUSED = "1" if fused else "0"
    K._TH_S16 = s16
    D = H * 48
    qkv = torch.randn(B, N, 3 * D, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    Wl = (torch.eye(H) + 0.2 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    Ww = (torch.eye(H) + 0.2 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    bl = torch.zeros(H, device=dev, requires_grad=True); bw = torch.zeros(H, device=dev, requires_grad=True)
    o = K.talking_heads_attention(qkv, Wl, bl, Ww, bw, H)
    o.float().sum().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(qkv.grad.float()).all()
    print("talking heads H=%d N=%d fused=%s s16=%s ok" % (H, N, fused, s16), flush=True)


for (H, N, B, fused, s16) in [(8, 77, 1, False, True), (8, 300, 2, False, True), (8, 264, 1, False, True), (16, 77, 1, False, True), (16, 130, 1, False, False),
                              (8, 100, 1, True, True), (4, 70, 1, True, True)]:
    th(H, N, B, fused, s16)

# CAM boxes (single + multi), ragged sizes
cams = torch.randn(2, 3, 9, 11, generator=g).to(dev)
pairs = torch.tensor([[0, 0], [0, 2], [1, 1]])
print(PL.cam_boxes(cams, pairs, (70, 90), 0.2).cpu().tolist())
b, c = PL.cam_boxes_multi(cams, pairs, (70, 90), 0.2, 0.1, max_boxes=8)
print(c.cpu().tolist())
# GT jitter
tg = [{"labels": torch.tensor([1, 2, 3]), "boxes": torch.tensor([[0.5, 0.5, 0.2, 0.3], [0.3, 0.4, 0.1, 0.1], [0.7, 0.6, 0.3, 0.2]])}, {"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros(0, 4)}]
E = CO.jitter_repeat(CO.pack_targets(tg, dev), 5, 0.1, CO.JitterRng(dev, seed=3))
print(E.boxes[:15].cpu().shape, E.offsets.cpu().tolist())
# optimizer
m = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).to(dev)
gb = FlatGradBuffer(m.parameters())
opt = FlatAdamW(m, gb, clip_max_norm=0.1, refresh_shadows=False)
for p in m.parameters():
    p.grad.copy_(torch.randn_like(p))
opt.step(); clip_grad_norm_(gb, 0.1)
torch.cuda.synchronize()
print("all ok")
