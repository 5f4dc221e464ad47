"""BASELINE configs[3] at full size on one GPU: TSCAM-M36 (D 768, depth 36, 16 heads) + conditional DETR, 300 queries, 81 logits,
3x800x1333 synthetic (N = 50 x 83 = 4150 tokens), batch 1: one fwd + both criteria + bwd, finite losses / gradients, time, peak memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from types import SimpleNamespace
from spe_b200 import factory
from spe_b200.dp import FlatGradBuffer

dev = torch.device("cuda")
cfg = SimpleNamespace(embed_dim=768, depth=36, num_heads=16, img_classes=80, patch=16, layer_to_det=35, depth_token_only=2, mlp_ratio=4.0,
                      pos_grid=(50, 84), det_heads=8, ffn=2048, enc_layers=6, dec_layers=6, num_queries=300, det_classes=81, num_refines=1,
                      ln_eps_backbone=1e-6, ln_eps_detr=1e-5)
B = int(os.environ.get("SPE_BATCH", "1"))
torch.manual_seed(0)
model = factory.build_detector(cfg, dev).train()
crit = factory.build_criterion(cfg, device=dev).eval()
crit_ref = factory.build_criterion(cfg, refine=True, device=dev).eval()
wd = crit.weight_dict
buf = FlatGradBuffer(model.parameters())
images = torch.randn(B, 3, 800, 1333, device=dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in bench.synth_targets(B, 7)]

def step():
    buf.zero_()
    out = model(images)
    ld, ld2 = crit(out[0], targets), crit_ref(out[1], targets)
    loss = sum(ld[k] * wd[k] for k in ld if k in wd) + sum(ld2[k] * wd[k] for k in ld2 if k in wd)
    loss.backward()
    return loss

for i in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = step()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    gn = float(buf.flat.norm())
    print("cfg4 step %d: loss %.4f finite=%s grad-norm %.4g  %.1f ms  (%.2f img/s)  peak mem %.1f GB  params %.1f M" % (
        i, float(loss), bool(torch.isfinite(loss)), gn, dt * 1e3, B / dt, torch.cuda.max_memory_allocated() / 1e9,
        sum(p.numel() for p in model.parameters()) / 1e6), flush=True)
assert torch.isfinite(loss) and gn == gn and gn > 0
