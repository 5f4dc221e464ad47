"""One profiled training step of the bench workload (cfg2, bs 8) between cudaProfilerStart/Stop, after 2 warm-up steps.
Used with:  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spe_b200 import factory
from spe_b200.dp import FlatGradBuffer

dev = torch.device("cuda")
cfg = bench.cfg2()
torch.manual_seed(42)
B = int(os.environ.get("SPE_BATCH", "8"))
model = factory.build_detector(cfg, dev).train()
crit = factory.build_criterion(cfg, device=dev).eval()
crit_ref = factory.build_criterion(cfg, refine=True, device=dev).eval()
wd = crit.weight_dict
buf = FlatGradBuffer(model.parameters())
images = torch.randn(B, 3, 640, 640, device=dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in bench.synth_targets(B, 7)]

def step():
    buf.zero_()
    out = model(images)
    ld, ld2 = crit(out[0], targets), crit_ref(out[1], targets)
    loss = sum(ld[k] * wd[k] for k in ld if k in wd) + sum(ld2[k] * wd[k] for k in ld2 if k in wd)
    loss.backward()
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
