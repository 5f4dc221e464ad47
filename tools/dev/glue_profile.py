"""Which Python lines launch the small torch glue kernels of one training step?  (torch.profiler, eager step, stacks aggregated)"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from spe_b200 import factory
from spe_b200.dp import FlatGradBuffer
from spe_b200.engine import TrainStep
dev = torch.device("cuda")
cfg = bench.cfg2()
torch.manual_seed(42)
model = factory.build_detector(cfg, dev).train()
crit = factory.build_criterion(cfg, device=dev).eval(); crit_ref = factory.build_criterion(cfg, refine=True, device=dev).eval()
gbuf = FlatGradBuffer(model.parameters())
step = TrainStep(model, crit, crit_ref, crit.weight_dict, gbuf, graph=False)
images = torch.randn(8, 3, 640, 640, device=dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in bench.synth_targets(8, 7)]
for _ in range(2): step(images, targets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step(images, targets)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_time_total <= 0 or not ev.name.startswith("aten::"):
        continue
    if ev.cpu_children:      # leaf ops only
        continue
    src = "?"
    for fr in (ev.stack or []):
        if "/spe_b200/" in fr or "bench.py" in fr:
            src = fr.split("/spe_b200/")[-1] if "/spe_b200/" in fr else fr
            break
    a = agg[(ev.name, src[:110])]
    a[0] += 1; a[1] += ev.device_time_total
tot = sum(a[1] for a in agg.values())
print("aten leaf ops with device time: %d launches, %.2f ms" % (sum(a[0] for a in agg.values()), tot / 1e3))
for (name, src), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6d x %8.1f us  %-28s %s" % (n, us, name, src))
