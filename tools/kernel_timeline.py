"""In-situ kernel durations of ONE graph-replayed bench step (torch.profiler / CUPTI activity records, warm caches, real order):
per-kernel-name totals + the idle time between kernels.  Writes a CSV summary to the path in argv[1] (optional)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from spe_b200 import factory
from spe_b200.engine import TrainStep

dev = torch.device("cuda")
cfg = bench.cfg2()
torch.manual_seed(42)
model = factory.build_detector(cfg, dev).train()
crit = factory.build_criterion(cfg, device=dev).eval()
crit_ref = factory.build_criterion(cfg, refine=True, device=dev).eval()
images = torch.randn(8, 3, 640, 640, device=dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in bench.synth_targets(8, 7)]
step = TrainStep(model, crit, crit_ref, graph=os.environ.get("SPE_EAGER") is None, max_gt=64)
for _ in range(4):
    step(images, targets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(images, targets)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0])
busy, gaps, last_end = 0.0, 0.0, None
for e in ev:
    d = e.time_range.end - e.time_range.start
    n = e.name
    if "gemm_tcgen05" in n:
        n = "gemm_tcgen05_kernel" + n[n.index("<"):n.index(">") + 1]
    n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = n.split("(")[0][:70]
    agg[n][0] += 1; agg[n][1] += d
    busy += d
    if last_end is not None and e.time_range.start > last_end:
        gaps += e.time_range.start - last_end
    last_end = max(last_end or 0, e.time_range.end)
span = ev[-1].time_range.end - ev[0].time_range.start
lines = ["span_us,%.1f" % span, "busy_us,%.1f" % busy, "idle_us,%.1f" % gaps, "kernels,%d" % len(ev), "name,launches,total_us,share"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("%s,%d,%.1f,%.4f" % (n.replace(",", ";"), c, t, t / span))
print("\n".join(lines[:60]))
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write("\n".join(lines) + "\n")
