"""Which torch (non-spe) kernels launch in one step, grouped by the aten op and input shapes (torch.profiler)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from spe_b200 import factory
from spe_b200.dp import FlatGradBuffer

dev = torch.device("cuda")
cfg = bench.cfg2()
torch.manual_seed(42)
B = 8
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::") and e.self_device_time_total > 0:
        st = [s for s in (e.stack or []) if "/repo/" in s and "torch_ops.py" not in s]
        k = (e.name, str(e.input_shapes)[:70], st[0][-60:] if st else "")
        agg[k][0] += 1; agg[k][1] += e.self_device_time_total
tot = sum(v[1] for v in agg.values())
print("torch op device time total %.1f us, %d op calls" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
    print("%5d %9.1f us  %-22s %-70s %s" % (v[0], v[1], k[0], k[1], k[2]))
