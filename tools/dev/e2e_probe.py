"""Where does the e2e step lose time against the device-resident step?  Times variants of the per-step host work."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
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
step = TrainStep(model, crit, crit_ref, crit.weight_dict, gbuf, graph=True, max_gt=64)
host_images = torch.randn(8, 3, 640, 640).pin_memory()
targets_host = bench.synth_targets(8, 7)
dev_images = host_images.to(dev); dev_targets = [{k: v.to(dev) for k, v in t.items()} for t in targets_host]
loss_host = torch.zeros(1).pin_memory()
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
print("device images + device targets      : %.3f ms (wall %.3f)" % timed(lambda: step(dev_images, dev_targets)))
print("device images + HOST targets        : %.3f ms (wall %.3f)" % timed(lambda: step(dev_images, targets_host)))
print("HOST images + device targets        : %.3f ms (wall %.3f)" % timed(lambda: step(host_images, dev_targets)))
print("HOST images + host targets          : %.3f ms (wall %.3f)" % timed(lambda: step(host_images, targets_host)))
def pf():
    l = step(host_images, targets_host)[0]; step.prefetch(host_images); loss_host.copy_(l.reshape(1), non_blocking=True)
print("HOST images prefetched + host tgts  : %.3f ms (wall %.3f)  hits %d" % (*timed(pf), step._prefetch_hits))
