"""Eager vs CUDA-graph TrainStep on the bench workload (cfg2, bs 8, 640x640): device ms per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
for graph in (False, True):
    step = TrainStep(model, crit, crit_ref, graph=graph, max_gt=64)
    for _ in range(3):
        loss = step(images, targets)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(6):
        loss = step(images, targets)[0]
    e1.record(); torch.cuda.synchronize()
    print("graph=%s: %.2f ms/step (device), wall %.2f ms/step, loss %.5f, mem %.1f GB" % (graph, e0.elapsed_time(e1) / 6, (time.perf_counter() - t0) / 6e-3,
                                                                            float(loss), torch.cuda.max_memory_allocated() / 2**30))
