"""Host enqueue time vs device time of one bench step (is the step launch-bound anywhere?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spe_b200 import factory, _lib
from spe_b200.dp import FlatGradBuffer

dev = torch.device("cuda")
cfg = bench.cfg2()
torch.manual_seed(42)
model = factory.build_detector(cfg, dev).train()
crit = factory.build_criterion(cfg, device=dev).eval()
crit_ref = factory.build_criterion(cfg, refine=True, device=dev).eval()
wd = crit.weight_dict
buf = FlatGradBuffer(model.parameters())
images = torch.randn(8, 3, 640, 640, device=dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in bench.synth_targets(8, 7)]

def step():
    buf.zero_()
    t0 = time.perf_counter()
    out = model(images)
    t1 = time.perf_counter()
    ld, ld2 = crit(out[0], targets), crit_ref(out[1], targets)
    loss = sum(ld[k] * wd[k] for k in ld if k in wd) + sum(ld2[k] * wd[k] for k in ld2 if k in wd)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2

for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    w0 = time.perf_counter(); e0.record()
    f, c, b = step()
    w1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    w2 = time.perf_counter()
    print("host enqueue: fwd %.1f ms  criterion %.1f ms  backward %.1f ms  total %.1f ms | device %.1f ms | wall to sync %.1f ms | spe launches %d"
          % (f * 1e3, c * 1e3, b * 1e3, (w1 - w0) * 1e3, e0.elapsed_time(e1), (w2 - w0) * 1e3, _lib.launch_count() - l0))

if os.environ.get("SPE_CPROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(2):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(45)
