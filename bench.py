#!/usr/bin/env python
"""bench.py -- images/sec of the SPE hot path (fwd + criterion incl. Hungarian matcher + bwd [+ grad all-reduce])
on synthetic 3x640x640 images, TSCAM-S24 + 6enc/6dec conditional DETR, 300 queries, 81 logits (BASELINE configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg4] [--batch B]
    (N > 1: launched by torch.distributed.run, one rank per GPU)

A step = model(images) -> criterion(out[0], targets) + criterion_refine(out[1], targets+scores) -> backward of the
weighted loss sum -> (N > 1) the flat gradient buffer all-reduced over NCCL in two buckets, the first one under the backbone's
backward (engine.TrainStep).  No optimizer step, no dataloader (the metric BASELINE.json names).  `--impl reference` times the
UNMODIFIED reference (staged verbatim under baseline/_ref by baseline/stage_reference.py, imported through oracle/ref_shim.py) on the
host CPU cores on the same config -- the oracle port when it is not staged -- and adds an informational eager-on-GPU timing of it.
`--config cfg4` = BASELINE configs[3] (TSCAM-M36, 800x1333, batch 1 per GPU); the default cfg2 is the metric's configuration.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "cfg2: TSCAM-S24(D384,depth24,H8)+6enc/6dec cond-DETR, 300 queries, 81 logits, 3x640x640 synthetic"


def cfg2():
    from types import SimpleNamespace
    return SimpleNamespace(embed_dim=384, depth=24, num_heads=8, img_classes=80, patch=16, layer_to_det=23, depth_token_only=2,
                           mlp_ratio=4.0, pos_grid=(50, 84), det_heads=8, ffn=2048, enc_layers=6, dec_layers=6, num_queries=300,
                           det_classes=81, num_refines=1, ln_eps_backbone=1e-6, ln_eps_detr=1e-5)


def cfg4():
    """BASELINE configs[3]: TSCAM-M36 (D 768, depth 36, 16 heads, tap after block 35), 800x1333 -> 50 x 83 = 4150 tokens."""
    c = cfg2()
    c.embed_dim, c.depth, c.num_heads, c.layer_to_det = 768, 36, 16, 35
    return c


CONFIGS = {
    "cfg2": dict(cfg=cfg2, hw=(640, 640), batch=8, workload=WORKLOAD),
    "cfg4": dict(cfg=cfg4, hw=(800, 1333), batch=1,
                 workload="cfg4: TSCAM-M36(D768,depth36,H16)+6enc/6dec cond-DETR, 300 queries, 81 logits, 3x800x1333 synthetic"),
}


def synth_targets(batch, seed, repeat=5, det_classes=81):
    """per image 1..10 GT boxes, labels 1..80, hung_match_ratio=5 exact repeats."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(batch):
        n = int(torch.randint(1, 11, (1,), generator=g))
        c = torch.rand(n, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.05
        out.append({"labels": torch.randint(1, det_classes, (n,), generator=g).repeat_interleave(repeat),
                    "boxes": torch.cat([c, wh], 1).repeat_interleave(repeat, 0),
                    "scores": (torch.rand(n, generator=g) * 0.8 + 0.1).repeat_interleave(repeat)})
    return out


def cfg5_inputs(batch=256, queries=300, n_gt=1000, classes=81, seed=5):
    """BASELINE configs[4] / SURVEY §8d cfg5: the Hungarian-matcher microbenchmark inputs (300 queries x 1000 GT, batch 256)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(batch, queries, classes, generator=g)
    boxes = torch.cat([torch.rand(batch, queries, 2, generator=g) * 0.8 + 0.1, torch.rand(batch, queries, 2, generator=g) * 0.48 + 0.02], -1)
    targets = []
    for _ in range(batch):
        targets.append({"labels": torch.randint(0, classes, (n_gt,), generator=g),
                        "boxes": torch.cat([torch.rand(n_gt, 2, generator=g) * 0.8 + 0.1, torch.rand(n_gt, 2, generator=g) * 0.48 + 0.02], -1)})
    return logits, boxes, targets


def matcher_microbench(dev, reps=5):
    """GPU leg of cfg5: cost-matrix build + batched assignment for 256 images, device-resident inputs, CUDA events."""
    from spe_b200 import criterion_ops as CO
    logits, boxes, targets = cfg5_inputs()
    lg, bx = logits.to(dev), boxes.to(dev)
    packed = CO.pack_targets(targets, dev)
    w = (2.0, 5.0, 2.0)
    for _ in range(2):
        CO.match(lg, bx, packed, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        CO.match(lg, bx, packed, w)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps / logits.shape[0]
    return {"us_per_img": us, "workload": "cfg5: 300 queries x 1000 GT, batch 256, cost build + assignment on the device (CUDA events, %d reps)" % reps}


def matcher_cpu_us_per_img(n_img=24):
    """reference matcher on the host: models/matcher.py:62-86 per image (cost matrix in torch + scipy LSAP), oracle port."""
    from oracle import spe_oracle as O
    logits, boxes, targets = cfg5_inputs(batch=n_img)
    t0 = time.perf_counter()
    O.hungarian_match(logits, boxes, targets)
    return (time.perf_counter() - t0) * 1e6 / n_img


def fused_talking_heads_microbench(dev, cfg, B, reps=3):
    """The fused talking-heads kernels (csrc/talking_fused.cu; SPE_TH_FUSED=1) at the benchmarked backbone shape, one layer forward +
    backward: per-kernel device time (library profiler, CUDA events on the launch stream) and the saved-activation footprint next to
    the unfused pipeline the timed step uses for training (ops._talking_fused_mode: auto)."""
    from spe_b200 import _lib, ops
    H, D = cfg.num_heads, cfg.embed_dim
    N = (640 // cfg.patch) ** 2
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, N, 3 * D, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    Wl = (torch.eye(H) + 0.1 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    Ww = (torch.eye(H) + 0.1 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    bl = torch.zeros(H, device=dev, requires_grad=True)
    bw = torch.zeros(H, device=dev, requires_grad=True)
    go = torch.randn(B, N, D, generator=g).to(torch.bfloat16).to(dev)
    res = {}
    for name, fn in (("fused", ops.TalkingHeadsFusedFn), ("unfused", ops.TalkingHeadsAttentionFn)):
        for it in range(reps + 1):
            if it == 1:
                torch.cuda.synchronize()
                _lib.prof_enable(True)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            torch.cuda.reset_peak_memory_stats()
            m0 = torch.cuda.memory_allocated()
            out = fn.apply(qkv, Wl, bl, Ww, bw, H)
            saved = torch.cuda.memory_allocated() - m0
            out.backward(go)
            qkv.grad = None
        e1.record()
        torch.cuda.synchronize()
        _lib.prof_enable(False)
        fam = _lib.prof_collect()
        res[name] = {"fwd_bwd_ms_per_layer": e0.elapsed_time(e1) / reps, "saved_activation_bytes_per_layer": int(saved),
                     "kernel_ms_per_layer": {k: v[0] / reps for k, v in fam.items() if v[2]}}
    res["shape"] = "B=%d H=%d N=%d dh=%d (cfg2 backbone block), one layer forward + backward" % (B, H, N, D // H)
    res["default"] = "training step uses the unfused pipeline (auto mode); fused forward is used when no gradient is required"
    return res


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "MEASURED_PEAKS.json (sustained bf16)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU baseline = the reference's algorithm on the host cores (oracle port), bounded sample: bs=1 steps
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_fn():
    from oracle import spe_oracle as O
    cfg = O.CFG2
    params = O.make_params(cfg, 0)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    wd = O.default_weight_dict(cfg)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(1, 3, 640, 640, generator=g)
    targets = synth_targets(1, 0)

    def step():
        for v in p.values():
            v.grad = None
        out = O.model_forward(p, cfg, images)
        ld = O.criterion_forward(out[0], targets, ("labels", "boxes", "cardinality"), gamma=2.0)
        ld2 = O.criterion_forward(out[1], targets, ("labels", "boxes", "cardinality"), gamma=2.0, refine=True)
        loss = O.total_loss(ld, wd) + O.total_loss(ld2, wd)
        loss.backward()
        return float(loss)

    return step


def reference_step_fn(device, batch):
    """The UNMODIFIED reference (models/, util/ staged verbatim under baseline/_ref by baseline/stage_reference.py, imported through
    oracle/ref_shim.py) on the same config, parameters (oracle.make_params(CFG2, 0)), losses and targets as our arm."""
    from oracle import ref_shim, spe_oracle as O
    cfg = O.CFG2
    ref, model = ref_shim.build_reference_model(cfg, O.make_params(cfg, 0))
    model.train().to(device)
    wd = O.default_weight_dict(cfg)
    losses = ("labels", "boxes", "cardinality")
    crit = ref_shim.build_reference_criterion(ref, cfg, wd, losses, gamma=2.0).to(device)
    crit_r = ref_shim.build_reference_criterion(ref, cfg, wd, losses, gamma=2.0, refine=True).to(device)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(batch, 3, 640, 640, generator=g).to(device)
    targets = [{k: v.to(device) for k, v in t.items()} for t in synth_targets(batch, 0)]

    def step():
        model.zero_grad(set_to_none=True)
        out = model(images)
        ld = crit(out[0], targets)
        ld2 = crit_r(out[1], targets)
        loss = sum(ld[k] * wd[k] for k in ld if k in wd) + sum(ld2[k] * wd[k] for k in ld2 if k in wd)
        loss.backward()
        return float(loss)

    return step


def reference_gpu_eager(batch, steps=3, warmup=2):
    """Informational same-box comparator: the reference's own eager PyTorch code on the B200, fp32 storage
    with TF32 matmuls allowed (the most favourable stock setting).  Not the reference arm's value: that is the CPU line."""
    if not torch.cuda.is_available():
        return None
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        b = batch
        while b >= 1:
            try:
                step = reference_step_fn(torch.device("cuda", 0), b)
                for _ in range(warmup):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                return {"value": 1e3 * b / ms, "unit": "images/s", "ms_per_step": ms, "batch": b, "steps": steps,
                        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                        "precision": "fp32 parameters/activations, TF32 tensor-core matmuls allowed, eager PyTorch, scipy LSAP on the host",
                        "note": "informational: the stock reference code on this B200; the arm's value/e2e are the CPU run"}
            except torch.OutOfMemoryError:
                step = None
                torch.cuda.empty_cache()
                b //= 2
        return {"unavailable": "out of memory at batch 1"}
    except Exception as e:                                        # informational leg: never fail the arm
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_shim
    torch.set_num_threads(os.cpu_count() or 1)
    real = ref_shim.available()
    if real:
        step = reference_step_fn(torch.device("cpu"), 1)
        kind, sample = "reference", "the unmodified reference (baseline/_ref via oracle/ref_shim.py): model fwd + SetCriterion + SetCriterionRefine (scipy LSAP) + bwd"
    else:
        step = cpu_reference_step_fn()
        kind, sample = "port", "oracle/spe_oracle.py fwd+criterion(scipy LSAP)+bwd (reference not staged: run baseline/stage_reference.py in the build container)"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "images/sec fwd+bwd", "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_step": "1 image (bounded sample of the bs=8 step)", "losses": "det(out[0]) + refine(out[1])",
                       "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": kind,
                             "sample": sample + ", bs=1 per step, fp32, torch threads=%d" % cores},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if real and args.ref_gpu:
        step = None
        line["reference_gpu_eager"] = reference_gpu_eager(args.batch or 8)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from spe_b200 import _lib, factory, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    conf = CONFIGS[args.config]
    cfg = conf["cfg"]()
    IH, IW = conf["hw"]
    workload = conf["workload"]
    if args.batch is None:
        args.batch = conf["batch"]
    torch.manual_seed(42 + rank)                                   # main.py:161 rank-dependent seed
    model = factory.build_detector(cfg, dev)
    # the reference zero-inits the last bbox layer and LayerScale=1e-5; keep the architecture's own random init
    model.train()
    crit = factory.build_criterion(cfg, ("labels", "boxes", "cardinality"), gamma=2.0, device=dev)
    crit_ref = factory.build_criterion(cfg, ("labels", "boxes", "cardinality"), gamma=2.0, refine=True, device=dev)
    crit.eval(); crit_ref.eval()                                    # deterministic targets: repeats pre-applied
    wd = crit.weight_dict
    # flat gradient buffer: ONE all-reduce per step covers every parameter.  p.grad are slices of it and the backward
    # kernels accumulate straight into them (wgrad GEMM reduce-add / atomics: no temporaries, no per-parameter add kernels).
    from spe_b200.dp import FlatGradBuffer
    from spe_b200.engine import TrainStep
    gbuf = FlatGradBuffer(model.parameters(), mode=os.environ.get("SPE_GRAD_MODE", "views"))
    # the step = engine.TrainStep: zero grads, refresh bf16 weight shadows, model, both criteria (12 Hungarian matchings), weighted
    # sum, backward, grad all-reduce.  Default: captured once into a CUDA graph and replayed (inputs copied into its static buffers
    # every step); --eager runs the same body launch by launch from Python.
    step_graph = TrainStep(model, crit, crit_ref, wd, gbuf, graph=not args.eager, max_gt=64)
    step_eager = TrainStep(model, crit, crit_ref, wd, gbuf, graph=False)

    B = args.batch
    g = torch.Generator().manual_seed(100 + rank)
    host_images = torch.randn(B, 3, IH, IW, generator=g).pin_memory()
    targets_host = synth_targets(B, 7 + rank)
    dev_images = host_images.to(dev)
    dev_targets = [{k: v.to(dev) for k, v in t.items()} for t in targets_host]
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def step(images, targets):
        return step_graph(images, targets)[0]

    def e2e_step():
        # H2D inside the timed region: the pinned host image batch and the host target lists go straight into the step
        loss = step_graph(host_images, targets_host)[0] if not args.eager else step_eager(
            host_images.to(dev, non_blocking=True), [{k: v.to(dev, non_blocking=True) for k, v in t.items()} for t in targets_host])[0]
        if not args.eager:
            step_graph.prefetch(host_images)         # the NEXT step's H2D image copy (39 MB, every step) runs under this step's compute
        loss_host.copy_(loss.reshape(1), non_blocking=True)                               # D2H of the step's result
        return loss

    def timed(fn, n, prof=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if prof:
            _lib.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - l0
        fam = None
        if prof:
            _lib.prof_enable(False)
            fam = _lib.prof_collect()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, fam

    for _ in range(args.warmup):
        step(dev_images, dev_targets)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, _ = timed(lambda: step(dev_images, dev_targets), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        e2e_step()
    ms_e2e, _, _ = timed(e2e_step, args.steps)
    # separate profiled EAGER pass (events around every launch of the main kernel families) -> roofline numbers; also counts the
    # library's launches per step (a graph replay issues the same kernels without passing through the counter)
    # (kernels are serialised for this pass -- the timed step runs the weight-gradient GEMMs next to the data-gradient GEMMs on a second
    #  stream, which would make the per-launch event times of overlapping kernels overlap too)
    ops._OVERLAP = False
    ms_prof, launches_prof, fam = timed(lambda: step_eager(dev_images, dev_targets), max(1, min(args.steps, 3)), prof=True)
    ops._OVERLAP = None
    if not args.eager:
        launches = launches_prof // max(1, min(args.steps, 3)) * args.steps
    nprof = max(1, min(args.steps, 3))

    if rank != 0:
        if world > 1:
            step_graph.close()          # captured NCCL all-reduces must be gone before the group is destroyed
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    imgs = B * world
    value = imgs * args.steps / (ms / 1e3)
    e2e = imgs * args.steps / (ms_e2e / 1e3)
    h2d = host_images.numel() * 4 + sum(v.numel() * v.element_size() for t in targets_host for v in t.values())
    breakdown = {}
    for k, (t, w, n) in fam.items():
        if n:
            breakdown[k] = {"ms_per_step": t / nprof, "launches_per_step": n / nprof, "work_per_step": w / nprof,
                            "share_of_step": (t / nprof) / (ms / args.steps)}       # of the TIMED (graph-replayed) step
    gm = fam["gemm"]
    g_tflops = gm[1] / (gm[0] * 1e-3) / 1e12 if gm[0] > 0 else 0.0
    hbm = {k: (fam[k][1] / (fam[k][0] * 1e-3) / 1e9 if fam[k][0] > 0 else 0.0)
           for k in ("gemm_attention", "talking_softmax_fwd", "talking_softmax_bwd", "softmax", "layernorm", "attention_fused")}
    dominant = max(breakdown, key=lambda k: breakdown[k]["ms_per_step"]) if breakdown else "gemm"
    if dominant == "gemm" or dominant not in hbm:
        roof = {"kernel": "gemm_tcgen05_kernel (all dense/linear GEMM launches of a step)", "bound": "tensor", "achieved": g_tflops, "peak": peaks["tflops"],
                "unit": "TFLOP/s", "frac": g_tflops / peaks["tflops"], "traffic": None}
    else:
        kname = "gemm_tcgen05_kernel (batched attention GEMMs: QK^T / PV and gradients)" if dominant == "gemm_attention" else dominant
        roof = {"kernel": kname, "bound": "hbm", "achieved": hbm[dominant], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": hbm[dominant] / peaks["hbm_gbs"], "traffic": None}
    # traffic: DRAM bytes per launch of that kernel from the committed ncu capture of one step (profiles/rNN_kernel_traffic.json)
    fam_sym = {"talking_softmax_fwd": "talking_fwd", "talking_softmax_bwd": "talking_bwd_rows", "softmax": "softmax_", "layernorm": "layernorm_",
               "attention_fused": "attn_fwd_kernel", "gemm": "gemm_tcgen05_kernel", "gemm_attention": "gemm_tcgen05_kernel"}
    try:
        import glob
        tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_kernel_traffic.json")))[-1]
        kt = json.load(open(tf))["kernels"]
        sel = [v for k, v in kt.items() if fam_sym.get(dominant, "gemm_tcgen05_kernel") in k]
        n = sum(v["launches"] for v in sel)
        roof["traffic"] = sum((v["dram_read_per_launch"] + v["dram_write_per_launch"]) * v["launches"] for v in sel) / max(1, n)
        roof["traffic_source"] = "%s: mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the %d launches of `%s*` in one step%s" % (
            os.path.basename(tf), n, fam_sym.get(dominant, "gemm_tcgen05_kernel"),
            " (dense and batched-attention GEMMs share the kernel symbol)" if dominant.startswith("gemm") else "")
        roof["algorithmic_per_launch"] = breakdown[dominant]["work_per_step"] / max(1.0, breakdown[dominant]["launches_per_step"])
    except Exception:
        pass
    # the fraction north_star asks for: attention-GEMM roofline.  Algorithmic FLOPs = SURVEY section 8(d): QK^T + PV only
    # (L 4N^2D + E 4N^2Dd + P Ld (4Q^2Dd + 6QNDd) per image forward, x3 for forward + backward; recompute inside the fused kernels and
    # the head mixes are NOT counted), over the device time of every kernel that implements attention (profiled pass, serialised).
    N_tok, Dm, Q = (IH // cfg.patch) * (IW // cfg.patch), cfg.embed_dim, cfg.num_queries
    att_fwd = cfg.depth * 4 * N_tok * N_tok * Dm + cfg.enc_layers * 4 * N_tok * N_tok * Dm + \
        (cfg.num_refines + 1) * cfg.dec_layers * (4 * Q * Q * Dm + 6 * Q * N_tok * Dm)
    att_fams = ("gemm_attention", "talking_softmax_fwd", "talking_softmax_bwd", "attention_fused", "softmax")
    att_ms = sum(fam[k][0] for k in att_fams) / nprof
    att_tflops = 3.0 * att_fwd * B / (att_ms * 1e-3) / 1e12 if att_ms > 0 else 0.0
    roof["attention_gemm"] = {"achieved": att_tflops, "unit": "TFLOP/s", "peak": peaks["tflops"], "frac": att_tflops / peaks["tflops"],
                              "algorithmic_flops_per_step": 3.0 * att_fwd * B, "ms_per_step": att_ms,
                              "families": {k: fam[k][0] / nprof for k in att_fams},
                              "note": "QK^T + PV flops only (SURVEY 8d) over all attention kernels incl. softmax / head-mix time"}
    roof["peak_source"] = peaks["src"] + " -- of measured"
    roof["how"] = "CUDA events on the launch stream around every launch of the family during %d profiled steps; achieved = sum(algorithmic work)/sum(time)" % nprof
    roof["gemm_tflops"] = g_tflops
    roof["gemm_frac_of_bf16_peak"] = g_tflops / peaks["tflops"]
    roof["hbm_families_gbs"] = hbm

    matcher = matcher_microbench(dev)
    fused_th = None
    try:
        fused_th = fused_talking_heads_microbench(dev, cfg, B) if args.config == "cfg2" else {"unsupported": "csrc/talking_fused.cu covers H in {4, 8}"}
    except Exception as e:                                        # informational block: never fails the bench line
        fused_th = {"error": str(e)[:200]}
    cpu = None
    if args.cpu_baseline and world == 1 and args.config == "cfg2":     # the CPU legs are cfg2's (a cfg4 CPU step takes minutes)
        torch.set_num_threads(os.cpu_count() or 1)
        from oracle import ref_shim
        real = ref_shim.available()                                # baseline/_ref staged: time the unmodified reference, else the oracle port
        cstep = reference_step_fn(torch.device("cpu"), 1) if real else cpu_reference_step_fn()
        cstep()                                                     # one warm-up step (allocator, thread pool), then two timed
        t0 = time.perf_counter()
        cstep(); cstep()
        dt = (time.perf_counter() - t0) / 2
        cpu = {"value": 1.0 / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "reference" if real else "port",
               "sample": ("the unmodified reference (baseline/_ref)" if real else "oracle (fp32 PyTorch restatement, scipy LSAP)")
                         + ": bs=1 fwd+criteria+bwd steps of the same config, 1 warm-up + 2 timed (%.1f s each)" % dt,
               "matcher_us_per_img": matcher_cpu_us_per_img(), "matcher_sample": "cfg5 shapes, 24 images, torch cost matrix + scipy LSAP per image"}
    line = {"metric": "images/sec fwd+bwd", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": B, "global_batch": imgs, "parallelism": "dp%d" % world,
                       "losses": "det(out[0]) + refine(out[1]), 12 Hungarian matchings/step, hung_match_ratio 5",
                       "l2": "working set per step >> 126 MB L2 (activations of one step: tens of GB), no explicit flush",
                       "weights": "random init (architecture default)",
                       "step": ("engine.TrainStep, whole step captured in a CUDA graph and replayed" if not args.eager else "engine.TrainStep, eager launches")},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "input_pipeline": "pinned host batch copied every step; the copy of step i+1 is issued on a copy stream while step i runs (TrainStep.prefetch, %d staged batches used)" % step_graph.__dict__.get("_prefetch_hits", 0)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernel_breakdown": breakdown, "matcher": matcher,
            "talking_heads_fused": fused_th, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        step_graph.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: 8 for cfg2, 1 for cfg4)")
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="cfg2 = the configuration BASELINE.json's metric is quoted on")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-ref-gpu", dest="ref_gpu", action="store_false", help="reference arm: skip the informational eager-on-GPU timing")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
