"""Turn the raw captures that tools_gpu_profiles.sh left in gpurun_out/ into the tracked round evidence under profiles/:

    python tools/make_profiles.py [round-tag, default r01]

  <tag>_launches_one_step.csv    every kernel launch of ONE training step: device time + DRAM bytes (ncu, cold cache / serialised)
  <tag>_kernel_traffic.json      per kernel symbol: launches, total us, mean DRAM read / write bytes per launch (bench.py reads
                                 `roofline.traffic` from this file)
  <tag>_ncu_hot_kernels.md       ncu --set full summary of the hot kernels (tools/prof_attn.py) incl. tensor pipe / MUFU / DRAM %
  <tag>_bench_n1.json / <tag>_bench_reference_cpu.json   the default bench line and the reference arm
  README.md                      what each file is + the digest the judge reads
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")


def launches():
    lines = [l for l in open(os.path.join(OUT, "launches.csv")) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {n: i for i, n in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rd:
        d = per.setdefault(r[ix["ID"]], {"name": short(r[ix["Kernel Name"]]), "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u, m = r[ix["Metric Unit"]], r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        else:
            d[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return list(per.values())


def main():
    os.makedirs(PROF, exist_ok=True)
    L = launches()
    with open(os.path.join(PROF, TAG + "_launches_one_step.csv"), "w") as f:
        f.write("# one training step (cfg2, bs 8) under ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n")
        f.write("idx,kernel,grid,block,time_us,dram_read_bytes,dram_write_bytes\n")
        for i, d in enumerate(L):
            f.write('%d,"%s","%s","%s",%.3f,%.0f,%.0f\n' % (i, d["name"][:90], d["grid"], d["block"], d["us"], d.get("dram__bytes_read.sum", 0),
                                                          d.get("dram__bytes_write.sum", 0)))
    agg = collections.OrderedDict()
    tot = sum(d["us"] for d in L)
    for d in L:
        a = agg.setdefault(d["name"][:70], {"launches": 0, "total_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        a["launches"] += 1
        a["total_us"] += d["us"]
        a["dram_read"] += d.get("dram__bytes_read.sum", 0)
        a["dram_write"] += d.get("dram__bytes_write.sum", 0)
    traffic = {"_note": "per kernel symbol over ONE step (ncu, cold cache / serialised): mean DRAM bytes per launch; small outputs that stay in the "
                        "126 MB L2 show ~0 write bytes", "_step_total_us": tot, "_launches": len(L), "kernels": {}}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["total_us"]):
        traffic["kernels"][k] = {"launches": a["launches"], "total_us": round(a["total_us"], 1), "share": round(a["total_us"] / tot, 4),
                                 "mean_us": round(a["total_us"] / a["launches"], 2),
                                 "dram_read_per_launch": round(a["dram_read"] / a["launches"]), "dram_write_per_launch": round(a["dram_write"] / a["launches"])}
    json.dump(traffic, open(os.path.join(PROF, TAG + "_kernel_traffic.json"), "w"), indent=1)

    # ---- ncu --set full summary
    hot_md = []
    raw = os.path.join(OUT, "prof_hot_raw.csv")
    cols = [("gpu__time_duration.sum", "time us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu(MUFU) %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("lts__t_sector_hit_rate.pct", "l2 hit %")]
    def table(raw, md, last_only=False):
        rows = list(csv.reader(open(raw)))
        hdr, units = rows[0], rows[1]
        ix = {n: i for i, n in enumerate(hdr)}
        use = [(c, t) for c, t in cols if c in ix]
        md.append("| kernel | " + " | ".join(t for _, t in use) + " |")
        md.append("|---|" + "---|" * len(use))
        body = rows[2:]
        if last_only:
            seen = collections.OrderedDict()
            for r in body:
                seen[r[ix["Kernel Name"]]] = r
            body = list(seen.values())
        for r in body:
            vals = []
            for c, _ in use:
                v = r[ix[c]]
                try:
                    fv = float(v.replace(",", ""))
                    if c.startswith("dram__bytes"):
                        fv *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(units[ix[c]], 1)
                    if c == "gpu__time_duration.sum":
                        fv *= {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(units[ix[c]], 1)
                    v = ("%.1f" % fv) if abs(fv) < 1e5 else ("%.3g" % fv)
                except ValueError:
                    pass
                vals.append(v)
            md.append("| `%s` | " % short(r[ix["Kernel Name"]])[:60] + " | ".join(vals) + " |")

    if os.path.exists(raw):
        table(raw, hot_md)
    fused_md = []
    raw2 = os.path.join(OUT, "prof_fused_raw.csv")
    if os.path.exists(raw2) and os.path.getsize(raw2) > 100:
        table(raw2, fused_md, last_only=True)
    th8_md = []
    raw3 = os.path.join(OUT, "prof_th8_raw.csv")
    if os.path.exists(raw3) and os.path.getsize(raw3) > 100:
        table(raw3, th8_md, last_only=True)
    with open(os.path.join(PROF, TAG + "_ncu_hot_kernels.md"), "w") as f:
        f.write("# %s -- ncu `--set full --clock-control none` on the hot kernels (tools/prof_attn.py: encoder self-attention + conditional\n"
                "# cross-attention fwd (fused) + bwd, one talking-heads attention fwd+bwd, one LayerScale FFN fwd+bwd, cfg2 shapes B=8, N=1600, D=384, H=8).\n"
                "# Per launch; cold-cache, serialised (compare shares, not absolutes).\n\n" % TAG)
        f.write("\n".join(hot_md) + "\n")
        if fused_md:
            f.write("\n## kernels added in round 2 (tools/prof_fused.py): fused talking-heads forward / recomputing backward (csrc/talking_fused.cu, cfg2 shape\n"
                    "## B=8, H=8, N=1600, dh=48) and the H=16 mix/softmax/mix kernels (csrc/talking_h16.cu, cfg4 shape B=1, N=4150, fp16 logits); last launch of each\n\n")
            f.write("\n".join(fused_md) + "\n")
        if th8_md:
            f.write("\n## the H = 8 mix / softmax / mix kernels of the training step (csrc/talking_h8.cu; captured after the table at the top, which still\n"
                    "## shows the rowwise.cu kernels they replaced)\n\n" + "\n".join(th8_md) + "\n")
        rep = os.path.join(OUT, "prof_attn_fused.ncu-rep")
        if os.path.exists(rep):
            try:
                txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), rep], capture_output=True, text=True, timeout=300).stdout
                f.write("\n## attn_fwd_kernel (fused attention forward, encoder shape: 832 CTAs), key metrics + warp stall reasons per issue\n\n```\n" + txt + "```\n")
            except Exception as e:                   # ncu missing: keep the table
                f.write("\n(ncu_stalls failed: %s)\n" % e)

    for src, dst in (("bench_default.json", TAG + "_bench_n1.json"), ("bench_reference.json", TAG + "_bench_reference_cpu.json"),
                     ("bench_cfg4.json", TAG + "_bench_cfg4_n1.json"), ("bench_n2_ov1.json", TAG + "_bench_n2.json"),
                     ("bench_n2_ov0.json", TAG + "_bench_n2_single_allreduce.json"), ("bench_n4_ov1.json", TAG + "_bench_n4.json"),
                     ("bench_n4_ov0.json", TAG + "_bench_n4_single_allreduce.json"), ("bench_n8.json", TAG + "_bench_n8.json")):
        p = os.path.join(OUT, src)
        if os.path.exists(p) and os.path.getsize(p) > 10:
            shutil.copy(p, os.path.join(PROF, dst))
    try:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py"), os.path.join(PROF, TAG + "_sass_opcodes.md")], capture_output=True, timeout=600)
    except Exception as e:
        print("sass summary failed", e)
    write_readme(traffic, tot, len(L))
    print("launches", len(L), "total us %.0f" % tot)
    for k, a in list(traffic["kernels"].items())[:12]:
        print("%-62s n=%4d %8.1f us %5.1f%%" % (k[:62], a["launches"], a["total_us"], 100 * a["share"]))


def write_readme(traffic, tot, n):
    b = None
    p = os.path.join(PROF, TAG + "_bench_n1.json")
    if os.path.exists(p):
        try:
            b = json.loads(open(p).read().strip().splitlines()[-1])
        except Exception:
            b = None
    with open(os.path.join(PROF, "README.md"), "w") as f:
        f.write("# profiles -- round %s\n\nAll captured on a B200 through `gpurun` (`./tools_gpu_profiles.sh`, then `python tools/make_profiles.py %s` here).\n\n" % (TAG[1:], TAG))
        f.write("| file | what | command |\n|---|---|---|\n")
        f.write("| `%s_bench_n1.json` | default `python bench.py` line (N=1, cfg2, bs 8, CUDA-graph step) | `python bench.py` |\n" % TAG)
        f.write("| `%s_bench_reference_cpu.json` | reference arm: the UNMODIFIED reference (baseline/_ref) on the host cores + informational eager-on-B200 timing of it | `python bench.py --impl reference --steps 2 --warmup 1` |\n" % TAG)
        f.write("| `%s_launches_one_step.csv` | every kernel launch of ONE training step: device time, DRAM read / write bytes | `ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python tools/step_once.py` |\n" % TAG)
        f.write("| `%s_kernel_traffic.json` | the launch list aggregated per kernel symbol (share of the step, mean DRAM bytes per launch); `bench.py` takes `roofline.traffic` from here | `python tools/make_profiles.py` |\n" % TAG)
        f.write("| `%s_ncu_hot_kernels.md` | `ncu --set full` of the hot kernels (fused attention fwd, attention / dense GEMMs, talking-heads kernels, softmax bwd, LayerNorm) + stall reasons of the fused attention kernel | `ncu --set full --clock-control none -k regex:... python tools/prof_attn.py 2` |\n" % TAG)
        f.write("| `%s_bench_cfg4_n1.json` | BASELINE configs[3] (TSCAM-M36, 800x1333, bs 1) | `python bench.py --config cfg4 --steps 5 --warmup 3` |\n" % TAG)
        f.write("| `%s_bench_n2.json`, `%s_bench_n2_single_allreduce.json` | N=2 data parallel: bucketed all-reduce overlapped with the backbone backward (default) vs one all-reduce after the step (`SPE_AR_OVERLAP=0`) | `torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 10` |\n" % (TAG, TAG))
        f.write("| `%s_bench_n4.json`, `%s_bench_n4_single_allreduce.json` | the same at N=4 | `torchrun --nproc-per-node 4 bench.py --gpus 4 --steps 20 --warmup 5` |\n" % (TAG, TAG))
        f.write("| `%s_bench_n8.json` | N=8 (default bucketed all-reduce) | `torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3` |\n" % TAG)
        f.write("| `%s_sass_opcodes.md` | static SASS census per kernel: UTCHMMA / LDTM / UTMALDG / HMMA / MOVM counts, spills | `python tools/sass_summary.py` |\n" % TAG)
        f.write("| `%s_ncu_qk_gemm.md` | `ncu --set full` of one K = 48 batched QK^T GEMM launch: metrics + stall reasons (DESIGN section 8 item 1) | `ncu --set full --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 python tools/dev/qk_once.py` |\n" % TAG)
        f.write("| `r01_*` | round-1 evidence (kept for comparison) | |\n\n")
        if b:
            f.write("## bench line (CUDA events, not under a profiler)\n\n")
            f.write("* value **%.1f images/s** (%.2f ms/step of %d images), e2e (pinned-host H2D + loss D2H inside the timed region) %.1f images/s\n" % (
                b["value"], b["ms_per_step"], b["config"]["batch_per_gpu"], b["e2e"]["value"]))
            f.write("* clocks under load: %s\n" % json.dumps(b.get("clocks")))
            f.write("* matcher microbenchmark (cfg5, 300 x 1000, batch 256): **%.1f us/image** on the GPU" % b["matcher"]["us_per_img"])
            if b.get("cpu_baseline") and b["cpu_baseline"].get("matcher_us_per_img"):
                f.write(" vs %.0f us/image for the reference matcher (torch cost + scipy) on the host" % b["cpu_baseline"]["matcher_us_per_img"])
            f.write("\n* roofline object: %s\n" % json.dumps({k: b["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic") if k in b["roofline"]}))
            f.write("* kernel families (CUDA events around every launch of the family, eager profiled steps; share = of the timed graph step):\n\n")
            f.write("| family | ms / step | launches / step | share |\n|---|---|---|---|\n")
            for k, v in sorted(b["kernel_breakdown"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
                f.write("| %s | %.2f | %.0f | %.3f |\n" % (k, v["ms_per_step"], v["launches_per_step"], v["share_of_step"]))
        f.write("\n## launch list of one step (ncu, cold-cache / serialised: compare SHARES)\n\n%d launches, %.1f ms summed device time.\n\n" % (n, tot / 1e3))
        f.write("| kernel | launches | total us | share | mean DRAM rd MB | mean DRAM wr MB |\n|---|---|---|---|---|---|\n")
        for k, a in list(traffic["kernels"].items())[:26]:
            f.write("| `%s` | %d | %.0f | %.1f%% | %.1f | %.1f |\n" % (k[:78], a["launches"], a["total_us"], 100 * a["share"], a["dram_read_per_launch"] / 1e6,
                                                                   a["dram_write_per_launch"] / 1e6))
        notes = os.path.join(PROF, TAG + "_notes.md")
        if os.path.exists(notes):
            f.write("\n" + open(notes).read())


if __name__ == "__main__":
    main()
