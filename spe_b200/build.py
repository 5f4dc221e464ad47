"""Builds spe_b200/libspe_b200.so (sm_100a only) with nvcc, in-tree.  `python -m spe_b200.build`."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SO = os.path.join(HERE, "libspe_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# per-file extra flags: the matcher cost must not contract a*b+c into FMAs (SURVEY.md H2 iii)
EXTRA = {"matcher.cu": ["-fmad=false"]}
SOURCES = ["api.cu", "gemm_tcgen05.cu", "matcher.cu", "criterion.cu", "rowwise.cu", "talking_generic.cu", "attn_fused.cu", "talking_fused.cu", "talking_h16.cu", "talking_h8.cu", "optim.cu", "targets.cu", "cam_boxes.cu"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "spe_b200.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((s, o, [NVCC] + ARCH + COMMON + EXTRA.get(src, []) + ["-c", s, "-o", o]))

    def run(job):
        s, o, cmd = job
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        with open(o + ".log", "w") as f:
            f.write(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, r.stdout))
        if verbose:
            print(r.stdout)
        return o

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(SO, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", SO] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return SO


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
