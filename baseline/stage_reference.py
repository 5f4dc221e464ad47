"""Stage the UNMODIFIED reference's hot-path packages (models/, util/) under baseline/_ref/ so that `bench.py --impl reference`
can run the real reference (through oracle/ref_shim.py: a fake `timm`, nothing else) on the GPU box, where /root/reference does
not exist.  baseline/_ref/ is git-ignored (reference sources never enter the history) but not gpurun-ignored (it travels with
the snapshot like the built .so files).

    python baseline/stage_reference.py            # also run by __graft_entry__.build() when /root/reference is present

The reference has no setup.py / pyproject.toml, so the contract's `pip install --target baseline/_ref /root/reference` cannot work
(pip: "does not appear to be a Python project"); a verbatim file copy is the equivalent: STAGED.json lists every file with its
sha256 so a reader can check that nothing was edited."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SPE_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("models", "util")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "models")):
        if verbose:
            print("stage_reference: %s not present, nothing staged" % SRC)
        return False
    files = {}
    for pkg in PACKAGES:
        for d, dirs, names in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = [x for x in dirs if x != "__pycache__"]
            for n in names:
                if not n.endswith(".py"):
                    continue
                src = os.path.join(d, n)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                files[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": files}, open(os.path.join(DST, "STAGED.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print("stage_reference: %d files -> %s" % (len(files), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
