"""SASS opcode census per kernel of the built library (cuobjdump -sass): the instructions that prove which hardware path a kernel uses.
  UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG/UTMAREDG = TMA tensor copies, HMMA = mma.sync,
  MOVM = movmatrix, MUFU.EX2 = exp2, SYNCS = mbarrier, UTCBAR = tcgen05.commit.
    python tools/sass_summary.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "spe_b200", "libspe_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MOVM", "MUFU.EX2", "SYNCS", "REDG", "RED.", "ATOMG", "LDGSTS", "STL", "LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEYS:
                if op.startswith(k) or (k.endswith(".") and op.startswith(k[:-1] + ".")):
                    cur[k] += 1
    names = list(per)
    try:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass

    def short(n):
        s = demangle.get(n, n)
        s = s.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        s = re.sub(r"\((int|bool|unsigned int)\)", "", s)
        s = re.sub(r"\(.*", "", s)
        return s[:86]

    rows = []
    for n, c in per.items():
        if any(c[k] for k in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "MOVM")):
            rows.append((short(n), c))
    rows.sort(key=lambda r: r[0])
    cols = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MOVM", "MUFU.EX2", "SYNCS", "STL", "LDL", "_total"]
    md = ["# SASS opcode census of spe_b200/libspe_b200.so (sm_100a), kernels that use tensor cores / TMEM / TMA",
          "", "`cuobjdump -sass` static instruction counts per kernel instantiation: UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st,",
          "UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, HMMA = mma.sync, MOVM = movmatrix, STL/LDL = local-memory spills.", "",
          "| kernel | " + " | ".join(c.replace("_total", "instructions") for c in cols) + " |", "|---|" + "---|" * len(cols)]
    for n, c in rows:
        md.append("| `%s` | " % n + " | ".join(str(c[k]) for k in cols) + " |")
    text = "\n".join(md) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text[:6000])


if __name__ == "__main__":
    main()
