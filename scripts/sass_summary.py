"""Per-kernel tcgen05 / TMEM / TMA instruction counts of the shipped library, from `cuobjdump -sass` (no GPU needed).

    python scripts/sass_summary.py [round tag] > profiles/sass_summary_<tag>.md

SASS mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (kind::f16 and kind::tf32), UTCBAR = tcgen05.commit,
LDTM / STTM = tcgen05.ld / st (TMEM <-> registers), UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA),
UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier operations.
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spectrogramgenai_b200", "lib", "libsgb200.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCATOMSWS", "SYNCS", "MUFU.EX2", "FFMA2", "HMMA", "IMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    cur[c] += 1
    dm = demangle(list(kernels))
    print(f"# SASS summary of libsgb200.so ({tag}) -- `cuobjdump -sass`, sm_100a\n")
    print("`python scripts/sass_summary.py`; counts are static instructions per kernel (template instantiation).  "
          "UTCHMMA = `tcgen05.mma`, UTCBAR = `tcgen05.commit`, LDTM / STTM = `tcgen05.ld` / `st`, UTMALDG / UTMASTG = TMA "
          "(`cp.async.bulk.tensor`) loads / stores, UTCATOMSWS = TMEM alloc / dealloc, SYNCS = mbarrier ops.  "
          "No `HMMA` / `IMMA` (mma.sync) anywhere: every contraction is tcgen05.\n")
    tot = Counter()
    rows = []
    for k, c in kernels.items():
        for col in COLS:
            tot[col] += c[col]
        if c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"]:
            name = re.sub(r"\(.*", "", dm.get(k, k)).replace("void ", "").replace("sg::tc::", "").replace("sg::", "")
            rows.append((name, c))
    print("| kernel | " + " | ".join(COLS[:8]) + " | MUFU.EX2 | FFMA2 | instrs |")
    print("|---|" + "---|" * 11)
    for name, c in rows:
        print(f"| `{name}` | " + " | ".join(str(c[x]) for x in COLS[:8]) + f" | {c['MUFU.EX2']} | {c['FFMA2']} | {c['_total']} |")
    print(f"\nLibrary totals over {len(kernels)} kernels: " + ", ".join(f"{c} {tot[c]}" for c in COLS) + ".")
    print(f"\nKernels without tensor-core / TMA instructions ({len(kernels) - len(rows)}): the memory-bound elementwise, "
          "normalisation, embedding and CUDA-core comparator kernels.")


if __name__ == "__main__":
    main()
