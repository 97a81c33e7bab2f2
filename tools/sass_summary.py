"""Blackwell evidence from the built library (no GPU needed): per kernel, how often the SASS mnemonics that
prove tcgen05 / TMEM / TMA appear (B200_PROFILING.md: `tcgen05.mma` -> UTC*MMA, `tcgen05.ld/st` -> LDTM / STTM,
`cp.async.bulk.tensor` -> UTMALDG / UTMASTG, `cp.async.bulk` -> UBLKCP, `tcgen05.commit` -> UTCBAR), plus the
first MMA issue block of each tcgen05 kernel as an excerpt.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "fastmax_experiments_b200", "libfastmax_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMUL2", "HFMA2", "FFMA", "HMMA"]
kern, cur, lines = collections.OrderedDict(), None, {}
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern[cur] = collections.Counter()
        lines[cur] = []
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
    if cur and m:
        op = m.group(1)
        lines[cur].append(ln.split("/*", 2)[1].split("*/", 1)[1].split(";")[0].strip())
        for k in KEYS:
            if op.startswith(k):
                kern[cur][k] += 1
print(f"# {os.path.relpath(lib, root)}: SASS mnemonic counts per kernel (sm_100a)\n")
print(f"{'kernel':58s} " + " ".join(f"{k:>8s}" for k in KEYS))
for name, c in kern.items():
    print(f"{name[:58]:58s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
print("\n# first tcgen05.mma issue block of each tensor-core kernel (uniform-register operands, back-to-back UTCHMMA)\n")
for name, ls in lines.items():
    idx = next((i for i, l in enumerate(ls) if l.startswith("UTCHMMA")), None)
    if idx is None:
        continue
    print(f"## {name}")
    for l in ls[max(0, idx - 6): idx + 8]:
        print("   ", l)
    print()
