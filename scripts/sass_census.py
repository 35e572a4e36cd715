"""Opcode census of the kernels in libocean_b200.so (cuobjdump -sass), one line per kernel:
    python scripts/sass_census.py [path/to/lib.so] > profiles/r02_sass_census.txt
Shows which Blackwell mechanisms each kernel really contains (UBLKCP = cp.async.bulk / TMA bulk copy,
SYNCS = mbarrier, SHFL = warp shuffle, MUFU = special-function unit, ...)."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "gfx_ocean_b200", "libocean_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n   # noqa: E731
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "BAR", "SHFL", "LDG.E.128", "LDG.E.64", "LDG", "STG.E.128", "STG", "LDS.128", "LDS", "STS",
         "LDL", "STL", "FFMA", "FADD", "FMUL", "MUFU", "ATOM", "CCTL", "NANOSLEEP", "ACQBULK", "FENCE"]
kern, counts = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (("." in w) and op.startswith(w)):
                counts[kern][w] += 1
print(f"# opcode census of {os.path.basename(lib)} (static SASS instruction counts per kernel; cuobjdump -sass)")
for k, c in sorted(counts.items(), key=lambda kv: demangle(kv[0])):
    name = demangle(k).replace("void ocean::", "").replace("ocean::", "")
    m = re.match(r"(\w+<[^>]*>|\w+)", name)
    name = (m.group(1) if m else name).replace("(int)", "").replace("(bool)", "")
    print(f"{name:60s} total={c['total']:5d} " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
