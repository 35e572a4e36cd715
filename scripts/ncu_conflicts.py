"""Per-instruction shared-memory bank conflicts of one ncu capture:
    python scripts/ncu_conflicts.py capture.ncu-rep   (groups consecutive instructions with excessive wavefronts)"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
tot = 0
groups = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(h):
        continue
    exc = int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    wf = int(r[ix["L1 Wavefronts Shared"]] or 0)
    if exc:
        tot += exc
        op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
        if groups and groups[-1][1] >= n - 6 and groups[-1][2] == op:
            g = groups[-1]; g[1] = n; g[3] += exc; g[4] += wf; g[5] += 1
        else:
            groups.append([n, n, op, exc, wf, 1, r[1].strip()])
print("total excessive wavefronts", tot)
for g in sorted(groups, key=lambda g: -g[3])[:14]:
    print(f"  sass[{g[0]}..{g[1]}] {g[2]:10s} x{g[5]:3d}  excessive={g[3]:8d} of {g[4]:8d} wavefronts   first: {g[6]}")
