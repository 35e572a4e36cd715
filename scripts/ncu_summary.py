"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, stall mix, hot SASS lines.
    python scripts/ncu_summary.py gpurun_out/prof_rows.ncu-rep [n_hot]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_hot = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, v = rows[0], rows[1], rows[-1]
m = dict(zip(h, v))
u = dict(zip(h, units))
print("kernel:", m.get("Kernel Name", "")[:100])
for k in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
          "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_fp64.sum"]:
    if k in m:
        print(f"  {k} = {m[k]} {u[k]}")
st = [(float(m[k]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
      for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
print("  stalls (warps per issue):", ", ".join(f"{n}={x:.2f}" for x, n in sorted(st, reverse=True) if x > 0.1))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data)
print(f"  SASS instructions {len(data)}, warp-instr executed {sum(int(r[iex]) for r in data)}, samples {tot}")
bars = [i for i, r in enumerate(data) if "BAR." in r[isrc] or "SYNCS" in r[isrc]]
print("  barriers/syncs at", bars[:40])
acc = collections.OrderedDict()
for i, r in enumerate(data):
    acc[i // 200] = acc.get(i // 200, 0) + int(r[isamp])
print("  samples per 200 instr:", {k * 200: x for k, x in acc.items()})
for idx, r in sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:n_hot]:
    print(f"  [{idx}] samples={r[isamp]} exec={r[iex]}  {r[isrc][:90]}")
c = collections.Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += int(r[iex])
print("  op mix:", [(k, x) for k, x in c.most_common(16)])
