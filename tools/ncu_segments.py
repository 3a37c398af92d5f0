"""Summarise an ncu report of one kernel: headline counters, then instructions / stall samples / shared-memory
wavefronts per barrier-delimited code segment (ncu --import-source on, read here without a GPU).

    python tools/ncu_segments.py gpurun_out/x.ncu-rep [--kernel REGEX [--skip N]] [--list FIRST LAST]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
# optional kernel filter for reports that hold several kernels: --kernel REGEX (first matching launch)
# (--skip N: the N+1-th matching launch)
sel = ["-k", "regex:" + sys.argv[sys.argv.index("--kernel") + 1], "-c", "1"] if "--kernel" in sys.argv else []
if "--skip" in sys.argv:
    sel += ["--launch-skip", sys.argv[sys.argv.index("--skip") + 1]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", *sel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]
for h, v in zip(hdr, vals):
    if h in want:
        print(f"{h} = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *sel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
print(rows[0][1][:140])
ix = {h: i for i, h in enumerate(rows[1])}
data = [r for r in rows[2:] if len(r) >= len(rows[1]) and r[0].startswith("0x")]  # SASS view only
seen = set()
for k, r in enumerate(data):  # some ncu versions print the listing twice
    if r[0] in seen:
        data = data[:k]
        break
    seen.add(r[0])
num = lambda r, k: int(r[ix[k]] or 0)  # noqa: E731
tot = sum(num(r, "Instructions Executed") for r in data)
tots = sum(num(r, "# Samples") for r in data)
cur, start = [0, 0, 0, 0], 0
for i, r in enumerate(data):
    cur[0] += num(r, "Instructions Executed"); cur[1] += num(r, "# Samples")
    cur[2] += num(r, "L1 Wavefronts Shared"); cur[3] += num(r, "L1 Wavefronts Shared Ideal")
    if "BAR.SYNC" in r[ix["Source"]] or i == len(data) - 1:
        print(f"seg {start:5d}-{i:5d}: inst {cur[0] / 1e6:8.1f}M ({100 * cur[0] / tot:5.1f}%)  samples {cur[1]:6d} "
              f"({100 * cur[1] / tots:5.1f}%)  smem wavefronts {cur[2] / 1e6:7.1f}M (ideal {cur[3] / 1e6:7.1f}M)")
        cur, start = [0, 0, 0, 0], i + 1
if "--list" in sys.argv:
    a, b = int(sys.argv[sys.argv.index("--list") + 1]), int(sys.argv[sys.argv.index("--list") + 2])
    for i in range(a, min(b + 1, len(data))):
        r = data[i]
        print(f"{i:5d} {num(r, 'Instructions Executed') / 1e6:7.2f}M s={num(r, '# Samples'):5d} "
              f"wf={num(r, 'L1 Wavefronts Shared') / 1e6:6.2f}/{num(r, 'L1 Wavefronts Shared Ideal') / 1e6:6.2f} "
              f"thr={r[ix['Avg. Threads Executed']]:>3s}  {r[ix['Source']].strip()[:100]}")
