"""Turns an ncu report into small CSV files (run on the GPU box so that only the summaries travel back):
  <out>_raw.csv      selected metrics of every captured launch, one column per launch
  <out>_stalls.txt   per-source-line instruction / stall-sample summary of the kernels whose name matches [regex ...]
usage: python tools/ncu_export.py report.ncu-rep out_prefix [kernel-regex ...]"""
import csv, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
pats = sys.argv[3:]
KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct']
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, data = rows[0], rows[1], rows[2:]
stall = [k for k in h if 'average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio')]
with open(out + "_raw.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
    for k in KEYS + stall:
        if k in h:
            i = h.index(k)
            w.writerow([k, u[i]] + [r[i][:100] for r in data])
lines = []
for pat in pats:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{pat}", "-c", "1"],
                         capture_output=True, text=True).stdout
    cur, hdr, mode, recs = None, None, None, []
    for rec in csv.reader(src.splitlines()):
        if not rec:
            continue
        if rec[0] == "File Path":
            cur = rec[1]; continue
        if rec[0] == "Line No":
            hdr, mode = rec, 'cuda'; continue
        if rec[0] == "Address":
            hdr, mode = rec, 'sass'; continue
        if rec[0] == "Function Name" or hdr is None:
            continue
        if mode == 'cuda' and rec[0].isdigit() and len(rec) == len(hdr):
            d = dict(zip(hdr[4:], rec[4:]))
            try:
                recs.append((cur, int(rec[0]), rec[1], int(d["Instructions Executed"]), int(d["# Samples"]), d))
            except ValueError:
                pass
    ti, ts = sum(r[3] for r in recs) or 1, sum(r[4] for r in recs) or 1
    lines.append(f"== {pat}: {ti} warp instructions with line info, {ts} stall samples; top lines by samples")
    keys = ["stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_mio", "stall_math", "stall_no_inst", "stall_lg",
            "stall_dispatch", "stall_not_selected", "stall_selected", "stall_sleep", "stall_membar", "stall_branch_resolving"]
    for f_, l, s, ni, ns, d in sorted(recs, key=lambda r: -r[4])[:25]:
        st = " ".join(f"{k[6:]}={d[k]}" for k in keys if k in d and d[k] not in ("0", "-"))
        lines.append(f"{f_.split('/')[-1][:16]}:{l:4d} inst={100 * ni / ti:5.1f}% samp={100 * ns / ts:5.1f}% | {s.strip()[:64]} | {st}")
open(out + "_stalls.txt", "w").write("\n".join(lines) + "\n")
print(f"{rep}: {len(data)} launches -> {out}_raw.csv, {out}_stalls.txt")
