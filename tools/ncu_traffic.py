"""Extracts per-launch DRAM traffic and pipe utilisation of named kernels from ncu reports into profiles/r02_traffic.json.
usage: python tools/ncu_traffic.py out.json  label=report.ncu-rep[:launch_index] ...
bench.py reads `dram_bytes` (dram__bytes_read.sum + dram__bytes_write.sum of ONE launch) from that file for roofline.traffic."""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_tensor.sum",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]

UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}

out = {}
for spec in sys.argv[2:]:
    label, rest = spec.split("=", 1)
    rep, _, idx = rest.partition(":")
    idx = int(idx) if idx else 0
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    rec = dict(zip(hdr, data[idx]))
    un = dict(zip(hdr, units))
    e = {"source": f"profiles/{rep.split('/')[-1]} (ncu --set full --clock-control none), launch {idx}: {rec.get('Kernel Name', '')[:80]}"}
    for k in KEYS:
        if k in rec and rec[k] not in ("", "n/a"):
            v = float(rec[k].replace(",", ""))
            e[k] = v * UNIT_SCALE.get(un.get(k, ""), 1)
    if "dram__bytes_read.sum" in e:
        e["dram_bytes"] = e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]
    out[label] = e
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
