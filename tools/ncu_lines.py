"""Per-source-line summary of an ncu report's source page (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [file-substring] [top-N]
Prints, for each CUDA source line of the matching file, the warp instructions executed and the stall samples."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file, hdr, rows = None, None, []
for rec in csv.reader(out.splitlines()):
    if not rec:
        continue
    if rec[0] == "File Path":
        cur_file = rec[1]; continue
    if rec[0] == "Line No":
        hdr = rec; continue
    if rec[0] == "Function Name" or hdr is None:
        continue
    if rec[0] != "" and rec[0].isdigit() and len(rec) == len(hdr):
        d = dict(zip(hdr[4:], rec[4:]))
        try:
            int(d["Instructions Executed"]); int(d["# Samples"])
        except ValueError:
            continue
        rows.append((cur_file, int(rec[0]), rec[1], d))
tot_inst = sum(int(d["Instructions Executed"]) for f, l, s, d in rows)
tot_samp = sum(int(d["# Samples"]) for f, l, s, d in rows)
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
sel = [(f, l, s, d) for f, l, s, d in rows if want in f]
sel.sort(key=lambda r: -int(r[3]["# Samples"]))
keys = ["stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_mio", "stall_math", "stall_no_inst", "stall_lg", "stall_dispatch", "stall_not_selected", "stall_selected"]
for f, l, s, d in sel[:top]:
    st = " ".join(f"{k[6:]}={d[k]}" for k in keys if k in d and d[k] not in ("0", "-"))
    print(f"{f.split('/')[-1]}:{l:4d} inst={int(d['Instructions Executed'])/tot_inst*100:5.1f}% samp={int(d['# Samples'])/tot_samp*100:5.1f}% | {s.strip()[:70]} | {st}")
