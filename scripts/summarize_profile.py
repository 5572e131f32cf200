"""Turns an ncu report (gpurun_out/*.ncu-rep) into the text summaries committed under profiles/.

usage: python scripts/summarize_profile.py gpurun_out/inflate_r1c.ncu-rep profiles/r1_inflate_v3 "note"
"""
import csv, io, subprocess, sys

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines = [f"# {rep}", f"# {note}", ""]
if len(rows) >= 3:
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else f"kernel {k}"
        lines.append(f"## launch {k}: {name}")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                lines.append(f"{h:90s} {v:>18s} {u}")
        lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg, cur_file, tot_i, tot_s = [], None, 0, 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 9 or r[0] in ("Line No", "Function Name"):
        continue
    if r[0] != "" and r[2] == "-":
        try:
            samp, inst, thr = int(r[4]), int(r[7]), int(r[8])
        except ValueError:
            continue
        agg.append((inst, samp, thr, cur_file, r[0], r[1].strip()[:110]))
        tot_i += inst
        tot_s += samp
if agg:
    lines.append(f"## source hot spots (warp instructions executed: {tot_i}, stall samples: {tot_s})")
    lines.append("  %inst  %samples  threads/inst  file:line  source")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    for inst, samp, thr, f, ln, s in sorted(agg, reverse=True)[:top]:
        lines.append(f"  {100*inst/max(tot_i,1):5.1f}  {100*samp/max(tot_s,1):5.1f}  {thr/max(inst,1):5.1f}  {f}:{ln}  {s}")
open(out + ".txt", "w").write("\n".join(lines) + "\n")
print("wrote", out + ".txt", len(lines), "lines")
