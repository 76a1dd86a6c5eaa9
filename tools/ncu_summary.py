"""Print the headline metrics and the hottest source lines of an .ncu-rep (one kernel)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
want += [c for c in h if 'issue_stalled' in c and c.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    for w in want:
        if w in h:
            print(f"{w:90s} {r[h.index(w)]} {rows[1][h.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
items, fname, hdr = [], "", None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        cs, ci = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and r and r[0].isdigit():
        try:
            items.append((float(r[cs] or 0), float(r[ci] or 0), fname, int(r[0]), r[1].strip()[:130]))
        except ValueError:
            pass
tot_s = sum(i[0] for i in items); tot_i = sum(i[1] for i in items)
print(f"total samples {tot_s:.0f}  total warp-instructions {tot_i:.0f}")
for s, i, f, l, x in sorted(items, reverse=True)[:top]:
    print(f"{100 * s / tot_s:5.1f}% smp {100 * i / max(tot_i, 1):5.1f}% ins  {f}:{l:<4d} {x}")
