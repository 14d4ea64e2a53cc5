#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel duration/DRAM/issue metrics and the hottest source lines.
usage: python scripts/ncu_lines.py <rep> [kernel-substring] [topN]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; ksub = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if ksub and ksub not in name: continue
    print("==", name[:80], "id", r[idx["ID"]])
    for w in want:
        if w in idx: print("   %-70s %s %s" % (w, r[idx[w]], rows[1][idx[w]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + (["-k", "regex:" + ksub] if ksub else []),
                     capture_output=True, text=True).stdout
cur = None; h = None; out = []; kern = None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": kern = r[1]; continue
    if r and r[0] == "Line No": h = r; continue
    if h and len(r) == len(h) and r[0] != "":
        try: out.append((kern, cur, int(r[0]), r[1], int(r[6]), int(r[7])))
        except ValueError: pass
bykern = collections.defaultdict(list)
for o in out: bykern[o[0]].append(o)
for kern, lst in bykern.items():
    tot = sum(o[5] for o in lst) or 1; tots = sum(o[4] for o in lst) or 1
    print("\n#### %s  inst=%d samples=%d" % (kern[:90], tot, tots))
    for o in sorted(lst, key=lambda o: -o[4])[:topn]:
        print("%-14s L%-4d inst %9d %4.1f%% samp %6d %4.1f%% | %s" % (o[1][-14:], o[2], o[5], 100 * o[5] / tot, o[4], 100 * o[4] / tots, o[3].strip()[:95]))
