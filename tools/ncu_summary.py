#!/usr/bin/env python
"""Summarise an .ncu-rep: headline raw metrics and the per-instruction stall picture.
usage: python tools/ncu_summary.py report.ncu-rep [ninstr]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__cycles_active.avg', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active']
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            print("  %-70s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; data = []
for r in rows:
    if r and r[0] == "Address": h = r
    elif r and r[0].startswith("0x") and h: data.append(r)
    elif r and r[0] == "Kernel Name" and data: break
H = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: sum(int(r[H[s]]) for r in data) for s in stalls}
total = sum(int(r[H["# Samples"]]) for r in data)
print("total samples", total, " instructions", len(data))
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]:
    print("  %-24s %8d %5.1f%%" % (s, v, 100 * v / total))
print("top instructions:")
for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][H["# Samples"]]))[:ntop]:
    top = sorted(stalls, key=lambda s: -int(r[H[s]]))[:2]
    print("  %4d %-58s %7s %s" % (i, r[H["Source"]].strip()[:58], r[H["# Samples"]], [(s[6:], r[H[s]]) for s in top]))
