#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): the metrics the roofline numbers come from."""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active']


def main(path):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("---", r[hdr.index("Kernel Name")][:60], "launch id", r[hdr.index("ID")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:70s} {r[i]:>18s} {units[i]}")
        stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                  if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct') and 'not_issued' not in h and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {h.split('stalled_')[1].split('_per_warp')[0]:40s} {v:8.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
