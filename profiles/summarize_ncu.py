"""Turns the ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
usage: python profiles/summarize_ncu.py launches <csv>   |   raw <ncu-rep> [kernel-regex]"""
import collections
import csv
import re
import subprocess
import sys

KEYS = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__shared_mem_per_block|"
                  r"launch__grid_size|launch__block_size|launch__occupancy_limit_(registers|shared_mem|warps)|"
                  r"l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|sm__inst_executed\.sum|smsp__inst_executed\.sum|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st\.sum|smsp__warps_eligible\.avg\.per_cycle_active|"
                  r"sm__cycles_elapsed\.avg|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < 5:
            continue
        name = r[4].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1].replace(",", ""))
    total = sum(v[1] for v in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:10.1f} {100 * v[1] / total:6.1f}%")


def raw(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ki]):
            continue
        print("==", r[ki][:100], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if KEYS.match(h):
                print(f"   {h:78s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
