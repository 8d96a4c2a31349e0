#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the text summary kept under profiles/.

usage: ncu_summary.py REPORT.ncu-rep OUT.txt [--traffic-key NAME --traffic-json profiles/roofline_traffic.json]

Writes one `metric | unit | value` line per selected metric of the first profiled launch and, when asked,
records dram__bytes_read.sum + dram__bytes_write.sum of that launch in the JSON bench.py reads for
`roofline.traffic`."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum(\.per_second)?|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__time_duration\.sum|gpu__dram_throughput.*|l1tex__t_(requests|sectors)_pipe_lsu_mem_global_op_(ld|st)\.sum|"
    r"l1tex__t_sector_hit_rate\.pct|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|launch__(occupancy_limit_.*|registers_per_thread|grid_size|block_size|waves_per_multiprocessor|shared_mem_per_block_static)|lts__t_sector_hit_rate\.pct|lts__t_sectors\.sum|"
    r"lts__t_sectors_srcunit_tex_op_(read|write)\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__cycles_elapsed\.max|sm__inst_executed_pipe_(alu|fma|lsu|xu|adu|cbu|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio|"
    r"sm__maximum_warps_per_active_cycle_pct|smsp__cycles_active\.avg)$")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines, m = [], {}
    for h, u, v in zip(hdr, units, vals):
        if h in ("Kernel Name", "Block Size", "Grid Size"):
            lines.append("%s | | %s" % (h, v))
        if KEEP.match(h):
            lines.append("%s | %s | %s" % (h, u, v))
            m[h] = (u, v)
    open(out, "w").write("\n".join(lines) + "\n")
    if "--traffic-key" in sys.argv:
        key = sys.argv[sys.argv.index("--traffic-key") + 1]
        path = sys.argv[sys.argv.index("--traffic-json") + 1]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u, v = m[k]
            tot += float(v.replace(",", "")) * scale[u]
        try:
            d = json.load(open(path))
        except Exception:
            d = {}
        d[key] = tot
        d[key + "_source"] = rep.split("/")[-1]
        json.dump(d, open(path, "w"), indent=1)
        print(key, tot)


if __name__ == "__main__":
    main()
