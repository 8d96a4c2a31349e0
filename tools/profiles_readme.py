#!/usr/bin/env python
"""Regenerate profiles/README.md: one table row per ncu summary (profiles/r*_ncu_*.txt) with the counters
BASELINE.json asks for (time, DRAM GB/s, L2 hit rate, sectors per request, issue activity, thread activity),
and one row per bench line (profiles/r*_bench_*.json)."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def load(path):
    m = {}
    for l in open(path):
        f = [x.strip() for x in l.split("|")]
        if len(f) == 3:
            m[f[0]] = (f[1], f[2])
    return m


def num(m, k, default=None):
    if k not in m:
        return default
    u, v = m[k]
    try:
        return float(v.replace(",", "")) * SCALE.get(u, 1)
    except ValueError:
        return default


def main():
    out = ["# profiles/", "",
           "ncu summaries (`ncu --set full --clock-control none`, one launch of 10 M reads after warm-up launches; "
           "`tools/gpu_profile.sh`, condensed by `tools/ncu_summary.py`) and bench lines (`bench.py`). "
           "Regenerate this file with `python tools/profiles_readme.py`.", "",
           "File name prefixes: `r01_` round 1; `r02_` round 2 up to the first full measurement set (and the round's "
           "experiment records: `r02_scan_experiments.txt`, `r02_hbm_random*`); `r02b_` the set after the staged seed "
           "scan (plus the 2-GPU line, the sanitizer runs and the exhaustive kernel before / after its shared-memory "
           "lists); `r02c_` the round's last full set (every config, ncu of every kernel, launch list); `r02d_` / `r02e_` "
           "single lines taken after the last two filter changes (configs[1] with the slice-addressed shared-memory "
           "filter; the default 3 Gbp line with four filter bits per key).", "",
           "## Kernels", "",
           "| file | kernel | ms | DRAM GB/s (rd+wr) | DRAM bytes/read | L2 hit % | ld sectors/request | issue active % | threads/inst | regs | top stall |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    for f in sorted(glob.glob(os.path.join(P, "r*_ncu_*.txt"))):
        m = load(f)
        t = num(m, "gpu__time_duration.sum")
        if not t:
            continue
        rd, wr = num(m, "dram__bytes_read.sum", 0), num(m, "dram__bytes_write.sum", 0)
        req, sec = num(m, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), num(m, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(v[1]) for k, v in m.items()
                  if "issue_stalled_" in k and "selected" not in k}
        top = max(stalls.items(), key=lambda kv: kv[1]) if stalls else ("", 0)
        kern = m.get("Kernel Name", ("", ""))[1]
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        out.append("| %s | `%s` | %.2f | %.0f | %.0f | %.1f | %s | %.1f | %s | %s | %s %.1f |" % (
            os.path.basename(f), kern, t * 1e3, (rd + wr) / t / 1e9, (rd + wr) / 1e7,
            num(m, "lts__t_sector_hit_rate.pct", 0), "%.1f" % (sec / req) if req else "-",
            num(m, "smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
            m.get("smsp__thread_inst_executed_per_inst_executed.ratio", ("", "-"))[1],
            m.get("launch__registers_per_thread", ("", "-"))[1], top[0], top[1]))
    out += ["", "## Bench lines", "",
            "| file | workload | GPUs | reads/s resident | reads/s end to end | scan ms | extend ms | roofline frac (r01: SURVEY model; r02: necessary bytes, dominant kernel) | CPU reference reads/s |",
            "|---|---|---|---|---|---|---|---|---|"]
    for f in sorted(glob.glob(os.path.join(P, "r*_bench_*.json"))):
        try:
            d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        except Exception:
            continue
        if "config" not in d or "e2e" not in d:
            continue
        if d.get("impl") == "reference":
            out.append("| %s | %s | - | - | - | - | - | - | %.3g (%s, %d cores) |" % (
                os.path.basename(f), d["config"]["workload"][:60], d["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"]))
            continue
        km = d.get("kernel_ms_per_pass") or d.get("kernel_ms", {})
        cpu = d.get("cpu_baseline")
        out.append("| %s | %s | %d | %.3g | %.3g | %.2f | %.2f | %.2f | %s |" % (
            os.path.basename(f), d["config"]["workload"][:60], d["n_gpus"], d["value"], d["e2e"]["value"],
            km.get("pack_scan", 0), km.get("extend", 0), d["roofline"]["frac"],
            ("%.3g" % cpu["value"]) if cpu else "-"))
    cli = sorted(glob.glob(os.path.join(P, "r*_cli_bench_*.json")))
    if cli:
        out += ["", "## Command line on files (`tools/cli_bench.py`)", "",
                "| file | reads | format | wall s | reads/s (wall) | reference reads/s (wall) | outputs equal to the reference's |", "|---|---|---|---|---|---|---|"]
        for f in cli:
            d = json.load(open(f))
            r = d.get("reference")
            out.append("| %s | %d | %s | %.2f | %.3g | %s | %s |" % (
                os.path.basename(f), d["reads"], d["format"], d["wall_s"], d["reads_per_s_wall"],
                ("%.3g" % r["reads_per_s_wall"]) if r else "-", json.dumps(d.get("parity")) if d.get("parity") else "-"))
    open(os.path.join(P, "README.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
