#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture: joins the SASS page of the report
(ncu -i rep --page source --csv) with the line table of the kernel's cubin (nvdisasm -g) by instruction order
and sums warp instructions, thread instructions and stall samples per source line.
  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [TOP]
Measurement aid (the build container has no GPU; reports come back from gpurun)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bgreat_b200", "lib", "libbgreat_b200.so")


def line_table(kernel):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=d, stdout=subprocess.DEVNULL)
    out = []
    for f in os.listdir(d):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        cur, inside, line = None, False, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", l)
            if m:
                inside = kernel in m.group(1)
                cur = m.group(1)
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                out.append((int(m.group(1), 16), line, m.group(2).strip()))
        if out:
            break
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    tab = line_table(kernel)
    if len(tab) != len(body):
        sys.stderr.write("warning: %d SASS rows in the report, %d in the cubin\n" % (len(body), len(tab)))
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for r, (_, line, _) in zip(body, tab):
        wi = int(r[col["Instructions Executed"]] or 0)
        ti = int(r[col["Thread Instructions Executed"]] or 0)
        ss = int(r[col["# Samples"]] or 0)
        a = agg[line]
        a[0] += wi
        a[1] += ti
        a[2] += ss
        a[3] += 1
        tot[0] += wi
        tot[1] += ti
        tot[2] += ss
    src = {}
    print("total: %d warp inst, %d thread inst, %d samples" % tuple(tot))
    if "--by-function" in sys.argv:
        # the same sums per enclosing function (the nearest preceding definition in the source file)
        fagg = defaultdict(lambda: [0, 0, 0])
        cache = {}

        def func_of(line):
            if not line:
                return "?"
            fn = os.path.join(ROOT, "bgreat_b200", "csrc", line[0])
            if fn not in cache:
                cache[fn] = open(fn).read().splitlines() if os.path.exists(fn) else []
            L = cache[fn]
            for i in range(min(line[1], len(L)) - 1, -1, -1):
                m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:BGR_HD|__device__ __forceinline__|__global__)\b.*?\b(\w+)\s*\(", L[i])
                if m:
                    return "%s:%s" % (line[0], m.group(1))
                m = re.match(r"\s*// ---- (.*)", L[i])
                if m and line[0] == "bgr_emit.cuh":
                    return "%s:[%s]" % (line[0], m.group(1)[:40])
            return line[0]
        for line, a in agg.items():
            f = fagg[func_of(line)]
            f[0] += a[0]
            f[1] += a[1]
            f[2] += a[2]
        print("%-64s %9s %6s %8s %6s %5s" % ("function", "warp-inst", "%", "samples", "%", "thr"))
        for f, a in sorted(fagg.items(), key=lambda kv: -kv[1][0])[:top]:
            print("%-64s %8.1fM %5.1f%% %8d %5.1f%% %5.1f" % (f, a[0] / 1e6, 100.0 * a[0] / max(tot[0], 1), a[2],
                                                              100.0 * a[2] / max(tot[2], 1), a[1] / max(a[0], 1)))
        return
    print("%-26s %8s %6s %8s %6s %5s  %s" % ("line", "warp-inst", "%", "samples", "%", "thr", "source"))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if line:
            fn = os.path.join(ROOT, "bgreat_b200", "csrc", line[0])
            if fn not in src and os.path.exists(fn):
                src[fn] = open(fn).read().splitlines()
            if fn in src and line[1] <= len(src[fn]):
                text = src[fn][line[1] - 1].strip()[:90]
        print("%-26s %8.1fM %5.1f%% %8d %5.1f%% %5.1f  %s" % ("%s:%d" % line if line else "?", a[0] / 1e6, 100.0 * a[0] / max(tot[0], 1),
                                                            a[2], 100.0 * a[2] / max(tot[2], 1), a[1] / max(a[0], 1), text))


if __name__ == "__main__":
    main()
