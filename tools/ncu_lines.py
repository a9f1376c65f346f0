#!/usr/bin/env python
"""Per-source-line profile of one kernel from an .ncu-rep captured with --import-source on.

ncu's CSV source page is per SASS instruction; the kernels here are one big inlined body from
rc_pipeline.h, so this joins the SASS rows (in address order) with `nvdisasm -g` line markers
of the same cubin (extracted from tochris_b200/libref_cuda.so, which must be the build that was
profiled) and sums samples / executed instructions per source line.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_draw [--top 40] [--launch N]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel, mangled=None):
    tmp = tempfile.mkdtemp(prefix="cub_")
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "tochris_b200", "libref_cuda.so")], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if "sm_100" in f][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines, cur, infn = [], ("?", 0), False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            infn = (mangled in ln) if mangled else re.search(r"\b%s\b|_Z\d+%s" % (kernel, kernel), ln) is not None
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append((cur, ln.split("*/", 1)[1].strip().rstrip(";")))
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0, help="which launch of the kernel in the report")
    ap.add_argument("--mangled", default=None, help="substring of the mangled name, to pick one template instance (e.g. k_drawILb0E)")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", a.kernel], capture_output=True, text=True).stdout
    # the page repeats per launch: split on the "Kernel Name" header rows
    chunks = re.split(r'(?m)^"Kernel Name",', txt)[1:]
    rows = list(csv.reader(io.StringIO(chunks[a.launch].split("\n", 1)[1])))
    hdr = rows[0]
    ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    body = [r for r in rows[1:] if len(r) > ci["Thread Instructions Executed"]]
    sl = sass_lines(a.kernel, a.mangled)
    if len(sl) != len(body):
        print(f"warning: {len(body)} SASS rows in the report vs {len(sl)} in the current cubin (different build?)")
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for i, r in enumerate(body):
        key = sl[i][0] if i < len(sl) else ("?", 0)
        s, ie, te = int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0)
        g = agg[key]
        g[0] += s; g[1] += ie; g[2] += te; g[3] += 1
        tot[0] += s; tot[1] += ie; tot[2] += te
    print(f"{a.kernel}: {len(body)} SASS instrs, {tot[1]} warp-instrs executed, {tot[0]} samples, avg active threads {tot[2] / max(tot[1], 1):.1f}")
    src = {}
    for (fn, ln), g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        if fn not in src:
            p = os.path.join(ROOT, "tochris_b200", "csrc", fn)
            src[fn] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[fn][ln - 1].strip() if 0 < ln <= len(src[fn]) else ""
        print(f"{100 * g[0] / max(tot[0], 1):5.1f}% smp {100 * g[1] / max(tot[1], 1):5.1f}% inst  sass {g[3]:4d}  thr {g[2] / max(g[1], 1):4.1f}  {fn}:{ln:<5d} {text[:110]}")


if __name__ == "__main__":
    main()
