#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture without the GUI.

ncu's CSV source page is SASS-only; this joins it with `nvdisasm --print-line-info` of the object file the kernel was
built from (same build!) and sums executed warp instructions and stall samples per CUDA source line.

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX OBJECT.o [--top N] [--launch K]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, regex, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"],
                         capture_output=True, text=True).stdout
    # the page repeats a ("Kernel Name", name) line + header per matching launch
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = dict(name=row[1], hdr=None, rows=[])
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    if not blocks:
        sys.exit("no kernel matched")
    b = blocks[min(launch, len(blocks) - 1)]
    h = {k: i for i, k in enumerate(b["hdr"])}
    res = []
    for r in b["rows"]:
        res.append(dict(addr=int(r[h["Address"]], 16), sass=r[h["Source"]].strip(), inst=int(r[h["Instructions Executed"]] or 0),
                        stall=int(r[h["Warp Stall Sampling (All Samples)"]] or 0)))
    return b["name"], res


def line_map(obj, mangled_part):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    text = subprocess.run(["nvdisasm", "--print-line-info", cubins[0]], capture_output=True, text=True).stdout
    lines, cur_fn, cur_line, m = [], None, None, []
    for ln in text.splitlines():
        if ln.startswith("//--------------------- .text."):
            cur_fn = ln
            cur_line = None
            continue
        if cur_fn is None or mangled_part not in cur_fn:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur_line = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        im = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if im:
            m.append((int(im.group(1), 16), cur_line, im.group(2).strip()))
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("regex")
    ap.add_argument("obj")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--mangled", default=None, help="substring of the mangled name (default: the regex)")
    a = ap.parse_args()
    name, rows = sass_rows(a.rep, a.regex, a.launch)
    lm = line_map(a.obj, a.mangled or a.regex)
    if len(lm) != len(rows):
        print(f"warning: {len(rows)} profiled SASS instructions vs {len(lm)} in the object (different build?)", file=sys.stderr)
    base = rows[0]["addr"]
    by_off = {off: line for off, line, _ in lm}
    agg = collections.defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for r in rows:
        line = by_off.get(r["addr"] - base)
        agg[line][0] += r["inst"]
        agg[line][1] += r["stall"]
        tot_i += r["inst"]
        tot_s += r["stall"]
    src_cache = {}

    def src(line):
        if line is None:
            return "?"
        f, n = line
        if f not in src_cache:
            for root, _, files in os.walk(os.path.dirname(os.path.abspath(a.obj))):
                if f in files:
                    src_cache[f] = open(os.path.join(root, f)).read().splitlines()
                    break
            else:
                src_cache[f] = []
        t = src_cache[f]
        return t[n - 1].strip()[:110] if 0 < n <= len(t) else ""
    print(f"{name[:100]}\n total warp instructions {tot_i}, stall samples {tot_s}")
    print(f"{'inst%':>6} {'stall%':>6}  line")
    for line, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        tag = f"{line[0]}:{line[1]}" if line else "?"
        print(f"{100.0 * i / max(tot_i, 1):6.2f} {100.0 * s / max(tot_s, 1):6.2f}  {tag:18s} {src(line)}")


if __name__ == "__main__":
    main()
