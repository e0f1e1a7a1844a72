#!/usr/bin/env python
"""Per-function / per-line instruction and stall-sample shares of one profiled kernel.

    python tools/ncu_lines.py gpurun_out/prof_jac.ncu-rep [out.txt]

Uses `ncu --page source --print-source cuda,sass` (needs -lineinfo and --import-source on)
and attributes every SASS instruction to the source line it was generated from, then to
the enclosing C function of that line (inlined code counts for the function it came from).
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def enclosing_functions(path):
    """line number -> function name, from a light scan of a C/CUDA source file."""
    out = {}
    try:
        lines = open(path).read().split("\n")
    except OSError:
        return out
    cur, depth = None, 0
    sig = re.compile(r"^(?:[A-Za-z_][\w\s\*&:<>,]*?)\b([A-Za-z_]\w*)\s*\([^;]*$")
    pending = None
    for i, l in enumerate(lines, 1):
        s = l.strip()
        if depth == 0 and not s.startswith(("#", "//", "/*", "*")):
            m = sig.match(l)
            if m and not s.startswith(("if", "for", "while", "switch", "return", "else")):
                pending = m.group(1)
        if depth == 0 and "{" in l and pending:
            cur = pending
        depth += l.count("{") - l.count("}")
        if cur:
            out[i] = cur
        if depth == 0 and "}" in l:
            cur, pending = None, None
    return out


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur_file, cur_line, hdr = None, None, None
    per_line = defaultdict(lambda: [0.0, 0.0, 0.0, ""])  # inst, thread inst, samples, text
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] not in ("", "..."):
            cur_line = int(r[0])
            per_line[(cur_file, cur_line)][3] = r[1]
            continue
        if not r[2].startswith("0x"):
            continue
        d = dict(zip(hdr[2:], r[2:]))
        try:
            e = per_line[(cur_file, cur_line)]
            e[0] += float(d["Instructions Executed"])
            e[1] += float(d["Thread Instructions Executed"])
            e[2] += float(d["# Samples"])
        except (ValueError, KeyError):
            pass
    tot_i = sum(v[0] for v in per_line.values()) or 1.0
    tot_s = sum(v[2] for v in per_line.values()) or 1.0
    funcs = {}
    per_fn = defaultdict(lambda: [0.0, 0.0, 0.0])
    for (f, ln), v in per_line.items():
        if f not in funcs:
            funcs[f] = enclosing_functions(f)
        fn = "%s:%s" % ((f or "?").split("/")[-1], funcs[f].get(ln, "?"))
        for k in range(3):
            per_fn[fn][k] += v[k]
    out = ["source: %s" % rep, "warp instructions executed: %.0f   stall samples: %.0f" % (tot_i, tot_s), "",
           "%-46s %8s %8s %9s" % ("function (inlined code attributed to its source)", "inst %", "smpl %", "lanes/32")]
    for fn, v in sorted(per_fn.items(), key=lambda kv: -kv[1][0])[:40]:
        out.append("%-46s %8.2f %8.2f %9.1f" % (fn, 100 * v[0] / tot_i, 100 * v[2] / tot_s, v[1] / max(v[0], 1)))
    out += ["", "top source lines:"]
    for (f, ln), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:40]:
        out.append("%6.2f%% inst %6.2f%% smpl %5.1f lanes  %s:%d  %s" % (
            100 * v[0] / tot_i, 100 * v[2] / tot_s, v[1] / max(v[0], 1), (f or "?").split("/")[-1], ln, v[3].strip()[:80]))
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
