"""Join an ncu source-page export (per SASS instruction: executed count + stall samples) with nvdisasm's
line info, and aggregate by innermost source line / by source function, so the hot spots of the step kernel can
be read without a GUI.

  python tools/ncu_lines.py gpurun_out/r1_prof.ncu-rep alphazero-pybind11_b200/libb2az.so k_stepILi8 [--top 40]

Runs here (no GPU): needs ncu, cuobjdump, nvdisasm on PATH. The .so must be the build the profile was taken on.
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


def sass_lines(so, kernel_pat):
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
    out = {}
    inside = False
    cur = []
    pend = []
    for line in txt.splitlines():
        if line.startswith("\t.section") or line.startswith("//-----"):
            inside = ".text." in line and kernel_pat in line
            cur, pend = [], []
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            pend.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"^(\$[\w$]+):\s*$", line)
        if m:  # an internal helper (division slow path, 64-bit remainder ...): no line info of its own
            cur, pend = [(m.group(1), 0)], []
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", line)
        if m:
            if pend:
                cur = pend  # innermost first, outermost last
                pend = []
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def ncu_rows(rep, kernel_pat):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    # several kernels (launches) may be in the report; take the first whose name matches
    blocks = re.split(r'(?m)^"Kernel Name",', txt)
    for b in blocks[1:]:
        name, rest = b.split("\n", 1)
        rd = csv.reader(io.StringIO(rest))
        hdr = next(rd)
        rows = [r for r in rd if len(r) == len(hdr)]
        return name, hdr, rows
    raise SystemExit("no kernel in report")


def load_funcs(src_dir):
    """line -> enclosing function name, by a crude scan of the headers (good enough for aggregation)."""
    table = {}
    for f in os.listdir(src_dir):
        p = os.path.join(src_dir, f)
        cur = "?"
        lines = open(p, errors="ignore").read().splitlines()
        m = {}
        for i, l in enumerate(lines, 1):
            g = re.match(r"^(?:template.*>\s*)?(?:AZ_HD|AZ_D|__global__|static|inline|int|void|u32|float|bool)[^;]*?\b(\w+)\s*\([^;]*$", l)
            if g and not l.startswith(" "):
                cur = g.group(1)
            m[i] = cur
        table[f] = m
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("so")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    sl = sass_lines(a.so, a.kernel)
    name, hdr, rows = ncu_rows(a.rep, a.kernel)
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = min(int(r[ia], 16) for r in rows)
    funcs = load_funcs(os.path.join(os.path.dirname(os.path.abspath(a.so)), "csrc"))
    by_line = collections.Counter()
    by_line_s = collections.Counter()
    by_func = collections.Counter()
    by_func_s = collections.Counter()
    by_op = collections.Counter()
    tot_e = tot_s = 0
    for r in rows:
        off = int(r[ia], 16) - base
        e, s = int(r[ie] or 0), int(r[isamp] or 0)
        tot_e += e
        tot_s += s
        stack, op = sl.get(off, ([], "?"))
        key = stack[0] if stack else ("?", 0)
        by_line[key] += e
        by_line_s[key] += s
        fn = funcs.get(key[0], {}).get(key[1], "?")
        by_func[(key[0], fn)] += e
        by_func_s[(key[0], fn)] += s
        by_op[op.split()[0] if not op.startswith("@") else op.split()[1]] += e
    print(f"kernel {name.strip()}  SASS instrs {len(rows)}  executed {tot_e}  samples {tot_s}")
    print("\n== by function (executed %, stall-sample %)")
    for k, e in by_func.most_common(a.top):
        print(f"{100 * e / tot_e:6.2f}%  {100 * by_func_s[k] / max(1, tot_s):6.2f}%  {k[0]}:{k[1]}")
    print("\n== by stall samples, function")
    for k, s in by_func_s.most_common(15):
        print(f"{100 * s / max(1, tot_s):6.2f}%  {k[0]}:{k[1]}")
    print("\n== by line (executed %, stall-sample %)")
    for k, e in by_line.most_common(a.top):
        print(f"{100 * e / tot_e:6.2f}%  {100 * by_line_s[k] / max(1, tot_s):6.2f}%  {k[0]}:{k[1]}")
    print("\n== by stall samples, line")
    for k, s in by_line_s.most_common(a.top):
        print(f"{100 * s / max(1, tot_s):6.2f}%  {100 * by_line[k] / tot_e:6.2f}%  {k[0]}:{k[1]}")
    print("\n== by opcode")
    for k, e in by_op.most_common(25):
        print(f"{100 * e / tot_e:6.2f}%  {k}")


if __name__ == "__main__":
    main()
