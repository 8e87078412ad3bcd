#!/usr/bin/env python
"""
Join an ncu report's per-SASS-instruction stall samples with nvdisasm line info.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep miosqp_b200/libbqp.so 'admm_tile_kernelILi8' [top]

Prints the share of warp-stall samples per source line (innermost line and the kernel-level call
site when the instruction was inlined), so hot spots can be read without the ncu GUI.
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    sass = ""
    for cub in glob.glob(os.path.join(tmp, "*.cubin")):
        sass += subprocess.run(["nvdisasm", "-gi", "-c", cub], stdout=subprocess.PIPE, text=True).stdout
    # offset -> (inner line, outer line) for the kernel matching pat
    line_of = {}
    in_k = False
    cur = (("", 0), ("", 0))
    for ln in sass.splitlines():
        if ln.startswith(".text."):
            in_k = pat in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)(.*)', ln)
        if m:
            inner = (os.path.basename(m.group(1)), int(m.group(2)))
            outers = re.findall(r'File "([^"]*)", line (\d+)', m.group(3))
            cur = (inner, (os.path.basename(outers[-1][0]), int(outers[-1][1])) if outers else inner)
            continue
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    base = None
    inner_s, outer_s, op_s = collections.Counter(), collections.Counter(), collections.Counter()
    total = 0
    for r in rows:
        if len(r) > 2 and r[0] == "Address":
            hdr = r
            si = hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        s = int(r[si] or 0)
        total += s
        (inner, outer), txt = line_of.get(addr - base, ((("", 0), ("", 0)), r[1]))
        inner_s[inner] += s
        outer_s[outer] += s
        op_s[txt.split()[0] if not txt.startswith("@") else txt.split()[1]] += s
    csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "miosqp_b200", "csrc")
    srcs = {}

    def text(fl):
        f, line = fl
        if f not in srcs:
            try:
                srcs[f] = open(os.path.join(csrc, f)).read().splitlines()
            except OSError:
                srcs[f] = []
        return srcs[f][line - 1].strip()[:100] if 0 < line <= len(srcs[f]) else ""
    print("total samples", total)
    for name, ctr in (("kernel-level line (call site)", outer_s), ("innermost line", inner_s)):
        print("== by", name)
        for fl, s in ctr.most_common(top):
            print("%6.2f%%  %s:%-4d %s" % (100.0 * s / max(total, 1), fl[0], fl[1], text(fl)))
    print("== by opcode")
    for op, s in op_s.most_common(15):
        print("%6.2f%%  %s" % (100.0 * s / max(total, 1), op))


if __name__ == "__main__":
    main()
