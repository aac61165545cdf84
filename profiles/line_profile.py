#!/usr/bin/env python3
"""Join an ncu SASS source page with nvdisasm line info: per-CUDA-line executed instructions.

usage: line_profile.py <report.ncu-rep> <kernel mangled-name prefix> [top]
(run where the .so that produced the report is still built: it is disassembled for line info)"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, prefix = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
tmp = tempfile.mkdtemp()
so = os.path.join(ROOT, "mchap_b200", "_lib", "libmchap_b200.so")
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
st = [i for i, l in enumerate(dis) if l.startswith(".text." + prefix)][0]
en = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and i > st]
en = en[0] if en else len(dis)
seq, cur = [], ("?", 0)
for l in dis[st:en]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s", l):
        seq.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]
ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > ci and r[0] != "Address"]
if len(data) > len(seq) and len(data) % len(seq) == 0:
    data = data[: len(seq)]  # several launches of the kernel in the report: take the first
assert len(data) == len(seq), (len(data), len(seq))
agg, samp = collections.Counter(), collections.Counter()
for key, r in zip(seq, data):
    agg[key] += int(r[ci])
    samp[key] += int(r[si])
tot, stot = sum(agg.values()), max(sum(samp.values()), 1)
print("total warp instructions: %d" % tot)
files = {}
for (f, ln), c in agg.most_common(top):
    if f not in files:
        p = os.path.join(ROOT, "mchap_b200", "csrc", f)
        files[f] = open(p).read().split("\n") if os.path.exists(p) else None
    t = files[f]
    txt = t[ln - 1].strip()[:80] if t and ln - 1 < len(t) else ""
    print("%5.2f%% instr %5.2f%% stall  %s:%d  %s" % (100 * c / tot, 100 * samp[(f, ln)] / stot, f, ln, txt))
