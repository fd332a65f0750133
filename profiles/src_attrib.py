#!/usr/bin/env python
"""Attributes the per-instruction counters of an ncu `--page source --csv` export (SASS view) to CUDA source lines.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-skip S --launch-count 1 > src.csv
    cuobjdump -xelf all libtaxidispatch.so ; nvdisasm -g td_pool.sm_100a.cubin > pool.sass
    python profiles/src_attrib.py src.csv pool.sass <mangled-kernel-substring> [top]

nvdisasm -g prints '//## File "...", line N' markers between the instructions of a function; the k-th instruction of
the function in the listing is the k-th row of the ncu export (same order), so the two are joined by position."""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines_of = []
inside = False
cur = None
inl = ""
for ln in open(sass, errors="replace"):
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        inside = kern in ln
        cur = None
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        inl = m.group(3)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines_of.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if r and r[0] == "Address")
body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
print("sass instructions: listing %d, ncu %d" % (len(lines_of), len(body)))
n = min(len(lines_of), len(body))
inst = defaultdict(int)
smp = defaultdict(int)
tot = stot = 0
for k in range(n):
    ie = int(body[k][ix["Instructions Executed"]] or 0)
    sp = int(body[k][ix["# Samples"]] or 0)
    inst[lines_of[k]] += ie
    smp[lines_of[k]] += sp
    tot += ie
    stot += sp
src = open("/root/repo/taxidispatcher_b200/csrc/td_pool.cu").read().split("\n") if "pool" in sass or True else []
print("total warp instructions %d, samples %d" % (tot, stot))
for ln, v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    fn, no = ln if ln else ("?", 0)
    text = src[no - 1].strip()[:100] if fn.endswith("td_pool.cu") and 0 < no <= len(src) else fn
    print("%6.2f%% inst  %6.2f%% samples  L%-5s %s" % (100.0 * v / tot, 100.0 * smp[ln] / max(stot, 1), no, text))
