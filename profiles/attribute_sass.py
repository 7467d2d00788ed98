"""Join an `ncu --page source --csv` export (SASS level, per-instruction counters) with the line table
of `nvdisasm -g -c <cubin>` so that executed instructions / stall samples can be summed per CUDA
source line.  Usage: attribute_sass.py <src.csv> <kernel.sass> <mangled kernel name> [top N]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# address offset -> (file, line) from nvdisasm
line_of = {}
cur = None
inside = False
for ln in open(sass):
    if ln.startswith(".text." + kname + ":"):
        inside = True
        continue
    if inside and ln.startswith("//---------------------"):
        break
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

rows = list(csv.reader(open(src_csv)))
# first kernel block only
hdr = None
base = None
agg = defaultdict(lambda: [0, 0, 0, 0])  # inst, thread inst, samples, no_inst
tot = [0, 0, 0, 0]
nblocks = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nblocks += 1
        if nblocks > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or not r or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    off = addr - base
    key, _ = line_of.get(off, (None, ""))
    inst = int(r[hdr["Instructions Executed"]])
    tinst = int(r[hdr["Thread Instructions Executed"]])
    samp = int(r[hdr["# Samples"]])
    noi = int(r[hdr["stall_no_inst"]]) if "stall_no_inst" in hdr else 0
    for k, v in enumerate((inst, tinst, samp, noi)):
        agg[key][k] += v
        tot[k] += v
print(f"total warp inst {tot[0]:,}  thread inst {tot[1]:,}  samples {tot[2]:,}  no_inst {tot[3]:,}  static SASS {len(line_of)}")
print(f"{'file:line':32s} {'warp inst':>12s} {'%':>6s} {'lanes':>6s} {'samples%':>8s}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    name = f"{key[0]}:{key[1]}" if key else "?"
    print(f"{name:32s} {v[0]:12,d} {100*v[0]/tot[0]:6.2f} {v[1]/max(v[0],1):6.1f} {100*v[2]/max(tot[2],1):8.2f}")
