"""Instruction-cache footprint of a kernel from an `ncu --page source --csv` export (SASS level, one row per
instruction with its execution count) joined with the line table of `nvdisasm -g -c <cubin>`.

    footprint.py <src.csv> <kernel.sass> <mangled kernel name> [calls per step of the hottest helper = 3]

Prints (1) how much static code runs how often per warp-step (the step count is taken from the most executed
instructions: a helper that is called N times per step), in instructions and in touched 128-byte lines,
(2) the hot code by source region, (3) the reconvergence points (BSSY / BSYNC) executed in most steps with the
`no instruction` stall samples that landed on them, (4) forward branches that jump over more than 384 bytes,
(5) the dynamic opcode mix.  This is the analysis behind profiles/README.md r01k-r01m: the advection kernel's
launch time follows the number of instruction-cache lines it makes the GPC-level cache serve.
"""
import collections
import csv
import re
import sys

src_csv, sass, kname = sys.argv[1:4]
calls_per_step = int(sys.argv[4]) if len(sys.argv) > 4 else 3

line_of, text_of, labels, order = {}, {}, {}, []
cur, inside = None, False
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
    m = re.match(r"^(\.L_x_\d+|\$\S+):", ln)
    if m:
        labels[m.group(1)] = len(order)
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:
        off = int(m.group(1), 16)
        line_of[off], text_of[off] = cur, m.group(2).strip()
        order.append(off)

rows = list(csv.reader(open(src_csv)))
hdr, base, recs = None, None, []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break  # first kernel of the report only
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or not r or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    recs.append((a - base, int(r[hdr["Instructions Executed"]]), int(r[hdr["stall_no_inst"]]), int(r[hdr["# Samples"]]), r[hdr["Source"]]))

counts = sorted(c for _, c, _, _, _ in recs)
steps = counts[-min(100, len(counts))] / calls_per_step
total = sum(c for _, c, _, _, _ in recs)
print(f"{len(recs)} SASS instructions, {steps:.0f} warp-steps, {total / steps:.0f} warp instructions per step")

bins = [0, 0.001, 0.01, 0.03, 0.1, 0.3, 0.9, 1.5, 1e9]
hist, dyn = collections.Counter(), collections.Counter()
for off, c, _, _, _ in recs:
    f = c / steps
    for i in range(len(bins) - 1):
        if bins[i] <= f < bins[i + 1]:
            hist[i] += 1
            dyn[i] += c
print("\nexecutions per warp-step   static instr     KB   share of executed instr")
for i in range(len(bins) - 1):
    hi = "inf" if bins[i + 1] > 1e8 else f"{bins[i + 1]:g}"
    print(f"  {bins[i]:>6g} .. {hi:<6s}        {hist[i]:6d}  {hist[i] * 16 / 1024:6.1f}   {100 * dyn[i] / total:5.1f} %")
lines = collections.defaultdict(float)
for off, c, _, _, _ in recs:
    lines[off // 128] = max(lines[off // 128], c / steps)
print("\n128-byte lines touched more often than f times per warp-step:")
for th in (0.9, 0.3, 0.1, 0.03, 0.01, 0.001):
    k = sum(v > th for v in lines.values())
    print(f"  f > {th:<6g} {k:5d} lines  {k * 128 / 1024:6.1f} KB")

print("\nhot code (more than 0.3 executions per warp-step) by source file:line, top 25:")
agg = collections.Counter()
for off, c, _, _, _ in recs:
    if c / steps > 0.3:
        agg[line_of.get(off)] += 1
for k, v in agg.most_common(25):
    print(f"  {str(k):40s} {v:5d} instr")

print("\nreconvergence points executed in more than 5 % of the warp-steps (per step, no_inst samples, samples, source):")
for off, c, ni, smp, srcline in recs:
    if ("BSYNC" in srcline or "BSSY" in srcline) and c / steps > 0.05:
        print(f"  {off:#7x} {srcline.strip()[:34]:34s} {c / steps:5.2f} {ni:5d} {smp:5d}  {line_of.get(off)}")

addr_of = {k: (order[v] if v < len(order) else None) for k, v in labels.items()}
print("\nforward branches over more than 384 bytes:")
for off in order:
    m = re.search(r"(BRA\S*)\s.*`\((\.L_x_\d+)\)", text_of[off])
    if m and addr_of.get(m.group(2)) is not None and addr_of[m.group(2)] - off > 384:
        print(f"  {off:#7x} -> {addr_of[m.group(2)]:#7x} {addr_of[m.group(2)] - off:6d} B  {text_of[off][:44]:44s} {line_of.get(off)}")

ops = collections.defaultdict(lambda: [0, 0, 0])
for off, c, ni, smp, srcline in recs:
    t = srcline.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op][0] += c
    ops[op][1] += ni
    ops[op][2] += smp
tot_ni = max(1, sum(o[1] for o in ops.values()))
tot_s = max(1, sum(o[2] for o in ops.values()))
print("\nopcode            per step   share   samples   no_inst samples")
for op, o in sorted(ops.items(), key=lambda kv: -kv[1][0])[:20]:
    print(f"  {op:12s} {o[0] / steps:10.1f}  {100 * o[0] / total:5.1f} %  {100 * o[2] / tot_s:5.1f} %   {100 * o[1] / tot_s * tot_s / tot_ni:5.1f} %")
