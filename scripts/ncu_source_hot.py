"""Hot-spot digest of an `ncu --page source --csv` export: executed-instruction share per opcode
and per contiguous SASS region (split at large changes of the execution count)."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    try:
        ins.append((r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]] or 0)))
    except Exception:
        pass
tot = sum(e for _, e, _ in ins); tots = sum(s for _, _, s in ins)
print("total warp-instr", tot, "samples", tots, "n_sass", len(ins))
ops = collections.Counter(); ops_s = collections.Counter()
for s, e, sm in ins:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    ops[op] += e; ops_s[op] += sm
print("by opcode (exec share, stall-sample share):")
for op, e in ops.most_common(18):
    print(f"  {op:12s} {100*e/tot:5.1f}%  {100*ops_s[op]/max(tots,1):5.1f}%")
# regions
print("regions (start idx, n instr, exec per instr, share, sample share):")
i = 0
while i < len(ins):
    j = i; e0 = ins[i][1]; acc = 0; sacc = 0
    while j < len(ins) and (0.5 * e0 <= ins[j][1] <= 2 * e0 or ins[j][1] == e0):
        acc += ins[j][1]; sacc += ins[j][2]; j += 1
    if acc > 0.01 * tot or sacc > 0.01 * tots:
        print(f"  [{i:4d}..{j:4d}) n={j-i:4d} exec/instr~{e0:10d} share={100*acc/tot:5.1f}% samples={100*sacc/max(tots,1):5.1f}%  first: {ins[i][0][:50]}")
    i = max(j, i + 1)
