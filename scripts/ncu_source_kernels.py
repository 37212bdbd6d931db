"""Per-kernel digest of an `ncu --page source --csv` export that holds several kernels: opcode mix
(executed warp-instructions, stall samples) and the hottest SASS lines of each kernel.
    python scripts/ncu_source_kernels.py file.src.csv [kernel-substring] [top-n]"""
import collections
import csv
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
kernels, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "ins": []}
        kernels.append(cur)
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if cur is None or hdr is None or len(r) < len(hdr):
        continue
    try:
        cur["ins"].append((r[hdr["Source"]].strip(), int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]] or 0),
                           {k: int(r[hdr[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
    except Exception:
        pass
for kd in kernels:
    if want not in kd["name"]:
        continue
    ins = kd["ins"]
    tot = sum(e for _, e, _, _ in ins) or 1
    tots = sum(s for _, _, s, _ in ins) or 1
    print(f"## {kd['name'][:100]}\n   warp-instr {tot:.4g}  samples {tots}  sass lines {len(ins)}")
    ops, ops_s = collections.Counter(), collections.Counter()
    stalls = collections.Counter()
    for s, e, sm, st in ins:
        parts = s.split()
        op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "LDG", "STG", "STS", "VIADDMNMX", "VIMNMX")) else op.split(".")[0]
        ops[op] += e
        ops_s[op] += sm
        for k, v in st.items():
            stalls[k] += v
    print("   opcode: exec share / sample share")
    for op, e in ops.most_common(16):
        print(f"     {op:18s} {100 * e / tot:5.1f}%  {100 * ops_s[op] / tots:5.1f}%")
    print("   stall reasons:", ", ".join(f"{k[6:]}={100 * v / tots:.0f}%" for k, v in stalls.most_common(7)))
    print(f"   hottest lines by samples:")
    for s, e, sm, st in sorted(ins, key=lambda t: -t[2])[:topn]:
        top = max(st, key=st.get) if st else ""
        print(f"     {100 * sm / tots:5.1f}%  exec {e:11d}  {top[6:]:12s} {s[:70]}")
