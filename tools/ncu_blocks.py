"""Per-basic-block issue share and lane utilisation from an ncu report's source page.
Usage: python tools/ncu_blocks.py report.ncu-rep [top_n]   (needs -lineinfo / --import-source on for nothing but SASS)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr, data = rows[1], rows[2:]
ia, it, isrc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
ins = [(r[isrc].strip(), int(r[ia]), int(r[it])) for r in data if len(r) > it]
tot = sum(x[1] for x in ins)
print(f"warp instructions executed: {tot}; SASS instructions: {len(ins)}; mean active lanes {sum(x[2] for x in ins) / tot:.2f}")


def opcode(s):
    p = s.split()
    return (p[1] if s.startswith("@") else p[0]).split(".")[0]


blocks, cur = [], None
for i, (s, a, t) in enumerate(ins):
    ends = i > 0 and opcode(ins[i - 1][0]) in ("BRA", "BSYNC", "CALL", "RET", "EXIT", "BREAK")
    if cur and cur["a"] == a and not ends:
        cur["n"] += 1; cur["t"] += t; cur["ops"].append(opcode(s))
    else:
        cur = {"i": i, "a": a, "n": 1, "t": t, "ops": [opcode(s)]}
        blocks.append(cur)
blocks.sort(key=lambda b: -b["a"] * b["n"])
acc = 0.0
print("first-instr  n_instr  executions  lanes  share  cumulative  dominant opcodes")
for b in blocks[:top]:
    share = b["a"] * b["n"] / tot
    acc += share
    c = collections.Counter(b["ops"])
    print(f"{b['i']:6d} {b['n']:4d} {b['a']:11d} {b['t'] / max(1, b['a'] * b['n']):5.1f} {share * 100:6.2f}% {acc * 100:6.1f}%  "
          + " ".join(f"{k}x{v}" for k, v in c.most_common(5)))
