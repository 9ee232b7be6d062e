"""SASS instruction histogram of the hot kernels in libptb200.so (cuobjdump -sass): opcode counts, registers / spills from
ptxas.log, and the mnemonics that prove the bulk-copy staging (UBLKCP + SYNCS = cp.async.bulk + mbarrier) and the 256-bit
node loads (LDG.E.ENL2.256).  usage: python tools/sass_hist.py > profiles/sass_hist_rNN.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oclpathtracer_b200", "libptb200.so")
HOT = [  # (mangled fragment, what)
    ("k_path_sm2ILb0ELi11ELi6", "k_path_sm2<11 CTAs/SM, 6 visits per vote>   C5: 2M-triangle scene, quantised binary nodes from L2/HBM, path state parked in shared memory"),
    ("k_path_smILi0ELb0ELi8ELi6", "k_path_sm<LARGE, 8 CTAs/SM, 6 visits per vote>   the registers-only form of the same (tune[12] = 1)"),
    ("k_mega_path_regenILb1ELi0ELb0ELb0", "k_mega_path_regen<BVH, LARGE>   the while-while form of the same (tune[5] = 2)"),
    ("k_mega_path_regenILb1ELi2ELb0ELb1", "k_mega_path_regen<BVH, FLAT, COOP>    C4: Cornell box, flat leaf boxes, pooled triangle phase"),
    ("k_megaILi1ELb1ELi2ELb0", "k_mega<AO, BVH, FLAT>           C2"),
    ("k_megaILi2ELb1ELi1ELb0", "k_mega<DIRECT, BVH, SMALL4>     C3: Cornell box, 4-wide tree"),
    ("k_megaILi0ELb1ELi2ELb0", "k_mega<PRIMARY, BVH, FLAT>      C1"),
    ("k_resolve", "k_resolve"),
    ("k_to_rgb8", "k_to_rgb8"),
]
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
ptxas = open(os.path.join(ROOT, "oclpathtracer_b200", "csrc", "ptxas.log")).read() if os.path.exists(os.path.join(ROOT, "oclpathtracer_b200", "csrc", "ptxas.log")) else ""
print("SASS histogram of libptb200.so (sm_100a), hot kernels\n")
for frag, what in HOT:
    names = [n for n in funcs if frag in n]
    if not names:
        print(f"== {what}: not found\n")
        continue
    name = names[0]
    ops = collections.Counter()
    for ln in funcs[name]:
        t = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op.rstrip(";")] += 1
    total = sum(ops.values())
    fam = collections.Counter()
    for op, n in ops.items():
        fam[op.split(".")[0]] += n
    regs = re.search(re.escape(name) + r"'.*?\n.*?\n\s*(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)\n.*?Used (\d+) registers", ptxas, re.S)
    print(f"== {what}\n   {name}\n   {total} SASS instructions" + (f", {regs.group(2)} registers, {regs.group(1)}" if regs else ""))
    print("   " + "  ".join(f"{k}x{v}" for k, v in fam.most_common(24)))
    proof = {k: v for k, v in ops.items() if any(s in k for s in ("UBLKCP", "SYNCS", "LDG.E.ENL2.256", "LDS.128", "LDL", "STL", "REDUX", "ATOM", "RED."))}
    print("   notable: " + ", ".join(f"{k}x{v}" for k, v in sorted(proof.items())) + "\n")
