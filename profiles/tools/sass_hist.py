"""Opcode histogram of a kernel from `cuobjdump -sass <object>`: python profiles/tools/sass_hist.py loam_livox_b200/csrc/knn.o knn_blocks_kernel > out.txt"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n")[0].strip()
    if pat not in name:
        continue
    ops = collections.Counter()
    for ins in re.findall(r"/\*[0-9a-f]+\*/\s+([^;/]+);", part):
        tok = ins.split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        ops[op.split(".")[0]] += 1
    tot = sum(ops.values())
    print(f"{name}: {tot} SASS instructions")
    print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
    print("  TMA / mbarrier opcodes: " + (", ".join(f"{k} {v}" for k, v in ops.items() if k in ("UBLKCP", "SYNCS", "UTMALDG", "FENCE")) or "none"))
