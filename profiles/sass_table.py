"""SASS opcode table per kernel of libviml_b200.so (run here: `python profiles/sass_table.py r2`): counts of the instructions that
prove which hardware path a kernel uses — DMMA (FP64 tensor pipe), UBLKCP (TMA bulk copy), LDL/STL (local memory = spills or
dynamically indexed arrays), ATOMS (shared-memory atomics), RED/ATOMG (global atomics), LDS/STS, DFMA/DMUL/DADD."""
import collections
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
so = os.path.join(root, "tc-viml_b200", "libviml_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ("DMMA", "UBLKCP", "LDL", "STL", "ATOMS", "RED", "ATOMG", "LDS", "STS", "DFMA", "DMUL", "DADD", "MUFU")
tab = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
        cur = tab.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for o in ops:
            if op == o or op.startswith(o + "."):
                cur[o] += 1
with open(os.path.join(here, f"sass_opcodes_{tag}.md"), "w") as f:
    f.write(f"# SASS opcode counts per kernel, round {tag[1:]} (`cuobjdump -sass tc-viml_b200/libviml_b200.so`, sm_100a; static counts)\n\n")
    f.write("| kernel | instr | " + " | ".join(ops) + " |\n|---|---|" + "---|" * len(ops) + "\n")
    for k, c in sorted(tab.items(), key=lambda kv: -kv[1]["total"]):
        f.write(f"| `{k}` | {c['total']} | " + " | ".join(str(c[o]) for o in ops) + " |\n")
print(open(os.path.join(here, f"sass_opcodes_{tag}.md")).read())
