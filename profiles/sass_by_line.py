"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info: samples / instructions per source line.
usage: python profiles/sass_by_line.py <src.csv from ncu> <nvdisasm -g -c output> <kernel mangled-name substring>"""
import csv, re, sys, collections
srccsv, dis, kname = sys.argv[1:4]
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
# parse disassembly: sequence of (file,line) for each instruction in the kernel's .text section
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
cur = ("?", 0); seq = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        if seq: break
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        seq.append(cur)
print("sass in ncu:", len(data), "sass in disasm:", len(seq))
agg = collections.defaultdict(lambda: [0, 0])
n = min(len(data), len(seq))
tot = 0
for r, loc in zip(data[:n], seq[:n]):
    s = int(r[ci["# Samples"]] or 0); ie = int(r[ci["Instructions Executed"]] or 0)
    agg[loc][0] += s; agg[loc][1] += ie; tot += s
for loc, (s, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    print(f"{loc[0]}:{loc[1]:<5d} samples {s:6d} ({100*s/tot:5.1f}%)  warp-inst {ie}")
