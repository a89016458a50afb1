"""Per-phase breakdown of one kernel from an `ncu --set full --import-source on` report: the SASS is cut at BAR.SYNC
instructions; per region: share of warp-stall samples, executed warp instructions and shared-memory wavefronts
(actual vs ideal), then the instructions with the most shared-memory wavefronts.
usage: python profiles/phase_breakdown.py <report.ncu-rep> <units per launch, e.g. 4096 windows>"""
import csv, io, subprocess, sys

rep, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]


def I(r, h):
    try:
        return int(r[ci[h]] or 0)
    except ValueError:
        return 0


tot_s = sum(I(r, "# Samples") for r in data) or 1
tot_i = sum(I(r, "Instructions Executed") for r in data) or 1
print(f"{rows[0][1][:100]}\nSASS instructions {len(data)}, samples {tot_s}, warp instructions per unit {tot_i / units:.0f}, "
      f"shared-memory wavefronts per unit {sum(I(r, 'L1 Wavefronts Shared') for r in data) / units:.0f} "
      f"(ideal {sum(I(r, 'L1 Wavefronts Shared Ideal') for r in data) / units:.0f})\n")
start, segs = 0, []
for k, r in enumerate(data):
    if "BAR.SYNC" in r[ci["Source"]] and I(r, "Instructions Executed") > 0:
        segs.append((start, k))
        start = k + 1
segs.append((start, len(data) - 1))
print("| SASS range (ends at a barrier) | samples | warp instr / unit | smem wavefronts / unit (ideal) |\n|---|---|---|---|")
for a, b in segs:
    s = sum(I(data[k], "# Samples") for k in range(a, b + 1))
    i = sum(I(data[k], "Instructions Executed") for k in range(a, b + 1))
    w = sum(I(data[k], "L1 Wavefronts Shared") for k in range(a, b + 1))
    wi = sum(I(data[k], "L1 Wavefronts Shared Ideal") for k in range(a, b + 1))
    if s / tot_s > 0.005 or i / tot_i > 0.005:
        print(f"| {a}-{b} | {100 * s / tot_s:.1f} % | {i / units:.0f} | {w / units:.0f} ({wi / units:.0f}) |")
print("\n| SASS index | instruction | executions / unit | smem wavefronts / unit | ideal |\n|---|---|---|---|---|")
for k in sorted(sorted(range(len(data)), key=lambda k: -I(data[k], "L1 Wavefronts Shared"))[:30]):
    r = data[k]
    print(f"| {k} | `{r[ci['Source']].strip()[:48]}` | {I(r, 'Instructions Executed') / units:.1f} | "
          f"{I(r, 'L1 Wavefronts Shared') / units:.1f} | {I(r, 'L1 Wavefronts Shared Ideal') / units:.1f} |")
