"""Summarise the round's ncu evidence into markdown (run here, on files gpurun brought back in gpurun_out/).
usage: python profiles/summarize.py <round tag, e.g. r1>
  gpurun_out/launches_<tag>.csv            -> profiles/launches_<tag>.csv / .md   (share of captured time per kernel)
  gpurun_out/full_<kernel>_<tag>.ncu-rep   -> profiles/ncu_summary_<tag>.md, profiles/traffic.json (via `ncu -i ... --page raw --csv`)"""
import collections, csv, glob, io, json, os, re, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
src = os.path.join(root, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(here, f"launches_{tag}.csv"))
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("<unnamed>::", "")
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(here, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list summary, round {tag[1:]} (`ncu --metrics gpu__time_duration.sum --clock-control none ... "
                f"python bench.py --steps 2 --warmup 3 --skip-cpu ...`, exact command in profiles/capture_{tag}.sh)\n\n"
                "Cold-cache, serialised per-launch times: compare SHARES, not absolutes. Raw list: "
                f"`profiles/launches_{tag}.csv`.\n\n| kernel | launches | total us | avg us | share of captured time |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f} % |\n")
        # the headline step = prep_windows + plan + assemble<0,0> on the full 4096-window batch: the same kernels also run
        # on the 512-window chunks of the e2e leg, so take the largest launches of each (one per timed step)
        per = collections.defaultdict(list)
        for r in rows[1:]:
            if len(r) >= len(hdr) and r[ci["Metric Name"]] == "gpu__time_duration.sum":
                v = float(r[ci["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ci["Metric Unit"]], 1e-3)
                per[re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("<unnamed>::", "")].append(v)
        pick = {}
        for name, key in (("assemble", "assemble_kernel<0, 0>"), ("plan", "plan_kernel"), ("prep", "prep_windows_kernel")):
            cand = [v for k, vs in per.items() if key in k for v in vs]
            if cand:
                top = sorted(cand, reverse=True)[:5]
                pick[name] = sum(top) / len(top)
        if len(pick) == 3:
            tot3 = sum(pick.values())
            events = ""
            for bj in (os.path.join(root, "gpurun_out", f"bench_{tag}.json"), os.path.join(here, f"bench_{tag}.json")):
                if os.path.exists(bj):
                    km = json.load(open(bj))["roofline"]["kernel_ms_per_step"]
                    te = sum(km.get(k, 0.0) for k in ("assemble_hb", "plan_windows", "prep_windows"))
                    events = (". `bench.py` (CUDA events, warm, same command without ncu) reports " + " / ".join(
                        f"{100 * km.get(k, 0.0) / te:.0f}" for k in ("assemble_hb", "plan_windows", "prep_windows")) +
                        " % (`roofline.kernel_ms_per_step`; there plan and prep run side by side on two streams and each takes a little longer)")
                    break
            f.write("\nHeadline step (full-size launches only, mean of the 5 largest of each kernel): " + ", ".join(
                f"{k} {v:.1f} us = {100 * v / tot3:.1f} %" for k, v in pick.items()) + events + ".\n")

want = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
        ("smsp__issue_active.avg.pct", "issue slots busy %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA sub-pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction")]
out = [f"# ncu --set full summaries, round {tag[1:]} (one launch per kernel, `--clock-control none`; profiles/capture_{tag}.sh)\n",
       "Timings under ncu are cold-cache and serialised: they are evidence of WHERE the time goes, never bench values.\n"]
traffic = {}
for rep in sorted(glob.glob(os.path.join(root, "gpurun_out", f"full_*_{tag}.ncu-rep"))):
    kname = os.path.basename(rep)[5:-len(f"_{tag}.ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out.append(f"\n## `{kname}`  ({m.get('Kernel Name', ('?', ''))[0][:110]})\n\n| metric | value |\n|---|---|")
    for key, label in want:
        hit = next((h for h in hdr if h.startswith(key)), None)
        if hit:
            out.append(f"| {label} | {m[hit][0]} {m[hit][1]} |")
    stalls = sorted(((float(v or 0), h) for h, (v, u) in m.items()
                     if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")), reverse=True)[:5]
    out.append("| top stall reasons (warps per issue) | " + ", ".join(
        f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in stalls) + " |")
    def num(key):
        hit = next((h for h in hdr if h.startswith(key)), None)
        if not hit:
            return None
        v, u = m[hit]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    r, w = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    if r is not None and w is not None:
        traffic[kname] = {"dram_bytes_read": r, "dram_bytes_write": w, "dram_bytes": r + w}
open(os.path.join(here, f"ncu_summary_{tag}.md"), "w").write("\n".join(out) + "\n")
tj = os.path.join(here, "traffic.json")
old = json.load(open(tj)) if os.path.exists(tj) else {}
# top-level keys are what bench.py reads for roofline.traffic (library kernel names)
for lib, k in (("assemble_hb", "assemble_kernel"), ("linearize_points", "points_kernel"), ("schur_landmarks", "schur_tma_kernel"),
               ("assoc_match", "match_kernel"), ("assoc_cull", "cull_tiles_kernel"), ("plan_windows", "plan_kernel")):
    if k in traffic:
        old[lib] = traffic[k]["dram_bytes"]
old["_source"] = f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, round {tag[1:]} (profiles/ncu_summary_{tag}.md)"
old["kernels_" + tag] = traffic
json.dump(old, open(tj, "w"), indent=1)
print("\n".join(out))
