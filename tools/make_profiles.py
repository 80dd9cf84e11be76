"""Turns the raw artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/:
    python tools/make_profiles.py r2
  gpurun_out/<tag>_blend.ncu-rep, <tag>_other.ncu-rep  -> profiles/<tag>_kernels.md, <tag>_traffic.json
  gpurun_out/<tag>_launches.csv                         -> profiles/<tag>_launches.csv, <tag>_launches_summary.md
  gpurun_out/stages_C*.json                             -> profiles/<tag>_stages.md
  gpurun_out/truth_C*.json                              -> profiles/<tag>_truth.md
  built objects (cuobjdump -sass)                       -> profiles/<tag>_sass.md
"""
import collections, csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"

METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem"]


def short(name):
    m = re.search(r"(?:dogs::)?(\w+)(?:<[^(]*>)?\(", name)
    return m.group(1) if m else name[:40]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def kernels():
    lines = [f"# ncu --set full counters of the hot kernels ({tag}; C3 step through the public API, tools/gpu_steps.py)", "",
             "One launch per kernel (cold-cache, serialised by ncu: compare with the live CUDA-event times in "
             f"{tag}_stages.md).  issue = warp instructions / (148 SMs x 4 schedulers x active cycles).", ""]
    traffic = {}
    for rep in (f"{tag}_blend.ncu-rep", f"{tag}_other.ncu-rep"):
        path = os.path.join(GO, rep)
        if not os.path.exists(path):
            continue
        hdr, units, rows = raw(path)
        ix = {h: i for i, h in enumerate(hdr)}
        stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        seen = set()
        for r in rows:
            name = short(r[ix["Kernel Name"]])
            key = (name, r[ix["Grid Size"]])
            if key in seen:
                continue
            seen.add(key)
            lines.append(f"## {name}  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}")
            lines.append("| metric | value | unit |\n|---|---|---|")
            for m in METRICS:
                if m in ix:
                    lines.append(f"| {m} | {r[ix[m]]} | {units[ix[m]]} |")
            inst = float(r[ix["smsp__inst_executed.sum"]].replace(",", ""))
            cyc = float(r[ix["sm__cycles_active.avg"]].replace(",", ""))
            lines.append(f"| issue slots used | {inst / (148 * 4 * cyc):.3f} | of 1 |")
            st = sorted(((float(r[ix[h]].replace(",", "")), h) for h in stall if r[ix[h]]), reverse=True)[:6]
            lines.append("| top stalls (cycles per issued instruction) | " + ", ".join(
                f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st) + " | |")
            lines.append("")
            unit = units[ix["dram__bytes_read.sum"]].lower()
            mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
            dram = (float(r[ix["dram__bytes_read.sum"]].replace(",", "")) + float(r[ix["dram__bytes_write.sum"]].replace(",", ""))) * mul
            if name not in traffic:
                traffic[name] = {"dram_bytes": int(dram), "warp_instructions": int(inst), "grid": r[ix["Grid Size"]],
                                 "time_us": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[ix["gpu__time_duration.sum"]], 1)}
    if traffic:
        open(os.path.join(OUT, f"{tag}_kernels.md"), "w").write("\n".join(lines) + "\n")
        json.dump(traffic, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
        print("kernels:", sorted(traffic))


def launches():
    src = os.path.join(GO, f"{tag}_launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(OUT, f"{tag}_launches.csv"))
    rows = [r for r in csv.DictReader(l for l in open(src) if l.startswith('"'))]
    # the capture holds every launch of the run; keep the LAST complete step (from its preprocess_fwd on)
    starts = [i for i, r in enumerate(rows) if "preprocess_fwd_kernel" in r["Kernel Name"]]
    n_all = len(rows)
    if len(starts) >= 2:
        rows = rows[starts[-1]:]
    t, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows:
        k = short(r["Kernel Name"])
        if "at::" in r["Kernel Name"]:
            k = "torch:" + k
        t[k] = t.get(k, 0.0) + float(r["Metric Value"].replace(",", ""))
        cnt[k] += 1
    tot = sum(t.values())
    unit = rows[0].get("Metric Unit", "ns") if rows else "ns"
    div = {"ns": 1e6, "us": 1e3, "ms": 1.0}.get(unit, 1e6)
    lines = [f"# ncu launch list of `bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-admm --no-extra --no-e2e --no-roofline` ({tag})", "",
             f"{n_all} launches captured over the whole run; below, the {len(rows)} launches of the last timed step "
             f"(forward, loss, backward): {tot / div:.3f} ms of kernel time (cold-cache, serialised: compare SHARES).", "",
             "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        lines.append(f"| {k} | {cnt[k]} | {v / div:.3f} | {100 * v / tot:.1f}% |")
    open(os.path.join(OUT, f"{tag}_launches_summary.md"), "w").write("\n".join(lines) + "\n")
    print("launches:", len(rows))


def stages():
    lines = [f"# Device time per stage (library stage timer, CUDA events, median of 15; tools/gpu_stages.py) — {tag}", "",
             "| stage | C1 ms | C2 ms | C3 ms |", "|---|---|---|---|"]
    data = {}
    for c in ("C1", "C2", "C3"):
        p = os.path.join(GO, f"stages_{c}.json")
        if os.path.exists(p):
            data[c] = json.load(open(p))
    if not data:
        return
    keys = [k for k in next(iter(data.values())) if k not in ("cfg", "depth_key_bits", "R")]
    for k in keys:
        lines.append(f"| {k} | " + " | ".join(f"{data[c][k]:.4f}" if c in data and data[c].get(k) is not None else "" for c in ("C1", "C2", "C3")) + " |")
    lines.append("| depth key bits | " + " | ".join(str(data[c].get("depth_key_bits", "")) if c in data else "" for c in ("C1", "C2", "C3")) + " |")
    lines += ["", "(gather_scan_sync includes the host read of the instance count: the stage tool uses the exact, synchronising path so",
              "that the stages line up; the training path does not wait there.  The photometric loss, ~0.16 ms at C3, is not a library stage.)"]
    open(os.path.join(OUT, f"{tag}_stages.md"), "w").write("\n".join(lines) + "\n")
    print("stages:", sorted(data))


def truth():
    lines = [f"# Gradient error against the FP64 truth (oracle/truth_bwd.cu; tools/gpu_truth_compare.py) — {tag}", "",
             "max |g - truth| / max |truth| per tensor; `n>1e-4` = elements beyond the 1e-4 bar.", ""]
    any_ = False
    for c in ("C2", "C3"):
        p = os.path.join(GO, f"truth_{c}.json")
        if not os.path.exists(p):
            continue
        any_ = True
        d = json.load(open(p))
        lines += [f"## {c}", "| tensor | this package | n>1e-4 | reference run 1 | run 2 | n>1e-4 | ours vs reference |", "|---|---|---|---|---|---|---|"]
        for k in ("means3D", "means2D", "dc", "sh", "opacities", "scales", "rotations"):
            e = d[k]
            lines.append(f"| {k} | {e['ours']:.2e} | {e['n_out_ours']} | {e['ref'][0]:.2e} | {e['ref'][1]:.2e} | {e['n_out_ref']} | {e['ours_vs_ref']:.2e} |")
        lines.append("")
    if any_:
        open(os.path.join(OUT, f"{tag}_truth.md"), "w").write("\n".join(lines) + "\n")
        print("truth ok")


def sass():
    hot = {"blend_bwd": "blend_bwd_kernelILb0", "blend_fwd": "blend_fwd_kernelILb0", "preprocess_fwd": "preprocess_fwd_kernel",
           "preprocess_bwd": "preprocess_bwd_kernel", "radix_sort": "radix_scatter_kernel", "admm": "admm_consensus_kernel"}
    files = {"blend_bwd": "blend_bwd.o", "blend_fwd": "blend_fwd.o", "preprocess_fwd": "rasterizer_fwd.o", "preprocess_bwd": "rasterizer_bwd.o",
             "radix_sort": "radix_sort.o", "admm": "admm.o"}
    lines = [f"# SASS evidence (cuobjdump -sass of the built objects, sm_100a) — {tag}", "",
             "Mnemonics that matter on sm_100a: `UBLKCP` = cp.async.bulk (TMA engine, 1-D bulk copy), `SYNCS` = mbarrier,",
             "`LDGSTS` = cp.async, `REDG` / `RED` = red.global, `MATCH` / `VOTE` / `SHFL` = warp primitives, `ACQBULK` = griddepcontrol.wait (programmatic dependent launch), `UTMACMDFLUSH` = cp.async.bulk.commit_group.",
             "No `UTC*MMA` / `HMMA`: no stage is a dense contraction.", ""]
    for name, sym in hot.items():
        obj = os.path.join(ROOT, "dogs_b200", "build", files[name])
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        # split per function
        funcs = re.split(r"\n\s*Function : ", out)
        for f in funcs[1:]:
            fname = f.split("\n", 1)[0]
            if sym not in fname:
                continue
            ops = collections.Counter()
            for l in f.splitlines():
                m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
                if m:
                    ops[m.group(1)] += 1
            tot = sum(ops.values())
            special = {k: v for k, v in ops.items() if k in ("UBLKCP", "SYNCS", "LDGSTS", "REDG", "RED", "MATCH", "VOTE", "SHFL", "MUFU", "DFMA", "DMUL", "DADD",
                                                              "ACQBULK", "UTMACMDFLUSH", "ATOMG", "ATOMS", "LDGDEPBAR", "DEPBAR", "BAR")}
            lines.append(f"## {fname[:100]}")
            lines.append(f"{tot} instructions; " + ", ".join(f"{k} {v}" for k, v in ops.most_common(12)))
            lines.append("notable: " + (", ".join(f"{k} {v}" for k, v in sorted(special.items())) or "-"))
            lines.append("")
            break
    open(os.path.join(OUT, f"{tag}_sass.md"), "w").write("\n".join(lines) + "\n")
    print("sass ok")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    kernels(); launches(); stages(); truth(); sass()
