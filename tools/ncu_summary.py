"""Summarises ncu outputs into profiles/: per-kernel time shares of one bench step from a launch list
(--metrics gpu__time_duration.sum CSV) and key counters of a --set full report.
  python tools/ncu_summary.py launches gpurun_out/launches.csv
  python tools/ncu_summary.py report gpurun_out/prof_r1.ncu-rep"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    m = re.search(r"(\w+)(?:<.*>)?\(", name)
    n = m.group(1) if m else name
    n = re.sub(r"^void\s+", "", name) if n in ("void",) else n
    if "at::" in name or "elementwise" in name or "reduce_kernel" in name:
        m = re.search(r"(\w*(?:elementwise|reduce|fill|copy|index|uniform|normal)\w*)", name)
        return "torch:" + (m.group(1) if m else "op")
    return n[:60]


def launches(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    t = collections.OrderedDict()
    cnt = collections.Counter()
    for r in rows:
        k = short(r["Kernel Name"])
        t[k] = t.get(k, 0.0) + float(r["Metric Value"])
        cnt[k] += 1
    tot = sum(t.values())
    print(f"# {len(rows)} launches, {tot/1e6:.3f} ms total kernel time (cold-cache, serialised: compare shares)")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {cnt[k]} | {v/1e6:.3f} | {100*v/tot:.1f}% |")


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in data:
        k = short(r[idx["Kernel Name"]])
        if k in seen:
            continue
        seen.add(k)
        print(f"\n## {k}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for m in METRICS:
            if m in idx:
                print(f"  {m:70s} {r[idx[m]]:>16s} {units[idx[m]]}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
