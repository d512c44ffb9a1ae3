#!/usr/bin/env python
"""Turn the artefacts of tools/gpu_session.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/:
  profiles/<tag>_ncu_summary.txt      key counters of the --set full capture of both kernels
  profiles/<tag>_launches_summary.txt per-kernel totals / shares of the launch list of the bench command
  profiles/<tag>_hotspots.txt         stall samples per source line and per phase
  profiles/traffic.json               dram bytes per launch of each kernel (read by bench.py for roofline.traffic)
usage: make_profile_summary.py <tag>"""
import csv, json, os, subprocess, sys, collections
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
rep = os.path.join(G, f"{tag}_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
out = [f"ncu --set full --clock-control none --import-source on, bench.py --snowpacks 256 --steps 1 --warmup 0 (1536 solves of cfg 2: "
       f"20 layers, 32 streams) on one B200; tag {tag}.", "NOTE: times under the profiler are serialised / cold-cache: compare shares, not absolutes.", ""]
traffic = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    out.append("-" * 100); out.append(f"{'Kernel Name':82s} {name}")
    for k in KEYS:
        if k in d: out.append(f"{k:82s} {d[k]} {units[hdr.index(k)]}")
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k:
            try:
                if float(d[k]) > 0.15: out.append(f"{k:82s} {d[k]}")
            except ValueError: pass
    def tobytes(key):
        v = float(d[key]); u = units[hdr.index(key)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    short = "eigen_kernel" if "eigen" in name else "boundary_kernel" if "boundary" in name else name
    # the capture holds the FIRST launch of each kernel: one chunk = max(4 x boundary grid, 1024) problems of the 1536
    bgrid = [float(dict(zip(hdr, q))["launch__grid_size"]) for q in rows[2:] if "boundary" in dict(zip(hdr, q))["Kernel Name"]]
    nsolves = min(1536.0, max(4.0 * (bgrid[0] if bgrid else 148.0), 1024.0))
    traffic[short] = {"dram_bytes_per_launch_in_capture": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
                      "solves_in_captured_launch": nsolves,
                      "dram_bytes_per_solve": (tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")) / nsolves}
open(os.path.join(P, f"{tag}_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
# launch list
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    tot = collections.OrderedDict(); n = collections.Counter()
    for r in csv.DictReader(l for l in open(lp) if not l.startswith("==")):
        try: t = float(r["Metric Value"])
        except (KeyError, ValueError): continue
        k = r["Kernel Name"].split("(")[0][:70]
        tot[k] = tot.get(k, 0.0) + t; n[k] += 1
    s = sum(tot.values())
    lines = [f"ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 1 --warmup 1 --no-cpu-baseline; tag {tag}",
             "per kernel: launches, total ms, share of the captured device time (cold-cache, serialised: shares only)", ""]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"{n[k]:5d}  {v/1e6:10.3f} ms  {100*v/s:6.2f} %  avg {v/n[k]/1e6:8.3f} ms  {k}")
    open(os.path.join(P, f"{tag}_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
# hotspots
cs = os.path.join(G, f"{tag}_cs.csv")
if not os.path.exists(cs):
    open(cs, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                                       capture_output=True, text=True).stdout)
hs = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hotspots.py"), cs, "25"], capture_output=True, text=True).stdout
st = ""
for kname in ("eigen", "boundary"):
    st += f"\n##### top SASS instructions by stall samples, {kname} kernel\n" + subprocess.run(
        [sys.executable, os.path.join(ROOT, "tools", "ncu_sass_top.py"), cs, kname, "25"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_hotspots.txt"), "w").write(hs + st)
print("\n".join(out[:3])); print(json.dumps(traffic, indent=1))
