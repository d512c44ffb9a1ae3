#!/usr/bin/env python
"""Turn the artefacts of tools/gpu_profile.sh (gpurun_out/<tag>_<case>_raw.csv, _hotspots.txt, _launches.csv — the
.ncu-rep reports are reduced to these CSV pages on the GPU box) into the committed summaries under profiles/:
  profiles/<tag>_<case>_ncu_summary.txt    key counters of the --set full capture of both kernels
  profiles/<tag>_<case>_launches_summary.txt per-kernel totals / shares of the launch list of the bench command
  profiles/<tag>_<case>_hotspots.txt       stall samples per source line / SASS instruction
  profiles/traffic.json                    dram bytes per launch of each cfg-2 kernel (read by bench.py: roofline.traffic)
usage: make_profile_summary.py <tag> [case ...]   (cases default to cfg2 cfg4 active32)"""
import collections
import csv
import json
import os
import shutil
import sys

tag = sys.argv[1]
cases = sys.argv[2:] or ["cfg2", "cfg4", "active32"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
DESC = {"cfg2": "bench.py --snowpacks 256 --steps 1 --warmup 0 (1536 solves of cfg 2: 20 layers, 32 streams)",
        "cfg4": "bench.py --workload cfg4 --snowpacks 13 --steps 1 --warmup 0 (156 solves of cfg 4: 50 layers, 64 streams; "
                "the 64 < h <= 128 instantiations)",
        "active32": "bench.py --workload active32 --snowpacks 60 --steps 1 --warmup 0 (180 solves: 10 layers, 32 streams, "
                    "active m_max = 2: blocks of 64 / 96 unknowns)"}
SOLVES = {"cfg2": 1536.0, "cfg4": 156.0, "active32": 180.0}
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
for case in cases:
    raw = os.path.join(G, f"{tag}_{case}_raw.csv")
    if not os.path.exists(raw):
        continue
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"ncu --set full --clock-control none --import-source on, {DESC[case]} on one B200; tag {tag}.",
           "NOTE: times under the profiler are serialised / cold-cache: compare shares, not absolutes.", ""]
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
        # the capture holds the FIRST launch of each kernel: one chunk of the batch
        bgrid = [float(dict(zip(hdr, q))["launch__grid_size"]) for q in rows[2:] if "boundary" in dict(zip(hdr, q))["Kernel Name"]]
        nsolves = min(SOLVES[case], max(4.0 * (bgrid[0] if bgrid else 148.0), 1024.0)) if case == "cfg2" else SOLVES[case]
        tot = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
        traffic[short] = {"dram_bytes_per_launch_in_capture": tot, "solves_in_captured_launch": nsolves,
                          "dram_bytes_per_solve": tot / nsolves}
        try:  # executed FP64 work: pipe utilisation x cycles x 64 FMA lanes x 2 flops per SM
            fp64 = float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]) / 100.0
            cyc = float(d["sm__cycles_elapsed.max"]); sms = 148
            out.append(f"{'derived: executed MFLOP per solve (fp64 pipe x cycles x 148 SM x 128 flop/clk)':82s} "
                       f"{fp64 * cyc * sms * 128 / nsolves / 1e6:.1f}")
        except (KeyError, ValueError):
            pass
    open(os.path.join(P, f"{tag}_{case}_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
    if case == "cfg2":
        json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    else:
        json.dump(traffic, open(os.path.join(P, f"{tag}_{case}_traffic.json"), "w"), indent=1)
    hs = os.path.join(G, f"{tag}_{case}_hotspots.txt")
    if os.path.exists(hs):
        shutil.copy(hs, os.path.join(P, f"{tag}_{case}_hotspots.txt"))
    lp = os.path.join(G, f"{tag}_{case}_launches.csv")
    if os.path.exists(lp):
        tot = collections.OrderedDict(); n = collections.Counter()
        for r in csv.DictReader(l for l in open(lp) if not l.startswith("==")):
            try: t = float(r["Metric Value"])
            except (KeyError, ValueError): continue
            k = r["Kernel Name"].split("(")[0][:70]
            tot[k] = tot.get(k, 0.0) + t; n[k] += 1
        s = sum(tot.values())
        lines = [f"ncu --metrics gpu__time_duration.sum --clock-control none, launch list of the {case} bench command "
                 f"(--steps 1 --warmup 1 --no-cpu-baseline); tag {tag}",
                 "per kernel: launches, total ms, share of the captured device time (cold-cache, serialised: shares only)", ""]
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            lines.append(f"{n[k]:5d}  {v/1e6:10.3f} ms  {100*v/s:6.2f} %  avg {v/n[k]/1e6:8.3f} ms  {k}")
        open(os.path.join(P, f"{tag}_{case}_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(out[:3])); print(json.dumps(traffic, indent=1))
