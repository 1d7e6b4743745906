"""Dev-time: one .ncu-rep holding many kernels (tools/profile_all.py under ncu --set full) -> profiles/<tag>_<kernel>.txt for
every distinct kernel (its longest launch), profiles/<tag>_inventory.txt (one table) and profiles/traffic.json (the per-launch
facts bench.py quotes: DRAM bytes, executed warp instructions, issue-slot utilisation).

    python tools/profile_inventory.py /tmp/r02_all.ncu-rep r02
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
os.makedirs("profiles", exist_ok=True)
# rep: an .ncu-rep, or the CSV `ncu -i rep --page raw --csv` printed (so the summaries can be redone without the report)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    if name not in col:
        return 0.0
    try:
        x = float(r[col[name]].replace(",", ""))
    except ValueError:
        return 0.0
    u = units[col[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    return x * scale


def short(name):
    n = re.sub(r"\(.*", "", name)
    n = n.replace("void ", "").replace("<unnamed>::", "").replace("at::native::", "at::")
    return n.strip()


WANT = [
    ("gpu__time_duration.sum", "duration (us)"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_static", "static smem/block"),
    ("dram__bytes_read.sum", "DRAM read (bytes)"), ("dram__bytes_write.sum", "DRAM write (bytes)"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak (active)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak (active)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (should be 0)"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (active)"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard (cycles/instr)"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait (fixed latency)"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "stall branch_resolving"),
    ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "stall no_instruction"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not_selected"),
]
launches = collections.OrderedDict()
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    launches.setdefault(short(r[col["Kernel Name"]]), []).append(r)

table = []
facts = {}
for name, rs in launches.items():
    if name.startswith("at::") or "vectorized_elementwise" in name or "elementwise_kernel" in name or "reduce_kernel" in name or "index" in name.lower() and "at" in name[:3]:
        continue          # torch plumbing (fills, reductions for mesh bounds)
    best = max(rs, key=lambda r: num(r, "gpu__time_duration.sum"))
    fname = re.sub(r"[^A-Za-z0-9]+", "_", name).strip("_")
    lines = [f"source: one ncu run over tools/profile_all.py (ncu --set full --clock-control none --import-source on), tag {tag}; the longest of {len(rs)} launch(es)",
             f"{'kernel':42s} {best[col['Kernel Name']][:160]}"]
    for key, label in WANT:
        if key in col:
            lines.append(f"{label:42s} {best[col[key]]} {units[col[key]]}")
    traffic = num(best, "dram__bytes_read.sum") + num(best, "dram__bytes_write.sum")
    lines.append(f"{'DRAM traffic per launch (read+write)':42s} {traffic / 1e6:.2f} MB")
    lines.append(f"{'all launches of this kernel, us':42s} " + " ".join(f"{num(r, 'gpu__time_duration.sum'):.1f}" for r in rs))
    open(f"profiles/{tag}_{fname}.txt", "w").write("\n".join(lines) + "\n")
    facts[name] = {"dram_bytes_per_launch": traffic, "warp_instructions": num(best, "smsp__inst_executed.sum"),
                   "issue_slots_busy_pct": num(best, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "duration_us": num(best, "gpu__time_duration.sum"), "launches": len(rs), "all": rs}
    table.append((name, len(rs), num(best, "gpu__time_duration.sum"), best[col["launch__registers_per_thread"]], num(best, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                  num(best, "smsp__issue_active.avg.pct_of_peak_sustained_active"), num(best, "smsp__inst_executed.sum"), traffic / 1e6,
                  num(best, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), num(best, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                  num(best, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))
with open(f"profiles/{tag}_inventory.txt", "w") as fh:
    fh.write(f"Every kernel of librendertoy_b200.so (and the NVRTC-compiled user-shader kernels), one ncu --set full capture each (tag {tag}, tools/profile_all.py):\n"
             "the longest launch per kernel.  Per-launch times are cold-cache and serialised.\n\n")
    fh.write(f"{'kernel':58s} {'n':>3s} {'us':>8s} {'regs':>5s} {'occ %':>6s} {'issue %':>8s} {'warp instr':>12s} {'DRAM MB':>8s} {'DRAM %':>7s} {'FMA %':>6s} {'tensor %':>8s}\n")
    for t in table:
        fh.write(f"{t[0][:58]:58s} {t[1]:3d} {t[2]:8.1f} {t[3]:>5s} {t[4]:6.1f} {t[5]:8.1f} {t[6]:12.0f} {t[7]:8.2f} {t[8]:7.1f} {t[9]:6.1f} {t[10]:8.1f}\n")
print(open(f"profiles/{tag}_inventory.txt").read())


def pick(pattern, nth=-1):
    """facts of the nth launch (default: last) of the first kernel whose name matches"""
    for name, f in facts.items():
        if re.search(pattern, name):
            r = f["all"][nth]
            return {"dram": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"), "instr": num(r, "smsp__inst_executed.sum"),
                    "issue": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "us": num(r, "gpu__time_duration.sum")}
    return None


out = {}
# cfg4 frame (lesson06 camera at 4K, t = 0.5): the 2nd launch of project, the 3rd+4th refit launches, the 2nd raycast launch
ray = [f for f in (pick(r"^project_kernel", 0), pick(r"^raycast_kernel<8, 0, 0, 1>", 0)) if f]
refit = facts.get(next((n for n in facts if n.startswith("view_refit_kernel")), ""), None)
if len(ray) == 2:
    extra = [{"dram": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"), "instr": num(r, "smsp__inst_executed.sum"), "us": num(r, "gpu__time_duration.sum")}
             for r in (refit["all"][0:1] if refit else [])]
    out["raycast_frame"] = {"dram_bytes_per_launch": sum(f["dram"] for f in ray + extra), "warp_instructions_per_frame": sum(f["instr"] for f in ray + extra),
                            "issue_slots_busy_pct": ray[1]["issue"], "us_serialised": sum(f["us"] for f in ray + extra),
                            "source": f"profiles/{tag}_inventory.txt: project_kernel + view_refit_kernel ({len(extra)} launch) + raycast_kernel<8> of the cfg4 frame (lesson06 camera, t = 0.5)"}
ras = [pick(r"^fill_u64_kernel", 0), pick(r"^raster_kernel<8, 0>", 0), pick(r"^coverage_kernel<8, 0>", 0), pick(r"^resolve_kernel<8>", 0)]
if all(ras):
    out["raster_frame"] = {"dram_bytes_per_launch": sum(f["dram"] for f in ras), "warp_instructions_per_frame": sum(f["instr"] for f in ras),
                           "us_serialised": sum(f["us"] for f in ras),
                           "source": f"profiles/{tag}_inventory.txt: fill_u64 + raster_kernel<8> + coverage_kernel<8> + resolve_kernel<8> of the cfg2 frame (t = 0.5)"}
for name, f in facts.items():
    out.setdefault("kernels", {})[name] = {k: v for k, v in f.items() if k != "all"}
out["_note"] = f"per-launch facts from one ncu --set full run (tag {tag}): dram__bytes_read.sum + dram__bytes_write.sum, sm__inst_executed.sum, smsp__issue_active"
if "raycast_frame" in out and "raster_frame" in out:
    json.dump(out, open("profiles/traffic.json", "w"), indent=1)
else:
    json.dump(out, open(f"profiles/{tag}_facts.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "kernels"}, indent=1))
