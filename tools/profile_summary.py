"""Dev-time: turn an .ncu-rep (one kernel, `ncu --set full --import-source on`) into a committed text summary.

    python tools/profile_summary.py gpurun_out/prof.ncu-rep <kernel-regex> profiles/r01_<name>.txt

Writes: headline metrics (duration, DRAM bytes/throughput, L2, pipe utilisation, occupancy, IPC, stall picture) and
the hottest source lines (via tools/ncu_lines.py).  Prints a JSON fragment with the per-launch DRAM traffic.
"""
import csv
import json
import subprocess
import sys

rep, pat, out = sys.argv[1], sys.argv[2], sys.argv[3]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def get(name, default="n/a"):
    return m.get(name, (default, ""))


WANT = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_static", "static smem/block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak (active)"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "FMA-heavy pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak (active)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (should be 0)"),
    ("sm__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (active)"), ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
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
    ("smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "stall dispatch"),
]
lines = [f"source: {rep}   (ncu --set full --clock-control none --import-source on, one launch)"]
for key, label in WANT:
    if key in m:
        v, u = m[key]
        lines.append(f"{label:42s} {v} {u}")


def num(name):
    v, u = get(name, "0")
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return 0.0
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return x * scale


traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
lines.append(f"{'DRAM traffic per launch (read+write)':42s} {traffic / 1e6:.2f} MB")
src = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, pat, "18"], capture_output=True, text=True).stdout
lines.append("")
lines.append("hottest source lines (share of executed warp instructions / of stall samples):")
lines.append(src)
open(out, "w").write("\n".join(lines) + "\n")
print(json.dumps({"kernel": pat, "dram_bytes_per_launch": traffic}))
