"""Dev-time: summarise gpurun_out/*_<tag>.* into profiles/ (committed) and profiles/traffic.json."""
import csv, collections, json, os, subprocess, sys
tag = sys.argv[1]
os.makedirs("profiles", exist_ok=True)
traffic = {}
if os.path.exists("profiles/traffic.json"):
    traffic = json.load(open("profiles/traffic.json"))
for k in ("raster_kernel", "coverage_kernel", "resolve_kernel", "raycast_kernel"):
    rep = f"gpurun_out/prof_{k}_{tag}.ncu-rep"
    if not os.path.exists(rep):
        continue
    out = subprocess.run([sys.executable, "tools/profile_summary.py", rep, k, f"profiles/{tag}_{k}.txt"], capture_output=True, text=True)
    print(out.stdout.strip(), out.stderr[-300:])
    traffic[k] = json.loads(out.stdout.strip().splitlines()[-1])["dram_bytes_per_launch"]
if all(k in traffic for k in ("raster_kernel", "coverage_kernel", "resolve_kernel")):
    traffic["raster_frame"] = traffic["raster_kernel"] + traffic["coverage_kernel"] + traffic["resolve_kernel"]
traffic["_note"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full captures tagged {tag} (profiles/{tag}_*.txt); raster_frame = sum of the three draw kernels (clears excluded)"
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
for name in ("raster", "raycast"):
    p = f"gpurun_out/launches_{name}_{tag}.csv"
    if not os.path.exists(p):
        continue
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")) / 1e3)
    with open(f"profiles/{tag}_launches_{name}.txt", "w") as fh:
        fh.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, python tools/quick_{name}_bench.py ncu (cold-cache, serialised: compare shares)\n")
        fh.write("per-launch device time in microseconds, in launch order per kernel\n\n")
        for k, v in agg.items():
            fh.write(f"{k[:100]}\n    n={len(v)}  " + " ".join(f"{x:.1f}" for x in v) + "\n")
    print(open(f"profiles/{tag}_launches_{name}.txt").read()[:3000])
