"""Dev-time: one-screen summary of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(f"N={d['n_gpus']} steps={d['steps']} ray {d['value']:.0f} Mrays/s ({1e3*r['kernel_ms']:.1f} us/frame, alone {1e3*r['kernel_ms_alone']:.1f}) e2e {d['e2e']['value']:.0f} dense {d['e2e']['dense_readback']['value']:.0f} | roofline {r['bound']} {r['frac']:.3f} (hbm {r.get('hbm', r)['frac']:.3f}) | ff {d['frame_filling']['value']:.0f} ({1e3*d['frame_filling']['ms_per_frame']:.0f} us)")
print("  ms_per_step", round(d["ms_per_step"], 2), "clocks", d["clocks"], "per-rank ms", [round(x, 1) for x in d["run"]["timed_region_ms_per_rank"]])
s = d["secondary"]
print(f"  raster {s['value']:.0f} Mtris/s ({1e3*s['roofline']['frame_ms']:.1f} us/frame, alone {1e3*s['roofline']['frame_ms_alone']:.1f}) e2e {s['e2e']['value']:.0f} hbm frac {s['roofline']['frac']:.3f} issue {(s['roofline'].get('issue') or {}).get('frac')} per-rank ms {[round(x, 1) for x in s['run']['timed_region_ms_per_rank']]}")
if "tiles" in d:
    t = d["tiles"]
    print(f"  tiles: ray {t['raycast']['value']:.0f} Mrays/s ({1e3*t['raycast']['ms_per_frame']:.1f} us/frame)  raster {t['raster']['value']:.0f} Mtris/s ({1e3*t['raster']['ms_per_frame']:.1f} us/frame)")
if "config4" in d:
    c = d["config4"]
    print(f"  config4: raster {c['raster']['value']:.0f} Mtris/s ({1e3*c['raster']['ms_per_frame_per_rank']:.0f} us/frame/rank, hbm {c['raster']['roofline']['frac']:.2f})  ray {c['raycast']['value']:.0f} Mrays/s ({1e3*c['raycast']['ms_per_frame_per_rank']:.0f} us/frame/rank, hbm {c['raycast']['roofline']['frac']:.2f})")
print("  cpu:", d.get("cpu_baseline", {}).get("value"), s.get("cpu_baseline", {}).get("value"), "cores", d.get("cpu_baseline", {}).get("cores"))
