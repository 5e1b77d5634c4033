import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e); continue
    print(f"== {path}: {d['config']['workload']}")
    print(f"   search {d['value']:.4g} q/s ({d['search']['ms']:.3f} ms, hits {d['search']['hits']}), build {d['build']['value']:.3f} Gbases/s ({d['build']['ms']:.3f} ms)")
    for k, v in d["kernels"].items():
        g = v['algorithmic_gbs']
        print(f"   {k:22s} x{v['launches_per_step']:<5.0f} {v['ms_per_step']:9.3f} ms  {('%.0f GB/s' % g) if g else ''}")
    print("   roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k != 'peak_source'})
    if "roofline_search" in d: print("   roofline_search", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline_search"].items() if k != 'peak_source'})
    if "e2e" in d: print("   e2e", d["e2e"])
    if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"])
    print("   clocks", d["clocks"], "launches", d["gpu_launches"])
