"""Print the per-step timing table written by `bench.py --profile-out` (developer helper)."""
import json, sys
d = json.load(open(sys.argv[1]))
agg = {}
for s in d["steps"]:
    k = s["name"]
    import re
    k = re.sub(r"blocks\.\d+\.", "blocks.*.", k)
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += s["ms"]; a[2] += s.get("algo_flops", 0); a[3] += s.get("algo_bytes", 0)
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print(f"{k:24s} x{a[0]:<2d} {a[1]*1000:8.1f} us {100*a[1]/tot:5.1f}%  {a[2]/a[1]/1e9 if a[1] else 0:8.1f} TF/s {a[3]/a[1]/1e6 if a[1] else 0:8.1f} GB/s")
print(f"total {tot*1000:.1f} us")
