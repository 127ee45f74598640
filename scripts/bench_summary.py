"""Short human-readable summary of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(f"c3: {d['ms_per_step']:.4f} ms/step, {d['value'] / 1e6:.1f} M solves/s, bwd {r['kernel_ms']:.4f} ms frac {r['frac']:.3f}, "
      f"e2e {d['e2e']['ms_per_step']:.2f} ms link_frac {d['e2e'].get('link_frac')}, parity {d.get('parity_rel_err')}")
for k, x in d.get("configs", {}).items():
    if "error" in x:
        print(k, "ERROR", x["error"])
        continue
    r = x["roofline"]
    extra = ""
    if "factorizing_kernel" in r:
        f = r["factorizing_kernel"]
        extra = f" | fact {f['kernel_ms']:.2f} ms hbm {f['frac']:.3f} fp64 {f['fp64']['frac']:.3f}"
    if "latency_vs_N_us" in x:
        extra += " | lat " + " ".join(f"{n}:{v['us']:.0f}" for n, v in x["latency_vs_N_us"].items())
    if "phase_us" in x:
        extra += " | phases " + " ".join(f"{n}:{v:.0f}" for n, v in x["phase_us"].items() if n != "note")
    print(f"{k}: {x['ms_per_step']:.4f} ms/step, kernel {r['kernel_ms']:.4f} ms hbm {r['frac']:.3f} fp64 {r['fp64']['frac']:.3f}, "
          f"e2e {x['e2e']['ms_per_step']:.3f} ms, parity {x.get('parity_rel_err')}, vs1gpu {x.get('parity_vs_1gpu')}, launches {x.get('gpu_launches')}{extra}")
