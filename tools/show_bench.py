import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "gpu_launches", "kernel_path"):
    print(k, d.get(k))
print("roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "traffic")})
print("e2e", d["e2e"]["value"])
for k in ("cpu_baseline", "roofline_hbm_case", "roofline_assembly", "assembly", "model_build", "clocks"):
    print(k, d.get(k))
