"""Pretty-prints the interesting parts of a bench.py JSON line."""
import json
import sys

for f in sys.argv[1:]:
    t = open(f).read().strip()
    if not t:
        print(f, "EMPTY")
        continue
    d = json.loads(t.splitlines()[-1])
    print(f, "value %.0f  ms %.4f  e2e %.0f  launches/step %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("launches_per_step")))
    r = d["roofline"]
    print("  class ms", r.get("conv_class_ms_per_step"), "TF/s", r.get("conv_class_tflops"), "frac %.4f" % r["frac"])
    lay = r["conv_us_per_layer"]
    for cls in ("fprop_tc", "dgrad_tc", "wgrad_tc", "fprop", "dgrad", "wgrad", "pack_weights"):
        row = [lay.get("%s[L%d]" % (cls, i)) for i in range(8)]
        if any(v is not None for v in row):
            print("  %-12s" % cls, row)
    print("  kernels", {k.replace("hmvae_", ""): v for k, v in list(r["kernel_ms_per_step"].items())[:14]})
    b = d.get("conv_large_batch")
    if b:
        print("  B=512:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in b.items() if k not in ("workload",)})
    if d.get("fk"):
        print("  fk:", d["fk"]["sizes"][str(d["fk"]["frames"])])
