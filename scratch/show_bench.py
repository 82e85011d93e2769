import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("no bench line:", e); sys.exit(0)
b = d.get("blocks_ms", {})
print("value %.4g  e2e %.4g  ms/step %.1f  | tree %.1f (sort %.1f lev %.1f stock %.1f) dens %.1f hydro %.1f grav %.1f | launches %s fp64frac %.3f" % (
    d["value"], d["e2e"]["value"], d["ms_per_step"], b.get("BUILD_TREE", 0), b.get("tree_sort", 0), b.get("tree_levels", 0), b.get("tree_stock", 0),
    b.get("density_kernel", 0), b.get("hydro_kernel", 0), b.get("gravity_kernel", 0), d.get("gpu_launches"), d.get("roofline_fp64", {}).get("frac", 0)))
print("stats", d.get("stats"))
if d.get("resident"):
    print("resident", d["resident"])
