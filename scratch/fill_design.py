"""fills the round-2 table of DESIGN.md (between the 8b-table markers) from profiles/r2_*.json"""
import json, re
def last(f):
    return json.loads(open("profiles/" + f).read().strip().splitlines()[-1])
d1, d2, d8, ref = last("r2_bench_cloud64m.json"), last("r2_bench_cloud64m_2gpu.json"), last("r2_bench_cloud64m_8gpu.json"), last("r2_bench_reference_arm.json")
d4 = last("r2_bench_cloud64m_4gpu.json")
e = lambda v: "%.3g" % v if v < 1e4 else ("%.2fe+08" % (v/1e8))
b = d1.get("blocks_ms") or d1.get("kernel_ms") or {}
blk = d1.get("breakdown_ms", {})
vals = {
 "V1": e(d1["value"]), "MS1": "%.0f" % d1["ms_per_step"], "E1": e(d1["e2e"]["value"]),
 "R1": e(d1["resident"]["value"]), "RMS1": "%.0f" % d1["resident"]["ms_per_step"],
 "REF": "%.2fe+05" % (ref["value"]/1e5), "SPD": "%.0f" % (d1["e2e"]["value"]/ref["value"]),
 "FP": "%.2f" % d1["roofline_fp64"]["achieved"], "FPF": "%.1f" % (100*d1["roofline_fp64"]["frac"]),
 "HB": "%.1f" % d1["roofline"]["achieved"], "HBF": "%.2f" % (100*d1["roofline"]["frac"]),
 "V2": e(d2["value"]), "MS2": "%.0f" % d2["ms_per_step"], "E2": e(d2["e2e"]["value"]), "R2": e(d2["resident"]["value"]),
 "V4": e(d4["value"]), "MS4": "%.0f" % d4["ms_per_step"], "E4": e(d4["e2e"]["value"]), "R4": e(d4["resident"]["value"]),
 "V8": e(d8["value"]), "MS8": "%.1f" % d8["ms_per_step"], "E8": e(d8["e2e"]["value"]), "R8": e(d8["resident"]["value"]),
}
km = d1.get("kernels_ms") or {}
import subprocess
out = subprocess.run(["python", "scratch/show_bench.py", "profiles/r2_bench_cloud64m.json"], stdout=subprocess.PIPE).stdout.decode()
m = re.search(r"tree ([0-9.]+) .*?dens ([0-9.]+) hydro ([0-9.]+) grav ([0-9.]+)", out)
vals["BLK"] = " / ".join(m.groups()) if m else "?"
tpl = open("scratch/design_8b.tpl").read()
for k, v in vals.items():
    tpl = tpl.replace("@%s@" % k, v)
s = open("DESIGN.md").read()
a = s.index("<!-- 8b-table-begin -->") + len("<!-- 8b-table-begin -->\n")
b = s.index("<!-- 8b-table-end -->")
open("DESIGN.md", "w").write(s[:a] + tpl + s[b:])
print(vals)
