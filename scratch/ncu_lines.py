"""per-source-line summary of one kernel of an ncu report: samples, instructions, L1 wavefronts"""
import csv, sys, subprocess, os
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
idx = sys.argv[4] if len(sys.argv) > 4 else None
cmd = ["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+kern]
if idx: cmd += ["--launch-skip", idx, "--launch-count", "1"]
raw = subprocess.run(cmd, stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(raw.splitlines()))
fname = "?"; hdr = None; seen = set(); data = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {}; [ix.setdefault(h,i) for i,h in enumerate(hdr)]; continue
    if hdr is None or not r[0].isdigit(): continue
    key = (fname, r[0])
    if key in seen: continue
    seen.add(key)
    def col(name):
        try: return float(r[ix[name]])
        except Exception: return 0.0
    data.append(dict(file=fname, line=int(r[0]), src=r[1].strip(), samples=col("# Samples"), inst=col("Instructions Executed"),
                     thr=col("Avg. Threads Executed"), l1g=col("L1 Tag Requests Global"), shwf=col("L1 Wavefronts Shared"),
                     shex=col("L1 Wavefronts Shared Excessive"), long_sb=col("stall_long_sb"), wait=col("stall_wait"), short_sb=col("stall_short_sb"),
                     math=col("stall_math"), mio=col("stall_mio"), lg=col("stall_lg"), br=col("stall_branch_resolving")))
ts = sum(d["samples"] for d in data) or 1; ti = sum(d["inst"] for d in data) or 1
print("total samples %d  warp-instructions %.4g  L1 global tag req %.4g  shared wavefronts %.4g (excess %.4g)" % (
    ts, ti, sum(d["l1g"] for d in data), sum(d["shwf"] for d in data), sum(d["shex"] for d in data)))
data.sort(key=lambda d: -d["samples"])
for d in data[:top]:
    print("%4.1f%% in%4.1f%% thr%3.0f L1g %8.3g shwf %8.3g ex %8.3g | lsb %4.0f wt %4.0f ssb %4.0f mio %3.0f lg %3.0f | %s:%d %s" % (
        100*d["samples"]/ts, 100*d["inst"]/ti, d["thr"], d["l1g"], d["shwf"], d["shex"],
        100*d["long_sb"]/max(d["samples"],1), 100*d["wait"]/max(d["samples"],1), 100*d["short_sb"]/max(d["samples"],1),
        100*d["mio"]/max(d["samples"],1), 100*d["lg"]/max(d["samples"],1), d["file"][:14], d["line"], d["src"][:90]))
# ---- aggregated by user-given line ranges:  name=file:lo-hi,... via env NCU_GROUPS
import os as _os
g = _os.environ.get("NCU_GROUPS")
if g:
    print("---- groups")
    for spec in g.split(";"):
        name, ranges = spec.split("=")
        s = i = l1 = sh = 0.0
        for rg in ranges.split(","):
            f, lh = rg.split(":"); lo, hi = map(int, lh.split("-"))
            for d in data:
                if d["file"].startswith(f) and lo <= d["line"] <= hi:
                    s += d["samples"]; i += d["inst"]; l1 += d["l1g"]; sh += d["shwf"]
        print("%-12s samples %5.1f%%  inst %5.1f%%  L1g %5.1f%%  shwf %5.1f%%" % (name, 100*s/ts, 100*i/ti,
              100*l1/max(sum(d["l1g"] for d in data),1), 100*sh/max(sum(d["shwf"] for d in data),1)))
