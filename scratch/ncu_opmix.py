"""dynamic opcode mix of one kernel of an ncu report (instructions executed / stall samples per opcode)"""
import csv, collections, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(raw.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 10: continue
    try: n = float(r[ix['Instructions Executed']]); s = float(r[ix['# Samples']])
    except Exception: continue
    data.append((r[ix['Source']].strip(), n, s))
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print('total instr %.3g samples %d' % (tot, ts))
op = collections.defaultdict(lambda: [0, 0])
for src, n, s in data:
    t = src.split()
    o = t[1] if t[0].startswith('@') else t[0]
    o = o.split('.')[0]
    op[o][0] += n; op[o][1] += s
for k, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    print('%-10s instr %5.1f%%  samples %5.1f%%' % (k, 100*v[0]/tot, 100*v[1]/ts))
