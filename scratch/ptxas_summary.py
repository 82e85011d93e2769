import re, sys, subprocess
t=open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv)>2 else ''
rows=[]
for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t, re.S):
    rows.append((m.group(1), m.group(2), m.group(3), m.group(5)))
names = subprocess.run(['c++filt']+[r[0] for r in rows], stdout=subprocess.PIPE).stdout.decode().splitlines()
for (n,st,sp,rg),d in zip(rows,names):
    d = re.sub(r'\(anonymous namespace\)::','',d); d=d.split('(')[0]
    if pat in d: print(d, 'stack',st,'spill',sp,'regs',rg)
