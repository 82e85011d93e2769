"""Extracts the defaults Parameters::SetDefaultValues assigns (src/Common/Parameters.cpp:166-536 of the
UNMODIFIED reference) into tests/golden/ref_defaults.json, so that the CPU test of the .dat mirror
(gandalf_b200/params.py DEFAULTS) can run where /root/reference does not exist.

    python tests/golden/make_defaults_golden.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "src", "Common", "Parameters.cpp")).read()
body = src[src.index("void Parameters::SetDefaultValues"):]
body = body[:body.index("\n}\n")]
out = {}
for kind, key, val in re.findall(r'(int|float|string)params\["([^"]+)"\]\s*=\s*([^;]+);', body):
    val = val.strip()
    if kind == "string":
        out[key] = val.strip('"')
    elif kind == "int":
        out[key] = int(val)
    else:
        out[key] = float(val)
here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "ref_defaults.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print("%d defaults" % len(out))
