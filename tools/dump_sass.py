"""Write the SASS of the default-dispatched hot kernels under profiles/ (one file per kernel, gzip for the big
ones) plus a mnemonic summary that shows the Blackwell-native instructions (UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG = TMA, FFMA2 = packed fp32, MUFU.EX2, USETMAXREG = setmaxnreg).
Usage: python tools/dump_sass.py [round_tag]"""
import collections
import gzip
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "tokenflow_b200", "libtokenflow_b200.so")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
WANT = {                      # demangled-name fragment -> file tag (the variants launch_ext_attn dispatches by default)
    "ext_attn_q4_kernel<3, true>": "ext_attn_q4_d40",
    "ext_attn_q4_kernel<4, false>": "ext_attn_q4_d64",
    "ext_attn_q4d_kernel<3>": "ext_attn_q4d_pairs",
    "ext_attn_h2_kernel<2, 3>": "ext_attn_h2_d80",
    "ext_attn_kernel<3, 64>": "ext_attn_v1_d160",
    "nn_field_kernel<1, 256, true>": "nn_field",
    "propagate_kernel<false, true>": "propagate",
    "layernorm_rows_kernel": "layernorm_rows",
    "cfg_ddim_kernel": "cfg_ddim",
}
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "FFMA2", "MUFU.EX2", "USETMAXREG", "SYNCS",
       "HMMA", "FMNMX3", "F2FP", "LDG", "STG", "BAR.SYNC")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
parts = re.split(r"\n\s*Function : ", sass)
summary = {}
for part in parts[1:]:
    mangled, _, body = part.partition("\n")
    name = subprocess.run(["c++filt", mangled.strip()], capture_output=True, text=True).stdout.strip()
    for frag, tag in WANT.items():
        if frag in name:
            lines = [ln for ln in body.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
            ops = collections.Counter()
            for ln in lines:
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", ln)
                if m:
                    ops[m.group(1)] += 1
            counts = {k: sum(v for op, v in ops.items() if op.startswith(k)) for k in KEY}
            summary[tag] = {"kernel": name[:160], "instructions": len(lines), **{k: v for k, v in counts.items() if v}}
            text = f"// {name}\n// cuobjdump -sass tokenflow_b200/libtokenflow_b200.so (sm_100a)\n" + body
            path = os.path.join(REPO, "profiles", f"{TAG}_sass_{tag}.txt")
            if len(text) > 200_000:
                with gzip.open(path + ".gz", "wt") as f:
                    f.write(text)
            else:
                with open(path, "w") as f:
                    f.write(text)
with open(os.path.join(REPO, "profiles", f"{TAG}_sass_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)
print(json.dumps(summary, indent=1))
