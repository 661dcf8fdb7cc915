"""SASS evidence for the hot kernels of libpgb.so (no GPU needed): per kernel the instruction count and a
histogram of the memory / synchronisation / arithmetic mnemonics that matter for the roofline discussion
(LDG/STG widths and cache hints, UBLKCP / SYNCS of the bulk-async variant, ATOMS of the hash table, SHFL,
DFMA ...), plus the lines of those instructions with their source line when -lineinfo is present.
usage: python scripts/sass_excerpts.py > profiles/r2_sass_excerpts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pygeon_b200", "libpgb.so")
WANT = [
    ("k_local<L1_STIFF, 3> (K1)", r"k_localILi1ELi3ELb0ELb0E"),
    ("k_numeric_t<4, 128, int> (K2 numeric)", r"k_numeric_tILi4ELi128EiLb0E"),
    ("k_numeric_bulk<4, 128, int> (K2, bulk-async variant)", r"k_numeric_bulkILi4ELi128EiLb0E"),
    ("k_numeric_g<4, 8, int> (K2, lane-group variant)", r"k_numeric_gILi4ELi8EiE"),
    ("k_spmv<4, int, CG> (K3)", r"6k_spmvILi4EiLi1EE"),
    ("k_spmv<2, int, MINRES> (K3)", r"6k_spmvILi2EiLi2EE"),
    ("kd_spmv<2, int, MINRES> (K4, peer memory)", r"7kd_spmvILi2EiLi2EE"),
    ("kd_mr_d (K4)", r"7kd_mr_dE"),
    ("kd_cg_p (K4)", r"7kd_cg_pE"),
    ("k_cg_xr (K3)", r"7k_cg_xrE"),
    ("k_rowins<4, RANGE, 24> (symbolic, P1 rows in 3D)", r"k_rowinsILi4ELb1ELi24EE"),
    ("k_rowhash<32, 1, 4, 128> (symbolic, round-1 kernel for those rows)", r"9k_rowhashILi32ELi1ELi4ELi128EE"),
    ("k_compact_ins<int> (symbolic)", r"k_compact_insIiE"),
    ("k_extract_fill<int> (system)", r"k_extract_fillIiE"),
]
KEY = re.compile(r"\b(LDG|STG|LDS|STS|ATOMS|ATOMG|ATOM|RED|UBLKCP|UTMALDG|SYNCS|LDGSTS|SHFL|DFMA|DADD|DMUL|MUFU|BAR|VIMNMX|MEMBAR|ERRBAR|CCTL)\S*")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
chunks = re.split(r"\n\s*Function : ", sass)
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), summarised by scripts/sass_excerpts.py\n")
for title, pat in WANT:
    body = next((c for c in chunks if re.search(pat, c.split("\n", 1)[0])), None)
    if body is None:
        print(f"## {title}: not found\n")
        continue
    lines = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m and KEY.match(m.group(1)):
            ops[m.group(1)] += 1
    print(f"## {title}\n   {body.split(chr(10), 1)[0][:110]}\n   {len(lines)} instructions")
    for op, n in sorted(ops.items(), key=lambda t: (-t[1], t[0])):
        print(f"   {n:5d}  {op}")
    print()
