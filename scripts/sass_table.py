"""`cuobjdump -sass libnbx.so` -> per-kernel instruction count and the counts of the mnemonics the design claims rest on
(profiles/r02_sass_hot_kernels.txt).  usage: python scripts/sass_table.py [path/to/libnbx.so] > profiles/r02_sass_hot_kernels.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "numbotics_b200/_lib/libnbx.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = ["UBLKCP", "LDGSTS", "STG", "LDG", "LDS", "STS", "FFMA", "FMUL", "FADD", "DFMA", "MUFU", "ATOMG", "RED", "BAR", "SHFL", "VOTE", "LDL", "STL", "CALL", "PREEXIT", "ACQBULK"]
rows, cur, k = [], None, -1
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = [names[k], collections.Counter(), 0]
        rows.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur[2] += 1
        op = m.group(1)
        for c in COLS:
            if op == c or op.startswith(c + "."):
                cur[1][c] += 1


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    return re.sub(r"\(.*", "", n)


keep = lambda n: not (("fkjac_wide_kernel<" in n and "<30," not in n) or ("fkjac_dense_kernel<" in n and ", 7, 128" not in n)
                      or "fkjac_generic" in n)
print("# SASS evidence for the hot kernels (cuobjdump -sass numbotics_b200/_lib/libnbx.so, sm_100a, built at this commit; scripts/sass_table.py)")
print("# Per kernel: instruction count, and the counts of the mnemonics that matter for the design claims:")
print("#   UBLKCP  = cp.async.bulk (TMA bulk copy shared -> global: the output tiles of the FK/J kernels leave the SM this way)")
print("#   LDGSTS  = cp.async (global -> shared without register staging: q tiles, scene arrays, edge windows)")
print("#   STG     = ordinary global stores (ragged / unaligned fallbacks, queue entries and small outputs only)")
print("#   FFMA/FMUL/FADD = the FP32 pipe; DFMA = float64 kernels; MUFU = special function unit; LDL/STL = local memory; CALL = out-of-line functions")
print("#   PREEXIT / ACQBULK = griddepcontrol.launch_dependents / griddepcontrol.wait (programmatic dependent launch of the verdict queries)")
print("#   no UTCHMMA / UTCQMMA / TMEM instructions anywhere: 3x4 rigid chains are not a dense contraction (north_star)")
print("#   (wide kernels other than 30 joints, dense kernels other than 7 joints and the legacy generic kernels are left out)")
print()
print(f"{'kernel':70s} {'instrs':>7s} " + " ".join(f"{c:>7s}" for c in COLS))
for name, cnt, n in sorted(rows, key=lambda r: short(r[0])):
    if keep(name):
        print(f"{short(name)[:70]:70s} {n:7d} " + " ".join(f"{cnt[c]:7d}" for c in COLS))
tensor = len(re.findall(r"UTC[A-Z]*MMA|\bLDTM\b|\bSTTM\b|\b[HIQ]MMA\b", sass))
flush = len(re.findall(r"UTMACMDFLUSH", sass))
print(f"\n# tensor-core / TMEM mnemonics (UTC*MMA, LDTM, STTM, HMMA / IMMA / QMMA) found in the whole library: {tensor}")
print(f"# UTMACMDFLUSH (cp.async.bulk.commit_group, the bulk stores' group commit): {flush}")
