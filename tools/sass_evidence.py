"""SASS evidence of the built library: per kernel, the instruction total and the counts of the mnemonics that prove the
hardware paths.   usage: python tools/sass_evidence.py > profiles/<tag>_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mobgs_b200", "libmobgs_b200.so")], capture_output=True, text=True).stdout
KEEP = ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "MUFU", "RED", "REDG", "ATOM", "ATOMG", "ATOMS", "SHFL", "STL", "LDL")
WANT = ("blend_fwd_kernel<10, true", "blend_bwd_tr_kernel<10, true", "hexplane", "tile_", "synth_project", "scan_kernel", "adam",
        "photo", "flow_warp", "camera_rays", "compact", "reg_loss", "knn", "midflow", "flow_records", "decode", "subframe_mean")
print("SASS evidence for the committed build (cuobjdump -sass mobgs_b200/libmobgs_b200.so, sm_100a; tools/sass_evidence.py).  Per kernel: total")
print("instructions and the counts of the mnemonics that prove the hardware paths: UTCHMMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld")
print("(TMEM -> registers), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, REDG / ATOMG / ATOMS =")
print("reductions / atomics (global / shared), MUFU = ex2 / rcp / rsq / sqrt, SHFL = warp shuffles, STL / LDL = spill stores / loads.\n")
agg = collections.Counter()
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if not any(s in dem for s in WANT):
        continue
    ops = collections.Counter(m.split(".")[0] for m in re.findall(r"^\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, flags=re.M))
    tot = sum(ops.values())
    for k in KEEP:
        agg[k] += ops.get(k, 0)
    short = dem.split("(")[0].replace("void mobgs::", "").replace("mobgs::", "")
    print(f"{short:58s} {tot:6d} instr  " + "  ".join(f"{k} {ops[k]}" for k in KEEP if ops.get(k)))
print("\nwhole library (kernels listed): " + "  ".join(f"{k} {agg[k]}" for k in ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS")))
