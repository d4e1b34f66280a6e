"""SASS opcode census of every object of libmtts (what proves a Blackwell-native kernel, B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk,
HMMA/IMMA/DMMA = mma.sync, LDGSTS = cp.async, REDUX = warp reductions, SYNCS = mbarrier.
Run here (no GPU needed): python scripts/sass_census.py > profiles/r02_sass_census.txt"""
import os, re, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "moss-ttsd_b200", "build")
PAT = collections.OrderedDict([
    ("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
    ("UTCCP/UTCBAR (tcgen05.cp/commit)", r"\bUTC(CP|BAR)"), ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"),
    ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (mma.sync f16/bf16/tf32)", r"\bHMMA"),
    ("LDSM (ldmatrix)", r"\bLDSM"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("REDUX", r"\bREDUX"), ("FFMA", r"\bFFMA"),
])
print("object".ljust(18) + "".join(k.split(" ")[0].rjust(10) for k in PAT) + "   kernels")
for f in sorted(os.listdir(BUILD)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, f)], capture_output=True, text=True).stdout
    kernels = len(re.findall(r"^\s*Function : ", sass, flags=re.M))
    print(f.ljust(18) + "".join(str(len(re.findall(p, sass))).rjust(10) for p in PAT.values()) + f"   {kernels}")
print()
for k in PAT:
    print(" ", k)
