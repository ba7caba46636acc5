"""Count the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA, TMEM loads, TMA loads / stores / reductions, bulk copies,
cluster barriers, packed fp32x2 FMA) per built object.  Run in the build container:  python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import glob
import os
import re
import subprocess

PAT = re.compile(r"\b(UTCHMMA[.\w]*|UTCBAR[.\w]*|LDTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAREDG[.\w]*|UBLKCP[.\w]*|UCGABAR[.\w]*|"
                 r"STAS[.\w]*|FFMA2|HMMA[.\w]*|LDSM[.\w]*|LDGSTS[.\w]*|SYNCS[.\w]*)")
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "molnextr_b200", "lib")
print("# SASS mnemonic counts per object (cuobjdump -sass, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit, `UTMALDG` = TMA tensor load, `UTMASTG` = TMA tensor store, "
      "`UTMAREDG` = TMA tensor reduce, `UBLKCP` = 1-D bulk copy, `UCGABAR` = cluster barrier, `STAS` = st.async (DSMEM), "
      "`FFMA2` = packed fp32x2 FMA, `HMMA` = mma.sync, `LDSM` = ldmatrix, `LDGSTS` = cp.async, `SYNCS` = mbarrier ops.\n")
print("| object | mnemonic: count |")
print("|---|---|")
for obj in sorted(glob.glob(os.path.join(root, "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    c = collections.Counter()
    for m in PAT.finditer(out):
        name = m.group(1)
        key = name if name.startswith(("UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "HMMA")) else name.split(".")[0]
        c[key] += 1
    if c:
        print(f"| `{os.path.basename(obj)}` | " + ", ".join(f"{k}: {v}" for k, v in sorted(c.items())) + " |")
