"""SASS instruction counts per kernel of libsdb200.so (cuobjdump -sass, sm_100a) -> profiles/r02_sass_summary.txt.
tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld/st -> LDTM/STTM, TMA loads -> UTMALDG, TMA stores -> UTMASTG,
L2 prefetch -> UTMAPF, tcgen05.commit -> UTCBAR, packed fp32x2 -> FFMA2/FADD2.   usage: python tools/sass_summary.py"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "stable-diffusion-from-scratch_b200", "libsdb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", ".2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "MUFU.EX2", "FFMA2", "FADD2", "HMMA", "LDG", "STG"]
rows, cur, cnt, n = [], None, None, 0
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, n, cnt))
        cur, cnt, n = m.group(1), dict.fromkeys(cols, 0), 0
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        n += 1
        for c in cols:
            if c == ".2CTA":
                if "UTCHMMA" in line and ".2CTA" in line:
                    cnt[c] += 1
            elif c == "HMMA":
                if re.search(r"\bHMMA", line):          # legacy mma.sync (must stay 0): UTCHMMA does not count
                    cnt[c] += 1
            elif c in line:
                cnt[c] += 1
if cur:
    rows.append((cur, n, cnt))
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
with open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w") as f:
    f.write("# SASS instruction counts per kernel of libsdb200.so (cuobjdump -sass, sm_100a), round 2 — tools/sass_summary.py\n")
    f.write("# tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld/st -> LDTM/STTM, TMA loads -> UTMALDG, TMA STORES -> UTMASTG, L2 prefetch -> UTMAPF, tcgen05.commit -> UTCBAR\n")
    f.write("%-96s %7s " % ("kernel", "instrs") + " ".join("%8s" % c for c in cols) + "\n")
    for (name, n, cnt), d in zip(rows, dem):
        d = re.sub(r"\(.*", "", d).replace("sdb::", "")
        f.write("%-96s %7d " % (d[:96], n) + " ".join("%8d" % cnt[c] for c in cols) + "\n")
print("kernels:", len(rows))
