"""Print the hottest SASS instructions (by warp-stall samples) of an exported `ncu --page source --print-source sass --csv` file,
with the dominant stall reason.  usage: ncu_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # first captured launch only
    if len(r) == len(hdr):
        body.append(r)
total = sum(int(r[isamp] or 0) for r in body)
print("total samples", total, "instructions", len(body))
ranked = sorted(enumerate(body), key=lambda t: -int(t[1][isamp] or 0))[:top]
for idx, r in sorted(ranked):
    st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% %-70s %s" % (idx, 100.0 * int(r[isamp]) / max(total, 1), r[isrc].strip()[:70], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
