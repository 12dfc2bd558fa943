"""Print a window of the ncu SASS source page: address, samples, inst executed, top stall reasons, smem wavefronts.
usage: ncu_sass.py <sass.csv> <start_hex_offset> <end_hex_offset>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
ix = {n: H.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal")}
stall = [i for i, n in enumerate(H) if n.startswith("stall_") and "Not Issued" not in n]
base = int(rows[hdr + 1][0], 16)
tot = 0
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    off = int(r[0], 16) - base
    if off < lo or off > hi: continue
    s = int(r[ix["# Samples"]] or 0); tot += s
    st = sorted(((int(r[i] or 0), H[i][6:]) for i in stall), reverse=True)[:3]
    print("%6x %-52s smp %6d inst %9s wf %8s/%8s  %s" % (off, r[ix["Source"]].strip()[:52], s, r[ix["Instructions Executed"]],
          r[ix["L1 Wavefronts Shared"]], r[ix["L1 Wavefronts Shared Ideal"]], " ".join("%s:%d" % (n, v) for v, n in st if v)))
print("window samples", tot)
