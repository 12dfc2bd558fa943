"""Join an ncu SASS source page (csv) with nvdisasm line info -> stall samples per source line.

usage: ncu_lines.py <report.ncu-rep> <lib.so> [kernel-substring] [top]

The kernel substring must select ONE instance of the mangled name (e.g. ipm_solve_kernelILb1ELb1ELb1ELi1:
the two-blocks-per-SM build in ipm_kernels_small.cu carries ...ELi2 and would shadow it).
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "ipm_solve_kernel"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
    if kern not in txt:
        continue
    cur, infn = None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
        if m:
            infn = kern in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ci = {n: H.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed")}
stall_cols = [i for i, n in enumerate(H) if n.startswith("stall_") and "Not Issued" not in n]
base = None
per_line = collections.Counter(); per_line_stall = collections.defaultdict(collections.Counter)
per_file = collections.Counter(); inst_line = collections.Counter()
total = 0
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    addr = int(r[ci["Address"]], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base, (("?", 0), ""))[0]
    s = int(r[ci["# Samples"]] or 0)
    total += s
    per_line[key] += s
    per_file[key[0]] += s
    inst_line[key] += int(r[ci["Instructions Executed"]] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            per_line_stall[key][H[i]] += v
print("total samples", total)
for f, s in per_file.most_common():
    print("  %-22s %6.2f%%" % (f, 100.0 * s / total))
print("top lines:")
for key, s in per_line.most_common(top):
    st = ", ".join("%s %d%%" % (k.replace("stall_", ""), 100 * v // max(s, 1)) for k, v in per_line_stall[key].most_common(3))
    print("  %-20s:%4d  %6.2f%%  inst %10d  [%s]" % (key[0], key[1], 100.0 * s / total, inst_line[key], st))
