"""SASS opcode histogram per kernel of the built library -> profiles/sass_rNN.txt
usage: sass_hist.py [lib.so] > profiles/sass_r02.txt"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                           "pycllp_b200", "libpycllp_b200.so")
out = subprocess.run("cuobjdump -sass %s | c++filt" % lib, shell=True, stdout=subprocess.PIPE, text=True).stdout
print("# SASS opcode histogram of pycllp_b200/libpycllp_b200.so (cuobjdump -sass, sm_100a), round-2 final build")
print("# FP64 has no tcgen05 kind: the tensor path is DMMA.8x8x4; UBLKCP = cp.async.bulk (TMA 1-D), SYNCS = mbarrier, "
      "LDGSTS = cp.async\n")
BASE = "DMMA DFMA DMUL DADD MUFU UBLKCP SYNCS LDGSTS LDS STS LDG STG LD ST LDL STL BAR SHFL REDUX CREDUX ATOMS ATOMG CALL".split()
FULL = ("ATOMS", "BAR", "DMMA", "LDGSTS", "MUFU", "SYNCS", "UBLKCP")
name, base, full, n = None, collections.Counter(), collections.Counter(), 0


def flush():
    if name and n:
        print(name[:110])
        print("  instructions", n)
        print("  " + "  ".join("%s %d" % (k, base[k]) for k in BASE if base[k]))
        print("  " + "  ".join("%s %d" % (k, v) for k, v in sorted(full.items())) + "\n")


for ln in out.splitlines():
    m = re.match(r"\s*Function : (.*)", ln)
    if m:
        flush()
        name, base, full, n = m.group(1), collections.Counter(), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        n += 1
        b = op.split(".")[0]
        base[b] += 1
        if b in FULL:
            full[".".join(op.split(".")[:4])] += 1
flush()
