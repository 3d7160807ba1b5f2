"""Instruction histogram per kernel of libtnb.so (cuobjdump -sass): python scratch/sass_hist.py > profiles/sass_r02.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "tncontract_b200", "lib", "libtnb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["DMMA", "DFMA", "DMUL", "DADD", "FFMA", "MUFU", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDS", "STS", "LDG", "STG",
       "BAR", "UCGABAR", "ACQBULK", "ATOM", "RED", "SHFL", "HMMA", "UTCHMMA", "LDTM"]
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["TOTAL"] += 1
        for k in KEY:
            if op == k or op.startswith(k + "."):
                hist[kern][k] += 1
demangle = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
print("# SASS instruction histogram per kernel, libtnb.so (sm_100a), `cuobjdump -sass` -- scratch/sass_hist.py")
print("# FP64 tensor = DMMA (mma.sync m8n8k4; tcgen05 has no f64 kind); TMA bulk copies = UBLKCP (+ SYNCS mbarrier ops);")
print("# cp.async = LDGSTS; tensor-map TMA = UTMALDG/UTMASTG")
for name, full in sorted(zip(demangle, hist), key=lambda x: -hist[x[1]]["TOTAL"]):
    h = hist[full]
    short = re.sub(r"\(.*", "", name)[:110]
    print("%-112s %s" % (short, " ".join("%s=%d" % (k, h[k]) for k in ["TOTAL"] + KEY if h[k])))
