"""Static SASS instruction-class counts of the hot kernels of libadmmtv.so (cuobjdump -sass), written as CSV.
    python tools/sass_summary.py [lib] > profiles/<round>_sass_summary.csv
UTMALDG / UTMASTG = TMA tensor tile load / store (cp.async.bulk.tensor), SYNCS = mbarrier operations, UBLKPF = TMA bulk L2
prefetch, FADD2 / FFMA2 / FMUL2 = packed dual-fp32 arithmetic, UCGABAR_* = cluster barrier (barrier.cluster.arrive / wait;
the distributed-shared-memory accesses of k_dim2c are the LD / ST through the mapa'd window, counted under "total")."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "admm_deconv_b200", "libadmmtv.so")
HOT = re.compile(r"k_dim1_fwd(_tma)?<(7|8|9|11), |k_dim1_bwd(_tma)?<(8|9), |k_dim2t?<(9|11), |k_small<7, 7|k_pack_fft1(_tma)?<9, 0|k_dim1_out(_tma)?<9, 1|"
                 r"k_dim1_bwd_last(_tma)?<9, 0|k_dim2c<12, 0|k_gmsd_(fwd|bwd)_s|k_ssim_(fwd|bwd)4<11")
CLASSES = ["UTMALDG", "UTMASTG", "SYNCS", "UBLKPF", "FADD2", "FFMA2", "FMUL2", "FADD", "FFMA", "FMUL", "LDG", "STG", "LDS", "STS", "RED",
           "ATOMG", "BAR", "WARPSYNC", "SHFL", "DADD", "LDGSTS", "UCGABAR_ARV", "UCGABAR_WAIT", "MUFU"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = {}
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        mangled = m.group(1)
        if mangled not in names:
            names[mangled] = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"^void admmtv::|\(.*$|admmtv::", "", names[mangled])
        cur = d if HOT.search(d) else None
        if cur:
            counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        if op in CLASSES:
            counts[cur][op] += 1
print("# SASS instruction-class counts of the hot kernels in admm_deconv_b200/libadmmtv.so (cuobjdump -sass, static counts; tools/sass_summary.py)")
print("# UTMALDG / UTMASTG = TMA tensor tile load / store (cp.async.bulk.tensor), SYNCS = mbarrier ops, UBLKPF = TMA bulk L2 prefetch, "
      "FADD2/FFMA2/FMUL2 = packed dual-fp32")
print("kernel," + ",".join(CLASSES) + ",total")
for k in sorted(counts):
    print('"' + k + '",' + ",".join(str(counts[k][c]) for c in CLASSES) + "," + str(counts[k]["total"]))
