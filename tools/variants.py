"""Builds tuning variants of the library (extra -D flags, size 512 only) as
admm_deconv_b200/libadmmtv_<tag>.so.  Used with tools/microbench.py on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_deconv_b200 import build as B

VARIANTS = {
    "base": ("ADMMTV_SWZ=0", "ADMMTV_PREFETCH=0"),
    "swz": ("ADMMTV_SWZ=1", "ADMMTV_PREFETCH=0"),
    "swz_pf": ("ADMMTV_SWZ=1", "ADMMTV_PREFETCH=1"),
    "swz_pf_u2": ("ADMMTV_SWZ=1", "ADMMTV_PREFETCH=1", "ADMMTV_UNROLL_ITEMS=2"),
    "tc10": ("ADMMTV_SWZ=1", "ADMMTV_PREFETCH=1", "ADMMTV_TC9=10"),
    "tc10_u2": ("ADMMTV_SWZ=1", "ADMMTV_PREFETCH=1", "ADMMTV_TC9=10", "ADMMTV_UNROLL_ITEMS=2"),
    "t18p0": ("ADMMTV_TC9=18",),
    "t18p1": ("ADMMTV_TC9=18", "ADMMTV_PREFETCH=1"),
    "t10p1": ("ADMMTV_TC9=10", "ADMMTV_PREFETCH=1"),
    "t10p1c2": ("ADMMTV_TC9=10", "ADMMTV_PREFETCH=1", "ADMMTV_CHUNK9=2"),
    "t10p1n128": ("ADMMTV_TC9=10", "ADMMTV_PREFETCH=1", "ADMMTV_NT9=128"),
    "t10p1n128c4": ("ADMMTV_TC9=10", "ADMMTV_PREFETCH=1", "ADMMTV_NT9=128", "ADMMTV_CHUNK9=4"),
    "t10p0n128": ("ADMMTV_TC9=10", "ADMMTV_PREFETCH=0", "ADMMTV_NT9=128"),
    "t6p1": ("ADMMTV_TC9=6", "ADMMTV_PREFETCH=1"),
    "t6p1n128": ("ADMMTV_TC9=6", "ADMMTV_PREFETCH=1", "ADMMTV_NT9=128"),
    "t18p1n128": ("ADMMTV_TC9=18", "ADMMTV_PREFETCH=1", "ADMMTV_NT9=128"),
    "d2p0": ("ADMMTV_D2_PERSIST=0",),
    "d2pa": ("ADMMTV_D2_PERSIST=2",),
    "d2pa3": ("ADMMTV_D2_PERSIST=2", "ADMMTV_D2_BLOCKS_PER_SM=3"),
    "d2p2": ("ADMMTV_D2_BLOCKS_PER_SM=2",),
    "d2p3": ("ADMMTV_D2_BLOCKS_PER_SM=3",),
    "d2p3mb3": ("ADMMTV_D2_BLOCKS_PER_SM=3", "ADMMTV_MINB2=3"),
    "d2p4mb4": ("ADMMTV_D2_BLOCKS_PER_SM=4", "ADMMTV_MINB2=4"),
    "c2mb3": ("ADMMTV_CHUNK9=2", "ADMMTV_MINB9=3"),
    "c4mb3": ("ADMMTV_CHUNK9=4", "ADMMTV_MINB9=3"),
    "c2mb4": ("ADMMTV_CHUNK9=2", "ADMMTV_MINB9=4"),
    "t10": ("ADMMTV_TC9=10",),
    "t6": ("ADMMTV_TC9=6",),
    "t10_n128": ("ADMMTV_TC9=10", "ADMMTV_NT9=128"),
    "t10_mb5": ("ADMMTV_TC9=10", "ADMMTV_MINB9=5"),
    "t18_n128": ("ADMMTV_TC9=18", "ADMMTV_NT9=128"),
    "t10_r8": ("ADMMTV_TC9=10", "ADMMTV_TR9=8"),
    "t10_r8_n128": ("ADMMTV_TC9=10", "ADMMTV_TR9=8", "ADMMTV_NT2=128"),
    "t10_r16_n128": ("ADMMTV_TC9=10", "ADMMTV_NT2=128"),
    "t10_r16_mb4": ("ADMMTV_TC9=10", "ADMMTV_MINB2=4"),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(B.build(tag=n, defines=VARIANTS[n], sizes=(9,)))
