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
    "n11_512": ("ADMMTV_NT11=512",),
    "n11_512_t10": ("ADMMTV_NT11=512", "ADMMTV_TC11=10"),
    "n11_256_t10": ("ADMMTV_NT11=256", "ADMMTV_TC11=10"),
    "n2max256": ("ADMMTV_NT2_MAX=256",),
    "tr11_4": ("ADMMTV_TR11=4",),
    "c8_4": ("ADMMTV_CHUNK8=4",),
    "c8_2": ("ADMMTV_CHUNK8=2",),
    "c8_4_t10": ("ADMMTV_CHUNK8=4", "ADMMTV_TC8=10"),
    "c8_8_t10": ("ADMMTV_CHUNK8=8", "ADMMTV_TC8=10"),
    "c8_4_t34": ("ADMMTV_CHUNK8=4", "ADMMTV_TC8=34"),
    "pfn0": ("ADMMTV_PF_NEXT=0",),
    "pfn296": ("ADMMTV_PF_NEXT=296",),
    "pfn592": ("ADMMTV_PF_NEXT=592",),
    "pre1": ("ADMMTV_PRELOAD1=1",),
    "pre1_mb2": ("ADMMTV_PRELOAD1=1", "ADMMTV_MINB9=2"),
    "isonoatom": ("ADMMTV_ISO_NOATOM=1",),
    "bh1": ("ADMMTV_BWD_HOIST=1",),
    "bh1_mb2": ("ADMMTV_BWD_HOIST=1", "ADMMTV_MINB9B=2"),
    "bh1_mb3": ("ADMMTV_BWD_HOIST=1", "ADMMTV_MINB9B=3"),
    "bh0_mb3": ("ADMMTV_BWD_HOIST=0", "ADMMTV_MINB9B=3"),
    "qpb1": ("ADMMTV_D2_ACC_QPB=1",),
    "qpb2": ("ADMMTV_D2_ACC_QPB=2",),
    "qpb8": ("ADMMTV_D2_ACC_QPB=8",),
    "d2n512": ("ADMMTV_NT2=512",),
    "d2n512mb2": ("ADMMTV_NT2=512", "ADMMTV_MINB2=2"),
    "d2n128": ("ADMMTV_NT2=128",),
    "d2n128mb6": ("ADMMTV_NT2=128", "ADMMTV_MINB2=6"),
    "d2mb3": ("ADMMTV_MINB2=3",),
    "n512_t18_c4": ("ADMMTV_NT9=512", "ADMMTV_TC9=18", "ADMMTV_CHUNK9=4", "ADMMTV_MINB9=2", "ADMMTV_MINB9B=1"),
    "n512_t10_c4": ("ADMMTV_NT9=512", "ADMMTV_TC9=10", "ADMMTV_CHUNK9=4", "ADMMTV_MINB9=2", "ADMMTV_MINB9B=1"),
    "n512_t10_c8": ("ADMMTV_NT9=512", "ADMMTV_TC9=10", "ADMMTV_CHUNK9=8", "ADMMTV_MINB9=2", "ADMMTV_MINB9B=1"),
    "n512_t10_c2": ("ADMMTV_NT9=512", "ADMMTV_TC9=10", "ADMMTV_CHUNK9=2", "ADMMTV_MINB9=3", "ADMMTV_MINB9B=1"),
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
    SIZES = tuple(int(a[2:]) for a in sys.argv[1:] if a.startswith("-s")) or (9,)
    names = [a for a in sys.argv[1:] if not a.startswith("-s")] or list(VARIANTS)
    for n in names:
        print(B.build(tag=n, defines=VARIANTS[n], sizes=SIZES))
