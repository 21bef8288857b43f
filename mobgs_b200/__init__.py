"""mobgs_b200 — B200-native render + deblur hot path for KAIST-VICLab/MoBGS.

Python here is host plumbing over libmobgs_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/mobgs_b200.h).  See DESIGN.md.
"""
__version__ = "0.1"
