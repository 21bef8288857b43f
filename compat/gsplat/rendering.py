"""`gsplat.rendering` stand-in (reference import: gaussian_renderer/__init__.py:15): the two
operators the MoBGS renderer uses, backed by libmobgs_b200.so.  There is no other backend."""
from mobgs_b200.rendering import fully_fused_projection, rasterization  # noqa: F401
