"""`gsplat.rendering` stand-in (reference import: gaussian_renderer/__init__.py:15)."""
import os

_BACKEND = os.environ.get("MOBGS_GSPLAT_BACKEND", "b200")

if _BACKEND == "oracle":   # test infrastructure only (golden generation / reference parity tests)
    from oracle.gsplat_ref import fully_fused_projection, rasterization  # noqa: F401
elif _BACKEND == "b200":
    from mobgs_b200.rendering import fully_fused_projection, rasterization  # noqa: F401
else:
    raise ImportError(f"unknown MOBGS_GSPLAT_BACKEND={_BACKEND!r}")
