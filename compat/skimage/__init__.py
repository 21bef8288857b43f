"""Import-only placeholder for `skimage` (the vendored LPIPS package imports it at module top,
models/__init__.py:7, models/dist_model.py:16; metrics.py uses its SSIM / PSNR — not on the hot path).
Used only when the real package is absent."""
from . import color, measure, metrics, transform  # noqa: F401
