"""placeholder for skimage.color (vendored LPIPS imports it at module top) — import only"""
