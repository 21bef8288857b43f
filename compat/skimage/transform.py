"""placeholder for skimage.transform (vendored LPIPS imports it at module top) — import only"""
