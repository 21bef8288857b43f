"""placeholder for skimage.metrics (metrics.py:6) — import only"""


def structural_similarity(*a, **k):
    raise NotImplementedError("skimage is not installed; this placeholder only satisfies the import")


def peak_signal_noise_ratio(*a, **k):
    raise NotImplementedError("skimage is not installed; this placeholder only satisfies the import")
