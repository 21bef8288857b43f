"""Import-only placeholder for `plyfile` (scene/gaussian_model.py:7, scene/dataset_readers.py:31);
PLY I/O is outside the hot path."""


class PlyData:  # pragma: no cover
    @staticmethod
    def read(*a, **k):
        raise NotImplementedError("plyfile is not installed in this image")


class PlyElement:  # pragma: no cover
    @staticmethod
    def describe(*a, **k):
        raise NotImplementedError("plyfile is not installed in this image")
