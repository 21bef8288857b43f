"""Minimal `imageio` stand-in on top of PIL: the reference reads PNGs with `imageio.imread` (eval.py:89,
scene/dataset_readers.py) and nothing else on the Stereo-Blur path.  Used only when the real package is absent."""
import numpy as np
from PIL import Image


def imread(uri, *args, **kwargs):
    with Image.open(uri) as im:
        return np.asarray(im)


def imwrite(uri, im, *args, **kwargs):
    Image.fromarray(np.asarray(im)).save(uri)


imsave = imwrite


class _V2:
    imread = staticmethod(imread)
    imwrite = staticmethod(imwrite)


v2 = _V2()
