"""Import-only placeholder for `lpips` (not on the hot path)."""


class EasyDict(dict):
    __getattr__ = dict.get


def __getattr__(name):
    raise AttributeError(name)
