"""Import-only placeholder (main_utils.py:3, scene/regulation.py:5, utils/scene_utils.py:4)."""
rcParams = {}


class _Any:
    def __getattr__(self, name):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


cm = _Any()
