"""mmengine.Config.fromfile stand-in (train.py:978): exec a python config with `_base_`."""
import os


class Config(dict):
    @staticmethod
    def fromfile(path):
        ns = {}
        with open(path) as f:
            exec(compile(f.read(), path, "exec"), ns)
        out = Config()
        base = ns.get("_base_")
        if base:
            out.update(Config.fromfile(os.path.join(os.path.dirname(path), base)))
        for k, v in ns.items():
            if k.startswith("_") or not isinstance(v, dict):
                continue
            merged = dict(out.get(k, {}))
            merged.update(v)
            out[k] = merged
        return out
