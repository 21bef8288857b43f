"""The C-ABI library loads and exports every symbol include/mobgs_b200.h declares, and the ctypes
mirror of every argument struct has the size gcc computes from the header.  No compute calls (CPU)."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "mobgs_b200.h")


def _declared_functions():
    src = open(HDR).read()
    return sorted(set(re.findall(r"\b(mobgs_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_plain_c():
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HDR], check=True)


def test_every_declared_symbol_is_exported_and_bound():
    from mobgs_b200 import _lib
    path = _lib.build()
    lib = ctypes.CDLL(path)
    names = _declared_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by {path}"
        assert n in _lib.ENTRY_POINTS or n in ("mobgs_subframe_mean", "mobgs_knn3_mean_dist2"), f"{n} has no ctypes binding in mobgs_b200/_lib.py"
    assert b"sm_100a" in _lib.load().mobgs_version()


def test_ctypes_struct_sizes_match_header(tmp_path):
    from mobgs_b200 import _lib
    pairs = {"MobgsCameras": _lib.Cameras, "MobgsProjectFwd": _lib.ProjectFwd, "MobgsProjectBwd": _lib.ProjectBwd,
             "MobgsStaticParams": _lib.StaticParams, "MobgsDynamicParams": _lib.DynamicParams,
             "MobgsSynthFwd": _lib.SynthFwd, "MobgsSynthBwd": _lib.SynthBwd, "MobgsPack": _lib.Pack,
             "MobgsTileCount": _lib.TileCount, "MobgsTileSort": _lib.TileSort, "MobgsBlendFwd": _lib.BlendFwd,
             "MobgsBlendBwd": _lib.BlendBwd, "MobgsDecodeFwd": _lib.DecodeFwd, "MobgsDecodeBwd": _lib.DecodeBwd,
             "MobgsHexMlpFwd": _lib.HexMlpFwd, "MobgsLists": _lib.Lists,
             "MobgsFlowRecFwd": _lib.FlowRecFwd, "MobgsFlowRecBwd": _lib.FlowRecBwd,
             "MobgsHexFeat": _lib.HexFeat}
    for name in re.findall(r"\}\s*(Mobgs[A-Za-z]+)\s*;", open(HDR).read()):
        assert name in pairs or hasattr(_lib, "EXTRA_STRUCTS") and name in _lib.EXTRA_STRUCTS, \
            f"struct {name} has no ctypes mirror listed in this test"
    pairs.update(getattr(_lib, "EXTRA_STRUCTS", {}))
    prog = '#include <stdio.h>\n#include "mobgs_b200.h"\nint main(void){\n'
    for name in pairs:
        prog += f'  printf("{name} %zu\\n", sizeof({name}));\n'
    prog += "  return 0;\n}\n"
    c = tmp_path / "sz.c"
    c.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert ctypes.sizeof(pairs[name]) == int(size), (name, ctypes.sizeof(pairs[name]), size)


def test_product_has_no_cpu_fallback():
    """ops refuse CPU tensors instead of routing around the CUDA library."""
    import pytest
    import torch
    from mobgs_b200 import ops
    with pytest.raises(RuntimeError):
        ops.project(torch.zeros(1, 3), torch.ones(1, 4), torch.ones(1, 3), torch.eye(4)[None], torch.eye(3)[None], 8, 8)
    # nothing under mobgs_b200/ may import the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mobgs_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
