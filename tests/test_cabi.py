"""The C-ABI shared library loads and exports every symbol include/rt_cuda.h
declares; without a GPU the product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from ray_tracing_b200 import host


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rt_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M):
        names.add(m.group(1))
    return names


def test_library_loads_and_exports_all_declared_symbols():
    lib = host.load_library()
    names = declared_functions()
    assert "render_frame_cuda" in names and "rt_parse_scene_file" in names and len(names) >= 30
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(host.EXPORTED_SYMBOLS) <= names


def test_struct_layouts_match_reference_sizes():
    # Object 68 B, Scene 69 636 B with num_objects at 69 632
    assert host.OBJECT_DTYPE.itemsize == 68
    assert ctypes.sizeof(host.RtScene) == 69636
    assert host.RtScene.num_objects.offset == 69632
    assert ctypes.sizeof(host.RtCamera) == 40
    assert host.OBJECT_DTYPE.fields["geom"][1] == 4 and host.OBJECT_DTYPE.fields["albedo"][1] == 28


def test_no_oracle_in_product():
    """The product never imports, links or loads anything under oracle/."""
    pkg = os.path.join(ROOT, "ray_tracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "rt_oracle" not in src and "libref_" not in src, f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
    import subprocess

    needed = subprocess.run(["ldd", host.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="this check is for boxes without a GPU")
def test_fails_loudly_without_gpu():
    with pytest.raises(host.RtError) as e:
        host.Renderer(num_gpus=1)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)
    import numpy as np

    sc = host.make_scene(host.parse_scene_string("sphere"))
    cam = host.Camera().as_struct()
    fb = np.zeros((4, 4, 3), np.float32)
    rc = host.load_library().render_frame_cuda(ctypes.byref(sc), ctypes.byref(cam), fb.ctypes.data, 4, 4, 1)
    assert rc == -1


def test_missing_library_fails_loudly(monkeypatch):
    """No pure-Python or CPU rendering path: without the built .so the binding raises."""
    monkeypatch.setattr(host, "_lib", None)
    monkeypatch.setattr(host, "LIB_PATH", os.path.join(ROOT, "ray_tracing_b200", "does_not_exist.so"))
    with pytest.raises(ImportError) as e:
        host.load_library()
    assert "no pure-Python or CPU rendering path" in str(e.value)
    with pytest.raises(ImportError):
        host.Renderer(num_gpus=1)


def test_bvh_builder_knob_validates_its_argument():
    """rt_cuda_set_bvh_builder needs no device: it only says who shapes the tree of the next upload."""
    lib = host.load_library()
    assert lib.rt_cuda_set_bvh_builder(host.RT_BVH_BUILDER_LBVH) == 0
    assert lib.rt_cuda_set_bvh_builder(host.RT_BVH_BUILDER_SAH) == 0
    assert lib.rt_cuda_set_bvh_builder(7) == -3            # RT_ERR_ARG
    assert b"builder" in lib.rt_cuda_last_error()
