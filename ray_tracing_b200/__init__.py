"""ray_tracing_b200 -- B200-native (sm_100a) render path for cozis/ray_tracing.

The product is the C-ABI library libraytrace_b200.so (include/rt_cuda.h,
ray_tracing_b200/csrc/); this package is its ctypes face plus scene helpers.
There is no CPU rendering path: `host.load_library()` raises if the CUDA
library has not been built.
"""
from . import host, scenes  # noqa: F401
from .host import (  # noqa: F401
    Camera,
    Renderer,
    RtError,
    parse_scene_file,
    parse_scene_string,
    parse_scene_file_large,
    parse_scene_string_large,
    move_camera,
    rotate_camera,
    get_camera_pos,
    camera_snapshot,
    camera_reset,
    quantize_frame,
)

__all__ = ["host", "scenes", "Camera", "Renderer", "RtError"]
