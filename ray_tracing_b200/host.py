"""ctypes binding of libraytrace_b200.so (include/rt_cuda.h).

This is the Python face of the C ABI -- plumbing for tests, bench.py and the
torchrun multi-GPU launcher.  The names mirror the reference's host functions
(parse_scene_file, move_camera, rotate_camera, get_camera_pos; src/scene.h:47,
src/camera.h:25-29) and the north_star entry point render_frame_cuda.

There is no CPU fallback: if the shared library is missing, importing the
binding raises; if no CUDA device is usable every render call raises RtError.
Nothing here imports anything from oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libraytrace_b200.so")

RT_MAX_OBJECTS = 1024
RT_OBJECT_CUBE, RT_OBJECT_SPHERE = 0, 1
RT_FB_F32X3, RT_FB_U8X4 = 0, 1
RT_VARIANT_EXACT, RT_VARIANT_FAST = 0, 1
RT_TRAVERSAL_AUTO, RT_TRAVERSAL_LINEAR, RT_TRAVERSAL_LBVH = 0, 1, 2
RT_MEM_AUTO, RT_MEM_HOST, RT_MEM_DEVICE = 0, 1, 2
RT_KERNEL_AUTO, RT_KERNEL_PIXEL, RT_KERNEL_PERSISTENT, RT_KERNEL_WAVEFRONT, RT_KERNEL_QUEUED = 0, 1, 2, 3, 4
RT_BVH_BUILDER_SAH, RT_BVH_BUILDER_LBVH = 0, 1
RT_UP, RT_DOWN, RT_LEFT, RT_RIGHT = 0, 1, 2, 3
RT_LBVH_THRESHOLD = 64

# byte-compatible with the reference `Object` (scene.h:24-31, 68 B)
OBJECT_DTYPE = np.dtype(
    [
        ("type", "<i4"),
        ("geom", "<f4", (6,)),
        ("albedo", "<f4", (3,)),
        ("roughness", "<f4"),
        ("reflectance", "<f4"),
        ("metallic", "<f4"),
        ("emission_power", "<f4"),
        ("emission_color", "<f4", (3,)),
    ]
)
assert OBJECT_DTYPE.itemsize == 68


class RtVector3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class RtCamera(C.Structure):
    _fields_ = [("pos", RtVector3), ("front", RtVector3), ("up", RtVector3), ("fov", C.c_float)]


class RtScene(C.Structure):
    _fields_ = [("objects", C.c_uint8 * (68 * RT_MAX_OBJECTS)), ("num_objects", C.c_int)]


class RtCubemap(C.Structure):
    _fields_ = [("data", C.c_void_p * 6), ("w", C.c_int), ("h", C.c_int), ("chan", C.c_int)]


class RtRenderOpts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("scale", C.c_int),
        ("num_columns", C.c_int),
        ("pass_index", C.c_uint64),
        ("fb_format", C.c_int),
        ("fb_memory", C.c_int),
        ("row_begin", C.c_int),
        ("row_end", C.c_int),
        ("accumulate", C.c_int),
        ("variant", C.c_int),
        ("traversal", C.c_int),
        ("kernel", C.c_int),
        ("band_only_fb", C.c_int),
        ("stream", C.c_void_p),
        ("interleave_count", C.c_int),
        ("interleave_index", C.c_int),
        ("pipeline", C.c_int),
        ("remote_fb", C.c_int),
        ("frame_seq", C.c_uint32),
        ("frame_ack", C.c_int),
    ]


class RtRenderStats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("pixels", C.c_uint64),
        ("render_ms", C.c_float),
        ("composite_ms", C.c_float),
        ("copy_ms", C.c_float),
        ("kernel_launches", C.c_int),
    ]


assert C.sizeof(RtScene) == 69636  # scene.h:33-36


class RtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rt_cuda error {code}: {message}")
        self.code = code


# every symbol include/rt_cuda.h declares (tests check the .so exports them all)
EXPORTED_SYMBOLS = [
    "rt_cuda_last_error",
    "rt_parse_scene_file",
    "rt_parse_scene_string",
    "rt_parse_scene_file_large",
    "rt_parse_scene_string_large",
    "rt_free_objects",
    "rt_camera_reset",
    "rt_move_camera",
    "rt_rotate_camera",
    "rt_get_camera_pos",
    "rt_camera_snapshot",
    "rt_quantize_frame",
    "rt_save_screenshot",
    "rt_cuda_init",
    "rt_cuda_init_device",
    "rt_cuda_shutdown",
    "rt_cuda_num_gpus",
    "rt_cuda_upload_scene",
    "rt_cuda_upload_objects",
    "rt_cuda_upload_skybox",
    "rt_render_opts_default",
    "render_frame_cuda",
    "render_frame_cuda_ex",
    "rt_cuda_accum_reset",
    "rt_cuda_accum_count",
    "rt_cuda_render_sweep",
    "rt_cuda_synchronize",
    "rt_cuda_debug_trace",
    "rt_cuda_debug_sample_cubemap",
    "rt_cuda_debug_camera_rays",
    "rt_cuda_debug_rng",
    "rt_cuda_debug_random_directions",
    "rt_pixel_key",
    "rt_cuda_debug_fp32_peak",
    "rt_cuda_debug_div_check",
    "rt_cuda_debug_set_sweep_threshold",
    "rt_cuda_debug_set_tile_schedule",
    "rt_cuda_debug_set_light_anyhit",
    "rt_cuda_set_bvh_builder",
    "rt_cuda_debug_set_concurrent_sweep",
    "rt_cuda_debug_set_sync_bands",
    "rt_cuda_debug_set_queued_dense",
    "rt_cuda_next_pass_index",
    "rt_cuda_param_bytes",
    "rt_cuda_set_progressive",
    "rt_cuda_invalidate_accumulation",
    "rt_cuda_accum_generation",
    "rt_cuda_update_frame",
    "rt_cuda_shared_frame_create",
    "rt_cuda_shared_frame_open",
    "rt_cuda_shared_frame_close",
    "rt_cuda_copy_to_host",
    "rt_cuda_gl_register_buffer",
    "rt_cuda_gl_update_frame",
    "rt_cuda_gl_render_frame",
    "rt_cuda_gl_unregister_buffer",
]

_lib = None


def load_library() -> C.CDLL:
    """Load the C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). "
            "ray_tracing_b200 has no pure-Python or CPU rendering path."
        )
    L = C.CDLL(LIB_PATH)
    L.rt_cuda_last_error.restype = C.c_char_p
    L.rt_parse_scene_file.restype = C.c_bool
    L.rt_parse_scene_file.argtypes = [C.c_char_p, C.POINTER(RtScene)]
    L.rt_parse_scene_string.restype = C.c_bool
    L.rt_parse_scene_string.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(RtScene)]
    L.rt_parse_scene_file_large.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.rt_parse_scene_string_large.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.rt_free_objects.argtypes = [C.c_void_p]
    L.rt_free_objects.restype = None
    L.rt_camera_reset.restype = None
    L.rt_move_camera.argtypes = [C.c_int, C.c_float]
    L.rt_move_camera.restype = None
    L.rt_rotate_camera.argtypes = [C.c_double, C.c_double]
    L.rt_rotate_camera.restype = None
    L.rt_get_camera_pos.restype = RtVector3
    L.rt_camera_snapshot.restype = RtCamera
    L.rt_quantize_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.rt_quantize_frame.restype = None
    L.rt_save_screenshot.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    L.rt_cuda_init.argtypes = [C.c_int]
    L.rt_cuda_init_device.argtypes = [C.c_int]
    L.rt_cuda_shutdown.restype = None
    L.rt_cuda_upload_scene.argtypes = [C.POINTER(RtScene)]
    L.rt_cuda_upload_objects.argtypes = [C.c_void_p, C.c_int]
    L.rt_cuda_update_objects.argtypes = [C.c_void_p, C.c_int]
    L.rt_cuda_upload_skybox.argtypes = [C.POINTER(RtCubemap)]
    L.rt_render_opts_default.argtypes = [C.POINTER(RtRenderOpts)]
    L.rt_render_opts_default.restype = None
    L.render_frame_cuda.argtypes = [C.POINTER(RtScene), C.POINTER(RtCamera), C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.render_frame_cuda_ex.argtypes = [C.POINTER(RtCamera), C.c_void_p, C.c_int, C.c_int, C.POINTER(RtRenderOpts), C.POINTER(RtRenderStats)]
    L.rt_cuda_accum_count.restype = C.c_float
    L.rt_cuda_render_sweep.argtypes = [C.POINTER(RtCamera), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(RtRenderOpts), C.POINTER(RtRenderStats)]
    L.rt_cuda_debug_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.rt_cuda_debug_sample_cubemap.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.rt_cuda_debug_camera_rays.argtypes = [C.POINTER(RtCamera), C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    L.rt_cuda_debug_rng.argtypes = [C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.rt_cuda_debug_random_directions.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
    L.rt_cuda_debug_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_float)]
    L.rt_cuda_debug_div_check.argtypes = [C.c_uint64, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.rt_cuda_shared_frame_create.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    L.rt_cuda_shared_frame_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.rt_cuda_shared_frame_close.argtypes = [C.c_void_p, C.c_int]
    L.rt_cuda_shared_frame_wait.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
    L.rt_cuda_shared_frame_release.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.rt_cuda_shared_frame_error.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.rt_cuda_copy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.rt_cuda_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.rt_cuda_debug_set_sweep_threshold.argtypes = [C.c_float]
    L.rt_cuda_debug_set_tile_schedule.argtypes = [C.c_int]
    L.rt_cuda_debug_set_light_anyhit.argtypes = [C.c_int]
    L.rt_cuda_set_bvh_builder.argtypes = [C.c_int]
    L.rt_cuda_debug_set_concurrent_sweep.argtypes = [C.c_int]
    L.rt_cuda_debug_set_sync_bands.argtypes = [C.c_int]
    L.rt_cuda_debug_set_queued_dense.argtypes = [C.c_int]
    L.rt_cuda_next_pass_index.restype = C.c_uint64
    L.rt_cuda_gl_register_buffer.argtypes = [C.c_uint, C.c_size_t]
    L.rt_cuda_gl_update_frame.argtypes = [C.POINTER(RtCamera), C.c_int, C.c_int, C.c_double, C.POINTER(RtRenderOpts), C.POINTER(RtRenderStats)]
    L.rt_cuda_gl_render_frame.argtypes = [C.POINTER(RtCamera), C.c_int, C.c_int, C.POINTER(RtRenderOpts), C.POINTER(RtRenderStats)]
    L.rt_cuda_debug_tile_order.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.rt_cuda_ray_counter.argtypes = [C.POINTER(C.c_uint64)]
    L.rt_cuda_param_bytes.restype = C.c_size_t
    L.rt_cuda_debug_walk_counts.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.rt_cuda_set_progressive.argtypes = [C.c_int, C.c_int]
    L.rt_cuda_accum_generation.restype = C.c_uint32
    L.rt_cuda_update_frame.argtypes = [C.POINTER(RtCamera), C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(RtRenderOpts), C.POINTER(RtRenderStats)]
    L.rt_pixel_key.restype = C.c_uint64
    L.rt_pixel_key.argtypes = [C.c_float, C.c_float, C.c_uint64]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise RtError(rc, load_library().rt_cuda_last_error().decode(errors="replace"))


# ----------------------------------------------------------------- host side


def parse_scene_file(path: str):
    """Reference `parse_scene_file` (scene.c:611-624): returns the object
    records (OBJECT_DTYPE array) or None when the reference would return false."""
    L = load_library()
    sc = RtScene()
    if not L.rt_parse_scene_file(os.fsencode(path), C.byref(sc)):
        return None
    return scene_objects(sc)


def parse_scene_string(text) -> "np.ndarray | None":
    L = load_library()
    data = text.encode() if isinstance(text, str) else bytes(text)
    sc = RtScene()
    if not L.rt_parse_scene_string(data, len(data), C.byref(sc)):
        return None
    return scene_objects(sc)


def parse_scene_string_partial(text):
    """(ok, objects parsed before the error) -- the reference leaves the partial
    count in scene->num_objects on failure (scene.c:208)."""
    L = load_library()
    data = text.encode() if isinstance(text, str) else bytes(text)
    sc = RtScene()
    ok = bool(L.rt_parse_scene_string(data, len(data), C.byref(sc)))
    return ok, scene_objects(sc)


def parse_scene_file_large(path: str) -> np.ndarray:
    L = load_library()
    ptr, n = C.c_void_p(), C.c_int()
    _check(L.rt_parse_scene_file_large(os.fsencode(path), C.byref(ptr), C.byref(n)))
    try:
        raw = (C.c_uint8 * (68 * n.value)).from_address(ptr.value) if n.value else b""
        return np.frombuffer(bytes(raw), dtype=OBJECT_DTYPE).copy()
    finally:
        L.rt_free_objects(ptr)


def parse_scene_string_large(text) -> np.ndarray:
    L = load_library()
    data = text.encode() if isinstance(text, str) else bytes(text)
    ptr, n = C.c_void_p(), C.c_int()
    _check(L.rt_parse_scene_string_large(data, len(data), C.byref(ptr), C.byref(n)))
    try:
        raw = (C.c_uint8 * (68 * n.value)).from_address(ptr.value) if n.value else b""
        return np.frombuffer(bytes(raw), dtype=OBJECT_DTYPE).copy()
    finally:
        L.rt_free_objects(ptr)


def scene_objects(sc: RtScene) -> np.ndarray:
    n = sc.num_objects
    return np.frombuffer(bytes(sc.objects)[: 68 * n], dtype=OBJECT_DTYPE).copy()


def make_scene(objects) -> RtScene:
    objects = np.ascontiguousarray(objects, dtype=OBJECT_DTYPE)
    if len(objects) > RT_MAX_OBJECTS:
        raise ValueError("RtScene holds at most 1024 objects (scene.h:3); use upload_objects")
    sc = RtScene()
    C.memmove(sc.objects, objects.ctypes.data, objects.nbytes)
    sc.num_objects = len(objects)
    return sc


@dataclass
class Camera:
    """Snapshot of the reference's file-static pose (camera.c:23-35)."""

    pos: tuple = (5.0, 5.0, 5.0)
    front: tuple = (-1.0, -1.0, -1.0)
    up: tuple = (0.0, 1.0, 0.0)
    fov: float = 30.0

    def as_struct(self) -> RtCamera:
        c = RtCamera()
        c.pos = RtVector3(*[float(v) for v in self.pos])
        c.front = RtVector3(*[float(v) for v in self.front])
        c.up = RtVector3(*[float(v) for v in self.up])
        c.fov = float(self.fov)
        return c

    def as_dict(self):
        return dict(pos=tuple(self.pos), front=tuple(self.front), up=tuple(self.up), fov=self.fov)


def _cam_from_struct(c: RtCamera) -> Camera:
    v = lambda p: (p.x, p.y, p.z)
    return Camera(v(c.pos), v(c.front), v(c.up), c.fov)


def camera_reset() -> None:
    load_library().rt_camera_reset()


def move_camera(direction: int, speed: float) -> None:
    load_library().rt_move_camera(direction, speed)


def rotate_camera(mouse_x: float, mouse_y: float) -> None:
    load_library().rt_rotate_camera(mouse_x, mouse_y)


def get_camera_pos():
    p = load_library().rt_get_camera_pos()
    return (p.x, p.y, p.z)


def camera_snapshot() -> Camera:
    return _cam_from_struct(load_library().rt_camera_snapshot())


SKYBOX_LIB_PATH = os.path.join(os.path.dirname(_HERE), "tools", "librt_skybox_stb.so")


def load_skybox_dir(directory: str) -> np.ndarray:
    """The caller's half of load_cubemap() (gpu_and_windowing.c:19-40): the six
    face JPEGs of `directory` (main.c:500-507) decoded by the reference's own
    decoder, stb_image, through tools/librt_skybox_stb.so.  Returns (6, h, w,
    chan) uint8 in CubeFace order.  Raises when the helper library or a face is
    missing -- callers that can live with other texels fall back explicitly."""
    if not os.path.exists(SKYBOX_LIB_PATH):
        raise FileNotFoundError(f"{SKYBOX_LIB_PATH} not built (make; needs the reference's 3p/stb at build time)")
    L = C.CDLL(SKYBOX_LIB_PATH)
    L.rt_skybox_load_dir.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.rt_skybox_free.argtypes = [C.c_void_p]
    L.rt_skybox_free.restype = None
    ptr, w, h, ch = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
    if L.rt_skybox_load_dir(os.fsencode(directory), C.byref(ptr), C.byref(w), C.byref(h), C.byref(ch)) != 0:
        raise FileNotFoundError(f"skybox faces not decodable in {directory}")
    try:
        n = 6 * w.value * h.value * ch.value
        faces = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).copy()
    finally:
        L.rt_skybox_free(ptr)
    return faces.reshape(6, h.value, w.value, ch.value)


def quantize_frame(frame: np.ndarray) -> np.ndarray:
    """screenshot()'s 8-bit rule (main.c:666-670)."""
    frame = np.ascontiguousarray(frame, dtype=np.float32)
    out = np.empty(frame.shape, np.uint8)
    load_library().rt_quantize_frame(frame.ctypes.data, frame.size // 3, out.ctypes.data)
    return out


def save_screenshot(path: str, frame: np.ndarray) -> None:
    """screenshot() of main.c:637-681: quantise, flip, write PNG (or PPM)."""
    frame = np.ascontiguousarray(frame, dtype=np.float32)
    _check(load_library().rt_save_screenshot(os.fsencode(path), frame.ctypes.data, frame.shape[1], frame.shape[0]))


def pixel_key(px: float, py: float, pass_index: int = 0) -> int:
    return load_library().rt_pixel_key(float(px), float(py), pass_index)


# --------------------------------------------------------------- device side


class Renderer:
    """Process-wide handle on the CUDA render path (the library keeps one
    context, like the reference keeps one set of globals)."""

    def __init__(self, num_gpus: int = 1, device: "int | None" = None):
        self.lib = load_library()
        if device is not None:
            _check(self.lib.rt_cuda_init_device(int(device)))
        else:
            _check(self.lib.rt_cuda_init(int(num_gpus)))
        self._keep = []

    def close(self) -> None:
        self.lib.rt_cuda_shutdown()

    @property
    def num_gpus(self) -> int:
        return self.lib.rt_cuda_num_gpus()

    # -- uploads
    def upload_scene(self, objects) -> None:
        objects = np.ascontiguousarray(objects, dtype=OBJECT_DTYPE)
        if len(objects) <= RT_MAX_OBJECTS:
            sc = make_scene(objects)
            _check(self.lib.rt_cuda_upload_scene(C.byref(sc)))
        else:
            _check(self.lib.rt_cuda_upload_objects(objects.ctypes.data, len(objects)))

    def upload_objects(self, objects) -> None:
        objects = np.ascontiguousarray(objects, dtype=OBJECT_DTYPE)
        _check(self.lib.rt_cuda_upload_objects(objects.ctypes.data, len(objects)))

    def update_objects(self, objects) -> None:
        """Same objects, changed in place: refresh device records, refit the LBVH (no rebuild)."""
        objects = np.ascontiguousarray(objects, dtype=OBJECT_DTYPE)
        _check(self.lib.rt_cuda_update_objects(objects.ctypes.data, len(objects)))

    def upload_skybox(self, faces: np.ndarray) -> None:
        """faces: (6, h, w, chan>=3) uint8 in CubeFace order (front, back, left,
        right, top, bottom), rows top first -- what stb_image returns."""
        faces = np.ascontiguousarray(faces, dtype=np.uint8)
        assert faces.ndim == 4 and faces.shape[0] == 6
        cm = RtCubemap()
        for i in range(6):
            cm.data[i] = faces[i].ctypes.data
        cm.h, cm.w, cm.chan = faces.shape[1], faces.shape[2], faces.shape[3]
        _check(self.lib.rt_cuda_upload_skybox(C.byref(cm)))

    # -- rendering
    def _opts(self, **kw) -> RtRenderOpts:
        o = RtRenderOpts()
        self.lib.rt_render_opts_default(C.byref(o))
        rows = kw.pop("rows", None)
        if rows is not None:
            o.row_begin, o.row_end = rows
        for k, v in kw.items():
            if v is None:
                continue
            if not hasattr(o, k):
                raise TypeError(f"unknown render option {k}")
            setattr(o, k, v)
        return o

    def render_frame(self, camera: Camera, w: int, h: int, scale: int = 1, *, out=None, stats: bool = True, **opts):
        """One pass into a host numpy frame (h, w, 3) f32 or (h, w, 4) u8, row 0 =
        bottom row.  Returns (frame, stats dict)."""
        o = self._opts(scale=scale, **opts)
        rows = (o.row_begin, o.row_end) if (o.row_begin or o.row_end) else (0, h)
        nrows = rows[1] - rows[0] if o.band_only_fb else h
        if out is None:
            out = np.zeros((nrows, w, 3), np.float32) if o.fb_format == RT_FB_F32X3 else np.zeros((nrows, w, 4), np.uint8)
        o.fb_memory = RT_MEM_HOST
        st = RtRenderStats()
        cam = camera.as_struct()
        _check(self.lib.render_frame_cuda_ex(C.byref(cam), out.ctypes.data, w, h, C.byref(o), C.byref(st) if stats else None))
        return out, _stats_dict(st)

    def render_frame_simple(self, objects, camera: Camera, w: int, h: int, scale: int = 1):
        """The north_star call: render_frame_cuda(scene, camera, fb, w, h, scale)."""
        sc = make_scene(objects) if objects is not None else None
        out = np.zeros((h, w, 3), np.float32)
        cam = camera.as_struct()
        _check(self.lib.render_frame_cuda(C.byref(sc) if sc is not None else None, C.byref(cam), out.ctypes.data, w, h, scale))
        return out

    def render_into(self, camera: Camera, ptr: int, w: int, h: int, *, stats: bool = False, host: bool = False, **opts):
        """One pass into caller memory at address `ptr` (device pointer unless
        host=True).  With stats=False and a device pointer the call is
        asynchronous on opts['stream'] (or the library stream)."""
        o = self._opts(**opts)
        o.fb_memory = RT_MEM_HOST if host else RT_MEM_DEVICE
        st = RtRenderStats()
        cam = camera.as_struct()
        _check(self.lib.render_frame_cuda_ex(C.byref(cam), C.c_void_p(ptr), w, h, C.byref(o), C.byref(st) if stats else None))
        return _stats_dict(st) if stats else None

    def render_sweep(self, camera: Camera, w: int, h: int, init_scale: int, first_pass: int = 0, *, out=None, ptr=None, stats=True, host: bool = False, **opts):
        """The 16 -> 1 progressive sweep (main.c:354, 402-403) into a numpy frame
        (`out`, or a fresh one) or into caller memory at `ptr` (device pointer
        unless host=True)."""
        o = self._opts(**opts)
        if ptr is None:
            if out is None:
                out = np.zeros((h, w, 3), np.float32) if o.fb_format == RT_FB_F32X3 else np.zeros((h, w, 4), np.uint8)
            o.fb_memory = RT_MEM_HOST
            dst = out.ctypes.data
        else:
            o.fb_memory = RT_MEM_HOST if host else RT_MEM_DEVICE
            dst = ptr
        st = RtRenderStats()
        cam = camera.as_struct()
        _check(self.lib.rt_cuda_render_sweep(C.byref(cam), C.c_void_p(dst), w, h, init_scale, first_pass, C.byref(o), C.byref(st) if stats else None))
        return out, _stats_dict(st)

    # -- cross-process frame on GPU 0 (fused P2P composite)
    def shared_frame_create(self, nbytes: int):
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _check(self.lib.rt_cuda_shared_frame_create(nbytes, C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def shared_frame_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        _check(self.lib.rt_cuda_shared_frame_open(buf, C.byref(ptr)))
        return ptr.value

    def shared_frame_close(self, ptr: int, owner: bool) -> None:
        _check(self.lib.rt_cuda_shared_frame_close(C.c_void_p(ptr), 1 if owner else 0))

    def shared_frame_wait(self, ptr: int, num_ranks: int, seq: int, stream=None) -> None:
        _check(self.lib.rt_cuda_shared_frame_wait(C.c_void_p(ptr), int(num_ranks), int(seq), C.c_void_p(stream) if stream else None))

    def shared_frame_release(self, ptr: int, seq: int, stream=None) -> None:
        _check(self.lib.rt_cuda_shared_frame_release(C.c_void_p(ptr), int(seq), C.c_void_p(stream) if stream else None))

    def shared_frame_error(self, ptr: int) -> int:
        e = C.c_uint32()
        _check(self.lib.rt_cuda_shared_frame_error(C.c_void_p(ptr), C.byref(e)))
        return int(e.value)

    def copy_to_host(self, host_ptr: int, dev_ptr: int, nbytes: int, stream=None) -> None:
        _check(self.lib.rt_cuda_copy_to_host(C.c_void_p(host_ptr), C.c_void_p(dev_ptr), nbytes, C.c_void_p(stream) if stream else None))

    def copy_async(self, dst_ptr: int, src_ptr: int, nbytes: int, stream=None) -> None:
        _check(self.lib.rt_cuda_copy_async(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), nbytes, C.c_void_p(stream) if stream else None))

    # -- the reference's frame loop (main.c:324-482)
    def set_progressive(self, init_scale: int, num_columns: int = 1) -> None:
        _check(self.lib.rt_cuda_set_progressive(init_scale, num_columns))

    def invalidate_accumulation(self) -> None:
        _check(self.lib.rt_cuda_invalidate_accumulation())

    def accum_generation(self) -> int:
        return self.lib.rt_cuda_accum_generation()

    def next_pass_index(self) -> int:
        return int(self.lib.rt_cuda_next_pass_index())

    def update_frame(self, camera: Camera, w: int, h: int, budget_ms: float = 0.0, *, out=None, **opts):
        o = self._opts(**opts)
        if out is None:
            out = np.zeros((h, w, 3), np.float32) if o.fb_format == RT_FB_F32X3 else np.zeros((h, w, 4), np.uint8)
        o.fb_memory = RT_MEM_HOST
        st = RtRenderStats()
        cam = camera.as_struct()
        _check(self.lib.rt_cuda_update_frame(C.byref(cam), out.ctypes.data, w, h, budget_ms, C.byref(o), C.byref(st)))
        return out, _stats_dict(st)

    def accum_reset(self) -> None:
        _check(self.lib.rt_cuda_accum_reset())

    def accum_count(self) -> float:
        return self.lib.rt_cuda_accum_count()

    def synchronize(self) -> None:
        _check(self.lib.rt_cuda_synchronize())

    # -- unit probes
    def debug_trace(self, rays, variant=RT_VARIANT_EXACT, traversal=RT_TRAVERSAL_AUTO):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        out = np.zeros((len(rays), 7), np.float32)
        obj = np.zeros(len(rays), np.int32)
        _check(self.lib.rt_cuda_debug_trace(rays.ctypes.data, len(rays), out.ctypes.data, obj.ctypes.data, variant, traversal))
        return out, obj

    def debug_sample_cubemap(self, dirs):
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(dirs), 3), np.float32)
        _check(self.lib.rt_cuda_debug_sample_cubemap(dirs.ctypes.data, len(dirs), out.ctypes.data))
        return out

    def debug_camera_rays(self, camera: Camera, pxpy, aspect: float):
        pxpy = np.ascontiguousarray(pxpy, dtype=np.float32).reshape(-1, 2)
        out = np.zeros((len(pxpy), 6), np.float32)
        cam = camera.as_struct()
        _check(self.lib.rt_cuda_debug_camera_rays(C.byref(cam), pxpy.ctypes.data, len(pxpy), aspect, out.ctypes.data))
        return out

    def debug_rng(self, state: int, n: int):
        u = np.zeros(n, np.uint64)
        f = np.zeros(n, np.float32)
        _check(self.lib.rt_cuda_debug_rng(state, n, u.ctypes.data, f.ctypes.data))
        return u, f

    def div_check(self, seed: int, blocks: int, per_thread: int, lo_b=-40, hi_b=40, lo_a=-60, hi_a=60) -> int:
        bad = C.c_uint64()
        _check(self.lib.rt_cuda_debug_div_check(seed, blocks, per_thread, lo_b, hi_b, lo_a, hi_a, C.byref(bad)))
        return bad.value

    def set_sweep_threshold(self, tau2: float) -> None:
        _check(self.lib.rt_cuda_debug_set_sweep_threshold(tau2))

    # -- CUDA-OpenGL presenter (needs a current GL context)
    def gl_register_buffer(self, gl_buffer: int, nbytes: int) -> None:
        _check(self.lib.rt_cuda_gl_register_buffer(gl_buffer, nbytes))

    def gl_update_frame(self, camera: Camera, w: int, h: int, budget_ms: float = 0.0, **opts):
        o = self._opts(**opts)
        st = RtRenderStats()
        cam = camera.as_struct()
        _check(self.lib.rt_cuda_gl_update_frame(C.byref(cam), w, h, budget_ms, C.byref(o), C.byref(st)))
        return _stats_dict(st)

    def gl_render_frame(self, camera: Camera, w: int, h: int, **opts):
        o = self._opts(**opts)
        st = RtRenderStats()
        cam = camera.as_struct()
        _check(self.lib.rt_cuda_gl_render_frame(C.byref(cam), w, h, C.byref(o), C.byref(st)))
        return _stats_dict(st)

    def gl_unregister_buffer(self) -> None:
        _check(self.lib.rt_cuda_gl_unregister_buffer())

    def set_tile_schedule(self, on) -> None:
        """False / 0: tiles in image order; True / 1 (default): longest tiles first from the costs recorded at
        the pass's own scale; 2: finer passes are also seeded by coarser ones."""
        _check(self.lib.rt_cuda_debug_set_tile_schedule(2 if (on == 2 and on is not True) else (1 if on else 0)))

    def set_queued_dense(self, on: bool) -> None:
        _check(self.lib.rt_cuda_debug_set_queued_dense(1 if on else 0))

    def set_sync_bands(self, bands: int) -> None:
        _check(self.lib.rt_cuda_debug_set_sync_bands(int(bands)))

    def set_concurrent_sweep(self, on: bool) -> None:
        _check(self.lib.rt_cuda_debug_set_concurrent_sweep(1 if on else 0))

    def set_light_anyhit(self, on: bool) -> None:
        _check(self.lib.rt_cuda_debug_set_light_anyhit(1 if on else 0))

    def set_bvh_builder(self, builder: int) -> None:
        """RT_BVH_BUILDER_SAH (host, default) or RT_BVH_BUILDER_LBVH (device) for the next scene upload."""
        _check(self.lib.rt_cuda_set_bvh_builder(int(builder)))

    def debug_tile_order(self, cost: np.ndarray, shift: int, tiles_x: int, tiles_y: int) -> np.ndarray:
        cost = np.ascontiguousarray(cost, dtype=np.uint32)
        out = np.zeros(tiles_x * tiles_y, np.uint32)
        _check(self.lib.rt_cuda_debug_tile_order(cost.ctypes.data, cost.shape[1], cost.shape[0], shift, tiles_x, tiles_y, out.ctypes.data))
        return out

    def ray_counter(self) -> int:
        """Rays traced on GPU 0 since the last call that returned statistics (waits for the device)."""
        n = C.c_uint64()
        _check(self.lib.rt_cuda_ray_counter(C.byref(n)))
        return int(n.value)

    def walk_counts(self):
        """(internal LBVH nodes visited, primitives tested) of the last call with stats; zeros unless built with -DRT_COUNT_WALK."""
        a, b = C.c_uint64(), C.c_uint64()
        _check(self.lib.rt_cuda_debug_walk_counts(C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def fp32_peak_tflops(self, fma: bool = True) -> float:
        out = C.c_float()
        _check(self.lib.rt_cuda_debug_fp32_peak(1 if fma else 0, C.byref(out)))
        return out.value

    def debug_random_directions(self, state: int, n: int):
        out = np.zeros((n, 3), np.float32)
        _check(self.lib.rt_cuda_debug_random_directions(state, n, out.ctypes.data))
        return out


def _stats_dict(st: RtRenderStats) -> dict:
    return dict(
        rays=int(st.rays),
        pixels=int(st.pixels),
        render_ms=float(st.render_ms),
        composite_ms=float(st.composite_ms),
        copy_ms=float(st.copy_ms),
        kernel_launches=int(st.kernel_launches),
    )
