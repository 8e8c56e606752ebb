"""ctypes bindings for the oracle -- TEST INFRASTRUCTURE.

Two checkers live here:

* ``Port``  -- oracle/librt_oracle.so, the CPU restatement (oracle/rt_oracle.c).
* ``Ref``   -- oracle/_ref/libref_*.so, the UNMODIFIED reference compiled by
  oracle/Makefile from /root/reference (prebuilt files travel to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(ray_tracing_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ASSETS = os.path.join(REF_DIR, "assets")

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

DEFAULT_CAMERA = dict(pos=(5, 5, 5), front=(-1, -1, -1), up=(0, 1, 0), fov=30.0)


def build(port: bool = True, ref: bool = True) -> None:
    targets = []
    if port:
        targets.append("port")
    if ref and os.path.isdir("/root/reference/src"):
        targets.append("ref")
    if targets:
        subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


class _Cam(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("front", C.c_float * 3), ("up", C.c_float * 3), ("fov", C.c_float)]


class _Sky(C.Structure):
    _fields_ = [("face", C.c_void_p * 6), ("w", C.c_int), ("h", C.c_int), ("chan", C.c_int)]


class _World(C.Structure):
    _fields_ = [("objects", C.c_void_p), ("num_objects", C.c_int), ("camera", _Cam), ("sky", _Sky)]


def _cam_struct(cam) -> _Cam:
    c = _Cam()
    c.pos[:] = [float(x) for x in cam["pos"]]
    c.front[:] = [float(x) for x in cam["front"]]
    c.up[:] = [float(x) for x in cam["up"]]
    c.fov = float(cam["fov"])
    return c


class Port:
    """The CPU restatement (oracle/rt_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "librt_oracle.so")
        if not os.path.exists(path):
            build(port=True, ref=False)
        L = self.lib = C.CDLL(path)
        L.rto_wyhash64.restype = C.c_uint64
        L.rto_wyhash64.argtypes = [C.POINTER(C.c_uint64)]
        L.rto_random_float.restype = C.c_float
        L.rto_random_float.argtypes = [C.POINTER(C.c_uint64)]
        L.rto_random_direction.argtypes = [C.POINTER(C.c_uint64), C.c_void_p]
        L.rto_pixel_key.restype = C.c_uint64
        L.rto_pixel_key.argtypes = [C.c_float, C.c_float, C.c_uint64]
        L.rto_camera_ray.argtypes = [C.POINTER(_Cam), C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.rto_trace_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.rto_sample_cubemap_many.argtypes = [C.POINTER(_Sky), C.c_void_p, C.c_int, C.c_void_p]
        L.rto_pixel.argtypes = [C.POINTER(_World), C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.rto_render.restype = C.c_uint64
        L.rto_render.argtypes = [C.POINTER(_World), C.c_void_p] + [C.c_int] * 4 + [C.c_uint64] + [C.c_int] * 3
        L.rto_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.rto_resolve.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float]
        self._keep = []

    # -- helpers
    def _sky(self, faces) -> _Sky:
        s = _Sky()
        faces = np.ascontiguousarray(faces, dtype=np.uint8)
        assert faces.ndim == 4 and faces.shape[0] == 6
        self._keep.append(faces)
        for i in range(6):
            s.face[i] = faces[i].ctypes.data
        s.h, s.w, s.chan = faces.shape[1], faces.shape[2], faces.shape[3]
        return s

    def world(self, objects, faces, camera=None) -> _World:
        objects = np.ascontiguousarray(objects, dtype=OBJECT_DTYPE)
        self._keep.append(objects)
        w = _World()
        w.objects = objects.ctypes.data
        w.num_objects = len(objects)
        w.camera = _cam_struct(camera or DEFAULT_CAMERA)
        w.sky = self._sky(faces)
        return w

    # -- unit probes
    def rng_u64(self, state: int, n: int):
        st = C.c_uint64(state)
        return [self.lib.rto_wyhash64(C.byref(st)) for _ in range(n)]

    def random_floats(self, state: int, n: int):
        st = C.c_uint64(state)
        return np.array([self.lib.rto_random_float(C.byref(st)) for _ in range(n)], np.float32)

    def random_direction(self, state: int):
        st = C.c_uint64(state)
        out = np.zeros(3, np.float32)
        self.lib.rto_random_direction(C.byref(st), out.ctypes.data)
        return out, st.value

    def pixel_key(self, px, py, pass_index=0) -> int:
        return self.lib.rto_pixel_key(float(px), float(py), pass_index)

    def camera_ray(self, px, py, aspect, camera=None):
        out = np.zeros(6, np.float32)
        cam = _cam_struct(camera or DEFAULT_CAMERA)
        self.lib.rto_camera_ray(C.byref(cam), float(px), float(py), float(aspect), out.ctypes.data)
        return out

    def trace_many(self, objects, rays):
        objects = np.ascontiguousarray(objects, dtype=OBJECT_DTYPE)
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        out = np.zeros((len(rays), 7), np.float32)
        obj = np.zeros(len(rays), np.int32)
        self.lib.rto_trace_many(objects.ctypes.data, len(objects), rays.ctypes.data, len(rays), out.ctypes.data, obj.ctypes.data)
        return out, obj

    def sample_cubemap_many(self, faces, dirs):
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        sky = self._sky(faces)
        out = np.zeros((len(dirs), 3), np.float32)
        self.lib.rto_sample_cubemap_many(C.byref(sky), dirs.ctypes.data, len(dirs), out.ctypes.data)
        return out

    def pixel(self, world, px, py, aspect, rng_state):
        out = np.zeros(3, np.float32)
        rays = C.c_uint64()
        self.lib.rto_pixel(C.byref(world), float(px), float(py), float(aspect), rng_state, out.ctypes.data, C.byref(rays))
        return out, rays.value

    def render(self, world, W, H, scale=1, num_columns=1, pass_index=0, rows=None, nthreads=None, out=None):
        if out is None:
            out = np.zeros((H, W, 3), np.float32)
        r0, r1 = rows if rows is not None else (0, H)
        nthreads = nthreads or (os.cpu_count() or 1)
        rays = self.lib.rto_render(C.byref(world), out.ctypes.data, W, H, scale, num_columns, pass_index, r0, r1, nthreads)
        return out, rays

    def accumulate(self, accum, data, scale):
        assert accum.dtype == np.float32 and data.dtype == np.float32
        self.lib.rto_accumulate(accum.ctypes.data, data.ctypes.data, accum.size, scale)

    def resolve(self, accum, count):
        frame = np.empty_like(accum)
        self.lib.rto_resolve(frame.ctypes.data, accum.ctypes.data, accum.size, float(count))
        return frame


def ref_available(variant: str = "pixel") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libref_{variant}.so"))


class Ref:
    """The unmodified reference behind oracle/ref_driver.c.

    variant: 'stream' (as shipped), 'count' (+ray counter), 'pixel' (per-pixel
    RNG key + ray counter), 'pixel_big' (MAX_OBJECTS raised to 131072).
    The reference keeps scene/skybox/camera in process globals, so one instance
    per variant per process.
    """

    def __init__(self, variant: str = "pixel"):
        path = os.path.join(REF_DIR, f"libref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        L = self.lib = C.CDLL(path)
        self.variant = variant
        L.refdrv_sizeof_scene.restype = C.c_size_t
        L.refdrv_sizeof_object.restype = C.c_size_t
        L.refdrv_scene_ptr.restype = C.c_void_p
        L.refdrv_parse_scene_file.argtypes = [C.c_char_p]
        L.refdrv_parse_scene_file_into.argtypes = [C.c_char_p, C.c_void_p]
        L.refdrv_set_scene.argtypes = [C.c_void_p, C.c_int]
        L.refdrv_load_skybox.argtypes = [C.c_char_p]
        L.refdrv_set_skybox.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.refdrv_get_skybox.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.refdrv_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.refdrv_get_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.refdrv_move_camera.argtypes = [C.c_int, C.c_float]
        L.refdrv_rotate_camera.argtypes = [C.c_double, C.c_double]
        L.refdrv_rng_seed.argtypes = [C.c_uint64]
        L.refdrv_rng_state.restype = C.c_uint64
        L.refdrv_rng_u64.restype = C.c_uint64
        L.refdrv_random_float.restype = C.c_float
        L.refdrv_random_direction.argtypes = [C.c_void_p]
        L.refdrv_pixel_key.restype = C.c_uint64
        L.refdrv_pixel_key.argtypes = [C.c_float, C.c_float, C.c_uint64]
        L.refdrv_camera_ray.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.refdrv_trace_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.refdrv_sample_cubemap_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.refdrv_pixel.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_void_p]
        L.refdrv_render.restype = C.c_double
        L.refdrv_render.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
        self._keep = []
        self.max_objects = L.refdrv_max_objects()

    # -- scene / skybox / camera
    def parse_scene_file(self, path: str) -> bool:
        return self.lib.refdrv_parse_scene_file(os.fsencode(path)) == 0

    def parse_scene_file_objects(self, path: str, fill: int = 0):
        """Parse into a fresh buffer pre-filled with `fill`; returns the object
        records (or None on parse failure)."""
        size = self.lib.refdrv_sizeof_scene()
        buf = np.full(size, fill, np.uint8)
        ok = self.lib.refdrv_parse_scene_file_into(os.fsencode(path), buf.ctypes.data) == 0
        if not ok:
            return None
        n = int(buf[self.max_objects * 68 : self.max_objects * 68 + 4].view("<i4")[0])
        return buf[: n * 68].view(OBJECT_DTYPE).copy()

    def scene_objects(self):
        size = self.lib.refdrv_sizeof_scene()
        raw = (C.c_uint8 * size).from_address(self.lib.refdrv_scene_ptr())
        buf = np.frombuffer(raw, np.uint8)
        n = int(buf[self.max_objects * 68 : self.max_objects * 68 + 4].view("<i4")[0])
        return buf[: n * 68].view(OBJECT_DTYPE).copy()

    def set_scene(self, objects) -> None:
        objects = np.ascontiguousarray(objects, dtype=OBJECT_DTYPE)
        if self.lib.refdrv_set_scene(objects.ctypes.data, len(objects)) != 0:
            raise ValueError("too many objects for this reference build")

    def load_skybox(self, directory: str = None):
        directory = directory or os.path.join(ASSETS, "skybox")
        if self.lib.refdrv_load_skybox(os.fsencode(directory)) != 0:
            raise FileNotFoundError(directory)
        return self.get_skybox()

    def get_skybox(self):
        ptrs = (C.c_void_p * 6)()
        w, h, ch = C.c_int(), C.c_int(), C.c_int()
        self.lib.refdrv_get_skybox(ptrs, C.byref(w), C.byref(h), C.byref(ch))
        n = w.value * h.value * ch.value
        faces = np.stack([np.frombuffer((C.c_uint8 * n).from_address(ptrs[i]), np.uint8) for i in range(6)])
        return faces.reshape(6, h.value, w.value, ch.value).copy()

    def set_skybox(self, faces) -> None:
        faces = np.ascontiguousarray(faces, dtype=np.uint8)
        self._keep = [faces]
        ptrs = (C.c_void_p * 6)(*[faces[i].ctypes.data for i in range(6)])
        self.lib.refdrv_set_skybox(ptrs, faces.shape[2], faces.shape[1], faces.shape[3])

    def set_camera(self, camera) -> None:
        a = [np.asarray(camera[k], np.float32) for k in ("pos", "front", "up")]
        self.lib.refdrv_set_camera(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, float(camera["fov"]))

    def get_camera(self):
        a = [np.zeros(3, np.float32) for _ in range(3)]
        fov = C.c_float()
        self.lib.refdrv_get_camera(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, C.byref(fov))
        return dict(pos=a[0], front=a[1], up=a[2], fov=fov.value)

    def reset_camera(self) -> None:
        self.lib.refdrv_reset_camera()

    def move_camera(self, direction: int, speed: float) -> None:
        self.lib.refdrv_move_camera(direction, speed)

    def rotate_camera(self, mx: float, my: float) -> None:
        self.lib.refdrv_rotate_camera(mx, my)

    # -- unit probes
    def rng_u64(self, state: int, n: int):
        self.lib.refdrv_rng_seed(state)
        return [self.lib.refdrv_rng_u64() for _ in range(n)]

    def random_floats(self, state: int, n: int):
        self.lib.refdrv_rng_seed(state)
        return np.array([self.lib.refdrv_random_float() for _ in range(n)], np.float32)

    def random_direction(self, state: int):
        self.lib.refdrv_rng_seed(state)
        out = np.zeros(3, np.float32)
        self.lib.refdrv_random_direction(out.ctypes.data)
        return out, self.lib.refdrv_rng_state()

    def pixel_key(self, px, py, pass_index=0) -> int:
        return self.lib.refdrv_pixel_key(float(px), float(py), pass_index)

    def camera_ray(self, px, py, aspect):
        out = np.zeros(6, np.float32)
        self.lib.refdrv_camera_ray(float(px), float(py), float(aspect), out.ctypes.data)
        return out

    def trace_many(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        out = np.zeros((len(rays), 7), np.float32)
        obj = np.zeros(len(rays), np.int32)
        self.lib.refdrv_trace_many(rays.ctypes.data, len(rays), out.ctypes.data, obj.ctypes.data)
        return out, obj

    def sample_cubemap_many(self, dirs):
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(dirs), 3), np.float32)
        self.lib.refdrv_sample_cubemap_many(dirs.ctypes.data, len(dirs), out.ctypes.data)
        return out

    def pixel(self, px, py, aspect, rng_state):
        out = np.zeros(3, np.float32)
        self.lib.refdrv_pixel(float(px), float(py), float(aspect), rng_state, out.ctypes.data)
        return out

    def render(self, W, H, scale=1, threads=1, pass_index=0, keyed=True):
        """One render_column() pass per column on `threads` fresh pthreads.
        Returns (frame[H,W,3] f32 bottom row first, wall seconds, trace_ray calls)."""
        out = np.zeros((H, W, 3), np.float32)
        rays = C.c_uint64()
        secs = self.lib.refdrv_render(out.ctypes.data, W, H, scale, threads, pass_index, 1 if keyed else 0, C.byref(rays))
        return out, secs, rays.value


def procedural_skybox(size: int = 64, seed: int = 7) -> np.ndarray:
    """Small deterministic cubemap (6, size, size, 3) u8 for fixtures: smooth
    gradients plus hash noise so neighbouring texels differ."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / max(size - 1, 1)
    faces = np.zeros((6, size, size, 3), np.uint8)
    for f in range(6):
        base = np.stack([(x * (f + 1) / 6.0), (y * (6 - f) / 6.0), ((x + y) * 0.5)], axis=-1)
        noise = rng.integers(0, 64, size=(size, size, 3))
        faces[f] = np.clip(base * 191 + noise, 0, 255).astype(np.uint8)
    return faces
