"""the CUDA-OpenGL presenter (include/rt_cuda.h: rt_cuda_gl_*),
which replaces the per-frame host upload of move_frame_to_the_gpu()
(gpu_and_windowing.c:371-376), against a REAL GL context: a headless EGL
device-platform context (tools/egl_probe.py), a pixel-unpack buffer registered
with CUDA, one frame rendered straight into it and read back with
glGetBufferSubData, compared with the oracle bit for bit.

The B200 pool's image has no graphics stack at all (profiles/r02_egl_probe.json:
libEGL.so.1 is not installed), so this test skips there with the probe's reason;
it is the test to run on a box that has one."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

GL_PIXEL_UNPACK_BUFFER = 0x88EC
GL_STREAM_DRAW = 0x88E0


def test_render_into_a_registered_gl_buffer(renderer, port, small_sky, builtin_objects):
    import egl_probe
    from ray_tracing_b200.host import Camera

    report = {}
    try:
        egl, dpy, ctx, surf, proc = egl_probe.egl_context(report)
    except RuntimeError as e:
        pytest.skip(f"no headless GL context on this box: {e}")
    gen = proc("glGenBuffers", None, C.c_int, C.POINTER(C.c_uint))
    bind = proc("glBindBuffer", None, C.c_uint, C.c_uint)
    data = proc("glBufferData", None, C.c_uint, C.c_ssize_t, C.c_void_p, C.c_uint)
    getsub = proc("glGetBufferSubData", None, C.c_uint, C.c_ssize_t, C.c_ssize_t, C.c_void_p)
    finish = proc("glFinish", None)
    W, H = 320, 180
    pbo = C.c_uint()
    gen(1, C.byref(pbo))
    bind(GL_PIXEL_UNPACK_BUFFER, pbo.value)
    data(GL_PIXEL_UNPACK_BUFFER, W * H * 12, None, GL_STREAM_DRAW)
    bind(GL_PIXEL_UNPACK_BUFFER, 0)
    renderer.upload_skybox(small_sky)
    renderer.upload_scene(builtin_objects[0])
    renderer.gl_register_buffer(pbo.value, W * H * 12)
    try:
        st = renderer.gl_render_frame(Camera(), W, H, scale=1)
        got = np.zeros((H, W, 3), np.float32)
        bind(GL_PIXEL_UNPACK_BUFFER, pbo.value)
        finish()
        getsub(GL_PIXEL_UNPACK_BUFFER, 0, got.nbytes, got.ctypes.data)
        bind(GL_PIXEL_UNPACK_BUFFER, 0)
    finally:
        renderer.gl_unregister_buffer()
    want, rays = port.render(port.world(builtin_objects[0], small_sky), W, H, 1, 1, 0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert st["rays"] == rays
