#!/usr/bin/env python
"""Is there a headless OpenGL context to be had on this box?  (the
CUDA-OpenGL presenter replaces gpu_and_windowing.c:371-376.)  Tries EGL with the
device platform (EGL_EXT_platform_device: no X server, no display needed),
creates a pbuffer surface and a desktop-GL context, and reports every step as
one JSON line.  tests/test_gl_interop.py uses egl_context() from here."""
import ctypes as C
import ctypes.util
import json
import sys

EGL_PLATFORM_DEVICE_EXT = 0x313F
EGL_NONE = 0x3038
EGL_SURFACE_TYPE, EGL_PBUFFER_BIT = 0x3033, 0x0001
EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT = 0x3040, 0x0008
EGL_RED_SIZE, EGL_GREEN_SIZE, EGL_BLUE_SIZE = 0x3024, 0x3023, 0x3022
EGL_WIDTH, EGL_HEIGHT = 0x3057, 0x3056
EGL_OPENGL_API = 0x30A2


def egl_context(report=None):
    """Returns (egl library, display, context, surface, getproc) with the context
    current on this thread, or raises RuntimeError saying which step failed."""
    report = report if report is not None else {}
    name = ctypes.util.find_library("EGL") or "libEGL.so.1"
    try:
        egl = C.CDLL(name)
    except OSError as e:
        report["libEGL"] = f"not loadable: {e}"
        raise RuntimeError(report["libEGL"])
    report["libEGL"] = name
    egl.eglGetProcAddress.restype = C.c_void_p
    egl.eglGetProcAddress.argtypes = [C.c_char_p]

    def proc(fn, restype, *argtypes):
        p = egl.eglGetProcAddress(fn.encode())
        if not p:
            raise RuntimeError(f"{fn} not available")
        return C.CFUNCTYPE(restype, *argtypes)(p)

    query = proc("eglQueryDevicesEXT", C.c_uint, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int))
    getdisp = proc("eglGetPlatformDisplayEXT", C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_int))
    devs = (C.c_void_p * 16)()
    n = C.c_int(0)
    if not query(16, devs, C.byref(n)) or n.value == 0:
        report["devices"] = 0
        raise RuntimeError("eglQueryDevicesEXT found no device")
    report["devices"] = n.value
    last = "no device initialised"
    for i in range(n.value):
        dpy = getdisp(EGL_PLATFORM_DEVICE_EXT, devs[i], None)
        major, minor = C.c_int(), C.c_int()
        egl.eglInitialize.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        if not dpy or not egl.eglInitialize(dpy, C.byref(major), C.byref(minor)):
            last = f"eglInitialize failed on device {i}"
            continue
        report["egl_version"] = f"{major.value}.{minor.value}"
        egl.eglBindAPI.argtypes = [C.c_uint]
        if not egl.eglBindAPI(EGL_OPENGL_API):
            last = "eglBindAPI(EGL_OPENGL_API) failed"
            continue
        attrs = (C.c_int * 11)(EGL_SURFACE_TYPE, EGL_PBUFFER_BIT, EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT, EGL_RED_SIZE, 8, EGL_GREEN_SIZE, 8, EGL_BLUE_SIZE, 8, EGL_NONE)
        cfg, ncfg = C.c_void_p(), C.c_int()
        egl.eglChooseConfig.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]
        if not egl.eglChooseConfig(dpy, attrs, C.byref(cfg), 1, C.byref(ncfg)) or ncfg.value == 0:
            last = "eglChooseConfig found no pbuffer + OpenGL config"
            continue
        pattrs = (C.c_int * 5)(EGL_WIDTH, 16, EGL_HEIGHT, 16, EGL_NONE)
        egl.eglCreatePbufferSurface.restype = C.c_void_p
        egl.eglCreatePbufferSurface.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        surf = egl.eglCreatePbufferSurface(dpy, cfg, pattrs)
        egl.eglCreateContext.restype = C.c_void_p
        egl.eglCreateContext.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        ctx = egl.eglCreateContext(dpy, cfg, None, None)
        egl.eglMakeCurrent.argtypes = [C.c_void_p] * 4
        if not surf or not ctx or not egl.eglMakeCurrent(dpy, surf, surf, ctx):
            last = "pbuffer surface / context / eglMakeCurrent failed"
            continue
        report["device_index"] = i
        return egl, dpy, ctx, surf, proc
    raise RuntimeError(last)


def main():
    report = {}
    try:
        egl, dpy, ctx, surf, proc = egl_context(report)
        gl_get_string = proc("glGetString", C.c_char_p, C.c_uint)
        report["gl_vendor"] = (gl_get_string(0x1F00) or b"").decode()
        report["gl_renderer"] = (gl_get_string(0x1F01) or b"").decode()
        report["gl_version"] = (gl_get_string(0x1F02) or b"").decode()
        report["ok"] = True
    except Exception as e:
        report["ok"] = False
        report["error"] = f"{type(e).__name__}: {e}"
    print(json.dumps(report))
    return 0


if __name__ == "__main__":
    sys.exit(main())
