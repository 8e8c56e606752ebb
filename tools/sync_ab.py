#!/usr/bin/env python
"""A/B of the banded host read-back (development tool, GPU box only): wall time of synchronous
render_frame_cuda_ex calls into a pinned host frame, scene_0 at 4K and 1080p, for several band
counts (rt_cuda_debug_set_sync_bands); pageable frame too."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    from ray_tracing_b200 import host, scenes
    r = host.Renderer(num_gpus=1)
    r.upload_skybox(scenes.procedural_skybox(256, seed=11))
    r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(0)))
    cam = host.Camera()
    for W, H in ((3840, 2160), (1920, 1080)):
        pinned = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        pageable = np.zeros((H, W, 3), np.float32)
        out = {"size": f"{W}x{H}"}
        for bands in (1, 2, 3, 4, 6, 8):
            r.set_sync_bands(-bands if bands > 1 else 1)
            for name, ptr in (("pinned", pinned.data_ptr()), ("pageable", pageable.ctypes.data)):
                for k in range(3):
                    r.render_into(cam, ptr, W, H, host=True, pass_index=k)
                t0 = time.perf_counter()
                n = 20
                for k in range(n):
                    r.render_into(cam, ptr, W, H, host=True, pass_index=k)
                out[f"{name}_b{bands}_ms"] = round((time.perf_counter() - t0) / n * 1e3, 3)
        print(json.dumps(out), flush=True)
    r.set_sync_bands(4)
    r.close()


if __name__ == "__main__":
    main()
