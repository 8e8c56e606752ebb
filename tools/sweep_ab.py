#!/usr/bin/env python
"""A/B of the progressive sweep (development tool, GPU box only): BASELINE config 4
(scene_0 1920x1080, 16 -> 1) with the passes side by side (rt_api.cu: sweep_concurrent) and one
after the other; device time per sweep over back-to-back stream-ordered sweeps, sha of the frame.
RT_SWEEP_SHARES="s2,s4,coarser" overrides the CTA-slot shares of the coarse passes."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from ray_tracing_b200 import host, scenes
    W, H = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
    r = host.Renderer(num_gpus=1)
    r.upload_skybox(scenes.procedural_skybox(256, seed=11))
    r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(0)))
    cam = host.Camera()
    frame = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    out = {"size": f"{W}x{H}", "shares": os.environ.get("RT_SWEEP_SHARES", "default")}
    for name, on in (("concurrent", True), ("sequential", False), ("concurrent_again", True)):
        r.set_concurrent_sweep(on)
        for k in range(5):
            r.render_sweep(cam, W, H, 16, 5 * k, ptr=frame.data_ptr(), stats=False, stream=stream.cuda_stream)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        e0.record(stream)
        for k in range(n):
            r.render_sweep(cam, W, H, 16, 5 * k, ptr=frame.data_ptr(), stats=False, stream=stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        r.render_sweep(cam, W, H, 16, 0, ptr=frame.data_ptr(), stats=False, stream=stream.cuda_stream)
        stream.synchronize()
        out[name] = {"ms": round(e0.elapsed_time(e1) / n, 4), "sha": hashlib.sha256(frame.cpu().numpy().tobytes()).hexdigest()[:16]}
    r.set_concurrent_sweep(True)
    r.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
