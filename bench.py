#!/usr/bin/env python
"""Headline benchmark of the per-pixel render path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

One *step* = one full pass of the hot path over one frame:
    scene_0 (9 objects), 3840x2160, scale 1, default camera pose, reference
    skybox (6 x 2048^2) -- BASELINE.json configs[2], the configuration its
    metric names for 1/2/4/8 B200; the same workload at every N so the
    driver's scaling numbers compare like with like (N=1 renders the whole
    frame, N>1 the rows in 16-row blocks dealt round robin and composited to
    rank 0: strong scaling).
Metric: Mrays/s = trace_ray-equivalent invocations (primary + bounce + shadow
rays, counted by the kernel that traced them) per second, whole job.

value  : frame resident on the device (render [+ NCCL gather to rank 0]),
         K steps back to back, CUDA events on the launching stream, max over
         ranks.
e2e    : the same metric through the reference-facing C-ABI call with a HOST
         framebuffer (render_frame_cuda_ex -> Vector3 frame in pinned host
         memory); camera/params go host->device and the frame comes back
         device->host inside the timed region, every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 3840, 2160
SCENE = 0
FLOPS_PER_RAY = {0: 141, 1: 105, 2: 69}     # 15 + 18*spheres + 12*cubes
WORKLOAD = "scene_0.txt 3840x2160 scale 1, default pose, pass 0 (BASELINE.json configs[2]); N>1: the frame's rows dealt to the ranks in 16-row blocks, round robin, composited to rank 0"
STAGED = os.path.join(ROOT, "oracle", "_ref", "assets")


def load_skybox_faces():
    """The reference's six 2048^2 JPEG faces if they were staged next to the
    oracle build (data files only), else a procedural 2048^2 cubemap of the
    same size and format.  Returns (faces u8 [6,h,w,3], description)."""
    from ray_tracing_b200 import scenes

    jpg = [os.path.join(STAGED, "skybox", f) for f in scenes.FACE_FILES]
    if all(os.path.exists(p) for p in jpg):
        try:
            from PIL import Image

            faces = np.stack([np.asarray(Image.open(p).convert("RGB")) for p in jpg])
            return np.ascontiguousarray(faces), "reference skybox JPEGs 6x2048x2048 (decoded with PIL for the bench; parity tests decode with the reference's stb_image)"
        except Exception:
            pass
    return scenes.procedural_skybox(2048, seed=11), "procedural 6x2048x2048 cubemap (reference JPEGs not staged)"


# --------------------------------------------------------------------- clocks


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.path = f"/tmp/rt_bench_clocks_{os.getpid()}.csv"
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for n, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------ reference


def reference_cpu_run(steps: int, warmup: int, budget_s: float = 150.0):
    """The reference's own CPU implementation of the path (oracle/_ref, the
    unmodified TUs compiled from /root/reference; else the oracle port) on all
    host cores, on a bounded sample of the workload."""
    from oracle import bindings
    from ray_tracing_b200 import host, scenes

    cores = os.cpu_count() or 1
    objs = host.parse_scene_string(scenes.builtin_scene_text(SCENE))
    faces, sky_desc = load_skybox_faces()
    use_ref = bindings.ref_available("stream") and bindings.ref_available("count")

    def threads_for(width):
        # render_column needs W % T == 0 to cover the frame (main.c:363)
        return max(t for t in range(1, min(cores, width) + 1) if width % t == 0)

    def run_once(w, h, counted=False):
        if use_ref:
            r = ref_count if counted else ref_stream
            frame, secs, rays = r.render(w, h, 1, threads_for(w), 0, keyed=False)
            return secs, rays
        t0 = time.perf_counter()
        _, rays = port.render(world, w, h, 1, 1, 0, nthreads=cores)
        return time.perf_counter() - t0, rays

    if use_ref:
        ref_stream, ref_count = bindings.Ref("stream"), bindings.Ref("count")
        for r in (ref_stream, ref_count):
            r.set_skybox(faces)
            r.reset_camera()
            r.set_scene(objs)
        kind = "reference"
    else:
        port = bindings.Port()
        world = port.world(objs, faces)
        kind = "port"

    # pick the largest sample of the workload that keeps the whole run bounded
    sizes = [(3840, 2160), (1920, 1080), (1280, 720), (640, 360)]
    probe_s, _ = run_once(640, 360)
    w, h = sizes[-1]
    for cw, ch in sizes:
        est = probe_s * (cw * ch) / (640 * 360)
        if est * (steps + warmup + 1) <= budget_s:
            w, h = cw, ch
            break
    _, rays = run_once(w, h, counted=True)        # ray count of this sample (untimed when kind == reference)
    for _ in range(warmup):
        run_once(w, h)
    times = []
    for _ in range(steps):
        s, r = run_once(w, h)
        times.append(s)
        if not use_ref:
            rays = r
    total = float(sum(times))
    threads = threads_for(w) if use_ref else cores
    return dict(
        mrays=rays * len(times) / total / 1e6, ms_per_step=1e3 * total / len(times), frames_per_s=len(times) / total,
        kind=kind, cores=threads, host_cores=cores, rays_per_step=int(rays),
        sample=f"scene_0 {w}x{h} scale 1 pass 0, full frame, {threads} threads (one render_column per thread), {sky_desc}",
    )


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        # RT_BENCH_REF_BUDGET_S bounds the CPU time of the whole run (tests use a few seconds)
        r = reference_cpu_run(args.steps, args.warmup, budget_s=float(os.environ.get("RT_BENCH_REF_BUDGET_S", "150")))
    except Exception as e:  # the oracle always exists; report rather than crash the driver
        print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}), file=args.out, flush=True)
        return 0
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": r["mrays"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": r["sample"]},
        "frames_per_s": r["frames_per_s"],
        "cpu_baseline": {"value": r["mrays"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["mrays"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


# ------------------------------------------------------------------------ GPU


def main_gpu(args):
    import torch
    import torch.distributed as dist

    from ray_tracing_b200 import host, scenes
    from ray_tracing_b200.distributed import band_rows, gather_bands

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    objs = host.parse_scene_string(scenes.builtin_scene_text(SCENE))
    faces, sky_desc = load_skybox_faces()
    r = host.Renderer(device=local)
    r.upload_skybox(faces)
    r.upload_scene(objs)
    cam = host.Camera()
    variant = host.RT_VARIANT_FAST if args.variant == "fast" else host.RT_VARIANT_EXACT
    kernel = {"auto": host.RT_KERNEL_AUTO, "pixel": host.RT_KERNEL_PIXEL, "persistent": host.RT_KERNEL_PERSISTENT,
              "wavefront": host.RT_KERNEL_WAVEFRONT, "queued": host.RT_KERNEL_QUEUED}[args.kernel]

    r0, r1 = band_rows(H, 1, rank, world)
    band = torch.empty((max(r1 - r0, 1), W, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    common = dict(scale=1, pass_index=0, rows=(r0, r1), band_only_fb=1, variant=variant, kernel=kernel)
    p2p = world > 1 and args.composite == "p2p"
    shared_ptr = None
    if p2p:
        # composite without a data-path collective: every rank renders its row blocks locally and
        # ships them with one strided peer copy into rank 0's frame (cudaIpc mapping)
        box = [None]
        if rank == 0:
            shared_ptr, handle = r.shared_frame_create(W * H * 12)
            box[0] = handle
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            shared_ptr = r.shared_frame_open(box[0])
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        # rows are dealt to ranks in blocks of 16, round robin (sky rows and scene rows cost very different amounts)
        p2p_opts = dict(scale=1, pass_index=0, interleave_count=world, interleave_index=rank, remote_fb=int(rank != 0), variant=variant, kernel=kernel)

    def step_device():
        if p2p:
            r.render_into(cam, shared_ptr, W, H, stream=stream, **p2p_opts)
            dist.all_reduce(flag)      # stream-ordered completion signal: rank 0's frame is whole after it
            return None
        # render this rank's band on torch's stream, then gather on rank 0 (NCCL)
        r.render_into(cam, band.data_ptr(), W, H, stream=stream, **common)
        if world > 1:
            return gather_bands(band, H, W, 1, rank, world, dist, dst=0)
        return band

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # rays of one step (exact count from the kernels; identical every step: pass 0)
    if p2p:
        st = r.render_into(cam, shared_ptr, W, H, stats=True, **p2p_opts)
    else:
        st = r.render_into(cam, band.data_ptr(), W, H, stats=True, **common)
    rays_t = torch.tensor([st["rays"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t)
    rays_per_step = int(rays_t.item())

    # ---- device-resident timing ------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())

    # kernel-only time of the dominant kernel (render), CUDA events on the same stream, this rank
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        if p2p:
            r.render_into(cam, shared_ptr, W, H, stream=stream, **p2p_opts)
        else:
            r.render_into(cam, band.data_ptr(), W, H, stream=stream, **common)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = torch.tensor([k0.elapsed_time(k1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    kern_ms = float(kern_ms.item())

    # ---- the same launches with tiles in image order (what a pose costs the first two
    # times it is rendered, and every time while the camera moves), for the record ----
    unscheduled = None
    if args.kernel in ("auto", "queued"):
        r.set_tile_schedule(False)
        for _ in range(3):
            step_device()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        u0.record()
        for _ in range(10):
            step_device()
        u1.record()
        barrier()
        ums = torch.tensor([u0.elapsed_time(u1) / 10], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ums, op=dist.ReduceOp.MAX)
        unscheduled = {"ms_per_step": float(ums.item()), "value": rays_per_step / (float(ums.item()) * 1e-3) / 1e6, "unit": "Mrays/s",
                       "note": "tiles handed out in image order (no cost-sorted schedule)"}
        r.set_tile_schedule(True)

    # ---- the other build of the same kernels, for the record (N=1 only) ----
    other = None
    if world == 1:
        ov = host.RT_VARIANT_EXACT if variant == host.RT_VARIANT_FAST else host.RT_VARIANT_FAST
        oc = dict(common, variant=ov)
        for _ in range(3):
            r.render_into(cam, band.data_ptr(), W, H, stream=stream, **oc)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        o0.record()
        for _ in range(10):
            r.render_into(cam, band.data_ptr(), W, H, stream=stream, **oc)
        o1.record()
        torch.cuda.synchronize()
        oms = o0.elapsed_time(o1) / 10
        orays = r.render_into(cam, band.data_ptr(), W, H, stats=True, **oc)["rays"]
        other = {"variant": "exact" if ov == host.RT_VARIANT_EXACT else "fast", "ms_per_step": oms, "value": orays / (oms * 1e-3) / 1e6, "unit": "Mrays/s",
                 "note": "fast = FMA contraction + approximate rcp/rsqrt, f32 sphere roots: <= 1 LSB/8-bit channel on >= 99.9 % of pixels; exact = bit-identical to the reference"}

    # ---- end to end through the C ABI with a host framebuffer --------------
    # (rank 0 owns the host frame; with N>1 the bands are gathered to GPU 0 first)
    host_frames = [torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)] if world == 1 else None
    e2e_opts = dict(scale=1, pass_index=0, variant=variant, kernel=kernel, stream=stream)
    shared_host = None
    if world > 1:
        # one host frame shared by all ranks (POSIX shared memory, page-locked in every process):
        # each rank copies the row blocks it rendered over its own PCIe link, no gather through GPU 0
        shm_path = f"/dev/shm/rt_bench_frame_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            np.lib.format.open_memmap(shm_path, mode="w+", dtype=np.float32, shape=(H, W, 3)).flush()
        dist.barrier()
        shared_host = np.load(shm_path, mmap_mode="r+")
        host_addr = shared_host.ctypes.data
        err = torch.cuda.cudart().cudaHostRegister(host_addr, shared_host.nbytes, 0)
        assert int(err) == 0, f"cudaHostRegister failed: {err}"

    def step_e2e(i, pipeline):
        if world == 1:
            # the drop-in call: params go H2D as kernel arguments, the Vector3 frame comes back D2H.
            # pipeline=1: the call returns once the copy is queued; frame i+1 renders while frame i drains
            r.render_into(cam, host_frames[i & 1].data_ptr(), W, H, host=True, pipeline=int(pipeline), **e2e_opts)
        else:
            r.render_into(cam, host_addr, W, H, host=True, pipeline=int(pipeline), interleave_count=world, interleave_index=rank, **e2e_opts)
            if not pipeline:
                dist.barrier()      # the frame is whole once every rank's copy has landed

    def run_e2e(n, pipeline):
        for i in range(3):
            step_e2e(i, pipeline)
        r.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(n):
            step_e2e(i, pipeline)
        r.synchronize()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_steps = max(3, min(args.steps, 20))
    e2e_sync_s = run_e2e(e2e_steps, False)
    e2e_s = run_e2e(e2e_steps, True)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        per_step_ms = total_ms / args.steps
        value = rays_per_step / (per_step_ms * 1e-3) / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # FP32 issue-rate ceiling measured live with register-only chains
        fma_peak = r.fp32_peak_tflops(True)
        muladd_peak = r.fp32_peak_tflops(False)
        peak = muladd_peak if variant == host.RT_VARIANT_EXACT else fma_peak
        kern_rays = st["rays"]                       # this rank's band
        achieved = kern_rays * FLOPS_PER_RAY[SCENE] / (kern_ms * 1e-3) / 1e12
        band_px = (W * H) // world
        algo_bytes = band_px * 12 + band_px * 32     # Vector3 store + one 32 B skybox sector per escaping path
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass

        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "variant": args.variant, "kernel": args.kernel, "skybox": sky_desc,
                "composite": ("none (1 GPU)" if world == 1 else ("P2P: ranks render their 16-row blocks (round robin) locally, then one strided peer copy per rank into rank 0's frame over NVLink (cudaIpc mapping) + 4-byte all_reduce as completion signal" if p2p else "NCCL gather of bands to rank 0")),
                "framebuffer": "Vector3 f32x3 (reference frame format), bottom row first",
                "l2": "no explicit flush: every step reads the 96 MiB RGBA8 skybox at random and writes a 99.5 MB frame (working set 196 MB > 126 MB L2)",
                "rays_per_step": rays_per_step, "pixels_per_step": W * H,
                "schedule": "every step renders the same pose, as the reference's progressive accumulation does (main.c:354-403); from the third launch of a pose the queued kernel hands out its 8x4 tiles longest-first, by the per-tile bounce counts the second launch recorded (warm-up). Scheduling only: frames are bit-identical. `unscheduled` = the same loop with tiles in image order",
            },
            "frames_per_s": 1e3 / per_step_ms,
            "mpix_per_s": W * H / (per_step_ms * 1e-3) / 1e6,
            "e2e": {
                "value": rays_per_step * e2e_steps / e2e_s / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": int(host.load_library().rt_cuda_param_bytes()) * world,   # kernel-argument block (camera frame, views, sizes) per rank; scene and skybox are resident
                "d2h_bytes_per_step": W * H * 12,
                "frames_per_s": e2e_steps / e2e_s, "steps": e2e_steps,
                "mode": "pipelined: call k+1 renders while the copy stream drains frame k into the other pinned host frame; timed until rt_cuda_synchronize()" if world == 1 else "pipelined like N=1; every rank renders its row blocks and copies them over its own PCIe link into one page-locked host frame shared by the ranks (POSIX shm); sync_value = with a barrier after every frame",
                "sync_value": rays_per_step * e2e_steps / e2e_sync_s / 1e6,
                "api": "render_frame_cuda_ex(cam, host Vector3 frame, w, h, opts)",
            },
            "gpu_launches": args.steps * world,        # one render kernel per rank per step (NCCL kernels not counted)
            "clocks": clocks,
            "roofline": {
                "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic, "kernel": {"auto": "render_queued_kernel", "queued": "render_queued_kernel", "persistent": "render_persistent_kernel",
                                                            "wavefront": "render_wavefront_kernel", "pixel": "render_pixel_kernel"}[args.kernel],
                "kernel_ms": kern_ms, "flops_per_ray": FLOPS_PER_RAY[SCENE], "rays_per_launch": kern_rays,
                "peak_source": "measured live: register-only %s chains on this GPU" % ("MUL+ADD (no-FMA exact build)" if variant == host.RT_VARIANT_EXACT else "FMA"),
                "fp32_fma_peak_tflops": fma_peak, "fp32_muladd_peak_tflops": muladd_peak,
            },
            "roofline_hbm": {
                "bound": "hbm", "achieved": algo_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": algo_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
            },
        }
        if other:
            line["other_variant"] = other
        if unscheduled:
            line["unscheduled"] = unscheduled
        if world == 1 and not args.no_cpu_baseline:
            try:
                c = reference_cpu_run(steps=1, warmup=0, budget_s=25.0)
                line["cpu_baseline"] = {"value": c["mrays"], "unit": "Mrays/s", "cores": c["cores"], "kind": c["kind"],
                                        "sample": c["sample"], "host_cores": c["host_cores"], "frames_per_s": c["frames_per_s"]}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(host_addr)
        del shared_host
        dist.barrier()
        if rank == 0:
            try:
                os.unlink(shm_path)
            except OSError:
                pass
    if p2p:
        torch.cuda.synchronize()
        if rank != 0:
            r.shared_frame_close(shared_ptr, owner=False)
        dist.barrier()
        if rank == 0:
            r.shared_frame_close(shared_ptr, owner=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries loaded later (NCCL
    prints its version banner there when NCCL_DEBUG is set) must not add to it:
    point fd 1 at stderr for the rest of the process and return a file on the real
    stdout for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="exact", choices=["exact", "fast"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "pixel", "persistent", "wavefront", "queued"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--composite", default="p2p", choices=["p2p", "nccl"], help="N>1: how bands reach rank 0")
    args = ap.parse_args()
    if args.gpus > 1 and "RANK" not in os.environ and args.impl != "reference":
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    args.out = _claim_stdout()
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
