#!/usr/bin/env python
"""Headline benchmark of the per-pixel render path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--config 3]   # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...      # the reference's CPU path

Workloads (BASELINE.json `configs`); one *step* = one pass of the
hot path over one frame (config 4: one 16->1 progressive sweep):

    1   scene_0   1280x720   scale 1                     (the reference's own CPU-runnable case)
    2a  scene_1   1920x1080  scale 1
    2b  scene_2   1920x1080  scale 1
    3   scene_0   3840x2160  scale 1                     <- default: the configuration the metric names
    4   scene_0   1920x1080  passes at scale 16,8,4,2,1 accumulated and resolved (frames/s = sweeps/s)
    5   100 000 spheres (LBVH) 3840x2160 scale 1

Default camera pose, pass 0, the reference's skybox (6 x 2048^2, decoded by the
reference's own stb_image).  The same workload at every N (strong scaling): N=1
renders the whole frame, N>1 deals its rows to the ranks in 16-row blocks, round
robin, composited into rank 0's frame over NVLink.

Metric: Mrays/s = trace_ray-equivalent invocations (primary + bounce + shadow
rays, counted by the kernel that traced them) per second, whole job.

value  : frame resident on the device, K steps back to back, CUDA events on
         the launching stream, max over ranks.
e2e    : the same metric through the reference-facing C-ABI call with a HOST
         framebuffer (render_frame_cuda_ex -> Vector3 frame in pinned host
         memory): camera/params go host->device and the frame comes back
         device->host inside the timed region, every step.
configs: the line of the default config also carries a short device-timed
         measurement of every other config (N>1: configs 3, 4, 5), each with its
         own roofline fraction, clock sample and frame hash.
frame_sha256 : sha256 of rank 0's composited frame (f32x3, bottom row first),
         identical at every N; tests/golden/bench_frame_hashes.json holds the
         hashes of the same frames rendered by the unmodified reference.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STAGED = os.path.join(ROOT, "oracle", "_ref", "assets")
GOLDEN_HASHES = os.path.join(ROOT, "tests", "golden", "bench_frame_hashes.json")
N_SPHERES = 100_000

# FLOPs per ray, SURVEY.md 8(d): 15 + 18*spheres + 12*cubes for the linear scan
CONFIGS = {
    "1": dict(scene=0, w=1280, h=720, kind="frame", flops_per_ray=141,
              workload="scene_0.txt 1280x720 scale 1, default pose, pass 0 (BASELINE.json configs[0])"),
    "2a": dict(scene=1, w=1920, h=1080, kind="frame", flops_per_ray=105,
               workload="scene_1.txt 1920x1080 scale 1, default pose, pass 0 (BASELINE.json configs[1])"),
    "2b": dict(scene=2, w=1920, h=1080, kind="frame", flops_per_ray=69,
               workload="scene_2.txt 1920x1080 scale 1, default pose, pass 0 (BASELINE.json configs[1])"),
    "3": dict(scene=0, w=3840, h=2160, kind="frame", flops_per_ray=141,
              workload="scene_0.txt 3840x2160 scale 1, default pose, pass 0 (BASELINE.json configs[2]); N>1: the frame's rows dealt to the ranks in 16-row blocks, round robin, composited to rank 0"),
    "4": dict(scene=0, w=1920, h=1080, kind="sweep", init_scale=16, flops_per_ray=141,
              workload="scene_0.txt 1920x1080 progressive sweep: passes at scale 16,8,4,2,1 (pass_index 0..4) accumulated and resolved (BASELINE.json configs[3]); one step = one sweep"),
    "5": dict(scene="spheres", w=3840, h=2160, kind="frame", flops_per_ray=None,
              workload="synthetic 100 000-sphere scene (seed 20261017) 3840x2160 scale 1, default pose, pass 0, device LBVH with index tie-break (BASELINE.json configs[4])"),
}
MULTI_GPU_CONFIGS = ("3", "4", "5")


def _scenes_module():
    """ray_tracing_b200/scenes.py loaded by path (numpy only), so that the
    reference arm can emit scene files without importing the product package."""
    spec = importlib.util.spec_from_file_location("_rt_scenes", os.path.join(ROOT, "ray_tracing_b200", "scenes.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def scene_text(cfg) -> str:
    sc = _scenes_module()
    return sc.synthetic_spheres_text(N_SPHERES) if cfg["scene"] == "spheres" else sc.builtin_scene_text(cfg["scene"])


def load_skybox_faces():
    """The reference's six 2048^2 JPEG faces decoded by the reference's own decoder
    (stb_image through tools/librt_skybox_stb.so: the texels the parity tests use),
    else a procedural 2048^2 cubemap of the same size and format.
    Returns (faces u8 [6,h,w,3], description)."""
    from ray_tracing_b200 import host, scenes

    try:
        return host.load_skybox_dir(os.path.join(STAGED, "skybox")), "reference skybox JPEGs 6x2048x2048 decoded by stb_image (the reference's decoder)"
    except (FileNotFoundError, OSError):
        return scenes.procedural_skybox(2048, seed=11), "procedural 6x2048x2048 cubemap (reference JPEGs or stb helper not staged)"


def sha256_frame(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def golden_hash(config: str, sky_desc: str):
    """Hash of the same frame rendered by the unmodified reference (tests/golden/
    make_bench_hashes.py); only meaningful with the reference's skybox."""
    if not sky_desc.startswith("reference skybox"):
        return None
    try:
        return json.load(open(GOLDEN_HASHES))["frames"].get(config)
    except Exception:
        return None


# --------------------------------------------------------------------- clocks


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int, period_ms: int = 100):
        self.path = f"/tmp/rt_bench_clocks_{os.getpid()}.csv"
        self.proc = None
        self.gpu = gpu_index
        self.period_ms = period_ms
        self.offset = 0

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms), "-i", str(self.gpu)],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def _parse(self, text):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in text.splitlines():
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
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out

    def window(self):
        """Samples written since the last call (one clock record per measured config)."""
        if self.proc is None:
            return self._parse("")
        try:
            self.f.flush()
            with open(self.path) as fh:
                fh.seek(self.offset)
                text = fh.read()
                self.offset = fh.tell()
            return self._parse(text)
        except Exception:
            return self._parse("")

    def stop(self):
        if self.proc is None:
            return self._parse("")
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        self.f.close()
        try:
            text = open(self.path).read()
            os.unlink(self.path)
        except Exception:
            text = ""
        return self._parse(text)


# ------------------------------------------------------------------ reference


def reference_cpu_run(config: str, steps: int, warmup: int, budget_s: float = 150.0):
    """The reference's own CPU implementation of the path (oracle/_ref: the
    unmodified TUs compiled from /root/reference; else the oracle port) on all
    host cores, on a bounded sample of the workload.  Imports nothing from the
    product package: scene files are parsed by the reference's parser and the
    skybox is decoded by the reference's loader."""
    from oracle import bindings

    cfg = CONFIGS[config]
    cores = os.cpu_count() or 1
    big = cfg["scene"] == "spheres"
    use_ref = bindings.ref_available("pixel_big") if big else (bindings.ref_available("stream") and bindings.ref_available("count"))
    scene_path = os.path.join(STAGED, f"scene_{cfg['scene']}.txt") if not big else f"/tmp/rt_bench_spheres_{N_SPHERES}.txt"
    if big or not os.path.exists(scene_path):
        if not big:
            scene_path = f"/tmp/rt_bench_scene_{cfg['scene']}.txt"
        with open(scene_path, "w") as fh:
            fh.write(scene_text(cfg))

    def threads_for(width):
        # render_column needs W % T == 0 to cover the frame (main.c:363)
        return max(t for t in range(1, min(cores, width) + 1) if width % t == 0)

    if use_ref:
        refs = [bindings.Ref("pixel_big")] if big else [bindings.Ref("stream"), bindings.Ref("count")]
        faces = None
        for r in refs:
            if not r.parse_scene_file(scene_path):
                raise RuntimeError(f"the reference could not parse {scene_path}")
            r.reset_camera()
            try:
                faces = r.load_skybox()
                sky_desc = "reference skybox JPEGs 6x2048x2048 loaded by the reference's load_cubemap"
            except FileNotFoundError:
                faces = bindings.procedural_skybox(512, seed=3)
                r.set_skybox(faces)
                sky_desc = "procedural 6x512x512 cubemap (reference JPEGs not staged)"
        kind = "reference"
    else:
        # oracle/_ref travels with the repo; without it the port has no scene parser of its own
        raise RuntimeError("oracle/_ref is not built (make -C oracle ref where /root/reference exists)")

    passes = [(16, 0), (8, 1), (4, 2), (2, 3), (1, 4)] if cfg["kind"] == "sweep" else [(1, 0)]

    def run_once(w, h, counted=False):
        secs, rays = 0.0, 0
        for scale, p in passes:
            if big:
                _, s, r = refs[0].render(w, h, scale, threads_for(w), p, keyed=True)
            else:
                _, s, r = (refs[1] if counted else refs[0]).render(w, h, scale, threads_for(w), p, keyed=False)
            secs += s
            rays += r
        return secs, rays

    # the largest sample of the workload that keeps the whole run bounded
    W, H = cfg["w"], cfg["h"]
    sizes = [(W, H)] + [(W // k, H // k) for k in (2, 3, 4, 6, 8, 12, 16, 24, 32) if W % k == 0 and H % k == 0]
    pw, ph = sizes[-1]
    probe_s, _ = run_once(pw, ph)
    w, h = pw, ph
    for cw, ch in sizes:
        est = probe_s * (cw * ch) / (pw * ph)
        if est * (steps + warmup + 1) <= budget_s:
            w, h = cw, ch
            break
    _, rays = run_once(w, h, counted=True)        # ray count of this sample (untimed for the stream build)
    for _ in range(warmup):
        run_once(w, h)
    times = []
    for _ in range(steps):
        s, r = run_once(w, h)
        times.append(s)
        if big:
            rays = r
    total = float(sum(times))
    threads = threads_for(w)
    what = "16->1 sweep (5 passes)" if cfg["kind"] == "sweep" else "scale 1 pass 0"
    name = f"{N_SPHERES} spheres (O(N) scan per ray: the reference has no acceleration structure)" if big else f"scene_{cfg['scene']}"
    return dict(
        mrays=rays * len(times) / total / 1e6, ms_per_step=1e3 * total / len(times), frames_per_s=len(times) / total,
        kind=kind, cores=threads, host_cores=cores, rays_per_step=int(rays),
        sample=f"{name} {w}x{h} {what}, full frame, {threads} threads (one render_column per thread), {sky_desc}",
    )


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        # RT_BENCH_REF_BUDGET_S bounds the CPU time of the whole run (tests use a few seconds)
        r = reference_cpu_run(args.config, args.steps, args.warmup, budget_s=float(os.environ.get("RT_BENCH_REF_BUDGET_S", "150")))
    except Exception as e:  # the oracle always exists; report rather than crash the driver
        print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}), file=args.out, flush=True)
        return 0
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": r["mrays"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config]["workload"], "config": args.config, "sample": r["sample"]},
        "frames_per_s": r["frames_per_s"],
        "cpu_baseline": {"value": r["mrays"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["mrays"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


# ------------------------------------------------------------------------ GPU


class GpuBench:
    """One process per GPU.  N>1: every rank renders the row blocks it owns into
    one of two local frames and ships them to rank 0's shared frame on its copy
    stream while the next frame renders (rt_cuda.h: opts->frame_seq); rank 0's
    consumer stream waits for all ranks' blocks of a frame and releases it."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from ray_tracing_b200 import host

        self.torch, self.dist, self.host, self.args = torch, dist, host, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.r = host.Renderer(device=self.local)
        self.faces, self.sky_desc = load_skybox_faces()
        self.r.upload_skybox(self.faces)
        self.cam = host.Camera()
        self.variant = host.RT_VARIANT_FAST if args.variant == "fast" else host.RT_VARIANT_EXACT
        self.kernel = {"auto": host.RT_KERNEL_AUTO, "pixel": host.RT_KERNEL_PIXEL, "persistent": host.RT_KERNEL_PERSISTENT,
                       "wavefront": host.RT_KERNEL_WAVEFRONT, "queued": host.RT_KERNEL_QUEUED}[args.kernel]
        self.render_stream = torch.cuda.current_stream()
        # frames in flight: consecutive frames are issued on these streams round robin (each into its own
        # frame buffer), so that the tail of frame k -- a few long paths on a nearly idle GPU -- runs under
        # the head of frame k+1.  render_streams[0] is the timing stream.
        self.render_streams = [self.render_stream] + [torch.cuda.Stream() for _ in range(max(args.streams, 1) - 1)]
        self.consumer_stream = torch.cuda.Stream(priority=-1)
        self.scene_loaded = None
        self.shared = {}        # frame bytes -> (ptr, seq)
        self.launches = 0

    # -- plumbing
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_int(self, x: int) -> int:
        t = self.torch.tensor([x], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def load_scene(self, cfg):
        key = cfg["scene"]
        if self.scene_loaded == key:
            return
        text = scene_text(cfg)
        objs = self.host.parse_scene_string_large(text) if key == "spheres" else self.host.parse_scene_string(text)
        self.r.upload_scene(objs)
        self.r.synchronize()
        self.scene_loaded = key

    def shared_frame(self, nbytes):
        """Rank 0's frame (device memory) mapped into every rank (cudaIpc)."""
        if nbytes in self.shared:
            return self.shared[nbytes]
        box = [None]
        if self.rank == 0:
            ptr, handle = self.r.shared_frame_create(nbytes)
            box[0] = handle
        self.dist.broadcast_object_list(box, src=0)
        if self.rank != 0:
            ptr = self.r.shared_frame_open(box[0])
        self.shared[nbytes] = [ptr, 0]
        return self.shared[nbytes]

    def close(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            for ptr, _ in self.shared.values():
                if self.rank != 0:
                    self.r.shared_frame_close(ptr, owner=False)
            self.dist.barrier()
            if self.rank == 0:
                for ptr, _ in self.shared.values():
                    self.r.shared_frame_close(ptr, owner=True)
        self.r.close()
        if self.world > 1:
            self.dist.destroy_process_group()

    # -- one config, device resident
    def make_step(self, cfg, fb_format=None):
        """Returns (step(), finish(), frame_to_host(), rays_per_step, launches_per_step)."""
        torch, host, r, cam = self.torch, self.host, self.r, self.cam
        W, H = cfg["w"], cfg["h"]
        common = dict(variant=self.variant, kernel=self.kernel)
        sweep = cfg["kind"] == "sweep"
        bpp = 12
        # a sweep owns the library's accumulation and cell buffers: one stream; plain frames: all of them
        nst = 1 if sweep else len(self.render_streams)
        streams = [s_.cuda_stream for s_ in self.render_streams[:nst]]
        if self.world == 1:
            frames = [torch.empty((H, W, 3), dtype=torch.float32, device=self.dev) for _ in range(nst)]
            frame = frames[0]

            def issue(k, stats=False):
                # pass k of the pose: the reference's accumulation draws fresh samples every pass (main.c:354-403)
                if sweep:
                    _, st = r.render_sweep(cam, W, H, cfg["init_scale"], 5 * k, ptr=frame.data_ptr(), stats=stats, stream=streams[0], **common)
                    return st
                return r.render_into(cam, frames[k % nst].data_ptr(), W, H, stats=stats, stream=streams[k % nst], scale=1, pass_index=k, **common)

            st = issue(0, stats=True)
            rays = st["rays"]
            launches = st["kernel_launches"]
            counter = [0]

            def step():
                issue(counter[0])
                counter[0] += 1

            def finish():
                pass

            def to_host():
                issue(0)                       # the frame that is hashed: pass 0
                torch.cuda.synchronize()
                return frame.cpu().numpy()

            return step, finish, to_host, rays, launches, rays

        shared = self.shared_frame(W * H * bpp)
        ptr = shared[0]
        il = dict(interleave_count=self.world, interleave_index=self.rank, remote_fb=1)
        rs = self.render_stream.cuda_stream
        cs = self.consumer_stream.cuda_stream

        def issue_stats():
            # synchronous call (copy follows the render on the same stream): exact ray count of this rank
            if sweep:
                _, st = r.render_sweep(cam, W, H, cfg["init_scale"], 0, ptr=ptr, stats=True, **il, **common)
                return st
            return r.render_into(cam, ptr, W, H, stats=True, scale=1, pass_index=0, **il, **common)

        st = issue_stats()
        my_rays = st["rays"]
        rays = self.allsum_int(my_rays)
        launches = st["kernel_launches"] + 3       # + arrive flag, ack poll, (rank 0) wait/release
        self.barrier()

        counter = [0]

        def step(k=None):
            shared[1] += 1
            seq = shared[1]
            if k is None:
                k = counter[0]
                counter[0] += 1
            if sweep:
                r.render_sweep(cam, W, H, cfg["init_scale"], 5 * k, ptr=ptr, stats=False, stream=rs, frame_seq=seq, frame_ack=1, **il, **common)
            else:
                r.render_into(cam, ptr, W, H, stream=streams[seq % nst], scale=1, pass_index=k, frame_seq=seq, frame_ack=1, **il, **common)
            if self.rank == 0:
                # the consumer: frame seq is whole once every rank's blocks have landed; hand it back at once
                r.shared_frame_wait(ptr, self.world, seq, stream=cs)
                r.shared_frame_release(ptr, seq, stream=cs)

        def finish():
            # the last frame is only complete when rank 0's consumer has seen it
            if self.rank == 0:
                self.render_stream.wait_stream(self.consumer_stream)

        def to_host():
            step(0)                            # the frame that is hashed: pass 0
            finish()
            self.barrier()
            r.synchronize()
            self.barrier()
            if self.rank != 0:
                return None
            out = np.empty((H, W, 3), np.float32)
            r.copy_to_host(out.ctypes.data, ptr, out.nbytes)
            err = r.shared_frame_error(ptr)
            if err:
                raise RuntimeError(f"pipelined composite: a flag poll timed out (error word {err})")
            return out

        return step, finish, to_host, rays, launches, my_rays

    def time_steps(self, step, finish, steps, warmup):
        """Returns (ms per step, rays per step): max over ranks of the device time, and the rays the
        kernels counted over exactly the timed steps (all ranks)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        finish()
        self.r.synchronize()
        self.barrier()
        rays0 = self.r.ray_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.render_stream)
        for s_ in self.render_streams[1:]:
            s_.wait_event(e0)
        for _ in range(steps):
            step()
        for s_ in self.render_streams[1:]:
            self.render_stream.wait_stream(s_)
        finish()
        e1.record(self.render_stream)
        self.r.synchronize()
        self.barrier()
        rays = self.allsum_int(self.r.ray_counter() - rays0)
        return self.allmax(e0.elapsed_time(e1)) / steps, rays / steps

    def measure_config(self, name, steps, warmup, sampler=None, min_ms=None):
        """Device-resident measurement of one config: ms per step, Mrays/s, frame hash.
        min_ms: lengthen the timed region to at least this long (so that the 50 ms clock
        sampler sees it): `steps` becomes a lower bound."""
        cfg = CONFIGS[name]
        self.load_scene(cfg)
        step, finish, to_host, rays, launches, my_rays = self.make_step(cfg)
        if min_ms:
            est, _ = self.time_steps(step, finish, 5, 3)
            steps = int(min(max(steps, min_ms / max(est, 1e-3)), 20000))
            self.launches += launches * 8
        if sampler:
            sampler.window()
        ms, rays = self.time_steps(step, finish, steps, max(warmup, 3))
        clocks = sampler.window() if sampler else None
        frame = to_host()
        out = dict(config=name, workload=cfg["workload"], ms_per_step=ms, rays_per_step=rays, value=rays / (ms * 1e-3) / 1e6, unit="Mrays/s",
                   frames_per_s=1e3 / ms, steps=steps, launches_per_step=launches, clocks=clocks,
                   passes="step k renders pass k of the pose (fresh samples every pass, as the reference's accumulation); rays_per_step = rays counted by the kernels over the timed steps / steps; the hashed frame is pass 0")
        if self.rank == 0:
            out["frame_sha256"] = sha256_frame(frame)
            want = golden_hash(name, self.sky_desc)
            if want is not None:
                out["frame_matches_reference"] = out["frame_sha256"] == want
        self.launches += launches * (steps + max(warmup, 3))
        if cfg["kind"] == "sweep" and self.world == 1:
            # the same sweeps with the passes one after the other (what every sweep did before the
            # concurrent sweep, and what ranks of a multi-GPU run still do), for the record
            self.r.set_concurrent_sweep(False)
            sms, _ = self.time_steps(step, finish, max(steps // 2, 5), 3)
            self.r.set_concurrent_sweep(True)
            self.launches += launches * (max(steps // 2, 5) + 3)
            out["schedule"] = "the five passes run side by side on separate streams, one resolve kernel folds them in pass order (rt_api.cu: sweep_concurrent)"
            out["passes_one_after_the_other"] = {"ms_per_step": sms, "frames_per_s": 1e3 / sms}
        return out, (step, finish, to_host, rays, my_rays)


def lbvh_flops_per_ray():
    """15 + 24*(internal nodes visited: two child boxes each, 12 flops per box) + 18*(sphere tests), from the
    counter build (profiles/r02_lbvh_counts.json, tools/lbvh_ab.py --counts); SURVEY.md 8(d)."""
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", "r02_lbvh_counts.json")))
        return 15 + 24 * c["nodes_per_ray"] + 18 * c["tests_per_ray"], c
    except Exception:
        return None, None


def main_gpu(args):
    b = GpuBench(args)
    torch, host, r = b.torch, b.host, b.r
    rank, world = b.rank, b.world
    if world != args.gpus and world > 1:
        args.gpus = world
    name = args.config
    cfg = CONFIGS[name]
    W, H = cfg["w"], cfg["h"]
    exact = b.variant == host.RT_VARIANT_EXACT

    sampler = ClockSampler(b.local, period_ms=50)
    if rank == 0:
        sampler.start()

    # ---- device-resident timing of the headline config --------------------
    main, (step, finish, to_host, rays_per_step, my_rays) = b.measure_config(name, args.steps, args.warmup, sampler if rank == 0 else None)
    per_step_ms = main["ms_per_step"]

    # kernel-only time of the dominant kernel (render) on this rank: CUDA events on the
    # launching stream around launches into a local frame, no composite
    b.load_scene(cfg)
    local = torch.empty((H, W, 3), dtype=torch.float32, device=b.dev)
    il = dict(interleave_count=world, interleave_index=rank) if world > 1 else {}
    kopts = dict(variant=b.variant, kernel=b.kernel, stream=b.render_stream.cuda_stream, **il)

    kcount = [0]

    def kernel_only(**over):
        k = kcount[0]
        kcount[0] += 1
        o = dict(kopts, **over)
        if cfg["kind"] == "sweep":
            r.render_sweep(b.cam, W, H, cfg["init_scale"], 5 * k, ptr=local.data_ptr(), stats=False, **o)
        else:
            r.render_into(b.cam, local.data_ptr(), W, H, scale=1, pass_index=k, **o)

    def time_local(fn, n):
        """(ms per launch, rays per launch) of n back-to-back launches on this rank."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        rays0 = r.ray_counter()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(b.render_stream)
        for _ in range(n):
            fn()
        k1.record(b.render_stream)
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / n, (r.ray_counter() - rays0) / n

    kern_ms_local, kern_rays = time_local(kernel_only, args.steps)
    kern_ms = b.allmax(kern_ms_local)
    my_rays = kern_rays

    # ---- the same launches with tiles in image order (what a pose costs the first two
    # times it is rendered, and every time while the camera moves), for the record ----
    unscheduled = None
    if args.kernel in ("auto", "queued") and cfg["kind"] == "frame":
        r.set_tile_schedule(False)
        ums, urays = b.time_steps(step, finish, 10, 3)
        unscheduled = {"ms_per_step": ums, "value": urays / (ums * 1e-3) / 1e6, "unit": "Mrays/s",
                       "note": "tiles handed out in image order (no cost-sorted schedule)"}
        r.set_tile_schedule(True)

    # ---- a moving camera: every pose is new, rendered coarse to fine as update_frame() does
    # (main.c:354, 402-403).  The scale-1 pass of a fresh pose is ordered by the tile costs its
    # scale-2 pass recorded; compared with the same pass with tiles in image order ----
    moving = None
    if world == 1 and cfg["kind"] == "frame" and args.kernel in ("auto", "queued"):
        def fresh_pose_scale1_ms(schedule):
            r.set_tile_schedule(schedule)
            ms = []
            for i in range(8):
                cam = host.Camera((5.0 + 0.03 * (i + 1), 5.0, 5.0), (-1.0, -1.0, -1.0), (0, 1, 0), 30.0)   # the same poses both ways; toggling the schedule forgets them
                for sc in (8, 4, 2):
                    r.render_into(cam, local.data_ptr(), W, H, stats=True, scale=sc, pass_index=0, variant=b.variant, kernel=b.kernel)
                ms.append(r.render_into(cam, local.data_ptr(), W, H, stats=True, scale=1, pass_index=0, variant=b.variant, kernel=b.kernel)["render_ms"])
            r.set_tile_schedule(True)
            return float(np.median(ms))
        seeded, plain, recording = fresh_pose_scale1_ms(2), fresh_pose_scale1_ms(False), fresh_pose_scale1_ms(True)
        moving = {"scale1_ms_seeded_by_scale2": seeded, "scale1_ms_image_order": plain, "scale1_ms_default": recording,
                  "note": "first scale-1 pass of a NEW pose (8 poses, median), preceded by its scale 8, 4, 2 passes as in update_frame(); device time of the pass.  default = image order while recording the tile costs that order the pose's later scale-1 passes; seeded = ordered by the costs of the scale-2 pass (rt_cuda_debug_set_tile_schedule(2): measured slower, not the default)"}

    # ---- the other build of the same kernels, for the record (N=1 only) ----
    other = None
    if world == 1 and cfg["kind"] == "frame":
        ov = host.RT_VARIANT_EXACT if not exact else host.RT_VARIANT_FAST
        oms, orays = time_local(lambda: kernel_only(variant=ov), 10)
        other = {"variant": "exact" if ov == host.RT_VARIANT_EXACT else "fast", "ms_per_step": oms, "value": orays / (oms * 1e-3) / 1e6, "unit": "Mrays/s",
                 "note": "fast = FMA contraction + approximate rcp/rsqrt, f32 sphere roots: <= 1 LSB/8-bit channel on >= 99.9 % of pixels; exact = bit-identical to the reference"}

    # ---- end to end through the C ABI with a host framebuffer --------------
    e2e = measure_e2e(b, cfg, rays_per_step, args)

    # ---- every other config, device resident, in the same process ----------
    others = {}
    if name == "3" and not args.no_other_configs:
        names = [c for c in CONFIGS if c != name and (world == 1 or c in MULTI_GPU_CONFIGS)]
        for c in names:
            m, _ = b.measure_config(c, 5, 3, sampler if rank == 0 else None, min_ms=300.0)
            others[c] = m
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        value = main["value"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # FP32 issue-rate ceiling measured live with register-only chains
        fma_peak = r.fp32_peak_tflops(True)
        muladd_peak = r.fp32_peak_tflops(False)
        peak = muladd_peak if exact else fma_peak

        def fp32_roofline(c, rays, ms, gpus=1):
            """`rays` in `ms` on `gpus` GPUs against `gpus` times the peak measured on this one"""
            fpr, counts = (CONFIGS[c]["flops_per_ray"], None) if CONFIGS[c]["flops_per_ray"] else lbvh_flops_per_ray()
            if not fpr:
                return None
            ach = rays * fpr / (ms * 1e-3) / 1e12
            d = {"bound": "fp32", "achieved": ach, "peak": peak * gpus, "unit": "TFLOP/s", "frac": ach / (peak * gpus) if peak else None, "flops_per_ray": fpr}
            if gpus > 1:
                d["gpus"] = gpus
            if counts:
                d["flops_per_ray_source"] = "15 + 24*nodes + 18*sphere tests per ray, counter build (profiles/r02_lbvh_counts.json): %.1f nodes, %.2f tests" % (counts["nodes_per_ray"], counts["tests_per_ray"])
            return d

        kernel_name = {"auto": "render_queued_kernel", "queued": "render_queued_kernel", "persistent": "render_persistent_kernel",
                       "wavefront": "render_wavefront_kernel", "pixel": "render_pixel_kernel"}[args.kernel]
        roof = fp32_roofline(name, my_rays, kern_ms) or {"bound": "fp32", "achieved": None, "peak": peak, "unit": "TFLOP/s", "frac": None}
        band_px = (W * H) // world
        algo_bytes = band_px * 12 + band_px * 32     # Vector3 store + one 32 B skybox sector per escaping path
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_src = None, None
        if world == 1:
            # DRAM bytes of one launch cannot be counted inside this process; the figure of the
            # committed `ncu --set full` capture of this kernel on this workload is quoted, with its source
            try:
                t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
                traffic, traffic_src = t.get(name, {}).get("dram_bytes_per_launch"), t.get(name, {}).get("source")
            except Exception:
                pass
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name, "kernel_ms": kern_ms, "rays_per_launch": my_rays,
                     "peak_source": "measured live: register-only %s chains on this GPU" % ("MUL+ADD (no-FMA exact build)" if exact else "FMA"),
                     "fp32_fma_peak_tflops": fma_peak, "fp32_muladd_peak_tflops": muladd_peak})
        for c, m in others.items():
            m["roofline"] = fp32_roofline(c, m["rays_per_step"], m["ms_per_step"], world)      # whole-job rays over all ranks
        composite = "none (1 GPU)" if world == 1 else (
            "pipelined P2P: every rank renders its 16-row blocks (round robin) into one of two local frames and ships them with one strided "
            "peer copy on its copy stream into rank 0's frame over NVLink (cudaIpc mapping) while the next frame renders; a flag word per rank in "
            "rank 0's memory marks a frame's blocks as arrived, rank 0's consumer stream waits for all of them and releases the frame (no collective)")
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": cfg["workload"], "config": name, "variant": args.variant, "kernel": args.kernel, "skybox": b.sky_desc,
                "composite": composite,
                "framebuffer": "Vector3 f32x3 (reference frame format), bottom row first",
                "l2": "no explicit flush: every step reads the 96 MiB RGBA8 skybox at random and writes a %.1f MB frame (working set > 126 MB L2 at 4K)" % (W * H * 12 / 1e6),
                "rays_per_step": rays_per_step, "pixels_per_step": W * H,
                "frames_in_flight": "consecutive frames are issued on %d CUDA streams round robin, each into its own frame buffer (N > 1: its own staging frame), so the tail of frame k overlaps the head of frame k+1; every frame is complete and composited before the timed region ends (--streams 1: one after the other)" % len(b.render_streams),
                "schedule": "every step renders the same pose, as the reference's progressive accumulation does (main.c:354-403); from the third launch of a pose the queued kernel hands out its 8x4 tiles longest-first, by the per-tile bounce counts the second launch recorded (warm-up). Scheduling only: frames are bit-identical. `unscheduled` = the same loop with tiles in image order",
            },
            "frames_per_s": 1e3 / per_step_ms,
            "mpix_per_s": W * H / (per_step_ms * 1e-3) / 1e6,
            "frame_sha256": main.get("frame_sha256"),
            "frame_matches_reference": main.get("frame_matches_reference"),
            "e2e": e2e,
            "gpu_launches": b.launches,
            "clocks": clocks,
            "clocks_timed_region": main["clocks"],
            "roofline": roof,
            "roofline_hbm": {
                "bound": "hbm", "achieved": algo_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": algo_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
            },
        }
        if others:
            line["configs"] = others
        if other:
            line["other_variant"] = other
        if unscheduled:
            line["unscheduled"] = unscheduled
        if moving:
            line["moving_camera"] = moving
        if world == 1 and not args.no_cpu_baseline:
            try:
                c = reference_cpu_run(name, steps=1, warmup=0, budget_s=25.0)
                line["cpu_baseline"] = {"value": c["mrays"], "unit": "Mrays/s", "cores": c["cores"], "kind": c["kind"],
                                        "sample": c["sample"], "host_cores": c["host_cores"], "frames_per_s": c["frames_per_s"]}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), file=args.out, flush=True)
    b.close()
    return 0


def measure_e2e(b, cfg, rays_per_step, args):
    """The same metric through render_frame_cuda_ex with a HOST frame: pipelined
    (headline), synchronous, and pipelined with the 8-bit frame format."""
    torch, host, r = b.torch, b.host, b.r
    W, H = cfg["w"], cfg["h"]
    world, rank = b.world, b.rank
    sweep = cfg["kind"] == "sweep"
    b.load_scene(cfg)
    opts = dict(variant=b.variant, kernel=b.kernel, stream=b.render_stream.cuda_stream)

    shm_path, shared_host, host_addr = None, None, None
    host_frames = None
    if world == 1:
        host_frames = {12: [torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)],
                       4: [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]}
    else:
        # one host frame shared by all ranks (POSIX shared memory, page-locked in every process):
        # each rank copies the row blocks it rendered over its own PCIe link, no gather through GPU 0
        shm_path = f"/dev/shm/rt_bench_frame_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            np.lib.format.open_memmap(shm_path, mode="w+", dtype=np.float32, shape=(H, W, 3)).flush()
        b.dist.barrier()
        shared_host = np.load(shm_path, mmap_mode="r+")
        host_addr = shared_host.ctypes.data
        err = torch.cuda.cudart().cudaHostRegister(host_addr, shared_host.nbytes, 0)
        assert int(err) == 0, f"cudaHostRegister failed: {err}"
        # the same for the 8-bit frame format (a third of the bytes into host memory)
        shm8_path = shm_path + "_u8"
        if rank == 0:
            np.lib.format.open_memmap(shm8_path, mode="w+", dtype=np.uint8, shape=(H, W, 4)).flush()
        b.dist.barrier()
        shared_host8 = np.load(shm8_path, mmap_mode="r+")
        host_addr8 = shared_host8.ctypes.data
        err = torch.cuda.cudart().cudaHostRegister(host_addr8, shared_host8.nbytes, 0)
        assert int(err) == 0, f"cudaHostRegister failed: {err}"

    def step(i, pipeline, fmt):
        bpp = 4 if fmt == host.RT_FB_U8X4 else 12
        if sweep:
            # the drop-in's update_frame(): the final pass of the sweep copies the resolved frame to the host
            if world == 1:
                out = host_frames[bpp][i & 1]
                r.render_sweep(b.cam, W, H, cfg["init_scale"], 0, ptr=out.data_ptr(), stats=False, host=True, fb_format=fmt, **opts)
            else:
                r.render_sweep(b.cam, W, H, cfg["init_scale"], 0, ptr=host_addr, stats=False, host=True, interleave_count=world, interleave_index=rank, **opts)
                b.dist.barrier()
            return
        if world == 1:
            # the drop-in call: params go H2D as kernel arguments, the frame comes back D2H.
            # pipeline=1: the call returns once the copy is queued; frame i+1 renders while frame i drains
            r.render_into(b.cam, host_frames[bpp][i & 1].data_ptr(), W, H, host=True, pipeline=int(pipeline), scale=1, pass_index=0, fb_format=fmt, **opts)
        else:
            r.render_into(b.cam, host_addr8 if bpp == 4 else host_addr, W, H, host=True, pipeline=int(pipeline), scale=1, pass_index=0, fb_format=fmt,
                          interleave_count=world, interleave_index=rank, **opts)
            if not pipeline:
                b.dist.barrier()      # the frame is whole once every rank's copy has landed

    def run(n, pipeline, fmt=host.RT_FB_F32X3):
        for i in range(3):
            step(i, pipeline, fmt)
        r.synchronize()
        b.barrier()
        t0 = time.perf_counter()
        for i in range(n):
            step(i, pipeline, fmt)
        r.synchronize()
        b.barrier()
        return b.allmax(time.perf_counter() - t0)

    n = max(3, args.steps)        # the K steps of the contract: the pipeline's drain (one frame copy) is inside the timed region
    sync_s = run(n, False)
    pipe_s = run(n, True) if not sweep else sync_s
    u8_s = run(n, True, host.RT_FB_U8X4) if (world == 1 or not sweep) else None
    # the frame the host ends up with, through the synchronous call
    step(0, False, host.RT_FB_F32X3)
    r.synchronize()
    b.barrier()
    frame_sha = None
    if rank == 0:
        frame_sha = sha256_frame(host_frames[12][0].numpy() if world == 1 else np.array(shared_host))
    if world > 1:
        b.dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(host_addr)
        torch.cuda.cudart().cudaHostUnregister(host_addr8)
        del shared_host, shared_host8
        b.dist.barrier()
        if rank == 0:
            for path in (shm_path, shm8_path):
                try:
                    os.unlink(path)
                except OSError:
                    pass
    mr = lambda s: rays_per_step * n / s / 1e6
    out = {
        "value": mr(pipe_s), "unit": "Mrays/s",
        "h2d_bytes_per_step": int(host.load_library().rt_cuda_param_bytes()) * world * (5 if sweep else 1),   # kernel-argument block (camera frame, views, sizes) per rank and launch; scene and skybox are resident
        "d2h_bytes_per_step": W * H * 12,
        "frames_per_s": n / pipe_s, "steps": n,
        "mode": ("synchronous update_frame(): the sweep's last pass copies the resolved frame to the host" if sweep else
                 ("pipelined (opts.pipeline = 1): call k+1 renders while the copy stream drains frame k into the other pinned host frame; timed until rt_cuda_synchronize()" if world == 1 else
                  "pipelined like N=1; every rank renders its row blocks and copies them over its own PCIe link into one page-locked host frame shared by the ranks (POSIX shm)")),
        "sync": {"value": mr(sync_s), "unit": "Mrays/s", "frames_per_s": n / sync_s,
                 "mode": "synchronous call (what INTEGRATION.md's update_frame() binding does): returns when the frame is in host memory" + ("; N>1: plus a barrier per frame" if world > 1 else "")},
        "sync_value": mr(sync_s),
        "frame_sha256": frame_sha,
        "api": "render_frame_cuda_ex(cam, host Vector3 frame, w, h, opts)" if not sweep else "rt_cuda_render_sweep(cam, host Vector3 frame, w, h, 16, ...)",
    }
    if u8_s:
        out["u8x4"] = {"value": mr(u8_s), "unit": "Mrays/s", "frames_per_s": n / u8_s, "d2h_bytes_per_step": W * H * 4,
                       "mode": "pipelined, RT_FB_U8X4 frame ((uint8_t)(x*255) per channel, what screenshot() and a display need): a third of the bytes"}
    return out


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
    ap.add_argument("--config", default="3", choices=sorted(CONFIGS))
    ap.add_argument("--variant", default="exact", choices=["exact", "fast"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "pixel", "persistent", "wavefront", "queued"])
    ap.add_argument("--streams", type=int, default=3, help="frames in flight: consecutive frames are issued on this many streams round robin (1 = one after the other)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="config 3 only: skip the short measurements of the other configs")
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
