"""Multi-GPU invariance: the frame is bit-identical whether it
is rendered by 1 GPU or by row bands on several, composited on GPU 0 --
(a) inside one process (rt_cuda_init(n): band kernels store into GPU 0 over P2P),
(b) one process per GPU (torch.distributed/NCCL ranks): NCCL gather of bands and
    the fused cudaIpc peer-store composite.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@needs2
@pytest.mark.parametrize("W,H,s", [(640, 360, 1), (644, 364, 4)])
def test_single_process_bands_equal_one_gpu(small_sky, builtin_objects, W, H, s):
    from ray_tracing_b200 import host

    r = host.Renderer(num_gpus=1)
    r.upload_skybox(small_sky)
    r.upload_scene(builtin_objects[0])
    want, st1 = r.render_frame(host.Camera(), W, H, s)
    sweep1, _ = r.render_sweep(host.Camera(), W, H, 8)
    r.close()
    n = min(_ngpu(), 8)
    r = host.Renderer(num_gpus=n)
    assert r.num_gpus == n
    r.upload_skybox(small_sky)
    r.upload_scene(builtin_objects[0])
    got, stn = r.render_frame(host.Camera(), W, H, s)
    sweepn, _ = r.render_sweep(host.Camera(), W, H, 8)
    r.close()
    assert np.array_equal(bits(got), bits(want))
    assert stn["rays"] == st1["rays"]
    assert np.array_equal(bits(sweepn), bits(sweep1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port_no, W, H, scale, tmp):
    import torch
    import torch.distributed as dist

    from oracle.bindings import procedural_skybox
    from ray_tracing_b200 import host, scenes
    from ray_tracing_b200.distributed import band_rows, gather_bands

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    r = host.Renderer(device=rank)
    r.upload_skybox(procedural_skybox(64, seed=7))
    r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(0)))
    cam = host.Camera()
    r0, r1 = band_rows(H, scale, rank, world)
    stream = torch.cuda.current_stream().cuda_stream
    # (1) NCCL gather of bands
    band = torch.zeros((max(r1 - r0, 1), W, 3), dtype=torch.float32, device=dev)
    r.render_into(cam, band.data_ptr(), W, H, scale=scale, rows=(r0, r1), band_only_fb=1, stream=stream)
    full = gather_bands(band[: r1 - r0], H, W, scale, rank, world, dist, dst=0)
    if rank == 0:
        np.save(os.path.join(tmp, "nccl.npy"), full.cpu().numpy())
    # (2) fused composite: peer stores into rank 0's frame
    box = [None]
    if rank == 0:
        ptr, handle = r.shared_frame_create(W * H * 12)
        box[0] = handle
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        ptr = r.shared_frame_open(box[0])
    r.render_into(cam, ptr, W, H, scale=scale, interleave_count=world, interleave_index=rank, remote_fb=int(rank != 0), stream=stream)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    if rank == 0:
        out = np.zeros((H, W, 3), np.float32)
        r.copy_to_host(out.ctypes.data, ptr, out.nbytes)
        np.save(os.path.join(tmp, "p2p.npy"), out)
    dist.barrier()
    if rank != 0:
        r.shared_frame_close(ptr, owner=False)
    dist.barrier()
    if rank == 0:
        r.shared_frame_close(ptr, owner=True)
    r.close()
    dist.destroy_process_group()


@needs2
@pytest.mark.parametrize("scale", [1, 4])
def test_rank_per_gpu_composites_equal_one_gpu(tmp_path, small_sky, builtin_objects, scale):
    import torch.multiprocessing as mp

    from ray_tracing_b200 import host

    W, H = 640, 364
    world = min(_ngpu(), 4)
    mp.spawn(_rank_main, args=(world, _free_port(), W, H, scale, str(tmp_path)), nprocs=world, join=True)
    r = host.Renderer(num_gpus=1)
    r.upload_skybox(small_sky)
    r.upload_scene(builtin_objects[0])
    want, _ = r.render_frame(host.Camera(), W, H, scale)
    r.close()
    assert np.array_equal(bits(np.load(tmp_path / "nccl.npy")), bits(want))
    assert np.array_equal(bits(np.load(tmp_path / "p2p.npy")), bits(want))
