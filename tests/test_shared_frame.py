"""The one-process-per-GPU composite (include/rt_cuda.h: rt_cuda_shared_frame_*,
opts->remote_fb / frame_seq / frame_ack) on whatever GPUs the box has -- ONE is
enough: ranks that share a device still go through cudaIpc, the staging frames,
the copy stream and the flag words, so the driver's single-GPU lease exercises
the code the 2/4/8-GPU bench runs (replaces the static column split of
src/main.c:363, 704-705).  Frames must be bit-identical to the one-rank frame,
including the pixels the reference's pass never writes (rows >= (h/scale)*scale,
main.c:285-290; columns >= T*column_w, main.c:363), whoever owns them."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CASES = {
    # name: (W, H, scale, num_columns, sweep_init_scale)
    "plain": (640, 364, 1, 1, 0),
    "rows_uncovered": (1920, 1080, 16, 1, 0),       # 1080 % 16 = 8 rows nobody renders; block 67 belongs to rank 67 % world
    "cols_uncovered": (1000, 360, 2, 3, 0),         # 1000 % 3 = 1 column nobody renders
    "sweep": (640, 360, 1, 1, 16),
}


def test_all_ranks_played_by_one_process(small_sky, builtin_objects):
    """Every rank's call issued by one process on one GPU: the owner path of the
    pipelined composite (local staging frames, copy stream, arrive / wait /
    release / ack flags) without any IPC."""
    from ray_tracing_b200 import host

    r = host.Renderer(num_gpus=1)
    try:
        r.upload_skybox(small_sky)
        r.upload_scene(builtin_objects[0])
        cam = host.Camera()
        for name, (W, H, scale, ncols, sweep) in CASES.items():
            if sweep:
                want, _ = r.render_sweep(cam, W, H, sweep)
            else:
                want, _ = r.render_frame(cam, W, H, scale, num_columns=ncols)
            for world in (3, 8):
                ptr, _ = r.shared_frame_create(W * H * 12)
                try:
                    # poison the frame: a pixel nobody writes (or clears) shows up
                    poison = np.full((H, W, 3), -7.0, np.float32)
                    r.copy_async(ptr, poison.ctypes.data, poison.nbytes)
                    r.synchronize()
                    for seq in (1, 2, 3):
                        for rank in range(world):
                            il = dict(interleave_count=world, interleave_index=rank, remote_fb=1, frame_seq=seq, frame_ack=1)
                            if sweep:
                                # a rank's sweep in the pipelined composite is self-contained (rt_api.cu: sweep_concurrent
                                # resolves the rank's own rows from its own cell buffers), so one process can play them in turn
                                r.render_sweep(cam, W, H, sweep, 0, ptr=ptr, stats=False, **il)
                            else:
                                r.render_into(cam, ptr, W, H, scale=scale, num_columns=ncols, **il)
                        r.shared_frame_wait(ptr, world, seq)
                        r.shared_frame_release(ptr, seq)
                    r.synchronize()
                    assert r.shared_frame_error(ptr) == 0
                    got = np.empty((H, W, 3), np.float32)
                    r.copy_to_host(got.ctypes.data, ptr, got.nbytes)
                    assert np.array_equal(bits(got), bits(want)), (name, world)
                finally:
                    r.shared_frame_close(ptr, owner=True)
    finally:
        r.close()


def _rank_main(rank, world, port_no, devices, case, tmp):
    import torch
    import torch.distributed as dist

    from oracle.bindings import procedural_skybox
    from ray_tracing_b200 import host, scenes

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dev = devices[rank]
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # control plane only: handle + barriers
    r = host.Renderer(device=dev)
    r.upload_skybox(procedural_skybox(64, seed=7))
    r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(0)))
    cam = host.Camera()
    W, H, scale, ncols, sweep = CASES[case]
    box = [None]
    if rank == 0:
        ptr, handle = r.shared_frame_create(W * H * 12)
        box[0] = handle
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        ptr = r.shared_frame_open(box[0])
    il = dict(interleave_count=world, interleave_index=rank, remote_fb=1)
    stream = torch.cuda.current_stream().cuda_stream
    consumer = torch.cuda.Stream(priority=-1)
    frames = []
    for seq in (1, 2, 3, 4):
        if sweep:
            r.render_sweep(cam, W, H, sweep, 0, ptr=ptr, stats=False, stream=stream, frame_seq=seq, frame_ack=1, **il)
        else:
            r.render_into(cam, ptr, W, H, scale=scale, num_columns=ncols, pass_index=seq - 1, stream=stream, frame_seq=seq, frame_ack=1, **il)
        if rank == 0:
            # consume every frame (copy it out) before handing the buffer back: frames differ (pass_index)
            r.shared_frame_wait(ptr, world, seq, stream=consumer.cuda_stream)
            out = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
            r.copy_async(out.data_ptr(), ptr, W * H * 12, stream=consumer.cuda_stream)
            r.shared_frame_release(ptr, seq, stream=consumer.cuda_stream)
            frames.append(out)
    r.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        assert r.shared_frame_error(ptr) == 0
        np.save(os.path.join(tmp, f"{case}.npy"), torch.stack(frames).cpu().numpy())
    dist.barrier()
    if rank != 0:
        r.shared_frame_close(ptr, owner=False)
    dist.barrier()
    if rank == 0:
        r.shared_frame_close(ptr, owner=True)
    r.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", list(CASES))
def test_rank_processes_pipelined_composite(tmp_path, small_sky, builtin_objects, case):
    """world ranks as separate processes (cudaIpc mapping of rank 0's frame); with
    fewer GPUs than ranks they share devices.  Four frames in flight back to back,
    each consumed by rank 0 before it releases the buffer."""
    import torch.multiprocessing as mp

    from ray_tracing_b200 import host

    n = max(_ngpu(), 1)
    world = 3 if n < 4 else 4
    devices = [i % n for i in range(world)]
    mp.spawn(_rank_main, args=(world, _free_port(), devices, case, str(tmp_path)), nprocs=world, join=True)
    W, H, scale, ncols, sweep = CASES[case]
    got = np.load(tmp_path / f"{case}.npy")
    r = host.Renderer(num_gpus=1)
    try:
        r.upload_skybox(small_sky)
        r.upload_scene(builtin_objects[0])
        for k in range(4):
            if sweep:
                want, _ = r.render_sweep(host.Camera(), W, H, sweep)
            else:
                want, _ = r.render_frame(host.Camera(), W, H, scale, num_columns=ncols, pass_index=k)
            assert np.array_equal(bits(got[k]), bits(want)), (case, k)
    finally:
        r.close()
