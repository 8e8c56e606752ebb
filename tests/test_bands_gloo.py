"""Host logic of the multi-GPU path: row-band partition and the gather to
rank 0, exercised with world_size-2/3 gloo process groups on the CPU.  The
bands are rendered by the oracle port here (tests may use the oracle); on the
GPU box tests/test_gpu_parity.py covers the same partition through the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ray_tracing_b200.distributed import all_bands, band_rows, gather_bands, owned_rows


@pytest.mark.parametrize("h,scale,world", [(1080, 1, 8), (1080, 16, 8), (2160, 1, 4), (90, 4, 3), (17, 2, 2), (8, 8, 4)])
def test_bands_tile_the_frame(h, scale, world):
    bands = all_bands(h, scale, world)
    assert bands[0][0] == 0 and bands[-1][1] == h
    for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
        assert a1 == b0 and a0 <= a1
    for r0, r1 in bands:
        assert r0 % scale == 0 and (r1 % scale == 0 or r1 == h)
    sizes = [(r1 - r0 + scale - 1) // scale for r0, r1 in bands]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("h,scale,world", [(2160, 1, 8), (1080, 16, 8), (1080, 4, 3), (360, 3, 2), (200, 8, 5)])
def test_interleaved_blocks_partition_the_frame(h, scale, world):
    seen = np.zeros(h, np.int32)
    sizes = []
    for r in range(world):
        rows = owned_rows(h, scale, r, world)
        seen[rows] += 1
        sizes.append(len(rows))
    covered = (h // scale) * scale
    assert (seen[:covered] == 1).all() and (seen[covered:] == 0).all()
    block = 16 if 16 % scale == 0 else 4 * scale
    assert max(sizes) - min(sizes) <= block
    # ownership in OUTPUT rows does not depend on the scale of a progressive sweep
    if 16 % scale == 0:
        assert owned_rows(h, scale, 0, world)[: 16] == owned_rows(h, 1, 0, world)[: 16]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, W, H, scale, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.bindings import Port, procedural_skybox
    from ray_tracing_b200 import host, scenes

    port = Port()
    objs = host.parse_scene_string(scenes.builtin_scene_text(0))
    world_desc = port.world(objs, procedural_skybox(32, seed=1))
    r0, r1 = band_rows(H, scale, rank, world)
    full = np.zeros((H, W, 3), np.float32)
    port.render(world_desc, W, H, scale, 1, 0, rows=(r0, r1), out=full, nthreads=1)
    band = torch.from_numpy(full[r0:r1].copy())
    frame = gather_bands(band, H, W, scale, rank, world, dist, dst=0)
    if rank == 0:
        np.save(os.path.join(tmp, "frame.npy"), frame.numpy())
    else:
        assert frame is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,scale", [(2, 1), (3, 4)])
def test_gather_bands_gloo(tmp_path, world, scale):
    W, H = 64, 44
    mp.spawn(_worker, args=(world, _free_port(), W, H, scale, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "frame.npy")
    from oracle.bindings import Port, procedural_skybox
    from ray_tracing_b200 import host, scenes

    port = Port()
    objs = host.parse_scene_string(scenes.builtin_scene_text(0))
    want, _ = port.render(port.world(objs, procedural_skybox(32, seed=1)), W, H, scale, 1, 0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
