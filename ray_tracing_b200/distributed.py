"""Row-band partition of a frame across ranks (one process per GPU) and the
composite to rank 0.

The render path has no exchange step while rendering: pixels are independent
once the RNG is keyed per pixel, scene and skybox are replicated.  The only
collective is the gather of finished bands to rank 0.  `band_rows` is the same
split `render_frame_cuda_ex` uses inside one process for several GPUs
(rt_api.cu: render_pass).
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def band_rows(h: int, scale: int, rank: int, world: int, row0: int = 0, row1: "int | None" = None) -> Tuple[int, int]:
    """Contiguous band [r0, r1) of output rows for `rank`, aligned to `scale`
    (a low-res row is never split; main.c:290 iterates low-res rows).  For a
    progressive sweep pass scale=16 (the coarsest scale) for every pass so each
    rank keeps one accumulation band."""
    row1 = h if row1 is None else row1
    lrows = (row1 - row0 + scale - 1) // scale
    base, extra = divmod(lrows, world)
    start = rank * base + min(rank, extra)
    take = base + (1 if rank < extra else 0)
    r0 = row0 + start * scale
    r1 = min(row0 + (start + take) * scale, row1)
    return r0, max(r0, r1)


def all_bands(h: int, scale: int, world: int) -> List[Tuple[int, int]]:
    return [band_rows(h, scale, r, world) for r in range(world)]


def gather_bands(band, h: int, w: int, scale: int, rank: int, world: int, dist, dst: int = 0):
    """Gather per-rank band tensors (rows x w x c) into the full frame on `dst`.

    `band` is a torch tensor on the rank's device (NCCL) or on the CPU (gloo).
    Bands may differ in height by one low-res row, so they are padded to the
    tallest band for the collective and trimmed on arrival.  Returns the full
    (h, w, c) tensor on `dst`, None elsewhere.
    """
    import torch

    bands = all_bands(h, scale, world)
    tallest = max(r1 - r0 for r0, r1 in bands)
    c = band.shape[-1]
    send = band
    if band.shape[0] != tallest:
        send = torch.zeros((tallest, w, c), dtype=band.dtype, device=band.device)
        send[: band.shape[0]] = band
    send = send.contiguous()
    if world == 1:
        return send[: bands[0][1] - bands[0][0]]
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    full = torch.empty((h, w, c), dtype=band.dtype, device=band.device)
    for (r0, r1), part in zip(bands, recv):
        full[r0:r1] = part[: r1 - r0]
    return full


INTERLEAVE_ROWS = 16  # rt_api.cu: RT_INTERLEAVE_ROWS


def owned_rows(h: int, scale: int, rank: int, world: int):
    """Output rows rendered by `rank` under the round-robin block split used for
    shared frames (`interleave_count` / `interleave_index` of RtRenderOpts,
    rt_api.cu: interleave_shift / launch_band): low-res rows are dealt in blocks
    of 16/scale rows (4 for scales that do not divide 16), block b to rank
    b % world.  Rows >= (h // scale) * scale are never written (main.c:285-290)."""
    per = INTERLEAVE_ROWS // scale if 1 <= scale <= INTERLEAVE_ROWS and INTERLEAVE_ROWS % scale == 0 else 4
    shift = 0
    while (1 << shift) < per:
        shift += 1
    per = 1 << shift
    lh = h // scale
    rows = []
    for j in range(lh):
        if (j // per) % world == rank:
            rows.extend(range(j * scale, (j + 1) * scale))
    return rows
