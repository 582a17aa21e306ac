"""Data-parallel plumbing of the multibox head: one process per GPU, the batch sharded by image.

The only exchange on the hot path is the pair of batch-global scalars of MultiBoxLoss — the max of conf
(box_utils.py:167) and the number of positives N (multibox_loss.py:117) — which the reference computes on
the batch gathered by DataParallel on GPU 0 (train_lesion_multiphase_v2.py:242-246).  Each rank's stage-1
kernel leaves them in a 16-byte header (`gssd_loss_stats`, include/gssd.h); the headers are all-gathered
(NCCL over NVLink on GPUs, gloo in the CPU tests) and stage 2 reduces them in-kernel (MAX / SUM).
Detect needs no exchange: its outputs are per image.
"""
import numpy as np
import torch

HEADER_BYTES = 16


def world(group=None):
    """-> (torch.distributed or None, world_size, rank) for `group` (None = default group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and group is not False:
        return dist, dist.get_world_size(group), dist.get_rank(group)
    return None, 1, 0


def shard(batch, rank, world_size):
    """contiguous image shard of a global batch: slice(start, stop) for `rank`."""
    per = (batch + world_size - 1) // world_size
    return slice(min(batch, rank * per), min(batch, (rank + 1) * per))


def all_gather_headers(header, group=None):
    """header: uint8[16] on any device -> uint8[world*16], rank-major, same device."""
    dist, ws, _ = world(group)
    if ws == 1:
        return header.clone()
    out = torch.empty((ws * HEADER_BYTES,), dtype=torch.uint8, device=header.device)
    dist.all_gather_into_tensor(out, header.contiguous(), group=group)
    return out


# ---- host mirrors of the device encodings (used by the tests and by tools) -----------------------------
def f2ord(x):
    """float32 -> order-preserving uint32 (csrc/common.cuh: f2ord)."""
    u = np.asarray(x, np.float32).view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def ord2f(o):
    o = np.asarray(o, np.uint32)
    return np.where(o & np.uint32(0x80000000), o & np.uint32(0x7fffffff), ~o).astype(np.uint32).view(np.float32)


def make_header(conf_max, num_pos):
    h = np.zeros(4, np.uint32)
    h[0] = f2ord(np.float32(conf_max))
    h[1] = np.uint32(num_pos)
    return torch.from_numpy(h.view(np.uint8).copy())


def combine_headers(headers):
    """uint8[world*16] -> (x_max float32, N int): what stage 2 computes in-kernel (csrc/loss.cu)."""
    w = headers.detach().cpu().numpy().view(np.uint32).reshape(-1, 4)
    return float(ord2f(w[:, 0].max())), int(w[:, 1].astype(np.int64).sum())
