"""Data-parallel plumbing of the multibox head: one process per GPU, the batch sharded by image.

The only exchange on the hot path is the pair of batch-global scalars of MultiBoxLoss — the max of conf
(box_utils.py:167) and the number of positives N (multibox_loss.py:117) — which the reference computes on
the batch gathered by DataParallel on GPU 0 (train_lesion_multiphase_v2.py:242-246).  Each rank's stage-1
kernel leaves them in a 16-byte header (`gssd_loss_stats`, include/gssd.h).  On one NVLink box the header travels as
peer stores from that kernel's last CTA into every rank's exchange buffer (`PeerExchange`); otherwise the headers are
all-gathered (NCCL, or gloo in the CPU tests).  Stage 2 reduces them in-kernel (MAX / SUM) either way.
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


# ---- peer exchange over NVLink peer memory (include/gssd.h: gssd_xchg) --------------------------------------
_exchanges = {}


class PeerExchange(object):
    """Every rank's 16-byte statistics slot buffer, peer-mapped into every other rank of the group (CUDA IPC): stage 1's
    last CTA stores into all peers, stage 2 spins on the local copy — no collective call, graph-capturable."""

    def __init__(self, group=None):
        """Collective over `group`.  Never raises between its collectives: a rank that fails (no IPC, ranks on several
        hosts, no peer access) records `self.error` and still takes part in every exchange, so that the others do not hang;
        `peer_exchange()` then agrees on the outcome with one all-reduce."""
        import ctypes as C
        from . import _lib
        lib = _lib.require_cuda()
        dist, ws, rank = world(group)
        self.lib, self.opened, self.own, self.x, self.error = lib, [], None, None, None
        handle = (C.c_ubyte * _lib.XCHG_HANDLE_BYTES)()
        try:
            if ws > _lib.XCHG_MAX_RANKS:
                raise RuntimeError("peer exchange supports up to %d ranks" % _lib.XCHG_MAX_RANKS)
            own = C.c_void_p()
            _lib.check(lib.gssd_xchg_create(C.byref(own), handle), "gssd_xchg_create")
            self.own = own
        except Exception as e:
            self.error = e
        mine = (bytes(handle) if self.error is None else None, torch.cuda.current_device(), _hostname())
        everyone = [None] * ws
        dist.all_gather_object(everyone, mine, group=group)
        if self.error is not None:
            return
        try:
            if any(h[0] is None for h in everyone):
                raise RuntimeError("a rank could not create its exchange buffer")
            if any(h[2] != mine[2] for h in everyone):
                raise RuntimeError("peer exchange needs all ranks on one host")
            x = _lib.Xchg()
            x.rank, x.world = rank, ws
            for r, (hb, _, _) in enumerate(everyone):
                if r == rank:
                    x.peers[r] = self.own.value
                else:
                    ptr = C.c_void_p()
                    buf = (C.c_ubyte * _lib.XCHG_HANDLE_BYTES).from_buffer_copy(hb)
                    _lib.check(lib.gssd_xchg_open(buf, C.byref(ptr)), "gssd_xchg_open")
                    self.opened.append(ptr)
                    x.peers[r] = ptr.value
            self.x = x
        except Exception as e:
            self.error = e

    def close(self):
        for ptr in self.opened:
            self.lib.gssd_xchg_close(ptr)
        self.opened = []
        if self.own is not None and self.own.value:
            self.lib.gssd_xchg_destroy(self.own)
            self.own = None


def _hostname():
    import socket
    return socket.gethostname()


def _group_key(dist, group):
    """the process-group OBJECT (kept referenced by the cache, so its id cannot be recycled): a group that is destroyed and
    created again is a different object and gets fresh IPC mappings"""
    return dist.group.WORLD if group is None else group


def exchange_timeout_ms():
    """GSSD_XCHG_TIMEOUT_S (default 10): how long a kernel waits for a peer's statistics before it traps.  A rank that never
    arrives must not wedge the GPU, but a job whose ranks can be further apart than this when they reach the criterion (a
    stalled data loader, validation on one rank) must raise it, or pass process_group=False for rank-local calls; 0 = for ever."""
    import os
    return int(float(os.environ.get("GSSD_XCHG_TIMEOUT_S", "10")) * 1000)


def peer_exchange(group=None, owner=None):
    """the PeerExchange of (`group`, `owner`) (created on first use — a collective over the group), or None when the exchange
    is off or unavailable: GSSD_PEER_XCHG=0, a non-NCCL backend, ranks on several hosts, or no peer access between the GPUs.
    Every consumer (a MultiBoxLoss module, a HostPipeline) passes itself as `owner` and gets its OWN buffer and epoch: two
    consumers that share one would interleave their steps' statistics."""
    import os
    dist, ws, _ = world(group)
    if ws <= 1 or os.environ.get("GSSD_PEER_XCHG", "1") == "0" or not torch.cuda.is_available():
        return None
    key = (_group_key(dist, group), id(owner) if owner is not None else 0)
    if key not in _exchanges:
        ex = None
        if dist.get_backend(group) == "nccl":                    # the same answer on every rank
            ex = PeerExchange(group)
            if ex.error is not None:
                import warnings
                warnings.warn("gssd: peer exchange unavailable (%s); using the NCCL all-gather" % ex.error)
            # every rank has mapped every buffer before first use, or nobody uses the exchange
            ok = torch.tensor([1 if ex.error is None else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok) == 0:
                ex.close()
                ex = None
            else:
                ex.x.timeout_ms = exchange_timeout_ms()
        _exchanges[key] = (ex, owner)                            # the owner stays referenced: its id cannot be recycled either
    return _exchanges[key][0]


def close_exchanges(group=None):
    """tear the exchanges down (all of them, or those of one group): call before destroy_process_group()"""
    dist, _, _ = world(group)
    for key in list(_exchanges):
        if group is None or (dist is not None and key[0] is _group_key(dist, group)):
            ex = _exchanges.pop(key)[0]
            if ex is not None:
                ex.close()


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
