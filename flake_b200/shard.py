"""Frame-range sharding of ONE stream over several ranks.

A frame's bytes depend only on its samples, the stream parameters and its header
number (encode.c:726-764, 969-975), so rank r encodes the contiguous block range
[b0, b1) with its context seeked to the counter the serial encoder would have had
at b0, and returns (frame bytes, frame lengths, max frame size).  There is no
data-path collective: the host of rank 0 concatenates in rank order (exclusive prefix
sum over the byte lengths gives every frame's file offset), takes the maximum of the
max-frame-sizes and finishes STREAMINFO with the MD5 of the whole PCM, which it
computes itself (MD5 is a serial chain over the stream, md5.c).

`torch.distributed` is used only to ship the per-rank results to rank 0
(gather_object over gloo/NCCL's CPU side); tests run it with gloo on CPU.
"""
from __future__ import annotations

import ctypes as C
import hashlib
from typing import List, Optional, Tuple

import numpy as np

from . import api
from .synth import pack_pcm


def block_ranges(nblocks: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced block ranges; rank r gets [b0, b1)."""
    base, extra = divmod(nblocks, world)
    out, b = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((b, b + n))
        b += n
    return out


def encode_range(lib, pcm: np.ndarray, sample_rate: int, bps: int, level: int, b0: int, b1: int,
                 total_samples: int, **overrides):
    """Encode blocks [b0, b1) of the stream `pcm` (whole-stream array or this rank's view
    starting at block b0, see `local`).  Returns (bytes, frame_len, frame_bs, max_frame_size)."""
    ch = pcm.shape[1]
    enc = api.Encoder(lib, ch, sample_rate, bps, total_samples, level, **overrides)
    if enc.validate() < 0:
        raise ValueError("invalid encoding parameters")
    enc.init()
    try:
        bs = int(enc.ctx.params.block_size)
        s0, s1 = b0 * bs, min(b1 * bs, pcm.shape[0])
        if s1 <= s0:
            return b"", np.zeros(0, np.uint32), np.zeros(0, np.uint32), 0
        counter = s0 if enc.ctx.params.allow_vbs else b0
        lib.flake_b200_seek(C.byref(enc.ctx), counter & 0xFFFFFFFF)
        data, flen, fbs = enc.encode_stream(pcm[s0:s1])
        mx = int(enc.stats().max_frame_size)
        return data.tobytes(), flen.copy(), fbs.copy(), mx
    finally:
        enc.close()


def assemble(parts, header: bytes, streaminfo_fn):
    """Host side of the sharded encode: prefix-sum offsets, max of max frame sizes."""
    lens = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.uint32)
    offsets = np.concatenate([[0], np.cumsum(lens.astype(np.int64))])[:-1] + len(header)
    body = b"".join(p[0] for p in parts)
    mx = max([p[3] for p in parts] + [0])
    h = bytearray(header)
    h[8:42] = streaminfo_fn(mx)
    return bytes(h) + body, offsets, lens


def encode_sharded(lib, pcm: np.ndarray, sample_rate: int, bps: int, level: int,
                   rank: int = 0, world: int = 1, group=None, **overrides) -> Optional[bytes]:
    """Every rank calls this with the same `pcm`; rank 0 returns the complete .flac bytes."""
    n, ch = pcm.shape
    enc = api.Encoder(lib, ch, sample_rate, bps, n, level, **overrides)
    if enc.validate() < 0:
        raise ValueError("invalid encoding parameters")
    bs = int(enc.ctx.params.block_size)
    nblocks = (n + bs - 1) // bs
    b0, b1 = block_ranges(nblocks, world)[rank]
    part = encode_range(lib, pcm, sample_rate, bps, level, b0, b1, n, **overrides)
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(part, gathered, dst=0, group=group)
    else:
        gathered = [part]
    if rank != 0:
        return None
    header = enc.init()          # rank 0 also owns the stream header
    try:
        md5 = hashlib.md5(pack_pcm(pcm, bps)).digest()

        def streaminfo(max_frame):
            si = api.FlakeStreaminfo()
            lib.flake_get_streaminfo(C.byref(enc.ctx), C.byref(si))
            si.max_frame_size = max(int(si.max_frame_size), max_frame)
            C.memmove(si.md5sum, md5, 16)
            buf = (C.c_ubyte * 34)()
            lib.flake_write_streaminfo(C.byref(si), buf)
            return bytes(buf)

        data, _, _ = assemble(gathered, header, streaminfo)
        return data
    finally:
        enc.close()
