"""Python face of flake_b200_encode_corpus (include/flake_b200.h, csrc/flake_corpus.c): many
streams -- or one long stream -- over the GPUs of one box from one process.

`Corpus` keeps the per-device engines between calls; `encode_corpus` is the one-shot form.
Buffers handed over as torch pinned tensors / numpy views of them are read and written by DMA
directly; ordinary numpy arrays go through the library's staging copy.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import api


@dataclass
class CorpusResult:
    data: np.ndarray          # the stream's frames, back to back (a view of the output buffer)
    frame_len: np.ndarray
    frame_bs: np.ndarray
    max_frame_size: int
    md5: bytes
    header: bytes             # "fLaC" + final STREAMINFO + vendor comment + padding

    def file_bytes(self) -> bytes:
        return self.header + self.data.tobytes()


def _context(lib, channels, rate, bps, level, samples=0, **overrides) -> api.FlakeContext:
    enc = api.Encoder(lib, channels, rate, bps, samples, level, **overrides)   # fills the public fields only
    if enc.validate() < 0:
        raise ValueError("invalid encoding parameters")
    return enc.ctx


class Corpus:
    def __init__(self, lib, channels: int, rate: int, bps: int, level: int, pcm_format: int,
                 devices: Optional[Sequence[int]] = None, longest: int = 0, threads_per_device: int = 0,
                 md5_threads: int = 0, chunk_blocks: int = 0, **overrides):
        self.lib = lib
        self.ctx = _context(lib, channels, rate, bps, level, longest, **overrides)
        self.fmt = pcm_format
        opt = api.FlakeB200CorpusOptions(threads_per_device, md5_threads, chunk_blocks)
        devs = (C.c_int * len(devices))(*devices) if devices else None
        self.handle = lib.flake_b200_corpus_open(C.byref(self.ctx), pcm_format, devs, len(devices) if devices else 0,
                                                 C.byref(opt))
        if not self.handle:
            raise api.FlakeLibraryError("flake_b200_corpus_open failed (no CUDA device, or bad parameters)")

    def max_encoded_size(self, nsamples: int) -> int:
        return int(self.lib.flake_b200_max_encoded_size(C.byref(self.ctx), nsamples))

    def frame_cap(self, nsamples: int) -> int:
        bs = int(self.ctx.params.block_size)
        return ((nsamples + bs - 1) // bs) * (8 if self.ctx.params.variable_block_size else 1) + 1

    def encode(self, pcms: Sequence[np.ndarray], nsamples: Optional[Sequence[int]] = None,
               outs: Optional[Sequence[np.ndarray]] = None, want_sizes: bool = True):
        """pcms[i]: interleaved samples of stream i in the corpus's container.  Returns
        (results, stats)."""
        n = len(pcms)
        items = (api.FlakeB200CorpusStream * max(n, 1))()
        keep = []
        for i, pcm in enumerate(pcms):
            ns = int(nsamples[i]) if nsamples is not None else int(pcm.shape[0])
            out = outs[i] if outs is not None else np.empty(self.max_encoded_size(ns), dtype=np.uint8)
            fcap = self.frame_cap(ns)
            flen = np.zeros(fcap, dtype=np.uint32)
            fbs = np.zeros(fcap, dtype=np.uint32)
            it = items[i]
            it.pcm = pcm.ctypes.data; it.nsamples = ns
            it.out = out.ctypes.data; it.out_cap = out.nbytes
            it.frame_len = flen.ctypes.data if want_sizes else None
            it.frame_bs = fbs.ctypes.data if want_sizes else None
            it.frame_cap = fcap
            keep.append((pcm, out, flen, fbs))
        stats = api.FlakeB200CorpusStats()
        rc = self.lib.flake_b200_corpus_encode(self.handle, items, n, C.byref(stats))
        if rc < 0:
            raise api.FlakeLibraryError("flake_b200_corpus_encode returned %d: %s" % (
                rc, stats.error.decode(errors="replace")))
        results = []
        for i in range(n):
            it = items[i]
            _, out, flen, fbs = keep[i]
            hl = self.lib.flake_b200_corpus_stream_header(C.byref(self.ctx), C.byref(it), None, 0)
            hb = (C.c_ubyte * hl)()
            self.lib.flake_b200_corpus_stream_header(C.byref(self.ctx), C.byref(it), hb, hl)
            results.append(CorpusResult(out[:it.bytes], flen[:it.nframes], fbs[:it.nframes], int(it.max_frame_size),
                                        bytes(it.md5sum), bytes(hb)))
        return results, stats

    def close(self):
        if self.handle:
            self.lib.flake_b200_corpus_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def encode_corpus(lib, pcms: Sequence[np.ndarray], channels: int, rate: int, bps: int, level: int,
                  pcm_format: int = api.PCM_S32, devices: Optional[Sequence[int]] = None, **kw):
    longest = max([int(p.shape[0]) for p in pcms] + [0])
    with Corpus(lib, channels, rate, bps, level, pcm_format, devices, longest, **kw) as co:
        return co.encode(pcms)
