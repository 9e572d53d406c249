"""ctypes mirror of libflake's C API (include/flake.h) and the batch extension
(include/flake_b200.h).

The functions keep the reference's names, argument meaning and error behaviour
(libflake/flake.h:217-295): negative ints for errors, library-owned header and
frame buffer.  The same binding class drives the compiled reference
(oracle/_ref/libflake_ref.so) in the tests, which is what makes the parity
tests read like a caller of the reference.

The product library is flake_b200/lib/libflake.so (nvcc, sm_100a).  Loading
fails loudly when it has not been built; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libflake.so")

PCM_S32, PCM_S16LE, PCM_S24LE, PCM_S8 = 0, 1, 2, 3

STAGES = ("frames", "prep", "lpc", "search", "pack")

ORDER_METHOD = {"max": 0, "est": 1, "2level": 2, "4level": 3, "8level": 4, "search": 5, "log": 6}


class FlakeEncodeParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "compression", "order_method", "stereo_method", "block_size", "padding_size",
        "min_prediction_order", "max_prediction_order", "prediction_type",
        "min_partition_order", "max_partition_order", "variable_block_size", "allow_vbs")]


class FlakeContext(C.Structure):
    _fields_ = [("channels", C.c_int), ("sample_rate", C.c_int), ("bits_per_sample", C.c_int),
                ("samples", C.c_uint), ("params", FlakeEncodeParams),
                ("header", C.POINTER(C.c_ubyte)), ("private_ctx", C.c_void_p)]


class FlakeStreaminfo(C.Structure):
    _fields_ = [("min_block_size", C.c_uint), ("max_block_size", C.c_uint),
                ("min_frame_size", C.c_uint), ("max_frame_size", C.c_uint),
                ("sample_rate", C.c_uint), ("channels", C.c_uint), ("bits_per_sample", C.c_uint),
                ("samples", C.c_uint), ("md5sum", C.c_ubyte * 16)]


class FlakeB200Stats(C.Structure):
    _fields_ = [("samples", C.c_ulonglong), ("frames", C.c_ulonglong), ("bytes", C.c_ulonglong),
                ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong),
                ("kernel_launches", C.c_ulonglong), ("max_frame_size", C.c_uint),
                ("verbatim_frames", C.c_uint), ("gpu_ms", C.c_double), ("md5_ms", C.c_double),
                ("wall_ms", C.c_double)]


MAX_DEVICES = 16


class FlakeB200CorpusStream(C.Structure):
    """One stream of a flake_b200_encode_corpus call (include/flake_b200.h)."""
    _fields_ = [("pcm", C.c_void_p), ("nsamples", C.c_ulonglong), ("out", C.c_void_p),
                ("out_cap", C.c_ulonglong), ("frame_len", C.c_void_p), ("frame_bs", C.c_void_p),
                ("frame_cap", C.c_uint), ("bytes", C.c_longlong), ("nframes", C.c_uint),
                ("max_frame_size", C.c_uint), ("min_frame_size", C.c_uint), ("verbatim_frames", C.c_uint),
                ("md5sum", C.c_ubyte * 16)]


class FlakeB200CorpusOptions(C.Structure):
    _fields_ = [("threads_per_device", C.c_int), ("md5_threads", C.c_int), ("chunk_blocks", C.c_int)]


class FlakeB200CorpusStats(C.Structure):
    _fields_ = [("wall_ms", C.c_double), ("md5_ms", C.c_double), ("streams", C.c_ulonglong),
                ("samples", C.c_ulonglong), ("bytes", C.c_ulonglong), ("chunks", C.c_ulonglong),
                ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong), ("kernel_launches", C.c_ulonglong),
                ("chunk_blocks", C.c_uint), ("devices", C.c_int), ("gpu_threads", C.c_int),
                ("md5_threads", C.c_int), ("md5_lanes", C.c_int),
                ("device_ms", C.c_double * MAX_DEVICES), ("device_samples", C.c_ulonglong * MAX_DEVICES),
                ("error", C.c_char * 256)]


class FbSub(C.Structure):
    """Per-subframe decision record (flake_b200/csrc/engine.h FbSub)."""
    _fields_ = [("type", C.c_int32), ("order", C.c_int32), ("obits", C.c_int32),
                ("wasted", C.c_int32), ("shift", C.c_int32), ("method", C.c_int32),
                ("porder", C.c_int32), ("est_order", C.c_int32), ("est_bits", C.c_uint32),
                ("first", C.c_int32), ("maxabs", C.c_uint32), ("is_const", C.c_int32),
                ("coefs", C.c_int32 * 32), ("params", C.c_uint8 * 256)]


class FlakeLibraryError(RuntimeError):
    pass


def load_library(path: Optional[str] = None, extension: bool = True) -> C.CDLL:
    """Open a libflake.  `extension=False` for a plain reference build."""
    path = path or os.environ.get("FLAKE_B200_LIB") or DEFAULT_LIB
    if not os.path.exists(path):
        raise FlakeLibraryError(
            "%s not found: build it with `python -m flake_b200.build product` "
            "(flake_b200 has no CPU fallback)" % path)
    lib = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
    P = C.POINTER
    lib.flake_set_defaults.argtypes = [P(FlakeEncodeParams)]; lib.flake_set_defaults.restype = C.c_int
    lib.flake_validate_params.argtypes = [P(FlakeContext)]; lib.flake_validate_params.restype = C.c_int
    lib.flake_encode_init.argtypes = [P(FlakeContext)]; lib.flake_encode_init.restype = C.c_int
    lib.flake_get_buffer.argtypes = [P(FlakeContext)]; lib.flake_get_buffer.restype = C.c_void_p
    lib.flake_encode_frame.argtypes = [P(FlakeContext), C.c_void_p, C.c_int]; lib.flake_encode_frame.restype = C.c_int
    lib.flake_encode_close.argtypes = [P(FlakeContext)]; lib.flake_encode_close.restype = None
    lib.flake_get_version.argtypes = []; lib.flake_get_version.restype = C.c_char_p
    lib.flake_get_streaminfo.argtypes = [P(FlakeContext), P(FlakeStreaminfo)]; lib.flake_get_streaminfo.restype = C.c_int
    lib.flake_write_streaminfo.argtypes = [P(FlakeStreaminfo), C.c_void_p]; lib.flake_write_streaminfo.restype = None
    if extension:
        lib.flake_b200_set_device.argtypes = [C.c_int]; lib.flake_b200_set_device.restype = C.c_int
        lib.flake_b200_set_chunk_blocks.argtypes = [P(FlakeContext), C.c_int]; lib.flake_b200_set_chunk_blocks.restype = C.c_int
        lib.flake_b200_encode_stream.argtypes = [P(FlakeContext), C.c_void_p, C.c_int, C.c_ulonglong,
                                                 C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_void_p,
                                                 C.c_uint, P(C.c_uint)]
        lib.flake_b200_encode_stream.restype = C.c_longlong
        lib.flake_b200_max_encoded_size.argtypes = [P(FlakeContext), C.c_ulonglong]
        lib.flake_b200_max_encoded_size.restype = C.c_ulonglong
        lib.flake_b200_seek.argtypes = [P(FlakeContext), C.c_uint]; lib.flake_b200_seek.restype = C.c_int
        lib.flake_b200_reset_stream.argtypes = [P(FlakeContext)]; lib.flake_b200_reset_stream.restype = C.c_int
        lib.flake_b200_tell.argtypes = [P(FlakeContext)]; lib.flake_b200_tell.restype = C.c_uint
        lib.flake_b200_encode_device.argtypes = [P(FlakeContext), C.c_void_p, C.c_int, C.c_ulonglong, C.c_uint,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.flake_b200_encode_device.restype = C.c_int
        lib.flake_b200_device_capacity.argtypes = [P(FlakeContext), P(C.c_ulonglong), P(C.c_ulonglong), P(C.c_uint)]
        lib.flake_b200_device_capacity.restype = C.c_int
        lib.flake_b200_last_subframes.argtypes = [P(FlakeContext), C.c_void_p, C.c_uint]
        lib.flake_b200_last_subframes.restype = C.c_int
        lib.flake_b200_subframe_record_size.argtypes = []; lib.flake_b200_subframe_record_size.restype = C.c_uint
        lib.flake_b200_set_profiling.argtypes = [P(FlakeContext), C.c_int]; lib.flake_b200_set_profiling.restype = C.c_int
        lib.flake_b200_stage_times.argtypes = [P(FlakeContext), P(C.c_double), P(C.c_ulonglong)]
        lib.flake_b200_stage_times.restype = C.c_int
        lib.flake_b200_write_seektable.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_ulonglong]
        lib.flake_b200_write_seektable.restype = C.c_longlong
        lib.flake_b200_set_streaminfo_sizes.argtypes = [P(FlakeContext), C.c_int]; lib.flake_b200_set_streaminfo_sizes.restype = C.c_int
        lib.flake_b200_get_stats.argtypes = [P(FlakeContext), P(FlakeB200Stats)]; lib.flake_b200_get_stats.restype = C.c_int
        lib.flake_b200_last_error.argtypes = [P(FlakeContext)]; lib.flake_b200_last_error.restype = C.c_char_p
        lib.flake_b200_version.argtypes = []; lib.flake_b200_version.restype = C.c_char_p
        lib.flake_b200_encode_corpus.argtypes = [P(FlakeContext), C.c_int, P(FlakeB200CorpusStream), C.c_uint,
                                                 P(C.c_int), C.c_int, P(FlakeB200CorpusOptions), P(FlakeB200CorpusStats)]
        lib.flake_b200_encode_corpus.restype = C.c_int
        lib.flake_b200_corpus_open.argtypes = [P(FlakeContext), C.c_int, P(C.c_int), C.c_int, P(FlakeB200CorpusOptions)]
        lib.flake_b200_corpus_open.restype = C.c_void_p
        lib.flake_b200_corpus_encode.argtypes = [C.c_void_p, P(FlakeB200CorpusStream), C.c_uint, P(FlakeB200CorpusStats)]
        lib.flake_b200_corpus_encode.restype = C.c_int
        lib.flake_b200_corpus_close.argtypes = [C.c_void_p]; lib.flake_b200_corpus_close.restype = None
        lib.flake_b200_corpus_error.argtypes = [C.c_void_p]; lib.flake_b200_corpus_error.restype = C.c_char_p
        lib.flake_b200_corpus_stream_header.argtypes = [P(FlakeContext), P(FlakeB200CorpusStream), C.c_void_p, C.c_uint]
        lib.flake_b200_corpus_stream_header.restype = C.c_int
        if lib.flake_b200_subframe_record_size() != C.sizeof(FbSub):
            raise FlakeLibraryError("FbSub layout mismatch between api.py and the library")
    return lib


@dataclass
class EncodedStream:
    header: bytes                 # bytes flake_encode_init produced
    frames: List[bytes]           # one entry per flake_encode_frame call / per frame (batch)
    streaminfo: bytes             # final 34-byte STREAMINFO body (flake_get_streaminfo)
    frame_bs: Optional[np.ndarray] = None

    @property
    def payload(self) -> bytes:
        return b"".join(self.frames)

    def file_bytes(self) -> bytes:
        """What flake/flake.c leaves on disk: header with STREAMINFO rewritten at offset 8."""
        h = bytearray(self.header)
        h[8:8 + 34] = self.streaminfo
        return bytes(h) + self.payload


class Encoder:
    """One FlakeContext.  Mirrors flake/flake.c's use of the API."""

    def __init__(self, lib: C.CDLL, channels: int, sample_rate: int, bits_per_sample: int,
                 samples: int = 0, compression: int = 5, **overrides):
        self.lib = lib
        self.ctx = FlakeContext()
        self.ctx.channels = channels
        self.ctx.sample_rate = sample_rate
        self.ctx.bits_per_sample = bits_per_sample
        self.ctx.samples = samples & 0xFFFFFFFF
        self.ctx.params.compression = compression
        if lib.flake_set_defaults(C.byref(self.ctx.params)):
            raise ValueError("invalid compression level %r" % (compression,))
        for k, v in overrides.items():
            if v is None:
                continue
            if not hasattr(self.ctx.params, k):
                raise AttributeError(k)
            setattr(self.ctx.params, k, int(v))
        self.header_len = -1
        self.open = False

    def validate(self) -> int:
        return self.lib.flake_validate_params(C.byref(self.ctx))

    def init(self) -> bytes:
        n = self.lib.flake_encode_init(C.byref(self.ctx))
        self.header_len = n
        if n < 0:
            self.lib.flake_encode_close(C.byref(self.ctx))
            raise FlakeLibraryError("flake_encode_init failed")
        self.open = True
        return C.string_at(self.ctx.header, n)

    def encode_frame(self, block: np.ndarray) -> bytes:
        """block: (n, channels) int32, C-contiguous."""
        blk = np.ascontiguousarray(block, dtype=np.int32)
        n = blk.shape[0]
        fs = self.lib.flake_encode_frame(C.byref(self.ctx), blk.ctypes.data, n)
        if fs < 0:
            raise FlakeLibraryError("flake_encode_frame returned %d" % fs)
        return C.string_at(self.lib.flake_get_buffer(C.byref(self.ctx)), fs)

    def streaminfo(self) -> Tuple[FlakeStreaminfo, bytes]:
        si = FlakeStreaminfo()
        if self.lib.flake_get_streaminfo(C.byref(self.ctx), C.byref(si)):
            raise FlakeLibraryError("flake_get_streaminfo failed")
        buf = (C.c_ubyte * 34)()
        self.lib.flake_write_streaminfo(C.byref(si), buf)
        return si, bytes(buf)

    # ---- flake_b200 extension ------------------------------------------
    def encode_stream(self, pcm: np.ndarray, pcm_format: int = PCM_S32,
                      nsamples: Optional[int] = None, want_sizes: bool = True):
        """Batch encode from host memory.  Returns (bytes, frame_len, frame_bs)."""
        if pcm_format == PCM_S32:
            arr = np.ascontiguousarray(pcm, dtype=np.int32)      # the C side reads nsamples * channels int32
        else:
            arr = np.ascontiguousarray(pcm)
        if nsamples is None:
            nsamples = arr.shape[0]
        need = int(nsamples) * int(self.ctx.channels) * {PCM_S32: 4, PCM_S16LE: 2, PCM_S24LE: 3, PCM_S8: 1}[pcm_format]
        if arr.nbytes < need:
            raise ValueError("pcm holds %d bytes, %d samples x %d channels need %d" % (
                arr.nbytes, nsamples, self.ctx.channels, need))
        cap = int(self.lib.flake_b200_max_encoded_size(C.byref(self.ctx), nsamples))
        out = np.empty(cap, dtype=np.uint8)
        bs = int(self.ctx.params.block_size)
        fcap = ((nsamples + bs - 1) // bs) * (8 if self.ctx.params.variable_block_size else 1) + 1
        flen = np.zeros(fcap, dtype=np.uint32)
        fbs = np.zeros(fcap, dtype=np.uint32)
        nf = C.c_uint(0)
        rc = self.lib.flake_b200_encode_stream(
            C.byref(self.ctx), arr.ctypes.data, pcm_format, nsamples, out.ctypes.data, cap,
            flen.ctypes.data if want_sizes else None, fbs.ctypes.data if want_sizes else None,
            fcap, C.byref(nf))
        if rc < 0:
            raise FlakeLibraryError("flake_b200_encode_stream returned %d: %s" % (
                rc, self.lib.flake_b200_last_error(C.byref(self.ctx)).decode()))
        return out[:rc], flen[:nf.value], fbs[:nf.value]

    def last_subframes(self, max_records: int = 1 << 16) -> List[FbSub]:
        arr = (FbSub * max_records)()
        n = self.lib.flake_b200_last_subframes(C.byref(self.ctx), arr, max_records)
        if n < 0:
            raise FlakeLibraryError("flake_b200_last_subframes failed")
        return [arr[i] for i in range(n)]

    def set_profiling(self, on: bool) -> None:
        if self.lib.flake_b200_set_profiling(C.byref(self.ctx), 1 if on else 0):
            raise FlakeLibraryError("flake_b200_set_profiling failed")

    def stage_times(self):
        """{stage: (cumulative ms, passes)} from CUDA events between the kernels."""
        ms = (C.c_double * 7)()
        cnt = (C.c_ulonglong * 7)()
        n = self.lib.flake_b200_stage_times(C.byref(self.ctx), ms, cnt)
        if n < 0:
            raise FlakeLibraryError("flake_b200_stage_times failed")
        return {STAGES[i]: (ms[i], cnt[i]) for i in range(n)}

    def stats(self) -> FlakeB200Stats:
        st = FlakeB200Stats()
        self.lib.flake_b200_get_stats(C.byref(self.ctx), C.byref(st))
        return st

    def close(self):
        if self.open:
            self.lib.flake_encode_close(C.byref(self.ctx))
            self.open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def encode_per_block(lib: C.CDLL, pcm: np.ndarray, sample_rate: int, bits_per_sample: int,
                     compression: int = 5, **overrides) -> EncodedStream:
    """The flake/flake.c:612-685 loop: one flake_encode_frame call per block."""
    n, ch = pcm.shape
    enc = Encoder(lib, ch, sample_rate, bits_per_sample, n, compression, **overrides)
    if enc.validate() < 0:
        raise ValueError("invalid encoding parameters")
    header = enc.init()
    try:
        bs = int(enc.ctx.params.block_size)
        frames = [enc.encode_frame(pcm[i:i + bs]) for i in range(0, n, bs)]
        _, si = enc.streaminfo()
    finally:
        enc.close()
    return EncodedStream(header, frames, si)


def encode_batch(lib: C.CDLL, pcm: np.ndarray, sample_rate: int, bits_per_sample: int,
                 compression: int = 5, pcm_format: int = PCM_S32, nsamples: Optional[int] = None,
                 channels: Optional[int] = None, chunk_blocks: Optional[int] = None,
                 **overrides) -> EncodedStream:
    """Whole stream through flake_b200_encode_stream (one call)."""
    if pcm_format == PCM_S32:
        n, ch = pcm.shape
    else:
        n, ch = nsamples, channels
    enc = Encoder(lib, ch, sample_rate, bits_per_sample, n, compression, **overrides)
    if enc.validate() < 0:
        raise ValueError("invalid encoding parameters")
    header = enc.init()
    try:
        if chunk_blocks:
            lib.flake_b200_set_chunk_blocks(C.byref(enc.ctx), chunk_blocks)
        data, flen, fbs = enc.encode_stream(pcm, pcm_format, n)
        _, si = enc.streaminfo()
    finally:
        enc.close()
    offs = np.concatenate([[0], np.cumsum(flen)]).astype(np.int64)
    raw = data.tobytes()
    frames = [raw[offs[i]:offs[i + 1]] for i in range(len(flen))]
    return EncodedStream(header, frames, si, fbs)
