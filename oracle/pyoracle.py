"""ctypes bindings of the TEST-ONLY oracle (oracle/libflake_oracle.so) and of the
compiled reference (oracle/_ref/libflake_ref.so, when it was built).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module (see oracle/flake_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libflake_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libflake_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "flake_ref")


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "channels", "sample_rate", "bps", "block_size", "order_method", "stereo_method",
        "prediction_type", "min_order", "max_order", "min_porder", "max_porder",
        "variable_block_size", "allow_vbs", "padding_size")] + [("total_samples", C.c_uint32)]


class OrcSubframe(C.Structure):
    _fields_ = [("type", C.c_int), ("order", C.c_int), ("obits", C.c_int), ("wasted", C.c_int),
                ("shift", C.c_int), ("coefs", C.c_int32 * 32), ("method", C.c_int),
                ("porder", C.c_int), ("params", C.c_int * 256), ("est_bits", C.c_uint32)]


class OrcFrameInfo(C.Structure):
    _fields_ = [("blocksize", C.c_int), ("ch_mode", C.c_int), ("verbatim_fallback", C.c_int),
                ("nbytes", C.c_int), ("sub", OrcSubframe * 8)]


class OrcDecInfo(C.Structure):
    _fields_ = [("channels", C.c_int), ("bps", C.c_int), ("sample_rate", C.c_int),
                ("total_samples", C.c_uint64), ("decoded_samples", C.c_uint64),
                ("nframes", C.c_uint32), ("min_bs", C.c_uint32), ("max_bs", C.c_uint32),
                ("max_frame_bytes", C.c_uint32), ("md5_ok", C.c_int), ("error", C.c_int),
                ("error_pos", C.c_uint64)]


_lib = None


def build_if_needed():
    srcs = [os.path.join(HERE, f) for f in ("flake_oracle.c", "flac_decode.c", "flake_oracle.h")]
    if (not os.path.exists(ORACLE_SO)
            or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_if_needed()
        L = C.CDLL(ORACLE_SO)
        P = C.POINTER
        L.orc_set_defaults.argtypes = [P(OrcParams), C.c_int]
        L.orc_validate.argtypes = [P(OrcParams)]
        L.orc_encode_frame.argtypes = [P(OrcParams), C.c_void_p, C.c_int, C.c_uint32, C.c_void_p,
                                       C.c_int, P(OrcFrameInfo)]
        L.orc_vbs_split.argtypes = [C.c_void_p, C.c_int, C.c_int, P(C.c_int * 8)]
        L.orc_encode_stream.argtypes = [P(OrcParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_void_p, C.c_uint32, P(C.c_uint32), P(C.c_uint32)]
        L.orc_encode_stream.restype = C.c_int64
        L.orc_write_header.argtypes = [P(OrcParams), C.c_void_p, C.c_int]
        L.orc_streaminfo.argtypes = [P(OrcParams), C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_streaminfo.restype = None
        L.orc_initial_max_frame_size.argtypes = [P(OrcParams)]
        L.orc_initial_max_frame_size.restype = C.c_uint32
        L.orc_md5_pcm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.orc_md5_pcm.restype = None
        L.orc_md5_zero_ctx.argtypes = [C.c_void_p]; L.orc_md5_zero_ctx.restype = None
        L.orc_autocorr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]; L.orc_autocorr.restype = None
        L.orc_lpc_calc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_rice_k.argtypes = [C.c_uint64, C.c_int]
        L.orc_rice_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    P(C.c_int), P(C.c_int), C.c_void_p]
        L.orc_rice_cost.restype = C.c_uint32
        L.orc_crc8.argtypes = [C.c_void_p, C.c_size_t]; L.orc_crc8.restype = C.c_uint8
        L.orc_crc16.argtypes = [C.c_void_p, C.c_size_t]; L.orc_crc16.restype = C.c_uint16
        L.orc_flac_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_uint64, P(OrcDecInfo)]
        L.orc_flac_decode.restype = C.c_int64
        _lib = L
    return _lib


def make_params(channels: int, sample_rate: int, bps: int, level: int, total_samples: int = 0,
                **overrides) -> OrcParams:
    """overrides use the FlakeEncodeParams field names (flake.h:59-161)."""
    p = OrcParams()
    if lib().orc_set_defaults(C.byref(p), level):
        raise ValueError("bad level")
    p.channels, p.sample_rate, p.bps = channels, sample_rate, bps
    p.total_samples = total_samples & 0xFFFFFFFF
    names = {"min_prediction_order": "min_order", "max_prediction_order": "max_order",
             "min_partition_order": "min_porder", "max_partition_order": "max_porder"}
    for k, v in overrides.items():
        if v is None:
            continue
        setattr(p, names.get(k, k), int(v))
    return p


def encode_stream(pcm: np.ndarray, sample_rate: int, bps: int, level: int, **overrides):
    """Returns (frame_bytes, frame_len, frame_bs, max_frame_size) from the restatement."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    n, ch = pcm.shape
    p = make_params(ch, sample_rate, bps, level, n, **overrides)
    cap = n * ch * 4 + (n // 16 + 16) * 64 + 4096
    out = np.empty(cap, dtype=np.uint8)
    fcap = (n // max(16, p.block_size) + 2) * 8
    flen = np.zeros(fcap, dtype=np.uint32)
    fbs = np.zeros(fcap, dtype=np.uint32)
    nf, mx = C.c_uint32(0), C.c_uint32(0)
    rc = lib().orc_encode_stream(C.byref(p), pcm.ctypes.data, n, out.ctypes.data, cap,
                                 flen.ctypes.data, fbs.ctypes.data, fcap, C.byref(nf), C.byref(mx))
    if rc < 0:
        raise RuntimeError("oracle encode failed")
    return out[:rc].tobytes(), flen[:nf.value].copy(), fbs[:nf.value].copy(), mx.value


def header(p: OrcParams) -> bytes:
    buf = (C.c_ubyte * (p.padding_size + 1024))()
    n = lib().orc_write_header(C.byref(p), buf, len(buf))
    return bytes(buf[:n])


def streaminfo(p: OrcParams, max_frame_size: int, md5: bytes) -> bytes:
    out = (C.c_ubyte * 34)()
    lib().orc_streaminfo(C.byref(p), max_frame_size, md5, out)
    return bytes(out)


def md5_pcm(pcm: np.ndarray, bps: int) -> bytes:
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    out = (C.c_ubyte * 16)()
    lib().orc_md5_pcm(pcm.ctypes.data, pcm.shape[1], bps, pcm.shape[0], out)
    return bytes(out)


def decode(data: bytes, has_header: bool = True, channels: int = 0, bps: int = 0,
           max_samples: Optional[int] = None):
    """Test decoder.  Returns (pcm (n, ch) int32, OrcDecInfo); raises on a bad stream."""
    info = OrcDecInfo()
    info.channels, info.bps = channels, bps
    buf = np.frombuffer(data, dtype=np.uint8)
    if max_samples is None:
        max_samples = max(65536, len(data) * 8)
    ch_guess = channels or 8
    pcm = np.zeros((max_samples, ch_guess), dtype=np.int32)
    if has_header and len(data) >= 22:
        ch_guess = ((data[20] >> 1) & 7) + 1
        pcm = np.zeros((max_samples, ch_guess), dtype=np.int32)
    n = lib().orc_flac_decode(buf.ctypes.data, len(data), 1 if has_header else 0,
                              pcm.ctypes.data, max_samples, C.byref(info))
    if n < 0:
        raise ValueError("FLAC decode error %d at byte %d (frame %d)" % (
            info.error, info.error_pos, info.nframes))
    return pcm[:n], info


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_library():
    """The reference libflake compiled from /root/reference (oracle/Makefile `ref`)."""
    from flake_b200.api import load_library
    if not have_ref():
        raise FileNotFoundError(REF_SO)
    return load_library(REF_SO, extension=False)


def ref_encode_parallel(pcm: np.ndarray, sample_rate: int, bps: int, level: int, threads: int = 0,
                        **overrides):
    """The compiled reference over a whole stream, in C (oracle/ref_shim.c): `threads`
    contiguous block ranges, one reference context per range with its frame counter set
    to the serial encoder's value there, frames concatenated in order -- i.e. exactly the
    bytes the flake/flake.c loop writes after the header.  threads=0: all host cores.
    Returns (frame_bytes (np.uint8), bytes_per_block (np.uint32), max_frame_size)."""
    from flake_b200.api import FlakeEncodeParams
    ref = ref_library()
    fn = ref.refshim_encode_parallel
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(FlakeEncodeParams), C.c_void_p, C.c_uint64, C.c_int,
                   C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    n, ch = pcm.shape
    prm = FlakeEncodeParams()
    prm.compression = level
    if ref.flake_set_defaults(C.byref(prm)):
        raise ValueError("bad level")
    for k, v in overrides.items():
        if v is not None:
            setattr(prm, k, int(v))
    bs = int(prm.block_size)
    nblocks = (n + bs - 1) // bs
    cap = n * ch * ((bps + 7) // 8) + (nblocks + 1) * (96 * (8 if prm.variable_block_size else 1) + 64) + n // 8 + 4096
    out = np.empty(cap, dtype=np.uint8)
    clen = np.zeros(nblocks + 1, dtype=np.uint32)
    nc, mx = C.c_uint32(0), C.c_uint32(0)
    if threads <= 0:
        threads = os.cpu_count() or 1
    rc = fn(ch, sample_rate, bps, C.byref(prm), pcm.ctypes.data, n, threads, out.ctypes.data, cap,
            clen.ctypes.data, nblocks + 1, C.byref(nc), C.byref(mx))
    if rc < 0:
        raise RuntimeError("reference encode failed")
    return out[:rc], clen[:nc.value], mx.value
