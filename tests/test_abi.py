"""The C-ABI boundary: the shipped library loads, exports every symbol include/*.h declares,
keeps the reference's struct layouts, and its host-side logic (presets, validation, metadata)
agrees with the oracle.  No GPU, no compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

from flake_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from flake_b200 import build
    build.build_product()
    return api.load_library()


def declared_symbols():
    names = set()
    for h in ("flake.h", "flake_b200.h"):
        with open(os.path.join(ROOT, "include", h)) as f:
            txt = f.read()
        names |= set(re.findall(r"FLAKE_API[^;(]*?\b(flake_\w+)\s*\(", txt))
    return sorted(names)


def test_headers_declare_the_reference_api():
    want = {"flake_set_defaults", "flake_validate_params", "flake_encode_init", "flake_get_buffer",
            "flake_encode_frame", "flake_encode_close", "flake_get_version", "flake_get_streaminfo",
            "flake_write_streaminfo", "flake_init_vorbiscomment", "flake_add_vorbiscomment_entry",
            "flake_get_vorbiscomment_size", "flake_write_vorbiscomment"}
    assert want <= set(declared_symbols())


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", api.DEFAULT_LIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    for s in declared_symbols():
        getattr(lib, s)


def test_library_is_sm100a_cuda(lib):
    out = subprocess.run(["cuobjdump", "-lelf", api.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts():
    assert C.sizeof(api.FlakeEncodeParams) == 48
    assert api.FlakeContext.params.offset == 16
    assert api.FlakeContext.header.offset == 64 and api.FlakeContext.private_ctx.offset == 72
    assert C.sizeof(api.FlakeStreaminfo) == 48


def test_presets_and_validation_match_oracle(lib, oracle):
    for level in range(13):
        p = api.FlakeEncodeParams()
        p.compression = level
        assert lib.flake_set_defaults(C.byref(p)) == 0
        o = oracle.make_params(2, 44100, 16, level)
        assert (p.block_size, p.prediction_type, p.min_prediction_order, p.max_prediction_order,
                p.order_method, p.min_partition_order, p.max_partition_order, p.stereo_method,
                p.variable_block_size, p.allow_vbs, p.padding_size) == (
                o.block_size, o.prediction_type, o.min_order, o.max_order, o.order_method,
                o.min_porder, o.max_porder, o.stereo_method, o.variable_block_size, o.allow_vbs,
                o.padding_size)
    bad = api.FlakeEncodeParams()
    bad.compression = 13
    assert lib.flake_set_defaults(C.byref(bad)) == -1
    assert lib.flake_set_defaults(None) == -1


@pytest.mark.parametrize("ch,rate,bps,level,ov,want", [
    (2, 44100, 16, 5, {}, 0), (2, 44100, 16, 11, {}, 1), (2, 96000, 24, 12, {}, 0),
    (0, 44100, 16, 5, {}, -1), (9, 44100, 16, 5, {}, -1), (2, 0, 16, 5, {}, -1),
    (2, 44100, 3, 5, {}, -1), (2, 44100, 32, 5, {}, 1), (2, 44100, 16, 8, {"variable_block_size": 1}, -1),
    (2, 44100, 16, 9, {"block_size": 64}, -1), (2, 44100, 16, 5, {"block_size": 15}, -1),
    (2, 44100, 16, 5, {"min_prediction_order": 9}, -1), (2, 44100, 16, 2, {"max_prediction_order": 5}, -1),
    (2, 44100, 16, 5, {"max_partition_order": 9}, -1), (2, 44100, 16, 5, {"padding_size": 1 << 24}, -1),
])
def test_validate_params(lib, oracle, ch, rate, bps, level, ov, want):
    enc = api.Encoder(lib, ch, rate, bps, 0, level, **ov)
    assert enc.validate() == want
    o = oracle.make_params(ch, rate, bps, level, 0, **ov)
    assert oracle.lib().orc_validate(C.byref(o)) == want


def test_vorbis_comment_and_streaminfo_serialisers(lib):
    class VC(C.Structure):
        _fields_ = [("vendor_string", C.c_char_p), ("num_entries", C.c_uint), ("entries", C.c_char_p * 1024)]
    vc = VC()
    lib.flake_init_vorbiscomment(C.byref(vc))
    assert vc.vendor_string == b"Flake SVN" and vc.num_entries == 0
    assert lib.flake_get_vorbiscomment_size(C.byref(vc)) == 17
    e1 = C.create_string_buffer(b"TITLE=test tone")
    assert lib.flake_add_vorbiscomment_entry(C.byref(vc), e1) == 0
    assert lib.flake_add_vorbiscomment_entry(C.byref(vc), C.create_string_buffer(b"no equals sign")) == 1
    assert lib.flake_add_vorbiscomment_entry(C.byref(vc), C.create_string_buffer(b"B\x7fD=x")) == 1
    size = lib.flake_get_vorbiscomment_size(C.byref(vc))
    assert size == 17 + 4 + 15
    buf = (C.c_ubyte * size)()
    assert lib.flake_write_vorbiscomment(C.byref(vc), buf) == 0
    raw = bytes(buf)
    assert raw[:4] == b"\x09\0\0\0" and raw[4:13] == b"Flake SVN" and raw[13:17] == b"\x01\0\0\0"
    assert raw[17:21] == b"\x0f\0\0\0" and raw[21:] == b"TITLE=test tone"

    si = api.FlakeStreaminfo(4096, 4096, 0, 16912, 44100, 2, 16, 158760000)
    for i in range(16):
        si.md5sum[i] = i
    out = (C.c_ubyte * 34)()
    lib.flake_write_streaminfo(C.byref(si), out)
    b = bytes(out)
    assert b[:4] == b"\x10\x00\x10\x00" and b[4:7] == b"\0\0\0" and b[7:10] == (16912).to_bytes(3, "big")
    assert b[10:14] == bytes([0x0a, 0xc4, 0x42, 0xf0]) and b[14:18] == (158760000).to_bytes(4, "big")
    assert b[18:] == bytes(range(16))


def test_no_cpu_fallback(lib):
    """Without a CUDA device the encoder must refuse to start, loudly; it must never encode on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    enc = api.Encoder(lib, 2, 44100, 16, 0, 8)
    with pytest.raises(api.FlakeLibraryError):
        enc.init()
    assert not enc.ctx.private_ctx and not enc.ctx.header
    lib.flake_encode_close(C.byref(enc.ctx))          # close after a failed init is safe (flake.c:560-563)


def test_product_does_not_link_the_oracle():
    out = subprocess.run(["nm", "-D", api.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "orc_" not in out
    out = subprocess.run(["ldd", api.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emu" not in out


def test_seektable_from_frame_lengths(lib):
    """flake_b200_write_seektable: FLAC seek points (sample u64, offset u64, samples u16, big endian)
    from the per-frame lengths / block sizes of a batch call; host-only code."""
    import numpy as np
    import struct
    flen = np.array([100, 250, 90, 300, 120, 80, 70], dtype=np.uint32)
    fbs = np.array([4096, 4096, 2048, 2048, 4096, 4096, 1000], dtype=np.uint32)

    def expect(interval):
        pts, sample, off, nxt = [], 0, 0, 0
        for l, b in zip(flen, fbs):
            if sample >= nxt:
                pts.append(struct.pack(">QQH", sample, off, int(b)))
                if interval:
                    while nxt <= sample:
                        nxt += interval
                else:
                    nxt = sample + 1
            sample += int(b); off += int(l)
        return b"".join(pts)

    for interval in (0, 4096, 6000, 10000, 1 << 30):
        need = lib.flake_b200_write_seektable(flen.ctypes.data, fbs.ctypes.data, len(flen), interval, None, 0)
        want = expect(interval)
        assert need == len(want)
        buf = (C.c_ubyte * need)()
        got = lib.flake_b200_write_seektable(flen.ctypes.data, fbs.ctypes.data, len(flen), interval, buf, need)
        assert got == need and bytes(buf) == want
        assert lib.flake_b200_write_seektable(flen.ctypes.data, fbs.ctypes.data, len(flen), interval, buf, need - 1) == -1


def test_reference_callers_link_against_the_library():
    """north_star: "the flake CLI and util/api_example.c link against it unchanged".  build()
    compiles the reference's own flake/flake.c (+ libpcm_io) and util/api_example.c, unchanged,
    against include/flake.h and links them with flake_b200/lib/libflake.so, plus the CLI with
    patches/flake_cli_batch.patch applied (oracle/Makefile).  Here: they exist, every libflake
    symbol they import is exported by the library, and the dynamic loader resolves them (ldd)."""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref")
    exes = [os.path.join(ref, n) for n in ("flake_cli_b200", "flake_cli_b200_batch", "api_example_b200")]
    if not all(os.path.exists(e) for e in exes):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    exported = set(l.split()[-1] for l in subprocess.run(
        ["nm", "-D", "--defined-only", os.path.join(ROOT, "flake_b200", "lib", "libflake.so")],
        stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines() if l.strip())
    for e in exes:
        und = [l.split()[-1] for l in subprocess.run(["nm", "-D", "-u", e], stdout=subprocess.PIPE, text=True,
                                                     check=True).stdout.splitlines() if "flake_" in l]
        assert und, e
        assert set(und) <= exported, (e, sorted(set(und) - exported))
        ldd = subprocess.run(["ldd", e], stdout=subprocess.PIPE, text=True).stdout
        assert "libflake.so" in ldd and "libflake.so => not found" not in ldd, ldd
    und = subprocess.run(["nm", "-D", "-u", exes[1]], stdout=subprocess.PIPE, text=True).stdout
    assert "flake_b200_encode_stream" in und and "flake_encode_frame" not in und
