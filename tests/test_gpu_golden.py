"""CUDA path against the committed golden vectors (outputs of the compiled reference) and,
at BASELINE.json sizes, through size-independent properties: encode -> decode round trip with
every CRC-8 / CRC-16 and the STREAMINFO MD5 verified, frame accounting, per-block API ==
batch API == frame-range-sharded encode."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from flake_b200 import api, shard, synth
from test_oracle_golden import STREAMS, golden_pcm, GOLD

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rec", STREAMS, ids=[r["name"] for r in STREAMS])
def test_gpu_reproduces_reference_files(rec, gpu_lib):
    pcm = golden_pcm(rec)
    got = api.encode_batch(gpu_lib, pcm, rec["rate"], rec["bps"], rec["level"], chunk_blocks=3,
                           **rec["overrides"])
    data = got.file_bytes()
    assert len(got.frames) == rec["nframes"]
    assert got.streaminfo.hex() == rec["streaminfo"]
    assert len(data) == rec["flac_len"]
    assert hashlib.sha256(data).hexdigest() == rec["flac_sha256"]
    if "file" in rec:
        with open(os.path.join(GOLD, rec["file"]), "rb") as f:
            assert f.read() == data


def _long_pcm(nsamples, channels, bps, rate, seed):
    return synth.long_pcm(nsamples, channels, bps, rate, seed=seed)


def compare_with_reference(got, pcm, rate, bps, level, oracle, **ov):
    """Byte-compare the WHOLE stream with the compiled reference (oracle/_ref, driven from C on
    all host cores by oracle/ref_shim.c).  Returns (frames compared, frames mismatching,
    size delta in percent)."""
    want, per_block, mx = oracle.ref_encode_parallel(pcm, rate, bps, level, **ov)
    mine = np.frombuffer(got.payload, dtype=np.uint8)
    flen = np.array(list(map(len, got.frames)), dtype=np.int64)
    nframes = len(flen)
    size_delta = 100.0 * (len(mine) - len(want)) / max(1, len(want))
    if len(mine) == len(want) and np.array_equal(mine, want):
        return nframes, 0, size_delta, mx
    # count per block: a block is one flake_encode_frame call (1..8 frames under VBS)
    bs_cum = np.cumsum(got.frame_bs)
    block = int(bs_cum[0]) if nframes else 1
    block = max(block, int(np.max(got.frame_bs)))
    mine_off = np.concatenate([[0], np.cumsum(flen)])
    ref_off = np.concatenate([[0], np.cumsum(per_block.astype(np.int64))])
    bad = 0
    first_of_block = np.searchsorted(bs_cum - got.frame_bs, np.arange(len(per_block)) * block)
    for b in range(len(per_block)):
        f0 = first_of_block[b]
        f1 = first_of_block[b + 1] if b + 1 < len(per_block) else nframes
        x = mine[mine_off[f0]:mine_off[f1]]
        y = want[ref_off[b]:ref_off[b + 1]]
        if len(x) != len(y) or not np.array_equal(x, y):
            bad += f1 - f0
    return nframes, int(bad), size_delta, mx


@pytest.mark.parametrize("name,nsamples,ch,bps,rate,level", [
    ("C1_10min_l5", 26_460_000, 2, 16, 44100, 5),
    ("C2_1h_l8", 158_760_000, 2, 16, 44100, 8),
    ("C3_10min_l12_s24_96k", 57_600_000, 2, 24, 96000, 12),
    ("C4_10min_l9_8ch_s24", 28_800_000, 8, 24, 48000, 9),
])
def test_full_size_parity_with_reference(name, nsamples, ch, bps, rate, level, gpu_lib, oracle):
    """BASELINE.json's configurations at their full sizes: every frame of the stream is
    byte-compared with the compiled reference, the stream is decoded back (all CRC-8/CRC-16
    and the STREAMINFO MD5 verified) and the frame accounting is checked."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    pcm = _long_pcm(nsamples, ch, bps, rate, seed=level)
    got = api.encode_batch(gpu_lib, pcm, rate, bps, level)
    compared, bad, delta, ref_max = compare_with_reference(got, pcm, rate, bps, level, oracle)
    print("\n%s: frames compared %d, mismatching %d, size delta %.4f %%, %d bytes"
          % (name, compared, bad, delta, len(got.payload)))
    assert bad == 0 and delta == 0.0
    bs = 8192 if level >= 11 else 4096
    assert compared >= (nsamples + bs - 1) // bs
    assert int(np.sum(got.frame_bs)) == nsamples
    # STREAMINFO: running maximum frame size as the reference reports it, MD5 of the raw PCM
    assert int.from_bytes(got.streaminfo[7:10], "big") == ref_max
    assert got.streaminfo[18:] == hashlib.md5(synth.pack_pcm(pcm, bps)).digest()
    dec, info = oracle.decode(got.file_bytes(), max_samples=nsamples + 16)
    assert info.md5_ok == 1, "STREAMINFO MD5 does not match the decoded PCM"
    assert info.decoded_samples == nsamples and info.total_samples == nsamples
    assert np.array_equal(dec, pcm)


def test_c2_full_hour_linearity_of_sharding(gpu_lib):
    """Frame-range independence at the full C2 length: encoding blocks [b0, b1) with a seeked
    context equals the corresponding byte range of the whole-stream encode."""
    n = 158_760_000
    pcm = _long_pcm(n, 2, 16, 44100, seed=8)
    whole = api.encode_batch(gpu_lib, pcm, 44100, 16, 8)
    assert len(whole.frames) == 38760
    lens = np.array(list(map(len, whole.frames)), dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    payload = whole.payload
    for (b0, b1) in [(0, 100), (19380, 19480), (38700, 38760)]:
        part = shard.encode_range(gpu_lib, pcm, 44100, 16, 8, b0, b1, n)
        assert part[0] == payload[offs[b0]:offs[b1]]
    # STREAMINFO: total samples and the MD5 of the raw little-endian PCM
    assert whole.streaminfo[18:] == hashlib.md5(synth.pack_pcm(pcm, 16)).digest()


def test_sharded_equals_single(gpu_lib, oracle):
    pcm = synth.synth_pcm(4096 * 37 + 1200, 2, 16, 44100, seed=77, kind="impulses")
    for level in (8, 9):
        single = api.encode_batch(gpu_lib, pcm, 44100, 16, level).file_bytes()
        n = pcm.shape[0]
        nblocks = (n + 4095) // 4096
        enc = api.Encoder(gpu_lib, 2, 44100, 16, n, level)
        header = enc.init()
        parts = [shard.encode_range(gpu_lib, pcm, 44100, 16, level, b0, b1, n)
                 for (b0, b1) in shard.block_ranges(nblocks, 8)]
        import ctypes as C
        md5 = hashlib.md5(synth.pack_pcm(pcm, 16)).digest()

        def si_fn(mx):
            si = api.FlakeStreaminfo()
            gpu_lib.flake_get_streaminfo(C.byref(enc.ctx), C.byref(si))
            si.max_frame_size = max(int(si.max_frame_size), mx)
            C.memmove(si.md5sum, md5, 16)
            buf = (C.c_ubyte * 34)()
            gpu_lib.flake_write_streaminfo(C.byref(si), buf)
            return bytes(buf)
        data, offsets, lens = shard.assemble(parts, header, si_fn)
        enc.close()
        assert data == single


def test_unmodified_reference_cli_links_and_matches(gpu_lib, oracle, tmp_path):
    """oracle/_ref/flake_cli_b200 = the reference's flake/flake.c + libpcm_io, UNCHANGED, linked
    against flake_b200's libflake.so.  Its output must equal the reference CLI's."""
    import subprocess
    from oracle import pyoracle
    cli = os.path.join(os.path.dirname(pyoracle.REF_SO), "flake_cli_b200")
    if not os.path.exists(cli) or not os.path.exists(pyoracle.REF_CLI):
        pytest.skip("oracle/_ref CLIs not built (needs /root/reference at build time)")
    pcm = synth.synth_pcm(4096 * 5 + 3136, 2, 16, 44100, seed=4)
    wav = tmp_path / "in.wav"
    wav.write_bytes(synth.wav_bytes(pcm, 16, 44100))
    for level in ("-5", "-8", "-9"):
        a, b = tmp_path / "a.flac", tmp_path / "b.flac"
        subprocess.run([cli, "-q", level, str(wav), "-o", str(a)], check=True)
        subprocess.run([pyoracle.REF_CLI, "-q", level, str(wav), "-o", str(b)], check=True)
        assert a.read_bytes() == b.read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("bps", [16, 24])
def test_patched_cli_batch_path_matches_reference_cli(gpu_lib, oracle, tmp_path, bps):
    """oracle/_ref/flake_cli_b200_batch = flake/flake.c with patches/flake_cli_batch.patch (the loop
    over flake_encode_frame replaced by flake_b200_encode_stream on 2048-block batches), linked
    against libflake.so: byte-identical files to the reference CLI, every level, s16 and s24 WAV,
    batches that end inside the stream and a short last block."""
    import subprocess
    from oracle import pyoracle
    cli = os.path.join(os.path.dirname(pyoracle.REF_SO), "flake_cli_b200_batch")
    if not os.path.exists(cli) or not os.path.exists(pyoracle.REF_CLI):
        pytest.skip("oracle/_ref CLIs not built (needs /root/reference at build time)")
    rate = 44100 if bps == 16 else 96000
    pcm = synth.synth_pcm(4096 * 9 + 1777, 2, bps, rate, seed=40 + bps)
    wav = tmp_path / "in.wav"
    wav.write_bytes(synth.wav_bytes(pcm, bps, rate))
    env = dict(os.environ, FLAKE_B200_BATCH_BLOCKS="4")      # three batches + the short block
    for level in ("-0", "-5", "-8", "-9", "-12"):
        a, b = tmp_path / "a.flac", tmp_path / "b.flac"
        subprocess.run([cli, "-q", level, str(wav), "-o", str(a)], check=True, env=env)
        subprocess.run([pyoracle.REF_CLI, "-q", level, str(wav), "-o", str(b)], check=True)
        assert a.read_bytes() == b.read_bytes(), level
    a, b = tmp_path / "a.flac", tmp_path / "b.flac"
    subprocess.run([cli, "-q", "-8", str(wav), "-o", str(a)], check=True)          # default batch size: one call
    subprocess.run([pyoracle.REF_CLI, "-q", "-8", str(wav), "-o", str(b)], check=True)
    assert a.read_bytes() == b.read_bytes()


def test_plain_c_caller(gpu_lib, oracle, tmp_path):
    """tests/c/batch_example.c: a C program in the shape of util/api_example.c, compiled against
    include/*.h and linked with libflake.so; per-block and batch outputs must be one valid file."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "batch_example"
    subprocess.run(["gcc", "-O1", os.path.join(root, "tests", "c", "batch_example.c"),
                    "-I", os.path.join(root, "include"), "-L", os.path.join(root, "flake_b200", "lib"),
                    "-lflake", "-Wl,-rpath," + os.path.join(root, "flake_b200", "lib"), "-o", str(exe)],
                   check=True)
    a, b = tmp_path / "a.flac", tmp_path / "b.flac"
    subprocess.run([str(exe), str(a), str(b), "8"], check=True)
    da, db = a.read_bytes(), b.read_bytes()
    assert da == db
    dec, info = oracle.decode(da)
    assert info.md5_ok == 1 and info.decoded_samples == 4096 * 20 + 1234


@pytest.mark.gpu
@pytest.mark.parametrize("level", [8, 9])
def test_seektable_and_frame_sizes_from_a_real_encode(gpu_lib, oracle, level):
    """SEEKTABLE built from the frame lengths / block sizes the batch call returns
    (level 9: variable block size, several frames per block), and STREAMINFO's opt-in minimum frame
    size.  Every seek point must be the byte offset of a frame whose header carries the point's
    sample number; the stream with the table inserted must still decode bit-exactly."""
    import struct
    ch, bps, rate = 2, 16, 44100
    n = 4096 * 40 + 1000
    pcm = synth.synth_pcm(n, ch, bps, rate, seed=11)
    if level == 9:        # bursts that make the VBS splitter cut blocks, then the same short tail
        pcm = np.concatenate([_vbs_pcm(4096 * 40, ch, bps, 13), pcm[4096 * 40:]])
    enc = api.Encoder(gpu_lib, ch, rate, bps, n, level)
    header = enc.init()
    try:
        assert gpu_lib.flake_b200_set_streaminfo_sizes(C.byref(enc.ctx), 1) == 0
        data, flen, fbs = enc.encode_stream(pcm)
        si, si_bytes = enc.streaminfo()
    finally:
        enc.close()
    frames = data.tobytes()
    assert si.min_frame_size == int(flen.min()) and si.max_frame_size >= int(flen.max())
    interval = 10 * 4096
    need = gpu_lib.flake_b200_write_seektable(flen.ctypes.data, fbs.ctypes.data, len(flen), interval, None, 0)
    assert need > 0 and need % 18 == 0
    buf = (C.c_ubyte * need)()
    assert gpu_lib.flake_b200_write_seektable(flen.ctypes.data, fbs.ctypes.data, len(flen), interval, buf, need) == need
    offs = np.concatenate([[0], np.cumsum(flen)]).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(fbs)]).astype(np.int64)
    pts = [struct.unpack(">QQH", bytes(buf)[i:i + 18]) for i in range(0, need, 18)]
    assert pts[0][:2] == (0, 0) and len(pts) >= n // interval
    for sample, off, count in pts:
        f = int(np.searchsorted(starts, sample))
        assert starts[f] == sample and offs[f] == off and fbs[f] == count          # a frame boundary, its size
        assert frames[off] == 0xff and (frames[off + 1] & 0xfe) == 0xf8              # frame sync code
    # the table as a metadata block (type 3) between STREAMINFO and the rest: still one valid stream
    hdr = bytearray(header)
    hdr[8:8 + 34] = si_bytes
    with_table = bytes(hdr[:42]) + bytes([3]) + struct.pack(">I", need)[1:] + bytes(buf) + bytes(hdr[42:]) + frames
    dec, info = oracle.decode(with_table)
    assert info.md5_ok == 1 and np.array_equal(dec, pcm)


def _vbs_pcm(n, ch, bps, seed):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.vbs_burst_pcm(n, ch, bps, seed, 4096)
