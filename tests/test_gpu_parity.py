"""Parity of the CUDA path (through the C ABI of flake_b200/lib/libflake.so) against the
oracle restatement (and the compiled reference when oracle/_ref travelled with the repo).

Bar: bit-exact frames, header and final STREAMINFO; the stream must also decode
back to the input with all CRCs and the MD5 valid.
"""
import ctypes as C
import zlib

import numpy as np
import pytest

from flake_b200 import api, synth

pytestmark = pytest.mark.gpu

CASES = [
    # (name, nsamples, channels, bps, rate, kind, level, overrides)
    ("l0_s16",   1152 * 9 + 77,  2, 16, 44100, "mix", 0, {}),
    ("l1_s16",   1152 * 9 + 78,  2, 16, 44100, "mix", 1, {}),
    ("l2_s16",   1152 * 9 + 80,  2, 16, 44100, "mix", 2, {}),
    ("l3_s16",   4096 * 5 + 100, 2, 16, 44100, "mix", 3, {}),
    ("l5_s16",   4096 * 9 + 3936 % 4096, 2, 16, 44100, "mix", 5, {}),
    ("l6_s16",   4096 * 5 + 200, 2, 16, 44100, "mix", 6, {}),
    ("l7_s16",   4096 * 5 + 300, 2, 16, 44100, "mix", 7, {}),
    ("l8_s16",   4096 * 9 + 3136, 2, 16, 44100, "mix", 8, {}),
    ("l9_s16",   4096 * 6, 2, 16, 44100, "impulses", 9, {}),
    ("l10_s16",  4096 * 4, 2, 16, 44100, "impulses", 10, {}),
    ("l11_s24",  8192 * 3 + 2048, 2, 24, 96000, "mix", 11, {}),
    ("l12_s24",  8192 * 3 + 2048, 2, 24, 96000, "impulses", 12, {}),
    # CD audio at the two highest presets: blocks of 8192 (four chunks through the two-stage TMA ring), the
    # order-32 kernel on 16-bit input
    ("l11_s16",  8192 * 3 + 1000, 2, 16, 44100, "mix", 11, {}),
    ("l12_s16",  8192 * 2 + 4096, 2, 16, 44100, "impulses", 12, {}),
    ("l8_bs8192_s16", 8192 * 3, 2, 16, 44100, "mix", 8, {"block_size": 8192}),
    ("l9_8ch",   4096 * 3 + 1024, 8, 24, 48000, "impulses", 9, {}),
    ("l8_mono",  4096 * 3 + 10, 1, 16, 44100, "mix", 8, {}),
    ("l8_noise", 4096 * 3, 2, 16, 44100, "noise", 8, {}),
    ("l0_noise", 1152 * 5, 2, 16, 44100, "noise", 0, {}),
    # full-scale 24-bit noise: zig-zag totals beyond 2^32 per block -> the 64-bit finish of k_search
    ("l8_noise_s24", 4096 * 3, 2, 24, 96000, "noise", 8, {}),
    ("l12_noise_s24", 8192 * 2, 2, 24, 96000, "noise", 12, {}),
    ("l5_wasted", 4096 * 3, 2, 16, 44100, "wasted", 5, {}),
    ("l5_silence", 4096 * 3, 2, 16, 44100, "silence", 5, {}),
    ("l8_tiny_tail", 4096 + 7, 2, 16, 44100, "mix", 8, {}),
    ("l8_tail3", 4096 + 3, 2, 16, 44100, "mix", 8, {}),
    ("l5_8bit", 4096 * 2 + 64, 2, 8, 22050, "mix", 5, {}),
    ("l8_custom_rate", 4096 * 2, 2, 16, 37800, "mix", 8, {}),
    ("l5_bs1000", 1000 * 5 + 10, 2, 16, 44100, "mix", 5, {"block_size": 1000}),
    ("l8_max", 4096 * 2, 2, 16, 44100, "mix", 8, {"order_method": 0}),
    ("l8_2level", 4096 * 2, 2, 16, 44100, "mix", 8, {"order_method": 2}),
    ("l8_8level", 4096 * 2, 2, 16, 44100, "mix", 8, {"order_method": 4}),
    ("l9_novbs", 4096 * 3, 2, 16, 44100, "mix", 9, {"variable_block_size": 0}),
    # maximum block sizes: blocks that do not fit shared memory take the global-memory paths
    ("bs65535_8ch_s24", 65535 + 4000, 8, 24, 48000, "mix", 8, {"block_size": 65535}),
    ("bs65535_mono", 65535 * 2, 1, 16, 44100, "mix", 5, {"block_size": 65535}),
    ("bs32768_l8", 32768 * 2 + 10, 2, 16, 44100, "mix", 8, {"block_size": 32768}),
    ("bs32768_l12_vbs", 32768 * 2, 2, 24, 96000, "impulses", 12, {"block_size": 32768}),
    ("bs16_min", 16 * 40 + 5, 2, 16, 44100, "mix", 5, {"block_size": 16}),
    ("bs128_vbs", 128 * 20, 2, 16, 44100, "impulses", 9, {"block_size": 128}),
    ("l5_7ch_20bit", 4096 * 2 + 50, 7, 20, 48000, "mix", 5, {}),
    ("l8_pmin3", 4096 * 2, 2, 16, 44100, "mix", 8, {"min_partition_order": 3}),
    ("l8_order32_log", 4096 * 2, 2, 16, 44100, "mix", 8, {"max_prediction_order": 32}),
    ("l5_odd_block", 4095 * 2, 2, 16, 44100, "mix", 2, {"block_size": 4095}),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_batch_matches_oracle(case, gpu_lib, oracle, monkeypatch):
    name, n, ch, bps, rate, kind, level, ov = case
    pcm = synth.synth_pcm(n, ch, bps, rate, seed=zlib.crc32(name.encode()) % 1000, kind=kind)
    # passes this small take the latency form of the LPC analysis (k_lpc_lat); the throughput form
    # (k_lpc) must produce the same bytes
    monkeypatch.setenv("FLAKE_B200_LPC_LAT_MAX", "0")
    bulk = api.encode_batch(gpu_lib, pcm, rate, bps, level, chunk_blocks=4, **ov)
    monkeypatch.delenv("FLAKE_B200_LPC_LAT_MAX")
    got = api.encode_batch(gpu_lib, pcm, rate, bps, level, chunk_blocks=4, **ov)
    assert got.payload == bulk.payload and got.streaminfo == bulk.streaminfo
    want, flen, fbs, mx = oracle.encode_stream(pcm, rate, bps, level, **ov)
    assert list(map(len, got.frames)) == list(flen)
    assert list(got.frame_bs) == list(fbs)
    assert got.payload == want
    p = oracle.make_params(ch, rate, bps, level, n, **ov)
    assert got.header == oracle.header(p)
    assert got.streaminfo == oracle.streaminfo(p, mx, oracle.md5_pcm(pcm, bps))
    dec, info = oracle.decode(got.file_bytes())
    assert info.md5_ok == 1
    assert np.array_equal(dec, pcm)


@pytest.mark.parametrize("level", [0, 5, 8, 9])
def test_per_block_api_matches_batch(level, gpu_lib):
    pcm = synth.synth_pcm(4096 * 3 + 512, 2, 16, 44100, seed=7, kind="impulses")
    a = api.encode_per_block(gpu_lib, pcm, 44100, 16, level)
    b = api.encode_batch(gpu_lib, pcm, 44100, 16, level, chunk_blocks=2)
    assert a.payload == b.payload
    assert a.header == b.header and a.streaminfo == b.streaminfo


@pytest.mark.parametrize("fmt,bps", [(api.PCM_S16LE, 16), (api.PCM_S24LE, 24)])
def test_packed_pcm_ingest(fmt, bps, gpu_lib):
    pcm = synth.synth_pcm(4096 * 2 + 100, 2, bps, 48000, seed=11)
    packed = np.frombuffer(synth.pack_pcm(pcm, bps), dtype=np.uint8)
    a = api.encode_batch(gpu_lib, pcm, 48000, bps, 8)
    b = api.encode_batch(gpu_lib, packed, 48000, bps, 8, pcm_format=fmt,
                         nsamples=pcm.shape[0], channels=2)
    assert a.payload == b.payload and a.streaminfo == b.streaminfo


def test_out_of_range_int32_input_is_passed_through(gpu_lib, oracle):
    """flake_encode_frame takes int32: samples wider than bits_per_sample are the caller's bug,
    but the encoder must still see exactly those values (the host layer packs to
    ceil(bps/8) bytes for the upload only when that is lossless) and the MD5 must hash the
    low bytes like md5.c:296-309."""
    pcm = synth.synth_pcm(4096 * 3, 2, 16, 44100, seed=5)
    pcm[1000:1100] *= 5          # beyond 16 bits
    pcm[5000, 1] = 1 << 20
    for level in (2, 8):
        got = api.encode_batch(gpu_lib, pcm, 44100, 16, level, chunk_blocks=2)
        want, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, level)
        assert got.payload == want
        p = oracle.make_params(2, 44100, 16, level, pcm.shape[0])
        assert got.streaminfo == oracle.streaminfo(p, mx, oracle.md5_pcm(pcm, 16))
        blk = api.encode_per_block(gpu_lib, pcm, 44100, 16, level)
        assert blk.payload == want and blk.streaminfo == got.streaminfo


def test_matches_compiled_reference(gpu_lib, oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ref = oracle.ref_library()
    for level, ch, bps, rate in [(5, 2, 16, 44100), (8, 2, 16, 44100), (12, 2, 24, 96000)]:
        pcm = synth.synth_pcm(8192 * 2 + 2048, ch, bps, rate, seed=level)
        want = api.encode_per_block(ref, pcm, rate, bps, level)
        got = api.encode_batch(gpu_lib, pcm, rate, bps, level)
        assert got.file_bytes() == want.file_bytes()


# order-search control flow (candidate groups + replay, see tests/test_emu_kernels.py) on the device
ORDER_SWEEP = [(6, 1, 12, 4096), (6, 1, 8, 4096), (6, 3, 9, 1024), (6, 1, 32, 4096), (6, 8, 32, 2048), (6, 7, 20, 576),
               (6, 1, 2, 4608), (6, 12, 12, 4096), (5, 1, 12, 4096), (5, 1, 32, 1024), (5, 5, 5, 576), (4, 1, 12, 4096),
               (4, 1, 32, 4608), (3, 1, 12, 4096), (3, 2, 3, 1024), (2, 1, 12, 4096), (2, 8, 32, 4096)]


@pytest.mark.parametrize("om,lo,hi,bs", ORDER_SWEEP, ids=["om%d_%d_%d_bs%d" % c for c in ORDER_SWEEP])
def test_order_search_sweep(om, lo, hi, bs, gpu_lib, oracle):
    ov = {"order_method": om, "min_prediction_order": lo, "max_prediction_order": hi, "block_size": bs,
          "prediction_type": 2, "variable_block_size": 0}
    pcm = synth.synth_pcm(bs * 5 + 40, 2, 16, 44100, seed=om * 100 + lo * 7 + hi, kind="mix")
    got = api.encode_batch(gpu_lib, pcm, 44100, 16, 8, chunk_blocks=4, **ov)
    want, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, 8, **ov)
    assert got.payload == want


def test_int32_input_beyond_24_bits_estimate_is_exact(gpu_lib, oracle):
    """The stereo estimate sums in 32 bits per run only for the packed formats; int32 input that
    exceeds its declared width must still give the reference's decision (64-bit sums)."""
    pcm = synth.synth_pcm(4096 * 2, 2, 16, 44100, seed=9)
    pcm[100:3000, 0] = (pcm[100:3000, 0].astype(np.int64) * 30000).clip(-2**31 + 1, 2**31 - 1).astype(np.int32)
    got = api.encode_batch(gpu_lib, pcm, 44100, 16, 8, chunk_blocks=2)
    want, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, 8)
    assert got.payload == want


def test_encode_stream_from_worker_threads(gpu_lib, oracle):
    """The CUDA current device is per host thread; a context must work from any thread (bench.py
    encodes several streams concurrently, one thread each)."""
    import threading
    pcm = synth.synth_pcm(4096 * 6 + 100, 2, 16, 44100, seed=21)
    want, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, 8)
    encs = [api.Encoder(gpu_lib, 2, 44100, 16, pcm.shape[0], 8) for _ in range(3)]
    for e in encs:
        e.init()
    got = [None] * 3

    def work(i):
        for _ in range(2):                      # second round: engine and lanes already exist
            gpu_lib.flake_b200_reset_stream(C.byref(encs[i].ctx))
            got[i] = encs[i].encode_stream(pcm, api.PCM_S32, pcm.shape[0])

    ths = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for e in encs:
        e.close()
    assert all(g is not None and g[0].tobytes() == want for g in got)
