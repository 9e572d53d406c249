"""Kernel LOGIC on the CPU: the CUDA sources of flake_b200/csrc compiled for the test-only
fiber emulator (tests/cuda_emu) and driven through the same C ABI.  This is not the product
path and not a fallback -- it lets the CPU-only CI exercise scans, the bit packer, the search
control flow and barrier placement; the GPU parity tests (-m gpu) are the gate for the real build."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from flake_b200 import api, synth


def _vbs_burst(n, ch, bps, seed, blk=4096):
    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.vbs_burst_pcm(n, ch, bps, seed, blk)


CASES = [
    ("l0", 1152 + 300, 2, 16, 44100, "mix", 0, {}),
    ("l2", 1152 + 301, 2, 16, 44100, "mix", 2, {}),
    ("l5", 2048 + 100, 2, 16, 44100, "mix", 5, {"block_size": 2048}),
    ("l7", 1024 * 2, 2, 16, 44100, "mix", 7, {"block_size": 1024}),
    ("l8", 4096 + 500, 2, 16, 44100, "mix", 8, {}),
    ("l8_mono_tail", 1024 + 7, 1, 16, 44100, "mix", 8, {"block_size": 1024}),
    ("l10_s24", 1024 * 2, 2, 24, 96000, "mix", 10, {"block_size": 1024}),
    ("l12_s24", 2048, 2, 24, 96000, "mix", 12, {"block_size": 2048}),
    ("l9_3ch", 1024 + 256, 3, 24, 48000, "impulses", 9, {"block_size": 1024}),
    ("noise_l0", 1152, 2, 16, 44100, "noise", 0, {}),
    # full-scale 24-bit noise: run sums of 2^28, partition sums beyond 2^32 -> the 64-bit finish of k_search
    ("noise_s24_l8", 1024 * 2, 2, 24, 96000, "noise", 8, {"block_size": 1024}),
    ("wasted_l5", 1024, 2, 16, 44100, "wasted", 5, {"block_size": 1024}),
    ("silence_l8", 1024, 2, 16, 44100, "silence", 8, {"block_size": 1024}),
    # the fast finish of k_search (16-bit, power-of-two blocks of 512..4096): partition order up to 8 (the level
    # of single runs), one span only, Rice parameters above 14 (RICE2) and the variable-block-size splits
    ("fused_l9_vbs", 4096 + 100, 2, 16, 44100, "mix", 9, {}),
    ("fused_l10_bs512", 512 * 3, 2, 16, 44100, "mix", 10, {"block_size": 512, "variable_block_size": 0}),
    ("fused_l8_bs2048", 2048 * 2 + 17, 2, 16, 44100, "mix", 8, {"block_size": 2048}),
    ("fused_noise_l8", 4096, 2, 16, 44100, "noise", 8, {}),
    ("fused_noise_l10", 2048, 1, 16, 44100, "noise", 10, {"block_size": 2048, "variable_block_size": 0}),
    # blocks of 8192: the order-32 kernel on 24-bit and on 16-bit input; 16-bit stereo frames of that size are four
    # chunks through the two-stage TMA ring (a stage is refilled while fibers run ahead of the issuing one)
    ("bs8192_l12_s24", 8192 + 300, 2, 24, 96000, "mix", 12, {"variable_block_size": 0}),
    ("bs8192_l11_s16", 8192, 2, 16, 44100, "mix", 11, {}),
    ("bs4096_l9_s24_8ch", 4096, 8, 24, 48000, "mix", 9, {"variable_block_size": 0}),
]


@pytest.mark.parametrize("lpc", ["throughput", "latency"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emulated_kernels_match_oracle(case, lpc, emu_lib, oracle, monkeypatch):
    # small passes would all take the latency form of the LPC analysis (k_lpc_lat): run both kernels
    monkeypatch.setenv("FLAKE_B200_LPC_LAT_MAX", "0" if lpc == "throughput" else "100000")
    name, n, ch, bps, rate, kind, level, ov = case
    pcm = synth.synth_pcm(n, ch, bps, rate, seed=len(name) * 7 + level, kind=kind)
    got = api.encode_batch(emu_lib, pcm, rate, bps, level, chunk_blocks=2, **ov)
    want, flen, fbs, mx = oracle.encode_stream(pcm, rate, bps, level, **ov)
    assert list(map(len, got.frames)) == list(flen)
    assert got.payload == want
    p = oracle.make_params(ch, rate, bps, level, n, **ov)
    assert got.header == oracle.header(p)
    assert got.streaminfo == oracle.streaminfo(p, mx, oracle.md5_pcm(pcm, bps))


def test_emulated_vbs_split(emu_lib, oracle):
    pcm = _vbs_burst(1024 * 3, 2, 16, seed=5, blk=1024)
    ov = {"block_size": 1024}
    got = api.encode_batch(emu_lib, pcm, 44100, 16, 9, chunk_blocks=2, **ov)
    want, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, 9, **ov)
    assert len(flen) > 3, "input should force VBS splits"
    assert list(got.frame_bs) == list(fbs)
    assert got.payload == want


def test_emulated_per_block_api(emu_lib, oracle):
    pcm = synth.synth_pcm(1024 * 2 + 100, 2, 16, 44100, seed=2)
    got = api.encode_per_block(emu_lib, pcm, 44100, 16, 5, block_size=1024)
    want, *_ = oracle.encode_stream(pcm, 44100, 16, 5, block_size=1024)
    assert got.payload == want
    # the short last block latches the stream shut (encode.c:989-994)
    enc = api.Encoder(emu_lib, 2, 44100, 16, 0, 5, block_size=1024)
    enc.init()
    enc.encode_frame(pcm[:100])
    with pytest.raises(api.FlakeLibraryError):
        enc.encode_frame(pcm[:1024])
    enc.close()


def test_subframe_decisions_exposed(emu_lib, oracle):
    """Stage-level parity: type / order / shift / coefficients / Rice parameters per subframe."""
    import ctypes as C
    pcm = synth.synth_pcm(2048, 2, 16, 44100, seed=9)
    enc = api.Encoder(emu_lib, 2, 44100, 16, 2048, 8, block_size=2048)
    enc.init()
    enc.encode_stream(pcm)
    subs = enc.last_subframes(16)
    enc.close()
    info = oracle.OrcFrameInfo()
    p = oracle.make_params(2, 44100, 16, 8, 2048, block_size=2048)
    out = np.zeros(1 << 16, dtype=np.uint8)
    oracle.lib().orc_encode_frame(C.byref(p), pcm.ctypes.data, 2048, 0, out.ctypes.data, len(out), C.byref(info))
    assert len(subs) == 2
    for c in range(2):
        s, o = subs[c], info.sub[c]
        assert (s.type, s.order, s.obits, s.wasted, s.method, s.porder, s.est_bits) == \
               (o.type, o.order, o.obits, o.wasted, o.method, o.porder, o.est_bits)
        if s.type == 32:
            assert s.shift == o.shift and list(s.coefs[:s.order]) == list(o.coefs[:o.order])
        assert list(s.params[:1 << s.porder]) == list(o.params[:1 << o.porder])


def test_rice_parameter_closed_form(emu_lib, oracle):
    """fb_rice_k's closed form (dev_common.cuh) against the reference's scan (rice.c:30-45)."""
    L = oracle.lib()
    emu_lib.fb_test_rice_k.argtypes = [C.c_uint64, C.c_int]
    rng = np.random.Generator(np.random.PCG64(3))
    cases = [(0, 0), (0, 1), (1, 1), (31, 64), (32, 64), (33, 64), (128, 64), (129, 64), (2**31 - 1, 4096),
             (2**31, 4096), (2**40, 4096), (2**43 + 12345, 65535), (5, 65535)]
    for _ in range(20000):
        n = int(rng.choice([0, 1, 2, 3, 15, 16, 17, 52, 64, 255, 256, 4084, 4096, 8192, 65535]))
        s = int(2 ** float(rng.uniform(0, 45))) + int(rng.integers(-600, 600))
        cases.append((max(0, s) if n else 0, n))
    for n in (1, 2, 3, 16, 64):
        for s in range(0, 40 * n):
            cases.append((s, n))
    for s, n in cases:
        assert emu_lib.fb_test_rice_k(s, n) == L.orc_rice_k(s, n), (s, n)


def test_crc_combination_algebra(emu_lib, oracle):
    """x^(8n) mod P multipliers used to merge per-thread CRC-16 chunks (k_pack.cuh)."""
    L = oracle.lib()
    emu_lib.fb_test_gf16_xpow8.restype = C.c_uint32
    emu_lib.fb_test_gf16_mul.restype = C.c_uint32
    rng = np.random.Generator(np.random.PCG64(4))
    for _ in range(50):
        a = bytes(rng.integers(0, 256, size=int(rng.integers(1, 200)), dtype=np.uint8))
        b = bytes(rng.integers(0, 256, size=int(rng.integers(0, 200)), dtype=np.uint8))
        ca, cb = L.orc_crc16(a, len(a)), L.orc_crc16(b, len(b))
        m = emu_lib.fb_test_gf16_xpow8(len(b))
        assert emu_lib.fb_test_gf16_mul(ca, m) ^ cb == L.orc_crc16(a + b, len(a) + len(b))


# Order-search control flow of k_search: candidates are costed in groups and the searches of
# optimize.c:205-261 replayed on the stored totals; the group planner of the log search merges
# steps whose candidate sets do not depend on pending results.  Sweep the (method, min, max)
# space, tileable and non-tileable block sizes.
ORDER_SWEEP = [(om, lo, hi, bs)
               for om in (2, 3, 4, 5, 6)
               for (lo, hi) in ((1, 12), (1, 8), (3, 9), (5, 5), (1, 32), (8, 32), (2, 3), (1, 2), (12, 12), (7, 20), (1, 17))
               for bs in (1024, 576)]


@pytest.mark.parametrize("om,lo,hi,bs", ORDER_SWEEP, ids=["om%d_%d_%d_bs%d" % c for c in ORDER_SWEEP])
def test_order_search_sweep(om, lo, hi, bs, emu_lib, oracle):
    ov = {"order_method": om, "min_prediction_order": lo, "max_prediction_order": hi, "block_size": bs,
          "prediction_type": 2, "variable_block_size": 0}
    pcm = synth.synth_pcm(bs * 2 + 40, 2, 16, 44100, seed=om * 100 + lo * 7 + hi, kind="mix")
    got = api.encode_batch(emu_lib, pcm, 44100, 16, 8, chunk_blocks=4, **ov)
    want, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, 8, **ov)
    assert got.payload == want


@pytest.mark.parametrize("lo,hi", [(0, 4), (0, 0), (2, 4), (1, 3), (4, 4)])
def test_fixed_order_sweep(lo, hi, emu_lib, oracle):
    ov = {"prediction_type": 1, "min_prediction_order": lo, "max_prediction_order": hi, "block_size": 1152}
    pcm = synth.synth_pcm(1152 * 2 + 9, 2, 16, 44100, seed=lo * 5 + hi, kind="mix")
    got = api.encode_batch(emu_lib, pcm, 44100, 16, 2, chunk_blocks=4, **ov)
    want, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, 2, **ov)
    assert got.payload == want
