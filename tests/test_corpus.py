"""flake_b200_encode_corpus (include/flake_b200.h, csrc/flake_corpus.c): many streams -- or one
stream cut into chunks -- over the GPUs of a box, MD5 by multi-buffer SIMD workers.

CPU part: the multi-buffer MD5 against hashlib (ragged lengths, every lane count), and the corpus
host logic (chunk prefix sums, lane refill, staged and direct copies, error paths) on the
fiber-emulated kernels with ONE GPU worker thread (the emulator is single-threaded).
GPU part (-m gpu): the same through the shipped library, compared with the per-stream batch call
(itself byte-checked against the reference in test_gpu_golden.py) and with the oracle.
"""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from flake_b200 import api, corpus, synth


class FbMd5(C.Structure):
    _fields_ = [("h", C.c_uint32 * 4), ("nbytes", C.c_uint64), ("tail", C.c_uint8 * 64)]


def _md5_mb(lib, bufs):
    n = len(bufs)
    ctx = (FbMd5 * n)()
    for i in range(n):
        lib.fb_md5_init(C.byref(ctx[i]))
    cp = (C.POINTER(FbMd5) * n)(*[C.pointer(ctx[i]) for i in range(n)])
    keep = [np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8) for b in bufs]
    dp = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
    ln = (C.c_size_t * n)(*[len(b) for b in bufs])
    lib.fb_md5_mb_update(cp, dp, ln, n)
    out = []
    for i in range(n):
        d = (C.c_uint8 * 16)()
        lib.fb_md5_final(C.byref(ctx[i]), d)
        out.append(bytes(d))
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 8, 15, 16, 17, 31, 32])
def test_multibuffer_md5_matches_hashlib(emu_lib, n):
    rng = np.random.default_rng(n)
    lens = [int(x) for x in rng.integers(0, 5000, size=n)]
    lens[0] = 64 * 40 + 3                      # at least one long stream, one exact multiple, one empty
    if n > 1:
        lens[1] = 64 * 7
    if n > 2:
        lens[2] = 0
    bufs = [rng.integers(0, 256, size=l, dtype=np.uint8).tobytes() for l in lens]
    assert emu_lib.fb_md5_mb_lanes() in (1, 16, 32)
    got = _md5_mb(emu_lib, bufs)
    assert got == [hashlib.md5(b).digest() for b in bufs]


def test_multibuffer_md5_continues_partial_blocks(emu_lib):
    """Two updates per stream, the first leaving every stream in the middle of a block."""
    rng = np.random.default_rng(5)
    n = 20
    bufs = [rng.integers(0, 256, size=1000 + 37 * i, dtype=np.uint8).tobytes() for i in range(n)]
    ctx = (FbMd5 * n)()
    for i in range(n):
        emu_lib.fb_md5_init(C.byref(ctx[i]))
    cp = (C.POINTER(FbMd5) * n)(*[C.pointer(ctx[i]) for i in range(n)])
    keep = [np.frombuffer(b, dtype=np.uint8) for b in bufs]
    for lo, hi in ((0, 333), (333, None)):
        parts = [k[lo:hi] for k in keep]
        dp = (C.c_void_p * n)(*[p.ctypes.data for p in parts])
        ln = (C.c_size_t * n)(*[p.size for p in parts])
        emu_lib.fb_md5_mb_update(cp, dp, ln, n)
    for i in range(n):
        d = (C.c_uint8 * 16)()
        emu_lib.fb_md5_final(C.byref(ctx[i]), d)
        assert bytes(d) == hashlib.md5(bufs[i]).digest()


def _streams(count, ch, bps, rate, block, seed0=0):
    """ragged corpus: lengths from a few samples to several blocks, one exact multiple, one empty"""
    lens = [block * 3 + 17, block * 2, 5, block + 1, 0, block * 5 - 1, block * 4 + 100][:count]
    while len(lens) < count:
        lens.append(block * (1 + len(lens) % 4) + 11 * len(lens))
    return [synth.synth_pcm(max(n, 1), ch, bps, rate, seed=seed0 + i)[:n] for i, n in enumerate(lens)]


def _check_against_single(lib, oracle, pcms, ch, bps, rate, level, fmt, results, **ov):
    for pcm, r in zip(pcms, results):
        n = pcm.shape[0]
        want_md5 = hashlib.md5(_digest_layout(pcm, bps)).digest()
        assert r.md5 == want_md5
        if n == 0:
            assert r.data.size == 0 and r.frame_len.size == 0
            continue
        want, flen, fbs, mx = oracle.encode_stream(pcm, rate, bps, level, **ov)
        assert r.data.tobytes() == want
        assert list(r.frame_len) == list(flen) and list(r.frame_bs) == list(fbs)
        assert r.max_frame_size == mx
        p = oracle.make_params(ch, rate, bps, level, n, **ov)
        hdr = bytearray(oracle.header(p))
        hdr[8:8 + 34] = oracle.streaminfo(p, mx, want_md5)
        assert r.header == bytes(hdr)


def _digest_layout(pcm, bps):
    nb = (bps + 7) // 8
    if nb == 2:
        return pcm.astype("<i2").tobytes()
    if nb == 1:
        return pcm.astype("i1").tobytes()
    if nb == 3:
        b = pcm.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        return np.ascontiguousarray(b).tobytes()
    return pcm.astype("<i4").tobytes()


def _packed(pcm, bps):
    return np.frombuffer(_digest_layout(pcm, bps), dtype=np.uint8)


@pytest.mark.parametrize("pinned", ["0", "1"], ids=["staged", "direct"])
def test_corpus_on_emulated_kernels(emu_lib, oracle, monkeypatch, pinned):
    """chunks of 2 blocks: streams span several chunks, prefix sums in chunk order, ragged ends"""
    monkeypatch.setenv("CUEMU_ALL_PINNED", pinned)
    ch, bps, rate, level, ov = 2, 16, 44100, 8, {"block_size": 1024}
    pcms = _streams(7, ch, bps, rate, 1024)
    with corpus.Corpus(emu_lib, ch, rate, bps, level, api.PCM_S16LE, devices=[0], threads_per_device=1,
                       md5_threads=3, chunk_blocks=2, **ov) as co:
        results, st = co.encode([_packed(p, bps) for p in pcms], nsamples=[p.shape[0] for p in pcms])
        assert st.streams == 7 and st.gpu_threads == 1 and st.md5_threads == 1 and st.md5_lanes == 7
        assert st.chunks == sum((p.shape[0] + 2047) // 2048 for p in pcms)
        _check_against_single(emu_lib, oracle, pcms, ch, bps, rate, level, api.PCM_S16LE, results, **ov)
        # the handle is reusable
        results2, _ = co.encode([_packed(p, bps) for p in pcms[:2]], nsamples=[p.shape[0] for p in pcms[:2]])
        assert results2[0].data.tobytes() == results[0].data.tobytes()


def test_corpus_int32_and_24bit_on_emulated_kernels(emu_lib, oracle):
    ch, bps, rate, level, ov = 2, 24, 96000, 5, {"block_size": 1024}
    pcms = _streams(4, ch, bps, rate, 1024, seed0=40)
    res, st = corpus.encode_corpus(emu_lib, [np.ascontiguousarray(p, dtype=np.int32) for p in pcms], ch, rate, bps,
                                   level, api.PCM_S32, devices=[0], threads_per_device=1, chunk_blocks=2, **ov)
    _check_against_single(emu_lib, oracle, pcms, ch, bps, rate, level, api.PCM_S32, res, **ov)
    with corpus.Corpus(emu_lib, ch, rate, bps, level, api.PCM_S24LE, devices=[0], threads_per_device=1,
                       chunk_blocks=3, **ov) as co:
        res, _ = co.encode([_packed(p, bps) for p in pcms], nsamples=[p.shape[0] for p in pcms])
    _check_against_single(emu_lib, oracle, pcms, ch, bps, rate, level, api.PCM_S24LE, res, **ov)


def test_corpus_vbs_on_emulated_kernels(emu_lib, oracle):
    """variable block size: frames per chunk are data dependent, header numbers count samples"""
    ch, bps, rate, level, ov = 2, 16, 44100, 9, {"block_size": 1024}
    pcms = _streams(3, ch, bps, rate, 1024, seed0=7)
    with corpus.Corpus(emu_lib, ch, rate, bps, level, api.PCM_S16LE, devices=[0], threads_per_device=1,
                       chunk_blocks=2, **ov) as co:
        res, _ = co.encode([_packed(p, bps) for p in pcms], nsamples=[p.shape[0] for p in pcms])
    _check_against_single(emu_lib, oracle, pcms, ch, bps, rate, level, api.PCM_S16LE, res, **ov)


def test_corpus_errors_on_emulated_kernels(emu_lib):
    ch, bps, rate, level, ov = 2, 16, 44100, 5, {"block_size": 1024}
    pcm = synth.synth_pcm(1024 * 4, ch, bps, rate, seed=1)
    with corpus.Corpus(emu_lib, ch, rate, bps, level, api.PCM_S16LE, devices=[0], threads_per_device=1,
                       chunk_blocks=2, **ov) as co:
        small = np.empty(100, dtype=np.uint8)              # output too small: -2 for that stream only
        good = np.empty(co.max_encoded_size(pcm.shape[0]), dtype=np.uint8)
        items = (api.FlakeB200CorpusStream * 2)()
        p = _packed(pcm, bps)
        for it, out in zip(items, (small, good)):
            it.pcm = p.ctypes.data; it.nsamples = pcm.shape[0]; it.out = out.ctypes.data; it.out_cap = out.nbytes
        st = api.FlakeB200CorpusStats()
        rc = emu_lib.flake_b200_corpus_encode(co.handle, items, 2, C.byref(st))
        assert rc == -2 and items[0].bytes == -2 and items[1].bytes > 0
    # container that is not the digest layout, bad device, bad parameters
    ctx = api.Encoder(emu_lib, ch, rate, bps, 0, level).ctx
    assert not emu_lib.flake_b200_corpus_open(C.byref(ctx), api.PCM_S24LE, None, 0, None)
    devs = (C.c_int * 1)(7)
    assert not emu_lib.flake_b200_corpus_open(C.byref(ctx), api.PCM_S16LE, devs, 1, None)
    ctx.channels = 9
    assert not emu_lib.flake_b200_corpus_open(C.byref(ctx), api.PCM_S16LE, None, 0, None)


def test_corpus_without_a_gpu_fails_loudly(gpu_lib):
    """no CUDA device -> no handle (there is no CPU path); with a device the GPU tests below run"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ctx = api.Encoder(gpu_lib, 2, 44100, 16, 0, 8).ctx
    assert not gpu_lib.flake_b200_corpus_open(C.byref(ctx), api.PCM_S16LE, None, 0, None)
    items = (api.FlakeB200CorpusStream * 1)()
    assert gpu_lib.flake_b200_encode_corpus(C.byref(ctx), api.PCM_S16LE, items, 0, None, 0, None, None) == -3


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [
    ("l8_s16", 2, 16, 44100, 8, {}),
    ("l9_vbs_8ch_s24", 8, 24, 48000, 9, {}),
    ("l5_mono_s8", 1, 8, 8000, 5, {}),
], ids=lambda c: c[0])
def test_corpus_matches_the_single_stream_call(gpu_lib, oracle, cfg):
    import torch
    name, ch, bps, rate, level, ov = cfg
    ndev = torch.cuda.device_count()
    block = 4096
    pcms = _streams(9, ch, bps, rate, block, seed0=100)
    fmt = {8: api.PCM_S8, 16: api.PCM_S16LE, 24: api.PCM_S24LE}[bps]
    # pinned inputs for the even streams (DMA straight from the caller's buffer), pageable for the odd ones
    packed = []
    for i, p in enumerate(pcms):
        b = _packed(p, bps)
        if i % 2 == 0 and b.size:
            t = torch.from_numpy(b.copy()).pin_memory()
            packed.append(t.numpy())
        else:
            packed.append(b)
    with corpus.Corpus(gpu_lib, ch, rate, bps, level, fmt, devices=list(range(ndev)), chunk_blocks=3, **ov) as co:
        res, st = co.encode(packed, nsamples=[p.shape[0] for p in pcms])
        assert st.devices == ndev and st.kernel_launches > 0
        _check_against_single(gpu_lib, oracle, pcms, ch, bps, rate, level, fmt, res, **ov)
        for p, r in zip(pcms, res):
            if p.shape[0]:
                one = api.encode_batch(gpu_lib, p, rate, bps, level, **ov)
                assert r.data.tobytes() == one.payload and r.file_bytes() == one.file_bytes()


@pytest.mark.gpu
def test_one_long_stream_split_over_the_devices(gpu_lib, oracle):
    """ONE stream cut into frame ranges, every GPU of the box takes chunks, the host
    prefix-sums the offsets: identical to the serial stream (decodes bit-exactly, MD5 verified)."""
    import torch
    ndev = torch.cuda.device_count()
    ch, bps, rate, level = 2, 16, 44100, 8
    n = 4096 * 148 * 3 + 1234
    pcm = synth.synth_pcm(n, ch, bps, rate, seed=77)
    t = torch.from_numpy(pcm.astype(np.int16)).pin_memory()
    with corpus.Corpus(gpu_lib, ch, rate, bps, level, api.PCM_S16LE, devices=list(range(ndev)), longest=n,
                       chunk_blocks=148, threads_per_device=2) as co:
        out = torch.empty(co.max_encoded_size(n), dtype=torch.uint8).pin_memory()
        res, st = co.encode([t.numpy()], nsamples=[n], outs=[out.numpy()])
    r = res[0]
    assert st.chunks == 4 and sum(1 for d in range(ndev) if st.device_samples[d]) == min(ndev, 4)
    one = api.encode_batch(gpu_lib, pcm, rate, bps, level)
    assert r.file_bytes() == one.file_bytes()
    dec, info = oracle.decode(r.file_bytes())
    assert info.md5_ok == 1 and np.array_equal(dec, pcm)
