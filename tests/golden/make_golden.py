#!/usr/bin/env python
"""Generate the golden vectors in this directory FROM THE COMPILED REFERENCE
(oracle/_ref/libflake_ref.so, built by `make -C oracle ref` out of the read-only
/root/reference tree).  The reference ships no known-answer tests of its own
(SURVEY.md section 4), so these fixtures are its outputs, captured here, on seeded
synthetic inputs; they travel to the GPU box where /root/reference does not exist.

  streams.json      per case: parameters, sha256 of the input PCM, sha256 + length of the
                    complete .flac the flake/flake.c loop leaves on disk, per-frame byte
                    lengths, final STREAMINFO hex
  small/<case>.flac the complete files of the small cases
  stages.json       stage-level vectors from the reference's internal (non-static)
                    functions: lpc_calc_coefs (lpc.c:224), find_optimal_rice_param
                    (rice.c:30), calc_rice_params_lpc/_fixed (rice.c:173-187), calc_crc8/16

Run:  python tests/golden/make_golden.py        (needs oracle/_ref)
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from flake_b200 import api, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

# (name, nsamples, channels, bps, rate, kind, seed, level, overrides, keep_file)
STREAM_CASES = [
    ("c1_l5_s16_stereo", 4096 * 24 + 3936, 2, 16, 44100, "mix", 1, 5, {}, False),
    ("c2_l8_s16_stereo", 4096 * 24 + 3136, 2, 16, 44100, "mix", 2, 8, {}, False),
    ("c3_l12_s24_96k", 8192 * 10 + 2048, 2, 24, 96000, "impulses", 3, 12, {}, False),
    ("c4_l9_8ch_s24", 4096 * 8 + 1024, 8, 24, 48000, "impulses", 4, 9, {}, False),
    ("c4_l10_8ch_s24", 4096 * 6 + 1024, 8, 24, 48000, "impulses", 5, 10, {}, False),
    ("l0_small", 1152 * 3 + 100, 2, 16, 44100, "mix", 10, 0, {}, True),
    ("l1_small", 1152 * 3 + 100, 2, 16, 44100, "mix", 11, 1, {}, True),
    ("l2_small", 1152 * 3 + 100, 2, 16, 44100, "mix", 12, 2, {}, True),
    ("l3_small", 4096 * 2 + 500, 2, 16, 44100, "mix", 13, 3, {}, True),
    ("l4_small", 4096 * 2 + 500, 2, 16, 44100, "mix", 14, 4, {}, False),
    ("l6_small", 4096 * 2 + 500, 2, 16, 44100, "mix", 16, 6, {}, False),
    ("l7_small", 4096 * 2 + 500, 2, 16, 44100, "mix", 17, 7, {}, True),
    ("l8_small", 4096 * 2 + 500, 2, 16, 44100, "mix", 18, 8, {}, True),
    ("l9_small_vbs", 4096 * 3, 2, 16, 44100, "impulses", 19, 9, {}, True),
    ("l11_small", 8192 * 2, 2, 16, 44100, "impulses", 21, 11, {}, False),
    ("noise_l0_verbatim", 1152 * 4, 2, 16, 44100, "noise", 30, 0, {}, True),
    ("noise_l8", 4096 * 2, 2, 16, 44100, "noise", 31, 8, {}, False),
    ("noise_s24_l8_rice2", 4096 * 2, 2, 24, 96000, "noise", 32, 8, {}, False),
    ("silence_l5", 4096 * 2, 2, 16, 44100, "silence", 33, 5, {}, True),
    ("wasted_l5", 4096 * 2, 2, 16, 44100, "wasted", 34, 5, {}, True),
    ("mono_l8", 4096 * 2 + 10, 1, 16, 44100, "mix", 35, 8, {}, False),
    ("tail7_l8", 4096 + 7, 2, 16, 44100, "mix", 36, 8, {}, True),
    ("tail3_l8", 4096 + 3, 2, 16, 44100, "mix", 37, 8, {}, True),
    ("u8_l5", 4096 * 2, 2, 8, 22050, "mix", 38, 5, {}, False),
    ("rate37800_l8", 4096 * 2, 2, 16, 37800, "mix", 39, 8, {}, False),
    ("bs1000_l5", 1000 * 4 + 10, 2, 16, 44100, "mix", 40, 5, {"block_size": 1000}, False),
    ("l8_order_max", 4096 * 2, 2, 16, 44100, "mix", 41, 8, {"order_method": 0}, False),
    ("l8_2level", 4096 * 2, 2, 16, 44100, "mix", 42, 8, {"order_method": 2}, False),
    ("l8_8level", 4096 * 2, 2, 16, 44100, "mix", 43, 8, {"order_method": 4}, False),
    ("l8_search", 4096 * 2, 2, 16, 44100, "mix", 44, 8, {"order_method": 5}, False),
    ("l9_allow_vbs_only", 4096 * 2, 2, 16, 44100, "mix", 45, 9, {"variable_block_size": 0}, False),
    ("l5_independent", 4096 * 2, 2, 16, 44100, "mix", 46, 5, {"stereo_method": 0}, False),
    ("l5_nopadding", 4096, 2, 16, 44100, "mix", 47, 5, {"padding_size": 0}, True),
]


def vbs_burst_pcm(n, ch, bps, seed, blk=4096):
    """quiet first half / loud second half inside each block: forces vbs.c:65-72 to split"""
    rng = np.random.Generator(np.random.PCG64(seed))
    full = (1 << (bps - 1)) - 1
    x = rng.standard_normal((n, ch)) * 0.0005 * full
    for b in range(0, n, blk):
        cut = b + int(rng.integers(1, 8)) * (blk // 8)
        x[cut:b + blk] = rng.standard_normal((min(n, b + blk) - cut, ch)) * 0.2 * full
    return np.clip(np.rint(x), -full - 1, full).astype(np.int32)


def main():
    ref = po.ref_library()
    os.makedirs(os.path.join(HERE, "small"), exist_ok=True)
    streams = []
    cases = list(STREAM_CASES)
    for name, n, ch, bps, rate, kind, seed, level, ov, keep in cases:
        pcm = synth.synth_pcm(n, ch, bps, rate, seed=seed, kind=kind)
        streams.append(run_case(ref, name, pcm, rate, bps, level, ov, keep,
                                {"gen": "synth", "kind": kind, "seed": seed}))
    for name, ch, bps, rate, level in [("vbs_split_l9", 2, 16, 44100, 9), ("vbs_split_l12_s24", 2, 24, 96000, 12),
                                       ("vbs_split_l10_8ch", 8, 24, 48000, 10)]:
        pcm = vbs_burst_pcm(4096 * 6, ch, bps, seed=99)
        streams.append(run_case(ref, name, pcm, rate, bps, level, {}, name == "vbs_split_l9",
                                {"gen": "vbs_burst", "seed": 99}))
    with open(os.path.join(HERE, "streams.json"), "w") as f:
        json.dump(streams, f, indent=1)
    stages = make_stages()
    with open(os.path.join(HERE, "stages.json"), "w") as f:
        json.dump(stages, f, indent=1)
    print("wrote %d stream cases, %d stage vectors" % (len(streams), sum(len(v) for v in stages.values())))


def run_case(ref, name, pcm, rate, bps, level, ov, keep, gen):
    r = api.encode_per_block(ref, pcm, rate, bps, level, **ov)
    data = r.file_bytes()
    # frame lengths per API call are not per frame under VBS; recover per-frame lengths with the decoder
    dec, info = po.decode(data)
    assert np.array_equal(dec, pcm) and info.md5_ok == 1, name
    rec = {
        "name": name, "nsamples": int(pcm.shape[0]), "channels": int(pcm.shape[1]), "bps": bps,
        "rate": rate, "level": level, "overrides": ov, "input": gen,
        "pcm_sha256": hashlib.sha256(pcm.tobytes()).hexdigest(),
        "flac_sha256": hashlib.sha256(data).hexdigest(), "flac_len": len(data),
        "header_len": len(r.header), "call_lens": [len(f) for f in r.frames],
        "nframes": int(info.nframes), "streaminfo": r.streaminfo.hex(),
    }
    if keep:
        with open(os.path.join(HERE, "small", name + ".flac"), "wb") as f:
            f.write(data)
        rec["file"] = "small/%s.flac" % name
    print("%-24s %8d bytes %4d frames" % (name, len(data), info.nframes))
    return rec


def make_stages():
    L = C.CDLL(po.REF_SO)
    L.crc_init()
    out = {"lpc": [], "rice_k": [], "rice_cost": [], "crc": [], "vbs": []}

    # lpc_calc_coefs(samples, blocksize, max_order, precision, omethod, coefs[][32], shift[])
    for (n, ch_kind, seed, max_order, om) in [(4096, "mix", 1, 8, 1), (4096, "mix", 2, 12, 6), (8192, "mix", 3, 32, 5),
                                              (4096, "noise", 4, 12, 6), (1000, "mix", 5, 8, 0), (4096, "sine", 6, 12, 6),
                                              (3136, "mix", 7, 12, 6), (512, "impulses", 8, 32, 5)]:
        pcm = synth.synth_pcm(n, 1, 16, 44100, seed=seed, kind=ch_kind)[:, 0].copy()
        coefs = (C.c_int32 * (32 * 32))()
        shift = (C.c_int * 32)()
        L.lpc_calc_coefs.restype = C.c_int
        est = L.lpc_calc_coefs(pcm.ctypes.data_as(C.c_void_p), n, max_order, 15, om, coefs, shift)
        rows = [max_order - 1] if om == 0 else ([est - 1] if om == 1 else list(range(max_order)))
        out["lpc"].append({
            "n": n, "kind": ch_kind, "seed": seed, "max_order": max_order, "omethod": om, "est": int(est),
            "pcm_sha256": hashlib.sha256(pcm.tobytes()).hexdigest(),
            "rows": {str(i): {"shift": int(shift[i]), "coefs": [int(coefs[i * 32 + j]) for j in range(i + 1)]}
                     for i in rows}})

    L.find_optimal_rice_param.argtypes = [C.c_uint64, C.c_int]
    L.find_optimal_rice_param.restype = C.c_int
    rng = np.random.Generator(np.random.PCG64(7))
    for _ in range(400):
        n = int(rng.choice([0, 1, 4, 16, 52, 64, 256, 4084, 4096, 65535]))
        e = float(rng.uniform(0, 44))
        s = int(2 ** e) + int(rng.integers(0, 1000)) - 500
        s = max(0, s)
        if n == 0:
            s = 0
        out["rice_k"].append([s, n, int(L.find_optimal_rice_param(s, n))])

    class RiceContext(C.Structure):
        _fields_ = [("method", C.c_int), ("porder", C.c_int), ("params", C.c_int * 256), ("esc_bps", C.c_int * 256)]
    L.calc_rice_params_lpc.restype = C.c_uint32
    L.calc_rice_params_fixed.restype = C.c_uint32
    for (n, scale, order, pmin, pmax, lpc, seed) in [(4096, 50.0, 8, 0, 6, 1, 1), (4096, 3000.0, 12, 0, 8, 1, 2),
                                                     (4608, 200.0, 2, 0, 8, 0, 3), (1152, 10.0, 4, 0, 3, 0, 4),
                                                     (8192, 1e6, 32, 0, 8, 1, 5), (4096, 0.2, 1, 0, 5, 1, 6),
                                                     (3136, 70.0, 12, 0, 6, 1, 7), (16, 9.0, 2, 0, 4, 0, 8)]:
        rg = np.random.Generator(np.random.PCG64(1000 + seed))
        res = np.rint(rg.laplace(0.0, scale, size=n)).astype(np.int32)
        rc = RiceContext()
        if lpc:
            bits = L.calc_rice_params_lpc(C.byref(rc), pmin, pmax, res.ctypes.data_as(C.c_void_p), n, order, 16, 15)
        else:
            bits = L.calc_rice_params_fixed(C.byref(rc), pmin, pmax, res.ctypes.data_as(C.c_void_p), n, order, 16)
        out["rice_cost"].append({"n": n, "scale": scale, "order": order, "pmin": pmin, "pmax": pmax, "lpc": lpc,
                                 "seed": seed, "res_sha256": hashlib.sha256(res.tobytes()).hexdigest(),
                                 "bits": int(bits), "method": rc.method, "porder": rc.porder,
                                 "params": [int(rc.params[i]) for i in range(1 << rc.porder)]})

    L.calc_crc8.restype = C.c_uint8
    L.calc_crc16.restype = C.c_uint16
    for ln in [0, 1, 2, 3, 5, 16, 255, 4097]:
        d = bytes(np.random.Generator(np.random.PCG64(ln)).integers(0, 256, size=ln, dtype=np.uint8))
        out["crc"].append({"len": ln, "seed": ln, "crc8": int(L.calc_crc8(d, ln)) if ln else 0,
                           "crc16": int(L.calc_crc16(d, ln)) if ln else 0})
    return out


if __name__ == "__main__":
    main()
