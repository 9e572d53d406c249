"""The oracle restatement (oracle/flake_oracle.c) against the golden vectors captured
from the compiled reference (tests/golden/, made by tests/golden/make_golden.py).
Runs without a GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

from flake_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

with open(os.path.join(GOLD, "streams.json")) as f:
    STREAMS = json.load(f)
with open(os.path.join(GOLD, "stages.json")) as f:
    STAGES = json.load(f)


def golden_pcm(rec):
    import importlib.util
    g = rec["input"]
    if g["gen"] == "synth":
        pcm = synth.synth_pcm(rec["nsamples"], rec["channels"], rec["bps"], rec["rate"],
                              seed=g["seed"], kind=g["kind"])
    else:
        spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
        mg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mg)
        pcm = mg.vbs_burst_pcm(rec["nsamples"], rec["channels"], rec["bps"], g["seed"])
    assert hashlib.sha256(pcm.tobytes()).hexdigest() == rec["pcm_sha256"], \
        "synthetic generator drifted: regenerate the golden vectors"
    return pcm


def oracle_file(oracle, rec, pcm):
    ov = rec["overrides"]
    body, flen, fbs, mx = oracle.encode_stream(pcm, rec["rate"], rec["bps"], rec["level"], **ov)
    p = oracle.make_params(rec["channels"], rec["rate"], rec["bps"], rec["level"], rec["nsamples"], **ov)
    hdr = bytearray(oracle.header(p))
    si = oracle.streaminfo(p, mx, oracle.md5_pcm(pcm, rec["bps"]))
    hdr[8:42] = si
    return bytes(hdr) + body, flen, si


@pytest.mark.parametrize("rec", STREAMS, ids=[r["name"] for r in STREAMS])
def test_stream_matches_reference_output(rec, oracle):
    pcm = golden_pcm(rec)
    data, flen, si = oracle_file(oracle, rec, pcm)
    assert len(flen) == rec["nframes"]
    assert si.hex() == rec["streaminfo"]
    assert len(data) == rec["flac_len"]
    assert hashlib.sha256(data).hexdigest() == rec["flac_sha256"]
    if "file" in rec:
        with open(os.path.join(GOLD, rec["file"]), "rb") as f:
            assert f.read() == data


@pytest.mark.parametrize("rec", [r for r in STREAMS if "file" in r], ids=lambda r: r["name"])
def test_golden_files_decode(rec, oracle):
    """The test decoder accepts the reference's own files (CRC-8, CRC-16, MD5) and gets the PCM back."""
    pcm = golden_pcm(rec)
    with open(os.path.join(GOLD, rec["file"]), "rb") as f:
        data = f.read()
    dec, info = oracle.decode(data)
    assert info.md5_ok == 1 and info.nframes == rec["nframes"]
    assert np.array_equal(dec, pcm)
    # a flipped bit must be caught by a CRC
    bad = bytearray(data)
    bad[min(rec["header_len"] + 40, len(bad) - 3)] ^= 0x10
    with pytest.raises(ValueError):
        oracle.decode(bytes(bad))


@pytest.mark.parametrize("rec", STAGES["lpc"], ids=lambda r: "n%d_o%d_m%d" % (r["n"], r["max_order"], r["omethod"]))
def test_lpc_stage(rec, oracle):
    import ctypes as C
    pcm = synth.synth_pcm(rec["n"], 1, 16, 44100, seed=rec["seed"], kind=rec["kind"])[:, 0].copy()
    assert hashlib.sha256(pcm.tobytes()).hexdigest() == rec["pcm_sha256"]
    coefs = np.zeros((32, 32), dtype=np.int32)
    shift = np.zeros(32, dtype=np.int32)
    est = oracle.lib().orc_lpc_calc(pcm.ctypes.data, rec["n"], rec["max_order"], rec["omethod"],
                                    coefs.ctypes.data, shift.ctypes.data)
    assert est == rec["est"]
    for i, row in rec["rows"].items():
        i = int(i)
        assert int(shift[i]) == row["shift"]
        assert coefs[i, :i + 1].tolist() == row["coefs"]


def test_rice_parameter_search(oracle):
    L = oracle.lib()
    for s, n, k in STAGES["rice_k"]:
        assert L.orc_rice_k(s, n) == k, (s, n)


@pytest.mark.parametrize("rec", STAGES["rice_cost"], ids=lambda r: "n%d_o%d" % (r["n"], r["order"]))
def test_rice_cost_stage(rec, oracle):
    import ctypes as C
    rg = np.random.Generator(np.random.PCG64(1000 + rec["seed"]))
    res = np.rint(rg.laplace(0.0, rec["scale"], size=rec["n"])).astype(np.int32)
    assert hashlib.sha256(res.tobytes()).hexdigest() == rec["res_sha256"]
    method, porder = C.c_int(), C.c_int()
    params = (C.c_int * 256)()
    bits = oracle.lib().orc_rice_cost(res.ctypes.data, rec["n"], rec["order"], 16, rec["pmin"], rec["pmax"],
                                      rec["lpc"], C.byref(method), C.byref(porder), params)
    assert (bits, method.value, porder.value) == (rec["bits"], rec["method"], rec["porder"])
    assert list(params[:1 << porder.value]) == rec["params"]


def test_crc_vectors(oracle):
    L = oracle.lib()
    for rec in STAGES["crc"]:
        d = bytes(np.random.Generator(np.random.PCG64(rec["seed"])).integers(0, 256, size=rec["len"], dtype=np.uint8))
        assert L.orc_crc8(d, len(d)) == rec["crc8"]
        assert L.orc_crc16(d, len(d)) == rec["crc16"]


def test_md5_known_answers(oracle):
    # RFC 1321 test suite through the PCM packer: 8-bit mono samples are the bytes themselves
    for msg, want in [(b"", "d41d8cd98f00b204e9800998ecf8427e"), (b"abc", "900150983cd24fb0d6963f7d28e17f72"),
                      (b"message digest", "f96b697d7cb7938d525a2f31aaf161d0")]:
        pcm = np.frombuffer(msg, dtype=np.int8).astype(np.int32).reshape(-1, 1)
        assert oracle.md5_pcm(pcm, 8).hex() == want
    # 16-bit stereo equals md5 of the little-endian WAV data chunk
    pcm = synth.synth_pcm(1000, 2, 16, 44100, seed=5)
    assert oracle.md5_pcm(pcm, 16).hex() == hashlib.md5(synth.pack_pcm(pcm, 16)).hexdigest()
    pcm = synth.synth_pcm(777, 2, 24, 96000, seed=6)
    assert oracle.md5_pcm(pcm, 24).hex() == hashlib.md5(synth.pack_pcm(pcm, 24)).hexdigest()
