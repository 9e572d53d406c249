"""Deterministic synthetic PCM for parity tests and benchmarks.

Signal recipe follows SURVEY.md section 8(d): a few sinusoids (one slow chirp)
under a slow envelope, low-passed Gaussian noise bursts, a white noise floor,
sparse filtered impulses (these trigger VBS splits, vbs.c:65-72), stretches of
digital silence and DC (CONSTANT subframes, optimize.c:143-151) and level steps
(wasted bits / low Rice parameters).  The right channel of a stereo pair is a
phase-shifted, attenuated copy of the left plus independent noise so that all
four stereo modes occur (encode.c:598-643).

The seed convention is ``0xF1A4E000 + file_index``.  Everything is plain numpy so
the same bytes are produced wherever this image runs.
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0xF1A4E000


def _lowpass(x: np.ndarray, a: float) -> np.ndarray:
    """One-pole low-pass via cumulative trick on short blocks (vectorised)."""
    # y[n] = (1-a) x[n] + a y[n-1]; do it with an FIR truncation, good enough
    taps = int(min(64, max(4, np.ceil(np.log(1e-3) / np.log(max(a, 1e-6))))))
    h = (1.0 - a) * a ** np.arange(taps)
    return np.convolve(x, h, mode="full")[: len(x)]


def synth_pcm(nsamples: int, channels: int = 2, bps: int = 16, sample_rate: int = 44100,
              seed: int = 0, kind: str = "mix") -> np.ndarray:
    """Return interleaved int32 PCM of shape (nsamples, channels), sign-extended to bps.

    kind: "mix" (default recipe), "noise" (full-scale white noise),
          "sine" (pure tone), "silence", "wasted" (multiples of 16), "impulses".
    """
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + seed))
    full = float((1 << (bps - 1)) - 1)
    n = int(nsamples)
    t = np.arange(n, dtype=np.float64) / float(sample_rate)
    out = np.zeros((n, channels), dtype=np.float64)

    if kind == "noise":
        out = rng.uniform(-full, full, size=(n, channels))
    elif kind == "silence":
        pass
    elif kind == "sine":
        for c in range(channels):
            out[:, c] = 0.6 * full * np.sin(2 * np.pi * (440.0 + 110.0 * c) * t + 0.3 * c)
    else:
        base = None
        for c in range(channels):
            if c % 2 == 1 and base is not None:
                # correlated copy of the previous channel
                shift = int(rng.integers(1, 40))
                x = 0.8 * np.roll(base, shift) + 0.02 * full * _lowpass(rng.standard_normal(n), 0.6)
            else:
                x = np.zeros(n)
                nt = int(rng.integers(3, 7))
                for k in range(nt):
                    f0 = float(rng.uniform(60.0, 5000.0))
                    amp = float(rng.uniform(0.03, 0.25))
                    ph = float(rng.uniform(0, 2 * np.pi))
                    if k == 0:   # slow chirp
                        f1 = f0 * float(rng.uniform(1.2, 2.5))
                        dur = max(t[-1], 1e-3) if n > 1 else 1.0
                        phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / dur)
                    else:
                        phase = 2 * np.pi * f0 * t
                    x += amp * np.sin(phase + ph)
                env = 0.55 + 0.45 * np.sin(2 * np.pi * float(rng.uniform(0.05, 0.4)) * t
                                           + float(rng.uniform(0, 6.28)))
                x *= env * full
                # shaped noise bursts at about -30 dBFS in random 4096-sample windows
                burst = np.zeros(n)
                nb = max(1, n // 32768)
                for s in rng.integers(0, max(1, n - 4096), size=nb):
                    burst[s:s + 4096] = 1.0
                x += burst * 0.03 * full * _lowpass(rng.standard_normal(n), 0.85)
                # white floor at about -50 dBFS
                x += 0.003 * full * rng.standard_normal(n)
                base = x
            # sparse filtered impulses
            ni = max(1, n // 20000) if kind in ("mix", "impulses") else 0
            if ni:
                imp = np.zeros(n)
                idx = rng.integers(0, n, size=ni)
                imp[idx] = rng.uniform(-0.7, 0.7, size=ni) * full
                x = x + _lowpass(imp, 0.7) * (4.0 if kind == "impulses" else 1.0)
            out[:, c] = x
        # silence, DC, level-step stretches (common to all channels)
        if n >= 65536:
            seg = max(4096 * 3, n // 40)
            a = int(rng.integers(0, n - seg)); out[a:a + seg, :] = 0.0
            b = int(rng.integers(0, n - seg)); out[b:b + seg, :] = np.round(0.1 * full)
            c0 = int(rng.integers(0, n - seg)); out[c0:c0 + seg, :] *= 0.05
        if kind == "wasted":
            out = np.round(out / 16.0) * 16.0

    q = np.clip(np.rint(out), -full - 1, full).astype(np.int32)
    if kind == "wasted":
        q &= ~np.int32(15)
    return np.ascontiguousarray(q)


def long_pcm(nsamples: int, channels: int, bps: int, sample_rate: int, seed: int = 0,
             base_seconds: float = 40.0) -> np.ndarray:
    """A long stream (minutes to hours) assembled from ONE `base_seconds` synthetic segment
    (synth_pcm, "mix") repeated with a different gain per tile, so that generating an hour of
    audio costs seconds of CPU.  Every tile is distinct data (the gain changes every sample
    value and tile boundaries do not fall on block boundaries).  This is the generator of the
    benchmark workloads: the GPU arm, the reference arm and the full-size parity tests all
    call it, so that they encode the same samples."""
    n0 = int(round(sample_rate * base_seconds))       # always the whole segment: a shorter request is a prefix
    base = synth_pcm(n0, channels, bps, sample_rate, seed=seed).astype(np.float64)
    reps = (nsamples + n0 - 1) // n0
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + 7919 * (seed + 1)))
    out = np.empty((nsamples, channels), dtype=np.int32)
    for r in range(reps):
        g = float(rng.uniform(0.3, 1.0))
        a, b = r * n0, min(nsamples, (r + 1) * n0)
        out[a:b] = np.rint(base[:b - a] * g).astype(np.int32)
    return out


_TRACK_BASES: dict = {}


def corpus_track(index: int, nsamples: int, channels: int, bps: int, sample_rate: int,
                 out: np.ndarray = None, nbases: int = 8, base_seconds: float = 40.0) -> np.ndarray:
    """Track `index` of a synthetic corpus (many files of one format): like long_pcm, tiles of a
    `base_seconds` "mix" segment with a different gain per tile, but the segment is one of
    `nbases` cached ones, entered at a track-specific offset, so that a corpus of hundreds of
    tracks costs seconds to generate and no two tracks hold the same samples.  Writes into `out`
    (any integer dtype that holds `bps` bits, shape (nsamples, channels)) when given."""
    key = (index % nbases, channels, bps, sample_rate, base_seconds)
    if key not in _TRACK_BASES:
        n0 = int(round(sample_rate * base_seconds))
        _TRACK_BASES[key] = synth_pcm(n0, channels, bps, sample_rate, seed=1000 + key[0]).astype(np.float32)
    base = _TRACK_BASES[key]
    n0 = base.shape[0]
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + 104729 * (index + 1)))
    if out is None:
        out = np.empty((nsamples, channels), dtype=np.int32)
    pos, off = 0, int(rng.integers(0, n0))
    while pos < nsamples:
        g = np.float32(rng.uniform(0.3, 1.0))
        take = min(nsamples - pos, n0 - off)
        out[pos:pos + take] = np.rint(base[off:off + take] * g)
        pos += take
        off = 0
    return out


def pack_pcm(pcm: np.ndarray, bps: int) -> bytes:
    """Little-endian packed sample bytes at ceil(bps/8) bytes per sample (WAV data chunk)."""
    nbytes = (bps + 7) // 8
    flat = pcm.reshape(-1).astype("<i4")
    raw = flat.view(np.uint8).reshape(-1, 4)[:, :nbytes]
    return np.ascontiguousarray(raw).tobytes()


def wav_bytes(pcm: np.ndarray, bps: int, sample_rate: int) -> bytes:
    """Plain PCM WAV (format tag 1) readable by the reference CLI (libpcm_io/wav.c)."""
    import struct
    n, ch = pcm.shape
    data = pack_pcm(pcm, bps)
    bytes_per = (bps + 7) // 8
    fmt = struct.pack("<HHIIHH", 1, ch, sample_rate, sample_rate * ch * bytes_per,
                      ch * bytes_per, bps)
    return (b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(data)) + b"WAVE" +
            b"fmt " + struct.pack("<I", len(fmt)) + fmt +
            b"data" + struct.pack("<I", len(data)) + data)
