"""dev tool: latency of the per-block API (flake_encode_frame, one block per synchronous call) and
the stage times of a one-block pass.

  python tools/per_block_time.py [level] [blocks]
"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flake_b200 import api, synth

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nblocks = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
lib = api.load_library()
B = 4096
pcm = np.ascontiguousarray(synth.synth_pcm(nblocks * B, 2, 16, 44100, seed=3), dtype=np.int32)
enc = api.Encoder(lib, 2, 44100, 16, nblocks * B, level)
enc.init()
ctx = C.byref(enc.ctx)
base, step = pcm.ctypes.data, B * 2 * 4
for b in range(50):
    lib.flake_encode_frame(ctx, base + b * step, B)
for rep in range(2):
    lib.flake_b200_reset_stream(ctx)
    t0 = time.perf_counter()
    nbytes = 0
    for b in range(nblocks):
        fs = lib.flake_encode_frame(ctx, base + b * step, B)
        assert fs > 0, fs
        nbytes += fs
    dt = time.perf_counter() - t0
    print("level %d: %.1f us per call, %.2f MSamples/s, %d bytes" % (level, dt / nblocks * 1e6, nblocks * B / dt / 1e6, nbytes), flush=True)
enc.set_profiling(True)
lib.flake_b200_reset_stream(ctx)
for b in range(200):
    lib.flake_encode_frame(ctx, base + b * step, B)
print({k: round(v[0] / 200 * 1e3, 1) for k, v in enc.stage_times().items()}, "us per block (events)")
enc.close()
