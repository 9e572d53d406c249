"""dev tool: time flake_b200_encode_stream on synthetic PCM (host buffers)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flake_b200 import api, synth
lib = api.load_library()
level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 60
base = synth.synth_pcm(int(44100 * 20), 2, 16, 44100, seed=1)
reps = max(1, int(secs / 20))
pcm = np.ascontiguousarray(np.tile(base, (reps, 1)))
n = pcm.shape[0] // 4096 * 4096
pcm = pcm[:n]
enc = api.Encoder(lib, 2, 44100, 16, n, level)
enc.init()
for it in range(3):
    t = time.time()
    lib.flake_b200_seek(enc.ctx, 0)
    data, flen, fbs = enc.encode_stream(pcm, api.PCM_S32, n, want_sizes=True)
    dt = time.time() - t
    st = enc.stats()
    print("level %d: %d samples in %.3fs = %.1f MSamples/s, %d bytes, ratio %.3f, gpu_ms %.1f md5_ms %.1f launches %d" % (
        level, n, dt, n / dt / 1e6, len(data), len(data) / (n * 4), st.gpu_ms, st.md5_ms, st.kernel_launches))
enc.close()
