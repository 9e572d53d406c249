"""dev tool: the e2e corpus leg alone, with knobs, to find what bounds it.

  python tools/corpus_time.py [tracks] [seconds] [gpu_threads] [md5_threads] [steps] [chunk_blocks] [devices]

ONE process drives `devices` GPUs (library threads).  tracks = 1 with a long duration is "one stream
split by frame range over the GPUs of the box".
"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flake_b200 import api, corpus as fc, synth

tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 128
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 225
gth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mth = int(sys.argv[4]) if len(sys.argv) > 4 else 0
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cb = int(sys.argv[6]) if len(sys.argv) > 6 else 0
ndev = int(sys.argv[7]) if len(sys.argv) > 7 else 1
n = int(secs * 44100)
lib = api.load_library()
co = fc.Corpus(lib, 2, 44100, 16, 8, api.PCM_S16LE, devices=list(range(ndev)), longest=n, md5_threads=mth, threads_per_device=gth,
               chunk_blocks=cb)
cap, fcap = co.max_encoded_size(n), co.frame_cap(n)
h_in = torch.empty((tracks, n, 2), dtype=torch.int16).pin_memory()
h_out = torch.empty((tracks, cap), dtype=torch.uint8).pin_memory()
inn, outn = h_in.numpy(), h_out.numpy()
for j in range(tracks):
    synth.corpus_track(j, n, 2, 16, 44100, out=inn[j])
items = (api.FlakeB200CorpusStream * tracks)()
for j in range(tracks):
    it = items[j]
    it.pcm = inn[j].ctypes.data; it.nsamples = n; it.out = outn[j].ctypes.data; it.out_cap = cap
st = api.FlakeB200CorpusStats()
ts = []
for k in range(steps + 1):
    t0 = time.perf_counter()
    rc = lib.flake_b200_corpus_encode(co.handle, items, tracks, C.byref(st))
    ts.append((time.perf_counter() - t0) * 1e3)
    assert rc == 0, rc
ms = float(np.mean(ts[1:]))
print("devices %d " % ndev + " ".join("dev%d %.0f ms %.0f%%" % (d, st.device_ms[d], 100.0 * st.device_samples[d] / max(1, st.samples)) for d in range(ndev)))
print("tracks %d x %.0f s gpu_threads %d md5_threads %d lanes %d chunk_blocks %d chunks %d: %.1f ms/step = %.0f MSamples/s; "
      "gpu_worker %.1f ms md5 %.1f ms h2d %.1f GB/s" % (tracks, secs, st.gpu_threads, st.md5_threads, st.md5_lanes,
      st.chunk_blocks, st.chunks, ms, tracks * n / ms / 1e3, st.device_ms[0], st.md5_ms, st.h2d_bytes / ms / 1e6), flush=True)
co.close()
