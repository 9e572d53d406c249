"""dev tool: per-stage device timing of the C2 pass (packed s16 resident in HBM) + an output
checksum, so library variants (FLAKE_B200_LIB) can be compared for speed AND identical bytes.

  python tools/stage_time.py [level] [seconds_of_audio] [passes]
"""
import ctypes as C, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from flake_b200 import api

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3600
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n = int(secs * 44100)
dev = torch.device("cuda", 0)
lib = api.load_library()
lib.flake_b200_set_device(0)
ov = {'variable_block_size': 0, 'allow_vbs': 0} if os.environ.get('FLAKE_TEST_NO_VBS') else {}
enc = api.Encoder(lib, 2, 44100, 16, n, level, **ov)
enc.init()
ctx = C.byref(enc.ctx)
d_pcm = torch.from_numpy(bench.workload_pcm(bench.C2, 0, n).astype(np.int16)).to(dev)
cs, cb, cf = C.c_ulonglong(), C.c_ulonglong(), C.c_uint()
lib.flake_b200_device_capacity(ctx, C.byref(cs), C.byref(cb), C.byref(cf))
chunk = int(cs.value); nch = (n + chunk - 1) // chunk
d_out = [torch.zeros(int(cb.value), dtype=torch.uint8, device=dev) for _ in range(nch)]
d_flen = torch.empty(int(cf.value), dtype=torch.int32, device=dev)
d_sum = torch.zeros((nch, 3), dtype=torch.int64, device=dev)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
def run():
    for k in range(nch):
        s0 = k * chunk; ns = min(chunk, n - s0)
        rc = lib.flake_b200_encode_device(ctx, d_pcm[s0:].data_ptr(), api.PCM_S16LE, ns, s0 // 4096,
                                          d_out[k].data_ptr(), d_flen.data_ptr(), None, d_sum[k].data_ptr(), st.cuda_stream)
        assert rc == 0, rc
for _ in range(3): run()
torch.cuda.synchronize()
enc.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(passes): run()
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / passes
stages = {k: round(v[0] / passes, 3) for k, v in enc.stage_times().items()}
h = hashlib.md5()
tot = 0
summ = d_sum.cpu().numpy()
for k in range(nch):
    nb = int(summ[k].view(np.uint64)[1]); tot += nb
    h.update(d_out[k][:nb].cpu().numpy().tobytes())
print("%s level %d: %.1f MSamples/s %.3f ms/pass chunk_blocks %d %s bytes %d md5 %s" % (
    os.environ.get("FLAKE_B200_LIB", "default"), level, n / ms / 1e3, ms, chunk // 4096, stages, tot, h.hexdigest()), flush=True)
