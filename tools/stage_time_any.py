"""dev tool: per-stage device timing for any (level, channels, bps, rate) on int32 device PCM.

  python tools/stage_time_any.py level channels bps rate seconds [passes]
"""
import ctypes as C, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flake_b200 import api

level, ch, bps, rate = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
secs = float(sys.argv[5]); passes = int(sys.argv[6]) if len(sys.argv) > 6 else 3
n = int(secs * rate)
dev = torch.device("cuda", 0)
lib = api.load_library(); lib.flake_b200_set_device(0)
enc = api.Encoder(lib, ch, rate, bps, n, level); enc.init()
B = enc.ctx.params.block_size
ctx = C.byref(enc.ctx)
g = torch.Generator(device=dev); g.manual_seed(1234)
full = float(2 ** (bps - 1) - 1)
t = torch.arange(n, device=dev, dtype=torch.float64) / rate
pcm = torch.empty((n, ch), dtype=torch.int32, device=dev)
for c in range(ch):
    x = 0.3 * torch.sin(2 * np.pi * (220.0 * (c + 1)) * t + c) + 0.2 * torch.sin(2 * np.pi * 1733.0 * t * (1 + 0.01 * c))
    x = x * (0.6 + 0.4 * torch.sin(2 * np.pi * 0.3 * t))
    x += 0.01 * torch.randn(n, generator=g, device=dev, dtype=torch.float32).double()
    imp = (torch.rand(n, generator=g, device=dev) < 1.0 / 30000.0).double()
    x += imp * 0.7
    pcm[:, c] = torch.clamp(torch.round(x * full), -full - 1, full).to(torch.int32)
cs, cb, cf = C.c_ulonglong(), C.c_ulonglong(), C.c_uint()
lib.flake_b200_device_capacity(ctx, C.byref(cs), C.byref(cb), C.byref(cf))
chunk = int(cs.value); nch = (n + chunk - 1) // chunk
d_out = torch.zeros(int(cb.value), dtype=torch.uint8, device=dev)
d_flen = torch.empty(int(cf.value), dtype=torch.int32, device=dev)
d_sum = torch.zeros((nch, 3), dtype=torch.int64, device=dev)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
def run():
    for k in range(nch):
        s0 = k * chunk; ns = min(chunk, n - s0)
        first = s0 if enc.ctx.params.allow_vbs else s0 // B
        rc = lib.flake_b200_encode_device(ctx, pcm[s0:].data_ptr(), api.PCM_S32, ns, first,
                                          d_out.data_ptr(), d_flen.data_ptr(), None, d_sum[k].data_ptr(), st.cuda_stream)
        assert rc == 0, rc
for _ in range(2): run()
torch.cuda.synchronize()
enc.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(passes): run()
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / passes
stages = {k: round(v[0] / passes, 3) for k, v in enc.stage_times().items()}
summ = d_sum.cpu().numpy()
tot = sum(int(summ[k].view(np.uint64)[1]) for k in range(nch))
frames = sum(int(summ[k].view(np.uint32)[0]) for k in range(nch))
print("level %d %dch %dbit %dHz block %d: %.1f MSamples/s (%.1f M channel-samples/s) %.3f ms/pass chunks %d %s frames %d ratio %.3f" % (
    level, ch, bps, rate, B, n / ms / 1e3, n * ch / ms / 1e3, ms, nch, stages, frames, tot / (n * ch * ((bps + 7) // 8))), flush=True)
