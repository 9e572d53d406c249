"""dev tool: device-resident pass of a BASELINE.json configuration (bench.CONFIGS: C1..C4) in its
packed layout, per-stage CUDA-event times.  The command the ncu captures of profiles/ run.

  python tools/stage_time_cfg.py C3 [passes] [warmup]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from flake_b200 import api

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = bench.CONFIGS[name]
dev = torch.device("cuda", 0)
lib = api.load_library(); lib.flake_b200_set_device(0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
dp = bench.DevicePass(lib, cfg, bench.workload_pcm(cfg, 0), dev, st)
for _ in range(warm):
    dp.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(passes):
    dp.run()
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / passes
stages = {k: round(v[0] / max(1, v[1]) * dp.nchunks, 3) for k, v in dp.stage_profile(1).items()}
frames, out_bytes = dp.totals()
print("%s: %.1f MSamples/s %.3f ms/pass chunks %d launches/pass %d %s frames %d bytes %d" % (
    name, dp.n / ms / 1e3, ms, dp.nchunks, 0, stages, frames, out_bytes), flush=True)
