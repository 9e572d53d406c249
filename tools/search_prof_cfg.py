"""dev tool: phase cycles of k_search for a BASELINE configuration (library built with -DFB_SEARCH_PROF).

  FLAKE_B200_LIB=flake_b200/lib/var/prof/libflake.so python tools/search_prof_cfg.py C3
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from flake_b200 import api

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = bench.CONFIGS[name]
dev = torch.device("cuda", 0)
lib = api.load_library(); lib.flake_b200_set_device(0)
lib.flake_b200_debug_search_prof.argtypes = [C.c_void_p, C.c_int]
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
dp = bench.DevicePass(lib, cfg, bench.workload_pcm(cfg, 0), dev, st)
dp.run(); torch.cuda.synchronize()
lib.flake_b200_debug_search_prof(None, 1)
dp.run(); torch.cuda.synchronize()
z = (C.c_ulonglong * 16)()
lib.flake_b200_debug_search_prof(z, 0)
v = np.array(list(z), dtype=np.float64)
ctas = max(1.0, v[15])
names = ["staging", "tiles", "wait A", "finish", "wait B", "plan", "replay", "final store", "store best"]
tot = v[:9].sum()
for i, nm in enumerate(names):
    print("%-12s %9.0f cycles per CTA  %5.1f%%" % (nm, v[i] / ctas, 100 * v[i] / tot))
print("%s: total %.0f cycles per CTA over %d CTAs" % (name, tot / ctas, int(ctas)))
