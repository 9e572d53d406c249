"""dev tool: phase cycles of k_search (library built with -DFB_SEARCH_PROF, see tools/variants.py)."""
import ctypes as C, os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flake_b200 import api
lib = api.load_library()
lib.flake_b200_debug_search_prof.argtypes = [C.c_void_p, C.c_int]
import runpy
sys.argv = ["stage_time.py", "8", "880", "1"]
z = (C.c_ulonglong * 16)()
# run the timing tool in-process so that the same library instance accumulates
import importlib.util
spec = importlib.util.spec_from_file_location("st", os.path.join(os.path.dirname(__file__), "stage_time.py"))
lib.flake_b200_debug_search_prof(None, 1)
st = importlib.util.module_from_spec(spec); spec.loader.exec_module(st)
lib2 = st.lib
lib2.flake_b200_debug_search_prof.argtypes = [C.c_void_p, C.c_int]
lib2.flake_b200_debug_search_prof(z, 0)
v = np.array(list(z), dtype=np.float64)
ctas = max(1.0, v[15])
names = ["staging", "tiles", "wait A", "finish", "wait B", "plan", "replay", "final store", "store best"]
tot = v[:9].sum()
for i, nm in enumerate(names):
    print("%-12s %9.0f cycles per CTA  %5.1f%%" % (nm, v[i] / ctas, 100 * v[i] / tot))
print("total %.0f cycles per CTA over %d CTAs" % (tot / ctas, int(ctas)))
