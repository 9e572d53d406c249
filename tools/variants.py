"""dev tool: build library variants with extra -D flags for A/B timing on the GPU box.

  python tools/variants.py name1:DEF1=V,DEF2=V name2:...   (built into flake_b200/lib/var/<name>/)
Then on the GPU:  FLAKE_B200_LIB=flake_b200/lib/var/<name>/libflake.so python tools/stage_time.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flake_b200 import build as b
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    print(b.build_product(force=True, variant=name, defines=[d for d in defs.split(",") if d]))
