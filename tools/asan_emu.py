"""dev tool: kernel logic under AddressSanitizer (compute-sanitizer is not available on the GPU pool).

Builds the fiber-emulated library (tests/cuda_emu) with -fsanitize=address into /tmp/asan and
encodes a spread of configurations, comparing with the oracle.  Run as
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_emu.py
Out-of-bounds accesses to "shared" (static) and "global" (heap) arrays abort with a report.
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
INC, CS, EMU, OUT = [os.path.join(ROOT, p) for p in ("include", "flake_b200/csrc", "tests/cuda_emu")] + ["/tmp/asan"]
os.makedirs(OUT, exist_ok=True)
common = ["-O1", "-g", "-fPIC", "-fsanitize=address", "-fno-omit-frame-pointer", "-DFLAKE_B200_CUDA_EMU",
          "-I", EMU, "-I", CS, "-I", INC]
def run(cmd):
    subprocess.run(cmd, check=True)
run(["g++", "-std=c++17", "-x", "c++", "-ffp-contract=off", "-Wno-unused-function", "-Wno-unknown-pragmas"] + common +
    ["-c", os.path.join(CS, "engine.cu"), "-o", OUT + "/engine.o"])
run(["g++", "-std=c++17"] + common + ["-c", os.path.join(EMU, "cuda_emu.cpp"), "-o", OUT + "/cuda_emu.o"])
for c in ("flake_host.c", "md5.c"):
    run(["gcc", "-std=gnu11"] + common + ["-c", os.path.join(CS, c), "-o", OUT + "/" + c + ".o"])
lib_path = OUT + "/libflake_emu_asan.so"
run(["g++", "-shared", "-fsanitize=address", "-o", lib_path, OUT + "/engine.o", OUT + "/cuda_emu.o",
     OUT + "/flake_host.c.o", OUT + "/md5.c.o", "-Wl,-Bsymbolic", "-lpthread"])

from flake_b200 import api, synth
from oracle import pyoracle as po
lib = api.load_library(lib_path)
cases = [(4096 + 500, 2, 16, 44100, 8, {}), (1024 + 7, 1, 16, 44100, 8, {"block_size": 1024}),
         (2048, 2, 24, 96000, 12, {"block_size": 2048}), (2048 + 100, 2, 16, 44100, 5, {"block_size": 2048}),
         (1024 + 256, 3, 24, 48000, 9, {"block_size": 1024}), (1152 + 301, 2, 16, 44100, 2, {}),
         (576 * 2 + 40, 2, 16, 44100, 8, {"block_size": 576, "order_method": 6, "max_prediction_order": 32}),
         (4095 * 2, 2, 16, 44100, 2, {"block_size": 4095}), (16 * 40 + 5, 2, 16, 44100, 5, {"block_size": 16}),
         (128 * 20, 2, 16, 44100, 9, {"block_size": 128})]
ok = True
for n, ch, bps, rate, level, ov in cases:
    pcm = synth.synth_pcm(n, ch, bps, rate, seed=level, kind="impulses" if level == 9 else "mix")
    got = api.encode_batch(lib, pcm, rate, bps, level, chunk_blocks=3, **ov)
    want = po.encode_stream(pcm, rate, bps, level, **ov)[0]
    print(n, ch, bps, level, ov, got.payload == want, flush=True)
    ok &= got.payload == want
sys.exit(0 if ok else 1)
