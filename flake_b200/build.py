"""Build recipes (plain nvcc / gcc command lines, no build system).

  build_product()  flake_b200/lib/libflake.so      nvcc, sm_100a  -- the shipped library
  build_oracle()   oracle/libflake_oracle.so (+ oracle/_ref/* when /root/reference exists)
  build_emu()      tests/cuda_emu/libflake_emu.so   g++ fiber emulation, TEST ONLY

Run as a script: ``python -m flake_b200.build [product|oracle|emu|all]``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "flake_b200", "csrc")
LIBDIR = os.path.join(ROOT, "flake_b200", "lib")
EMUDIR = os.path.join(ROOT, "tests", "cuda_emu")
ORACLE = os.path.join(ROOT, "oracle")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-fmad=false",              # FP64 LPC stage must not contract mul+add (lpc.c is built -std=c99)
    "-Xcompiler", "-fPIC", "-std=c++17",
]


HOST_C = ("flake_host.c", "flake_corpus.c", "md5.c", "md5_mb.c")      # the C host layer (gcc)


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _csrc_files():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def build_product(force: bool = False, variant: str = "", defines=()) -> str:
    """The shipped library.  `variant`/`defines` (dev only, tools/variants.py) build
    flake_b200/lib/var/libflake_<variant>.so with extra -D flags for A/B timing on the GPU."""
    libdir = os.path.join(LIBDIR, "var", variant) if variant else LIBDIR
    os.makedirs(libdir, exist_ok=True)
    out = os.path.join(libdir, "libflake.so")
    inc = os.path.join(ROOT, "include")
    srcs = _csrc_files() + [os.path.join(inc, f) for f in os.listdir(inc)]
    if not force and not _newer(out, srcs):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    obj_cu = os.path.join(libdir, "engine.o")
    cu_srcs = [f for f in srcs if f.endswith((".cu", ".cuh", ".h"))]
    if force or defines or _newer(obj_cu, cu_srcs):          # a change of the C host files alone only relinks
        _run([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] +
             ["-I", inc, "-c", os.path.join(CSRC, "engine.cu"), "-o", obj_cu])
    objs = [obj_cu]
    for c in HOST_C:
        o = os.path.join(libdir, c.replace(".c", ".o"))
        _run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-I", inc, "-I", CSRC,
              "-c", os.path.join(CSRC, c), "-o", o])
        objs.append(o)
    _run([nvcc, "-shared", "-o", out] + objs +
         ["-Xlinker", "-Bsymbolic", "-Xlinker", "--version-script=" + os.path.join(CSRC, "libflake.map"),
          "-cudart", "static", "-lpthread"])
    return out


def build_emu(force: bool = False) -> str:
    out = os.path.join(EMUDIR, "libflake_emu.so")
    inc = os.path.join(ROOT, "include")
    srcs = _csrc_files() + [os.path.join(EMUDIR, "cuda_emu.h"), os.path.join(EMUDIR, "cuda_emu.cpp")]
    if not force and not _newer(out, srcs):
        return out
    common = ["-O1", "-g", "-fPIC", "-ffp-contract=off", "-DFLAKE_B200_CUDA_EMU", "-I", EMUDIR, "-I", CSRC,
              "-I", inc, "-Wno-unused-function", "-Wno-unknown-pragmas"]
    objs = []
    o = os.path.join(EMUDIR, "engine_emu.o")
    _run(["g++", "-std=c++17", "-x", "c++"] + common + ["-c", os.path.join(CSRC, "engine.cu"), "-o", o])
    objs.append(o)
    o = os.path.join(EMUDIR, "cuda_emu.o")
    _run(["g++", "-std=c++17"] + common + ["-c", os.path.join(EMUDIR, "cuda_emu.cpp"), "-o", o])
    objs.append(o)
    for c in HOST_C:
        o = os.path.join(EMUDIR, c.replace(".c", "_emu.o"))
        _run(["gcc", "-std=gnu11", "-O1", "-g", "-fPIC", "-Wall", "-I", inc, "-I", CSRC,
              "-c", os.path.join(CSRC, c), "-o", o])
        objs.append(o)
    _run(["g++", "-shared", "-o", out] + objs + ["-Wl,-Bsymbolic", "-lpthread"])
    return out


def build_oracle() -> str:
    _run(["make", "-C", ORACLE, "oracle"])
    if os.path.isdir("/root/reference/libflake"):
        _run(["make", "-C", ORACLE, "ref"])
        if os.path.exists(os.path.join(LIBDIR, "libflake.so")):
            _run(["make", "-C", ORACLE, "_ref/flake_cli_b200", "_ref/flake_cli_b200_batch", "_ref/api_example_b200"])
    return os.path.join(ORACLE, "libflake_oracle.so")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("product", "all"):
        print(build_product(force=True))
    if what in ("oracle", "all"):
        print(build_oracle())
    if what in ("emu", "all"):
        print(build_emu(force=True))
