/*
 * cuda_compat.h -- the one switch between the real CUDA toolchain (product
 * build, nvcc, sm_100a) and the test-only fiber emulation that lets the host
 * CI run kernel logic without a GPU (tests/cuda_emu/, -DFLAKE_B200_CUDA_EMU;
 * never defined by the product build).
 */
#ifndef FLAKE_B200_CUDA_COMPAT_H
#define FLAKE_B200_CUDA_COMPAT_H

#ifdef FLAKE_B200_CUDA_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define FB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#endif
