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

/* 16-byte asynchronous global->shared copy (LDGSTS); both addresses 16-byte aligned */
#ifdef FLAKE_B200_CUDA_EMU
static inline void fb_cp_async16(void *dst_shared, const void *src_global) { memcpy(dst_shared, src_global, 16); }
static inline void fb_cp_async_wait_all(void) {}
#else
__device__ __forceinline__ void fb_cp_async16(void *dst_shared, const void *src_global)
{
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst_shared);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src_global) : "memory");
}
__device__ __forceinline__ void fb_cp_async_wait_all(void)
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
#endif

#endif
