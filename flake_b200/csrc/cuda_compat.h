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

/* ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier ----------------------
 * cp.async.bulk (SASS: UBLKCP) moves a contiguous byte range with ONE instruction issued by
 * one thread; the mbarrier counts the bytes that land (complete_tx).  Source, destination and
 * size are multiples of 16 bytes.  The emulation copies at issue time. */
#ifdef FLAKE_B200_CUDA_EMU
/* the barrier word counts completed phases: a phase = one expect_tx (+ its copy, done on the spot by
 * the issuing fiber, which does not yield in between); a waiter yields until the phase of the parity
 * it names is over, like mbarrier.try_wait.parity -- a fiber that runs ahead of the issuing one must
 * not read the stage before its refill */
typedef unsigned long long fb_mbar_t;
static inline void fb_mbar_init(fb_mbar_t *bar, unsigned) { *bar = 0; }
static inline void fb_mbar_init_fence(void) {}
static inline void fb_mbar_expect_tx(fb_mbar_t *bar, unsigned) { *(volatile fb_mbar_t *)bar += 1; }
static inline void fb_bulk_g2s(void *dst_shared, const void *src_global, unsigned bytes, fb_mbar_t *) { memcpy(dst_shared, src_global, bytes); }
static inline void fb_mbar_wait(fb_mbar_t *bar, unsigned parity) { while ((*(volatile fb_mbar_t *)bar & 1u) == parity) cuemu::yield(); }
#else
typedef unsigned long long fb_mbar_t;
__device__ __forceinline__ void fb_mbar_init(fb_mbar_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
/* makes the initialised barrier visible to the async proxy before the first bulk copy */
__device__ __forceinline__ void fb_mbar_init_fence(void)
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fb_mbar_expect_tx(fb_mbar_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fb_bulk_g2s(void *dst_shared, const void *src_global, unsigned bytes, fb_mbar_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fb_mbar_wait(fb_mbar_t *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
#endif

/* 2^-s as a double, 0 <= s < 1023 */
#ifdef FLAKE_B200_CUDA_EMU
static inline double fb_exp2_neg(int s) { return ldexp(1.0, -s); }
#else
__device__ __forceinline__ double fb_exp2_neg(int s) { return __hiloint2double((1023 - s) << 20, 0); }
#endif

/* marks a block as not speculatable, so that a rarely taken `if` stays a branch instead of
 * being turned into selects executed by every thread */
#ifdef FLAKE_B200_CUDA_EMU
#define FB_COLD_BLOCK() do { } while (0)
#else
#define FB_COLD_BLOCK() asm volatile("" ::: "memory")
#endif

/* 128-bit load from shared memory at a compile-time byte offset from a shared-space address.
 * Pointers that reach a function as arguments are generic to the compiler; going through the
 * 32-bit shared address makes the loads LDS instead of generic LD. */
#ifdef FLAKE_B200_CUDA_EMU
typedef const unsigned char *fb_sptr;
static inline fb_sptr fb_to_sptr(const void *p) { return (const unsigned char *)p; }
template <int OFF> static inline int4 fb_lds128(fb_sptr a) { int4 v; memcpy(&v, a + OFF, 16); return v; }
static inline int4 fb_lds128_at(fb_sptr a, int byte_off) { int4 v; memcpy(&v, a + byte_off, 16); return v; }
#else
typedef unsigned fb_sptr;
__device__ __forceinline__ fb_sptr fb_to_sptr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int OFF> __device__ __forceinline__ int4 fb_lds128(fb_sptr a)
{
    int4 v;
    asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF));
    return v;
}
__device__ __forceinline__ int4 fb_lds128_at(fb_sptr a, int byte_off)
{
    int4 v;
    asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a + (unsigned)byte_off));
    return v;
}
#endif

#endif
