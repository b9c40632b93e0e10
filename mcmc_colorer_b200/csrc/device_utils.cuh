// Device helpers: Philox4x32-10, cache-hinted loads, warp reductions.  sm_100a only.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// -DMCMCB200_BOUNDS_CHECK=1 builds a checking variant of the library: every index the hot kernels compute themselves (stage positions,
// granule destinations, slot tables, palette indices, queue slots) is tested before use and a violation raises the sticky device error
// flag 4, which every host call that reads results reports as an error.  (compute-sanitizer is not available on the pool these kernels
// were developed on; scripts/sanitize_driver.py runs every kernel family against the checking build.)
#ifndef MCMCB200_BOUNDS_CHECK
#define MCMCB200_BOUNDS_CHECK 0
#endif
#if MCMCB200_BOUNDS_CHECK
#define MCMCB200_CHECK(cond, st) do { if (!(cond)) *reinterpret_cast<volatile uint32_t *>(&(st)->errorFlag) = 4u; } while (0)
#else
#define MCMCB200_CHECK(cond, st) ((void)0)
#endif

namespace mcmcb200 {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Stateless: replaces the 48-byte-per-vertex curandState of
// the reference (GPUutils/GPURandomizer.cu:8-13) -- 0 bytes of RNG state traffic per sweep.
// RNG contract v2 (include/mcmcb200.h): ONE call serves FOUR consecutive vertices --
//   counter = (vertex >> 2, purpose, sweep, 0), key = (seed_lo, seed_hi), draw of vertex v = output word (v & 3).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint32_t sweep, uint32_t group, uint32_t purpose) {
	uint32_t c0 = group, c1 = purpose, c2 = sweep, c3 = 0u;
	uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
	for (int r = 0; r < 10; ++r) {
		const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	return make_uint4(c0, c1, c2, c3);
}

// the draw of ONE vertex (direct kernels: thread per vertex in arbitrary order; the blocked sweep fills a per-tile table instead)
__device__ __forceinline__ uint32_t philox_draw(uint64_t seed, uint32_t sweep, uint32_t vertex, uint32_t purpose) {
	const uint4 w = philox4(seed, sweep, vertex >> 2, purpose);
	const uint32_t k = vertex & 3u;
	return k == 0u ? w.x : k == 1u ? w.y : k == 2u ? w.z : w.w;
}

// u in [0,1) (UNIFORM, like uniform_real_distribution<float>) or (0,1] (DYNAMIC, like curand_uniform)
__device__ __forceinline__ float draw_to_uniform(uint32_t x, bool openAtZero) {
	const uint32_t m = (x >> 8) + (openAtZero ? 1u : 0u);
	return __uint2float_rn(m) * 5.9604644775390625e-8f;   // exact: m <= 2^24, power-of-two scale
}

// ---------------------------------------------------------------------------------------------
// Loads.  The CSR neighbour stream is read exactly once per sweep: 256-bit, L1 no-allocate, L2 evict-first
// (SASS: LDG.E.NA.EFL2.256.CONSTANT).  The colour gathers are the only data with reuse: L2 evict-last policy.
// ---------------------------------------------------------------------------------------------
struct U32x8 { uint32_t v[8]; };

__device__ __forceinline__ U32x8 ld_stream_256(const uint32_t * p /* 32-byte aligned */) {
	U32x8 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
		: "l"(p));
	return r;
}

__device__ __forceinline__ uint64_t make_policy_evict_last() {
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

template <typename ColT> __device__ __forceinline__ uint32_t ld_color(const ColT * p, uint64_t pol);
template <> __device__ __forceinline__ uint32_t ld_color<uint8_t>(const uint8_t * p, uint64_t pol) {
	uint32_t r;
	asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
	return r;
}
template <> __device__ __forceinline__ uint32_t ld_color<uint16_t>(const uint16_t * p, uint64_t pol) {
	uint32_t r;
	asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
	return r;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk copies (TMA engine; SASS: SYNCS.* and UBLKCP).  One elected thread arms the barrier with the
// byte count and issues the copies; every thread of the CTA (or of a warp group) waits on the phase parity.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long * bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long * bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long * bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long * bar, uint32_t parity) {
	asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
	             :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses (ld/st, the acquire of a hand-over flag) ordered before the async-proxy copies that follow
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(void * dstSmem, const void * srcGmem, uint32_t bytes, unsigned long long * bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dstSmem)), "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void * dstSmem, const void * srcGmem, uint32_t bytes, unsigned long long * bar, uint64_t pol) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
	             :: "r"(smem_u32(dstSmem)), "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// pull a contiguous range (16-byte aligned, a multiple of 16 bytes) into L2 ahead of the loads that will stream it
__device__ __forceinline__ void bulk_prefetch_l2(const void * srcGmem, uint32_t bytes) {
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(srcGmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t make_policy_evict_first() {
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

__device__ __forceinline__ uint64_t warp_reduce_or64(uint64_t x) {
	const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)x);
	const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(x >> 32));
	return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint64_t warp_reduce_add64(uint64_t x) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
	return x;
}

} // namespace mcmcb200
