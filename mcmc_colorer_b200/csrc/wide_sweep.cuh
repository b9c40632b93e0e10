// Wide-palette sweep: nCol > 512 (up to 65 535, u16 colours).  The reference takes any palette -- its scratch is a bool[n * nCol]
// array (coloringMCMC_main.cu:5-60, coloringMCMC_balance.cu:96-108) -- while the kernels of sweep_kernel.cuh / binned_sweep.cuh /
// blocked_sweep.cuh keep a vertex' occupancy in at most eight 64-bit registers.
//
// With a palette this wide a vertex rarely conflicts (probability ~ degree / nCol), and a vertex that does NOT conflict needs no
// occupancy information at all: its proposal is the "stay" distribution, a function of its own colour and the draw.  So a sweep is
//   wide_tables_kernel   the proposal table of this colouring (UNIFORM: free-colour weight by number of occupied colours; DYNAMIC:
//                        dist[k]) in global memory -- nCol floats no longer fit a CTA's shared memory next to everything else;
//   wide_count_kernel    one pass over the degree-binned rows (thread / warp / CTA rows of binned_sweep.cuh) that only COUNTS the
//                        neighbours sharing the vertex' colour: id load, colour gather, compare.  Warp rows are walked as one flat
//                        range per batch of 32 rows (lane -> (row, offset) by a binary search over the scanned row lengths), so short
//                        rows cost no idle lanes.  Non-conflicting vertices are finished here; the others -- and rows at least as long
//                        as the palette, whose free-colour count DYNAMIC needs -- are queued;
//   wide_walk_kernel     the queued rows: occupancy BITMAPS of nCol bits in shared memory (a batch of rows per warp, one bitmap for
//                        the whole CTA for hub rows), filled with shared-memory atomic ORs, then free-colour count (popc), proposal
//                        and the sequential float32 CDF walk over the bitmap words by dense lanes; last CTA finalizes the sweep.
// Class-size deltas go straight to the global scratch (only vertices that change colour).  Same result, bit for bit, as the narrow
// kernels' commit_vertex: identical float32 operations in identical order.
#pragma once
#include "binned_sweep.cuh"

namespace mcmcb200 {

constexpr uint32_t kWideWarpBytes = 12288; // shared memory of one warp of wide_walk_kernel: the bitmaps of a batch of queued rows

struct WideArgs {
	uint8_t *     fp;       // [nGlobal (padded)] low byte of every vertex' colour: what the counting pass gathers (see wide_fingerprint_kernel)
	const float * S;        // [nCol + 1]
	float *       tab;      // [nCol + 1]  UNIFORM: freeW[Zn];  DYNAMIC: dist[k]
	uint32_t      bmWords;  // 32-bit words of one bitmap
	uint32_t      bmStride; // words between the bitmaps of a batch (odd: lane j reading word w of bitmap j is conflict free)
	uint32_t      batch;    // warp rows per batch (1..32): as many bitmaps as fit the warp's shared memory
	uint32_t      warpBytes;// shared memory per warp
	uint32_t      words64;  // 64-bit words of a debug mask row
};

__host__ inline void wide_geometry(uint32_t nCol, WideArgs & wa) {
	wa.bmWords = (nCol + 31u) / 32u;
	wa.bmStride = wa.bmWords | 1u;
	uint32_t budget = kWideWarpBytes;
	uint32_t batch = (budget - 128u) / (4u * wa.bmStride);          // (128 bytes: the batch's "same colour" counters)
	if (batch > 32u) batch = 32u;
	if (batch < 1u) { batch = 1u; budget = 4u * wa.bmStride + 128u; }
	wa.batch = batch;
	wa.warpBytes = (4u * wa.bmStride * batch + 128u + 15u) & ~15u;
}
__host__ inline size_t wide_smem_bytes(const WideArgs & wa) {
	return sizeof(uint32_t) * 16 + (size_t)wa.warpBytes * (kThreadsBin / 32);
}

// The counting pass only asks "does this neighbour have my colour?".  It gathers from a one-byte fingerprint of the colouring (the
// low byte) instead of the u16 colour array: half the footprint -- 50 MB instead of 100 MB for 5e7 vertices, which is what decides
// whether the random gathers hit L2 (ncu on config 4: 57 % hit rate and 17 GB of DRAM sector fills per sweep with the u16 array) --
// and re-checks the full colour on the rare fingerprint match (1/256 of the edges + the true matches).  Rebuilt from the current
// colouring at the start of every sweep (150 MB of streaming), so no other code path has to maintain it.
__global__ void wide_fingerprint_kernel(const SweepArgs a, const WideArgs wa) {
	const DevState * st = a.st;
	const uint16_t * cur = a.colorsOverride ? static_cast<const uint16_t *>(a.colorsOverride) : static_cast<const uint16_t *>(a.colors[st->sweep & 1]);
	const uint32_t n8 = (a.nGlobal + 7u) >> 3;                   // (the colour buffers are padded by 64 Ki entries)
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += gridDim.x * blockDim.x) {
		const uint4 c = __ldg(reinterpret_cast<const uint4 *>(cur) + i);
		uint2 o;
		o.x = (c.x & 0xffu) | ((c.x >> 8) & 0xff00u) | ((c.y & 0xffu) << 16) | ((c.y >> 16) << 24);
		o.y = (c.z & 0xffu) | ((c.z >> 8) & 0xff00u) | ((c.w & 0xffu) << 16) | ((c.w >> 16) << 24);
		reinterpret_cast<uint2 *>(wa.fp)[i] = o;
	}
}

// S table: once per handle (one thread: the sum is sequential by definition)
__global__ void wide_S_kernel(float * S, uint32_t nCol, float eps) {
	if (blockIdx.x || threadIdx.x) return;
	float s = 0.0f; S[0] = 0.0f;
	for (uint32_t k = 0; k < nCol; ++k) { s = __fadd_rn(s, eps); S[k + 1] = s; }
}

// proposal table of the colouring about to be swept (same expressions as fill_proposal_table)
__global__ void wide_tables_kernel(const SweepArgs a, const WideArgs wa) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= a.nCol) return;
	if (a.proposal == MCMCB200_PROPOSAL_DYNAMIC) {
		const unsigned long long * hc = a.hist[a.st->sweep & 1];
		const float nF = __uint2float_rn(a.nGlobal), dF = __uint2float_rn(a.nCol - 1u);
		wa.tab[k] = __fdiv_rn(__fsub_rn(1.0f, __fdiv_rn(__uint2float_rn((uint32_t)hc[k]), nF)), dF);
	} else {
		wa.tab[k] = __fdiv_rn(__fsub_rn(1.0f, __fmul_rn(a.eps, __uint2float_rn(k))), __uint2float_rn(a.nCol - k));
	}
}

// ---- occupancy view of one row: nCol bits in shared memory ----
struct WideBitmap {
	const uint32_t * w; uint32_t words;
	__device__ __forceinline__ uint32_t next(uint32_t pos, uint32_t nCol) const {
		if (pos >= nCol) return nCol;
		uint32_t i = pos >> 5;
		uint32_t bits = w[i] & (0xffffffffu << (pos & 31u));
		while (!bits) { if (++i >= words) return nCol; bits = w[i]; }
		const uint32_t c = (i << 5) + (uint32_t)(__ffs((int)bits) - 1);
		return c < nCol ? c : nCol;
	}
	template <typename F> __device__ __forceinline__ void each(uint32_t nCol, F f) const {      // occupied colours in ascending order
		for (uint32_t i = 0; i < words; ++i) {
			uint32_t bits = w[i];
			while (bits) {
				const uint32_t c = (i << 5) + (uint32_t)(__ffs((int)bits) - 1);
				bits &= bits - 1u;
				if (c < nCol) f(c);
			}
		}
	}
};

// colour write + taboo + class-size deltas (finish_vertex of sweep_kernel.cuh; the deltas go straight to the global scratch)
__device__ __forceinline__ void finish_wide(const SweepArgs & a, uint16_t * __restrict__ nxt, uint32_t lv, uint32_t myOwn, uint32_t newc, bool touchTaboo) {
	MCMCB200_CHECK(newc < a.nCol && myOwn < a.nCol && lv < a.nLocal, a.st);
	if (touchTaboo && a.tabooIter) a.taboo[lv] = (uint16_t)((newc == myOwn) ? a.tabooIter : 0u);
	nxt[a.vBegin + lv] = (uint16_t)newc;
	if (newc != myOwn) { atomicAdd(a.scratch + 2 + myOwn, ~0ull); atomicAdd(a.scratch + 2 + newc, 1ull); }
}

// sequential float32 CDF walk of a conflicting vertex with free colours (x = free-colour weight (UNIFORM) or r (DYNAMIC)): exactly the
// reference's sequence of partial sums, straight over the bitmap words, 4 colours per stop test (walk_conflicting of sweep_kernel.cuh: the running sum is
// non-decreasing, so the first crossing inside a block of 4 is located afterwards; DYNAMIC's r < 0 corner takes single steps).
// ~4.5 instructions per colour whatever the number of occupied colours -- hub rows have nearly all of them occupied.
template <bool kDyn>
__device__ __forceinline__ uint32_t walk_wide(const SweepArgs & a, const WideArgs & wa, const WideBitmap & occ, float u, float x) {
	const uint32_t nCol = a.nCol;
	const float eps = a.eps;
	float cdf = 0.0f;
	const bool blocks = !kDyn || x >= 0.0f;
	for (uint32_t i = 0; i < occ.words; ++i) {
		uint32_t bits = occ.w[i];
		const uint32_t base = i << 5;
		const uint32_t lim = min(32u, nCol - base);
		uint32_t b = 0;
		if (blocks) {
			for (; b + 4u <= lim; b += 4u) {
				float q0, q1, q2, q3;
				if (kDyn) {
					const float * d = wa.tab + base + b;
					q0 = (bits & 1u) ? eps : __fadd_rn(__ldg(d + 0), x); q1 = (bits & 2u) ? eps : __fadd_rn(__ldg(d + 1), x);
					q2 = (bits & 4u) ? eps : __fadd_rn(__ldg(d + 2), x); q3 = (bits & 8u) ? eps : __fadd_rn(__ldg(d + 3), x);
				} else {
					q0 = (bits & 1u) ? eps : x; q1 = (bits & 2u) ? eps : x; q2 = (bits & 4u) ? eps : x; q3 = (bits & 8u) ? eps : x;
				}
				const float c1 = __fadd_rn(cdf, q0), c2 = __fadd_rn(c1, q1), c3 = __fadd_rn(c2, q2), c4 = __fadd_rn(c3, q3);
				if (kDyn ? (c4 >= u) : (c4 > u)) {
					const uint32_t first = (kDyn ? (c1 >= u) : (c1 > u)) ? 0u : (kDyn ? (c2 >= u) : (c2 > u)) ? 1u : (kDyn ? (c3 >= u) : (c3 > u)) ? 2u : 3u;
					return base + b + first;
				}
				cdf = c4;
				bits >>= 4;
			}
		}
		for (; b < lim; ++b) {
			const float q = (bits & 1u) ? eps : (kDyn ? __fadd_rn(__ldg(wa.tab + base + b), x) : x);
			bits >>= 1;
			cdf = __fadd_rn(cdf, q);
			if (kDyn ? (cdf >= u) : (cdf > u)) return base + b;
		}
	}
	return nCol - 1u;                                           // overflow contract: clamp to nCol-1
}

// phase 3 of one vertex up to the walk (commit_vertex of sweep_kernel.cuh with the occupancy behind `occ`).  ZnKnown: number of
// occupied colours if the caller has it (bitmaps: popc), 0xffffffff otherwise (lists: counted here, only when it is needed).
// Returns true when the vertex still has to WALK (u, x, Zn filled in); otherwise the vertex is finished.
template <bool kDyn, typename Occ>
__device__ __forceinline__ bool prepare_wide(const SweepArgs & a, const WideArgs & wa, uint32_t t, uint16_t * __restrict__ nxt, uint32_t v, uint32_t lv,
                                             uint32_t myOwn, const Occ & occ, uint32_t same, uint32_t ZnKnown, float stayW, float & u, float & x) {
	const uint32_t nCol = a.nCol;
	const float eps = a.eps;
	const bool viol = same > 0u;                                  // (counters and the violator list: wide_count_kernel)
	if (a.dbgMasks) {
		unsigned long long * row = a.dbgMasks + (size_t)lv * wa.words64;
		for (uint32_t w = 0; w < wa.words64; ++w) row[w] = 0ull;
		occ.each(nCol, [&](uint32_t c) { row[c >> 6] |= 1ull << (c & 63u); });
		a.dbgSame[lv] = same;
	}
	if (a.countOnly) return false;
	if (a.tabooIter) {
		const uint32_t tb = a.taboo[lv];
		if (tb > 0u) { a.taboo[lv] = (uint16_t)(tb - 1u); finish_wide(a, nxt, lv, myOwn, myOwn, false); return false; }
	}
	// Zn is only needed by conflicting vertices (and by DYNAMIC's "no free colour: no draw" rule, which a row shorter than the
	// palette can never meet)
	uint32_t Zn = ZnKnown;
	if (Zn == 0xffffffffu && viol) { Zn = 0u; occ.each(nCol, [&](uint32_t) { ++Zn; }); }
	const uint32_t Zp = (Zn == 0xffffffffu) ? nCol : nCol - Zn;    // (unknown: a non-conflicting list row, Zp > 0 for sure)
	if (kDyn && Zp == 0u) { finish_wide(a, nxt, lv, myOwn, myOwn, false); return false; }   // coloringMCMC_balance.cu:111-115: no draw, taboo untouched
	if (a.tape) u = a.tape[(size_t)(t - a.tapeBase) * a.nGlobal + v];
	else u = draw_to_uniform(philox_draw(a.seed, t + 1u, v, 0u), kDyn);
	if (!viol || Zp == 0u) {                                        // "stay" distribution
		const float sOwn = __ldg(wa.S + myOwn);
		const float tOwn = __fadd_rn(sOwn, stayW);
		const bool notBefore = kDyn ? (sOwn < u) : (sOwn <= u);
		const bool hit = kDyn ? (tOwn >= u) : (tOwn > u);
		const uint32_t newc = (notBefore && hit) ? myOwn : walk_stay<kDyn>(nCol, myOwn, eps, stayW, u);
		finish_wide(a, nxt, lv, myOwn, newc, true);
		return false;
	}
	if (!kDyn) x = __ldg(wa.tab + Zn);
	else {
		float rem = 0.0f;                                           // ascending colour order, like the reference's loop (:104-107)
		occ.each(nCol, [&](uint32_t c) { rem = __fadd_rn(rem, __fsub_rn(__ldg(wa.tab + c), eps)); });
		x = __fdiv_rn(rem, __uint2float_rn(Zp));
	}
	return true;
}

// queue of the rows wide_walk_kernel has to look at: q[0] entries of the warp queue, q[1] entries of the CTA queue, q[2] / q[3] the
// two work counters of wide_walk_kernel; warp-queue entries from q[8], CTA-queue entries in qCta.  Entry = local vertex id.
struct WideQueues { uint32_t * q; uint32_t * qCta; };

// fast path: a vertex without a same-coloured neighbour whose row is shorter than the palette (so it has free colours) -- the "stay"
// distribution (coloringMCMC_CPU.cpp:472-478), no occupancy needed
template <bool kDyn>
__device__ __forceinline__ void commit_quiet(const SweepArgs & a, const WideArgs & wa, uint32_t t, uint16_t * __restrict__ nxt, uint32_t v, uint32_t lv,
                                             uint32_t myOwn, float stayW) {
	if (a.tabooIter) {
		const uint32_t tb = a.taboo[lv];
		if (tb > 0u) { a.taboo[lv] = (uint16_t)(tb - 1u); finish_wide(a, nxt, lv, myOwn, myOwn, false); return; }
	}
	float u;
	if (a.tape) u = a.tape[(size_t)(t - a.tapeBase) * a.nGlobal + v];
	else u = draw_to_uniform(philox_draw(a.seed, t + 1u, v, 0u), kDyn);
	const float sOwn = __ldg(wa.S + myOwn);
	const float tOwn = __fadd_rn(sOwn, stayW);
	const bool notBefore = kDyn ? (sOwn < u) : (sOwn <= u);
	const bool hit = kDyn ? (tOwn >= u) : (tOwn > u);
	const uint32_t newc = (notBefore && hit) ? myOwn : walk_stay<kDyn>(a.nCol, myOwn, a.eps, stayW, u);
	finish_wide(a, nxt, lv, myOwn, newc, true);
}

// bookkeeping of one counted vertex: counters, violator list (tail cutting), then finish it here or queue it.  Returns true if the
// vertex has to go to wide_walk_kernel.
// rows the counting pass does not even look at when a walk pass follows: their occupancy is needed whatever they count (a row at
// least as long as the palette may have no free colour; the debug interface wants every bitmap), so wide_walk_kernel counts them
__device__ __forceinline__ bool wide_row_deferred(const SweepArgs & a, uint32_t deg) { return deg >= a.nCol || a.dbgMasks != nullptr; }

// counters + violator list (tail cutting) of one vertex whose same-colour count is known
__device__ __forceinline__ void wide_book(const SweepArgs & a, uint32_t v, uint32_t lv, uint32_t same,
                                          unsigned long long & accDirected, unsigned long long & accViol) {
	const bool viol = same > 0u;
	accDirected += same;
	accViol += viol ? 1ull : 0ull;
	if (a.dbgSame) a.dbgSame[lv] = same;
	if (viol && a.violList != nullptr) {
		if (a.forceEmit || __ldcg(&a.st->emitNow)) {
			const uint32_t idx = atomicAdd(a.violCount, 1u);
			if (idx < a.violCap) a.violList[idx] = v;
		}
	}
}

template <bool kDyn>
__device__ __forceinline__ bool counted_vertex(const SweepArgs & a, const WideArgs & wa, uint32_t t, uint16_t * __restrict__ nxt, uint32_t v, uint32_t lv,
                                               uint32_t own, uint32_t same, uint32_t deg, float stayW,
                                               unsigned long long & accDirected, unsigned long long & accViol) {
	const bool viol = same > 0u;
	wide_book(a, v, lv, same, accDirected, accViol);
	if (a.dbgMasks) return true;                                  // debug interface: every row builds its bitmap
	if (a.countOnly) return false;
	if (viol || deg >= a.nCol) return true;                       // needs the occupancy (free colours may be 0 only if deg >= nCol)
	commit_quiet<kDyn>(a, wa, t, nxt, v, lv, own, stayW);
	return false;
}

template <bool kDyn>
__global__ void __launch_bounds__(kThreadsBin, 4)
wide_count_kernel(const SweepArgs a, const BinnedArgs bn, const WideArgs wa, const WideQueues wq, const uint32_t finalizeHere) {
	using ColT = uint16_t;
	__shared__ uint32_t s_ctl[8];
	__shared__ unsigned long long s_red[16];
	__shared__ uint32_t s_sameAll[kThreadsBin];                  // per warp: the batch's "same colour" counters
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int nWarps = kThreadsBin / 32;
	DevState * st = a.st;
	if (!a.countOnly && st->convergedAt >= 0) return;
	const uint32_t t = st->sweep;
	const ColT * __restrict__ cur = a.colorsOverride ? static_cast<const ColT *>(a.colorsOverride) : static_cast<const ColT *>(a.colors[t & 1]);
	ColT * __restrict__ nxt = static_cast<ColT *>(a.colors[(t + 1) & 1]);
	const float stayW = stay_weight<kDyn>(a.nCol, a.eps);
	const uint64_t polLast = make_policy_evict_last();
	const uint8_t * __restrict__ fp = wa.fp;
	unsigned long long accDirected = 0ull, accViol = 0ull;

	// ---------------- CTA rows ----------------
	for (;;) {
		__syncthreads();
		if (tid == 0) { s_ctl[0] = atomicAdd(bn.counters + 2, 1u); s_ctl[1] = 0u; }
		__syncthreads();
		const uint32_t i = s_ctl[0];
		if (i >= bn.n[2]) break;
		const uint32_t lv = bn.list[2][i];
		const uint32_t e0 = a.rowptr[lv], e1 = a.rowptr[lv + 1];
		if (!finalizeHere && wide_row_deferred(a, e1 - e0)) {      // (uniform) counted by the walk pass, which needs the row anyway
			if (tid == 0) wq.qCta[atomicAdd(wq.q + 1, 1u)] = lv;
			continue;
		}
		const uint32_t own = (uint32_t)cur[a.vBegin + lv];
		uint32_t same = 0;
		for (uint32_t e = e0 + tid; e < e1; e += (uint32_t)kBinUnroll * kThreadsBin) {
			uint32_t nb[kBinUnroll], c[kBinUnroll];
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) nb[k] = (e + k * kThreadsBin < e1) ? __ldcs(a.neighs + e + k * kThreadsBin) : 0u;
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) c[k] = (e + k * kThreadsBin < e1) ? ld_color<uint8_t>(fp + nb[k], polLast) : 0xffffffffu;
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) if (c[k] == (own & 0xffu)) same += ((uint32_t)cur[nb[k]] == own);
		}
		same = __reduce_add_sync(0xffffffffu, same);
		if (lane == 0 && same) atomicAdd(&s_ctl[1], same);
		__syncthreads();
		if (tid == 0) {
			if (counted_vertex<kDyn>(a, wa, t, nxt, a.vBegin + lv, lv, own, s_ctl[1], e1 - e0, stayW, accDirected, accViol))
				wq.qCta[atomicAdd(wq.q + 1, 1u)] = lv;
		}
	}

	// ---------------- warp rows: the edges of a batch of 32 rows as one flat range ----------------
	{
		uint32_t * s_same = s_sameAll + warp * 32;
		for (;;) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(bn.counters + 1, 32u);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= bn.n[1]) break;
			const uint32_t cntB = min(32u, bn.n[1] - base);
			const bool valid = (uint32_t)lane < cntB;
			const uint32_t lvMine = valid ? bn.list[1][base + lane] : 0u;
			const uint32_t begMine = valid ? a.rowptr[lvMine] : 0u, endMine = valid ? a.rowptr[lvMine + 1] : 0u;
			const uint32_t ownMine = valid ? (uint32_t)cur[a.vBegin + lvMine] : 0xffffffffu;
			const bool deferred = valid && !finalizeHere && wide_row_deferred(a, endMine - begMine);   // counted by the walk pass
			const uint32_t lenMine = deferred ? 0u : endMine - begMine;
			uint32_t P = lenMine;                                     // inclusive scan of the row lengths
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t tt = __shfl_up_sync(0xffffffffu, P, o); if (lane >= o) P += tt; }
			const uint32_t T = __shfl_sync(0xffffffffu, P, 31);
			const uint32_t startMine = P - lenMine;
			__syncwarp();
			s_same[lane] = 0u;
			__syncwarp();
			for (uint32_t f0 = 0; f0 < T; f0 += 32u * kBinUnroll) {
				uint32_t nb[kBinUnroll], c[kBinUnroll], jj[kBinUnroll];
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) {
					const uint32_t f = f0 + 32u * k + lane;
					uint32_t j = 0;                                   // smallest j with P_j > f
#pragma unroll
					for (int step = 16; step; step >>= 1) { const uint32_t pj = __shfl_sync(0xffffffffu, P, (int)(j + step - 1u)); if (pj <= f) j += step; }
					j = min(j, 31u);
					const uint32_t sj = __shfl_sync(0xffffffffu, startMine, (int)j), bj = __shfl_sync(0xffffffffu, begMine, (int)j);
					jj[k] = j;
					nb[k] = (f < T) ? __ldcs(a.neighs + bj + (f - sj)) : 0u;
				}
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) c[k] = (f0 + 32u * k + lane < T) ? ld_color<uint8_t>(fp + nb[k], polLast) : 0xfffffffeu;
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) {
					const uint32_t ownJ = __shfl_sync(0xffffffffu, ownMine, (int)jj[k]);
					if (c[k] == (ownJ & 0xffu)) {                      // fingerprint match (1/256 + the true ones): look at the colour itself
						if ((uint32_t)cur[nb[k]] == ownJ) atomicAdd(&s_same[jj[k]], 1u);
					}
				}
			}
			__syncwarp();
			bool push = deferred;
			if (valid && !deferred) push = counted_vertex<kDyn>(a, wa, t, nxt, a.vBegin + lvMine, lvMine, ownMine, s_same[lane], endMine - begMine, stayW, accDirected, accViol);
			const uint32_t pm = __ballot_sync(0xffffffffu, push);
			if (pm) {
				uint32_t qb = 0;
				if (lane == 0) qb = atomicAdd(wq.q + 0, (uint32_t)__popc(pm));
				qb = __shfl_sync(0xffffffffu, qb, 0);
				if (push) wq.q[8u + qb + (uint32_t)__popc(pm & ((1u << lane) - 1u))] = lvMine;
			}
			__syncwarp();
		}
	}

	// ---------------- thread rows ----------------
	for (;;) {
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(bn.counters + 0, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= bn.n[0]) break;
		const bool valid = base + lane < bn.n[0];
		const uint32_t lv = valid ? bn.list[0][base + lane] : 0u;
		const uint32_t beg = valid ? a.rowptr[lv] : 0u;
		const bool deferred = valid && !finalizeHere && a.dbgMasks != nullptr;
		const uint32_t deg = (valid && !deferred) ? (a.rowptr[lv + 1] - beg) : 0u;
		const uint32_t own = valid ? (uint32_t)cur[a.vBegin + lv] : 0u;
		uint32_t same = 0;
		for (uint32_t i = 0; i < deg; i += 4u) {
			uint32_t nb[4], c[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) nb[k] = (i + k < deg) ? __ldg(a.neighs + beg + i + k) : 0u;
#pragma unroll
			for (int k = 0; k < 4; ++k) c[k] = (i + k < deg) ? ld_color<uint8_t>(fp + nb[k], polLast) : 0xffffffffu;
#pragma unroll
			for (int k = 0; k < 4; ++k) if (c[k] == (own & 0xffu)) same += ((uint32_t)cur[nb[k]] == own);
		}
		bool push = deferred;
		if (valid && !deferred) push = counted_vertex<kDyn>(a, wa, t, nxt, a.vBegin + lv, lv, own, same, deg, stayW, accDirected, accViol);
		const uint32_t pm = __ballot_sync(0xffffffffu, push);
		if (pm) {
			uint32_t qb = 0;
			if (lane == 0) qb = atomicAdd(wq.q + 0, (uint32_t)__popc(pm));
			qb = __shfl_sync(0xffffffffu, qb, 0);
			if (push) wq.q[8u + qb + (uint32_t)__popc(pm & ((1u << lane) - 1u))] = lv;
		}
	}

	// ---- epilogue: counters; the sweep is finalized here only if no wide_walk_kernel follows (plain counting pass) ----
	accDirected = warp_reduce_add64(accDirected);
	accViol = warp_reduce_add64(accViol);
	__syncthreads();
	if (lane == 0) { s_red[warp] = accDirected; s_red[8 + warp] = accViol; }
	__syncthreads();
	if (tid == 0) {
		unsigned long long d = 0, vv = 0;
		for (int w = 0; w < nWarps; ++w) { d += s_red[w]; vv += s_red[8 + w]; }
		if (d) atomicAdd(a.scratch + 0, d);
		if (vv) atomicAdd(a.scratch + 1, vv);
	}
	if (finalizeHere && a.fuseFinalize) {
		__threadfence();
		__syncthreads();
		if (tid == 0) s_ctl[3] = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
		__syncthreads();
		if (s_ctl[3]) {
			__threadfence();
			finalize_sweep_device(a);
		}
	}
}

template <bool kDyn>
__global__ void __launch_bounds__(kThreadsBin, 2)
wide_walk_kernel(const SweepArgs a, const WideArgs wa, const WideQueues wq) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	using ColT = uint16_t;
	const uint32_t nCol = a.nCol, bmWords = wa.bmWords;
	uint32_t * s_ctl = reinterpret_cast<uint32_t *>(smem_raw);
	unsigned char * s_warp = reinterpret_cast<unsigned char *>(s_ctl + 16);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	unsigned char * mineRaw = s_warp + (size_t)warp * wa.warpBytes;       // this warp's bitmaps
	uint32_t * s_bm = reinterpret_cast<uint32_t *>(s_warp);               // hub rows: warp 0's area holds the CTA's bitmap
	DevState * st = a.st;
	if (!a.countOnly && st->convergedAt >= 0) return;
	const uint32_t t = st->sweep;
	const ColT * __restrict__ cur = a.colorsOverride ? static_cast<const ColT *>(a.colorsOverride) : static_cast<const ColT *>(a.colors[t & 1]);
	ColT * __restrict__ nxt = static_cast<ColT *>(a.colors[(t + 1) & 1]);
	const float stayW = stay_weight<kDyn>(nCol, a.eps);
	const uint64_t polLast = make_policy_evict_last();
	const uint32_t nCtaQ = __ldcg(wq.q + 1), nWarpQ = __ldcg(wq.q + 0);
	unsigned long long accDirected = 0ull, accViol = 0ull;        // of the rows the counting pass deferred (wide_row_deferred)

	// ---------------- queued hub rows: one bitmap for the CTA ----------------
	for (;;) {
		__syncthreads();
		if (tid == 0) { s_ctl[0] = (nCtaQ ? atomicAdd(wq.q + 2, 1u) : 0u); s_ctl[1] = 0u; s_ctl[2] = 0u; }
		for (uint32_t w = tid; w < bmWords; w += kThreadsBin) s_bm[w] = 0u;
		__syncthreads();
		const uint32_t i = s_ctl[0];
		if (i >= nCtaQ) break;
		const uint32_t lv = wq.qCta[i];
		const uint32_t e0 = a.rowptr[lv], e1 = a.rowptr[lv + 1];
		const uint32_t own = (uint32_t)cur[a.vBegin + lv];
		uint32_t same = 0;
		for (uint32_t e = e0 + tid; e < e1; e += (uint32_t)kBinUnroll * kThreadsBin) {
			uint32_t nb[kBinUnroll], c[kBinUnroll];
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) nb[k] = (e + k * kThreadsBin < e1) ? __ldcs(a.neighs + e + k * kThreadsBin) : 0u;
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) c[k] = (e + k * kThreadsBin < e1) ? ld_color<ColT>(cur + nb[k], polLast) : 0xffffffffu;
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) {
				if (e + k * kThreadsBin < e1) {
					same += (c[k] == own);
					const uint32_t bit = 1u << (c[k] & 31u);
					MCMCB200_CHECK(c[k] < nCol, st);
					if (!(s_bm[c[k] >> 5] & bit)) atomicOr(&s_bm[c[k] >> 5], bit);      // (hub rows hit the same few words: test first)
				}
			}
		}
		same = __reduce_add_sync(0xffffffffu, same);
		if (lane == 0 && same) atomicAdd(&s_ctl[1], same);
		__syncthreads();
		uint32_t zn = 0;
		for (uint32_t w = tid; w < bmWords; w += kThreadsBin) zn += (uint32_t)__popc(s_bm[w]);
		zn = __reduce_add_sync(0xffffffffu, zn);
		if (lane == 0 && zn) atomicAdd(&s_ctl[2], zn);
		__syncthreads();
		if (tid == 0) {
			const WideBitmap occ{s_bm, bmWords};
			float u = 0.0f, x = 0.0f;
			if (wide_row_deferred(a, e1 - e0)) wide_book(a, a.vBegin + lv, lv, s_ctl[1], accDirected, accViol);
			if (prepare_wide<kDyn>(a, wa, t, nxt, a.vBegin + lv, lv, own, occ, s_ctl[1], s_ctl[2], stayW, u, x))
				finish_wide(a, nxt, lv, own, walk_wide<kDyn>(a, wa, occ, u, x), true);
		}
	}

	// ---------------- queued warp / thread rows: a batch of rows per warp, one bitmap each, edges as one flat range ----------------
	{
		uint32_t * bms = reinterpret_cast<uint32_t *>(mineRaw);
		const uint32_t B = wa.batch, stride = wa.bmStride;
		uint32_t * s_same = bms + (size_t)B * stride;                 // [32]
		for (;;) {
			uint32_t base = 0;
			if (lane == 0) base = nWarpQ ? atomicAdd(wq.q + 3, B) : 0u;
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= nWarpQ) break;
			const uint32_t cntB = min(B, nWarpQ - base);
			const bool valid = (uint32_t)lane < cntB;
			const uint32_t lvMine = valid ? wq.q[8u + base + lane] : 0u;
			const uint32_t begMine = valid ? a.rowptr[lvMine] : 0u, endMine = valid ? a.rowptr[lvMine + 1] : 0u;
			const uint32_t ownMine = valid ? (uint32_t)cur[a.vBegin + lvMine] : 0xffffffffu;
			uint32_t P = endMine - begMine;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t tt = __shfl_up_sync(0xffffffffu, P, o); if (lane >= o) P += tt; }
			const uint32_t T = __shfl_sync(0xffffffffu, P, 31);
			const uint32_t startMine = P - (endMine - begMine);
			__syncwarp();
			for (uint32_t w = lane; w < cntB * stride; w += 32u) bms[w] = 0u;
			s_same[lane] = 0u;
			__syncwarp();
			for (uint32_t f0 = 0; f0 < T; f0 += 32u * kBinUnroll) {
				uint32_t nb[kBinUnroll], c[kBinUnroll], jj[kBinUnroll];
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) {
					const uint32_t f = f0 + 32u * k + lane;
					uint32_t j = 0;
#pragma unroll
					for (int step = 16; step; step >>= 1) { const uint32_t pj = __shfl_sync(0xffffffffu, P, (int)(j + step - 1u)); if (pj <= f) j += step; }
					j = min(j, 31u);
					const uint32_t sj = __shfl_sync(0xffffffffu, startMine, (int)j), bj = __shfl_sync(0xffffffffu, begMine, (int)j);
					jj[k] = j;
					nb[k] = (f < T) ? __ldcs(a.neighs + bj + (f - sj)) : 0u;
				}
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) c[k] = (f0 + 32u * k + lane < T) ? ld_color<ColT>(cur + nb[k], polLast) : 0xfffffffeu;
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) {
					const uint32_t ownJ = __shfl_sync(0xffffffffu, ownMine, (int)jj[k]);
					if (f0 + 32u * k + lane < T) {
						MCMCB200_CHECK(c[k] < nCol && jj[k] < cntB && nb[k] < a.nGlobal, st);
						if (c[k] == ownJ) atomicAdd(&s_same[jj[k]], 1u);
						atomicOr(&bms[(size_t)jj[k] * stride + (c[k] >> 5)], 1u << (c[k] & 31u));
					}
				}
			}
			__syncwarp();
			if (valid) {
				const uint32_t * bm = bms + (size_t)lane * stride;    // odd stride: word w of the 32 bitmaps sits in 32 different banks
				uint32_t zn = 0;
				for (uint32_t w = 0; w < bmWords; ++w) zn += (uint32_t)__popc(bm[w]);
				const WideBitmap occ{bm, bmWords};
				float u = 0.0f, x = 0.0f;
				if (wide_row_deferred(a, endMine - begMine)) wide_book(a, a.vBegin + lvMine, lvMine, s_same[lane], accDirected, accViol);
				if (prepare_wide<kDyn>(a, wa, t, nxt, a.vBegin + lvMine, lvMine, ownMine, occ, s_same[lane], zn, stayW, u, x))
					finish_wide(a, nxt, lvMine, ownMine, walk_wide<kDyn>(a, wa, occ, u, x), true);
			}
			__syncwarp();
		}
	}

	if (accDirected) atomicAdd(a.scratch + 0, accDirected);       // (few rows: one atomic per thread that owned one)
	if (accViol) atomicAdd(a.scratch + 1, accViol);
	if (a.fuseFinalize) {
		__threadfence();
		__syncthreads();
		if (tid == 0) s_ctl[3] = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
		__syncthreads();
		if (s_ctl[3]) {
			__threadfence();
			finalize_sweep_device(a);
		}
	}
}

} // namespace mcmcb200
