// Wide-palette sweep: nCol > 512 (up to 65 535, u16 colours).  The reference takes any palette -- its scratch is a bool[n * nCol]
// array (coloringMCMC_main.cu:5-60, coloringMCMC_balance.cu:96-108) -- while the kernels of sweep_kernel.cuh / binned_sweep.cuh /
// blocked_sweep.cuh keep a vertex' occupancy in at most eight 64-bit registers.  Here the occupancy of a vertex is
//   * thread rows (degree <= 32): the LIST of its neighbours' colours, parked lane-interleaved in shared memory (64 B per vertex
//     instead of nCol/8): "same colour" count while gathering; everything else is only needed by the few vertices that conflict,
//     which walk the colours in ascending order through the list (next occupied colour >= position: O(degree) per step);
//   * warp rows / CTA rows (degree-binned lists of binned_sweep.cuh): a bitmap of nCol bits per warp (or per CTA) in shared
//     memory, filled with shared-memory atomic ORs while the lanes stride the row with coalesced neighbour-id loads.
// Tables that no longer fit a CTA's shared memory live in global memory (L2 resident, read through the read-only path): S[k] = eps
// added k times, and the proposal table (UNIFORM: free-colour weight by number of occupied colours; DYNAMIC: dist[k], refreshed by
// wide_tables_kernel before every sweep).  Class-size deltas go straight to the global scratch (only vertices that change colour).
// Same result, bit for bit, as the narrow kernels' commit_vertex: identical float32 operations in identical order.
#pragma once
#include "binned_sweep.cuh"

namespace mcmcb200 {

constexpr uint32_t kWideListCap = 32;     // == kBinThreadMax: colours a thread row parks
constexpr uint32_t kWideQueueCap = 64;    // parked walks of thread rows per warp: < 32 left over + up to 32 new ones; drained 32 at a time by dense lanes
constexpr uint32_t kWideWarpBytes = 7680; // shared memory of one warp: max(colour lists + walk queue, the bitmaps of a batch of warp rows)

struct WideArgs {
	const float * S;        // [nCol + 1]
	float *       tab;      // [nCol + 1]  UNIFORM: freeW[Zn];  DYNAMIC: dist[k]
	uint32_t      bmWords;  // 32-bit words of one bitmap
	uint32_t      bmStride; // words between the bitmaps of a batch (odd: lane j reading word w of bitmap j is conflict free)
	uint32_t      batch;    // warp rows per batch (1..32): as many bitmaps as fit the warp's shared memory
	uint32_t      warpBytes;// shared memory per warp
	uint32_t      words64;  // 64-bit words of a debug mask row
};

__host__ __device__ inline uint32_t wide_list_queue_bytes() {
	return (uint32_t)(sizeof(uint16_t) * kWideListCap * 32u                       // colour lists of the 32 lanes
	                  + sizeof(uint16_t) * kWideListCap * kWideQueueCap           // parked lists
	                  + 6u * sizeof(uint32_t) * kWideQueueCap);                   // parked lv, own, deg, u, x (+ 1 spare)
}
__host__ inline void wide_geometry(uint32_t nCol, WideArgs & wa) {
	wa.bmWords = (nCol + 31u) / 32u;
	wa.bmStride = wa.bmWords | 1u;
	const uint32_t lq = wide_list_queue_bytes();
	uint32_t budget = kWideWarpBytes > lq ? kWideWarpBytes : lq;
	uint32_t batch = (budget - 128u) / (4u * wa.bmStride);          // (128 bytes: the batch's "same colour" counters)
	if (batch > 32u) batch = 32u;
	if (batch < 1u) { batch = 1u; budget = 4u * wa.bmStride + 128u; }
	wa.batch = batch;
	wa.warpBytes = (budget + 15u) & ~15u;
}
__host__ inline size_t wide_smem_bytes(const WideArgs & wa) {
	return sizeof(uint32_t) * 16 + sizeof(unsigned long long) * 16 + (size_t)wa.warpBytes * (kThreadsBin / 32);
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

// ---- occupancy views: "smallest occupied colour >= pos" (nCol when there is none) ----
struct WideList {                      // entries at base[j * stride], j < deg
	const uint16_t * base; uint32_t deg, stride;
	__device__ __forceinline__ uint32_t next(uint32_t pos, uint32_t nCol) const {
		uint32_t best = nCol;
		for (uint32_t j = 0; j < deg; ++j) { const uint32_t c = base[j * stride]; if (c >= pos && c < best) best = c; }
		return best;
	}
	template <typename F> __device__ __forceinline__ void each(uint32_t nCol, F f) const {      // occupied colours in ascending order
		for (uint32_t c = next(0u, nCol); c < nCol; c = next(c + 1u, nCol)) f(c);
	}
};
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

// sequential float32 CDF walk of a conflicting vertex with free colours: runs of free colours between consecutive occupied ones;
// x = free-colour weight (UNIFORM) or r (DYNAMIC).  Exactly the reference's sequence of partial sums.
template <bool kDyn, typename Occ>
__device__ __forceinline__ uint32_t walk_wide(const SweepArgs & a, const WideArgs & wa, const Occ & occ, float u, float x) {
	const uint32_t nCol = a.nCol;
	const float eps = a.eps;
	float cdf = 0.0f;
	uint32_t pos = 0u;
	while (pos < nCol) {
		const uint32_t oc = occ.next(pos, nCol);
		for (uint32_t k = pos; k < oc; ++k) {
			cdf = __fadd_rn(cdf, kDyn ? __fadd_rn(__ldg(wa.tab + k), x) : x);
			if (kDyn ? (cdf >= u) : (cdf > u)) return k;
		}
		if (oc >= nCol) break;
		cdf = __fadd_rn(cdf, eps);
		if (kDyn ? (cdf >= u) : (cdf > u)) return oc;
		pos = oc + 1u;
	}
	return nCol - 1u;                                           // overflow contract: clamp to nCol-1
}

// the same walk straight over the bitmap words, 4 colours per stop test (walk_conflicting of sweep_kernel.cuh: the running sum is
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
                                             uint32_t myOwn, const Occ & occ, uint32_t same, uint32_t ZnKnown, float stayW,
                                             unsigned long long & accDirected, unsigned long long & accViol, float & u, float & x) {
	const uint32_t nCol = a.nCol;
	const float eps = a.eps;
	const bool viol = same > 0u;
	accDirected += same;
	accViol += viol ? 1ull : 0ull;
	if (a.dbgMasks) {
		unsigned long long * row = a.dbgMasks + (size_t)lv * wa.words64;
		for (uint32_t w = 0; w < wa.words64; ++w) row[w] = 0ull;
		occ.each(nCol, [&](uint32_t c) { row[c >> 6] |= 1ull << (c & 63u); });
		a.dbgSame[lv] = same;
	}
	if (viol && a.violList != nullptr) {
		if (a.forceEmit || __ldcg(&a.st->emitNow)) {
			const uint32_t idx = atomicAdd(a.violCount, 1u);
			if (idx < a.violCap) a.violList[idx] = v;
		}
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

template <bool kDyn>
__global__ void __launch_bounds__(kThreadsBin, 2)
wide_sweep_kernel(const SweepArgs a, const BinnedArgs bn, const WideArgs wa) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	using ColT = uint16_t;
	const uint32_t nCol = a.nCol, bmWords = wa.bmWords;
	uint32_t * s_ctl = reinterpret_cast<uint32_t *>(smem_raw);
	unsigned long long * s_red = reinterpret_cast<unsigned long long *>(s_ctl + 16);
	unsigned char * s_warp = reinterpret_cast<unsigned char *>(s_red + 16);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int nWarps = kThreadsBin / 32;
	unsigned char * mineRaw = s_warp + (size_t)warp * wa.warpBytes;       // this warp's shared memory (bitmaps OR lists + queue)
	uint32_t * s_bm = reinterpret_cast<uint32_t *>(s_warp);               // CTA rows: warp 0's area holds the CTA's bitmap

	DevState * st = a.st;
	if (!a.countOnly && st->convergedAt >= 0) return;
	const uint32_t t = st->sweep;
	const ColT * __restrict__ cur = a.colorsOverride ? static_cast<const ColT *>(a.colorsOverride) : static_cast<const ColT *>(a.colors[t & 1]);
	ColT * __restrict__ nxt = static_cast<ColT *>(a.colors[(t + 1) & 1]);
	const float stayW = stay_weight<kDyn>(nCol, a.eps);
	const uint64_t polLast = make_policy_evict_last();
	unsigned long long accDirected = 0ull, accViol = 0ull;

	// ---------------- CTA rows: one bitmap for the CTA ----------------
	for (;;) {
		__syncthreads();
		if (tid == 0) { s_ctl[0] = atomicAdd(bn.counters + 2, 1u); s_ctl[1] = 0u; s_ctl[2] = 0u; }
		for (uint32_t w = tid; w < bmWords; w += kThreadsBin) s_bm[w] = 0u;
		__syncthreads();
		const uint32_t i = s_ctl[0];
		if (i >= bn.n[2]) break;
		const uint32_t lv = bn.list[2][i];
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
			if (prepare_wide<kDyn>(a, wa, t, nxt, a.vBegin + lv, lv, own, occ, s_ctl[1], s_ctl[2], stayW, accDirected, accViol, u, x))
				finish_wide(a, nxt, lv, own, walk_wide<kDyn>(a, wa, occ, u, x), true);
		}
	}

	// ---------------- warp rows: a batch of rows per warp, one bitmap each.  The edges of the whole batch are walked as ONE flat
	//                  range (lane -> (row, offset) by a binary search over the scanned row lengths), so the id loads and colour
	//                  gathers of consecutive rows overlap -- most warp rows of a power-law graph are only a few dozen entries long,
	//                  and row-by-row processing would pay two dependent memory latencies per row.  Then the lanes commit together ----
	{
		uint32_t * bms = reinterpret_cast<uint32_t *>(mineRaw);
		const uint32_t B = wa.batch, stride = wa.bmStride;
		uint32_t * s_same = bms + (size_t)B * stride;                 // [32]
		for (;;) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(bn.counters + 1, B);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= bn.n[1]) break;
			const uint32_t cntB = min(B, bn.n[1] - base);
			const bool valid = (uint32_t)lane < cntB;
			const uint32_t lvMine = valid ? bn.list[1][base + lane] : 0u;
			const uint32_t begMine = valid ? a.rowptr[lvMine] : 0u, endMine = valid ? a.rowptr[lvMine + 1] : 0u;
			const uint32_t ownMine = valid ? (uint32_t)cur[a.vBegin + lvMine] : 0xffffffffu;
			uint32_t P = endMine - begMine;                           // inclusive scan of the row lengths
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
					uint32_t j = 0;                                   // smallest j with P_j > f
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
				if (prepare_wide<kDyn>(a, wa, t, nxt, a.vBegin + lvMine, lvMine, ownMine, occ, s_same[lane], zn, stayW, accDirected, accViol, u, x))
					finish_wide(a, nxt, lvMine, ownMine, walk_wide<kDyn>(a, wa, occ, u, x), true);
			}
			__syncwarp();
		}
	}

	// ---------------- thread rows: colour lists; conflicting vertices park their walk, dense lanes drain the queue ----------------
	{
		__syncwarp();
		uint16_t * s_list = reinterpret_cast<uint16_t *>(mineRaw);                       // [kWideListCap][32]
		uint16_t * q_list = s_list + kWideListCap * 32u;                                 // [kWideListCap][kWideQueueCap]
		uint32_t * q_hdr = reinterpret_cast<uint32_t *>(q_list + kWideListCap * kWideQueueCap);   // lv, own, deg | u, x   [5][cap]
		uint32_t qn = 0;                                                                 // entries parked (warp-uniform)
		auto drain = [&](uint32_t first, uint32_t count) {
			if ((uint32_t)lane < count) {
				const uint32_t i = first + lane;
				const WideList occ{q_list + i, q_hdr[2 * kWideQueueCap + i], kWideQueueCap};
				const uint32_t lv = q_hdr[i], own = q_hdr[kWideQueueCap + i];
				const float u = __uint_as_float(q_hdr[3 * kWideQueueCap + i]), x = __uint_as_float(q_hdr[4 * kWideQueueCap + i]);
				finish_wide(a, nxt, lv, own, walk_wide<kDyn>(a, wa, occ, u, x), true);
			}
		};
		for (;;) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(bn.counters + 0, 32u);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= bn.n[0]) break;
			const bool valid = base + lane < bn.n[0];
			const uint32_t lv = valid ? bn.list[0][base + lane] : 0u;
			const uint32_t beg = valid ? a.rowptr[lv] : 0u, deg = valid ? (a.rowptr[lv + 1] - beg) : 0u;
			const uint32_t own = valid ? (uint32_t)cur[a.vBegin + lv] : 0u;
			uint16_t * mine = s_list + lane;
			uint32_t same = 0;
			for (uint32_t i = 0; i < deg; i += 4u) {
				uint32_t nb[4], c[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) nb[k] = (i + k < deg) ? __ldg(a.neighs + beg + i + k) : 0u;
#pragma unroll
				for (int k = 0; k < 4; ++k) c[k] = (i + k < deg) ? ld_color<ColT>(cur + nb[k], polLast) : 0xffffffffu;
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					if (i + k < deg) { MCMCB200_CHECK(c[k] < nCol && i + k < kWideListCap, st); same += (c[k] == own); mine[(i + k) * 32u] = (uint16_t)c[k]; }
				}
			}
			bool walk = false;
			float u = 0.0f, x = 0.0f;
			if (valid) {
				const WideList occ{mine, deg, 32u};
				walk = prepare_wide<kDyn>(a, wa, t, nxt, a.vBegin + lv, lv, own, occ, same, 0xffffffffu, stayW, accDirected, accViol, u, x);
			}
			const uint32_t wm = __ballot_sync(0xffffffffu, walk);
			const uint32_t nw = (uint32_t)__popc(wm);
			if (nw) {
				if (walk) {                                           // (qn < 32 here, nw <= 32: always room)
					const uint32_t slot = qn + (uint32_t)__popc(wm & ((1u << lane) - 1u));
					MCMCB200_CHECK(slot < kWideQueueCap, st);
					for (uint32_t j = 0; j < deg; ++j) q_list[j * kWideQueueCap + slot] = mine[j * 32u];
					q_hdr[slot] = lv; q_hdr[kWideQueueCap + slot] = own; q_hdr[2 * kWideQueueCap + slot] = deg;
					q_hdr[3 * kWideQueueCap + slot] = __float_as_uint(u); q_hdr[4 * kWideQueueCap + slot] = __float_as_uint(x);
				}
				qn += nw;
				__syncwarp();
				if (qn >= 32u) { drain(qn - 32u, 32u); __syncwarp(); qn -= 32u; }
			}
			__syncwarp();
		}
		if (qn) { drain(0u, qn); __syncwarp(); }
	}

	// ---- epilogue (same protocol as binned_sweep_kernel; class-size deltas went to the scratch directly) ----
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
