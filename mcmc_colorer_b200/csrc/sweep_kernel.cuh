// The fused MCMC sweep kernel (sm_100a).  One launch = one synchronous sweep C_t -> C_{t+1} over the owned
// vertex range, including the convergence counters and the colour-class-size update -- no host round trip.
//
// Reference functions replaced by this one kernel (paths relative to reference src/):
//   conflictCounter + sumReduction + host sum      graph_coloring/coloringMCMC_utils.cu:103-198
//   cudaMemset(colorsChecker, n*nCol)              graph_coloring/coloringMCMC_main.cu:174
//   D2H colouring + host histogram + H2D           graph_coloring/coloringMCMC_main.cu:211-214
//   genDynamicDistribution                         graph_coloring/coloringMCMC_utils.cu:64-70
//   selectStarColoring (UNIFORM)                   graph_coloring/coloringMCMC_standard.cu:9-82
//   selectStarColoringBalanceDynamic (DYNAMIC)     graph_coloring/coloringMCMC_balance.cu:79-143
//   CPU twin: count_free_colors / fill_p / extract_new_color / violation_count
//                                                  graph_coloring/coloringMCMC_CPU.cpp:328-528
//
// Work decomposition ("degree binning" happens inside the kernel, per tile, no preprocessing pass):
//   * persistent CTAs pull tiles of kTileV consecutive vertices from an atomic counter;
//   * PHASE 1 (edge-parallel, CTA-wide): the tile's contiguous slice of `neighs` is streamed with 256-bit
//     evict-first loads; every thread immediately gathers the colours of its 8 neighbours (evict-last hint,
//     narrow u8/u16 colours) and parks them in shared memory in CSR order.  Perfectly coalesced and load
//     balanced whatever the degree distribution;
//   * PHASE 2 builds each vertex' neighbour-colour occupancy bitmask from shared memory:
//       deg <= kLightMaxDeg  : one thread per vertex,
//       deg <= kCapEdges     : one warp per vertex (ballot-free OR + __reduce_or_sync),
//       deg >  kCapEdges     : whole CTA per vertex (hub), gathers straight from global;
//   * PHASE 3 (thread per vertex): conflict flag, free-colour count (popc), taboo gate, Philox draw,
//     sequential float32 CDF walk over the bitmask (bit-exact with the reference's running sum), colour write;
//   * epilogue: warp-shuffle + one atomic per CTA for the counters, shared-memory class-size deltas, and the
//     last CTA to finish applies the deltas, records the history and decides convergence on the device.
#pragma once
#include "device_utils.cuh"

namespace mcmcb200 {

constexpr int kThreads     = 256;   // threads per CTA
constexpr int kTileV       = 256;   // vertices per tile (one per thread in phase 3)
constexpr int kCapEdges    = 8192;  // neighbour colours staged per sub-tile
constexpr int kLightMaxDeg = 64;    // thread-per-vertex up to here, warp-per-vertex above
constexpr int kMaxColWords = 8;     // nCol <= 512 on the bitmask-in-registers path
constexpr int kMaxPeers    = 8;     // GPUs of one NVSwitch box whose colour replicas a sweep can store into directly

struct DevState {
	uint32_t sweep;          // index t of the current colouring C_t
	int32_t  convergedAt;    // -1, or t once the selected count of C_t was <= z
	uint32_t ticket;         // CTAs finished in the running launch
	uint32_t tileCounter;    // dynamic tile scheduler
	uint32_t convergence;    // 0: violating vertices, 1: conflicting edges
	uint32_t countsSweep;    // colouring index lastDirected/lastViol describe (0xffffffff: none)
	uint64_t z;              // threshold
	uint64_t lastDirected;   // sum_v #{u in N(v): C[u]==C[v]}  == 2 * conflicting edges
	uint64_t lastViol;       // violating vertices
	uint32_t errorFlag;      // sticky device-side error: 1 colour out of range, 2 blocked sweep aborted (producer missing), 3 a peer GPU never arrived
	uint32_t xseq;           // multi-GPU: collective finalize steps performed so far (parity selects the exchange slot)
	uint32_t emitNow;        // tail cutting: the running sweep appends its violating vertices to SweepArgs::violList
	uint32_t violListSweep;  // colouring index that list describes (0xffffffff: none / overflowed)
	uint32_t violListCount;  // entries of the list
	uint32_t pad2;
};

struct SweepArgs {
	const uint32_t * rowptr;     // [nLocal+1], rowptr[0]==0
	const uint32_t * neighs;     // owned rows, GLOBAL neighbour ids, 32-byte aligned
	uint32_t nLocal, vBegin, nGlobal, nCol;
	uint32_t numTiles;
	float    eps;
	uint32_t tabooIter;
	uint32_t proposal;
	uint64_t seed;
	void *   colors[2];          // ColT[nGlobal (padded)], colouring t lives in colors[t&1]
	const void * colorsOverride; // count-only: evaluate this colouring instead
	uint16_t * taboo;            // [nLocal] or nullptr
	const float * tape;          // [tapeSweeps][nGlobal] or nullptr
	uint32_t tapeBase;           // sweep index of tape row 0
	DevState * st;
	unsigned long long * scratch; // [2 + nCol]: directed conflicts, violating vertices, class-size deltas (two's complement)
	unsigned long long * hist[2]; // class sizes of colouring t in hist[t&1]
	unsigned long long * history; // [historyCap][2]
	uint32_t historyCap;
	uint32_t countOnly;          // 1: counters only, nothing is modified
	unsigned long long * countOut; // count-only result [2]
	uint32_t fuseFinalize;       // last CTA runs finalize_sweep
	uint32_t noEarlyStop;        // keep sweeping even when C_t is already proper (replay / benchmarking)
	void *   peerColors[2][kMaxPeers]; // fused exchange: every rank's two colour buffers (IPC-mapped; own entries = local)
	unsigned long long * peerXchg[kMaxPeers]; // fused exchange: every rank's counter-exchange block (layout: xchg_* below)
	uint32_t nPeers;             // 0: no fused exchange (single GPU, or NCCL all-gather by the caller)
	uint32_t myRank;             // index of this handle in the peer tables
	unsigned long long * dbgMasks; // optional [nLocal][W]
	uint32_t * dbgSame;          // optional [nLocal]
	// tail cutting (params.tailcut): when the chain gets close to the threshold z the sweeps emit the violating vertices of the
	// colouring they evaluate, so that the repair never has to rescan the graph (tailcut_kernel.cuh)
	uint32_t * violList;         // [violCap] global vertex ids, or nullptr
	uint32_t * violCount;        // entries appended by the running launch
	uint32_t violCap;
	uint32_t forceEmit;          // emit regardless of DevState::emitNow (the count-only pass mcmcb200_tailcut falls back to)
	unsigned long long emitThreshold; // finalize: the NEXT sweep emits iff this colouring has at most so many violating vertices
};

__host__ __device__ inline size_t sweep_smem_bytes(uint32_t nCol, int W, int colBytes) {
	size_t b = 0;
	b += sizeof(uint32_t) * (kTileV + 4);                 // s_rp
	b += sizeof(float) * (size_t)((nCol + 1 + 3) & ~3u);  // s_S
	b += sizeof(float) * (size_t)((nCol + 3) & ~3u);      // s_dist
	b += sizeof(int) * (size_t)((nCol + 3) & ~3u);        // s_hist
	b += sizeof(uint32_t) * kTileV;                       // s_same
	b += sizeof(uint32_t) * 8;                            // s_ctl
	b += sizeof(uint16_t) * kTileV;                       // s_heavy
	b = (b + 15) & ~(size_t)15;
	b += sizeof(unsigned long long) * (size_t)kTileV * W; // s_mask
	if (W > 2) b += sizeof(uint32_t) * (kThreads / 32) * 2 * W;   // s_wm: per-warp mask accumulators of the wide palettes
	b += (size_t)colBytes * (kCapEdges + 16);             // s_col
	b = (b + 15) & ~(size_t)15;
	b += (size_t)kTileV * (8 * W + 16);                   // deferred CDF walks of one sub-tile (mask, vertex/own, draw/weight)
	return (b + 15) & ~(size_t)15;
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU, fused exchange: the per-sweep all-reduce of {directed conflicts, violating vertices, class-size deltas} and the
// inter-rank barrier, done by the sweep kernel itself over NVLink -- no NCCL call, no host in the loop.
// Every rank owns an exchange block (IPC-mapped by all peers):  [0..1] arrival counters, then two slots of 2 + nCol sums.
// The LAST CTA of a rank's sweep adds its local counters into the current slot of EVERY rank (system-scope reductions on peer
// memory), fences, bumps every rank's arrival counter, waits until all nPeers ranks have arrived at its own block, and only then
// runs finalize_sweep_device on the global sums.  Because nobody passes this point before everybody's last CTA got here, it is
// also the barrier that orders sweep t+1 after every peer's colour stores of sweep t.  Slots alternate with st->xseq; a slot is
// re-used two collectives later, which a rank can only reach after all ranks finished the collective in between.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t xchg_words(uint32_t nCol) { return 2 + 2 * (size_t)(nCol + 2); }
__device__ __forceinline__ unsigned long long * xchg_slot(unsigned long long * block, uint32_t nCol, uint32_t par) { return block + 2 + (size_t)par * (nCol + 2); }

__device__ __forceinline__ void cross_rank_reduce(const SweepArgs & a) {
	DevState * st = a.st;
	const uint32_t par = st->xseq & 1u, nw = a.nCol + 2u;
	for (uint32_t k = threadIdx.x; k < nw; k += blockDim.x) {
		const unsigned long long v = __ldcg(a.scratch + k);
		if (v) for (uint32_t r = 0; r < a.nPeers; ++r) atomicAdd_system(xchg_slot(a.peerXchg[r], a.nCol, par) + k, v);
	}
	__threadfence_system();                               // my sums (and, through the ticket chain, every CTA's colour stores) before my arrival
	__syncthreads();
	if (threadIdx.x == 0) {
		for (uint32_t r = 0; r < a.nPeers; ++r) atomicAdd_system(a.peerXchg[r] + par, 1ull);
		unsigned long long * mine = a.peerXchg[a.myRank] + par;
		unsigned long long v;
		long long t0 = clock64();
		for (uint32_t spins = 0;; ++spins) {
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
			if (v >= (unsigned long long)a.nPeers) break;
			if ((spins & 1023u) == 1023u && clock64() - t0 > 20000000000ll) { st->errorFlag = 3u; break; }   // ~10 s: a peer never launched
			__nanosleep(200);
		}
	}
	__syncthreads();
	unsigned long long * slot = xchg_slot(a.peerXchg[a.myRank], a.nCol, par);
	for (uint32_t k = threadIdx.x; k < nw; k += blockDim.x) {
		unsigned long long v;
		asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(slot + k) : "memory");
		a.scratch[k] = v;                                 // the global sums take the place of the local ones
		slot[k] = 0ull;                                   // slot is clean for the collective after the next
	}
	if (threadIdx.x == 0) { a.peerXchg[a.myRank][par] = 0ull; st->xseq = st->xseq + 1u; }
	__threadfence();
	__syncthreads();
}

// finalize: executed by ONE CTA after all counters of the sweep are in `scratch` (fused: the last CTA of the
// sweep kernel; multi-GPU: a 1-CTA kernel after the cross-rank all-reduce).
__device__ __forceinline__ void finalize_sweep_device(const SweepArgs & a) {
	DevState * st = a.st;
	const unsigned long long directed = __ldcg(a.scratch + 0);
	const unsigned long long viol = __ldcg(a.scratch + 1);
	const uint32_t t = st->sweep;
	// errorFlag 2: the blocked sweep was aborted (pass B never saw pass A's output, blocked_sweep.cuh wait_part_ready) -- nothing
	// of this launch counts: no history entry, no class sizes, the sweep index does not advance.  The host retries (read_state).
	const bool aborted = *reinterpret_cast<volatile uint32_t *>(&st->errorFlag) == 2u;
	if (aborted) {
		// fall through to the scratch reset only
	} else if (a.countOut != nullptr) {                // count-only on an arbitrary colouring
		if (threadIdx.x == 0) { a.countOut[0] = directed; a.countOut[1] = viol; }
	} else {
		const unsigned long long metric = (st->convergence == 0) ? viol : (directed >> 1);
		const bool conv = (metric <= st->z) && !a.noEarlyStop;
		if (!a.countOnly && !conv) {
			const unsigned long long * hc = a.hist[t & 1];
			unsigned long long * hn = a.hist[(t + 1) & 1];
			for (uint32_t k = threadIdx.x; k < a.nCol; k += blockDim.x) hn[k] = hc[k] + __ldcg(a.scratch + 2 + k);
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			if (t < a.historyCap) { a.history[2 * (size_t)t] = directed >> 1; a.history[2 * (size_t)t + 1] = viol; }
			st->lastDirected = directed; st->lastViol = viol; st->countsSweep = t;
			if (a.violList != nullptr) {                   // the violating vertices of C_t, if this launch emitted them
				const uint32_t cnt = __ldcg(a.violCount);
				const bool emitted = a.forceEmit || st->emitNow;
				if (a.countOnly || conv) {                 // C_t stays the current colouring: keep the list for the repair
					st->violListSweep = (emitted && cnt <= a.violCap) ? t : 0xffffffffu;
					st->violListCount = cnt;
				} else {
					st->violListSweep = 0xffffffffu;
					st->emitNow = (viol <= a.emitThreshold) ? 1u : 0u;
				}
				*a.violCount = 0u;
			}
			if (!a.countOnly) {
				if (conv) st->convergedAt = (int32_t)t;
				else st->sweep = t + 1;
			}
		}
	}
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < a.nCol + 2; k += blockDim.x) a.scratch[k] = 0ull;
	if (threadIdx.x == 0) { st->ticket = 0; st->tileCounter = 0; }
}

// (a chain that already stopped at its threshold: the sweep launch before this was a no-op and so is its finalize)
__global__ void finalize_kernel(SweepArgs a) {
	if (!a.countOnly && a.st->convergedAt >= 0) return;
	finalize_sweep_device(a);
}

// ---------------------------------------------------------------------------------------------
// Sequential CDF walks.  The reference selects the first colour whose float32 running sum exceeds the draw
// (coloringMCMC_CPU.cpp:510-514; ">=" in coloringMCMC_balance.cu:123-136).  The sum is order dependent, so it is
// reproduced addend by addend with __fadd_rn (never contracted); the loops below are kept as tight as possible
// because they are the kernel's instruction hot spot (65 % of the warp instructions before this rewrite).
// ---------------------------------------------------------------------------------------------
// Per-CTA proposal table in shared memory (nCol floats).
//   DYNAMIC: dist[k] = (1 - hist[k]/n) / (nCol-1)                        (genDynamicDistribution, coloringMCMC_utils.cu:69)
//   UNIFORM: the same storage holds freeW[Zn] = (1 - eps*Zn) / (nCol-Zn), the weight of a free colour of a vertex with Zn
//            occupied colours (coloringMCMC_CPU.cpp:416) -- one table lookup instead of a correctly rounded division per
//            conflicting vertex; identical operations, identical bits.
template <bool kDyn>
__device__ __forceinline__ void fill_proposal_table(const SweepArgs & a, uint32_t t, float * s_dist, int tid, int nThreads) {
	const uint32_t nCol = a.nCol;
	if (kDyn) {
		const unsigned long long * hc = a.hist[t & 1];
		const float nF = __uint2float_rn(a.nGlobal), dF = __uint2float_rn(nCol - 1u);
		for (uint32_t k = tid; k < nCol; k += nThreads)
			s_dist[k] = __fdiv_rn(__fsub_rn(1.0f, __fdiv_rn(__uint2float_rn((uint32_t)hc[k]), nF)), dF);
	} else {
		for (uint32_t k = tid; k < nCol; k += nThreads)     // Zn = k occupied, Zp = nCol - k >= 1 free
			s_dist[k] = __fdiv_rn(__fsub_rn(1.0f, __fmul_rn(a.eps, __uint2float_rn(k))), __uint2float_rn(nCol - k));
	}
}

// weight of the own colour in the "stay" distribution, 1 - (nCol-1)*eps.
//   UNIFORM follows the CPU colourer's x86-64 build: product and difference are two roundings (coloringMCMC_CPU.cpp:406,474);
//   DYNAMIC follows the reference GPU kernel as nvcc compiles it (default -fmad=true): ONE fused multiply-add,
//   `FFMA R, -(float)(nCol-1), eps, 1` in the SASS of selectStarColoringBalanceDynamic (coloringMCMC_balance.cu:88,132) built
//   for sm_100a with the reference's flags; the GPU test suite replays that kernel against this one.  The two differ for 17 palettes >= 9499 colours at eps = 1e-8 and from 89 colours on at eps = 1e-4.
template <bool kDyn>
__device__ __forceinline__ float stay_weight(uint32_t nCol, float eps) {
	const float k = __uint2float_rn(nCol - 1u);
	return kDyn ? __fmaf_rn(-k, eps, 1.0f) : __fsub_rn(1.0f, __fmul_rn(k, eps));
}

// "stay" distribution (all colours eps, own colour 1-(nCol-1)eps) when the draw fell into an epsilon tail:
// probability ~ nCol*eps per vertex, kept out of line.
template <bool kDyn>
__device__ __noinline__ uint32_t walk_stay(uint32_t nCol, uint32_t own, float eps, float stayW, float u) {
	float cdf = 0.0f;
	for (uint32_t k = 0; k < nCol; ++k) {
		cdf = __fadd_rn(cdf, k == own ? stayW : eps);
		if (kDyn ? (cdf >= u) : (cdf > u)) return k;
	}
	return nCol - 1u;                                       // overflow contract: clamp to nCol-1
}

// conflicting vertex with free colours: occupied colours weigh eps, free ones freeW (UNIFORM) or dist[k]+r (DYNAMIC).
// The running sum is non-decreasing (all addends >= 0; the DYNAMIC corner r < 0 takes the one-step loop), so the
// stop test is made once per 4 colours and the first crossing inside the block is located afterwards: ~4.5 instead of
// 10 instructions per colour, with exactly the reference's sequence of float32 partial sums.
template <int W, bool kDyn>
__device__ __forceinline__ uint32_t walk_conflicting(const unsigned long long (&m)[W], uint32_t nCol, float eps,
                                                     float freeW, float r, const float * s_dist, float u) {
	float cdf = 0.0f;
	const bool blocks = !kDyn || r >= 0.0f;
#pragma unroll
	for (int h = 0; h < 2 * W; ++h) {
		if ((uint32_t)(h * 32) < nCol) {
			uint32_t bits = (h & 1) ? (uint32_t)(m[h >> 1] >> 32) : (uint32_t)m[h >> 1];
			const uint32_t lim = min(32u, nCol - (uint32_t)(h * 32));
			uint32_t b = 0;
			if (blocks) {
				for (; b + 4u <= lim; b += 4u) {
					float q0, q1, q2, q3;
					if (kDyn) {
						const float * d = s_dist + h * 32 + b;
						q0 = (bits & 1u) ? eps : __fadd_rn(d[0], r); q1 = (bits & 2u) ? eps : __fadd_rn(d[1], r);
						q2 = (bits & 4u) ? eps : __fadd_rn(d[2], r); q3 = (bits & 8u) ? eps : __fadd_rn(d[3], r);
					} else {
						q0 = (bits & 1u) ? eps : freeW; q1 = (bits & 2u) ? eps : freeW;
						q2 = (bits & 4u) ? eps : freeW; q3 = (bits & 8u) ? eps : freeW;
					}
					const float c1 = __fadd_rn(cdf, q0), c2 = __fadd_rn(c1, q1), c3 = __fadd_rn(c2, q2), c4 = __fadd_rn(c3, q3);
					if (kDyn ? (c4 >= u) : (c4 > u)) {
						const uint32_t first = (kDyn ? (c1 >= u) : (c1 > u)) ? 0u : (kDyn ? (c2 >= u) : (c2 > u)) ? 1u
						                     : (kDyn ? (c3 >= u) : (c3 > u)) ? 2u : 3u;
						return (uint32_t)(h * 32) + b + first;
					}
					cdf = c4;
					bits >>= 4;
				}
			}
			for (; b < lim; ++b) {
				float q;
				if (kDyn) q = (bits & 1u) ? eps : __fadd_rn(s_dist[h * 32 + b], r);
				else q = (bits & 1u) ? eps : freeW;
				bits >>= 1;
				cdf = __fadd_rn(cdf, q);
				if (kDyn ? (cdf >= u) : (cdf > u)) return (uint32_t)(h * 32) + b;
			}
		}
	}
	return nCol - 1u;                                       // overflow contract: clamp to nCol-1
}

// DYNAMIC: r = (sum over the occupied colours, ascending, of dist[c] - eps) / Zp   (coloringMCMC_balance.cu:104-109,124)
template <int W>
__device__ __forceinline__ float dynamic_reminder(const unsigned long long (&m)[W], uint32_t Zp, float eps, const float * s_dist) {
	float rem = 0.0f;
#pragma unroll
	for (int w = 0; w < W; ++w) {
		unsigned long long bits = m[w];
		while (bits) {
			const int b = __ffsll((long long)bits) - 1;
			bits &= bits - 1ull;
			rem = __fadd_rn(rem, __fsub_rn(s_dist[w * 64 + b], eps));
		}
	}
	return __fdiv_rn(rem, __uint2float_rn(Zp));
}

// Deferred CDF walks: conflicting vertices are parked in shared memory and walked later by densely packed lanes
// (in the first sweeps ~1/3 of the vertices conflict, later almost none: without the queue nearly every warp would
// run the full walk for a handful of active lanes).
template <int W>
struct WalkQueue {
	uint32_t * count;               // shared counter
	uint32_t cap;
	unsigned long long * mask;      // [cap][W]
	uint32_t * lvOwn;               // [cap][2]: local vertex index, own colour
	float * uw;                     // [cap][2]: draw, freeW (UNIFORM) or r (DYNAMIC)
	static __host__ __device__ constexpr size_t bytes_per_entry() { return 8 * W + 16; }
};

// colour write + taboo + class-size deltas of a vertex whose new colour is known
// nxtTile != nullptr: the new colour goes to the tile's shared-memory staging row (index lv - tileV0) and is written out
// coalesced -- to the local replica and, in the fused multi-GPU exchange, to every peer's -- once the tile is finished
template <typename ColT>
__device__ __forceinline__ void finish_vertex(const SweepArgs & a, ColT * __restrict__ nxt, uint32_t lv, uint32_t myOwn, uint32_t newc,
                                              int * s_hist, bool touchTaboo, ColT * nxtTile = nullptr, uint32_t tileV0 = 0) {
	MCMCB200_CHECK(newc < a.nCol && myOwn < a.nCol && lv < a.nLocal, a.st);
	if (touchTaboo && a.tabooIter) a.taboo[lv] = (uint16_t)((newc == myOwn) ? a.tabooIter : 0u);   // coloringMCMC_CPU.cpp:526
	if (nxtTile) nxtTile[lv - tileV0] = (ColT)newc;
	else nxt[a.vBegin + lv] = (ColT)newc;
	if (newc != myOwn) { atomicAdd(&s_hist[myOwn], -1); atomicAdd(&s_hist[newc], 1); }
}

// ---------------------------------------------------------------------------------------------
// PHASE 3 for one vertex (shared by the direct-gather kernel and the source-blocked kernel): conflict flag, free
// colour count, taboo gate, draw, proposal, colour write, class-size deltas.
//   v = global vertex id, lv = local (owned) index, m = occupancy mask, same = #neighbours with v's colour.
// ---------------------------------------------------------------------------------------------
// kPlain: the instance for the plain production sweep -- the caller guarantees no replay tape, no taboo, no debug masks and not a
// count-only pass, so their (warp-uniform, but issued per vertex) tests are compiled out: 5 % of pass B's instructions on config 3.
template <int W, typename ColT, bool kDyn, bool kPlain = false>
__device__ __forceinline__ void commit_vertex(const SweepArgs & a, uint32_t t, ColT * __restrict__ nxt, uint32_t v, uint32_t lv,
                                              uint32_t myOwn, const unsigned long long (&m)[W], uint32_t same,
                                              const float * s_S, const float * s_dist, int * s_hist, float stayW,
                                              unsigned long long & accDirected, unsigned long long & accViol,
                                              const WalkQueue<W> * queue = nullptr, ColT * nxtTile = nullptr, uint32_t tileV0 = 0,
                                              const uint32_t * drawTab = nullptr) {
	constexpr bool isDyn = kDyn;
	const uint32_t nCol = a.nCol;
	const float eps = a.eps;
	const bool viol = same > 0u;                              // occ[C[v]]  (violation_count, coloringMCMC_CPU.cpp:342-348)
	accDirected += same;
	accViol += viol ? 1ull : 0ull;
	if (!kPlain && a.dbgMasks) {
#pragma unroll
		for (int w = 0; w < W; ++w) a.dbgMasks[(size_t)lv * W + w] = m[w];
		a.dbgSame[lv] = same;
	}
	if (viol && a.violList != nullptr) {                      // tail cutting: remember who violates (one atomic per group of lanes)
		if (a.forceEmit || __ldcg(&a.st->emitNow)) {
			const unsigned act = __activemask();
			const int leader = __ffs((int)act) - 1;
			uint32_t base = 0;
			if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(a.violCount, (uint32_t)__popc(act));
			base = __shfl_sync(act, base, leader);
			const uint32_t idx = base + (uint32_t)__popc(act & ((1u << (threadIdx.x & 31)) - 1u));
			if (idx < a.violCap) a.violList[idx] = v;
		}
	}
	if (!kPlain && a.countOnly) return;
	uint32_t newc = myOwn;
	bool tabooed = false;
	if (!kPlain && a.tabooIter) {                             // TABOO gate, coloringMCMC_CPU.cpp:496-501
		const uint32_t tb = a.taboo[lv];
		if (tb > 0u) { a.taboo[lv] = (uint16_t)(tb - 1u); tabooed = true; }
	}
	if (!tabooed) {
		uint32_t Zn = 0;
#pragma unroll
		for (int w = 0; w < W; ++w) Zn += __popcll(m[w]);
		const uint32_t Zp = nCol - Zn;                        // free colours (count_free_colors :382)
		if (isDyn && Zp == 0u) {
			newc = myOwn;                                     // coloringMCMC_balance.cu:111-115: no draw, taboo untouched
		} else {
			float u;
			if (!kPlain && a.tape) u = a.tape[(size_t)(t - a.tapeBase) * a.nGlobal + v];
			else if (drawTab) u = draw_to_uniform(drawTab[lv - tileV0], isDyn);      // the tile's Philox words, one call per 4 vertices
			else u = draw_to_uniform(philox_draw(a.seed, t + 1u, v, 0u), isDyn);
			const bool stay = !viol || Zp == 0u;              // :472-478 / :402-411
			if (stay) {
				// fast path: everything before `own` weighs eps (table S), own weighs stayW
				const float sOwn = s_S[myOwn];
				const float tOwn = __fadd_rn(sOwn, stayW);
				const bool notBefore = isDyn ? (sOwn < u) : (sOwn <= u);
				const bool hit = isDyn ? (tOwn >= u) : (tOwn > u);
				newc = (notBefore && hit) ? myOwn : walk_stay<isDyn>(nCol, myOwn, eps, stayW, u);
			} else {
				float freeW = 0.0f, r = 0.0f;
				if (!isDyn) {                                 // (1 - eps*Zv) / Zvcomp, coloringMCMC_CPU.cpp:416
					freeW = s_dist[Zn];                       // fill_proposal_table: (1 - eps*Zn) / Zp, Zp >= 1 here
				}
				if (queue != nullptr) {
					// one atomic per converged group of lanes instead of one per lane
					const unsigned act = __activemask();
					const int leader = __ffs((int)act) - 1;
					const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
					uint32_t qbase = 0;
					if ((int)(threadIdx.x & 31) == leader) qbase = atomicAdd(queue->count, (uint32_t)__popc(act));
					qbase = __shfl_sync(act, qbase, leader);
					const uint32_t qi = qbase + (uint32_t)__popc(act & lt);
					if (qi < queue->cap) {                    // park the walk; drained by dense lanes (drain_walk_queue)
						MCMCB200_CHECK(Zn < nCol && lv < a.nLocal, a.st);
#pragma unroll
						for (int w = 0; w < W; ++w) queue->mask[(size_t)qi * W + w] = m[w];
						queue->lvOwn[2 * qi] = lv; queue->lvOwn[2 * qi + 1] = myOwn;
						queue->uw[2 * qi] = u; queue->uw[2 * qi + 1] = freeW;           // (DYNAMIC: r is computed by the dense lanes of the drain)
						return;
					}
				}
				if (isDyn) r = dynamic_reminder<W>(m, nCol - Zn, eps, s_dist);
				newc = walk_conflicting<W, isDyn>(m, nCol, eps, freeW, r, s_dist, u);
			}
			finish_vertex<ColT>(a, nxt, lv, myOwn, newc, s_hist, true, nxtTile, tileV0);
			return;
		}
	}
	finish_vertex<ColT>(a, nxt, lv, myOwn, newc, s_hist, false, nxtTile, tileV0);
}

// entries [first, first+count) of a queue, one per lane (called by the owning warp; count <= 32)
template <int W, typename ColT, bool kDyn>
__device__ __forceinline__ void drain_walk_queue(const SweepArgs & a, ColT * __restrict__ nxt, const WalkQueue<W> & q, uint32_t first,
                                                 uint32_t count, const float * s_dist, int * s_hist, int lane,
                                                 ColT * nxtTile = nullptr, uint32_t tileV0 = 0) {
	if ((uint32_t)lane < count) {
		const uint32_t i = first + lane;
		unsigned long long m[W];
#pragma unroll
		for (int w = 0; w < W; ++w) m[w] = q.mask[(size_t)i * W + w];
		const uint32_t lv = q.lvOwn[2 * i], own = q.lvOwn[2 * i + 1];
		const float u = q.uw[2 * i];
		float x = q.uw[2 * i + 1];
		if (kDyn) {                                               // the reminder of a parked vertex: here the lanes are dense
			uint32_t Zn = 0;
#pragma unroll
			for (int w = 0; w < W; ++w) Zn += __popcll(m[w]);
			x = dynamic_reminder<W>(m, a.nCol - Zn, a.eps, s_dist);
		}
		const uint32_t newc = walk_conflicting<W, kDyn>(m, a.nCol, a.eps, x, x, s_dist, u);
		finish_vertex<ColT>(a, nxt, lv, own, newc, s_hist, true, nxtTile, tileV0);
	}
}

template <int W, typename ColT, bool kDyn>
__global__ void __launch_bounds__(kThreads, (W <= 2 ? 4 : 2))
sweep_kernel(const SweepArgs a) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t nCol = a.nCol;
	uint32_t * s_rp   = reinterpret_cast<uint32_t *>(smem_raw);
	float *    s_S    = reinterpret_cast<float *>(s_rp + kTileV + 4);
	float *    s_dist = s_S + ((nCol + 1 + 3) & ~3u);
	int *      s_hist = reinterpret_cast<int *>(s_dist + ((nCol + 3) & ~3u));
	uint32_t * s_same = reinterpret_cast<uint32_t *>(s_hist + ((nCol + 3) & ~3u));
	uint32_t * s_ctl  = s_same + kTileV;
	uint16_t * s_heavy = reinterpret_cast<uint16_t *>(s_ctl + 8);
	size_t off = (size_t)(reinterpret_cast<unsigned char *>(s_heavy + kTileV) - smem_raw);
	off = (off + 15) & ~(size_t)15;
	unsigned long long * s_mask = reinterpret_cast<unsigned long long *>(smem_raw + off);
	// Wide palettes (W > 2, more than 128 colours): a mask held in W 64-bit registers costs a W-way select per edge (~5 W
	// instructions), so the masks are accumulated in shared memory instead: per-thread rows of 32-bit words, transposed
	// (word w of slot s at s_m32[w * kTileV + s]: conflict free), and per-warp accumulators s_wm for the rows a warp or the
	// whole CTA shares (one shared-memory atomic OR per edge).
	constexpr bool kWide = W > 2;
	uint32_t * s_m32 = reinterpret_cast<uint32_t *>(s_mask);
	uint32_t * s_wm = reinterpret_cast<uint32_t *>(s_mask + (size_t)kTileV * W);
	ColT * s_col = reinterpret_cast<ColT *>(reinterpret_cast<unsigned char *>(s_mask + (size_t)kTileV * W) + (kWide ? sizeof(uint32_t) * (kThreads / 32) * 2 * W : 0));
	auto load_wide_mask = [&](uint32_t slot, unsigned long long (&mm)[W]) {
#pragma unroll
		for (int w = 0; w < W; ++w)
			mm[w] = (unsigned long long)s_m32[(2 * w) * kTileV + slot] | ((unsigned long long)s_m32[(2 * w + 1) * kTileV + slot] << 32);
	};

	// deferred CDF walks: conflicting vertices are parked here in phase 3 and walked by densely packed threads at the end of
	// the sub-tile (a walk is ~5 instructions per colour; inline it would run with a handful of active lanes per warp)
	WalkQueue<W> wq{};
	{
		size_t qoff = (size_t)(reinterpret_cast<unsigned char *>(s_col + kCapEdges + 16) - smem_raw);
		qoff = (qoff + 15) & ~(size_t)15;
		wq.count = s_ctl + 3;
		wq.cap = kTileV;
		wq.mask = reinterpret_cast<unsigned long long *>(smem_raw + qoff);
		wq.lvOwn = reinterpret_cast<uint32_t *>(wq.mask + (size_t)kTileV * W);
		wq.uw = reinterpret_cast<float *>(wq.lvOwn + 2 * kTileV);
	}
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	DevState * st = a.st;
	if (!a.countOnly && st->convergedAt >= 0) return;      // converged earlier in this batch of launches: no-op

	const uint32_t t = st->sweep;
	const ColT * __restrict__ cur = a.colorsOverride ? static_cast<const ColT *>(a.colorsOverride)
	                                                 : static_cast<const ColT *>(a.colors[t & 1]);
	ColT * __restrict__ nxt = static_cast<ColT *>(a.colors[(t + 1) & 1]);
	const float eps = a.eps;
	// "stay" weight 1 - (nCol-1)*eps, two roundings like the reference's x86 build (coloringMCMC_CPU.cpp:406,474)
	const float stayW = stay_weight<kDyn>(nCol, eps);
	const uint64_t polLast = make_policy_evict_last();

	// ---- prologue: per-CTA tables ----
	for (uint32_t k = tid; k < nCol; k += kThreads) s_hist[k] = 0;
	if (tid == 0) {   // S[k] = eps added k times sequentially (the running sum over k epsilon-weighted colours)
		float s = 0.0f; s_S[0] = 0.0f;
		for (uint32_t k = 0; k < nCol; ++k) { s = __fadd_rn(s, eps); s_S[k + 1] = s; }
	}
	if (!a.countOnly) fill_proposal_table<kDyn>(a, t, s_dist, tid, kThreads);
	if (tid == 0) *wq.count = 0u;
	unsigned long long accDirected = 0ull, accViol = 0ull;

	// ---- persistent tile loop ----
	for (;;) {
		__syncthreads();
		if (tid == 0) s_ctl[0] = atomicAdd(&st->tileCounter, 1u);
		__syncthreads();
		const uint32_t tile = s_ctl[0];
		if (tile >= a.numTiles) break;
		const uint32_t v0 = tile * kTileV;
		const uint32_t nv = min((uint32_t)kTileV, a.nLocal - v0);
		for (uint32_t i = tid; i <= nv; i += kThreads) s_rp[i] = a.rowptr[v0 + i];
		const bool haveV = (uint32_t)tid < nv;
		const uint32_t gv = a.vBegin + v0 + tid;                   // global vertex id of this thread's slot
		const uint32_t own = haveV ? (uint32_t)cur[gv] : 0u;
		__syncthreads();
		const uint32_t myBeg = haveV ? s_rp[tid] : 0u;
		const uint32_t deg = haveV ? (s_rp[tid + 1] - myBeg) : 0u;

		uint32_t done = 0;
		while (done < nv) {
			const uint32_t base = s_rp[done];
			const bool fits = haveV && (uint32_t)tid >= done && (s_rp[tid + 1] - base) <= (uint32_t)kCapEdges;
			const uint32_t cnt = __syncthreads_count(fits);

			unsigned long long m[W];
#pragma unroll
			for (int w = 0; w < W; ++w) m[w] = 0ull;
			uint32_t same = 0;
			bool mine;

			if (cnt == 0) {
				// ---------------- hub: one vertex, whole CTA, no staging ----------------
				const uint32_t hubSlot = done;
				const uint32_t hOwn = (uint32_t)cur[a.vBegin + v0 + hubSlot];
				const uint32_t e0 = base, e1 = s_rp[hubSlot + 1];
				if (kWide) { if (tid < 2 * W) s_m32[tid * kTileV + hubSlot] = 0u; if (lane < 2 * W) s_wm[warp * 2 * W + lane] = 0u; }
				else if (tid < W) s_mask[(size_t)hubSlot * W + tid] = 0ull;
				if (tid == 0) s_same[hubSlot] = 0u;
				__syncthreads();
				for (uint32_t e = e0 + tid; e < e1; e += kThreads) {
					const uint32_t c = ld_color<ColT>(cur + a.neighs[e], polLast);
					same += (c == hOwn);
					if (kWide) atomicOr(&s_wm[warp * 2 * W + (c >> 5)], 1u << (c & 31u));
					else {
#pragma unroll
						for (int w = 0; w < W; ++w) m[w] |= ((int)(c >> 6) == w) ? (1ull << (c & 63u)) : 0ull;
					}
				}
				if (kWide) {
					__syncwarp();
					if (lane < 2 * W) { const uint32_t r = s_wm[warp * 2 * W + lane]; if (r) atomicOr(&s_m32[lane * kTileV + hubSlot], r); }
				} else {
#pragma unroll
					for (int w = 0; w < W; ++w) {
						const unsigned long long r = warp_reduce_or64(m[w]);
						if (lane == 0 && r) atomicOr(&s_mask[(size_t)hubSlot * W + w], r);
					}
				}
				same = __reduce_add_sync(0xffffffffu, same);
				if (lane == 0 && same) atomicAdd(&s_same[hubSlot], same);
				__syncthreads();
				mine = (uint32_t)tid == hubSlot;
				if (mine) {
					if (kWide) load_wide_mask(hubSlot, m);
					else {
#pragma unroll
						for (int w = 0; w < W; ++w) m[w] = s_mask[(size_t)hubSlot * W + w];
					}
					same = s_same[hubSlot];
				}
				done += 1;
			} else {
				// ---------------- PHASE 1: stream neighs, gather colours into shared memory ----------------
				const uint32_t e0 = base, e1 = s_rp[done + cnt];
				const uint32_t ea = e0 & ~7u;                       // 32-byte aligned start of the stream
				const uint32_t nOct = (e1 - ea + 7u) >> 3;
				for (uint32_t o = tid; o < nOct; o += 2 * kThreads) {
					const uint32_t o2 = o + kThreads;
					const bool has2 = o2 < nOct;
					const U32x8 nbA = ld_stream_256(a.neighs + ea + 8u * o);
					U32x8 nbB;
					if (has2) nbB = ld_stream_256(a.neighs + ea + 8u * o2);
					uint32_t cA[8], cB[8];
#pragma unroll
					for (int j = 0; j < 8; ++j)
						cA[j] = (ea + 8u * o + j < e1) ? ld_color<ColT>(cur + nbA.v[j], polLast) : 0u;
					if (has2) {
#pragma unroll
						for (int j = 0; j < 8; ++j)
							cB[j] = (ea + 8u * o2 + j < e1) ? ld_color<ColT>(cur + nbB.v[j], polLast) : 0u;
					}
					if (sizeof(ColT) == 1) {
						uint2 pk;
						pk.x = cA[0] | (cA[1] << 8) | (cA[2] << 16) | (cA[3] << 24);
						pk.y = cA[4] | (cA[5] << 8) | (cA[6] << 16) | (cA[7] << 24);
						*reinterpret_cast<uint2 *>(s_col + 8u * o) = pk;
						if (has2) {
							pk.x = cB[0] | (cB[1] << 8) | (cB[2] << 16) | (cB[3] << 24);
							pk.y = cB[4] | (cB[5] << 8) | (cB[6] << 16) | (cB[7] << 24);
							*reinterpret_cast<uint2 *>(s_col + 8u * o2) = pk;
						}
					} else {
						uint4 pk;
						pk.x = cA[0] | (cA[1] << 16); pk.y = cA[2] | (cA[3] << 16);
						pk.z = cA[4] | (cA[5] << 16); pk.w = cA[6] | (cA[7] << 16);
						*reinterpret_cast<uint4 *>(s_col + 8u * o) = pk;
						if (has2) {
							pk.x = cB[0] | (cB[1] << 16); pk.y = cB[2] | (cB[3] << 16);
							pk.z = cB[4] | (cB[5] << 16); pk.w = cB[6] | (cB[7] << 16);
							*reinterpret_cast<uint4 *>(s_col + 8u * o2) = pk;
						}
					}
				}
				if (tid == 0) s_ctl[1] = 0u;
				__syncthreads();

				// ---------------- PHASE 2a: thread-per-vertex masks (light) ----------------
				mine = haveV && (uint32_t)tid >= done && (uint32_t)tid < done + cnt;
				const bool heavy = mine && deg > (uint32_t)kLightMaxDeg;
				if (mine && !heavy) {
					const ColT * p = s_col + (myBeg - ea);
					if (kWide) {
#pragma unroll
						for (int w = 0; w < 2 * W; ++w) s_m32[w * kTileV + tid] = 0u;
						for (uint32_t i = 0; i < deg; ++i) {
							const uint32_t c = p[i];
							same += (c == own);
							s_m32[(c >> 5) * kTileV + tid] |= 1u << (c & 31u);
						}
						load_wide_mask((uint32_t)tid, m);
					} else {
						for (uint32_t i = 0; i < deg; ++i) {
							const uint32_t c = p[i];
							same += (c == own);
							if (W == 1) m[0] |= 1ull << c;
							else {
#pragma unroll
								for (int w = 0; w < W; ++w) m[w] |= ((int)(c >> 6) == w) ? (1ull << (c & 63u)) : 0ull;
							}
						}
					}
				} else if (heavy) {
					s_heavy[atomicAdd(&s_ctl[1], 1u)] = (uint16_t)tid;
				}
				__syncthreads();
				const uint32_t nHeavy = s_ctl[1];
				// ---------------- PHASE 2b: warp-per-vertex masks (heavy) ----------------
				if (nHeavy) {
					for (uint32_t h = warp; h < nHeavy; h += kThreads / 32) {
						const uint32_t slot = s_heavy[h];
						const uint32_t hb = s_rp[slot], hd = s_rp[slot + 1] - hb;
						const uint32_t hOwn = (uint32_t)cur[a.vBegin + v0 + slot];
						const ColT * p = s_col + (hb - ea);
						uint32_t hs = 0;
						if (kWide) {
							if (lane < 2 * W) s_wm[warp * 2 * W + lane] = 0u;
							__syncwarp();
							for (uint32_t i = lane; i < hd; i += 32) {
								const uint32_t c = p[i];
								hs += (c == hOwn);
								atomicOr(&s_wm[warp * 2 * W + (c >> 5)], 1u << (c & 31u));
							}
							__syncwarp();
							if (lane < 2 * W) s_m32[lane * kTileV + slot] = s_wm[warp * 2 * W + lane];
							__syncwarp();
						} else {
							unsigned long long hm[W];
#pragma unroll
							for (int w = 0; w < W; ++w) hm[w] = 0ull;
							for (uint32_t i = lane; i < hd; i += 32) {
								const uint32_t c = p[i];
								hs += (c == hOwn);
#pragma unroll
								for (int w = 0; w < W; ++w) hm[w] |= ((int)(c >> 6) == w) ? (1ull << (c & 63u)) : 0ull;
							}
#pragma unroll
							for (int w = 0; w < W; ++w) {
								const unsigned long long r = warp_reduce_or64(hm[w]);
								if (lane == 0) s_mask[(size_t)slot * W + w] = r;
							}
						}
						hs = __reduce_add_sync(0xffffffffu, hs);
						if (lane == 0) s_same[slot] = hs;
					}
					__syncthreads();
					if (heavy) {
						if (kWide) load_wide_mask((uint32_t)tid, m);
						else {
#pragma unroll
							for (int w = 0; w < W; ++w) m[w] = s_mask[(size_t)tid * W + w];
						}
						same = s_same[tid];
					}
				}
				done += cnt;
			}

			// ---------------- PHASE 3: thread-per-vertex proposal, draw, colour write ----------------
			if (mine)
				commit_vertex<W, ColT, kDyn>(a, t, nxt, a.vBegin + v0 + tid, v0 + tid, own, m, same, s_S, s_dist, s_hist, stayW,
				                             accDirected, accViol, &wq);
			__syncthreads();
			{
				const uint32_t qn = min(*wq.count, wq.cap);
				if (qn) drain_walk_queue<W, ColT, kDyn>(a, nxt, wq, 0u, qn, s_dist, s_hist, tid);
				__syncthreads();
				if (tid == 0) *wq.count = 0u;
			}
		} // sub-tiles
	} // tiles

	// ---- epilogue: counters ----
	accDirected = warp_reduce_add64(accDirected);
	accViol = warp_reduce_add64(accViol);
	__syncthreads();
	unsigned long long * s_red = reinterpret_cast<unsigned long long *>(s_mask);   // reuse
	if (lane == 0) { s_red[warp] = accDirected; s_red[8 + warp] = accViol; }
	__syncthreads();
	if (tid == 0) {
		unsigned long long d = 0, vv = 0;
		for (int w = 0; w < kThreads / 32; ++w) { d += s_red[w]; vv += s_red[8 + w]; }
		if (d) atomicAdd(a.scratch + 0, d);
		if (vv) atomicAdd(a.scratch + 1, vv);
	}
	if (!a.countOnly) {
		for (uint32_t k = tid; k < nCol; k += kThreads) {
			const int dlt = s_hist[k];
			if (dlt) atomicAdd(a.scratch + 2 + k, (unsigned long long)(long long)dlt);
		}
	}
	if (a.fuseFinalize) {
		__threadfence();
		__syncthreads();
		if (tid == 0) s_ctl[2] = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
		__syncthreads();
		if (s_ctl[2]) {
			__threadfence();
			finalize_sweep_device(a);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
template <typename ColT>
__global__ void init_colors_philox_kernel(ColT * colors, uint32_t n, uint32_t nCol, uint64_t seed) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;      // one Philox call = the colours of vertices 4g .. 4g+3
	const uint32_t v = 4u * g;
	if (v >= n) return;
	// uniform colour in [0,nCol) -- replaces initColoring (coloringMCMC_utils.cu:24-33) without its nCol overshoot
	const uint4 w = philox4(seed, 0u, g, 1u);
	colors[v] = (ColT)__umulhi(w.x, nCol);
	if (v + 1u < n) colors[v + 1u] = (ColT)__umulhi(w.y, nCol);
	if (v + 2u < n) colors[v + 2u] = (ColT)__umulhi(w.z, nCol);
	if (v + 3u < n) colors[v + 3u] = (ColT)__umulhi(w.w, nCol);
}

template <typename ColT>
__global__ void narrow_colors_kernel(const uint32_t * src, ColT * dst, uint32_t n, uint32_t nCol, DevState * st) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	const uint32_t c = src[v];
	if (c >= nCol) { st->errorFlag = 1u; dst[v] = 0; }
	else dst[v] = (ColT)c;
}

template <typename ColT>
__global__ void widen_colors_kernel(const ColT * src, uint32_t * dst, uint32_t n) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v < n) dst[v] = (uint32_t)src[v];
}

// class sizes of a colouring (used at init only; the sweeps keep them current through deltas); also the range check of
// colourings that arrive in the narrow device format (mcmcb200_init_colors_narrow)
template <typename ColT>
__global__ void class_sizes_kernel(const ColT * colors, uint32_t n, uint32_t nCol, unsigned long long * hist, DevState * st) {
	extern __shared__ unsigned int s_h[];
	if (nCol > 8192u) {                                           // palettes too wide for a shared-memory histogram (launched without one)
		for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
			const uint32_t c = colors[v];
			if (c < nCol) atomicAdd(hist + c, 1ull); else st->errorFlag = 1u;
		}
		return;
	}
	for (uint32_t k = threadIdx.x; k < nCol; k += blockDim.x) s_h[k] = 0u;
	__syncthreads();
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
		const uint32_t c = colors[v];
		if (c < nCol) atomicAdd(&s_h[c], 1u); else st->errorFlag = 1u;
	}
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < nCol; k += blockDim.x)
		if (s_h[k]) atomicAdd(hist + k, (unsigned long long)s_h[k]);
}

} // namespace mcmcb200
