// Source-blocked two-pass sweep.  Same result as sweep_kernel, different data movement.
//
// Why: the direct kernel issues one random 1-byte gather per directed edge into the L2-resident colour array; ncu
// shows it pinned on the L1->L2 request port (one 32-byte sector request per cycle per SM, profiles/r01a_*).  Here
// every random access is served from SHARED memory and DRAM sees streams:
//
//   static, once per graph (build_blocked_layout):
//     the directed edges (v <- u) are binned by SOURCE chunk b = u / 65536 (stable radix sort, so inside a bucket
//     they stay in CSR order, i.e. grouped by destination tile T and vertex); each (bucket, tile) run is padded to
//     4 entries.  srcLocal[pos] = u % 65536 (u16, bucket-major).  ecol, the gathered neighbour colours, is TILE-major:
//     tile T's stage image (its P runs back to back) is one contiguous block; granDst maps every 4-entry granule of
//     srcLocal to its place in ecol.  gidxS (SELL-32, u16) says where each edge's colour sits in its tile's image.
//   pass A  blocked_gather_kernel: per (part, bucket) item, load the 64 Ki colours of the chunk into shared memory
//     (coalesced), stream srcLocal (2 B/edge) + granDst (1 B/edge), gather from shared memory, scatter 4-byte granules
//     into ecol (1 B/edge; neighbouring runs are written by concurrently running CTAs and meet in L2).
//   pass B  blocked_sweep_kernel: per destination tile, ONE contiguous cp.async block copy of the tile's image into
//     shared memory (+ slot table, slice starts, own colours), occupancy masks through the SELL index words
//     (2 B/edge, shared-memory gather), then exactly phases 2-3 of the direct kernel (proposal, draw, colour write,
//     counters, device-side finalize).
//   overlap: pass A is DRAM bound, pass B instruction-issue bound.  They run concurrently on two streams; pass A
//     works through the tiles in `parts` and counts finished buckets per part, pass B takes tiles in the same order and
//     starts a tile when its part is complete (bl.sync).  Two 384-thread pass-B CTAs and one 256-thread pass-A CTA fill
//     an SM's registers exactly.
//
// DRAM bytes per directed edge: 2 + 1 + 1 (pass A) + 2 + 1 (pass B) = 7 versus the 8 of the reference layout, all of it
// streaming; measured per sweep on config 3: pass A 6.9 GB read (incl. ~1 GB partial-sector fills) + 2.1 GB written,
// pass B 5.7 GB read + 0.1 GB written (profiles/r01g_*).
#pragma once
#include <cub/cub.cuh>
#include <type_traits>

#include "sweep_kernel.cuh"

namespace mcmcb200 {

#ifndef MCMCB200_TIMING
#define MCMCB200_TIMING 0     /* 1: kernels record start/end/wait times (experiments; host prints them after every sweep) */
#endif
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr uint32_t kChunkBits = 16;
constexpr uint32_t kChunkV    = 1u << kChunkBits;    // source chunk: 65536 vertices, u16 local ids
#ifndef MCMCB200_THREADS_A
#define MCMCB200_THREADS_A 256
#endif
constexpr int      kThreadsA  = MCMCB200_THREADS_A;
// pass-B geometry: MCMCB200_THREADS_B threads per CTA, two CTAs per SM for the narrow-palette instances.  Every thread helps
// staging a tile (one contiguous cp.async block), the second CTA of the SM hides the wait.  Together with the 64-register cap
// below, 2 x 384 pass-B threads leave exactly the registers, threads and shared memory one 256-thread pass-A CTA needs on the
// same SM -- that is what lets the two passes overlap.  (A warp-specialised producer/consumer variant with double-buffered
// tiles was measured slower -- 4.80 vs 4.59 ms per sweep on config 3 -- and removed.)
#ifndef MCMCB200_THREADS_B
#define MCMCB200_THREADS_B 384
#endif
#ifndef MCMCB200_QUEUE_CAP
#define MCMCB200_QUEUE_CAP 48
#endif
#ifndef MCMCB200_HEAVY_CAP
#define MCMCB200_HEAVY_CAP 1024
#endif
// palettes wider than 128 colours keep their masks in 4-8 64-bit registers per lane: those instances run 512 threads, one CTA
#ifndef MCMCB200_REGS_B
#define MCMCB200_REGS_B 64
#endif
template <int W> struct PassB { static constexpr int threads = (W <= 2) ? MCMCB200_THREADS_B : 512;
                                static constexpr int maxRegs = (W <= 2) ? MCMCB200_REGS_B : 128; };

struct BlockedLayout {
	bool      valid = false;
	uint32_t  P = 0;             // source chunks
	uint32_t  TV = 0;            // vertices per destination tile (multiple of 256)
	uint32_t  numTiles = 0;
	uint32_t  stageCap = 0;      // entries a tile stages at most
	uint32_t  totalPadded = 0;   // entries in srcLocal / ecol
	uint16_t * srcLocal = nullptr;   // [totalPadded], bucket-major
	void *     ecol = nullptr;       // ColT[totalPadded]
	uint16_t * gidx = nullptr;       // [nnzLocal (+16)], CSR order (heavy rows and layout construction)
	uint2 *    gidxS = nullptr;      // SELL-32-sigma copy of gidx for the light rows: slice-interleaved 4-entry words
	uint16_t * order = nullptr;      // (construction only) [numTiles*TV] slot -> vertex (local to the tile), degree-descending inside each tile
	uint16_t * slotInfo = nullptr;   // [numTiles*TV] slot -> local vertex (bits 0-12) | kSlotHeavy for rows longer than kLightMaxDeg; 0xffff = empty slot.  Staged per tile (TMA)
	uint32_t * sliceOff = nullptr;   // [numTiles*TV/32 + 1] start of each 32-slot slice in gidxS (uint2 units)
	uint32_t * granDst = nullptr;    // [totalPadded/4]  chunk-major granule (4 entries of srcLocal) -> its granule in the tile-major ecol
	uint32_t * tileBase = nullptr;   // [numTiles+1]     first entry of each tile's stage image in ecol
	uint32_t * items = nullptr;      // [numItems][3] = bucket, begin, end (entries); item = part * P + bucket (empty items allowed)
	uint32_t  numItems = 0;
	uint32_t  numParts = 0;          // pass A works through the tiles in numParts stretches (short ones first and last) ...
	uint8_t * tilePart = nullptr;    // [numTiles] part of each tile
	uint32_t * sync = nullptr;       // [2 + numParts]: next item, next tile, buckets finished per part  (... and pass B follows behind)
	unsigned long long * dbgTimes = nullptr;
	uint32_t  nbuf = 1;              // stage buffers per pass-B CTA: 2 = tile T+1 is copied in (TMA) while tile T is computed
	size_t    smemA = 0, smemB = 0;
	size_t    bytes = 0;             // device memory the layout holds (mcmcb200_layout_bytes)
	int       gridA = 0, gridB = 0;
};

struct BlockedArgs {
	uint32_t P, TV, numTiles, stageCap;
	const uint16_t * srcLocal;
	void * ecol;
	const uint16_t * gidx;
	const uint2 * gidxS;
	const uint16_t * slotInfo;
	const uint32_t * sliceOff;
	const uint32_t * granDst;
	const uint32_t * tileBase;
	const uint32_t * items;
	uint32_t numItems, numParts, nbuf;
	const uint8_t * tilePart;
	unsigned long long * dbgTimes;   // (MCMCB200_TIMING builds only) [0] A first start [1] A last end [2] B first start [3] B last end [4] B wait ns
	uint32_t * sync;
#if MCMCB200_BOUNDS_CHECK
	uint32_t totalPadded;            // entries of srcLocal / ecol.  (Only in the checking build: pass B sits right at its 64-register cap and
	                                 //  ptxas starts spilling when this struct grows.)
#endif
};

// ------------------------------------------------------------------------------------------------------------------
// layout construction kernels
// ------------------------------------------------------------------------------------------------------------------
__global__ void blk_max_tile_edges_kernel(const uint32_t * rowptr, uint32_t nLocal, uint32_t TV, uint32_t numTiles, uint32_t * out) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= numTiles) return;
	const uint32_t v0 = k * TV, v1 = min(nLocal, v0 + TV);
	atomicMax(out, rowptr[v1] - rowptr[v0]);
}

__global__ void blk_tile_edge_starts_kernel(const uint32_t * rowptr, uint32_t nLocal, uint32_t TV, uint32_t numTiles, uint32_t * tileE) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > numTiles) return;
	tileE[k] = rowptr[min(nLocal, k * TV)];
}

__global__ void blk_edge_keys_kernel(const uint32_t * neighs, uint32_t nnz, uint16_t * keys, uint32_t * vals) {
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nnz) return;
	keys[e] = (uint16_t)(neighs[e] >> kChunkBits);
	vals[e] = e;
}

// tile of a CSR edge position: tileE is ascending, tileE[T] <= e < tileE[T+1] (empty tiles are skipped).  Interpolation first --
// tiles of a graph that fits this layout hold similar numbers of edges, so the guess e * numTiles / nnz is almost always within
// a step or two (the plain 16-step binary search per edge made blk_fill_entries_kernel the longest kernel of the layout build:
// 108 ms on config 3) -- then a bounded walk, then binary search on what is left.
__device__ __forceinline__ uint32_t blk_tile_of_edge(const uint32_t * __restrict__ tileE, uint32_t numTiles, uint32_t e) {
	const uint32_t nnz = __ldg(tileE + numTiles);
	uint32_t T = (uint32_t)(((unsigned long long)e * numTiles) / (nnz ? nnz : 1u));
	if (T >= numTiles) T = numTiles - 1u;
	uint32_t lo = 0, hi = numTiles;                         // answer in [lo, hi)
	for (int step = 0; step < 6; ++step) {
		if (__ldg(tileE + T) > e) { hi = T; if (T == 0) break; --T; }
		else if (__ldg(tileE + T + 1) <= e) { lo = T + 1; ++T; if (T >= numTiles) { T = numTiles - 1u; break; } }
		else return T;
	}
	if (lo >= hi) return min(lo, numTiles - 1u);
	hi = hi - 1;                                            // first index with tileE[idx + 1] > e in [lo, hi]
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (__ldg(tileE + mid + 1) > e) hi = mid; else lo = mid + 1;
	}
	return lo;
}

__global__ void blk_run_count_kernel(const uint16_t * keys, const uint32_t * vals, uint32_t nnz, const uint32_t * tileE,
                                     uint32_t numTiles, uint32_t * cnt /* [P][numTiles] */) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nnz) return;
	const uint32_t T = blk_tile_of_edge(tileE, numTiles, vals[i]);
	atomicAdd(&cnt[(size_t)keys[i] * numTiles + T], 1u);
}

// plen[b][T] = cnt padded to 4; plenT[T][b] = the same, transposed
__global__ void blk_pad_kernel(const uint32_t * cnt, uint32_t P, uint32_t numTiles, uint32_t * plen, uint32_t * plenT) {
	const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (size_t)P * numTiles) return;
	const uint32_t b = (uint32_t)(idx / numTiles), T = (uint32_t)(idx % numTiles);
	const uint32_t p = (cnt[idx] + 3u) & ~3u;
	plen[idx] = p;
	plenT[(size_t)T * P + b] = p;
}

// runStart[T][b] = gs[b][T];  stageOff[T][b] = scanT[T][b] - scanT[T][0], stageOff[T][P] = tile total
__global__ void blk_tables_kernel(const uint32_t * gs, const uint32_t * scanT, const uint32_t * plenT, uint32_t P, uint32_t numTiles,
                                  uint32_t * runStart, uint32_t * stageOff, uint32_t * maxStage) {
	const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (size_t)P * numTiles) return;
	const uint32_t T = (uint32_t)(idx / P), b = (uint32_t)(idx % P);
	runStart[idx] = gs[(size_t)b * numTiles + T];
	const uint32_t off = scanT[idx] - scanT[(size_t)T * P];
	stageOff[(size_t)T * (P + 1) + b] = off;
	if (b == P - 1) {
		const uint32_t tot = off + plenT[idx];
		stageOff[(size_t)T * (P + 1) + P] = tot;
		atomicMax(maxStage, tot);
	}
}

__global__ void blk_fill_entries_kernel(const uint16_t * keys, const uint32_t * vals, uint32_t nnz, const uint32_t * neighs,
                                        const uint32_t * tileE, uint32_t numTiles, uint32_t P, const uint32_t * us /* [P][numTiles] first sorted index */,
                                        const uint32_t * gs /* [P][numTiles] */, const uint32_t * stageOff, const uint32_t * scanT, uint32_t alignMask,
                                        uint16_t * srcLocal, uint16_t * gidx) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nnz) return;
	const uint32_t e = vals[i], b = keys[i];
	const uint32_t T = blk_tile_of_edge(tileE, numTiles, e);
	const size_t idx = (size_t)b * numTiles + T;
	const uint32_t r = i - us[idx];
	srcLocal[gs[idx] + r] = (uint16_t)(neighs[e] & (kChunkV - 1u));
	// position in the tile's stage: pass B copies the tile's block from the 16-byte boundary below its first entry
	gidx[e] = (uint16_t)((scanT[(size_t)T * P] & alignMask) + stageOff[(size_t)T * (P + 1) + b] + r);
}

// ---- tile-local construction (P <= kTileLocalMaxP source chunks): a WARP per destination tile walks the tile's contiguous CSR slice in
//      order and keeps one counter per source chunk in shared memory -- run lengths in a first pass; in a second pass the position of
//      every edge inside its (chunk, tile) run = the counter before it, which is exactly the rank a stable sort by chunk would give it.
//      No global sort of the 1.6e9 edges, no per-edge tile search, and both outputs are written (nearly) in order: the sort-based
//      construction spent 110 ms of config 3's 207 ms in one kernel that gathered neighs[e] and scattered gidx[e] at random
//      (2 x 51 GB of 32-byte sectors for 2-byte payloads); this one reads the CSR twice. ----
constexpr uint32_t kTileLocalMaxP = 4096;       // 16 KiB of counters per warp at most
constexpr int kTileLocalWarps = 4;

__global__ void __launch_bounds__(kTileLocalWarps * 32)
blk_tile_hist_kernel(const uint32_t * __restrict__ neighs, const uint32_t * __restrict__ tileE, uint32_t numTiles, uint32_t P, uint32_t * cnt /* [P][numTiles] */) {
	extern __shared__ uint32_t s_tl[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t * h = s_tl + (size_t)warp * P;
	const uint32_t T = blockIdx.x * kTileLocalWarps + warp;
	if (T >= numTiles) return;                                   // (no CTA barrier below)
	for (uint32_t b = lane; b < P; b += 32) h[b] = 0u;
	__syncwarp();
	const uint32_t e0 = tileE[T], e1 = tileE[T + 1];
	for (uint32_t e = e0; e < e1; e += 32u) {
		const bool active = e + lane < e1;
		const uint32_t b = active ? (__ldcs(neighs + e + lane) >> kChunkBits) : 0xffffffffu;
		const uint32_t mask = __match_any_sync(0xffffffffu, b);
		if (active && lane == __ffs((int)mask) - 1) h[b] += (uint32_t)__popc(mask);
		__syncwarp();
	}
	for (uint32_t b = lane; b < P; b += 32) cnt[(size_t)b * numTiles + T] = h[b];
}

// second pass: ONE warp per CTA and tile.  The tile's source-local ids are first put in STAGE ORDER in shared memory (position =
// the run's offset in the tile's stage image + rank), then every run leaves as whole 8-byte granules -- 4 stores per 16-entry run
// instead of 16 scattered 2-byte ones (the 2-byte scatter made the first version of this kernel 96 ms on config 3).
__host__ __device__ inline size_t blk_tile_rank_smem(uint32_t P, uint32_t stageCap) { return sizeof(uint32_t) * (size_t)P + sizeof(uint16_t) * ((size_t)stageCap + 16); }

__global__ void __launch_bounds__(32)
blk_tile_rank_kernel(const uint32_t * __restrict__ neighs, const uint32_t * __restrict__ tileE, uint32_t numTiles, uint32_t P, uint32_t stageCap,
                     const uint32_t * __restrict__ runStart /* [T][b] */, const uint32_t * __restrict__ stageOff /* [T][P+1] */,
                     const uint32_t * __restrict__ scanT /* [T][b] */, uint32_t alignMask, uint16_t * srcLocal, uint16_t * gidx) {
	extern __shared__ __align__(16) uint32_t s_tl[];
	const int lane = threadIdx.x;
	uint32_t * h = s_tl;
	uint16_t * sbuf = reinterpret_cast<uint16_t *>(s_tl + P);     // (P * 4 bytes: 8-byte aligned for even P; odd P is padded below)
	if (P & 1u) sbuf += 2;
	const uint32_t T = blockIdx.x;
	if (T >= numTiles) return;
	const uint32_t * rs = runStart + (size_t)T * P;
	const uint32_t * so = stageOff + (size_t)T * (P + 1);
	const uint32_t tot = __ldg(so + P);                           // padded entries of the tile's stage image (<= stageCap)
	for (uint32_t b = lane; b < P; b += 32) h[b] = 0u;
	for (uint32_t i = lane; 2u * i < tot; i += 32) reinterpret_cast<uint32_t *>(sbuf)[i] = 0u;    // run padding stays 0
	__syncwarp();
	const uint32_t e0 = tileE[T], e1 = tileE[T + 1];
	const uint32_t base = __ldg(scanT + (size_t)T * P) & alignMask;     // pass B copies the tile's block from the 16-byte boundary below its first entry
	const uint32_t lt = (1u << lane) - 1u;
	constexpr int kU = 8;                                         // chunks of 32 edges loaded ahead (one warp per SM sub-partition: latency is everything)
	for (uint32_t eb = e0; eb < e1; eb += 32u * kU) {
		uint32_t nbv[kU];
#pragma unroll
		for (int u = 0; u < kU; ++u) { const uint32_t e = eb + 32u * u + lane; nbv[u] = (e < e1) ? __ldcs(neighs + e) : 0xffffffffu; }
#pragma unroll
		for (int u = 0; u < kU; ++u) {
			const uint32_t e = eb + 32u * u + lane;
			const bool active = e < e1;
			const uint32_t nb = nbv[u];
			const uint32_t b = active ? (nb >> kChunkBits) : 0xffffffffu;
			const uint32_t mask = __match_any_sync(0xffffffffu, b);
			uint32_t r = 0;
			if (active) r = h[b] + (uint32_t)__popc(mask & lt);          // edges of the same chunk earlier in CSR order, in this tile
			__syncwarp();
			if (active && lane == __ffs((int)mask) - 1) h[b] += (uint32_t)__popc(mask);
			__syncwarp();
			if (active) {
				const uint32_t pos = __ldg(so + b) + r;
				sbuf[pos] = (uint16_t)(nb & (kChunkV - 1u));
				gidx[e] = (uint16_t)(base + pos);
			}
		}
	}
	__syncwarp();
	// runs out: lane per source chunk, whole granules (runs are padded to 4 entries = 8 bytes, both here and in srcLocal)
	for (uint32_t b = lane; b < P; b += 32) {
		const uint32_t p0 = __ldg(so + b), p1 = __ldg(so + b + 1);
		const uint2 * src = reinterpret_cast<const uint2 *>(sbuf + p0);
		uint2 * dst = reinterpret_cast<uint2 *>(srcLocal + __ldg(rs + b));
		for (uint32_t g = 0; g < ((p1 - p0) >> 2); ++g) dst[g] = src[g];
	}
}

// pass-A work items: item (part p, bucket b) = the entries of bucket b that belong to tiles [partStart[p], partStart[p+1])
__global__ void blk_items_kernel(const uint32_t * gs /* [P][numTiles] */, uint32_t P, uint32_t numTiles, const uint32_t * partStart, uint32_t numParts,
                                 uint32_t total, uint32_t * items, uint8_t * tilePart) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numParts * P) return;
	const uint32_t p = i / P, b = i % P;
	const uint32_t t0 = partStart[p], t1 = partStart[p + 1];
	const uint32_t nextBucket = (b + 1 < P) ? gs[(size_t)(b + 1) * numTiles] : total;
	items[3 * (size_t)i] = b;
	items[3 * (size_t)i + 1] = gs[(size_t)b * numTiles + t0];
	items[3 * (size_t)i + 2] = (t1 < numTiles) ? gs[(size_t)b * numTiles + t1] : nextBucket;
	if (b == 0) for (uint32_t T = t0; T < t1; ++T) tilePart[T] = (uint8_t)p;
}

// granDst: pass A reads srcLocal chunk-major (so that the gather runs out of one 64 Ki-colour chunk in shared memory) and
// writes ecol TILE-major, so that pass B finds the whole stage image of a tile as one contiguous block.  One entry per
// 4-entry granule: where the granule of run (chunk b, tile T) lands.
__global__ void blk_gran_kernel(const uint32_t * runStart /* [T][b] chunk-major start */, const uint32_t * scanT /* [T][b] tile-major start */,
                                const uint32_t * plenT /* [T][b] */, size_t cells, uint32_t * granDst) {
	const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= cells) return;
	const uint32_t g0 = runStart[idx] >> 2, ng = plenT[idx] >> 2, d0 = scanT[idx] >> 2;
	for (uint32_t g = 0; g < ng; ++g) granDst[g0 + g] = d0 + g;
}

__global__ void blk_tile_base_kernel(const uint32_t * scanT, uint32_t P, uint32_t numTiles, uint32_t totalPadded, uint32_t * tileBase) {
	const uint32_t T = blockIdx.x * blockDim.x + threadIdx.x;
	if (T < numTiles) tileBase[T] = scanT[(size_t)T * P];
	if (T == numTiles) tileBase[T] = totalPadded;
}

// SELL-32-sigma (sigma = one tile) construction -------------------------------------------------------------------
// key = (tile << 32) | (0xffffffff - degree): one stable radix sort orders every tile's vertices by descending degree
__global__ void blk_sell_keys_kernel(const uint32_t * rowptr, uint32_t nLocal, uint32_t TV, unsigned long long * keys, uint32_t * vals) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nLocal) return;
	const uint32_t deg = rowptr[v + 1] - rowptr[v];
	keys[v] = ((unsigned long long)(v / TV) << 32) | (unsigned long long)(0xffffffffu - deg);
	vals[v] = v;
}

// order[T*TV + s] = local index of the s-th vertex of tile T (sorted); slots past the tile's vertex count get 0xffff
__global__ void blk_sell_order_kernel(const uint32_t * sortedV, uint32_t nLocal, uint32_t TV, uint32_t numTiles, uint16_t * order) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (size_t)numTiles * TV) return;
	const uint32_t T = (uint32_t)(i / TV), s = (uint32_t)(i % TV);
	const uint32_t v0 = T * TV, nv = min(TV, nLocal - v0);
	order[i] = (s < nv) ? (uint16_t)(sortedV[v0 + s] - v0) : (uint16_t)0xffffu;
}

// words[slice] = 32 * ceil(max light degree in the slice / 4)   (uint2 words the slice occupies in gidxS)
__global__ void blk_sell_width_kernel(const uint32_t * rowptr, const uint16_t * order, uint32_t TV, uint32_t numSlices, uint32_t * words) {
	const uint32_t sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (sl >= numSlices) return;
	const size_t slot = (size_t)sl * 32 + lane;
	const uint32_t T = (uint32_t)(slot / TV);
	const uint32_t o = order[slot];
	uint32_t deg = 0;
	if (o != 0xffffu) { const uint32_t v = T * TV + o; deg = rowptr[v + 1] - rowptr[v]; if (deg > (uint32_t)kLightMaxDeg) deg = 0; }
	deg = __reduce_max_sync(0xffffffffu, deg);
	if (lane == 0) words[sl] = 32u * ((deg + 3u) >> 2);
}

// slotInfo[T*TV + s] = local vertex (TV <= 8192: 13 bits) | kSlotHeavy when the row is longer than kLightMaxDeg; 0xffff = empty
constexpr uint32_t kSlotHeavy = 0x8000u, kSlotVertexMask = 0x1fffu;
__global__ void blk_sell_slotinfo_kernel(const uint32_t * rowptr, const uint16_t * order, uint32_t TV, uint32_t numTiles, uint16_t * slotInfo) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (size_t)numTiles * TV) return;
	const uint32_t T = (uint32_t)(i / TV), o = order[i];
	uint32_t info = 0xffffu;
	if (o != 0xffffu) { const uint32_t v = T * TV + o; info = o | ((rowptr[v + 1] - rowptr[v] > (uint32_t)kLightMaxDeg) ? kSlotHeavy : 0u); }
	slotInfo[i] = (uint16_t)info;
}

// Every row of a slice is filled to the slice's width: positions past the row's degree (and the whole row of an empty
// slot or of a heavy vertex) hold `dummy`, the index of a stage byte that pass B keeps at the all-ones colour, which
// matches no palette entry -- the mask loop needs no per-edge bounds test.
__global__ void blk_sell_fill_kernel(const uint32_t * rowptr, const uint16_t * order, const uint16_t * gidx, uint32_t TV, uint32_t numSlices,
                                     const uint32_t * sliceOff, uint32_t dummy, uint2 * gidxS) {
	const uint32_t sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (sl >= numSlices) return;
	const size_t slot = (size_t)sl * 32 + lane;
	const uint32_t T = (uint32_t)(slot / TV);
	const uint32_t o = order[slot];
	uint32_t beg = 0, deg = 0;
	if (o != 0xffffu) { const uint32_t v = T * TV + o; beg = rowptr[v]; deg = rowptr[v + 1] - beg; if (deg > (uint32_t)kLightMaxDeg) deg = 0; }
	const uint32_t base = sliceOff[sl], width = (sliceOff[sl + 1] - base) >> 5;
	for (uint32_t j = 0; j < width; ++j) {
		uint32_t e[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) e[k] = (4u * j + k < deg) ? (uint32_t)gidx[beg + 4u * j + k] : dummy;
		gidxS[(size_t)base + (size_t)j * 32 + lane] = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
	}
}

#ifndef MCMCB200_ST_LAST
#define MCMCB200_ST_LAST 1      /* ecol stores keep their L2 lines (evict_last) until the neighbouring run completes the sector */
#endif
__device__ __forceinline__ void st_ecol32(void * p, uint32_t v, unsigned long long pol) {
	if (MCMCB200_ST_LAST) asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
	else *reinterpret_cast<uint32_t *>(p) = v;
}
__device__ __forceinline__ void st_ecol64(void * p, uint2 v, unsigned long long pol) {
	if (MCMCB200_ST_LAST) asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" :: "l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
	else *reinterpret_cast<uint2 *>(p) = v;
}

// ------------------------------------------------------------------------------------------------------------------
// pass A: ecol[pos] = cur[bucket * 65536 + srcLocal[pos]]  -- the gather runs out of shared memory
// ------------------------------------------------------------------------------------------------------------------
// (64 registers at most: one pass-A CTA must fit next to two pass-B CTAs -- 2 x 384 x 64 + 256 x 64 registers = the whole file)
template <typename ColT>
__global__ void __launch_bounds__(kThreadsA, 4)
blocked_gather_kernel(const SweepArgs a, const BlockedArgs bl) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	ColT * chunk = reinterpret_cast<ColT *>(smem_raw);
	const DevState * st = a.st;
	if (!a.countOnly && st->convergedAt >= 0) return;
	const uint32_t t = st->sweep;
	const ColT * __restrict__ cur = a.colorsOverride ? static_cast<const ColT *>(a.colorsOverride)
	                                                 : static_cast<const ColT *>(a.colors[t & 1]);
	ColT * __restrict__ ecol = static_cast<ColT *>(bl.ecol);
	const int tid = threadIdx.x;
	// Items are handed out dynamically in part-major order (item = part * P + bucket): the CTAs running at the same time hold
	// neighbouring source chunks and write the same stretch of tiles, so the runs of (T, b) and (T, b+1) -- adjacent in the
	// tile-major ecol -- meet in L2 instead of going to DRAM half written (3.7 GB of read-for-fill traffic per sweep otherwise).
	// Every finished item bumps its part's counter; pass B, running concurrently on another stream, starts a tile as soon as
	// all P buckets of the tile's part are in.
	unsigned long long pol = 0ull;
	if (MCMCB200_ST_LAST) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	__shared__ uint32_t s_item;
	__shared__ __align__(8) unsigned long long s_bar;           // mbarrier of the chunk copy
	if (tid == 0) { mbar_init(&s_bar, 1u); mbar_fence_init(); }
	uint32_t have = 0xffffffffu, loads = 0u;
	if (MCMCB200_TIMING && tid == 0) atomicMin(bl.dbgTimes + 0, global_ns());
	for (;;) {
		__syncthreads();                                      // everybody is done with the previous item (and its chunk)
		if (tid == 0) s_item = atomicAdd(bl.sync + 0, 1u);
		__syncthreads();
		const uint32_t it = s_item;
		if (it >= bl.numItems) { if (MCMCB200_TIMING && tid == 0) atomicMax(bl.dbgTimes + 1, global_ns()); break; }
		const uint32_t b = bl.items[3 * it], beg = bl.items[3 * it + 1], end = bl.items[3 * it + 2];
		if (b != have && beg < end) {
			// the chunk's 64 Ki colours: ONE bulk copy issued by one thread (TMA engine, no LDG/STS through the LSU pipe);
			// the colour buffers are padded by 64 Ki entries, so a whole chunk is always readable
			if (tid == 0) {
				fence_proxy_async();                              // the previous chunk's shared-memory reads precede the async-proxy write
				constexpr uint32_t bytes = kChunkV * (uint32_t)sizeof(ColT);
				mbar_arrive_expect_tx(&s_bar, bytes);
				tma_bulk_g2s(chunk, cur + (size_t)b * kChunkV, bytes, &s_bar);
			}
			mbar_wait(&s_bar, loads & 1u);
			loads++;
			have = b;
		}
		// one 4-entry granule per lane per step: a warp reads 256 contiguous bytes of local ids and 128 of destinations, and
		// its store covers whole 32-byte sectors wherever a run spans them (full-sector first touches need no fill in L2)
#ifndef MCMCB200_A_KU
#define MCMCB200_A_KU 13     /* config 3: 8 -> 3.62 ms, 10 -> 3.52, 12 -> 3.49, 13 -> 3.40 (with the plain pass-B instances), 14 -> 3.42, 15 -> 3.45, 16 (spills) -> 3.66 per sweep */
#endif
		constexpr uint32_t kU = MCMCB200_A_KU;   // granules in flight per thread: 12 bytes of loads each (pass A is bound by bytes in flight)
		const uint32_t g0 = beg >> 2, g1 = end >> 2;           // runs are padded to 4 entries: items are whole granules
		for (uint32_t j0 = g0 + tid; j0 < g1; j0 += kThreadsA * kU) {
			uint2 ids[kU];
			uint32_t gd[kU];
#pragma unroll
			for (uint32_t k = 0; k < kU; ++k) {
				const uint32_t j = j0 + k * kThreadsA;
				if (j < g1) { ids[k] = __ldcs(reinterpret_cast<const uint2 *>(bl.srcLocal) + j); gd[k] = __ldcs(bl.granDst + j); }
			}
#pragma unroll
			for (uint32_t k = 0; k < kU; ++k) {
				const uint32_t j = j0 + k * kThreadsA;
				if (j < g1) {
					const uint2 d = ids[k];
					MCMCB200_CHECK(4ull * gd[k] + 4ull <= bl.totalPadded && j < (bl.totalPadded >> 2), a.st);
					const uint32_t c0 = chunk[d.x & 0xffffu], c1 = chunk[d.x >> 16], c2 = chunk[d.y & 0xffffu], c3 = chunk[d.y >> 16];
					if (sizeof(ColT) == 1) st_ecol32(ecol + 4u * (size_t)gd[k], c0 | (c1 << 8) | (c2 << 16) | (c3 << 24), pol);
					else st_ecol64(ecol + 4u * (size_t)gd[k], make_uint2(c0 | (c1 << 16), c2 | (c3 << 16)), pol);
				}
			}
		}
		__threadfence();                                      // this item's ecol stores are visible device-wide ...
		__syncthreads();
		if (tid == 0) atomicAdd(bl.sync + 2 + it / bl.P, 1u);  // ... before its part is reported
	}
}

// ------------------------------------------------------------------------------------------------------------------
// pass B: per destination tile -- stage the tile (TMA), occupancy masks through the SELL index words, then phases 2-3 of the
// direct kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kWarpQueueCap = MCMCB200_QUEUE_CAP;     // deferred CDF walks parked per warp (drained 32 at a time, no CTA barrier; overflow walks inline)
__host__ __device__ constexpr uint32_t warp_queue_cap(int W) { return W == 1 ? kWarpQueueCap : (kWarpQueueCap * 5u) / 6u; }   // 24- vs 32-byte entries
constexpr uint32_t kHeavyCap = MCMCB200_HEAVY_CAP;        // warp-per-vertex work list per tile (overflow is handled by the owning thread)

// bytes of one stage buffer set (everything the TMA copies bring in for one tile); every part is a multiple of 16 bytes
__host__ __device__ inline size_t blocked_buf_bytes(uint32_t TV, uint32_t stageCap, int colBytes) {
	size_t b = 0;
	b += sizeof(uint16_t) * (size_t)TV;                    // slotTab: slot -> vertex | heavy flag
	b += sizeof(uint32_t) * (size_t)((TV >> 5) + 4);       // sliceTab: SELL slice starts of the tile
	b += (size_t)colBytes * (size_t)(TV + 16);             // ownCol: the tile's current colours
	b += (size_t)colBytes * (size_t)(stageCap + 16);       // stage; [stageCap, +16) = dummy colour
	return (b + 15) & ~(size_t)15;
}

__host__ __device__ inline size_t blocked_smem_bytes_B(uint32_t nCol, uint32_t nbuf, uint32_t TV, uint32_t stageCap, int colBytes, int W) {
	const int warps = ((W <= 2) ? MCMCB200_THREADS_B : 512) >> 5;
	size_t b = 0;
	b += sizeof(float) * (size_t)((nCol + 1 + 3) & ~3u);   // s_S
	b += sizeof(float) * (size_t)((nCol + 3) & ~3u);       // s_dist (DYNAMIC) / free-colour weight table (UNIFORM)
	b += sizeof(int) * (size_t)((nCol + 3) & ~3u);         // s_hist
	b += sizeof(uint32_t) * 8;                             // s_ctl
	b += sizeof(unsigned long long) * 4;                   // mbarriers (one per stage buffer) + tile ids
	b += sizeof(uint32_t) * 128;                           // s_red (64 x u64)
	b += sizeof(uint32_t) * 32;                            // per-warp queue counters
	b += sizeof(uint16_t) * (size_t)kHeavyCap;             // s_heavy
	b = (b + 15) & ~(size_t)15;
	if (W <= 2) b += (size_t)warps * warp_queue_cap(W) * (8 * W + 16);   // per-warp walk queues (mask, lv/own, u/w)
	b = (b + 15) & ~(size_t)15;
	b += sizeof(uint32_t) * (size_t)TV;                    // s_draw: the tile's Philox words (one call per 4 vertices)
	b += (size_t)nbuf * (((size_t)colBytes * (TV + 16) + 15) & ~(size_t)15);   // s_new: the tile's new colours, written out coalesced (local + peers)
	b += (size_t)nbuf * blocked_buf_bytes(TV, stageCap, colBytes);
	return (b + 127) & ~(size_t)127;
}

// pass B waits until pass A has delivered all P buckets of the tile's part.  Bounded: if the producer never shows up (it was
// not co-scheduled: another tenant on the device, a profiler serialising kernels the wrong way round) the sweep is ABORTED --
// errorFlag 2, no tile of this CTA is processed or written, the finalize step does not advance, and every host call that
// reads results returns MCMCB200_ECUDA (mcmcb200_sweep then retries the sweep with the two passes back to back).
__device__ __forceinline__ bool wait_part_ready(const BlockedArgs & bl, uint32_t T, DevState * st) {
	const uint32_t * flag = bl.sync + 2 + bl.tilePart[T];
	uint32_t v;
	long long t0 = 0;
	for (uint32_t spins = 0;; ++spins) {
		asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
		if (v >= bl.P) return true;
		if (spins == 0) t0 = clock64();
		else if ((spins & 255u) == 0) {
			if (*reinterpret_cast<volatile uint32_t *>(&st->errorFlag) == 2u) return false;      // another CTA gave up already
			if (clock64() - t0 > 4000000000ll) { st->errorFlag = 2u; __threadfence(); return false; }   // ~2 s
		}
		__nanosleep(100);
	}
}
#ifndef MCMCB200_EARLY_TICKET
#define MCMCB200_EARLY_TICKET 0     /* measured on config 3: 3.65 ms with, 3.64 without (the L2 prefetches compete with pass A) */
#endif
__device__ __forceinline__ bool part_ready_now(const BlockedArgs & bl, uint32_t T) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bl.sync + 2 + bl.tilePart[T]) : "memory");
	return v >= bl.P;
}
// L2 prefetch of everything tile T streams or copies (ranges widened to 16-byte boundaries; all arrays have slack behind their end)
template <typename ColT>
__device__ __forceinline__ void prefetch_tile_l2(const SweepArgs & a, const BlockedArgs & bl, const ColT * cur, uint32_t T) {
	const uint32_t TV = bl.TV, spt = TV >> 5;
	const uint32_t v0 = T * TV, nv = min(TV, a.nLocal - v0);
	const uint32_t tb = __ldg(bl.tileBase + T), te = __ldg(bl.tileBase + T + 1);
	constexpr uint32_t alignE = 16u / (uint32_t)sizeof(ColT);
	const uint32_t a0 = tb & ~(alignE - 1u);
	const uint32_t bytesStage = ((te - a0) * (uint32_t)sizeof(ColT) + 15u) & ~15u;
	if (bytesStage) bulk_prefetch_l2(static_cast<const ColT *>(bl.ecol) + a0, bytesStage);
	bulk_prefetch_l2(bl.slotInfo + (size_t)T * TV, TV * (uint32_t)sizeof(uint16_t));
	bulk_prefetch_l2(reinterpret_cast<const unsigned char *>(cur) + (size_t)(a.vBegin + v0) * sizeof(ColT), (nv * (uint32_t)sizeof(ColT) + 15u) & ~15u);
	const uint32_t s0 = __ldg(bl.sliceOff + (size_t)T * spt), s1 = __ldg(bl.sliceOff + (size_t)(T + 1) * spt);
	if (s1 > s0) bulk_prefetch_l2(bl.gidxS + s0, (s1 - s0) * (uint32_t)sizeof(uint2));
}
// 1 << c with PTX semantics: shift amounts >= 64 give 0 (the dummy colour sets no bit)
__device__ __forceinline__ unsigned long long bit64_clamped(uint32_t c) {
	unsigned long long r;
	asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(1ull), "r"(c));
	return r;
}

// shared-memory views of one pass-B CTA
template <int W, typename ColT>
struct PassBShared {
	float * S; float * dist; int * hist; uint32_t * ctl; unsigned long long * bars; uint32_t * tileOf; unsigned long long * red;
	uint32_t * qcnt; uint16_t * heavy; unsigned char * queues; uint32_t * draw; unsigned char * newBase; unsigned char * bufBase;
	uint32_t newBytes, bufBytes, TV, stageCap, soffStride;
	__device__ __forceinline__ PassBShared(unsigned char * raw, uint32_t nCol, uint32_t nbuf, uint32_t TV_, uint32_t stageCap_) {
		TV = TV_; stageCap = stageCap_;
		soffStride = (TV >> 5) + 4;
		S = reinterpret_cast<float *>(raw);
		dist = S + ((nCol + 1 + 3) & ~3u);
		hist = reinterpret_cast<int *>(dist + ((nCol + 3) & ~3u));
		ctl = reinterpret_cast<uint32_t *>(hist + ((nCol + 3) & ~3u));
		bars = reinterpret_cast<unsigned long long *>(ctl + 8);
		tileOf = reinterpret_cast<uint32_t *>(bars + 2);
		red = bars + 4;
		qcnt = reinterpret_cast<uint32_t *>(red + 64);
		heavy = reinterpret_cast<uint16_t *>(qcnt + 32);
		size_t off = (size_t)(reinterpret_cast<unsigned char *>(heavy + kHeavyCap) - raw);
		off = (off + 15) & ~(size_t)15;
		queues = raw + off;
		if (W <= 2) off += (size_t)(PassB<W>::threads / 32) * warp_queue_cap(W) * (8 * W + 16);
		off = (off + 15) & ~(size_t)15;
		draw = reinterpret_cast<uint32_t *>(raw + off);
		off += sizeof(uint32_t) * (size_t)TV;
		newBytes = (uint32_t)((sizeof(ColT) * (size_t)(TV + 16) + 15) & ~(size_t)15);
		newBase = raw + off;
		off += (size_t)nbuf * newBytes;
		bufBytes = (uint32_t)blocked_buf_bytes(TV, stageCap, (int)sizeof(ColT));
		bufBase = raw + off;
	}
	__device__ __forceinline__ uint16_t * slotTab(uint32_t buf) const { return reinterpret_cast<uint16_t *>(bufBase + (size_t)buf * bufBytes); }
	__device__ __forceinline__ uint32_t * sliceTab(uint32_t buf) const { return reinterpret_cast<uint32_t *>(slotTab(buf) + TV); }
	__device__ __forceinline__ ColT * ownCol(uint32_t buf) const { return reinterpret_cast<ColT *>(sliceTab(buf) + soffStride); }
	__device__ __forceinline__ ColT * stageBuf(uint32_t buf) const { return ownCol(buf) + (TV + 16); }
	__device__ __forceinline__ ColT * newCol(uint32_t buf) const { return reinterpret_cast<ColT *>(newBase + (size_t)buf * newBytes); }
};

// ONE thread: four bulk copies (TMA engine) bring in everything tile T needs -- slot table, slice starts, current colours and
// the gathered neighbour colours (pass A left the tile's whole stage image contiguous in ecol) -- and complete on `bar`.
template <int W, typename ColT>
__device__ __forceinline__ void stage_tile_tma(const SweepArgs & a, const BlockedArgs & bl, const PassBShared<W, ColT> & sm, const ColT * cur,
                                               uint32_t T, uint32_t buf, unsigned long long * bar, uint64_t polFirst, bool prefetchIdx = true) {
	const uint32_t TV = bl.TV, spt = TV >> 5;
	const uint32_t v0 = T * TV, nv = min(TV, a.nLocal - v0);
	const uint32_t tb = __ldg(bl.tileBase + T), te = __ldg(bl.tileBase + T + 1);
	const uint32_t bytesSlot = TV * (uint32_t)sizeof(uint16_t);
	const uint32_t bytesSoff = (spt + 4u) * (uint32_t)sizeof(uint32_t);     // spt + 1 used; the table has slack behind its end
	const uint32_t bytesOwn = (nv * (uint32_t)sizeof(ColT) + 15u) & ~15u;   // (colour arrays are padded; tiles start 256-aligned)
	// copy the stage image from the 16-byte boundary below its first entry (the static indices in gidx / gidxS include that offset)
	constexpr uint32_t alignE = 16u / (uint32_t)sizeof(ColT);
	const uint32_t a0 = tb & ~(alignE - 1u);
	const uint32_t bytesStage = ((te - a0) * (uint32_t)sizeof(ColT) + 15u) & ~15u;
	mbar_arrive_expect_tx(bar, bytesSlot + bytesSoff + bytesOwn + bytesStage);
	tma_bulk_g2s(sm.slotTab(buf), bl.slotInfo + (size_t)T * TV, bytesSlot, bar);
	tma_bulk_g2s(sm.sliceTab(buf), bl.sliceOff + (size_t)T * spt, bytesSoff, bar);
	tma_bulk_g2s(sm.ownCol(buf), reinterpret_cast<const unsigned char *>(cur) + (size_t)(a.vBegin + v0) * sizeof(ColT), bytesOwn, bar);
	if (bytesStage) tma_bulk_g2s_hint(sm.stageBuf(buf), static_cast<const ColT *>(bl.ecol) + a0, bytesStage, bar, polFirst);   // read once: evict_first
#ifndef MCMCB200_PREFETCH_IDX
#define MCMCB200_PREFETCH_IDX 1
#endif
	if (MCMCB200_PREFETCH_IDX && prefetchIdx) {
		// the tile's SELL index words (2 B per edge, streamed by the mask loop with plain loads: ncu showed their first use as the
		// kernel's largest stall, ~15 % of the samples waiting on DRAM): ask for the whole range in L2 now, while the copies fly
		const uint32_t s0 = __ldg(bl.sliceOff + (size_t)T * spt), s1 = __ldg(bl.sliceOff + (size_t)(T + 1) * spt);
		if (s1 > s0) bulk_prefetch_l2(bl.gidxS + s0, (s1 - s0) * (uint32_t)sizeof(uint2));
	}
}

// per-tile views handed to the slot / heavy-list routines
template <int W, typename ColT>
struct TileView {
	const uint16_t * slot; const uint32_t * soff; const ColT * own; const ColT * stage; ColT * snew; uint32_t * heavyCount; const uint32_t * draw;
	uint32_t v0, nv;
};

// ---- phases 1'+2+3 for one slot (thread per vertex): occupancy mask straight from the stage buffer through the static SELL
//      index words (u16: where each edge's colour sits in this tile's stage), then commit_vertex.  A warp's 32 rows are one
//      slice of uniform width, so the mask loop is warp-uniform and branch free. ----
template <int W, typename ColT, bool kDyn, bool kPlain>
__device__ __forceinline__ void sweep_slot(const SweepArgs & a, const BlockedArgs & bl, const PassBShared<W, ColT> & sm, const TileView<W, ColT> & tv,
                                           uint32_t t, ColT * __restrict__ nxt, uint32_t slot, int lane, float stayW,
                                           const WalkQueue<W> * wq, unsigned long long & accDirected, unsigned long long & accViol) {
	// the padded SELL rows need a colour value outside every palette of this instance: all-ones (255 / 65535).
	// u8 with W == 4 covers nCol up to 256, where 255 is a real colour: that instance keeps the per-edge degree test.
	constexpr bool kPad = !(sizeof(ColT) == 1 && W == 4);
	const uint32_t TV = bl.TV;
	const ColT * stage = tv.stage;
	const bool inTile = slot < TV;                            // warp-uniform (TV is a multiple of 32)
	const uint32_t info = inTile ? (uint32_t)tv.slot[slot] : 0xffffu;
	const bool valid = info != 0xffffu;
	const uint32_t lv = info & kSlotVertexMask;
	bool light = valid && !(info & kSlotHeavy);
	bool inlineHeavy = false;
	if (valid && !light) {                                    // warp-per-vertex list; if it is full the thread does the row itself
		const uint32_t hi = atomicAdd(tv.heavyCount, 1u);
		if (hi < kHeavyCap) sm.heavy[hi] = (uint16_t)lv; else inlineHeavy = true;
	}
	MCMCB200_CHECK(!valid || lv < tv.nv, a.st);
	const uint32_t own = valid ? (uint32_t)tv.own[lv] : 0u;
	unsigned long long m[W];
#pragma unroll
	for (int w = 0; w < W; ++w) m[w] = 0ull;
	uint32_t same = 0;
	auto addc = [&](uint32_t idx) {
		MCMCB200_CHECK(idx < bl.stageCap + 16u, a.st);
		const uint32_t c = stage[idx];
		MCMCB200_CHECK(c < a.nCol || (kPad && c == (uint32_t)(ColT)~(ColT)0), a.st);
		asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(same) : "r"(c), "r"(own));   // same += (c == own)
		if (W == 1) m[0] |= kPad ? bit64_clamped(c) : (1ull << c);
		else {
#pragma unroll
			for (int w = 0; w < W; ++w) m[w] |= ((int)(c >> 6) == w) ? (1ull << (c & 63u)) : 0ull;
		}
	};
	if (inTile) {
		// SELL slice: word j of the 32 rows is one contiguous 256-byte load for the warp
		const uint32_t so0 = tv.soff[slot >> 5], nW = (tv.soff[(slot >> 5) + 1] - so0) >> 5;
		const uint2 * gq = bl.gidxS + so0 + lane;
		uint32_t degL = 0u;                                   // (only the kPad == false instance tests it)
		if (!kPad && light) degL = a.rowptr[tv.v0 + lv + 1] - a.rowptr[tv.v0 + lv];
#ifndef MCMCB200_MASK_PF
#define MCMCB200_MASK_PF 6
#endif
		constexpr uint32_t kMP = MCMCB200_MASK_PF;           // words fetched up front (covers degree <= 4*kMP in one round trip)
		uint2 q[kMP];
#pragma unroll
		for (uint32_t j = 0; j < kMP; ++j) if (j < nW) q[j] = __ldcs(gq + (size_t)j * 32);
#pragma unroll
		for (uint32_t j = 0; j < kMP; ++j) {
			if (j < nW) {
				const uint32_t p0 = 4u * j;
				if (kPad || p0 < degL) addc(q[j].x & 0xffffu);
				if (kPad || p0 + 1u < degL) addc(q[j].x >> 16);
				if (kPad || p0 + 2u < degL) addc(q[j].y & 0xffffu);
				if (kPad || p0 + 3u < degL) addc(q[j].y >> 16);
			}
		}
		for (uint32_t j = kMP; j < nW; j += 2) {
			const uint2 q0 = __ldcs(gq + (size_t)j * 32);
			uint2 q1 = make_uint2(0u, 0u);
			const bool two = j + 1u < nW;
			if (two) q1 = __ldcs(gq + (size_t)(j + 1u) * 32);
			const uint32_t p0 = 4u * j;
			if (kPad || p0 < degL) addc(q0.x & 0xffffu);
			if (kPad || p0 + 1u < degL) addc(q0.x >> 16);
			if (kPad || p0 + 2u < degL) addc(q0.y & 0xffffu);
			if (kPad || p0 + 3u < degL) addc(q0.y >> 16);
			if (kPad ? two : (p0 + 4u < degL)) addc(q1.x & 0xffffu);
			if (kPad ? two : (p0 + 5u < degL)) addc(q1.x >> 16);
			if (kPad ? two : (p0 + 6u < degL)) addc(q1.y & 0xffffu);
			if (kPad ? two : (p0 + 7u < degL)) addc(q1.y >> 16);
		}
	}
	if (inlineHeavy) {                                        // overflow of the heavy list: plain CSR-order indices
		const uint32_t myBeg = a.rowptr[tv.v0 + lv], deg = a.rowptr[tv.v0 + lv + 1] - myBeg;
		for (uint32_t i = 0; i < deg; ++i) addc(__ldg(bl.gidx + myBeg + i));
		light = true;
	}
	if (light)
		commit_vertex<W, ColT, kDyn, kPlain>(a, t, nxt, a.vBegin + tv.v0 + lv, tv.v0 + lv, own, m, same, sm.S, sm.dist, sm.hist, stayW, accDirected, accViol,
		                             wq, tv.snew, tv.v0, tv.draw);
	if (wq != nullptr) {                                      // this warp walks 32 parked vertices at a time: dense lanes, no CTA barrier
		__syncwarp();
		const uint32_t qn = min(*wq->count, wq->cap);
		if (qn >= 32u) {
			drain_walk_queue<W, ColT, kDyn>(a, nxt, *wq, qn - 32u, 32u, sm.dist, sm.hist, lane, tv.snew, tv.v0);
			__syncwarp();
			if (lane == 0) *wq->count = qn - 32u;
		}
		__syncwarp();
	}
}

// warp per heavy vertex (deg > kLightMaxDeg): lanes stride the row through the plain CSR-order index array
template <int W, typename ColT, bool kDyn, bool kPlain>
__device__ __forceinline__ void sweep_heavy_list(const SweepArgs & a, const BlockedArgs & bl, const PassBShared<W, ColT> & sm, const TileView<W, ColT> & tv,
                                                 uint32_t t, ColT * __restrict__ nxt, uint32_t nHeavy, int warp, int nWarps, int lane, float stayW,
                                                 unsigned long long & accDirected, unsigned long long & accViol) {
	for (uint32_t h = warp; h < nHeavy; h += nWarps) {
		const uint32_t hv = sm.heavy[h];
		const uint32_t hb = a.rowptr[tv.v0 + hv], hd = a.rowptr[tv.v0 + hv + 1] - hb;
		const uint32_t gv = a.vBegin + tv.v0 + hv;
		const uint32_t own = (uint32_t)tv.own[hv];
		unsigned long long m[W];
#pragma unroll
		for (int w = 0; w < W; ++w) m[w] = 0ull;
		uint32_t same = 0;
		for (uint32_t i = lane; i < hd; i += 32) {
			const uint32_t c = tv.stage[__ldg(bl.gidx + hb + i)];
			MCMCB200_CHECK(__ldg(bl.gidx + hb + i) < bl.stageCap + 16u && c < a.nCol, a.st);
			same += (c == own);
#pragma unroll
			for (int w = 0; w < W; ++w) m[w] |= ((int)(c >> 6) == w) ? (1ull << (c & 63u)) : 0ull;
		}
#pragma unroll
		for (int w = 0; w < W; ++w) m[w] = warp_reduce_or64(m[w]);
		same = __reduce_add_sync(0xffffffffu, same);
		if (lane == 0)
			commit_vertex<W, ColT, kDyn, kPlain>(a, t, nxt, gv, tv.v0 + hv, own, m, same, sm.S, sm.dist, sm.hist, stayW, accDirected, accViol,
			                             nullptr, tv.snew, tv.v0, tv.draw);
	}
}

// the tile's new colours, coalesced 16-byte stores: to the local replica and -- fused exchange -- straight into every peer
// GPU's replica over NVLink (peer pointers from cudaIpcOpenMemHandle); tiles start 256-vertex aligned
template <typename ColT>
__device__ __forceinline__ void write_out_tile(const SweepArgs & a, uint32_t t, ColT * nxt, const ColT * snew, uint32_t v0, uint32_t nv,
                                               uint32_t thr, uint32_t nThr) {
	const size_t byteOff = (size_t)(a.vBegin + v0) * sizeof(ColT);
	const uint32_t nBytes = nv * (uint32_t)sizeof(ColT), nVec = nBytes >> 4;
	const uint4 * src4 = reinterpret_cast<const uint4 *>(snew);
	const uint32_t nDest = a.nPeers ? a.nPeers : 1u;
	for (uint32_t d = 0; d < nDest; ++d) {
		unsigned char * dst = a.nPeers ? static_cast<unsigned char *>(a.peerColors[(t + 1) & 1][d]) : reinterpret_cast<unsigned char *>(nxt);
		uint4 * dst4 = reinterpret_cast<uint4 *>(dst + byteOff);
		for (uint32_t i = thr; i < nVec; i += nThr) dst4[i] = src4[i];
		for (uint32_t i = (nVec << 4) + thr; i < nBytes; i += nThr) dst[byteOff + i] = reinterpret_cast<const unsigned char *>(snew)[i];
	}
}

template <int W, typename ColT, bool kDyn, bool kPlain = false>
__global__ void __launch_bounds__(PassB<W>::threads) __maxnreg__(PassB<W>::maxRegs)
blocked_sweep_kernel(const SweepArgs a, const BlockedArgs bl) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int kT = PassB<W>::threads;
	const uint32_t nCol = a.nCol, TV = bl.TV, nbuf = bl.nbuf;
	const PassBShared<W, ColT> sm(smem_raw, nCol, nbuf, TV, bl.stageCap);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr bool useQueue = W <= 2;
	WalkQueue<W> wq{};
	if (useQueue) {                                           // this warp's private queue
		constexpr uint32_t qcap = warp_queue_cap(W);
		constexpr size_t perWarp = (size_t)qcap * (8 * W + 16);
		unsigned char * qb = sm.queues + (size_t)warp * perWarp;
		wq.count = sm.qcnt + warp;
		wq.cap = qcap;
		wq.mask = reinterpret_cast<unsigned long long *>(qb);
		wq.lvOwn = reinterpret_cast<uint32_t *>(wq.mask + (size_t)qcap * W);
		wq.uw = reinterpret_cast<float *>(wq.lvOwn + 2 * qcap);
	}
	DevState * st = a.st;
	if (!a.countOnly && st->convergedAt >= 0) return;
	const uint32_t t = st->sweep;
	const ColT * __restrict__ cur = a.colorsOverride ? static_cast<const ColT *>(a.colorsOverride)
	                                                 : static_cast<const ColT *>(a.colors[t & 1]);
	ColT * __restrict__ nxt = static_cast<ColT *>(a.colors[(t + 1) & 1]);
	const float eps = a.eps;
	const float stayW = stay_weight<kDyn>(nCol, eps);
	const uint64_t polFirst = make_policy_evict_first();

	for (uint32_t k = tid; k < nCol; k += kT) sm.hist[k] = 0;
	if (tid == 0) {
		float s = 0.0f; sm.S[0] = 0.0f;
		for (uint32_t k = 0; k < nCol; ++k) { s = __fadd_rn(s, eps); sm.S[k + 1] = s; }
		sm.ctl[1] = 0u;                                       // heavy-list counter
		for (uint32_t b = 0; b < nbuf; ++b) mbar_init(sm.bars + b, 1u);
		mbar_fence_init();
	}
	if (tid < 16) for (uint32_t b = 0; b < nbuf; ++b) sm.stageBuf(b)[bl.stageCap + tid] = (ColT)~(ColT)0;   // the dummy colour of the padded SELL rows
	if (!a.countOnly) fill_proposal_table<kDyn>(a, t, sm.dist, tid, kT);
	unsigned long long accDirected = 0ull, accViol = 0ull;
	if (MCMCB200_TIMING && tid == 0) atomicMin(bl.dbgTimes + 2, global_ns());
	__syncthreads();

	// thread 0 is the producer: it takes the next tile (ascending order: the order pass A completes them in), waits until pass A
	// has delivered the tile's part and issues the bulk copies into stage buffer `buf`; the CTA consumes the buffers in turn.
	// (MCMCB200_EARLY_TICKET, single stage buffer) the ticket of the NEXT tile is taken while the current one is computed, and -- if
	// pass A has already delivered its part -- everything the tile will need (stage image, slot table, colours, SELL index words) is
	// pulled into L2 a whole tile ahead, so that the bulk copies at the end of the tile and the first index loads hit L2.
	uint32_t nextT = 0xffffffffu; bool nextPrefetched = false;
	auto produce = [&](uint32_t buf) {
		uint32_t Tn;
		bool idxDone = false;
		if (MCMCB200_EARLY_TICKET && nextT != 0xffffffffu) { Tn = nextT; idxDone = nextPrefetched; nextT = 0xffffffffu; }
		else Tn = atomicAdd(bl.sync + 1, 1u);
		if (Tn < bl.numTiles) {
			bool ok;
			if (MCMCB200_TIMING) { const unsigned long long w0 = global_ns(); ok = wait_part_ready(bl, Tn, st); atomicAdd(bl.dbgTimes + 4, global_ns() - w0); }
			else ok = wait_part_ready(bl, Tn, st);
			if (!ok) Tn = 0xffffffffu;                            // producer missing: abort (see wait_part_ready)
		}
		sm.tileOf[buf] = Tn;
		if (Tn < bl.numTiles) {
			fence_proxy_async();                                  // pass A's stores (acquired above) and this CTA's reads of the buffer precede the copies
			stage_tile_tma<W, ColT>(a, bl, sm, cur, Tn, buf, sm.bars + buf, polFirst, !idxDone);
		} else mbar_arrive(sm.bars + buf);                        // nothing to copy: complete the phase so that the consumers see tileOf
	};
	if (tid == 0) for (uint32_t b = 0; b < nbuf; ++b) produce(b);

	for (uint32_t it = 0;; ++it) {
		const uint32_t buf = (nbuf == 2u) ? (it & 1u) : 0u;
		mbar_wait(sm.bars + buf, ((nbuf == 2u) ? (it >> 1) : it) & 1u);      // the tile is staged
		const uint32_t T = sm.tileOf[buf];
		if (T >= bl.numTiles) { if (MCMCB200_TIMING && tid == 0) atomicMax(bl.dbgTimes + 3, global_ns()); break; }
		TileView<W, ColT> tv;
		tv.v0 = T * TV; tv.nv = min(TV, a.nLocal - tv.v0);
		tv.slot = sm.slotTab(buf); tv.soff = sm.sliceTab(buf); tv.own = sm.ownCol(buf); tv.stage = sm.stageBuf(buf); tv.snew = sm.newCol(buf);
		tv.heavyCount = sm.ctl + 1;
		tv.draw = (a.tape || a.countOnly) ? nullptr : sm.draw;
		if (MCMCB200_EARLY_TICKET && nbuf == 1u && tid == 0) {
			nextT = atomicAdd(bl.sync + 1, 1u);
			nextPrefetched = false;
			if (nextT < bl.numTiles && part_ready_now(bl, nextT)) { prefetch_tile_l2<ColT>(a, bl, cur, nextT); nextPrefetched = true; }
		}
		// the tile's draws: one Philox4x32-10 call per 4 consecutive vertices (tiles start 256-aligned), parked in shared memory
		if (tv.draw) {
			const uint32_t g0 = (a.vBegin + tv.v0) >> 2;
			uint4 * d4 = reinterpret_cast<uint4 *>(sm.draw);
			for (uint32_t i = tid; i < ((tv.nv + 3u) >> 2); i += kT) d4[i] = philox4(a.seed, t + 1u, g0 + i, 0u);
		}
		if (useQueue && lane == 0) *wq.count = 0u;
		__syncthreads();                                      // draws are in; the previous tile's write-out has left s_new (nbuf == 1)
		for (uint32_t g = 0; g < tv.nv; g += kT)
			sweep_slot<W, ColT, kDyn, kPlain>(a, bl, sm, tv, t, nxt, g + tid, lane, stayW, useQueue ? &wq : nullptr, accDirected, accViol);
		if (useQueue) {
			__syncwarp();
			const uint32_t qn = min(*wq.count, wq.cap);
			drain_walk_queue<W, ColT, kDyn>(a, nxt, wq, 0u, qn, sm.dist, sm.hist, lane, tv.snew, tv.v0);
			__syncwarp();
		}
		__syncthreads();
		const uint32_t nHeavy = min(sm.ctl[1], kHeavyCap);
		sweep_heavy_list<W, ColT, kDyn, kPlain>(a, bl, sm, tv, t, nxt, nHeavy, warp, kT / 32, lane, stayW, accDirected, accViol);
		__syncthreads();                                      // the tile is finished: everybody has read the heavy list and the stage buffer, s_new is complete
		if (tid == 0) { sm.ctl[1] = 0u; produce(buf); }       // refill this buffer (nbuf == 2: the other one is already in flight)
		if (!a.countOnly) write_out_tile<ColT>(a, t, nxt, tv.snew, tv.v0, tv.nv, (uint32_t)tid, (uint32_t)kT);
	}

	// ---- epilogue (same protocol as sweep_kernel) ----
	accDirected = warp_reduce_add64(accDirected);
	accViol = warp_reduce_add64(accViol);
	__syncthreads();
	if (lane == 0) { sm.red[warp] = accDirected; sm.red[32 + warp] = accViol; }
	__syncthreads();
	if (tid == 0) {
		unsigned long long d = 0, vv = 0;
		for (int w = 0; w < kT / 32; ++w) { d += sm.red[w]; vv += sm.red[32 + w]; }
		if (d) atomicAdd(a.scratch + 0, d);
		if (vv) atomicAdd(a.scratch + 1, vv);
	}
	if (!a.countOnly) {
		for (uint32_t k = tid; k < nCol; k += kT) {
			const int dlt = sm.hist[k];
			if (dlt) atomicAdd(a.scratch + 2 + k, (unsigned long long)(long long)dlt);
		}
	}
	if (a.fuseFinalize) {
		if (a.nPeers) __threadfence_system(); else __threadfence();   // (peer colour stores of this CTA: visible system-wide before the ticket)
		__syncthreads();
		if (tid == 0) sm.ctl[3] = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
		__syncthreads();
		if (sm.ctl[3]) {
			__threadfence();
			if (a.nPeers) cross_rank_reduce(a);               // all-reduce + inter-rank barrier over NVLink, inside the kernel
			finalize_sweep_device(a);
		}
	}
}

} // namespace mcmcb200
