// Degree-binned direct-gather sweep: the path for large SKEWED graphs (R-MAT / power-law, BASELINE config 4), where a
// 256-vertex tile can hold a hub row and neither the tile-synchronous direct kernel (ncu on config 4: 53 % of the stall
// samples are CTA barriers behind a few long rows) nor the source-blocked layout (a row must fit one tile's stage) works.
//
// Static, once per graph: three ascending vertex lists by degree -- thread rows (deg <= 32), warp rows (33..4096), CTA rows.
// One launch per sweep, persistent CTAs, work taken from three atomic counters, no CTA barrier outside the CTA-row phase:
//   CTA rows    all 256 threads stride the row (coalesced neighbour ids, colour gathers with the evict_last policy);
//   warp rows   a warp takes 32 vertices; the 32 lanes stride each row in turn, lane j keeps the reduced mask of row j, then
//               all 32 lanes run commit_vertex together (dense phase 3);
//   thread rows a lane walks its own row, 4 neighbour ids + 4 colour gathers in flight.
// Conflicting vertices park their CDF walk in a per-warp shared-memory queue that is walked 32 entries at a time (with 512
// colours a walk is ~2400 instructions).  Wide palettes accumulate their masks in shared memory (see sweep_kernel.cuh).
// Result: identical to sweep_kernel / blocked_sweep_kernel (same commit_vertex, same counters, same finalize).
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "sweep_kernel.cuh"

namespace mcmcb200 {

constexpr int      kThreadsBin   = 256;
constexpr uint32_t kBinThreadMax = 32;      // thread rows: degree <= 32
constexpr uint32_t kBinWarpMax   = 4096;    // warp rows: degree <= 4096; longer rows take the whole CTA
constexpr uint32_t kBinQueueCap  = 48;
#ifndef MCMCB200_BIN_UNROLL
#define MCMCB200_BIN_UNROLL 4
#endif
constexpr int      kBinUnroll    = MCMCB200_BIN_UNROLL;   // coalesced id loads (then colour gathers) in flight per lane in warp / CTA rows

struct BinnedLayout {
	bool       valid = false;
	uint32_t   n[3] = {0, 0, 0};              // thread / warp / CTA rows
	uint32_t * list[3] = {nullptr, nullptr, nullptr};
	uint32_t * counters = nullptr;           // [4] next index of each list
	int        grid = 0;
	size_t     smem = 0;
};

struct BinnedArgs {
	const uint32_t * list[3];
	uint32_t n[3];
	uint32_t * counters;
};

struct DegreeInBin {
	const uint32_t * rowptr; uint32_t lo, hi;   // lo <= degree <= hi
	__host__ __device__ bool operator()(uint32_t v) const { const uint32_t d = rowptr[v + 1] - rowptr[v]; return d >= lo && d <= hi; }
};

inline void free_binned_layout(BinnedLayout & L) {
	for (int i = 0; i < 3; ++i) cudaFree(L.list[i]);
	cudaFree(L.counters);
	L = BinnedLayout{};
}

inline cudaError_t build_binned_layout(BinnedLayout & L, const uint32_t * d_rowptr, uint32_t nLocal, cudaStream_t stream, uint64_t * launches) {
	L = BinnedLayout{};
	cudaError_t err = cudaSuccess;
	if (nLocal == 0) {                                            // an empty partition (a rank of a very skewed graph): three empty lists
		for (int b = 0; b < 3; ++b) if ((err = cudaMalloc(&L.list[b], sizeof(uint32_t))) != cudaSuccess) { free_binned_layout(L); return err; }
		if ((err = cudaMalloc(&L.counters, 4 * sizeof(uint32_t))) != cudaSuccess) { free_binned_layout(L); return err; }
		L.valid = true;
		return cudaSuccess;
	}
	uint32_t * d_num = nullptr;
	void * d_tmp = nullptr; size_t tmpBytes = 0;
	const uint32_t lo[3] = {0u, kBinThreadMax + 1u, kBinWarpMax + 1u}, hi[3] = {kBinThreadMax, kBinWarpMax, 0xffffffffu};
	thrust::counting_iterator<uint32_t> ids(0u);
	uint32_t * scratch = nullptr;
	if ((err = cudaMalloc(&d_num, sizeof(uint32_t))) != cudaSuccess) goto done;
	if ((err = cudaMalloc(&scratch, sizeof(uint32_t) * (size_t)nLocal)) != cudaSuccess) goto done;
	if ((err = cub::DeviceSelect::If(nullptr, tmpBytes, ids, scratch, d_num, (int)nLocal, DegreeInBin{d_rowptr, 0u, 0u}, stream)) != cudaSuccess) goto done;
	if ((err = cudaMalloc(&d_tmp, tmpBytes)) != cudaSuccess) goto done;
	for (int b = 0; b < 3; ++b) {
		if ((err = cub::DeviceSelect::If(d_tmp, tmpBytes, ids, scratch, d_num, (int)nLocal, DegreeInBin{d_rowptr, lo[b], hi[b]}, stream)) != cudaSuccess) goto done;
		(*launches) += 2;
		if ((err = cudaMemcpyAsync(&L.n[b], d_num, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) goto done;
		if ((err = cudaStreamSynchronize(stream)) != cudaSuccess) goto done;
		if ((err = cudaMalloc(&L.list[b], sizeof(uint32_t) * std::max<size_t>(L.n[b], 1))) != cudaSuccess) goto done;
		if ((err = cudaMemcpyAsync(L.list[b], scratch, sizeof(uint32_t) * (size_t)L.n[b], cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) goto done;
	}
	if ((err = cudaMalloc(&L.counters, 4 * sizeof(uint32_t))) != cudaSuccess) goto done;
	if ((err = cudaStreamSynchronize(stream)) != cudaSuccess) goto done;
	L.valid = (uint64_t)L.n[0] + L.n[1] + L.n[2] == nLocal;
done:
	cudaFree(d_num); cudaFree(scratch); cudaFree(d_tmp);
	if (err != cudaSuccess || !L.valid) { cudaError_t keep = err; free_binned_layout(L); err = keep; }
	if (err == cudaErrorMemoryAllocation) { cudaGetLastError(); err = cudaSuccess; }   // no room for the lists: the tile-synchronous kernel needs none
	return err;
}

inline BinnedArgs make_binned_args(const BinnedLayout & L) {
	BinnedArgs b{};
	for (int i = 0; i < 3; ++i) { b.list[i] = L.list[i]; b.n[i] = L.n[i]; }
	b.counters = L.counters;
	return b;
}

__host__ __device__ inline size_t binned_smem_bytes(uint32_t nCol, int W) {
	size_t b = 0;
	b += sizeof(float) * (size_t)((nCol + 1 + 3) & ~3u);   // s_S
	b += sizeof(float) * (size_t)((nCol + 3) & ~3u);       // s_dist
	b += sizeof(int) * (size_t)((nCol + 3) & ~3u);         // s_hist
	b += sizeof(uint32_t) * 16;                            // s_ctl (8) + per-warp queue counters (8)
	b += sizeof(unsigned long long) * 16;                  // s_red
	b += sizeof(unsigned long long) * (size_t)W;           // s_hub: mask of the current CTA row
	b = (b + 15) & ~(size_t)15;
	b += (size_t)(kThreadsBin / 32) * kBinQueueCap * (8 * W + 16);   // per-warp walk queues
	if (W > 2) {
		b += sizeof(uint32_t) * (kThreadsBin / 32) * 2 * W;            // s_wm: per-warp accumulators
		b += sizeof(uint32_t) * (size_t)kThreadsBin * 2 * W;           // s_m32: per-thread mask rows (transposed)
	}
	return (b + 15) & ~(size_t)15;
}

template <int W, typename ColT, bool kDyn>
__global__ void __launch_bounds__(kThreadsBin, (W <= 2 ? 4 : 3))
binned_sweep_kernel(const SweepArgs a, const BinnedArgs bn) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr bool kWide = W > 2;
	const uint32_t nCol = a.nCol;
	float *    s_S    = reinterpret_cast<float *>(smem_raw);
	float *    s_dist = s_S + ((nCol + 1 + 3) & ~3u);
	int *      s_hist = reinterpret_cast<int *>(s_dist + ((nCol + 3) & ~3u));
	uint32_t * s_ctl  = reinterpret_cast<uint32_t *>(s_hist + ((nCol + 3) & ~3u));
	uint32_t * s_qcnt = s_ctl + 8;
	unsigned long long * s_red = reinterpret_cast<unsigned long long *>(s_ctl + 16);
	unsigned long long * s_hub = s_red + 16;
	size_t off = (size_t)(reinterpret_cast<unsigned char *>(s_hub + W) - smem_raw);
	off = (off + 15) & ~(size_t)15;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int nWarps = kThreadsBin / 32;
	WalkQueue<W> wq{};
	{
		constexpr size_t perWarp = (size_t)kBinQueueCap * (8 * W + 16);
		unsigned char * qb = smem_raw + off + (size_t)warp * perWarp;
		wq.count = s_qcnt + warp;
		wq.cap = kBinQueueCap;
		wq.mask = reinterpret_cast<unsigned long long *>(qb);
		wq.lvOwn = reinterpret_cast<uint32_t *>(wq.mask + (size_t)kBinQueueCap * W);
		wq.uw = reinterpret_cast<float *>(wq.lvOwn + 2 * kBinQueueCap);
		off += (size_t)nWarps * perWarp;
	}
	uint32_t * s_wm = reinterpret_cast<uint32_t *>(smem_raw + off);                  // [nWarps][2W]   (wide only)
	uint32_t * s_m32 = s_wm + (kWide ? nWarps * 2 * W : 0);                         // [2W][kThreadsBin] (wide only)

	DevState * st = a.st;
	if (!a.countOnly && st->convergedAt >= 0) return;
	const uint32_t t = st->sweep;
	const ColT * __restrict__ cur = a.colorsOverride ? static_cast<const ColT *>(a.colorsOverride)
	                                                 : static_cast<const ColT *>(a.colors[t & 1]);
	ColT * __restrict__ nxt = static_cast<ColT *>(a.colors[(t + 1) & 1]);
	const float eps = a.eps;
	const float stayW = stay_weight<kDyn>(nCol, eps);
	const uint64_t polLast = make_policy_evict_last();

	for (uint32_t k = tid; k < nCol; k += kThreadsBin) s_hist[k] = 0;
	if (tid == 0) {
		float s = 0.0f; s_S[0] = 0.0f;
		for (uint32_t k = 0; k < nCol; ++k) { s = __fadd_rn(s, eps); s_S[k + 1] = s; }
	}
	if (lane == 0) *wq.count = 0u;
	if (!a.countOnly) fill_proposal_table<kDyn>(a, t, s_dist, tid, kThreadsBin);
	unsigned long long accDirected = 0ull, accViol = 0ull;
	__syncthreads();

	auto set_bit = [&](unsigned long long (&mm)[W], uint32_t c) {
		if (W == 1) mm[0] |= 1ull << c;
		else {
#pragma unroll
			for (int w = 0; w < W; ++w) mm[w] |= ((int)(c >> 6) == w) ? (1ull << (c & 63u)) : 0ull;
		}
	};
	auto drain_if_full = [&]() {                                  // this warp walks 32 parked vertices at a time
		__syncwarp();
		const uint32_t qn = min(*wq.count, wq.cap);
		if (qn >= 32u) {
			drain_walk_queue<W, ColT, kDyn>(a, nxt, wq, qn - 32u, 32u, s_dist, s_hist, lane);
			__syncwarp();
			if (lane == 0) *wq.count = qn - 32u;
		}
		__syncwarp();
	};

	// ---------------- CTA rows ----------------
	for (;;) {
		__syncthreads();
		if (tid == 0) s_ctl[0] = atomicAdd(bn.counters + 2, 1u);
		if (tid < W) s_hub[tid] = 0ull;
		if (tid == 0) s_ctl[1] = 0u;
		if (kWide && lane < 2 * W) s_wm[warp * 2 * W + lane] = 0u;
		__syncthreads();
		const uint32_t i = s_ctl[0];
		if (i >= bn.n[2]) break;
		const uint32_t lv = bn.list[2][i];
		const uint32_t e0 = a.rowptr[lv], e1 = a.rowptr[lv + 1];
		const uint32_t own = (uint32_t)cur[a.vBegin + lv];
		unsigned long long m[W];
#pragma unroll
		for (int w = 0; w < W; ++w) m[w] = 0ull;
		uint32_t same = 0;
		for (uint32_t e = e0 + tid; e < e1; e += (uint32_t)kBinUnroll * kThreadsBin) {   // kBinUnroll coalesced id loads, then as many gathers, in flight
			uint32_t nb[kBinUnroll], c[kBinUnroll];
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) nb[k] = (e + k * kThreadsBin < e1) ? __ldcs(a.neighs + e + k * kThreadsBin) : 0u;
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) c[k] = (e + k * kThreadsBin < e1) ? ld_color<ColT>(cur + nb[k], polLast) : 0xffffffffu;
#pragma unroll
			for (int k = 0; k < kBinUnroll; ++k) {
				if (e + k * kThreadsBin < e1) {
					same += (c[k] == own);
					if (kWide) atomicOr(&s_wm[warp * 2 * W + (c[k] >> 5)], 1u << (c[k] & 31u));
					else set_bit(m, c[k]);
				}
			}
		}
		if (kWide) {
			__syncwarp();
			if (lane < 2 * W) {
				const uint32_t r = s_wm[warp * 2 * W + lane];
				if (r) atomicOr(reinterpret_cast<uint32_t *>(s_hub) + lane, r);      // little endian: word 2w = low half of mask word w
			}
		} else {
#pragma unroll
			for (int w = 0; w < W; ++w) {
				const unsigned long long r = warp_reduce_or64(m[w]);
				if (lane == 0 && r) atomicOr(&s_hub[w], r);
			}
		}
		same = __reduce_add_sync(0xffffffffu, same);
		if (lane == 0 && same) atomicAdd(&s_ctl[1], same);
		__syncthreads();
		if (tid == 0) {
#pragma unroll
			for (int w = 0; w < W; ++w) m[w] = s_hub[w];
			commit_vertex<W, ColT, kDyn>(a, t, nxt, a.vBegin + lv, lv, own, m, s_ctl[1], s_S, s_dist, s_hist, stayW, accDirected, accViol);
		}
	}

	// ---------------- warp rows: 32 vertices per warp and batch ----------------
	for (;;) {
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(bn.counters + 1, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= bn.n[1]) break;
		const uint32_t cntB = min(32u, bn.n[1] - base);
		const bool valid = (uint32_t)lane < cntB;
		const uint32_t lvMine = valid ? bn.list[1][base + lane] : 0u;
		const uint32_t begMine = valid ? a.rowptr[lvMine] : 0u, endMine = valid ? a.rowptr[lvMine + 1] : 0u;
		const uint32_t ownMine = valid ? (uint32_t)cur[a.vBegin + lvMine] : 0u;
		unsigned long long m[W];
#pragma unroll
		for (int w = 0; w < W; ++w) m[w] = 0ull;
		uint32_t same = 0;
		for (uint32_t j = 0; j < cntB; ++j) {
			const uint32_t e0 = __shfl_sync(0xffffffffu, begMine, j), e1 = __shfl_sync(0xffffffffu, endMine, j);
			const uint32_t ownJ = __shfl_sync(0xffffffffu, ownMine, j);
			unsigned long long mm[W];
#pragma unroll
			for (int w = 0; w < W; ++w) mm[w] = 0ull;
			uint32_t ss = 0;
			if (kWide) { if (lane < 2 * W) s_wm[warp * 2 * W + lane] = 0u; __syncwarp(); }
			for (uint32_t e = e0 + lane; e < e1; e += 32u * kBinUnroll) {   // kBinUnroll coalesced id loads, then as many gathers, in flight
				uint32_t nb[kBinUnroll], c[kBinUnroll];
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) nb[k] = (e + 32u * k < e1) ? __ldcs(a.neighs + e + 32u * k) : 0u;
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) c[k] = (e + 32u * k < e1) ? ld_color<ColT>(cur + nb[k], polLast) : 0xffffffffu;
#pragma unroll
				for (int k = 0; k < kBinUnroll; ++k) {
					if (e + 32u * k < e1) {
						MCMCB200_CHECK(c[k] < nCol && nb[k] < a.nGlobal, st);
						ss += (c[k] == ownJ);
						if (kWide) atomicOr(&s_wm[warp * 2 * W + (c[k] >> 5)], 1u << (c[k] & 31u));
						else set_bit(mm, c[k]);
					}
				}
			}
			ss = __reduce_add_sync(0xffffffffu, ss);
			if (kWide) {
				__syncwarp();
				if ((uint32_t)lane == j) {
#pragma unroll
					for (int w = 0; w < W; ++w)
						m[w] = (unsigned long long)s_wm[warp * 2 * W + 2 * w] | ((unsigned long long)s_wm[warp * 2 * W + 2 * w + 1] << 32);
					same = ss;
				}
				__syncwarp();
			} else {
#pragma unroll
				for (int w = 0; w < W; ++w) mm[w] = warp_reduce_or64(mm[w]);
				if ((uint32_t)lane == j) {
#pragma unroll
					for (int w = 0; w < W; ++w) m[w] = mm[w];
					same = ss;
				}
			}
		}
		if (valid)
			commit_vertex<W, ColT, kDyn>(a, t, nxt, a.vBegin + lvMine, lvMine, ownMine, m, same, s_S, s_dist, s_hist, stayW, accDirected, accViol, &wq);
		drain_if_full();
	}

	// ---------------- thread rows ----------------
	for (;;) {
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(bn.counters + 0, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= bn.n[0]) break;
		const bool valid = base + lane < bn.n[0];
		const uint32_t lv = valid ? bn.list[0][base + lane] : 0u;
		const uint32_t beg = valid ? a.rowptr[lv] : 0u, deg = valid ? (a.rowptr[lv + 1] - beg) : 0u;
		const uint32_t own = valid ? (uint32_t)cur[a.vBegin + lv] : 0u;
		unsigned long long m[W];
#pragma unroll
		for (int w = 0; w < W; ++w) m[w] = 0ull;
		uint32_t same = 0;
		if (kWide) {
#pragma unroll
			for (int w = 0; w < 2 * W; ++w) s_m32[w * kThreadsBin + tid] = 0u;
		}
		for (uint32_t i = 0; i < deg; i += 4u) {
			uint32_t nb[4], c[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) nb[k] = (i + k < deg) ? __ldg(a.neighs + beg + i + k) : 0u;   // (the row's sector serves the next iterations from L1)
#pragma unroll
			for (int k = 0; k < 4; ++k) c[k] = (i + k < deg) ? ld_color<ColT>(cur + nb[k], polLast) : 0xffffffffu;
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				if (i + k < deg) {
					MCMCB200_CHECK(c[k] < nCol && nb[k] < a.nGlobal, st);
					same += (c[k] == own);
					if (kWide) s_m32[(c[k] >> 5) * kThreadsBin + tid] |= 1u << (c[k] & 31u);
					else set_bit(m, c[k]);
				}
			}
		}
		if (kWide) {
#pragma unroll
			for (int w = 0; w < W; ++w)
				m[w] = (unsigned long long)s_m32[(2 * w) * kThreadsBin + tid] | ((unsigned long long)s_m32[(2 * w + 1) * kThreadsBin + tid] << 32);
		}
		if (valid)
			commit_vertex<W, ColT, kDyn>(a, t, nxt, a.vBegin + lv, lv, own, m, same, s_S, s_dist, s_hist, stayW, accDirected, accViol, &wq);
		drain_if_full();
	}
	{                                                             // remainder of this warp's queue
		__syncwarp();
		const uint32_t qn = min(*wq.count, wq.cap);
		drain_walk_queue<W, ColT, kDyn>(a, nxt, wq, 0u, qn, s_dist, s_hist, lane);
		__syncwarp();
	}

	// ---- epilogue (same protocol as sweep_kernel) ----
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
	if (!a.countOnly) {
		for (uint32_t k = tid; k < nCol; k += kThreadsBin) {
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

} // namespace mcmcb200
