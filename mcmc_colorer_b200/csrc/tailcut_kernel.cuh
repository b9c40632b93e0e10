// Tail cutting on the device.  Replaces the reference's single-thread greedy repair
//   tailCutting<<<1,1>>>                       graph_coloring/coloringMCMC_utils.cu:73-101
//   driver loop + colour ordering              graph_coloring/coloringMCMC_main.cu:271-290
// Reference semantics of ONE pass: vertices flagged by conflictCounter (a same-coloured neighbour with a LARGER
// id, _utils.cu:115) are visited in ascending id order; each takes the first colour, in ascending class-size
// order, that no neighbour currently has (in-place, Gauss-Seidel).  The parallel version below produces the SAME
// result: a flagged vertex is processed in the round in which it has no still-pending flagged neighbour with a
// smaller id (that neighbour would have been visited before it by the sequential loop); two vertices processed
// in the same round are therefore never adjacent, and each sees exactly the colours the sequential loop would.
#pragma once
#include "sweep_kernel.cuh"

namespace mcmcb200 {

// rows longer than this are repaired by a whole CTA with an occupancy bitmap in shared memory (nCol bits); shorter ones by a warp
// that re-scans the row for every candidate colour (a hub row of 10^6 neighbours times a palette of 10^3 candidates would not end)
constexpr uint32_t kTcHeavyDeg = 2048;
constexpr uint32_t kTcThreads = 1024;

// warp: does any neighbour of the row [e0, e1) satisfy pred(u)?  (lanes stride the row, coalesced; early exit per 32 entries)
template <typename Pred>
__device__ __forceinline__ bool tc_warp_any(const uint32_t * __restrict__ neighs, uint32_t e0, uint32_t e1, int lane, Pred pred) {
	for (uint32_t e = e0; e < e1; e += 32u) {
		const bool hit = (e + lane < e1) && pred(neighs[e + lane]);
		if (__any_sync(0xffffffffu, hit)) return true;
	}
	return false;
}

// the reference's colour choice (_utils.cu:89-97): keep the colour if no neighbour has it, else the first of order[0 .. nCol-2]
// that no neighbour has, else order[nCol-1] untested.  Warp version: one row scan per candidate.  *inexact: the result clashes.
template <typename ColT>
__device__ __forceinline__ uint32_t tc_pick_warp(const uint32_t * __restrict__ neighs, uint32_t e0, uint32_t e1, const ColT * colors,
                                                 uint32_t old, uint32_t nCol, const uint32_t * __restrict__ order, int lane, bool * inexact) {
	auto occupied = [&](uint32_t c) { return tc_warp_any(neighs, e0, e1, lane, [&](uint32_t u) { return (uint32_t)colors[u] == c; }); };
	uint32_t nodeCol = old, j = 0;
	while (occupied(nodeCol) && j < nCol) { nodeCol = order[j]; j++; }
	*inexact = (j == nCol) && occupied(nodeCol);
	return nodeCol;
}

// CTA version: occupancy bitmap of the row in shared memory (bm: (nCol+31)/32 words; s_best: one word), then a parallel search
// of the first free candidate.  Every thread of the CTA must call it; returns the same value to all.
template <typename ColT>
__device__ __forceinline__ uint32_t tc_pick_cta(const uint32_t * __restrict__ neighs, uint32_t e0, uint32_t e1, const ColT * colors,
                                                uint32_t old, uint32_t nCol, const uint32_t * __restrict__ order, uint32_t * bm, uint32_t * s_best,
                                                bool * inexact) {
	const uint32_t words = (nCol + 31u) / 32u;
	for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) bm[w] = 0u;
	if (threadIdx.x == 0) *s_best = 0xffffffffu;
	__syncthreads();
	for (uint32_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
		const uint32_t c = colors[neighs[e]], bit = 1u << (c & 31u);
		if (!(bm[c >> 5] & bit)) atomicOr(&bm[c >> 5], bit);
	}
	__syncthreads();
	auto occ = [&](uint32_t c) { return (bm[c >> 5] >> (c & 31u)) & 1u; };
	uint32_t res;
	if (!occ(old)) { res = old; *inexact = false; }
	else {
		for (uint32_t j = threadIdx.x; j + 1u < nCol; j += blockDim.x)
			if (!occ(order[j])) { atomicMin(s_best, j); break; }
		__syncthreads();
		const uint32_t b = *s_best;
		res = (b != 0xffffffffu) ? order[b] : order[nCol - 1u];
		if (MCMCB200_BOUNDS_CHECK && res >= nCol) __trap();
		*inexact = (b == 0xffffffffu) && occ(res);
	}
	__syncthreads();                                          // bm / s_best may be reused right away
	return res;
}

// ---- full-scan path: one reference pass over the whole graph ----
template <typename ColT>
__global__ void tailcut_flag_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, uint32_t n,
                                    const ColT * __restrict__ colors, uint8_t * pending, uint32_t * list, uint32_t * listCount) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	const uint32_t c = colors[v];
	bool flag = false;
	for (uint32_t e = rowptr[v]; e < rowptr[v + 1] && !flag; ++e) {
		const uint32_t u = neighs[e];
		flag = (u > v) && (colors[u] == c);
	}
	pending[v] = flag ? 1 : 0;
	if (flag) list[atomicAdd(listCount, 1u)] = v;
}

// warp per listed vertex
__global__ void tailcut_ready_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs,
                                     const uint8_t * __restrict__ pending, const uint32_t * __restrict__ list, uint32_t listCount,
                                     uint8_t * ready) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	if (!pending[v]) { if (lane == 0) ready[i] = 0; return; }
	const bool blocked = tc_warp_any(neighs, rowptr[v], rowptr[v + 1], lane, [&](uint32_t u) { return u < v && pending[u]; });
	if (lane == 0) ready[i] = blocked ? 0 : 1;
}

// warp per listed vertex; rows longer than kTcHeavyDeg are handed to tailcut_apply_heavy_kernel (heavy[0] = count, then ids)
template <typename ColT>
__global__ void tailcut_apply_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, uint32_t nCol,
                                     ColT * colors, uint8_t * pending, const uint32_t * __restrict__ list, uint32_t listCount,
                                     const uint8_t * __restrict__ ready, const uint32_t * __restrict__ order,
                                     unsigned long long * hist, uint32_t * remaining, uint32_t * heavy, uint32_t * changed,
                                     uint32_t * outIds = nullptr, uint32_t * outCols = nullptr, uint32_t * outCount = nullptr, uint32_t * inexactFlag = nullptr) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	if (!pending[v]) return;
	if (!ready[i]) { if (lane == 0) atomicAdd(remaining, 1u); return; }
	const uint32_t e0 = rowptr[v], e1 = rowptr[v + 1];
	if (e1 - e0 > kTcHeavyDeg) { if (lane == 0) heavy[1u + atomicAdd(heavy, 1u)] = v; return; }
	const uint32_t old = colors[v];
	bool inexact;
	const uint32_t nodeCol = tc_pick_warp<ColT>(neighs, e0, e1, colors, old, nCol, order, lane, &inexact);
	__syncwarp();
	if (lane == 0) {
		colors[v] = (ColT)nodeCol;                                           // :97
		pending[v] = 0;
		if (nodeCol != old) { atomicAdd(hist + old, ~0ull); atomicAdd(hist + nodeCol, 1ull); atomicAdd(changed, 1u); }
		if (outIds) { const uint32_t k = atomicAdd(outCount, 1u); outIds[k] = v; outCols[k] = nodeCol; }   // (distributed repair: told to the other ranks)
		if (inexact && inexactFlag) *inexactFlag = 1u;
	}
}

template <typename ColT>
__global__ void __launch_bounds__(kTcThreads)
tailcut_apply_heavy_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, uint32_t nCol, ColT * colors, uint8_t * pending,
                           const uint32_t * __restrict__ order, unsigned long long * hist, const uint32_t * heavy, uint32_t * changed,
                           uint32_t * outIds = nullptr, uint32_t * outCols = nullptr, uint32_t * outCount = nullptr, uint32_t * inexactFlag = nullptr) {
	extern __shared__ uint32_t tc_smem[];
	uint32_t * bm = tc_smem + 4;
	const uint32_t cnt = heavy[0];
	for (uint32_t k = blockIdx.x; k < cnt; k += gridDim.x) {
		const uint32_t v = heavy[1u + k];
		const uint32_t old = colors[v];
		bool inexact;
		const uint32_t nodeCol = tc_pick_cta<ColT>(neighs, rowptr[v], rowptr[v + 1], colors, old, nCol, order, bm, tc_smem, &inexact);
		if (threadIdx.x == 0) {
			colors[v] = (ColT)nodeCol;
			pending[v] = 0;
			if (nodeCol != old) { atomicAdd(hist + old, ~0ull); atomicAdd(hist + nodeCol, 1ull); atomicAdd(changed, 1u); }
			if (outIds) { const uint32_t k = atomicAdd(outCount, 1u); outIds[k] = v; outCols[k] = nodeCol; }
			if (inexact && inexactFlag) *inexactFlag = 1u;
		}
		__syncthreads();
	}
}

inline size_t tc_smem_bytes(uint32_t nCol) { return sizeof(uint32_t) * (4 + (size_t)(nCol + 31u) / 32u); }

// One reference pass.  Returns a cudaError_t as int; *flaggedOut = number of vertices the pass visited.
// d_heavy: [1 + n] scratch for the ready hub rows of a round (may alias nothing else).
inline int launch_tailcut_pass(cudaStream_t stream, int colBytes, const uint32_t * rowptr, const uint32_t * neighs, uint32_t n,
                               uint32_t nCol, void * colors, unsigned long long * hist, const uint32_t * d_order,
                               uint8_t * d_pending, uint8_t * d_ready, uint32_t * d_list, uint32_t * d_heavy, uint32_t * d_counters /* [3] */,
                               uint32_t * flaggedOut, uint32_t * changedOut, uint64_t * launches) {
	cudaError_t e;
	*changedOut = 0;
	if ((e = cudaMemsetAsync(d_counters, 0, 3 * sizeof(uint32_t), stream)) != cudaSuccess) return (int)e;
	const uint32_t blocks = (n + 255) / 256;
	if (colBytes == 1) tailcut_flag_kernel<uint8_t><<<blocks, 256, 0, stream>>>(rowptr, neighs, n, (const uint8_t *)colors, d_pending, d_list, d_counters);
	else tailcut_flag_kernel<uint16_t><<<blocks, 256, 0, stream>>>(rowptr, neighs, n, (const uint16_t *)colors, d_pending, d_list, d_counters);
	(*launches)++;
	uint32_t listCount = 0;
	if ((e = cudaMemcpyAsync(&listCount, d_counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
	if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
	*flaggedOut = listCount;
	if (listCount == 0) return 0;
	const uint32_t lb = (uint32_t)(((uint64_t)listCount * 32u + 255u) / 256u);       // warp per listed vertex
	const size_t smem = tc_smem_bytes(nCol);
	for (uint32_t guard = 0; guard <= listCount; ++guard) {
		if ((e = cudaMemsetAsync(d_counters + 1, 0, sizeof(uint32_t), stream)) != cudaSuccess) return (int)e;
		if ((e = cudaMemsetAsync(d_heavy, 0, sizeof(uint32_t), stream)) != cudaSuccess) return (int)e;
		tailcut_ready_kernel<<<lb, 256, 0, stream>>>(rowptr, neighs, d_pending, d_list, listCount, d_ready);
		if (colBytes == 1) {
			tailcut_apply_kernel<uint8_t><<<lb, 256, 0, stream>>>(rowptr, neighs, nCol, (uint8_t *)colors, d_pending, d_list, listCount, d_ready, d_order, hist, d_counters + 1, d_heavy, d_counters + 2);
			tailcut_apply_heavy_kernel<uint8_t><<<64, kTcThreads, smem, stream>>>(rowptr, neighs, nCol, (uint8_t *)colors, d_pending, d_order, hist, d_heavy, d_counters + 2);
		} else {
			tailcut_apply_kernel<uint16_t><<<lb, 256, 0, stream>>>(rowptr, neighs, nCol, (uint16_t *)colors, d_pending, d_list, listCount, d_ready, d_order, hist, d_counters + 1, d_heavy, d_counters + 2);
			tailcut_apply_heavy_kernel<uint16_t><<<64, kTcThreads, smem, stream>>>(rowptr, neighs, nCol, (uint16_t *)colors, d_pending, d_order, hist, d_heavy, d_counters + 2);
		}
		(*launches) += 3;
		uint32_t remaining = 0;
		if ((e = cudaMemcpyAsync(&remaining, d_counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
		if (remaining == 0) break;
	}
	if ((e = cudaMemcpyAsync(changedOut, d_counters + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
	if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
	return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// List-driven tail cutting (the production path).  The chain stops sweeping once at most z = max(50, n/2000) vertices violate
// (coloringMCMC_main.cu:150-170); the sweep that established this also emitted those vertices (SweepArgs::violList), so the
// repair touches z rows instead of the whole CSR and needs no host round trip per round:
//   tc_filter_kernel   the reference's flags (conflictCounter, _utils.cu:115: a same-coloured neighbour with a LARGER id)
//                      for the listed vertices only -- an unlisted vertex has no same-coloured neighbour at all;
//   tc_rounds_kernel   ONE CTA runs all rounds of the dependency-ordered repair above (ready / apply, __syncthreads between):
//                      a warp per vertex, the whole CTA with a shared-memory occupancy bitmap for rows beyond kTcHeavyDeg;
//   tc_recount_kernel  conflicts and violations of the repaired colouring, again from the listed rows only (a vertex that was
//                      not violating can only become so if a repaired neighbour found every colour taken -- flagged `inexact`,
//                      the caller then recounts with a full pass), and the violators that are left (the next pass's list).
// All three walk a row with the 32 lanes of a warp (coalesced), so hub rows cost deg/32 steps.
// ---------------------------------------------------------------------------------------------------------------
struct TailcutCounters {
	uint32_t flagged;            // entries of flist (this pass)
	uint32_t inexact;            // a repaired vertex found all nCol colours occupied
	uint32_t nextCount;          // violating vertices after the pass (entries of the next list)
	uint32_t nextFlagged;        // of those, vertices the reference would flag again
	unsigned long long directed; // sum over vertices of same-coloured neighbours (= 2 x conflicting edges)
	unsigned long long viol;     // violating vertices
	uint32_t changed;            // vertices the pass gave a new colour (0: the pass made no progress -- hubs with every colour taken)
	uint32_t pad;
};

template <typename ColT>
__global__ void tc_filter_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, const ColT * __restrict__ colors,
                                 const uint32_t * __restrict__ list, uint32_t listCount, uint8_t * pending, uint32_t * flist, TailcutCounters * cnt) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	const uint32_t c = colors[v];
	const bool flag = tc_warp_any(neighs, rowptr[v], rowptr[v + 1], lane, [&](uint32_t u) { return (u > v) && ((uint32_t)colors[u] == c); });
	if (flag && lane == 0) { pending[v] = 1; flist[atomicAdd(&cnt->flagged, 1u)] = v; }
}

template <typename ColT>
__global__ void __launch_bounds__(kTcThreads)
tc_rounds_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, uint32_t nCol, ColT * colors, uint8_t * pending,
                 const uint32_t * __restrict__ flist, uint32_t * resume /* [flagged], zeroed */, TailcutCounters * cnt, const uint32_t * __restrict__ order,
                 unsigned long long * hist) {
	extern __shared__ uint32_t tc_smem[];                        // [0] best, [1] left, [2] heavy ready this round, [4..] bitmap
	uint32_t * bm = tc_smem + 4;
	const uint32_t n = cnt->flagged;
	const int lane = threadIdx.x & 31;
	const uint32_t warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
	for (;;) {
		// ready: no still-pending flagged neighbour with a smaller id (the sequential loop would have visited it first)
		for (uint32_t i = warp; i < n; i += nWarps) {
			const uint32_t v = flist[i];
			if (pending[v] != 1) continue;
			// (flags only ever drop during a pass, so the part of the row that held no pending smaller neighbour in an earlier round
			//  need not be looked at again: a hub row is scanned once per pass, not once per round)
			const uint32_t e1 = rowptr[v + 1];
			uint32_t e = rowptr[v] + resume[i];
			for (; e < e1; e += 32u) {
				bool hit = false;
				if (e + lane < e1) { const uint32_t u = neighs[e + lane]; hit = u < v && pending[u]; }
				if (__any_sync(0xffffffffu, hit)) break;
			}
			__syncwarp();
			if (lane == 0) {
				if (e < e1) resume[i] = e - rowptr[v];
				else pending[v] = 3;                               // pending AND ready (still non-zero for its larger neighbours)
			}
		}
		if (threadIdx.x == 0) { tc_smem[1] = 0u; tc_smem[2] = 0u; }
		__syncthreads();
		// apply: ready vertices are pairwise non-adjacent, each sees exactly the colours the sequential loop would
		for (uint32_t i = warp; i < n; i += nWarps) {
			const uint32_t v = flist[i];
			const uint32_t p = pending[v];
			if (p == 1) { if (lane == 0) atomicAdd(&tc_smem[1], 1u); continue; }
			if (p != 3) continue;
			const uint32_t e0 = rowptr[v], e1 = rowptr[v + 1];
			if (e1 - e0 > kTcHeavyDeg) { if (lane == 0) atomicAdd(&tc_smem[2], 1u); continue; }   // the whole CTA does these below
			const uint32_t old = colors[v];
			bool inexact;
			const uint32_t nodeCol = tc_pick_warp<ColT>(neighs, e0, e1, colors, old, nCol, order, lane, &inexact);
			__syncwarp();
			if (lane == 0) {
				if (inexact) cnt->inexact = 1u;                        // every colour taken: may disturb an unlisted neighbour
				colors[v] = (ColT)nodeCol;                             // :97
				pending[v] = 4;                                        // done this round (flag drops after the barrier)
				if (nodeCol != old) { atomicAdd(hist + old, ~0ull); atomicAdd(hist + nodeCol, 1ull); atomicAdd(&cnt->changed, 1u); }
			}
		}
		__syncthreads();
		if (tc_smem[2]) {                                         // hub rows that are ready: one after the other, all threads
			for (uint32_t i = 0; i < n; ++i) {
				const uint32_t v = flist[i];
				if (pending[v] != 3) continue;                        // (uniform: every thread reads the same flag)
				const uint32_t old = colors[v];
				bool inexact;
				const uint32_t nodeCol = tc_pick_cta<ColT>(neighs, rowptr[v], rowptr[v + 1], colors, old, nCol, order, bm, tc_smem, &inexact);
				if (threadIdx.x == 0) {
					if (inexact) cnt->inexact = 1u;
					colors[v] = (ColT)nodeCol;
					pending[v] = 4;
					if (nodeCol != old) { atomicAdd(hist + old, ~0ull); atomicAdd(hist + nodeCol, 1ull); atomicAdd(&cnt->changed, 1u); }
				}
				__syncthreads();
			}
		}
		__syncthreads();                                          // colour writes of this round before the flags drop
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { const uint32_t v = flist[i]; if (pending[v] == 4) pending[v] = 0; }
		__threadfence_block();
		__syncthreads();
		if (tc_smem[1] == 0u) break;
		__syncthreads();
	}
}

template <typename ColT>
__global__ void tc_recount_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, const ColT * __restrict__ colors,
                                  const uint32_t * __restrict__ list, uint32_t listCount, uint8_t * pending, uint32_t * nextList, TailcutCounters * cnt) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	if (lane == 0) pending[v] = 0;
	const uint32_t c = colors[v];
	uint32_t same = 0; bool flag = false;
	for (uint32_t e = rowptr[v] + lane; e < rowptr[v + 1]; e += 32u) {
		const uint32_t u = neighs[e];
		const bool eq = (uint32_t)colors[u] == c;
		same += eq;
		flag = flag || (eq && u > v);
	}
	same = __reduce_add_sync(0xffffffffu, same);
	flag = __any_sync(0xffffffffu, flag);
	if (same && lane == 0) {
		atomicAdd(&cnt->directed, (unsigned long long)same);
		atomicAdd(&cnt->viol, 1ull);
		nextList[atomicAdd(&cnt->nextCount, 1u)] = v;
		if (flag) atomicAdd(&cnt->nextFlagged, 1u);
	}
}

// distributed repair (multi-GPU): flags of the vertices any rank is going to visit; colours chosen by another rank
__global__ void tc_mark_kernel(const uint32_t * __restrict__ ids, uint32_t count, uint8_t * pending) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) pending[ids[i]] = 1;
}
template <typename ColT>
__global__ void tc_remote_kernel(const uint32_t * __restrict__ ids, const uint32_t * __restrict__ cols, uint32_t count, ColT * colors, uint8_t * pending,
                                 unsigned long long * hist) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const uint32_t v = ids[i], c = cols[i], old = colors[v];
	colors[v] = (ColT)c;
	pending[v] = 0;
	if (c != old) { atomicAdd(hist + old, ~0ull); atomicAdd(hist + c, 1ull); }
}
__global__ void tc_commit_global_kernel(DevState * st, unsigned long long directed, unsigned long long viol, uint32_t listCount) {
	st->lastDirected = directed; st->lastViol = viol; st->countsSweep = st->sweep;
	st->convergedAt = -1;
	st->violListSweep = st->sweep; st->violListCount = listCount;
}

// the repaired colouring's counters become the chain's (no full recount needed); the surviving violators are its list
__global__ void tc_commit_kernel(DevState * st, const TailcutCounters * cnt) {
	st->lastDirected = cnt->directed; st->lastViol = cnt->viol; st->countsSweep = st->sweep;
	st->convergedAt = -1;                                         // the next sweep (if any) re-evaluates the threshold
	st->violListSweep = st->sweep; st->violListCount = cnt->nextCount;
}

} // namespace mcmcb200
