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

__global__ void tailcut_ready_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs,
                                     const uint8_t * __restrict__ pending, const uint32_t * __restrict__ list, uint32_t listCount,
                                     uint8_t * ready) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	if (!pending[v]) { ready[i] = 0; return; }
	bool ok = true;
	for (uint32_t e = rowptr[v]; e < rowptr[v + 1] && ok; ++e) {
		const uint32_t u = neighs[e];
		ok = !(u < v && pending[u]);
	}
	ready[i] = ok ? 1 : 0;
}

template <typename ColT>
__global__ void tailcut_apply_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, uint32_t nCol,
                                     ColT * colors, uint8_t * pending, const uint32_t * __restrict__ list, uint32_t listCount,
                                     const uint8_t * __restrict__ ready, const uint32_t * __restrict__ order,
                                     unsigned long long * hist, uint32_t * remaining) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	if (!pending[v]) return;
	if (!ready[i]) { atomicAdd(remaining, 1u); return; }
	const uint32_t e0 = rowptr[v], e1 = rowptr[v + 1];
	auto occupied = [&](uint32_t c) {
		for (uint32_t e = e0; e < e1; ++e) if ((uint32_t)colors[neighs[e]] == c) return true;
		return false;
	};
	const uint32_t old = colors[v];
	uint32_t nodeCol = old, j = 0;
	while (occupied(nodeCol) && j < nCol) { nodeCol = order[j]; j++; }      // _utils.cu:91-95
	colors[v] = (ColT)nodeCol;                                               // :97
	pending[v] = 0;
	if (nodeCol != old) { atomicAdd(hist + old, ~0ull); atomicAdd(hist + nodeCol, 1ull); }
}

// One reference pass.  Returns a cudaError_t as int; *flaggedOut = number of vertices the pass visited.
inline int launch_tailcut_pass(cudaStream_t stream, int colBytes, const uint32_t * rowptr, const uint32_t * neighs, uint32_t n,
                               uint32_t nCol, void * colors, unsigned long long * hist, const uint32_t * d_order,
                               uint8_t * d_pending, uint8_t * d_ready, uint32_t * d_list, uint32_t * d_counters /* [2] */,
                               uint32_t * flaggedOut, uint64_t * launches) {
	cudaError_t e;
	if ((e = cudaMemsetAsync(d_counters, 0, 2 * sizeof(uint32_t), stream)) != cudaSuccess) return (int)e;
	const uint32_t blocks = (n + 255) / 256;
	if (colBytes == 1) tailcut_flag_kernel<uint8_t><<<blocks, 256, 0, stream>>>(rowptr, neighs, n, (const uint8_t *)colors, d_pending, d_list, d_counters);
	else tailcut_flag_kernel<uint16_t><<<blocks, 256, 0, stream>>>(rowptr, neighs, n, (const uint16_t *)colors, d_pending, d_list, d_counters);
	(*launches)++;
	uint32_t listCount = 0;
	if ((e = cudaMemcpyAsync(&listCount, d_counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
	if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
	*flaggedOut = listCount;
	if (listCount == 0) return 0;
	const uint32_t lb = (listCount + 255) / 256;
	for (uint32_t guard = 0; guard <= listCount; ++guard) {
		if ((e = cudaMemsetAsync(d_counters + 1, 0, sizeof(uint32_t), stream)) != cudaSuccess) return (int)e;
		tailcut_ready_kernel<<<lb, 256, 0, stream>>>(rowptr, neighs, d_pending, d_list, listCount, d_ready);
		if (colBytes == 1) tailcut_apply_kernel<uint8_t><<<lb, 256, 0, stream>>>(rowptr, neighs, nCol, (uint8_t *)colors, d_pending, d_list, listCount, d_ready, d_order, hist, d_counters + 1);
		else tailcut_apply_kernel<uint16_t><<<lb, 256, 0, stream>>>(rowptr, neighs, nCol, (uint16_t *)colors, d_pending, d_list, listCount, d_ready, d_order, hist, d_counters + 1);
		(*launches) += 2;
		uint32_t remaining = 0;
		if ((e = cudaMemcpyAsync(&remaining, d_counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
		if (remaining == 0) break;
	}
	return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// List-driven tail cutting (the production path).  The chain stops sweeping once at most z = max(50, n/2000) vertices violate
// (coloringMCMC_main.cu:150-170); the sweep that established this also emitted those vertices (SweepArgs::violList), so the
// repair touches z rows instead of the whole CSR and needs no host round trip per round:
//   tc_filter_kernel   the reference's flags (conflictCounter, _utils.cu:115: a same-coloured neighbour with a LARGER id)
//                      for the listed vertices only -- an unlisted vertex has no same-coloured neighbour at all;
//   tc_rounds_kernel   ONE CTA runs all rounds of the dependency-ordered repair above (ready / apply, __syncthreads between);
//   tc_recount_kernel  conflicts and violations of the repaired colouring, again from the listed rows only (a vertex that was
//                      not violating can only become so if a repaired neighbour found every colour taken -- flagged `inexact`,
//                      the caller then recounts with a full pass), and the violators that are left (the next pass's list).
// ---------------------------------------------------------------------------------------------------------------
struct TailcutCounters {
	uint32_t flagged;            // entries of flist (this pass)
	uint32_t inexact;            // a repaired vertex found all nCol colours occupied
	uint32_t nextCount;          // violating vertices after the pass (entries of the next list)
	uint32_t nextFlagged;        // of those, vertices the reference would flag again
	unsigned long long directed; // sum over vertices of same-coloured neighbours (= 2 x conflicting edges)
	unsigned long long viol;     // violating vertices
};

template <typename ColT>
__global__ void tc_filter_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, const ColT * __restrict__ colors,
                                 const uint32_t * __restrict__ list, uint32_t listCount, uint8_t * pending, uint32_t * flist, TailcutCounters * cnt) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	const uint32_t c = colors[v];
	bool flag = false;
	for (uint32_t e = rowptr[v]; e < rowptr[v + 1] && !flag; ++e) {
		const uint32_t u = neighs[e];
		flag = (u > v) && ((uint32_t)colors[u] == c);
	}
	if (flag) { pending[v] = 1; flist[atomicAdd(&cnt->flagged, 1u)] = v; }
}

template <typename ColT>
__global__ void __launch_bounds__(1024)
tc_rounds_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, uint32_t nCol, ColT * colors, uint8_t * pending,
                 const uint32_t * __restrict__ flist, TailcutCounters * cnt, const uint32_t * __restrict__ order, unsigned long long * hist) {
	__shared__ uint32_t s_left;
	const uint32_t n = cnt->flagged;
	for (;;) {
		// ready: no still-pending flagged neighbour with a smaller id (the sequential loop would have visited it first)
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
			const uint32_t v = flist[i];
			if (pending[v] != 1) continue;
			bool ok = true;
			for (uint32_t e = rowptr[v]; e < rowptr[v + 1] && ok; ++e) { const uint32_t u = neighs[e]; ok = !(u < v && pending[u]); }
			if (ok) pending[v] = 3;                                // pending AND ready (still non-zero for its larger neighbours)
		}
		if (threadIdx.x == 0) s_left = 0u;
		__syncthreads();
		// apply: ready vertices are pairwise non-adjacent, each sees exactly the colours the sequential loop would
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
			const uint32_t v = flist[i];
			const uint32_t p = pending[v];
			if (p == 1) { atomicAdd(&s_left, 1u); continue; }
			if (p != 3) continue;
			const uint32_t e0 = rowptr[v], e1 = rowptr[v + 1];
			auto occupied = [&](uint32_t c) {
				for (uint32_t e = e0; e < e1; ++e) if ((uint32_t)colors[neighs[e]] == c) return true;
				return false;
			};
			const uint32_t old = colors[v];
			uint32_t nodeCol = old, j = 0;
			while (occupied(nodeCol) && j < nCol) { nodeCol = order[j]; j++; }      // _utils.cu:91-95
			if (j == nCol && occupied(nodeCol)) cnt->inexact = 1u;                   // every colour taken: may disturb an unlisted neighbour
			colors[v] = (ColT)nodeCol;                                               // :97
			if (nodeCol != old) { atomicAdd(hist + old, ~0ull); atomicAdd(hist + nodeCol, 1ull); }
		}
		__syncthreads();                                          // colour writes of this round before the flags drop
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { const uint32_t v = flist[i]; if (pending[v] == 3) pending[v] = 0; }
		__threadfence_block();
		__syncthreads();
		if (s_left == 0u) break;
		__syncthreads();
	}
}

template <typename ColT>
__global__ void tc_recount_kernel(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ neighs, const ColT * __restrict__ colors,
                                  const uint32_t * __restrict__ list, uint32_t listCount, uint8_t * pending, uint32_t * nextList, TailcutCounters * cnt) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= listCount) return;
	const uint32_t v = list[i];
	pending[v] = 0;
	const uint32_t c = colors[v];
	uint32_t same = 0; bool flag = false;
	for (uint32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
		const uint32_t u = neighs[e];
		const bool eq = (uint32_t)colors[u] == c;
		same += eq;
		flag = flag || (eq && u > v);
	}
	if (same) {
		atomicAdd(&cnt->directed, (unsigned long long)same);
		atomicAdd(&cnt->viol, 1ull);
		nextList[atomicAdd(&cnt->nextCount, 1u)] = v;
		if (flag) atomicAdd(&cnt->nextFlagged, 1u);
	}
}

// the repaired colouring's counters become the chain's (no full recount needed); the surviving violators are its list
__global__ void tc_commit_kernel(DevState * st, const TailcutCounters * cnt) {
	st->lastDirected = cnt->directed; st->lastViol = cnt->viol; st->countsSweep = st->sweep;
	st->convergedAt = -1;                                         // the next sweep (if any) re-evaluates the threshold
	st->violListSweep = st->sweep; st->violListCount = cnt->nextCount;
}

} // namespace mcmcb200
