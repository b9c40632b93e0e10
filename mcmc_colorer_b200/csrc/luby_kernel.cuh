// Luby-style MIS colourer, cross-check only (north star: "the Luby GPU colorer is re-implemented only as a
// cross-check").  Restates the host-driven loop of the reference's ColoringLuby::run (graph_coloring/coloringLuby.cu:364-501)
// and its kernels (:222-341): for every new colour, a maximal independent set of the still uncoloured vertices is grown
// by rounds of { fair coin per candidate (:236-248), drop the lower-degree endpoint of every edge whose two endpoints
// were both chosen, both on a tie (:252-283), add the survivors and retire them and their neighbours (:289-316) }.
// Differences, all deliberate: the coin is stateless Philox (vertex, purpose 2, round) instead of curandState; the
// conflict rule reads a frozen snapshot of the choices (the reference reads flags other threads are clearing, a benign
// race by its own account), so the result is deterministic; no dynamic parallelism (run_fast needs device-side
// cudaDeviceSynchronize, gone since CUDA 12).  Colours are 1-based like the reference's; 0 = uncoloured.
#pragma once
#include "device_utils.cuh"

namespace mcmcb200 {

__global__ void luby_prune_kernel(uint32_t n, const uint32_t * colors, uint8_t * cands, uint8_t * is) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	cands[v] = colors[v] == 0u;                                   // prune_eligible, coloringLuby.cu:223-228
	is[v] = 0;
}

__global__ void luby_choose_kernel(uint32_t n, uint64_t seed, uint32_t round, const uint8_t * cands, uint8_t * chosen) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	const float u = draw_to_uniform(philox_draw(seed, round, v, 2u), true);     // (0,1] like curand_uniform
	chosen[v] = (u < 0.5f) ? cands[v] : 0;                       // set_initial_distr_k, :236-248
}

__global__ void luby_resolve_kernel(uint32_t n, const uint32_t * rowptr, const uint32_t * neighs, const uint8_t * chosen, uint8_t * keep) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	uint8_t k = chosen[v];
	if (k) {
		const uint32_t b = rowptr[v], dv = rowptr[v + 1] - b;
		for (uint32_t j = 0; j < dv && k; ++j) {                 // check_conflicts_k, :252-283
			const uint32_t u = neighs[b + j];
			if (chosen[u] && dv <= rowptr[u + 1] - rowptr[u]) k = 0;
		}
	}
	keep[v] = k;
}

__global__ void luby_update_kernel(uint32_t n, const uint32_t * rowptr, const uint32_t * neighs, const uint8_t * keep, uint8_t * cands,
                                   uint8_t * is, uint32_t * anyLeft) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	if (keep[v]) {                                               // update_eligible_k, :289-316
		is[v] = 1;
		cands[v] = 0;
		for (uint32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) cands[neighs[e]] = 0;
	}
}

__global__ void luby_left_kernel(uint32_t n, const uint8_t * cands, uint32_t * anyLeft) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v < n && cands[v]) *anyLeft = 1u;                        // check_finished_k, :320-328
}

__global__ void luby_color_kernel(uint32_t n, uint32_t color, const uint8_t * is, uint32_t * colors, uint32_t * anyUncolored) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	if (is[v]) colors[v] = color;                                // add_color_and_check_uncolored_k, :332-346
	if (colors[v] == 0u) *anyUncolored = 1u;
}

} // namespace mcmcb200
