// Host side of the source-blocked sweep: builds the static layout once per graph (all on the device, CUB radix sort
// + scans) and launches the two passes.
#pragma once
#include <cstdlib>
#include <vector>

#include "blocked_sweep.cuh"

namespace mcmcb200 {

inline void free_blocked_layout(BlockedLayout & L) {
	cudaFree(L.srcLocal); cudaFree(L.ecol); cudaFree(L.gidx); cudaFree(L.gidxS); cudaFree(L.order); cudaFree(L.slotInfo); cudaFree(L.sliceOff);
	cudaFree(L.granDst); cudaFree(L.tileBase); cudaFree(L.items); cudaFree(L.sync); cudaFree(L.tilePart); cudaFree(L.dbgTimes);
	L = BlockedLayout{};
}

#define BLK_CU(call) do { err = (call); if (err != cudaSuccess) goto done; } while (0)

// Returns cudaSuccess with L.valid == false when the layout does not apply (hub rows larger than a tile, > 2^31 edges...).
#ifndef MCMCB200_BUILD_BY_SORT
#define MCMCB200_BUILD_BY_SORT 0     /* 1: always the sort-based construction (the path of graphs with more than 4096 source chunks) */
#endif
inline cudaError_t build_blocked_layout(BlockedLayout & L, const uint32_t * d_rowptr, const uint32_t * d_neighs, uint32_t nLocal,
                                        uint64_t nnzLocal, uint32_t nGlobal, int colBytes, uint32_t stageCapBytes, uint32_t itemEntries, uint32_t roundV,
                                        cudaStream_t stream, uint64_t * launches) {
	cudaError_t err = cudaSuccess;
	L = BlockedLayout{};
	if (nnzLocal == 0 || nnzLocal >= (1ull << 31) || nLocal == 0) return cudaSuccess;
	const uint32_t nnz = (uint32_t)nnzLocal;
	const uint32_t P = (nGlobal + kChunkV - 1) / kChunkV;
	const uint32_t stageCap = (stageCapBytes / (uint32_t)colBytes) & ~15u;
	if (stageCap < 1024 || stageCap > 65504) return cudaSuccess;    // stage positions (and the dummy slot at stageCap) are 16-bit

	uint32_t * d_tmp = nullptr;          // [2]: scratch scalars
	uint32_t * d_tileE = nullptr;
	uint16_t * d_keys[2] = {nullptr, nullptr};
	uint32_t * d_vals[2] = {nullptr, nullptr};
	void * d_cub = nullptr; size_t cubBytes = 0;
	uint32_t * d_cnt = nullptr, * d_us = nullptr, * d_plen = nullptr, * d_gs = nullptr, * d_plenT = nullptr, * d_scanT = nullptr, * d_bs = nullptr;
	uint32_t TV = 0, numTiles = 0, h2[2] = {0, 0};
	size_t cells = 0;
	unsigned long long * d_k64[2] = {nullptr, nullptr};
	uint32_t * d_v32[2] = {nullptr, nullptr};
	uint32_t * d_words = nullptr, * d_runStart = nullptr, * d_stageOff = nullptr;
	uint32_t numSlices = 0, sellTotal = 0;
	bool tileLocal = false; size_t tlSmem = 0;

	BLK_CU(cudaMalloc(&d_tmp, 2 * sizeof(uint32_t)));
	// ---- tile size: a multiple of 128 vertices whose worst tile (edges + run padding) fits the stage.  Pass B walks a tile in
	//      rounds of one vertex per thread, so a tile of k * threads vertices leaves no lane idle in the last round (measured on
	//      config 3: 1536 = 4 x 384 vertices 3.81 ms, 1792 vertices 4.35 ms): take the largest multiple of `roundV` that fits, or,
	//      for small stages / graphs, the largest multiple of 128 ----
	{
		auto fits = [&](uint32_t tv, bool & ok) -> cudaError_t {
			const uint32_t nt = (nLocal + tv - 1) / tv;
			cudaError_t e = cudaMemsetAsync(d_tmp, 0, sizeof(uint32_t), stream);
			if (e != cudaSuccess) return e;
			blk_max_tile_edges_kernel<<<(nt + 255) / 256, 256, 0, stream>>>(d_rowptr, nLocal, tv, nt, d_tmp); (*launches)++;
			e = cudaMemcpyAsync(h2, d_tmp, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
			if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
			const uint64_t worst = (uint64_t)h2[0] + 3ull * std::min<uint64_t>(P, h2[0]) + 16ull;   // + run padding + start misalignment
			ok = worst <= stageCap;
			return e;
		};
		const uint32_t maxTv = std::min<uint32_t>(8192u, ((nLocal + 127u) / 128u) * 128u);      // (13-bit local vertex ids in the slot table)
		bool ok = false;
		uint32_t lo = 0, hi = maxTv / 128u;                        // binary search on the multiple of 128 (worst-tile size grows with tv)
		while (lo < hi) {
			const uint32_t mid = (lo + hi + 1) / 2;
			BLK_CU(fits(mid * 128u, ok));
			if (ok) lo = mid; else hi = mid - 1;
		}
		uint32_t tv = lo * 128u;
		if (tv >= roundV && tv % roundV != 0) {
			const uint32_t tr = (tv / roundV) * roundV;
			BLK_CU(fits(tr, ok));
			if (ok) tv = tr;
		}
		if (tv) { BLK_CU(fits(tv, ok)); if (!ok) tv = 0; }        // (re-check: the search assumes monotonicity)
		if (tv) { TV = tv; numTiles = (nLocal + tv - 1) / tv; }
	}
	if (TV == 0) goto done;                                     // a row (or 256 of them) exceeds the stage: direct kernel only
	cells = (size_t)P * numTiles;
	if (cells >= (1ull << 31)) goto done;

	// ---- bin the directed edges by source chunk (stable radix sort of (chunk, edge id)) ----
	BLK_CU(cudaMalloc(&d_tileE, sizeof(uint32_t) * ((size_t)numTiles + 1)));
	blk_tile_edge_starts_kernel<<<(numTiles + 1 + 255) / 256, 256, 0, stream>>>(d_rowptr, nLocal, TV, numTiles, d_tileE); (*launches)++;
	// tile-local construction (blocked_sweep.cuh: blk_tile_hist_kernel / blk_tile_rank_kernel) whenever the per-warp counters fit
	tileLocal = P <= kTileLocalMaxP && !MCMCB200_BUILD_BY_SORT;
	tlSmem = sizeof(uint32_t) * (size_t)P * kTileLocalWarps;
	if (tileLocal && tlSmem > 48 * 1024) BLK_CU(cudaFuncSetAttribute(blk_tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tlSmem));
	if (tileLocal) BLK_CU(cudaFuncSetAttribute(blk_tile_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(blk_tile_rank_smem(P, stageCap) + 16)));
	if (!tileLocal) {
	for (int i = 0; i < 2; ++i) {
		BLK_CU(cudaMalloc(&d_keys[i], sizeof(uint16_t) * (size_t)nnz));
		BLK_CU(cudaMalloc(&d_vals[i], sizeof(uint32_t) * (size_t)nnz));
	}
	blk_edge_keys_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(d_neighs, nnz, d_keys[0], d_vals[0]); (*launches)++;
	{
		cub::DoubleBuffer<uint16_t> kb(d_keys[0], d_keys[1]);
		cub::DoubleBuffer<uint32_t> vb(d_vals[0], d_vals[1]);
		int endBit = 1;
		while ((1u << endBit) < P) endBit++;
		BLK_CU(cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, kb, vb, (int)nnz, 0, endBit, stream));
		BLK_CU(cudaMalloc(&d_cub, cubBytes));
		BLK_CU(cub::DeviceRadixSort::SortPairs(d_cub, cubBytes, kb, vb, (int)nnz, 0, endBit, stream)); (*launches) += 4;
		if (kb.Current() != d_keys[0]) { std::swap(d_keys[0], d_keys[1]); }
		if (vb.Current() != d_vals[0]) { std::swap(d_vals[0], d_vals[1]); }
		cudaFree(d_cub); d_cub = nullptr;
	}
	cudaFree(d_keys[1]); d_keys[1] = nullptr; cudaFree(d_vals[1]); d_vals[1] = nullptr;
	}

	// ---- run lengths per (bucket, tile), padded to 4; three exclusive scans ----
	BLK_CU(cudaMalloc(&d_cnt, sizeof(uint32_t) * cells));
	BLK_CU(cudaMalloc(&d_us, sizeof(uint32_t) * cells));
	BLK_CU(cudaMalloc(&d_plen, sizeof(uint32_t) * cells));
	BLK_CU(cudaMalloc(&d_gs, sizeof(uint32_t) * cells));
	BLK_CU(cudaMalloc(&d_plenT, sizeof(uint32_t) * cells));
	BLK_CU(cudaMalloc(&d_scanT, sizeof(uint32_t) * cells));
	if (tileLocal) {
		blk_tile_hist_kernel<<<(numTiles + kTileLocalWarps - 1) / kTileLocalWarps, kTileLocalWarps * 32, tlSmem, stream>>>(d_neighs, d_tileE, numTiles, P, d_cnt); (*launches)++;
	} else {
		BLK_CU(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * cells, stream));
		blk_run_count_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(d_keys[0], d_vals[0], nnz, d_tileE, numTiles, d_cnt); (*launches)++;
	}
	blk_pad_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(d_cnt, P, numTiles, d_plen, d_plenT); (*launches)++;
	{
		size_t need = 0;
		BLK_CU(cub::DeviceScan::ExclusiveSum(nullptr, need, d_cnt, d_us, (int)cells, stream));
		BLK_CU(cudaMalloc(&d_cub, need));
		if (!tileLocal) BLK_CU(cub::DeviceScan::ExclusiveSum(d_cub, need, d_cnt, d_us, (int)cells, stream));
		BLK_CU(cub::DeviceScan::ExclusiveSum(d_cub, need, d_plen, d_gs, (int)cells, stream));
		BLK_CU(cub::DeviceScan::ExclusiveSum(d_cub, need, d_plenT, d_scanT, (int)cells, stream)); (*launches) += 6;
	}
	// total padded entries = gs[last] + plen[last]; must fit 32-bit offsets
	{
		uint32_t lastG = 0, lastP = 0;
		BLK_CU(cudaMemcpyAsync(&lastG, d_gs + cells - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
		BLK_CU(cudaMemcpyAsync(&lastP, d_plen + cells - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
		BLK_CU(cudaStreamSynchronize(stream));
		const uint64_t total = (uint64_t)lastG + lastP;
		if (total < nnz || total >= 0xfffffff0ull || total > (uint64_t)nnz + 3ull * cells) goto done;   // wrapped: too many runs
		L.totalPadded = (uint32_t)total;
	}
	// ---- static tables and entry arrays ----
	BLK_CU(cudaMalloc(&d_runStart, sizeof(uint32_t) * cells));
	BLK_CU(cudaMalloc(&d_stageOff, sizeof(uint32_t) * ((size_t)numTiles * (P + 1))));
	BLK_CU(cudaMemsetAsync(d_tmp, 0, 2 * sizeof(uint32_t), stream));
	blk_tables_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(d_gs, d_scanT, d_plenT, P, numTiles, d_runStart, d_stageOff, d_tmp); (*launches)++;
	BLK_CU(cudaMalloc(&L.granDst, sizeof(uint32_t) * ((size_t)(L.totalPadded >> 2) + 16)));
	BLK_CU(cudaMalloc(&L.tileBase, sizeof(uint32_t) * ((size_t)numTiles + 1)));
	blk_gran_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(d_runStart, d_scanT, d_plenT, cells, L.granDst); (*launches)++;
	blk_tile_base_kernel<<<(numTiles + 1 + 255) / 256, 256, 0, stream>>>(d_scanT, P, numTiles, L.totalPadded, L.tileBase); (*launches)++;
	BLK_CU(cudaMalloc(&L.srcLocal, sizeof(uint16_t) * ((size_t)L.totalPadded + 16)));
	BLK_CU(cudaMemsetAsync(L.srcLocal, 0, sizeof(uint16_t) * ((size_t)L.totalPadded + 16), stream));
	BLK_CU(cudaMalloc(&L.gidx, sizeof(uint16_t) * ((size_t)nnz + 16)));
	BLK_CU(cudaMemsetAsync(L.gidx, 0, sizeof(uint16_t) * ((size_t)nnz + 16), stream));
	if (tileLocal) {
		blk_tile_rank_kernel<<<numTiles, 32, blk_tile_rank_smem(P, stageCap) + 16, stream>>>(
			d_neighs, d_tileE, numTiles, P, stageCap, d_runStart, d_stageOff, d_scanT, (uint32_t)(16 / colBytes) - 1u, L.srcLocal, L.gidx); (*launches)++;
	} else {
		blk_fill_entries_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(d_keys[0], d_vals[0], nnz, d_neighs, d_tileE, numTiles, P, d_us, d_gs,
		                                                                d_stageOff, d_scanT, (uint32_t)(16 / colBytes) - 1u, L.srcLocal, L.gidx); (*launches)++;
	}
	BLK_CU(cudaMalloc(&L.ecol, (size_t)colBytes * ((size_t)L.totalPadded + 16)));
	BLK_CU(cudaMemsetAsync(L.ecol, 0, (size_t)colBytes * ((size_t)L.totalPadded + 16), stream));
	// ---- SELL-32-sigma copy of gidx for the light rows: per tile, vertices by descending degree; 32-slot slices interleaved ----
	// (the sort buffers of the edge binning are released first)
	cudaFree(d_keys[0]); d_keys[0] = nullptr; cudaFree(d_vals[0]); d_vals[0] = nullptr;
	cudaFree(d_cnt); d_cnt = nullptr; cudaFree(d_us); d_us = nullptr; cudaFree(d_plen); d_plen = nullptr; cudaFree(d_plenT); d_plenT = nullptr;
	for (int i = 0; i < 2; ++i) {
		BLK_CU(cudaMalloc(&d_k64[i], sizeof(unsigned long long) * (size_t)nLocal));
		BLK_CU(cudaMalloc(&d_v32[i], sizeof(uint32_t) * (size_t)nLocal));
	}
	blk_sell_keys_kernel<<<(nLocal + 255) / 256, 256, 0, stream>>>(d_rowptr, nLocal, TV, d_k64[0], d_v32[0]); (*launches)++;
	{
		cub::DoubleBuffer<unsigned long long> kb(d_k64[0], d_k64[1]);
		cub::DoubleBuffer<uint32_t> vb(d_v32[0], d_v32[1]);
		int endBit = 33;
		while (endBit < 64 && (1ull << (endBit - 32)) < (unsigned long long)numTiles) endBit++;
		size_t need = 0;
		cudaFree(d_cub); d_cub = nullptr;
		BLK_CU(cub::DeviceRadixSort::SortPairs(nullptr, need, kb, vb, (int)nLocal, 0, endBit, stream));
		BLK_CU(cudaMalloc(&d_cub, need));
		BLK_CU(cub::DeviceRadixSort::SortPairs(d_cub, need, kb, vb, (int)nLocal, 0, endBit, stream)); (*launches) += 6;
		if (vb.Current() != d_v32[0]) std::swap(d_v32[0], d_v32[1]);
	}
	numSlices = (uint32_t)(((size_t)numTiles * TV) / 32);
	BLK_CU(cudaMalloc(&L.order, sizeof(uint16_t) * (size_t)numTiles * TV));
	blk_sell_order_kernel<<<(unsigned)(((size_t)numTiles * TV + 255) / 256), 256, 0, stream>>>(d_v32[0], nLocal, TV, numTiles, L.order); (*launches)++;
	BLK_CU(cudaMalloc(&d_words, sizeof(uint32_t) * ((size_t)numSlices + 1)));
	BLK_CU(cudaMemsetAsync(d_words, 0, sizeof(uint32_t) * ((size_t)numSlices + 1), stream));
	blk_sell_width_kernel<<<(numSlices * 32 + 255) / 256, 256, 0, stream>>>(d_rowptr, L.order, TV, numSlices, d_words); (*launches)++;
	BLK_CU(cudaMalloc(&L.sliceOff, sizeof(uint32_t) * ((size_t)numSlices + 8)));    // (+ slack: pass B copies spt + 4 entries per tile)
	BLK_CU(cudaMemsetAsync(L.sliceOff, 0, sizeof(uint32_t) * ((size_t)numSlices + 8), stream));
	{
		size_t need = 0;
		cudaFree(d_cub); d_cub = nullptr;
		BLK_CU(cub::DeviceScan::ExclusiveSum(nullptr, need, d_words, L.sliceOff, (int)numSlices + 1, stream));
		BLK_CU(cudaMalloc(&d_cub, need));
		BLK_CU(cub::DeviceScan::ExclusiveSum(d_cub, need, d_words, L.sliceOff, (int)numSlices + 1, stream)); (*launches) += 2;
	}
	BLK_CU(cudaMemcpyAsync(&sellTotal, L.sliceOff + numSlices, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	BLK_CU(cudaStreamSynchronize(stream));
	if ((uint64_t)sellTotal > (uint64_t)nnz + 32ull * numSlices + 64ull * nLocal) goto done;              // 32-bit scan wrapped
	BLK_CU(cudaMalloc(&L.gidxS, sizeof(uint2) * ((size_t)sellTotal + 64)));
	BLK_CU(cudaMemsetAsync(L.gidxS, 0, sizeof(uint2) * ((size_t)sellTotal + 64), stream));
	blk_sell_fill_kernel<<<(numSlices * 32 + 255) / 256, 256, 0, stream>>>(d_rowptr, L.order, L.gidx, TV, numSlices, L.sliceOff, stageCap, L.gidxS); (*launches)++;
	BLK_CU(cudaMalloc(&L.slotInfo, sizeof(uint16_t) * (size_t)numTiles * TV));
	blk_sell_slotinfo_kernel<<<(unsigned)(((size_t)numTiles * TV + 255) / 256), 256, 0, stream>>>(d_rowptr, L.order, TV, numTiles, L.slotInfo); (*launches)++;
	// ---- pass-A work items: (part, bucket) -> entry range; parts are stretches of tiles so that "part p finished" means the
	//      stage images of its tiles are complete (pass B follows pass A part by part) ----
	BLK_CU(cudaMemcpyAsync(h2, d_tmp, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	BLK_CU(cudaStreamSynchronize(stream));
	if (h2[0] + 16u > stageCap) goto done;                             // (cannot happen given the TV choice; keeps the kernel's bound honest)
	{
		// about itemEntries entries per item in the body: K = tiles per body part.  The first and the last K tiles are cut
		// into K/4, K/4, K/2 (and mirrored): pass B can start after 1/4 of a part and has only 1/4 of a part left when pass A ends.
		const uint64_t perBucket = ((uint64_t)L.totalPadded + P - 1) / P;
		const uint32_t np0 = (uint32_t)std::min<uint64_t>(32, std::max<uint64_t>(1, (perBucket + itemEntries - 1) / itemEntries));
		const uint32_t K = std::max<uint32_t>(1u, (numTiles + np0 - 1) / np0);
		std::vector<uint32_t> ps;
		ps.push_back(0);
		auto cut = [&](uint32_t len) { const uint32_t nx = std::min<uint64_t>(numTiles, (uint64_t)ps.back() + std::max<uint32_t>(1u, len)); if (nx > ps.back()) ps.push_back(nx); };
		if (numTiles >= 4u * K && K >= 8u) {         // (the same cuts on partitions of only 2 parts were measured slower: 0.62 vs 0.55 ms on an 8-GPU rank)
			cut(K / 4); cut(K / 4); cut(K - 2 * (K / 4));
			while (numTiles - ps.back() > 2u * K) cut(K);
			const uint32_t rest = numTiles - ps.back();               // in (K, 2K]
			const uint32_t tail = std::min(K, rest);
			if (rest > tail) cut(rest - tail);
			cut(tail - 2 * (tail / 4)); cut(tail / 4); cut(tail / 4);
			if (ps.back() < numTiles) ps.push_back(numTiles);
		} else {
			// (few parts: e.g. one rank of an 8-GPU run.  A short first part -- K/4, K/8, K/16 tiles -- so that pass B need not wait for
			//  half of pass A was measured slower on such a partition: 0.563 / 0.570 / 0.575 vs 0.540 ms, its small items pay the chunk reload)
			while (ps.back() < numTiles) cut(K);
		}
		L.numParts = (uint32_t)ps.size() - 1;
		if (L.numParts > 255u) goto done;
		L.numItems = L.numParts * P;
		uint32_t * d_ps = nullptr;
		BLK_CU(cudaMalloc(&d_ps, sizeof(uint32_t) * ps.size()));
		err = cudaMemcpyAsync(d_ps, ps.data(), sizeof(uint32_t) * ps.size(), cudaMemcpyHostToDevice, stream);
		if (err == cudaSuccess) err = cudaMalloc(&L.items, sizeof(uint32_t) * 3 * (size_t)L.numItems);
		if (err == cudaSuccess) err = cudaMalloc(&L.tilePart, numTiles);
		if (err == cudaSuccess) {
			blk_items_kernel<<<(L.numItems + 255) / 256, 256, 0, stream>>>(d_gs, P, numTiles, d_ps, L.numParts, L.totalPadded, L.items, L.tilePart); (*launches)++;
			err = cudaStreamSynchronize(stream);
		}
		cudaFree(d_ps);
		if (err != cudaSuccess) goto done;
		BLK_CU(cudaMalloc(&L.dbgTimes, 8 * sizeof(unsigned long long)));
		BLK_CU(cudaMalloc(&L.sync, sizeof(uint32_t) * (2 + (size_t)L.numParts)));
		BLK_CU(cudaMemsetAsync(L.sync, 0, sizeof(uint32_t) * (2 + (size_t)L.numParts), stream));
		BLK_CU(cudaStreamSynchronize(stream));
	}
	cudaFree(L.order); L.order = nullptr;                        // construction only
	L.P = P; L.TV = TV; L.numTiles = numTiles; L.stageCap = stageCap;
	L.bytes = sizeof(uint16_t) * ((size_t)L.totalPadded + 16) + (size_t)colBytes * ((size_t)L.totalPadded + 16)       // srcLocal, ecol
	        + sizeof(uint32_t) * ((size_t)(L.totalPadded >> 2) + 16) + sizeof(uint16_t) * ((size_t)nnz + 16)            // granDst, gidx
	        + sizeof(uint2) * ((size_t)sellTotal + 64) + sizeof(uint16_t) * (size_t)numTiles * TV                       // gidxS, slotInfo
	        + sizeof(uint32_t) * ((size_t)numSlices + 8) + sizeof(uint32_t) * ((size_t)numTiles + 1)                    // sliceOff, tileBase
	        + sizeof(uint32_t) * 3 * (size_t)L.numItems + numTiles;                                                      // items, tilePart
	L.valid = true;
done:
	cudaFree(d_tmp); cudaFree(d_tileE); cudaFree(d_keys[0]); cudaFree(d_keys[1]); cudaFree(d_vals[0]); cudaFree(d_vals[1]); cudaFree(d_cub);
	cudaFree(d_cnt); cudaFree(d_us); cudaFree(d_plen); cudaFree(d_gs); cudaFree(d_plenT); cudaFree(d_scanT); cudaFree(d_bs);
	cudaFree(d_k64[0]); cudaFree(d_k64[1]); cudaFree(d_v32[0]); cudaFree(d_v32[1]); cudaFree(d_words); cudaFree(d_runStart); cudaFree(d_stageOff);
	if (err != cudaSuccess || !L.valid) { cudaError_t keep = err; free_blocked_layout(L); err = keep; }
	if (err == cudaErrorMemoryAllocation) { cudaGetLastError(); err = cudaSuccess; }   // not enough room for the layout: direct kernel
	return err;
}
#undef BLK_CU

inline BlockedArgs make_blocked_args(const BlockedLayout & L) {
	BlockedArgs b{};
	b.P = L.P; b.TV = L.TV; b.numTiles = L.numTiles; b.stageCap = L.stageCap;
#if MCMCB200_BOUNDS_CHECK
	b.totalPadded = L.totalPadded;
#endif
	b.srcLocal = L.srcLocal; b.ecol = L.ecol; b.gidx = L.gidx; b.gidxS = L.gidxS; b.slotInfo = L.slotInfo; b.sliceOff = L.sliceOff;
	b.granDst = L.granDst; b.tileBase = L.tileBase;
	b.items = L.items; b.numItems = L.numItems; b.numParts = L.numParts; b.nbuf = L.nbuf; b.tilePart = L.tilePart; b.sync = L.sync; b.dbgTimes = L.dbgTimes;
	return b;
}

} // namespace mcmcb200
