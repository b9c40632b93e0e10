// libmcmcb200.so -- C ABI (include/mcmcb200.h) over the sm_100a sweep kernels.  No torch types, no CPU fallback.
#include "../../include/mcmcb200.h"
#include "sweep_kernel.cuh"
#include "tailcut_kernel.cuh"
#include "blocked_build.cuh"
#include "binned_sweep.cuh"
#include "wide_sweep.cuh"
#include "csr_build.cuh"
#include "luby_kernel.cuh"

#include <nvtx3/nvToolsExt.h>    // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx; no-ops without a tool attached

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

using namespace mcmcb200;

static_assert(sizeof(mcmcb200_params) == 72, "ABI: mcmcb200_params layout (mirrored by capi.Params)");
static_assert(sizeof(mcmcb200_status_t) == 40, "ABI: mcmcb200_status_t layout (mirrored by capi.Status)");

namespace {

thread_local char g_lastCudaError[512] = "";

int cuda_fail(cudaError_t e, const char * what, int line) {
	snprintf(g_lastCudaError, sizeof(g_lastCudaError), "%s failed at mcmcb200.cu:%d: %s (%s)", what, line,
	         cudaGetErrorName(e), cudaGetErrorString(e));
	return (e == cudaErrorMemoryAllocation) ? MCMCB200_ENOMEM
	     : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) ? MCMCB200_ENODEVICE
	     : MCMCB200_ECUDA;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call, __LINE__); } while (0)

constexpr uint32_t kColorPad = 65536;   // slack so equal-sized all-gather chunks always fit

} // namespace

struct mcmcb200_handle {
	mcmcb200_params p{};
	int device = 0;
	cudaStream_t stream = nullptr;
	uint32_t nGlobal = 0, vBegin = 0, vEnd = 0, nLocal = 0;
	uint64_t nnzLocal = 0;
	int W = 1, colBytes = 1;
	uint32_t * d_rowptr = nullptr;
	uint32_t * d_neighs = nullptr;
	bool ownsCsr = true;
	void * d_colors[2] = {nullptr, nullptr};
	void * d_colorsTmp = nullptr;        // conflicts_of scratch colouring
	uint16_t * d_taboo = nullptr;
	float * d_tape = nullptr;
	uint32_t tapeSweeps = 0, tapeBase = 0;
	DevState * d_state = nullptr;
	unsigned long long * d_scratch = nullptr, * d_hist[2] = {nullptr, nullptr}, * d_history = nullptr, * d_countOut = nullptr;
	uint32_t historyCap = 0;
	uint32_t * d_stage32 = nullptr;      // nGlobal u32 staging for the uint32 host interface
	uint32_t * h_pinned = nullptr;       // pinned bounce buffer for colours (nGlobal u32)
	bool colorsInit = false;
	bool pendingCountOnly = false;
	uint32_t hostSweepUpper = 0;         // upper bound of the device-side sweep index
	int gridBlocks = 0;
	size_t smemBytes = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaStream_t streamA = nullptr;      // pass A of the blocked sweep runs here, concurrently with pass B on `stream`
	cudaEvent_t evFork = nullptr, evReset = nullptr;
	bool overlap = false;                // both passes co-resident on every SM (decided in configure_blocked)
	bool timed = false;
	uint64_t launches = 0;
	uint64_t z = 0;
	int smCount = 0;
	BlockedLayout bl;                    // source-blocked two-pass layout (valid => the sweeps use it)
	BinnedLayout bn;                     // degree-binned direct sweep (valid => used when bl is not)
	bool wide = false;                   // nCol > 512: wide_sweep_kernel over the binned lists (wide_sweep.cuh)
	float * d_wideS = nullptr, * d_wideTab = nullptr;
	uint8_t * d_wideFp = nullptr;                          // one-byte fingerprint of the current colouring (wide_fingerprint_kernel)
	uint32_t * d_wideQ = nullptr, * d_wideQcta = nullptr;   // rows queued by wide_count_kernel for wide_walk_kernel
	int wideGridCount = 0;
	uint32_t maskWords64 = 1;            // 64-bit words of one vertex' occupancy mask (debug interface)
	void * peerColors[2][kMaxPeers] = {};  // fused multi-GPU exchange: IPC-mapped colour buffers of every rank (own = local)
	uint32_t * d_violList[2] = {nullptr, nullptr};   // tail cutting: violating vertices emitted by the sweeps / left by the last repair pass
	uint32_t * d_violCount = nullptr; uint32_t violCap = 0;
	uint8_t * d_pending = nullptr;         // tail cutting: per-vertex flags (all zero between calls)
	uint32_t * d_flist = nullptr, * d_tcResume = nullptr;
	// distributed repair (mcmcb200_tailcut_dist_*): colour order, ready flags, hub rows of a round, (id, colour) pairs in and out
	uint32_t * d_tcOrder = nullptr, * d_tcHeavy = nullptr, * d_tcIo = nullptr, * d_tcCounters = nullptr; uint8_t * d_tcReady = nullptr;
	uint32_t tcIoCap = 0, tcFlagged = 0, tcListCount = 0; int tcSrc = 0; bool tcActive = false;
	TailcutCounters * d_tcCnt = nullptr;
	unsigned long long * d_xchg = nullptr; // this rank's counter-exchange block (sweep_kernel.cuh: cross_rank_reduce)
	unsigned long long * peerXchg[kMaxPeers] = {};
	uint32_t nPeers = 0, myPeerIndex = 0;
	bool split() const { return (p.flags & MCMCB200_FLAG_NO_FUSED_FINALIZE) != 0 && nPeers == 0; }   // caller reduces the counters (NCCL path)
};

namespace {

// NVTX range over a host entry point (the reference has no tracing at all)
struct NvtxRange {
	explicit NvtxRange(const char * name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
	NvtxRange(const NvtxRange &) = delete;
	NvtxRange & operator=(const NvtxRange &) = delete;
};

template <int W, typename ColT>
cudaError_t launch_sweep_t(mcmcb200_handle * h, const SweepArgs & a) {
	if (h->bl.valid) {
		// source-blocked path: pass A (gather through shared memory) + pass B (tile sweep); two launches per sweep
		const BlockedArgs b = make_blocked_args(h->bl);
		const size_t syncBytes = sizeof(uint32_t) * (2 + (size_t)h->bl.numParts);
		cudaStream_t sA = h->stream;
		cudaError_t e = cudaSuccess;
		if (h->overlap) {
			// pass A on its own stream, after everything queued on the main stream so far (previous sweep, colour exchange);
			// pass B on the main stream follows A part by part through the counters in bl.sync
			sA = h->streamA;
			e = cudaEventRecord(h->evFork, h->stream);
			if (e == cudaSuccess) e = cudaStreamWaitEvent(sA, h->evFork, 0);
		}
		if (e == cudaSuccess) e = cudaMemsetAsync(h->bl.sync, 0, syncBytes, sA);
#if MCMCB200_TIMING
		{ const unsigned long long init[8] = {~0ull, 0, ~0ull, 0, 0, 0, 0, 0}; cudaMemcpyAsync(h->bl.dbgTimes, init, sizeof(init), cudaMemcpyHostToDevice, sA); }
#endif
		if (e == cudaSuccess && h->overlap) e = cudaEventRecord(h->evReset, sA);   // B may start as soon as the counters are clean
		if (e != cudaSuccess) return e;
		// (measured and dropped: the first part as a separate launch at full occupancy, the rest co-resident with pass B --
		//  config 3 3.64 vs 3.62 ms, an 8-GPU-sized partition 0.544 vs 0.551 ms)
		blocked_gather_kernel<ColT><<<h->bl.gridA, kThreadsA, h->bl.smemA, sA>>>(a, b);
		h->launches++;
		if ((e = cudaGetLastError()) != cudaSuccess) return e;    // (pass B must not be launched without its producer)
		if (h->overlap && (e = cudaStreamWaitEvent(h->stream, h->evReset, 0)) != cudaSuccess) return e;
		// the plain production sweep (no replay tape, no taboo, no debug masks, not a count-only pass) has its own instances
		const bool plain = W <= 2 && !a.countOnly && !a.tape && !a.tabooIter && !a.dbgMasks;
		const bool dyn = a.proposal == MCMCB200_PROPOSAL_DYNAMIC;
		if (plain && dyn) blocked_sweep_kernel<W, ColT, true, (W <= 2)><<<h->bl.gridB, PassB<W>::threads, h->bl.smemB, h->stream>>>(a, b);
		else if (plain) blocked_sweep_kernel<W, ColT, false, (W <= 2)><<<h->bl.gridB, PassB<W>::threads, h->bl.smemB, h->stream>>>(a, b);
		else if (dyn) blocked_sweep_kernel<W, ColT, true><<<h->bl.gridB, PassB<W>::threads, h->bl.smemB, h->stream>>>(a, b);
		else blocked_sweep_kernel<W, ColT, false><<<h->bl.gridB, PassB<W>::threads, h->bl.smemB, h->stream>>>(a, b);
#if MCMCB200_TIMING
		{ unsigned long long tm[8]; cudaStreamSynchronize(h->stream); cudaStreamSynchronize(sA); cudaMemcpy(tm, h->bl.dbgTimes, sizeof(tm), cudaMemcpyDeviceToHost);
		  fprintf(stderr, "[timing] A %.3f ms (start +0) | B start +%.3f ms, end +%.3f ms | A end +%.3f ms | B waited %.3f ms summed over %d CTAs\n", (tm[1] - tm[0]) * 1e-6,
		          ((double)tm[2] - (double)tm[0]) * 1e-6, ((double)tm[3] - (double)tm[0]) * 1e-6, ((double)tm[1] - (double)tm[0]) * 1e-6, tm[4] * 1e-6, h->bl.gridB); }
#endif
		return cudaGetLastError();
	}
	if (h->bn.valid) {
		const BinnedArgs b = make_binned_args(h->bn);
		cudaError_t e = cudaMemsetAsync(h->bn.counters, 0, 4 * sizeof(uint32_t), h->stream);
		if (e != cudaSuccess) return e;
		if (a.proposal == MCMCB200_PROPOSAL_DYNAMIC) binned_sweep_kernel<W, ColT, true><<<h->bn.grid, kThreadsBin, h->bn.smem, h->stream>>>(a, b);
		else binned_sweep_kernel<W, ColT, false><<<h->bn.grid, kThreadsBin, h->bn.smem, h->stream>>>(a, b);
		return cudaGetLastError();
	}
	if (a.proposal == MCMCB200_PROPOSAL_DYNAMIC) sweep_kernel<W, ColT, true><<<h->gridBlocks, kThreads, h->smemBytes, h->stream>>>(a);
	else sweep_kernel<W, ColT, false><<<h->gridBlocks, kThreads, h->smemBytes, h->stream>>>(a);
	return cudaGetLastError();
}

// shared-memory opt-in and grid sizes of the two blocked kernels for this handle's template instance
template <int W, typename ColT>
cudaError_t configure_blocked_t(mcmcb200_handle * h) {
	BlockedLayout & L = h->bl;
	L.smemA = (size_t)kChunkV * sizeof(ColT);
	L.smemB = blocked_smem_bytes_B(h->p.nCol, L.nbuf, L.TV, L.stageCap, (int)sizeof(ColT), W);
	int dev = 0, optin = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	if (e != cudaSuccess) return e;
	if (L.smemA > (size_t)optin || L.smemB > (size_t)optin) { L.valid = false; return cudaSuccess; }   // does not fit: direct kernel
	e = cudaFuncSetAttribute(blocked_gather_kernel<ColT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemA);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(blocked_sweep_kernel<W, ColT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemB);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(blocked_sweep_kernel<W, ColT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemB);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(blocked_sweep_kernel<W, ColT, true, (W <= 2)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemB);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(blocked_sweep_kernel<W, ColT, false, (W <= 2)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemB);
	if (e != cudaSuccess) return e;
	int oa = 0, ob0 = 0, ob1 = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oa, blocked_gather_kernel<ColT>, kThreadsA, L.smemA);
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ob0, blocked_sweep_kernel<W, ColT, false>, PassB<W>::threads, L.smemB);
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ob1, blocked_sweep_kernel<W, ColT, true>, PassB<W>::threads, L.smemB);
	if (e != cudaSuccess) return e;
	int ob = ob0 < ob1 ? ob0 : ob1;
	if (oa < 1 || ob < 1) { L.valid = false; return cudaSuccess; }
	L.gridB = (int)std::max<uint32_t>(1u, std::min<uint32_t>(L.numTiles, (uint32_t)(ob * h->smCount)));
	// Overlap: pass A is DRAM bound, pass B issue bound.  With `ob` pass-B CTAs resident an SM must still have room
	// (shared memory, registers, threads) for at least one pass-A CTA -- otherwise a grid of waiting B CTAs could keep A out.
	cudaFuncAttributes fa{}, fb{};
	e = cudaFuncGetAttributes(&fa, blocked_gather_kernel<ColT>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&fb, blocked_sweep_kernel<W, ColT, false>);
	if (e != cudaSuccess) return e;
	int smemSM = 0, regsSM = 0, thrSM = 0;
	cudaDeviceGetAttribute(&smemSM, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
	cudaDeviceGetAttribute(&regsSM, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
	cudaDeviceGetAttribute(&thrSM, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
	auto regsOf = [](int perThread, int threads) { return ((perThread + 7) / 8 * 8) * ((threads + 31) / 32 * 32); };
	const long freeSmem = (long)smemSM - (long)ob * (long)(L.smemB + 1024);
	const long freeRegs = (long)regsSM - (long)ob * regsOf(fb.numRegs, PassB<W>::threads);
	const long freeThr = (long)thrSM - (long)ob * PassB<W>::threads;
	long aFit = std::min<long>(freeSmem / (long)(L.smemA + 1024), std::min<long>(freeRegs / regsOf(fa.numRegs, kThreadsA), freeThr / kThreadsA));
	h->overlap = aFit >= 1 && !(h->p.flags & MCMCB200_FLAG_NO_OVERLAP);
	if (h->overlap) {
		L.gridA = (int)std::max<uint32_t>(1u, std::min<uint32_t>(L.numItems, (uint32_t)(std::min<long>(aFit, oa) * h->smCount)));
		if (!h->streamA) e = cudaStreamCreateWithFlags(&h->streamA, cudaStreamNonBlocking);
		if (e == cudaSuccess && !h->evFork) e = cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming);
		if (e == cudaSuccess && !h->evReset) e = cudaEventCreateWithFlags(&h->evReset, cudaEventDisableTiming);
		if (e != cudaSuccess) return e;
	} else {
		L.gridA = (int)std::max<uint32_t>(1u, std::min<uint32_t>(L.numItems, (uint32_t)(oa * h->smCount)));
	}
	return cudaSuccess;
}

template <int W, typename ColT>
cudaError_t configure_binned_t(mcmcb200_handle * h) {
	BinnedLayout & L = h->bn;
	L.smem = binned_smem_bytes(h->p.nCol, W);
	cudaError_t e = cudaFuncSetAttribute(binned_sweep_kernel<W, ColT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(binned_sweep_kernel<W, ColT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
	int o0 = 0, o1 = 0;
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, binned_sweep_kernel<W, ColT, false>, kThreadsBin, L.smem);
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, binned_sweep_kernel<W, ColT, true>, kThreadsBin, L.smem);
	if (e != cudaSuccess) return e;
	const int o = o0 < o1 ? o0 : o1;
	if (o < 1) { L.valid = false; return cudaSuccess; }
	L.grid = o * h->smCount;
	return cudaSuccess;
}

cudaError_t configure_binned(mcmcb200_handle * h) {
	switch (h->W) {
	case 1: return configure_binned_t<1, uint8_t>(h);
	case 2: return configure_binned_t<2, uint8_t>(h);
	case 4: return configure_binned_t<4, uint8_t>(h);
	default: return configure_binned_t<8, uint16_t>(h);
	}
}

cudaError_t launch_sweep_wide(mcmcb200_handle * h, const SweepArgs & a) {
	const BinnedArgs b = make_binned_args(h->bn);
	WideArgs wa{};
	wide_geometry(h->p.nCol, wa);
	wa.S = h->d_wideS; wa.tab = h->d_wideTab; wa.words64 = h->maskWords64; wa.fp = h->d_wideFp;
	const WideQueues wq{h->d_wideQ, h->d_wideQcta};
	cudaError_t e = cudaMemsetAsync(h->bn.counters, 0, 4 * sizeof(uint32_t), h->stream);
	if (e == cudaSuccess) e = cudaMemsetAsync(h->d_wideQ, 0, 8 * sizeof(uint32_t), h->stream);
	if (e != cudaSuccess) return e;
	const bool walk = !a.countOnly || a.dbgMasks != nullptr;      // a plain counting pass needs no occupancy at all
	if (!a.countOnly) { wide_tables_kernel<<<(h->p.nCol + 255) / 256, 256, 0, h->stream>>>(a, wa); h->launches++; }
	wide_fingerprint_kernel<<<h->smCount * 8, 256, 0, h->stream>>>(a, wa); h->launches++;
	const bool dyn = a.proposal == MCMCB200_PROPOSAL_DYNAMIC;
	if (dyn) wide_count_kernel<true><<<h->wideGridCount, kThreadsBin, 0, h->stream>>>(a, b, wa, wq, walk ? 0u : 1u);
	else wide_count_kernel<false><<<h->wideGridCount, kThreadsBin, 0, h->stream>>>(a, b, wa, wq, walk ? 0u : 1u);
	if ((e = cudaGetLastError()) != cudaSuccess || !walk) return e;
	h->launches++;
	if (dyn) wide_walk_kernel<true><<<h->bn.grid, kThreadsBin, h->bn.smem, h->stream>>>(a, wa, wq);
	else wide_walk_kernel<false><<<h->bn.grid, kThreadsBin, h->bn.smem, h->stream>>>(a, wa, wq);
	return cudaGetLastError();
}

cudaError_t configure_wide(mcmcb200_handle * h) {
	BinnedLayout & L = h->bn;
	WideArgs wa{};
	wide_geometry(h->p.nCol, wa);
	L.smem = wide_smem_bytes(wa);
	cudaError_t e = cudaFuncSetAttribute(wide_walk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(wide_walk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
	int o0 = 0, o1 = 0, c0 = 0, c1 = 0;
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, wide_walk_kernel<false>, kThreadsBin, L.smem);
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, wide_walk_kernel<true>, kThreadsBin, L.smem);
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c0, wide_count_kernel<false>, kThreadsBin, 0);
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c1, wide_count_kernel<true>, kThreadsBin, 0);
	if (e != cudaSuccess) return e;
	const int o = o0 < o1 ? o0 : o1, c = c0 < c1 ? c0 : c1;
	if (o < 1 || c < 1) { L.valid = false; return cudaSuccess; }
	L.grid = o * h->smCount;
	h->wideGridCount = c * h->smCount;
	// the rows the counting pass hands to the walk pass: at most every owned vertex (debug interface), hub rows separately
	if ((e = cudaMalloc(&h->d_wideFp, (size_t)h->nGlobal + kColorPad)) != cudaSuccess) return e;
	if ((e = cudaMalloc(&h->d_wideQ, sizeof(uint32_t) * ((size_t)h->nLocal + 8))) != cudaSuccess) return e;
	if ((e = cudaMalloc(&h->d_wideQcta, sizeof(uint32_t) * ((size_t)L.n[2] + 1))) != cudaSuccess) return e;
	return cudaSuccess;
}

cudaError_t launch_sweep(mcmcb200_handle * h, const SweepArgs & a) {
	h->launches++;   // (the blocked path counts its first pass itself)
	if (h->wide) return launch_sweep_wide(h, a);
	switch (h->W) {
	case 1: return launch_sweep_t<1, uint8_t>(h, a);
	case 2: return launch_sweep_t<2, uint8_t>(h, a);
	case 4: return launch_sweep_t<4, uint8_t>(h, a);
	default: return launch_sweep_t<8, uint16_t>(h, a);
	}
}

cudaError_t configure_blocked(mcmcb200_handle * h) {
	switch (h->W) {
	case 1: return configure_blocked_t<1, uint8_t>(h);
	case 2: return configure_blocked_t<2, uint8_t>(h);
	case 4: return configure_blocked_t<4, uint8_t>(h);
	default: return configure_blocked_t<8, uint16_t>(h);
	}
}

template <int W, typename ColT>
cudaError_t occupancy_t(int * blocks, size_t smem) {
	int a = 0, b = 0;
	// (more than the 48 KiB default for the wide palettes: mask rows + walk queue)
	cudaError_t e = cudaFuncSetAttribute(sweep_kernel<W, ColT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(sweep_kernel<W, ColT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sweep_kernel<W, ColT, false>, kThreads, smem);
	if (e != cudaSuccess) return e;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sweep_kernel<W, ColT, true>, kThreads, smem);
	*blocks = a < b ? a : b;
	return e;
}

cudaError_t sweep_occupancy(mcmcb200_handle * h, int * blocks) {
	switch (h->W) {
	case 1: return occupancy_t<1, uint8_t>(blocks, h->smemBytes);
	case 2: return occupancy_t<2, uint8_t>(blocks, h->smemBytes);
	case 4: return occupancy_t<4, uint8_t>(blocks, h->smemBytes);
	default: return occupancy_t<8, uint16_t>(blocks, h->smemBytes);
	}
}

SweepArgs make_args(mcmcb200_handle * h) {
	SweepArgs a{};
	a.rowptr = h->d_rowptr; a.neighs = h->d_neighs;
	a.nLocal = h->nLocal; a.vBegin = h->vBegin; a.nGlobal = h->nGlobal; a.nCol = h->p.nCol;
	a.numTiles = (h->nLocal + kTileV - 1) / kTileV;
	a.eps = h->p.epsilon; a.tabooIter = h->p.tabooIteration; a.proposal = h->p.proposal; a.seed = h->p.seed;
	a.colors[0] = h->d_colors[0]; a.colors[1] = h->d_colors[1];
	a.colorsOverride = nullptr;
	a.taboo = h->d_taboo;
	a.tape = h->tapeSweeps ? h->d_tape : nullptr; a.tapeBase = h->tapeBase;
	a.st = h->d_state; a.scratch = h->d_scratch; a.hist[0] = h->d_hist[0]; a.hist[1] = h->d_hist[1];
	a.history = h->d_history; a.historyCap = h->historyCap;
	a.countOnly = 0; a.countOut = nullptr;
	a.fuseFinalize = h->split() ? 0u : 1u;
	a.noEarlyStop = (h->p.flags & MCMCB200_FLAG_NO_EARLY_STOP) ? 1u : 0u;
	a.nPeers = h->bl.valid ? h->nPeers : 0u;
	for (int b = 0; b < 2; ++b) for (uint32_t q = 0; q < kMaxPeers; ++q) a.peerColors[b][q] = h->peerColors[b][q];
	for (uint32_t q = 0; q < kMaxPeers; ++q) a.peerXchg[q] = h->peerXchg[q];
	a.myRank = h->myPeerIndex;
	a.dbgMasks = nullptr; a.dbgSame = nullptr;
	a.violList = h->d_violList[0]; a.violCount = h->d_violCount; a.violCap = h->violCap; a.forceEmit = 0;
	a.emitThreshold = h->violCap ? (unsigned long long)h->violCap : 0ull;
	return a;
}

int check_params(const mcmcb200_params * p, uint32_t nGlobal) {
	if (!p || p->nCol == 0 || nGlobal == 0) return MCMCB200_EINVAL;
	if (p->proposal > 1u || p->convergence > 1u) return MCMCB200_EINVAL;
	if (!(p->epsilon >= 0.0f)) return MCMCB200_EINVAL;
	if (p->nCol > 65535u) return MCMCB200_EUNSUPPORTED;               // colours are at most 16 bits on the device
	if (p->nCol > 64u * kMaxColWords && (p->flags & (MCMCB200_FLAG_FORCE_BLOCKED | MCMCB200_FLAG_FORCE_DIRECT))) return MCMCB200_EUNSUPPORTED;   // wide palettes: wide_sweep_kernel only
	if (p->tabooIteration > 65535u) return MCMCB200_EUNSUPPORTED;
	if (p->proposal == MCMCB200_PROPOSAL_DYNAMIC && p->nCol < 2) return MCMCB200_EINVAL;
	if (p->stageBuffers > 2u || (p->itemBits && (p->itemBits < 12u || p->itemBits > 24u))) return MCMCB200_EINVAL;
	return MCMCB200_OK;
}

int select_device(const mcmcb200_params * p, int * devOut, int * smCount) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		snprintf(g_lastCudaError, sizeof(g_lastCudaError), "no CUDA device: %s", cudaGetErrorString(e));
		cudaGetLastError();
		return MCMCB200_ENODEVICE;
	}
	int dev = p->device;
	if (dev < 0) CU(cudaGetDevice(&dev));
	if (dev >= count) return MCMCB200_ENODEVICE;
	CU(cudaSetDevice(dev));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, dev));
	if (prop.major != 10) {   // the library holds sm_100a SASS only
		snprintf(g_lastCudaError, sizeof(g_lastCudaError), "device %d is sm_%d%d; libmcmcb200 is built for sm_100a only",
		         dev, prop.major, prop.minor);
		return MCMCB200_ENODEVICE;
	}
	*devOut = dev; *smCount = prop.multiProcessorCount;
	return MCMCB200_OK;
}

// common tail of the three create variants: everything except the CSR
int alloc_chain_state(mcmcb200_handle * h) {
	const uint32_t nCol = h->p.nCol;
	h->W = nCol <= 64 ? 1 : nCol <= 128 ? 2 : nCol <= 256 ? 4 : 8;
	h->wide = nCol > 64u * kMaxColWords;
	h->maskWords64 = h->wide ? (nCol + 63u) / 64u : (uint32_t)h->W;
	h->colBytes = nCol <= 256 ? 1 : 2;
	h->z = h->p.tailcut ? std::max<uint64_t>(50, h->nGlobal / 2000) : 0;   // coloringMCMC_main.cu:150-157
	h->historyCap = std::max<uint32_t>(h->p.maxRip + 2u, 1024u);
	CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	CU(cudaEventCreate(&h->ev0)); CU(cudaEventCreate(&h->ev1));
	const size_t colElems = (size_t)h->nGlobal + kColorPad;
	for (int i = 0; i < 2; ++i) {
		CU(cudaMalloc(&h->d_colors[i], colElems * h->colBytes));
		CU(cudaMemsetAsync(h->d_colors[i], 0, colElems * h->colBytes, h->stream));
	}
	if (h->p.tabooIteration) CU(cudaMalloc(&h->d_taboo, sizeof(uint16_t) * std::max<size_t>(h->nLocal, 1)));
	CU(cudaMalloc(&h->d_state, sizeof(DevState)));
	CU(cudaMalloc(&h->d_scratch, sizeof(unsigned long long) * (nCol + 2)));
	CU(cudaMalloc(&h->d_hist[0], sizeof(unsigned long long) * nCol));
	CU(cudaMalloc(&h->d_hist[1], sizeof(unsigned long long) * nCol));
	CU(cudaMalloc(&h->d_history, sizeof(unsigned long long) * 2 * h->historyCap));
	CU(cudaMalloc(&h->d_countOut, sizeof(unsigned long long) * 2));
	if (h->p.tailcut) {
		// room for the violators of a colouring up to 32 z away from the threshold (sweep_kernel.cuh: emitThreshold); a partition
		// lists the violators it owns (distributed repair: mcmcb200_tailcut_dist_*)
		h->violCap = (uint32_t)std::min<uint64_t>(h->nGlobal, 32ull * h->z + 4096ull);
		for (int i = 0; i < 2; ++i) CU(cudaMalloc(&h->d_violList[i], sizeof(uint32_t) * (size_t)h->violCap));
		CU(cudaMalloc(&h->d_flist, sizeof(uint32_t) * (size_t)h->violCap));
		CU(cudaMalloc(&h->d_tcResume, sizeof(uint32_t) * (size_t)h->violCap));
		CU(cudaMalloc(&h->d_violCount, sizeof(uint32_t)));
		CU(cudaMemsetAsync(h->d_violCount, 0, sizeof(uint32_t), h->stream));
		CU(cudaMalloc(&h->d_tcCnt, sizeof(TailcutCounters)));
		CU(cudaMalloc(&h->d_pending, std::max<size_t>(h->nGlobal, 1)));
		CU(cudaMemsetAsync(h->d_pending, 0, std::max<size_t>(h->nGlobal, 1), h->stream));
	}
	CU(cudaMalloc(&h->d_xchg, sizeof(unsigned long long) * xchg_words(nCol)));
	CU(cudaMemsetAsync(h->d_xchg, 0, sizeof(unsigned long long) * xchg_words(nCol), h->stream));
	CU(cudaMemsetAsync(h->d_scratch, 0, sizeof(unsigned long long) * (nCol + 2), h->stream));
	CU(cudaMemsetAsync(h->d_history, 0, sizeof(unsigned long long) * 2 * h->historyCap, h->stream));
	if (h->wide) {
		CU(cudaMalloc(&h->d_wideS, sizeof(float) * ((size_t)nCol + 1)));
		CU(cudaMalloc(&h->d_wideTab, sizeof(float) * ((size_t)nCol + 1)));
		CU(cudaMemsetAsync(h->d_wideTab, 0, sizeof(float) * ((size_t)nCol + 1), h->stream));
		wide_S_kernel<<<1, 32, 0, h->stream>>>(h->d_wideS, nCol, h->p.epsilon);
		h->launches++;
		CU(cudaGetLastError());
	} else {
		h->smemBytes = sweep_smem_bytes(nCol, h->W, h->colBytes);
		int perSM = 0;
		CU(sweep_occupancy(h, &perSM));
		if (perSM < 1) return MCMCB200_EUNSUPPORTED;
		const uint32_t numTiles = (h->nLocal + kTileV - 1) / kTileV;
		h->gridBlocks = (int)std::max<uint32_t>(1u, std::min<uint32_t>(numTiles, (uint32_t)(perSM * h->smCount)));
	}
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int reset_state(mcmcb200_handle * h) {
	DevState s{};
	s.sweep = 0; s.convergedAt = -1; s.ticket = 0; s.tileCounter = 0;
	s.convergence = h->p.convergence; s.countsSweep = 0xffffffffu; s.z = h->z;
	s.lastDirected = 0; s.lastViol = 0; s.errorFlag = 0; s.xseq = 0;
	s.emitNow = 0; s.violListSweep = 0xffffffffu; s.violListCount = 0;
	if (h->d_violCount) CU(cudaMemsetAsync(h->d_violCount, 0, sizeof(uint32_t), h->stream));
	CU(cudaMemcpyAsync(h->d_state, &s, sizeof(s), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemsetAsync(h->d_xchg, 0, sizeof(unsigned long long) * xchg_words(h->p.nCol), h->stream));   // (all ranks reset together: init_colors is collective)
	CU(cudaMemsetAsync(h->d_scratch, 0, sizeof(unsigned long long) * (h->p.nCol + 2), h->stream));
	if (h->d_taboo) CU(cudaMemsetAsync(h->d_taboo, 0, sizeof(uint16_t) * std::max<size_t>(h->nLocal, 1), h->stream));
	h->hostSweepUpper = 0; h->tapeBase = 0; h->pendingCountOnly = false;
	h->tcActive = false;                                          // (a distributed repair that was left half-way does not survive a new colouring)
	return MCMCB200_OK;
}

int ensure_stage(mcmcb200_handle * h) {
	if (!h->d_stage32) CU(cudaMalloc(&h->d_stage32, sizeof(uint32_t) * (size_t)h->nGlobal));
	return MCMCB200_OK;
}

int narrow_into(mcmcb200_handle * h, const uint32_t * hostColors, void * dst) {
	int rc = ensure_stage(h); if (rc) return rc;
	CU(cudaMemcpyAsync(h->d_stage32, hostColors, sizeof(uint32_t) * (size_t)h->nGlobal, cudaMemcpyHostToDevice, h->stream));
	const uint32_t n = h->nGlobal, blocks = (n + 255) / 256;
	if (h->colBytes == 1) narrow_colors_kernel<uint8_t><<<blocks, 256, 0, h->stream>>>(h->d_stage32, (uint8_t *)dst, n, h->p.nCol, h->d_state);
	else narrow_colors_kernel<uint16_t><<<blocks, 256, 0, h->stream>>>(h->d_stage32, (uint16_t *)dst, n, h->p.nCol, h->d_state);
	h->launches++;
	CU(cudaGetLastError());
	return MCMCB200_OK;
}

int compute_class_sizes(mcmcb200_handle * h, const void * colors, unsigned long long * hist) {
	CU(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * h->p.nCol, h->stream));
	const int blocks = std::max(1, std::min<int>(h->smCount * 4, (int)((h->nGlobal + 255) / 256)));
	const size_t smem = h->p.nCol <= 8192u ? sizeof(unsigned int) * h->p.nCol : 0;   // (palettes beyond 8 Ki colours: global atomics)
	if (h->colBytes == 1) class_sizes_kernel<uint8_t><<<blocks, 256, smem, h->stream>>>((const uint8_t *)colors, h->nGlobal, h->p.nCol, hist, h->d_state);
	else class_sizes_kernel<uint16_t><<<blocks, 256, smem, h->stream>>>((const uint16_t *)colors, h->nGlobal, h->p.nCol, hist, h->d_state);
	h->launches++;
	CU(cudaGetLastError());
	return MCMCB200_OK;
}

int read_state_raw(mcmcb200_handle * h, DevState * s) {
	CU(cudaMemcpyAsync(s, h->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

cudaError_t launch_sweep(mcmcb200_handle * h, const SweepArgs & a);
SweepArgs make_args(mcmcb200_handle * h);

// Every host call that reads results goes through here.  If an overlapped blocked sweep was aborted on the device (pass B timed
// out waiting for pass A: the two kernels were not co-scheduled -- another tenant on the GPU, MPS, a tool serialising launches)
// nothing of it was committed; the handle drops to the back-to-back mode for good and the lost sweeps are run again.
int read_state(mcmcb200_handle * h, DevState * s) {
	int rc = read_state_raw(h, s); if (rc) return rc;
	if (s->errorFlag != 2u) return MCMCB200_OK;
	if (!h->overlap) {
		snprintf(g_lastCudaError, sizeof(g_lastCudaError), "blocked sweep aborted on the device (sweep %u)", s->sweep);
		return MCMCB200_ECUDA;
	}
	h->overlap = false;
	if (h->streamA) CU(cudaStreamSynchronize(h->streamA));
	if (h->p.flags & MCMCB200_FLAG_NO_FUSED_FINALIZE) {       // multi-GPU: the ranks run in lock step, one of them cannot redo a sweep alone
		snprintf(g_lastCudaError, sizeof(g_lastCudaError), "overlapped blocked sweep aborted on the device (sweep %u); handle switched to back-to-back passes", s->sweep);
		return MCMCB200_ECUDA;
	}
	const uint32_t zero = 0;
	CU(cudaMemcpyAsync(&h->d_state->errorFlag, &zero, sizeof(zero), cudaMemcpyHostToDevice, h->stream));
	const uint32_t lost = (s->convergedAt < 0 && h->hostSweepUpper > s->sweep) ? h->hostSweepUpper - s->sweep : 0u;
	for (uint32_t i = 0; i < lost; ++i) { SweepArgs a = make_args(h); CU(launch_sweep(h, a)); }
	rc = read_state_raw(h, s); if (rc) return rc;
	if (s->errorFlag == 2u) {
		snprintf(g_lastCudaError, sizeof(g_lastCudaError), "blocked sweep aborted on the device (sweep %u)", s->sweep);
		return MCMCB200_ECUDA;
	}
	return MCMCB200_OK;
}

int create_common(mcmcb200_handle ** out, uint32_t nGlobal, uint32_t vBegin, uint32_t vEnd, uint64_t nnzLocal,
                  const uint32_t * cumulDegs, const uint32_t * neighs /* first owned edge */, bool deviceCsr,
                  const mcmcb200_params * p) {
	NvtxRange nvtx_("mcmcb200_create");
	if (!out) return MCMCB200_EINVAL;
	*out = nullptr;
	int rc = check_params(p, nGlobal); if (rc) return rc;
	if (vBegin > vEnd || vEnd > nGlobal || !cumulDegs || (nnzLocal && !neighs)) return MCMCB200_EINVAL;
	if (nnzLocal >= 0xfffffff0ull) return MCMCB200_EUNSUPPORTED;   // 32-bit CSR offsets, like the reference (graph.h:19-20)
	if (vBegin & 255u) return MCMCB200_EINVAL;                      // owned colours move as 16-byte vectors / bulk copies (mcmcb200.h)
	int dev = 0, sms = 0;
	rc = select_device(p, &dev, &sms); if (rc) return rc;
	mcmcb200_handle * h = new (std::nothrow) mcmcb200_handle;
	if (!h) return MCMCB200_ENOMEM;
	h->p = *p; h->device = dev; h->smCount = sms;
	h->nGlobal = nGlobal; h->vBegin = vBegin; h->vEnd = vEnd; h->nLocal = vEnd - vBegin; h->nnzLocal = nnzLocal;
	auto fail = [&](int code) { mcmcb200_destroy(h); return code; };
	if (deviceCsr) {
		if ((reinterpret_cast<uintptr_t>(neighs) & 31u) != 0) return fail(MCMCB200_EINVAL);
		h->ownsCsr = false;
		h->d_rowptr = const_cast<uint32_t *>(cumulDegs);
		h->d_neighs = const_cast<uint32_t *>(neighs);
	} else {
		h->ownsCsr = true;
		// rebase the offsets to 0 and validate monotonicity / ids on the host (cheap, once)
		std::vector<uint32_t> rp((size_t)h->nLocal + 1);
		const uint32_t base = cumulDegs[0];
		for (size_t i = 0; i <= h->nLocal; ++i) {
			if (cumulDegs[i] < base || (i && cumulDegs[i] < cumulDegs[i - 1])) return fail(MCMCB200_EINVAL);
			rp[i] = cumulDegs[i] - base;
		}
		if (rp[h->nLocal] != nnzLocal) return fail(MCMCB200_EINVAL);
		for (uint64_t e = 0; e < nnzLocal; ++e) if (neighs[e] >= nGlobal) return fail(MCMCB200_EINVAL);
		cudaError_t e1 = cudaMalloc(&h->d_rowptr, sizeof(uint32_t) * ((size_t)h->nLocal + 1));
		if (e1 != cudaSuccess) return fail(cuda_fail(e1, "cudaMalloc(rowptr)", __LINE__));
		const size_t padded = ((size_t)nnzLocal + 15) & ~(size_t)7;   // 256-bit loads may touch up to 7 ids past the end
		e1 = cudaMalloc(&h->d_neighs, sizeof(uint32_t) * std::max<size_t>(padded, 8));
		if (e1 != cudaSuccess) return fail(cuda_fail(e1, "cudaMalloc(neighs)", __LINE__));
		e1 = cudaMemset(h->d_neighs, 0, sizeof(uint32_t) * std::max<size_t>(padded, 8));
		if (e1 == cudaSuccess) e1 = cudaMemcpy(h->d_rowptr, rp.data(), sizeof(uint32_t) * rp.size(), cudaMemcpyHostToDevice);
		if (e1 == cudaSuccess && nnzLocal) e1 = cudaMemcpy(h->d_neighs, neighs, sizeof(uint32_t) * nnzLocal, cudaMemcpyHostToDevice);
		if (e1 != cudaSuccess) return fail(cuda_fail(e1, "CSR upload", __LINE__));
	}
	rc = alloc_chain_state(h);
	if (rc) return fail(rc);
	{
		// kernel choice: the source-blocked two-pass sweep pays off once the random colour gathers dominate (large sparse
		// graphs); small graphs keep the single-pass direct kernel.  MCMCB200_FLAG_FORCE_{DIRECT,BLOCKED} override.
		const bool forceDirect = (p->flags & MCMCB200_FLAG_FORCE_DIRECT) != 0, forceBlocked = (p->flags & MCMCB200_FLAG_FORCE_BLOCKED) != 0;
		const bool forceBinned = (p->flags & MCMCB200_FLAG_FORCE_BINNED) != 0;
		const bool large = nnzLocal >= (1ull << 22) && nGlobal >= (1u << 18);
		// a handle that will only run a few sweeps does not amortise the layout build (mcmcb200.h: expectedSweeps)
		const bool fewSweeps = p->expectedSweeps != 0u && p->expectedSweeps < 16u;
		const bool want = !h->wide && (forceBlocked || (!forceDirect && !forceBinned && large && !fewSweeps));
		if (want) {
			// colour bytes a tile stages in shared memory.  Large partitions: 32 KiB (tiles of 4 x 384 vertices on a mean-degree-16
			// graph), so that two pass-B CTAs and one pass-A CTA (64 KiB chunk) share an SM and the two passes overlap; measured on
			// config 3 (profiles/r02*): 24 KiB 4.38 ms, 32 KiB 3.64, 36 KiB 4.35, 44 KiB 4.01 -- short tiles keep pass B's per-tile
			// phases short, too short ones cut pass A's runs.  Smaller partitions (fill/drain of the A->B pipeline would eat the
			// gain; measured on config 5): the largest stage the 16-bit positions allow, passes back to back
			// (round 2, pass A with 12 granules in flight: the overlapped pair also wins on 1.6e8-edge partitions -- config 5: 0.362 ms
			//  against 0.409 back to back -- so the threshold moved from 2^29 to 2^27 edges: 8-GPU ranks of config 3 overlap too)
			uint32_t capBytes = nnzLocal >= (1ull << 27) ? 32768u : 65504u;
			uint32_t nbuf = 1u;
			// pass-A work item: 2^16 entries of one bucket when the passes overlap (the chunk reload is one bulk copy now; 2^15:
			// 4.04 ms, 2^16: 3.64, 2^17: 3.70, 2^18: 3.81 on config 3), 2^17 back to back
			uint32_t itemEntries = nnzLocal >= (1ull << 27) ? (1u << 16) : (1u << 17);
			if (p->itemBits) itemEntries = 1u << p->itemBits;
			if (p->stageCapBytes) capBytes = p->stageCapBytes;
			if (p->stageBuffers) nbuf = p->stageBuffers;
			NvtxRange nvtxBuild("build_blocked_layout");
			cudaError_t e = build_blocked_layout(h->bl, h->d_rowptr, h->d_neighs, h->nLocal, nnzLocal, nGlobal, h->colBytes, capBytes, itemEntries,
			                                     (uint32_t)(h->W <= 2 ? PassB<1>::threads : PassB<4>::threads), h->stream, &h->launches);
			h->bl.nbuf = nbuf;
			if (e == cudaSuccess && h->bl.valid) e = configure_blocked(h);
			if (e == cudaSuccess && !h->bl.valid) free_blocked_layout(h->bl);
			if (e != cudaSuccess) return fail(cuda_fail(e, "build_blocked_layout", __LINE__));
			if (forceBlocked && !h->bl.valid) return fail(MCMCB200_EUNSUPPORTED);
		}
		// large graph whose rows do not fit the blocked layout (hubs): degree-binned direct sweep instead of the tile-synchronous one
		if (!h->bl.valid && !forceDirect && !forceBlocked && (forceBinned || large || h->wide)) {
			cudaError_t e = build_binned_layout(h->bn, h->d_rowptr, h->nLocal, h->stream, &h->launches);
			if (e == cudaSuccess && h->bn.valid) e = h->wide ? configure_wide(h) : configure_binned(h);
			if (e != cudaSuccess) return fail(cuda_fail(e, "build_binned_layout", __LINE__));
			if (!h->bn.valid) free_binned_layout(h->bn);
			if ((forceBinned || h->wide) && !h->bn.valid) return fail(h->nLocal == 0 && h->wide ? MCMCB200_EUNSUPPORTED : MCMCB200_ENOMEM);
		}
	}
	rc = reset_state(h);
	if (rc) return fail(rc);
	*out = h;
	return MCMCB200_OK;
}

int run_count_pass(mcmcb200_handle * h, const void * overrideColors, unsigned long long * countOut,
                   unsigned long long * dbgMasks, uint32_t * dbgSame) {
	SweepArgs a = make_args(h);
	a.countOnly = 1; a.colorsOverride = overrideColors; a.countOut = countOut;
	a.dbgMasks = dbgMasks; a.dbgSame = dbgSame; a.fuseFinalize = 1; a.tape = nullptr;
	if (overrideColors || countOut || dbgMasks) { a.nPeers = 0; a.violList = nullptr; }   // a pure function of a colouring: this rank's rows only, no collective, no list
	CU(launch_sweep(h, a));
	return MCMCB200_OK;
}

// List-driven tail cutting (tailcut_kernel.cuh).  Returns MCMCB200_OK / an error, or 1 when the violator list cannot be had
// (more violators than the list holds): the caller then uses the full-scan path.
template <typename ColT>
int tailcut_from_list_t(mcmcb200_handle * h, DevState & s, void * curV, unsigned long long * hist, const std::vector<uint32_t> & order,
                        uint32_t maxRounds, uint32_t * rounds) {
	ColT * cur = static_cast<ColT *>(curV);
	const uint32_t nCol = h->p.nCol;
	if (s.violListSweep != s.sweep || s.violListCount > h->violCap) {
		// no list for this colouring (chain stopped by maxRip, colours set by the caller, ...): one count-only pass emits it
		SweepArgs a = make_args(h);
		a.countOnly = 1; a.fuseFinalize = 1; a.tape = nullptr; a.forceEmit = 1; a.nPeers = 0;
		CU(cudaMemsetAsync(h->d_violCount, 0, sizeof(uint32_t), h->stream));
		CU(launch_sweep(h, a));
		int rc = read_state(h, &s); if (rc) return rc;
		if (s.violListSweep != s.sweep || s.violListCount > h->violCap) return 1;
	}
	uint32_t * d_order = nullptr;
	CU(cudaMalloc(&d_order, sizeof(uint32_t) * nCol));
	cudaError_t e = cudaMemcpyAsync(d_order, order.data(), sizeof(uint32_t) * nCol, cudaMemcpyHostToDevice, h->stream);
	uint32_t listCount = s.violListCount, used = 0;
	int src = 0;
	TailcutCounters c{};
	for (; e == cudaSuccess && used < maxRounds; ++used) {                 // while (conflictCounter > 0), _main.cu:279
		if (listCount == 0) break;
		if ((e = cudaMemsetAsync(h->d_tcCnt, 0, sizeof(TailcutCounters), h->stream)) != cudaSuccess) break;
		const uint32_t lb = (uint32_t)(((uint64_t)listCount * 32u + 255u) / 256u);   // a warp per listed vertex
		tc_filter_kernel<ColT><<<lb, 256, 0, h->stream>>>(h->d_rowptr, h->d_neighs, cur, h->d_violList[src], listCount, h->d_pending, h->d_flist, h->d_tcCnt);
		if ((e = cudaMemsetAsync(h->d_tcResume, 0, sizeof(uint32_t) * (size_t)h->violCap, h->stream)) != cudaSuccess) break;
		tc_rounds_kernel<ColT><<<1, kTcThreads, tc_smem_bytes(nCol), h->stream>>>(h->d_rowptr, h->d_neighs, nCol, cur, h->d_pending, h->d_flist, h->d_tcResume, h->d_tcCnt, d_order, hist);
		tc_recount_kernel<ColT><<<lb, 256, 0, h->stream>>>(h->d_rowptr, h->d_neighs, cur, h->d_violList[src], listCount, h->d_pending, h->d_violList[src ^ 1], h->d_tcCnt);
		h->launches += 3;
		if ((e = cudaMemcpyAsync(&c, h->d_tcCnt, sizeof(c), cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) break;
		if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) break;
		src ^= 1;
		listCount = c.nextCount;
		if (c.flagged == 0) break;                                         // the pass found nothing to repair
		if (c.inexact) { ++used; break; }
		if (c.changed == 0) { ++used; break; }                             // no vertex could be given another colour: further passes would repeat this one
		if (c.nextFlagged == 0) { ++used; break; }
	}
	if (e == cudaSuccess && src == 1) e = cudaMemcpyAsync(h->d_violList[0], h->d_violList[1], sizeof(uint32_t) * (size_t)listCount, cudaMemcpyDeviceToDevice, h->stream);
	if (e == cudaSuccess) {
		if (c.inexact) {                                                   // a repaired vertex found every colour taken: recount with a full pass at the next status
			const uint32_t stale = 0xffffffffu; const int32_t notConv = -1;
			e = cudaMemcpyAsync(&h->d_state->countsSweep, &stale, sizeof(stale), cudaMemcpyHostToDevice, h->stream);
			if (e == cudaSuccess) e = cudaMemcpyAsync(&h->d_state->convergedAt, &notConv, sizeof(notConv), cudaMemcpyHostToDevice, h->stream);
			if (e == cudaSuccess) e = cudaMemcpyAsync(&h->d_state->violListSweep, &stale, sizeof(stale), cudaMemcpyHostToDevice, h->stream);
		} else if (used > 0 || s.violListCount > 0) {
			tc_commit_kernel<<<1, 1, 0, h->stream>>>(h->d_state, h->d_tcCnt);
			h->launches++;
			e = cudaGetLastError();
		}
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
	cudaFree(d_order);
	if (e != cudaSuccess) return cuda_fail(e, "tailcut (list)", __LINE__);
	if (rounds) *rounds = used;
	// a repaired vertex found every colour taken and kept a clashing one: vertices outside the list may violate now, so the
	// remaining passes (the reference keeps going while conflicts are left, _main.cu:279) rescan the graph
	if (c.inexact && c.nextFlagged > 0 && used < maxRounds) return 1;
	return MCMCB200_OK;
}

int tailcut_from_list(mcmcb200_handle * h, DevState & s, void * cur, unsigned long long * hist, const std::vector<uint32_t> & order,
                      uint32_t maxRounds, uint32_t * rounds) {
	return h->colBytes == 1 ? tailcut_from_list_t<uint8_t>(h, s, cur, hist, order, maxRounds, rounds)
	                        : tailcut_from_list_t<uint16_t>(h, s, cur, hist, order, maxRounds, rounds);
}

} // namespace

extern "C" {

int mcmcb200_abi_version(void) { return MCMCB200_ABI_VERSION; }

const char * mcmcb200_last_cuda_error(void) { return g_lastCudaError; }

const char * mcmcb200_strerror(int code) {
	switch (code) {
	case MCMCB200_OK: return "ok";
	case MCMCB200_EINVAL: return "invalid argument";
	case MCMCB200_ENODEVICE: return "no usable sm_100a CUDA device (libmcmcb200 has no CPU fallback)";
	case MCMCB200_ECUDA: return "CUDA runtime error (see mcmcb200_last_cuda_error)";
	case MCMCB200_ENOMEM: return "out of memory";
	case MCMCB200_EUNSUPPORTED: return "configuration not supported by this build";
	case MCMCB200_ETAPE: return "replay tape exhausted";
	case MCMCB200_ESTATE: return "call out of order";
	default: return "unknown error";
	}
}

int mcmcb200_create(mcmcb200_handle ** out, uint32_t n, uint64_t nnz, const uint32_t * cumulDegs,
                    const uint32_t * neighs, const mcmcb200_params * p) {
	if (!cumulDegs) return MCMCB200_EINVAL;
	return create_common(out, n, 0, n, nnz, cumulDegs, neighs ? neighs + cumulDegs[0] : nullptr, false, p);
}

int mcmcb200_create_partition(mcmcb200_handle ** out, uint32_t nGlobal, uint32_t vBegin, uint32_t vEnd,
                              const uint32_t * cumulDegs, const uint32_t * neighs, const mcmcb200_params * p) {
	if (!cumulDegs || vBegin > vEnd) return MCMCB200_EINVAL;
	const uint64_t nnzLocal = (uint64_t)cumulDegs[vEnd - vBegin] - cumulDegs[0];
	return create_common(out, nGlobal, vBegin, vEnd, nnzLocal, cumulDegs, neighs, false, p);
}

int mcmcb200_create_device_csr(mcmcb200_handle ** out, uint32_t nGlobal, uint32_t vBegin, uint32_t vEnd,
                               uint64_t nnzLocal, const uint32_t * d_cumulDegs, const uint32_t * d_neighs,
                               const mcmcb200_params * p) {
	return create_common(out, nGlobal, vBegin, vEnd, nnzLocal, d_cumulDegs, d_neighs, true, p);
}

int mcmcb200_csr_from_edges(uint32_t n, uint64_t m, const uint32_t * src, const uint32_t * dst, int device,
                            uint32_t ** d_cumulDegs, uint32_t ** d_neighs, uint64_t * nnz) {
	NvtxRange nvtx_("mcmcb200_csr_from_edges");
	if (!d_cumulDegs || !d_neighs || !nnz || n == 0 || (m && (!src || !dst))) return MCMCB200_EINVAL;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		snprintf(g_lastCudaError, sizeof(g_lastCudaError), "no CUDA device: %s", cudaGetErrorString(e));
		cudaGetLastError();
		return MCMCB200_ENODEVICE;
	}
	if (device >= 0) CU(cudaSetDevice(device));
	// host arrays are staged in device memory first
	uint32_t * tmp[2] = {nullptr, nullptr};
	const uint32_t * in[2] = {src, dst};
	for (int i = 0; i < 2 && m; ++i) {
		cudaPointerAttributes at{};
		e = cudaPointerGetAttributes(&at, in[i]);
		if (e != cudaSuccess) { cudaGetLastError(); at.type = cudaMemoryTypeUnregistered; }
		if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
			e = cudaMalloc(&tmp[i], sizeof(uint32_t) * m);
			if (e == cudaSuccess) e = cudaMemcpy(tmp[i], in[i], sizeof(uint32_t) * m, cudaMemcpyHostToDevice);
			if (e != cudaSuccess) { cudaFree(tmp[0]); cudaFree(tmp[1]); return cuda_fail(e, "edge list upload", __LINE__); }
			in[i] = tmp[i];
		}
	}
	uint64_t bad = 0;
	e = build_csr_from_edges(n, m, in[0], in[1], (cudaStream_t)0, d_cumulDegs, d_neighs, nnz, &bad);
	cudaFree(tmp[0]); cudaFree(tmp[1]);
	if (e == cudaErrorInvalidValue) { cudaGetLastError(); return bad ? MCMCB200_EINVAL : MCMCB200_EUNSUPPORTED; }   // id >= n / more than 2^32 entries
	if (e != cudaSuccess) return cuda_fail(e, "build_csr_from_edges", __LINE__);
	return MCMCB200_OK;
}

void mcmcb200_csr_free(uint32_t * d_cumulDegs, uint32_t * d_neighs) {
	cudaFree(d_cumulDegs); cudaFree(d_neighs);
	cudaGetLastError();
}

void mcmcb200_destroy(mcmcb200_handle * h) {
	if (!h) return;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->ownsCsr) { cudaFree(h->d_rowptr); cudaFree(h->d_neighs); }
	cudaFree(h->d_colors[0]); cudaFree(h->d_colors[1]); cudaFree(h->d_colorsTmp); cudaFree(h->d_taboo);
	cudaFree(h->d_tape); cudaFree(h->d_state); cudaFree(h->d_scratch); cudaFree(h->d_hist[0]); cudaFree(h->d_hist[1]);
	cudaFree(h->d_history); cudaFree(h->d_countOut); cudaFree(h->d_stage32);
	free_blocked_layout(h->bl);
	free_binned_layout(h->bn);
	cudaFree(h->d_wideS); cudaFree(h->d_wideTab); cudaFree(h->d_wideQ); cudaFree(h->d_wideQcta); cudaFree(h->d_wideFp);
	mcmcb200_ipc_detach(h);
	cudaFree(h->d_xchg);
	cudaFree(h->d_violList[0]); cudaFree(h->d_violList[1]); cudaFree(h->d_violCount); cudaFree(h->d_pending); cudaFree(h->d_flist); cudaFree(h->d_tcResume); cudaFree(h->d_tcCnt);
	cudaFree(h->d_tcOrder); cudaFree(h->d_tcHeavy); cudaFree(h->d_tcIo); cudaFree(h->d_tcCounters); cudaFree(h->d_tcReady);
	if (h->h_pinned) cudaFreeHost(h->h_pinned);
	if (h->streamA) { cudaStreamSynchronize(h->streamA); cudaStreamDestroy(h->streamA); }
	if (h->evFork) cudaEventDestroy(h->evFork);
	if (h->evReset) cudaEventDestroy(h->evReset);
	if (h->ev0) cudaEventDestroy(h->ev0);
	if (h->ev1) cudaEventDestroy(h->ev1);
	if (h->stream) cudaStreamDestroy(h->stream);
	cudaGetLastError();
	delete h;
}

int mcmcb200_init_colors(mcmcb200_handle * h, const uint32_t * colors) {
	NvtxRange nvtx_("mcmcb200_init_colors");
	if (!h) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	int rc = reset_state(h); if (rc) return rc;
	if (colors) {
		rc = narrow_into(h, colors, h->d_colors[0]); if (rc) return rc;
	} else {
		const uint32_t n = h->nGlobal, blocks = ((n + 3) / 4 + 255) / 256;
		if (h->colBytes == 1) init_colors_philox_kernel<uint8_t><<<blocks, 256, 0, h->stream>>>((uint8_t *)h->d_colors[0], n, h->p.nCol, h->p.seed);
		else init_colors_philox_kernel<uint16_t><<<blocks, 256, 0, h->stream>>>((uint16_t *)h->d_colors[0], n, h->p.nCol, h->p.seed);
		h->launches++;
		CU(cudaGetLastError());
	}
	rc = compute_class_sizes(h, h->d_colors[0], h->d_hist[0]); if (rc) return rc;
	DevState s;
	rc = read_state(h, &s); if (rc) return rc;
	if (s.errorFlag) { h->colorsInit = false; return MCMCB200_EINVAL; }   // a colour >= nCol
	h->colorsInit = true;
	return MCMCB200_OK;
}

int mcmcb200_color_bytes(mcmcb200_handle * h, uint32_t * elemBytes) {
	if (!h || !elemBytes) return MCMCB200_EINVAL;
	*elemBytes = (uint32_t)h->colBytes;
	return MCMCB200_OK;
}

int mcmcb200_init_colors_narrow(mcmcb200_handle * h, const void * colors, uint32_t elemBytes) {
	if (!h || !colors) return MCMCB200_EINVAL;
	if (elemBytes != (uint32_t)h->colBytes) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	int rc = reset_state(h); if (rc) return rc;
	CU(cudaMemcpyAsync(h->d_colors[0], colors, (size_t)h->nGlobal * h->colBytes, cudaMemcpyHostToDevice, h->stream));
	rc = compute_class_sizes(h, h->d_colors[0], h->d_hist[0]); if (rc) return rc;    // (also the colour < nCol check)
	DevState s;
	rc = read_state(h, &s); if (rc) return rc;
	if (s.errorFlag) { h->colorsInit = false; return MCMCB200_EINVAL; }
	h->colorsInit = true;
	return MCMCB200_OK;
}

int mcmcb200_get_colors_narrow(mcmcb200_handle * h, void * out, uint32_t elemBytes) {
	if (!h || !out) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	if (elemBytes != (uint32_t)h->colBytes) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	CU(cudaMemcpyAsync(out, h->d_colors[s.sweep & 1], (size_t)h->nGlobal * h->colBytes, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_init_colors_slice(mcmcb200_handle * h, const uint32_t * ownedColors) {
	if (!h || !ownedColors) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	int rc = reset_state(h); if (rc) return rc;
	rc = ensure_stage(h); if (rc) return rc;
	h->colorsInit = false;
	if (h->nLocal) {
		CU(cudaMemcpyAsync(h->d_stage32, ownedColors, sizeof(uint32_t) * (size_t)h->nLocal, cudaMemcpyHostToDevice, h->stream));
		const uint32_t blocks = (h->nLocal + 255) / 256;
		if (h->colBytes == 1) narrow_colors_kernel<uint8_t><<<blocks, 256, 0, h->stream>>>(h->d_stage32, (uint8_t *)h->d_colors[0] + h->vBegin, h->nLocal, h->p.nCol, h->d_state);
		else narrow_colors_kernel<uint16_t><<<blocks, 256, 0, h->stream>>>(h->d_stage32, (uint16_t *)h->d_colors[0] + h->vBegin, h->nLocal, h->p.nCol, h->d_state);
		h->launches++;
		CU(cudaGetLastError());
	}
	return MCMCB200_OK;
}

int mcmcb200_init_colors_slice_narrow(mcmcb200_handle * h, const void * ownedColors, uint32_t elemBytes) {
	if (!h || !ownedColors) return MCMCB200_EINVAL;
	if (elemBytes != (uint32_t)h->colBytes) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	int rc = reset_state(h); if (rc) return rc;
	h->colorsInit = false;
	if (h->nLocal)   // straight into the owned slice of the replica; the range check happens in _finish (class-size pass over the whole colouring)
		CU(cudaMemcpyAsync(static_cast<unsigned char *>(h->d_colors[0]) + (size_t)h->vBegin * h->colBytes, ownedColors, (size_t)h->nLocal * h->colBytes,
		                   cudaMemcpyHostToDevice, h->stream));
	return MCMCB200_OK;
}

int mcmcb200_get_colors_slice_narrow(mcmcb200_handle * h, void * out, uint32_t elemBytes) {
	if (!h || !out) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	if (elemBytes != (uint32_t)h->colBytes) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	if (h->nLocal)
		CU(cudaMemcpyAsync(out, static_cast<const unsigned char *>(h->d_colors[s.sweep & 1]) + (size_t)h->vBegin * h->colBytes, (size_t)h->nLocal * h->colBytes,
		                   cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_init_colors_finish(mcmcb200_handle * h) {
	if (!h) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	int rc = compute_class_sizes(h, h->d_colors[0], h->d_hist[0]); if (rc) return rc;
	DevState s;
	rc = read_state(h, &s); if (rc) return rc;
	if (s.errorFlag) { h->colorsInit = false; return MCMCB200_EINVAL; }
	h->colorsInit = true;
	return MCMCB200_OK;
}

int mcmcb200_get_colors_slice(mcmcb200_handle * h, uint32_t * out) {
	if (!h || !out) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	rc = ensure_stage(h); if (rc) return rc;
	if (h->nLocal) {
		const uint32_t blocks = (h->nLocal + 255) / 256;
		if (h->colBytes == 1) widen_colors_kernel<uint8_t><<<blocks, 256, 0, h->stream>>>((const uint8_t *)h->d_colors[s.sweep & 1] + h->vBegin, h->d_stage32, h->nLocal);
		else widen_colors_kernel<uint16_t><<<blocks, 256, 0, h->stream>>>((const uint16_t *)h->d_colors[s.sweep & 1] + h->vBegin, h->d_stage32, h->nLocal);
		h->launches++;
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(out, h->d_stage32, sizeof(uint32_t) * (size_t)h->nLocal, cudaMemcpyDeviceToHost, h->stream));
	}
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_set_tape(mcmcb200_handle * h, const float * u, uint32_t sweeps) {
	if (!h) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	cudaFree(h->d_tape); h->d_tape = nullptr; h->tapeSweeps = 0;
	if (!u || sweeps == 0) return MCMCB200_OK;
	const size_t bytes = sizeof(float) * (size_t)sweeps * h->nGlobal;
	CU(cudaMalloc(&h->d_tape, bytes));
	CU(cudaMemcpy(h->d_tape, u, bytes, cudaMemcpyHostToDevice));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	h->tapeSweeps = sweeps; h->tapeBase = s.sweep;
	return MCMCB200_OK;
}

int mcmcb200_sweep(mcmcb200_handle * h, uint32_t k) {
	NvtxRange nvtx_("mcmcb200_sweep");
	if (!h) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	const bool split = h->split();
	if (split && k != 1) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	h->tcActive = false;                                          // a sweep ends any distributed repair in progress
	CU(cudaEventRecord(h->ev0, h->stream));
	for (uint32_t i = 0; i < k; ++i) {
		if (h->tapeSweeps && (h->hostSweepUpper - h->tapeBase) >= h->tapeSweeps) {
			CU(cudaEventRecord(h->ev1, h->stream)); h->timed = true;
			return MCMCB200_ETAPE;
		}
		SweepArgs a = make_args(h);
		CU(launch_sweep(h, a));
		h->hostSweepUpper++;
	}
	CU(cudaEventRecord(h->ev1, h->stream));
	h->timed = true;
	h->pendingCountOnly = false;
	return MCMCB200_OK;
}

int mcmcb200_finalize_sweep(mcmcb200_handle * h) {
	NvtxRange nvtx_("mcmcb200_finalize_sweep (multi-GPU: after the counter all-reduce)");
	if (!h) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	SweepArgs a = make_args(h);
	a.countOnly = h->pendingCountOnly ? 1u : 0u;
	finalize_kernel<<<1, kThreads, 0, h->stream>>>(a);
	h->launches++;
	CU(cudaGetLastError());
	h->pendingCountOnly = false;
	return MCMCB200_OK;
}

int mcmcb200_status(mcmcb200_handle * h, mcmcb200_status_t * out) {
	NvtxRange nvtx_("mcmcb200_status");
	if (!h || !out) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	const bool split = h->split();
	if (s.countsSweep != s.sweep) {
		if (split) {
			// multi-GPU: launch the local counting pass; the caller all-reduces COUNTERS, calls finalize_sweep and asks again
			SweepArgs a = make_args(h);
			a.countOnly = 1; a.fuseFinalize = 0; a.tape = nullptr;
			CU(launch_sweep(h, a));
			h->pendingCountOnly = true;
			memset(out, 0, sizeof(*out));
			out->sweep = s.sweep; out->countsSweep = 0xffffffffu; out->z = s.z;
			return MCMCB200_OK;
		}
		rc = run_count_pass(h, nullptr, nullptr, nullptr, nullptr); if (rc) return rc;
		rc = read_state(h, &s); if (rc) return rc;
	}
	std::vector<unsigned long long> hist(h->p.nCol);
	CU(cudaMemcpyAsync(hist.data(), h->d_hist[s.sweep & 1], sizeof(unsigned long long) * h->p.nCol, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	uint32_t used = 0;
	for (unsigned long long c : hist) used += c != 0;
	out->sweep = s.sweep;
	out->conflictEdges = s.lastDirected >> 1;
	out->violatingVertices = s.lastViol;
	out->usedColors = used;
	out->countsSweep = s.countsSweep;
	out->z = s.z;
	const uint64_t metric = s.convergence == 0 ? s.lastViol : (s.lastDirected >> 1);
	out->converged = metric <= s.z ? 1 : 0;
	if (s.errorFlag) return MCMCB200_EINVAL;
	return MCMCB200_OK;
}

int mcmcb200_get_colors(mcmcb200_handle * h, uint32_t * out) {
	NvtxRange nvtx_("mcmcb200_get_colors");
	if (!h || !out) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	rc = ensure_stage(h); if (rc) return rc;
	const uint32_t n = h->nGlobal, blocks = (n + 255) / 256;
	if (h->colBytes == 1) widen_colors_kernel<uint8_t><<<blocks, 256, 0, h->stream>>>((const uint8_t *)h->d_colors[s.sweep & 1], h->d_stage32, n);
	else widen_colors_kernel<uint16_t><<<blocks, 256, 0, h->stream>>>((const uint16_t *)h->d_colors[s.sweep & 1], h->d_stage32, n);
	h->launches++;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out, h->d_stage32, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_get_class_sizes(mcmcb200_handle * h, uint64_t * out) {
	if (!h || !out) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	CU(cudaMemcpyAsync(out, h->d_hist[s.sweep & 1], sizeof(uint64_t) * h->p.nCol, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_get_history(mcmcb200_handle * h, uint64_t * out, uint32_t cap, uint32_t * count) {
	if (!h || !out || !count) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	uint32_t have = (s.countsSweep == 0xffffffffu) ? 0u : std::min<uint32_t>(s.countsSweep + 1u, h->historyCap);
	have = std::min(have, cap);
	if (have) CU(cudaMemcpy(out, h->d_history, sizeof(uint64_t) * 2 * have, cudaMemcpyDeviceToHost));
	*count = have;
	return MCMCB200_OK;
}

int mcmcb200_conflicts_of(mcmcb200_handle * h, const uint32_t * colors, uint64_t * edges, uint64_t * vertices) {
	if (!h || !colors) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	if (!h->d_colorsTmp) CU(cudaMalloc(&h->d_colorsTmp, ((size_t)h->nGlobal + kColorPad) * h->colBytes));
	CU(cudaStreamSynchronize(h->stream));
	uint32_t zero = 0;
	CU(cudaMemcpyAsync(&h->d_state->errorFlag, &zero, sizeof(zero), cudaMemcpyHostToDevice, h->stream));
	int rc = narrow_into(h, colors, h->d_colorsTmp); if (rc) return rc;
	rc = run_count_pass(h, h->d_colorsTmp, h->d_countOut, nullptr, nullptr); if (rc) return rc;
	unsigned long long res[2];
	CU(cudaMemcpyAsync(res, h->d_countOut, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
	DevState s;
	rc = read_state(h, &s); if (rc) return rc;
	if (s.errorFlag) {
		CU(cudaMemcpy(&h->d_state->errorFlag, &zero, sizeof(zero), cudaMemcpyHostToDevice));
		return MCMCB200_EINVAL;
	}
	const bool whole = h->vBegin == 0 && h->vEnd == h->nGlobal;
	if (edges) *edges = whole ? (res[0] >> 1) : res[0];   // a partition reports its directed count (sum over ranks = 2*edges)
	if (vertices) *vertices = res[1];
	return MCMCB200_OK;
}

int mcmcb200_debug_all_occupancy(mcmcb200_handle * h, uint64_t * masks, uint32_t * same) {
	if (!h || !masks || !same) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	unsigned long long * d_m = nullptr; uint32_t * d_s = nullptr;
	const size_t nm = (size_t)std::max<uint32_t>(h->nLocal, 1) * h->maskWords64;
	CU(cudaMalloc(&d_m, sizeof(unsigned long long) * nm));
	cudaError_t e = cudaMalloc(&d_s, sizeof(uint32_t) * std::max<uint32_t>(h->nLocal, 1));
	if (e != cudaSuccess) { cudaFree(d_m); return cuda_fail(e, "cudaMalloc(dbgSame)", __LINE__); }
	DevState s;
	int rc = read_state(h, &s);
	if (!rc) rc = run_count_pass(h, h->d_colors[s.sweep & 1], h->d_countOut, d_m, d_s);
	if (!rc) {
		e = cudaStreamSynchronize(h->stream);
		if (e == cudaSuccess) e = cudaMemcpy(masks, d_m, sizeof(unsigned long long) * (size_t)h->nLocal * h->maskWords64, cudaMemcpyDeviceToHost);
		if (e == cudaSuccess) e = cudaMemcpy(same, d_s, sizeof(uint32_t) * h->nLocal, cudaMemcpyDeviceToHost);
		if (e != cudaSuccess) rc = cuda_fail(e, "debug copy", __LINE__);
	}
	cudaFree(d_m); cudaFree(d_s);
	return rc;
}

int mcmcb200_debug_occupancy(mcmcb200_handle * h, uint32_t v, uint32_t * maskWords) {
	if (!h || !maskWords || v < h->vBegin || v >= h->vEnd) return MCMCB200_EINVAL;
	std::vector<uint64_t> masks((size_t)h->nLocal * h->maskWords64);
	std::vector<uint32_t> same(h->nLocal);
	int rc = mcmcb200_debug_all_occupancy(h, masks.data(), same.data()); if (rc) return rc;
	const uint32_t words = (h->p.nCol + 31) / 32;
	for (uint32_t w = 0; w < words; ++w) {
		const uint64_t m = masks[(size_t)(v - h->vBegin) * h->maskWords64 + (w >> 1)];
		maskWords[w] = (uint32_t)(m >> (32 * (w & 1)));
	}
	return MCMCB200_OK;
}

int mcmcb200_tailcut(mcmcb200_handle * h, uint32_t maxRounds, uint32_t * rounds) {
	NvtxRange nvtx_("mcmcb200_tailcut");
	if (!h) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	if (h->vBegin != 0 || h->vEnd != h->nGlobal) return MCMCB200_EUNSUPPORTED;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	void * cur = h->d_colors[s.sweep & 1];
	unsigned long long * hist = h->d_hist[s.sweep & 1];
	const uint32_t nCol = h->p.nCol, n = h->nLocal;
	// colours in ascending class size (coloringMCMC_main.cu:272-277); ties by colour index (contract)
	std::vector<unsigned long long> hh(nCol);
	CU(cudaMemcpy(hh.data(), hist, sizeof(unsigned long long) * nCol, cudaMemcpyDeviceToHost));
	std::vector<uint32_t> order(nCol);
	for (uint32_t i = 0; i < nCol; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hh[a] < hh[b]; });
	uint32_t usedByList = 0;
	if (h->d_violList[0]) {
		// ---- list-driven path (params.tailcut): the sweeps emitted the violating vertices; no rescan of the graph, no host
		//      round trip per round.  Falls through to the full-scan path below when the list is not available. ----
		uint32_t done = 0;
		int lrc = tailcut_from_list(h, s, cur, hist, order, maxRounds, &done);
		if (lrc != 1) { if (rounds) *rounds = done; return lrc; }  // 1 = list unavailable / overflowed / no longer exact
		usedByList = done;
	}
	uint32_t * d_order = nullptr, * d_list = nullptr, * d_heavy = nullptr, * d_counters = nullptr; uint8_t * d_pending = nullptr, * d_ready = nullptr;
	auto cleanup = [&]() { cudaFree(d_order); cudaFree(d_list); cudaFree(d_heavy); cudaFree(d_counters); cudaFree(d_pending); cudaFree(d_ready); };
	cudaError_t e = cudaMalloc(&d_order, sizeof(uint32_t) * nCol);
	if (e == cudaSuccess) e = cudaMalloc(&d_list, sizeof(uint32_t) * std::max<uint32_t>(n, 1));
	if (e == cudaSuccess) e = cudaMalloc(&d_heavy, sizeof(uint32_t) * ((size_t)n + 1));
	if (e == cudaSuccess) e = cudaMalloc(&d_counters, sizeof(uint32_t) * 3);
	if (e == cudaSuccess) e = cudaMalloc(&d_pending, std::max<uint32_t>(n, 1));
	if (e == cudaSuccess) e = cudaMalloc(&d_ready, std::max<uint32_t>(n, 1));
	if (e == cudaSuccess) e = cudaMemcpy(d_order, order.data(), sizeof(uint32_t) * nCol, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "tailcut workspace", __LINE__); }
	uint32_t used = usedByList;
	for (; used < maxRounds; ++used) {                                     // while (conflictCounter > 0), _main.cu:279
		uint32_t flagged = 0, changed = 0;
		const int ce = launch_tailcut_pass(h->stream, h->colBytes, h->d_rowptr, h->d_neighs, n, nCol, cur, hist, d_order,
		                                   d_pending, d_ready, d_list, d_heavy, d_counters, &flagged, &changed, &h->launches);
		if (ce) { cleanup(); return cuda_fail((cudaError_t)ce, "tailcut pass", __LINE__); }
		if (flagged == 0) break;                                           // no conflicting edge left
		if (changed == 0) { ++used; break; }                               // no progress (every colour taken around the vertices that are left): the
		                                                                   // reference would repeat this pass for ever (_main.cu:279)
	}
	cleanup();
	// the colouring changed under the counters: force a recount at the next status
	const uint32_t stale = 0xffffffffu; const int32_t notConv = -1;
	CU(cudaMemcpyAsync(&h->d_state->countsSweep, &stale, sizeof(stale), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(&h->d_state->convergedAt, &notConv, sizeof(notConv), cudaMemcpyHostToDevice, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	if (rounds) *rounds = used;
	return MCMCB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Distributed tail cutting: the repair of a multi-GPU chain (the reference is single-GPU; same sequential-greedy result as
// mcmcb200_tailcut / coloringMCMC_utils.cu:73-101).  Every rank repairs the violators it OWNS, in globally synchronised rounds:
// a flagged vertex is ready when no flagged neighbour with a smaller id -- on any rank -- is still pending, ready vertices are
// pairwise non-adjacent, and after each round the ranks tell each other the (vertex, colour) pairs they decided.  The caller
// (multigpu.py DistributedSweeper.tailcut) moves the small lists between the ranks; nothing here talks to another GPU.
// ---------------------------------------------------------------------------------------------------------------
namespace {
int tc_dist_ready(mcmcb200_handle * h, uint32_t ioCap) {
	if (!h->d_violList[0]) return MCMCB200_ESTATE;                // handle was not created with params.tailcut
	if (!h->d_tcOrder) CU(cudaMalloc(&h->d_tcOrder, sizeof(uint32_t) * h->p.nCol));
	if (!h->d_tcHeavy) CU(cudaMalloc(&h->d_tcHeavy, sizeof(uint32_t) * ((size_t)h->violCap + 1)));
	if (!h->d_tcReady) CU(cudaMalloc(&h->d_tcReady, std::max<uint32_t>(h->violCap, 1)));
	if (!h->d_tcCounters) CU(cudaMalloc(&h->d_tcCounters, sizeof(uint32_t) * 4));
	if (h->tcIoCap < ioCap) {
		cudaFree(h->d_tcIo); h->d_tcIo = nullptr; h->tcIoCap = 0;
		CU(cudaMalloc(&h->d_tcIo, sizeof(uint32_t) * 2 * (size_t)ioCap));
		h->tcIoCap = ioCap;
	}
	return MCMCB200_OK;
}
// rowptr indexed by GLOBAL vertex id (the repair kernels address rows by the ids in the lists; only ids in [vBegin, vEnd) -- the
// vertices this rank listed itself -- are ever used as a row index, so every access stays inside the allocation)
const uint32_t * tc_rowptr_global(const mcmcb200_handle * h) {
	return reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(h->d_rowptr) - sizeof(uint32_t) * (uintptr_t)h->vBegin);
}
} // namespace

int mcmcb200_tailcut_dist_begin(mcmcb200_handle * h, const uint32_t * order, uint32_t * outFlagged, uint32_t cap, uint32_t * count) {
	if (!h || !order || !outFlagged || !count) return MCMCB200_EINVAL;
	if (!h->colorsInit) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	int rc = tc_dist_ready(h, std::max<uint32_t>(h->violCap, 1)); if (rc) return rc;
	DevState s;
	rc = read_state(h, &s); if (rc) return rc;
	if (!h->tcActive) {                                           // first pass: the list the converging sweep emitted
		if (s.violListSweep != s.sweep || s.violListCount > h->violCap) return MCMCB200_ESTATE;
		h->tcSrc = 0; h->tcListCount = s.violListCount; h->tcActive = true;
	}
	CU(cudaMemcpyAsync(h->d_tcOrder, order, sizeof(uint32_t) * h->p.nCol, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemsetAsync(h->d_tcCnt, 0, sizeof(TailcutCounters), h->stream));
	void * cur = h->d_colors[s.sweep & 1];
	const uint32_t lb = (uint32_t)(((uint64_t)h->tcListCount * 32u + 255u) / 256u);
	if (h->tcListCount) {
		if (h->colBytes == 1) tc_filter_kernel<uint8_t><<<lb, 256, 0, h->stream>>>(tc_rowptr_global(h), h->d_neighs, (const uint8_t *)cur, h->d_violList[h->tcSrc], h->tcListCount, h->d_pending, h->d_flist, h->d_tcCnt);
		else tc_filter_kernel<uint16_t><<<lb, 256, 0, h->stream>>>(tc_rowptr_global(h), h->d_neighs, (const uint16_t *)cur, h->d_violList[h->tcSrc], h->tcListCount, h->d_pending, h->d_flist, h->d_tcCnt);
		h->launches++;
		CU(cudaGetLastError());
	}
	TailcutCounters c{};
	CU(cudaMemcpyAsync(&c, h->d_tcCnt, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	h->tcFlagged = c.flagged;
	if (c.flagged > cap) return MCMCB200_EINVAL;
	CU(cudaMemcpy(outFlagged, h->d_flist, sizeof(uint32_t) * (size_t)c.flagged, cudaMemcpyDeviceToHost));
	*count = c.flagged;
	return MCMCB200_OK;
}

int mcmcb200_tailcut_dist_mark(mcmcb200_handle * h, const uint32_t * ids, uint32_t count) {
	if (!h || (count && !ids)) return MCMCB200_EINVAL;
	if (!h->tcActive) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	if (!count) return MCMCB200_OK;
	for (uint32_t i = 0; i < count; ++i) if (ids[i] >= h->nGlobal) return MCMCB200_EINVAL;
	int rc = tc_dist_ready(h, std::max<uint32_t>(count, h->tcIoCap)); if (rc) return rc;
	CU(cudaMemcpyAsync(h->d_tcIo, ids, sizeof(uint32_t) * (size_t)count, cudaMemcpyHostToDevice, h->stream));
	tc_mark_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(h->d_tcIo, count, h->d_pending);
	h->launches++;
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_tailcut_dist_round(mcmcb200_handle * h, uint32_t * outIds, uint32_t * outCols, uint32_t cap, uint32_t * processed, uint32_t * remaining,
                                uint32_t * inexact) {
	if (!h || !outIds || !outCols || !processed || !remaining || !inexact) return MCMCB200_EINVAL;
	if (!h->tcActive) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	*processed = 0; *remaining = 0; *inexact = 0;
	if (h->tcFlagged == 0) return MCMCB200_OK;
	int rc = tc_dist_ready(h, std::max<uint32_t>(h->tcFlagged, h->tcIoCap)); if (rc) return rc;
	DevState s;
	rc = read_state(h, &s); if (rc) return rc;
	void * cur = h->d_colors[s.sweep & 1];
	unsigned long long * hist = h->d_hist[s.sweep & 1];
	const uint32_t n = h->tcFlagged, nCol = h->p.nCol;
	const uint32_t lb = (uint32_t)(((uint64_t)n * 32u + 255u) / 256u);
	uint32_t * d_out = h->d_tcIo, * d_outCols = h->d_tcIo + h->tcIoCap;
	CU(cudaMemsetAsync(h->d_tcCounters, 0, 4 * sizeof(uint32_t), h->stream));      // [0] remaining [1] changed [2] processed [3] inexact
	CU(cudaMemsetAsync(h->d_tcHeavy, 0, sizeof(uint32_t), h->stream));
	const uint32_t * rp = tc_rowptr_global(h);
	tailcut_ready_kernel<<<lb, 256, 0, h->stream>>>(rp, h->d_neighs, h->d_pending, h->d_flist, n, h->d_tcReady);
	const size_t smem = tc_smem_bytes(nCol);
	if (h->colBytes == 1) {
		tailcut_apply_kernel<uint8_t><<<lb, 256, 0, h->stream>>>(rp, h->d_neighs, nCol, (uint8_t *)cur, h->d_pending, h->d_flist, n, h->d_tcReady, h->d_tcOrder, hist,
		                                                          h->d_tcCounters + 0, h->d_tcHeavy, h->d_tcCounters + 1, d_out, d_outCols, h->d_tcCounters + 2, h->d_tcCounters + 3);
		tailcut_apply_heavy_kernel<uint8_t><<<64, kTcThreads, smem, h->stream>>>(rp, h->d_neighs, nCol, (uint8_t *)cur, h->d_pending, h->d_tcOrder, hist, h->d_tcHeavy,
		                                                                         h->d_tcCounters + 1, d_out, d_outCols, h->d_tcCounters + 2, h->d_tcCounters + 3);
	} else {
		tailcut_apply_kernel<uint16_t><<<lb, 256, 0, h->stream>>>(rp, h->d_neighs, nCol, (uint16_t *)cur, h->d_pending, h->d_flist, n, h->d_tcReady, h->d_tcOrder, hist,
		                                                           h->d_tcCounters + 0, h->d_tcHeavy, h->d_tcCounters + 1, d_out, d_outCols, h->d_tcCounters + 2, h->d_tcCounters + 3);
		tailcut_apply_heavy_kernel<uint16_t><<<64, kTcThreads, smem, h->stream>>>(rp, h->d_neighs, nCol, (uint16_t *)cur, h->d_pending, h->d_tcOrder, hist, h->d_tcHeavy,
		                                                                          h->d_tcCounters + 1, d_out, d_outCols, h->d_tcCounters + 2, h->d_tcCounters + 3);
	}
	h->launches += 3;
	CU(cudaGetLastError());
	uint32_t cnt[4];
	CU(cudaMemcpyAsync(cnt, h->d_tcCounters, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	if (cnt[2] > cap) return MCMCB200_EINVAL;
	CU(cudaMemcpy(outIds, d_out, sizeof(uint32_t) * (size_t)cnt[2], cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(outCols, d_outCols, sizeof(uint32_t) * (size_t)cnt[2], cudaMemcpyDeviceToHost));
	*processed = cnt[2]; *remaining = cnt[0]; *inexact = cnt[3];
	return MCMCB200_OK;
}

int mcmcb200_tailcut_dist_apply(mcmcb200_handle * h, const uint32_t * ids, const uint32_t * cols, uint32_t count) {
	if (!h || (count && (!ids || !cols))) return MCMCB200_EINVAL;
	if (!h->tcActive) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	if (!count) return MCMCB200_OK;
	for (uint32_t i = 0; i < count; ++i) if (ids[i] >= h->nGlobal || cols[i] >= h->p.nCol) return MCMCB200_EINVAL;
	int rc = tc_dist_ready(h, std::max<uint32_t>(count, h->tcIoCap)); if (rc) return rc;
	DevState s;
	rc = read_state(h, &s); if (rc) return rc;
	CU(cudaMemcpyAsync(h->d_tcIo, ids, sizeof(uint32_t) * (size_t)count, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(h->d_tcIo + h->tcIoCap, cols, sizeof(uint32_t) * (size_t)count, cudaMemcpyHostToDevice, h->stream));
	if (h->colBytes == 1) tc_remote_kernel<uint8_t><<<(count + 255) / 256, 256, 0, h->stream>>>(h->d_tcIo, h->d_tcIo + h->tcIoCap, count, (uint8_t *)h->d_colors[s.sweep & 1], h->d_pending, h->d_hist[s.sweep & 1]);
	else tc_remote_kernel<uint16_t><<<(count + 255) / 256, 256, 0, h->stream>>>(h->d_tcIo, h->d_tcIo + h->tcIoCap, count, (uint16_t *)h->d_colors[s.sweep & 1], h->d_pending, h->d_hist[s.sweep & 1]);
	h->launches++;
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_tailcut_dist_recount(mcmcb200_handle * h, uint64_t * directedLocal, uint64_t * violLocal, uint32_t * nextFlagged) {
	if (!h || !directedLocal || !violLocal || !nextFlagged) return MCMCB200_EINVAL;
	if (!h->tcActive) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	void * cur = h->d_colors[s.sweep & 1];
	CU(cudaMemsetAsync(h->d_tcCnt, 0, sizeof(TailcutCounters), h->stream));
	if (h->tcListCount) {
		const uint32_t lb = (uint32_t)(((uint64_t)h->tcListCount * 32u + 255u) / 256u);
		if (h->colBytes == 1) tc_recount_kernel<uint8_t><<<lb, 256, 0, h->stream>>>(tc_rowptr_global(h), h->d_neighs, (const uint8_t *)cur, h->d_violList[h->tcSrc], h->tcListCount, h->d_pending, h->d_violList[h->tcSrc ^ 1], h->d_tcCnt);
		else tc_recount_kernel<uint16_t><<<lb, 256, 0, h->stream>>>(tc_rowptr_global(h), h->d_neighs, (const uint16_t *)cur, h->d_violList[h->tcSrc], h->tcListCount, h->d_pending, h->d_violList[h->tcSrc ^ 1], h->d_tcCnt);
		h->launches++;
		CU(cudaGetLastError());
	}
	TailcutCounters c{};
	CU(cudaMemcpyAsync(&c, h->d_tcCnt, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	h->tcSrc ^= 1; h->tcListCount = c.nextCount;
	*directedLocal = c.directed; *violLocal = c.viol; *nextFlagged = c.nextFlagged;
	return MCMCB200_OK;
}

int mcmcb200_tailcut_dist_end(mcmcb200_handle * h, uint64_t directedGlobal, uint64_t violGlobal, uint32_t exact) {
	if (!h) return MCMCB200_EINVAL;
	if (!h->tcActive) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	h->tcActive = false;
	if (h->tcSrc == 1) CU(cudaMemcpyAsync(h->d_violList[0], h->d_violList[1], sizeof(uint32_t) * (size_t)h->tcListCount, cudaMemcpyDeviceToDevice, h->stream));
	h->tcSrc = 0;
	if (exact) {
		tc_commit_global_kernel<<<1, 1, 0, h->stream>>>(h->d_state, directedGlobal, violGlobal, h->tcListCount);
		h->launches++;
		CU(cudaGetLastError());
	} else {                                                       // a repaired vertex found every colour taken: recount with a full pass at the next status
		const uint32_t stale = 0xffffffffu; const int32_t notConv = -1;
		CU(cudaMemcpyAsync(&h->d_state->countsSweep, &stale, sizeof(stale), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(&h->d_state->convergedAt, &notConv, sizeof(notConv), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(&h->d_state->violListSweep, &stale, sizeof(stale), cudaMemcpyHostToDevice, h->stream));
	}
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_luby_color(uint32_t n, uint64_t nnz, const uint32_t * cumulDegs, const uint32_t * neighs, uint64_t seed, int32_t device,
                        uint32_t * colorsOut, uint32_t * numColors, uint32_t * rounds) {
	NvtxRange nvtx_("mcmcb200_luby_color");
	if (!cumulDegs || (nnz && !neighs) || !colorsOut || n == 0) return MCMCB200_EINVAL;
	// the same host-side CSR validation as mcmcb200_create: monotone offsets starting at 0, nnz consistent, ids < n
	if (cumulDegs[0] != 0 || cumulDegs[n] != nnz || nnz >= 0xfffffff0ull) return MCMCB200_EINVAL;
	for (uint32_t i = 0; i < n; ++i) if (cumulDegs[i + 1] < cumulDegs[i]) return MCMCB200_EINVAL;
	for (uint64_t e2 = 0; e2 < nnz; ++e2) if (neighs[e2] >= n) return MCMCB200_EINVAL;
	mcmcb200_params p{}; p.device = device;
	int dev = 0, sms = 0;
	int rc = select_device(&p, &dev, &sms); if (rc) return rc;
	uint32_t * d_rp = nullptr, * d_nb = nullptr, * d_col = nullptr, * d_flag = nullptr;
	uint8_t * d_c = nullptr, * d_is = nullptr, * d_ch = nullptr, * d_keep = nullptr;
	cudaStream_t st = nullptr;
	CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	auto cleanup = [&]() { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); st = nullptr; } cudaFree(d_rp); cudaFree(d_nb); cudaFree(d_col); cudaFree(d_flag); cudaFree(d_c); cudaFree(d_is); cudaFree(d_ch); cudaFree(d_keep); };
	cudaError_t e = cudaMalloc(&d_rp, sizeof(uint32_t) * ((size_t)n + 1));
	if (e == cudaSuccess) e = cudaMalloc(&d_nb, sizeof(uint32_t) * std::max<uint64_t>(nnz, 1));
	if (e == cudaSuccess) e = cudaMalloc(&d_col, sizeof(uint32_t) * (size_t)n);
	if (e == cudaSuccess) e = cudaMalloc(&d_flag, sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMalloc(&d_c, n);
	if (e == cudaSuccess) e = cudaMalloc(&d_is, n);
	if (e == cudaSuccess) e = cudaMalloc(&d_ch, n);
	if (e == cudaSuccess) e = cudaMalloc(&d_keep, n);
	if (e == cudaSuccess) e = cudaMemcpy(d_rp, cumulDegs, sizeof(uint32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice);
	if (e == cudaSuccess && nnz) e = cudaMemcpy(d_nb, neighs, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemsetAsync(d_col, 0, sizeof(uint32_t) * (size_t)n, st);
	if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "luby setup", __LINE__); }
	const uint32_t blocks = (n + 127) / 128;                      // block 128 like the reference (coloringLuby.cu)
	uint32_t color = 0, round = 0, flag = 1;
	while (flag) {                                                // CICLO_1: one colour per iteration (coloringLuby.cu:389-470)
		color++;
		luby_prune_kernel<<<blocks, 128, 0, st>>>(n, d_col, d_c, d_is);
		uint32_t left = 1;
		while (left) {                                            // CICLO_2: grow the independent set until no candidate is left
			round++;
			luby_choose_kernel<<<blocks, 128, 0, st>>>(n, seed, round, d_c, d_ch);
			luby_resolve_kernel<<<blocks, 128, 0, st>>>(n, d_rp, d_nb, d_ch, d_keep);
			luby_update_kernel<<<blocks, 128, 0, st>>>(n, d_rp, d_nb, d_keep, d_c, d_is, d_flag);
			cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), st);
			luby_left_kernel<<<blocks, 128, 0, st>>>(n, d_c, d_flag);
			e = cudaMemcpyAsync(&left, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);
			if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "luby round", __LINE__); }
		}
		cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), st);
		luby_color_kernel<<<blocks, 128, 0, st>>>(n, color, d_is, d_col, d_flag);
		e = cudaMemcpyAsync(&flag, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "luby colour", __LINE__); }
	}
	e = cudaStreamSynchronize(st);
	if (e == cudaSuccess) e = cudaMemcpy(colorsOut, d_col, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost);
	cleanup();
	if (e != cudaSuccess) return cuda_fail(e, "luby copy", __LINE__);
	if (numColors) *numColors = color;
	if (rounds) *rounds = round;
	return MCMCB200_OK;
}

int mcmcb200_device_view(mcmcb200_handle * h, int which, void ** devPtr, uint64_t * bytes, uint32_t * elemBytes) {
	if (!h || !devPtr) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	DevState s;
	int rc = read_state(h, &s); if (rc) return rc;
	const uint64_t colBytesTotal = ((uint64_t)h->nGlobal + kColorPad) * h->colBytes;
	switch (which) {
	case MCMCB200_VIEW_COLORS_CUR:  *devPtr = h->d_colors[s.sweep & 1]; if (bytes) *bytes = colBytesTotal; if (elemBytes) *elemBytes = h->colBytes; break;
	case MCMCB200_VIEW_COLORS_NEXT: *devPtr = h->d_colors[(s.sweep + 1) & 1]; if (bytes) *bytes = colBytesTotal; if (elemBytes) *elemBytes = h->colBytes; break;
	case MCMCB200_VIEW_COUNTERS:    *devPtr = h->d_scratch; if (bytes) *bytes = sizeof(unsigned long long) * (h->p.nCol + 2); if (elemBytes) *elemBytes = 8; break;
	default: return MCMCB200_EINVAL;
	}
	return MCMCB200_OK;
}

int mcmcb200_ipc_export(mcmcb200_handle * h, unsigned char * handles /* [3][64] */) {
	if (!h || !handles) return MCMCB200_EINVAL;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	if (!h->bl.valid) return MCMCB200_EUNSUPPORTED;          // the fused exchange lives in the source-blocked sweep
	CU(cudaSetDevice(h->device));
	void * ptrs[3] = {h->d_colors[0], h->d_colors[1], h->d_xchg};
	for (int b = 0; b < 3; ++b) {
		cudaIpcMemHandle_t mh;
		CU(cudaIpcGetMemHandle(&mh, ptrs[b]));
		memcpy(handles + 64 * b, &mh, 64);
	}
	return MCMCB200_OK;
}

int mcmcb200_ipc_attach(mcmcb200_handle * h, uint32_t nRanks, uint32_t myRank, const unsigned char * handles /* [nRanks][3][64] */) {
	NvtxRange nvtx_("mcmcb200_ipc_attach (fused exchange set-up)");
	if (!h || !handles || nRanks < 2 || nRanks > (uint32_t)kMaxPeers || myRank >= nRanks) return MCMCB200_EINVAL;
	if (!h->bl.valid) return MCMCB200_EUNSUPPORTED;
	if (h->nPeers) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	void * opened[kMaxPeers][3] = {};
	for (uint32_t q = 0; q < nRanks; ++q) {
		if (q == myRank) { opened[q][0] = h->d_colors[0]; opened[q][1] = h->d_colors[1]; opened[q][2] = h->d_xchg; continue; }
		for (int b = 0; b < 3; ++b) {
			cudaIpcMemHandle_t mh;
			memcpy(&mh, handles + ((size_t)q * 3 + b) * 64, 64);
			cudaError_t e = cudaIpcOpenMemHandle(&opened[q][b], mh, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) {                               // close what was opened so far: nothing half attached
				for (uint32_t q2 = 0; q2 <= q; ++q2)
					if (q2 != myRank) for (int b2 = 0; b2 < 3; ++b2) if (opened[q2][b2] && !(q2 == q && b2 >= b)) cudaIpcCloseMemHandle(opened[q2][b2]);
				return cuda_fail(e, "cudaIpcOpenMemHandle", __LINE__);
			}
		}
	}
	for (uint32_t q = 0; q < nRanks; ++q) {
		h->peerColors[0][q] = opened[q][0]; h->peerColors[1][q] = opened[q][1];
		h->peerXchg[q] = static_cast<unsigned long long *>(opened[q][2]);
	}
	h->nPeers = nRanks; h->myPeerIndex = myRank;
	return MCMCB200_OK;
}

int mcmcb200_ipc_detach(mcmcb200_handle * h) {
	if (!h) return MCMCB200_EINVAL;
	if (!h->nPeers) return MCMCB200_OK;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	for (uint32_t q = 0; q < h->nPeers; ++q) {
		if (q != h->myPeerIndex) {
			for (int b = 0; b < 2; ++b) if (h->peerColors[b][q]) cudaIpcCloseMemHandle(h->peerColors[b][q]);
			if (h->peerXchg[q]) cudaIpcCloseMemHandle(h->peerXchg[q]);
		}
		h->peerColors[0][q] = h->peerColors[1][q] = nullptr; h->peerXchg[q] = nullptr;
	}
	h->nPeers = 0;
	cudaGetLastError();
	return MCMCB200_OK;
}

int mcmcb200_stream(mcmcb200_handle * h, void ** cudaStreamOut) {
	if (!h || !cudaStreamOut) return MCMCB200_EINVAL;
	*cudaStreamOut = (void *)h->stream;
	return MCMCB200_OK;
}

int mcmcb200_synchronize(mcmcb200_handle * h) {
	if (!h) return MCMCB200_EINVAL;
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	return MCMCB200_OK;
}

int mcmcb200_last_sweep_ms(mcmcb200_handle * h, float * ms) {
	if (!h || !ms) return MCMCB200_EINVAL;
	if (!h->timed) return MCMCB200_ESTATE;
	CU(cudaSetDevice(h->device));
	CU(cudaEventSynchronize(h->ev1));
	CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
	return MCMCB200_OK;
}

int mcmcb200_kernel_mode(mcmcb200_handle * h, int * mode) {
	if (!h || !mode) return MCMCB200_EINVAL;
	*mode = h->bl.valid ? (h->overlap ? MCMCB200_MODE_BLOCKED_OVERLAPPED : MCMCB200_MODE_BLOCKED)
	                    : h->wide ? MCMCB200_MODE_WIDE_BINNED : h->bn.valid ? MCMCB200_MODE_DIRECT_BINNED : MCMCB200_MODE_DIRECT;
	return MCMCB200_OK;
}

int mcmcb200_layout_bytes(mcmcb200_handle * h, uint64_t * bytes) {
	if (!h || !bytes) return MCMCB200_EINVAL;
	uint64_t b = 0;
	if (h->bl.valid) b += h->bl.bytes;
	if (h->bn.valid) b += sizeof(uint32_t) * ((uint64_t)h->bn.n[0] + h->bn.n[1] + h->bn.n[2] + 3);
	if (h->wide) b += (uint64_t)h->nGlobal + kColorPad + sizeof(uint32_t) * ((uint64_t)h->nLocal + 8 + h->bn.n[2] + 1) + 2 * sizeof(float) * ((uint64_t)h->p.nCol + 1);
	*bytes = b;
	return MCMCB200_OK;
}

int mcmcb200_launch_count(mcmcb200_handle * h, uint64_t * launches) {
	if (!h || !launches) return MCMCB200_EINVAL;
	*launches = h->launches;
	return MCMCB200_OK;
}

} // extern "C"
