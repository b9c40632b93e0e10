// TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// Harness around the UNMODIFIED reference GPU MCMC colourer (ColoringMCMC<float,float>), compiled for sm_100a from the
// sources where they lie under /root/reference/src by oracle/Makefile (target `refgpu`); nothing is copied.  The
// reference translation units it links:
//     graph_coloring/coloringMCMC_{main,balance,standard,decrease,utils,prints}.cu, graph_coloring/colorer.cpp,
//     graph/graphGPU.cu (+ graph/graphCPU.cpp textually, like main.cu:10-11), GPUutils/GPURandomizer.cu, utils/*.cpp
// It exists because the shipped GPU variant (COLOR_BALANCE_DYNAMIC_DISTR, coloringMCMC.h:39) and tailCutting have no CPU
// twin in the reference: the only way to pin the DYNAMIC proposal and the tail cut of libmcmcb200 to the reference is to
// run the reference's own kernels.  Used by tests/test_gpu_refgpu.py (-m gpu) only.
//
// What the harness adds (no reference source is modified):
//   * struct RefGPU : ColoringMCMC<float,float>: the data members are `protected` (coloringMCMC.h:52-139), so a derived
//     class can set/read coloring_d / taboo_d and drive ONE loop body of run() (coloringMCMC_main.cu:174,211-221,263-265)
//     with the reference's own launch shapes;
//   * draw capture: before a step the per-vertex curandState array is copied and curand_uniform is evaluated on the COPY
//     -- the value selectStarColoringBalanceDynamic's `curand_uniform(&states[idx])` (coloringMCMC_balance.cu:120) will
//     see if vertex idx draws in this step.  That array is the tape handed to mcmcb200_set_tape;
//   * taboo early return: the shipped kernel leaves starColoring_d[idx] untouched (coloringMCMC_balance.cu:84-90), i.e.
//     stale from two sweeps ago after the pointer swap; the STANDARD kernel writes coloring_d[idx] (coloringMCMC_standard.cu:
//     14-20).  The contract follows STANDARD; `prefill` = 1 emulates it on the reference side by copying
//     coloring_d into starColoring_d BEFORE the unmodified kernel runs.  With tabooIteration == 0 (the CLI default) the two
//     are identical.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include <algorithm>

#include "graph/graph.h"
#include "graph/graphCPU.cpp"          // template bodies live in the .cpp (main.cu:10 does the same)
#include "graph_coloring/coloring.h"
#include "graph_coloring/coloringMCMC.h"
#include "GPUutils/GPURandomizer.h"

bool g_traceLogEn = false;             // declared in utils/miscUtils.h
template class Graph<float, float>;

namespace {

__global__ void peek_uniform_kernel(curandState * copy, uint32_t n, float * out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = curand_uniform(&copy[i]);
}

struct RefGPU : public ColoringMCMC<float, float> {
	RefGPU(Graph<float, float> * g_d, curandState * st, ColoringMCMCParams p) : ColoringMCMC<float, float>(g_d, st, p) {
		// run() uploads the identity permutation (coloringMCMC_main.cu:129-132); the kernel reads it (_balance.cu:106)
		for (uint32_t i = 0; i < param.nCol; i++) orderedIndex_h[i] = i;
		cudaMemcpy(orderedIndex_d, orderedIndex_h, param.nCol * sizeof(uint32_t), cudaMemcpyHostToDevice);
		cudaMemset(coloring_d, 0, nnodes * sizeof(uint32_t));
		cudaMemset(starColoring_d, 0, nnodes * sizeof(uint32_t));
		cudaMemset(taboo_d, 0, nnodes * sizeof(uint32_t));
	}
	uint32_t n() const { return nnodes; }
	uint32_t * colors() { return coloring_d; }
	uint32_t * taboo() { return taboo_d; }
	curandState * states() { return randStates; }

	// one loop body of run() for the shipped configuration: coloringMCMC_main.cu:174, 211-221, 263-265
	int step_dynamic(int prefill) {
		cudaMemset(colorsChecker_d, 0, (size_t)nnodes * param.nCol * sizeof(bool));                     // :174
		if (prefill) cudaMemcpy(starColoring_d, coloring_d, nnodes * sizeof(uint32_t), cudaMemcpyDeviceToDevice);
		cudaMemcpy(coloring_h, coloring_d, nnodes * sizeof(uint32_t), cudaMemcpyDeviceToHost);          // :211
		memset(statsColors_h, 0, nnodes * sizeof(uint32_t));                                            // :212
		for (uint32_t i = 0; i < nnodes; i++) statsColors_h[coloring_h[i]]++;                           // :213
		cudaMemcpy(statsColors_d, statsColors_h, param.nCol * sizeof(uint32_t), cudaMemcpyHostToDevice); // :214
		ColoringMCMC_k::genDynamicDistribution<<<blocksPerGrid_nCol, threadsPerBlock>>>(probDistributionDynamic_d, param.nCol, nnodes, statsColors_d);   // :216
		ColoringMCMC_k::selectStarColoringBalanceDynamic<<<blocksPerGrid, threadsPerBlock>>>(nnodes, starColoring_d, qStar_d, param.nCol, coloring_d,
			graphStruct_d->cumulDegs, graphStruct_d->neighs, colorsChecker_d, taboo_d, param.tabooIteration, probDistributionDynamic_d, orderedIndex_d,
			randStates, param.lambda, param.epsilon, statsColors_d);                                                                                     // :218
		cudaError_t e = cudaDeviceSynchronize();                                                        // :221
		switchPointer = coloring_d; coloring_d = starColoring_d; starColoring_d = switchPointer;        // :263-265
		return (int)e;
	}

	// Conflicting-EDGE count of the current colouring: the reference's conflictCounter kernel (coloringMCMC_utils.cu:103-119)
	// followed by a HOST sum of its per-vertex counts.  calcConflicts (:184-198) finishes with sumReduction, which adds the
	// uint32 counts as denormal floats, reads up to 127 words past nnodes and lets block b overwrite word b while other
	// blocks may still read it: same integer when it works, not a sound checker -- so the harness sums on the host.
	int conflicts() {
		ColoringMCMC_k::conflictCounter<<<blocksPerGrid, threadsPerBlock>>>(nnodes, conflictCounter_d, coloring_d, graphStruct_d->cumulDegs, graphStruct_d->neighs);
		cudaMemcpy(conflictCounter_h, conflictCounter_d, nnodes * sizeof(uint32_t), cudaMemcpyDeviceToHost);
		long long c = 0;
		for (uint32_t i = 0; i < nnodes; i++) c += conflictCounter_h[i];
		return (int)c;
	}

	// tail cutting, coloringMCMC_main.cu:271-290, without the log prints
	int tailcut(int maxRounds, int * roundsOut) {
		cudaMemcpy(coloring_h, coloring_d, nnodes * sizeof(uint32_t), cudaMemcpyDeviceToHost);          // :272
		memset(statsColors_h, 0, nnodes * sizeof(uint32_t));
		for (uint32_t i = 0; i < nnodes; i++) statsColors_h[coloring_h[i]]++;
		for (uint32_t i = 0; i < param.nCol; i++) orderedIndex_h[i] = i;
		// the reference uses std::sort (:276), which leaves the order of equal class sizes to the library; the contract
		// fixes ties by ascending colour index == std::stable_sort.  Same comparator.
		std::stable_sort(&orderedIndex_h[0], &orderedIndex_h[param.nCol], [&](int i, int j) { return statsColors_h[i] < statsColors_h[j]; });
		cudaMemcpy(orderedIndex_d, orderedIndex_h, param.nCol * sizeof(uint32_t), cudaMemcpyHostToDevice);
		conflictCounter = conflicts();
		int rounds = 0;
		while (conflictCounter > 0 && rounds < maxRounds) {                                              // :279
			ColoringMCMC_k::conflictCounter<<<blocksPerGrid, threadsPerBlock>>>(nnodes, conflictCounter_d, coloring_d, graphStruct_d->cumulDegs, graphStruct_d->neighs);
			cudaMemset(colorsChecker_d, 0, (size_t)nnodes * param.nCol * sizeof(bool));
			ColoringMCMC_k::tailCutting<<<1, 1>>>(nnodes, param.nCol, coloring_d, graphStruct_d->cumulDegs, graphStruct_d->neighs, colorsChecker_d,
				conflictCounter, conflictCounter_d, orderedIndex_d);
			conflictCounter = conflicts();
			rounds++;
		}
		// restore the identity permutation the DYNAMIC kernel expects
		for (uint32_t i = 0; i < param.nCol; i++) orderedIndex_h[i] = i;
		cudaMemcpy(orderedIndex_d, orderedIndex_h, param.nCol * sizeof(uint32_t), cudaMemcpyHostToDevice);
		if (roundsOut) *roundsOut = rounds;
		return conflictCounter;
	}
	uint32_t ripCount() const { return rip; }
	bool hitMax() const { return maxIterReached; }
};

struct Handle {
	Graph<float, float> * g_h = nullptr;
	Graph<float, float> * g_d = nullptr;
	GPURand * rnd = nullptr;
	RefGPU * col = nullptr;
};

} // namespace

extern "C" {

// Graph from a CSR (host) -> Graph(Graph*) device copy (graphGPU.cu:210-226) -> GPURand(n, seed) (GPURandomizer.cu:85-96)
// -> ColoringMCMC<float,float>(graph_d, randStates, params) (coloringMCMC_main.cu:5-60); params as main.cu:160-168.
void * refgpu_create(uint32_t n, uint32_t nnz, const uint32_t * cumulDegs, const uint32_t * neighs, float prob, uint32_t nCol,
                     float epsilon, float lambda, uint32_t tabooIteration, int tailcut, uint32_t maxRip, float numColorRatio, long curandSeed) {
	Handle * h = new Handle;
	h->g_h = new Graph<float, float>(n, false);
	GraphStruct<float, float> * s = h->g_h->getStruct();
	memcpy(s->cumulDegs, cumulDegs, sizeof(uint32_t) * ((size_t)n + 1));
	s->nEdges = nnz;
	s->neighs = new node[nnz ? nnz : 1];
	memcpy(s->neighs, neighs, sizeof(uint32_t) * (size_t)nnz);
	h->g_h->prob = prob;
	h->g_h->doStats();
	h->g_d = new Graph<float, float>(h->g_h);
	h->rnd = new GPURand(n, curandSeed);
	ColoringMCMCParams p;
	p.maxRip = maxRip; p.nCol = nCol; p.numColorRatio = numColorRatio; p.lambda = lambda; p.epsilon = epsilon;
	p.ratioFreezed = 1e-2f; p.tabooIteration = tabooIteration; p.tailcut = tailcut != 0;
	h->col = new RefGPU(h->g_d, h->rnd->randStates, p);
	if (cudaDeviceSynchronize() != cudaSuccess) return nullptr;
	return h;
}

void refgpu_destroy(void * hp) {
	Handle * h = (Handle *)hp;
	if (!h) return;
	delete h->col; delete h->rnd;
	// (the reference's Graph destructor frees the host copy; the device copy is left to process exit like main.cu does)
	delete h;
}

int refgpu_set_colors(void * hp, const uint32_t * c) { RefGPU * r = ((Handle *)hp)->col; return (int)cudaMemcpy(r->colors(), c, r->n() * sizeof(uint32_t), cudaMemcpyHostToDevice); }
int refgpu_get_colors(void * hp, uint32_t * c) { RefGPU * r = ((Handle *)hp)->col; return (int)cudaMemcpy(c, r->colors(), r->n() * sizeof(uint32_t), cudaMemcpyDeviceToHost); }
int refgpu_set_taboo(void * hp, const uint32_t * t) { RefGPU * r = ((Handle *)hp)->col; return (int)cudaMemcpy(r->taboo(), t, r->n() * sizeof(uint32_t), cudaMemcpyHostToDevice); }
int refgpu_get_taboo(void * hp, uint32_t * t) { RefGPU * r = ((Handle *)hp)->col; return (int)cudaMemcpy(t, r->taboo(), r->n() * sizeof(uint32_t), cudaMemcpyDeviceToHost); }

// the draw every vertex would take in the NEXT step (curand_uniform on a copy of the states; the real states are untouched)
int refgpu_peek_draws(void * hp, float * out) {
	RefGPU * r = ((Handle *)hp)->col;
	const uint32_t n = r->n();
	curandState * copy = nullptr; float * d_out = nullptr;
	cudaError_t e = cudaMalloc(&copy, sizeof(curandState) * (size_t)n);
	if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(float) * (size_t)n);
	if (e == cudaSuccess) e = cudaMemcpy(copy, r->states(), sizeof(curandState) * (size_t)n, cudaMemcpyDeviceToDevice);
	if (e == cudaSuccess) { peek_uniform_kernel<<<(n + 255) / 256, 256>>>(copy, n, d_out); e = cudaDeviceSynchronize(); }
	if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost);
	cudaFree(copy); cudaFree(d_out);
	return (int)e;
}

int refgpu_step_dynamic(void * hp, int prefill) { return ((Handle *)hp)->col->step_dynamic(prefill); }
int refgpu_conflicts(void * hp) { return ((Handle *)hp)->col->conflicts(); }
int refgpu_tailcut(void * hp, int maxRounds, int * rounds) { return ((Handle *)hp)->col->tailcut(maxRounds, rounds); }

// The reference's own run(iteration) (coloringMCMC_main.cu:100-298): writes <dir>.log and <dir>-colors.txt
// (coloringMCMC_prints.cu:27-49,96-230).  Returns rip; *maxIter = "Max iteration reached".
int refgpu_run(void * hp, int iteration, const char * directory, int * maxIter) {
	RefGPU * r = ((Handle *)hp)->col;
	r->setDirectoryPath(std::string(directory));
	std::streambuf * old = std::cout.rdbuf(nullptr);
	r->run(iteration);
	std::cout.rdbuf(old);
	if (maxIter) *maxIter = r->hitMax() ? 1 : 0;
	return (int)r->ripCount();
}

} // extern "C"
