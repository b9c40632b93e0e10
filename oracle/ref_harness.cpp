// TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// Harness around the UNMODIFIED reference CPU MCMC colourer.  The reference translation
// units (graph_coloring/coloringMCMC_CPU.cpp, coloringMCMC_CPUutils.cpp, utils/*.cpp) are
// compiled where they lie under /root/reference by oracle/Makefile; nothing is copied.
// This file only adds what the reference needs to link without a GPU / network and a
// C interface (ctypes) through which tests drive the reference's own public methods:
//
//   * no-op specialisations of the four Graph<float,float> GPU members that graph.h:101
//     references (their bodies live in graphGPU.cu, which needs a device);
//   * a stub `dbg` (utils/dbg.cpp forks `stty` twice per sweep and has UB on Linux,
//     dbg.cpp:41-67; ColoringMCMC_CPU::run dereferences g_debugger unconditionally,
//     coloringMCMC_CPU.cpp:244);
//   * struct RefMCMC : ColoringMCMC_CPU<float,float> -- data members are `protected`
//     (coloringMCMC_CPU.h:52-105), so a derived class may set the colouring, replay a
//     draw tape and read C / Cviols / freeColors.  One sweep = exactly the calls of
//     ColoringMCMC_CPU::run's loop body (coloringMCMC_CPU.cpp:152, 183-204, 259-260).
//
// Overflow contract: when the sequential CDF walk of
// extract_new_color (coloringMCMC_CPU.cpp:510-514) runs off the end, the reference
// calls libc rand() (:517-520).  The harness detects that case on the reference-filled
// `p` *before* calling extract_new_color and applies the contract rule (idx = nCol-1,
// the rule of the reference GPU kernels, coloringMCMC_balance.cu:128,138).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <fstream>

#include "graph/graph.h"
#include "graph/graphCPU.cpp"          // template bodies live in the .cpp (main.cu:10 does the same)
#include "graph_coloring/coloring.h"
#include "graph_coloring/coloringMCMC_CPU.h"
#include "utils/dbg.h"

bool   g_traceLogEn = false;           // declared in utils/miscUtils.h
dbg  * g_debugger   = nullptr;         // main.cu:24

// ---- GPU members of Graph<> : never reached on the CPU path -----------------------
template<> void Graph<float, float>::setMemGPU(node_sz, int) {}
template<> void Graph<float, float>::deleteMemGPU() {}
template<> void Graph<float, float>::setupImporterGPU() {}
template<> void Graph<float, float>::setupReduxGPU(const uint32_t * const, const uint32_t, const int32_t * const,
	GraphStruct<float, float> * const, const uint32_t * const, const uint32_t * const, const float * const) {}
template class Graph<float, float>;

// ---- stub debugger ------------------------------------------------------------------
dbg::dbg() {}
dbg::dbg(Graph<float, float> * gg, ColoringMCMC_CPU<float, float> * colMCMC) : gr(gg), col(colMCMC) {}
dbg::~dbg() {}
bool dbg::check_F12keypress() { return false; }
void dbg::stop_and_debug() {}

// ---- derived colourer ---------------------------------------------------------------
struct RefMCMC : public ColoringMCMC_CPU<float, float> {
	RefMCMC(Graph<float, float> * g, ColoringMCMCParams p, uint32_t seed) : ColoringMCMC_CPU<float, float>(g, p, seed) {
		// run() sets the identity colour permutation (coloringMCMC_CPU.cpp:131-132); fill_p reads it (:473)
		size_t ii = 0;
		for (auto & v : colorIdx) v = ii++;
	}

	// Sweep vertices [vb, ve) of the current colouring with an external draw tape u[0..n).
	// do_swap=1 performs the std::swap(C, Cstar) of coloringMCMC_CPU.cpp:259.
	// Returns the violating-vertex count of C *before* the sweep (the value run() tests, :136/:152).
	double lastLoopSeconds = 0;   // wall time of the per-vertex loop of the last sweep_tape (the :183-204 loop only)
	size_t sweep_tape(const float * u, size_t vb, size_t ve, int do_swap, uint64_t * overflowCount) {
		Cviol = violation_count(C, Cviols);                                   // :152
		for (size_t i = 0; i < nNodes; i++) nodeProbab[i] = u[i];              // :139 (draws replaced by the tape)
		if (vb > 0 || ve < nNodes) Cstar = C;                                  // partial sweeps leave the rest unchanged
		auto t0 = std::chrono::steady_clock::now();
		for (size_t i = vb; i < ve; i++) {                                     // :183
			size_t Zvcomp = count_free_colors(i, C, freeColors);               // :191
			size_t Zv = nCol - Zvcomp;                                         // :192
			fill_p(i, Zv);                                                     // :195
			if (taboo[i] == 0) {
				// overflow detection on the reference-filled p, same walk as :510-514
				float cdf = 0; size_t idx;
				for (idx = 0; idx < p.size(); idx++) { cdf += p[idx]; if (cdf > nodeProbab[i]) break; }
				if (idx >= nCol) {
					if (overflowCount) (*overflowCount)++;
					idx = nCol - 1;                                            // contract rule
					q[i] = p[idx];                                             // :522
					Cstar[i] = (uint32_t)idx;                                  // :523
					taboo[i] = (Cstar[i] == C[i]) * tabooIteration;            // :526
					continue;
				}
			}
			extract_new_color(i, p, nodeProbab, q, Cstar);                     // :198
		}
		lastLoopSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		size_t before = Cviol;
		if (do_swap) std::swap(C, Cstar);                                      // :259
		return before;
	}

	// The loop of run() (coloringMCMC_CPU.cpp:136-270) with the object's own generator,
	// without the g_debugger call and without the (non-terminating, :296) tail-cut loop.
	size_t run_native(size_t * sweepsOut) {
		Cviol = violation_count(C, Cviols);                                    // :127
		while (Cviol > z) {                                                    // :136
			for (auto & val : nodeProbab) val = unifDistr(gen);                // :139
			Cviol = violation_count(C, Cviols);                                // :152
			for (size_t i = 0; i < nNodes; i++) {
				size_t Zvcomp = count_free_colors(i, C, freeColors);
				fill_p(i, nCol - Zvcomp);
				extract_new_color(i, p, nodeProbab, q, Cstar);
			}
			Cstarviol = violation_count(Cstar, Cstarviols);                    // :211
			std::swap(C, Cstar); std::swap(Cviol, Cstarviol);                  // :259-260
			iter++;                                                            // :264
			if (iter > maxiter) { maxIterReached = true; break; }              // :265-269
		}
		if (sweepsOut) *sweepsOut = iter;
		return Cviol;
	}

	// diagnostic: run_native's loop, reporting every CDF-walk overflow (vertex, draw, final cdf) before it happens
	size_t scan_overflows(size_t maxSweeps, double * rec /* [4*cap] */, size_t cap) {
		size_t found = 0;
		Cviol = violation_count(C, Cviols);
		for (size_t s = 0; s < maxSweeps && Cviol > z; s++) {
			for (auto & val : nodeProbab) val = unifDistr(gen);
			Cviol = violation_count(C, Cviols);
			for (size_t i = 0; i < nNodes; i++) {
				size_t Zvcomp = count_free_colors(i, C, freeColors);
				fill_p(i, nCol - Zvcomp);
				float cdf = 0; size_t idx;
				for (idx = 0; idx < p.size(); idx++) { cdf += p[idx]; if (cdf > nodeProbab[i]) break; }
				if (idx >= nCol && found < cap) {
					rec[4 * found] = (double)s; rec[4 * found + 1] = (double)i;
					rec[4 * found + 2] = (double)nodeProbab[i]; rec[4 * found + 3] = (double)cdf; found++;
				}
				extract_new_color(i, p, nodeProbab, q, Cstar);
			}
			Cstarviol = violation_count(Cstar, Cstarviols);
			std::swap(C, Cstar); std::swap(Cviol, Cstarviol);
		}
		return found;
	}

	std::vector<uint32_t> & colours() { return C; }
	std::vector<uint32_t> & tabooVec() { return taboo; }
	std::vector<bool> & viols() { return Cviols; }
	std::vector<bool> & freeCols() { return freeColors; }
	std::vector<float> & pVec() { return p; }
	size_t iterations() const { return iter; }
	bool hitMaxIter() const { return maxIterReached; }
	uint32_t nColours() const { return nCol; }
};

struct RefHandle {
	Graph<float, float> * graph;
	RefMCMC * mcmc;
};

extern "C" {

// Graph(n, prob, seed) -> setupRnd2 (graphCPU.cpp:29-31, 290-404).  Uses libc rand();
// reseed!=0 calls srand(reseed) first (srand(1) == glibc's initial state).
void * ref_graph_simulate(uint32_t n, float prob, uint32_t reseed) {
	if (reseed) srand(reseed);
	// setupRnd2 prints progress bars to stdout; silence them
	std::streambuf * old = std::cout.rdbuf(nullptr);
	Graph<float, float> * g = new Graph<float, float>(n, prob, 0u);
	std::cout.rdbuf(old);
	return g;
}

// Wrap an arbitrary CSR in the reference's Graph (graph.h:94 ctor + public getStruct()).
void * ref_graph_from_csr(uint32_t n, uint32_t nnz, const uint32_t * cumulDegs, const uint32_t * neighs, float prob) {
	Graph<float, float> * g = new Graph<float, float>(n, false);               // setup(): allocs cumulDegs[n+1]
	GraphStruct<float, float> * s = g->getStruct();
	memcpy(s->cumulDegs, cumulDegs, sizeof(uint32_t) * ((size_t)n + 1));
	s->nEdges = nnz;
	s->neighs = new node[nnz ? nnz : 1];
	memcpy(s->neighs, neighs, sizeof(uint32_t) * (size_t)nnz);
	g->prob = prob;
	g->doStats();                                                              // graphCPU.cpp:432-450
	return g;
}

// Graph(fileImporter*, bool) -> setupImporterNew (graphCPU.cpp:19-26,112-170) through utils/fileImporter.cpp
void * ref_graph_from_file(const char * path) {
	std::streambuf * old = std::cout.rdbuf(nullptr);
	fileImporter * imp = new fileImporter(std::string(path), "");
	Graph<float, float> * g = new Graph<float, float>(imp, false);
	std::cout.rdbuf(old);
	delete imp;
	return g;
}

void ref_graph_info(void * gp, uint32_t * n, uint32_t * nnz, uint32_t * maxDeg, uint32_t * minDeg, float * meanDeg) {
	Graph<float, float> * g = (Graph<float, float> *)gp;
	*n = g->getStruct()->nNodes; *nnz = g->getStruct()->nEdges;
	*maxDeg = g->getMaxNodeDeg(); *minDeg = g->getMinNodeDeg(); *meanDeg = g->getMeanNodeDeg();
}

void ref_graph_copy_csr(void * gp, uint32_t * cumulDegs, uint32_t * neighs) {
	GraphStruct<float, float> * s = ((Graph<float, float> *)gp)->getStruct();
	memcpy(cumulDegs, s->cumulDegs, sizeof(uint32_t) * ((size_t)s->nNodes + 1));
	memcpy(neighs, s->neighs, sizeof(uint32_t) * (size_t)s->nEdges);
}

void ref_graph_free(void * gp) { delete (Graph<float, float> *)gp; }

// ColoringMCMC_CPU(Graph*, ColoringMCMCParams, seed) (coloringMCMC_CPU.cpp:7-98); params as main.cu:160-168.
void * ref_mcmc_create(void * gp, uint32_t nCol, float numColorRatio, float lambda, float epsilon, float ratioFreezed,
		uint32_t maxRip, uint32_t tabooIteration, int tailcut, uint32_t seed) {
	ColoringMCMCParams p;
	p.maxRip = maxRip; p.nCol = nCol; p.numColorRatio = numColorRatio; p.lambda = lambda; p.epsilon = epsilon;
	p.ratioFreezed = ratioFreezed; p.tabooIteration = tabooIteration; p.tailcut = tailcut != 0;
	RefHandle * h = new RefHandle;
	h->graph = (Graph<float, float> *)gp;
	h->mcmc = new RefMCMC(h->graph, p, seed);
	return h;
}

void ref_mcmc_free(void * hp) { RefHandle * h = (RefHandle *)hp; delete h->mcmc; delete h; }

void ref_mcmc_get_colors(void * hp, uint32_t * out) {
	auto & C = ((RefHandle *)hp)->mcmc->colours();
	memcpy(out, C.data(), C.size() * sizeof(uint32_t));
}
void ref_mcmc_set_colors(void * hp, const uint32_t * in) {
	auto & C = ((RefHandle *)hp)->mcmc->colours();
	memcpy(C.data(), in, C.size() * sizeof(uint32_t));
}
void ref_mcmc_get_taboo(void * hp, uint32_t * out) {
	auto & T = ((RefHandle *)hp)->mcmc->tabooVec();
	memcpy(out, T.data(), T.size() * sizeof(uint32_t));
}
void ref_mcmc_set_taboo(void * hp, const uint32_t * in) {
	auto & T = ((RefHandle *)hp)->mcmc->tabooVec();
	memcpy(T.data(), in, T.size() * sizeof(uint32_t));
}

// violation_count (coloringMCMC_CPU.cpp:328-351) on an arbitrary colouring; viol[v] in {0,1}.
uint64_t ref_mcmc_violations(void * hp, const uint32_t * colors, uint8_t * viol) {
	RefMCMC * m = ((RefHandle *)hp)->mcmc;
	std::vector<uint32_t> c(colors, colors + m->colours().size());
	std::vector<bool> v(c.size());
	uint64_t r = m->violation_count(c, v);
	if (viol) for (size_t i = 0; i < c.size(); i++) viol[i] = v[i];
	return r;
}

// count_free_colors (coloringMCMC_CPU.cpp:361-383): occ[c] = !freeColors[c]; returns #free.
uint64_t ref_mcmc_occupancy(void * hp, const uint32_t * colors, uint32_t v, uint8_t * occ) {
	RefMCMC * m = ((RefHandle *)hp)->mcmc;
	std::vector<uint32_t> c(colors, colors + m->colours().size());
	std::vector<bool> fc(m->nColours());
	uint64_t freeCnt = m->count_free_colors(v, c, fc);
	for (size_t i = 0; i < fc.size(); i++) occ[i] = !fc[i];
	return freeCnt;
}

// fill_p on the object's current colouring (after a violation_count) -- exposes p for vertex v.
void ref_mcmc_fill_p(void * hp, uint32_t v, float * pOut) {
	RefMCMC * m = ((RefHandle *)hp)->mcmc;
	m->violation_count(m->colours(), m->viols());
	size_t Zvcomp = m->count_free_colors(v, m->colours(), m->freeCols());
	m->fill_p(v, m->nColours() - Zvcomp);
	memcpy(pOut, m->pVec().data(), m->pVec().size() * sizeof(float));
}

uint64_t ref_mcmc_sweep_tape(void * hp, const float * u, uint64_t * overflowCount) {
	RefMCMC * m = ((RefHandle *)hp)->mcmc;
	return m->sweep_tape(u, 0, m->colours().size(), 1, overflowCount);
}

// partial sweep over [vb, ve) (bounded CPU-baseline samples); returns the seconds spent in the reference's
// per-vertex loop (count_free_colors + fill_p + extract_new_color, coloringMCMC_CPU.cpp:183-204) for those vertices
double ref_mcmc_sweep_range_timed(void * hp, const float * u, uint64_t vb, uint64_t ve) {
	RefMCMC * m = ((RefHandle *)hp)->mcmc;
	m->sweep_tape(u, vb, ve, 1, nullptr);
	return m->lastLoopSeconds;
}

// free-running chain with the object's own std::default_random_engine (pins of SURVEY 8c)
uint64_t ref_mcmc_run_native(void * hp, uint64_t * sweeps, int * maxIterReached) {
	RefMCMC * m = ((RefHandle *)hp)->mcmc;
	size_t s = 0;
	uint64_t viol = m->run_native(&s);
	if (sweeps) *sweeps = s;
	if (maxIterReached) *maxIterReached = m->hitMaxIter();
	return viol;
}

// the reference's own run() -- ONLY for seeds known to converge (otherwise :281-311 never terminates)
void ref_mcmc_run(void * hp) {
	RefHandle * h = (RefHandle *)hp;
	dbg stub(h->graph, h->mcmc);
	g_debugger = &stub;
	h->mcmc->run();
	g_debugger = nullptr;
}
uint64_t ref_mcmc_scan_overflows(void * hp, uint64_t maxSweeps, double * rec, uint64_t cap) {
	return ((RefHandle *)hp)->mcmc->scan_overflows(maxSweeps, rec, cap);
}
uint64_t ref_mcmc_iterations(void * hp) { return ((RefHandle *)hp)->mcmc->iterations(); }

// saveStats / saveColor (coloringMCMC_CPUutils.cpp:69-109) -- pins the log formats
void ref_mcmc_save_stats(void * hp, uint64_t it, float duration, const char * path) {
	std::ofstream f(path);
	((RefHandle *)hp)->mcmc->saveStats(it, duration, f);
}
void ref_mcmc_save_colors(void * hp, const char * path) {
	std::ofstream f(path);
	((RefHandle *)hp)->mcmc->saveColor(f);
}

} // extern "C"
