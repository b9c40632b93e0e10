// Host-side colourer classes over the C ABI (include/mcmcb200.h), source-compatible with the reference:
//   ColoringMCMC<nodeW,edgeW>      graph_coloring/coloringMCMC.h:43-140       (ctor(graph_d, randStates, params),
//                                   setDirectoryPath, run(iteration); writes <dir>.log and <dir>-colors.txt)
//   ColoringMCMC_CPU<nodeW,edgeW>  graph_coloring/coloringMCMC_CPU.h:12-31     (ctor(graph, params, seed), run(),
//                                   saveStats(it, duration, ofstream&), saveColor(ofstream&))
// Both classes drive the same sm_100a sweep kernels.  ColoringMCMC_CPU keeps the *semantics* of the reference CPU
// class (uniform proposal of fill_p, violating-vertex convergence test, its log format) but there is no CPU compute
// path in this build: without a B200 the constructors throw.
#pragma once
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "coloring.h"
#include "graph.h"

// Additions of the B200 build that have no slot in the reference struct; passed next to it.
struct ColoringMCMCOptions {
	uint32_t proposal = 1;     // MCMCB200_PROPOSAL_DYNAMIC: the shipped GPU variant (coloringMCMC.h:39)
	uint32_t convergence = 1;  // conflicting edges (coloringMCMC_main.cu:169)
	uint64_t seed = 0;
	int      device = -1;
	uint32_t sweepsPerCheck = 1;
};

struct mcmcb200_handle;
struct curandStateXORWOW;                 // the reference passes curandState*; accepted and ignored (stateless Philox)
typedef struct curandStateXORWOW curandState;

struct McmcError : public std::runtime_error {
	int code;
	McmcError(int c, const std::string & what) : std::runtime_error(what), code(c) {}
};

template <typename nodeW, typename edgeW> class ColoringMCMC {
public:
	ColoringMCMC(Graph<nodeW, edgeW> * inGraph_d, curandState * randStates, ColoringMCMCParams params);
	ColoringMCMC(Graph<nodeW, edgeW> * inGraph_d, curandState * randStates, ColoringMCMCParams params, ColoringMCMCOptions opt);
	~ColoringMCMC();
	void run(int iteration);
	void setDirectoryPath(std::string directory) { this->directory = directory; }

	// results of the last run (the reference keeps them private and only writes files)
	const std::vector<uint32_t> & getColors() const { return colors; }
	uint64_t getConflictEdges() const { return conflictEdges; }
	uint32_t getRip() const { return rip; }
	bool     getMaxIterReached() const { return maxIterReached; }
	double   getDuration() const { return duration; }
	float    getStdDev() const { return stdDev; }
	float    getBalancingIndex() const { return balancingIndex; }
	uint32_t getUsedColors() const { return usedColors; }

protected:
	Graph<nodeW, edgeW> * graph;
	ColoringMCMCParams param;
	ColoringMCMCOptions opt;
	mcmcb200_handle * h{nullptr};
	uint32_t nnodes;
	float prob;
	uint32_t rip{0};
	bool maxIterReached{false};
	double duration{0};
	uint64_t conflictEdges{0};
	float stdDev{0}, balancingIndex{0};
	uint32_t usedColors{0};
	std::vector<uint32_t> colors;
	std::string directory;
};

template <typename nodeW, typename edgeW> class ColoringMCMC_CPU {
public:
	ColoringMCMC_CPU(Graph<nodeW, edgeW> * g, ColoringMCMCParams params, uint32_t seed);
	~ColoringMCMC_CPU();
	void run();
	void saveStats(size_t iter, float duration, std::ofstream & outFile);   // coloringMCMC_CPUutils.cpp:69-102
	void saveColor(std::ofstream & outfile);                                // coloringMCMC_CPUutils.cpp:105-109
	std::vector<uint32_t> * getC() { return &C; }
	size_t getIterations() const { return iter; }
	bool   getMaxIterReached() const { return maxIterReached; }

protected:
	Graph<nodeW, edgeW> * graph;
	ColoringMCMCParams param;
	mcmcb200_handle * h{nullptr};
	std::vector<uint32_t> C;
	std::vector<uint64_t> hist;
	size_t nNodes;
	uint32_t seed;
	size_t iter{0};
	bool maxIterReached{false};
};

// ColoringLuby<nodeW,edgeW>  graph_coloring/coloringLuby.h:30-93 (ctor(graph_d, randStates), run()/run_fast(), saveStats, saveColor,
// getColoringGPU()->nCol): the cross-check colourer, host-driven MIS rounds on the device (mcmcb200_luby_color; run_fast's
// dynamic parallelism does not exist on sm_100, both entry points run the same loop).  Colours are 1-based like the reference's.
template <typename nodeW, typename edgeW> class ColoringLuby {
public:
	ColoringLuby(Graph<nodeW, edgeW> * inGraph_d, curandState * randStates, uint64_t seed = 0, int device = -1);
	void run();
	void run_fast() { run(); }
	void saveStats(size_t it, float duration, std::ofstream & outFile);    // coloringLuby.cu:179-211
	void saveColor(std::ofstream & outfile);                               // coloringLuby.cu:214-219
	uint32_t getNumOfColors() const { return numOfColors; }
	const std::vector<uint32_t> & getColors() const { return C; }

protected:
	Graph<nodeW, edgeW> * graph;
	uint32_t nnodes;
	uint64_t seed;
	int device;
	uint32_t numOfColors{0}, rounds{0};
	std::vector<uint32_t> C;
};

