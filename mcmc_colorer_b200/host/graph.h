// CSR graph container, source-compatible with the reference's graph/graph.h for the members the MCMC path uses.
//   GraphStruct  : graph/graph.h:37-79  (cumulDegs[n+1], neighs[nEdges]; nEdges counts both directions)
//   Graph        : graph/graph.h:84-133 (ctor(n, prob, seed) = --simulate, ctor(fileImporter*, bool) = --graph)
// The device copy the reference makes with Graph(Graph*) (graph/graphGPU.cu:210-226) happens inside
// mcmcb200_create(); Graph(Graph*) here only aliases the host graph so that main.cu-shaped code keeps compiling.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

typedef uint32_t node;     // graph node   (graph.h:19)
typedef uint32_t node_sz;  //              (graph.h:20)

class fileImporter;

template <typename nodeW, typename edgeW> struct GraphStruct {
	node      nNodes{0};
	node_sz   nEdges{0};
	node_sz * cumulDegs{nullptr};
	node    * neighs{nullptr};
	nodeW   * nodeWeights{nullptr};
	edgeW   * edgeWeights{nullptr};
	nodeW   * nodeThresholds{nullptr};
	~GraphStruct() { delete[] neighs; delete[] cumulDegs; delete[] nodeWeights; delete[] edgeWeights; delete[] nodeThresholds; }
	bool is_valid() const {                                    // graph.h:56-63
		for (uint32_t i = 0; i < nEdges; i++) if (neighs[i] > nNodes - 1) return false;
		return cumulDegs[nNodes] == nEdges;
	}
	node_sz deg(node i) const { return cumulDegs[i + 1] - cumulDegs[i]; }
};

template <typename nodeW, typename edgeW> class Graph {
public:
	// n <= kExactRandLimit reproduces Graph::setupRnd2 bit for bit (libc rand(), graphCPU.cpp:290-404); larger n uses
	// an O(E) sampler of the same G(n,p) family seeded with `seed` (the reference needs n(n+1)/2 bits: 62 GB at 1 M).
	static constexpr node kExactRandLimit = 50000;
	Graph(node nn, float prob, uint32_t seed);
	Graph(fileImporter * imp, bool GPUEnb);
	Graph(Graph<nodeW, edgeW> * const fullGraph);                  // "device" alias, see header comment
	Graph(node nn, const node_sz * cumulDegs, const node * neighs, float prob);   // adopt-by-copy of an existing CSR
	~Graph();
	// the O(E) sampler for any n (the (n, prob, seed) ctor selects it above kExactRandLimit)
	static Graph<nodeW, edgeW> * makeFast(node nn, float prob, uint32_t seed) {
		Graph<nodeW, edgeW> * g = new Graph<nodeW, edgeW>();
		g->prob = prob; g->setupRndFast(nn, prob, seed);
		return g;
	}

	GraphStruct<nodeW, edgeW> * getStruct() { return str; }
	GraphStruct<nodeW, edgeW> * getStruct() const { return str; }
	node getMaxNodeDeg() { return maxDeg; }
	node getMinNodeDeg() { return minDeg; }
	float getMeanNodeDeg() { return meanDeg; }
	bool isGPUEnabled() { return alias; }
	void doStats();                                                 // graphCPU.cpp:432-450
	void setupRnd2(node nn, float prob, uint32_t seed);
	void setupRndFast(node nn, float prob, uint32_t seed);
	void setupImporterNew();                                        // graphCPU.cpp:112-170

	float prob{0.0f};                                               // graph.h:129

private:
	Graph() {}
	float density{0.0f};
	GraphStruct<nodeW, edgeW> * str{nullptr};
	node maxDeg{0}, minDeg{0};
	float meanDeg{0.0f};
	bool connected{true};
	bool alias{false};
	fileImporter * fImport{nullptr};
};
