#include "graph.h"
#include "fileImporter.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>

template <typename nodeW, typename edgeW>
Graph<nodeW, edgeW>::Graph(node nn, float prob, uint32_t seed) : prob{prob} {
	if (nn <= kExactRandLimit) setupRnd2(nn, prob, seed);
	else setupRndFast(nn, prob, seed);
}

template <typename nodeW, typename edgeW>
Graph<nodeW, edgeW>::Graph(fileImporter * imp, bool) : fImport{imp} {
	setupImporterNew();
	// reference graphCPU.cpp:25: file graphs get prob = nEdges / n^2 (uint32 product, then float)
	prob = str->nEdges / (float)(str->nNodes * str->nNodes);
}

template <typename nodeW, typename edgeW>
Graph<nodeW, edgeW>::Graph(Graph<nodeW, edgeW> * const g)
	: prob{g->prob}, density{g->density}, str{g->str}, maxDeg{g->maxDeg}, minDeg{g->minDeg}, meanDeg{g->meanDeg},
	  connected{g->connected}, alias{true} {}

template <typename nodeW, typename edgeW>
Graph<nodeW, edgeW>::Graph(node nn, const node_sz * cumulDegs, const node * neighs, float p) : prob{p} {
	str = new GraphStruct<nodeW, edgeW>;
	str->nNodes = nn;
	str->cumulDegs = new node_sz[(size_t)nn + 1];
	std::memcpy(str->cumulDegs, cumulDegs, sizeof(node_sz) * ((size_t)nn + 1));
	str->nEdges = cumulDegs[nn];
	str->neighs = new node[std::max<size_t>(str->nEdges, 1)];
	std::memcpy(str->neighs, neighs, sizeof(node) * (size_t)str->nEdges);
	doStats();
}

template <typename nodeW, typename edgeW>
Graph<nodeW, edgeW>::~Graph() { if (!alias) delete str; }

// Bit-for-bit Graph::setupRnd2 (graphCPU.cpp:290-404): n(n+1)/2 libc rand() draws over the upper triangle incl. the
// diagonal (row j, column i >= j), diagonal cleared, both directions stored, neighbour lists ascending.
// `seed` is unused there too: the generator is libc rand() in whatever state the process left it (ArgHandle only
// calls srand when no --seed is given, ArgHandle.cpp:272-276).
template <typename nodeW, typename edgeW>
void Graph<nodeW, edgeW>::setupRnd2(node n, float prob, uint32_t) {
	const size_t nn = n, vecSize = nn * (nn + 1) / 2;
	std::vector<bool> tri(vecSize);
	for (size_t k = 0; k < vecSize; k++) tri[k] = ((double)rand() / (RAND_MAX)) >= prob ? 0 : 1;
	str = new GraphStruct<nodeW, edgeW>;
	str->cumulDegs = new node_sz[nn + 1];
	std::fill(str->cumulDegs, str->cumulDegs + nn + 1, 0);
	str->nNodes = n;
	size_t i = 0, j = 0;
	for (size_t k = 0; k < vecSize; k++) {
		if (j == i) tri[k] = 0;
		if (tri[k]) { str->cumulDegs[i + 1]++; str->cumulDegs[j + 1]++; str->nEdges += 2; }
		if (++i == nn) { j++; i = j; }
	}
	for (size_t v = 1; v <= nn; v++) str->cumulDegs[v] += str->cumulDegs[v - 1];
	str->neighs = new node[std::max<size_t>(str->nEdges, 1)];
	std::vector<node_sz> fill(nn, 0);
	i = j = 0;
	for (size_t k = 0; k < vecSize; k++) {
		if (tri[k]) {
			str->neighs[str->cumulDegs[j] + fill[j]++] = (node)i;
			str->neighs[str->cumulDegs[i] + fill[i]++] = (node)j;
		}
		if (++i == nn) { j++; i = j; }
	}
	doStats();
}

// O(E) sampler for large n: m = round(p * n(n-1)/2) undirected pairs drawn uniformly (mt19937_64(seed)), canonicalised,
// de-duplicated, symmetrised; neighbour lists ascending, no self loops.
template <typename nodeW, typename edgeW>
void Graph<nodeW, edgeW>::setupRndFast(node n, float prob, uint32_t seed) {
	const uint64_t nn = n;
	const uint64_t m = (uint64_t)((double)prob * (double)nn * (double)(nn - 1) / 2.0 + 0.5);
	std::mt19937_64 eng(seed ? seed : 1);
	std::vector<uint64_t> keys;
	keys.reserve(m);
	for (uint64_t e = 0; e < m; e++) {
		uint64_t a = eng() % nn, b = eng() % nn;
		if (a == b) continue;
		if (a > b) std::swap(a, b);
		keys.push_back(a * nn + b);
	}
	std::sort(keys.begin(), keys.end());
	keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
	str = new GraphStruct<nodeW, edgeW>;
	str->nNodes = n;
	str->cumulDegs = new node_sz[nn + 1];
	std::fill(str->cumulDegs, str->cumulDegs + nn + 1, 0);
	for (uint64_t k : keys) { str->cumulDegs[k / nn + 1]++; str->cumulDegs[k % nn + 1]++; }
	for (uint64_t v = 1; v <= nn; v++) str->cumulDegs[v] += str->cumulDegs[v - 1];
	str->nEdges = str->cumulDegs[nn];
	str->neighs = new node[std::max<size_t>(str->nEdges, 1)];
	std::vector<node_sz> fill(nn, 0);
	// keys are sorted by (lo, hi): for vertex v, neighbours lo < v arrive in ascending lo, then hi > v ascending
	for (uint64_t k : keys) { const uint64_t hi = k % nn, lo = k / nn; str->neighs[str->cumulDegs[hi] + fill[hi]++] = (node)lo; }
	for (uint64_t k : keys) { const uint64_t hi = k % nn, lo = k / nn; str->neighs[str->cumulDegs[lo] + fill[lo]++] = (node)hi; }
	doStats();
}

// graphCPU.cpp:112-170: two passes over the edge list, self loops dropped, the back edge added for every edge.
template <typename nodeW, typename edgeW>
void Graph<nodeW, edgeW>::setupImporterNew() {
	const uint32_t nn = fImport->nNodes;
	str = new GraphStruct<nodeW, edgeW>;
	str->cumulDegs = new node_sz[(size_t)nn + 1];
	std::fill(str->cumulDegs, str->cumulDegs + (nn + 1), 0);
	str->nNodes = nn;
	fImport->fRewind();
	while (fImport->getNextEdge()) {
		if (fImport->edgeIsValid && fImport->srcIdx != fImport->dstIdx) {
			str->cumulDegs[fImport->srcIdx + 1]++; str->cumulDegs[fImport->dstIdx + 1]++; str->nEdges += 2;
		}
	}
	for (uint32_t i = 1; i < nn + 1; i++) str->cumulDegs[i] += str->cumulDegs[i - 1];
	str->neighs = new node[std::max<size_t>(str->nEdges, 1)];
	std::vector<size_t> fill(nn, 0);
	fImport->fRewind();
	while (fImport->getNextEdge()) {
		if (fImport->edgeIsValid && fImport->srcIdx != fImport->dstIdx) {
			str->neighs[str->cumulDegs[fImport->srcIdx] + fill[fImport->srcIdx]++] = fImport->dstIdx;
			str->neighs[str->cumulDegs[fImport->dstIdx] + fill[fImport->dstIdx]++] = fImport->srcIdx;
		}
	}
	doStats();
}

template <typename nodeW, typename edgeW>
void Graph<nodeW, edgeW>::doStats() {                              // graphCPU.cpp:432-450
	const size_t nn = str->nNodes;
	maxDeg = 0; minDeg = (node)nn;
	for (uint32_t i = 0; i < nn; i++) {
		maxDeg = std::max<node>(maxDeg, str->deg(i));
		minDeg = std::min<node>(minDeg, str->deg(i));
	}
	density = (float)str->nEdges / (float)(nn * (nn - 1) / 2);
	meanDeg = (float)str->nEdges / (float)nn;
	connected = minDeg != 0;
}

template class Graph<float, float>;
