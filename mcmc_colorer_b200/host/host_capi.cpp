// Small C interface to the C++ host graph layer (libmcmcb200_host.so) so that the Python tests can compare the
// host-side CSR construction (--simulate / --graph paths) with the reference's, array for array.
#include <cstdint>
#include <cstring>
#include <string>

#include "fileImporter.h"
#include "graph.h"

extern "C" {

void * mcmchost_graph_simulate(uint32_t n, float prob, uint32_t seed) { return new Graph<float, float>(n, prob, seed); }

void * mcmchost_graph_simulate_fast(uint32_t n, float prob, uint32_t seed) { return Graph<float, float>::makeFast(n, prob, seed); }

void * mcmchost_graph_from_file(const char * path) {
	fileImporter * imp = new fileImporter(std::string(path), "");
	Graph<float, float> * g = new Graph<float, float>(imp, false);
	delete imp;
	return g;
}

void mcmchost_graph_info(void * gp, uint32_t * n, uint32_t * nnz, uint32_t * maxDeg, uint32_t * minDeg, float * meanDeg, float * prob) {
	Graph<float, float> * g = (Graph<float, float> *)gp;
	*n = g->getStruct()->nNodes; *nnz = g->getStruct()->nEdges;
	*maxDeg = g->getMaxNodeDeg(); *minDeg = g->getMinNodeDeg(); *meanDeg = g->getMeanNodeDeg(); *prob = g->prob;
}

void mcmchost_graph_copy_csr(void * gp, uint32_t * cumulDegs, uint32_t * neighs) {
	auto * s = ((Graph<float, float> *)gp)->getStruct();
	std::memcpy(cumulDegs, s->cumulDegs, sizeof(uint32_t) * ((size_t)s->nNodes + 1));
	std::memcpy(neighs, s->neighs, sizeof(uint32_t) * (size_t)s->nEdges);
}

void mcmchost_graph_free(void * gp) { delete (Graph<float, float> *)gp; }

} // extern "C"
