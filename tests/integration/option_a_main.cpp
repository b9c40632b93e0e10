// INTEGRATION.md, Option A, as a compile-and-link check (tests/test_host_layer.py::test_integration_option_a_compiles):
// the colourer section of the reference's main.cu (:160-198), statement for statement, compiled against the REFERENCE's own
// graph/graph.h, graph/graphCPU.cpp and graph_coloring/coloring.h -- with host/coloringMCMC.{h,cpp} of this repository dropped in
// for graph_coloring/coloringMCMC.h + coloringMCMC_{main,balance,standard,decrease,utils,prints}.cu and coloringMCMC_CPU.{h,cpp}.
// Built only where /root/reference exists; on a box without a B200 it must fail LOUDLY at mcmcb200_create (no CPU fallback).
#include <ctime>
#include <fstream>
#include <iostream>

#include "graph/graph.h"
#include "graph/graphCPU.cpp"            // template bodies live in the .cpp (main.cu:10 does the same)
#include "graph_coloring/coloring.h"
#include "coloringMCMC.h"                // <- the drop-in (copied next to this file by the test)

bool g_traceLogEn = false;               // utils/miscUtils.h
// GPU members of the reference's Graph<> live in graphGPU.cu; this check links without nvcc
template<> void Graph<float, float>::setMemGPU(node_sz, int) {}
template<> void Graph<float, float>::deleteMemGPU() {}
template<> void Graph<float, float>::setupImporterGPU() {}
template<> void Graph<float, float>::setupReduxGPU(const uint32_t * const, const uint32_t, const int32_t * const,
	GraphStruct<float, float> * const, const uint32_t * const, const uint32_t * const, const float * const) {}
template class Graph<float, float>;

int main(int argc, char * argv[]) {
	const uint32_t N = 300, seed = 7, i = 0;
	const float prob = 0.1f, numColorRatio = 1.0f;
	std::string outDir = argc > 1 ? argv[1] : ".";
	std::streambuf * old = std::cout.rdbuf(nullptr);                     // setupRnd2 prints progress bars
	Graph<float, float> * test = new Graph<float, float>(N, prob, seed);  // main.cu:63
	std::cout.rdbuf(old);
	curandState * randStates = nullptr;                                  // GPURandGen.randStates (main.cu:80): accepted, ignored

	ColoringMCMCParams params;                                           // main.cu:160-168
	params.numColorRatio = numColorRatio;
	params.nCol = test->getMaxNodeDeg() * numColorRatio;
	params.epsilon = 1e-8f;
	params.lambda = 1.0f;
	params.ratioFreezed = 1e-2;
	params.maxRip = 250;
	params.tabooIteration = 0;
	params.tailcut = false;
	try {
		ColoringMCMC_CPU<float, float> mcmc_cpu(test, params, seed + i); // main.cu:171
		mcmc_cpu.run();                                                  // :175
		std::ofstream cpuFileLog(outDir + "/optA-MCMC_CPU-0.log");
		mcmc_cpu.saveStats(i, 0.0f, cpuFileLog);                         // :185
		std::ofstream cpuFileColors(outDir + "/optA-MCMC_CPU-0-colors.txt");
		mcmc_cpu.saveColor(cpuFileColors);                               // :188

		ColoringMCMC<float, float> colMCMC(test, randStates, params);    // :192 (host graph: the upload happens in mcmcb200_create)
		colMCMC.setDirectoryPath(outDir + "/optA-MCMC_GPU-0");           // :194
		colMCMC.run(i);                                                  // :197
	} catch (const std::exception & e) {
		std::cerr << "option_a: " << e.what() << std::endl;
		return 2;
	}
	std::cout << "option_a ok" << std::endl;
	return 0;
}
