// MCMC_Colorer driver: same flow, flags, defaults and output files as the reference's main.cu:28-215, with the
// colourers running on libmcmcb200 (sm_100a).
#include <chrono>
#include <fstream>
#include <iostream>

#include "ArgHandle.h"
#include "coloring.h"
#include "coloringMCMC.h"
#include "fileImporter.h"
#include "graph.h"

int main(int argc, char * argv[]) {
	ArgHandle commandLine(argc, argv);
	commandLine.processCommandLine();
	uint32_t N;
	float prob;
	uint32_t seed = commandLine.seed;
	uint32_t nColFromC = commandLine.nCol;
	std::string outDir = commandLine.outDir;
	float numColorRatio = 1.0f / (float)commandLine.numColRatio;        // main.cu:53
	uint32_t repet = commandLine.repetitions;

	Graph<float, float> * test;
	fileImporter * fImport = nullptr;
	if (commandLine.simulate) {                                         // main.cu:60-63
		N = commandLine.n;
		prob = (float)commandLine.prob;
		test = new Graph<float, float>(N, prob, seed);
	} else {                                                            // main.cu:64-70
		fImport = new fileImporter(commandLine.graphFilename, "");
		test = new Graph<float, float>(fImport, false);
		prob = test->getStruct()->nEdges / (float)(test->getStruct()->nNodes * test->getStruct()->nNodes);
		N = test->getStruct()->nNodes;
	}
	std::cout << "Nodes: " << test->getStruct()->nNodes << " - Edges: " << test->getStruct()->nEdges << std::endl;
	std::cout << "Min Degree: " << test->getMinNodeDeg() << " - Max Degree: " << test->getMaxNodeDeg() << " - Mean Degree: "
	          << test->getMeanNodeDeg() << std::endl;

	Graph<float, float> graph_d(test);                                   // main.cu:78 (device upload happens in the colourer)
	int rcode = EXIT_SUCCESS;

	for (uint32_t i = 0; i < repet; i++) {
		std::cout << "Repetition: " << i << std::endl;
		if (commandLine.greedyff || commandLine.rebalanced_greedyff)
			std::cout << "--grdffgpu/--vffgpu: not part of this build (different algorithms, out of the hot-path scope, see DESIGN.md)" << std::endl;

		ColoringMCMCParams params;                                       // main.cu:160-168
		params.numColorRatio  = numColorRatio;
		params.nCol           = (nColFromC != 0) ? nColFromC : test->getMaxNodeDeg() * numColorRatio;
		params.epsilon        = 1e-8f;
		params.lambda         = 1.0f;
		params.ratioFreezed   = 1e-2;
		params.maxRip         = 250;
		params.tabooIteration = commandLine.tabooIteration;
		params.tailcut        = commandLine.tailcut;

		try {
			if (commandLine.lubygpu) {                                   // main.cu:90-109 (cross-check colourer)
				ColoringLuby<float, float> colLuby(&graph_d, nullptr, seed + i, commandLine.device);
				auto t0 = std::chrono::steady_clock::now();
				colLuby.run_fast();
				double duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				std::cout << "LubyGPU - number of colors: " << colLuby.getNumOfColors() << std::endl;
				std::cout << "LubyGPU elapsed time: " << duration << std::endl;
				std::ofstream lubyFileLog, lubyFileColors;
				lubyFileLog.open(outDir + "/" + commandLine.graphName + "-LUBY-" + std::to_string(i) + ".log");
				colLuby.saveStats(i, duration, lubyFileLog);
				lubyFileLog.close();
				lubyFileColors.open(outDir + "/" + commandLine.graphName + "-LUBY-" + std::to_string(i) + "-colors.txt");
				colLuby.saveColor(lubyFileColors);
				lubyFileColors.close();
			}
			if (commandLine.mcmccpu) {                                   // main.cu:170-190
				ColoringMCMC_CPU<float, float> mcmc_cpu(test, params, seed + i);
				auto t0 = std::chrono::steady_clock::now();
				mcmc_cpu.run();
				double duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				std::cout << "MCMC_CPU elapsed time: " << duration << std::endl;
				std::ofstream cpuFileLog, cpuFileColors;
				cpuFileLog.open(outDir + "/" + commandLine.graphName + "-MCMC_CPU-" + std::to_string(i) + ".log");
				mcmc_cpu.saveStats(i, duration, cpuFileLog);
				cpuFileLog.close();
				cpuFileColors.open(outDir + "/" + commandLine.graphName + "-MCMC_CPU-" + std::to_string(i) + "-colors.txt");
				mcmc_cpu.saveColor(cpuFileColors);
				cpuFileColors.close();
			}
			if (commandLine.mcmcgpu) {                                   // main.cu:192-202
				ColoringMCMCOptions opt;
				opt.proposal = commandLine.proposal == "uniform" ? 0u : 1u;
				opt.convergence = 1;
				opt.seed = seed + i;
				opt.device = commandLine.device;
				opt.sweepsPerCheck = commandLine.sweepsPerCheck;
				ColoringMCMC<float, float> colMCMC(&graph_d, nullptr, params, opt);
				colMCMC.setDirectoryPath(outDir + "/" + commandLine.graphName + "-MCMC_GPU-" + std::to_string(i));
				auto t0 = std::chrono::steady_clock::now();
				colMCMC.run(i);
				double duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				std::cout << "MCMC GPU elapsed time: " << duration << std::endl << std::endl;
			}
		} catch (const McmcError & e) {
			std::cerr << "MCMC_Colorer: " << e.what() << std::endl;
			rcode = EXIT_FAILURE;
			break;
		}
	}
	delete test;
	delete fImport;
	return rcode;
}
