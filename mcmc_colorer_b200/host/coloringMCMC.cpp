#include "coloringMCMC.h"

#include <chrono>
#include <cmath>
#include <iostream>

#include "../../include/mcmcb200.h"

namespace {

void check(int rc, const char * what) {
	if (rc != MCMCB200_OK) {
		std::string msg = std::string(what) + ": " + mcmcb200_strerror(rc);
		if (rc == MCMCB200_ECUDA || rc == MCMCB200_ENODEVICE || rc == MCMCB200_ENOMEM) msg += std::string(" [") + mcmcb200_last_cuda_error() + "]";
		throw McmcError(rc, msg);
	}
}

mcmcb200_params to_abi(const ColoringMCMCParams & p, uint32_t proposal, uint32_t convergence, uint64_t seed, int device) {
	mcmcb200_params a{};
	a.nCol = p.nCol; a.epsilon = p.epsilon; a.lambda = p.lambda; a.numColorRatio = p.numColorRatio;
	a.ratioFreezed = p.ratioFreezed; a.tabooIteration = p.tabooIteration; a.maxRip = p.maxRip; a.tailcut = p.tailcut ? 1u : 0u;
	a.proposal = proposal; a.convergence = convergence; a.seed = seed; a.device = device; a.flags = 0;
	return a;
}

template <typename G> mcmcb200_handle * make_handle(G * graph, const mcmcb200_params & a) {
	auto * s = graph->getStruct();
	mcmcb200_handle * h = nullptr;
	check(mcmcb200_create(&h, s->nNodes, s->nEdges, s->cumulDegs, s->neighs, &a), "mcmcb200_create");
	return h;
}

} // namespace

// ------------------------------------------------------------------------------------------------------------------
// ColoringMCMC  (GPU semantics: coloringMCMC_main.cu:100-298, log format coloringMCMC_prints.cu:27-230)
// ------------------------------------------------------------------------------------------------------------------
template <typename nodeW, typename edgeW>
ColoringMCMC<nodeW, edgeW>::ColoringMCMC(Graph<nodeW, edgeW> * g, curandState * rs, ColoringMCMCParams params)
	: ColoringMCMC(g, rs, params, ColoringMCMCOptions{}) {}

template <typename nodeW, typename edgeW>
ColoringMCMC<nodeW, edgeW>::ColoringMCMC(Graph<nodeW, edgeW> * g, curandState *, ColoringMCMCParams params, ColoringMCMCOptions o)
	: graph(g), param(params), opt(o), nnodes(g->getStruct()->nNodes), prob(g->prob) {
	h = make_handle(g, to_abi(params, o.proposal, o.convergence, o.seed, o.device));
}

template <typename nodeW, typename edgeW>
ColoringMCMC<nodeW, edgeW>::~ColoringMCMC() { mcmcb200_destroy(h); }

template <typename nodeW, typename edgeW>
void ColoringMCMC<nodeW, edgeW>::run(int /*iteration*/) {
	std::ofstream logFile, colorsFile;
	if (!directory.empty()) { logFile.open(directory + ".log"); colorsFile.open(directory + "-colors.txt"); }
	// __customPrintRun0_start (coloringMCMC_prints.cu:40-48); the cudaMemGetInfo line is host specific and omitted
	logFile << "numCol: " << param.nCol << std::endl;
	logFile << "epsilon: " << param.epsilon << std::endl;
	logFile << "lambda: " << param.lambda << std::endl;
	logFile << "ratioFreezed: " << param.ratioFreezed << std::endl;
	logFile << "maxRip: " << param.maxRip << std::endl << std::endl;
	logFile << "numColorRatio: " << param.numColorRatio << std::endl;

	check(mcmcb200_init_colors(h, nullptr), "mcmcb200_init_colors");
	const bool useEdges = opt.convergence == 1;
	auto t0 = std::chrono::steady_clock::now();
	mcmcb200_status_t st{};
	check(mcmcb200_status(h, &st), "mcmcb200_status");
	rip = 0;
	for (;;) {                                                     // do { rip++; ... } while (rip < maxRip), _main.cu:160-269
		rip++;
		if (st.converged) break;                                   // conflictCounter <= z, _main.cu:169
		logFile << "***** Tentativo numero: " << rip << std::endl;  // __customPrintRun2_conflicts
		logFile << "conflitti rilevati: " << (useEdges ? st.conflictEdges : st.violatingVertices) << std::endl;
		check(mcmcb200_sweep(h, opt.sweepsPerCheck), "mcmcb200_sweep");
		check(mcmcb200_status(h, &st), "mcmcb200_status");
		logFile << "nuovi conflitti rilevati: " << (useEdges ? st.conflictEdges : st.violatingVertices) << std::endl;
		if (rip >= param.maxRip) break;
	}
	if (param.tailcut && st.conflictEdges > 0) {                   // _main.cu:271-290
		logFile << "***** Tentativo numero: " << rip << std::endl << "---> TailCutting" << std::endl;
		logFile << "conflitti rilevati: " << st.conflictEdges << std::endl;
		uint32_t rounds = 0;
		check(mcmcb200_tailcut(h, 64, &rounds), "mcmcb200_tailcut");
		check(mcmcb200_status(h, &st), "mcmcb200_status");
		logFile << "nuovi conflitti rilevati: " << st.conflictEdges << std::endl;
	}
	duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	maxIterReached = rip == param.maxRip;                          // _main.cu:294-295
	conflictEdges = st.conflictEdges;

	// __customPrintRun7_end + getStatsNumColors("end_") (coloringMCMC_prints.cu:96-230)
	colors.resize(nnodes);
	check(mcmcb200_get_colors(h, colors.data()), "mcmcb200_get_colors");
	std::vector<uint64_t> hist(param.nCol);
	check(mcmcb200_get_class_sizes(h, hist.data()), "mcmcb200_get_class_sizes");
	int counter = 0, max_i = 0, min_i = nnodes, max_c = 0, min_c = nnodes;
	const int numberOfCol = param.nCol;
	float average = (float)nnodes / numberOfCol, variance = 0, standardDeviation, bal = 0;
	for (int i = 0; i < numberOfCol; i++) {
		const uint32_t s = (uint32_t)hist[i];
		if (s > 0) {
			counter++;
			if ((int)s > max_c) { max_i = i; max_c = s; }
			if ((int)s < min_c) { min_i = i; min_c = s; }
			bal += powf(s - average, 2.f);
		}
	}
	bal /= (nnodes * prob);
	bal = sqrtf(bal);
	for (int i = 0; i < numberOfCol; i++) variance += powf(((uint32_t)hist[i] - average), 2.f);
	variance /= numberOfCol;
	standardDeviation = sqrtf(variance);
	stdDev = standardDeviation; balancingIndex = bal; usedColors = counter;

	logFile << "COLORAZIONE FINALE" << std::endl;
	logFile << "Time " << duration << std::endl;
	logFile << "Max iteration reached " << (rip < param.maxRip ? "no" : "yes") << std::endl;
	for (uint32_t i = 0; i < nnodes; i++) colorsFile << i << " " << colors[i] << "\n";
	logFile << "Number of used colors is " << counter << " on " << numberOfCol << " available" << std::endl;
	logFile << "Most used colors is " << max_i << " used " << max_c << " times" << std::endl;
	logFile << "Least used colors is " << min_i << " used " << min_c << " times" << std::endl;
	logFile << std::endl;
	logFile << "Average " << average << std::endl;
	logFile << "Variance " << variance << std::endl;
	logFile << "StandardDeviation " << standardDeviation << std::endl;
	logFile << "BalancingIndex " << bal << std::endl;
	logFile << std::endl;
	logFile << std::endl << "end colorazione finale -------------------------------------------------------------------" << std::endl << std::endl;
}

// ------------------------------------------------------------------------------------------------------------------
// ColoringMCMC_CPU semantics (coloringMCMC_CPU.cpp:115-321) on the device
// ------------------------------------------------------------------------------------------------------------------
template <typename nodeW, typename edgeW>
ColoringMCMC_CPU<nodeW, edgeW>::ColoringMCMC_CPU(Graph<nodeW, edgeW> * g, ColoringMCMCParams params, uint32_t seed)
	: graph(g), param(params), nNodes(g->getStruct()->nNodes), seed(seed) {
	h = make_handle(g, to_abi(params, MCMCB200_PROPOSAL_UNIFORM, 0 /* violating vertices */, seed, -1));
	check(mcmcb200_init_colors(h, nullptr), "mcmcb200_init_colors");   // the ctor draws the start colouring (:61)
}

template <typename nodeW, typename edgeW>
ColoringMCMC_CPU<nodeW, edgeW>::~ColoringMCMC_CPU() { mcmcb200_destroy(h); }

template <typename nodeW, typename edgeW>
void ColoringMCMC_CPU<nodeW, edgeW>::run() {
	mcmcb200_status_t st{};
	check(mcmcb200_status(h, &st), "mcmcb200_status");
	iter = 0; maxIterReached = false;
	while (!st.converged) {                                        // while (Cviol > z), :136
		check(mcmcb200_sweep(h, 1), "mcmcb200_sweep");
		check(mcmcb200_status(h, &st), "mcmcb200_status");
		iter++;                                                    // :264
		if (iter > param.maxRip) { maxIterReached = true; break; } // :265-269
	}
	if (st.conflictEdges > 0 && param.tailcut) {                   // :281-311 (GPU tail-cut semantics; the CPU loop never terminates)
		uint32_t rounds = 0;
		check(mcmcb200_tailcut(h, 64, &rounds), "mcmcb200_tailcut");
	}
	C.resize(nNodes);
	check(mcmcb200_get_colors(h, C.data()), "mcmcb200_get_colors");
	hist.resize(param.nCol);
	check(mcmcb200_get_class_sizes(h, hist.data()), "mcmcb200_get_class_sizes");
}

template <typename nodeW, typename edgeW>
void ColoringMCMC_CPU<nodeW, edgeW>::saveStats(size_t it, float duration, std::ofstream & outFile) {
	const uint32_t nCol = param.nCol;
	outFile << "MCMC Colorer - CPU version - Report" << std::endl;
	outFile << "-------------------------------------------" << std::endl;
	outFile << "GRAPH INFO" << std::endl;
	outFile << "Nodes: " << nNodes << " - Edges: " << graph->getStruct()->nEdges << std::endl;
	outFile << "Max deg: " << graph->getMaxNodeDeg() << " - Min deg: " << graph->getMinNodeDeg() << " - Avg deg: " << graph->getMeanNodeDeg() << std::endl;
	outFile << "Edge probability (for randomly generated graphs): " << graph->prob << std::endl;
	outFile << "Seed: " << seed << std::endl;
	outFile << "-------------------------------------------" << std::endl;
	outFile << "EXECUTION INFO" << std::endl;
	outFile << "Repetition: " << it << std::endl;
	outFile << "Execution time: " << duration << std::endl;
	outFile << "Iteration performed: " << iter << std::endl;
	outFile << "Max iteration reached: " << (maxIterReached ? "yes" : "no") << std::endl;
	outFile << "-------------------------------------------" << std::endl;
	outFile << "Color histogram:" << std::endl;
	size_t usedCols = 0; int sum = 0;
	for (uint32_t c = 0; c < nCol; c++) { outFile << c << ": " << hist[c] << std::endl; if (hist[c]) usedCols++; sum += (int)hist[c]; }
	outFile << "Number of colors: " << nCol << " - Used colors: " << usedCols << std::endl;
	outFile << "Color ratio: " << param.numColorRatio << std::endl;
	float mean = sum / (float)nCol;
	float variance = 0;
	for (uint32_t c = 0; c < nCol; c++) { float val = (float)(size_t)hist[c]; variance += ((val - mean) * (val - mean)); }
	variance /= (float)nCol;
	float std = sqrtf(variance);
	outFile << "Average number of nodes for each color: " << mean << std::endl;
	outFile << "Variance: " << variance << std::endl;
	outFile << "StD: " << std << std::endl;
}

template <typename nodeW, typename edgeW>
void ColoringMCMC_CPU<nodeW, edgeW>::saveColor(std::ofstream & outfile) {
	for (size_t i = 0; i < C.size(); i++) outfile << i << " " << C[i] << "\n";
}

// ------------------------------------------------------------------------------------------------------------------
// ColoringLuby (cross-check): coloringLuby.cu:364-501 behaviour, log format :179-219
// ------------------------------------------------------------------------------------------------------------------
template <typename nodeW, typename edgeW>
ColoringLuby<nodeW, edgeW>::ColoringLuby(Graph<nodeW, edgeW> * g, curandState *, uint64_t seed, int device)
	: graph(g), nnodes(g->getStruct()->nNodes), seed(seed), device(device) {}

template <typename nodeW, typename edgeW>
void ColoringLuby<nodeW, edgeW>::run() {
	auto * s = graph->getStruct();
	C.assign(nnodes, 0u);
	check(mcmcb200_luby_color(s->nNodes, s->nEdges, s->cumulDegs, s->neighs, seed, device, C.data(), &numOfColors, &rounds), "mcmcb200_luby_color");
}

template <typename nodeW, typename edgeW>
void ColoringLuby<nodeW, edgeW>::saveStats(size_t it, float duration, std::ofstream & outFile) {
	outFile << "Luby Colorer - GPU version - Report" << std::endl;
	outFile << "-------------------------------------------" << std::endl;
	outFile << "GRAPH INFO" << std::endl;
	outFile << "Nodes: " << nnodes << " - Edges: " << graph->getStruct()->nEdges << std::endl;
	outFile << "Max deg: " << graph->getMaxNodeDeg() << " - Min deg: " << graph->getMinNodeDeg() << " - Avg deg: " << graph->getMeanNodeDeg() << std::endl;
	outFile << "Edge probability (for randomly generated graphs): " << graph->prob << std::endl;
	outFile << "-------------------------------------------" << std::endl;
	outFile << "EXECUTION INFO" << std::endl;
	outFile << "Repetition: " << it << std::endl;
	outFile << "Execution time: " << duration << std::endl;
	outFile << "-------------------------------------------" << std::endl;
	outFile << "Number of colors: " << numOfColors << std::endl;
	outFile << "Color histogram:" << std::endl;
	std::vector<size_t> histBins(numOfColors, 0);
	for (uint32_t val : C) histBins[val - 1]++;                            // colours are 1-based (:199)
	long sum = 0;
	for (size_t idx = 0; idx < histBins.size(); idx++) { outFile << idx << ": " << histBins[idx] << std::endl; sum += (long)histBins[idx]; }
	float mean = (int)sum / (float)numOfColors;
	float variance = 0;
	for (size_t val : histBins) variance += ((val - mean) * (val - mean));
	variance /= (float)numOfColors;
	outFile << "Average number of nodes for each color: " << mean << std::endl;
	outFile << "Variance: " << variance << std::endl;
	outFile << "StD: " << sqrtf(variance) << std::endl;
}

template <typename nodeW, typename edgeW>
void ColoringLuby<nodeW, edgeW>::saveColor(std::ofstream & outfile) {
	size_t idx = 0;
	for (uint32_t val : C) outfile << idx++ << " " << val << std::endl;
}

template class ColoringMCMC<float, float>;
template class ColoringMCMC_CPU<float, float>;
template class ColoringLuby<float, float>;
