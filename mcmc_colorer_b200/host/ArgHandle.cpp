#include "ArgHandle.h"

#include <getopt.h>
#include <sys/stat.h>
#include <sys/types.h>

#include <cstdlib>
#include <ctime>
#include <iostream>
#include <vector>

namespace {
std::vector<std::string> split_str(const std::string & s, const std::string & delims) {   // utils/miscUtils.cpp:30-41
	std::vector<std::string> out;
	size_t cur, next = (size_t)-1;
	do {
		cur = next + 1;
		next = s.find_first_of(delims, cur);
		if (s.substr(cur, next - cur) != "") out.push_back(s.substr(cur, next - cur));
	} while (next != std::string::npos);
	return out;
}
[[noreturn]] void die(const std::string & msg) { std::cout << msg << std::endl; exit(-1); }
}

ArgHandle::ArgHandle(int argc, char ** argv)
	: graphFilename(""), outDir(""), prob(0.0), numColRatio(0.0), n(0), nCol(0), seed(0), verboseLevel(0), repetitions(1),
	  tabooIteration(0), simulate(false), mcmccpu(false), mcmcgpu(false), lubygpu(false), tailcut(false), greedyff(false),
	  rebalanced_greedyff(false), proposal("dynamic"), device(-1), sweepsPerCheck(1), quiet(false), argc(argc), argv(argv) {}

ArgHandle::~ArgHandle() {}

void ArgHandle::processCommandLine() {
	char const * short_options = "g:o:s:n:12345k:r:t:lR:S:v:hMP:D:C:q";
	const struct option long_options[] = {
		{"graph", required_argument, 0, 'g'}, {"outDir", required_argument, 0, 'o'},
		{"simulate", required_argument, 0, 's'}, {"nodes", required_argument, 0, 'n'},
		{"mcmccpu", no_argument, 0, '1'}, {"mcmcgpu", no_argument, 0, '2'}, {"lubygpu", no_argument, 0, '3'},
		{"grdffgpu", no_argument, 0, '4'}, {"vffgpu", no_argument, 0, '5'},
		{"nCol", required_argument, 0, 'k'}, {"numColRatio", required_argument, 0, 'r'},
		{"tabooIteration", required_argument, 0, 't'}, {"tabooIterations", required_argument, 0, 't'},   // README spelling too
		{"tailcut", no_argument, 0, 'l'},
		{"repet", required_argument, 0, 'R'}, {"seed", required_argument, 0, 'S'},
		{"verbose-level", required_argument, 0, 'v'}, {"help", no_argument, 0, 'h'}, {"cite-me", no_argument, 0, 'M'},
		{"proposal", required_argument, 0, 'P'}, {"device", required_argument, 0, 'D'},
		{"sweepsPerCheck", required_argument, 0, 'C'}, {"quiet", no_argument, 0, 'q'},
		{0, 0, 0, 0}};
	for (int i = 1; i < argc; i++) if (std::string(argv[i]) == "--quiet" || std::string(argv[i]) == "-q") quiet = true;
	if (!quiet) printLogo();
	auto toInt = [](const char * s, const char * err) { try { return std::stoi(s); } catch (...) { die(err); } };
	auto toDbl = [](const char * s, const char * err) { try { return std::stod(s); } catch (...) { die(err); } };
	while (1) {
		int idx = 0;
		int c = getopt_long(argc, argv, short_options, long_options, &idx);
		if (c == -1) break;
		switch (c) {
		case 'g': graphFilename = optarg; break;
		case 'o': outDir = optarg; break;
		case 's': {
			simulate = true;
			double t = toDbl(optarg, "Argument missing: specify the probabilty for positive class.");
			if ((t < 0) | (t > 1)) die("Simulation: probabilty of positive class must be 0 < prob < 1.");
			prob = t; break; }
		case 'n': { int t = toInt(optarg, "n must be a positive integer."); if (t < 1) die("n must be a positive integer."); n = t; break; }
		case '1': mcmccpu = true; break;
		case '2': mcmcgpu = true; break;
		case '3': lubygpu = true; break;
		case '4': greedyff = true; break;
		case '5': rebalanced_greedyff = true; break;
		case 'k': { int t = toInt(optarg, "nCol must be a positive integer."); if (t < 1) die("nCol must be a positive integer."); nCol = t; break; }
		case 'r': {
			// the reference accepts [1,16] (ArgHandle.cpp:148-156); BASELINE config 5 needs ratios below 1 (nCol > maxDeg)
			double t = toDbl(optarg, "Argument missing: specify color ratio 0 < numColRatio <= 16.");
			if (!(t > 0.0) || t > 16.0) die("Color ratio must be 0 < numColRatio <= 16.0.");
			numColRatio = t; break; }
		case 't': { int t = toInt(optarg, "tabooIteration must be a positive integer."); if (t < 1) die("tabooIteration must be a positive integer."); tabooIteration = t; break; }
		case 'l': tailcut = true; break;
		case 'R': { int t = toInt(optarg, "repetitions must be a positive integer."); if (t < 1) die("repetitions must be a positive integer."); repetitions = t; break; }
		case 'S': seed = toInt(optarg, "seed argument must be integer."); break;
		case 'v': verboseLevel = toInt(optarg, "verbose-level argument must be integer."); break;
		case 'h': displayHelp(); exit(0);
		case 'M': citeMe(); exit(0);
		case 'P': proposal = optarg; if (proposal != "dynamic" && proposal != "uniform") die("--proposal must be dynamic or uniform."); break;
		case 'D': device = toInt(optarg, "device must be an integer."); break;
		case 'C': { int t = toInt(optarg, "sweepsPerCheck must be a positive integer."); if (t < 1) die("sweepsPerCheck must be a positive integer."); sweepsPerCheck = t; break; }
		case 'q': quiet = true; break;
		default: break;
		}
	}
	if ((!simulate) && graphFilename.empty()) die("Graph file undefined (--graph). Specify a graph file or enable simulation mode.");
	if ((!mcmccpu) && (!mcmcgpu) && (!lubygpu) && (!greedyff) && (!rebalanced_greedyff)) {     // ArgHandle.cpp:247-250
		std::cout << "No coloring algorithm specified: enabling MCMC CPU by default (--mcmccpu | --mcmcgpu | --lubygpu)" << std::endl;
		mcmccpu = true;
	}
	if (simulate && (n == 0)) die("Simualtion enabled: specify the number of nodes (-n).");
	if ((mcmccpu || mcmcgpu) && (nCol == 0) && !quiet)
		std::cout << "No number of colors specified (--nCol): enabling default value: maxDeg / numColRatio." << std::endl;
	if (numColRatio == 0.0) {
		if (!quiet) std::cout << "Using default color ratio (1.0) (--numColRatio)" << std::endl;
		numColRatio = 1.0;
	}
	if (seed == 0) {                                               // ArgHandle.cpp:272-276: srand only on this path
		seed = (uint32_t)time(NULL);
		std::cout << "No seed specified. Generating a random seed: " << seed << " (--seed)." << std::endl;
		srand(seed);
	}
	if (verboseLevel > 3) verboseLevel = 3;
	if (!simulate) {                                               // ArgHandle.cpp:289-301
		std::vector<std::string> just = split_str(graphFilename, "/\\");
		std::vector<std::string> parts = split_str(just[just.size() - 1], ".");
		graphName = parts[0];
		for (size_t i = 1; i + 1 < parts.size(); i++) graphName += "." + parts[i];
	} else {
		graphName = std::to_string(n) + "_" + std::to_string(prob) + "_" + std::to_string(numColRatio);
	}
	if (outDir.empty()) {
		outDir = graphName + "_out";
		if (!quiet) std::cout << "No output directory defined. Saving to: " << outDir << " (--outDir)." << std::endl;
	}
	mkdir(outDir.c_str(), 0775);                                   // the reference only creates the default dir (:303-307)
}

void ArgHandle::displayHelp() {
	std::cout << "Usage: " << std::endl << "    " << argv[0] << " [options]" << std::endl << std::endl;
	std::cout << "Options:" << std::endl;
	std::cout << "    --help               Print this help." << std::endl;
	std::cout << "  Dataset" << std::endl;
	std::cout << "    --graph file.txt     Input graph specified as a list of edges (mandatory if not in simulation mode)." << std::endl;
	std::cout << "    --outDir             Output directory." << std::endl;
	std::cout << "    --simulate P         Enable simulation of a random Erdos graph. Edges are generated with probability" << std::endl;
	std::cout << "                         P (0 < P < 1). -n parameter is mandatory." << std::endl;
	std::cout << "    -n N                 Number of nodes to be generated. Enabled only if --simulate option is specified." << std::endl;
	std::cout << "  Coloring algorithm" << std::endl;
	std::cout << "    --mcmccpu            MCMC colorer with the CPU colorer's semantics (uniform proposal, CPU log format), run on the GPU." << std::endl;
	std::cout << "    --mcmcgpu            Enables MCMC GPU colorer." << std::endl;
	std::cout << "    --lubygpu            Enables Luby GPU colorer (cross-check)." << std::endl;
	std::cout << "    --grdffgpu, --vffgpu Not part of this build (different algorithms; see DESIGN.md)." << std::endl;
	std::cout << "  Coloring options (only for MCMC CPU and MCMC GPU)" << std::endl;
	std::cout << "    --nCol N             Number of colors" << std::endl;
	std::cout << "    --numColRatio N.N    Optional divider for number of colors (default = 1.0, 0 < numColRatio <= 16.0)" << std::endl;
	std::cout << "    --tabooIterations N  Optional number of iteration for the taboo strategy" << std::endl;
	std::cout << "    --tailcut            Enables tail cutting strategy (default = disabled)" << std::endl;
	std::cout << "    --proposal P         dynamic (default, shipped GPU variant) | uniform (CPU colorer's proposal)" << std::endl;
	std::cout << "    --device D           CUDA device ordinal (default: current)" << std::endl;
	std::cout << "    --sweepsPerCheck K   sweeps launched back to back between two convergence reads (default 1)" << std::endl;
	std::cout << "  General options" << std::endl;
	std::cout << "    --repet N            Number of repetitions for each coloring (optional, default = 1)." << std::endl;
	std::cout << "    --seed N             Seed for random number generator(optional, default = random)." << std::endl;
	std::cout << "    --quiet              No banner." << std::endl << std::endl;
}

void ArgHandle::citeMe() {
	std::cout << std::endl << "This work can be cited by adding the following items to your bibliografy:" << std::endl << std::endl;
	std::cout << "@inproceedings{colorerGbR2019," << std::endl;
	std::cout << "	author    = {Conte, Donatello and Grossi, Giuliano and Lanzarotti, Raffaella and Lin, Jianyi and Petrini, Alessandro}," << std::endl;
	std::cout << "	title     = {A parallel MCMC algorithm for the Balanced Graph Coloring problem}," << std::endl;
	std::cout << "	booktitle = {IAPR International workshop on Graph-Based Representation in Pattern Recognition, Tours, France}," << std::endl;
	std::cout << "	year      = {2019}," << std::endl << "	month     = {Jul}," << std::endl << "	day       = {19-21}" << std::endl << "}" << std::endl << std::endl;
}

void ArgHandle::printLogo() {
	std::cout << "_______________________________________________________________" << std::endl;
	std::cout << "  MCMC Colorer -- B200 (sm_100a) build of the MCMC balanced colouring sampler" << std::endl;
	std::cout << "  after PhuseLab / AnacletoLab MCMC_Colorer (Conte, Grossi, Lanzarotti, Lin, Petrini, GbR 2019)" << std::endl;
	std::cout << "  '--help' for the list of command line options, '--cite-me' for citation info" << std::endl;
	std::cout << "_______________________________________________________________" << std::endl << std::endl;
}
