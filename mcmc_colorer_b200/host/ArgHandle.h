// Command-line handling of MCMC_Colorer: same option table, defaults and public fields as the reference's
// utils/ArgHandle.{h,cpp} (getopt_long table ArgHandle.cpp:29-57), plus additive options of the B200 build
// (--proposal, --device, --sweepsPerCheck, --tabooIterations alias, numColRatio accepted in (0,16]).
#pragma once
#include <cstdint>
#include <string>

class ArgHandle {
public:
	ArgHandle(int argc, char ** argv);
	virtual ~ArgHandle();
	void processCommandLine();

	std::string graphFilename;
	std::string outDir;
	double   prob;
	double   numColRatio;
	uint32_t n;
	uint32_t nCol;
	uint32_t seed;
	uint32_t verboseLevel;
	uint32_t repetitions;
	uint32_t tabooIteration;
	bool     simulate;
	bool     mcmccpu;
	bool     mcmcgpu;
	bool     lubygpu;
	bool     tailcut;
	bool     greedyff;
	bool     rebalanced_greedyff;
	std::string graphName;
	// additions
	std::string proposal;        // "dynamic" (shipped GPU default) | "uniform"
	int      device;
	uint32_t sweepsPerCheck;
	bool     quiet;

	void displayHelp();
private:
	void printLogo();
	void citeMe();
	int     argc;
	char ** argv;
};
