// Host-side types of the colourer layer, source-compatible with the reference's graph_coloring/coloring.h.
#pragma once
#include <cstdint>

typedef uint32_t col;      // node colour            (reference coloring.h:7)
typedef uint32_t col_sz;   //                        (reference coloring.h:8)

// Field-for-field the reference struct (graph_coloring/coloring.h:65-74); defaults are set by main (main.cu:160-168).
struct ColoringMCMCParams {
	uint32_t maxRip;
	col_sz   nCol;
	float    numColorRatio;
	float    lambda;
	float    epsilon;
	float    ratioFreezed;
	uint32_t tabooIteration;
	bool     tailcut;
};

// Additions of the B200 build that have no slot in the reference struct; passed next to it.
struct ColoringMCMCOptions {
	uint32_t proposal = 1;     // MCMCB200_PROPOSAL_DYNAMIC: the shipped GPU variant (coloringMCMC.h:39)
	uint32_t convergence = 1;  // conflicting edges (coloringMCMC_main.cu:169)
	uint64_t seed = 0;
	int      device = -1;
	uint32_t sweepsPerCheck = 1;
};
