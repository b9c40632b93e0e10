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
