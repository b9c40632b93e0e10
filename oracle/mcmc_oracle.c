/* TEST INFRASTRUCTURE ONLY -- see mcmc_oracle.h.  Plain-C restatement of the reference CPU sampler
 * (and of the two GPU-only pieces that have no CPU twin: the DYNAMIC proposal and the edge-conflict
 * metric).  Compiled with -ffp-contract=off: every float operation below is one IEEE-754 binary32
 * rounding, exactly like the reference's x86-64 SSE build (no FMA contraction).
 * "parity pinned": validated against oracle/_ref (the unmodified reference) and tests/golden/. */
#include "mcmc_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10.  Third-party algorithm (Random123 v1.14, include/Random123/philox.h; Salmon, Moraes,
 * Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).  Not vendored in the reference
 * (which uses cuRAND XORWOW state, GPUutils/GPURandomizer.cu:8-13, and std::default_random_engine,
 * coloringMCMC_CPU.cpp:53); north_star replaces both with counter-based Philox.
 * ---------------------------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
		uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += PHILOX_W0; k1 += PHILOX_W1;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* RNG contract v2 (include/mcmcb200.h): one Philox call serves four consecutive vertices --
 * counter = (vertex >> 2, purpose, sweep, 0), key = (seed_lo, seed_hi); the draw of vertex v is output word v & 3.
 * purpose 0 = sweep draw, 1 = initial colour, 2 = Luby cross-check. */
uint32_t orc_draw_bits(uint64_t seed, uint32_t sweep, uint32_t vertex, uint32_t purpose) {
	uint32_t ctr[4] = { vertex >> 2, purpose, sweep, 0u };
	uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
	uint32_t out[4];
	orc_philox4x32_10(ctr, key, out);
	return out[vertex & 3u];
}

/* UNIFORM proposal: u in [0,1) like std::uniform_real_distribution<float>(0,1) (coloringMCMC_CPU.cpp:55,139);
 * DYNAMIC proposal: u in (0,1] like curand_uniform (coloringMCMC_balance.cu:120).  24 random bits, exact. */
float orc_draw_uniform(uint64_t seed, uint32_t sweep, uint32_t vertex, int proposal) {
	uint32_t m = orc_draw_bits(seed, sweep, vertex, 0u) >> 8;
	if (proposal == ORC_PROPOSAL_DYNAMIC) m += 1u;
	return (float)m * 5.9604644775390625e-8f; /* 2^-24, exact product */
}

/* initial colour uniform in [0,nCol): coloringMCMC_CPU.cpp:54,61 (uniform_int_distribution) /
 * coloringMCMC_utils.cu:24-33 (whose (int)(u*nCol) can emit nCol -- not reproduced) */
uint32_t orc_init_color(uint64_t seed, uint32_t vertex, uint32_t nCol) {
	return (uint32_t)(((uint64_t)orc_draw_bits(seed, 0u, vertex, 1u) * (uint64_t)nCol) >> 32);
}

void orc_init_colors(uint64_t seed, uint32_t vb, uint32_t ve, uint32_t nCol, uint32_t * out) {
	for (uint32_t v = vb; v < ve; v++) out[v - vb] = orc_init_color(seed, v, nCol);
}

void orc_fill_bits(uint64_t seed, uint32_t sweep, uint32_t vb, uint32_t ve, uint32_t purpose, uint32_t * out) {
	for (uint32_t v = vb; v < ve; v++) out[v - vb] = orc_draw_bits(seed, sweep, v, purpose);
}

void orc_fill_tape(uint64_t seed, uint32_t sweep, uint32_t vb, uint32_t ve, int proposal, float * u) {
	for (uint32_t v = vb; v < ve; v++) u[v - vb] = orc_draw_uniform(seed, sweep, v, proposal);
}

/* ------------------------------------------------------------------------------------------------
 * graph/graphCPU.cpp:290-404  Graph::setupRnd2
 * ---------------------------------------------------------------------------------------------- */
int orc_setup_rnd2(uint32_t n, float prob, uint32_t * cumulDegs, uint32_t * neighs, uint64_t neighsCap, uint64_t * nnzOut) {
	static uint8_t * bits = NULL; static uint64_t bitsN = 0; static uint32_t bitsFor = 0;
	const uint64_t nn = n, vecSize = nn * (nn + 1) / 2;                    /* :294-295 */
	if (neighs == NULL) {
		free(bits);
		bits = (uint8_t *)malloc(vecSize ? vecSize : 1);
		if (!bits) return -1;
		bitsN = vecSize; bitsFor = n;
		for (uint64_t i = 0; i < vecSize; i++)                             /* :307-308 */
			bits[i] = ((double)rand() / (RAND_MAX)) >= prob ? 0 : 1;
		memset(cumulDegs, 0, sizeof(uint32_t) * (nn + 1));                 /* :323 */
		uint64_t nEdges = 0, i = 0, j = 0;
		for (uint64_t k = 0; k < vecSize; k++) {                           /* :333-347 */
			if (j == i) bits[k] = 0;                                       /* no self loops, :334-335 */
			if (bits[k]) { cumulDegs[i + 1]++; cumulDegs[j + 1]++; nEdges += 2; }
			i++;
			if (i == nn) { j++; i = j; }
		}
		for (uint64_t v = 1; v < nn + 1; v++) cumulDegs[v] += cumulDegs[v - 1]; /* :359-360 */
		*nnzOut = nEdges;
		return 0;
	}
	if (!bits || bitsFor != n || bitsN != vecSize) return -2;
	uint64_t * tempDegs = (uint64_t *)calloc(nn ? nn : 1, sizeof(uint64_t));
	if (!tempDegs) return -1;
	uint64_t i = 0, j = 0;
	for (uint64_t k = 0; k < vecSize; k++) {                               /* :374-390 */
		if (bits[k]) {
			uint64_t idx = cumulDegs[j] + tempDegs[j];
			if (idx >= neighsCap) { free(tempDegs); return -3; }
			neighs[idx] = (uint32_t)i; tempDegs[j]++;
			idx = cumulDegs[i] + tempDegs[i];
			if (idx >= neighsCap) { free(tempDegs); return -3; }
			neighs[idx] = (uint32_t)j; tempDegs[i]++;
		}
		i++;
		if (i == nn) { j++; i = j; }
	}
	free(tempDegs); free(bits); bits = NULL; bitsN = 0; bitsFor = 0;
	*nnzOut = cumulDegs[n];
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * counting
 * ---------------------------------------------------------------------------------------------- */
/* coloringMCMC_CPU.cpp:328-351 */
uint64_t orc_violation_count(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, const uint32_t * colors,
                             uint32_t vb, uint32_t ve, uint8_t * viol) {
	(void)n;
	uint64_t total = 0;
	for (uint32_t i = vb; i < ve; i++) {
		const uint32_t nodeColor = colors[i];                              /* :335 */
		uint64_t nodeViolations = 0;
		for (uint32_t k = cumulDegs[i]; k < cumulDegs[i + 1]; k++)         /* :342-343 */
			nodeViolations += (nodeColor == colors[neighs[k]]);
		if (viol) viol[i] = nodeViolations > 0;
		if (nodeViolations > 0) total++;                                   /* :345-348 */
	}
	return total;
}

/* coloringMCMC_utils.cu:103-119 (+ host sum :194-197) */
uint64_t orc_conflict_edges(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, const uint32_t * colors,
                            uint32_t vb, uint32_t ve) {
	(void)n;
	uint64_t total = 0;
	for (uint32_t idx = vb; idx < ve; idx++) {
		const uint32_t nodeCol = colors[idx];
		for (uint32_t k = cumulDegs[idx]; k < cumulDegs[idx + 1]; k++)
			total += (colors[neighs[k]] == nodeCol) && (idx < neighs[k]);  /* :115 */
	}
	return total;
}

/* coloringMCMC_CPU.cpp:361-383 */
uint32_t orc_occupancy(uint32_t v, const uint32_t * cumulDegs, const uint32_t * neighs, const uint32_t * colors,
                       uint32_t nCol, uint8_t * occ) {
	memset(occ, 0, nCol);                                                  /* :366 (freeColors all 1) */
	for (uint32_t k = cumulDegs[v]; k < cumulDegs[v + 1]; k++)             /* :376-379 */
		occ[colors[neighs[k]]] = 1;
	uint32_t freeCnt = 0;
	for (uint32_t c = 0; c < nCol; c++) freeCnt += !occ[c];                /* :382 */
	return freeCnt;
}

/* coloringMCMC_CPU.cpp:392-481 (baseline branch; the linear/exp variants are commented out there) */
void orc_fill_p_uniform(uint32_t nCol, float eps, const uint8_t * occ, uint32_t ownColor, float * p) {
	uint32_t Zvcomp = 0;
	for (uint32_t c = 0; c < nCol; c++) Zvcomp += !occ[c];
	const size_t Zv = (size_t)nCol - Zvcomp;
	const int viol = occ[ownColor];                                        /* Cviols[v], :400 */
	if (viol && Zvcomp > 0) {
		const float freeW = (1.0f - eps * Zv) / (float)Zvcomp;             /* :416 */
		for (uint32_t c = 0; c < nCol; c++) p[c] = occ[c] ? eps : freeW;   /* :414-420 */
	} else {
		/* not violating (:472-478), or violating with no free colour (:402-411): stay put */
		const float stay = 1.0f - (nCol - 1) * eps;                        /* :406 / :474 */
		for (uint32_t c = 0; c < nCol; c++) p[c] = (c == ownColor) ? stay : eps;
	}
}

/* coloringMCMC_utils.cu:64-70 genDynamicDistribution */
static void dynamic_distribution(uint32_t n, uint32_t nCol, const uint32_t * hist, float * dist) {
	for (uint32_t c = 0; c < nCol; c++)
		dist[c] = (1 - ((float)hist[c] / (float)n)) / (float)(nCol - 1);   /* :69 */
}

uint64_t orc_sweep(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, uint32_t nCol, float eps,
                   uint32_t tabooIteration, int proposal, const uint32_t * C, uint32_t * Cstar, uint32_t * taboo,
                   const float * u, const uint32_t * hist, uint32_t vb, uint32_t ve) {
	uint64_t overflow = 0;
	uint8_t * occ = (uint8_t *)malloc(nCol ? nCol : 1);
	float * p = (float *)malloc(sizeof(float) * (nCol ? nCol : 1));
	float * dist = NULL;
	if (proposal == ORC_PROPOSAL_DYNAMIC) {
		dist = (float *)malloc(sizeof(float) * (nCol ? nCol : 1));
		dynamic_distribution(n, nCol, hist, dist);
	}
	for (uint32_t v = vb; v < ve; v++) {
		const uint32_t own = C[v];
		/* TABOO gate: coloringMCMC_CPU.cpp:496-501 == coloringMCMC_standard.cu:14-20 (keeps C[v]) */
		if (taboo && taboo[v] > 0) { taboo[v]--; Cstar[v] = own; continue; }
		const uint32_t Zp = orc_occupancy(v, cumulDegs, neighs, C, nCol, occ);
		uint32_t idx;
		if (proposal == ORC_PROPOSAL_UNIFORM) {
			orc_fill_p_uniform(nCol, eps, occ, own, p);                    /* :195 */
			float cdf = 0;                                                 /* :505 */
			for (idx = 0; idx < nCol; idx++) {                             /* :510-514 */
				cdf += p[idx];
				if (cdf > u[v]) break;
			}
			if (idx >= nCol) { idx = nCol - 1; overflow++; }               /* contract vs :517-520 */
		} else {
			/* coloringMCMC_balance.cu:101-139 */
			if (!Zp) { Cstar[v] = own; continue; }                         /* :111-115 no draw, taboo untouched */
			float reminder = 0;
			for (uint32_t i = 0; i < nCol; i++)                            /* :104-107 */
				if (occ[i]) reminder += (dist[i] - eps);
			const float denomReminder = (float)Zp;                         /* :109 */
			const float randnum = u[v];
			float threshold = 0, q;
			uint32_t i = 0;
			if (occ[own]) {                                                /* :122-129 */
				do {
					float r = reminder / denomReminder;
					q = occ[i] ? eps : (dist[i] + r);
					threshold += q;
					i++;
				} while (threshold < randnum && i < nCol);
			} else {                                                       /* :130-136 */
				/* the reference GPU kernel as nvcc builds it (default -fmad=true) evaluates 1.0f - (nCol-1)*epsilon with ONE
				 * rounding: `FFMA R, -R(nCol-1), R(eps), 1` in the SASS of selectStarColoringBalanceDynamic (oracle/_ref_gpu);
				 * tests/test_gpu_refgpu.py pins this against the real kernel at eps = 1e-4, nCol = 89, where the two differ */
				const float stayW = fmaf(-(float)(nCol - 1), eps, 1.0f);
				do {
					q = (own == i) ? stayW : eps;
					threshold += q;
					i++;
				} while (threshold < randnum && i < nCol);
			}
			if (i == nCol && threshold < randnum) overflow++;
			idx = i - 1;                                                   /* :138 */
		}
		Cstar[v] = idx;                                                    /* :523 */
		if (taboo) taboo[v] = (idx == own) * tabooIteration;               /* :526 / _balance.cu:141 */
	}
	free(occ); free(p); free(dist);
	return overflow;
}

void orc_class_sizes(uint32_t n, const uint32_t * colors, uint32_t nCol, uint32_t * hist) {
	memset(hist, 0, sizeof(uint32_t) * nCol);
	for (uint32_t v = 0; v < n; v++) hist[colors[v]]++;                    /* coloringMCMC_CPUutils.cpp:88 */
}

void orc_color_stats(uint32_t n, uint32_t nCol, const uint32_t * hist, float prob, orc_color_stats_t * out) {
	memset(out, 0, sizeof(*out));
	/* coloringMCMC_CPUutils.cpp:90-98 */
	int sum = 0; uint32_t used = 0;
	for (uint32_t c = 0; c < nCol; c++) { sum += (int)hist[c]; if (hist[c]) used++; }
	float mean = sum / (float)nCol;                                        /* :94 */
	float variance = 0;
	for (uint32_t c = 0; c < nCol; c++) {                                  /* :96 */
		float val = (float)(size_t)hist[c];
		variance += ((val - mean) * (val - mean));
	}
	variance /= (float)nCol;                                               /* :97 */
	out->usedColors = used; out->meanCPU = mean; out->varianceCPU = variance; out->stdCPU = sqrtf(variance);

	/* coloringMCMC_prints.cu:140-174 */
	int max_i = 0, min_i = (int)n, max_c = 0, min_c = (int)n;              /* :141-142 */
	float average = (float)n / nCol, var2 = 0, bal = 0;                    /* :148 */
	for (uint32_t i = 0; i < nCol; i++) {
		if (hist[i] > 0) {                                                 /* :152-163 */
			if ((int)hist[i] > max_c) { max_i = (int)i; max_c = (int)hist[i]; }
			if ((int)hist[i] < min_c) { min_i = (int)i; min_c = (int)hist[i]; }
			bal += powf(hist[i] - average, 2.f);
		}
	}
	bal /= (n * prob);                                                     /* :166 */
	bal = sqrtf(bal);                                                      /* :167 */
	for (uint32_t i = 0; i < nCol; i++) var2 += powf((hist[i] - average), 2.f); /* :169-171 */
	var2 /= nCol;                                                          /* :172 */
	out->mostUsed = (uint32_t)max_i; out->mostUsedCount = (uint32_t)max_c;
	out->leastUsed = (uint32_t)min_i; out->leastUsedCount = (uint32_t)min_c;
	out->averageGPU = average; out->varianceGPU = var2; out->stdGPU = sqrtf(var2); out->balancingIndex = bal;
}

uint32_t orc_run(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, uint32_t nCol, float eps,
                 uint32_t tabooIteration, int proposal, uint64_t seed, uint32_t maxRip, uint64_t z,
                 uint32_t * colors, uint64_t * finalCount, int * maxIterReached) {
	uint32_t * C = colors;
	uint32_t * Cstar = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t * taboo = tabooIteration ? (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t)) : NULL;
	uint32_t * hist = (uint32_t *)malloc(sizeof(uint32_t) * (nCol ? nCol : 1));
	float * u = (float *)malloc(sizeof(float) * (n ? n : 1));
	uint32_t sweeps = 0; int hitMax = 0; uint64_t count;
	for (;;) {
		/* loop test: violating vertices (CPU :136) or conflicting edges (GPU _main.cu:163-170) */
		count = (proposal == ORC_PROPOSAL_UNIFORM)
			? orc_violation_count(n, cumulDegs, neighs, C, 0, n, NULL)
			: orc_conflict_edges(n, cumulDegs, neighs, C, 0, n);
		if (count <= z) break;
		if (sweeps >= maxRip) { hitMax = 1; break; }
		orc_fill_tape(seed, sweeps + 1, 0, n, proposal, u);
		if (proposal == ORC_PROPOSAL_DYNAMIC) orc_class_sizes(n, C, nCol, hist);
		orc_sweep(n, cumulDegs, neighs, nCol, eps, tabooIteration, proposal, C, Cstar, taboo, u, hist, 0, n);
		memcpy(C, Cstar, sizeof(uint32_t) * n);                            /* swap, :259 */
		sweeps++;
	}
	if (finalCount) *finalCount = count;
	if (maxIterReached) *maxIterReached = hitMax;
	free(Cstar); free(taboo); free(hist); free(u);
	return sweeps;
}

/* ------------------------------------------------------------------------------------------------
 * tail cutting (GPU semantics): coloringMCMC_main.cu:271-290, coloringMCMC_utils.cu:73-101
 * Colour order: ascending class size; ties by ascending colour index (the reference's std::sort at
 * _main.cu:276 leaves tie order to the library -- the contract fixes it).
 * ---------------------------------------------------------------------------------------------- */
uint32_t orc_tailcut(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, uint32_t nCol, uint32_t * colors,
                     uint32_t maxRounds, uint64_t * conflictEdgesOut) {
	uint32_t * hist = (uint32_t *)malloc(sizeof(uint32_t) * (nCol ? nCol : 1));
	uint32_t * order = (uint32_t *)malloc(sizeof(uint32_t) * (nCol ? nCol : 1));
	uint8_t * occ = (uint8_t *)malloc(nCol ? nCol : 1);
	orc_class_sizes(n, colors, nCol, hist);                                /* _main.cu:272-274 */
	for (uint32_t i = 0; i < nCol; i++) order[i] = i;
	for (uint32_t i = 1; i < nCol; i++) {                                  /* insertion sort == stable sort by hist */
		uint32_t x = order[i]; uint32_t j = i;
		while (j > 0 && hist[order[j - 1]] > hist[x]) { order[j] = order[j - 1]; j--; }
		order[j] = x;
	}
	uint32_t rounds = 0;
	uint64_t conflicts = orc_conflict_edges(n, cumulDegs, neighs, colors, 0, n);
	while (conflicts > 0 && rounds < maxRounds) {                          /* _main.cu:279 */
		/* conflictCounter flags: vertex idx has a same-coloured neighbour u > idx (_utils.cu:115), frozen before the pass */
		uint8_t * flag = (uint8_t *)calloc(n ? n : 1, 1);
		for (uint32_t idx = 0; idx < n; idx++)
			for (uint32_t k = cumulDegs[idx]; k < cumulDegs[idx + 1]; k++)
				if (colors[neighs[k]] == colors[idx] && idx < neighs[k]) { flag[idx] = 1; break; }
		uint64_t resolved = 0;
		for (uint32_t idx = 0; idx < n && resolved < conflicts; idx++) {   /* _utils.cu:78 */
			if (!flag[idx]) continue;
			resolved++;
			uint32_t nodeCol = colors[idx];
			memset(occ, 0, nCol);
			for (uint32_t k = cumulDegs[idx]; k < cumulDegs[idx + 1]; k++) /* :88-89, current (in-place) colouring */
				occ[colors[neighs[k]]] = 1;
			uint32_t j = 0;
			while (occ[nodeCol] && j < nCol) { nodeCol = order[j]; j++; }  /* :91-95 */
			colors[idx] = nodeCol;                                         /* :97 */
		}
		free(flag);
		conflicts = orc_conflict_edges(n, cumulDegs, neighs, colors, 0, n);
		rounds++;
	}
	if (conflictEdgesOut) *conflictEdgesOut = conflicts;
	free(hist); free(order); free(occ);
	return rounds;
}
