/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's MCMC balanced-colouring
 * sweep.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker.  The product (libmcmcb200.so) never links it.
 *
 * Parity pin: this port is checked function by function against oracle/_ref/libmcmc_ref.so (the
 * UNMODIFIED reference CPU colourer compiled from /root/reference, see oracle/Makefile) in
 * tests/test_oracle.py, and against the committed fixtures in tests/golden/ that were
 * generated from that library (tests/golden/make_golden.py).  The reference's own test-suite holds
 * no golden vectors for this path, so those are the pins.
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it restates.
 */
#ifndef MCMC_ORACLE_H
#define MCMC_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PROPOSAL_UNIFORM = 0, ORC_PROPOSAL_DYNAMIC = 1 };

/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 philox.h).  Third-party algorithm restated
 * from the publication; pinned by the Random123 known-answer vectors in tests/test_oracle.py. */
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* draw conventions shared with the device code (include/mcmcb200.h "RNG contract") */
uint32_t orc_draw_bits(uint64_t seed, uint32_t sweep, uint32_t vertex, uint32_t purpose);
float    orc_draw_uniform(uint64_t seed, uint32_t sweep, uint32_t vertex, int proposal);
uint32_t orc_init_color(uint64_t seed, uint32_t vertex, uint32_t nCol);
void     orc_init_colors(uint64_t seed, uint32_t vb, uint32_t ve, uint32_t nCol, uint32_t * out /* [ve-vb] */);
void     orc_fill_bits(uint64_t seed, uint32_t sweep, uint32_t vb, uint32_t ve, uint32_t purpose, uint32_t * out /* [ve-vb] */);
void     orc_fill_tape(uint64_t seed, uint32_t sweep, uint32_t vb, uint32_t ve, int proposal, float * u /* [ve-vb] */);

/* graph/graphCPU.cpp:290-404  Graph::setupRnd2 -- libc rand() exact Erdos-Renyi generator.
 * Call once with neighs==NULL to get *nnzOut (cumulDegs must hold n+1), then with storage.
 * The libc rand() state is consumed exactly like the reference: n(n+1)/2 calls. */
int      orc_setup_rnd2(uint32_t n, float prob, uint32_t * cumulDegs, uint32_t * neighs, uint64_t neighsCap, uint64_t * nnzOut);

/* graph_coloring/coloringMCMC_CPU.cpp:328-351 violation_count: #vertices with a same-coloured neighbour */
uint64_t orc_violation_count(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, const uint32_t * colors,
                             uint32_t vb, uint32_t ve, uint8_t * viol /* [n] or NULL */);
/* graph_coloring/coloringMCMC_utils.cu:103-119 conflictCounter (+ the sum of :184-198): #edges {v,u}, v<u, same colour */
uint64_t orc_conflict_edges(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, const uint32_t * colors,
                            uint32_t vb, uint32_t ve);
/* graph_coloring/coloringMCMC_CPU.cpp:361-383 count_free_colors: occ[c]=1 iff a neighbour has colour c; returns #free */
uint32_t orc_occupancy(uint32_t v, const uint32_t * cumulDegs, const uint32_t * neighs, const uint32_t * colors,
                       uint32_t nCol, uint8_t * occ /* [nCol] */);
/* fill_p (coloringMCMC_CPU.cpp:392-481) for one vertex; p[nCol] */
void     orc_fill_p_uniform(uint32_t nCol, float eps, const uint8_t * occ, uint32_t ownColor, float * p);

/* One synchronous sweep C -> Cstar over vertices [vb,ve).
 *   proposal UNIFORM : coloringMCMC_CPU.cpp:183-204 (count_free_colors, fill_p, extract_new_color) == GPU _standard.cu:9-82
 *   proposal DYNAMIC : coloringMCMC_balance.cu:79-143 + coloringMCMC_utils.cu:64-70 (needs hist of C over ALL n vertices)
 * taboo may be NULL when tabooIteration==0.  u[v] is the draw of vertex v (index by global vertex id).
 * Returns the number of CDF-walk overflows (contract: idx = nCol-1). */
uint64_t orc_sweep(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, uint32_t nCol, float eps,
                   uint32_t tabooIteration, int proposal, const uint32_t * C, uint32_t * Cstar, uint32_t * taboo,
                   const float * u, const uint32_t * hist /* [nCol], DYNAMIC only */, uint32_t vb, uint32_t ve);

void     orc_class_sizes(uint32_t n, const uint32_t * colors, uint32_t nCol, uint32_t * hist);

typedef struct {
	uint32_t usedColors;
	uint32_t mostUsed, mostUsedCount, leastUsed, leastUsedCount;
	float    meanCPU, varianceCPU, stdCPU;          /* coloringMCMC_CPUutils.cpp:94-98 */
	float    averageGPU, varianceGPU, stdGPU, balancingIndex; /* coloringMCMC_prints.cu:146-174 */
} orc_color_stats_t;
void     orc_color_stats(uint32_t n, uint32_t nCol, const uint32_t * hist, float prob, orc_color_stats_t * out);

/* Free-running chain with the Philox draw convention: loop of coloringMCMC_CPU.cpp:136-270 (metric = violating vertices)
 * or coloringMCMC_main.cu:159-269 (metric = conflicting edges, proposal DYNAMIC).  Returns sweeps performed;
 * colors is updated in place to the colouring whose count was tested last. */
uint32_t orc_run(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, uint32_t nCol, float eps,
                 uint32_t tabooIteration, int proposal, uint64_t seed, uint32_t maxRip, uint64_t z,
                 uint32_t * colors, uint64_t * finalCount, int * maxIterReached);

/* Tail cutting, GPU semantics: coloringMCMC_main.cu:271-290 + coloringMCMC_utils.cu:73-101.
 * Greedy sequential repair of conflicted vertices in ascending vertex order, trying colours in ascending class size.
 * Returns the number of repair rounds; colors updated in place. */
uint32_t orc_tailcut(uint32_t n, const uint32_t * cumulDegs, const uint32_t * neighs, uint32_t nCol, uint32_t * colors,
                     uint32_t maxRounds, uint64_t * conflictEdgesOut);

#ifdef __cplusplus
}
#endif
#endif
