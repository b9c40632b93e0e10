/*
 * mcmcb200.h -- C ABI of the B200-native MCMC balanced-colouring sweep (libmcmcb200.so).
 *
 * The reference (Topopiccione/MCMC_Colorer) has no FFI layer: main.cu constructs the C++ classes
 * directly.  This header is the boundary a maintainer binds instead of those classes' device code;
 * each entry point cites the reference interface it replaces (paths relative to reference src/).
 * INTEGRATION.md shows the reference-side patch (the ColoringMCMC<nodeW,edgeW> shim over this ABI).
 *
 * Conventions
 *   - every function returns 0 on success or a negative MCMCB200_E* code; nothing aborts, no exception
 *     crosses the boundary (the reference's cudaCheck prints and abort()s, GPUutils/GPUutils.h:20-26);
 *   - plain pointers and sizes only; host arrays stay owned by the caller and may be freed after the call;
 *   - one handle = one Markov chain on one CUDA device; a handle is not thread-safe, distinct handles
 *     are independent;
 *   - there is NO CPU fallback: without a CUDA device every entry point that needs one fails with
 *     MCMCB200_ENODEVICE.
 *
 * RNG contract, version 2 (replaces GPURand/curandState, GPUutils/GPURandomizer.cu:8-13,85-96, and
 * std::default_random_engine, graph_coloring/coloringMCMC_CPU.cpp:53-55): stateless Philox4x32-10, ONE call per FOUR
 * consecutive vertices:
 *   counter = (global vertex id >> 2, purpose, sweep, 0), key = (seed & 0xffffffff, seed >> 32),
 *   x = output word (global vertex id & 3);
 *   purpose 0: sweep draw, sweep = 1,2,...   UNIFORM: u = (x >> 8) * 2^-24 in [0,1)
 *                                           DYNAMIC: u = ((x >> 8) + 1) * 2^-24 in (0,1]
 *   purpose 1: initial colour, sweep = 0    colour = (x * nCol) >> 32
 *   purpose 2: Luby cross-check, sweep = round
 * so trajectories do not depend on the GPU count or on the launch shape.  (Version 1 spent a whole call per vertex and
 * threw three of the four words away.)
 */
#ifndef MCMCB200_H
#define MCMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCMCB200_ABI_VERSION 2

enum {
	MCMCB200_OK          =  0,
	MCMCB200_EINVAL      = -1,  /* bad argument (null pointer, nCol == 0, colour >= nCol, ...) */
	MCMCB200_ENODEVICE   = -2,  /* no usable CUDA device / wrong architecture */
	MCMCB200_ECUDA       = -3,  /* a CUDA runtime call failed; mcmcb200_last_cuda_error() has the text */
	MCMCB200_ENOMEM      = -4,  /* host or device allocation failed */
	MCMCB200_EUNSUPPORTED= -5,  /* configuration outside this build (e.g. nnz >= 2^32) */
	MCMCB200_ETAPE       = -6,  /* replay tape exhausted */
	MCMCB200_ESTATE      = -7   /* call out of order (e.g. sweep before init_colors) */
};

enum { MCMCB200_PROPOSAL_UNIFORM = 0,   /* coloringMCMC_standard.cu:9-82 == CPU fill_p, coloringMCMC_CPU.cpp:392-481 */
       MCMCB200_PROPOSAL_DYNAMIC = 1 }; /* coloringMCMC_balance.cu:79-143 + coloringMCMC_utils.cu:64-70 (shipped GPU default) */

/* Superset of ColoringMCMCParams (graph_coloring/coloring.h:65-74; defaults main.cu:160-168). */
typedef struct mcmcb200_params {
	uint32_t nCol;            /* number of colours, coloring.h:67 */
	float    epsilon;         /* coloring.h:70, main.cu:163 (1e-8f) */
	float    lambda;          /* coloring.h:69 -- carried for the log; dead unless HASTINGS (coloringMCMC.h:41) */
	float    numColorRatio;   /* coloring.h:68 -- carried for the log */
	float    ratioFreezed;    /* coloring.h:71 -- carried for the log */
	uint32_t tabooIteration;  /* coloring.h:72 */
	uint32_t maxRip;          /* coloring.h:66 */
	uint32_t tailcut;         /* coloring.h:73: z = max(50, n/2000) (coloringMCMC_main.cu:150-157) */
	uint32_t proposal;        /* MCMCB200_PROPOSAL_* */
	uint32_t convergence;     /* 0: stop on violating VERTICES <= z (CPU, coloringMCMC_CPU.cpp:136)
	                             1: stop on conflicting EDGES <= z (GPU, coloringMCMC_main.cu:169) */
	uint64_t seed;
	int32_t  device;          /* CUDA device ordinal; -1 = current device */
	uint32_t flags;           /* MCMCB200_FLAG_* */
	/* tuning of the source-blocked sweep; 0 = chosen from the graph size (what every production caller passes) */
	uint32_t stageCapBytes;   /* bytes of gathered neighbour colours a destination tile stages in shared memory */
	uint32_t itemBits;        /* log2 of the entries of one pass-A work item */
	uint32_t stageBuffers;    /* 1 or 2 stage buffers per pass-B CTA (2: the next tile is copied in while this one is computed) */
	uint32_t expectedSweeps;  /* how many sweeps the caller expects to run on this handle, 0 = many / unknown.  The source-blocked
	                             layout costs ~0.09 ns per directed edge to build and saves ~6 ps per edge and sweep (config 3: 0.14 s
	                             against 9 ms per sweep), so a handle that will run fewer than 16 sweeps (one short chain to a proper
	                             colouring) keeps the direct kernels, which need no layout */
} mcmcb200_params;

#define MCMCB200_FLAG_NO_FUSED_FINALIZE 1u  /* caller reduces the sweep counters across ranks itself (multi-GPU) */
#define MCMCB200_FLAG_FORCE_DIRECT      4u  /* always use the single-pass direct-gather sweep kernel */
#define MCMCB200_FLAG_FORCE_BLOCKED     8u  /* always use the source-blocked two-pass sweep (EUNSUPPORTED if a row exceeds a tile) */
#define MCMCB200_FLAG_FORCE_BINNED     16u  /* always use the degree-binned direct sweep (the path of large skewed graphs) */
#define MCMCB200_FLAG_NO_OVERLAP       32u  /* source-blocked sweep: run the two passes back to back even where they could overlap
                                               (a device shared with other work; see mcmcb200_kernel_mode) */
#define MCMCB200_FLAG_NO_EARLY_STOP     2u  /* sweeps keep advancing after C_t became proper (tape replay, benchmarking);
                                               mcmcb200_status still reports `converged` for the current colouring */

typedef struct mcmcb200_status_s {
	uint32_t sweep;              /* index t of the current colouring C_t (= sweeps performed) */
	int32_t  converged;          /* 1 iff the count selected by params.convergence of C_t is <= z */
	uint64_t conflictEdges;      /* #{ {v,u} in E : C[v]==C[u] }   (coloringMCMC_utils.cu:103-119,184-198) */
	uint64_t violatingVertices;  /* #{ v : exists u in N(v), C[u]==C[v] }   (coloringMCMC_CPU.cpp:328-351) */
	uint32_t usedColors;         /* #{ c : classSize[c] > 0 } */
	uint32_t countsSweep;        /* colouring index the two counts above refer to (== sweep unless no sweep ran yet) */
	uint64_t z;                  /* tail-cut threshold in force */
} mcmcb200_status_t;

typedef struct mcmcb200_handle mcmcb200_handle;

/* Replaces Graph(Graph*) H2D copy (graph/graphGPU.cu:210-226) + ColoringMCMC ctor allocations
 * (coloringMCMC_main.cu:5-60).  CSR as GraphStruct (graph/graph.h:37-45): cumulDegs[n+1], neighs[nnz] with
 * nnz counting both directions.  Copies the CSR to the device. */
int mcmcb200_create(mcmcb200_handle ** out, uint32_t n, uint64_t nnz, const uint32_t * cumulDegs,
                    const uint32_t * neighs, const mcmcb200_params * p);

/* Vertex-partitioned variant (one handle per GPU/rank): this handle owns global vertices [vBegin, vEnd) of an
 * nGlobal-vertex graph.  cumulDegs holds vEnd-vBegin+1 entries (any base; differences are used), neighs the
 * owned rows only, with GLOBAL neighbour ids.  The colour array is replicated (nGlobal entries).
 * vBegin must be a multiple of 256 (the sweep moves the owned colours with 16-byte vector and bulk copies): EINVAL otherwise. */
int mcmcb200_create_partition(mcmcb200_handle ** out, uint32_t nGlobal, uint32_t vBegin, uint32_t vEnd,
                              const uint32_t * cumulDegs, const uint32_t * neighs, const mcmcb200_params * p);

/* Same, adopting a CSR that already lives in DEVICE memory (no copy; the caller keeps it alive and unchanged
 * until destroy).  d_cumulDegs[0] must be 0 and d_neighs must be 32-byte aligned and readable up to the next
 * multiple of 8 elements (the sweep streams it with 256-bit loads). */
int mcmcb200_create_device_csr(mcmcb200_handle ** out, uint32_t nGlobal, uint32_t vBegin, uint32_t vEnd,
                               uint64_t nnzLocal, const uint32_t * d_cumulDegs, const uint32_t * d_neighs,
                               const mcmcb200_params * p);

/* GPU-side CSR construction from an edge list: the device twin of Graph::setupImporterNew (graph/graphCPU.cpp:112-170) for
 * inputs too large for the host loop -- self-loops dropped, the back-edge of every edge added, duplicates kept, neighbours of a
 * row in file order: array for array the CSR the reference builds.  src / dst: m entries each, HOST or DEVICE memory, ids < n.
 * The two result arrays live in device memory (device = CUDA ordinal, -1 = current), are laid out the way
 * mcmcb200_create_device_csr wants them (d_cumulDegs[0] == 0, d_neighs padded) and are released with mcmcb200_csr_free. */
int mcmcb200_csr_from_edges(uint32_t n, uint64_t m, const uint32_t * src, const uint32_t * dst, int device,
                            uint32_t ** d_cumulDegs, uint32_t ** d_neighs, uint64_t * nnz);
void mcmcb200_csr_free(uint32_t * d_cumulDegs, uint32_t * d_neighs);

void mcmcb200_destroy(mcmcb200_handle * h);

/* Replaces initColoring (coloringMCMC_utils.cu:24-33) / the ctor draw (coloringMCMC_CPU.cpp:54,61).
 * colors == NULL: Philox uniform initial colouring (RNG contract above).  Otherwise n (nGlobal) host colours,
 * each < nCol.  Resets sweep to 0, taboo to 0 and the replay position. */
int mcmcb200_init_colors(mcmcb200_handle * h, const uint32_t * colors);

/* The same two transfers in the device's own narrow colour format -- one byte per vertex for palettes of up to 256 colours, two
 * above (mcmcb200_color_bytes) -- for callers that do not need the reference's uint32 layout (coloring.h:7): a quarter of the
 * PCIe bytes, no conversion kernels.  elemBytes must equal mcmcb200_color_bytes (EINVAL otherwise); colours >= nCol are EINVAL. */
int mcmcb200_color_bytes(mcmcb200_handle * h, uint32_t * elemBytes);
int mcmcb200_init_colors_narrow(mcmcb200_handle * h, const void * colors /* [nGlobal] u8 or u16 */, uint32_t elemBytes);
int mcmcb200_get_colors_narrow(mcmcb200_handle * h, void * out /* [nGlobal] u8 or u16 */, uint32_t elemBytes);

/* Tail cutting of a multi-GPU chain (the reference is single-GPU; same result as mcmcb200_tailcut, i.e. as the sequential greedy
 * repair of coloringMCMC_utils.cu:73-101 + coloringMCMC_main.cu:271-290).  Every rank repairs the violating vertices it OWNS in
 * globally synchronised rounds; the CALLER moves the small lists between the ranks (multigpu.py: DistributedSweeper.tailcut), these
 * calls never talk to another GPU.  Handles must be created with params.tailcut.  One pass:
 *   begin   the owned vertices the reference would visit (a same-coloured neighbour with a larger id), from the violator list the
 *           converging sweep emitted (first pass) or the previous pass left; `order` = colours by ascending class size
 *   mark    the flagged vertices of ALL ranks (the ready test looks at neighbours on other ranks)
 *   round   ready = no flagged neighbour with a smaller id is still pending; the ready owned vertices take the first colour of
 *           `order` no neighbour has; returns the (vertex, colour) pairs decided, how many owned vertices are still waiting, and
 *           whether a vertex found every colour taken (`inexact`: it may now clash with a vertex outside the lists)
 *   apply   (vertex, colour) pairs decided by OTHER ranks in this round
 *   recount local counters of the repaired colouring over the owned list; the surviving violators become the next pass's list
 *   end     global counters (summed by the caller) become the chain's; exact = 0 forces a full recount at the next status */
int mcmcb200_tailcut_dist_begin(mcmcb200_handle * h, const uint32_t * order /* [nCol] */, uint32_t * outFlagged, uint32_t cap, uint32_t * count);
int mcmcb200_tailcut_dist_mark(mcmcb200_handle * h, const uint32_t * ids, uint32_t count);
int mcmcb200_tailcut_dist_round(mcmcb200_handle * h, uint32_t * outIds, uint32_t * outCols, uint32_t cap, uint32_t * processed, uint32_t * remaining,
                                uint32_t * inexact);
int mcmcb200_tailcut_dist_apply(mcmcb200_handle * h, const uint32_t * ids, const uint32_t * cols, uint32_t count);
int mcmcb200_tailcut_dist_recount(mcmcb200_handle * h, uint64_t * directedLocal, uint64_t * violLocal, uint32_t * nextFlagged);
int mcmcb200_tailcut_dist_end(mcmcb200_handle * h, uint64_t directedGlobal, uint64_t violGlobal, uint32_t exact);

/* Replay mode (parity tests): sweep k (k = 0,1,...) draws u[k*n + v] for vertex v instead of Philox.
 * sweeps == 0 or u == NULL returns to Philox. */
int mcmcb200_set_tape(mcmcb200_handle * h, const float * u, uint32_t sweeps);

/* Run up to k synchronous sweeps C_t -> C_{t+1} (the loop body of ColoringMCMC::run, coloringMCMC_main.cu:159-269,
 * == ColoringMCMC_CPU::run, coloringMCMC_CPU.cpp:136-270), asynchronously on the handle's stream, with no host
 * round trip in between.  A sweep that finds C_t converged leaves C_t current and turns the rest into no-ops. */
int mcmcb200_sweep(mcmcb200_handle * h, uint32_t k);

/* Synchronises and reports the counters of the current colouring (replaces calcConflicts' D2H + host sum,
 * coloringMCMC_utils.cu:184-198, and violation_count).  If no sweep has looked at the current colouring yet,
 * a counting pass runs first. */
int mcmcb200_status(mcmcb200_handle * h, mcmcb200_status_t * out);

int mcmcb200_get_colors(mcmcb200_handle * h, uint32_t * out /* [nGlobal] */);
/* class sizes of the current colouring (replaces the D2H + host histogram of coloringMCMC_main.cu:211-214) */
int mcmcb200_get_class_sizes(mcmcb200_handle * h, uint64_t * out /* [nCol] */);
/* per-sweep history of the counters: out[2*t] = conflictEdges(C_t), out[2*t+1] = violatingVertices(C_t), t < count */
int mcmcb200_get_history(mcmcb200_handle * h, uint64_t * out, uint32_t cap, uint32_t * count);

/* Tail cutting (coloringMCMC_main.cu:271-290 + tailCutting<<<1,1>>>, coloringMCMC_utils.cu:73-101): greedy repair
 * of the remaining conflicts, colours tried in ascending class size.  rounds returns the passes used. */
int mcmcb200_tailcut(mcmcb200_handle * h, uint32_t maxRounds, uint32_t * rounds);

/* -------- cross-check -------- */
/* Luby-style MIS colourer (replaces ColoringLuby::run, graph_coloring/coloringLuby.cu:364-501; cross-check only, not
 * performance work): greedy colour classes as maximal independent sets.  colorsOut[n] receives 1-based colours like the
 * reference's; numColors the number of classes, rounds the MIS rounds used.  Stand-alone: no handle needed. */
int mcmcb200_luby_color(uint32_t n, uint64_t nnz, const uint32_t * cumulDegs, const uint32_t * neighs, uint64_t seed, int32_t device,
                        uint32_t * colorsOut, uint32_t * numColors, uint32_t * rounds);

/* -------- parity / debug -------- */
/* Pure function of a colouring, evaluated by the sweep kernel's own counting path on the handle's graph. */
int mcmcb200_conflicts_of(mcmcb200_handle * h, const uint32_t * colors, uint64_t * edges, uint64_t * vertices);
/* Occupancy bitmask of vertex v (global id, must be owned) w.r.t. the current colouring, as built by the sweep
 * kernel: bit c of word c/32 set iff a neighbour has colour c.  words = ceil(nCol/32). */
int mcmcb200_debug_occupancy(mcmcb200_handle * h, uint32_t v, uint32_t * maskWords);
/* All owned vertices at once: masks[(v-vBegin)*words64 + w] (64-bit words; words64 = 1, 2, 4 or 8 for palettes of up to
 * 64 / 128 / 256 / 512 colours, ceil(nCol/64) above) and
 * same[v-vBegin] = number of neighbours sharing v's colour. */
int mcmcb200_debug_all_occupancy(mcmcb200_handle * h, uint64_t * masks, uint32_t * same);

/* -------- multi-GPU plumbing (one process per GPU; the exchange itself is torch.distributed / NCCL) -------- */
enum { MCMCB200_VIEW_COLORS_CUR = 0,   /* device colour array of the current colouring (element size below) */
       MCMCB200_VIEW_COLORS_NEXT = 1,  /* device colour array the next sweep writes (owned slice only) */
       MCMCB200_VIEW_COUNTERS = 2 };   /* int64[2 + nCol]: directed conflicts, violating vertices, class-size deltas */
int mcmcb200_device_view(mcmcb200_handle * h, int which, void ** devPtr, uint64_t * bytes, uint32_t * elemBytes);
/* With MCMCB200_FLAG_NO_FUSED_FINALIZE: mcmcb200_sweep(h,1) stops after the local pass; the caller all-gathers the
 * NEXT colour slices, all-reduces COUNTERS (both on mcmcb200_stream) and then calls mcmcb200_finalize_sweep. */
int mcmcb200_finalize_sweep(mcmcb200_handle * h);
/* Sliced host interface for partitioned handles: a rank uploads / downloads only the colours of the vertices it owns
 * (uint32, vEnd-vBegin entries).  After mcmcb200_init_colors_slice on every rank the caller all-gathers the
 * MCMCB200_VIEW_COLORS_CUR buffers (owned slices) and then calls mcmcb200_init_colors_finish (class sizes, validation). */
int mcmcb200_init_colors_slice(mcmcb200_handle * h, const uint32_t * ownedColors);
int mcmcb200_init_colors_finish(mcmcb200_handle * h);
int mcmcb200_get_colors_slice(mcmcb200_handle * h, uint32_t * out /* [vEnd - vBegin] */);
/* The sliced transfers in the device's narrow colour format (elemBytes = mcmcb200_color_bytes): no staging, no conversion. */
int mcmcb200_init_colors_slice_narrow(mcmcb200_handle * h, const void * ownedColors /* [vEnd - vBegin] u8 or u16 */, uint32_t elemBytes);
int mcmcb200_get_colors_slice_narrow(mcmcb200_handle * h, void * out /* [vEnd - vBegin] u8 or u16 */, uint32_t elemBytes);

/* Fused exchange (one box, NVLink/NVSwitch): no collective library call and no host in the sweep loop.  The sweep kernel
 * itself stores every finished tile's new colours into ALL ranks' colour replicas through peer pointers, and the last CTA of
 * every rank's sweep all-reduces the counters {conflicts, violations, class-size deltas} into all ranks' exchange blocks with
 * system-scope reductions over NVLink, waits for all ranks to arrive (the inter-rank barrier) and finalizes the sweep on the
 * device.  After attach, mcmcb200_sweep(h, k) runs k sweeps on N GPUs as 2k launches per rank; the ranks must issue the same
 * sequence of mcmcb200_init_colors* / mcmcb200_sweep / mcmcb200_status calls (SPMD: each of them is a collective).
 * Each rank exports three cudaIpcMemHandle_t (64 bytes each: two colour buffers, the exchange block), the caller exchanges
 * them (any transport) and attaches the full table [nRanks][3][64].  Needs the source-blocked sweep on every rank
 * (MCMCB200_EUNSUPPORTED otherwise: agree on eligibility BEFORE attaching anywhere); nRanks <= 8.  mcmcb200_ipc_detach closes
 * the mappings again (back to MCMCB200_FLAG_NO_FUSED_FINALIZE semantics: the caller exchanges). */
int mcmcb200_ipc_export(mcmcb200_handle * h, unsigned char * handles /* [3][64] */);
int mcmcb200_ipc_attach(mcmcb200_handle * h, uint32_t nRanks, uint32_t myRank, const unsigned char * handles);
int mcmcb200_ipc_detach(mcmcb200_handle * h);
int mcmcb200_stream(mcmcb200_handle * h, void ** cudaStream);
int mcmcb200_synchronize(mcmcb200_handle * h);
/* elapsed milliseconds (CUDA events on the handle's stream) of the kernels launched by the last mcmcb200_sweep */
int mcmcb200_last_sweep_ms(mcmcb200_handle * h, float * ms);
/* number of kernels this library launched on behalf of the handle so far */
int mcmcb200_launch_count(mcmcb200_handle * h, uint64_t * launches);
/* Which sweep implementation this handle runs (chosen at create from the graph size, the palette and the FORCE flags):
 * MCMCB200_MODE_DIRECT  one launch per sweep, neighbour colours gathered from the L2-resident colour array;
 * MCMCB200_MODE_BLOCKED two launches per sweep (source-blocked gather, tile sweep), one after the other;
 * MCMCB200_MODE_BLOCKED_OVERLAPPED the same two kernels running concurrently on two streams. */
#define MCMCB200_MODE_DIRECT 0
#define MCMCB200_MODE_BLOCKED 1
#define MCMCB200_MODE_BLOCKED_OVERLAPPED 2
#define MCMCB200_MODE_DIRECT_BINNED 3      /* one launch per sweep, thread / warp / CTA rows by degree (large skewed graphs) */
#define MCMCB200_MODE_WIDE_BINNED 4        /* palettes above 512 colours: the degree-binned rows with colour lists / shared-memory bitmaps
                                              instead of register masks (one table launch + one sweep launch per sweep) */
int mcmcb200_kernel_mode(mcmcb200_handle * h, int * mode);
/* Device memory held by the sweep layout built at create time, on top of the CSR and the colour buffers: the source-blocked layout
 * (about 8.7 bytes per directed edge), the degree-bin lists, the wide-palette tables and queues; 0 for the small-graph kernel. */
int mcmcb200_layout_bytes(mcmcb200_handle * h, uint64_t * bytes);

const char * mcmcb200_strerror(int code);
const char * mcmcb200_last_cuda_error(void);
int mcmcb200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCMCB200_H */
