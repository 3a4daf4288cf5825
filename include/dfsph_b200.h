/* dfsph_b200.h -- C ABI of the B200-native DFSPH hot path (neighbourhood search + DFSPH pressure-solver loop).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  One `Real` per shared object,
 * matching the reference's compile-time Real (SPlisHSPlasH/Common.h:6-24):
 *     libdfsph_b200_f32.so  Real = float   -> behaviour of the reference's float+AVX solver variant
 *                                             (SPlisHSPlasH/DFSPH/TimeStepDFSPH.cpp:733-1100)
 *     libdfsph_b200_f64.so  Real = double  -> behaviour of the reference's scalar solver variant
 *                                             (SPlisHSPlasH/DFSPH/TimeStepDFSPH.cpp:1104-1421)
 * All Real* arguments are `float*` in the f32 library and `double*` in the f64 library (dfsph_b200_sizeof_real()).
 *
 * Every entry point returns 0 on success or a negative dfsph_b200_status; dfsph_b200_last_error() gives text.
 * A context is single-caller (like the reference: TimeStep::step() is called from one thread,
 * Simulator/SimulatorBase.cpp:920,970) and owns all device memory; host pointers are never retained past a call.
 * There is NO CPU fallback: without a CUDA device dfsph_b200_create() fails with DFSPH_B200_ERR_CUDA.
 *
 * Reference interfaces each entry point replaces are cited per function (paths relative to the reference root).
 */
#ifndef DFSPH_B200_H
#define DFSPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dfsph_b200_ctx dfsph_b200_ctx;

typedef enum {
    DFSPH_B200_OK = 0,
    DFSPH_B200_ERR_INVALID = -1,      /* bad argument / call order */
    DFSPH_B200_ERR_CUDA = -2,         /* CUDA runtime error (sticky in the ctx) */
    DFSPH_B200_ERR_CAPACITY = -3,     /* neighbour-list / particle capacity exceeded (raise it in the config) */
    DFSPH_B200_ERR_UNSUPPORTED = -4,  /* feature outside the hot-path scope (dynamic bodies, 2-D, multiphase, ...) */
    DFSPH_B200_ERR_COMM = -5          /* multi-GPU exchange failed */
} dfsph_b200_status;

/* Kernel ids use the reference's enum values: Simulation "kernel" / "gradKernel" of the 3-D build
 * (SPlisHSPlasH/Simulation.cpp:221-228, 241-248). */
enum { DFSPH_B200_KERNEL_CUBIC = 0, DFSPH_B200_KERNEL_WENDLAND_QUINTIC_C2 = 1, DFSPH_B200_KERNEL_POLY6 = 2,
       DFSPH_B200_KERNEL_SPIKY = 3, DFSPH_B200_KERNEL_PRECOMPUTED_CUBIC = 4 };

/* Particle fields, named after the reference's FieldDescription names
 * (SPlisHSPlasH/FluidModel.cpp:60-66, SPlisHSPlasH/DFSPH/TimeStepDFSPH.cpp:49-53). */
typedef enum {
    DFSPH_B200_FIELD_POSITION = 0,       /* "position"               Real[3] */
    DFSPH_B200_FIELD_VELOCITY = 1,       /* "velocity"               Real[3] */
    DFSPH_B200_FIELD_DENSITY = 2,        /* "density"                Real    */
    DFSPH_B200_FIELD_FACTOR = 3,         /* "factor"                 Real    */
    DFSPH_B200_FIELD_DENSITY_ADV = 4,    /* "advected density"       Real    */
    DFSPH_B200_FIELD_KAPPA = 5,          /* "p / rho^2"              Real    (m_pressure_rho2)   */
    DFSPH_B200_FIELD_KAPPA_V = 6,        /* "p_v / rho^2"            Real    (m_pressure_rho2_V) */
    DFSPH_B200_FIELD_PRESSURE_ACCEL = 7, /* "pressure acceleration"  Real[3] */
    DFSPH_B200_FIELD_ID = 8,             /* "id"                     uint32  (original particle index) */
    DFSPH_B200_FIELD_STATE = 9,          /* "state"                  uint32  (0 Active, 1 AnimatedByEmitter, 2 Fixed) */
    DFSPH_B200_FIELD_BOUNDARY_VOLUME = 10, /* BoundaryModel_Akinci2012::m_V, Real, in the order boundaries were added */
    DFSPH_B200_FIELD_NUM_NEIGHBORS = 11  /* uint32: fluid + boundary neighbour count used by the deficiency test */
} dfsph_b200_field;

typedef struct {
    int32_t device;                 /* CUDA device ordinal */
    int32_t kernel;                 /* Simulation "kernel": 0 cubic, 1 Wendland quintic C2, 2 Poly6, 3 Spiky, 4 precomputed
                                       cubic (DFSPH default, Simulation.cpp:579-582; setKernel :338-393).  Honoured by
                                       the f64 library everywhere (the pairs 0/0 and 4/4 have dedicated kernels, every
                                       other kernel/grad_kernel pair runs a run-time switched variant); the f32 library
                                       mirrors the AVX build: solver sums always use the analytic cubic kernel
                                       (TimeStep.cpp:80, TimeStepDFSPH.cpp:789) and `kernel` only selects the kernel of
                                       the boundary-volume initialisation (BoundaryModel_Akinci2012.cpp:61-72). */
    double particle_radius;         /* Simulation "particleRadius"; support radius = 4 r (Simulation.cpp:283) */
    uint64_t max_fluid_particles;   /* device capacity; 0 = size of the first set_fluid call */
    int32_t max_fluid_neighbors;    /* per-particle neighbour-table capacity, fluid set   (0 = default 64) */
    int32_t max_boundary_neighbors; /* per-particle neighbour-table capacity, boundary set (0 = default 64) */
    double domain_min[3];           /* cell-grid extent; if min == max it is derived from the particle sets at */
    double domain_max[3];           /*   the first step (particles that leave it are clamped to the edge cells) */
    int32_t rank;                   /* multi-GPU slab decomposition (axis and bounds: dfsph_b200_comm_init): this context's slab (0 for single GPU) */
    int32_t world_size;             /* number of slabs (1 for single GPU) */
    int32_t grad_kernel;            /* Simulation "gradKernel" (setGradKernel, Simulation.cpp:306-336), same ids;
                                       -1 (dfsph_b200_default_config) = same as `kernel` */
} dfsph_b200_config;

/* TimeStepDFSPH / Simulation / TimeManager parameters, same names and defaults as the reference
 * (TimeStepDFSPH.cpp:28-41,73-115; Simulation.cpp:67-88,163-277; TimeManager.cpp:12). */
typedef struct {
    double time_step_size;          /* TimeManager h, default 0.001 */
    double gravitation[3];          /* default (0,-9.81,0) */
    uint32_t min_iterations;        /* 2 */
    uint32_t max_iterations;        /* 100 */
    double max_error;               /* 0.01  (percent) */
    uint32_t max_iterations_v;      /* 100 */
    double max_error_v;             /* 0.1   (percent) */
    int32_t enable_divergence_solver; /* 1 */
    int32_t cfl_method;             /* 0 none, 1 standard, 2 iter (Simulation.cpp:395-413) */
    double cfl_factor;              /* 0.5 */
    double cfl_min_time_step_size;  /* 1e-4 */
    double cfl_max_time_step_size;  /* 5e-3 */
    /* next-row f1: FluidModel "viscosityMethod" (0 none, 1 "Standard viscosity") and the Viscosity_Standard
     * parameters "viscosity" / "viscosityBoundary" (Viscosity/Viscosity_Standard.cpp:21-22,33-44).  Other methods: */
    int32_t viscosity_method;       /* default here 0; the reference's FluidModel default is 1 (FluidModel.cpp:98) */
    double viscosity;               /* 0.01 */
    double viscosity_boundary;      /* 0.0 */
} dfsph_b200_params;

typedef struct {
    uint32_t iterations;            /* "iterations"  (pressure solver)   TimeStepDFSPH.cpp:343 */
    uint32_t iterations_v;          /* "iterationsV" (divergence solver) TimeStepDFSPH.cpp:497 */
    double avg_density_error;       /* last avg_density_err of the pressure solve   */
    double avg_density_error_v;     /* last avg_density_err of the divergence solve */
    double time_step_size;          /* h after updateTimeStepSize (used by the pressure solve of this step) */
    double time;                    /* TimeManager time after the step */
    uint32_t num_particles;
    uint32_t max_neighbors;         /* largest fluid-neighbour count seen this step */
    uint32_t gpu_launches;          /* kernels launched by this step */
    float ms_search;                /* device time: cell sort + reorder + neighbour table (CUDA events) */
    float ms_solver;                /* device time: everything after the search */
} dfsph_b200_step_stats;

int dfsph_b200_sizeof_real(void);
const char* dfsph_b200_version(void);
void dfsph_b200_default_config(dfsph_b200_config* cfg);
void dfsph_b200_default_params(dfsph_b200_params* p);

/* Replaces: `new TimeStepDFSPH()` + `new NeighborhoodSearch(supportRadius)` (Simulation.cpp:149,577-583). */
int dfsph_b200_create(const dfsph_b200_config* cfg, dfsph_b200_ctx** out);
int dfsph_b200_destroy(dfsph_b200_ctx* ctx);
const char* dfsph_b200_last_error(const dfsph_b200_ctx* ctx);  /* ctx may be NULL: error of the last failed create */

/* Replaces: FluidModel::initModel -> add_point_set(x, n, dynamic, search, find) (FluidModel.cpp:285-327) and
 * SimulationDataDFSPH::init (DFSPH/SimulationDataDFSPH.cpp:22-60).  x, v: AoS Real[3*n] (Eigen DontAlign layout,
 * Common.h:27); v, id, state may be NULL (zero velocity, id = index, Active).  kappa/kappa_v start at 0.
 * May be called again to restart a run (SimulatorBase::reset -> TimeStep::reset): device capacity, boundary sets and,
 * in multi-GPU contexts, the communicator and peer mappings are kept (every rank must re-submit its slab). */
int dfsph_b200_set_fluid(dfsph_b200_ctx* ctx, uint64_t n, const void* x, const void* v,
                         const uint32_t* id, const uint32_t* state, double density0, double volume);

/* Replaces: BoundaryModel_Akinci2012::initModel -> add_point_set(x, n, dynamic=false, search=false, find=true)
 * (BoundaryModel_Akinci2012.cpp:77-110).  May be called several times (one call per rigid body); bodies are
 * concatenated on the device.  V may be NULL (then call dfsph_b200_compute_boundary_volume).  Dynamic / animated
 * bodies are outside the scope: is_dynamic != 0 returns DFSPH_B200_ERR_UNSUPPORTED. */
int dfsph_b200_add_boundary(dfsph_b200_ctx* ctx, uint64_t n, const void* x, const void* V, int is_dynamic);

/* Replaces: Simulation::updateBoundaryVolume + BoundaryModel_Akinci2012::computeBoundaryVolume
 * (Simulation.cpp:696-756, BoundaryModel_Akinci2012.cpp:48-75):  V_i = 1 / (W(0) + sum_{boundary j} W(x_i - x_j)). */
int dfsph_b200_compute_boundary_volume(dfsph_b200_ctx* ctx);

/* Replaces: GenParam setters of TimeStepDFSPH / Simulation / TimeManager (see dfsph_b200_params). */
int dfsph_b200_set_params(dfsph_b200_ctx* ctx, const dfsph_b200_params* p);
int dfsph_b200_get_params(const dfsph_b200_ctx* ctx, dfsph_b200_params* p);

/* Replaces: TimeStepDFSPH::step() (DFSPH/TimeStepDFSPH.cpp:117-249) including
 * Simulation::performNeighborhoodSearch (Simulation.cpp:606-619).  State stays resident on the device.
 * stats may be NULL (then the call does not synchronise with the device). */
int dfsph_b200_step(dfsph_b200_ctx* ctx, dfsph_b200_step_stats* stats);

/* The same step through HOST buffers, as the reference-facing plugin does when host code touched the particle
 * arrays: uploads x and v (AoS, in id order, i.e. row k = particle with id k), steps, downloads x and v into the
 * same buffers and (if non-NULL) density into `density`.  Copies are inside the call (the velocity upload overlaps the
 * neighbour search, the density download overlaps the solver).
 * Multi-GPU contexts (ids are global): rows are in this rank's device order, the buffers must hold
 * dfsph_b200_capacity() rows; the first num_particles() rows are read, the first stats->num_particles rows (the count
 * after this step's migration) are written; DFSPH_B200_FIELD_ID (by_id = 0) names the rows. */
int dfsph_b200_step_host(dfsph_b200_ctx* ctx, void* x_inout, void* v_inout, void* density_out,
                         dfsph_b200_step_stats* stats);

/* Replaces: the host-pointer accessors FluidModel::getPosition(i) ... / FieldDescription::getFct(i)
 * (FluidModel.h:57-73,247-400).  Rows are in CURRENT device order (the order the search sorted the particles into;
 * download DFSPH_B200_FIELD_ID to map rows to original indices) unless by_id != 0, in which case row k belongs to the
 * particle whose id is k.  `bytes` must equal n * elemsize of the field. */
int dfsph_b200_download(dfsph_b200_ctx* ctx, dfsph_b200_field field, void* dst, size_t bytes, int by_id);
int dfsph_b200_upload(dfsph_b200_ctx* ctx, dfsph_b200_field field, const void* src, size_t bytes, int by_id);

/* Replaces: NeighborhoodSearch::find_neighbors + PointSet::n_neighbors / neighbor_list
 * (Simulation.cpp:617, Simulation.h:456-473) for set 0 (fluid) against set `other` (0 fluid, 1 boundary).
 * Runs the search on the current positions if needed.  counts[n]; if idx != NULL, offsets[n+1] and idx[cap] receive a
 * CSR table (lists ascending).  Fluid rows/indices are in current device order (see FIELD_ID); boundary indices are
 * the order in which boundary particles were added.  Only for tests and non-ported host code -- the solver never
 * materialises host-visible lists.  A list that exceeds max_fluid_neighbors / max_boundary_neighbors makes the call fail
 * with DFSPH_B200_ERR_CAPACITY (like the step): truncated lists are never handed out. */
int dfsph_b200_neighbors(dfsph_b200_ctx* ctx, int other, uint32_t* counts, uint64_t* offsets, uint32_t* idx, uint64_t cap);

/* Neighbour search + density only (what ReadWriteStateTests.cpp:349-350 exercises):
 * Simulation::performNeighborhoodSearch(); TimeStep::computeDensities(0). */
int dfsph_b200_search_and_density(dfsph_b200_ctx* ctx);

uint64_t dfsph_b200_num_particles(const dfsph_b200_ctx* ctx);
uint64_t dfsph_b200_num_boundary_particles(const dfsph_b200_ctx* ctx);
uint64_t dfsph_b200_capacity(const dfsph_b200_ctx* ctx);   /* fluid rows the device arrays (and step_host buffers) hold */

/* Evaluate the device kernel functions W and gradW at n points r (AoS Real[3n]) -> W[n], gradW[3n].
 * Lets the reference's Tests/Kernel/KernelTests.cpp checks run against the device implementations. */
int dfsph_b200_eval_kernel(dfsph_b200_ctx* ctx, int kernel, uint64_t n, const void* r, void* W, void* gradW);

/* ---- multi-GPU slab decomposition along one axis (0 x, 1 y, 2 z): one context per GPU / process (DESIGN.md "Multi-GPU") -------------------
 * The reference has no distributed path at all; these entry points are new.  Rank 0 obtains an id with
 * dfsph_b200_comm_get_unique_id and hands the 256 bytes to every rank (the host layer uses torch.distributed for
 * that); every rank then calls dfsph_b200_comm_init BEFORE dfsph_b200_set_fluid with the slab axis and its slab [slab_lo, slab_hi) (use
 * +-1e300 for the outermost faces) and passes only the particles inside its slab to set_fluid (ids = global indices).
 * All ranks must share config.domain_min/max.  dfsph_b200_step then also performs: migration of particles that left
 * the slab and the one-support-radius ghost exchange of x and v once per step (NCCL send/recv, message sizes from an
 * exchange of counts), and per solver iteration the refresh of the ghosts' kappa and pressure acceleration and the
 * all-reduce of the density-error sum -- over NVLink peer memory once dfsph_b200_p2p_import has run (below), over NCCL
 * send/recv + ncclAllReduce otherwise -- plus the all-reduce of the particle count and the CFL maximum, so that every
 * rank takes identical iteration and time-step decisions.  Transfers in multi-GPU runs address rows in this rank's
 * device order (by_id = 0, together with DFSPH_B200_FIELD_ID: ids are global); dfsph_b200_step_host works on slab
 * contexts with buffers of dfsph_b200_capacity() rows in that order. */
int dfsph_b200_comm_get_unique_id(void* id256);   /* 256 bytes: two NCCL ids (reductions/migration + halo refresh) */
int dfsph_b200_comm_init(dfsph_b200_ctx* ctx, const void* id256, int rank, int world_size, int axis, double slab_lo, double slab_hi);

/* Optional: refresh the ghost particles with direct NVLink peer stores instead of NCCL send/recv.  After set_fluid
 * every rank exports a 512-byte blob of CUDA IPC handles (its particle arrays, flag words and all-reduce table), the
 * host layer all-gathers the blobs and hands every rank the full set, and from then on (a) every ghost refresh is one
 * small kernel that stores the export values straight into the neighbour's ghost slots and publishes a sequence number
 * there -- a one-thread kernel in front of the consumer spins on the local flag; the sequence numbers, export counts and
 * slot offsets live in device memory -- and (b) the per-iteration all-reduce of the density error is fused into the tail
 * of pass B: every rank stores its partial sum into every rank's table, waits for all of them and adds them up in rank
 * order.  A solver iteration is then six kernels and no NCCL call, and the solver loops run as CUDA graphs with a WHILE
 * node exactly as on one GPU (no host synchronisation inside a solve).
 * Requires peer access between the GPUs (NVLink/NVSwitch), at most 16 ranks; if p2p_import is never called the NCCL
 * path is used. */
int dfsph_b200_p2p_export(dfsph_b200_ctx* ctx, void* blob512);
int dfsph_b200_p2p_import(dfsph_b200_ctx* ctx, const void* blobs_all /* world_size x 512 bytes, in rank order */);
int dfsph_b200_p2p_disable(dfsph_b200_ctx* ctx);   /* all ranks together, between steps: back to the NCCL path */

/* Device timing.  timer_start/stop bracket any number of calls with two CUDA events on the context's own stream
 * (the stream every kernel of this library is launched on).  With profiling on, every launch of the kernel classes
 * below is additionally bracketed by its own event pair; get_profile returns accumulated milliseconds and launch
 * counts per class since profiling was switched on.  (Mirrors the reference's START_TIMING/STOP_TIMING_AVG timers
 * "neighborhood_search", "divergenceSolve", "pressureSolve": Simulation.cpp:616, TimeStepDFSPH.cpp:166,213.) */
enum {
    DFSPH_B200_PROF_SORT = 0,         /* cell counting sort + reorder */
    DFSPH_B200_PROF_BUILD = 1,        /* neighbour table */
    DFSPH_B200_PROF_INIT = 2,         /* K1+K2+K3 density / factor / divergence source */
    DFSPH_B200_PROF_ACCEL = 3,        /* pass A of an iteration */
    DFSPH_B200_PROF_JACOBI_DIV = 4,   /* pass B, divergence solve */
    DFSPH_B200_PROF_JACOBI_PRESS = 5, /* pass B, pressure solve */
    DFSPH_B200_PROF_DIV_FINAL = 6,
    DFSPH_B200_PROF_PRESS_INIT = 7,
    DFSPH_B200_PROF_PRESS_FINAL = 8,
    DFSPH_B200_PROF_CLASSES = 9
};
int dfsph_b200_set_profiling(dfsph_b200_ctx* ctx, int on);
int dfsph_b200_get_profile(dfsph_b200_ctx* ctx, double* ms /*[DFSPH_B200_PROF_CLASSES]*/, uint64_t* count);
int dfsph_b200_timer_start(dfsph_b200_ctx* ctx);
int dfsph_b200_timer_stop(dfsph_b200_ctx* ctx, float* ms);

/* Multi-socket hosts: move the calling thread to the CPUs of `device`'s NUMA node and prefer that node's memory for later
 * allocations (call before dfsph_b200_alloc_pinned / before the FluidModel arrays are allocated and registered; one rank =
 * one process = one GPU).  The reference has no counterpart (its arrays live wherever the OpenMP first touch put them).
 * Returns the node (>= 0) or -1 if the topology is unknown or the binding is not permitted; never fails the caller. */
int dfsph_b200_bind_host_numa(int device);

/* Pinned host buffers for the host-buffer path (dfsph_b200_step_host) and an explicit stream synchronise. */
void* dfsph_b200_alloc_pinned(size_t bytes);
void dfsph_b200_free_pinned(void* p);
/* Page-lock host arrays the caller already owns (the FluidModel's std::vector storage of x, v, density:
 * SPlisHSPlasH/FluidModel.h:113-124) so that dfsph_b200_step_host moves them at PCIe speed instead of through the
 * driver's pageable staging (measured: 16.5 instead of 46.9 ms per step at 10 M particles).  The caller must unregister
 * before the memory is freed or reallocated.  Returns DFSPH_B200_OK, or DFSPH_B200_ERR_CUDA (the array then simply
 * stays pageable). */
int dfsph_b200_host_register(void* p, size_t bytes);
int dfsph_b200_host_unregister(void* p);
int dfsph_b200_synchronize(dfsph_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DFSPH_B200_H */
