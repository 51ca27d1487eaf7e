/*
 * epseon_cuda.h -- C ABI of the B200-native Numerov hot path (libepseon_cuda.so).
 *
 * This is the drop-in boundary under the reference's algorithm slot
 *     virtual void Algorithm<FP>::run(const std::stop_token&, TaskHandle<FP>*)
 *         (cpp/gpu/include/epseon/gpu/algorithms/algorithm.hpp:34)
 * as implemented by VibwaAlgorithm<FP>::run
 *         (cpp/gpu/include/epseon/gpu/algorithms/vibwa.hpp:605-637),
 * whose Vulkan logical-device / VMA / descriptor-set body it replaces.  The
 * reference has no FFI of its own for this path (the slot does no arithmetic);
 * the entry points below are what that slot binds instead.  Each one cites the
 * reference interface it stands in for.  Plain C: pointers and sizes only, no
 * C++/torch/Python types, no exceptions across the boundary.
 *
 * Conventions
 *   - every call returns an int status (EPS_OK == 0); the text of the last
 *     failure is eps_last_error(ctx) (ctx may be NULL for creation failures);
 *   - one eps_ctx per CUDA device and per host thread (thread-compatible, like
 *     the reference's one-jthread-per-TaskHandle model, task_handle.hpp:100);
 *   - the caller owns every host buffer; the ctx owns every device buffer;
 *   - work is stream-ordered on the ctx's stream; calls that return data to a
 *     host buffer synchronise before returning, the others do not
 *     (eps_sync() waits explicitly);
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with EPS_ERR_CUDA.
 *
 * Units follow the build's spec: energies in cm^-1,
 * lengths in Angstrom, masses in amu; `scale` = h^2 * (2 mu / hbar^2) / 12.
 */
#ifndef EPSEON_CUDA_H
#define EPSEON_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPS_ABI_VERSION 2

enum {
    EPS_OK          = 0,
    EPS_ERR_INVALID = 1, /* bad argument                                   */
    EPS_ERR_CUDA    = 2, /* CUDA runtime failure / no device               */
    EPS_ERR_RANGE   = 3, /* energy or table outside the validity window    */
    EPS_ERR_STATE   = 4, /* call order (e.g. sweep before set_potentials)  */
    EPS_ERR_NOMEM   = 5,
    EPS_ERR_CANCELLED = 6 /* interrupted by eps_request_stop                */
};

typedef struct eps_ctx eps_ctx;

/* Device description.  Replaces what the reference reads from
 * vk::PhysicalDeviceProperties / vk::PhysicalDeviceMemoryProperties
 * (compute_context.hpp:45-48, compute_context.cpp:94-104). */
typedef struct eps_device_props {
    char     name[256];
    int32_t  ordinal;               /* CUDA ordinal: the unique device_id */
    int32_t  cc_major, cc_minor;
    int32_t  sm_count;
    int32_t  clock_khz;
    int32_t  driver_version;        /* cudaDriverGetVersion  */
    int32_t  runtime_version;       /* cudaRuntimeGetVersion */
    int32_t  pci_domain, pci_bus, pci_device;
    int32_t  integrated;
    int32_t  max_threads_per_block;
    int32_t  max_grid[3];
    int32_t  max_block[3];
    int32_t  l2_bytes;
    uint64_t total_global_mem;
    uint64_t shared_mem_per_block_optin;
    uint64_t shared_mem_per_sm;
    uint8_t  uuid[16];
} eps_device_props;

/* Per-curve facts computed by eps_set_potentials. */
typedef struct eps_curve_info {
    uint32_t i0;       /* first integrated grid index (energy independent)  */
    uint32_t n_steps;  /* recurrence steps per trial energy                 */
    double   scale;    /* s: E[cm^-1] -> dimensionless Numerov energy       */
    double   v_min;    /* minimum of the table                              */
    double   v_last;   /* last table value (taken as the asymptote)         */
} eps_curve_info;

/* Level-search parameters.  v_min/v_max are VibwaAlgorithmConfig::min_level /
 * max_level (algorithm_config.hpp:82-83); level_count = v_max-v_min+1 is the
 * reference's planned output-buffer length (algorithm_config.hpp:177). */
typedef struct eps_solve_params {
    uint32_t v_min, v_max;
    uint32_t n_coarse;       /* trial energies of the bracketing sweep (per curve)  */
    uint32_t refine_points;  /* interior trial energies per level per round (M)     */
    uint32_t max_rounds;     /* k-section rounds / Cooley iterations per level      */
    uint32_t flags;          /* EPS_SOLVE_* (0: k-section on node counts)           */
    double   rel_tol;        /* stop when hi-lo <= rel_tol*max(|lo|,|hi|)           */
} eps_solve_params;

/* eps_solve_params.flags.
 * EPS_SOLVE_COOLEY: after the coarse sweep has bracketed the levels, refine every level with the
 *   outward/inward matching (Cooley) iteration instead of k-section sweeps (spec
 *   DESIGN.md section 3.9): march out from the left wall and in from the right wall to the outer
 *   classical turning point, take the Rayleigh-quotient correction of section 3.8 as a Newton step,
 *   keep a rigorous bracket from the node count of the full outward solution (a step that leaves it
 *   is replaced by the midpoint), stop when |correction| <= rel_tol |E|.  One CTA per (curve, level)
 *   iterates on the device; 3-5 iterations from a coarse bracket.  Needs the accurate tables
 *   (EPS_OPT_FORM = 1) and windows of >= 256 steps (else the call uses k-section).  levels = E,
 *   widths = |last correction|.  Agrees with the k-section result to the tolerance (a few 1e-14
 *   measured), not bit for bit; bit-identical to the oracle's statement of the same iteration.
 * EPS_SOLVE_OPEN_TAIL (with EPS_SOLVE_COOLEY): the inward branch starts on the DECAYING solution of
 *   the recurrence (last coefficient frozen) instead of psi = 0 at the end of the table: the levels
 *   of the unbounded problem.  For states near dissociation, whose tail still matters at r_max --
 *   within VibwaAlgorithmConfig::min_distance_to_asymptote of the asymptote
 *   (algorithm_config.hpp:81) -- the box pushes the level up (docs-example curve, v = 26:
 *   +2.6e-6 relative); the open tail removes that.  Node counts are a property of the box problem, so
 *   the bracket becomes a soft window [lo - (hi - lo), hi] around the box level. */
enum { EPS_SOLVE_COOLEY = 1, EPS_SOLVE_OPEN_TAIL = 2 };
/* Grid steps per segment of the Cooley kernel for windows of n_steps steps when n_items (curve, level)
 * pairs are searched in one call (one thread per segment; the summation order, hence the result's
 * last bits, depend on it: the oracle is handed the same number). */
uint32_t eps_cooley_segment_length(uint32_t n_steps, uint64_t n_items);

/* Counters since the last eps_stats_reset(). */
typedef struct eps_stats {
    uint64_t sweep_launches;   /* Numerov sweep kernel launches                    */
    uint64_t other_launches;   /* prep / bracketing / bookkeeping kernel launches  */
    uint64_t grid_steps;       /* grid-steps x trial-energies executed by sweeps   */
    double   sweep_ms;         /* summed CUDA-event time of the sweep launches     */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t kernel_launches;  /* every kernel launched: each sweep kernel (a constant-bank sweep
                                  is one launch per 3968-step chunk) + other_launches          */
} eps_stats;

/* ---- devices and contexts (replaces ComputeContext::getPhysicalDevicesInfo /
 * getDeviceInterface, compute_context.cpp:94-117, and createLogicalDevice,
 * vibwa.hpp:654-674) ------------------------------------------------------- */
int         eps_abi_version(void);
int         eps_device_count(int* count);
int         eps_device_get_props(int device, eps_device_props* out);
int         eps_ctx_create(int device, eps_ctx** out);
int         eps_ctx_destroy(eps_ctx* ctx);
const char* eps_last_error(const eps_ctx* ctx);
int         eps_sync(eps_ctx* ctx);
/* Device memory held by the context's grow-only buffers; eps_ctx_trim releases the scratch ones
 * (and, with drop_potentials != 0, the resident tables too: a new eps_set_potentials* is then
 * needed).  The reference frees its VMA buffers when run() returns (vibwa.hpp:236-283); a pooled
 * context is trimmed before it is parked. */
int         eps_ctx_device_bytes(eps_ctx* ctx, uint64_t* bytes);
int         eps_ctx_trim(eps_ctx* ctx, int drop_potentials);

/* ---- cooperative cancellation (binds TaskHandle<FP>::cancel, task_handle.hpp:136-144; the
 * reference polls its stop_token three times in run(), vibwa.hpp:606,617,632).
 * eps_request_stop is the ONE entry point that may be called from another thread while a compute
 * call is running on the context: eps_solve_levels* test the flag between refinement rounds, every
 * sweep before it returns, and sweep CTAs on entry (queued launches drain in microseconds).  The
 * interrupted call returns EPS_ERR_CANCELLED; the flag stays up until eps_reset_stop. */
int eps_request_stop(eps_ctx* ctx);
int eps_reset_stop(eps_ctx* ctx);

/* ---- potentials (replaces the staging->device upload the reference plans in
 * ShaderResources, vibwa.hpp:57-234; input is the table that
 * PotentialSource<FP>::get_potential_data() returns, potential_source.hpp:38)
 * V: host, n_curves rows of n_points doubles; scale: host, one per curve. */
int eps_set_potentials(eps_ctx* ctx, const double* V, uint32_t n_curves, uint32_t n_points,
                       const double* scale);
int eps_get_curve_info(eps_ctx* ctx, uint32_t curve, eps_curve_info* out);

/* ---- rotational states (the reference's parameter set has no J,
 * algorithm_config.hpp:78-83 -- additive).  Every table row is expanded ON THE DEVICE into n_J
 * effective curves
 *     V_J(r_i) = V(r_i) + J(J+1) * (hbar^2/2mu) / r_i^2,   r_i = r_min[c] + i*grid_step[c],
 * hbar^2/2mu = grid_step[c]^2 / (12*scale[c]).  Resident curve index = c*n_J + j, so every later
 * call (eps_sweep*, eps_solve_levels*, eps_wavefunctions, eps_get_curve_info) sees
 * n_curves*n_J curves and takes / returns arrays of that many rows.  J[j] = 0 keeps the row bit
 * for bit; a grid point at r = 0 becomes a 1e300 wall.  V: host [n_curves][n_points];
 * scale, r_min, grid_step: host [n_curves]; J: host [n_J]. */
int eps_set_potentials_rot(eps_ctx* ctx, const double* V, uint32_t n_curves, uint32_t n_points,
                           const double* scale, const double* r_min, const double* grid_step,
                           const uint32_t* J, uint32_t n_J);

/* ---- tabulated sources (N1): natural cubic spline through n_knots points (r strictly increasing),
 * resampled on the uniform grid r_i = r_min + i*(r_max-r_min)/(n_points-1).  Makes real the
 * tabulated-file source the reference declares and leaves empty
 * (PotentialFileLoader<FP>::get_potential_data, potential_source.hpp:89-91) for ab initio tables
 * on non-uniform grids.  eps_spline_coefficients is host-only (no context): coef receives
 * a,b,c,d of S = a + dx(b + dx(c + dx d)) per interval, 4*(n_knots-1) doubles.
 * eps_spline_resample evaluates on the device and returns V_out (host, n_points doubles). */
int eps_spline_coefficients(const double* r, const double* V, uint32_t n_knots, double* coef);
int eps_spline_resample(eps_ctx* ctx, const double* r, const double* V, uint32_t n_knots, double r_min,
                        double r_max, uint32_t n_points, double* V_out);

/* ---- Numerov sweep + node count (the missing compute dispatch of
 * vibwa.hpp:605-637).  Per curve, n_energies trial energies; outputs are host
 * buffers [n_curves][n_energies] and any of them may be NULL (results then
 * stay resident on the device).
 *   eps_sweep:          explicit energies E[n_curves][n_energies] (host).
 *   eps_sweep_uniform:  E_j = E_lo[c] + j*dE, dE=(E_hi[c]-E_lo[c])/(n-1),
 *                       generated on the device. */
int eps_sweep(eps_ctx* ctx, const double* E, uint64_t n_energies, uint32_t* nodes,
              double* tail_mant, int32_t* tail_exp);
int eps_sweep_uniform(eps_ctx* ctx, const double* E_lo, const double* E_hi, uint64_t n_energies,
                      uint32_t* nodes, double* tail_mant, int32_t* tail_exp);

/* Affine grid: E_j = E0[c] + (j0 + j) * dE[c], j = 0..n-1 -- a slice of a larger ("global")
 * uniform grid with the global grid's energies reproduced bit for bit.  This is what a rank of an
 * energy-range-sharded job calls. */
int eps_sweep_grid(eps_ctx* ctx, const double* E0, const double* dE, uint32_t j0, uint64_t n_energies,
                   uint32_t* nodes, double* tail_mant, int32_t* tail_exp);

/* ---- bracketing + k-section refinement of levels v_min..v_max of every
 * resident curve inside [E_lo[c], E_hi[c]].  Outputs (host):
 *   levels[n_curves][level_count]  (NaN where the level is not in range) --
 *     the reference's planned output buffer (algorithm_config.hpp:183-190); may be NULL: the
 *     results then stay on the device (for eps_mailbox_post_levels / eps_group_*);
 *   widths[n_curves][level_count]  final bracket widths (may be NULL);
 *   n_below[n_curves]              levels below E_hi[c]   (may be NULL). */
int eps_solve_levels(eps_ctx* ctx, const eps_solve_params* p, const double* E_lo,
                     const double* E_hi, double* levels, double* widths, uint32_t* n_below);

/* Same search with the coarse sweep on the affine grid E_j = E0[c] + (j0 + j) * dE[c],
 * j = 0..p->n_coarse-1.  n_last / n_first (may be NULL) receive the node counts at the last /
 * first grid point of every curve.  Energy-range sharding: rank r passes j0 = r * per and
 * n_coarse = per + 1 (one shared point with its neighbour); every bracket of the global grid then
 * falls into exactly one rank's slice and the union of the ranks' levels is bit-identical to a
 * single-device eps_solve_levels over the whole range. */
int eps_solve_levels_grid(eps_ctx* ctx, const eps_solve_params* p, const double* E0, const double* dE,
                          uint32_t j0, double* levels, double* widths, uint32_t* n_last,
                          uint32_t* n_first);

/* ---- options / counters.
 * EPS_OPT_SCAN_SEGMENTS: transfer-matrix (scan) path of the sweep for the few-energy, long-grid
 *   regime: 0 = automatic (default), 1 = never, n >= 2 = always cut the grid into n segments.
 *   On the scan path node counts equal the sequential march's (energies whose result is
 *   ill-conditioned are detected and recomputed sequentially); tails agree to rounding.
 *   Up to 512 segments of whole 2048-step tiles.
 * EPS_OPT_SCAN_COMBINE: how the segment transfer matrices of an energy are chained.  0 = automatic:
 *   a serial loop per energy below 8 segments, from 8 on a block-level parallel prefix over the
 *   2x2 matrices (16 segment lanes x 32 energies per CTA, Hillis-Steele in shared memory); 1 = serial
 *   loop always, 2 = prefix always.  Same node counts; tails agree to rounding.
 * EPS_OPT_SCAN_EXACT: 1 (default) = eps_solve_levels also recomputes flagged energies
 *   sequentially, so levels carry the sequential march's bits (late refinement rounds, whose
 *   energies all crowd an eigenvalue, then run at sequential speed); 0 = it does not: faster,
 *   and the levels agree with the sequential ones to the rounding-noise floor of the FP64
 *   recurrence (a few 1e-9 relative).
 * EPS_OPT_CBANK: sweeps over ONE resident curve can feed the coefficient table through the
 *   kernel-parameter constant bank (uniform-register operand, chunked launches) instead of the
 *   TMA / shared-memory ring; identical results.  0 = automatic (large sweeps), 1 = whenever
 *   the launch qualifies, 2 = never.  EPS_OPT_CBANK_SHAPE, EPS_OPT_CBANK_PDL: tuning knobs.
 * EPS_OPT_CBANK_GROUP: the constant-bank sweep carries 28 B of state per energy from chunk launch to
 *   chunk launch.  n > 0 runs all chunk launches of one group of n resident waves of CTAs before
 *   the next group's, so that the group's state stays in L2 instead of crossing HBM once per chunk;
 *   0 = one group (every launch covers all energies); default 2 (34 MB of state on a B200:
 *   0.7 MB instead of 940 MB of DRAM traffic per chunk launch on C5, for 0.7 % more sweep time).
 *   Identical results.
 * EPS_OPT_PREP_PARTS: eps_set_potentials* prepares few long curves (<= 64 curves of >= 65 536
 *   points) with every curve cut into chunks over many CTAs; 0 = automatic, 1 = never (one CTA
 *   per curve).  Identical results.
 * EPS_OPT_PACK128: refinement rounds of many-curve batches with <= 64 points per level pack their rows
 *   into 128-energy CTAs (four resident per SM) instead of 256-energy ones when that balances the
 *   SMs better (a small per-device batch, e.g. 512 curves: 3.46 CTAs per SM); 0 = automatic,
 *   1 = always, 2 = never.  Identical results.
 * EPS_OPT_FORM: the recurrence every later eps_set_potentials* prepares its tables for, and every
 *   sweep on them then runs.  0 (default) = X form: 4 FP64 operations per
 *   grid step, eigenvalue rounding-noise floor ~1e-9 relative at 2e5 grid points and ~2e-8 at 1e6.
 *   1 = D form ("accurate"): the chain carries the first difference, 5 operations per step,
 *   eigenvalues within ~1e-14 of the binary128 solution of the discrete problem at every grid size
 *   (tests/test_accuracy_floor.py).  Node counts and levels of each form are bit-identical to the
 *   oracle's same form; the two forms agree with each other to the X form's noise floor. */
enum { EPS_OPT_SCAN_SEGMENTS = 1, EPS_OPT_SCAN_EXACT = 2, EPS_OPT_CBANK = 3, EPS_OPT_CBANK_SHAPE = 4, EPS_OPT_CBANK_PDL = 5,
       EPS_OPT_PREP_PARTS = 6, EPS_OPT_FORM = 7, EPS_OPT_PACK128 = 8, EPS_OPT_CBANK_GROUP = 9,
       EPS_OPT_SCAN_COMBINE = 10 };
enum { EPS_CNT_SCAN_LAUNCHES = 1, EPS_CNT_SCAN_FLAGGED = 2, EPS_CNT_CBANK_LAUNCHES = 3 };
int eps_set_option(eps_ctx* ctx, int option, int64_t value);
int eps_get_counter(eps_ctx* ctx, int counter, uint64_t* value);

/* ---- normalised wavefunctions of located levels (N7; the reference plans no such
 * output -- additive).  E[n_curves][n_levels] (host; NaN entries give
 * a zero row), grid_step[n_curves] = h of every curve (for the norm
 * integral h * sum psi^2 = 1).  Outputs (host):
 *   psi[n_curves][n_levels][n_points]  on the full r grid (zero outside the
 *     integration window), positive on the first lobe;
 *   match_index[n_curves][n_levels]    grid index of the outward/inward matching
 *     point, UINT32_MAX for skipped rows (may be NULL). */
int eps_wavefunctions(eps_ctx* ctx, const double* E, uint32_t n_levels, const double* grid_step,
                      double* psi, uint32_t* match_index);

/* ---- a-posteriori energy correction of located levels (the Cooley step of an
 * outward/inward-matching search; additive).  For every E[n_curves][n_levels] (NaN rows give NaN) the
 * matched outward/inward solution leaves a residual of the Numerov equation at the matching point;
 * dE = -psi_m r_m / (psi^T B psi) / scale is the first-order correction of E (error quadratic in the
 * energy error): |dE| is an error estimate of a located level, and E + dE polishes a level found
 * with a loose tolerance.  dE[n_curves][n_levels] (host). */
int eps_level_corrections(eps_ctx* ctx, const double* E, uint32_t n_levels, const double* grid_step,
                          double* dE);

/* ---- several devices in ONE process ("one process, G host threads, one eps_ctx per
 * device"; the reference gives one device per task, device_interface.hpp:22 -- additive).  A group
 * owns one context and one host thread per device.  The path shards without any exchange during
 * compute; only the located levels are gathered: every device writes them into a buffer on the
 * group's first device with cudaMemcpyPeerAsync (NVLink when peer access exists), one
 * device->host copy returns them.
 *   EPS_SHARD_CURVES: contiguous blocks of curves per device.
 *   EPS_SHARD_ENERGY: every device holds all curves and works on a contiguous slice of the ONE
 *     global energy grid (slices of neighbours share one grid point), so sweeps and level searches
 *     return exactly the bits of a single-device call.
 * Arrays are host pointers for the WHOLE job ([n_curves] / [n_curves][n]).  eps_group_ctx lends a
 * device's context (options, counters, stats); eps_group_last_ms = max over the devices of the
 * CUDA-event time of the last group call. */
typedef struct eps_group eps_group;
enum { EPS_SHARD_CURVES = 0, EPS_SHARD_ENERGY = 1 };
int         eps_group_create(const int* devices, uint32_t n_devices, eps_group** out);
int         eps_group_destroy(eps_group* g);
uint32_t    eps_group_size(const eps_group* g);
eps_ctx*    eps_group_ctx(eps_group* g, uint32_t rank);
const char* eps_group_last_error(const eps_group* g);
int         eps_group_last_ms(eps_group* g, float* ms);
int         eps_group_set_option(eps_group* g, int option, int64_t value);
int         eps_group_set_potentials(eps_group* g, const double* V, uint32_t n_curves, uint32_t n_points,
                                     const double* scale, int shard);
int         eps_group_sweep_uniform(eps_group* g, const double* E_lo, const double* E_hi, uint64_t n_energies,
                                    uint32_t* nodes);
int         eps_group_solve_levels(eps_group* g, const eps_solve_params* p, const double* E_lo, const double* E_hi,
                                   double* levels, double* widths, uint32_t* n_below);

/* ---- one process per device (e.g. under torchrun): a mailbox in rank 0's device memory, shared
 * through a CUDA IPC handle (64 bytes; the caller moves it between the processes).  Every rank
 * writes its small result into its slot device-to-device (a peer write over NVLink) followed by a
 * sequence number in stream order; rank 0 waits for all sequence numbers, fetches the first
 * bytes_per_rank bytes of every slot with one strided device->host copy (out: [world][bytes_per_rank])
 * and acknowledges, which lets the senders post again (one payload in flight per rank).  Slot layout of eps_mailbox_post_levels: [levels | widths] of the
 * rank's last level search. */
typedef struct eps_mailbox eps_mailbox;
int    eps_mailbox_create(eps_ctx* ctx, uint32_t world, size_t bytes_per_rank, eps_mailbox** out, unsigned char* handle64);
int    eps_mailbox_open(eps_ctx* ctx, const unsigned char* handle64, uint32_t world, uint32_t rank, size_t bytes_per_rank,
                        eps_mailbox** out);
int    eps_mailbox_destroy(eps_mailbox* mb);
int    eps_mailbox_post_levels(eps_mailbox* mb, uint32_t seq);
int    eps_mailbox_post(eps_mailbox* mb, const void* src, size_t bytes, uint32_t seq);
int    eps_mailbox_collect(eps_mailbox* mb, uint32_t seq, void* out, size_t bytes_per_rank, double timeout_s);
size_t eps_mailbox_slot_bytes(const eps_mailbox* mb);

/* ---- page-locked host buffers (optional) ---------------------------------
 * Every entry point accepts ANY host pointer.  Tables and result arrays that live in memory
 * from eps_host_alloc are page-locked, so their host<->device copies run at the full PCIe rate
 * without the driver's staging copy (a 328 MB batch of curves: ~6 ms instead of ~20 ms).
 * ctx may be NULL for eps_host_free. */
int eps_host_alloc(eps_ctx* ctx, size_t bytes, void** out);
int eps_host_free(eps_ctx* ctx, void* p);

/* ---- measurement helpers (bench.py / tests) ------------------------------- */
int eps_timer_start(eps_ctx* ctx);
int eps_timer_stop(eps_ctx* ctx, float* ms);
int eps_stats_get(eps_ctx* ctx, eps_stats* out);
int eps_stats_reset(eps_ctx* ctx);
int eps_l2_flush(eps_ctx* ctx);
int eps_fp64_probe(eps_ctx* ctx, double* tflops, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* EPSEON_CUDA_H */
