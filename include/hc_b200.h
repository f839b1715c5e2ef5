/* hc_b200 -- C ABI of the B200-native batched path tracker for HomotopyContinuation.jl.
 *
 * This is the drop-in boundary: the Julia host (`ccall`) builds systems, start
 * solutions and ModelKit instruction tapes and calls these entry points; the library implements
 * the `track -> PathResult` contract of the reference's trackers on the GPU.
 *
 * What each entry point replaces in the reference (file:line under /root/reference):
 *   hc_system_create            InterpretedSystem(F): src/model_kit/interpreted_system.jl:23-51; takes
 *                               ModelKit.instruction_sequence(F) / jacobian_instruction_sequence(F)
 *                               (src/model_kit/instruction_sequence.jl:133-143, 256-272) verbatim
 *   hc_homotopy_create          StraightLineHomotopy / ParameterHomotopy / CoefficientHomotopy /
 *                               ToricHomotopy constructors (src/homotopies/*.jl) incl. FixedParameterSystem
 *   hc_track_batch   mode 0     track(::EndgameTracker, x, t1) for every start  (src/endgame_tracker.jl:902-914)
 *                               driven like threaded_solve (src/solve.jl:628-709); with path_q it is
 *                               many_solve (src/solve.jl:815-881) inverted to (parameter point x start)
 *                    mode 1     track(::Tracker, x, t1, t0)                     (src/tracker.jl:1023-1026)
 *   hc_polyhedral_track_batch   track(::PolyhedralTracker, (cell, x))           (src/polyhedral.jl:414-530)
 *   hc_polyhedral_track_cells   the same with the start solutions of every mixed cell made on the device
 *                               (PolyhedralStartSolutionsIterator, src/polyhedral.jl:104-144; src/binomial_system.jl:55-106)
 *   hc_evaluate, hc_evaluate_dd, hc_evaluate_and_jacobian, hc_taylor
 *                               the operator API evaluate!/evaluate_and_jacobian!/taylor!
 *                               (src/model_kit/abstract_system_homotopy.jl:96-120), single point test hooks
 *   hc_options                  TrackerOptions + TrackerParameters (src/tracker.jl:45-62, 94-140),
 *                               EndgameOptions (src/endgame_tracker.jl:47-72), WeightedNormOptions (src/norm.jl:36-40)
 *   hc_results                  struct-of-arrays of PathResult (src/path_result.jl:76-98)
 *
 * Conventions: every array is caller-owned and must stay alive for the duration of the (blocking)
 * call; complex numbers are (re, im) pairs of doubles (Julia ComplexF64); matrices and batches are
 * column-major as in Julia (starts: n x N, i.e. path k at starts[2*n*k ...]).  Functions returning
 * int32 give 0 on success and a negative code on library errors (hc_last_error() describes it);
 * per-path numerical failures are not errors but return codes in hc_results, as in the reference.
 * There is no CPU fallback: without a usable CUDA device every compute entry point fails. */
#ifndef HC_B200_H
#define HC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const int32_t* instructions; /* 6 int32 per Instruction: input[4], op, output (1-based); last = OP_STOP */
    int32_t n_instructions;
    const double* constants;     /* tape[1..C] */
    int32_t n_constants;
    int32_t param_offset, n_params; /* parameters_range = param_offset+1 : param_offset+n_params */
    int32_t t_index;                /* continuation_parameter_index, 0 = nothing */
    int32_t var_offset, n_vars;     /* variables_range */
    const int32_t* u_assign;        /* u_assignments (i, k) */
    int32_t n_u;
    const int32_t* U_assign;        /* U_assignments (j, k) */
    int32_t n_U;
    int32_t out_dim, tape_space;    /* output_dim, tape_space_needed */
} hc_program_desc;

enum { HC_STRAIGHT_LINE = 0, HC_PARAMETER = 1, HC_COEFFICIENT = 2, HC_TORIC = 3 };

typedef struct {
    int32_t kind;
    void* F;            /* hc_system: target system / the parametrised system */
    void* G;            /* hc_system: start system (straight line only) */
    double gamma[2];    /* straight line */
    const double* G_params; int32_t n_G_params; /* fixed parameters of G (total degree `scaling`) */
    const double* F_params; int32_t n_F_params; /* fixed parameters of F */
    const double* p;    /* start parameters (t = 1); toric: start coefficients */
    const double* q;    /* target parameters (t = 0) */
    int32_t n_pq;
} hc_homotopy_desc;

typedef struct {
    int32_t max_steps; double max_step_size, max_initial_step_size; int32_t extended_precision;
    double min_step_size, min_rel_step_size;
    double a, beta_a, beta_omega_p, beta_tau, strict_beta_tau; int32_t min_newton_iters;
    double endgame_start; int32_t max_endgame_steps, max_endgame_extended_steps;
    double min_cond, min_cond_growth, min_coord_growth;
    int32_t zero_is_at_infinity, at_infinity_check, only_nonsingular;
    double singular_min_accuracy; int32_t max_winding_number;
    double val_finite_tol, val_at_infinity_tol, sing_cond, sing_accuracy, scaling_threshold;
    int32_t refine_steps;
    double scale_min, scale_abs_min, scale_max;
} hc_options;

typedef struct {
    int32_t* return_code;        /* EndgameTrackerCode order (mode 0 / polyhedral), TrackerCode order (mode 1) */
    double* solution;            /* 2n x N */
    double* t;
    double* accuracy;
    double* residual;
    uint8_t* singular;
    double* condition_jacobian;  /* mode 1: TrackerResult.tau */
    int32_t* winding_number;     /* 0 = nothing */
    uint8_t* extended_precision;
    double* last_point;          /* 2n x N */
    double* last_t;
    double* valuation;           /* n x N */
    uint8_t* has_valuation;
    double* omega;
    double* mu;
    int32_t* accepted_steps;
    int32_t* rejected_steps;
    int32_t* steps_eg;
    uint8_t* extended_precision_used;
    int64_t* counters;           /* optional 8 x N: factorizations, ldivs, evaljac, eval, eval_dd, taylor K=1, K=2, K=3 */
} hc_results;

typedef struct {                 /* device timing of the last batch call on this thread */
    double h2d_ms, kernel_ms, d2h_ms;
    int64_t h2d_bytes, d2h_bytes;
    int32_t grid, block, lanes;
    int64_t slab_bytes;
    int32_t devices;             /* devices the batch was split over */
    int32_t engine;              /* 0 lane group per path, 1 thread per path (interpreter), 2 thread per path (specialised kernel) */
    int64_t handoff_paths;       /* two-pass batches: paths the thread-per-path kernel handed to the lane-group engine ... */
    double handoff_ms;           /* ... and the time of that second kernel (included in kernel_ms) */
} hc_timing;

int32_t hc_init(int32_t device);            /* the process drives this one CUDA device (= hc_init_devices(&device, 1)) */
/* The process drives these CUDA devices with ONE host call per batch: every hc_track_* call splits its path index
 * range into contiguous shards, one per device (sweeps: whole parameter points), launches them concurrently on
 * per-device streams and copies each device's PathResult slice into the caller's arrays at the shard offset
 * (reference: threaded_solve stores results by path index, src/solve.jl:628-709).  Batches with fewer than
 * HC_B200_MIN_PATHS_PER_DEVICE (2048) paths per device use fewer devices.  Handles created afterwards hold one copy of
 * their programs per device; call it before creating handles. */
int32_t hc_init_devices(const int32_t* devices, int32_t n_devices);
int32_t hc_device_count(void);
/* stop_early_cb / catch_interrupt (reference src/solve.jl:618, 677, 685-707): on != 0 makes the running (and any
 * later) batch kernel stop handing out new paths -- paths in flight finish, paths never started keep return_code 0
 * (tracking).  May be called from another thread / a signal handler; every hc_track_* call clears it on entry. */
void hc_request_cancel(int32_t on);
const char* hc_last_error(void);
void hc_options_default(hc_options* o);
void* hc_system_create(const hc_program_desc* eval, const hc_program_desc* jac);
void hc_system_destroy(void* s);
void* hc_homotopy_create(const hc_homotopy_desc* d);
void hc_homotopy_destroy(void* h);
/* start_parameters!(T, p), target_parameters!(T, q), parameters!(T, p, q) of a parameter / coefficient homotopy
 * (reference src/endgame_tracker.jl:831-840, src/homotopies/parameter_homotopy.jl:47-64): P complex values each,
 * NULL = keep.  Takes effect for the next batch call. */
int32_t hc_homotopy_set_parameters(void* H, const double* p, const double* q);

int32_t hc_track_batch(void* H, const hc_options* o, int32_t mode, int64_t N, const double* starts,
                       const double* t1, const double* t0, const double* path_p, const double* path_q,
                       const double* omega_mu, hc_results* out, int32_t reserved);
int32_t hc_polyhedral_track_batch(void* Htoric, void* Hcoeff, const hc_options* o, int64_t N, const double* starts,
                                  const int32_t* cell_index, const double* cell_weights, int32_t ncells,
                                  hc_results* out, int32_t reserved);

/* Start solutions produced on the device: nothing but the degrees crosses the bus.
 * Tracks the paths first .. first + N - 1 of the total-degree start system of H = StraightLineHomotopy(G, F):
 * path k starts from x_i = cis(2 pi j_i / d_i), (j_1, ..., j_n) = mixed-radix digits of k, first index fastest --
 * the order of TotalDegreeStartSolutionsIterator (reference src/total_degree.jl:235-262; replaces
 * collect(starts), src/solve.jl:543).  `first` lets every GPU take an index range without a start matrix. */
int32_t hc_track_total_degree(void* H, const hc_options* o, const int32_t* degrees, int64_t first, int64_t N,
                              hc_results* out);
/* Polyhedral start solutions produced on the device (second half): only per-CELL data crosses
 * the bus.  Replaces PolyhedralStartSolutionsIterator + BinomialSystemSolver.X (reference src/polyhedral.jl:104-144,
 * src/binomial_system.jl:55-106, 238-261): the host keeps hnf! and the two n x n solves per cell and sends, per mixed
 * cell c, its volume, the Hermite normal form H_c (n x n int64, row-major, lower triangular: A U = H), the transformed
 * angles mu_c = rem(U^T angle(b) / 2 pi, 2) and the moduli r_c = exp(A^-T log|b|) (both n doubles).  Path k of the call
 * is start solution (first + k) mod (sum of the volumes) in the iterator's order: cells in order, inside a cell the
 * unit-root combinations of fill_unit_roots_combinations! (first coordinate slowest); the device solves the triangular
 * system for the angles in double-double arithmetic as compute_angular_part! does and sets x_j = r_j cis(2 pi alpha_j).
 * cell_weights: ncells x P as in hc_polyhedral_track_batch. */
int32_t hc_polyhedral_track_cells(void* Htoric, void* Hcoeff, const hc_options* o, int64_t first, int64_t N, int32_t ncells,
                                  const int64_t* cell_volume, const int64_t* bin_H, const double* bin_mu, const double* bin_r,
                                  const double* cell_weights, hc_results* out);
/* Many-parameter solve (reference many_solve, src/solve.jl:815-881: the same S start solutions tracked to each
 * of M target parameter vectors): starts is n x S, target_params P x M (column per point).  Path j * S + s =
 * start s to parameter point j; `out` holds S * M paths in that order.  Only the S starts and one parameter
 * column per POINT are copied to the device. */
int32_t hc_track_sweep(void* H, const hc_options* o, int64_t S, const double* starts, int64_t M,
                       const double* target_params, hc_results* out);
/* Duplicate filter of a monodromy round (reference: the UniquePoints lookup of add_tracked_result!,
 * src/monodromy.jl:1176-1200, src/unique_points.jl:247-285): match[i] = index of the first of the M known points
 * (n complex each, point-major) within max(atol, rtol * ||cand_i||_2) of candidate i, or -1.  The loop driver itself
 * stays on the host (monodromy.py mirrors it): three hc_track_batch calls per loop, p -> p1 -> p2 -> p. */
int32_t hc_unique_points_filter(int32_t n, int64_t M, const double* known, int64_t N, const double* cand, double atol, double rtol,
                                int64_t* match);
/* The same sweep with the results reduced on the device (reference: many_solve with a `transform_result`
 * that counts, src/solve.jl:422-430): counts is 5 x M int32 -- per parameter point [nonsingular, singular, real (max |imag|
 * < real_tol, 0 = 1e-6), at infinity, failed] over its S paths, without multiplicity clustering.  20 bytes per point leave
 * the device instead of the PathResults of its paths. */
int32_t hc_track_sweep_counts(void* H, const hc_options* o, int64_t S, const double* starts, int64_t M,
                              const double* target_params, double real_tol, int32_t* counts);
void hc_get_timing(hc_timing* t);

/* Device-resident variant for throughput measurement: inputs are uploaded once, results stay on
 * the device; hc_resident_run may be called repeatedly (kernel only). */
void* hc_resident_create(void* H, void* Hcoeff_or_null, const hc_options* o, int32_t mode, int64_t N, const double* starts,
                         const double* t1, const double* t0, const double* path_p, const double* path_q,
                         const int32_t* cell_index, const double* cell_weights, int32_t ncells);
/* the same for a polyhedral batch whose start solutions are made on the device (arguments of hc_polyhedral_track_cells) */
void* hc_resident_create_cells(void* Htoric, void* Hcoeff, const hc_options* o, int64_t first, int64_t N, int32_t ncells,
                               const int64_t* cell_volume, const int64_t* bin_H, const double* bin_mu, const double* bin_r,
                               const double* cell_weights);
int32_t hc_resident_run(void* r, double* kernel_ms);
int32_t hc_resident_fetch(void* r, hc_results* out);
void hc_resident_destroy(void* r);

/* single point test hooks of the operator API */
int32_t hc_evaluate(void* H, const double* x, const double* t, double* u);
int32_t hc_evaluate_dd(void* H, const double* x_hi, const double* x_lo, const double* t, double* u);
int32_t hc_evaluate_and_jacobian(void* H, const double* x, const double* t, double* u, double* U);
int32_t hc_taylor(void* H, int32_t K, const double* tx, const double* t, double* u);
int32_t hc_toric_set_weights(void* H, const double* w);
/* evaluate! / evaluate_and_jacobian! at N points at once (x: n x N, u: n x N, U: n x n x N or NULL, one t): the batched
 * form of the operator API, e.g. for residual checks of many endpoints (reference src/overdetermined.jl:3-15) */
int32_t hc_evaluate_batch(void* H, int64_t N, const double* x, const double* t, double* u, double* U);

/* Specialised kernels (reference: `compile = true`, src/model_kit/compiled_system_homotopy.jl:178-243).  Large batches
 * (HC_B200_JIT_MIN_PATHS, default 8192; HC_B200_JIT = 0 | 1 | auto) are tracked by a kernel that is generated and
 * compiled for the system at run time (NVRTC, sm_100a): evaluate / Jacobian / Taylor become straight-line code with the
 * tape slots in registers.  hc_jit_prepare builds that kernel ahead of the first batch (it needs no CUDA device and
 * fills the on-disk cache); info, if not NULL, receives {cubin bytes, milliseconds, hot lane-state bytes, cold
 * lane-state bytes, 1 if it came from the cache}.  flags: bit 0 = the variant hc_polyhedral_track_batch runs, bit 1 = the
 * variant for batches with per-path parameter rows (path_p / path_q, hc_track_sweep). */
int32_t hc_jit_prepare(void* H, int32_t flags, double* info);

/* Page-locks (cudaHostRegister) / releases caller-owned host buffers -- start solutions, parameters, the arrays of
 * hc_results -- so that the copies of hc_track_batch run as DMA transfers.  Optional; the Julia host would pin
 * the arrays of its result struct once per solve (reference: results are plain Julia Vectors, src/solve.jl:637). */
int32_t hc_host_register(void* p, int64_t bytes);
int32_t hc_host_unregister(void* p);

/* fp64 FMA pipe microbenchmark (roofline denominator): returns achieved GFLOP/s */
double hc_dfma_peak(int32_t iters);

#ifdef __cplusplus
}
#endif
#endif
