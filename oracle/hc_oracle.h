/* ORACLE (test infrastructure, NOT product code): C interface of the CPU restatement of
 * HomotopyContinuation.jl's path-tracking hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * The plain-data structs deliberately have the same layout as include/hc_b200.h so that the
 * Python test harness can drive both libraries with the same ctypes definitions. */
#ifndef HC_ORACLE_H
#define HC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const int32_t* instructions; /* 6 int32 per instruction: in[4], op, out (1-based); last = OP_STOP */
    int32_t n_instructions;
    const double* constants; /* re,im interleaved */
    int32_t n_constants;
    int32_t param_offset, n_params; /* parameters at tape[param_offset+1 ...] */
    int32_t t_index;                /* 0 = none */
    int32_t var_offset, n_vars;
    const int32_t* u_assign; /* pairs (i, k) */
    int32_t n_u;
    const int32_t* U_assign; /* pairs (j, k), j column-major over (out_dim, n_vars) */
    int32_t n_U;
    int32_t out_dim, tape_space;
} orc_program_desc;

typedef struct {
    int32_t kind; /* 0 straight line, 1 parameter, 2 coefficient, 3 toric */
    void* F;
    void* G;
    double gamma[2];
    const double* G_params; int32_t n_G_params;
    const double* F_params; int32_t n_F_params;
    const double* p; /* start params (t=1) / toric system coefficients */
    const double* q; /* target params (t=0) */
    int32_t n_pq;
} orc_homotopy_desc;

typedef struct {
    /* TrackerOptions (src/tracker.jl:94-140) */
    int32_t max_steps;
    double max_step_size, max_initial_step_size;
    int32_t extended_precision;
    double min_step_size, min_rel_step_size;
    /* TrackerParameters (src/tracker.jl:45-62) */
    double a, beta_a, beta_omega_p, beta_tau, strict_beta_tau;
    int32_t min_newton_iters;
    /* EndgameOptions (src/endgame_tracker.jl:47-72) */
    double endgame_start;
    int32_t max_endgame_steps, max_endgame_extended_steps;
    double min_cond, min_cond_growth, min_coord_growth;
    int32_t zero_is_at_infinity, at_infinity_check, only_nonsingular;
    double singular_min_accuracy;
    int32_t max_winding_number;
    double val_finite_tol, val_at_infinity_tol, sing_cond, sing_accuracy, scaling_threshold;
    int32_t refine_steps;
    /* WeightedNormOptions (src/norm.jl:36-40) */
    double scale_min, scale_abs_min, scale_max;
} orc_options;

typedef struct { /* caller-allocated SoA; fields of PathResult (src/path_result.jl:76-98) */
    int32_t* return_code;
    double* solution; /* 2n x N */
    double* t;
    double* accuracy;
    double* residual;
    uint8_t* singular;
    double* condition_jacobian;
    int32_t* winding_number; /* 0 = nothing */
    uint8_t* extended_precision;
    double* last_point; /* 2n x N */
    double* last_t;
    double* valuation; /* n x N */
    uint8_t* has_valuation;
    double* omega;
    double* mu;
    int32_t* accepted_steps;
    int32_t* rejected_steps;
    int32_t* steps_eg;
    uint8_t* extended_precision_used;
    int64_t* counters; /* optional, 8 x N: factorizations, ldivs, 0... */
} orc_results;

void orc_options_default(orc_options* o);
void* orc_system_create(const orc_program_desc* eval, const orc_program_desc* jac);
void orc_system_destroy(void* s);
void* orc_homotopy_create(const orc_homotopy_desc* d);
void orc_homotopy_destroy(void* h);

/* mode 0: EndgameTracker.track(x, t1 real) ; mode 1: Tracker.track(x, t1, t0) (codes = TrackerCode)
 * path_p / path_q: optional per-path start / target parameters (P x N complex), else NULL. */
int32_t orc_track_batch(void* H, const orc_options* o, int32_t mode, int64_t N, const double* starts,
                        const double* t1, const double* t0, const double* path_p, const double* path_q,
                        const double* omega_mu, /* optional 2 x N (omega, mu), NaN = unset */
                        orc_results* out, int32_t nthreads);
int32_t orc_polyhedral_track_batch(void* Htoric, void* Hcoeff, const orc_options* o, int64_t N, const double* starts,
                                   const int32_t* cell_index, const double* cell_weights, int32_t ncells,
                                   orc_results* out, int32_t nthreads);

/* test hooks */
int32_t orc_evaluate(void* H, const double* x, const double* t, double* u);
int32_t orc_evaluate_dd(void* H, const double* x_hi, const double* x_lo, const double* t, double* u);
int32_t orc_evaluate_and_jacobian(void* H, const double* x, const double* t, double* u, double* U);
int32_t orc_taylor(void* H, int32_t K, const double* tx, const double* t, double* u);
int32_t orc_toric_set_weights(void* H, const double* w);
void orc_la_solve(int32_t n, const double* A, const double* b, const double* weights, int32_t refine, double* x);
double orc_la_cond(int32_t n, const double* A, const double* d_l, const double* d_r);
double orc_la_inverse_inf_norm_est(int32_t n, const double* A, const double* d_l, const double* d_r);
void orc_dd_op(int32_t op, const double* a, const double* b, int32_t p, double* out);
void orc_stepper_trace(const double* start, const double* target, const double* ds, int32_t k, double* out);
void orc_norm_test(int32_t n, const double* x, const double* y, const double* w, double* out);
double orc_nthroot(double x, int32_t n);
void orc_taylor_op(int32_t op, int32_t K, const double* a, const double* b, const double* c, const double* d, int32_t p, double* out);
void orc_valuation_trace(int32_t n, int32_t steps, const double* tx, const double* t, double* out);

#ifdef __cplusplus
}
#endif
#endif
