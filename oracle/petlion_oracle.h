/*
 * petlion_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the PETLION.jl hot path (reference @ /root/reference,
 * v1.0.6): generated residual/Jacobian callback surface, Newton initialisation of
 * the algebraic block, the SUNDIALS-IDA variable-order BDF stepper the reference
 * drives one step at a time, and the host-side stop/interpolation logic.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (petlion.jl_b200/) never links or imports it.
 *
 * Parity status: the reference cannot be executed here (no Julia; IDA and KLU are
 * un-vendored third-party C libraries).  The oracle is pinned against the reference's
 * own executed-notebook outputs (tests/golden/reference_goldens.json): I1C, V(0+) and
 * V[1:13] of the 2C charge, the c_e rows, the IDA step ladders of three 1C discharges, and
 * (temperature = true) every printed digit and all 76 steps of the thermal 4C charge plus the
 * first 28 steps of its dT = :hold continuation.  See DESIGN.md section 2.
 * PARITY UNPINNED for the aging = :SEI rows (film, SOH, j_s: residuals.jl:260-297, 519-552): the
 * reference ships no executed SEI example, so that part is a restatement without a known answer.
 */
#ifndef PETLION_ORACLE_H
#define PETLION_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Full parameter table, oracle order (ASCII names; reference unicode names in comments).
 * Values: src/params.jl:5-117 (LCO, LiC6), :180-226 (system_LCO_LiC6),
 *         :295-367 (NMC, LiC6_NMC), :428-445 (system_NMC_LiC6).                      */
#define ORC_THETA_FIELDS(X)                                                             \
    X(D_n) X(D_p) X(D_s) X(D_sn) X(D_sp) X(Ea_D_sn) X(Ea_D_sp) X(Ea_k_n) X(Ea_k_p)      \
    X(Rp_n) X(Rp_p) X(T0) /* T₀ */ X(T_amb) X(brugg_n) X(brugg_p) X(brugg_s)            \
    X(c_e0) /* c_e₀ */ X(c_max_n) X(c_max_p) X(k_n) X(k_p)                              \
    X(l_a) X(l_n) X(l_p) X(l_s) X(l_z) X(t_plus) /* t₊ */                               \
    X(theta_max_n) X(theta_max_p) X(theta_min_n) X(theta_min_p) /* θ_* */               \
    X(sigma_a) X(sigma_n) X(sigma_p) X(sigma_z) /* σ_* */                               \
    X(eps_fn) X(eps_fp) X(eps_n) X(eps_p) X(eps_s) /* ϵ_* */                            \
    X(Cp_a) X(Cp_n) X(Cp_p) X(Cp_s) X(Cp_z) X(h_cell)                                   \
    X(lambda_a) X(lambda_n) X(lambda_p) X(lambda_s) X(lambda_z) /* λ_* */               \
    X(rho_a) X(rho_n) X(rho_p) X(rho_s) X(rho_z) /* ρ_* */                              \
    X(M_n) X(R_SEI) X(Uref_s) X(i_0_jside) X(k_n_aging) X(w)                            \
    X(lambda_MHC_n) X(lambda_MHC_p) /* λ_MHC_* (rxn_MHC only) */                        \
    X(D_e) /* electrolyte diffusivity scale of D_eff_LGM50 (NMC_LGM50 only) */

typedef struct {
#define X(n) double n;
    ORC_THETA_FIELDS(X)
#undef X
} orc_theta;

#define ORC_NTHETA ((int)(sizeof(orc_theta) / sizeof(double)))

enum { ORC_CATHODE_LCO = 0, ORC_CATHODE_NMC = 1,
       ORC_CATHODE_LGM50 = 2 /* NMC_LGM50 + LiC6_LGM50 + system_LGM50_NMC_LiC6 (Chen et al. 2020), params.jl:514-849 */ };
enum { ORC_RXN_BV = 0, ORC_RXN_MHC = 1 };   /* custom_functions.jl:212-231, 241-298 */
/* method_I / method_V / method_P (scalar_residual.jl:167-202) and `dT` = the constant_temperature
 * residual of input_methods.jl:182-189 (control row  val - temperature_weighting(Y'[T])).
 * ORC_METHOD_DT_ALG is internal: the same row inside newtons_method!, where the reference substitutes
 * Y'_diff -> rhs_diff(Y) (scalar_residual.jl:347-363). */
/* ORC_METHOD_ETA: method_eta_p, the plating overpotential Phi_s.n[1] - Phi_e.n[1] (scalar_residual.jl:92, 199-203) */
enum { ORC_METHOD_I = 0, ORC_METHOD_V = 1, ORC_METHOD_P = 2, ORC_METHOD_DT = 3, ORC_METHOD_ETA = 4, ORC_METHOD_DT_ALG = 5,
       /* the concentration-rate inputs dc_s_p_max/min, dc_s_n_max/min, dc_e_max/min (input_methods.jl:190-245): a
        * run_residual  val - Y'[ind]  with ind = the arg-max / arg-min state of the previous solution's last point
        * (orc_run::dc_kind selects which, orc_run::dc_ind is resolved at the start of the run); _ALG: the same row
        * inside newtons_method! (Y'[ind] -> rhs[ind], scalar_residual.jl:347-363) */
       ORC_METHOD_DC = 6, ORC_METHOD_DC_ALG = 7 };
enum { ORC_DC_S_P_MAX = 0, ORC_DC_S_P_MIN, ORC_DC_S_N_MAX, ORC_DC_S_N_MIN, ORC_DC_E_MAX, ORC_DC_E_MIN };

/* model structure: petlion(cathode; N_p, ..., temperature, aging) -- src/params.jl:119-174 */
typedef struct {
    int N_p, N_s, N_n, N_a, N_z, N_r_p, N_r_n;
    int temperature; /* 0 isothermal, 1 thermal   */
    int aging;       /* 0 none, 1 :SEI            */
    int cathode;     /* ORC_CATHODE_*             */
    int rxn_p, rxn_n; /* ORC_RXN_*: rxn_BV (default) or rxn_MHC, per electrode (params.jl:51, 114) */
    int fickian_spectral; /* 0: Fickian_method = :finite_difference (default), 1: :spectral (params.jl:142, residuals.jl:181-235) */
} orc_model;

/* index layout -- src/external.jl:275-365, SURVEY App. A (0-based here) */
typedef struct {
    int Nx;              /* N_p+N_s+N_n */
    int c_e, c_s_p, c_s_n, T, film, SOH; /* differential starts (-1 if inactive) */
    int j, phi_e, phi_s, j_s, iI;        /* algebraic starts */
    int N_diff, N_alg, N_tot;
} orc_layout;

struct orc_dense;
/* run = run_constant{method,value}: src/structures.jl:46-54 */
typedef struct {
    int method;   /* ORC_METHOD_* */
    double value; /* applied I [C-rate], V [V] or P [W/m^2] */
    double tf;    /* final (local) time */
    int input_kind; /* 0 number; 1 :hold (value from previous state); 2 :rest (checks.jl:12,388) */
    int new_run;  /* 1: fresh simulate(); 0: simulate!() continuation (adds tstop 1.0) */
    double t0;    /* global time offset (run.t0) */
    /* run_function{method,func} (structures.jl:55) with func(t) restricted to a piecewise-linear table:
     * knots tab_t[tab_n] non-decreasing in the run's LOCAL time (variable_input_functions.ipynb: "all times
     * start at t = 0"), values tab_v[tab_n]; a repeated knot time is a jump and the function is
     * right-continuous there (`t < 100 ? 1 : 0.5` == knots (0,1),(100,1),(100,0.5)); constant outside the
     * table.  value(t) = scale * table(t).  tab_n == 0: run_constant.  tdiscon = opts.tdiscon
     * (structures.jl:279), ascending.  last_value mirrors run.value[] (scalar_residual.jl:170). */
    int tab_n;
    const double *tab_t, *tab_v;
    double scale;
    int n_tdiscon;
    const double *tdiscon;
    double *last_value;
    int n_tstops;            /* opts.tstops (params.jl:272): explicit stop times, local to the run */
    const double *tstops;
    /* dense output (`tf::AbstractVector`, model_evaluation.jl:80, 148-149): the integrator's own BDF
     * interpolant (IDAGetSolution) evaluated at the requested GLOBAL times, ascending; rows of system
     * `dense_sys` in the [B][n] output arrays (any of them may be NULL).  dense_done[sys] = rows filled
     * (times past the end of the run are left untouched). */
    const struct orc_dense *dense;
    int dense_sys;
    int dc_kind;             /* ORC_DC_* (method ORC_METHOD_DC) */
    int dc_ind;              /* the state index the row holds; filled in by the run (0-based) */
} orc_run;

typedef struct orc_dense {
    int n;
    const double *t;                 /* [n] */
    double *V, *Icur, *SOC, *T;      /* [B][n] */
    double *Y;                       /* [B][n][N_tot] */
    int *done;                       /* [B] */
} orc_dense;

/* options_simulation: src/structures.jl:266-285, defaults src/params.jl:256-280 */
typedef struct {
    double abstol, reltol, abstol_init, reltol_init;
    int maxiters;
    int check_bounds, interp_final;
    /* IDA knobs as set by Sundials.jl's IDA() constructor */
    int ida_maxord;  /* 5  */
    int ida_maxcor;  /* max_nonlinear_iters */
    int ida_maxnef;  /* max_error_test_failures */
    int ida_maxncf;  /* max_convergence_failures */
    int skip_alg_deriv; /* 1: newtons_method!(...; initialize_algebraic_derivatives=false) (model_evaluation.jl:433):
                           Y'_alg = 0 at the start of a run, as PETLION versions before that estimate did */
} orc_opts;

/* boundary_stop_conditions: src/structures.jl:237-251 (NaN deactivates) */
typedef struct {
    double V_max, V_min, SOC_max, SOC_min, T_max, c_s_n_max, I_max, I_min, eta_plating_min,
        c_e_min, dfilm_max;
} orc_bounds;

/* per-simulation result summary */
typedef struct {
    double t_end, V_end, I_end, SOC_end;
    double T_end;  /* temperature_weighting(T) of the final state (T0 for isothermal models) */
    int flag;      /* 0..11 as checks.jl ; <0 hard failure */
    int n_steps;   /* accepted IDA steps (= saved points - 1) */
    int n_res, n_jac, n_netf, n_ncfn, n_newton_init;
    int n_reinit;  /* re-initialisations at input discontinuities (checks.jl:341-364) */
} orc_summary;

/* negative (hard-failure) flags */
#define ORC_FAIL_NEWTON_INIT (-1)
#define ORC_FAIL_CONV (-2)
#define ORC_FAIL_ERRTEST (-3)
#define ORC_FAIL_MAXITERS (-4)
#define ORC_FAIL_NONFINITE (-5)
#define ORC_FAIL_INIT_BOUNDS (-6)
#define ORC_FAIL_PREVIOUS (-7)    /* continuation of a system whose earlier segment failed (state_t is NaN) */

int orc_ntheta(void);
const char *orc_theta_name(int i);
void orc_theta_defaults(int cathode, double *theta /*[ORC_NTHETA]*/);
void orc_bounds_defaults(int cathode, orc_bounds *b);
void orc_opts_defaults(orc_opts *o);
void orc_layout_make(const orc_model *m, orc_layout *L);
double orc_calc_I1C(const double *theta);

/* callback surface */
void orc_initial_guess(const orc_model *m, const double *theta, double SOC, double *Y0);
void orc_residual(const orc_model *m, const double *theta, const orc_run *run, double t,
                  const double *Y, const double *YP, double *res);
/* CSC pattern (0-based) of [J_sp_base; J_sp_scalar'] -- scalar_residual.jl:501.
 * Call with colptr==NULL to get nnz. */
int orc_jac_pattern(const orc_model *m, int method, int *colptr, int *rowval);
/* nzval of dF/dY + gamma dF/dY' in that CSC order (coloured complex-step: the reference's
 * jacobian=:AD route, generate_functions.jl:166-235, exact to round-off) */
void orc_jacobian(const orc_model *m, const double *theta, const orc_run *run, double t,
                  const double *Y, const double *YP, double gamma, double *nzval);

/* newtons_method! -- src/model_evaluation.jl:430-480.  returns #iterations or <0 */
int orc_newton_init(const orc_model *m, const double *theta, const orc_run *run,
                    const orc_opts *o, double *Y, double *YP);

/* simulate()/simulate!() for one system.  state_Y/state_YP/state_SOC/state_t are in/out
 * (continuation state).  traj_* may be NULL; otherwise length n_save_max.             */
int orc_simulate(const orc_model *m, const double *theta, const orc_run *run,
                 const orc_opts *o, const orc_bounds *b, double SOC0, double *state_Y,
                 double *state_YP, double *state_SOC, double *state_t, orc_summary *out,
                 int n_save_max, double *traj_t, double *traj_V, double *traj_I,
                 double *traj_SOC, int *traj_n);

/* batch driver (OpenMP over systems): theta[B][ORC_NTHETA]; value[B] */
int orc_simulate_batch(const orc_model *m, int B, const double *theta, const orc_run *run,
                       const double *values, const orc_opts *o, const orc_bounds *b,
                       const double *SOC0, double *state_Y, double *state_YP,
                       double *state_SOC, double *state_t, orc_summary *out, int n_save_max,
                       double *traj_t, double *traj_V, double *traj_I, double *traj_SOC,
                       int *traj_n, int nthreads);

/* counter-based RNG shared by CPU and GPU sides: u in [0,1) */
/* piecewise-linear table of a run_function (see orc_run) */
double orc_table_eval(int n, const double *tt, const double *vv, double t);

double orc_rng_u01(unsigned long long seed, unsigned long long system_id, unsigned param_id);

#ifdef __cplusplus
}
#endif
#endif
