/*
 * petlion_b200.h -- C ABI of the B200-native batched DFN integrator (libpetlion_b200.so).
 *
 * Drop-in boundary for PETLION.jl's hot path.  The reference has no FFI of its own; the seam is
 * the set of Julia callables in `p.funcs` / `Jac_and_res` (src/structures.jl:315-334,
 * src/physics_equations/scalar_residual.jl:435-487) consumed by `initialize_simulation!` and
 * `solve!` (src/model_evaluation.jl:174-232, 312-333).  A Julia maintainer binds these entry
 * points with `ccall`.  Every array is batch-major: row = one system,
 * contiguous doubles.  Nothing returned is library-allocated; the caller owns all buffers.
 *
 * Return value: 0 ok, <0 error (text via plb_last_error()).  Per-system soft exits use the
 * reference's run.info.flag codes 0..11 (src/checks.jl); per-system hard failures are negative.
 * A handle is used by one host thread at a time; calls return after the stream has synchronised.
 */
#ifndef PETLION_B200_H
#define PETLION_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plb_handle_s *plb_handle;

enum { PLB_CATHODE_LCO = 0, PLB_CATHODE_NMC = 1,
       PLB_CATHODE_NMC_LGM50 = 2 /* NMC_LGM50 + LiC6_LGM50 + system_LGM50_NMC_LiC6 (params.jl:514-849): aging = false only */ };
/* reaction rate laws, chosen per electrode (params.jl:51, 114; custom_functions.jl:212-231 rxn_BV, :233-298 rxn_MHC) */
enum { PLB_RXN_BV = 0, PLB_RXN_MHC = 1 };
/* method_I / method_V / method_P (scalar_residual.jl:167-202); PLB_METHOD_DT = the `dT` input of thermal
 * models (constant spatially-averaged temperature: control row val - temperature_weighting(Y'[T]),
 * src/physics_equations/input_methods.jl:182-189; dT=:hold == dT=0);
 * PLB_METHOD_ETA_P = method_eta_p: the plating overpotential Phi_s.n[1] - Phi_e.n[1] held at a value
 * (scalar_residual.jl:92, 199-203; input_methods.jl:108-143) */
enum { PLB_METHOD_I = 0, PLB_METHOD_V = 1, PLB_METHOD_P = 2, PLB_METHOD_DT = 3, PLB_METHOD_ETA_P = 4,
       /* the concentration-rate inputs (input_methods.jl:190-245): the time derivative of the largest / smallest surface
        * concentration of an electrode, or electrolyte concentration, of the previous solution's last point is held at a
        * value (`:hold` = 0).  plb_simulate continuation runs of isothermal models without aging only; not available at
        * operator level (the row depends on the state the run starts from). */
       PLB_METHOD_DC_S_P_MAX = 6, PLB_METHOD_DC_S_P_MIN = 7, PLB_METHOD_DC_S_N_MAX = 8, PLB_METHOD_DC_S_N_MIN = 9,
       PLB_METHOD_DC_E_MAX = 10, PLB_METHOD_DC_E_MIN = 11 };
enum { PLB_MEM_HOST = 0, PLB_MEM_DEVICE = 1 };                /* where the caller's buffers live */

/* petlion(cathode; N_p, N_s, N_n, N_a, N_z, N_r_p, N_r_n, temperature, aging) -- src/params.jl:119-174 */
typedef struct {
    int cathode;
    int N_p, N_s, N_n, N_a, N_z, N_r_p, N_r_n;   /* N_r_p = N_r_n = 10 (every family), 12 or 14 (iso / thermal / SEI, <= 32 x-nodes) */
    int temperature; /* 0: isothermal.  1: temperature=true (LCO only: NMC has no thermal parameters;
                        needs N_p, N_n >= 5 and N_a + N_z <= N_p + N_s + N_n)                  */
    int aging;       /* 0: none.  1: aging=:SEI (LCO, isothermal; adds film, SOH, j_s: N = 322 for 10/10/10) */
    int device;      /* CUDA device ordinal */
    int rxn_p, rxn_n; /* PLB_RXN_*: reaction rate law of the positive / negative electrode (0 = rxn_BV, the default).
                        rxn_MHC adds the keys lambda_MHC_p / lambda_MHC_n to theta (LCO parameter set only) */
    int fickian_spectral; /* 0: Fickian_method = :finite_difference (default); 1: :spectral (params.jl:142, residuals.jl:181-235):
                        Chebyshev collocation in the particles; isothermal, thermal and SEI families on up to 32 x-nodes, N_r = 10 */
} plb_model_desc;

/* run_constant{method,value}: src/structures.jl:46-54; input kinds: src/physics_equations/input_methods.jl:5-74 */
enum { PLB_INPUT_VALUE = 0, PLB_INPUT_HOLD = 1, PLB_INPUT_REST = 2 };
typedef struct {
    int method;      /* PLB_METHOD_* */
    int input_kind;  /* PLB_INPUT_VALUE: a number; PLB_INPUT_HOLD: `:hold` (value taken from the previous
                        state, incl. the reference's V=:hold quirk input_methods.jl:53-63);
                        PLB_INPUT_REST: I = :rest (value 0; bounds not checked, src/checks.jl:12,388) */
    double value;    /* used for PLB_INPUT_VALUE when `values` passed to the call is NULL */
    double tf;       /* final time of this run (seconds, local to the run); reference default 1e6 */
    int new_run;     /* 1: simulate(); 0: simulate!() continuation from state_* */
    int reserved;
} plb_run;

/* options_simulation: src/structures.jl:266-285; defaults src/params.jl:256-280 */
typedef struct {
    double abstol, reltol, abstol_init, reltol_init;
    int maxiters;
    int check_bounds;
    int interp_final;
    int skip_alg_deriv;  /* 0 (default): newtons_method! estimates dY_alg/dt (model_evaluation.jl:462-477);
                            1: its keyword initialize_algebraic_derivatives=false (:433) -- Y'_alg = 0 at the start
                            of a run, which is also how PETLION versions before that estimate behaved */
} plb_opts;

/* boundary_stop_conditions: src/structures.jl:237-251 (NaN deactivates a bound) */
typedef struct {
    double V_max, V_min, SOC_max, SOC_min, T_max, c_s_n_max, I_max, I_min, eta_plating_min,
        c_e_min, dfilm_max;
} plb_bounds;

/* per-system summary record (80 bytes) */
typedef struct {
    double t_end, V_end, I_end, SOC_end;
    double T_end;      /* temperature_weighting(T) of the final state [K] (T0 for isothermal models) */
    double aux_end;    /* SOH of the final state when aging=:SEI, else 0 */
    int flag;          /* run.info.flag: 0 tf, 1 V_min, 2 V_max, 3 SOC_min, 4 SOC_max, 5 T_max, 6 c_s_n,
                          7 I_max, 8 I_min, 9 c_e_min, 10 dfilm, 11 eta_plating; <0 hard failure */
    int n_steps;       /* accepted integrator steps */
    int n_res, n_jac;  /* residual / Jacobian evaluations */
    int n_netf, n_ncfn;/* error-test / Newton-convergence failures */
    int n_newton_init; /* iterations of the algebraic initialisation */
    int n_reinit;      /* re-initialisations at input discontinuities (checks.jl:341-364; table runs) */
} plb_summary;

#define PLB_FAIL_NEWTON_INIT (-1) /* "Could not initialize DAE" model_evaluation.jl:456 */
#define PLB_FAIL_CONV (-2)        /* "Model failed to converge"  checks.jl:233-236        */
#define PLB_FAIL_ERRTEST (-3)
#define PLB_FAIL_MAXITERS (-4)    /* checks.jl:239 */
#define PLB_FAIL_NONFINITE (-5)
#define PLB_FAIL_INIT_BOUNDS (-6) /* check_initial_SOC, checks.jl:327-339 */
#define PLB_FAIL_PREVIOUS (-7)    /* simulate!() of a system whose earlier segment failed hard: a hard failure leaves
                                     state_t = NaN, and a continuation passes such a system through untouched (the
                                     reference would have thrown at the first failure) */

const char *plb_last_error(void);

/* petlion(...) : build the model handle (device workspaces, index tables). */
int plb_create(const plb_model_desc *desc, plb_handle *out);
int plb_destroy(plb_handle h);
/* optional: CUDA stream (cudaStream_t passed as void*) used for PLB_MEM_DEVICE calls */
int plb_set_stream(plb_handle h, void *cuda_stream);

/* sizes: N.tot, N.diff, number of used parameters, nnz of J_full.sp for a method */
int plb_nstates(plb_handle h);
int plb_ndiff(plb_handle h);
int plb_ntheta(plb_handle h);
int plb_jac_nnz(plb_handle h, int method);

/* theta_keys of the generated functions: alphabetically sorted used keys, UTF-8
 * (src/generate_functions.jl:327-363, 387).  keys[i] points to static storage. */
int plb_theta_keys(plb_handle h, const char **keys);
int plb_theta_index(plb_handle h, const char *key_utf8);
/* default parameter row (LCO()/LiC6()/system_LCO_LiC6 ...: src/params.jl) and default bounds/opts */
int plb_theta_defaults(plb_handle h, double *theta_row);
int plb_bounds_defaults(plb_handle h, plb_bounds *b);
int plb_opts_defaults(plb_handle h, plb_opts *o);
/* calc_I1C: src/physics_equations/auxiliary_states_and_coefficients.jl:631-647 */
int plb_calc_I1C(plb_handle h, int B, const double *theta, double *I1C);

/* J_full.sp: CSC pattern of [J_sp_base; J_sp_scalar'] (scalar_residual.jl:501); index base 0 or 1 */
int plb_jac_pattern(plb_handle h, int method, int *colptr, int *rowval, int one_based);

/* initial_guess!(Y0, SOC, theta_tot, X_applied) -- src/model_evaluation.jl:204 */
int plb_initial_guess(plb_handle h, int B, const double *soc, const double *theta, double *Y0, int mem);

/* R_full(res,t,Y,YP,p,run) -- model_evaluation.jl:263 ;  J_full(J,t,Y,YP,gamma,p,run) -- :264
 * values: per-system control value [B] or NULL (then run->value).  Either output may be NULL.
 * There is no `t` argument: the generated residual depends on time only through a run_function's value
 * (scalar_residual.jl:169-170), and a closure cannot cross this boundary -- the caller evaluates its input at t and
 * passes the numbers in `values` (that is what plb_simulate_table does on the device, per step). */
int plb_resjac(plb_handle h, int B, const double *Y, const double *YP, const double *gamma,
               const double *theta, const plb_run *run, const double *values, double *res,
               double *nzval, int mem);

/* newtons_method!(p,Y,YP,run,opts,R_alg,R_diff,J_alg) -- model_evaluation.jl:430-480.
 * Y in/out [B x N], YP out [B x N]; status[B] = iterations (>0) or PLB_FAIL_NEWTON_INIT */
int plb_newton_init(plb_handle h, int B, double *Y, double *YP, const double *theta,
                    const plb_run *run, const double *values, const plb_opts *opts, int *status,
                    int mem);

/* KLU's role inside IDA (model_evaluation.jl:265-271, 417-428): evaluate J = dF/dY + gamma dF/dY' at
 * (Y, Y'), factorise it with the structured solver the integrator uses, and solve J x = rhs.
 * All arrays [B x N] in the reference state order; gamma[B]; status[B] (optional) 0 ok / -1 singular. */
int plb_linear_solve(plb_handle h, int B, const double *Y, const double *YP, const double *gamma,
                     const double *theta, const plb_run *run, const double *values,
                     const double *rhs, double *x, int *status, int mem);

/* simulate(p, tf; I|V|P=..., SOC, bounds..., opts...) / simulate!(sol, p, ...):
 * replaces initialize_simulation! + IDA + solve! + exit_simulation!
 * (model_evaluation.jl:10-97, 174-382; checks.jl:1-249; save_outputs.jl:11-40).
 *   soc0[B]                      initial SOC (new runs)
 *   state_Y[BxN], state_YP[BxN], state_SOC[B], state_t[B]   continuation state, in/out
 *   summary[B]
 *   traj_*[B x n_save_max] (optional, may be NULL; n_save_max may be 0), traj_n[B];
 *   traj_T = temperature_weighting(T) per saved step (thermal models; T0 otherwise)
 *   traj_Y[B x n_save_max x N] (optional): the full state of every saved row in the reference state order
 *   -- what set_vars! stores for outputs = :all / (:c_e, :c_s_avg, :j, ...) (save_outputs.jl:11-40)
 */
int plb_simulate(plb_handle h, int B, const double *theta, const plb_run *run,
                 const double *values, const plb_opts *opts, const plb_bounds *bounds,
                 const double *soc0, double *state_Y, double *state_YP, double *state_SOC,
                 double *state_t, plb_summary *summary, int n_save_max, double *traj_t,
                 double *traj_V, double *traj_I, double *traj_SOC, double *traj_T, double *traj_Y,
                 int *traj_n, int mem);

/* run_function{method,func} (src/structures.jl:55-63; examples/variable_input_functions.ipynb): a
 * time-varying I(t) / V(t) / P(t) / eta_p(t).  A Julia closure cannot cross the C ABI, so the function is
 * a piecewise-linear table of the run's LOCAL time ("all times start at t = 0, even if the simulation
 * follows another one"): knots t[n] non-decreasing, values v[n]; a repeated knot time is a jump, the
 * function being right-continuous there (`t < 100 ? 1 : 0.5` is t = {0,100,100}, v = {1,1,0.5});
 * constant outside the table.  tdiscon = opts.tdiscon (src/structures.jl:279): the integrator is made to
 * stop at tdiscon - reltol/2 (model_evaluation.jl:295-297).  Table and tdiscon are HOST arrays. */
typedef struct {
    int n;
    const double *t, *v;
    int n_tdiscon;
    const double *tdiscon;   /* ascending; may be NULL when n_tdiscon == 0 */
} plb_input_table;

/* simulate(p, tf; I = I_fun, tdiscon = [...]) with I_fun given by `table`.  Same arguments as plb_simulate;
 * run->input_kind must be PLB_INPUT_VALUE, run->value is ignored and `scale[B]` (optional, `mem` decides
 * host/device) multiplies the table per system: value_b(t) = scale[b] * table(t).
 * Mirrors scalar_residual! for run_function (scalar_residual.jl:169-170: the function is evaluated at
 * every residual evaluation), initial_current! (input_methods.jl:27-29, 64-74, 104-107), the stop list of
 * postfix_integrator! (model_evaluation.jl:288-310), check_solve for run_function (checks.jl:251-268: a
 * failed step is not an error) and check_reinitialization! (checks.jl:341-364). */
int plb_simulate_table(plb_handle h, int B, const double *theta, const plb_run *run,
                       const plb_input_table *table, const double *scale, const plb_opts *opts,
                       const plb_bounds *bounds, const double *soc0, double *state_Y,
                       double *state_YP, double *state_SOC, double *state_t, plb_summary *summary,
                       int n_save_max, double *traj_t, double *traj_V, double *traj_I,
                       double *traj_SOC, double *traj_T, double *traj_Y, int *traj_n, int mem);

/* p.opts.tstops (src/params.jl:272, "times when the DAE solver explicitly stops"; merged into the stop list by
 * postfix_integrator!, model_evaluation.jl:291-293): local times of the run, applied to every later simulate call
 * of this handle; n = 0 clears them.  HOST array. */
int plb_set_tstops(plb_handle h, int n, const double *tstops);

/* simulate(p, tf::AbstractVector; ...) (src/model_evaluation.jl:13, 80 -> interp_sol, :148-149): results at
 * user-requested times.  The reference re-interpolates the saved steps with a cubic spline on the host afterwards
 * (src/save_outputs.jl:74-133; the host layer keeps that as sol(t)); here the integrator itself evaluates its BDF
 * interpolant (IDAGetSolution) of the step that covers each requested time, on the device, while it steps.
 * t_global[n]: ascending GLOBAL times (HOST array, copied).  V/I/SOC/T [B x n], Y [B x n x N] (reference state
 * order), n_done[B] = rows filled per system: any of them may be NULL; they live where the `mem` of the simulate
 * call says and must stay valid until it returns.  Rows past the end of a run are NaN.  Applies to the NEXT
 * plb_simulate / plb_simulate_table call of this handle only; n = 0 clears a pending request. */
int plb_set_dense_output(plb_handle h, int n, const double *t_global, double *V, double *I, double *SOC,
                         double *T, double *Y, int *n_done, int mem);

/* ---- multi-GPU fan-out inside the library ("multi-GPU fan-out is internal"; 8(e)) -----------------
 * One model on several GPUs of the box.  Simulations are independent, so the path shards with NO data-path collective:
 * device k owns the contiguous block [k*R, min(B, (k+1)*R)) of systems, R = ceil(B / n_dev); every shard runs
 * concurrently on its own device, driven by its own host thread, and writes straight into the caller's HOST arrays.
 * The one collective of the path follows: an ncclAllGather of the fixed-size (80-byte) summaries, after which EVERY
 * device holds the summaries of the whole batch (plb_group_device_summaries) -- what a device-side consumer of a
 * sweep (a reduction, a parameter update) needs.  NCCL is loaded at run time (libnccl.so.2); without it the group still
 * simulates and only the device-side gather is unavailable.  tstops / dense-output / table requests are per-device
 * features (plb_group_handle) and are not fanned out. */
typedef struct plb_group_s *plb_group;
int plb_group_create(const plb_model_desc *desc, int n_dev, const int *devices, plb_group *out);  /* desc->device is ignored */
int plb_group_destroy(plb_group g);
int plb_group_size(plb_group g);
plb_handle plb_group_handle(plb_group g, int k);   /* device k's handle: sizes, keys, patterns, defaults */
/* plb_simulate over the group; every array is a HOST array of the whole batch (traj_Y is not fanned out) */
int plb_group_simulate(plb_group g, int B, const double *theta, const plb_run *run, const double *values,
                       const plb_opts *opts, const plb_bounds *bounds, const double *soc0, double *state_Y,
                       double *state_YP, double *state_SOC, double *state_t, plb_summary *summary,
                       int n_save_max, double *traj_t, double *traj_V, double *traj_I, double *traj_SOC,
                       double *traj_T, int *traj_n);
/* after plb_group_simulate: device pointer, on device k, to the all-gathered summaries [n_dev x R] (row b of the batch
 * is entry (b / R) * R + b % R = b; the tail rows of the last block are zero); returns -1 if NCCL is unavailable */
int plb_group_device_summaries(plb_group g, int k, const plb_summary **out, int *rows);
/* milliseconds of the last all-gather (CUDA events on device 0's stream), 0 if none */
float plb_group_last_gather_ms(plb_group g);

/* diagnostic: launch geometry of a compiled model family (0 isothermal, 1 thermal, 2 SEI, 3 wide, 4 wide SEI): out[8] = {integrator warps
 * per CTA, CTAs per SM, dynamic shared memory per CTA [B], K1 warps per CTA, K1 CTAs per SM, K1 shared
 * memory per CTA [B], workspace vector stride, Jacobian slots per lane}.  Needs no GPU. */
int plb_variant_info(int family, long long *out);

/* kernel launch counter (number of CUDA kernels this handle has launched) */
long long plb_launch_count(plb_handle h);
/* last kernel time in ms measured with CUDA events on the launching stream (0 if none) */
float plb_last_kernel_ms(plb_handle h);

#ifdef __cplusplus
}
#endif
#endif
