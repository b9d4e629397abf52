// plb_common.cuh -- plain types shared by the host side (plb_kernels.cu) and the two device variants.
//
// The device code (plb_device.cuh, plb_integrator.cuh, plb_tick.cuh, plb_variant.cuh) is compiled once per
// model family, into namespaces plb::iso (isothermal, N = 301), plb::th (temperature = true, N = 351) and
// plb::sei (aging = :SEI, N = 322): plb_variant_iso.cu / plb_variant_th.cu / plb_variant_sei.cu.  The reference does the same thing at model-build
// time: `petlion(...; temperature=true)` generates a different residual/Jacobian
// (/root/reference/src/generate_functions.jl:102-164, src/params.jl:119-174).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plb {

constexpr double kF = 96485.3321233;     // const_Faradays,  structures.jl:10
constexpr double kR = 8.31446261815324;  // const_Ideal_Gas, structures.jl:11
constexpr double kTref = 298.15;
constexpr unsigned FULL = 0xffffffffu;

enum { CHEM_LCO = 0, CHEM_NMC = 1,
       CHEM_LGM = 2 /* NMC_LGM50 / LiC6_LGM50 (params.jl:514-849): compiled in sibling builds only (plb_variant_{iso,th}lgm.cu) */ };
// control row (scalar_residual.jl:167-202): applied current / voltage / power, and the constant-temperature
// mode dT (input_methods.jl:182-189): val - temperature_weighting(Y'[T]).  METHOD_DT_ALG is the same row as
// newtons_method! sees it, with Y'_T replaced by the right-hand side of the T rows (scalar_residual.jl:347-363).
// METHOD_ETA: method_eta_p, the plating overpotential Phi_s.n[1] - Phi_e.n[1] held at a value (scalar_residual.jl:92, 199-203)
enum { METHOD_I = 0, METHOD_V = 1, METHOD_P = 2, METHOD_DT = 3, METHOD_ETA = 4, METHOD_DT_ALG = 5,
       // the concentration-rate inputs dc_s_p_max .. dc_e_min (input_methods.jl:190-245): control row val - Y'[ind] with ind
       // chosen from the previous solution at the start of the run; _ALG: its newtons_method! form (Y'[ind] -> rhs[ind]).
       // Only the families compiled with PLB_DC carry them (plb_variant_isodc.cu, plb_variant_widedc.cu).  A lane_eval
       // call passes  method | target lane << 8 | component << 16  (component 0: surface c_s, 1: c_e).
       METHOD_DC = 6, METHOD_DC_ALG = 7 };
enum { DC_S_P_MAX = 0, DC_S_P_MIN, DC_S_N_MAX, DC_S_N_MIN, DC_E_MAX, DC_E_MIN };
constexpr int N_METHODS = 5;

// canonical parameter fields (ASCII names of the reference keys); the thermal block is only part of a
// theta row when temperature = true
enum ThetaField {
    TF_D_n, TF_D_p, TF_D_s, TF_D_sn, TF_D_sp, TF_Ea_D_sn, TF_Ea_D_sp, TF_Ea_k_n, TF_Ea_k_p,
    TF_Rp_n, TF_Rp_p, TF_T0, TF_brugg_n, TF_brugg_p, TF_brugg_s, TF_c_e0, TF_c_max_n, TF_c_max_p,
    TF_k_n, TF_k_p, TF_l_n, TF_l_p, TF_l_s, TF_t_plus, TF_theta_max_n, TF_theta_max_p,
    TF_theta_min_n, TF_theta_min_p, TF_sigma_n, TF_sigma_p, TF_eps_fn, TF_eps_fp, TF_eps_n,
    TF_eps_p, TF_eps_s,
    // thermal (params.jl:5-117, 177-226)
    TF_Cp_a, TF_Cp_n, TF_Cp_p, TF_Cp_s, TF_Cp_z, TF_T_amb, TF_h_cell, TF_l_a, TF_l_z,
    TF_lambda_a, TF_lambda_n, TF_lambda_p, TF_lambda_s, TF_lambda_z,
    TF_rho_a, TF_rho_n, TF_rho_p, TF_rho_s, TF_rho_z, TF_sigma_a, TF_sigma_z,
    // aging = :SEI (params.jl:58-117); rho_n above is shared with the thermal block
    TF_M_n, TF_R_SEI, TF_Uref_s, TF_i_0_jside, TF_k_n_aging, TF_w,
    // rxn_MHC (params.jl:16, 67)
    TF_lambda_MHC_n, TF_lambda_MHC_p,
    // NMC_LGM50: scale of D_eff_LGM50 (params.jl:648, 714)
    TF_D_e,
    TF_COUNT
};

// model descriptor (by value into kernels)
struct ModelDesc {
    int Np, Ns, Nn, Nx, Ne;      // nodes per section, Nx = Np+Ns+Nn <= 32, Ne = Np+Nn
    int Na, Nz;                  // current-collector nodes (thermal models), else 0
    int thermal;                 // temperature = true
    int aging;                   // aging = :SEI
    int chem;                    // CHEM_*
    int rxn_mhc;                 // bit 0 / 1: rxn_MHC instead of rxn_BV in the positive / negative electrode
    int ntheta;                  // length of one theta row (reference order, used keys only)
    int theta_stride;            // row stride in doubles
    int mid;                     // meeting node of the twisted block elimination
    // reference layout offsets (external.jl:275-365):
    //   c_e | c_s (particle-major) | [T: a|p|s|n|z] | [film | SOH] | j | Phi_e | Phi_s | [j_s] | I
    int off_cs, off_T, off_film, off_SOH, off_j, off_pe, off_ps, off_js, off_I, N_diff, N_tot;
    double inv_n[4];             // 1/Np, 1/Ns, 1/Nn (Delta x of a section, numerical_tools.jl:216)
    double soh_geo[64];          // aging = :SEI: d trapz(extrapolate_section(y, :n)) / d y_k for a section of unit length
                                 // (residuals.jl:278-297, external.jl:469-522); depends on N_n only
    int8_t slot[TF_COUNT];       // theta field -> position in the row (-1: not a key of this variant)
};

struct Opts {
    double abstol, reltol, abstol_init, reltol_init;
    int maxiters, check_bounds, interp_final;
    int maxord, maxcor, maxnef, maxncf;   // Sundials.jl IDA(): 5, 3, 7, 10
    int skip_alg_deriv;                   // newtons_method!(...; initialize_algebraic_derivatives=false)
};
struct Bounds {
    double V_max, V_min, SOC_max, SOC_min, T_max, c_s_n_max, I_max, I_min, eta_plating_min, c_e_min,
        dfilm_max;
};
struct Summary {   // == plb_summary (include/petlion_b200.h), 80 bytes
    double t_end, V_end, I_end, SOC_end, T_end, aux_end;
    int flag, n_steps, n_res, n_jac, n_netf, n_ncfn, n_newton_init, n_reinit;
};

constexpr int FAIL_NEWTON_INIT = -1, FAIL_CONV = -2, FAIL_ERRTEST = -3, FAIL_MAXITERS = -4,
              FAIL_NONFINITE = -5, FAIL_INIT_BOUNDS = -6, FAIL_PREVIOUS = -7;

struct ResJacArgs {
    ModelDesc m;
    int B;
    const double *Y, *YP, *gamma, *theta, *values;
    int method;
    double value;
    double *res, *nzval;
    int nnz;
    const int* src;       // [nnz] recipe per CSC position (see plb_variant.cuh)
    int use_tma;          // rows staged with bulk copies (needs 16-byte aligned Y, Y', theta arrays)
};

struct AuxArgs {
    ModelDesc m;
    int B;
    const double *theta, *soc, *values;
    int method;
    double value;
    Opts o;
    double *Y, *YP;
    int* status;
    double* gws;
    // linear-solve operator (plb_linear_solve): J(Y, YP, gamma) x = rhs
    const double *gamma, *rhs;
    double* x;
};

struct SimArgs {
    ModelDesc m;
    int B;
    const double* theta;
    const double* values;     // per-system control value or nullptr
    int method;
    double value, tf;
    int input_kind, new_run;   // 0 value, 1 :hold, 2 :rest
    Opts o;
    Bounds b;
    const double* soc0;
    double *sY, *sYP, *sSOC, *st;
    Summary* out;
    int n_save_max;
    double *tr_t, *tr_V, *tr_I, *tr_SOC, *tr_T;
    double* tr_Y;             // optional: every saved row's full state [B][n_save_max][N] (outputs = :all, save_outputs.jl:11-40)
    int* tr_n;
    int* counter;
    double* gws;              // global workspace: [grid * systems per CTA][gws_per_slot]: the parked history vectors, then the factored blocks
    // run_function restricted to a piecewise-linear table of the run's local time (structures.jl:55,
    // scalar_residual.jl:169-170): value(t) = scale[sys] * table(t); `values` then holds the scales.
    // tab_n == 0: run_constant.  tstops: the merged, sorted stop list of postfix_integrator!
    // (model_evaluation.jl:288-310) when a table is given.
    int tab_n, n_tstops;
    const double *tab_t, *tab_v, *tstops;
    // dense output (`tf::AbstractVector`, model_evaluation.jl:80, 148-149): the BDF interpolant of the step that
    // covers each requested GLOBAL time (ascending), rows [B][n_dense]; any output may be null
    int n_dense;
    const double* dense_t;
    double *dn_V, *dn_I, *dn_SOC, *dn_T, *dn_Y;
    int* dn_n;
    int dc_kind;              // METHOD_DC: DC_S_P_MAX .. DC_E_MIN
};

// what the host needs to know about a compiled variant
struct VariantInfo {
    int sim_warps, sim_ctas;        // warps per CTA / CTAs per SM of the persistent integrator
    int nr;                         // N_r (radial nodes per particle) this family is compiled for
    int gws_per_slot;               // doubles of global workspace per system in flight
    int k1_warps, k1_ctas;
    size_t sim_smem, k1_smem;       // dynamic shared memory per CTA
    int vs, nglobal;                // workspace vector stride, history vectors parked in global memory
    int n_slots, n_stage, k1_src_max;
    int lanes;                      // lanes per system: 32, or 64 in the wide families
    int k1_tma;                     // this family has the TMA-staged K1 (k_resjac_tma)
};

// launchers, one set per variant (defined in plb_variant.cuh)
#define PLB_DECLARE_VARIANT(NS)                                                                     \
    namespace NS {                                                                                  \
    VariantInfo info();                                                                             \
    /* structural enumeration of one lane's Jacobian entries: (row, col) in the reference layout */ \
    bool slot_rc(const ModelDesc& m, int method, int slot, int lane, int& row, int& col);           \
    int slot_recipe(const ModelDesc& m, int slot, int lane);                                        \
    cudaError_t launch_resjac(const ResJacArgs& a, int grid, cudaStream_t s);                       \
    cudaError_t launch_initguess(const AuxArgs& a, int grid, cudaStream_t s);                       \
    cudaError_t launch_newton(const AuxArgs& a, int grid, cudaStream_t s);                          \
    cudaError_t launch_linsolve(const AuxArgs& a, int grid, cudaStream_t s);                        \
    cudaError_t launch_simulate(const SimArgs& a, int grid, cudaStream_t s);                        \
    }
PLB_DECLARE_VARIANT(iso)
PLB_DECLARE_VARIANT(th)
PLB_DECLARE_VARIANT(sei)
PLB_DECLARE_VARIANT(wide)
PLB_DECLARE_VARIANT(wsei)
PLB_DECLARE_VARIANT(wth)
PLB_DECLARE_VARIANT(thsei)
PLB_DECLARE_VARIANT(wthsei)
PLB_DECLARE_VARIANT(isomhc)
PLB_DECLARE_VARIANT(thmhc)
PLB_DECLARE_VARIANT(seimhc)
PLB_DECLARE_VARIANT(isolgm)
PLB_DECLARE_VARIANT(thlgm)
PLB_DECLARE_VARIANT(isodc)
PLB_DECLARE_VARIANT(widedc)
PLB_DECLARE_VARIANT(iso12)
PLB_DECLARE_VARIANT(th12)
PLB_DECLARE_VARIANT(sei12)
PLB_DECLARE_VARIANT(iso14)
PLB_DECLARE_VARIANT(th14)
PLB_DECLARE_VARIANT(sei14)
PLB_DECLARE_VARIANT(isosp)
PLB_DECLARE_VARIANT(thsp)
PLB_DECLARE_VARIANT(seisp)
PLB_DECLARE_VARIANT(widemhc)
PLB_DECLARE_VARIANT(wseimhc)
PLB_DECLARE_VARIANT(wthmhc)
PLB_DECLARE_VARIANT(thseimhc)
PLB_DECLARE_VARIANT(wthseimhc)
PLB_DECLARE_VARIANT(widelgm)
PLB_DECLARE_VARIANT(wthlgm)

}  // namespace plb
