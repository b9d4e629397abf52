// plb_kernels.cu -- kernels and the C ABI (include/petlion_b200.h) of libpetlion_b200.so
//
// Kernels (all FP64, sm_100a, warp-per-system):
//   k_resjac     K1: batched residual + CSC Jacobian values  (R_full / J_full callback surface,
//                /root/reference/src/physics_equations/scalar_residual.jl:558-602)
//   k_initguess  initial_guess!                                (states_definition.jl:80-121)
//   k_newton     K3: newtons_method!                           (model_evaluation.jl:430-480)
//   k_simulate   K4: fused persistent integrator               (model_evaluation.jl:312-382 + IDA)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/petlion_b200.h"
#include "plb_integrator.cuh"

using namespace plb;

// =================================================================================================
// K1: residual + Jacobian (CSC nzval) over a batch
// =================================================================================================
// canonical enumeration of one lane's Jacobian entries ("slots"); the host builds, per slot and
// lane, the position in the reference's CSC ordering (or -1).
enum JacSlot {
    JS_CE_L = 0, JS_CE_D, JS_CE_U, JS_CE_J,
    JS_J_CS, JS_J_CE, JS_J_PE, JS_J_PS, JS_J_J,
    JS_PE_L, JS_PE_D, JS_PE_U, JS_PC_L, JS_PC_D, JS_PC_U, JS_PE_J,
    JS_PS_L, JS_PS_D, JS_PS_U, JS_PS_J, JS_PS_I,
    JS_CS_J,
    JS_CS0,                      // 100 particle-block slots r*NR+c
    JS_CTRL_PS0 = JS_CS0 + NR * NR, JS_CTRL_PSN, JS_CTRL_I,
    JS_COUNT
};

struct ResJacArgs {
    ModelDesc m;
    int B;
    const double *Y, *YP, *gamma, *theta, *values;
    int method;
    double value;
    double *res, *nzval;
    int nnz;
    const int* src;       // [nnz] recipe per CSC position: bits 0-15 index into the warp's value table
                          //   (lane-computed entries: slot*32+lane; particle entries: K1_NSTAGE*32 + r*NR+c),
                          //   bit 16 particle-block entry, bit 17 anode, bit 18 diagonal
};

#ifndef PLB_K1_CTAS
#define PLB_K1_CTAS 3
#endif
constexpr int K1_WARPS = 4;
constexpr int K1_NSTAGE = JS_CS0 + 3;   // lane-computed slots: 0..JS_CS_J, then the three control-row slots
constexpr int K1_SRC_MAX = 2304;        // >= nnz of every built variant
__host__ __device__ constexpr int k1_stage_slot(int js) { return js < JS_CS0 ? js : JS_CS0 + (js - JS_CTRL_PS0); }

struct K1Warp {
    double S[K1_NSTAGE][32];   // lane-computed Jacobian entries
    double MCs[NR * NR];       // particle stencil coefficients (contiguous with S: one value table)
    WarpConst C;
};

// K1: one warp evaluates F and the CSC values of dF/dY + gamma dF/dY' of one system at a time.
// HBM traffic per system is exactly the algorithmic 8*(3N + n_theta + nnz) bytes: Y, Y', theta rows
// are read once (lane-mapped, L1-coalesced), res and nzval rows are written once with lane-consecutive
// 8-byte stores.  77 % of nzval is the constant particle stencil scaled by D_s/Rp^2 (minus gamma on
// the diagonal): those entries are produced in the coalesced, branch-free write loop from a recipe
// table held in shared memory, never staged; only the ~500 lane-computed entries go through a
// 6.4 KB shared-memory stage.
template <int CHEM>
__global__ void __launch_bounds__(K1_WARPS * 32, PLB_K1_CTAS) k_resjac(ResJacArgs a) {
    __shared__ K1Warp ws[K1_WARPS];
    __shared__ int src_s[K1_SRC_MAX];
    for (int i = threadIdx.x; i < a.nnz; i += blockDim.x) src_s[i] = a.src[i];
    {
        K1Warp& w0 = ws[threadIdx.x >> 5];
        for (int i = threadIdx.x & 31; i < NR * NR; i += 32) w0.MCs[i] = laws::MC[i / NR][i % NR];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ModelDesc& m = a.m;
    const int N = m.N_tot;
    K1Warp& w = ws[warp];
    const LaneRole ro = make_role(m, lane);
    const int nwarps = gridDim.x * K1_WARPS;
    for (int sys = blockIdx.x * K1_WARPS + warp; sys < a.B; sys += nwarps) {
        const double* __restrict__ gY = a.Y + (size_t)sys * N;
        const double* __restrict__ gYP = a.YP + (size_t)sys * N;
        LaneVec y, yp, res;
        y.ce = ro.act ? gY[ro.x] : 0.0; yp.ce = ro.act ? gYP[ro.x] : 0.0;
        y.pe = ro.act ? gY[m.off_pe + ro.x] : 0.0; yp.pe = 0.0;
        if (ro.elec) {
#pragma unroll
            for (int r = 0; r < NR; r++) { y.cs[r] = gY[m.off_cs + ro.e * NR + r]; yp.cs[r] = gYP[m.off_cs + ro.e * NR + r]; }
            y.j = gY[m.off_j + ro.e]; y.ps = gY[m.off_ps + ro.e];
        } else {
#pragma unroll
            for (int r = 0; r < NR; r++) { y.cs[r] = 0.0; yp.cs[r] = 0.0; }
            y.j = 0.0; y.ps = 0.0;
        }
        yp.j = 0.0; yp.ps = 0.0;
        const double Iapp = gY[m.off_I];
        const double value = a.values ? a.values[sys] : a.value;
        setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
        LaneJac J;
        CtrlRow ctrl;
        if (a.nzval) lane_eval<CHEM, true>(m, w.C, ro, y, yp, Iapp, a.method, value, res, ctrl, J);
        else lane_eval<CHEM, false>(m, w.C, ro, y, yp, Iapp, a.method, value, res, ctrl, J);
        if (a.res) {
            double* __restrict__ gR = a.res + (size_t)sys * N;
            if (ro.act) { gR[ro.x] = res.ce; gR[m.off_pe + ro.x] = res.pe; }
            if (ro.elec) {
#pragma unroll
                for (int r = 0; r < NR; r++) gR[m.off_cs + ro.e * NR + r] = res.cs[r];
                gR[m.off_j + ro.e] = res.j;
                gR[m.off_ps + ro.e] = res.ps;
            }
            if (lane == 0) gR[m.off_I] = ctrl.res;
        }
        if (a.nzval) {
            const double g = a.gamma ? a.gamma[sys] : 0.0;
            w.S[JS_CE_L][lane] = J.ceL; w.S[JS_CE_D][lane] = J.ceD - g; w.S[JS_CE_U][lane] = J.ceU; w.S[JS_CE_J][lane] = J.ce_j;
            w.S[JS_J_CS][lane] = J.j_cs; w.S[JS_J_CE][lane] = J.j_ce; w.S[JS_J_PE][lane] = J.j_pe; w.S[JS_J_PS][lane] = J.j_ps;
            w.S[JS_J_J][lane] = -1.0;
            w.S[JS_PE_L][lane] = J.peL; w.S[JS_PE_D][lane] = J.peD; w.S[JS_PE_U][lane] = J.peU;
            w.S[JS_PC_L][lane] = J.pcL; w.S[JS_PC_D][lane] = J.pcD; w.S[JS_PC_U][lane] = J.pcU; w.S[JS_PE_J][lane] = J.pe_j;
            w.S[JS_PS_L][lane] = J.psL; w.S[JS_PS_D][lane] = J.psD; w.S[JS_PS_U][lane] = J.psU; w.S[JS_PS_J][lane] = J.ps_j;
            w.S[JS_PS_I][lane] = J.ps_I;
            w.S[JS_CS_J][lane] = J.cs_j;
            w.S[k1_stage_slot(JS_CTRL_PS0)][lane] = ctrl.g_ps0;
            w.S[k1_stage_slot(JS_CTRL_PSN)][lane] = ctrl.g_psN;
            w.S[k1_stage_slot(JS_CTRL_I)][lane] = ctrl.g_I;
            __syncwarp();
            const double kap_p = w.C.sec[SC_kap][0], kap_n = w.C.sec[SC_kap][2];
            const double* tab = &w.S[0][0];
            double* __restrict__ gN = a.nzval + (size_t)sys * a.nnz;
#pragma unroll 4
            for (int p = lane; p < a.nnz; p += 32) {
                const int rc = src_s[p];
                const double t = tab[rc & 0xffff];
                const double kap = (rc & (1 << 17)) ? kap_n : kap_p;
                const double gd = (rc & (1 << 18)) ? g : 0.0;
                gN[p] = (rc & (1 << 16)) ? fma(kap, t, -gd) : t;
            }
        }
        __syncwarp();
    }
}

// =================================================================================================
// initial_guess!, newtons_method!, simulate
// =================================================================================================
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
};

#ifndef PLB_SIM_WARPS
#define PLB_SIM_WARPS 6           // warps (systems in flight) per CTA
#endif
#ifndef PLB_SIM_CTAS
#define PLB_SIM_CTAS 1            // CTAs per SM the register/shared-memory budget is sized for
#endif
constexpr int SIM_WARPS = PLB_SIM_WARPS;
constexpr int SIM_CTAS = PLB_SIM_CTAS;

__device__ __forceinline__ WarpWS make_ws(unsigned char* smem_raw, double* gws, int warp) {
    WarpSmem& sm = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
    double* g = gws + ((size_t)blockIdx.x * SIM_WARPS + warp) * (size_t)(NGLOBAL > 0 ? NGLOBAL : 1) * VS;
    return WarpWS{g, &sm.svec[0][0], sm.C, sm.Fa, sm.K};
}

#include "plb_tick.cuh"

template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * 32, SIM_CTAS) k_initguess(AuxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpWS w = make_ws(smem_raw, a.gws, warp);
    const ModelDesc& m = a.m;
    const LaneRole ro = make_role(m, lane);
    for (int sys = blockIdx.x * SIM_WARPS + warp; sys < a.B; sys += gridDim.x * SIM_WARPS) {
        setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
        const double* th = w.C.theta;
        const double SOC = a.soc[sys];
        const double csp = th[TF_c_max_p] * (SOC * (th[TF_theta_max_p] - th[TF_theta_min_p]) + th[TF_theta_min_p]);
        const double csn = th[TF_c_max_n] * (SOC * (th[TF_theta_max_n] - th[TF_theta_min_n]) + th[TF_theta_min_n]);
        double* Y = a.Y + (size_t)sys * m.N_tot;
        const double cs0 = ro.sec == 0 ? csp : csn;
        if (ro.act) { Y[ro.x] = th[TF_c_e0]; Y[m.off_pe + ro.x] = 0.0; }
        if (ro.elec) {
            const double thx = cs0 * w.C.sec[SC_inv_cmax][ro.sec];
            double U, dU, dUdT = 0.0, ddUdT = 0.0;
            if (CHEM == CHEM_LCO) {
                if (ro.sec == 0) laws::OCV_LCO(thx, U, dU, dUdT, ddUdT);
                else laws::OCV_LiC6(thx, sqrt(fmax(thx, 1e-4)), U, dU, dUdT, ddUdT);
                if (w.C.g[GC_dUdT_on] != 0.0) U += dUdT * (w.C.g[GC_T] - kTref);
            } else {
                if (ro.sec == 0) laws::OCV_NMC(thx, U, dU);
                else laws::OCV_LiC6_NMC(thx, U, dU);
            }
            for (int r = 0; r < NR; r++) Y[m.off_cs + ro.e * NR + r] = cs0;
            Y[m.off_j + ro.e] = 0.0;
            Y[m.off_ps + ro.e] = U;
        }
        if (lane == 0) Y[m.off_I] = 0.0;
        __syncwarp();
    }
}

template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * 32, SIM_CTAS) k_newton(AuxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpWS w = make_ws(smem_raw, a.gws, warp);
    const ModelDesc& m = a.m;
    const LaneRole ro = make_role(m, lane);
    const int N = m.N_tot;
    for (int sys = blockIdx.x * SIM_WARPS + warp; sys < a.B; sys += gridDim.x * SIM_WARPS) {
        setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
        for (int i = lane; i < N; i += 32) w.v(V_PHI0)[i] = a.Y[(size_t)sys * N + ref_index(m, i)];
        __syncwarp();
        RunCtl rc;
        rc.method = a.method;
        rc.value = a.values ? a.values[sys] : a.value;
        int nres = 0, njac = 0;
        const int it = newton_init<CHEM>(m, w, ro, rc, a.o, w.v(V_PHI0), w.v(V_PHI1), lane, nres, njac);
        for (int i = lane; i < N; i += 32) {
            a.Y[(size_t)sys * N + ref_index(m, i)] = w.v(V_PHI0)[i];
            a.YP[(size_t)sys * N + ref_index(m, i)] = it > 0 ? w.v(V_PHI1)[i] : 0.0;
        }
        if (lane == 0 && a.status) a.status[sys] = it;
        __syncwarp();
    }
}

template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * 32, SIM_CTAS) k_simulate(SimArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // persistent CTAs; every warp pulls systems from a global queue (step counts vary ~1.5x across
    // a batch) and all warps of the CTA tick in lockstep through the heavy phases (plb_tick.cuh)
    simulate_cta<CHEM>(a, smem_raw);
}

// =================================================================================================
// host side
// =================================================================================================
static thread_local std::string g_err;
static int fail(const std::string& s) { g_err = s; return -1; }
#define CUDA_OK(x)                                                                        \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct KeyDef { const char* utf8; const char* ascii; int field; double lco, nmc; };
// reference keys (UTF-8) <-> canonical fields, with the defaults of src/params.jl:5-117,177-226 (LCO)
// and :295-367, 428-445 (NMC).  nmc = NaN: key not part of the NMC parameter set.
static const double NA = NAN;
static const KeyDef KEYS[] = {
    {"D_n", "D_n", TF_D_n, 7.5e-10, NA}, {"D_p", "D_p", TF_D_p, 7.5e-10, NA}, {"D_s", "D_s", TF_D_s, 7.5e-10, NA},
    {"D_sn", "D_sn", TF_D_sn, 3.9e-14, 1.5e-14}, {"D_sp", "D_sp", TF_D_sp, 1e-14, 2e-14},
    {"Ea_D_sn", "Ea_D_sn", TF_Ea_D_sn, 5000.0, 4e4}, {"Ea_D_sp", "Ea_D_sp", TF_Ea_D_sp, 5000.0, 2.5e4},
    {"Ea_k_n", "Ea_k_n", TF_Ea_k_n, 5000.0, 3e4}, {"Ea_k_p", "Ea_k_p", TF_Ea_k_p, 5000.0, 3e4},
    {"Rp_n", "Rp_n", TF_Rp_n, 2e-6, 10e-6}, {"Rp_p", "Rp_p", TF_Rp_p, 2e-6, 7.5e-6},
    {"T\xe2\x82\x80", "T0", TF_T0, 25 + 273.15, 25 + 273.15},
    {"brugg_n", "brugg_n", TF_brugg_n, 4.0, 1.5}, {"brugg_p", "brugg_p", TF_brugg_p, 4.0, 1.5},
    {"brugg_s", "brugg_s", TF_brugg_s, 4.0, 1.5},
    {"c_e\xe2\x82\x80", "c_e0", TF_c_e0, 1000.0, 1200.0},
    {"c_max_n", "c_max_n", TF_c_max_n, 30555.0, 31080.0}, {"c_max_p", "c_max_p", TF_c_max_p, 51554.0, 51830.0},
    {"k_n", "k_n", TF_k_n, 5.0310e-11, 6.3466e-10}, {"k_p", "k_p", TF_k_p, 2.334e-11, 6.3066e-10},
    {"l_n", "l_n", TF_l_n, 88e-6, 48e-6}, {"l_p", "l_p", TF_l_p, 80e-6, 41.6e-6}, {"l_s", "l_s", TF_l_s, 25e-6, 25e-6},
    {"t\xe2\x82\x8a", "t_plus", TF_t_plus, 0.364, 0.38},
    {"\xce\xb8_max_n", "theta_max_n", TF_theta_max_n, 0.85510, 0.790813},
    {"\xce\xb8_max_p", "theta_max_p", TF_theta_max_p, 0.49550, 0.359749},
    {"\xce\xb8_min_n", "theta_min_n", TF_theta_min_n, 0.01429, 0.001},
    {"\xce\xb8_min_p", "theta_min_p", TF_theta_min_p, 0.99174, 0.955473},
    {"\xcf\x83_n", "sigma_n", TF_sigma_n, 100.0, 100.0}, {"\xcf\x83_p", "sigma_p", TF_sigma_p, 100.0, 100.0},
    {"\xcf\xb5_fn", "eps_fn", TF_eps_fn, 0.0326, 0.038}, {"\xcf\xb5_fp", "eps_fp", TF_eps_fp, 0.025, 0.12},
    {"\xcf\xb5_n", "eps_n", TF_eps_n, 0.485, 0.3}, {"\xcf\xb5_p", "eps_p", TF_eps_p, 0.385, 0.3},
    {"\xcf\xb5_s", "eps_s", TF_eps_s, 0.724, 0.4},
};

struct plb_handle_s {
    plb_model_desc desc;
    ModelDesc m;
    std::vector<int> keys;               // indices into KEYS, reference (sorted) order
    // CSC patterns per method
    std::vector<int> colptr[3], rowval[3];
    int* d_src[3] = {nullptr, nullptr, nullptr};
    int* d_counter = nullptr;
    double* d_gws = nullptr;             // global workspace of the persistent warps
    int sim_grid = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long launches = 0;
    float last_ms = 0.f;
    int num_sms = 0;
};

const char* plb_last_error(void) { return g_err.c_str(); }

// structural enumeration of the Jacobian: (row, col) in the reference layout for slot/lane
static bool slot_rc(const ModelDesc& m, int method, int slot, int lane, int& row, int& col) {
    const int Np = m.Np, Ns = m.Ns, Nx = m.Nx;
    if (lane >= Nx) return false;
    const int x = lane;
    const bool isp = x < Np, isn = x >= Np + Ns, elec = isp || isn;
    const int e = isp ? x : x - Ns;
    const bool first_e = (isp && x == 0) || (isn && x == Np + Ns);
    const bool last_e = (isp && x == Np - 1) || (isn && x == Nx - 1);
    const int r_ce = x, r_pe = m.off_pe + x, r_j = m.off_j + e, r_ps = m.off_ps + e, I = m.off_I;
    auto cs = [&](int r) { return m.off_cs + e * NR + r; };
    const bool last = x == Nx - 1;
    switch (slot) {
        case JS_CE_L: row = r_ce; col = x - 1; return x > 0;
        case JS_CE_D: row = r_ce; col = x; return true;
        case JS_CE_U: row = r_ce; col = x + 1; return x < Nx - 1;
        case JS_CE_J: row = r_ce; col = r_j; return elec;
        case JS_J_CS: row = r_j; col = cs(NR - 1); return elec;
        case JS_J_CE: row = r_j; col = x; return elec;
        case JS_J_PE: row = r_j; col = r_pe; return elec;
        case JS_J_PS: row = r_j; col = r_ps; return elec;
        case JS_J_J: row = r_j; col = r_j; return elec;
        case JS_PE_L: row = r_pe; col = r_pe - 1; return x > 0 && !last;
        case JS_PE_D: row = r_pe; col = r_pe; return true;
        case JS_PE_U: row = r_pe; col = r_pe + 1; return !last;
        case JS_PC_L: row = r_pe; col = x - 1; return x > 0 && !last;
        case JS_PC_D: row = r_pe; col = x; return !last;
        case JS_PC_U: row = r_pe; col = x + 1; return !last;
        case JS_PE_J: row = r_pe; col = r_j; return elec && !last;
        case JS_PS_L: row = r_ps; col = r_ps - 1; return elec && !first_e;
        case JS_PS_D: row = r_ps; col = r_ps; return elec;
        case JS_PS_U: row = r_ps; col = r_ps + 1; return elec && !last_e;
        case JS_PS_J: row = r_ps; col = r_j; return elec;
        case JS_PS_I: row = r_ps; col = I; return (isp && first_e) || (isn && last_e);
        case JS_CS_J: row = cs(NR - 1); col = r_j; return elec;
        case JS_CTRL_PS0: row = I; col = m.off_ps; return lane == 0 && method != PLB_METHOD_I;
        case JS_CTRL_PSN: row = I; col = m.off_ps + m.Ne - 1; return lane == Nx - 1 && method != PLB_METHOD_I;
        case JS_CTRL_I: row = I; col = I; return lane == 0 && method != PLB_METHOD_V;
        default: break;
    }
    if (slot >= JS_CS0 && slot < JS_CS0 + NR * NR) {
        const int r = (slot - JS_CS0) / NR, c = (slot - JS_CS0) % NR;
        row = cs(r); col = cs(c);
        return elec && (laws::mc_mask(r) & (1u << c));
    }
    return false;
}

static int build_patterns(plb_handle_s* h) {
    const ModelDesc& m = h->m;

    for (int method = 0; method < 3; method++) {
        std::vector<std::pair<int, int>> ent;   // (col, row)
        for (int lane = 0; lane < 32; lane++)
            for (int s = 0; s < JS_COUNT; s++) {
                int r, c;
                if (slot_rc(m, method, s, lane, r, c)) ent.push_back({c, r});
            }
        std::sort(ent.begin(), ent.end());
        ent.erase(std::unique(ent.begin(), ent.end()), ent.end());
        std::map<std::pair<int, int>, int> idx;
        h->colptr[method].assign(m.N_tot + 1, 0);
        h->rowval[method].resize(ent.size());
        for (size_t k = 0; k < ent.size(); k++) {
            idx[ent[k]] = (int)k;
            h->rowval[method][k] = ent[k].second;
            h->colptr[method][ent[k].first + 1]++;
        }
        for (int c = 0; c < m.N_tot; c++) h->colptr[method][c + 1] += h->colptr[method][c];
        // per CSC position: where K1 takes the value from
        if ((int)ent.size() > K1_SRC_MAX) return fail("internal: K1_SRC_MAX too small");
        std::vector<int> src(ent.size(), -1);
        for (int lane = 0; lane < 32; lane++)
            for (int s = 0; s < JS_COUNT; s++) {
                int r, c;
                if (!slot_rc(m, method, s, lane, r, c)) continue;
                const int p = idx[{c, r}];
                if (s >= JS_CS0 && s < JS_CS0 + NR * NR) {
                    const int rr = (s - JS_CS0) / NR, cc = (s - JS_CS0) % NR;
                    const int el = lane >= m.Np + m.Ns ? 1 : 0;
                    src[p] = (K1_NSTAGE * 32 + rr * NR + cc) | (1 << 16) | (el << 17) | ((rr == cc ? 1 : 0) << 18);
                } else {
                    src[p] = k1_stage_slot(s) * 32 + lane;
                }
            }
        for (int v : src) if (v < 0) return fail("internal: Jacobian position without a source");
        CUDA_OK(cudaMalloc(&h->d_src[method], src.size() * sizeof(int)));
        CUDA_OK(cudaMemcpy(h->d_src[method], src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return 0;
}

int plb_create(const plb_model_desc* d, plb_handle* out) {
    if (!d || !out) return fail("plb_create: null argument");
    if (d->temperature) return fail("plb_create: temperature=true is not built yet (isothermal variants only)");
    if (d->aging) return fail("plb_create: aging=:SEI is not built yet");
    if (d->N_r_p != NR || d->N_r_n != NR) return fail("plb_create: only N_r_p = N_r_n = 10 is built");
    if (d->N_p < 2 || d->N_s < 2 || d->N_n < 2 || d->N_p + d->N_s + d->N_n > 32)
        return fail("plb_create: need 2 <= N_p,N_s,N_n and N_p+N_s+N_n <= 32 (one lane per node)");
    if (d->cathode != PLB_CATHODE_LCO && d->cathode != PLB_CATHODE_NMC) return fail("plb_create: unknown cathode");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("plb_create: no CUDA device available (this library has no CPU fallback)");
    CUDA_OK(cudaSetDevice(d->device));
    plb_handle_s* h = new plb_handle_s();
    h->desc = *d;
    ModelDesc& m = h->m;
    memset(&m, 0, sizeof m);
    m.Np = d->N_p; m.Ns = d->N_s; m.Nn = d->N_n; m.Nx = m.Np + m.Ns + m.Nn; m.Ne = m.Np + m.Nn;
    m.chem = d->cathode == PLB_CATHODE_LCO ? CHEM_LCO : CHEM_NMC;
    m.off_cs = m.Nx; m.off_j = m.off_cs + NR * m.Ne; m.N_diff = m.off_j;
    m.off_pe = m.off_j + m.Ne; m.off_ps = m.off_pe + m.Nx; m.off_I = m.off_ps + m.Ne; m.N_tot = m.off_I + 1;
    if (m.N_tot > VS) { delete h; return fail("plb_create: system too large for the workspace stride"); }
    for (int f = 0; f < TF_COUNT; f++) m.slot[f] = -1;
    // used keys in the reference's (code-point sorted) order; KEYS[] is already sorted that way
    for (int k = 0; k < (int)(sizeof(KEYS) / sizeof(KEYS[0])); k++) {
        const double dv = d->cathode == PLB_CATHODE_LCO ? KEYS[k].lco : KEYS[k].nmc;
        if (dv != dv) continue;
        m.slot[KEYS[k].field] = (int8_t)h->keys.size();
        h->keys.push_back(k);
    }
    m.ntheta = (int)h->keys.size();
    m.theta_stride = m.ntheta;
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, d->device));
    h->num_sms = prop.multiProcessorCount;
    if (build_patterns(h)) { delete h; return -1; }
    CUDA_OK(cudaMalloc(&h->d_counter, sizeof(int)));
    h->sim_grid = h->num_sms * SIM_CTAS;
    CUDA_OK(cudaMalloc(&h->d_gws, (size_t)h->sim_grid * SIM_WARPS * (NGLOBAL > 0 ? NGLOBAL : 1) * VS * sizeof(double)));
    CUDA_OK(cudaEventCreate(&h->ev0));
    CUDA_OK(cudaEventCreate(&h->ev1));
    *out = h;
    return 0;
}

int plb_destroy(plb_handle h) {
    if (!h) return 0;
    for (int i = 0; i < 3; i++) cudaFree(h->d_src[i]);
    cudaFree(h->d_counter);
    cudaFree(h->d_gws);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return 0;
}
int plb_set_stream(plb_handle h, void* s) { h->stream = (cudaStream_t)s; return 0; }
int plb_nstates(plb_handle h) { return h->m.N_tot; }
int plb_ndiff(plb_handle h) { return h->m.N_diff; }
int plb_ntheta(plb_handle h) { return h->m.ntheta; }
int plb_jac_nnz(plb_handle h, int method) { return (method < 0 || method > 2) ? -1 : (int)h->rowval[method].size(); }
long long plb_launch_count(plb_handle h) { return h->launches; }
float plb_last_kernel_ms(plb_handle h) { return h->last_ms; }

int plb_theta_keys(plb_handle h, const char** keys) {
    for (size_t i = 0; i < h->keys.size(); i++) keys[i] = KEYS[h->keys[i]].utf8;
    return (int)h->keys.size();
}
int plb_theta_index(plb_handle h, const char* key) {
    for (size_t i = 0; i < h->keys.size(); i++)
        if (!strcmp(KEYS[h->keys[i]].utf8, key) || !strcmp(KEYS[h->keys[i]].ascii, key)) return (int)i;
    return -1;
}
int plb_theta_defaults(plb_handle h, double* row) {
    for (size_t i = 0; i < h->keys.size(); i++)
        row[i] = h->desc.cathode == PLB_CATHODE_LCO ? KEYS[h->keys[i]].lco : KEYS[h->keys[i]].nmc;
    return 0;
}
int plb_bounds_defaults(plb_handle h, plb_bounds* b) {
    // src/params.jl:233-253 (LCO), :451-471 (NMC)
    const double nan_ = NAN;
    if (h->desc.cathode == PLB_CATHODE_LCO) { b->V_min = 2.5; b->V_max = 4.3; b->T_max = 55 + 273.15; }
    else { b->V_min = 2.8; b->V_max = 4.2; b->T_max = nan_; }
    b->SOC_min = 0.0; b->SOC_max = 1.0; b->c_s_n_max = nan_; b->I_max = nan_; b->I_min = nan_;
    b->eta_plating_min = nan_; b->c_e_min = nan_; b->dfilm_max = nan_;
    return 0;
}
int plb_opts_defaults(plb_handle, plb_opts* o) {
    // src/params.jl:256-280
    o->abstol = 1e-6; o->reltol = 1e-3; o->abstol_init = 1e-6; o->reltol_init = 1e-3;
    o->maxiters = 10000; o->check_bounds = 1; o->interp_final = 1; o->reserved = 0;
    return 0;
}
int plb_calc_I1C(plb_handle h, int B, const double* theta, double* I1C) {
    // host-side: update_theta! recomputes I1C from the dict (generate_functions.jl:364-372)
    const ModelDesc& m = h->m;
    for (int s = 0; s < B; s++) {
        const double* t = theta + (size_t)s * m.ntheta;
        auto g = [&](int f) { return t[m.slot[f]]; };
        const double eps_sp = 1.0 - (g(TF_eps_fp) + g(TF_eps_p)), eps_sn = 1.0 - (g(TF_eps_fn) + g(TF_eps_n));
        const double qp = eps_sp * g(TF_l_p) * g(TF_c_max_p) * (g(TF_theta_min_p) - g(TF_theta_max_p));
        const double qn = eps_sn * g(TF_l_n) * g(TF_c_max_n) * (g(TF_theta_max_n) - g(TF_theta_min_n));
        I1C[s] = (kF / 3600.0) * std::min(qp, qn);
    }
    return 0;
}
int plb_jac_pattern(plb_handle h, int method, int* colptr, int* rowval, int one_based) {
    if (method < 0 || method > 2) return fail("plb_jac_pattern: bad method");
    const int o = one_based ? 1 : 0;
    for (size_t i = 0; i < h->colptr[method].size(); i++) colptr[i] = h->colptr[method][i] + o;
    for (size_t i = 0; i < h->rowval[method].size(); i++) rowval[i] = h->rowval[method][i] + o;
    return 0;
}

// device staging helper for PLB_MEM_HOST calls
struct DevBuf {
    void* p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { n = bytes; return bytes ? (cudaMalloc(&p, bytes) == cudaSuccess ? 0 : -1) : 0; }
};
template <class T>
static int stage_in(DevBuf& b, const T*& ptr, size_t count, int mem, cudaStream_t s) {
    if (!ptr || mem == PLB_MEM_DEVICE) return 0;
    if (b.alloc(count * sizeof(T))) return fail("cudaMalloc failed");
    if (cudaMemcpyAsync(b.p, ptr, count * sizeof(T), cudaMemcpyHostToDevice, s) != cudaSuccess) return fail("H2D failed");
    ptr = (const T*)b.p;
    return 0;
}
template <class T>
static int stage_inout(DevBuf& b, T*& ptr, T*& host, size_t count, int mem, bool copy_in, cudaStream_t s) {
    host = nullptr;
    if (!ptr || mem == PLB_MEM_DEVICE) return 0;
    if (b.alloc(count * sizeof(T))) return fail("cudaMalloc failed");
    if (copy_in && cudaMemcpyAsync(b.p, ptr, count * sizeof(T), cudaMemcpyHostToDevice, s) != cudaSuccess) return fail("H2D failed");
    host = ptr;
    ptr = (T*)b.p;
    return 0;
}
template <class T>
static int stage_out(T* dev, T* host, size_t count, cudaStream_t s) {
    if (!host) return 0;
    if (cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, s) != cudaSuccess) return fail("D2H failed");
    return 0;
}

static Opts to_opts(const plb_opts* o) {
    Opts r;
    r.abstol = o->abstol; r.reltol = o->reltol; r.abstol_init = o->abstol_init; r.reltol_init = o->reltol_init;
    r.maxiters = o->maxiters; r.check_bounds = o->check_bounds; r.interp_final = o->interp_final;
    // Sundials.jl IDA() constructor values (third-party): max_order 5, max_nonlinear_iters 3,
    // max_error_test_failures 7, max_convergence_failures 10
    r.maxord = 5; r.maxcor = 3; r.maxnef = 7; r.maxncf = 10;
    return r;
}

int plb_initial_guess(plb_handle h, int B, const double* soc, const double* theta, double* Y0, int mem) {
    if (B <= 0) return 0;
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    DevBuf b1, b2, b3;
    double* hostY;
    if (stage_in(b1, soc, (size_t)B, mem, s) || stage_in(b2, theta, (size_t)B * m.ntheta, mem, s) ||
        stage_inout(b3, Y0, hostY, (size_t)B * m.N_tot, mem, false, s)) return -1;
    AuxArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.B = B; a.theta = theta; a.soc = soc; a.Y = Y0; a.gws = h->d_gws;
    const size_t smem = sizeof(WarpSmem) * SIM_WARPS;
    const int grid = std::min((B + SIM_WARPS - 1) / SIM_WARPS, h->sim_grid);
    if (m.chem == CHEM_LCO) {
        CUDA_OK(cudaFuncSetAttribute(k_initguess<CHEM_LCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_initguess<CHEM_LCO><<<grid, SIM_WARPS * 32, smem, s>>>(a);
    } else {
        CUDA_OK(cudaFuncSetAttribute(k_initguess<CHEM_NMC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_initguess<CHEM_NMC><<<grid, SIM_WARPS * 32, smem, s>>>(a);
    }
    h->launches++;
    CUDA_OK(cudaGetLastError());
    if (stage_out(Y0, hostY, (size_t)B * m.N_tot, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int plb_resjac(plb_handle h, int B, const double* Y, const double* YP, const double* gamma,
               const double* theta, const plb_run* run, const double* values, double* res,
               double* nzval, int mem) {
    if (B <= 0) return 0;
    if (!run || run->method < 0 || run->method > 2) return fail("plb_resjac: bad run");
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    const int nnz = (int)h->rowval[run->method].size();
    DevBuf b1, b2, b3, b4, b5, b6, b7;
    double *hostR, *hostN;
    if (stage_in(b1, Y, (size_t)B * m.N_tot, mem, s) || stage_in(b2, YP, (size_t)B * m.N_tot, mem, s) ||
        stage_in(b3, gamma, (size_t)B, mem, s) || stage_in(b4, theta, (size_t)B * m.ntheta, mem, s) ||
        stage_in(b5, values, (size_t)B, mem, s) ||
        stage_inout(b6, res, hostR, (size_t)B * m.N_tot, mem, false, s) ||
        stage_inout(b7, nzval, hostN, (size_t)B * nnz, mem, false, s)) return -1;
    ResJacArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.B = B; a.Y = Y; a.YP = YP; a.gamma = gamma; a.theta = theta; a.values = values;
    a.method = run->method; a.value = run->value; a.res = res; a.nzval = nzval; a.nnz = nnz;
    a.src = h->d_src[run->method];
    const size_t smem = 0;
    const int grid = std::min((B + K1_WARPS - 1) / K1_WARPS, h->num_sms * PLB_K1_CTAS * 2);
    CUDA_OK(cudaEventRecord(h->ev0, s));
    if (m.chem == CHEM_LCO) {
        k_resjac<CHEM_LCO><<<grid, K1_WARPS * 32, smem, s>>>(a);
    } else {
        k_resjac<CHEM_NMC><<<grid, K1_WARPS * 32, smem, s>>>(a);
    }
    CUDA_OK(cudaEventRecord(h->ev1, s));
    h->launches++;
    CUDA_OK(cudaGetLastError());
    if (stage_out(res, hostR, (size_t)B * m.N_tot, s) || stage_out(nzval, hostN, (size_t)B * nnz, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
    return 0;
}

int plb_newton_init(plb_handle h, int B, double* Y, double* YP, const double* theta, const plb_run* run,
                    const double* values, const plb_opts* opts, int* status, int mem) {
    if (B <= 0) return 0;
    if (!run || !opts) return fail("plb_newton_init: null run/opts");
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    DevBuf b1, b2, b3, b4, b5;
    double *hostY, *hostYP;
    int* hostS;
    if (stage_in(b1, theta, (size_t)B * m.ntheta, mem, s) || stage_in(b2, values, (size_t)B, mem, s) ||
        stage_inout(b3, Y, hostY, (size_t)B * m.N_tot, mem, true, s) ||
        stage_inout(b4, YP, hostYP, (size_t)B * m.N_tot, mem, false, s) ||
        stage_inout(b5, status, hostS, (size_t)B, mem, false, s)) return -1;
    AuxArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.B = B; a.theta = theta; a.values = values; a.method = run->method; a.value = run->value;
    a.o = to_opts(opts); a.Y = Y; a.YP = YP; a.status = status; a.gws = h->d_gws;
    const size_t smem = sizeof(WarpSmem) * SIM_WARPS;
    const int grid = std::min((B + SIM_WARPS - 1) / SIM_WARPS, h->sim_grid);
    if (m.chem == CHEM_LCO) {
        CUDA_OK(cudaFuncSetAttribute(k_newton<CHEM_LCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_newton<CHEM_LCO><<<grid, SIM_WARPS * 32, smem, s>>>(a);
    } else {
        CUDA_OK(cudaFuncSetAttribute(k_newton<CHEM_NMC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_newton<CHEM_NMC><<<grid, SIM_WARPS * 32, smem, s>>>(a);
    }
    h->launches++;
    CUDA_OK(cudaGetLastError());
    if (stage_out(Y, hostY, (size_t)B * m.N_tot, s) || stage_out(YP, hostYP, (size_t)B * m.N_tot, s) ||
        stage_out(status, hostS, (size_t)B, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int plb_simulate(plb_handle h, int B, const double* theta, const plb_run* run, const double* values,
                 const plb_opts* opts, const plb_bounds* bounds, const double* soc0, double* sY,
                 double* sYP, double* sSOC, double* st, plb_summary* summary, int n_save_max,
                 double* tr_t, double* tr_V, double* tr_I, double* tr_SOC, int* tr_n, int mem) {
    if (B <= 0) return 0;
    if (!run || !opts || !bounds || !theta || !sY || !sSOC || !st || !summary)
        return fail("plb_simulate: null required argument");
    if (run->method < 0 || run->method > 2) return fail("plb_simulate: bad method");
    if (run->input_kind != PLB_INPUT_VALUE && run->new_run && run->input_kind == PLB_INPUT_HOLD)
        return fail("plb_simulate: Cannot use `:hold` without a previous simulation.");   // checks.jl:385
    if (run->input_kind == PLB_INPUT_REST && run->method == PLB_METHOD_V) return fail("plb_simulate: Unsupported input symbol.");
    static_assert(sizeof(plb_summary) == sizeof(Summary), "summary layout");
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    const size_t BN = (size_t)B * m.N_tot, BS = (size_t)B * (n_save_max > 0 ? n_save_max : 0);
    DevBuf b[16];
    double *hY, *hYP, *hSOC, *ht, *htt, *htV, *htI, *htS;
    int* htn;
    plb_summary* hsum;
    const bool cont = !run->new_run;
    if (stage_in(b[0], theta, (size_t)B * m.ntheta, mem, s) || stage_in(b[1], values, (size_t)B, mem, s) ||
        stage_in(b[2], soc0, (size_t)B, mem, s) || stage_inout(b[3], sY, hY, BN, mem, cont, s) ||
        stage_inout(b[4], sYP, hYP, BN, mem, false, s) || stage_inout(b[5], sSOC, hSOC, (size_t)B, mem, cont, s) ||
        stage_inout(b[6], st, ht, (size_t)B, mem, cont, s) ||
        stage_inout(b[7], summary, hsum, (size_t)B, mem, false, s) ||
        stage_inout(b[8], tr_t, htt, BS, mem, false, s) || stage_inout(b[9], tr_V, htV, BS, mem, false, s) ||
        stage_inout(b[10], tr_I, htI, BS, mem, false, s) || stage_inout(b[11], tr_SOC, htS, BS, mem, false, s) ||
        stage_inout(b[12], tr_n, htn, (size_t)B, mem, false, s)) return -1;
    SimArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.B = B; a.theta = theta; a.values = values; a.method = run->method; a.value = run->value;
    a.tf = run->tf; a.input_kind = run->input_kind; a.new_run = run->new_run; a.o = to_opts(opts);
    memcpy(&a.b, bounds, sizeof(Bounds));
    static_assert(sizeof(plb_bounds) == sizeof(Bounds), "bounds layout");
    a.soc0 = soc0; a.sY = sY; a.sYP = sYP; a.sSOC = sSOC; a.st = st; a.out = (Summary*)summary;
    a.n_save_max = n_save_max > 0 ? n_save_max : 0;
    a.tr_t = tr_t; a.tr_V = tr_V; a.tr_I = tr_I; a.tr_SOC = tr_SOC; a.tr_n = tr_n;
    a.counter = h->d_counter;
    a.gws = h->d_gws;
    CUDA_OK(cudaMemsetAsync(h->d_counter, 0, sizeof(int), s));
    const size_t smem = sizeof(WarpSmem) * SIM_WARPS;
    const int grid = std::min((B + SIM_WARPS - 1) / SIM_WARPS, h->sim_grid);
    CUDA_OK(cudaEventRecord(h->ev0, s));
    if (m.chem == CHEM_LCO) {
        CUDA_OK(cudaFuncSetAttribute(k_simulate<CHEM_LCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_simulate<CHEM_LCO><<<grid, SIM_WARPS * 32, smem, s>>>(a);
    } else {
        CUDA_OK(cudaFuncSetAttribute(k_simulate<CHEM_NMC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_simulate<CHEM_NMC><<<grid, SIM_WARPS * 32, smem, s>>>(a);
    }
    CUDA_OK(cudaEventRecord(h->ev1, s));
    h->launches++;
    CUDA_OK(cudaGetLastError());
    if (stage_out(sY, hY, BN, s) || stage_out(sYP, hYP, BN, s) || stage_out(sSOC, hSOC, (size_t)B, s) ||
        stage_out(st, ht, (size_t)B, s) || stage_out(summary, hsum, (size_t)B, s) || stage_out(tr_t, htt, BS, s) ||
        stage_out(tr_V, htV, BS, s) || stage_out(tr_I, htI, BS, s) || stage_out(tr_SOC, htS, BS, s) ||
        stage_out(tr_n, htn, (size_t)B, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
    return 0;
}
