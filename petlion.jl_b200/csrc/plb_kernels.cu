// plb_kernels.cu -- host side of libpetlion_b200.so: the C ABI declared in include/petlion_b200.h.
//
// The device code lives in plb_variant.cuh and is compiled once per model family (plb_variant_iso.cu,
// plb_variant_th.cu); this file owns the handle, the parameter-key tables, the CSC pattern of the
// Jacobian and the staging of host buffers.
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/petlion_b200.h"
#include "plb_common.cuh"

using namespace plb;

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

struct KeyDef { const char* utf8; const char* ascii; int field; double lco, nmc, lgm; int only; };   // only: bit 0 thermal, bit 1 aging, bit 2 / 3 rxn_p / rxn_n = rxn_MHC (0: always)
// reference keys (UTF-8) <-> canonical fields, with the defaults of src/params.jl:5-117,177-226 (LCO)
// and :295-367, 428-445 (NMC), :514-745 (lgm: NMC_LGM50 / LiC6_LGM50).  NaN: key not part of that parameter set.  The table is in the
// reference's key order (Symbols sorted by code point); `only` marks the keys that the generated functions
// use only when temperature = true (bit 0) and/or aging = :SEI (bit 1).
static const double NA = NAN;
static const KeyDef KEYS[] = {
    {"Cp_a", "Cp_a", TF_Cp_a, 897.0, NA, 897.0, 1}, {"Cp_n", "Cp_n", TF_Cp_n, 700.0, NA, 700.0, 1}, {"Cp_p", "Cp_p", TF_Cp_p, 700.0, NA, 700.0, 1},
    {"Cp_s", "Cp_s", TF_Cp_s, 700.0, NA, 700.0, 1}, {"Cp_z", "Cp_z", TF_Cp_z, 385.0, NA, 385.0, 1},
    {"D_e", "D_e", TF_D_e, NA, NA, 8.794e-11, 0},
    {"D_n", "D_n", TF_D_n, 7.5e-10, NA, NA, 0}, {"D_p", "D_p", TF_D_p, 7.5e-10, NA, NA, 0}, {"D_s", "D_s", TF_D_s, 7.5e-10, NA, NA, 0},
    {"D_sn", "D_sn", TF_D_sn, 3.9e-14, 1.5e-14, 3.3e-14, 0}, {"D_sp", "D_sp", TF_D_sp, 1e-14, 2e-14, 4e-15, 0},
    {"Ea_D_sn", "Ea_D_sn", TF_Ea_D_sn, 5000.0, 4e4, 3.03e4, 0}, {"Ea_D_sp", "Ea_D_sp", TF_Ea_D_sp, 5000.0, 2.5e4, 0.0, 0},
    {"Ea_k_n", "Ea_k_n", TF_Ea_k_n, 5000.0, 3e4, 35000.0, 0}, {"Ea_k_p", "Ea_k_p", TF_Ea_k_p, 5000.0, 3e4, 17800.0, 0},
    {"M_n", "M_n", TF_M_n, 7.3e-4, NA, NA, 2}, {"R_SEI", "R_SEI", TF_R_SEI, 0.01, NA, NA, 2},
    {"Rp_n", "Rp_n", TF_Rp_n, 2e-6, 10e-6, 5.86e-6, 0}, {"Rp_p", "Rp_p", TF_Rp_p, 2e-6, 7.5e-6, 5.22e-06, 0},
    {"T_amb", "T_amb", TF_T_amb, 25 + 273.15, NA, 25 + 273.15, 1},
    {"T\xe2\x82\x80", "T0", TF_T0, 25 + 273.15, 25 + 273.15, 25 + 273.15, 0},
    {"Uref_s", "Uref_s", TF_Uref_s, 0.4, NA, NA, 2},
    {"brugg_n", "brugg_n", TF_brugg_n, 4.0, 1.5, 1.5, 0}, {"brugg_p", "brugg_p", TF_brugg_p, 4.0, 1.5, 1.5, 0},
    {"brugg_s", "brugg_s", TF_brugg_s, 4.0, 1.5, 1.5, 0},
    {"c_e\xe2\x82\x80", "c_e0", TF_c_e0, 1000.0, 1200.0, 1000.0, 0},
    {"c_max_n", "c_max_n", TF_c_max_n, 30555.0, 31080.0, 33133.0, 0}, {"c_max_p", "c_max_p", TF_c_max_p, 51554.0, 51830.0, 63104.0, 0},
    {"h_cell", "h_cell", TF_h_cell, 1.0, NA, 1.0, 1},
    {"i_0_jside", "i_0_jside", TF_i_0_jside, 1.5e-6, NA, NA, 2},
    {"k_n", "k_n", TF_k_n, 5.0310e-11, 6.3466e-10, 6.716046737258585e-12, 0}, {"k_n_aging", "k_n_aging", TF_k_n_aging, 1.0, NA, NA, 2},
    {"k_p", "k_p", TF_k_p, 2.334e-11, 6.3066e-10, 3.5445802224420315e-11, 0},
    {"l_a", "l_a", TF_l_a, 10e-6, NA, 16e-6, 1},
    {"l_n", "l_n", TF_l_n, 88e-6, 48e-6, 85.2e-6, 0}, {"l_p", "l_p", TF_l_p, 80e-6, 41.6e-6, 75.6e-6, 0}, {"l_s", "l_s", TF_l_s, 25e-6, 25e-6, 12e-6, 0},
    {"l_z", "l_z", TF_l_z, 10e-6, NA, 12e-6, 1},
    {"t\xe2\x82\x8a", "t_plus", TF_t_plus, 0.364, 0.38, 0.2594, 0},
    {"w", "w", TF_w, 2.0, NA, NA, 2},
    {"\xce\xb8_max_n", "theta_max_n", TF_theta_max_n, 0.85510, 0.790813, 29866.0 / 33133, 0},
    {"\xce\xb8_max_p", "theta_max_p", TF_theta_max_p, 0.49550, 0.359749, 17038.0 / 63104.0, 0},
    {"\xce\xb8_min_n", "theta_min_n", TF_theta_min_n, 0.01429, 0.001, 0.0481727, 0},
    {"\xce\xb8_min_p", "theta_min_p", TF_theta_min_p, 0.99174, 0.955473, 0.8395, 0},
    {"\xce\xbb_MHC_n", "lambda_MHC_n", TF_lambda_MHC_n, 6.26e-20, NA, 0.0, 8}, {"\xce\xbb_MHC_p", "lambda_MHC_p", TF_lambda_MHC_p, 6.26e-20, NA, 0.0, 4},
    {"\xce\xbb_a", "lambda_a", TF_lambda_a, 237.0, NA, 237.0, 1}, {"\xce\xbb_n", "lambda_n", TF_lambda_n, 1.7, NA, 1.7, 1},
    {"\xce\xbb_p", "lambda_p", TF_lambda_p, 2.1, NA, 2.1, 1}, {"\xce\xbb_s", "lambda_s", TF_lambda_s, 0.16, NA, 0.16, 1},
    {"\xce\xbb_z", "lambda_z", TF_lambda_z, 401.0, NA, 401.0, 1},
    {"\xcf\x81_a", "rho_a", TF_rho_a, 2700.0, NA, 2700.0, 1}, {"\xcf\x81_n", "rho_n", TF_rho_n, 2500.0, NA, 1657.0, 3},
    {"\xcf\x81_p", "rho_p", TF_rho_p, 2500.0, NA, 3262.0, 1}, {"\xcf\x81_s", "rho_s", TF_rho_s, 1100.0, NA, 397.0, 1},
    {"\xcf\x81_z", "rho_z", TF_rho_z, 8940.0, NA, 8960.0, 1},
    {"\xcf\x83_a", "sigma_a", TF_sigma_a, 3.55e7, NA, 36.914e6, 1},
    {"\xcf\x83_n", "sigma_n", TF_sigma_n, 100.0, 100.0, 215.0, 0}, {"\xcf\x83_p", "sigma_p", TF_sigma_p, 100.0, 100.0, 0.18, 0},
    {"\xcf\x83_z", "sigma_z", TF_sigma_z, 5.96e7, NA, 58.41e6, 1},
    {"\xcf\xb5_fn", "eps_fn", TF_eps_fn, 0.0326, 0.038, 0.0, 0}, {"\xcf\xb5_fp", "eps_fp", TF_eps_fp, 0.025, 0.12, 0.0, 0},
    {"\xcf\xb5_n", "eps_n", TF_eps_n, 0.485, 0.3, 0.25, 0}, {"\xcf\xb5_p", "eps_p", TF_eps_p, 0.385, 0.3, 0.335, 0},
    {"\xcf\xb5_s", "eps_s", TF_eps_s, 0.724, 0.4, 0.47, 0},
};

// one compiled model family
struct Variant {
    VariantInfo (*info)();
    bool (*slot_rc)(const ModelDesc&, int, int, int, int&, int&);
    int (*slot_recipe)(const ModelDesc&, int, int);
    cudaError_t (*resjac)(const ResJacArgs&, int, cudaStream_t);
    cudaError_t (*initguess)(const AuxArgs&, int, cudaStream_t);
    cudaError_t (*newton)(const AuxArgs&, int, cudaStream_t);
    cudaError_t (*linsolve)(const AuxArgs&, int, cudaStream_t);
    cudaError_t (*simulate)(const SimArgs&, int, cudaStream_t);
};
static const Variant V_ISO = {iso::info, iso::slot_rc, iso::slot_recipe, iso::launch_resjac, iso::launch_initguess,
                              iso::launch_newton, iso::launch_linsolve, iso::launch_simulate};
static const Variant V_TH = {th::info, th::slot_rc, th::slot_recipe, th::launch_resjac, th::launch_initguess,
                             th::launch_newton, th::launch_linsolve, th::launch_simulate};
static const Variant V_SEI = {sei::info, sei::slot_rc, sei::slot_recipe, sei::launch_resjac, sei::launch_initguess,
                              sei::launch_newton, sei::launch_linsolve, sei::launch_simulate};
// grids with 33..64 x-nodes: two warps per system
static const Variant V_WIDE = {wide::info, wide::slot_rc, wide::slot_recipe, wide::launch_resjac, wide::launch_initguess,
                               wide::launch_newton, wide::launch_linsolve, wide::launch_simulate};
static const Variant V_WSEI = {wsei::info, wsei::slot_rc, wsei::slot_recipe, wsei::launch_resjac, wsei::launch_initguess,
                               wsei::launch_newton, wsei::launch_linsolve, wsei::launch_simulate};

static const Variant V_WTH = {wth::info, wth::slot_rc, wth::slot_recipe, wth::launch_resjac, wth::launch_initguess,
                              wth::launch_newton, wth::launch_linsolve, wth::launch_simulate};

static const Variant V_THSEI = {thsei::info, thsei::slot_rc, thsei::slot_recipe, thsei::launch_resjac, thsei::launch_initguess,
                                thsei::launch_newton, thsei::launch_linsolve, thsei::launch_simulate};

static const Variant V_WTHSEI = {wthsei::info, wthsei::slot_rc, wthsei::slot_recipe, wthsei::launch_resjac, wthsei::launch_initguess,
                                 wthsei::launch_newton, wthsei::launch_linsolve, wthsei::launch_simulate};

// the 32-node families with rxn_MHC compiled in (used by models that select it)
#define PLB_VARIANT_TABLE(NS) {NS::info, NS::slot_rc, NS::slot_recipe, NS::launch_resjac, NS::launch_initguess, NS::launch_newton, NS::launch_linsolve, NS::launch_simulate}
static const Variant V_ISO12 = PLB_VARIANT_TABLE(iso12);        // N_r = 12 / 14 sibling builds
static const Variant V_TH12 = PLB_VARIANT_TABLE(th12);
static const Variant V_SEI12 = PLB_VARIANT_TABLE(sei12);
static const Variant V_ISO14 = PLB_VARIANT_TABLE(iso14);
static const Variant V_TH14 = PLB_VARIANT_TABLE(th14);
static const Variant V_SEI14 = PLB_VARIANT_TABLE(sei14);
static const Variant V_ISOSP = PLB_VARIANT_TABLE(isosp);        // Fickian_method = :spectral sibling builds
static const Variant V_THSP = PLB_VARIANT_TABLE(thsp);
static const Variant V_SEISP = PLB_VARIANT_TABLE(seisp);
static const Variant V_ISOLGM = PLB_VARIANT_TABLE(isolgm);      // NMC_LGM50 chemistry (its own instantiation of the iso / th families)
static const Variant V_THLGM = PLB_VARIANT_TABLE(thlgm);
static const Variant V_ISOMHC = PLB_VARIANT_TABLE(isomhc);
static const Variant V_THMHC = PLB_VARIANT_TABLE(thmhc);
static const Variant V_SEIMHC = PLB_VARIANT_TABLE(seimhc);
// the isothermal families with the concentration-rate inputs compiled in (used by those runs only)
static const Variant V_ISODC = {isodc::info, isodc::slot_rc, isodc::slot_recipe, isodc::launch_resjac, isodc::launch_initguess,
                                isodc::launch_newton, isodc::launch_linsolve, isodc::launch_simulate};
static const Variant V_WIDEDC = {widedc::info, widedc::slot_rc, widedc::slot_recipe, widedc::launch_resjac, widedc::launch_initguess,
                                 widedc::launch_newton, widedc::launch_linsolve, widedc::launch_simulate};

// the rest of the option matrix: rxn_MHC on the two-warp grids and with temperature + aging, NMC_LGM50 on the two-warp grids
static const Variant V_WIDEMHC = PLB_VARIANT_TABLE(widemhc);
static const Variant V_WSEIMHC = PLB_VARIANT_TABLE(wseimhc);
static const Variant V_WTHMHC = PLB_VARIANT_TABLE(wthmhc);
static const Variant V_THSEIMHC = PLB_VARIANT_TABLE(thseimhc);
static const Variant V_WTHSEIMHC = PLB_VARIANT_TABLE(wthseimhc);
static const Variant V_WIDELGM = PLB_VARIANT_TABLE(widelgm);
static const Variant V_WTHLGM = PLB_VARIANT_TABLE(wthlgm);

// which compiled family runs an option set: (temperature, aging, two warps per system, rxn_MHC compiled in, NMC_LGM50 chemistry,
// N_r, spectral particle scheme).  Like the reference, which generates one set of functions per option string
// (strings_directory_func, external.jl:417-456), every row is its own instantiation of the same templates.
struct VEntry { int th, sei, wide, mhc, lgm, nr, sp; const Variant* v; };
static const VEntry VTAB[] = {
    {0, 0, 0, 0, 0, 10, 0, &V_ISO},    {1, 0, 0, 0, 0, 10, 0, &V_TH},    {0, 1, 0, 0, 0, 10, 0, &V_SEI},    {1, 1, 0, 0, 0, 10, 0, &V_THSEI},
    {0, 0, 1, 0, 0, 10, 0, &V_WIDE},   {1, 0, 1, 0, 0, 10, 0, &V_WTH},   {0, 1, 1, 0, 0, 10, 0, &V_WSEI},   {1, 1, 1, 0, 0, 10, 0, &V_WTHSEI},
    {0, 0, 0, 1, 0, 10, 0, &V_ISOMHC}, {1, 0, 0, 1, 0, 10, 0, &V_THMHC}, {0, 1, 0, 1, 0, 10, 0, &V_SEIMHC}, {1, 1, 0, 1, 0, 10, 0, &V_THSEIMHC},
    {0, 0, 1, 1, 0, 10, 0, &V_WIDEMHC}, {1, 0, 1, 1, 0, 10, 0, &V_WTHMHC}, {0, 1, 1, 1, 0, 10, 0, &V_WSEIMHC}, {1, 1, 1, 1, 0, 10, 0, &V_WTHSEIMHC},
    {0, 0, 0, 0, 1, 10, 0, &V_ISOLGM}, {1, 0, 0, 0, 1, 10, 0, &V_THLGM}, {0, 0, 1, 0, 1, 10, 0, &V_WIDELGM}, {1, 0, 1, 0, 1, 10, 0, &V_WTHLGM},
    {0, 0, 0, 0, 0, 12, 0, &V_ISO12},  {1, 0, 0, 0, 0, 12, 0, &V_TH12},  {0, 1, 0, 0, 0, 12, 0, &V_SEI12},
    {0, 0, 0, 0, 0, 14, 0, &V_ISO14},  {1, 0, 0, 0, 0, 14, 0, &V_TH14},  {0, 1, 0, 0, 0, 14, 0, &V_SEI14},
    {0, 0, 0, 0, 0, 10, 1, &V_ISOSP},  {1, 0, 0, 0, 0, 10, 1, &V_THSP},  {0, 1, 0, 0, 0, 10, 1, &V_SEISP},
};
static const Variant* find_variant(int th, int sei, int wide, int mhc, int lgm, int nr, int sp) {
    for (const VEntry& e : VTAB)
        if (e.th == th && e.sei == sei && e.wide == wide && e.mhc == mhc && e.lgm == lgm && e.nr == nr && e.sp == sp) return e.v;
    return nullptr;
}

struct plb_handle_s {
    plb_model_desc desc;
    ModelDesc m;
    const Variant* v = nullptr;
    const Variant* v_dc = nullptr;       // the same family with METHOD_DC compiled in (isothermal, no aging), else null
    VariantInfo vi;
    std::vector<int> keys;               // indices into KEYS, reference (sorted) order
    // CSC patterns per method
    std::vector<int> colptr[N_METHODS], rowval[N_METHODS];
    int* d_src[N_METHODS] = {};
    bool has_dT = false;                 // methods: I, V, P, eta_p everywhere; dT only in thermal models
    bool method_ok(int method) const { return method >= 0 && method < N_METHODS && (method != METHOD_DT || has_dT); }
    int* d_counter = nullptr;
    double* d_gws = nullptr;             // global workspace of the persistent warps
    int sim_grid = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long launches = 0;
    std::vector<double> opt_tstops;       // p.opts.tstops (params.jl:272): applied to every simulate call
    float last_ms = 0.f;
    int num_sms = 0;
    bool k1_no_tma = getenv("PLB_K1_NO_TMA") != nullptr;    // A/B knob (profiles/k1_probe.py)
    // device staging buffers of PLB_MEM_HOST calls: grow-only, reused from call to call (a cudaMalloc /
    // cudaFree pair per argument and call costs more than the transfers themselves)
    static constexpr int NPOOL = 28;
    void* pool_ptr[NPOOL] = {};
    size_t pool_cap[NPOOL] = {};
    // dense-output request of the next simulate call (plb_set_dense_output): one-shot
    struct {
        int n = 0, mem = 0;
        std::vector<double> t;
        double *V = nullptr, *I = nullptr, *SOC = nullptr, *T = nullptr, *Y = nullptr;
        int* n_done = nullptr;
    } dense;
};

// Every entry point that touches CUDA runs on the handle's device, whatever the caller's current device is
// (a process may hold handles on several GPUs, and a host such as torch moves the current device around).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess; else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define PLB_ENTER(h, name)                                                       \
    if (!(h)) return fail(name ": null handle");                                  \
    DeviceGuard guard_((h)->desc.device);                                         \
    if (!guard_.ok) return fail(name ": cudaSetDevice failed")

const char* plb_last_error(void) { return g_err.c_str(); }

static int build_patterns(plb_handle_s* h) {
    const ModelDesc& m = h->m;
    const int n_slots = h->vi.n_slots;
    for (int method = 0; method < N_METHODS; method++) {
        if (!h->method_ok(method)) continue;
        std::vector<std::pair<int, int>> ent;   // (col, row)
        for (int lane = 0; lane < h->vi.lanes; lane++)
            for (int s = 0; s < n_slots; s++) {
                int r, c;
                if (h->v->slot_rc(m, method, s, lane, r, c)) ent.push_back({c, r});
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
        if ((int)ent.size() > h->vi.k1_src_max) return fail("internal: K1_SRC_MAX too small");
        std::vector<int> src(ent.size(), -1);
        for (int lane = 0; lane < h->vi.lanes; lane++)
            for (int s = 0; s < n_slots; s++) {
                int r, c;
                if (!h->v->slot_rc(m, method, s, lane, r, c)) continue;
                src[idx[{c, r}]] = h->v->slot_recipe(m, s, lane);
            }
        for (int v : src) if (v < 0) return fail("internal: Jacobian position without a source");
        CUDA_OK(cudaMalloc(&h->d_src[method], src.size() * sizeof(int)));
        CUDA_OK(cudaMemcpy(h->d_src[method], src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return 0;
}

int plb_create(const plb_model_desc* d, plb_handle* out) {
    if (!d || !out) return fail("plb_create: null argument");
    if (d->aging && d->cathode != PLB_CATHODE_LCO) return fail("plb_create: aging=:SEI needs the LCO parameter set");
    // radial nodes per particle (params.jl:134-136): compile-time in every family (the stencil's eigen-basis, a lane's
    // registers, the recipe tables); 10 everywhere, 12 and 14 as sibling builds of the 32-node iso / th / sei families
    const int NR_HOST = d->N_r_p;
    if (d->N_r_p != d->N_r_n || (NR_HOST != 10 && NR_HOST != 12 && NR_HOST != 14))
        return fail("plb_create: built for N_r_p = N_r_n = 10 (every family), 12 or 14 (isothermal, thermal and SEI families on up to 32 x-nodes)");
    const int Nx_ = d->N_p + d->N_s + d->N_n;
    if (d->N_p < 2 || d->N_s < 2 || d->N_n < 2 || Nx_ > 64)
        return fail("plb_create: need 2 <= N_p,N_s,N_n and N_p+N_s+N_n <= 64 (one lane per node, one or two warps per system)");
    if (d->cathode != PLB_CATHODE_LCO && d->cathode != PLB_CATHODE_NMC && d->cathode != PLB_CATHODE_NMC_LGM50) return fail("plb_create: unknown cathode");
    const bool lgm = d->cathode == PLB_CATHODE_NMC_LGM50;
    if ((d->rxn_p != PLB_RXN_BV && d->rxn_p != PLB_RXN_MHC) || (d->rxn_n != PLB_RXN_BV && d->rxn_n != PLB_RXN_MHC))
        return fail("plb_create: unknown reaction rate law (built: rxn_BV, rxn_MHC)");
    // NMC() / LiC6_NMC() define no lambda_MHC_* (params.jl:295-367): the reference throws a KeyError there
    if ((d->rxn_p == PLB_RXN_MHC || d->rxn_n == PLB_RXN_MHC) && d->cathode != PLB_CATHODE_LCO)
        return fail("plb_create: rxn_MHC is built for the LCO parameter set (the NMC set has no lambda_MHC_p / lambda_MHC_n, NMC_LGM50 sets them to 0)");
    if (d->temperature) {
        // NMC()/LiC6_NMC() carry no thermal parameters (params.jl:295-367): the reference cannot build it either
        if (d->cathode == PLB_CATHODE_NMC) return fail("plb_create: temperature=true needs the LCO or NMC_LGM50 parameter set");
        if (d->N_p < 5 || d->N_n < 5) return fail("plb_create: temperature=true needs N_p, N_n >= 5");
        if (d->N_a < 1 || d->N_z < 1 || d->N_a + d->N_z > d->N_p + d->N_s + d->N_n)
            return fail("plb_create: temperature=true needs 1 <= N_a, N_z and N_a+N_z <= N_p+N_s+N_n (one collector node per lane)");
    }
    // one warp per system up to 32 x-nodes, unless the state vector outgrows that family's workspace stride
    const int Ntot_ = 2 * Nx_ + (NR_HOST + 2) * (d->N_p + d->N_n) + 1 + (d->aging ? 2 * d->N_n + 1 : 0);
    const int NtotT_ = Ntot_ + (d->temperature ? d->N_a + Nx_ + d->N_z : 0);
    const int want_mhc = (d->rxn_p == PLB_RXN_MHC || d->rxn_n == PLB_RXN_MHC) ? 1 : 0;
    if (d->fickian_spectral != 0 && d->fickian_spectral != 1) return fail("plb_create: unknown Fickian_method (0 = :finite_difference, 1 = :spectral)");
    const Variant* narrow = find_variant(d->temperature ? 1 : 0, d->aging ? 1 : 0, 0, want_mhc, lgm ? 1 : 0, NR_HOST, d->fickian_spectral);
    const bool wide = Nx_ > 32 || !narrow || NtotT_ > narrow->info().vs;
    const Variant* chosen = wide ? find_variant(d->temperature ? 1 : 0, d->aging ? 1 : 0, 1, want_mhc, lgm ? 1 : 0, NR_HOST, d->fickian_spectral) : narrow;
    if (chosen && NtotT_ > chosen->info().vs) chosen = nullptr;
    if (!chosen) {
        if (NR_HOST != 10)
            return fail("plb_create: N_r = 12 / 14 is built for the isothermal, thermal and SEI families on up to 32 x-nodes (LCO / NMC, rxn_BV)");
        if (d->fickian_spectral)
            // Fickian_method = :spectral (params.jl:142; residuals.jl:181-235, "BETA"): sibling builds of the 32-node families
            return fail("plb_create: Fickian_method = :spectral is built for the isothermal, thermal and SEI families on up to 32 x-nodes (LCO / NMC, rxn_BV, N_r = 10)");
        return fail("plb_create: no compiled family for this combination of options (system too large for the workspace stride?)");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("plb_create: no CUDA device available (this library has no CPU fallback)");
    if (d->device < 0 || d->device >= ndev) return fail("plb_create: no such CUDA device");
    plb_handle_s* h = new plb_handle_s();
    h->desc = *d;
    DeviceGuard guard_(d->device);
    if (!guard_.ok) { delete h; return fail("plb_create: cudaSetDevice failed"); }
    // every failure below releases what was allocated so far
#define CREATE_OK(x)                                                                                \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) { plb_destroy(h); return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } \
    } while (0)
    // one warp per system up to 32 x-nodes, unless the state vector outgrows that family's workspace stride
    // (many electrode nodes: N = 2 Nx + 12 Ne + 1): then the two-warp family runs it with its upper lanes idle
    h->v = chosen;
    h->has_dT = d->temperature != 0;
    h->vi = h->v->info();
    if (h->v == &V_ISO) h->v_dc = &V_ISODC;
    if (h->v == &V_WIDE) h->v_dc = &V_WIDEDC;
    if (h->v_dc) {
        const VariantInfo di = h->v_dc->info();      // same geometry and workspace: the two builds share every buffer
        if (di.sim_warps != h->vi.sim_warps || di.sim_ctas != h->vi.sim_ctas || di.vs != h->vi.vs || di.nglobal != h->vi.nglobal || di.gws_per_slot != h->vi.gws_per_slot) h->v_dc = nullptr;
    }
    ModelDesc& m = h->m;
    memset(&m, 0, sizeof m);
    m.Np = d->N_p; m.Ns = d->N_s; m.Nn = d->N_n; m.Nx = m.Np + m.Ns + m.Nn; m.Ne = m.Np + m.Nn;
    m.thermal = d->temperature ? 1 : 0;
    m.aging = d->aging ? 1 : 0;
    m.Na = m.thermal ? d->N_a : 0; m.Nz = m.thermal ? d->N_z : 0;
    m.chem = d->cathode == PLB_CATHODE_LCO ? CHEM_LCO : (lgm ? CHEM_LGM : CHEM_NMC);
    m.rxn_mhc = (d->rxn_p == PLB_RXN_MHC ? 1 : 0) | (d->rxn_n == PLB_RXN_MHC ? 2 : 0);
    m.mid = m.thermal ? m.Np + m.Ns / 2 : m.Nx / 2;
    m.inv_n[0] = 1.0 / m.Np; m.inv_n[1] = 1.0 / m.Ns; m.inv_n[2] = 1.0 / m.Nn; m.inv_n[3] = 0.0;
    if (m.aging) {
        // residuals_SOH! (residuals.jl:278-297): rhs = F a_n/(3600 I1C) * trapz over the anode of j_s after a quadratic
        // extrapolation to both ends (extrapolate_section / extrap_x_0, external.jl:496-522): linear in j_s, so node k
        // has a fixed weight.  For a section of unit length:
        const int N = m.Nn;
        auto xs = [&](int i) -> double { return i <= 0 ? 0.0 : (i > N ? 1.0 : (1.0 / (2.0 * N)) + (i - 1) * ((1.0 - 1.0 / N) / (N - 1))); };
        const double x1 = xs(1), x2 = xs(2), x3 = xs(3);
        const double r = (x3 - x1) / (x2 - x1);
        const double den = x3 * x3 - x1 * x1 - (x2 * x2 - x1 * x1) / (x2 - x1) * (x3 - x1);
        auto ext = [&](int i) -> double {     // weight of the i-th node from the end (1..3) in the end value
            const double c = (i == 1 ? r - 1.0 : (i == 2 ? -r : 1.0)) / den;
            const double b = ((i == 2 ? 1.0 : 0.0) - (i == 1 ? 1.0 : 0.0) - c * (x2 * x2 - x1 * x1)) / (x2 - x1);
            return (i == 1 ? 1.0 : 0.0) - c * x1 * x1 - b * x1;
        };
        const double w0 = 0.5 * (xs(1) - xs(0)), wN = 0.5 * (xs(N + 1) - xs(N));
        for (int k = 1; k <= N && k <= 64; k++) {
            double wk = 0.5 * (xs(k + 1) - xs(k - 1));
            if (k <= 3) wk += w0 * ext(k);
            if (k >= N - 2) wk += wN * ext(N + 1 - k);
            m.soh_geo[k - 1] = wk;
        }
    }
    m.off_cs = m.Nx; m.off_T = m.off_cs + NR_HOST * m.Ne;
    m.off_film = m.off_T + (m.thermal ? m.Na + m.Nx + m.Nz : 0);
    m.off_SOH = m.off_film + (m.aging ? m.Nn : 0);
    m.off_j = m.off_SOH + (m.aging ? 1 : 0); m.N_diff = m.off_j;
    m.off_pe = m.off_j + m.Ne; m.off_ps = m.off_pe + m.Nx; m.off_js = m.off_ps + m.Ne;
    m.off_I = m.off_js + (m.aging ? m.Nn : 0); m.N_tot = m.off_I + 1;
    if (m.N_tot > h->vi.vs) { plb_destroy(h); return fail("plb_create: system too large for the workspace stride"); }
    for (int f = 0; f < TF_COUNT; f++) m.slot[f] = -1;
    // used keys in the reference's (code-point sorted) order; KEYS[] is already sorted that way
    for (int k = 0; k < (int)(sizeof(KEYS) / sizeof(KEYS[0])); k++) {
        const double dv = d->cathode == PLB_CATHODE_LCO ? KEYS[k].lco : (d->cathode == PLB_CATHODE_NMC_LGM50 ? KEYS[k].lgm : KEYS[k].nmc);
        if (dv != dv) continue;
        if (KEYS[k].only && !((KEYS[k].only & 1) && m.thermal) && !((KEYS[k].only & 2) && m.aging) &&
            !((KEYS[k].only & 4) && (m.rxn_mhc & 1)) && !((KEYS[k].only & 8) && (m.rxn_mhc & 2))) continue;
        m.slot[KEYS[k].field] = (int8_t)h->keys.size();
        h->keys.push_back(k);
    }
    m.ntheta = (int)h->keys.size();
    m.theta_stride = m.ntheta;
    cudaDeviceProp prop;
    CREATE_OK(cudaGetDeviceProperties(&prop, d->device));
    h->num_sms = prop.multiProcessorCount;
    if (build_patterns(h)) { const std::string e = g_err; plb_destroy(h); return fail(e); }
    CREATE_OK(cudaMalloc(&h->d_counter, sizeof(int)));
    h->sim_grid = h->num_sms * h->vi.sim_ctas;
    CREATE_OK(cudaMalloc(&h->d_gws, (size_t)h->sim_grid * h->vi.sim_warps * h->vi.gws_per_slot * sizeof(double)));
    CREATE_OK(cudaEventCreate(&h->ev0));
    CREATE_OK(cudaEventCreate(&h->ev1));
#undef CREATE_OK
    *out = h;
    return 0;
}

int plb_variant_info(int family, long long* out) {
    const VariantInfo v = family == 1 ? th::info() : (family == 2 ? sei::info() : (family == 3 ? wide::info() : (family == 4 ? wsei::info() : (family == 5 ? wth::info() : (family == 6 ? thsei::info() : (family == 7 ? wthsei::info() : iso::info()))))));
    out[0] = v.sim_warps; out[1] = v.sim_ctas; out[2] = (long long)v.sim_smem; out[3] = v.k1_warps;
    out[4] = v.k1_ctas; out[5] = (long long)v.k1_smem; out[6] = v.vs; out[7] = v.n_slots;
    return 0;
}

int plb_destroy(plb_handle h) {
    if (!h) return 0;
    DeviceGuard guard_(h->desc.device);
    for (int i = 0; i < N_METHODS; i++) cudaFree(h->d_src[i]);
    for (int i = 0; i < plb_handle_s::NPOOL; i++) if (h->pool_ptr[i]) cudaFree(h->pool_ptr[i]);
    cudaFree(h->d_counter);
    cudaFree(h->d_gws);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return 0;
}
int plb_set_stream(plb_handle h, void* s) { if (!h) return fail("plb_set_stream: null handle"); h->stream = (cudaStream_t)s; return 0; }
int plb_nstates(plb_handle h) { return h ? h->m.N_tot : fail("plb_nstates: null handle"); }
int plb_ndiff(plb_handle h) { return h ? h->m.N_diff : fail("plb_ndiff: null handle"); }
int plb_ntheta(plb_handle h) { return h ? h->m.ntheta : fail("plb_ntheta: null handle"); }
int plb_jac_nnz(plb_handle h, int method) { return (!h || !h->method_ok(method)) ? -1 : (int)h->rowval[method].size(); }
long long plb_launch_count(plb_handle h) { return h ? h->launches : -1; }
float plb_last_kernel_ms(plb_handle h) { return h ? h->last_ms : 0.f; }

int plb_theta_keys(plb_handle h, const char** keys) {
    if (!h || !keys) return fail("plb_theta_keys: null argument");
    for (size_t i = 0; i < h->keys.size(); i++) keys[i] = KEYS[h->keys[i]].utf8;
    return (int)h->keys.size();
}
int plb_theta_index(plb_handle h, const char* key) {
    if (!h || !key) return -1;
    for (size_t i = 0; i < h->keys.size(); i++)
        if (!strcmp(KEYS[h->keys[i]].utf8, key) || !strcmp(KEYS[h->keys[i]].ascii, key)) return (int)i;
    return -1;
}
int plb_theta_defaults(plb_handle h, double* row) {
    if (!h || !row) return fail("plb_theta_defaults: null argument");
    for (size_t i = 0; i < h->keys.size(); i++)
        row[i] = h->desc.cathode == PLB_CATHODE_LCO ? KEYS[h->keys[i]].lco
                 : (h->desc.cathode == PLB_CATHODE_NMC_LGM50 ? KEYS[h->keys[i]].lgm : KEYS[h->keys[i]].nmc);
    return 0;
}
int plb_bounds_defaults(plb_handle h, plb_bounds* b) {
    if (!h || !b) return fail("plb_bounds_defaults: null argument");
    // src/params.jl:233-253 (LCO), :451-471 (NMC)
    const double nan_ = NAN;
    if (h->desc.cathode == PLB_CATHODE_LCO) { b->V_min = 2.5; b->V_max = 4.3; b->T_max = 55 + 273.15; }
    else if (h->desc.cathode == PLB_CATHODE_NMC_LGM50) { b->V_min = 2.5; b->V_max = 4.2; b->T_max = 55 + 273.15; }   // params.jl:760-773
    else { b->V_min = 2.8; b->V_max = 4.2; b->T_max = nan_; }
    b->SOC_min = 0.0; b->SOC_max = 1.0; b->c_s_n_max = nan_; b->I_max = nan_; b->I_min = nan_;
    b->eta_plating_min = nan_; b->c_e_min = nan_; b->dfilm_max = nan_;
    return 0;
}
int plb_opts_defaults(plb_handle, plb_opts* o) {
    // src/params.jl:256-280
    o->abstol = 1e-6; o->reltol = 1e-3; o->abstol_init = 1e-6; o->reltol_init = 1e-3;
    o->maxiters = 10000; o->check_bounds = 1; o->interp_final = 1; o->skip_alg_deriv = 0;
    return 0;
}
int plb_calc_I1C(plb_handle h, int B, const double* theta, double* I1C) {
    if (!h || !theta || !I1C) return fail("plb_calc_I1C: null argument");
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
    if (!h || !colptr || !rowval) return fail("plb_jac_pattern: null argument");
    if (!h->method_ok(method)) return fail("plb_jac_pattern: bad method");
    const int o = one_based ? 1 : 0;
    for (size_t i = 0; i < h->colptr[method].size(); i++) colptr[i] = h->colptr[method][i] + o;
    for (size_t i = 0; i < h->rowval[method].size(); i++) rowval[i] = h->rowval[method][i] + o;
    return 0;
}

// device staging helper for PLB_MEM_HOST calls: slot `k` of the handle's grow-only pool
struct DevBuf {
    void* p = nullptr;
    size_t n = 0;
    plb_handle_s* h = nullptr;
    int k = 0;
    int alloc(size_t bytes) {
        n = bytes;
        if (!bytes) return 0;
        if (h->pool_cap[k] < bytes) {
            if (h->pool_ptr[k]) cudaFree(h->pool_ptr[k]);
            h->pool_ptr[k] = nullptr; h->pool_cap[k] = 0;
            if (cudaMalloc(&h->pool_ptr[k], bytes) != cudaSuccess) return -1;
            h->pool_cap[k] = bytes;
        }
        p = h->pool_ptr[k];
        return 0;
    }
};
// the staging slots of one API call
struct DevBufs {
    DevBuf b[plb_handle_s::NPOOL];
    explicit DevBufs(plb_handle_s* h) { for (int i = 0; i < plb_handle_s::NPOOL; i++) { b[i].h = h; b[i].k = i; } }
    DevBuf& operator[](int i) { return b[i]; }
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
    r.skip_alg_deriv = o->skip_alg_deriv;
    // Sundials.jl IDA() constructor values (third-party): max_order 5, max_nonlinear_iters 3,
    // max_error_test_failures 7, max_convergence_failures 10
    r.maxord = 5; r.maxcor = 3; r.maxnef = 7; r.maxncf = 10;
    return r;
}

int plb_initial_guess(plb_handle h, int B, const double* soc, const double* theta, double* Y0, int mem) {
    PLB_ENTER(h, "plb_initial_guess");
    if (B <= 0) return 0;
    if (!soc || !theta || !Y0) return fail("plb_initial_guess: null argument");
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    DevBufs db(h);
    DevBuf &b1 = db[0], &b2 = db[1], &b3 = db[2];
    double* hostY;
    if (stage_in(b1, soc, (size_t)B, mem, s) || stage_in(b2, theta, (size_t)B * m.ntheta, mem, s) ||
        stage_inout(b3, Y0, hostY, (size_t)B * m.N_tot, mem, false, s)) return -1;
    AuxArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.B = B; a.theta = theta; a.soc = soc; a.Y = Y0; a.gws = h->d_gws;
    const int grid = std::min((B + h->vi.sim_warps - 1) / h->vi.sim_warps, h->sim_grid);
    CUDA_OK(h->v->initguess(a, grid, s));
    h->launches++;
    if (stage_out(Y0, hostY, (size_t)B * m.N_tot, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int plb_resjac(plb_handle h, int B, const double* Y, const double* YP, const double* gamma,
               const double* theta, const plb_run* run, const double* values, double* res,
               double* nzval, int mem) {
    PLB_ENTER(h, "plb_resjac");
    if (B <= 0) return 0;
    if (!run || !h->method_ok(run->method)) return fail("plb_resjac: bad run (dT needs temperature=true)");
    if (!Y || !YP || !theta) return fail("plb_resjac: null required argument");
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    const int nnz = (int)h->rowval[run->method].size();
    DevBufs db(h);
    DevBuf &b1 = db[0], &b2 = db[1], &b3 = db[2], &b4 = db[3], &b5 = db[4], &b6 = db[5], &b7 = db[6];
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
    // the bulk-copy (TMA) kernel needs 16-byte aligned input arrays; anything else takes the per-lane loads
    a.use_tma = h->vi.k1_tma && !h->k1_no_tma && (((uintptr_t)Y | (uintptr_t)YP | (uintptr_t)theta) & 15) == 0;
    // one resident wave: every CTA loads the recipe tables once and then streams its share of the batch
    const int grid = std::min((B + h->vi.k1_warps - 1) / h->vi.k1_warps, h->num_sms * h->vi.k1_ctas * (a.use_tma ? 1 : 2));
    CUDA_OK(cudaEventRecord(h->ev0, s));
    CUDA_OK(h->v->resjac(a, grid, s));
    CUDA_OK(cudaEventRecord(h->ev1, s));
    h->launches++;
    if (stage_out(res, hostR, (size_t)B * m.N_tot, s) || stage_out(nzval, hostN, (size_t)B * nnz, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
    return 0;
}

int plb_newton_init(plb_handle h, int B, double* Y, double* YP, const double* theta, const plb_run* run,
                    const double* values, const plb_opts* opts, int* status, int mem) {
    PLB_ENTER(h, "plb_newton_init");
    if (B <= 0) return 0;
    if (!run || !opts || !Y || !YP || !theta) return fail("plb_newton_init: null required argument");
    if (!h->method_ok(run->method)) return fail("plb_newton_init: bad method");
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    DevBufs db(h);
    DevBuf &b1 = db[0], &b2 = db[1], &b3 = db[2], &b4 = db[3], &b5 = db[4];
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
    const int grid = std::min((B + h->vi.sim_warps - 1) / h->vi.sim_warps, h->sim_grid);
    CUDA_OK(h->v->newton(a, grid, s));
    h->launches++;
    if (stage_out(Y, hostY, (size_t)B * m.N_tot, s) || stage_out(YP, hostYP, (size_t)B * m.N_tot, s) ||
        stage_out(status, hostS, (size_t)B, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int plb_linear_solve(plb_handle h, int B, const double* Y, const double* YP, const double* gamma,
                     const double* theta, const plb_run* run, const double* values, const double* rhs,
                     double* x, int* status, int mem) {
    PLB_ENTER(h, "plb_linear_solve");
    if (B <= 0) return 0;
    if (!run || !h->method_ok(run->method)) return fail("plb_linear_solve: bad run");
    if (!Y || !YP || !theta || !rhs || !x) return fail("plb_linear_solve: null required argument");
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    const size_t BN = (size_t)B * m.N_tot;
    DevBufs db(h);
    DevBuf &b1 = db[0], &b2 = db[1], &b3 = db[2], &b4 = db[3], &b5 = db[4], &b6 = db[5], &b7 = db[6], &b8 = db[7];
    double* hostX;
    int* hostS;
    if (stage_in(b1, Y, BN, mem, s) || stage_in(b2, YP, BN, mem, s) || stage_in(b3, gamma, (size_t)B, mem, s) ||
        stage_in(b4, theta, (size_t)B * m.ntheta, mem, s) || stage_in(b5, values, (size_t)B, mem, s) ||
        stage_in(b6, rhs, BN, mem, s) || stage_inout(b7, x, hostX, BN, mem, false, s) ||
        stage_inout(b8, status, hostS, (size_t)B, mem, false, s)) return -1;
    AuxArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.B = B; a.theta = theta; a.values = values; a.method = run->method; a.value = run->value;
    a.Y = const_cast<double*>(Y); a.YP = const_cast<double*>(YP); a.gamma = gamma; a.rhs = rhs; a.x = x;
    a.status = status; a.gws = h->d_gws;
    const int grid = std::min((B + h->vi.sim_warps - 1) / h->vi.sim_warps, h->sim_grid);
    CUDA_OK(h->v->linsolve(a, grid, s));
    h->launches++;
    if (stage_out(x, hostX, BN, s) || stage_out(status, hostS, (size_t)B, s)) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

static int simulate_impl(plb_handle h, int B, const double* theta, const plb_run* run, const plb_input_table* tab,
                         const double* values, const plb_opts* opts, const plb_bounds* bounds, const double* soc0,
                         double* sY, double* sYP, double* sSOC, double* st, plb_summary* summary, int n_save_max,
                         double* tr_t, double* tr_V, double* tr_I, double* tr_SOC, double* tr_T, double* tr_Y, int* tr_n, int mem) {
    PLB_ENTER(h, "plb_simulate");
    // the dense-output request is one-shot: whatever happens to this call, the next one starts without it
    auto dense = h->dense;
    h->dense = {};
    if (B <= 0) return 0;
    if (!run || !opts || !bounds || !theta || !sY || !sSOC || !st || !summary)
        return fail("plb_simulate: null required argument");
    if (dense.n && dense.mem != mem) return fail("plb_simulate: the dense-output buffers must live where the other buffers live");
    const bool dc = run->method >= PLB_METHOD_DC_S_P_MAX && run->method <= PLB_METHOD_DC_E_MIN;
    if (dc) {
        if (!h->v_dc) return fail("plb_simulate: the concentration-rate inputs dc_s_* / dc_e_* are built for isothermal models without aging");
        if (run->new_run) return fail("plb_simulate: dc_s_* / dc_e_* need a previous solution (input_methods.jl:196)");
        if (run->input_kind == PLB_INPUT_REST || tab) return fail("plb_simulate: dc_s_* / dc_e_* take a number or :hold");
    } else if (!h->method_ok(run->method))
        return fail(run->method == PLB_METHOD_DT ? "plb_simulate: Temperature must be enabled when using `dT`."   // input_methods.jl:183
                                                 : "plb_simulate: bad method");
    if (run->input_kind != PLB_INPUT_VALUE && run->new_run && run->input_kind == PLB_INPUT_HOLD)
        return fail("plb_simulate: Cannot use `:hold` without a previous simulation.");   // checks.jl:385
    if (run->input_kind == PLB_INPUT_REST && run->method != PLB_METHOD_I && run->method != PLB_METHOD_P)
        return fail("plb_simulate: Unsupported input symbol.");
    static_assert(sizeof(plb_summary) == sizeof(Summary), "summary layout");
    // run_function as a table: validate, and merge the stop list of postfix_integrator!
    // (model_evaluation.jl:288-310): sort([tdiscon .- reltol/2; 1.0 if continuing; tf]), non-positive entries
    // dropped; entries at or past tf are never reached (the run ends at tf) and are left out
    std::vector<double> tstops;
    const double *tab_t = nullptr, *tab_v = nullptr, *d_tstops = nullptr;
    if (tab) {
        if (tab->n_tdiscon < 0 || (tab->n_tdiscon > 0 && !tab->tdiscon)) return fail("plb_simulate_table: bad tdiscon");
        for (int k = 1; k < tab->n_tdiscon; k++)
            if (tab->tdiscon[k] < tab->tdiscon[k - 1]) return fail("plb_simulate_table: tdiscon must be ascending");
    }
    if (tab || !h->opt_tstops.empty()) {
        // sort([opts.tstops; tdiscon .- reltol/2; 1.0 if continuing; tf]) without the entries <= 0 (and those >= tf)
        for (double ts : h->opt_tstops) if (ts > 0.0 && ts < run->tf) tstops.push_back(ts);
        for (int k = 0; tab && k < tab->n_tdiscon; k++) {
            const double ts = tab->tdiscon[k] - opts->reltol / 2;
            if (ts > 0.0 && ts < run->tf) tstops.push_back(ts);
        }
        if (!run->new_run && 1.0 < run->tf) tstops.push_back(1.0);
        std::sort(tstops.begin(), tstops.end());
        tstops.push_back(run->tf);
        d_tstops = tstops.data();
    }
    if (tab) {
        if (run->input_kind != PLB_INPUT_VALUE) return fail("plb_simulate_table: a table input cannot be :hold or :rest");
        if (run->method == PLB_METHOD_DT) return fail("plb_simulate_table: dT takes a number or :hold");
        if (tab->n < 1 || !tab->t || !tab->v) return fail("plb_simulate_table: empty table");
        for (int k = 0; k < tab->n; k++) {
            if (!std::isfinite(tab->t[k]) || !std::isfinite(tab->v[k])) return fail("plb_simulate_table: non-finite knot");
            if (k && tab->t[k] < tab->t[k - 1]) return fail("plb_simulate_table: knot times must be non-decreasing");
            if (k >= 2 && tab->t[k] == tab->t[k - 2]) return fail("plb_simulate_table: more than two knots at one time");
        }
        tab_t = tab->t; tab_v = tab->v;
    }
    const ModelDesc& m = h->m;
    cudaStream_t s = h->stream;
    const size_t BN = (size_t)B * m.N_tot, BS = (size_t)B * (n_save_max > 0 ? n_save_max : 0);
    DevBufs b(h);
    double *hY, *hYP, *hSOC, *ht, *htt, *htV, *htI, *htS, *htT, *htY;
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
        stage_inout(b[12], tr_n, htn, (size_t)B, mem, false, s) ||
        stage_inout(b[13], tr_T, htT, BS, mem, false, s) || stage_inout(b[17], tr_Y, htY, BS * m.N_tot, mem, false, s)) return -1;
    if (tab && (stage_in(b[14], tab_t, (size_t)tab->n, PLB_MEM_HOST, s) || stage_in(b[15], tab_v, (size_t)tab->n, PLB_MEM_HOST, s))) return -1;
    if (d_tstops && stage_in(b[16], d_tstops, tstops.size(), PLB_MEM_HOST, s)) return -1;
    const double* dn_t = dense.n ? dense.t.data() : nullptr;
    double *hdV, *hdI, *hdS, *hdT, *hdY;
    int* hdn;
    const size_t BD = (size_t)B * dense.n;
    if (dense.n && (stage_in(b[18], dn_t, (size_t)dense.n, PLB_MEM_HOST, s) || stage_inout(b[19], dense.V, hdV, BD, mem, false, s) ||
                    stage_inout(b[20], dense.I, hdI, BD, mem, false, s) || stage_inout(b[21], dense.SOC, hdS, BD, mem, false, s) ||
                    stage_inout(b[22], dense.T, hdT, BD, mem, false, s) || stage_inout(b[23], dense.Y, hdY, BD * m.N_tot, mem, false, s) ||
                    stage_inout(b[24], dense.n_done, hdn, (size_t)B, mem, false, s))) return -1;
    SimArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.B = B; a.theta = theta; a.values = values; a.method = dc ? METHOD_DC : run->method; a.value = run->value;
    a.dc_kind = dc ? run->method - PLB_METHOD_DC_S_P_MAX : 0;
    a.tf = run->tf; a.input_kind = run->input_kind; a.new_run = run->new_run; a.o = to_opts(opts);
    memcpy(&a.b, bounds, sizeof(Bounds));
    static_assert(sizeof(plb_bounds) == sizeof(Bounds), "bounds layout");
    a.soc0 = soc0; a.sY = sY; a.sYP = sYP; a.sSOC = sSOC; a.st = st; a.out = (Summary*)summary;
    a.n_save_max = n_save_max > 0 ? n_save_max : 0;
    a.tr_t = tr_t; a.tr_V = tr_V; a.tr_I = tr_I; a.tr_SOC = tr_SOC; a.tr_T = tr_T; a.tr_Y = tr_Y; a.tr_n = tr_n;
    a.counter = h->d_counter;
    a.gws = h->d_gws;
    if (tab) { a.tab_n = tab->n; a.tab_t = tab_t; a.tab_v = tab_v; }
    if (d_tstops) { a.n_tstops = (int)tstops.size(); a.tstops = d_tstops; }
    if (dense.n) {
        a.n_dense = dense.n; a.dense_t = dn_t;
        a.dn_V = dense.V; a.dn_I = dense.I; a.dn_SOC = dense.SOC; a.dn_T = dense.T; a.dn_Y = dense.Y; a.dn_n = dense.n_done;
        // rows past the end of a run stay NaN (0xff bytes), the row counters start at zero
        for (double* q : {dense.V, dense.I, dense.SOC, dense.T}) if (q) CUDA_OK(cudaMemsetAsync(q, 0xff, BD * sizeof(double), s));
        if (dense.Y) CUDA_OK(cudaMemsetAsync(dense.Y, 0xff, BD * m.N_tot * sizeof(double), s));
        if (dense.n_done) CUDA_OK(cudaMemsetAsync(dense.n_done, 0, (size_t)B * sizeof(int), s));
    }
    CUDA_OK(cudaMemsetAsync(h->d_counter, 0, sizeof(int), s));
    const int grid = std::min((B + h->vi.sim_warps - 1) / h->vi.sim_warps, h->sim_grid);
    CUDA_OK(cudaEventRecord(h->ev0, s));
    CUDA_OK((dc ? h->v_dc : h->v)->simulate(a, grid, s));
    CUDA_OK(cudaEventRecord(h->ev1, s));
    h->launches++;
    if (stage_out(sY, hY, BN, s) || stage_out(sYP, hYP, BN, s) || stage_out(sSOC, hSOC, (size_t)B, s) ||
        stage_out(st, ht, (size_t)B, s) || stage_out(summary, hsum, (size_t)B, s) || stage_out(tr_t, htt, BS, s) ||
        stage_out(tr_V, htV, BS, s) || stage_out(tr_I, htI, BS, s) || stage_out(tr_SOC, htS, BS, s) ||
        stage_out(tr_n, htn, (size_t)B, s) || stage_out(tr_T, htT, BS, s) || stage_out(tr_Y, htY, BS * m.N_tot, s)) return -1;
    if (dense.n && (stage_out(dense.V, hdV, BD, s) || stage_out(dense.I, hdI, BD, s) || stage_out(dense.SOC, hdS, BD, s) ||
                    stage_out(dense.T, hdT, BD, s) || stage_out(dense.Y, hdY, BD * m.N_tot, s) || stage_out(dense.n_done, hdn, (size_t)B, s))) return -1;
    CUDA_OK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
    return 0;
}

int plb_simulate(plb_handle h, int B, const double* theta, const plb_run* run, const double* values,
                 const plb_opts* opts, const plb_bounds* bounds, const double* soc0, double* sY,
                 double* sYP, double* sSOC, double* st, plb_summary* summary, int n_save_max,
                 double* tr_t, double* tr_V, double* tr_I, double* tr_SOC, double* tr_T, double* tr_Y, int* tr_n, int mem) {
    return simulate_impl(h, B, theta, run, nullptr, values, opts, bounds, soc0, sY, sYP, sSOC, st, summary, n_save_max,
                         tr_t, tr_V, tr_I, tr_SOC, tr_T, tr_Y, tr_n, mem);
}

int plb_simulate_table(plb_handle h, int B, const double* theta, const plb_run* run, const plb_input_table* table,
                       const double* scale, const plb_opts* opts, const plb_bounds* bounds, const double* soc0,
                       double* sY, double* sYP, double* sSOC, double* st, plb_summary* summary, int n_save_max,
                       double* tr_t, double* tr_V, double* tr_I, double* tr_SOC, double* tr_T, double* tr_Y, int* tr_n, int mem) {
    if (!table) return fail("plb_simulate_table: null table");
    return simulate_impl(h, B, theta, run, table, scale, opts, bounds, soc0, sY, sYP, sSOC, st, summary, n_save_max,
                         tr_t, tr_V, tr_I, tr_SOC, tr_T, tr_Y, tr_n, mem);
}

int plb_set_dense_output(plb_handle h, int n, const double* t_global, double* V, double* I, double* SOC, double* T,
                         double* Y, int* n_done, int mem) {
    if (!h) return fail("plb_set_dense_output: null handle");
    h->dense = {};
    if (n == 0) return 0;
    if (n < 0 || !t_global) return fail("plb_set_dense_output: bad arguments");
    for (int k = 0; k < n; k++) {
        if (!std::isfinite(t_global[k])) return fail("plb_set_dense_output: non-finite time");
        if (k && t_global[k] < t_global[k - 1]) return fail("plb_set_dense_output: times must be ascending");
    }
    h->dense.n = n; h->dense.mem = mem; h->dense.t.assign(t_global, t_global + n);
    h->dense.V = V; h->dense.I = I; h->dense.SOC = SOC; h->dense.T = T; h->dense.Y = Y; h->dense.n_done = n_done;
    return 0;
}

int plb_set_tstops(plb_handle h, int n, const double* tstops) {
    if (!h) return fail("plb_set_tstops: null handle");
    if (n < 0 || (n > 0 && !tstops)) return fail("plb_set_tstops: bad arguments");
    for (int k = 0; k < n; k++) if (!std::isfinite(tstops[k])) return fail("plb_set_tstops: non-finite stop time");
    h->opt_tstops.assign(tstops, tstops + n);
    return 0;
}


// =================================================================================================
// multi-GPU fan-out: one handle per device, contiguous batch shards, one NCCL all-gather of the summaries
// =================================================================================================
namespace {
// the few NCCL entry points this path needs, bound at run time (the product has no link-time dependency on NCCL)
struct Nccl {
    void* lib = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString;
    }
};
Nccl g_nccl;
}  // namespace

struct plb_group_s {
    std::vector<plb_handle_s*> h;
    std::vector<int> dev;
    std::vector<void*> comm;                 // ncclComm_t per device (empty: NCCL unavailable)
    std::vector<cudaStream_t> stream;
    std::vector<plb_summary*> d_send, d_all; // per device: its block [R], the gathered batch [n_dev x R]
    size_t cap_rows = 0;
    int rows = 0;                            // R of the last simulate
    bool gathered = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float gather_ms = 0.f;
};

int plb_group_create(const plb_model_desc* desc, int n_dev, const int* devices, plb_group* out) {
    if (!desc || !devices || !out || n_dev < 1) return fail("plb_group_create: bad arguments");
    plb_group_s* g = new plb_group_s();
    for (int k = 0; k < n_dev; k++) {
        plb_model_desc d = *desc;
        d.device = devices[k];
        plb_handle hk = nullptr;
        if (plb_create(&d, &hk)) { const std::string e = g_err; plb_group_destroy(g); return fail(e); }
        g->h.push_back(hk);
        g->dev.push_back(devices[k]);
    }
    g->stream.assign(n_dev, nullptr);
    g->d_send.assign(n_dev, nullptr);
    g->d_all.assign(n_dev, nullptr);
    for (int k = 0; k < n_dev; k++) {
        DeviceGuard guard(devices[k]);
        if (cudaStreamCreateWithFlags(&g->stream[k], cudaStreamNonBlocking) != cudaSuccess) { plb_group_destroy(g); return fail("plb_group_create: cudaStreamCreate failed"); }
        g->h[k]->stream = g->stream[k];
    }
    if (n_dev > 1 && g_nccl.load()) {
        g->comm.assign(n_dev, nullptr);
        const int rc = g_nccl.CommInitAll(g->comm.data(), n_dev, devices);
        if (rc != 0) { const std::string e = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(rc); g->comm.clear(); plb_group_destroy(g); return fail(e); }
    }
    {
        DeviceGuard guard(devices[0]);
        cudaEventCreate(&g->ev0); cudaEventCreate(&g->ev1);
    }
    *out = g;
    return 0;
}

int plb_group_destroy(plb_group g) {
    if (!g) return 0;
    for (size_t k = 0; k < g->comm.size(); k++) if (g->comm[k]) g_nccl.CommDestroy(g->comm[k]);
    for (size_t k = 0; k < g->h.size(); k++) {
        DeviceGuard guard(g->dev[k]);
        if (k < g->d_send.size()) { cudaFree(g->d_send[k]); cudaFree(g->d_all[k]); }
        if (k < g->stream.size() && g->stream[k]) { g->h[k]->stream = nullptr; cudaStreamDestroy(g->stream[k]); }
        if (k == 0) { if (g->ev0) cudaEventDestroy(g->ev0); if (g->ev1) cudaEventDestroy(g->ev1); }
        plb_destroy(g->h[k]);
    }
    delete g;
    return 0;
}
int plb_group_size(plb_group g) { return g ? (int)g->h.size() : fail("plb_group_size: null group"); }
plb_handle plb_group_handle(plb_group g, int k) { return (g && k >= 0 && k < (int)g->h.size()) ? g->h[k] : nullptr; }
float plb_group_last_gather_ms(plb_group g) { return g ? g->gather_ms : 0.f; }

int plb_group_simulate(plb_group g, int B, const double* theta, const plb_run* run, const double* values, const plb_opts* opts,
                       const plb_bounds* bounds, const double* soc0, double* sY, double* sYP, double* sSOC, double* st,
                       plb_summary* summary, int n_save_max, double* tr_t, double* tr_V, double* tr_I, double* tr_SOC,
                       double* tr_T, int* tr_n) {
    if (!g) return fail("plb_group_simulate: null group");
    if (B <= 0) return 0;
    if (!run || !opts || !bounds || !theta || !sY || !sSOC || !st || !summary) return fail("plb_group_simulate: null required argument");
    const int G = (int)g->h.size();
    const int R = (B + G - 1) / G;
    const int N = g->h[0]->m.N_tot, nth = g->h[0]->m.ntheta;
    const size_t ns = n_save_max > 0 ? n_save_max : 0;
    g->gathered = false;
    g->rows = R;
    // every shard from its own host thread: the per-device calls are synchronous, the devices run side by side
    std::vector<int> rc(G, 0);
    std::vector<std::string> err(G);
    std::vector<std::thread> th;
    for (int k = 0; k < G; k++) {
        const int lo = k * R, n = std::max(0, std::min(B, lo + R) - lo);
        th.emplace_back([=, &rc, &err]() {
            if (n == 0) return;
            auto off = [&](auto* p, size_t stride) { return p ? p + (size_t)lo * stride : p; };
            rc[k] = plb_simulate(g->h[k], n, theta + (size_t)lo * nth, run, off(values, 1), opts, bounds, off(soc0, 1), sY + (size_t)lo * N,
                                 off(sYP, N), sSOC + lo, st + lo, summary + lo, n_save_max, off(tr_t, ns), off(tr_V, ns), off(tr_I, ns),
                                 off(tr_SOC, ns), off(tr_T, ns), nullptr, off(tr_n, 1), PLB_MEM_HOST);
            if (rc[k]) err[k] = g_err;      // (thread-local in the worker: hand it to the caller)
        });
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < G; k++) if (rc[k]) return fail("device " + std::to_string(g->dev[k]) + ": " + err[k]);
    if (g->comm.empty()) return 0;
    // ---- the one collective: all-gather of the summaries, so that every device holds the whole batch's ----
    if (g->cap_rows < (size_t)R) {
        for (int k = 0; k < G; k++) {
            DeviceGuard guard(g->dev[k]);
            cudaFree(g->d_send[k]); cudaFree(g->d_all[k]);
            g->d_send[k] = g->d_all[k] = nullptr;
            CUDA_OK(cudaMalloc(&g->d_send[k], (size_t)R * sizeof(plb_summary)));
            CUDA_OK(cudaMalloc(&g->d_all[k], (size_t)R * G * sizeof(plb_summary)));
        }
        g->cap_rows = R;
    }
    for (int k = 0; k < G; k++) {
        DeviceGuard guard(g->dev[k]);
        const int lo = k * R, n = std::max(0, std::min(B, lo + R) - lo);
        CUDA_OK(cudaMemsetAsync(g->d_send[k], 0, (size_t)R * sizeof(plb_summary), g->stream[k]));
        // the device copy of the shard's summaries is still in the handle's staging pool (slot 7 of simulate_impl)
        if (n) CUDA_OK(cudaMemcpyAsync(g->d_send[k], g->h[k]->pool_ptr[7], (size_t)n * sizeof(plb_summary), cudaMemcpyDeviceToDevice, g->stream[k]));
    }
    {
        DeviceGuard guard(g->dev[0]);
        CUDA_OK(cudaEventRecord(g->ev0, g->stream[0]));
    }
    int nrc = g_nccl.GroupStart();
    for (int k = 0; k < G && nrc == 0; k++)
        nrc = g_nccl.AllGather(g->d_send[k], g->d_all[k], (size_t)R * sizeof(plb_summary), /*ncclUint8*/ 1, g->comm[k], g->stream[k]);
    const int erc = g_nccl.GroupEnd();
    if (nrc == 0) nrc = erc;
    if (nrc != 0) return fail(std::string("ncclAllGather: ") + g_nccl.GetErrorString(nrc));
    {
        DeviceGuard guard(g->dev[0]);
        CUDA_OK(cudaEventRecord(g->ev1, g->stream[0]));
    }
    for (int k = 0; k < G; k++) {
        DeviceGuard guard(g->dev[k]);
        CUDA_OK(cudaStreamSynchronize(g->stream[k]));
    }
    cudaEventElapsedTime(&g->gather_ms, g->ev0, g->ev1);
    g->gathered = true;
    return 0;
}

int plb_group_device_summaries(plb_group g, int k, const plb_summary** out, int* rows) {
    if (!g || !out || k < 0 || k >= (int)g->h.size()) return fail("plb_group_device_summaries: bad arguments");
    if (!g->gathered) return fail(g->comm.empty() ? "plb_group_device_summaries: NCCL is not available (libnccl.so.2 not found, or one device)"
                                                  : "plb_group_device_summaries: no gathered batch yet");
    *out = g->d_all[k];
    if (rows) *rows = g->rows;
    return 0;
}
