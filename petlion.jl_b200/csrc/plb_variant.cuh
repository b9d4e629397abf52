// plb_variant.cuh -- kernels and launchers of ONE model family; included by plb_variant_iso.cu (PLB_TH=0,
// namespace plb::iso) and plb_variant_th.cu (PLB_TH=1, namespace plb::th).
//
// Kernels (all FP64, sm_100a, warp-per-system):
//   k_resjac     K1: batched residual + CSC Jacobian values  (R_full / J_full callback surface,
//                /root/reference/src/physics_equations/scalar_residual.jl:558-602)
//   k_initguess  initial_guess!                                (states_definition.jl:80-121)
//   k_newton     K3: newtons_method!                           (model_evaluation.jl:430-480)
//   k_linsolve   factorize + solve of the Newton matrix        (KLU inside IDA, model_evaluation.jl:265-271, 417-428)
//   k_simulate   K4: fused persistent integrator               (model_evaluation.jl:312-382 + IDA)
#pragma once
#include "plb_integrator.cuh"

namespace plb {
namespace PLB_NS {

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
#if PLB_SPECTRAL
    JS_CS_J_LAST = JS_CS_J + NR - 1,     // Fickian_method = :spectral: NR slots, the j column reaches every radial row
#endif
#if PLB_SEI
    // aging = :SEI (anode lanes): j row d/dfilm; j_s row; film row; SOH row; d/dj_s of the c_e, Phi_e, Phi_s rows
    JS_J_FILM, JS_JS_PS, JS_JS_PE, JS_JS_J, JS_JS_JS, JS_JS_FILM, JS_JS_I, JS_FILM_JS, JS_FILM_D,
    JS_SOH_JS, JS_SOH_D, JS_CE_JS, JS_PE_JS, JS_PS_JS,
#endif
#if PLB_TH
    JS_CS_T0,                                       // 10 slots: d res_cs[r] / dT
    JS_J_T = JS_CS_T0 + NR, JS_PE_TL, JS_PE_TD, JS_PE_TU,
    JS_T_TL, JS_T_TD, JS_T_TU, JS_T_J, JS_T_CS,
    JS_T_CE0,                                       // 5 slots each: nodes x-2 .. x+2
    JS_T_PE0 = JS_T_CE0 + 5, JS_T_PS0 = JS_T_PE0 + 5,
    JS_TX_L = JS_T_PS0 + 5, JS_TX_D, JS_TX_U, JS_TX_I,
#if PLB_SEI
    JS_JS_T, JS_T_JS, JS_T_FILM,                    // temperature = true with aging = :SEI
#endif
    JS_KAP,                                         // staging only: D_s(T_x)/Rp^2 of the node
    JS_CS0,
#else
    JS_CS0,                      // 100 particle-block slots r*NR+c
#endif
    JS_CTRL_PS0 = JS_CS0 + NR * NR, JS_CTRL_PSN, JS_CTRL_I,
    JS_CTRL_T, JS_CTRL_TX,       // dT control row: entries on this lane's T / collector T (thermal variant)
    JS_CTRL_EPS, JS_CTRL_EPE,    // eta_p control row: Phi_s / Phi_e of the first anode node
    JS_COUNT
};

#ifndef PLB_K1_UNROLL
#define PLB_K1_UNROLL 8      // independent recipe -> value -> store chains in flight in the write loop
#endif
#ifndef PLB_K1_CTAS
#define PLB_K1_CTAS (PLB_WIDE ? 1 : (PLB_TH ? 2 : 3))
#endif
#ifndef PLB_K1_WARPS
#define PLB_K1_WARPS (PLB_WIDE ? 2 : 4)
#endif
#ifndef PLB_SEI_TMA
#define PLB_SEI_TMA 0          // A/B knob: the TMA-staged K1 in the SEI family (needs PLB_K1_WARPS=3 to keep three CTAs per SM)
#endif
constexpr int K1_WARPS = PLB_K1_WARPS;            // systems (lane groups) per CTA
constexpr int K1_NSTAGE = JS_CS0 + 7;   // lane-computed slots: 0..JS_CS0-1, then the seven control-row slots
constexpr int K1_SRC_MAX = (WIDE ? (TH ? (SEI ? 7168 : 6656) : 4864) : (TH ? (SEI ? 3328 : 3072) : 2304))         // >= nnz of every built variant
                           + (NR - 10) * 10 * (WIDE ? 64 : 32) + (PLB_SPECTRAL ? 27 * (WIDE ? 64 : 32) : 0);   // N_r siblings: 9 (+1 thermal) more block entries per radial node and particle
__host__ __device__ constexpr int k1_stage_slot(int js) { return js < JS_CS0 ? js : JS_CS0 + (js - JS_CTRL_PS0); }

// pitch of the value table: one double of padding per slot row.  Consecutive CSC entries of a column come from
// different slots of the SAME lane (the rows of column c_e[x] are slots CE_U/CE_D/CE_L/J_CE/PC_* of lanes x-1..x+1),
// which at pitch 32 all sit in one bank pair: ncu counted 23.6 M shared-memory bank conflicts per launch in the gather
constexpr int K1_PITCH = LW + 1;
struct alignas(16) K1Warp {
    double S[K1_NSTAGE][K1_PITCH];   // lane-computed Jacobian entries
#if PLB_TH
    double MCs[NR * NR];       // particle stencil coefficients (contiguous with S: one value table)
#else
    double PB[2][NR * NR];     // the particle blocks D_s/Rp^2 * M - gamma*I of the two electrodes (contiguous with S)
#endif
    WarpConst C;
};
constexpr size_t K1_MC_BYTES = TH ? 0 : sizeof(double) * NR * NR;   // CTA-shared copy of the stencil (non-thermal)
constexpr size_t K1_SMEM = XCH_BYTES_PER_GROUP * K1_WARPS + sizeof(K1Warp) * K1_WARPS + sizeof(int) * K1_SRC_MAX + K1_MC_BYTES;

// K1: one warp evaluates F and the CSC values of dF/dY + gamma dF/dY' of one system at a time.
// HBM traffic per system is exactly the algorithmic 8*(3N + n_theta + nnz) bytes: Y, Y', theta rows
// are read once (lane-mapped, L1-coalesced), res and nzval rows are written once with lane-consecutive
// 8-byte stores.  Most of nzval is the constant particle stencil scaled by D_s/Rp^2 (minus gamma on
// the diagonal).  Isothermal families: D_s is one number per electrode, so the two 10x10 blocks are staged once
// per system (200 fma) and the write loop is a pure gather  nzval[p] = table[recipe[p]]  -- coalesced,
// branch-free, ~5 instructions per 32 entries.  Thermal family: D_s(T) differs per node, the block entries are
// formed in the write loop from the recipe's flag bits.
//   recipe: index into the warp's value table (lane-computed: slot*K1_PITCH+lane; particle entries:
//   K1_NSTAGE*K1_PITCH + electrode*NR*NR + r*NR+c); thermal: bits 0-15 index (K1_NSTAGE*K1_PITCH + r*NR+c),
//   16 particle-block entry, 18 diagonal, 19-23 node
template <int CHEM>
__global__ void __launch_bounds__(K1_WARPS * LW, PLB_K1_CTAS) k_resjac(const __grid_constant__ ResJacArgs a) {
    extern __shared__ __align__(16) unsigned char k1_raw0[];
    unsigned char* k1_raw = k1_raw0 + XCH_BYTES_PER_GROUP * K1_WARPS;   // wide: the exchange scratch comes first
    K1Warp* ws = reinterpret_cast<K1Warp*>(k1_raw);
    int* src_s = reinterpret_cast<int*>(k1_raw + sizeof(K1Warp) * K1_WARPS);
    for (int i = threadIdx.x; i < a.nnz; i += blockDim.x) src_s[i] = a.src[i];
#if PLB_TH
    {
        K1Warp& w0 = ws[grp_id()];
        for (int i = grp_lane(); i < NR * NR; i += LW) w0.MCs[i] = laws::MC[i / NR][i % NR];
    }
#else
    double* mc_s = reinterpret_cast<double*>(k1_raw + sizeof(K1Warp) * K1_WARPS + sizeof(int) * K1_SRC_MAX);
    for (int i = threadIdx.x; i < NR * NR; i += blockDim.x) mc_s[i] = laws::MC[i / NR][i % NR];
#endif
    __syncthreads();
    const int warp = grp_id(), lane = grp_lane();
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
        y.T = 0.0; y.Tx = 0.0; yp.T = 0.0; yp.Tx = 0.0;
        y.js = 0.0; y.film = 0.0; y.soh = 0.0; yp.js = 0.0; yp.film = 0.0; yp.soh = 0.0;
        if (SEI) {
            const int k = ro.x - (m.Np + m.Ns);
            if (ro.sec == 2) { y.js = gY[m.off_js + k]; y.film = gY[m.off_film + k]; yp.film = gYP[m.off_film + k]; }
            y.soh = gY[m.off_SOH]; yp.soh = gYP[m.off_SOH];
        }
        if (TH) {
            if (ro.act) { y.T = gY[m.off_T + m.Na + ro.x]; yp.T = gYP[m.off_T + m.Na + ro.x]; }
            if (ro.ix >= 0) { y.Tx = gY[m.off_T + ro.ix]; yp.Tx = gYP[m.off_T + ro.ix]; }
        }
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
            if (TH) {
                if (ro.act) gR[m.off_T + m.Na + ro.x] = res.T;
                if (ro.ix >= 0) gR[m.off_T + ro.ix] = res.Tx;
            }
            if (SEI) {
                const int k = ro.x - (m.Np + m.Ns);
                if (ro.sec == 2) { gR[m.off_js + k] = res.js; gR[m.off_film + k] = res.film; }
                if (lane == 0) gR[m.off_SOH] = res.soh;
            }
            if (lane == 0) gR[m.off_I] = ctrl.res;
        }
        if (a.nzval) {
            const double g = a.gamma ? a.gamma[sys] : 0.0;
            w.S[JS_CE_L][lane] = J.ceL; w.S[JS_CE_D][lane] = J.ceD - g; w.S[JS_CE_U][lane] = J.ceU; w.S[JS_CE_J][lane] = J.ce_j;
            w.S[JS_J_CS][lane] = J.j_cs; w.S[JS_J_CE][lane] = J.j_ce; w.S[JS_J_PE][lane] = J.j_pe; w.S[JS_J_PS][lane] = J.j_ps;
#if PLB_SEI
            w.S[JS_J_J][lane] = J.j_j;
            w.S[JS_J_FILM][lane] = J.j_film;
            w.S[JS_JS_PS][lane] = J.js_ps; w.S[JS_JS_PE][lane] = J.js_pe; w.S[JS_JS_J][lane] = J.js_j;
            w.S[JS_JS_JS][lane] = J.js_js; w.S[JS_JS_FILM][lane] = J.js_film; w.S[JS_JS_I][lane] = J.js_I;
            w.S[JS_FILM_JS][lane] = J.film_js; w.S[JS_FILM_D][lane] = -g;
            w.S[JS_SOH_JS][lane] = J.soh_js; w.S[JS_SOH_D][lane] = -g;
            w.S[JS_CE_JS][lane] = J.ce_j; w.S[JS_PE_JS][lane] = J.pe_j; w.S[JS_PS_JS][lane] = J.ps_j;
#else
            w.S[JS_J_J][lane] = -1.0;
#endif
            w.S[JS_PE_L][lane] = J.peL; w.S[JS_PE_D][lane] = J.peD; w.S[JS_PE_U][lane] = J.peU;
            w.S[JS_PC_L][lane] = J.pcL; w.S[JS_PC_D][lane] = J.pcD; w.S[JS_PC_U][lane] = J.pcU; w.S[JS_PE_J][lane] = J.pe_j;
            w.S[JS_PS_L][lane] = J.psL; w.S[JS_PS_D][lane] = J.psD; w.S[JS_PS_U][lane] = J.psU; w.S[JS_PS_J][lane] = J.ps_j;
            w.S[JS_PS_I][lane] = J.ps_I;
#if PLB_SPECTRAL
#pragma unroll
            for (int r = 0; r < NR; r++) w.S[JS_CS_J + r][lane] = J.cs_j * laws::GJ[r];
#else
            w.S[JS_CS_J][lane] = J.cs_j;
#endif
#if PLB_TH
#pragma unroll
            for (int r = 0; r < NR; r++) w.S[JS_CS_T0 + r][lane] = J.csT[r];
            w.S[JS_J_T][lane] = J.j_T;
            w.S[JS_PE_TL][lane] = J.peTL; w.S[JS_PE_TD][lane] = J.peTD; w.S[JS_PE_TU][lane] = J.peTU;
            w.S[JS_T_TL][lane] = J.T_TL; w.S[JS_T_TD][lane] = J.T_TD - g; w.S[JS_T_TU][lane] = J.T_TU;
            w.S[JS_T_J][lane] = J.T_j; w.S[JS_T_CS][lane] = J.T_cs;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                w.S[JS_T_CE0 + k][lane] = J.T_ce[k]; w.S[JS_T_PE0 + k][lane] = J.T_pe[k]; w.S[JS_T_PS0 + k][lane] = J.T_ps[k];
            }
            w.S[JS_TX_L][lane] = J.Tx_L; w.S[JS_TX_D][lane] = J.Tx_D - g; w.S[JS_TX_U][lane] = J.Tx_U; w.S[JS_TX_I][lane] = J.Tx_I;
#if PLB_SEI
            w.S[JS_JS_T][lane] = J.js_T; w.S[JS_T_JS][lane] = J.T_js; w.S[JS_T_FILM][lane] = J.T_film;
#endif
            w.S[JS_KAP][lane] = J.kap;
            if (lane == 0) w.S[JS_KAP][LW] = 1.0;
#endif
            w.S[k1_stage_slot(JS_CTRL_PS0)][lane] = ctrl.g_ps0;
            w.S[k1_stage_slot(JS_CTRL_PSN)][lane] = ctrl.g_psN;
            w.S[k1_stage_slot(JS_CTRL_I)][lane] = ctrl.g_I;
            w.S[k1_stage_slot(JS_CTRL_T)][lane] = g * ctrl.gTn;
            w.S[k1_stage_slot(JS_CTRL_TX)][lane] = g * ctrl.gTx;
            w.S[k1_stage_slot(JS_CTRL_EPS)][lane] = ctrl.g_eta;
            w.S[k1_stage_slot(JS_CTRL_EPE)][lane] = -ctrl.g_eta;
#if !PLB_TH
            {   // the two particle blocks of this system
                const double kap_p = w.C.sec[SC_kap][0], kap_n = w.C.sec[SC_kap][2];
#pragma unroll
                for (int i = lane; i < NR * NR; i += LW) {
                    const double mc = mc_s[i];
                    const double gd = (i % (NR + 1) == 0) ? g : 0.0;
                    w.PB[0][i] = fma(kap_p, mc, -gd);
                    w.PB[1][i] = fma(kap_n, mc, -gd);
                }
            }
#endif
            grp_sync();
            const double* tab = &w.S[0][0];
            double* __restrict__ gN = a.nzval + (size_t)sys * a.nnz;
#if PLB_TH
#pragma unroll 4
            for (int p = lane; p < a.nnz; p += LW) {
                // every entry is kap * t - gd: lane-computed entries point kap at the 1.0 kept in the padding
                // element of the KAP row (fma(1, t, -0) == t), particle entries at their node's D_s(T)/Rp^2
                const int rc = src_s[p];
                const double t = tab[rc & 0xffff];
                const double kap = tab[JS_KAP * K1_PITCH + ((rc >> 19) & 127)];
                const double gd = (rc & (1 << 18)) ? g : 0.0;
                gN[p] = fma(kap, t, -gd);
            }
#else
            _Pragma(PLB_STR(unroll PLB_K1_UNROLL))
            for (int p = lane; p < a.nnz; p += LW) gN[p] = tab[src_s[p]];
#endif
        }
        grp_sync();
    }
}

#if !PLB_WIDE && !PLB_TH
// =================================================================================================
// K1, TMA-staged (isothermal and SEI families; the default whenever the caller's arrays are 16-byte aligned)
// =================================================================================================
// What the round-1 ncu source page of k_resjac showed (profiles/k1_r1i + profiles/ncu_by_region.py): 18 % of the
// kernel's time waiting on the theta row (a dependent global load at the head of setup_consts), 20 % in the serial
// per-section constants, 7 % on the lane-mapped Y / Y' loads (ten 80-byte-strided LDG per lane: 4x more L1 sectors
// than bytes), 19 % in the gather loop.  Here:
//   * the Y, Y', theta rows of a warp's NEXT system are bulk-copied (cp.async.bulk global -> shared, completion on
//     an mbarrier) while the current system is evaluated and written: no lane ever waits on an input row, and L1
//     sees no strided traffic at all;
//   * the rows are read from shared memory with 128-bit accesses (a lane's ten radial values: five LDS.128);
//   * the residual row is transposed through shared memory and leaves with 128-bit coalesced stores;
//   * the gather loop moves two CSC entries per lane and iteration (one 32-bit recipe pair, two table reads, one
//     STG.128).
// Rows have odd length (N = 301, nnz = 2139, 35 parameters): row `sys` of a [B][len] array starts 16-byte aligned
// only when sys*len is even.  A bulk copy needs 16-byte aligned addresses and sizes, so each row is split into its
// largest aligned even-length part (bulk) and one edge element (fetched by a lane into a register at issue time);
// in shared memory the row is shifted by `sh = (sys*len) & 1` doubles, which keeps the (global, shared) pairs
// aligned alike.
// nzval is NOT staged for a bulk store: a 17 KB row per warp would leave 6 warps per SM (measured in round 1: the
// kernel is issue/latency bound, occupancy is what it lives on); its stores are already full 128-byte lines.
constexpr int K1_VSIN = VS + 2;                       // staged row: N + shift, even
constexpr int K1_THIN = 64;
struct alignas(16) K1In {
    double Y[K1_VSIN], YP[K1_VSIN], TH[K1_THIN];
    unsigned long long bar;
    unsigned long long pad;
};
constexpr size_t K1T_SMEM = sizeof(K1Warp) * K1_WARPS + sizeof(K1In) * K1_WARPS + sizeof(uint32_t) * K1_SRC_MAX + K1_MC_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "K1_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra K1_DONE;\n"
        "bra K1_WAIT;\n"
        "K1_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the aligned part of row `sys` of a [B][len] array: shift sh, first bulk element, number of bulk elements,
// index of the edge element (-1: none)
struct RowSplit { int sh, cnt, edge; };
__device__ __forceinline__ RowSplit row_split(long long sys, int len) {
    RowSplit r;
    r.sh = (int)((sys * len) & 1);
    r.cnt = (len - r.sh) & ~1;
    r.edge = (r.sh + r.cnt < len) ? len - 1 : (r.sh ? 0 : -1);
    return r;
}

template <int CHEM>
__global__ void __launch_bounds__(K1_WARPS * LW, PLB_K1_CTAS) k_resjac_tma(const __grid_constant__ ResJacArgs a) {
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Warp* ws = reinterpret_cast<K1Warp*>(k1_raw);
    K1In* ins = reinterpret_cast<K1In*>(k1_raw + sizeof(K1Warp) * K1_WARPS);
    // recipe pairs: tabA[q] = (src[2q], src[2q+1]) for rows that start aligned, tabB[q] = (src[2q+1], src[2q+2]) for
    // rows that start on an odd double; 16 bits each (table indices < 2^16)
    uint32_t* tabA = reinterpret_cast<uint32_t*>(k1_raw + sizeof(K1Warp) * K1_WARPS + sizeof(K1In) * K1_WARPS);
    uint32_t* tabB = tabA + K1_SRC_MAX / 2;
    double* mc_s = reinterpret_cast<double*>(tabA + K1_SRC_MAX);
    const int nnz = a.nnz;
    for (int q = threadIdx.x; 2 * q < nnz; q += blockDim.x) {
        const uint32_t s0 = a.src[2 * q], s1 = 2 * q + 1 < nnz ? a.src[2 * q + 1] : s0, s2 = 2 * q + 2 < nnz ? a.src[2 * q + 2] : s1;
        tabA[q] = s0 | (s1 << 16);
        tabB[q] = s1 | (s2 << 16);
    }
    for (int i = threadIdx.x; i < NR * NR; i += blockDim.x) mc_s[i] = laws::MC[i / NR][i % NR];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    K1Warp& w = ws[warp];
    K1In& in = ins[warp];
    if (lane == 0) mbar_init(&in.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const ModelDesc& m = a.m;
    const int N = m.N_tot, NT = m.ntheta;
    const LaneRole ro = make_role(m, lane);
    const int nwarps = gridDim.x * K1_WARPS;
    const uint32_t src_first = a.src[0], src_last = a.src[nnz - 1];
    // ---- issue the bulk copies (and the edge / scalar fetches) of one system into this warp's staging buffer ----
    double edge = 0.0, g_next = 0.0, v_next = a.value;
    int edge_slot = -1;
    auto issue = [&](int sys) {
        const RowSplit rY = row_split(sys, N), rT = row_split(sys, NT);
        const double* gY = a.Y + (size_t)sys * N;
        const double* gYP = a.YP + (size_t)sys * N;
        const double* gT = a.theta + (size_t)sys * m.theta_stride;
        if (lane == 0) {
            // the generic-proxy reads of the buffer (previous system) are complete: order them before the async writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&in.bar, (uint32_t)((2 * rY.cnt + rT.cnt) * sizeof(double)));
            bulk_g2s(in.Y + 2 * rY.sh, gY + rY.sh, rY.cnt * sizeof(double), &in.bar);
            bulk_g2s(in.YP + 2 * rY.sh, gYP + rY.sh, rY.cnt * sizeof(double), &in.bar);
            bulk_g2s(in.TH + 2 * rT.sh, gT + rT.sh, rT.cnt * sizeof(double), &in.bar);
        }
        // edge elements: lanes 0, 1, 2 hold the one of Y, Y', theta in a register until the buffer is consumed next
        edge_slot = -1;
        if (lane < 2 && rY.edge >= 0) { edge = (lane == 0 ? gY : gYP)[rY.edge]; edge_slot = rY.edge + rY.sh; }
        if (lane == 2 && rT.edge >= 0) { edge = gT[rT.edge]; edge_slot = rT.edge + rT.sh; }
        g_next = a.gamma ? a.gamma[sys] : 0.0;
        v_next = a.values ? a.values[sys] : a.value;
    };
    int sys = blockIdx.x * K1_WARPS + warp;
    uint32_t phase = 0;
    if (sys < a.B) issue(sys);
    for (; sys < a.B; sys += nwarps) {
        const int shY = (int)(((long long)sys * N) & 1), shT = (int)(((long long)sys * NT) & 1);
        // a row shifted by sh lives at [sh .. sh+len): its bulk part starts at element sh, i.e. at slot 2*sh
        // (sh = 1: element 1 at slot 2, 16-byte aligned; element 0 at slot 1)
        if (edge_slot >= 0) (lane == 0 ? in.Y : (lane == 1 ? in.YP : in.TH))[edge_slot] = edge;
        const double g = g_next, value = v_next;
        __syncwarp();
        mbar_wait(&in.bar, phase);
        phase ^= 1;
        const double* __restrict__ sY = in.Y + shY;
        const double* __restrict__ sYP = in.YP + shY;
        LaneVec y, yp, res;
        y.ce = ro.act ? sY[ro.x] : 0.0; yp.ce = ro.act ? sYP[ro.x] : 0.0;
        y.pe = ro.act ? sY[m.off_pe + ro.x] : 0.0; yp.pe = 0.0;
        if (ro.elec) {
            const double* pc = sY + m.off_cs + ro.e * NR;
            const double* pp = sYP + m.off_cs + ro.e * NR;
            if (((m.off_cs + shY) & 1) == 0) {
#pragma unroll
                for (int k = 0; k < NR / 2; k++) {
                    const double2 t = reinterpret_cast<const double2*>(pc)[k], u = reinterpret_cast<const double2*>(pp)[k];
                    y.cs[2 * k] = t.x; y.cs[2 * k + 1] = t.y; yp.cs[2 * k] = u.x; yp.cs[2 * k + 1] = u.y;
                }
            } else {
#pragma unroll
                for (int r = 0; r < NR; r++) { y.cs[r] = pc[r]; yp.cs[r] = pp[r]; }
            }
            y.j = sY[m.off_j + ro.e]; y.ps = sY[m.off_ps + ro.e];
        } else {
#pragma unroll
            for (int r = 0; r < NR; r++) { y.cs[r] = 0.0; yp.cs[r] = 0.0; }
            y.j = 0.0; y.ps = 0.0;
        }
        yp.j = 0.0; yp.ps = 0.0;
        y.T = 0.0; y.Tx = 0.0; yp.T = 0.0; yp.Tx = 0.0;
        y.js = 0.0; y.film = 0.0; y.soh = 0.0; yp.js = 0.0; yp.film = 0.0; yp.soh = 0.0;
        if (SEI) {
            const int k = ro.x - (m.Np + m.Ns);
            if (ro.sec == 2) { y.js = sY[m.off_js + k]; y.film = sY[m.off_film + k]; yp.film = sYP[m.off_film + k]; }
            y.soh = sY[m.off_SOH]; yp.soh = sYP[m.off_SOH];
        }
        const double Iapp = sY[m.off_I];
        setup_consts(m, in.TH + shT, w.C, lane);
        __syncwarp();
        // every lane has taken what it needs from the staging buffer: refill it with the warp's next system
        if (sys + nwarps < a.B) issue(sys + nwarps);
        LaneJac J;
        CtrlRow ctrl;
        if (a.nzval) lane_eval<CHEM, true>(m, w.C, ro, y, yp, Iapp, a.method, value, res, ctrl, J);
        else lane_eval<CHEM, false>(m, w.C, ro, y, yp, Iapp, a.method, value, res, ctrl, J);
        if (a.res) {
            // transpose through shared memory (the value table is free until the Jacobian is staged), then
            // 128-bit coalesced stores; slot = element + shift keeps the global pairs 16-byte aligned
            double* row = &w.S[0][0] + shY;
            if (ro.act) { row[ro.x] = res.ce; row[m.off_pe + ro.x] = res.pe; }
            if (ro.elec) {
                double* pc = row + m.off_cs + ro.e * NR;
                if (((m.off_cs + shY) & 1) == 0) {
#pragma unroll
                    for (int k = 0; k < NR / 2; k++) reinterpret_cast<double2*>(pc)[k] = make_double2(res.cs[2 * k], res.cs[2 * k + 1]);
                } else {
#pragma unroll
                    for (int r = 0; r < NR; r++) pc[r] = res.cs[r];
                }
                row[m.off_j + ro.e] = res.j;
                row[m.off_ps + ro.e] = res.ps;
            }
            if (SEI) {
                const int k = ro.x - (m.Np + m.Ns);
                if (ro.sec == 2) { row[m.off_js + k] = res.js; row[m.off_film + k] = res.film; }
                if (lane == 0) row[m.off_SOH] = res.soh;
            }
            if (lane == 0) row[m.off_I] = ctrl.res;
            __syncwarp();
            double* __restrict__ gR = a.res + (size_t)sys * N;
            const double2* row2 = reinterpret_cast<const double2*>(&w.S[0][0]);
            // pair q = slots (2q, 2q+1) = elements (2q - sh, 2q + 1 - sh)
#pragma unroll 2
            for (int q = lane; 2 * q - shY + 1 < N; q += 32) {
                const int j0 = 2 * q - shY;
                const double2 v = row2[q];
                if (j0 >= 0) *reinterpret_cast<double2*>(gR + j0) = v;
                else gR[0] = v.y;
            }
            if (lane == 0 && ((N - shY) & 1)) gR[N - 1] = row[N - 1];
            __syncwarp();
        }
        if (a.nzval) {
            w.S[JS_CE_L][lane] = J.ceL; w.S[JS_CE_D][lane] = J.ceD - g; w.S[JS_CE_U][lane] = J.ceU; w.S[JS_CE_J][lane] = J.ce_j;
            w.S[JS_J_CS][lane] = J.j_cs; w.S[JS_J_CE][lane] = J.j_ce; w.S[JS_J_PE][lane] = J.j_pe; w.S[JS_J_PS][lane] = J.j_ps;
#if PLB_SEI
            w.S[JS_J_J][lane] = J.j_j;
            w.S[JS_J_FILM][lane] = J.j_film;
            w.S[JS_JS_PS][lane] = J.js_ps; w.S[JS_JS_PE][lane] = J.js_pe; w.S[JS_JS_J][lane] = J.js_j;
            w.S[JS_JS_JS][lane] = J.js_js; w.S[JS_JS_FILM][lane] = J.js_film; w.S[JS_JS_I][lane] = J.js_I;
            w.S[JS_FILM_JS][lane] = J.film_js; w.S[JS_FILM_D][lane] = -g;
            w.S[JS_SOH_JS][lane] = J.soh_js; w.S[JS_SOH_D][lane] = -g;
            w.S[JS_CE_JS][lane] = J.ce_j; w.S[JS_PE_JS][lane] = J.pe_j; w.S[JS_PS_JS][lane] = J.ps_j;
#else
            w.S[JS_J_J][lane] = -1.0;
#endif
            w.S[JS_PE_L][lane] = J.peL; w.S[JS_PE_D][lane] = J.peD; w.S[JS_PE_U][lane] = J.peU;
            w.S[JS_PC_L][lane] = J.pcL; w.S[JS_PC_D][lane] = J.pcD; w.S[JS_PC_U][lane] = J.pcU; w.S[JS_PE_J][lane] = J.pe_j;
            w.S[JS_PS_L][lane] = J.psL; w.S[JS_PS_D][lane] = J.psD; w.S[JS_PS_U][lane] = J.psU; w.S[JS_PS_J][lane] = J.ps_j;
            w.S[JS_PS_I][lane] = J.ps_I;
#if PLB_SPECTRAL
#pragma unroll
            for (int r = 0; r < NR; r++) w.S[JS_CS_J + r][lane] = J.cs_j * laws::GJ[r];
#else
            w.S[JS_CS_J][lane] = J.cs_j;
#endif
            w.S[k1_stage_slot(JS_CTRL_PS0)][lane] = ctrl.g_ps0;
            w.S[k1_stage_slot(JS_CTRL_PSN)][lane] = ctrl.g_psN;
            w.S[k1_stage_slot(JS_CTRL_I)][lane] = ctrl.g_I;
            w.S[k1_stage_slot(JS_CTRL_T)][lane] = g * ctrl.gTn;
            w.S[k1_stage_slot(JS_CTRL_TX)][lane] = g * ctrl.gTx;
            w.S[k1_stage_slot(JS_CTRL_EPS)][lane] = ctrl.g_eta;
            w.S[k1_stage_slot(JS_CTRL_EPE)][lane] = -ctrl.g_eta;
            {   // the two particle blocks of this system
                const double kap_p = w.C.sec[SC_kap][0], kap_n = w.C.sec[SC_kap][2];
#pragma unroll
                for (int i = lane; i < NR * NR; i += LW) {
                    const double mc = mc_s[i];
                    const double gd = (i % (NR + 1) == 0) ? g : 0.0;
                    w.PB[0][i] = fma(kap_p, mc, -gd);
                    w.PB[1][i] = fma(kap_n, mc, -gd);
                }
            }
            __syncwarp();
            const double* tab = &w.S[0][0];
            double* __restrict__ gN = a.nzval + (size_t)sys * nnz;
            const int par = (int)(((long long)sys * nnz) & 1);
            const uint32_t* __restrict__ tp = par ? tabB : tabA;
            const int nfull = (nnz - par) >> 1;
            double2* __restrict__ g2 = reinterpret_cast<double2*>(gN + par);
            _Pragma(PLB_STR(unroll PLB_K1_UNROLL))
            for (int q = lane; q < nfull; q += 32) {
                const uint32_t u = tp[q];
                g2[q] = make_double2(tab[u & 0xffffu], tab[u >> 16]);
            }
            if (lane == 0 && par) gN[0] = tab[src_first];
            if (lane == 1 && ((nnz - par) & 1)) gN[nnz - 1] = tab[src_last];
        }
        __syncwarp();
    }
}
#endif

// structural enumeration of the Jacobian: (row, col) in the reference layout for slot/lane
bool slot_rc(const ModelDesc& m, int method, int slot, int lane, int& row, int& col) {
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
#if !PLB_SPECTRAL
        case JS_CS_J: row = cs(NR - 1); col = r_j; return elec;
#endif
        case JS_CTRL_PS0: row = I; col = m.off_ps; return lane == 0 && (method == METHOD_V || method == METHOD_P);
        case JS_CTRL_PSN: row = I; col = m.off_ps + m.Ne - 1; return lane == Nx - 1 && (method == METHOD_V || method == METHOD_P);
        case JS_CTRL_I: row = I; col = I; return lane == 0 && (method == METHOD_I || method == METHOD_P);
        case JS_CTRL_EPS: row = I; col = r_ps; return method == METHOD_ETA && x == Np + Ns;
        case JS_CTRL_EPE: row = I; col = r_pe; return method == METHOD_ETA && x == Np + Ns;
        default: break;
    }
#if PLB_SPECTRAL
    if (slot >= JS_CS_J && slot < JS_CS_J + NR) { row = cs(slot - JS_CS_J); col = r_j; return elec; }
#endif
    if (slot >= JS_CS0 && slot < JS_CS0 + NR * NR) {
        const int r = (slot - JS_CS0) / NR, c = (slot - JS_CS0) % NR;
        row = cs(r); col = cs(c);
        return elec && (laws::mc_mask(r) & (1u << c));
    }
#if PLB_SEI
    {
        const int k = x - (Np + Ns);
        const int r_js = m.off_js + k, r_film = m.off_film + k, r_soh = m.off_SOH;
        switch (slot) {
            case JS_J_FILM: row = r_j; col = r_film; return isn;
            case JS_JS_PS: row = r_js; col = r_ps; return isn;
            case JS_JS_PE: row = r_js; col = r_pe; return isn;
            case JS_JS_J: row = r_js; col = r_j; return isn;
            case JS_JS_JS: row = r_js; col = r_js; return isn;
            case JS_JS_FILM: row = r_js; col = r_film; return isn;
            case JS_JS_I: row = r_js; col = I; return isn;
            case JS_FILM_JS: row = r_film; col = r_js; return isn;
            case JS_FILM_D: row = r_film; col = r_film; return isn;
            case JS_SOH_JS: row = r_soh; col = r_js; return isn;
            case JS_SOH_D: row = r_soh; col = r_soh; return lane == 0;
            case JS_CE_JS: row = r_ce; col = r_js; return isn;
            case JS_PE_JS: row = r_pe; col = r_js; return isn && !last;
            case JS_PS_JS: row = r_ps; col = r_js; return isn;
            default: break;
        }
    }
#endif
#if PLB_TH
    const int rT = m.off_T + m.Na + x;
    const bool cha = lane < m.Na, chz = lane >= Nx - m.Nz;
    const int kx = cha ? lane : lane - (Nx - m.Nz);
    const int rX = cha ? m.off_T + lane : m.off_T + m.Na + Nx + kx;
    if (slot >= JS_CS_T0 && slot < JS_CS_T0 + NR) { row = cs(slot - JS_CS_T0); col = rT; return elec; }
    switch (slot) {
        case JS_J_T: row = r_j; col = rT; return elec;
        case JS_PE_TL: row = r_pe; col = rT - 1; return x > 0 && !last;
        case JS_PE_TD: row = r_pe; col = rT; return !last;
        case JS_PE_TU: row = r_pe; col = rT + 1; return !last;
        case JS_T_TL: row = rT; col = rT - 1; return true;     // node 0: last node of the positive collector
        case JS_T_TD: row = rT; col = rT; return true;
        case JS_T_TU: row = rT; col = rT + 1; return true;     // node Nx-1: first node of the negative collector
        case JS_T_J: row = rT; col = r_j; return elec;
        case JS_T_CS: row = rT; col = cs(NR - 1); return elec;
        case JS_TX_L: row = rX; col = rX - 1; return (cha && kx > 0) || chz;
        case JS_TX_D: row = rX; col = rX; return cha || chz;
        case JS_TX_U: row = rX; col = rX + 1; return cha || (chz && kx < m.Nz - 1);
        case JS_TX_I: row = rX; col = I; return cha || chz;
#if PLB_SEI
        case JS_JS_T: row = m.off_js + x - (Np + Ns); col = rT; return isn;
        case JS_T_JS: row = rT; col = m.off_js + x - (Np + Ns); return isn;
        case JS_T_FILM: row = rT; col = m.off_film + x - (Np + Ns); return isn;
#endif
        case JS_CTRL_T: row = I; col = rT; return method == METHOD_DT;
        case JS_CTRL_TX: row = I; col = rX; return method == METHOD_DT && (cha || chz);
        default: break;
    }
    // stencils of thermal_derivatives: one-sided (own node and two inward) at the ends, central elsewhere
    auto stencil = [&](int d, bool lo_end, bool hi_end, bool own) -> bool {
        if (lo_end) return d >= 0;
        if (hi_end) return d <= 0;
        return d == -1 || d == 1 || (d == 0 && own);
    };
    if (slot >= JS_T_CE0 && slot < JS_T_CE0 + 5) {
        const int d = slot - JS_T_CE0 - 2;
        row = rT; col = x + d;
        return stencil(d, x == 0, last, true);               // own node: K_eff(c_e) and dc_e/c_e
    }
    if (slot >= JS_T_PE0 && slot < JS_T_PE0 + 5) {
        const int d = slot - JS_T_PE0 - 2;
        row = rT; col = r_pe + d;
        return stencil(d, x == 0, last, elec);               // own node only through eta
    }
    if (slot >= JS_T_PS0 && slot < JS_T_PS0 + 5) {
        const int d = slot - JS_T_PS0 - 2;
        row = rT; col = r_ps + d;
        return elec && stencil(d, first_e, last_e, true);    // own node through eta
    }
#endif
    return false;
}

// where K1 takes the value of the CSC position fed by (slot, lane) from
int slot_recipe(const ModelDesc& m, int slot, int lane) {
    if (slot >= JS_CS0 && slot < JS_CS0 + NR * NR) {
        const int rr = (slot - JS_CS0) / NR, cc = (slot - JS_CS0) % NR;
        const int el = lane >= m.Np + m.Ns ? 1 : 0;
#if PLB_TH
        (void)el;
        return (K1_NSTAGE * K1_PITCH + rr * NR + cc) | (1 << 16) | ((rr == cc ? 1 : 0) << 18) | (lane << 19);
#else
        return K1_NSTAGE * K1_PITCH + el * NR * NR + rr * NR + cc;
#endif
    }
#if PLB_TH
    return (k1_stage_slot(slot) * K1_PITCH + lane) | (LW << 19);      // kap slot LW: the constant 1.0
#else
    return k1_stage_slot(slot) * K1_PITCH + lane;
#endif
}

// =================================================================================================
// initial_guess!, newtons_method!, linear solve, simulate
// =================================================================================================
// Systems (lane groups) in flight per CTA, and CTAs per SM.  One barrier per tick (plb_tick.cuh) keeps the systems of a CTA on
// one instruction stream; with only that barrier left, one large CTA per SM is best for the 32-lane families (measured, iso,
// round 1: 6x1 229 k sims/s, 3x2 213 k, 2x3 200 k).  How many systems fit is a matter of what lives in shared memory
// (plb_integrator.cuh: PLB_NGLOBAL; plb_device.cuh: the factored blocks in the global workspace).
// Final geometry, each entry measured against its neighbours (profiles/ab_r3_*geometry*.txt):
//   family   groups x CTAs   history vectors in global memory      sims/s per segment, start of the sitting -> final
//   iso          8 x 1       none                                   331 k -> 426 k   (register-bound at 8: 230 x 256 threads)
//   sei          7 x 1       none                                   258 k -> 368 k   (6 systems before)
//   th           8 x 1       phi_4, phi_5                           136 k -> 206 k   (5 before)
//   thsei        5 x 1       phi_5                                   92 k -> 145 k   (3 before)
//   wide         1 x 4       phi_5                                  137 k -> 158 k   (3 CTAs before)
//   wsei         4 x 1       phi_5   (one tick barrier for all)     115 k -> 164 k   (1 x 3 before)
//   wth          1 x 3       phi_5                                   52 k -> 65.8 k  (2 CTAs before)
//   wthsei       1 x 2       none                                  28.7 k -> 53.7 k  (1 before)
#ifndef PLB_SIM_WARPS
#if PLB_NR == 10
#define PLB_SIM_WARPS (PLB_WIDE ? ((PLB_SEI && !PLB_TH) ? 4 : 1) : (PLB_TH ? (PLB_SEI ? 5 : 8) : (PLB_SEI ? 7 : 8)))
#else   // N_r = 12 / 14 siblings: longer vectors and larger particle inverses per system
#define PLB_SIM_WARPS (PLB_WIDE ? 1 : (PLB_TH ? (PLB_SEI ? 4 : 7) : (PLB_SEI ? (PLB_NR == 12 ? 6 : 5) : (PLB_NR == 12 ? 7 : 6))))
#endif
#endif
#ifndef PLB_SIM_CTAS
#define PLB_SIM_CTAS (PLB_WIDE ? (PLB_TH ? (PLB_SEI ? 2 : 3) : (PLB_SEI ? 1 : 4)) : 1)
#endif
constexpr int SIM_WARPS = PLB_SIM_WARPS;
constexpr int SIM_CTAS = PLB_SIM_CTAS;
// optional: the warp-uniform integrator state (SimState, plb_tick.cuh) in shared memory instead of registers
#ifndef PLB_STATE_SMEM
#define PLB_STATE_SMEM 1          // measured (iso): 229 k sims/s in registers/local memory, 246 k in shared memory
#endif
constexpr size_t STATE_BYTES = PLB_STATE_SMEM ? 480 : 0;           // per physical warp
constexpr size_t STATE_OFFSET = XCH_BYTES_PER_GROUP * SIM_WARPS + sizeof(WarpSmem) * SIM_WARPS;
constexpr size_t SIM_SMEM = STATE_OFFSET + STATE_BYTES * SIM_WARPS * (LW / 32);
// doubles per system slot of the global workspace: the history vectors parked there + the factored blocks of the linear solve
constexpr int GWS_PER_SLOT = (NGLOBAL > 0 ? NGLOBAL : 1) * VS + FA_GLOBAL;
static_assert(SIM_SMEM <= 227 * 1024, "the integrator's shared memory exceeds one SM: lower PLB_SIM_WARPS for this family");

__device__ __forceinline__ WarpWS make_ws(unsigned char* smem_raw, double* gws, int warp) {
    WarpSmem& sm = reinterpret_cast<WarpSmem*>(smem_raw + XCH_BYTES_PER_GROUP * SIM_WARPS)[warp];
    double* g = gws + ((size_t)blockIdx.x * SIM_WARPS + warp) * (size_t)GWS_PER_SLOT;
#if (PLB_TH && PLB_TH_BLOCKS_GLOBAL) || (!PLB_TH && PLB_BLOCKS_GLOBAL)
    sm.Fa.blk = g + (size_t)(NGLOBAL > 0 ? NGLOBAL : 1) * VS;     // (every lane of the group writes the same pointer)
    grp_sync();
#endif
    return WarpWS{g, &sm.svec[0][0], sm.C, sm.Fa, sm.K};
}

}  // namespace PLB_NS
}  // namespace plb

#include "plb_tick.cuh"

namespace plb {
namespace PLB_NS {

// initial_guess of one node (states_definition.jl:80-121): c_s from the SOC, c_e0, T0, Phi_s = U(c_s*)
template <int CHEM>
__device__ __forceinline__ void initial_lane(const ModelDesc& m, const WarpConst& C, const LaneRole& ro,
                                             double SOC, LaneVec& y0) {
    const double* th = C.theta;
    const double csp = th[TF_c_max_p] * (SOC * (th[TF_theta_max_p] - th[TF_theta_min_p]) + th[TF_theta_min_p]);
    const double csn = th[TF_c_max_n] * (SOC * (th[TF_theta_max_n] - th[TF_theta_min_n]) + th[TF_theta_min_n]);
    y0.ce = th[TF_c_e0]; y0.j = 0.0; y0.pe = 0.0; y0.ps = 0.0;
    y0.T = th[TF_T0]; y0.Tx = th[TF_T0];
    y0.js = 0.0; y0.film = 0.0; y0.soh = 1.0;
    const double cs0 = ro.sec == 0 ? csp : csn;
#pragma unroll
    for (int r = 0; r < NR; r++) y0.cs[r] = cs0;
    if (ro.elec) {
        const double thx = cs0 * C.sec[SC_inv_cmax][ro.sec];
        double U, dU, dUdT = 0.0, ddUdT = 0.0;
        if (CHEM == CHEM_LCO) {
            if (ro.sec == 0) laws::OCV_LCO(thx, U, dU, dUdT, ddUdT);
            else laws::OCV_LiC6(thx, sqrt(fmax(thx, 1e-4)), U, dU, dUdT, ddUdT);
            if (C.g[GC_dUdT_on] != 0.0) U += dUdT * (C.g[GC_T] - kTref);
        } else if (CHEM == CHEM_LGM) {
            if (ro.sec == 0) laws::OCV_NMC811(thx, U, dU);
            else laws::OCV_LiC6_LGM50(thx, U, dU);
        } else {
            if (ro.sec == 0) laws::OCV_NMC(thx, U, dU);
            else laws::OCV_LiC6_NMC(thx, U, dU);
        }
        y0.ps = U;
    }
}

template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * LW, SIM_CTAS) k_initguess(const __grid_constant__ AuxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = grp_id(), lane = grp_lane();
    WarpWS w = make_ws(smem_raw, a.gws, warp);
    const ModelDesc& m = a.m;
    const LaneRole ro = make_role(m, lane);
    const int N = m.N_tot;
    for (int sys = blockIdx.x * SIM_WARPS + warp; sys < a.B; sys += gridDim.x * SIM_WARPS) {
        setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
        LaneVec y0;
        initial_lane<CHEM>(m, w.C, ro, a.soc[sys], y0);
        store_lane(m, ro, w.v(V_PHI0), y0, 0.0, lane);
        grp_sync();
        for (int i = lane; i < N; i += LW) a.Y[(size_t)sys * N + ref_index(m, i)] = w.v(V_PHI0)[i];
        grp_sync();
    }
}

template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * LW, SIM_CTAS) k_newton(const __grid_constant__ AuxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = grp_id(), lane = grp_lane();
    WarpWS w = make_ws(smem_raw, a.gws, warp);
    const ModelDesc& m = a.m;
    const LaneRole ro = make_role(m, lane);
    const int N = m.N_tot;
    for (int sys = blockIdx.x * SIM_WARPS + warp; sys < a.B; sys += gridDim.x * SIM_WARPS) {
        setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
        for (int i = lane; i < N; i += LW) w.v(V_PHI0)[i] = a.Y[(size_t)sys * N + ref_index(m, i)];
        grp_sync();
        RunCtl rc;
        rc.method = a.method; rc.dc = 0;
        rc.value = a.values ? a.values[sys] : a.value;
        int nres = 0, njac = 0;
        const int it = newton_init<CHEM>(m, w, ro, rc, a.o, w.v(V_PHI0), w.v(V_PHI1), lane, nres, njac);
        for (int i = lane; i < N; i += LW) {
            a.Y[(size_t)sys * N + ref_index(m, i)] = w.v(V_PHI0)[i];
            a.YP[(size_t)sys * N + ref_index(m, i)] = it > 0 ? w.v(V_PHI1)[i] : 0.0;
        }
        if (lane == 0 && a.status) a.status[sys] = it;
        grp_sync();
    }
}

// x = (dF/dY + gamma dF/dY')^{-1} rhs at the state (Y, Y'): Jacobian evaluation, structured
// factorisation and one solve -- what KLU does for IDA (model_evaluation.jl:265-271)
template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * LW, SIM_CTAS) k_linsolve(const __grid_constant__ AuxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = grp_id(), lane = grp_lane();
    WarpWS w = make_ws(smem_raw, a.gws, warp);
    const ModelDesc& m = a.m;
    const LaneRole ro = make_role(m, lane);
    const int N = m.N_tot;
    for (int sys = blockIdx.x * SIM_WARPS + warp; sys < a.B; sys += gridDim.x * SIM_WARPS) {
        setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
        for (int i = lane; i < N; i += LW) {
            w.v(V_PHI0)[i] = a.Y[(size_t)sys * N + ref_index(m, i)];
            w.v(V_PHI1)[i] = a.YP[(size_t)sys * N + ref_index(m, i)];
            w.v(V_EE)[i] = a.rhs[(size_t)sys * N + ref_index(m, i)];
        }
        grp_sync();
        LaneVec y, yp, res, g;
        double Iy, Ip, gI;
        load_lane(m, ro, w.v(V_PHI0), y, Iy);
        load_lane(m, ro, w.v(V_PHI1), yp, Ip);
        load_lane(m, ro, w.v(V_EE), g, gI);
        LaneJac J;
        CtrlRow ctrl;
        lane_eval_ni<CHEM, true>(m, w.C, ro, y, yp, Iy, a.method, a.values ? a.values[sys] : a.value, res, ctrl, J);
        const double cj = a.gamma ? a.gamma[sys] : 0.0;
        warp_factor(m, ro, J, ctrl, cj, false, w.Fa, lane);
        const double dI = warp_solve(m, ro, w.Fa, false, g, gI, lane);
        grp_sync();
        store_lane(m, ro, w.v(V_EE), g, dI, lane);
        grp_sync();
        for (int i = lane; i < N; i += LW) a.x[(size_t)sys * N + ref_index(m, i)] = w.v(V_EE)[i];
        if (lane == 0 && a.status) a.status[sys] = (w.Fa.schur_inv == w.Fa.schur_inv) ? 0 : -1;
        grp_sync();
    }
}

template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * LW, SIM_CTAS) k_simulate(const __grid_constant__ SimArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // persistent CTAs; every warp pulls systems from a global queue (step counts vary ~1.5x across
    // a batch) and all warps of the CTA tick in lockstep through the heavy phases (plb_tick.cuh)
    simulate_cta<CHEM, false>(a, smem_raw);
}
// the same integrator with the optional features compiled in: tabulated inputs (run_function) and per-step state rows
template <int CHEM>
__global__ void __launch_bounds__(SIM_WARPS * LW, SIM_CTAS) k_simulate_ext(const __grid_constant__ SimArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    simulate_cta<CHEM, true>(a, smem_raw);
}

// =================================================================================================
// launchers (host)
// =================================================================================================
VariantInfo info() {
    VariantInfo v;
    v.sim_warps = SIM_WARPS; v.sim_ctas = SIM_CTAS; v.k1_warps = K1_WARPS; v.k1_ctas = PLB_K1_CTAS;
    v.sim_smem = SIM_SMEM; v.k1_smem = K1_SMEM; v.vs = VS; v.nglobal = NGLOBAL;
    v.n_slots = JS_COUNT; v.n_stage = K1_NSTAGE; v.k1_src_max = K1_SRC_MAX; v.lanes = LW;
    v.k1_tma = (!WIDE && !TH && (!SEI || PLB_SEI_TMA)) ? 1 : 0;
    v.nr = NR;
    v.gws_per_slot = GWS_PER_SLOT;
    return v;
}

#ifdef PLB_ONLY_CHEM
// sibling build of one more chemistry: only that instantiation exists here
#define PLB_LAUNCH(KERNEL, ARGS, GRID, BLOCK, SMEM, STREAM)                                                         \
    do {                                                                                                            \
        if ((ARGS).m.chem != PLB_ONLY_CHEM) return cudaErrorInvalidValue;                                           \
        cudaError_t e_ = cudaFuncSetAttribute(KERNEL<PLB_ONLY_CHEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)); \
        if (e_ != cudaSuccess) return e_;                                                                           \
        KERNEL<PLB_ONLY_CHEM><<<(GRID), (BLOCK), (SMEM), (STREAM)>>>(ARGS);                                         \
        return cudaGetLastError();                                                                                  \
    } while (0)
#else
#define PLB_LAUNCH(KERNEL, ARGS, GRID, BLOCK, SMEM, STREAM)                                                         \
    do {                                                                                                            \
        cudaError_t e_;                                                                                             \
        if ((ARGS).m.chem == CHEM_LCO) {                                                                            \
            e_ = cudaFuncSetAttribute(KERNEL<CHEM_LCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)); \
            if (e_ != cudaSuccess) return e_;                                                                       \
            KERNEL<CHEM_LCO><<<(GRID), (BLOCK), (SMEM), (STREAM)>>>(ARGS);                                          \
        } else {                                                                                                    \
            if (TH || SEI) return cudaErrorInvalidValue; /* NMC carries no thermal / aging parameters */           \
            e_ = cudaFuncSetAttribute(KERNEL<(PLB_TH || PLB_SEI) ? CHEM_LCO : CHEM_NMC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)); \
            if (e_ != cudaSuccess) return e_;                                                                       \
            KERNEL<(PLB_TH || PLB_SEI) ? CHEM_LCO : CHEM_NMC><<<(GRID), (BLOCK), (SMEM), (STREAM)>>>(ARGS);          \
        }                                                                                                           \
        return cudaGetLastError();                                                                                  \
    } while (0)
#endif

cudaError_t launch_resjac(const ResJacArgs& a, int grid, cudaStream_t s) {
#if !PLB_WIDE && !PLB_TH
    // (SEI: its larger value table leaves two CTAs per SM next to the staging buffers; the per-lane loads stay ahead: 0.65 vs 1.08 ms)
    if (a.use_tma) PLB_LAUNCH(k_resjac_tma, a, grid, K1_WARPS * LW, K1T_SMEM, s);
#endif
    PLB_LAUNCH(k_resjac, a, grid, K1_WARPS * LW, K1_SMEM, s);
}
cudaError_t launch_initguess(const AuxArgs& a, int grid, cudaStream_t s) { PLB_LAUNCH(k_initguess, a, grid, SIM_WARPS * LW, SIM_SMEM, s); }
cudaError_t launch_newton(const AuxArgs& a, int grid, cudaStream_t s) { PLB_LAUNCH(k_newton, a, grid, SIM_WARPS * LW, SIM_SMEM, s); }
cudaError_t launch_linsolve(const AuxArgs& a, int grid, cudaStream_t s) { PLB_LAUNCH(k_linsolve, a, grid, SIM_WARPS * LW, SIM_SMEM, s); }
cudaError_t launch_simulate(const SimArgs& a, int grid, cudaStream_t s) {
    if (a.tab_n || a.tr_Y || a.n_tstops || a.n_dense) PLB_LAUNCH(k_simulate_ext, a, grid, SIM_WARPS * LW, SIM_SMEM, s);
    PLB_LAUNCH(k_simulate, a, grid, SIM_WARPS * LW, SIM_SMEM, s);
}

}  // namespace PLB_NS
}  // namespace plb
