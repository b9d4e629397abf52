// plb_integrator.cuh -- fused per-warp DAE integrator: algebraic initialisation, variable-order
// variable-step BDF (fixed-leading-coefficient form, orders 1..5) with modified Newton, stop
// conditions with linear back-interpolation, output capture.
//
// Replaces, for one system per warp and without leaving the SM:
//   newtons_method!                 /root/reference/src/model_evaluation.jl:430-480
//   IDA (Sundials.jl `step!`)       call sites model_evaluation.jl:249-250, 262-280, 320
//                                   (third-party SUNDIALS 5.x IDA algorithm; decision logic restated
//                                   so that step/order sequences coincide with the reference's)
//   solve! / check_simulation_stop! / check_solve / interp_final_points!
//                                   model_evaluation.jl:312-382, checks.jl:1-249
//   set_vars! (t, V, I, SOC rows)   save_outputs.jl:11-40
#pragma once
#include <float.h>

#include "plb_device.cuh"

namespace plb {

struct Opts {
    double abstol, reltol, abstol_init, reltol_init;
    int maxiters, check_bounds, interp_final;
    int maxord, maxcor, maxnef, maxncf;   // Sundials.jl IDA(): 5, 3, 7, 10
};
struct Bounds {
    double V_max, V_min, SOC_max, SOC_min, T_max, c_s_n_max, I_max, I_min, eta_plating_min, c_e_min,
        dfilm_max;
};
struct Summary {
    double t_end, V_end, I_end, SOC_end;
    int flag, n_steps, n_res, n_jac, n_netf, n_ncfn, n_newton_init, reserved;
};

constexpr int FAIL_NEWTON_INIT = -1, FAIL_CONV = -2, FAIL_ERRTEST = -3, FAIL_MAXITERS = -4,
              FAIL_NONFINITE = -5, FAIL_INIT_BOUNDS = -6;

constexpr int VS = 304;   // vector stride in doubles (N_tot = 301 padded)
enum VecId { V_PHI0 = 0, V_PHI1, V_PHI2, V_PHI3, V_PHI4, V_PHI5, V_YPRED, V_YPPRED, V_EWT, V_EE, V_COUNT };

struct IdaCoef {
    double psi[6], alpha[6], beta[6], sigma[6], gamma[6];
    double cvals[6], dvals[6];     // interpolation weights at the return time
    double cprev[6];               // interpolation weights at the previous return time
};

// Number of workspace vectors that live in global memory (L2-resident, L1-cached) instead of
// shared memory: ids [0, PLB_NGLOBAL).  6 = the BDF history phi_0..phi_5; the predictor, weights and
// correction stay in shared memory.  Fewer shared-memory bytes per system = more systems per SM.
#ifndef PLB_NGLOBAL
#define PLB_NGLOBAL 0
#endif
constexpr int NGLOBAL = PLB_NGLOBAL;
constexpr int NSHARED = V_COUNT - NGLOBAL;

struct WarpSmem {
    double svec[NSHARED > 0 ? NSHARED : 1][VS];
    WarpConst C;
    WarpFactor Fa;
    IdaCoef K;
};

struct WarpWS {
    double* gbase;     // this warp's slice of the global workspace: NGLOBAL vectors of VS doubles
    double* sbase;     // shared-memory vectors
    WarpConst& C;
    WarpFactor& Fa;
    IdaCoef& K;
    __device__ __forceinline__ double* v(int id) const {
        return id < NGLOBAL ? gbase + id * VS : sbase + (id - NGLOBAL) * VS;
    }
};

// internal vector layout: c_e[Nx] | c_s radial-major [NR][Ne] | j[Ne] | Phi_e[Nx] | Phi_s[Ne] | I
// (the reference layout is particle-major for c_s; radial-major makes lane accesses conflict-free)
__device__ __forceinline__ void load_lane(const ModelDesc& m, const LaneRole& ro, const double* v,
                                          LaneVec& y, double& I) {
    y.ce = ro.act ? v[ro.x] : 0.0;
    y.pe = ro.act ? v[m.off_pe + ro.x] : 0.0;
    if (ro.elec) {
#pragma unroll
        for (int r = 0; r < NR; r++) y.cs[r] = v[m.off_cs + r * m.Ne + ro.e];
        y.j = v[m.off_j + ro.e];
        y.ps = v[m.off_ps + ro.e];
    } else {
#pragma unroll
        for (int r = 0; r < NR; r++) y.cs[r] = 0.0;
        y.j = 0.0; y.ps = 0.0;
    }
    I = v[m.off_I];
}
__device__ __forceinline__ void store_lane(const ModelDesc& m, const LaneRole& ro, double* v,
                                           const LaneVec& y, double I, int lane) {
    if (ro.act) { v[ro.x] = y.ce; v[m.off_pe + ro.x] = y.pe; }
    if (ro.elec) {
#pragma unroll
        for (int r = 0; r < NR; r++) v[m.off_cs + r * m.Ne + ro.e] = y.cs[r];
        v[m.off_j + ro.e] = y.j;
        v[m.off_ps + ro.e] = y.ps;
    }
    if (lane == 0) v[m.off_I] = I;
}
// reference (particle-major) layout <-> internal index
__device__ __forceinline__ int ref_index(const ModelDesc& m, int i) {
    if (i < m.off_cs || i >= m.off_j) return i;
    const int k = i - m.off_cs, r = k / m.Ne, e = k % m.Ne;
    return m.off_cs + e * NR + r;
}

__device__ __forceinline__ double wrms(const ModelDesc& m, const double* v, const double* w, int lane) {
    double s = 0.0;
    for (int i = lane; i < m.N_tot; i += 32) { const double p = v[i] * w[i]; s = fma(p, p, s); }
    return sqrt(warp_sum(s) / m.N_tot);
}

struct RunCtl {
    int method;
    double value;
};

// ------------------------------------------------------------------------------------------------
// newtons_method! -- model_evaluation.jl:430-480.  Y: vector in workspace (in/out), YP: vector (out)
// returns iterations (>0) or FAIL_NEWTON_INIT
// ------------------------------------------------------------------------------------------------
template <int CHEM>
__device__ __forceinline__ int newton_init(const ModelDesc& m, WarpWS& w, const LaneRole& ro,
                                           const RunCtl& rc, const Opts& o, double* Y, double* YP,
                                           int lane, int& n_res, int& n_jac) {
    LaneVec y, yp, res;
    LaneJac J;
    CtrlRow ctrl;
    double I;
    load_lane(m, ro, Y, y, I);
    yp.ce = 0.0; yp.j = 0.0; yp.pe = 0.0; yp.ps = 0.0;
#pragma unroll
    for (int r = 0; r < NR; r++) yp.cs[r] = 0.0;
    int iter;
    bool ok = false;
    for (iter = 1; iter <= 100; iter++) {
        lane_eval_ni<CHEM, true>(m, w.C, ro, y, yp, I, rc.method, rc.value, res, ctrl, J);   // R_alg, J_alg
        n_res++; n_jac++;
        warp_factor(m, ro, J, ctrl, 0.0, true, w.Fa, lane);
        const double dI = warp_solve(m, ro, w.Fa, true, res, ctrl.res, lane);
        // Y_new .-= factor \ res ; stop on the absolute 2-norm of the update (:451-454)
        double s = 0.0;
        if (ro.elec) { y.j -= res.j; y.ps -= res.ps; s += res.j * res.j + res.ps * res.ps; }
        if (ro.act) { y.pe -= res.pe; s += res.pe * res.pe; }
        I -= dI;
        s = warp_sum(s) + dI * dI;
        if (!(s == s) || isinf(s)) return FAIL_NEWTON_INIT;
        if (sqrt(s) < o.reltol_init) { ok = true; break; }
    }
    if (!ok) return FAIL_NEWTON_INIT;
    // R_diff(YP,t,Y,YP): YP_diff = rhs of the differential rows (:460)
    lane_eval_ni<CHEM, false>(m, w.C, ro, y, yp, I, rc.method, rc.value, res, ctrl, J);
    n_res++;
    LaneVec ypo;
    ypo.ce = res.ce;
#pragma unroll
    for (int r = 0; r < NR; r++) ypo.cs[r] = res.cs[r];
    // estimate dY_alg/dt (:462-477): Delta_t = max(10 reltol_init, sqrt(eps(c_e0)))
    {
        const double c0 = fabs(w.C.theta[TF_c_e0]);
        const double epsv = ::nextafter(c0, DBL_MAX) - c0;
        const double dt = fmax(10.0 * o.reltol_init, sqrt(epsv));
        LaneVec yn = y;
        yn.ce = y.ce + dt * ypo.ce;
#pragma unroll
        for (int r = 0; r < NR; r++) yn.cs[r] = y.cs[r] + dt * ypo.cs[r];
        lane_eval_ni<CHEM, false>(m, w.C, ro, yn, yp, I, rc.method, rc.value, res, ctrl, J);
        n_res++;
        const double dI = warp_solve(m, ro, w.Fa, true, res, ctrl.res, lane);
        ypo.j = -res.j / dt; ypo.pe = -res.pe / dt; ypo.ps = -res.ps / dt;
        store_lane(m, ro, Y, y, I, lane);
        store_lane(m, ro, YP, ypo, -dI / dt, lane);
    }
    __syncwarp();
    return iter;
}

// ------------------------------------------------------------------------------------------------
// IDA state (uniform across the warp, in registers) + helpers
// ------------------------------------------------------------------------------------------------
struct Ida {
    double tn, hh, hused, cj, cjlast, cjold, cjratio, ss, rr, hin;
    double tstop, tretlast;
    int kk, kused, knew, phase, ns, nst, tstopset;
    int nre, nje, netf, ncfn;
};

__device__ __forceinline__ void ewt_set(const ModelDesc& m, WarpWS& w, const Opts& o, int lane) {
    for (int i = lane; i < m.N_tot; i += 32)
        w.v(V_EWT)[i] = 1.0 / (o.reltol * fabs(w.v(V_PHI0)[i]) + o.abstol);
    __syncwarp();
}

// interpolation weights of IDAGetSolution at time t -> c[0..kord], d[0..kord-1]
__device__ __forceinline__ int getsol_weights(const Ida& M, const IdaCoef& K, double t, double* c, double* d) {
    int kord = M.kused; if (kord == 0) kord = 1;
    const double delt = t - M.tn;
    double cc = 1.0, dd = 0.0, gam = delt / K.psi[0];
    c[0] = cc;
    for (int j = 1; j <= kord; j++) {
        dd = dd * gam + cc / K.psi[j - 1];
        cc = cc * gam;
        gam = (delt + K.psi[j - 1]) / K.psi[j];
        c[j] = cc; d[j - 1] = dd;
    }
    return kord;
}

__device__ __forceinline__ double ida_set_coeffs(const ModelDesc& m, WarpWS& w, Ida& M, int lane) {
    IdaCoef& K = w.K;
    if (M.hh != M.hused || M.kk != M.kused) M.ns = 0;
    M.ns = min(M.ns + 1, M.kused + 2);
    __syncwarp();
    if (M.kk + 1 >= M.ns) {
        if (lane == 0) {
            K.beta[0] = 1.0; K.alpha[0] = 1.0; K.gamma[0] = 0.0; K.sigma[0] = 1.0;
            double temp1 = M.hh;
            for (int i = 1; i <= M.kk; i++) {
                const double temp2 = K.psi[i - 1];
                K.psi[i - 1] = temp1;
                K.beta[i] = K.beta[i - 1] * K.psi[i - 1] / temp2;
                temp1 = temp2 + M.hh;
                K.alpha[i] = M.hh / temp1;
                K.sigma[i] = i * K.sigma[i - 1] * K.alpha[i];
                K.gamma[i] = K.gamma[i - 1] + K.alpha[i - 1] / M.hh;
            }
            K.psi[M.kk] = temp1;
        }
    }
    __syncwarp();
    double alphas = 0.0, alpha0 = 0.0;
    for (int i = 0; i < M.kk; i++) { alphas -= 1.0 / (i + 1); alpha0 -= K.alpha[i]; }
    M.cjlast = M.cj;
    M.cj = -alphas / M.hh;
    double ck = fabs(K.alpha[M.kk] + alphas - alpha0);
    ck = fmax(ck, K.alpha[M.kk]);
    for (int k = M.ns; k <= M.kk; k++) {
        const double bk = K.beta[k];
        for (int i = lane; i < m.N_tot; i += 32) w.v(V_PHI0 + k)[i] *= bk;
    }
    M.tn += M.hh;
    __syncwarp();
    return ck;
}

// nonlinear solve: 0 ok, >0 recoverable failure
template <int CHEM>
__device__ __forceinline__ int ida_nls(const ModelDesc& m, WarpWS& w, const LaneRole& ro, const RunCtl& rc,
                                       const Opts& o, Ida& M, int lane) {
    const IdaCoef& K = w.K;
    bool callLSetup = false;
    if (M.nst == 0) { M.cjold = M.cj; M.ss = 20.0; callLSetup = true; }
    // predictor
    for (int i = lane; i < m.N_tot; i += 32) {
        double yv = 0.0, ypv = 0.0;
        for (int j = 0; j <= M.kk; j++) yv += w.v(V_PHI0 + j)[i];
        for (int j = 1; j <= M.kk; j++) ypv += K.gamma[j] * w.v(V_PHI0 + j)[i];
        w.v(V_YPRED)[i] = yv; w.v(V_YPPRED)[i] = ypv; w.v(V_EE)[i] = 0.0;
    }
    __syncwarp();
    M.cjratio = M.cj / M.cjold;
    {
        const double temp1 = (1.0 - 0.25) / (1.0 + 0.25), temp2 = 1.0 / temp1;
        if (M.cjratio < temp1 || M.cjratio > temp2) callLSetup = true;
        if (M.cj != M.cjlast) M.ss = 100.0;
    }
    LaneVec yp0, ypp0;     // predictor in registers (node mapping)
    double Ip0, Ipp0;
    load_lane(m, ro, w.v(V_YPRED), yp0, Ip0);
    load_lane(m, ro, w.v(V_YPPRED), ypp0, Ipp0);
    LaneVec ee;
    double eeI = 0.0;
    ee.ce = ee.j = ee.pe = ee.ps = 0.0;
#pragma unroll
    for (int r = 0; r < NR; r++) ee.cs[r] = 0.0;
    bool jcur = false;
    int retval = 0;
    double oldnrm = 0.0;
    LaneVec ewt;
    double ewtI;
    load_lane(m, ro, w.v(V_EWT), ewt, ewtI);
    for (;;) {
        LaneVec y, yp, res;
        LaneJac J;
        CtrlRow ctrl;
        y.ce = yp0.ce + ee.ce; y.j = yp0.j + ee.j; y.pe = yp0.pe + ee.pe; y.ps = yp0.ps + ee.ps;
        yp.ce = ypp0.ce + M.cj * ee.ce; yp.j = 0.0; yp.pe = 0.0; yp.ps = 0.0;
#pragma unroll
        for (int r = 0; r < NR; r++) { y.cs[r] = yp0.cs[r] + ee.cs[r]; yp.cs[r] = ypp0.cs[r] + M.cj * ee.cs[r]; }
        double Iy = Ip0 + eeI;
        if (callLSetup) {
            lane_eval_ni<CHEM, true>(m, w.C, ro, y, yp, Iy, rc.method, rc.value, res, ctrl, J);
            M.nre++; M.nje++;
            warp_factor(m, ro, J, ctrl, M.cj, false, w.Fa, lane);
            // a non-finite factorisation is a recoverable lsetup failure
            const double chk = w.Fa.schur_inv;
            if (!(chk == chk) || isinf(chk)) { retval = 1; break; }
            M.cjold = M.cj; M.cjratio = 1.0; M.ss = 20.0;
            jcur = true;
        } else {
            lane_eval_ni<CHEM, false>(m, w.C, ro, y, yp, Iy, rc.method, rc.value, res, ctrl, J);
            M.nre++;
        }
        int mi = 0;
        for (;;) {
            // delta = -J^{-1} F, scaled by 2/(1+cjratio) when the Jacobian is stale
            res.ce = -res.ce; res.j = -res.j; res.pe = -res.pe; res.ps = -res.ps;
#pragma unroll
            for (int r = 0; r < NR; r++) res.cs[r] = -res.cs[r];
            double dI = warp_solve(m, ro, w.Fa, false, res, -ctrl.res, lane);
            if (M.cjratio != 1.0) {
                const double sc = 2.0 / (1.0 + M.cjratio);
                res.ce *= sc; res.j *= sc; res.pe *= sc; res.ps *= sc; dI *= sc;
#pragma unroll
                for (int r = 0; r < NR; r++) res.cs[r] *= sc;
            }
            double s = 0.0;
            if (ro.act) {
                ee.ce += res.ce; ee.pe += res.pe;
                s = fma(res.ce * ewt.ce, res.ce * ewt.ce, s);
                s = fma(res.pe * ewt.pe, res.pe * ewt.pe, s);
            }
            if (ro.elec) {
                ee.j += res.j; ee.ps += res.ps;
                s = fma(res.j * ewt.j, res.j * ewt.j, s);
                s = fma(res.ps * ewt.ps, res.ps * ewt.ps, s);
#pragma unroll
                for (int r = 0; r < NR; r++) { ee.cs[r] += res.cs[r]; s = fma(res.cs[r] * ewt.cs[r], res.cs[r] * ewt.cs[r], s); }
            }
            eeI += dI;
            const double delnrm = sqrt((warp_sum(s) + (dI * ewtI) * (dI * ewtI)) / m.N_tot);
            // idaNlsConvTest
            retval = -99;
            if (mi == 0) {
                oldnrm = delnrm;
                if (delnrm <= 1e-4 * (1e-4 * 0.33)) retval = 0;
            } else {
                const double rate = pow(delnrm / oldnrm, 1.0 / mi);
                if (rate > 0.9) retval = 2;
                else M.ss = rate / (1.0 - rate);
            }
            if (retval == -99 && M.ss * delnrm <= 0.33) retval = 0;
            if (retval >= 0) break;
            mi++;
            if (mi >= o.maxcor) { retval = 2; break; }
            y.ce = yp0.ce + ee.ce; y.j = yp0.j + ee.j; y.pe = yp0.pe + ee.pe; y.ps = yp0.ps + ee.ps;
            yp.ce = ypp0.ce + M.cj * ee.ce;
#pragma unroll
            for (int r = 0; r < NR; r++) { y.cs[r] = yp0.cs[r] + ee.cs[r]; yp.cs[r] = ypp0.cs[r] + M.cj * ee.cs[r]; }
            Iy = Ip0 + eeI;
            lane_eval_ni<CHEM, false>(m, w.C, ro, y, yp, Iy, rc.method, rc.value, res, ctrl, J);
            M.nre++;
        }
        if (retval == 0) break;
        if (retval > 0 && !jcur) {
            callLSetup = true;
            ee.ce = ee.j = ee.pe = ee.ps = 0.0; eeI = 0.0;
#pragma unroll
            for (int r = 0; r < NR; r++) ee.cs[r] = 0.0;
            continue;
        }
        break;
    }
    store_lane(m, ro, w.v(V_EE), ee, eeI, lane);
    __syncwarp();
    return retval;
}

__device__ __forceinline__ bool ida_test_error(const ModelDesc& m, WarpWS& w, Ida& M, double ck,
                                               double& err_k, double& err_km1, int lane) {
    const IdaCoef& K = w.K;
    const double* ee = w.v(V_EE);
    const double* ewt = w.v(V_EWT);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const double* pk = w.v(V_PHI0 + M.kk);
    const double* pk1 = w.v(V_PHI0 + (M.kk > 0 ? M.kk - 1 : 0));
    for (int i = lane; i < m.N_tot; i += 32) {
        const double e = ee[i], wt = ewt[i];
        const double a = e * wt; s0 = fma(a, a, s0);
        const double d1 = (pk[i] + e); const double b = d1 * wt; s1 = fma(b, b, s1);
        const double d2 = (d1 + pk1[i]); const double c = d2 * wt; s2 = fma(c, c, s2);
    }
    const double enorm_k = sqrt(warp_sum(s0) / m.N_tot);
    err_k = K.sigma[M.kk] * enorm_k;
    const double terr_k = (M.kk + 1) * err_k;
    M.knew = M.kk;
    if (M.kk > 1) {
        err_km1 = K.sigma[M.kk - 1] * sqrt(warp_sum(s1) / m.N_tot);
        const double terr_km1 = M.kk * err_km1;
        if (M.kk > 2) {
            const double err_km2 = K.sigma[M.kk - 2] * sqrt(warp_sum(s2) / m.N_tot);
            const double terr_km2 = (M.kk - 1) * err_km2;
            if (fmax(terr_km1, terr_km2) <= terr_k) M.knew = M.kk - 1;
        } else {
            if (terr_km1 <= 0.5 * terr_k) M.knew = M.kk - 1;
        }
    }
    return ck * enorm_k > 1.0;
}

__device__ __forceinline__ void ida_restore(const ModelDesc& m, WarpWS& w, Ida& M, double saved_t, int lane) {
    IdaCoef& K = w.K;
    M.tn = saved_t;
    __syncwarp();
    if (lane == 0)
        for (int j = 1; j <= M.kk; j++) K.psi[j - 1] = K.psi[j] - M.hh;
    for (int j = M.ns; j <= M.kk; j++) {
        const double s = 1.0 / K.beta[j];
        for (int i = lane; i < m.N_tot; i += 32) w.v(V_PHI0 + j)[i] *= s;
    }
    __syncwarp();
}

__device__ __forceinline__ void ida_complete_step(const ModelDesc& m, WarpWS& w, const Opts& o, Ida& M,
                                                  double err_k, double err_km1, int lane) {
    const double* ee = w.v(V_EE);
    const double* ewt = w.v(V_EWT);
    M.nst++;
    const int kdiff = M.kk - M.kused;
    M.kused = M.kk;
    M.hused = M.hh;
    if (M.knew == M.kk - 1 || M.kk == o.maxord) M.phase = 1;
    if (M.phase == 0) {
        if (M.nst > 1) { M.kk++; M.hh = 2.0 * M.hh; }
    } else {
        int action = 0;   // 0 unset, 1 lower, 2 maintain, 3 raise
        double err_kp1 = 0.0, err_knew;
        if (M.knew == M.kk - 1) action = 1;
        else if (M.kk == o.maxord) action = 2;
        else if (M.kk + 1 >= M.ns || kdiff == 1) action = 2;
        if (action == 0) {
            double s = 0.0;
            const double* pk = w.v(V_PHI0 + M.kk + 1);
            for (int i = lane; i < m.N_tot; i += 32) { const double a = (ee[i] - pk[i]) * ewt[i]; s = fma(a, a, s); }
            const double enorm = sqrt(warp_sum(s) / m.N_tot);
            err_kp1 = enorm / (M.kk + 2);
            const double terr_k = (M.kk + 1) * err_k, terr_kp1 = (M.kk + 2) * err_kp1;
            if (M.kk == 1) {
                action = (terr_kp1 >= 0.5 * terr_k) ? 2 : 3;
            } else {
                const double terr_km1 = M.kk * err_km1;
                if (terr_km1 <= fmin(terr_k, terr_kp1)) action = 1;
                else if (terr_kp1 >= terr_k) action = 2;
                else action = 3;
            }
        }
        if (action == 3) { M.kk++; err_knew = err_kp1; }
        else if (action == 1) { M.kk--; err_knew = err_km1; }
        else err_knew = err_k;
        double hnew = M.hh;
        M.rr = pow(2.0 * err_knew + 1e-4, -1.0 / (M.kk + 1));
        if (M.rr >= 2.0) hnew = 2.0 * M.hh;
        else if (M.rr <= 1.0) { M.rr = fmax(0.5, fmin(0.9, M.rr)); hnew = M.hh * M.rr; }
        M.hh = hnew;
    }
    // phi updates
    for (int i = lane; i < m.N_tot; i += 32) {
        const double e = ee[i];
        if (M.kused < o.maxord) w.v(V_PHI0 + M.kused + 1)[i] = e;
        double acc = w.v(V_PHI0 + M.kused)[i] + e;
        w.v(V_PHI0 + M.kused)[i] = acc;
        for (int j = M.kused - 1; j >= 0; j--) { acc += w.v(V_PHI0 + j)[i]; w.v(V_PHI0 + j)[i] = acc; }
    }
    __syncwarp();
}

// IDAStep: 0 ok, <0 failure code
template <int CHEM>
__device__ __forceinline__ int ida_step(const ModelDesc& m, WarpWS& w, const LaneRole& ro, const RunCtl& rc,
                                        const Opts& o, Ida& M, int lane) {
    IdaCoef& K = w.K;
    const double saved_t = M.tn;
    int ncf = 0, nef = 0;
    if (M.nst == 0) {
        M.kk = 1; M.kused = 0; M.hused = 0.0; M.cj = 1.0 / M.hh; M.phase = 0; M.ns = 0;
        __syncwarp();
        if (lane == 0) K.psi[0] = M.hh;
        __syncwarp();
    }
    double err_k = 0.0, err_km1 = 0.0;
    for (;;) {
        const double ck = ida_set_coeffs(m, w, M, lane);
        const int nflag = ida_nls<CHEM>(m, w, ro, rc, o, M, lane);
        bool errfail = false;
        if (nflag == 0) errfail = ida_test_error(m, w, M, ck, err_k, err_km1, lane);
        if (nflag == 0 && !errfail) break;
        ida_restore(m, w, M, saved_t, lane);
        M.phase = 1;
        if (nflag != 0) {
            M.ncfn++; ncf++;
            M.rr = 0.25;
            M.hh *= M.rr;
            if (ncf >= o.maxncf) return FAIL_CONV;
        } else {
            nef++; M.netf++;
            if (nef == 1) {
                const double err_knew = (M.kk == M.knew) ? err_k : err_km1;
                M.kk = M.knew;
                M.rr = 0.9 * pow(2.0 * err_knew + 1e-4, -1.0 / (M.kk + 1));
                M.rr = fmax(0.25, fmin(0.9, M.rr));
                M.hh *= M.rr;
            } else if (nef == 2) {
                M.kk = M.knew; M.rr = 0.25; M.hh *= M.rr;
            } else if (nef < o.maxnef) {
                M.kk = 1; M.rr = 0.25; M.hh *= M.rr;
            } else return FAIL_ERRTEST;
        }
        if (M.nst == 0) {
            __syncwarp();
            if (lane == 0) K.psi[0] = M.hh;
            for (int i = lane; i < m.N_tot; i += 32) w.v(V_PHI1)[i] *= M.rr;
            __syncwarp();
        }
        if (!(fabs(M.hh) > 0.0) || isinf(M.hh)) return FAIL_CONV;
    }
    ida_complete_step(m, w, o, M, err_k, err_km1, lane);
    return 0;
}

// IDASolve(ONE_STEP) with a stop time.  Returns 0 / 1 (tstop return) / <0.  *tret = return time.
template <int CHEM>
__device__ __forceinline__ int ida_solve_one_step(const ModelDesc& m, WarpWS& w, const LaneRole& ro,
                                                  const RunCtl& rc, const Opts& o, Ida& M, double tout,
                                                  double& tret, int lane) {
    const double ur = DBL_EPSILON;
    if (M.nst == 0) {
        ewt_set(m, w, o, lane);
        const double tdist = fabs(tout - M.tn);
        M.hh = M.hin;
        if (M.hh == 0.0) {
            M.hh = 0.001 * tdist;
            const double ypnorm = wrms(m, w.v(V_PHI1), w.v(V_EWT), lane);
            if (ypnorm > 0.5 / M.hh) M.hh = 0.5 / ypnorm;
            if (tout < M.tn) M.hh = -M.hh;
        }
        if (M.tstopset) {
            if ((M.tn + M.hh - M.tstop) * M.hh > 0.0) M.hh = (M.tstop - M.tn) * (1.0 - 4.0 * ur);
        }
        M.kk = 0; M.kused = 0;
        for (int i = lane; i < m.N_tot; i += 32) w.v(V_PHI1)[i] *= M.hh;
        __syncwarp();
    } else {
        if (M.tstopset) {
            const double troundoff = 100.0 * ur * (fabs(M.tn) + fabs(M.hh));
            if (fabs(M.tn - M.tstop) <= troundoff) {
                tret = M.tretlast = M.tstop; M.tstopset = 0;
                return 1;
            }
            if ((M.tn + M.hh - M.tstop) * M.hh > 0.0) M.hh = (M.tstop - M.tn) * (1.0 - 4.0 * ur);
        }
        ewt_set(m, w, o, lane);
    }
    {
        const double nrm = wrms(m, w.v(V_PHI0), w.v(V_EWT), lane);
        if (ur * nrm > 1.0) { tret = M.tn; return FAIL_CONV; }
    }
    const int sflag = ida_step<CHEM>(m, w, ro, rc, o, M, lane);
    if (sflag != 0) { tret = M.tretlast = M.tn; return sflag; }
    if (M.tstopset) {
        const double troundoff = 100.0 * ur * (fabs(M.tn) + fabs(M.hh));
        if (fabs(M.tn - M.tstop) <= troundoff) {
            tret = M.tretlast = M.tstop; M.tstopset = 0;
            return 1;
        }
        if ((M.tn + M.hh - M.tstop) * M.hh > 0.0) M.hh = (M.tstop - M.tn) * (1.0 - 4.0 * ur);
    }
    tret = M.tretlast = M.tn;
    return 0;
}

// interpolated value / derivative of component i (internal index) with weights (c, d), order kord
__device__ __forceinline__ double interp_y(const WarpWS& w, const double* c, int kord, int i) {
    double y = 0.0;
    for (int j = 0; j <= kord; j++) y = fma(c[j], w.v(V_PHI0 + j)[i], y);
    return y;
}
__device__ __forceinline__ double interp_yp(const WarpWS& w, const double* d, int kord, int i) {
    double y = 0.0;
    for (int j = 1; j <= kord; j++) y = fma(d[j - 1], w.v(V_PHI0 + j)[i], y);
    return y;
}

struct PrevVals {   // boundary_stop_prev_values, structures.jl:174-184
    double frac, V, SOC, c_s_n, I, eta_plating, c_e_min;
};

// check_simulation_stop! -- checks.jl:1-224 (isothermal, no SEI: T and dfilm checks inactive)
__device__ __forceinline__ void check_stop(const ModelDesc& m, const WarpWS& w, const RunCtl& rc,
                                           const Opts& o, const Bounds& b, bool is_rest, double tf,
                                           PrevVals& pv, int& flag, double t, const double* c,
                                           const double* d, int kord, double SOC, int lane) {
    const double eps = t < 1.0 ? o.reltol : 0.0;
    if (t >= tf) { flag = 0; return; }
    if (!o.check_bounds || is_rest) return;
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    const double Ic = interp_y(w, c, kord, m.off_I), dIc = interp_yp(w, d, kord, m.off_I);
    if (rc.method != METHOD_I) {   // check_stop_I :31-54
        if ((Ic - b.I_max > eps) && dIc > 0) {
            const double tf_ = (pv.I - b.I_max) / (pv.I - Ic);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 7; }
        } else if ((b.I_min - Ic > eps) && dIc < 0) {
            const double tf_ = (pv.I - b.I_min) / (pv.I - Ic);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 8; }
        }
        pv.I = Ic;
    }
    if (rc.method != METHOD_V) {   // check_stop_V :56-80
        const double V = interp_y(w, c, kord, iP0) - interp_y(w, c, kord, iPN);
        const double dV = interp_yp(w, d, kord, iP0) - interp_yp(w, d, kord, iPN);
        if ((b.V_min - V > eps) && dV < 0) {
            const double tf_ = (pv.V - b.V_min) / (pv.V - V);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 1; }
        } else if ((V - b.V_max > eps) && dV > 0) {
            const double tf_ = (pv.V - b.V_max) / (pv.V - V);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 2; }
        }
        pv.V = V;
    }
    // check_stop_SOC :82-104
    if ((b.SOC_min - SOC > eps) && Ic < 0) {
        const double tf_ = (pv.SOC - b.SOC_min) / (pv.SOC - SOC);
        if (tf_ < pv.frac) { pv.frac = tf_; flag = 3; }
    } else if ((SOC - b.SOC_max > eps) && Ic > 0) {
        const double tf_ = (pv.SOC - b.SOC_max) / (pv.SOC - SOC);
        if (tf_ < pv.frac) { pv.frac = tf_; flag = 4; }
    }
    pv.SOC = SOC;
    // check_stop_c_s_surf :141-161
    if (b.c_s_n_max == b.c_s_n_max) {
        double mx = -INFINITY;
        for (int e = m.Np + lane; e < m.Ne; e += 32) mx = fmax(mx, interp_y(w, c, kord, m.off_cs + (NR - 1) * m.Ne + e));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, off));
        const double lim = b.c_s_n_max * w.C.theta[TF_c_max_n];
        if (Ic > 0 && mx - lim > eps) {
            const double tf_ = (pv.c_s_n - lim) / (pv.c_s_n - mx);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 6; }
        }
        pv.c_s_n = mx;
    }
    // check_stop_c_e :163-183
    if (b.c_e_min == b.c_e_min) {
        double mn = INFINITY;
        for (int i = lane; i < m.Nx; i += 32) mn = fmin(mn, interp_y(w, c, kord, i));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL, mn, off));
        if (b.c_e_min - mn > eps) {
            const double tf_ = (pv.c_e_min - b.c_e_min) / (pv.c_e_min - mn);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 9; }
        }
        pv.c_e_min = mn;
    }
    // check_stop_eta_plating :185-202
    if (b.eta_plating_min == b.eta_plating_min) {
        const int in_ = m.off_ps + m.Np, ie = m.off_pe + m.Np + m.Ns;
        const double ep = interp_y(w, c, kord, in_) - interp_y(w, c, kord, ie);
        const double dep = interp_yp(w, d, kord, in_) - interp_yp(w, d, kord, ie);
        if (b.eta_plating_min - ep > eps && dep < 0) {
            const double tf_ = (pv.eta_plating - b.eta_plating_min) / (pv.eta_plating - ep);
            if (tf_ < pv.frac) { pv.frac = tf_; flag = 11; }
        }
        pv.eta_plating = ep;
    }
}

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
    double *tr_t, *tr_V, *tr_I, *tr_SOC;
    int* tr_n;
    int* counter;
    double* gws;              // global workspace: [grid * warps_per_cta][NGLOBAL][VS]
};

// simulate / simulate! for one system -- model_evaluation.jl:10-97, 174-232, 312-382
template <int CHEM>
__device__ void simulate_system(const SimArgs& a, int sys, WarpWS& w, int lane) {
    const ModelDesc& m = a.m;
    const LaneRole ro = make_role(m, lane);
    const int N = m.N_tot;
    setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
    RunCtl rc;
    rc.method = a.method;
    rc.value = a.values ? a.values[sys] : a.value;
    Summary out;
    out.t_end = 0; out.V_end = 0; out.I_end = 0; out.SOC_end = 0; out.flag = -1; out.n_steps = 0;
    out.n_res = 0; out.n_jac = 0; out.n_netf = 0; out.n_ncfn = 0; out.n_newton_init = 0; out.reserved = 0;
    double* Y0 = w.v(V_PHI0);
    double* YP0 = w.v(V_PHI1);
    double SOC, t0;
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    double I_prev_state = 0.0;
    // ---- initialize_simulation! :174-232 ------------------------------------------------------------
    if (a.new_run) {
        // initial_guess! (states_definition.jl:80-121)
        SOC = a.soc0 ? a.soc0[sys] : 1.0;
        const double* th = w.C.theta;
        const double csp = th[TF_c_max_p] * (SOC * (th[TF_theta_max_p] - th[TF_theta_min_p]) + th[TF_theta_min_p]);
        const double csn = th[TF_c_max_n] * (SOC * (th[TF_theta_max_n] - th[TF_theta_min_n]) + th[TF_theta_min_n]);
        LaneVec y0;
        y0.ce = th[TF_c_e0]; y0.j = 0.0; y0.pe = 0.0; y0.ps = 0.0;
        const double cs0 = ro.sec == 0 ? csp : csn;
#pragma unroll
        for (int r = 0; r < NR; r++) y0.cs[r] = cs0;
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
            y0.ps = U;
        }
        store_lane(m, ro, Y0, y0, 0.0, lane);
        t0 = 0.0;
    } else {
        for (int i = lane; i < N; i += 32) Y0[i] = a.sY[(size_t)sys * N + ref_index(m, i)];
        SOC = a.sSOC[sys];
        t0 = ::nextafter(a.st[sys], DBL_MAX);   // initial_time, model_evaluation.jl:112
    }
    __syncwarp();
    I_prev_state = Y0[m.off_I];
    // initial_current! (input_methods.jl:11-107)
    {
        double Ig;
        const double V0 = Y0[iP0] - Y0[iPN];
        if (a.input_kind == 1) {              // :hold -- value from the previous state
            if (rc.method == METHOD_I) { rc.value = I_prev_state; Ig = I_prev_state; }
            else if (rc.method == METHOD_V) { rc.value = V0; Ig = V0; }   // sic: input_methods.jl:58
            else { rc.value = I_prev_state * w.C.g[GC_I1C] * V0; Ig = I_prev_state; }
        } else if (a.input_kind == 2) {       // :rest
            rc.value = 0.0; Ig = 0.0;
        } else if (rc.method == METHOD_I) Ig = rc.value;
        else if (rc.method == METHOD_V) {
            if (!a.new_run && I_prev_state != 0.0) Ig = I_prev_state;
            else Ig = rc.value > V0 ? 1.0 : -1.0;
        } else Ig = rc.value / (V0 * w.C.g[GC_I1C]);
        __syncwarp();
        if (lane == 0) Y0[m.off_I] = Ig;
        __syncwarp();
    }
    int nres = 0, njac = 0;
    const int nit = newton_init<CHEM>(m, w, ro, rc, a.o, Y0, YP0, lane, nres, njac);
    out.n_newton_init = nit;
    out.n_res = nres; out.n_jac = njac;
    bool done = false;
    if (nit < 0) { out.flag = FAIL_NEWTON_INIT; done = true; }
    if (!done && a.new_run) {   // check_initial_SOC, checks.jl:327-339
        const double I0 = Y0[m.off_I];
        if (I0 != 0 && ((SOC >= a.b.SOC_max && I0 > 0) || (SOC <= a.b.SOC_min && I0 < 0))) {
            out.flag = FAIL_INIT_BOUNDS; done = true;
        }
    }
    const size_t so = (size_t)sys * a.n_save_max;
    int nsave = 0;
    double t = 0.0, tprev = 0.0;
    double SOC_end = SOC, t_end = t0, V_end = 0.0, I_end = 0.0;
    if (!done) {
        Ida M;
        M.tn = 0.0; M.hh = 0.0; M.hused = 0.0; M.cj = 0.0; M.cjlast = 0.0; M.cjold = 0.0; M.cjratio = 1.0;
        M.ss = 20.0; M.rr = 0.0; M.hin = 0.0; M.tstop = 0.0; M.tretlast = 0.0;
        M.kk = 0; M.kused = 0; M.knew = 0; M.phase = 0; M.ns = 0; M.nst = 0; M.tstopset = 0;
        M.nre = nres; M.nje = njac; M.netf = 0; M.ncfn = 0;
        for (int k = 2; k < 6; k++)
            for (int i = lane; i < N; i += 32) w.v(V_PHI0 + k)[i] = 0.0;
        __syncwarp();
        // tstops: [1.0 if continuing; tf] (:288-310)
        double tstops[2];
        int ntstops = 0, itstop = 0;
        if (!a.new_run && 1.0 < a.tf) tstops[ntstops++] = 1.0;
        tstops[ntstops++] = a.tf;
        // first output row and t = 0 stop check (:225-230)
        double cw[6] = {1, 0, 0, 0, 0, 0}, dw[6] = {1, 0, 0, 0, 0, 0};   // y = phi0, yp = phi1 (= YP0 before scaling)
        double Vc = Y0[iP0] - Y0[iPN], Ic = Y0[m.off_I];
        double I_prev = Ic;
        if (lane == 0 && nsave < a.n_save_max) {
            if (a.tr_t) a.tr_t[so + nsave] = t0;
            if (a.tr_V) a.tr_V[so + nsave] = Vc;
            if (a.tr_I) a.tr_I[so + nsave] = Ic;
            if (a.tr_SOC) a.tr_SOC[so + nsave] = SOC;
        }
        nsave++;
        PrevVals pv;
        pv.frac = 1.0; pv.V = -1; pv.SOC = -1; pv.c_s_n = -1; pv.I = -1; pv.eta_plating = -1; pv.c_e_min = -1;
        int flag = -1;
        check_stop(m, w, rc, a.o, a.b, a.input_kind == 2, a.tf, pv, flag, 0.0, cw, dw, 1, SOC, lane);
        int iter = 1, hard = 0;
        bool retried = false;
        double tg_prev = t0;
        int kord = 1, kord_prev = 1;
        if (flag == -1) {
            // solve! (:312-333)
            for (;;) {
                tprev = t;
                // remember the interpolation weights of the previous return point for the final interp
                M.tstop = tstops[itstop]; M.tstopset = 1;
                double tret = t;
                const int fl = ida_solve_one_step<CHEM>(m, w, ro, rc, a.o, M, tstops[itstop], tret, lane);
                if (fl == 1 || tret >= tstops[itstop]) { if (itstop < ntstops - 1) itstop++; }
                t = tret;
                iter++;
                if (fl < 0 || t == tprev) {
                    // check_solve (checks.jl:226-237): one retry of the very first step with h = reltol
                    if (t == 0.0 && iter == 2 && !retried && M.nst == 0) {
                        retried = true;
                        // phi1 currently holds h_failed-scaled YP0: undo and restart the first step
                        const double sc = 1.0 / w.K.psi[0];
                        for (int i = lane; i < N; i += 32) w.v(V_PHI1)[i] *= sc;
                        __syncwarp();
                        M.hin = a.o.reltol;
                        continue;
                    }
                    hard = (fl == FAIL_ERRTEST) ? FAIL_ERRTEST : FAIL_CONV;
                    break;
                }
                kord_prev = kord;
                kord = getsol_weights(M, w.K, t, cw, dw);
                Ic = interp_y(w, cw, kord, m.off_I);
                Vc = interp_y(w, cw, kord, iP0) - interp_y(w, cw, kord, iPN);
                // set_vars! -> SOC trapezoid (save_outputs.jl:31, scalar_residual.jl:103-111)
                const double tg = t + t0;
                SOC = SOC + 0.5 * (tg - tg_prev) * (Ic + I_prev) / 3600.0;
                if (lane == 0 && nsave < a.n_save_max) {
                    if (a.tr_t) a.tr_t[so + nsave] = tg;
                    if (a.tr_V) a.tr_V[so + nsave] = Vc;
                    if (a.tr_I) a.tr_I[so + nsave] = Ic;
                    if (a.tr_SOC) a.tr_SOC[so + nsave] = SOC;
                }
                nsave++;
                check_stop(m, w, rc, a.o, a.b, a.input_kind == 2, a.tf, pv, flag, t, cw, dw, kord, SOC, lane);
                if (iter == a.o.maxiters) { hard = FAIL_MAXITERS; break; }
                if (!(Ic == Ic) || !(Vc == Vc) || isinf(Ic) || isinf(Vc)) { hard = FAIL_NONFINITE; break; }
                if (flag != -1) break;
                I_prev = Ic;
                tg_prev = tg;
            }
        }
        (void)kord_prev;
        // exit_simulation! (:335-382)
        t_end = t + t0;
        SOC_end = SOC;
        double fr = 1.0;
        bool do_interp = false;
        if (hard) flag = hard;
        else if (a.o.interp_final && flag != 0 && flag != -1 && t > 1.0) { do_interp = true; fr = pv.frac; }
        // weights of the previous return point (Y_prev = interpolant at tprev; exact at mesh points)
        double cp[6] = {1, 0, 0, 0, 0, 0}, dp[6];
        if (do_interp) getsol_weights(M, w.K, tprev, cp, dp);
        // final state (reference layout) and outputs
        double ps0 = 0.0, psN = 0.0, If = 0.0;
        for (int i = lane; i < N; i += 32) {
            const double yn = interp_y(w, cw, kord, i);
            double yf = yn;
            if (do_interp) { const double ypv = interp_y(w, cp, kord, i); yf = fr * (yn - ypv) + ypv; }
            a.sY[(size_t)sys * N + ref_index(m, i)] = yf;
            if (a.sYP) a.sYP[(size_t)sys * N + ref_index(m, i)] = interp_yp(w, dw, kord, i);
            if (i == iP0) ps0 = yf;
            if (i == iPN) psN = yf;
            if (i == m.off_I) If = yf;
        }
        ps0 = warp_sum(ps0); psN = warp_sum(psN); If = warp_sum(If);
        V_end = ps0 - psN; I_end = If;
        if (do_interp) {
            const double ti = fr * (t - tprev) + tprev;
            const double tgi = ti + t0, tgl = t + t0;
            SOC_end = SOC + 0.5 * (tgi - tgl) * (If + If) / 3600.0;
            t_end = tgi;
            if (lane == 0 && nsave - 1 < a.n_save_max && nsave >= 1) {
                if (a.tr_t) a.tr_t[so + nsave - 1] = tgi;
                if (a.tr_V) a.tr_V[so + nsave - 1] = V_end;
                if (a.tr_I) a.tr_I[so + nsave - 1] = I_end;
                if (a.tr_SOC) a.tr_SOC[so + nsave - 1] = SOC_end;
            }
        }
        out.flag = flag;
        out.n_res = M.nre; out.n_jac = M.nje; out.n_netf = M.netf; out.n_ncfn = M.ncfn;
    } else {
        // failed before integration: hand the initial state back
        for (int i = lane; i < N; i += 32) {
            a.sY[(size_t)sys * N + ref_index(m, i)] = Y0[i];
            if (a.sYP) a.sYP[(size_t)sys * N + ref_index(m, i)] = 0.0;
        }
        nsave = 1;
    }
    out.t_end = t_end; out.V_end = V_end; out.I_end = I_end; out.SOC_end = SOC_end;
    out.n_steps = nsave - 1;
    if (lane == 0) {
        a.out[sys] = out;
        a.sSOC[sys] = SOC_end;
        a.st[sys] = t_end;
        if (a.tr_n) a.tr_n[sys] = nsave < a.n_save_max ? nsave : a.n_save_max;
    }
    __syncwarp();
}

}  // namespace plb
