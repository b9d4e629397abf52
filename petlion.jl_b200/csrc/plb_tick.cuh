// plb_tick.cuh -- the fused integrator as a CTA-synchronous "tick" machine.
//
// Why: the per-system program (residual ~1.4k, factorisation ~1.7k, solve ~1.2k SASS instructions of
// straight-line FP64 code) does not fit the SM instruction cache.  With every warp at a different
// point of it, ncu showed sm__icc_request_hit_rate ~50 % and the GPC-level instruction cache at 70 % of
// peak: the kernel was instruction-fetch bound and adding warps made it slower.  Here every warp of
// a CTA still owns one system end to end, but all warps execute the three heavy phases in lockstep
// (one residual evaluation per warp per tick), so a single instruction stream per SM serves all of
// its systems.  The light per-state glue (predictor, error test, stop checks, ...) runs between the
// barriers, unaligned.
//
// One tick per warp = exactly one evaluation of F (optionally with the Jacobian + factorisation)
// followed by one linear solve.  The nested loops of the reference
//   solve! -> IDASolve(ONE_STEP) -> IDAStep retry loop -> IDANls lsetup retry -> Newton iteration
// (model_evaluation.jl:312-333 and SUNDIALS IDA) and of newtons_method! (model_evaluation.jl:430-480)
// are flattened into the states below; decision logic is unchanged from plb_integrator.cuh.
#pragma once
#include "plb_integrator.cuh"

namespace plb {
namespace PLB_NS {

// alignment barriers inside a tick (a performance heuristic only: they keep the warps of a CTA on one
// instruction stream through the factorisation and the solve; results do not depend on them)
#ifndef PLB_TICK_SYNC_JAC
#define PLB_TICK_SYNC_JAC 0
#endif
#ifndef PLB_TICK_SYNC_SOLVE
#define PLB_TICK_SYNC_SOLVE 0     // measured (iso, 6 warps x 1 CTA): 172 k sims/s with it, 225 k without
#endif
#ifndef PLB_TICK_SYNC_START
#define PLB_TICK_SYNC_START 1     // the one barrier per tick that keeps a CTA's systems on one instruction stream
#endif
// once-per-simulation code (fetch + parameter setup, first row, summary).  A/B knob: out of line (__noinline__) it
// leaves the tick loop spill-free, but the calls put the workspace descriptors on the stack: 272 k vs 298 k sims/s
#ifndef PLB_COLD
#define PLB_COLD __forceinline__
#endif
#ifndef PLB_TICK_OUTLINE
#define PLB_TICK_OUTLINE 0        // residual / factorisation / solve as out-of-line calls inside the tick (A/B knob)
#endif
#ifndef PLB_TICK_SYNC_EVERY
#define PLB_TICK_SYNC_EVERY 1     // barrier every n-th tick (A/B knob)
#endif
// Every evaluation of a tick runs the residual+Jacobian instantiation lane_eval<true>, whether its warp factorises in this tick or
// not: ONE copy of the heaviest phase in the tick's instruction stream, and the factorising warps no longer leave the common
// stream at its start (alone on their own copy they were the stragglers every other warp waited for at the next barrier).
// ~800 more instructions for the warps that do not need the Jacobian, and still (sims/s, one B200, 0 -> 1): iso 331 k -> 363 k,
// thermal 136 k -> 154 k, sei 258 k -> 286 k, thsei 92 k -> 104 k, wide SEI 115 k -> 127 k, NMC_LGM50 thermal 100 k -> 120 k
// (profiles/README.md).  A warp's arithmetic still depends on its own state only.
#ifndef PLB_TICK_ONE_EVAL
#define PLB_TICK_ONE_EVAL 1
#endif
// A Newton iteration that needs lsetup takes TWO ticks: evaluate + factorise in the first, evaluate + solve + glue in the second
// (the same evaluation again: same state, same instantiation, same bits -- results do not change).  The factorisation then
// runs NEXT TO the other warps' solve + glue instead of before the factorising warp's own, and no warp is left alone on the tail
// of a tick with every other warp of the CTA waiting for it at the next barrier.
#ifndef PLB_TICK_SPLIT_LSETUP
#define PLB_TICK_SPLIT_LSETUP 0
#endif
#ifndef PLB_TICK_SYNC_STEP
#define PLB_TICK_SYNC_STEP 0
#endif
#ifndef PLB_TICK_VOTE_JAC
#define PLB_TICK_VOTE_JAC 0       // CTA-wide vote "does any warp factorise in this tick?": 225 k with, 229 k without
#endif

enum TickState { ST_FETCH = 0, ST_INIT_ITER, ST_INIT_RDIFF, ST_INIT_DT, ST_NLS, ST_EXHAUSTED };
enum Pending { PEND_NONE = 0, PEND_RETURNED, PEND_BEGIN, PEND_FINISH };

struct SimState {
    int state, sys, pending;
    RunCtl rc;
    Ida M;
    double SOC, t0, t, tprev, tg_prev, I_prev;
    PrevVals pv;
    int flag, iter, hard, nsave, kord, retried;
    double tstop0, tstop1;
    int ntstops, itstop;
    int ni_iter, n_newton_init;
    // nonlinear solve (IDANls / Newton iteration)
    int callLSetup, jcur, mi;
    double oldnrm;
    // step attempt (IDAStep)
    double saved_t, ck, err_k, err_km1;
    int ncf, nef;
    // last solver return
    int ret_fl;
    double ret_t;
    double dt_init;
    // run_function (tabulated input): per-system scale, time of the algebraic initialisation in flight
    // (0, or t + reltol for a re-initialisation at a discontinuity: checks.jl:341-364)
    double scale, t_init;
    int reinit, n_reinit;
    // dense output: next requested time to fill, SOC before the last accepted step's trapezoid update
    int idense;
    double SOC_before;
};

// value of the tabulated input at local time t: last knot k with tab_t[k] <= t (right-continuous at a
// repeated knot = jump), linear to the next knot, constant outside the table
__device__ __noinline__ double table_eval(const double* __restrict__ tt, const double* __restrict__ vv, int n, double t) {
    int lo = -1, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(tt + mid) <= t) lo = mid; else hi = mid;
    }
    if (lo < 0) return __ldg(vv);
    if (lo == n - 1) return __ldg(vv + n - 1);
    const double t0 = __ldg(tt + lo), t1 = __ldg(tt + lo + 1), v0 = __ldg(vv + lo), v1 = __ldg(vv + lo + 1);
    return v0 + (v1 - v0) * ((t - t0) / (t1 - t0));
}
template <bool EXT>
__device__ __forceinline__ double cur_tstop(const SimArgs& a, const SimState& S) {
    return (EXT && a.n_tstops) ? __ldg(a.tstops + S.itstop) : (S.itstop == 0 ? S.tstop0 : S.tstop1);
}

// dense output (EXT kernels): fill the rows of the requested global times in (.., tg_limit] that are still open.
// The state at a requested time is the interpolant of the step that covers it (IDAGetSolution; `initial`: the
// start state phi_0); SOC follows the trapezoid of set_vars! (save_outputs.jl:31) from the last accepted point.
__device__ __noinline__ void dense_emit(const SimArgs& a, WarpWS& w, SimState& S, double tg_limit, bool initial, int lane) {
    const ModelDesc& m = a.m;
    const int N = m.N_tot;
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    __syncwarp();
    int k = S.idense;
    while (k < a.n_dense) {
        const double tg = __ldg(a.dense_t + k);
        if (!(tg <= tg_limit)) break;
        grp_sync();
        int kord = 1;
        if (lane == 0) {
            if (initial) { w.K.cprev[0] = 1.0; w.K.cprev[1] = 0.0; }
            else { double dp[6]; kord = getsol_weights(S.M, w.K, tg - S.t0, w.K.cprev, dp); }
        }
        kord = grp_bcast_int(kord, 0);
        grp_sync();
        const double* c = w.K.cprev;
        const double Id = interp_y(w, c, kord, m.off_I);
        const double Vd = interp_y(w, c, kord, iP0) - interp_y(w, c, kord, iPN);
        const size_t row = (size_t)S.sys * a.n_dense + k;
        if (a.dn_T) {
            const double Tw = weighted_T(m, w, c, kord, false, lane);
            if (lane == 0) a.dn_T[row] = Tw;
        }
        if (lane == 0) {
            if (a.dn_V) a.dn_V[row] = Vd;
            if (a.dn_I) a.dn_I[row] = Id;
            if (a.dn_SOC) a.dn_SOC[row] = initial ? S.SOC : S.SOC_before + 0.5 * (tg - S.tg_prev) * (Id + S.I_prev) / 3600.0;
        }
        if (a.dn_Y) {
            double* out = a.dn_Y + row * N;
#pragma unroll 1
            for (int i = lane; i < N; i += LW) out[ref_index(m, i)] = interp_y(w, c, kord, i);
        }
        k++;
    }
    __syncwarp();
    S.idense = k;
    if (lane == 0 && a.dn_n) a.dn_n[S.sys] = k;
    grp_sync();
}

// ---- pieces of ida_nls -------------------------------------------------------------------------------
__device__ __forceinline__ void nls_begin(const ModelDesc& m, WarpWS& w, SimState& S, int lane) {
    Ida& M = S.M;
    __syncwarp();   // S lives in shared memory and is updated by all lanes together: converge first
    S.callLSetup = 0;
    if (M.nst == 0) { M.cjold = M.cj; M.ss = 20.0; S.callLSetup = 1; }
    predict_pass(m, w, M, lane);
    M.cjratio = M.cj / M.cjold;
    const double temp1 = (1.0 - 0.25) / (1.0 + 0.25), temp2 = 1.0 / temp1;
    if (M.cjratio < temp1 || M.cjratio > temp2) S.callLSetup = 1;
    if (M.cj != M.cjlast) M.ss = 100.0;
    S.jcur = 0; S.mi = 0; S.oldnrm = 0.0;
}

// after the solve of one Newton iteration: returns -99 continue, 0 converged, >0 recoverable failure
__device__ __forceinline__ int nls_post(const ModelDesc& m, WarpWS& w, const LaneRole& ro, const Opts& o,
                                        SimState& S, LaneVec& d, double dI, int lane) {
    Ida& M = S.M;
    if (M.cjratio != 1.0) {
        const double sc = 2.0 / (1.0 + M.cjratio);
        d.ce *= sc; d.j *= sc; d.pe *= sc; d.ps *= sc; dI *= sc;
        if (TH) { d.T *= sc; d.Tx *= sc; }
        if (SEI) { d.js *= sc; d.film *= sc; d.soh *= sc; }
#pragma unroll
        for (int r = 0; r < NR; r++) d.cs[r] *= sc;
    }
    LaneVec ee, ewt;
    double eeI, ewtI;
    load_lane(m, ro, w.v(V_EE), ee, eeI);
    load_lane(m, ro, w.v(V_EWT), ewt, ewtI);
    double s = 0.0;
    if (ro.act) {
        ee.ce += d.ce; ee.pe += d.pe;
        s = fma(d.ce * ewt.ce, d.ce * ewt.ce, s);
        s = fma(d.pe * ewt.pe, d.pe * ewt.pe, s);
    }
    if (ro.elec) {
        ee.j += d.j; ee.ps += d.ps;
        s = fma(d.j * ewt.j, d.j * ewt.j, s);
        s = fma(d.ps * ewt.ps, d.ps * ewt.ps, s);
#pragma unroll
        for (int r = 0; r < NR; r++) { ee.cs[r] += d.cs[r]; s = fma(d.cs[r] * ewt.cs[r], d.cs[r] * ewt.cs[r], s); }
    }
    if (TH) {
        if (ro.act) { ee.T += d.T; s = fma(d.T * ewt.T, d.T * ewt.T, s); }
        if (ro.ix >= 0) { ee.Tx += d.Tx; s = fma(d.Tx * ewt.Tx, d.Tx * ewt.Tx, s); }
    }
    if (SEI) {
        if (ro.sec == 2) {
            ee.js += d.js; s = fma(d.js * ewt.js, d.js * ewt.js, s);
            ee.film += d.film; s = fma(d.film * ewt.film, d.film * ewt.film, s);
        }
        ee.soh += d.soh;
        if (lane == 0) s = fma(d.soh * ewt.soh, d.soh * ewt.soh, s);
    }
    eeI += dI;
    grp_sync();   // every lane has read the I slot of EE (load_lane) before lane 0 rewrites it (two-warp groups)
    store_lane(m, ro, w.v(V_EE), ee, eeI, lane);
    grp_sync();
    const double delnrm = sqrt((warp_sum(s) + (dI * ewtI) * (dI * ewtI)) / m.N_tot);
    __syncwarp();
    int retval = -99;
    if (S.mi == 0) {
        S.oldnrm = delnrm;
        if (delnrm <= 1e-4 * (1e-4 * 0.33)) retval = 0;
    } else {
        const double rate = pow(delnrm / S.oldnrm, 1.0 / S.mi);
        if (rate > 0.9) retval = 2;
        else M.ss = rate / (1.0 - rate);
    }
    if (retval == -99 && M.ss * delnrm <= 0.33) retval = 0;
    if (retval == -99) {
        S.mi++;
        if (S.mi >= o.maxcor) retval = 2;
    }
    return retval;
}

// start one attempt of IDAStep: coefficients, predictor, lsetup decision
__device__ __forceinline__ void attempt_begin(const ModelDesc& m, WarpWS& w, SimState& S, int lane) {
    S.ck = ida_set_coeffs(m, w, S.M, lane);
    nls_begin(m, w, S, lane);
    S.state = ST_NLS;
}

// IDASolve(ONE_STEP) up to the first residual evaluation.  Returns true if an evaluation is needed
// (state = ST_NLS), false if the call returned (S.ret_fl / S.ret_t set).
template <bool EXT>
__device__ __forceinline__ bool solve_begin(const SimArgs& a, WarpWS& w, SimState& S, int lane) {
    const ModelDesc& m = a.m;
    const Opts& o = a.o;
    Ida& M = S.M;
    __syncwarp();   // S lives in shared memory and is updated by all lanes together: converge first
    const double ur = DBL_EPSILON;
    const double tout = cur_tstop<EXT>(a, S);
    M.tstop = tout; M.tstopset = 1;
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
        if ((M.tn + M.hh - M.tstop) * M.hh > 0.0) M.hh = (M.tstop - M.tn) * (1.0 - 4.0 * ur);
        M.kk = 0; M.kused = 0;
        { double* p1 = w.v(V_PHI1); PLB_FOR_ELEMS(i, m.N_tot) p1[i] *= M.hh; }
        grp_sync();
    } else {
        const double troundoff = 100.0 * ur * (fabs(M.tn) + fabs(M.hh));
        if (fabs(M.tn - M.tstop) <= troundoff) {
            S.ret_t = M.tretlast = M.tstop; M.tstopset = 0; S.ret_fl = 1;
            return false;
        }
        if ((M.tn + M.hh - M.tstop) * M.hh > 0.0) M.hh = (M.tstop - M.tn) * (1.0 - 4.0 * ur);
        ewt_set(m, w, o, lane);
    }
    if (o.reltol < ur) {
        // IDA's "too much accuracy requested" test: |y_i| ewt_i < 1/rtol, so uround*||y|| > 1 needs rtol < uround
        const double nrm = wrms(m, w.v(V_PHI0), w.v(V_EWT), lane);
        if (ur * nrm > 1.0) { S.ret_t = M.tn; S.ret_fl = FAIL_CONV; return false; }
    }
    // IDAStep prologue
    S.saved_t = M.tn; S.ncf = 0; S.nef = 0; S.err_k = 0.0; S.err_km1 = 0.0;
    if (M.nst == 0) {
        M.kk = 1; M.kused = 0; M.hused = 0.0; M.cj = 1.0 / M.hh; M.phase = 0; M.ns = 0;
        grp_sync();
        if (lane == 0) w.K.psi[0] = M.hh;
        grp_sync();
    }
    attempt_begin(m, w, S, lane);
    return true;
}

// IDAStopTest2 after a successful step
__device__ __forceinline__ void solve_end(SimState& S) {
    Ida& M = S.M;
    __syncwarp();   // S lives in shared memory and is updated by all lanes together: converge first
    const double ur = DBL_EPSILON;
    if (M.tstopset) {
        const double troundoff = 100.0 * ur * (fabs(M.tn) + fabs(M.hh));
        if (fabs(M.tn - M.tstop) <= troundoff) {
            S.ret_t = M.tretlast = M.tstop; M.tstopset = 0; S.ret_fl = 1;
            return;
        }
        if ((M.tn + M.hh - M.tstop) * M.hh > 0.0) M.hh = (M.tstop - M.tn) * (1.0 - 4.0 * ur);
    }
    S.ret_t = M.tretlast = M.tn; S.ret_fl = 0;
}

// The Newton loop of one attempt ended with `retval` (0 ok, >0 recoverable failure): error test,
// step completion or retry bookkeeping (IDAStep / IDAHandleNFlag).  Sets S.pending.
__device__ __forceinline__ void after_nls(const ModelDesc& m, WarpWS& w, const Opts& o, SimState& S, int retval, int lane) {
    Ida& M = S.M;
    __syncwarp();   // S lives in shared memory and is updated by all lanes together: converge first
    bool errfail = false;
    if (retval == 0) errfail = ida_test_error(m, w, M, S.ck, S.err_k, S.err_km1, lane);
    if (retval == 0 && !errfail) {
        ida_complete_step(m, w, o, M, S.err_k, S.err_km1, lane);
        solve_end(S);
        S.pending = PEND_RETURNED;
        return;
    }
    ida_restore(m, w, M, S.saved_t, lane);
    M.phase = 1;
    int fail = 0;
    if (retval != 0) {
        if (lane == 0) M.ncfn++;     // statistics: written by one elected lane, read after the simulation's last barrier
        S.ncf++;
        M.rr = 0.25;
        M.hh *= M.rr;
        if (S.ncf >= o.maxncf) fail = FAIL_CONV;
    } else {
        S.nef++;
        if (lane == 0) M.netf++;
        if (S.nef == 1) {
            const double err_knew = (M.kk == M.knew) ? S.err_k : S.err_km1;
            M.kk = M.knew;
            M.rr = 0.9 * pow(2.0 * err_knew + 1e-4, -1.0 / (M.kk + 1));
            M.rr = fmax(0.25, fmin(0.9, M.rr));
            M.hh *= M.rr;
        } else if (S.nef == 2) {
            M.kk = M.knew; M.rr = 0.25; M.hh *= M.rr;
        } else if (S.nef < o.maxnef) {
            M.kk = 1; M.rr = 0.25; M.hh *= M.rr;
        } else fail = FAIL_ERRTEST;
    }
    if (!fail) {
        if (M.nst == 0) {
            grp_sync();
            if (lane == 0) w.K.psi[0] = M.hh;
#pragma unroll 1
            for (int i = lane; i < m.N_tot; i += LW) w.v(V_PHI1)[i] *= M.rr;
            grp_sync();
        }
        if (!(fabs(M.hh) > 0.0) || isinf(M.hh)) fail = FAIL_CONV;
    }
    if (fail) {
        S.ret_t = M.tretlast = M.tn; S.ret_fl = fail;
        S.pending = PEND_RETURNED;
        return;
    }
    attempt_begin(m, w, S, lane);   // PREDICT_AGAIN
}

// one turn of solve! after step!(int) returned (model_evaluation.jl:320-328, checks.jl:226-249)
// returns 1 to continue stepping, 0 to finish, 2 when a re-initialisation (Newton on the algebraic block at
// t + reltol, then IDAReInit) has been set up and the next tick is its first evaluation
template <bool EXT>
__device__ __forceinline__ int host_after_return(const SimArgs& a, WarpWS& w, SimState& S, int lane) {
    const ModelDesc& m = a.m;
    Ida& M = S.M;
    __syncwarp();   // S lives in shared memory and is updated by all lanes together: converge first
    const int N = m.N_tot;
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    const double tcur_stop = cur_tstop<EXT>(a, S);
    if (S.ret_fl == 1 || S.ret_t >= tcur_stop) { if (S.itstop < S.ntstops - 1) S.itstop++; }
    S.t = S.ret_t;
    S.iter++;
    // check_solve(run::run_function, ...) (checks.jl:251-268) has no "failed to converge" branch: a failed
    // step stores its point again and the discontinuity check at the end of this function decides
    if (!(EXT && a.tab_n) && (S.ret_fl < 0 || S.t == S.tprev)) {
        if (S.t == 0.0 && S.iter == 2 && !S.retried && M.nst == 0) {
            S.retried = 1;
            const double sc = 1.0 / w.K.psi[0];
#pragma unroll 1
            for (int i = lane; i < N; i += LW) w.v(V_PHI1)[i] *= sc;
            grp_sync();
            M.hin = a.o.reltol;
            return 1;
        }
        S.hard = (S.ret_fl == FAIL_ERRTEST) ? FAIL_ERRTEST : FAIL_CONV;
        return 0;
    }
    grp_sync();
    if (lane == 0) S.kord = getsol_weights(M, w.K, S.t, w.K.cvals, w.K.dvals);
    S.kord = grp_bcast_int(S.kord, 0);
    grp_sync();
    double Ic = 0.0, vp = 0.0, vn = 0.0;
#pragma unroll 1
    for (int j = 0; j <= S.kord; j++) {
        const double cj_ = w.K.cvals[j];
        const double* ph = w.v(V_PHI0 + j);
        Ic = fma(cj_, ph[m.off_I], Ic); vp = fma(cj_, ph[iP0], vp); vn = fma(cj_, ph[iPN], vn);
    }
    const double Vc = vp - vn;
    const double tg = S.t + S.t0;
    if (EXT) S.SOC_before = S.SOC;
    S.SOC = S.SOC + 0.5 * (tg - S.tg_prev) * (Ic + S.I_prev) / 3600.0;
    const size_t so = (size_t)S.sys * a.n_save_max;
    if (lane == 0 && S.nsave < a.n_save_max) {
        if (a.tr_t) a.tr_t[so + S.nsave] = tg;
        if (a.tr_V) a.tr_V[so + S.nsave] = Vc;
        if (a.tr_I) a.tr_I[so + S.nsave] = Ic;
        if (a.tr_SOC) a.tr_SOC[so + S.nsave] = S.SOC;
    }
    if (a.tr_T && S.nsave < a.n_save_max) {
        const double Tw = weighted_T(m, w, w.K.cvals, S.kord, false, lane);
        if (lane == 0) a.tr_T[so + S.nsave] = Tw;
    }
    if (EXT && a.tr_Y && S.nsave < a.n_save_max) {
        double* row = a.tr_Y + (so + S.nsave) * N;
#pragma unroll 1
        for (int i = lane; i < N; i += LW) row[ref_index(m, i)] = interp_y(w, w.K.cvals, S.kord, i);
    }
    S.nsave++;
    {   // copies: check_stop is out of line, and S must not have its address taken (it would live in local memory)
        PrevVals pv = S.pv;
        int flag = S.flag;
        const RunCtl rc = S.rc;
        check_stop(m, w, rc, a.o, a.b, a.input_kind == 2, a.tf, pv, flag, S.t, w.K.cvals, w.K.dvals, S.kord, S.SOC, Ic, Vc, lane);
        S.pv = pv; S.flag = flag;
    }
    if (S.iter == a.o.maxiters) { S.hard = FAIL_MAXITERS; return 0; }
    if (!(Ic == Ic) || !(Vc == Vc) || isinf(Ic) || isinf(Vc)) { S.hard = FAIL_NONFINITE; return 0; }
    if (S.flag != -1) return 0;
    if (EXT && a.n_dense && S.t > S.tprev) dense_emit(a, w, S, tg, false, lane);
    S.I_prev = Ic;
    S.tg_prev = tg;
    const double dt_step = S.t - S.tprev;
    S.tprev = S.t;
    if (EXT && a.tab_n && dt_step < 1e-3 * a.o.reltol) {
        // check_reinitialization! (checks.jl:341-364): value(run) is what the last residual evaluation
        // computed; a jump within the next reltol seconds restarts the DAE there
        const double t_new = S.t + a.o.reltol;
        const double v_old = S.rc.value, v_new = S.scale * table_eval(a.tab_t, a.tab_v, a.tab_n, t_new);
        const double tol = fmax(a.o.abstol, a.o.reltol * fmax(fabs(v_old), fabs(v_new)));
        if (!(fabs(v_old - v_new) <= tol)) {
            // Y = int.u (the interpolant at t) becomes the Newton start; the history is discarded
            grp_sync();
#pragma unroll 1
            for (int i = lane; i < N; i += LW) w.v(V_PHI0)[i] = interp_y(w, w.K.cvals, S.kord, i);
            grp_sync();
            S.t_init = t_new; S.reinit = 1; S.ni_iter = 0;
            S.state = ST_INIT_ITER;
            return 2;
        }
    }
    return 1;
}

// IDAReInit(mem, t_new, Y, YP) after the re-initialisation's newtons_method! (checks.jl:358-361)
__device__ PLB_COLD void reinit_integration(const SimArgs& a, WarpWS& w, SimState& S, int lane) {
    const ModelDesc& m = a.m;
    for (int k = 2; k < 6; k++) {
#pragma unroll 1
        for (int i = lane; i < m.N_tot; i += LW) w.v(V_PHI0 + k)[i] = 0.0;
    }
    grp_sync();
    Ida& M = S.M;
    M.tn = S.t_init; M.tretlast = S.t_init;
    M.hh = 0.0; M.hused = 0.0; M.cj = 0.0; M.cjlast = 0.0; M.cjold = 0.0; M.cjratio = 1.0;
    M.ss = 20.0; M.rr = 0.0; M.tstop = 0.0;
    M.kk = 0; M.kused = 0; M.knew = 0; M.phase = 0; M.ns = 0; M.nst = 0; M.tstopset = 0;
    S.reinit = 0;
    if (lane == 0) S.n_reinit++;
    S.pending = PEND_BEGIN;
}

// exit_simulation! (model_evaluation.jl:335-382) + summary / state hand-back
template <bool EXT>
__device__ PLB_COLD void finish(const SimArgs& a, WarpWS& w, SimState& S, bool integrated, int lane) {
    const ModelDesc& m = a.m;
    const Ida& M = S.M;
    const int N = m.N_tot;
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    Summary out;
    out.n_reinit = S.n_reinit; out.aux_end = 0.0;
    double t_end = S.t + S.t0, SOC_end = S.SOC, V_end = 0.0, I_end = 0.0, T_end = TH ? 0.0 : w.C.g[GC_T];
    const size_t so = (size_t)S.sys * a.n_save_max;
    if (integrated) {
        double fr = 1.0;
        bool do_interp = false;
        if (S.hard) S.flag = S.hard;
        else if (a.o.interp_final && S.flag != 0 && S.flag != -1 && S.t > 1.0) { do_interp = true; fr = S.pv.frac; }
        grp_sync();
        if (lane == 0) {
            if (S.nsave <= 1) { w.K.cvals[0] = 1.0; w.K.cvals[1] = 0.0; w.K.dvals[0] = M.nst == 0 ? 1.0 : 1.0 / w.K.psi[0]; }
            if (do_interp) {
                // Y_prev (sol.Y[end] of the reference) = the interpolant at the previous return time; when that was the
                // previous step point, tn - hused, use the exact step
                double dp[6];
                const bool at_step = fabs((M.tn - M.hused) - S.tprev) <= 100.0 * DBL_EPSILON * (fabs(M.tn) + fabs(M.hused));
                getsol_weights_delt(M, w.K, at_step ? -M.hused : S.tprev - M.tn, w.K.cprev, dp);
            }
        }
        grp_sync();
        double ps0 = 0.0, psN = 0.0, If = 0.0, Tw = 0.0, aux = 0.0;
#pragma unroll 1
        for (int i = lane; i < N; i += LW) {
            const double yn = interp_y(w, w.K.cvals, S.kord, i);
            double yf = yn;
            if (do_interp) { const double ypv = interp_y(w, w.K.cprev, S.kord, i); yf = fr * (yn - ypv) + ypv; }
            a.sY[(size_t)S.sys * N + ref_index(m, i)] = yf;
            if (EXT && a.tr_Y && do_interp && S.nsave >= 1 && S.nsave - 1 < a.n_save_max) a.tr_Y[(so + S.nsave - 1) * N + ref_index(m, i)] = yf;
            if (a.sYP) a.sYP[(size_t)S.sys * N + ref_index(m, i)] = interp_yp(w, w.K.dvals, S.kord, i);
            if (i == iP0) ps0 = yf;
            if (i == iPN) psN = yf;
            if (i == m.off_I) If = yf;
#if PLB_SEI
            if (i == m.off_SOH) aux = yf;
#endif
#if PLB_TH
            if (i >= m.off_T && i < m.off_film) {   // temperature_weighting of the final state (film | SOH follow the T block)
                const int k = i - m.off_T, x = k - m.Na;
                const int q = k < m.Na ? 0 : (x < m.Np ? 1 : (x < m.Np + m.Ns ? 2 : (x < m.Nx ? 3 : 4)));
                Tw = fma(yf, w.C.s5[0][q], Tw);
            }
#endif
        }
        ps0 = warp_sum(ps0); psN = warp_sum(psN); If = warp_sum(If);
        V_end = ps0 - psN; I_end = If;
        if (SEI) out.aux_end = warp_sum(aux);       // SOH of the final state
        (void)Tw; (void)aux;
#if PLB_TH
        {
            const double* th = w.C.theta;
            T_end = warp_sum(Tw) / (th[TF_l_a] + th[TF_l_p] + th[TF_l_s] + th[TF_l_n] + th[TF_l_z]);
        }
#endif
        if (do_interp) {
            const double ti = fr * (S.t - S.tprev) + S.tprev;
            const double tgi = ti + S.t0, tgl = S.t + S.t0;
            SOC_end = S.SOC + 0.5 * (tgi - tgl) * (If + If) / 3600.0;
            t_end = tgi;
            if (lane == 0 && S.nsave - 1 < a.n_save_max && S.nsave >= 1) {
                if (a.tr_t) a.tr_t[so + S.nsave - 1] = tgi;
                if (a.tr_V) a.tr_V[so + S.nsave - 1] = V_end;
                if (a.tr_I) a.tr_I[so + S.nsave - 1] = I_end;
                if (a.tr_SOC) a.tr_SOC[so + S.nsave - 1] = SOC_end;
                if (a.tr_T) a.tr_T[so + S.nsave - 1] = T_end;
            }
        }
        // dense output of the last (possibly shortened) step: requested times up to the end of the run
        if (EXT && a.n_dense && !S.hard && S.flag != -1 && S.t > S.tprev) dense_emit(a, w, S, t_end, false, lane);
    } else {
        // failed before integration: hand the (initial) state back
#pragma unroll 1
        for (int i = lane; i < N; i += LW) {
            a.sY[(size_t)S.sys * N + ref_index(m, i)] = w.v(V_PHI0)[i];
            if (a.sYP) a.sYP[(size_t)S.sys * N + ref_index(m, i)] = 0.0;
        }
        S.nsave = 1;
        t_end = S.t0;
        if (TH) T_end = w.C.g[GC_T];
    }
    out.t_end = t_end; out.V_end = V_end; out.I_end = I_end; out.SOC_end = SOC_end; out.T_end = T_end;
    out.flag = S.flag; out.n_steps = S.nsave - 1;
    out.n_res = M.nre; out.n_jac = M.nje; out.n_netf = M.netf; out.n_ncfn = M.ncfn;
    out.n_newton_init = S.n_newton_init;
    if (lane == 0) {
        a.out[S.sys] = out;
        a.sSOC[S.sys] = SOC_end;
        a.st[S.sys] = S.flag < 0 ? NAN : t_end;     // a hard failure poisons the continuation state (PLB_FAIL_PREVIOUS)
        if (a.tr_n) a.tr_n[S.sys] = S.nsave < a.n_save_max ? S.nsave : a.n_save_max;
    }
    grp_sync();
    S.state = ST_FETCH;
}

// initialize_simulation! up to the first Newton-init evaluation (model_evaluation.jl:174-214)
template <int CHEM, bool EXT>
__device__ PLB_COLD void fetch_and_setup(const SimArgs& a, WarpWS& w, const LaneRole& ro, SimState& S, int lane) {
    const ModelDesc& m = a.m;
    int sys = 0;
    if (lane == 0) sys = atomicAdd(a.counter, 1);
    sys = grp_bcast_int(sys, 0);
    if (sys >= a.B) { S.state = ST_EXHAUSTED; return; }
    S.sys = sys;
    const int N = m.N_tot;
    setup_consts(m, a.theta + (size_t)sys * m.theta_stride, w.C, lane);
    S.rc.method = a.method; S.rc.dc = 0;
    S.t_init = 0.0; S.reinit = 0; S.n_reinit = 0; S.scale = 1.0;
    if (EXT && a.tab_n) {   // run_function: initial_current! evaluates the function at t = 0 (input_methods.jl:27-29, 64-74, 104-107)
        S.scale = a.values ? a.values[sys] : 1.0;
        S.rc.value = S.scale * table_eval(a.tab_t, a.tab_v, a.tab_n, 0.0);
    } else S.rc.value = a.values ? a.values[sys] : a.value;
    double* Y0 = w.v(V_PHI0);
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    if (a.new_run) {
        S.SOC = a.soc0 ? a.soc0[sys] : 1.0;
        const double* th = w.C.theta;
        const double csp = th[TF_c_max_p] * (S.SOC * (th[TF_theta_max_p] - th[TF_theta_min_p]) + th[TF_theta_min_p]);
        const double csn = th[TF_c_max_n] * (S.SOC * (th[TF_theta_max_n] - th[TF_theta_min_n]) + th[TF_theta_min_n]);
        LaneVec y0;
        y0.ce = th[TF_c_e0]; y0.j = 0.0; y0.pe = 0.0; y0.ps = 0.0;
        y0.T = th[TF_T0]; y0.Tx = th[TF_T0];
        y0.js = 0.0; y0.film = 0.0; y0.soh = 1.0;
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
            } else if (CHEM == CHEM_LGM) {
                if (ro.sec == 0) laws::OCV_NMC811(thx, U, dU);
                else laws::OCV_LiC6_LGM50(thx, U, dU);
            } else {
                if (ro.sec == 0) laws::OCV_NMC(thx, U, dU);
                else laws::OCV_LiC6_NMC(thx, U, dU);
            }
            y0.ps = U;
        }
        store_lane(m, ro, Y0, y0, 0.0, lane);
        S.t0 = 0.0;
    } else {
#pragma unroll 1
        for (int i = lane; i < N; i += LW) Y0[i] = a.sY[(size_t)sys * N + ref_index(m, i)];
        S.SOC = a.sSOC[sys];
        S.t0 = ::nextafter(a.st[sys], DBL_MAX);   // initial_time, model_evaluation.jl:112
        if (!(S.t0 == S.t0)) {
            // an earlier segment of this system failed hard: pass it through untouched
            grp_sync();
            S.flag = FAIL_PREVIOUS; S.hard = 0; S.nsave = 0; S.t = 0.0; S.n_newton_init = 0; S.n_reinit = 0;
            S.M.nre = 0; S.M.nje = 0; S.M.netf = 0; S.M.ncfn = 0; S.M.nst = 0;
            finish<EXT>(a, w, S, false, lane);
            return;
        }
    }
    grp_sync();
    const double I_prev_state = Y0[m.off_I];
    // initial_current! (input_methods.jl:11-107)
    {
        double Ig;
        const double V0 = Y0[iP0] - Y0[iPN];
#if PLB_DC
        if (S.rc.method == METHOD_DC) {
            // input_method(::Val{:dc_s_p_max}, ...) etc. (input_methods.jl:195-245): the state with the largest / smallest
            // value in the last point of the previous solution; argmax / argmin return the first one
            const int k = a.dc_kind;
            const bool is_e = k >= DC_E_MAX, is_n = (k == DC_S_N_MAX || k == DC_S_N_MIN);
            const bool want_max = (k == DC_S_P_MAX || k == DC_S_N_MAX || k == DC_E_MAX);
            const LaneRole ro_ = make_role(m, lane);
            const bool cand = is_e ? ro_.act : (ro_.sec == (is_n ? 2 : 0));
            double v = 0.0;
            if (cand) v = is_e ? Y0[lane] : Y0[m.off_cs + ro_.e * NR + NR - 1];
            const double best = want_max ? grp_max(cand ? v : -INFINITY) : grp_min(cand ? v : INFINITY);
            const int tgt = (int)grp_min((cand && v == best) ? (double)lane : 1e9);
            S.rc.dc = tgt | ((is_e ? 1 : 0) << 8);
            if (a.input_kind == 1) S.rc.value = 0.0;          // custom_res!: hold_val = 0
            Ig = I_prev_state;                                // initial_current! of a run_residual (input_methods.jl:173-176)
        } else
#endif
        if (S.rc.method == METHOD_DT) {
            // run_residual: custom_res! (model_evaluation.jl:155-170) -> value, or 0 for :hold;
            // initial_current! (input_methods.jl:173-176): the previous current, else 1
            if (a.input_kind == 1) S.rc.value = 0.0;
            Ig = a.new_run ? 1.0 : I_prev_state;
        } else if (S.rc.method == METHOD_ETA) {
            // initial_current! for method_eta_p (input_methods.jl:118-143)
            const int ln = m.Np + m.Ns;
            if (a.input_kind == 1) { S.rc.value = Y0[m.off_ps + m.Np] - Y0[m.off_pe + ln]; Ig = I_prev_state; }
            else if (!a.new_run) Ig = I_prev_state;
            else Ig = S.rc.value > V0 ? 1.0 : -1.0;
        } else if (a.input_kind == 1) {
            if (S.rc.method == METHOD_I) { S.rc.value = I_prev_state; Ig = I_prev_state; }
            else if (S.rc.method == METHOD_V) { S.rc.value = V0; Ig = V0; }   // sic: input_methods.jl:58
            else { S.rc.value = I_prev_state * w.C.g[GC_I1C] * V0; Ig = I_prev_state; }
        } else if (a.input_kind == 2) {
            S.rc.value = 0.0; Ig = 0.0;
        } else if (S.rc.method == METHOD_I) Ig = S.rc.value;
        else if (S.rc.method == METHOD_V) {
            if (!a.new_run && ((EXT && a.tab_n) || I_prev_state != 0.0)) Ig = I_prev_state;   // :40-52 / :64-74
            else Ig = S.rc.value > V0 ? 1.0 : -1.0;
        } else Ig = S.rc.value / (V0 * w.C.g[GC_I1C]);
        grp_sync();
        if (lane == 0) Y0[m.off_I] = Ig;
        grp_sync();
    }
    Ida& M = S.M;
    M.tn = 0.0; M.hh = 0.0; M.hused = 0.0; M.cj = 0.0; M.cjlast = 0.0; M.cjold = 0.0; M.cjratio = 1.0;
    M.ss = 20.0; M.rr = 0.0; M.hin = 0.0; M.tstop = 0.0; M.tretlast = 0.0;
    M.kk = 0; M.kused = 0; M.knew = 0; M.phase = 0; M.ns = 0; M.nst = 0; M.tstopset = 0;
    M.nre = 0; M.nje = 0; M.netf = 0; M.ncfn = 0;
    S.t = 0.0; S.tprev = 0.0; S.flag = -1; S.iter = 1; S.hard = 0; S.nsave = 0; S.kord = 1; S.retried = 0;
    S.ni_iter = 0; S.n_newton_init = 0; S.pending = PEND_NONE;
    S.idense = 0; S.SOC_before = S.SOC;
    S.state = ST_INIT_ITER;
}

// after newtons_method!: rest of initialize_simulation! (model_evaluation.jl:216-231)
template <bool EXT>
__device__ PLB_COLD void begin_integration(const SimArgs& a, WarpWS& w, SimState& S, int lane) {
    const ModelDesc& m = a.m;
    const int N = m.N_tot;
    const int iP0 = m.off_ps, iPN = m.off_ps + m.Ne - 1;
    const double* Y0 = w.v(V_PHI0);
    if (a.new_run) {   // check_initial_SOC, checks.jl:327-339
        const double I0 = Y0[m.off_I];
        if (I0 != 0 && ((S.SOC >= a.b.SOC_max && I0 > 0) || (S.SOC <= a.b.SOC_min && I0 < 0))) {
            S.flag = FAIL_INIT_BOUNDS;
            finish<EXT>(a, w, S, false, lane);
            return;
        }
    }
    for (int k = 2; k < 6; k++) {
#pragma unroll 1
        for (int i = lane; i < N; i += LW) w.v(V_PHI0 + k)[i] = 0.0;
    }
    grp_sync();
    S.ntstops = 0; S.itstop = 0;
    if (!a.new_run && 1.0 < a.tf) { S.tstop0 = 1.0; S.tstop1 = a.tf; S.ntstops = 2; }
    else { S.tstop0 = a.tf; S.tstop1 = a.tf; S.ntstops = 1; }
    if (EXT && a.n_tstops) S.ntstops = a.n_tstops;
    grp_sync();
    if (lane == 0) { w.K.cvals[0] = 1.0; w.K.cvals[1] = 0.0; w.K.dvals[0] = 1.0; }   // y = phi0, yp = phi1 = YP0
    grp_sync();
    const double Vc = Y0[iP0] - Y0[iPN], Ic = Y0[m.off_I];
    S.I_prev = Ic;
    const size_t so = (size_t)S.sys * a.n_save_max;
    if (lane == 0 && S.nsave < a.n_save_max) {
        if (a.tr_t) a.tr_t[so + S.nsave] = S.t0;
        if (a.tr_V) a.tr_V[so + S.nsave] = Vc;
        if (a.tr_I) a.tr_I[so + S.nsave] = Ic;
        if (a.tr_SOC) a.tr_SOC[so + S.nsave] = S.SOC;
    }
    if (a.tr_T && S.nsave < a.n_save_max) {
        const double Tw = weighted_T(m, w, w.K.cvals, 1, false, lane);
        if (lane == 0) a.tr_T[so + S.nsave] = Tw;
    }
    if (EXT && a.tr_Y && S.nsave < a.n_save_max) {
        double* row = a.tr_Y + (so + S.nsave) * N;
#pragma unroll 1
        for (int i = lane; i < N; i += LW) row[ref_index(m, i)] = Y0[i];
    }
    S.nsave++;
    if (EXT && a.n_dense) {
        // requested times before the start of a continuation belong to the earlier runs; the start itself is a row
        if (lane == 0 && a.dn_n) a.dn_n[S.sys] = 0;
        if (!a.new_run) { int k = 0; while (k < a.n_dense && __ldg(a.dense_t + k) < S.t0) k++; S.idense = k; }
        dense_emit(a, w, S, S.t0, true, lane);
    }
    S.pv.T = -1; S.pv.dfilm = -1;
    S.pv.frac = 1.0; S.pv.V = -1; S.pv.SOC = -1; S.pv.c_s_n = -1; S.pv.I = -1; S.pv.eta_plating = -1; S.pv.c_e_min = -1;
    S.kord = 1;
    {
        PrevVals pv = S.pv;
        int flag = S.flag;
        const RunCtl rc = S.rc;
        check_stop(m, w, rc, a.o, a.b, a.input_kind == 2, a.tf, pv, flag, 0.0, w.K.cvals, w.K.dvals, 1, S.SOC, Ic, Vc, lane);
        S.pv = pv; S.flag = flag;
    }
    S.tg_prev = S.t0;
    S.pending = (S.flag == -1) ? PEND_BEGIN : PEND_FINISH;
}

// EXT: the run uses a tabulated input and/or keeps every saved state row (compiled out of the plain kernel,
// whose instruction footprint is what bounds it)
template <int CHEM, bool EXT>
__device__ __forceinline__ void simulate_cta(const SimArgs& a, unsigned char* smem_raw) {
    const int warp = grp_id(), lane = grp_lane();
    WarpWS w = make_ws(smem_raw, a.gws, warp);
    const ModelDesc& m = a.m;
    const LaneRole ro = make_role(m, lane);
#if PLB_STATE_SMEM
    static_assert(sizeof(SimState) <= STATE_BYTES, "STATE_BYTES too small");
    // every lane of a warp writes the same values at the same instruction: one copy per physical warp
    SimState& S = *reinterpret_cast<SimState*>(smem_raw + STATE_OFFSET + STATE_BYTES * (threadIdx.x >> 5));
#else
    SimState S;
#endif
    S.state = ST_FETCH; S.pending = PEND_NONE; S.sys = 0;
#if PLB_VEC2
    // the paired passes reach one component into the padding of the vectors when N is odd: zero it once
    for (int k = 0; k < V_COUNT; k++) {
#pragma unroll 1
        for (int i = m.N_tot + lane; i < VS; i += LW) w.v(k)[i] = 0.0;
    }
    grp_sync();
#endif
#if PLB_TICK_SYNC_EVERY > 1
    unsigned tick_no = 0;
#endif
    for (;;) {
        // ------------------------------ PRE: get to an evaluation point ---------------------------------
        while (S.state == ST_FETCH) fetch_and_setup<CHEM, EXT>(a, w, ro, S, lane);   // (again after a passed-through system)
        if (EXT && a.tab_n && S.state != ST_EXHAUSTED) {
            // run.func(t) of this tick's evaluation: IDA evaluates F at the trial time tn; newtons_method! at its
            // t, except for the algebraic-derivative estimate, which passes dt itself as the time
            // (model_evaluation.jl:469).  Done here, before the lane vectors are live, to keep the call cheap.
            const double te = S.state == ST_NLS ? S.M.tn : (S.state == ST_INIT_DT ? S.dt_init : S.t_init);
            __syncwarp();
            S.rc.value = S.scale * table_eval(a.tab_t, a.tab_v, a.tab_n, te);
        }
        LaneVec y, yp, res;
        double Iy = 0.0;
        bool do_eval = S.state != ST_EXHAUSTED, need_jac = false, alg_only = false, do_solve = false;
        yp.ce = 0.0; yp.j = 0.0; yp.pe = 0.0; yp.ps = 0.0; yp.T = 0.0; yp.Tx = 0.0; yp.js = 0.0; yp.film = 0.0; yp.soh = 0.0;
#pragma unroll
        for (int r = 0; r < NR; r++) yp.cs[r] = 0.0;
        if (do_eval) {
            if (S.state == ST_NLS) {
                LaneVec e;
                double eI;
                predictor_lane(m, ro, w, S.M.kk, y, Iy, yp);       // y_pred, y'_pred from the history
                load_lane(m, ro, w.v(V_EE), e, eI);
                const double cj = S.M.cj;
                y.ce += e.ce; y.j += e.j; y.pe += e.pe; y.ps += e.ps; Iy += eI;
                yp.ce = yp.ce + cj * e.ce;
                if (TH) { y.T += e.T; y.Tx += e.Tx; yp.T = yp.T + cj * e.T; yp.Tx = yp.Tx + cj * e.Tx; }
                if (SEI) {
                    y.js += e.js; y.film += e.film; y.soh += e.soh;
                    yp.film = yp.film + cj * e.film; yp.soh = yp.soh + cj * e.soh;
                }
#pragma unroll
                for (int r = 0; r < NR; r++) { y.cs[r] += e.cs[r]; yp.cs[r] = yp.cs[r] + cj * e.cs[r]; }
                // (the algebraic components of y' never enter a residual)
                yp.j = 0.0; yp.pe = 0.0; yp.ps = 0.0; yp.js = 0.0;
                need_jac = S.callLSetup != 0; do_solve = !(PLB_TICK_SPLIT_LSETUP && need_jac);
            } else {
                load_lane(m, ro, w.v(V_PHI0), y, Iy);
                alg_only = true;
                if (S.state == ST_INIT_ITER) { need_jac = true; do_solve = true; }
                else if (S.state == ST_INIT_DT) {
                    LaneVec p;
                    double pI;
                    load_lane(m, ro, w.v(V_PHI1), p, pI);
                    y.ce += S.dt_init * p.ce;
                    if (TH) { y.T += S.dt_init * p.T; y.Tx += S.dt_init * p.Tx; }
                    if (SEI) { y.film += S.dt_init * p.film; y.soh += S.dt_init * p.soh; }
#pragma unroll
                    for (int r = 0; r < NR; r++) y.cs[r] += S.dt_init * p.cs[r];
                    do_solve = true;
                }
            }
        }
        // ------------------------------ aligned heavy phases --------------------------------------------
#if PLB_TICK_SYNC_START
#if PLB_TICK_SYNC_EVERY > 1
        if ((tick_no++ % PLB_TICK_SYNC_EVERY) == 0) { if (!__syncthreads_or(do_eval)) break; }
#elif PLB_TICK_SYNC_STEP
        // one barrier per step ATTEMPT, not per evaluation: the second and later Newton iterations of an attempt follow
        // the first without one (every warp passes the same number of barriers: between two of them it runs one attempt,
        // or one evaluation of the initialisation, or nothing)
        if (!(S.state == ST_NLS && S.mi > 0)) { if (!__syncthreads_or(do_eval)) break; }
#else
        if (!__syncthreads_or(do_eval)) break;
#endif
#else
        if (!do_eval) break;
#endif
#if PLB_TICK_VOTE_JAC
        const int any_jac = __syncthreads_or(do_eval && need_jac);
#else
        const int any_jac = do_eval && need_jac;
#endif
        LaneJac J;
        CtrlRow ctrl;
        ctrl.res = 0.0; ctrl.g_ps0 = 0.0; ctrl.g_psN = 0.0; ctrl.g_I = 1.0; ctrl.gTn = 0.0; ctrl.gTx = 0.0; ctrl.g_eta = 0.0;
        if (do_eval) {
            // each warp picks the variant from its OWN state only, so a system's arithmetic (and therefore
            // its bits) never depends on which other systems share the CTA
            // the dT control row takes its newtons_method! form during the algebraic initialisation
            const int meth = method_word(S.rc, alg_only);
#if PLB_TICK_OUTLINE
            if (need_jac) lane_eval_ni<CHEM, true>(m, w.C, ro, y, yp, Iy, meth, S.rc.value, res, ctrl, J);
            else lane_eval_ni<CHEM, false>(m, w.C, ro, y, yp, Iy, meth, S.rc.value, res, ctrl, J);
#else
#if PLB_TICK_ONE_EVAL
            lane_eval<CHEM, true>(m, w.C, ro, y, yp, Iy, meth, S.rc.value, res, ctrl, J);
#else
            if (need_jac) lane_eval<CHEM, true>(m, w.C, ro, y, yp, Iy, meth, S.rc.value, res, ctrl, J);
            else lane_eval<CHEM, false>(m, w.C, ro, y, yp, Iy, meth, S.rc.value, res, ctrl, J);
#endif
#endif
            if (lane == 0 && (do_solve || S.state != ST_NLS)) S.M.nre++;      // (a split lsetup tick repeats its evaluation: counted once)
        }
        bool lsetup_bad = false;
        if (any_jac) {
#if PLB_TICK_SYNC_JAC
            __syncthreads();
#endif
            if (do_eval && need_jac) {
#if PLB_TICK_OUTLINE
                warp_factor(m, ro, J, ctrl, alg_only ? 0.0 : S.M.cj, alg_only, w.Fa, lane);
#else
                warp_factor_impl(m, ro, J, ctrl, alg_only ? 0.0 : S.M.cj, alg_only, w.Fa, lane);
#endif
                if (lane == 0) S.M.nje++;
                const double chk = w.Fa.schur_inv;
                lsetup_bad = !(chk == chk) || isinf(chk);
            }
        }
#if PLB_TICK_SYNC_SOLVE
        __syncthreads();
#endif
        double dI = 0.0;
        if (do_eval && do_solve && !lsetup_bad) {
            double gI = ctrl.res;
            if (S.state == ST_NLS) {   // Newton update of the BDF step: delta = -J^{-1} F
                res.ce = -res.ce; res.j = -res.j; res.pe = -res.pe; res.ps = -res.ps; gI = -gI;
                if (TH) { res.T = -res.T; res.Tx = -res.Tx; }
                if (SEI) { res.js = -res.js; res.film = -res.film; res.soh = -res.soh; }
#pragma unroll
                for (int r = 0; r < NR; r++) res.cs[r] = -res.cs[r];
            }
#if PLB_TICK_OUTLINE
            dI = warp_solve(m, ro, w.Fa, alg_only, res, gI, lane);
#else
            dI = warp_solve_impl(m, ro, w.Fa, alg_only, res, gI, lane);
#endif
        }
        // ------------------------------ POST: per-state glue --------------------------------------------
        __syncwarp();
        if (do_eval) {
            bool start_now = false;
            if (S.state == ST_NLS) {
                int retval;
                if (lsetup_bad) retval = 1;
                else {
                    if (need_jac) { S.M.cjold = S.M.cj; S.M.cjratio = 1.0; S.M.ss = 20.0; S.jcur = 1; S.callLSetup = 0; }
                    if (PLB_TICK_SPLIT_LSETUP && need_jac) retval = -99;      // the iteration itself runs in the next tick, on these factors
                    else retval = nls_post(m, w, ro, a.o, S, res, dI, lane);
                }
                if (retval > 0 && !S.jcur && !lsetup_bad) {
                    // recoverable failure with a stale Jacobian: redo once with a fresh one
                    S.callLSetup = 1; S.mi = 0;
#pragma unroll 1
                    for (int i = lane; i < m.N_tot; i += LW) w.v(V_EE)[i] = 0.0;
                    grp_sync();
                } else if (retval != -99) {
                    after_nls(m, w, a.o, S, retval, lane);
                }
            } else if (S.state == ST_INIT_ITER) {
                // Y_new .-= factor \ res ; absolute 2-norm of the update (model_evaluation.jl:451-454)
                // (Y is re-read from phi_0 here rather than kept live across the factorisation and the solve:
                // 36+ registers less pressure on the tick's hot phases)
                load_lane(m, ro, w.v(V_PHI0), y, Iy);
                double s = 0.0;
                if (ro.elec) { y.j -= res.j; y.ps -= res.ps; s += res.j * res.j + res.ps * res.ps; }
                if (ro.act) { y.pe -= res.pe; s += res.pe * res.pe; }
                if (SEI && ro.sec == 2) { y.js -= res.js; s += res.js * res.js; }
                Iy -= dI;
                s = warp_sum(s) + dI * dI;
                store_lane(m, ro, w.v(V_PHI0), y, Iy, lane);
                grp_sync();
                S.ni_iter++;
                const bool bad = lsetup_bad || !(s == s) || isinf(s);
                if (bad || (S.ni_iter >= 100 && !(sqrt(s) < a.o.reltol_init))) {
                    S.flag = FAIL_NEWTON_INIT; S.n_newton_init = FAIL_NEWTON_INIT;
                    finish<EXT>(a, w, S, false, lane);   // (also for a failed re-initialisation: the run is reported from its start)
                } else if (sqrt(s) < a.o.reltol_init) {
                    if (!(EXT && S.reinit)) S.n_newton_init = S.ni_iter;
                    S.state = ST_INIT_RDIFF;
                }
            } else if (S.state == ST_INIT_RDIFF) {
                // R_diff(YP,t,Y,YP): YP_diff = rhs (model_evaluation.jl:460); algebraic part still zero
                LaneVec ypo;
                ypo.ce = res.ce; ypo.j = 0.0; ypo.pe = 0.0; ypo.ps = 0.0;
                ypo.T = TH ? res.T : 0.0; ypo.Tx = TH ? res.Tx : 0.0;
                ypo.film = SEI ? res.film : 0.0; ypo.soh = SEI ? res.soh : 0.0; ypo.js = 0.0;
#pragma unroll
                for (int r = 0; r < NR; r++) ypo.cs[r] = res.cs[r];
                store_lane(m, ro, w.v(V_PHI1), ypo, 0.0, lane);
                grp_sync();
                const double c0 = fabs(w.C.theta[TF_c_e0]);
                const double epsv = ::nextafter(c0, DBL_MAX) - c0;
                S.dt_init = fmax(10.0 * a.o.reltol_init, sqrt(epsv));   // :464
                S.state = ST_INIT_DT;
                start_now = a.o.skip_alg_deriv != 0;   // initialize_algebraic_derivatives = false (:433): Y'_alg stays 0
            } else {   // ST_INIT_DT: YP_alg = -(factor \ R_alg(Y + dt YP)) / dt  (:466-476)
                LaneVec ypo;
                double pI;
                load_lane(m, ro, w.v(V_PHI1), ypo, pI);
                ypo.j = -res.j / S.dt_init; ypo.pe = -res.pe / S.dt_init; ypo.ps = -res.ps / S.dt_init;
                if (SEI) ypo.js = -res.js / S.dt_init;
                grp_sync();   // (as above: the I slot of phi_1 is read by every lane, written by lane 0)
                store_lane(m, ro, w.v(V_PHI1), ypo, -dI / S.dt_init, lane);
                grp_sync();
                start_now = true;
            }
            if (start_now) {
                if (EXT && S.reinit) reinit_integration(a, w, S, lane);
                else begin_integration<EXT>(a, w, S, lane);
            }
            // between evaluations: host loop of solve! until the next evaluation is needed
            while (S.pending != PEND_NONE) {
                if (S.pending == PEND_RETURNED) { const int r = host_after_return<EXT>(a, w, S, lane); S.pending = r == 1 ? PEND_BEGIN : (r == 0 ? PEND_FINISH : PEND_NONE); }
                if (S.pending == PEND_BEGIN) S.pending = solve_begin<EXT>(a, w, S, lane) ? PEND_NONE : PEND_RETURNED;
                if (S.pending == PEND_FINISH) { finish<EXT>(a, w, S, true, lane); S.pending = PEND_NONE; }
            }
        }
    }
}

}  // namespace PLB_NS
}  // namespace plb
