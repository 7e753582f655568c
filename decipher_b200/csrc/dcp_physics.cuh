// dcp_physics.cuh — one HRU, one timestep: root zone -> unsaturated zone -> analytical
// saturated-zone solver -> results_ac.  Shared by the streamed and the resident kernels.
//
// Two arithmetic policies:
//   EXACT (FAST=false): the reference's operations in the reference's order, IEEE division, no FMA
//          contraction (the TU is built with -fmad=false).  Differs from the CPU only through
//          CUDA's exp/log (<= 1 ulp each).
//   FAST  (FAST=true):  divisions by per-run / per-HRU constants become multiplications by hoisted
//          reciprocals and 1/exp(-x) becomes exp(x).  Each substitution is <= ~1.5 ulp; measured
//          drift on flows stays far inside the 1e-9 contract (tests/test_gpu_parity.py).
#pragma once
#include "dcp_device.cuh"
#include "dcp_fastmath.cuh"

struct HruConst {            // per (HRU, member) constants for one step
    double ac, inv_ac;       // area, 1/area (FAST only)
    double q0, q1, q2;       // q0, q0*cos, q0*cos*exp(-cos*smax/szm)  (dyna_satzone_analyticalsolver.f90:49-51)
    double m2, inv_m2;       // szm/cos (:52), cos/szm (FAST only)
    double td_dt, dtt_inv_m2; // td/dt and dtt*cos/szm: the folded constants of hru_step_lean<., true>
    double srmax, inv_srmax, td, smax;
    int ims_rz, ims_uz;
};

struct StepConst {
    double dt, dtt, inv_dtt;
    int ntt;
    const double *etab = nullptr, *ltab = nullptr;   // shared-memory tables of fm_exp_tab / fm_log_tab (hru_step_lean<., ., true>)
};

struct HruState {
    double sd, suz, srz, qint1;
    double e1;               // lean step with CARRY: exp(sd/m2) carried from the previous step
};

struct HruFlux {             // outputs of one step
    double qbf;              // new baseflow per unit area
    double ex;               // exs + exus   (dyna_results_ac.f90:34)
    double ea;               // actual evapotranspiration of the step
    double pex;              // root-zone excess
    double exs1, exs2, exus; // checked mode: -sd part, (qbf-q0)*dtt part (dyna_satzone_analyticalsolver.f90:107-119)
    int rz_bad;              // root-zone balance STOP condition (dyna_root_zone1.f90:73-77)
};

template <bool FAST, bool CHECK>
__device__ __forceinline__ void hru_step(const HruConst &c, const StepConst &g, double p, double ep,
                                         double qin, HruState &s, HruFlux &f) {
    // ---------------- root_zone1 (RRMODEL/dyna_root_zone1.f90:32-77) ----------------
    double pex = 0.0, ea = 0.0;
    f.rz_bad = 0;
    if (c.ims_rz == 1) {
        const double storeini = s.srz;
        double srz = s.srz + p;
        if (srz > c.srmax) {
            pex = srz - c.srmax;
            srz = c.srmax;
        }
        double storeout = pex;
        if (ep > 0.0) {
            ea = FAST ? ep * (srz * c.inv_srmax) : ep * (srz / c.srmax);
            if (ea > ep) ea = ep;
            if (ea >= srz) ea = srz;
            srz = srz - ea;
            storeout = storeout + ea;
        }
        if (CHECK) {
            const double storebal = storeini + p - storeout - srz;
            if (storebal > DCP_F32_1EM10 || storebal < -1.0) f.rz_bad = 1;
        }
        s.srz = srz;
    }
    // ---------------- unsat_zone1 (RRMODEL/dyna_unsat_zone1.f90:40-86) ----------------
    double uz = 0.0, exus = 0.0;
    if (c.ims_uz == 1) {
        double uz2 = 0.0;
        double suz = s.suz + pex;
        const double sd = s.sd;
        if (sd <= 0.0 && suz > 0.0) {
            exus = suz;
            suz = 0.0;
        } else if (sd > 0.0 && suz > sd) {
            exus = suz - sd;
            suz = sd;
        }
        if (sd > 0.0 && suz > 0.0) {
            uz2 = g.dt * (suz / (sd * c.td));
            if (uz2 > suz) uz2 = suz;
            suz = suz - uz2;
            if (suz < 0.0000001) {
                uz2 = uz2 + suz;
                suz = 0.0;
            }
        }
        if (uz2 > 0.0) uz = FAST ? uz2 * g.inv_dtt : uz2 / g.dtt;
        s.suz = suz;
    }
    // ---------------- satzone_analyticalsolver (RRMODEL/dyna_satzone_analyticalsolver.f90:49-131) ----
    const double storeini = s.sd;
    double sd = s.sd, qbf = 0.0;
    const double qin_ac = FAST ? qin * c.inv_ac : qin / c.ac;
    for (int itt = 1; itt <= g.ntt; ++itt) {
        double qin2;
        if (FAST)
            qin2 = (g.ntt == 1) ? (s.qint1 + (qin - s.qint1)) * c.inv_ac
                                : (s.qint1 + (double)itt * (qin - s.qint1) / (double)g.ntt) * c.inv_ac;
        else
            qin2 = (s.qint1 + (double)itt * (qin - s.qint1) / (double)g.ntt) / c.ac;      // :63
        const double u = qin2 + uz;
        const double u2 = c.q2 + u;
        if (sd <= c.smax) {
            if (u2 > 0.0) {
                const double a = FAST ? (u2 * g.dtt) * c.inv_m2 : u2 * g.dtt / c.m2;
                const double e1 = FAST ? exp(sd * c.inv_m2) : 1.0 / exp(-sd / c.m2);
                if (a < DCP_F32_1EM10) {                                                   // :72-77
                    const double b = FAST ? (c.q1 * g.dtt) * c.inv_m2 : c.q1 * g.dtt / c.m2;
                    const double hb = FAST ? 0.5 * b : 0.5 * c.q1 * g.dtt / c.m2;
                    sd = c.m2 * log((e1 + b) - a * (e1 - hb));
                } else {                                                                   // :80
                    const double r = c.q1 / u2;
                    sd = c.m2 * log(r + (e1 - r) * exp(-a));
                }
            } else {                                                                       // :84
                const double b = FAST ? (c.q1 * g.dtt) * c.inv_m2 : c.q1 * g.dtt / c.m2;
                sd = c.m2 * log((FAST ? exp(sd * c.inv_m2) : exp(sd / c.m2)) + b);
            }
        } else {                                                                           // :86-95
            const double tc = (sd - c.smax) / u;
            if (tc > g.dtt) {
                sd = sd - u * g.dtt;
            } else {
                const double t_temp = g.dtt - tc;
                const double r = c.q1 / u2;
                const double e1 = FAST ? exp(c.smax * c.inv_m2) : 1.0 / exp(-c.smax / c.m2);
                const double a = FAST ? (u2 * t_temp) * c.inv_m2 : u2 * t_temp / c.m2;
                sd = c.m2 * log(r + (e1 - r) * exp(-a));
            }
        }
        qbf = (FAST ? (sd - storeini) * g.inv_dtt : (sd - storeini) / g.dtt) + uz + qin_ac;   // :98
    }
    double exs = 0.0, exs1 = 0.0, exs2 = 0.0;
    if (sd < 0.0) {                                                                        // :107-112
        exs = 0.0 - sd;
        exs1 = exs;
        sd = 0.0;
    }
    if (qbf > c.q0) {                                                                      // :115-119
        exs2 = (qbf - c.q0) * g.dtt;
        exs = exs + exs2;
        qbf = c.q0;
    }
    s.sd = sd;
    s.qint1 = qin;                                                                         // dyna_results_ac.f90:27
    f.qbf = qbf;
    f.ex = exs + exus;
    f.ea = ea;
    f.pex = pex;
    f.exs1 = exs1;
    f.exs2 = exs2;
    f.exus = exus;
}

// ------------------------------------------------------------------------------------------------
// Generic saturated-zone step alone (slow path of the lean step below): the sat-zone statements of
// hru_step<true, false>, kept out of line and fed by value so the hot loop holds no addressable structs.
// ------------------------------------------------------------------------------------------------
struct SatOut { double sd, qbf, exs; };

static __device__ __noinline__ SatOut sat_zone_slow(double inv_ac, double q0, double q1, double q2, double m2, double inv_m2,
                                             double smax, double dtt, double inv_dtt, int ntt, double qin, double uz,
                                             double qint1, double sd) {
    const double storeini = sd;
    double qbf = 0.0;
    const double qin_ac = qin * inv_ac;
    for (int itt = 1; itt <= ntt; ++itt) {
        const double qin2 = (ntt == 1) ? (qint1 + (qin - qint1)) * inv_ac
                                       : (qint1 + (double)itt * (qin - qint1) / (double)ntt) * inv_ac;
        const double u = qin2 + uz;
        const double u2 = q2 + u;
        if (sd <= smax) {
            if (u2 > 0.0) {
                const double a = (u2 * dtt) * inv_m2;
                const double e1 = exp(sd * inv_m2);
                if (a < DCP_F32_1EM10) {
                    const double b = (q1 * dtt) * inv_m2;
                    sd = m2 * log((e1 + b) - a * (e1 - 0.5 * b));
                } else {
                    const double r = q1 / u2;
                    sd = m2 * log(r + (e1 - r) * exp(-a));
                }
            } else {
                sd = m2 * log(exp(sd * inv_m2) + (q1 * dtt) * inv_m2);
            }
        } else {
            const double tc = (sd - smax) / u;
            if (tc > dtt) {
                sd = sd - u * dtt;
            } else {
                const double r = q1 / u2;
                sd = m2 * log(r + (exp(smax * inv_m2) - r) * exp(-((u2 * (dtt - tc)) * inv_m2)));
            }
        }
        qbf = (sd - storeini) * inv_dtt + uz + qin_ac;
    }
    double exs = 0.0;
    if (sd < 0.0) { exs = 0.0 - sd; sd = 0.0; }
    if (qbf > q0) { exs = exs + (qbf - q0) * dtt; qbf = q0; }
    SatOut o; o.sd = sd; o.qbf = qbf; o.exs = exs;
    return o;
}

// ------------------------------------------------------------------------------------------------
// LEAN path of the FAST policy: the same model equations written branch-free (selects instead of
// branches, fm_exp / fm_log / fm_rcp instead of the library calls) so that several member slots
// interleave in one basic block and the FP64 pipe stays busy.  Preconditions (checked by the caller,
// who otherwise uses hru_step<true,...>): ims_rz == ims_uz == 1, ntt == 1, epp = max(ep, 0).
// The regular and the small-u2 branch of the solver (dyna_satzone_analyticalsolver.f90:72-84) are both
// evaluated and selected; everything else (sd > smax, arguments outside the fast exp range) raises
// `slow`, and the caller re-runs the saturated zone of that lane with sat_zone_generic.
// Outputs uz and exus are returned so that the slow path can be completed.
// ------------------------------------------------------------------------------------------------
struct LeanOut {
    double qbf, exac, ea, uz, exus, sd0;
    bool slow;
};

// accuracy experiments (tools/err_probe.py): swap single primitives of the lean step for the library ones
#ifdef DCP_LEAN_LIBM_EXP
#define LEAN_EXP(x) exp(x)
#else
#define LEAN_EXP(x) fm_exp(x)
#endif
#ifdef DCP_LEAN_LIBM_LOG
#define LEAN_LOG(x) log(x)
#else
#define LEAN_LOG(x) fm_log(x)
#endif
#ifdef DCP_LEAN_TRUE_DIV
#define LEAN_DIV(a, b) ((a) / (b))
#else
#define LEAN_DIV(a, b) ((a) * fm_rcp(b))
#endif

// CARRY: the solver's new deficit is sd = m2*log(arg) (:76-84), so the next step's exp(sd/m2) IS this step's
// `arg` (or 1 when the deficit is clipped at zero, :107-112): the exponential is carried in the state instead of
// being recomputed (one exp and the quotient correction less per step).  The stored sd stays the rounded
// m2*log(arg), consistent with the carried value to an ulp each step, so nothing drifts; callers re-derive e1
// with exp() whenever they replace sd (slow path, initial state).
// TAB: table-driven exp / log (dcp_fastmath.cuh: 22 FP64 instructions instead of 41 for the pair)
template <bool CARRY = false, bool FOLD = false, bool TAB = false>
__device__ __forceinline__ void hru_step_lean(const HruConst &c, const StepConst &g, double p, double epp,
                                              double qin, HruState &s, LeanOut &o) {
    // root_zone1 (dyna_root_zone1.f90:39-66)
    double srz = s.srz + p;
    const double pex = fm_relu(srz - c.srmax);
    srz = srz - pex;                                      // min(srz, srmax): one DADD instead of a compare and two selects
                                                          // (exact whenever srz <= 2 srmax, Sterbenz; an ulp of srz otherwise)
    double ea = epp * (srz * c.inv_srmax);
    ea = fm_min(ea, srz);                                 // srz <= srmax already, so ea <= ep up to an ulp: min(.., ep) dropped
    srz = srz - ea;
    // unsat_zone1 (dyna_unsat_zone1.f90:49-78)
    const double sd0 = s.sd;
    double suz = s.suz + pex;
    const double lim = fm_relu(sd0);
    const double exus = fm_relu(suz - lim);
    suz = suz - exus;                                     // min(suz, lim), as above
    // FOLD: dt and dtt are folded into per-member constants (td/dt, dtt*cos/szm): three multiplications less per unit
    const double den = sd0 * (FOLD ? c.td_dt : c.td);
    double uz2 = FOLD ? LEAN_DIV(suz, den) : g.dt * LEAN_DIV(suz, den);
    // :60-68 drain at most what is there.  No guards are needed around the quotient: with sd <= 0 the store was just
    // emptied (lim = 0, so suz = 0 and min(x, 0) = 0 for x = +-0 or NaN, because fm_min returns its second operand
    // unless the first is smaller); with a vanishing sd*td the quotient is +inf or NaN and the same min returns suz,
    // which is the reference's clip; with suz = 0 the quotient is 0
    uz2 = fm_min(uz2, suz);
    suz = suz - uz2;
    const double fl = fm_sel_lt(suz, 0.0000001, suz, 0.0);   // :70-73 flush: what is left goes with the drainage
    uz2 = uz2 + fl;
    suz = suz - fl;
    const double uz = uz2 * g.inv_dtt;
    // satzone_analyticalsolver (dyna_satzone_analyticalsolver.f90:60-119), ntt == 1 so qin2 == qin/ac
    const double qin_ac = qin * c.inv_ac;
    const double u = qin_ac + uz;
    const double u2 = c.q2 + u;
    const double a = FOLD ? u2 * c.dtt_inv_m2 : (u2 * g.dtt) * c.inv_m2;
    // sd/m2 enters an exponential, so its rounding error is amplified by |sd/m2| (up to a few hundred):
    // one Newton correction makes the quotient faithful to the reference's sd/m2
    double x1 = 0.0, e1;
    if (CARRY) e1 = s.e1;
    else {
        x1 = sd0 * c.inv_m2;
        x1 = fma(fma(-x1, c.m2, sd0), c.inv_m2, x1);
        e1 = LEAN_EXP(x1);
    }
    const double r = LEAN_DIV(c.q1, u2);
    const double e2 = TAB ? fm_exp_tab(-a, g.etab) : LEAN_EXP(-a);
    const double b = FOLD ? c.q1 * c.dtt_inv_m2 : (c.q1 * g.dtt) * c.inv_m2;
    const double arg_small = (e1 + b) - a * (e1 - 0.5 * b);      // :76-77 (and :84 when u2 == 0)
    const double arg_reg = fma(e1 - r, e2, r);                    // :80
    const double arg = fm_sel_lt(a, DCP_F32_1EM10, arg_small, arg_reg);
    double sd = c.m2 * (TAB ? fm_log_tab(arg, g.ltab) : LEAN_LOG(arg));
    // rare cases: sd0 > smax, exponential arguments beyond 700, an argument of the logarithm that is not a positive
    // finite number in [2^-963, 2^964) (one unsigned compare of the exponent bits; it rejects the sign bit too)
    const bool arg_ok = (uint32_t)(fm_hi(arg) - 0x03c00000) < (uint32_t)(0x7c300000 - 0x03c00000);
    o.slow = !(sd0 <= c.smax) || !(x1 <= 700.0) || !(a <= 700.0) || !arg_ok;
    double qbf = (sd - sd0) * g.inv_dtt + uz + qin_ac;            // :98
    if (CARRY) s.e1 = fm_sel_lt(sd, 0.0, 1.0, arg);               // exp(max(sd, 0)/m2)
    double exs = fm_relu(-sd);                                    // :107-112
    sd = sd + exs;                                                // max(sd, 0), exact
    const double over = fm_relu(qbf - c.q0);                      // :115-119
    exs = fma(over, g.dtt, exs);
    qbf = qbf - over;                                             // min(qbf, q0), as above
    s.sd = sd; s.suz = suz; s.srz = srz; s.qint1 = qin;
    o.qbf = qbf;
    o.exac = (exs + exus) * c.ac;
    o.ea = ea; o.uz = uz; o.exus = exus; o.sd0 = sd0;
}

// completes a lane whose lean step raised `slow`: redo the saturated zone from the saved inputs
__device__ __forceinline__ void hru_step_lean_fixup(const HruConst &c, const StepConst &g, double qin,
                                                    double qint1_old, double &sd_state, LeanOut &o) {
    const SatOut r = sat_zone_slow(c.inv_ac, c.q0, c.q1, c.q2, c.m2, c.inv_m2, c.smax, g.dtt, g.inv_dtt, g.ntt, qin,
                                   o.uz, qint1_old, o.sd0);
    sd_state = r.sd;
    o.qbf = r.qbf;
    const double ex = r.exs + o.exus;
    o.exac = (ex > 0.0) ? ex * c.ac : 0.0;
}


