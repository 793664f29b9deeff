/*
 * rr_oracle.c -- CPU restatement of the RRMPG ensemble hot path (TEST INFRASTRUCTURE).
 *
 * This file is the parity oracle for rrmpg_b200.  It is NOT product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (rrmpg_b200/) never links or calls it.
 *
 * Every function restates, statement by statement, one numba @njit function of
 * the reference (kratzert/RRMPG @ 7de78c2).  Citations are file:line under
 * /root/reference.  The numba lowering rules the restatement depends on were
 * probed against numba 0.65 (see tests/golden/make_golden.py, which also PINS this
 * oracle bit-for-bit against the live numba reference and against the
 * reference's four golden fixtures):
 *
 *   - no FMA contraction (numba fastmath is off)  -> build with -ffp-contract=off
 *   - max(0, x)  ==  (x > 0) ? x : 0.0            (max(0,NaN)=0, max(0,-0.0)=+0.0)
 *   - min(a, b)  ==  (b < a) ? b : a              (min(NaN,1)=NaN, min(1,NaN)=1)
 *   - x**2 = x*x ; x**4 = (x*x)*(x*x)             (integer literal exponents)
 *   - x**y with float y = libm pow ; np.tanh = libm tanh ; np.exp = libm exp
 *   - np.mean(v) = (sequential left-to-right sum) / n
 *   - 4/9 is the Python double constant; ((4/9 * S) / x1)
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC -pthread).
 * The *_batch drivers run independent ensemble members on all host threads
 * (pthreads); the reference itself is single-threaded (Python loop over members,
 * e.g. rrmpg/models/hbvedu.py:199-209), so the batch drivers are a generous CPU
 * baseline, not a slower one.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define NB_MAX0(x) (((x) > 0.0) ? (x) : 0.0)
#define NB_MIN(a, b) (((b) < (a)) ? (b) : (a))

/* ------------------------------------------------------------------------- */
/* ABC model: rrmpg/models/abcmodel_model.py:16-60                            */
/* qsim / storage are written with a stride (elements) so a member can be      */
/* scattered straight into column i of a C-order [T,N] array                   */
/* (rrmpg/models/abcmodel.py:174-181).  storage may be NULL.                   */
/* ------------------------------------------------------------------------- */
void oracle_abc(const double *prec, int64_t T, double initial_state,
                const double *p /* a,b,c */, double *qsim, int64_t ldq,
                double *storage, int64_t lds)
{
    const double a = p[0], b = p[1], c = p[2];
    if (T <= 0) return;
    double s_prev = initial_state;                 /* abcmodel_model.py:50 */
    qsim[0] = 0.0;                                 /* np.zeros, t=0 skipped (:53) */
    if (storage) storage[0] = initial_state;
    for (int64_t t = 1; t < T; ++t) {
        /* abcmodel_model.py:56 */
        qsim[t * ldq] = (1 - a - b) * prec[t] + c * s_prev;
        /* abcmodel_model.py:59 */
        double s = (1 - c) * s_prev + a * prec[t];
        if (storage) storage[t * lds] = s;
        s_prev = s;
    }
}

/* ------------------------------------------------------------------------- */
/* HBV-Edu: rrmpg/models/hbvedu_model.py:16-129                               */
/* month is 0-based int8 (wrapper subtracts 1, rrmpg/models/hbvedu.py:164)     */
/* params order = _dtype order (rrmpg/models/hbvedu.py:63-66):                 */
/*   T_t, DD, FC, Beta, C, PWP, K_0, K_1, K_2, K_p, L                          */
/* ------------------------------------------------------------------------- */
void oracle_hbvedu(const double *temp, const double *prec, const int8_t *month,
                   const double *PE_m, const double *T_m, int64_t T,
                   double snow_init, double soil_init, double s1_init, double s2_init,
                   const double *p, double *qsim, int64_t ld,
                   double *snow_o, double *soil_o, double *s1_o, double *s2_o)
{
    const double T_t = p[0], DD = p[1], FC = p[2], Beta = p[3], C = p[4], PWP = p[5];
    const double K_0 = p[6], K_1 = p[7], K_2 = p[8], K_p = p[9], L = p[10];
    if (T <= 0) return;
    double snow = snow_init, soil = soil_init, s1 = s1_init, s2 = s2_init; /* :78-81 */
    qsim[0] = 0.0;
    if (snow_o) { snow_o[0] = snow; soil_o[0] = soil; s1_o[0] = s1; s2_o[0] = s2; }
    for (int64_t t = 1; t < T; ++t) {              /* :84 */
        double snow_new, liquid_water;
        if (temp[t] < T_t) {                       /* :87 */
            snow_new = snow + prec[t];             /* :89 */
            liquid_water = 0.0;                    /* :91 */
        } else {
            double m = DD * (temp[t] - T_t);
            double d = snow - m;
            snow_new = NB_MAX0(d);                 /* :94 */
            liquid_water = prec[t] + NB_MIN(snow, m); /* :96 */
        }
        double prec_eff = liquid_water * pow(soil / FC, Beta);      /* :99 */
        int mi = month[t];
        double pe = (1 + C * (temp[t] - T_m[mi])) * PE_m[mi];       /* :102 */
        double ea;
        if (soil > PWP) ea = pe;                   /* :105-106 */
        else ea = pe * (soil / PWP);               /* :108 */
        double soil_new = soil + liquid_water - prec_eff - ea;      /* :111 */
        double ex = s1 - L;
        double over = NB_MAX0(ex);
        double s1_new = (s1 + prec_eff - over * K_0 - s1 * K_1 - s1 * K_p); /* :114-118 */
        double s2_new = (s2 + s1 * K_p - s2 * K_2);                 /* :121-123 */
        double q = (over * K_0 + s1_new * K_1 + s2_new * K_2);      /* :125-127 */
        qsim[t * ld] = q;
        if (snow_o) {
            snow_o[t * ld] = snow_new; soil_o[t * ld] = soil_new;
            s1_o[t * ld] = s1_new; s2_o[t * ld] = s2_new;
        }
        snow = snow_new; soil = soil_new; s1 = s1_new; s2 = s2_new;
    }
}

/* ------------------------------------------------------------------------- */
/* GR4J S-curves: rrmpg/models/gr4j_model.py:159-192 (t is an int there)       */
/* ------------------------------------------------------------------------- */
static double s_curve1(int64_t t, double x4)
{
    if (t <= 0) return 0.0;
    else if ((double)t < x4) return pow((double)t / x4, 2.5);
    else return 1.0;
}
static double s_curve2(int64_t t, double x4)
{
    if (t <= 0) return 0.0;
    else if ((double)t <= x4) return 0.5 * pow((double)t / x4, 2.5);
    else if ((double)t < 2 * x4) return 1 - 0.5 * pow(2 - (double)t / x4, 2.5);
    else return 1.0;
}

#define ORACLE_UH_CAP 4096

/* ------------------------------------------------------------------------- */
/* GR4J: rrmpg/models/gr4j_model.py:16-157.  params = x1,x2,x3,x4.             */
/* All T inputs are simulated (the reference prepends a 0 and loops 1..T).     */
/* Returns 0, or -1 if the unit hydrograph length exceeds ORACLE_UH_CAP.       */
/* ------------------------------------------------------------------------- */
int oracle_gr4j(const double *prec, const double *etp, int64_t T,
                double s_init, double r_init, const double *p,
                double *qsim, int64_t ld, double *s_o, double *r_o)
{
    const double x1 = p[0], x2 = p[1], x3 = p[2], x4 = p[3];
    double S = s_init * x1;                        /* :64 */
    double R = r_init * x3;                        /* :65 */
    int64_t n1 = (int64_t)ceil(x4);                /* :68 */
    int64_t n2 = (int64_t)ceil(2 * x4 + 1);        /* :69 */
    if (n1 < 1 || n2 < 1 || n1 > ORACLE_UH_CAP || n2 > ORACLE_UH_CAP) return -1;
    double *o1 = (double *)calloc((size_t)(2 * (n1 + n2)), sizeof(double));
    double *o2 = o1 + n1, *uh1 = o2 + n2, *uh2 = uh1 + n1;
    for (int64_t j = 1; j <= n1; ++j) o1[j - 1] = s_curve1(j, x4) - s_curve1(j - 1, x4); /* :75-76 */
    for (int64_t j = 1; j <= n2; ++j) o2[j - 1] = s_curve2(j, x4) - s_curve2(j - 1, x4); /* :78-79 */

    for (int64_t t = 0; t < T; ++t) {              /* :86 (index shifted by the prepended 0) */
        double p_n, pe_n, p_s, e_s;
        if (prec[t] >= etp[t]) {                   /* :89 */
            p_n = prec[t] - etp[t];
            pe_n = 0.0;
            double sr = S / x1;
            double th = tanh(p_n / x1);
            p_s = ((x1 * (1 - sr * sr) * th) / (1 + sr * th));       /* :95-96 */
            e_s = 0.0;
        } else {
            p_n = 0.0;
            pe_n = etp[t] - prec[t];
            double sr = S / x1;
            double th = tanh(pe_n / x1);
            e_s = ((S * (2 - sr) * th) / (1 + (1 - sr) * th));       /* :107-108 */
            p_s = 0.0;
        }
        S = S - e_s + p_s;                         /* :114 */
        double u = 4.0 / 9.0 * S / x1;
        double u2 = u * u;
        double perc = S * (1 - pow(1 + u2 * u2, -0.25));             /* :117 */
        S = S - perc;                              /* :120 */
        double p_r = perc + (p_n - p_s);           /* :123 */
        double p1 = 0.9 * p_r;                     /* :126 */
        double p2 = 0.1 * p_r;                     /* :127 */
        for (int64_t j = 0; j < n1 - 1; ++j) uh1[j] = uh1[j + 1] + o1[j] * p1; /* :130-131 */
        uh1[n1 - 1] = o1[n1 - 1] * p1;             /* :132 */
        for (int64_t j = 0; j < n2 - 1; ++j) uh2[j] = uh2[j + 1] + o2[j] * p2; /* :134-135 */
        uh2[n2 - 1] = o2[n2 - 1] * p2;             /* :136 */
        double F = x2 * pow(R / x3, 3.5);          /* :139 */
        double rr = R + uh1[0] + F;
        R = NB_MAX0(rr);                           /* :142 */
        double v = R / x3;
        double v2 = v * v;
        double q_r = R * (1 - pow(1 + v2 * v2, -0.25));              /* :145 */
        R = R - q_r;                               /* :148 */
        double qd = uh2[0] + F;
        double q_d = NB_MAX0(qd);                  /* :151 */
        qsim[t * ld] = q_r + q_d;                  /* :154 */
        if (s_o) { s_o[t * ld] = S; r_o[t * ld] = R; }
    }
    free(o1);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Cemaneige: rrmpg/models/cemaneige_model.py:16-127.  params = CTG, Kf.       */
/* prec/mean_temp/frac are C-order [T,L].  lw is [T,L] scratch (caller owned). */
/* G_o / eTG_o (nullable) are [T,L,N]-style: element (t,l) at (t*L+l)*ldg.     */
/* outflow is written with stride ld.                                          */
/* ------------------------------------------------------------------------- */
void oracle_cemaneige(const double *prec, const double *mean_temp, const double *frac,
                      int64_t T, int64_t L, double snow_pack_init, double thermal_state_init,
                      const double *p, double *lw, double *outflow, int64_t ld,
                      double *G_o, double *eTG_o, int64_t ldg)
{
    const double CTG = p[0], Kf = p[1];
    for (int64_t l = 0; l < L; ++l) {              /* :73 (prange is inert under plain @njit) */
        /* :76-80  snow = prec*frac ; rain = prec - snow ; G_tresh = 0.9*365.25*mean(snow) */
        double acc = 0.0;
        for (int64_t t = 0; t < T; ++t) acc += prec[t * L + l] * frac[t * L + l];
        double G_tresh = 0.9 * 365.25 * (acc / (double)T);
        double G = 0.0, eTG = 0.0;
        for (int64_t t = 0; t < T; ++t) {
            double snow = prec[t * L + l] * frac[t * L + l];
            double rain = prec[t * L + l] - snow;
            double Tm = mean_temp[t * L + l];
            if (t == 0) G = snow_pack_init;        /* :85-86 */
            else G = G + snow;                     /* :88 */
            if (t == 0) eTG = thermal_state_init;  /* :91-92 */
            else eTG = CTG * eTG + (1 - CTG) * Tm; /* :94 */
            if (eTG > 0) eTG = 0.0;                /* :95-96 */
            double pot_melt;
            if (eTG == 0 && Tm > 0) {              /* :99 */
                pot_melt = Kf * Tm;                /* :100 */
                if (pot_melt > G) pot_melt = G;    /* :103-104 */
            } else pot_melt = 0.0;
            double G_ratio;
            if (G < G_tresh) G_ratio = G / G_tresh; /* :109-110 */
            else G_ratio = 1.0;
            double melt = (0.9 * G_ratio + 0.1) * pot_melt;          /* :115 */
            G = G - melt;                          /* :118 */
            lw[t * L + l] = rain + melt;           /* :121 */
            if (G_o) { G_o[(t * L + l) * ldg] = G; eTG_o[(t * L + l) * ldg] = eTG; }
        }
    }
    for (int64_t t = 0; t < T; ++t) {              /* :124-125  np.mean over layers */
        double acc = 0.0;
        for (int64_t l = 0; l < L; ++l) acc += lw[t * L + l];
        outflow[t * ld] = acc / (double)L;
    }
}

/* ------------------------------------------------------------------------- */
/* CemaneigeGR4J: rrmpg/models/cemaneigegr4j_model.py:17-64                   */
/* params = CTG, Kf, x1, x2, x3, x4  (rrmpg/models/cemaneigegr4j.py:67-72)     */
/* ------------------------------------------------------------------------- */
int oracle_cemaneigegr4j(const double *prec, const double *mean_temp, const double *etp,
                         const double *frac, int64_t T, int64_t L,
                         double snow_pack_init, double thermal_state_init,
                         double s_init, double r_init, const double *p,
                         double *lw /* [T,L] */, double *liq /* [T] */,
                         double *qsim, int64_t ld, double *G_o, double *eTG_o, int64_t ldg,
                         double *s_o, double *r_o)
{
    oracle_cemaneige(prec, mean_temp, frac, T, L, snow_pack_init, thermal_state_init,
                     p, lw, liq, 1, G_o, eTG_o, ldg);                /* :57 */
    return oracle_gr4j(liq, etp, T, s_init, r_init, p + 2, qsim, ld, s_o, r_o); /* :62 */
}

/* ------------------------------------------------------------------------- */
/* Cemaneige with SWE-SCA hysteresis: rrmpg/models/cemaneigehyst_model.py:5-166 */
/* params (first four fields of the Hyst records) = CTG, Kf, Thacc, Rsp          */
/* numba lowers max(a,b) to (b > a) ? b : a and min(a,b) to (b < a) ? b : a.     */
/* Quirk kept: at t = 0 the accumulation branch reads sca[t-1] = sca[T-1], which */
/* still holds its np.zeros value (or sca_init itself when T == 1), :126.        */
/* sca_o / G_o / eTG_o are [T,L,N]-style (stride ldg); rain_o [T,L] dense.       */
/* ------------------------------------------------------------------------- */
#define NB_MAX(a, b) (((b) > (a)) ? (b) : (a))
void oracle_cemaneigehyst(const double *prec, const double *mean_temp, const double *frac,
                          int64_t T, int64_t L, double snow_pack_init, double thermal_state_init,
                          double sca_init, const double *p, double *lw, double *outflow, int64_t ld,
                          double *G_o, double *eTG_o, double *sca_o, int64_t ldg, double *rain_o)
{
    const double CTG = p[0], Kf = p[1], Thacc = p[2], Rsp = p[3];
    for (int64_t l = 0; l < L; ++l) {
        double acc = 0.0;
        for (int64_t t = 0; t < T; ++t) acc += prec[t * L + l] * frac[t * L + l];
        const double Psolannual = 365.25 * (acc / (double)T);          /* :102 */
        double G = 0.0, eTG = 0.0, swe_max = 0.0, Thmax = 0.0;
        double sca_prev = (T == 1) ? sca_init : 0.0;                     /* sca[-1] at t = 0 */
        for (int64_t t = 0; t < T; ++t) {
            const double snow = prec[t * L + l] * frac[t * L + l];
            const double rain = prec[t * L + l] - snow;                  /* :99 */
            const double Tm = mean_temp[t * L + l];
            double sca;
            if (t == 0) { G = snow_pack_init; sca = sca_init; }          /* :107-109 */
            else G = G + snow;                                           /* :111 */
            if (t == 0) eTG = thermal_state_init;                        /* :114-115 */
            else eTG = CTG * eTG + (1 - CTG) * Tm;                       /* :117 */
            if (eTG > 0) eTG = 0.0;                                      /* :118-119 */
            double pot_melt;
            if (eTG == 0 && Tm > 0) {                                    /* :122 */
                pot_melt = Kf * Tm;
                if (pot_melt > G) pot_melt = G;
            } else pot_melt = 0.0;
            const double snow_balance = snow - pot_melt;                 /* :131 */
            if (snow_balance >= 0) {                                     /* :133 */
                sca = sca_prev + snow_balance / Thacc;                   /* :135 */
                swe_max = NB_MAX(swe_max, G);                            /* :136 */
            } else {
                const double Thmelt = Psolannual * Rsp;                  /* :139 */
                if (swe_max > Thmelt) Thmax = Thmelt;                    /* :142-145 */
                else Thmax = swe_max;
                if (Thmax > 0) sca = G / Thmax;                          /* :148-151 */
                else sca = 0.0;
            }
            { const double m = NB_MAX(sca, 0.0); sca = NB_MIN(m, 1.0); } /* :154 */
            double melt = (0.9 * sca + 0.1) * pot_melt;                  /* :157 */
            melt = NB_MIN(melt, G);                                      /* :160 */
            G = G - melt;                                                /* :163 */
            if (G == 0) swe_max = 0.0;                                   /* :166-167 */
            lw[t * L + l] = rain + melt;                                 /* :171 */
            if (G_o) { G_o[(t * L + l) * ldg] = G; eTG_o[(t * L + l) * ldg] = eTG; sca_o[(t * L + l) * ldg] = sca; }
            if (rain_o) rain_o[t * L + l] = rain;
            sca_prev = sca;
        }
    }
    (void)NB_MAX(0.0, 0.0);
    for (int64_t t = 0; t < T; ++t) {
        double acc = 0.0;
        for (int64_t l = 0; l < L; ++l) acc += lw[t * L + l];
        outflow[t * ld] = acc / (double)L;
    }
}

/* ------------------------------------------------------------------------- */
/* Snow + ice + GR4J couplings:                                                */
/*   run_cemaneigegr4jice       rrmpg/models/cemaneigegr4jice_model.py:16-93    */
/*   run_cemaneigehystgr4j      rrmpg/models/cemaneigehystgr4j_model.py:17-79   */
/*   run_cemaneigehystgr4jice   rrmpg/models/cemaneigehystgr4jice_model.py:18-104 */
/*   run_icemelt                rrmpg/models/icemelt_model.py:16-65             */
/* One restatement with two switches.  Record layouts:                          */
/*   !hyst, ice : CTG Kf x1 x2 x3 x4 DDF          hyst, !ice: CTG Kf Thacc Rsp x1..x4 */
/*    hyst, ice : CTG Kf Thacc Rsp x1 x2 x3 x4 DDF                                 */
/* scratch: lw [T,L], Gs [T,L] (final snow pack per step), es [T,L], liq [T], snowmelt [T] */
/* ------------------------------------------------------------------------- */
int oracle_snowice_gr4j(int hyst, int ice, const double *prec, const double *mean_temp,
                        const double *etp, const double *frac_ice, const double *frac, int64_t T,
                        int64_t L, const double *inits /* g0,e0,sca0,s,r */, const double *p,
                        double *scratch, double *qsim, int64_t ld, double *G_o, double *eTG_o,
                        double *sca_o, int64_t ldg, double *s_o, double *r_o, double *icemelt_o,
                        double *snowmelt_o, double *rain_o)
{
    double *lw = scratch, *Gs = lw + T * L, *es = Gs + T * L, *ss = es + T * L;
    double *liq = ss + T * L, *sm = liq + T;
    if (hyst) oracle_cemaneigehyst(prec, mean_temp, frac, T, L, inits[0], inits[1], inits[2], p, lw, sm, 1,
                                   Gs, es, ss, 1, rain_o);
    else oracle_cemaneige(prec, mean_temp, frac, T, L, inits[0], inits[1], p, lw, sm, 1, Gs, es, 1);
    const int goff = hyst ? 4 : 2;
    for (int64_t t = 0; t < T; ++t) {
        double total = 0.0;
        if (ice) {
            const double ddf = p[goff + 4];
            for (int64_t l = 0; l < L; ++l) {                        /* icemelt_model.py:52-62 */
                double melt = ddf * (mean_temp[t * L + l] - 0);
                if (melt < 0) melt = 0.0;
                const double w = (Gs[t * L + l] > 1) ? 0.0 : melt;
                total += w * frac_ice[l];                            /* np.sum(icemelt * frac_ice, axis=1) */
            }
            liq[t] = sm[t] + total;                                  /* cemaneigegr4jice_model.py:87 */
        } else liq[t] = sm[t];
        if (icemelt_o) icemelt_o[t * ld] = total;
        if (snowmelt_o) snowmelt_o[t * ld] = sm[t];
        if (G_o) for (int64_t l = 0; l < L; ++l) {
            G_o[(t * L + l) * ldg] = Gs[t * L + l]; eTG_o[(t * L + l) * ldg] = es[t * L + l];
            if (sca_o) sca_o[(t * L + l) * ldg] = hyst ? ss[t * L + l] : 0.0;
        }
    }
    return oracle_gr4j(liq, etp, T, inits[3], inits[4], p + goff, qsim, ld, s_o, r_o);
}

/* ------------------------------------------------------------------------- */
/* Forcing preprocessors: rrmpg/models/cemaneige_utils.py                     */
/* ------------------------------------------------------------------------- */
/* extrapolate_precipitation :101-158 */
void oracle_extrapolate_precipitation(const double *prec, int64_t T, const double *alt, int64_t L,
                                      double station, double *out /* [T,L] */)
{
    const double beta_altitude = 0.0004;           /* :127 */
    const double z_thresh = 4000;                  /* :130 */
    for (int64_t l = 0; l < L; ++l) {
        if (alt[l] <= z_thresh) {                  /* :143 */
            double f = exp((alt[l] - station) * beta_altitude);
            for (int64_t t = 0; t < T; ++t) out[t * L + l] = prec[t] * f;
        } else if (station <= z_thresh) {          /* :150 */
            double f = exp((z_thresh - station) * beta_altitude);
            for (int64_t t = 0; t < T; ++t) out[t * L + l] = prec[t] * f;
        } else {
            for (int64_t t = 0; t < T; ++t) out[t * L + l] = prec[t]; /* :156 */
        }
    }
}

/* extrapolate_temperature :161-208 */
void oracle_extrapolate_temperature(const double *tmin, const double *tmean, const double *tmax,
                                    int64_t T, const double *alt, int64_t L, double station,
                                    double *omin, double *omean, double *omax)
{
    const double theta_temp = -0.0065;             /* :188 */
    for (int64_t l = 0; l < L; ++l) {
        double d = (alt[l] - station) * theta_temp; /* :201 */
        for (int64_t t = 0; t < T; ++t) {
            omin[t * L + l] = tmin[t] + d;
            omean[t * L + l] = tmean[t] + d;
            omax[t * L + l] = tmax[t] + d;
        }
    }
}

/* calculate_solid_fraction :16-98 (prec is only used for its shape there) */
void oracle_solid_fraction(const double *alt, int64_t L, const double *tmean, const double *tmin,
                           const double *tmax, int64_t T, double *out /* [T,L] */)
{
    const double z_thresh = 1500;                  /* :51 */
    for (int64_t l = 0; l < L; ++l) {
        if (alt[l] < z_thresh) {                   /* :64 */
            for (int64_t t = 0; t < T; ++t) {
                double mx = tmax[t * L + l], mn = tmin[t * L + l];
                if (mx <= 0) out[t * L + l] = 1.0;
                else if (mn >= 0) out[t * L + l] = 0.0;
                else out[t * L + l] = 1 - (mx / (mx - mn));          /* :77-79 */
            }
        } else {
            for (int64_t t = 0; t < T; ++t) {
                double m = tmean[t * L + l];
                if (m >= 3) out[t * L + l] = 0.0;
                else if (m <= 0) out[t * L + l] = 1.0;
                else out[t * L + l] = 1 - (m + 1) / 4;               /* :96 */
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Ensemble drivers = the member loops of the wrappers                         */
/* (abcmodel.py:174-181, hbvedu.py:199-209, gr4j.py:169-178 minus its early    */
/* return bug at :178, cemaneige.py:227-240, cemaneigegr4j.py:249-268),        */
/* spread over host threads (pthreads, contiguous member blocks).  params is   */
/* the packed AoS record array [N,k].  Outputs are C-order [T,N] (storages     */
/* [T,L,N]); storages are nullable.                                            */
/* ------------------------------------------------------------------------- */
int oracle_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        int c = CPU_COUNT(&set);
        if (c > 0) n = c;
    }
    return n > 0 ? (int)n : 1;
}

typedef void (*member_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { member_fn fn; void *ctx; int64_t lo, hi; } span_t;
static void *span_main(void *arg)
{
    span_t *s = (span_t *)arg;
    s->fn(s->lo, s->hi, s->ctx);
    return NULL;
}
static void parallel_members(int64_t N, int nthreads, member_fn fn, void *ctx)
{
    if (nthreads <= 0) nthreads = oracle_num_threads();
    if (nthreads > N) nthreads = (int)(N > 0 ? N : 1);
    if (nthreads <= 1) { fn(0, N, ctx); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    span_t *sp = (span_t *)malloc(sizeof(span_t) * (size_t)nthreads);
    for (int k = 0; k < nthreads; ++k) {
        sp[k].fn = fn; sp[k].ctx = ctx;
        sp[k].lo = (N * k / nthreads) & ~(int64_t)7; sp[k].hi = (k + 1 == nthreads) ? N : ((N * (k + 1) / nthreads) & ~(int64_t)7);
        pthread_create(&th[k], NULL, span_main, &sp[k]);
    }
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
    free(th); free(sp);
}

typedef struct {
    const double *a0, *a1, *a2, *a3, *a4; const int8_t *m;
    int64_t T, L, N, pstride; const double *inits; const double *params;
    double *o0, *o1, *o2, *o3, *o4; int rc;
} bctx_t;

/* Members are computed MB at a time into contiguous thread-local series and then written to
 * the [rows, N] result as MB consecutive doubles per row: one cache line per row instead of
 * one strided store per member-timestep (the reference's `qsim[:, i] = ...` column scatter,
 * rrmpg/models/hbvedu.py:202, pays that stride; this baseline does not). */
#define MB 8
static void scatter_rows(const double *local, int64_t rows, int64_t nb, double *dst, int64_t N, int64_t i0)
{
    if (!dst) return;
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t b = 0; b < nb; ++b) dst[r * N + i0 + b] = local[b * rows + r];
}
static double *local_alloc(int64_t rows, int nbuf)
{
    return (double *)malloc(sizeof(double) * (size_t)(rows > 0 ? rows : 1) * MB * (size_t)nbuf);
}

static void abc_span(int64_t lo, int64_t hi, void *v)
{
    bctx_t *c = (bctx_t *)v;
    const int64_t T = c->T;
    double *buf = local_alloc(T, 2), *q = buf, *s = buf + MB * T;
    for (int64_t i0 = lo; i0 < hi; i0 += MB) {
        const int64_t nb = (hi - i0 < MB) ? hi - i0 : MB;
        for (int64_t b = 0; b < nb; ++b)
            oracle_abc(c->a0, T, c->inits[0], c->params + 3 * (i0 + b), q + b * T, 1,
                       c->o1 ? s + b * T : NULL, 1);
        scatter_rows(q, T, nb, c->o0, c->N, i0);
        scatter_rows(s, T, nb, c->o1, c->N, i0);
    }
    free(buf);
}
void oracle_abc_batch(const double *prec, int64_t T, double s0, const double *params, int64_t N,
                      double *qsim, double *storage, int nthreads)
{
    bctx_t c; memset(&c, 0, sizeof(c));
    c.a0 = prec; c.T = T; c.N = N; c.inits = &s0; c.params = params; c.o0 = qsim; c.o1 = storage;
    parallel_members(N, nthreads, abc_span, &c);
}

static void hbv_span(int64_t lo, int64_t hi, void *v)
{
    bctx_t *c = (bctx_t *)v;
    const int64_t T = c->T;
    double *buf = local_alloc(T, 5);
    double *o[5];
    for (int k = 0; k < 5; ++k) o[k] = buf + (int64_t)k * MB * T;
    for (int64_t i0 = lo; i0 < hi; i0 += MB) {
        const int64_t nb = (hi - i0 < MB) ? hi - i0 : MB;
        for (int64_t b = 0; b < nb; ++b)
            oracle_hbvedu(c->a0, c->a1, c->m, c->a2, c->a3, T, c->inits[0], c->inits[1],
                          c->inits[2], c->inits[3], c->params + 11 * (i0 + b), o[0] + b * T, 1,
                          c->o1 ? o[1] + b * T : NULL, c->o1 ? o[2] + b * T : NULL,
                          c->o1 ? o[3] + b * T : NULL, c->o1 ? o[4] + b * T : NULL);
        scatter_rows(o[0], T, nb, c->o0, c->N, i0);
        scatter_rows(o[1], T, nb, c->o1, c->N, i0);
        scatter_rows(o[2], T, nb, c->o2, c->N, i0);
        scatter_rows(o[3], T, nb, c->o3, c->N, i0);
        scatter_rows(o[4], T, nb, c->o4, c->N, i0);
    }
    free(buf);
}
void oracle_hbvedu_batch(const double *temp, const double *prec, const int8_t *month,
                         const double *PE_m, const double *T_m, int64_t T, const double *inits,
                         const double *params, int64_t N, double *qsim, double *snow, double *soil,
                         double *s1, double *s2, int nthreads)
{
    bctx_t c; memset(&c, 0, sizeof(c));
    c.a0 = temp; c.a1 = prec; c.m = month; c.a2 = PE_m; c.a3 = T_m; c.T = T; c.N = N;
    c.inits = inits; c.params = params; c.o0 = qsim; c.o1 = snow; c.o2 = soil; c.o3 = s1; c.o4 = s2;
    parallel_members(N, nthreads, hbv_span, &c);
}

static void gr4j_span(int64_t lo, int64_t hi, void *v)
{
    bctx_t *c = (bctx_t *)v;
    const int64_t T = c->T;
    double *buf = local_alloc(T, 3);
    double *q = buf, *s = buf + MB * T, *r = buf + 2 * MB * T;
    for (int64_t i0 = lo; i0 < hi; i0 += MB) {
        const int64_t nb = (hi - i0 < MB) ? hi - i0 : MB;
        for (int64_t b = 0; b < nb; ++b) {
            int rc = oracle_gr4j(c->a0, c->a1, T, c->inits[0], c->inits[1], c->params + 4 * (i0 + b),
                                 q + b * T, 1, c->o1 ? s + b * T : NULL, c->o1 ? r + b * T : NULL);
            if (rc) __atomic_store_n(&c->rc, rc, __ATOMIC_RELAXED);
        }
        scatter_rows(q, T, nb, c->o0, c->N, i0);
        scatter_rows(s, T, nb, c->o1, c->N, i0);
        scatter_rows(r, T, nb, c->o2, c->N, i0);
    }
    free(buf);
}
int oracle_gr4j_batch(const double *prec, const double *etp, int64_t T, double s_init, double r_init,
                      const double *params, int64_t N, double *qsim, double *s_o, double *r_o,
                      int nthreads)
{
    double inits[2] = { s_init, r_init };
    bctx_t c; memset(&c, 0, sizeof(c));
    c.a0 = prec; c.a1 = etp; c.T = T; c.N = N; c.inits = inits; c.params = params;
    c.o0 = qsim; c.o1 = s_o; c.o2 = r_o;
    parallel_members(N, nthreads, gr4j_span, &c);
    return c.rc;
}

static void cema_span(int64_t lo, int64_t hi, void *v)
{
    bctx_t *c = (bctx_t *)v;
    const int64_t T = c->T, L = c->L, TL = T * L;
    double *lw = (double *)malloc(sizeof(double) * (size_t)(TL > 0 ? TL : 1));
    double *q = local_alloc(T, 1), *st = local_alloc(TL, 2), *G = st, *E = st + MB * TL;
    for (int64_t i0 = lo; i0 < hi; i0 += MB) {
        const int64_t nb = (hi - i0 < MB) ? hi - i0 : MB;
        for (int64_t b = 0; b < nb; ++b)
            oracle_cemaneige(c->a0, c->a1, c->a2, T, L, c->inits[0], c->inits[1],
                             c->params + c->pstride * (i0 + b), lw, q + b * T, 1,
                             c->o1 ? G + b * TL : NULL, c->o1 ? E + b * TL : NULL, 1);
        scatter_rows(q, T, nb, c->o0, c->N, i0);
        scatter_rows(G, TL, nb, c->o1, c->N, i0);
        scatter_rows(E, TL, nb, c->o2, c->N, i0);
    }
    free(lw); free(q); free(st);
}
void oracle_cemaneige_batch(const double *prec, const double *mean_temp, const double *frac,
                            int64_t T, int64_t L, double g0, double e0, const double *params,
                            int64_t pstride, int64_t N, double *outflow, double *G_o, double *eTG_o,
                            int nthreads)
{
    double inits[2] = { g0, e0 };
    bctx_t c; memset(&c, 0, sizeof(c));
    c.a0 = prec; c.a1 = mean_temp; c.a2 = frac; c.T = T; c.L = L; c.N = N; c.pstride = pstride;
    c.inits = inits; c.params = params; c.o0 = outflow; c.o1 = G_o; c.o2 = eTG_o;
    parallel_members(N, nthreads, cema_span, &c);
}

static void cg_span(int64_t lo, int64_t hi, void *v)
{
    bctx_t *c = (bctx_t *)v;
    const int64_t T = c->T, L = c->L, TL = T * L;
    double *lw = (double *)malloc(sizeof(double) * (size_t)(TL > 0 ? TL : 1));
    double *liq = (double *)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
    double *buf = local_alloc(T, 3), *st = local_alloc(TL, 2);
    double *q = buf, *s = buf + MB * T, *r = buf + 2 * MB * T, *G = st, *E = st + MB * TL;
    for (int64_t i0 = lo; i0 < hi; i0 += MB) {
        const int64_t nb = (hi - i0 < MB) ? hi - i0 : MB;
        for (int64_t b = 0; b < nb; ++b) {
            int rc = oracle_cemaneigegr4j(c->a0, c->a1, c->a3, c->a2, T, L, c->inits[0], c->inits[1],
                                          c->inits[2], c->inits[3], c->params + 6 * (i0 + b), lw, liq,
                                          q + b * T, 1, c->o1 ? G + b * TL : NULL,
                                          c->o1 ? E + b * TL : NULL, 1, c->o1 ? s + b * T : NULL,
                                          c->o1 ? r + b * T : NULL);
            if (rc) __atomic_store_n(&c->rc, rc, __ATOMIC_RELAXED);
        }
        scatter_rows(q, T, nb, c->o0, c->N, i0);
        scatter_rows(G, TL, nb, c->o1, c->N, i0);
        scatter_rows(E, TL, nb, c->o2, c->N, i0);
        scatter_rows(s, T, nb, c->o3, c->N, i0);
        scatter_rows(r, T, nb, c->o4, c->N, i0);
    }
    free(lw); free(liq); free(buf); free(st);
}
int oracle_cemaneigegr4j_batch(const double *prec, const double *mean_temp, const double *etp,
                               const double *frac, int64_t T, int64_t L, const double *inits,
                               const double *params, int64_t N, double *qsim, double *G_o,
                               double *eTG_o, double *s_o, double *r_o, int nthreads)
{
    bctx_t c; memset(&c, 0, sizeof(c));
    c.a0 = prec; c.a1 = mean_temp; c.a2 = frac; c.a3 = etp; c.T = T; c.L = L; c.N = N;
    c.inits = inits; c.params = params; c.o0 = qsim; c.o1 = G_o; c.o2 = eTG_o; c.o3 = s_o; c.o4 = r_o;
    parallel_members(N, nthreads, cg_span, &c);
    return c.rc;
}

typedef struct {
    int hyst, ice; const double *prec, *mt, *etp, *fice, *frac; int64_t T, L, N, k; const double *inits, *params;
    double *q, *G, *E, *S, *s, *r, *im, *smelt; int rc;
} sictx_t;
static void snowice_span(int64_t lo, int64_t hi, void *v)
{
    sictx_t *c = (sictx_t *)v;
    const int64_t T = c->T, L = c->L, TL = T * L;
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(4 * TL + 2 * T + 8));
    for (int64_t i = lo; i < hi; ++i) {
        int rc = oracle_snowice_gr4j(c->hyst, c->ice, c->prec, c->mt, c->etp, c->fice, c->frac, T, L, c->inits,
                                     c->params + c->k * i, scratch, c->q + i, c->N,
                                     c->G ? c->G + i : NULL, c->E ? c->E + i : NULL,
                                     c->S ? c->S + i : NULL, c->N, c->s ? c->s + i : NULL,
                                     c->r ? c->r + i : NULL, c->im ? c->im + i : NULL,
                                     c->smelt ? c->smelt + i : NULL, NULL);
        if (rc) __atomic_store_n(&c->rc, rc, __ATOMIC_RELAXED);
    }
    free(scratch);
}
/* member loops of cemaneigegr4jice.py:262-284, cemaneigehystgr4j.py:262-286, cemaneigehystgr4jice.py:276-304 */
int oracle_snowice_gr4j_batch(int hyst, int ice, const double *prec, const double *mean_temp, const double *etp,
                              const double *frac_ice, const double *frac, int64_t T, int64_t L,
                              const double *inits, const double *params, int64_t N, double *qsim, double *G_o,
                              double *eTG_o, double *sca_o, double *s_o, double *r_o, double *icemelt_o,
                              double *snowmelt_o, int nthreads)
{
    sictx_t c; memset(&c, 0, sizeof(c));
    c.hyst = hyst; c.ice = ice; c.prec = prec; c.mt = mean_temp; c.etp = etp; c.fice = frac_ice; c.frac = frac;
    c.T = T; c.L = L; c.N = N; c.k = 6 + (hyst ? 2 : 0) + (ice ? 1 : 0); c.inits = inits; c.params = params;
    c.q = qsim; c.G = G_o; c.E = eTG_o; c.S = sca_o; c.s = s_o; c.r = r_o; c.im = icemelt_o; c.smelt = snowmelt_o;
    parallel_members(N, nthreads, snowice_span, &c);
    return c.rc;
}

/* calc_mse per ensemble column: rrmpg/tools/monte_carlo.py:70-71 with            */
/* rrmpg/utils/metrics.py:110-136 (np.mean((obs-sim)**2)); the sum here is        */
/* sequential, numpy's is pairwise -> compare with a tolerance, not bit-exact.    */
void oracle_mse_columns(const double *qobs, const double *qsim, int64_t T, int64_t N, double *mse)
{
    for (int64_t i = 0; i < N; ++i) {
        double acc = 0.0;
        for (int64_t t = 0; t < T; ++t) {
            double d = qobs[t] - qsim[t * N + i];
            acc += d * d;
        }
        mse[i] = acc / (double)T;
    }
}

/* ------------------------------------------------------------------------------------------
 * Test helper, not part of the reference: the division by a loop-invariant divisor that the CUDA
 * kernels use in place of IEEE division where bit-exactness is required (rr_common.cuh:
 * div_by_invariant; rr_cemaneige.cuh: the contract step evaluates G / G_tresh with it).  With
 * y = RN(1/b): q = RN(a y); r = fma(-b, q, a); q = fma(r, y, q); r = fma(-b, q, a); q = fma(r, y, q).
 * Returns how many of n pseudo-random operand pairs give a result different from a / b.
 *   mode 0: b log-uniform in [2^-60, 2^60], a log-uniform in [2^-960, 2^960)   (the range test of the kernel)
 *   mode 1: the snow-cover ratio: b log-uniform in [2^-20, 2^20], a = u b with u in (0, 1]
 *   mode 2: adversarial: a = k b (+- a few ulp) for small integers k and their reciprocals
 * fma() is exact (hardware FMA or glibc's software fma), the TU is built with -ffp-contract=off.
 * ------------------------------------------------------------------------------------------ */
static uint64_t oracle_rng(uint64_t *s)
{
    uint64_t x = *s;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return *s = x;
}
static double oracle_u01(uint64_t *s) { return (double)(oracle_rng(s) >> 11) * 0x1.0p-53; }
static double oracle_invariant_div(double a, double b, double y)
{
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    q = fma(r, y, q);
    return q;
}
int64_t oracle_check_invariant_division(int64_t n, uint64_t seed, int mode)
{
    uint64_t s = seed ? seed : 88172645463325252ULL;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        double a, b;
        if (mode == 0) {
            b = ldexp(1.0 + oracle_u01(&s), (int)(oracle_rng(&s) % 120) - 60);
            a = ldexp(1.0 + oracle_u01(&s), (int)(oracle_rng(&s) % 1920) - 960);
        } else if (mode == 1) {
            b = ldexp(1.0 + oracle_u01(&s), (int)(oracle_rng(&s) % 40) - 20);
            a = b * (1.0 - oracle_u01(&s));
        } else {
            b = ldexp(1.0 + oracle_u01(&s), (int)(oracle_rng(&s) % 120) - 60);
            const double k = (double)(1 + oracle_rng(&s) % 64);
            a = (oracle_rng(&s) & 1) ? b * k : b / k;
            const int nudge = (int)(oracle_rng(&s) % 5) - 2;
            for (int j = 0; j < (nudge < 0 ? -nudge : nudge); ++j) a = nextafter(a, nudge < 0 ? 0.0 : INFINITY);
        }
        const double y = 1.0 / b;
        if (oracle_invariant_div(a, b, y) != a / b) ++bad;
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------
 * Test helper, not part of the reference: the CONTRACT step of the CUDA snow routine
 * (rr_cemaneige.cuh, cema_kernel with CONTRACT = true) restated on the CPU with exact fma(), so that its
 * bit-identity with run_cemaneige (oracle_cemaneige above) can be checked over long series and extreme
 * in-contract parameters without a GPU: G / G_tresh evaluated unconditionally by the unchecked
 * division-by-an-invariant sequence, potential melt and ratio by selects, mean over the layers by IEEE
 * division.  outflow [T][N]; G, eTG [T][L][N] (nullable).  L <= 16.
 * ------------------------------------------------------------------------------------------ */
void oracle_cemaneige_contract_twin(const double *prec, const double *mean_temp, const double *frac, int64_t T,
                                    int64_t L, double g0, double e0, const double *params, int64_t pstride,
                                    int64_t N, double *outflow, double *Gout, double *Eout)
{
    double gt[16], inv_gt[16];
    for (int64_t l = 0; l < L; ++l) {           /* the packer: snow = prec * frac, sequential mean */
        double acc = 0.0;
        for (int64_t t = 0; t < T; ++t) acc += prec[t * L + l] * frac[t * L + l];
        gt[l] = 0.9 * 365.25 * (acc / (double)T);
        inv_gt[l] = 1.0 / gt[l];
    }
    for (int64_t i = 0; i < N; ++i) {
        const double CTG = params[i * pstride], Kf = params[i * pstride + 1];
        const double omCTG = 1 - CTG;
        double G[16], eTG[16];
        for (int64_t t = 0; t < T; ++t) {
            double lw_sum = 0.0;
            for (int64_t l = 0; l < L; ++l) {
                const double p = prec[t * L + l];
                const double snow = p * frac[t * L + l], rain = p - snow, Tm = mean_temp[t * L + l];
                double g = (t == 0) ? g0 : G[l] + snow;
                double e = (t == 0) ? e0 : CTG * eTG[l] + omCTG * Tm;
                e = (e > 0) ? 0.0 : e;
                const double kt = Kf * Tm;
                const double capped = (kt > g) ? g : kt;
                const double pot = (e == 0 && Tm > 0) ? capped : 0.0;
                const double q = oracle_invariant_div(g, gt[l], inv_gt[l]);
                const double ratio = (g < gt[l]) ? q : 1.0;
                const double melt = (0.9 * ratio + 0.1) * pot;
                g = g - melt;
                lw_sum += rain + melt;
                G[l] = g;
                eTG[l] = e;
                if (Gout) Gout[(t * L + l) * N + i] = g;
                if (Eout) Eout[(t * L + l) * N + i] = e;
            }
            outflow[t * N + i] = (L == 1) ? lw_sum : lw_sum / (double)L;
        }
    }
}
