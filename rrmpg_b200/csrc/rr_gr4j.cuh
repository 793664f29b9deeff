// rr_gr4j.cuh -- per-member GR4J state machine shared by the GR4J and CemaneigeGR4J kernels.
// Restates run_gr4j (rrmpg/models/gr4j_model.py:16-157) and its S-curves (:159-192).
//
// The reference sizes the two unit-hydrograph buffers per member (num_uh1 = ceil(x4),
// num_uh2 = ceil(2 x4 + 1), :68-69).  Here they are compile-time-sized register arrays
// (C1, C2 >= the largest length in the batch, chosen by the host from max x4): ordinates
// beyond a member's own length are exactly 0 (S-curve differences 1 - 1) and the member's last
// real slot is written without reading its padded neighbour, so the padded arithmetic is
// value-identical to the reference's variable-length loops (:130-136).
#pragma once
#include "rr_common.cuh"
#include "rr_math.cuh"

namespace rrb {

static __device__ __noinline__ double gr4j_s_curve1(int t, double x4) {  // gr4j_model.py:159-173
    if (t <= 0) return 0.0;
    else if ((double)t < x4) return pow((double)t / x4, 2.5);
    else return 1.0;
}
static __device__ __noinline__ double gr4j_s_curve2(int t, double x4) {  // gr4j_model.py:176-192
    if (t <= 0) return 0.0;
    else if ((double)t <= x4) return 0.5 * pow((double)t / x4, 2.5);
    else if ((double)t < 2 * x4) return 1 - 0.5 * pow(2 - (double)t / x4, 2.5);
    else return 1.0;
}

// OSMEM: the unit hydrograph ordinates (per member, constant over the series) live in shared memory instead of
// registers -- slot j of thread tid at ord + j * kOrdStride, a conflict-free column layout for CTAs of at most
// kOrdThreads threads.  Used by the long-hydrograph class (x4 <= 10: 31 ordinates + 31 state slots per member would
// otherwise take 124 registers and push the coupled kernels into local-memory spills).
constexpr int kOrdThreads = 64;

template <int C1, int C2, int MATH, bool OSMEM = false>
struct Gr4jMember {
    double x1, x2, x3;
    double inv_x1, inv_x3, k_tanh, k49;  // FAST: 1/x1, 1/x3, 2 log2(e) / x1, (4/9) / x1
    int n1, n2;
    double o1[OSMEM ? 1 : C1], o2[OSMEM ? 1 : C2];  // unit hydrograph ordinates (registers)
    uint32_t ord;                                   // OSMEM: shared address of this thread's ordinate column
    double u1[C1], u2[C2];  // routed water still in the unit hydrographs
    double S, R;            // production / routing store

    static constexpr int kStateSlots = 2 + C1 + C2;
    static constexpr uint32_t kOrdStride = kOrdThreads * 8;
    static constexpr size_t kOrdSmemBytes = OSMEM ? (size_t)(C1 + C2) * kOrdStride : 0;

    __device__ __forceinline__ double O1(int j) const {
        if constexpr (OSMEM) return lds_f64(ord + (uint32_t)j * kOrdStride);
        else return o1[j];
    }
    __device__ __forceinline__ double O2(int j) const {
        if constexpr (OSMEM) return lds_f64(ord + (uint32_t)(C1 + j) * kOrdStride);
        else return o2[j];
    }
    __device__ __forceinline__ void set_ord(int slot, double v) {  // slot < C1: UH1, else UH2
        if constexpr (OSMEM) {
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(ord + (uint32_t)slot * kOrdStride), "d"(v) : "memory");
        } else {
            if (slot < C1) o1[slot] = v;
            else o2[slot - C1] = v;
        }
    }

    // ord_addr: shared address of the CTA's ordinate block (OSMEM only; every thread of the CTA passes the same)
    __device__ __forceinline__ void init(const double* p /* x1,x2,x3,x4 */, double s_init, double r_init,
                                         uint32_t ord_addr = 0) {
        ord = ord_addr + threadIdx.x * 8u;
        x1 = p[0]; x2 = p[1]; x3 = p[2];
        const double x4 = p[3];
        inv_x1 = 1.0 / x1; inv_x3 = 1.0 / x3;
        k_tanh = 2.8853900817779268 / x1;  // 2 / ln 2
        k49 = (4.0 / 9.0) / x1;
        if (MATH == RRB_MATH_FAST_) { pin(inv_x1); pin(inv_x3); pin(k_tanh); pin(k49); }
        judge(p, s_init, r_init);
        S = s_init * x1;  // :64
        R = r_init * x3;  // :65
        n1 = (int)ceil(x4);          // :68
        n2 = (int)ceil(2 * x4 + 1);  // :69
        n1 = n1 < 1 ? 1 : (n1 > C1 ? C1 : n1);
        n2 = n2 < 1 ? 1 : (n2 > C2 ? C2 : n2);
#pragma unroll
        for (int j = 1; j <= C1; ++j) {
            set_ord(j - 1, gr4j_s_curve1(j, x4) - gr4j_s_curve1(j - 1, x4));  // :75-76
            u1[j - 1] = 0.0;
        }
#pragma unroll
        for (int j = 1; j <= C2; ++j) {
            set_ord(C1 + j - 1, gr4j_s_curve2(j, x4) - gr4j_s_curve2(j - 1, x4));  // :78-79
            u2[j - 1] = 0.0;
        }
    }

    __device__ __forceinline__ void load(const double* state, int64_t N, int64_t i) {
        S = state[0 * N + i];
        R = state[1 * N + i];
#pragma unroll
        for (int j = 0; j < C1; ++j) u1[j] = state[(2 + j) * N + i];
#pragma unroll
        for (int j = 0; j < C2; ++j) u2[j] = state[(2 + C1 + j) * N + i];
    }
    __device__ __forceinline__ void save(double* state, int64_t N, int64_t i) const {
        state[0 * N + i] = S;
        state[1 * N + i] = R;
#pragma unroll
        for (int j = 0; j < C1; ++j) state[(2 + j) * N + i] = u1[j];
#pragma unroll
        for (int j = 0; j < C2; ++j) state[(2 + C1 + j) * N + i] = u2[j];
    }

    // ---- FAST path contract -----------------------------------------------------------------------
    // step_fast() carries no special-value handling: it is only entered when every member of the CTA is
    // "sane" (finite parameters of moderate magnitude, initial fractions within [0, 1]) and the packed forcing is
    // finite and moderately ranged (flag written by the pack kernel).  Then every store stays finite
    // (S <= 2.25 x1 + P and R <= x3 + UH water after each step), no operand can reach the ranges where the
    // branch-free sequences of rr_math.cuh differ from libm, and NaN never appears.  Otherwise the CTA runs
    // step_precise(): the reference's operations, special values included.
    bool sane;
    __device__ __forceinline__ void judge(const double* p, double s_init, double r_init) {
        // initial fractions within [0, 1] like the wrapper enforces (gr4j.py:136-144): a negative routing store would
        // make (R/x3)^3.5 a NaN that numba's max(0, NaN) = 0 silently absorbs -- reference-order arithmetic only
        sane = p[0] >= 1e-3 && p[0] <= 1e6 && p[2] >= 1e-3 && p[2] <= 1e6 && fabs(p[1]) <= 1e4 && p[3] > 0.0 &&
               p[3] <= 64.0 && s_init >= 0.0 && s_init <= 1.0 && r_init >= 0.0 && r_init <= 1.0;
    }
    // FAST folds the 0.9 / 0.1 split of the routed water (:126-127) into the unit hydrograph ordinates
    __device__ __forceinline__ void enter_fast() {
#pragma unroll
        for (int j = 0; j < C1; ++j) set_ord(j, O1(j) * 0.9);
#pragma unroll
        for (int j = 0; j < C2; ++j) set_ord(C1 + j, O2(j) * 0.1);
    }

    // one timestep: P = precipitation (or Cemaneige liquid outflow), E = potential evapotranspiration
    __device__ __forceinline__ double step(double P, double E, uint32_t) { return step_precise(P, E); }

    __device__ __forceinline__ double step_precise(double P, double E) {
        const bool wet = P >= E;                   // :89
        const double arg = wet ? P - E : E - P;    // p_n (:90) or pe_n (:101)
        const double p_n = wet ? arg : 0.0;
        const double sr = S / x1;
        const double th = tanh(arg / x1);
        const double num = wet ? (x1 * (1 - sr * sr)) * th : (S * (2 - sr)) * th;  // :95 / :107
        const double den = wet ? 1 + sr * th : 1 + (1 - sr) * th;                  // :96 / :108
        const double frac = num / den;
        const double p_s = wet ? frac : 0.0;
        const double e_s = wet ? 0.0 : frac;
        S = S - e_s + p_s;  // :114
        const double u = 4.0 / 9.0 * S / x1;
        const double uu = u * u;
        const double perc = S * (1 - pow(1 + uu * uu, -0.25));  // :117
        S = S - perc;                           // :120
        const double p_r = perc + (p_n - p_s);  // :123
        const double p1 = 0.9 * p_r;            // :126
        const double p2 = 0.1 * p_r;            // :127
        // unit hydrographs, :130-136
#pragma unroll
        for (int j = 0; j < C1 - 1; ++j) {
            const double add = O1(j) * p1;
            u1[j] = (j == n1 - 1) ? add : u1[j + 1] + add;
        }
        u1[C1 - 1] = O1(C1 - 1) * p1;
#pragma unroll
        for (int j = 0; j < C2 - 1; ++j) {
            const double add = O2(j) * p2;
            u2[j] = (j == n2 - 1) ? add : u2[j + 1] + add;
        }
        u2[C2 - 1] = O2(C2 - 1) * p2;
        const double gw = x2 * pow(R / x3, 3.5);  // :139
        R = nb_max0(R + u1[0] + gw);              // :142
        const double v = R / x3;
        const double vv = v * v;
        const double q_r = R * (1 - pow(1 + vv * vv, -0.25));  // :145
        R = R - q_r;                             // :148
        const double q_d = nb_max0(u2[0] + gw);  // :151
        return q_r + q_d;                        // :154
    }

    // FAST (sane operands only, see above): same recurrence, ~80 fp64 instructions instead of ~107
    // (every one of them costs two issue slots), one basic block:
    //  * wet / dry through the sign of d = P - E (P >= E <=> d >= +0; d = -0 needs P = -0 and gives a zero
    //    argument, for which both branches of the reference coincide),
    //  * tanh(a) = m/(m+2) with m = expm1(2a): both production-store formulas collapse to
    //    num m / (2 + c m) with num = base - S sr, base = x1 | 2S and c = 1 + sr | 2 - sr,
    //  * reciprocals of x1 / x3 hoisted; (1+u^4)^(-1/4), w^3.5 and 1/den from the reciprocal / reciprocal
    //    square root seed units with one third-order correction each,
    //  * S (1 - y) as one FMA, the unit hydrographs as one FMA per slot on pre-scaled ordinates.  A padded
    //    slot holds 0 * p = 0: identical to the reference for finite p.
    __device__ __forceinline__ double step_fast(double P, double E, uint32_t tb, const Exp2Regs& k) {
        const double d = P - E;
        const int hd = __double2hiint(d);
        const bool wet = hd >= 0;
        const double arg = fabs(d);
        const double sr = S * inv_x1;
        const double m = exp2m1_sane(arg * k_tanh, tb, k);
        const double sgn = __hiloint2double(0x3FF00000 | (hd & (int)0x80000000), 0);  // +1 wet, -1 dry
        const double cbase = __hiloint2double(wet ? 0x3FF00000 : 0x40000000, 0);      // 1 wet, 2 dry
        const double c = fma(sgn, sr, cbase);
        const double base = wet ? x1 : S + S;
        const double num = fma(-S, sr, base);
        const double frac = (num * m) * rcp_sane(fma(c, m, 2.0));  // p_s (wet) or e_s (dry)
        S = fma(sgn, frac, S);
        const double u = S * k49;
        const double uu = u * u;
        const double perc = fma(-S, rsqrt4_sane(fma(uu, uu, 1.0), k), S);
        S = S - perc;
        const double p_n = wet ? arg - frac : 0.0;
        const double p_r = perc + p_n;
#pragma unroll
        for (int j = 0; j < C1 - 1; ++j) u1[j] = fma(O1(j), p_r, u1[j + 1]);
        u1[C1 - 1] = O1(C1 - 1) * p_r;
#pragma unroll
        for (int j = 0; j < C2 - 1; ++j) u2[j] = fma(O2(j), p_r, u2[j + 1]);
        u2[C2 - 1] = O2(C2 - 1) * p_r;
        const double gw = x2 * pow35_sane(R * inv_x3, k);
        R = max0_sane((R + u1[0]) + gw);
        const double v = R * inv_x3;
        const double vv = v * v;
        const double q_r = fma(-R, rsqrt4_sane(fma(vv, vv, 1.0), k), R);
        R = R - q_r;
        return q_r + max0_sane(u2[0] + gw);
    }
};

// step dispatch on a compile-time tag (ic<1> = FAST step), usable from generic lambdas for member types
// without a FAST step
template <class M>
__device__ __forceinline__ double gr4j_step(ic<1>, M& m, double P, double E, uint32_t tb, const Exp2Regs& k) {
    return m.step_fast(P, E, tb, k);
}
template <class M>
__device__ __forceinline__ double gr4j_step(ic<0>, M& m, double P, double E, uint32_t tb, const Exp2Regs&) {
    return m.step(P, E, tb);
}

// unit-hydrograph capacity classes used by the kernels
using Gr4jUh3F = Gr4jMember<3, 7, RRB_MATH_FAST_>;
using Gr4jUh3P = Gr4jMember<3, 7, RRB_MATH_PRECISE_>;
using Gr4jUh4F = Gr4jMember<4, 9, RRB_MATH_FAST_>;
using Gr4jUh4P = Gr4jMember<4, 9, RRB_MATH_PRECISE_>;
using Gr4jUh10F = Gr4jMember<10, 21, RRB_MATH_FAST_, true>;
using Gr4jUh10P = Gr4jMember<10, 21, RRB_MATH_PRECISE_, true>;

// generic fallback for very long unit hydrographs (x4 beyond the register variants): buffers
// in per-thread local memory with run-time lengths, PRECISE arithmetic only.  x4 <= 64.
constexpr int kGr4jGenericC1 = 64, kGr4jGenericC2 = 129;

struct Gr4jMemberDyn {
    double x1, x2, x3;
    int n1, n2;
    double o1[kGr4jGenericC1], o2[kGr4jGenericC2];
    double u1[kGr4jGenericC1], u2[kGr4jGenericC2];
    double S, R;

    static constexpr int kStateSlots = 2 + kGr4jGenericC1 + kGr4jGenericC2;
    static constexpr size_t kOrdSmemBytes = 0;

    __device__ void init(const double* p, double s_init, double r_init, uint32_t = 0) {
        x1 = p[0]; x2 = p[1]; x3 = p[2];
        const double x4 = p[3];
        S = s_init * x1;
        R = r_init * x3;
        n1 = (int)ceil(x4);
        n2 = (int)ceil(2 * x4 + 1);
        n1 = n1 < 1 ? 1 : (n1 > kGr4jGenericC1 ? kGr4jGenericC1 : n1);
        n2 = n2 < 1 ? 1 : (n2 > kGr4jGenericC2 ? kGr4jGenericC2 : n2);
        for (int j = 1; j <= kGr4jGenericC1; ++j) {
            o1[j - 1] = (j <= n1) ? gr4j_s_curve1(j, x4) - gr4j_s_curve1(j - 1, x4) : 0.0;
            u1[j - 1] = 0.0;
        }
        for (int j = 1; j <= kGr4jGenericC2; ++j) {
            o2[j - 1] = (j <= n2) ? gr4j_s_curve2(j, x4) - gr4j_s_curve2(j - 1, x4) : 0.0;
            u2[j - 1] = 0.0;
        }
    }
    __device__ void load(const double* state, int64_t N, int64_t i) {
        S = state[0 * N + i];
        R = state[1 * N + i];
        for (int j = 0; j < kGr4jGenericC1; ++j) u1[j] = state[(2 + j) * N + i];
        for (int j = 0; j < kGr4jGenericC2; ++j) u2[j] = state[(2 + kGr4jGenericC1 + j) * N + i];
    }
    __device__ void save(double* state, int64_t N, int64_t i) const {
        state[0 * N + i] = S;
        state[1 * N + i] = R;
        for (int j = 0; j < kGr4jGenericC1; ++j) state[(2 + j) * N + i] = u1[j];
        for (int j = 0; j < kGr4jGenericC2; ++j) state[(2 + kGr4jGenericC1 + j) * N + i] = u2[j];
    }
    __device__ double step(double P, double E, uint32_t) {
        const bool wet = P >= E;
        const double arg = wet ? P - E : E - P;
        const double p_n = wet ? arg : 0.0;
        const double sr = S / x1;
        const double th = tanh(arg / x1);
        const double num = wet ? (x1 * (1 - sr * sr)) * th : (S * (2 - sr)) * th;
        const double den = wet ? 1 + sr * th : 1 + (1 - sr) * th;
        const double frac = num / den;
        const double p_s = wet ? frac : 0.0;
        const double e_s = wet ? 0.0 : frac;
        S = S - e_s + p_s;
        const double u = 4.0 / 9.0 * S / x1;
        const double uu = u * u;
        const double perc = S * (1 - pow(1 + uu * uu, -0.25));
        S = S - perc;
        const double p_r = perc + (p_n - p_s);
        const double p1 = 0.9 * p_r;
        const double p2 = 0.1 * p_r;
        for (int j = 0; j < n1 - 1; ++j) u1[j] = u1[j + 1] + o1[j] * p1;
        u1[n1 - 1] = o1[n1 - 1] * p1;
        for (int j = 0; j < n2 - 1; ++j) u2[j] = u2[j + 1] + o2[j] * p2;
        u2[n2 - 1] = o2[n2 - 1] * p2;
        const double gw = x2 * pow(R / x3, 3.5);
        R = nb_max0(R + u1[0] + gw);
        const double v = R / x3;
        const double vv = v * v;
        const double q_r = R * (1 - pow(1 + vv * vv, -0.25));
        R = R - q_r;
        const double q_d = nb_max0(u2[0] + gw);
        return q_r + q_d;
    }
};

}  // namespace rrb
