// rr_math.cuh -- FAST-path fp64 transcendentals for the recurrence kernels.
//
// The reference calls glibc pow/tanh once or more per member-timestep
// (rrmpg/models/hbvedu_model.py:99, rrmpg/models/gr4j_model.py:95-96,117,139,145).  CUDA's libm
// pow costs ~95 fp64 instructions, which makes the kernels fp64-issue-bound far below the HBM
// roofline.  These replacements trade the last ulp for a 4x shorter instruction sequence:
// table-driven log2 / exp2 (128-entry tables staged in shared memory, tables from
// gen_math_tables.py), ~1e-15 relative accuracy, with the exotic operand ranges (zero,
// negative, denormal, inf/nan, overflow) routed to the libm slow path so special-value
// behaviour is unchanged.  The functions are __host__ __device__ so the algorithm and tables
// are also checked on the CPU against glibc (rrb_host_fast_pow in rr_api.cu, tests/test_fastmath.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stddef.h>

#include "rr_math_tables.h"

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace rrb {

enum { RRB_MATH_FAST_ = 0, RRB_MATH_PRECISE_ = 1 };

struct FastTables {
    double log2tab[2 * tables::kLogN];           // {invc, log2c}
    unsigned long long exp2tab[tables::kExpN];   // bits(2^(j/128)) - (j << 45)
    double exp2m1tab[tables::kExpN];             // 2^(j/128) - 1
    double log2poly[8];                          // log2(1+r)/r coefficients (7 used)
    double exp2poly[8];                          // (2^r - 1)/r coefficients (6 used)
};

#ifdef __CUDACC__
static __device__ const FastTables d_fast_tables = {RRB_LOG2_TABLE, RRB_EXP2_TABLE, RRB_EXP2M1_TABLE, RRB_LOG2_POLY, RRB_EXP2_POLY};
#endif
static const FastTables h_fast_tables = {RRB_LOG2_TABLE, RRB_EXP2_TABLE, RRB_EXP2M1_TABLE, RRB_LOG2_POLY, RRB_EXP2_POLY};

__host__ __device__ __forceinline__ uint64_t f64_bits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
__host__ __device__ __forceinline__ double bits_f64(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

// log2(x) for positive normal finite x.  |error| ~ 1e-16 * max(1, |log2 x|).
__host__ __device__ __forceinline__ double fast_log2_normal(double x, const FastTables* tb) {
    constexpr double A[7] = RRB_LOG2_POLY;
    const uint64_t ix = f64_bits(x);
    const uint64_t tmp = ix - tables::kLogOff;
    const int i = (int)((tmp >> 45) & (tables::kLogN - 1));
    const int k = (int)((int64_t)tmp >> 52);
    const double z = bits_f64(ix - (tmp & 0xFFF0000000000000ULL));
    const double invc = tb->log2tab[2 * i], log2c = tb->log2tab[2 * i + 1];
    const double r = fma(z, invc, -1.0);  // |r| <= 2^-8
    const double base = (double)k + log2c;
    // Estrin: r*(A1 + r A2 + r^2 (A3 + r A4) + r^4 (A5 + r A6))
    const double r2 = r * r;
    const double a = fma(r, A[1], A[0]);
    const double b = fma(r, A[3], A[2]);
    const double c = fma(r, A[5], A[4]);
    const double r4 = r2 * r2;
    double t = fma(r2, b, a);
    t = fma(r4, c, t);
    return fma(r, t, base);
}

// 2^z for |z| < 1020.  relative error ~ 2e-16.
__host__ __device__ __forceinline__ double fast_exp2_bounded(double z, const FastTables* tb) {
    constexpr double C[6] = RRB_EXP2_POLY;
    constexpr double kShift = 0x1.8p52 / tables::kExpN;  // rounds z to multiples of 1/128
    double kd = z + kShift;
    const uint64_t ki = f64_bits(kd);
    kd -= kShift;
    const double r = z - kd;  // |r| <= 2^-8
    const uint64_t sbits = tb->exp2tab[ki & (tables::kExpN - 1)] + (ki << 45);
    const double scale = bits_f64(sbits);
    // 2^r - 1 = r*(C1 + r C2 + r^2 (C3 + r C4 + r^2 C5))
    const double r2 = r * r;
    const double a = fma(r, C[1], C[0]);
    double b = fma(r, C[3], C[2]);
    b = fma(r2, C[4], b);
    const double t = fma(r2, b, a);
    return fma(scale, r * t, scale);
}

// 2^z - 1 for 0 <= z (z is clamped to 64; NaN propagates).  Relative error ~3e-16 for every z,
// including z -> 0 (no cancellation: the table holds 2^(j/128) - 1 itself).
__host__ __device__ __forceinline__ double fast_exp2m1_nonneg(double z, const FastTables* tb) {
    constexpr double C[6] = RRB_EXP2_POLY;
    constexpr double kShift = 0x1.8p52 / tables::kExpN;
    z = (z > 64.0) ? 64.0 : z;
    double kd = z + kShift;
    const uint64_t ki = f64_bits(kd);
    kd -= kShift;
    const double r = z - kd;
    const int j = (int)(ki & (tables::kExpN - 1));
    const double scale = bits_f64(tb->exp2tab[j] + (ki << 45));
    const double base = ((ki & 0xFFFFFFFFULL) < (uint64_t)tables::kExpN) ? tb->exp2m1tab[j] : scale - 1.0;
    const double r2 = r * r;
    const double a = fma(r, C[1], C[0]);
    double b = fma(r, C[3], C[2]);
    b = fma(r2, C[4], b);
    const double t = fma(r2, b, a);
    return fma(scale, r * t, base);
}

// x^y.  Fast path for positive normal x with a comfortably finite result, libm otherwise.
__host__ __device__ __forceinline__ double fast_pow(double x, double y, const FastTables* tb) {
    const uint64_t ix = f64_bits(x);
    if (ix - 0x0010000000000000ULL < 0x7FE0000000000000ULL) {  // 0 < x < inf, normal
        const double z = y * fast_log2_normal(x, tb);
        if (fabs(z) < 1000.0) return fast_exp2_bounded(z, tb);  // also rejects NaN
    }
    return pow(x, y);
}

#ifdef __CUDACC__
// Device-side variants reading the tables through a 32-bit shared-window address (see lds_f64x2 in
// rr_common.cuh for why).  No range checks: the caller guarantees 2^-16 < x < 2^16 and |y| < 32,
// hence x positive normal and |y log2 x| <= 512.  Out-of-contract operands give garbage, never a fault
// (table indices are masked).
// Coefficients held in registers for the whole kernel (load once from the parameter bank, then pin()):
// an FMA can take only one constant-bank operand, so polynomial steps of the form fma(r, c1, c0) would
// otherwise need a load per step.
struct PowRegs {
    double a0, a1, a2, a3, a4, a5;  // log2(1+r)/r, degree 5
    double c0, c1, c2, c3, c4;      // (2^r - 1)/r, degree 4
};

__device__ __forceinline__ double fast_pow_unchecked_smem(double x, double y, uint32_t tb_addr, const PowRegs& pc) {
    constexpr double kShift = 0x1.8p52 / tables::kExpN;
    // ---- log2(x); all index / exponent work on the high word (the low word passes through unchanged)
    const uint32_t hx = (uint32_t)__double2hiint(x);
    const uint32_t ht = hx - (uint32_t)(tables::kLogOff >> 32);
    const int k = (int)ht >> 20;
    const double z = __hiloint2double((int)(hx - (ht & 0xFFF00000u)), __double2loint(x));
    double invc, log2c;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(log2c) : "r"(tb_addr + ((ht >> 9) & 0x7F0u)));
    const double r = fma(z, invc, -1.0);
    const double base = (double)k + log2c;
    const double r2 = r * r;
    const double a = fma(r, pc.a1, pc.a0);
    const double b = fma(r, pc.a3, pc.a2);
    const double c = fma(r, pc.a5, pc.a4);
    const double r4 = r2 * r2;
    double t = fma(r2, b, a);
    t = fma(r4, c, t);
    const double zz = y * fma(r, t, base);
    // ---- 2^zz
    double kd = zz + kShift;
    const uint32_t ki = (uint32_t)__double2loint(kd);  // round(128 zz), two's complement
    kd -= kShift;
    const double rr = zz - kd;
    uint32_t tlo, thi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                 : "=r"(tlo), "=r"(thi)
                 : "r"(tb_addr + (uint32_t)(2 * tables::kLogN * 8) + (ki & (tables::kExpN - 1)) * 8u));
    const double scale = __hiloint2double((int)(thi + (ki << 13)), (int)tlo);  // bits + (ki << 45)
    const double q2 = rr * rr;
    const double e = fma(rr, pc.c1, pc.c0);
    double f = fma(rr, pc.c3, pc.c2);
    f = fma(q2, pc.c4, f);
    const double g = fma(q2, f, e);
    return fma(scale, rr * g, scale);
}

// coefficients come from the shared-memory copy of the tables through volatile loads: values the
// assembler cannot re-materialise, so they stay in registers for the whole time loop
__device__ __forceinline__ double lds_f64_at(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ PowRegs load_pow_regs(uint32_t tb_addr) {
    const uint32_t a = tb_addr + (uint32_t)offsetof(FastTables, log2poly);
    const uint32_t c = tb_addr + (uint32_t)offsetof(FastTables, exp2poly);
    return PowRegs{lds_f64_at(a), lds_f64_at(a + 8), lds_f64_at(a + 16), lds_f64_at(a + 24), lds_f64_at(a + 32),
                   lds_f64_at(a + 40), lds_f64_at(c), lds_f64_at(c + 8), lds_f64_at(c + 16), lds_f64_at(c + 24),
                   lds_f64_at(c + 32)};
}

// ----------------------------------------------------------------------------------------
// Branch-free variants for SANE operands (finite, moderately ranged; the kernels establish that once
// per CTA before the time loop -- Gr4jMember::sane -- and run the reference-order arithmetic otherwise).
// No special-value handling lives in the time loop: a step is one basic block, so the scheduler can
// overlap the production-store and routing-store chains, and every fp64 instruction saved is two
// issue slots.  Seeds come from MUFU.RCP64H / MUFU.RSQ64H (~2^-20 relative); one third-order
// correction takes them below 2 ulp.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_approx_f64(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
}
struct Exp2Regs {
    double c0, c1, c2, c3, c4;  // (2^r - 1)/r, degree 4
    double k5_32, k3_8;         // 5/32, 3/8: multiplier constants of the third-order corrections (an FMA takes one immediate)
};
__device__ __forceinline__ Exp2Regs load_exp2_regs(uint32_t tb_addr) {
    const uint32_t c = tb_addr + (uint32_t)offsetof(FastTables, exp2poly);
    double k5_32 = 0.15625, k3_8 = 0.375;
    asm volatile("" : "+d"(k5_32), "+d"(k3_8));  // opaque: held in registers instead of re-materialised per use
    return Exp2Regs{lds_f64_at(c), lds_f64_at(c + 8), lds_f64_at(c + 16), lds_f64_at(c + 24), lds_f64_at(c + 32), k5_32, k3_8};
}
// clamps on the high word of a non-negative double (integer min / max: one ALU instruction)
__device__ __forceinline__ double clamp_hi_le(double x, int hi_cap) {
    return __hiloint2double(min(__double2hiint(x), hi_cap), __double2loint(x));
}
__device__ __forceinline__ double clamp_hi_ge(double x, int hi_floor) {
    return __hiloint2double(max(__double2hiint(x), hi_floor), __double2loint(x));
}
// numba's max(0, x) for a non-NaN x through the sign bit: ISETP + 2 SEL instead of DSETP + 2 SEL
__device__ __forceinline__ double max0_sane(double x) { return (__double2hiint(x) < 0) ? 0.0 : x; }

// 2^z - 1 for finite z >= 0 (z is capped just above 64, where tanh has long been 1)
__device__ __forceinline__ double exp2m1_sane(double z, uint32_t tb_addr, const Exp2Regs& k) {
    constexpr double kShift = 0x1.8p52 / tables::kExpN;
    z = clamp_hi_le(z, 0x40500000);  // 64.0
    double kd = z + kShift;
    const uint32_t ki = (uint32_t)__double2loint(kd);
    kd -= kShift;
    const double r = z - kd;
    const uint32_t j8 = (ki & (tables::kExpN - 1)) * 8u;
    uint32_t tlo, thi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                 : "=r"(tlo), "=r"(thi)
                 : "r"(tb_addr + (uint32_t)offsetof(FastTables, exp2tab) + j8));
    const double scale = __hiloint2double((int)(thi + (ki << 13)), (int)tlo);
    double m1;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(m1) : "r"(tb_addr + (uint32_t)offsetof(FastTables, exp2m1tab) + j8));
    const double base = (ki < (uint32_t)tables::kExpN) ? m1 : scale - 1.0;
    const double r2 = r * r;
    const double a = fma(r, k.c1, k.c0);
    double b = fma(r, k.c3, k.c2);
    b = fma(r2, k.c4, b);
    const double t = fma(r2, b, a);
    return fma(scale, r * t, base);
}
// 1 / den for a positive normal den: y (1 + e + e^2), e = 1 - den y
__device__ __forceinline__ double rcp_sane(double den) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
    const double e = fma(-den, y, 1.0);
    return fma(y, fma(e, e, e), y);
}
// v^(-1/4) for 1 <= v < 2^1000: y (1 + e/4 + 5 e^2/32), e = 1 - v y^4
__device__ __forceinline__ double rsqrt4_sane(double v, const Exp2Regs& k) {
    const double a = rsqrt_approx_f64(v);
    const double y = a * rsqrt_approx_f64(a);
    const double y2 = y * y;
    const double e = fma(-v, y2 * y2, 1.0);
    return fma(y, e * fma(e, k.k5_32, 0.25), y);
}
// w^3.5 = w^4 w^(-1/2) for finite w >= 0 (w = 0 -> 0: the seed operand is floored at 2^-1021, w^4 underflows)
__device__ __forceinline__ double pow35_sane(double w, const Exp2Regs& k) {
    const double wc = clamp_hi_ge(w, 0x00200000);
    const double y0 = rsqrt_approx_f64(wc);
    const double t = wc * y0;
    const double e = fma(-t, y0, 1.0);                 // 1 - w y^2
    const double y = fma(y0 * e, fma(e, k.k3_8, 0.5), y0);  // y (1 + e/2 + 3 e^2/8)
    const double w2 = w * w;
    return (w2 * w2) * y;
}

// ----------------------------------------------------------------------------------------
// Round 2: the HBV-Edu soil chain.  prec_eff = liquid (soil/FC)^Beta (rrmpg/models/hbvedu_model.py:99) sits on the
// only long loop-carried dependency of the model (soil -> log2 -> x Beta -> exp2 -> soil); with 3.5 warps per SM
// sub-partition at 65 536 members the kernel runs at the latency of that chain, not at an issue or pipe limit
// (profiles/r02_fp64_probe.txt).  The sequence below is organised for DEPTH:
//   * log2(soil/FC) = log2(soil) - log2(FC) with the second term hoisted (one DMUL off the chain),
//   * 512-entry log2 table indexed by the top mantissa bits (|r| < 2^-10: degree 4 instead of 6, Estrin depth 2),
//   * kd = fma(Beta, L, shift) / r = fma(Beta, L, -(kd - shift)): the product Beta L is never rounded on its own,
//   * 1024-entry exp2 table (|r| <= 2^-11: (2^r - 1)/r at degree 2),
//   * the table scale 2^(k/1024) arrives late (LDS behind the shift), so it enters last:
//     w = liquid (1 + r g) first, then soil_new = fma(-scale, w, soil_partial) and prec_eff = scale w side by side.
// Host/device twin of the exact device sequence (tests/test_fastmath.py checks it against glibc pow on the CPU).
// ----------------------------------------------------------------------------------------
namespace hbvpow {
// log2(1 + r) / r = A1 + A2 r + A3 r^2 + A4 r^3 (|r| < 2^-10: the r^5 term is below 2^-51.8 absolute)
constexpr double A1 = 0x1.71547652b82fep+0, A2 = -0x1.71547652b82fep-1, A3 = 0x1.ec709dc3a03fdp-2, A4 = -0x1.71547652b82fep-2;
// (2^r - 1) / r = C1 + C2 r + C3 r^2 (|r| <= 2^-11: the r^4 term is below 5.5e-16 relative)
constexpr double C1 = 0x1.62e42fefa39efp-1, C2 = 0x1.ebfbdff82c58fp-3, C3 = 0x1.c6b08d704a0c0p-5;
constexpr double kShift = 0x1.8p52 / tables::kExpKN;  // rounds to multiples of 1/1024
}  // namespace hbvpow
struct HbvTables {
    double log2m[2 * tables::kLogMN];           // {invc, log2c}, index = top 9 mantissa bits
    unsigned long long exp2k[tables::kExpKN];   // bits(2^(j/1024)) - (j << 42)
    double poly[8];                             // A1..A4, C1..C3: read back through volatile shared loads so that the
                                                // assembler keeps them in registers instead of re-materialising them
};
#define RRB_HBV_POLY {hbvpow::A1, hbvpow::A2, hbvpow::A3, hbvpow::A4, hbvpow::C1, hbvpow::C2, hbvpow::C3, 0.0}
#ifdef __CUDACC__
static __device__ const HbvTables d_hbv_tables = {RRB_LOG2M_TABLE, RRB_EXP2K_TABLE, RRB_HBV_POLY};
#endif
static const HbvTables h_hbv_tables = {RRB_LOG2M_TABLE, RRB_EXP2K_TABLE, RRB_HBV_POLY};

// liquid * (soil/FC)^Beta and the new soil moisture soil_partial - that, for a positive normal soil with
// soil/FC in [2^-15, 2^15), |Beta| < 32, log2FC = log2(FC).  Returns prec_eff; *soil_new = fma(-scale, w, soil_partial).
__host__ __device__ __forceinline__ double hbv_pow_step_twin(double soil, double log2FC, double Beta, double liquid,
                                                            double soil_partial, const HbvTables* tb, double* soil_new) {
    using namespace hbvpow;
    const uint64_t ix = f64_bits(soil);
    const uint32_t hs = (uint32_t)(ix >> 32);
    const int i = (int)((hs >> 11) & (tables::kLogMN - 1));
    const double m = bits_f64((ix & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL);
    const double kml = (double)((int)(hs >> 20) - 1023) - log2FC;
    const double invc = tb->log2m[2 * i], log2c = tb->log2m[2 * i + 1];
    const double r = fma(m, invc, -1.0);
    const double base = kml + log2c;
    const double r2 = r * r;
    const double a = fma(r, A2, A1);
    const double b = fma(r, A4, A3);
    const double t = fma(r2, b, a);
    const double L = fma(r, t, base);                 // log2(soil / FC)
    const double kd = fma(Beta, L, kShift);
    const uint32_t ki = (uint32_t)f64_bits(kd);       // round(1024 Beta L), two's complement
    const double kdm = kd - kShift;
    const double rr = fma(Beta, L, -kdm);             // |rr| <= 2^-11
    const double scale = bits_f64(tb->exp2k[ki & (tables::kExpKN - 1)] + ((uint64_t)ki << 42));
    const double q2 = rr * rr;
    const double e = fma(rr, C2, C1);
    const double g = fma(q2, C3, e);
    const double u = rr * g;                          // 2^rr - 1
    const double w = fma(liquid, u, liquid);          // liquid 2^rr
    *soil_new = fma(-scale, w, soil_partial);
    return scale * w;
}

// stage the HBV tables into shared memory (call by every thread, before a __syncthreads())
__device__ __forceinline__ const HbvTables* hbv_tables_to_smem(unsigned char* smem_at) {
    HbvTables* dst = reinterpret_cast<HbvTables*>(smem_at);
    const double2* src = reinterpret_cast<const double2*>(&d_hbv_tables);
    double2* d = reinterpret_cast<double2*>(dst);
    for (int k = threadIdx.x; k < (int)(sizeof(HbvTables) / 16); k += blockDim.x) d[k] = src[k];
    return dst;
}
__host__ __device__ constexpr size_t hbv_tables_smem_bytes() { return sizeof(HbvTables); }

// stage the tables into shared memory (call by every thread, before a __syncthreads())
__device__ __forceinline__ const FastTables* fastmath_tables_to_smem(unsigned char* smem_at) {
    FastTables* dst = reinterpret_cast<FastTables*>(smem_at);
    const uint64_t* src = reinterpret_cast<const uint64_t*>(&d_fast_tables);
    uint64_t* d = reinterpret_cast<uint64_t*>(dst);
    for (int k = threadIdx.x; k < (int)(sizeof(FastTables) / 8); k += blockDim.x) d[k] = src[k];
    return dst;
}
#endif

__host__ __device__ constexpr size_t fastmath_smem_bytes() { return sizeof(FastTables); }

}  // namespace rrb
