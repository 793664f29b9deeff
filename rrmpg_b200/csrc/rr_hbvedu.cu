// rr_hbvedu.cu -- HBV-Edu ensemble kernel (the model BASELINE.json's metric is quoted on).
// Restates run_hbvedu (rrmpg/models/hbvedu_model.py:16-129) for N members at once and replaces
// the member loop of HBVEdu.simulate (rrmpg/models/hbvedu.py:199-209).
//
// Packed forcing per timestep (member independent): F[t] = { temp, prec, dT, PEm } with
//   dT  = temp[t] - T_m[month[t]]   (the inner subtraction of hbvedu_model.py:102, bit-identical)
//   PEm = PE_m[month[t]]
// (FAST packs dT*PEm instead of dT: pe = (1 + C dT) PEm = fma(C, dT*PEm, PEm) is then one instruction.)
// Per member: 4 stores (snow, soil, s1, s2) and 11 parameters in registers.
//
// MATH = PRECISE: IEEE divisions, CUDA libm pow (<= 2 ulp from glibc's), no contraction: every
//                 operation of the reference in the reference's order.
// MATH = FAST   : same recurrence with the fp64 instruction count cut ~4x: reciprocals of FC / PWP and
//                 the linear-store coefficients hoisted, explicit FMAs, table-driven pow (rr_math.cuh,
//                 ~1e-15 relative), and the pow skipped for warps whose members all have
//                 liquid_water == 0 (then prec_eff = 0 * finite = +0 exactly as in the reference).
//                 Discharge stays within rtol 1e-10 of the reference (tests/test_parity_gpu.py).
// Both: branch-free snow routine, forcing of step t+1 prefetched into registers during step t,
//       running output pointers, t = 0 peeled out of the time loop.
#include "rr_common.cuh"
#include "rr_kernels.h"
#include "rr_math.cuh"

#include <cstdlib>

#ifndef RRB_HBV_DEFAULT_VARIANT
#define RRB_HBV_DEFAULT_VARIANT 2
#endif

namespace rrb {

__global__ void hbv_pack_kernel(const double* __restrict__ temp, const double* __restrict__ prec,
                                const int8_t* __restrict__ month0, const double* __restrict__ PE_m,
                                const double* __restrict__ T_m, int64_t T, int64_t Tpad, int fast,
                                double* __restrict__ F, uint32_t* __restrict__ fflag) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= Tpad) return;
    const int64_t c = blockIdx.y;  // catchment
    temp += c * T; prec += c * T; month0 += c * T; PE_m += c * 12; T_m += c * 12;
    F += c * Tpad * kHbvR;
    double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
    if (t < T) {
        const int m = month0[t];
        v.x = temp[t];
        v.y = prec[t];
        v.z = temp[t] - T_m[m];
        v.w = PE_m[m];
        if (fast) v.z = v.z * v.w;
        // inf / NaN precipitation or temperature: FAST contract broken (prec - prec, sign test of temp - T_t)
        if (!(fabs(v.y) <= 1e300) || !(fabs(v.x) <= 1e300)) atomicOr(fflag, 1u);
    }
    reinterpret_cast<double4*>(F)[t] = v;
}

cudaError_t pack_hbvedu(const double* temp, const double* prec, const int8_t* month0, const double* PE_m,
                        const double* T_m, int64_t T, double* F, int math, int count, cudaStream_t s) {
    int64_t Tpad = padded_steps(T, kHbvTT);
    uint32_t* fflag = reinterpret_cast<uint32_t*>(F + (int64_t)count * Tpad * kHbvR);
    cudaError_t e = cudaMemsetAsync(fflag, 0, kForcingFlagBytes, s);
    if (e != cudaSuccess) return e;
    hbv_pack_kernel<<<dim3((unsigned)((Tpad + 255) / 256), (unsigned)count), 256, 0, s>>>(
        temp, prec, month0, PE_m, T_m, T, Tpad, math == RRB_MATH_FAST_, F, fflag);
    return cudaGetLastError();
}

struct HbvOut {
    double *qsim, *snow, *soil, *s1, *s2;
};

// the per-CTA flag words follow the forcing flag slot (kForcingFlagBytes behind it; rr_kernels.h: hbv_scratch_bytes)
__device__ __forceinline__ const uint32_t* hbv_cta_flags(const uint32_t* fflag) { return fflag + kForcingFlagBytes / 4; }
__device__ __forceinline__ uint32_t* hbv_cta_flags(uint32_t* fflag) { return fflag + kForcingFlagBytes / 4; }

// blockIdx.y = catchment: shift every per-catchment pointer (a no-op for the single-catchment launch)
#define HBV_BATCH_PROLOGUE                                                        \
    if (batch.count > 1) {                                                        \
        const int64_t c = blockIdx.y;                                             \
        F += c * batch.forcing_stride;                                            \
        params += c * N * 11;                                                     \
        if (out.qsim) out.qsim += c * batch.out_stride;                           \
        if (out.snow) {                                                           \
            out.snow += c * batch.out_stride; out.soil += c * batch.out_stride;   \
            out.s1 += c * batch.out_stride; out.s2 += c * batch.out_stride;       \
        }                                                                         \
        if (obj.qobs) { obj.qobs += c * obj.T; obj.mse += c * N; }                \
    }                                                                             \
    if (batch.inits) { /* also for a batch (or a chunk of a batch) of ONE catchment */ \
        const int64_t c = batch.count > 1 ? (int64_t)blockIdx.y : 0;              \
        snow0 = batch.inits[4 * c]; soil0 = batch.inits[4 * c + 1];               \
        s10 = batch.inits[4 * c + 2]; s20 = batch.inits[4 * c + 3];               \
    }

struct HbvF {  // forcing of one timestep
    double temp, prec, dT, PEm;
    static __device__ __forceinline__ HbvF load(uint32_t addr) {
        const double2 a = lds_f64x2(addr), b = lds_f64x2(addr + 16);
        return HbvF{a.x, a.y, b.x, b.y};
    }
};

// rare operands (soil/FC outside [2^-15, 2^15), |Beta| >= 32, FC <= 0, non-finite values): the reference's
// own operations, out of line so they cost the time loop one predicated branch
static __device__ __noinline__ double hbv_slow_pow(double soil, double FC, double Beta) { return pow(soil / FC, Beta); }

// ------------------------------------------------------------------------------------------------
// PRECISE: every operation of the reference, in the reference's order.
// ------------------------------------------------------------------------------------------------
template <bool WRITEQ, bool STORAGE, bool OBJ>
__global__ void hbv_precise_kernel(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
                                   const double* __restrict__ params, int64_t N, HbvOut out, Slab slab,
                                   Objective obj, Batch batch, const uint32_t* __restrict__ only_if_flag, int fast_packing,
                                   int fast_grid_x, int flag_div) {
    // launched behind the FAST kernel with the same grid: do the work only when that one declined it -- the whole
    // launch (forcing flag: non-finite rain / temperature) or this CTA (its flag word: a member outside the FAST
    // contract, or a soil moisture that left the range of the table-driven pow; see hbv_fast2_kernel)
    // (the FAST launch has fast_grid_x CTAs per catchment, each covering flag_div CTAs of this launch)
    if (only_if_flag && *only_if_flag == 0u &&
        (flag_div == 0 || hbv_cta_flags(only_if_flag)[blockIdx.y * fast_grid_x + blockIdx.x / flag_div] == 0u)) return;
    HBV_BATCH_PROLOGUE
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // threads past the end of the ensemble recompute member N-1 and store the same values to the same
    // addresses: no predicate lives in the time loop
    const int64_t i = gi < N ? gi : N - 1;
    // record order = HBVEdu._dtype (rrmpg/models/hbvedu.py:63-66)
    const double* p = params + 11 * i;
    const double T_t = p[0], DD = p[1], FC = p[2], Beta = p[3], C = p[4], PWP = p[5];
    const double K_0 = p[6], K_1 = p[7], K_2 = p[8], K_p = p[9], L = p[10];

    double snow = snow0, soil = soil0, s1 = s10, s2 = s20;  // hbvedu_model.py:78-81
    double acc = 0.0;
    int64_t t_first = slab.t_begin;
    int64_t off = i + (slab.t_begin - slab.row0) * N;  // row r of the buffers = timestep row0 + r
    if (slab.t_begin > 0) {
        snow = slab.state[0 * N + i];
        soil = slab.state[1 * N + i];
        s1 = slab.state[2 * N + i];
        s2 = slab.state[3 * N + i];
        if (OBJ) acc = slab.state[4 * N + i];
    } else {
        // t = 0 is not simulated (the reference loop starts at 1, hbvedu_model.py:84): qsim[0] = 0 and the
        // storages hold the initial states
        if (WRITEQ) st_stream(out.qsim + off, 0.0);
        if (STORAGE) {
            st_stream(out.snow + off, snow);
            st_stream(out.soil + off, soil);
            st_stream(out.s1 + off, s1);
            st_stream(out.s2 + off, s2);
        }
        if (OBJ) {
            const double d = obj.qobs[0];
            acc = d * d;
        }
        off += N;
        t_first = 1;
    }
    double* q_o = WRITEQ ? out.qsim + off : nullptr;
    double* snow_o = STORAGE ? out.snow + off : nullptr;
    double* soil_o = STORAGE ? out.soil + off : nullptr;
    double* s1_o = STORAGE ? out.s1 + off : nullptr;
    double* s2_o = STORAGE ? out.s2 + off : nullptr;

    stream_forcing_regs<kHbvR, kHbvTT, HbvF>(F, t_first, slab.t_end, [&](int64_t t, const HbvF& f) {
        // snow routine, both branches evaluated and selected (hbvedu_model.py:87-96)
        const bool cold = f.temp < T_t;
        const double m = DD * (f.temp - T_t);
        const double snow_new = cold ? snow + f.prec : nb_max0(snow - m);
        const double liquid = cold ? 0.0 : f.prec + nb_min(snow, m);
        const double prec_eff = liquid * pow(soil / FC, Beta);              // :99
        // (FAST packing holds dT * PEm in the dT slot: pe = PEm + C (dT PEm) -- only reached for non-finite rain)
        const double pe = fast_packing ? f.PEm + C * f.dT : (1 + C * f.dT) * f.PEm;  // :102
        const double ea = (soil > PWP) ? pe : pe * (soil / PWP);            // :105-108
        const double soil_new = soil + liquid - prec_eff - ea;              // :111
        const double over = nb_max0(s1 - L);
        const double s1_new = s1 + prec_eff - over * K_0 - s1 * K_1 - s1 * K_p;  // :114-118
        const double s2_new = s2 + s1 * K_p - s2 * K_2;                     // :121-123
        const double qv = over * K_0 + s1_new * K_1 + s2_new * K_2;         // :125-127
        snow = snow_new;
        soil = soil_new;
        s1 = s1_new;
        s2 = s2_new;
        if (WRITEQ) {
            st_stream(q_o, qv);
            q_o += N;
        }
        if (STORAGE) {
            st_stream(snow_o, snow);
            st_stream(soil_o, soil);
            st_stream(s1_o, s1);
            st_stream(s2_o, s2);
            snow_o += N;
            soil_o += N;
            s1_o += N;
            s2_o += N;
        }
        if (OBJ) {
            const double d = obj.qobs[t] - qv;
            acc += d * d;
        }
    });

    if (gi < N) {
        if (slab.save_state) {
            slab.state[0 * N + i] = snow;
            slab.state[1 * N + i] = soil;
            slab.state[2 * N + i] = s1;
            slab.state[3 * N + i] = s2;
            if (OBJ) slab.state[4 * N + i] = acc;
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i] = acc / (double)obj.T;
    }
}

// ------------------------------------------------------------------------------------------------
// FAST: the same recurrence reorganised for the fp64 pipe and for the few warps per scheduler that a
// 65 536-member ensemble leaves (one thread per member = 3.5 warps per SM sub-partition).
//   A(t)  snow routine + potential evapotranspiration: needs forcing[t], snow[t-1] and parameters only
//   B(t)  soil moisture (the pow), response routine, discharge: the long dependent chain
// Timesteps are processed in groups of G: first A for the whole group (independent work, high ILP),
// then one warp vote per step ("does any member have liquid water?", known without touching the soil
// chain), then the B chains, each either with or without the pow.
// ------------------------------------------------------------------------------------------------
#ifndef RRB_HBV_GROUP
#define RRB_HBV_GROUP 2
#endif
constexpr int kHbvGroup = RRB_HBV_GROUP;

// Cost model (measured, profiles/r01_*): with one thread per member the 65 536-member workload leaves 3.5
// warps per SM sub-partition and the kernel is ISSUE bound -- every fp64 instruction holds the issue port
// for 2 cycles (16 fp64 lanes per sub-partition), every other instruction for 1.  The body below is written
// to minimise (2 x fp64 + other) instructions per member-timestep: ~21 fp64 + ~22 other without the pow,
// ~27 fp64 + ~11 other more with it (61.6 warp-instructions on average on the bench forcing, ncu).
template <bool WRITEQ, bool STORAGE, bool OBJ>
__global__ void hbv_fast_kernel(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
                                const double* __restrict__ params, int64_t N, HbvOut out, Slab slab,
                                Objective obj, Batch batch, const uint32_t* __restrict__ fflag) {
    // FAST contract: finite precipitation and temperature (the snow routine below forms prec - prec for "no liquid
    // water" and reads temp < T_t off the sign of temp - T_t).  The
    // packer flags anything else and the PRECISE kernel launched right behind this one takes the whole launch.
    if (*fflag != 0u) return;
    HBV_BATCH_PROLOGUE
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = gi < N ? gi : N - 1;  // see hbv_precise_kernel
    const double* p = params + 11 * i;
    const double T_t = p[0], DD = p[1], FC = p[2], Beta = p[3], C = p[4], PWP = p[5];
    const double K_0 = p[6], K_1 = p[7], K_2 = p[8], K_p = p[9], L = p[10];
    double inv_FC = 1.0 / FC, inv_PWP = 1.0 / PWP;
    // max(0, s1 - L) is read off the sign bit below; numba's max(0, NaN) = 0 for a NaN threshold = an infinite one
    const double Lq = (L == L) ? L : __longlong_as_double(0x7FF0000000000000LL);
    // temp < T_t is read off the sign of temp - T_t: a NaN threshold of either sign means "never cold", like the
    // reference's comparison (the subtraction then yields the canonical, positive NaN)
    // ... and a zero threshold is taken as -0.0: temp - (-0.0) is +0 for temp = +-0, like the reference's -0.0 < 0.0 = False
    const double Tt = (T_t == T_t) ? ((T_t == 0.0) ? -0.0 : T_t) : __longlong_as_double(0x7FF8000000000000LL);
    double c1 = 1.0 - K_1 - K_p;  // s1 (1 - K_1 - K_p)
    double c2 = 1.0 - K_2;        // s2 (1 - K_2)
    // The table-driven pow is used when soil/FC is within [2^-15, 2^15) and |Beta| < 32 (then
    // |Beta log2 x| < 512 and x is a positive normal).  The range test is one unsigned compare on the high
    // word of soil against per-member bounds derived from FC (never true for FC <= 0, NaN, inf, denormal).
    uint32_t safe_lo = 0u, safe_span = 0u;
    if (fabs(Beta) < 32.0 && FC > 0x1p-900 && FC < 0x1p900) {
        safe_lo = (uint32_t)__double2hiint(FC * 0x1p-15) + 1u;
        safe_span = (uint32_t)__double2hiint(FC * 0x1p15) - safe_lo;
    }
    int64_t stride = N;
    pin(inv_FC); pin(inv_PWP); pin(c1); pin(c2); pin(safe_lo); pin(safe_span); pin(stride);

    double snow = snow0, soil = soil0, s1 = s10, s2 = s20;  // hbvedu_model.py:78-81
    double acc = 0.0;
    int64_t t_first = slab.t_begin;
    int64_t off = i + (slab.t_begin - slab.row0) * N;
    if (slab.t_begin > 0) {
        snow = slab.state[0 * N + i];
        soil = slab.state[1 * N + i];
        s1 = slab.state[2 * N + i];
        s2 = slab.state[3 * N + i];
        if (OBJ) acc = slab.state[4 * N + i];
    } else {
        if (WRITEQ) st_stream(out.qsim + off, 0.0);
        if (STORAGE) {
            st_stream(out.snow + off, snow);
            st_stream(out.soil + off, soil);
            st_stream(out.s1 + off, s1);
            st_stream(out.s2 + off, s2);
        }
        if (OBJ) {
            const double d = obj.qobs[0];
            acc = d * d;
        }
        off += N;
        t_first = 1;
    }
    double* q_o = WRITEQ ? out.qsim + off : nullptr;
    double* snow_o = STORAGE ? out.snow + off : nullptr;
    double* soil_o = STORAGE ? out.soil + off : nullptr;
    double* s1_o = STORAGE ? out.s1 + off : nullptr;
    double* s2_o = STORAGE ? out.s2 + off : nullptr;

    extern __shared__ __align__(128) unsigned char rrb_smem[];
    uint32_t tb = smem_u32(fastmath_tables_to_smem(rrb_smem + forcing_smem_bytes<kHbvR, kHbvTT>()));
    pin(tb);
    __syncthreads();  // the staged tables are visible
    const PowRegs pr = load_pow_regs(tb);

    stream_forcing_grouped<kHbvR, kHbvTT, kHbvGroup, HbvF>(
        F, t_first, slab.t_end, [&](auto gc, int64_t t0, const HbvF* f) {
            constexpr int G = decltype(gc)::value;
            double liquid[G], pe[G], snow_g[G];
            bool need[G];
            // ---- A: snow routine (hbvedu_model.py:87-96), potential evapotranspiration (:102).
            // Both branches are evaluated and selected.  max(0, snow - m) and min(snow, m) share one
            // predicate: snow - m > 0 <=> m < snow for every operand pair (NaN and inf included).
#pragma unroll
            for (int g = 0; g < G; ++g) {
                // melt = min(snow, DD (temp - T_t)) serves both max(0, snow - m) = snow - melt and the liquid water
                // prec + melt; the cold branch (snow + prec, no liquid water) is the same two additions with
                // -prec in place of melt: snow - (-prec) and prec + (-prec) = +0 for finite precipitation
                const double dtt = f[g].temp - Tt;
                const double m = DD * dtt;
                const double melt = (m < snow) ? m : snow;
                const bool cold = __double2hiint(dtt) < 0;  // temp < T_t through the sign of the (finite) difference
                const double nprec = __hiloint2double(__double2hiint(f[g].prec) ^ (int)0x80000000, __double2loint(f[g].prec));
                const double sel = cold ? nprec : melt;
                snow = snow - sel;
                liquid[g] = f[g].prec + sel;
                snow_g[g] = snow;
                pe[g] = fma(C, f[g].dT, f[g].PEm);  // FAST packing: dT holds dT * PEm
            }
            // ---- prec_eff = liquid * (soil/FC)^Beta (:99) is +0 whenever liquid == 0 and the power is
            // finite, so a warp evaluates the pow only if one of its members has liquid water
#pragma unroll
            for (int g = 0; g < G; ++g)  // liquid != 0 on the bit pattern (one LOP3; -0 and NaN count as water)
                need[g] = __any_sync(0xffffffffu, (__double2hiint(liquid[g]) | __double2loint(liquid[g])) != 0);
            // ---- B: soil moisture, response routine, discharge
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const bool safe = ((uint32_t)__double2hiint(soil) - safe_lo) < safe_span;
                const double ea = (soil > PWP) ? pe[g] : pe[g] * (soil * inv_PWP);  // :105-108
                const double oK = max0_sane(s1 - Lq) * K_0;  // sign-bit select: a NaN s1 poisons s1_new and q either way
                const double s2_new = fma(s1, K_p, s2 * c2);                         // :121-123
                double s1_new = fma(s1, c1, -oK);                                    // :114-118 without prec_eff
                double soil_new = (soil + liquid[g]) - ea;                           // :111 without prec_eff
                if (need[g]) {
                    double pw = fast_pow_unchecked_smem(soil * inv_FC, Beta, tb, pr);
                    if (!safe) pw = hbv_slow_pow(soil, FC, Beta);
                    const double prec_eff = liquid[g] * pw;
                    soil_new -= prec_eff;
                    s1_new += prec_eff;
                } else if (!safe) {
                    const double prec_eff = liquid[g] * hbv_slow_pow(soil, FC, Beta);  // 0 * (inf | nan)
                    soil_new -= prec_eff;
                    s1_new += prec_eff;
                }
                soil = soil_new;
                s1 = s1_new;
                s2 = s2_new;
                const double qv = fma(s2_new, K_2, fma(s1_new, K_1, oK));            // :125-127
                if (WRITEQ) {
                    st_stream(q_o, qv);
                    q_o += stride;
                }
                if (STORAGE) {
                    st_stream(snow_o, snow_g[g]);
                    st_stream(soil_o, soil);
                    st_stream(s1_o, s1);
                    st_stream(s2_o, s2);
                    snow_o += stride;
                    soil_o += stride;
                    s1_o += stride;
                    s2_o += stride;
                }
                if (OBJ) {
                    const double d = obj.qobs[t0 + g] - qv;
                    acc += d * d;
                }
            }
        });

    if (gi < N) {
        if (slab.save_state) {
            slab.state[0 * N + i] = snow;
            slab.state[1 * N + i] = soil;
            slab.state[2 * N + i] = s1;
            slab.state[3 * N + i] = s2;
            if (OBJ) slab.state[4 * N + i] = acc;
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i] = acc / (double)obj.T;
    }
}


// ------------------------------------------------------------------------------------------------
// FAST, round 2 (hbv_fast2_kernel): organised for the DEPTH of the loop-carried soil chain.
//
// What bounds the round-1 kernel (profiles/r02_fp64_probe.txt, profiles/r01_ncu_full_hbv_v9_*): with one thread per
// member a 65 536-member ensemble leaves 3.5 warps per SM sub-partition; ncu shows issue 57 % and the fp64 pipe 48 %
// busy with `wait` (fixed-latency dependency) as the top stall -- the kernel runs at the latency of
// soil -> log2 -> x Beta -> exp2 -> soil (~215 cycles for a lone warp on a wet step), not at a pipe or issue limit.
// This kernel therefore
//   * evaluates the pow with the depth-organised sequence of rr_math.cuh (hbv_pow_step_twin is its CPU twin): hoisted
//     log2(FC), 512/1024-entry tables (one polynomial degree less per half), fused shift, late table scale: ~135
//     cycles of dependent latency and 19 instead of 25 fp64 instructions per wet member-step;
//   * has no per-step range check and no slow-path call in the time loop: a sticky per-thread maximum tracks whether
//     any soil moisture left the range of the table-driven pow; at the end the CTA votes and, if so, sets its flag
//     word and leaves its members to the PRECISE kernel queued behind (same for members outside the contract below,
//     decided before the loop).  The CTA's carry state / objective are then left untouched for that kernel;
//   * optionally runs TWO members per thread (MPT = 2: members 2j, 2j+1; one 16-byte streaming store per step, the
//     forcing loads, loop control, vote and addressing shared by both chains, which the compiler interleaves);
//   * addresses output rows as base + row * stride with a 32-bit row counter (one IMAD.WIDE per store).
// FAST contract per member (else the CTA is left to PRECISE): all parameters and initial states finite, FC and PWP in
// [2^-500, 2^500], |Beta| < 32.  Per launch (forcing flag, set by the packer): finite precipitation and temperature.
// ------------------------------------------------------------------------------------------------
struct HbvPowK {  // polynomial coefficients held in registers (an FMA takes one constant-bank operand)
    double a1, a2, a3, a4, c1, c2, c3;
};

template <int MPT, bool WRITEQ, bool STORAGE, bool OBJ>
__global__ void hbv_fast2_kernel(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
                                 const double* __restrict__ params, int64_t N, HbvOut out, Slab slab, Objective obj,
                                 Batch batch, uint32_t* __restrict__ fflag) {
    if (*fflag != 0u) return;  // non-finite forcing: the PRECISE kernel behind takes the whole launch
    HBV_BATCH_PROLOGUE
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nthreads = (N + MPT - 1) / MPT;   // MPT = 2 is launched for even N only
    // threads past the end of the ensemble recompute the last member(s) and store the same values again
    const int64_t i0 = MPT * (gi < nthreads ? gi : nthreads - 1);
    double T_t[MPT], DD[MPT], Beta[MPT], C[MPT], PWP[MPT], K_0[MPT], K_1[MPT], K_2[MPT], K_p[MPT], Lq[MPT];
    double inv_PWP[MPT], log2FC[MPT], c1[MPT], c2[MPT], Tt[MPT];
    uint32_t safe_lo[MPT], safe_span[MPT], worst[MPT];
    bool sane = true;
#pragma unroll
    for (int m = 0; m < MPT; ++m) {
        const double* p = params + 11 * (i0 + m);  // record order = HBVEdu._dtype (rrmpg/models/hbvedu.py:63-66)
        T_t[m] = p[0]; DD[m] = p[1]; const double FC = p[2]; Beta[m] = p[3]; C[m] = p[4]; PWP[m] = p[5];
        K_0[m] = p[6]; K_1[m] = p[7]; K_2[m] = p[8]; K_p[m] = p[9]; const double L = p[10];
#pragma unroll
        for (int k = 0; k < 11; ++k) sane = sane && (fabs(p[k]) <= 1e300);
        sane = sane && FC >= 0x1p-500 && FC <= 0x1p500 && PWP[m] >= 0x1p-500 && PWP[m] <= 0x1p500 && fabs(Beta[m]) < 32.0;
        inv_PWP[m] = 1.0 / PWP[m];
        log2FC[m] = log2(sane ? FC : 1.0);
        Lq[m] = L;
        // temp < T_t is read off the sign of temp - T_t; a zero threshold is taken as -0.0 so that the difference is
        // +0 for temp = +-0, like the reference's (-0.0 < 0.0) = False
        Tt[m] = (T_t[m] == 0.0) ? -0.0 : T_t[m];
        c1[m] = 1.0 - K_1[m] - K_p[m];  // s1 (1 - K_1 - K_p)
        c2[m] = 1.0 - K_2[m];           // s2 (1 - K_2)
        // soil/FC within [2^-15, 2^15): one unsigned compare on the high word of soil
        safe_lo[m] = (uint32_t)__double2hiint(FC * 0x1p-15) + 1u;
        safe_span[m] = (uint32_t)__double2hiint(FC * 0x1p15) - safe_lo[m];
        worst[m] = 0u;
        pin(inv_PWP[m]); pin(log2FC[m]); pin(c1[m]); pin(c2[m]); pin(safe_lo[m]);
    }
    sane = sane && fabs(snow0) <= 1e300 && fabs(soil0) <= 1e300 && fabs(s10) <= 1e300 && fabs(s20) <= 1e300;
    uint32_t* my_flag = hbv_cta_flags(fflag) + (blockIdx.y * gridDim.x + blockIdx.x);
    if (!__syncthreads_and(sane)) {  // CTA-uniform: a member outside the contract
        if (threadIdx.x == 0) *my_flag = 1u;
        return;
    }

    double snow[MPT], soil[MPT], s1[MPT], s2[MPT], acc[MPT];
#pragma unroll
    for (int m = 0; m < MPT; ++m) { snow[m] = snow0; soil[m] = soil0; s1[m] = s10; s2[m] = s20; acc[m] = 0.0; }  // hbvedu_model.py:78-81
    int64_t t_first = slab.t_begin;
    int64_t off = i0 + (slab.t_begin - slab.row0) * N;  // row r of the buffers = timestep row0 + r
    if (slab.t_begin > 0) {
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            snow[m] = slab.state[0 * N + i0 + m];
            soil[m] = slab.state[1 * N + i0 + m];
            s1[m] = slab.state[2 * N + i0 + m];
            s2[m] = slab.state[3 * N + i0 + m];
            if (OBJ) acc[m] = slab.state[4 * N + i0 + m];
        }
    } else {
        // t = 0 is not simulated (the reference loop starts at 1, hbvedu_model.py:84): qsim[0] = 0, storages = initial states
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            if (WRITEQ) st_stream(out.qsim + off + m, 0.0);
            if (STORAGE) {
                st_stream(out.snow + off + m, snow[m]);
                st_stream(out.soil + off + m, soil[m]);
                st_stream(out.s1 + off + m, s1[m]);
                st_stream(out.s2 + off + m, s2[m]);
            }
            if (OBJ) {
                const double d = obj.qobs[0];
                acc[m] = d * d;
            }
        }
        off += N;
        t_first = 1;
    }
    char* q_o = reinterpret_cast<char*>(WRITEQ ? out.qsim + off : nullptr);
    char* snow_o = reinterpret_cast<char*>(STORAGE ? out.snow + off : nullptr);
    char* soil_o = reinterpret_cast<char*>(STORAGE ? out.soil + off : nullptr);
    char* s1_o = reinterpret_cast<char*>(STORAGE ? out.s1 + off : nullptr);
    char* s2_o = reinterpret_cast<char*>(STORAGE ? out.s2 + off : nullptr);
    uint32_t row_bytes = (uint32_t)N * 8u;  // N < 2^29 (launch_hbvedu)
    uint32_t row = 0u;                      // rows written since t_first
    pin(row_bytes);
    auto put = [&](char* base, uint32_t r, const double* v) {
        double* p = reinterpret_cast<double*>(base + (uint64_t)r * (uint64_t)row_bytes);  // IMAD.WIDE.U32
        if (MPT == 2) st_stream_pair(p, v[0], v[MPT - 1]);
        else st_stream(p, v[0]);
    };

    extern __shared__ __align__(128) unsigned char rrb_smem[];
    uint32_t tb = smem_u32(hbv_tables_to_smem(rrb_smem + forcing_smem_bytes<kHbvR, kHbvTT>()));
    pin(tb);
    __syncthreads();  // the staged tables are visible
    const uint32_t pa = tb + (uint32_t)offsetof(HbvTables, poly);
    const HbvPowK pk{lds_f64_at(pa), lds_f64_at(pa + 8), lds_f64_at(pa + 16), lds_f64_at(pa + 24),
                     lds_f64_at(pa + 32), lds_f64_at(pa + 40), lds_f64_at(pa + 48)};

    stream_forcing_grouped<kHbvR, kHbvTT, kHbvGroup, HbvF>(
        F, t_first, slab.t_end, [&](auto gc, int64_t t0, const HbvF* f) {
            constexpr int G = decltype(gc)::value;
            double liquid[G][MPT], pe[G][MPT], pew[G][MPT], snow_g[G][MPT];
            bool need[G];
            // ---- A: snow routine (hbvedu_model.py:87-96), potential evapotranspiration (:102); off the soil chain.
            // melt = min(snow, DD (temp - T_t)) serves both max(0, snow - m) = snow - melt and the liquid water
            // prec + melt; the cold branch (snow + prec, no liquid water) is the same two additions with -prec in
            // place of melt: snow - (-prec) and prec + (-prec) = +0 for finite precipitation
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double nprec = __hiloint2double(__double2hiint(f[g].prec) ^ (int)0x80000000, __double2loint(f[g].prec));
                bool wet = false;
#pragma unroll
                for (int m = 0; m < MPT; ++m) {
                    const double dtt = f[g].temp - Tt[m];
                    const double mm = DD[m] * dtt;
                    const double melt = (mm < snow[m]) ? mm : snow[m];
                    const bool cold = __double2hiint(dtt) < 0;  // temp < T_t through the sign of the (finite) difference
                    const double sel = cold ? nprec : melt;
                    snow[m] = snow[m] - sel;
                    liquid[g][m] = f[g].prec + sel;
                    snow_g[g][m] = snow[m];
                    pe[g][m] = fma(C[m], f[g].dT, f[g].PEm);  // FAST packing: dT holds dT * PEm
                    pew[g][m] = pe[g][m] * inv_PWP[m];
                    wet = wet || ((__double2hiint(liquid[g][m]) | __double2loint(liquid[g][m])) != 0);  // -0 and NaN count as water
                }
                // prec_eff = liquid * (soil/FC)^Beta (:99) is +0 whenever liquid == 0 and the power is finite, so a warp
                // evaluates the pow only if one of its members has liquid water
                need[g] = __any_sync(0xffffffffu, wet);
            }
            // ---- B: soil moisture, response routine, discharge
#pragma unroll
            for (int g = 0; g < G; ++g) {
                double qv[MPT], sp[MPT], oK[MPT], s1_new[MPT], s2_new[MPT];
                uint32_t hs[MPT];
#pragma unroll
                for (int m = 0; m < MPT; ++m) {
                    hs[m] = (uint32_t)__double2hiint(soil[m]);
                    worst[m] = max(worst[m], hs[m] - safe_lo[m]);  // sticky range check, judged after the time loop
                    const double ea = (soil[m] > PWP[m]) ? pe[g][m] : pew[g][m] * soil[m];  // :105-108
                    sp[m] = (soil[m] + liquid[g][m]) - ea;                                   // :111 without prec_eff
                    oK[m] = max0_sane(s1[m] - Lq[m]) * K_0[m];
                    s2_new[m] = fma(s1[m], K_p[m], s2[m] * c2[m]);                           // :121-123
                    s1_new[m] = fma(s1[m], c1[m], -oK[m]);                                   // :114-118 without prec_eff
                }
                if (need[g]) {  // warp-uniform; the chains of the thread's members share one basic block and interleave
                    using namespace hbvpow;
#pragma unroll
                    for (int m = 0; m < MPT; ++m) {
                        // log2(soil) - log2(FC)
                        const double mant = __hiloint2double((int)((hs[m] & 0x000FFFFFu) | 0x3FF00000u), __double2loint(soil[m]));
                        double invc, log2c;
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(log2c) : "r"(tb + ((hs[m] >> 7) & 0x1FF0u)));
                        const double kml = (double)((int)(hs[m] >> 20) - 1023) - log2FC[m];
                        const double r = fma(mant, invc, -1.0);
                        const double base = kml + log2c;
                        const double r2 = r * r;
                        const double a = fma(r, pk.a2, pk.a1);
                        const double b = fma(r, pk.a4, pk.a3);
                        const double t = fma(r2, b, a);
                        const double Lg = fma(r, t, base);
                        // 2^(Beta Lg) = scale (1 + rr gg)
                        const double kd = fma(Beta[m], Lg, kShift);
                        const uint32_t ki = (uint32_t)__double2loint(kd);
                        uint32_t tlo, thi;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                                     : "=r"(tlo), "=r"(thi)
                                     : "r"(tb + (uint32_t)offsetof(HbvTables, exp2k) + ((ki & (uint32_t)(tables::kExpKN - 1)) << 3)));
                        const double kdm = kd - kShift;
                        const double rr = fma(Beta[m], Lg, -kdm);
                        const double scale = __hiloint2double((int)(thi + (ki << 10)), (int)tlo);  // bits + (ki << 42)
                        const double q2 = rr * rr;
                        const double e = fma(rr, pk.c2, pk.c1);
                        const double gg = fma(q2, pk.c3, e);
                        const double u = rr * gg;
                        const double w = fma(liquid[g][m], u, liquid[g][m]);
                        soil[m] = fma(-scale, w, sp[m]);
                        s1_new[m] += scale * w;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < MPT; ++m) soil[m] = sp[m];
                }
#pragma unroll
                for (int m = 0; m < MPT; ++m) {
                    s1[m] = s1_new[m];
                    s2[m] = s2_new[m];
                    qv[m] = fma(s2_new[m], K_2[m], fma(s1_new[m], K_1[m], oK[m]));           // :125-127
                    if (OBJ) {
                        const double d = obj.qobs[t0 + g] - qv[m];
                        acc[m] += d * d;
                    }
                }
                if (WRITEQ) put(q_o, row, qv);
                if (STORAGE) {
                    put(snow_o, row, snow_g[g]);
                    put(soil_o, row, soil);
                    put(s1_o, row, s1);
                    put(s2_o, row, s2);
                }
                ++row;
            }
        });

    // did any soil moisture leave the range of the table-driven pow?  Then this CTA's results are void: flag it for
    // the PRECISE kernel behind and leave the carry state / objective as they were.
    bool bad = false;
#pragma unroll
    for (int m = 0; m < MPT; ++m) bad = bad || (worst[m] >= safe_span[m]);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) *my_flag = bad ? 1u : 0u;
    if (bad) return;
    if (gi < nthreads) {
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            if (slab.save_state) {
                slab.state[0 * N + i0 + m] = snow[m];
                slab.state[1 * N + i0 + m] = soil[m];
                slab.state[2 * N + i0 + m] = s1[m];
                slab.state[3 * N + i0 + m] = s2[m];
                if (OBJ) slab.state[4 * N + i0 + m] = acc[m];
            }
            if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i0 + m] = acc[m] / (double)obj.T;
        }
    }
}

int state_slots_hbvedu() { return 5; }

// Which FAST kernel runs: 1 / 2 = hbv_fast2_kernel with one / two members per thread (round 2), 0 = the round-1
// kernel (kept for A/B timing).  Default: two members per thread when the rows allow 16-byte stores.
static int hbv_variant() {
    static const int v = [] {
        const char* e = getenv("RRMPG_B200_HBV_VARIANT");
        return e ? atoi(e) : -1;
    }();
    return v;
}

cudaError_t launch_hbvedu(const double* F, int64_t T, const double* inits4, const double* params, int64_t N,
                          double* qsim, double* snow, double* soil, double* s1, double* s2, const Slab& slab,
                          const Objective& obj, const LaunchCfg& cfg, const uint32_t* fflag, const Batch& batch) {
    (void)T;
    if (N <= 0 || batch.count <= 0) return cudaSuccess;
    if (N >= (int64_t(1) << 29)) return cudaErrorInvalidValue;  // 32-bit row pitch in bytes (hbv_fast2_kernel)
    const bool fast = cfg.math == RRB_MATH_FAST_;
    const bool st = snow != nullptr, ob = obj.qobs != nullptr, wq = qsim != nullptr;
    HbvOut out{qsim, snow, soil, s1, s2};
    int variant = cfg.variant > 0 ? (cfg.variant == 3 ? 0 : cfg.variant) : hbv_variant();
    auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; };
    const bool pair_ok = (N % 2) == 0 && aligned16(qsim) && aligned16(snow) && aligned16(soil) && aligned16(s1) && aligned16(s2) &&
                         (!slab.state || aligned16(slab.state));
    if (variant < 0) variant = RRB_HBV_DEFAULT_VARIANT;
    if (variant == 2 && !pair_ok) variant = 1;
    const int mpt = (fast && variant == 2) ? 2 : 1;
    const int64_t nthreads = (N + mpt - 1) / mpt;
    const int block = cfg.block > 0 ? cfg.block : pick_block(nthreads * batch.count, cfg.sm_count, nthreads >= 256 ? 256 : 64);
    const dim3 grid((unsigned)((nthreads + block - 1) / block), (unsigned)batch.count);
    const dim3 grid_p((unsigned)((N + block - 1) / block), (unsigned)batch.count);  // PRECISE: one member per thread
    const size_t smem_ring = forcing_smem_bytes<kHbvR, kHbvTT>();
#define RRB_HBV_ARGS F, inits4[0], inits4[1], inits4[2], inits4[3], params, N, out, slab, obj, batch
#define RRB_HBV_DISPATCH(LAUNCH_)                          \
    do {                                                   \
        if (wq && st && ob) LAUNCH_(true, true, true);     \
        else if (wq && st) LAUNCH_(true, true, false);     \
        else if (wq && ob) LAUNCH_(true, false, true);     \
        else if (wq) LAUNCH_(true, false, false);          \
        else if (st && ob) LAUNCH_(false, true, true);     \
        else if (st) LAUNCH_(false, true, false);          \
        else LAUNCH_(false, false, true);                  \
    } while (0)
#define RRB_HBV_PRECISE(Q_, S_, O_) \
    hbv_precise_kernel<Q_, S_, O_><<<grid_p, block, smem_ring, cfg.stream>>>(RRB_HBV_ARGS, p_flag, p_packing, (int)grid.x, p_div)
    const uint32_t* p_flag = nullptr;
    int p_packing = 0, p_div = 0;
    if (!fast) {
        RRB_HBV_DISPATCH(RRB_HBV_PRECISE);
        return cudaGetLastError();
    }
    uint32_t* wflag = const_cast<uint32_t*>(fflag);
    if (variant == 0) {
        const size_t smem = smem_ring + fastmath_smem_bytes();
#define RRB_HBV_FAST1(Q_, S_, O_) hbv_fast_kernel<Q_, S_, O_><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, fflag)
        RRB_HBV_DISPATCH(RRB_HBV_FAST1);
#undef RRB_HBV_FAST1
    } else {
        const size_t smem = smem_ring + hbv_tables_smem_bytes();
#define RRB_HBV_FAST2A(Q_, S_, O_) hbv_fast2_kernel<1, Q_, S_, O_><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag)
#define RRB_HBV_FAST2B(Q_, S_, O_) hbv_fast2_kernel<2, Q_, S_, O_><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag)
        if (mpt == 2) RRB_HBV_DISPATCH(RRB_HBV_FAST2B);
        else RRB_HBV_DISPATCH(RRB_HBV_FAST2A);
#undef RRB_HBV_FAST2A
#undef RRB_HBV_FAST2B
        p_div = mpt;  // the PRECISE launch honours the per-CTA flags
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the fallback for a flagged forcing block / flagged CTAs: exits at once otherwise (FAST packing: dT slot = dT * PEm)
    p_flag = fflag;
    p_packing = 1;
    RRB_HBV_DISPATCH(RRB_HBV_PRECISE);
#undef RRB_HBV_PRECISE
#undef RRB_HBV_DISPATCH
#undef RRB_HBV_ARGS
    return cudaGetLastError();
}

}  // namespace rrb
