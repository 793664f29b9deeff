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
//                 ~1e-15 relative), the pow skipped for warps whose members all have liquid_water == 0
//                 (then prec_eff = 0 * finite = +0 exactly as in the reference), the snow routine skipped for
//                 groups of steps that are warm-and-bare or cold for a whole warp (hbv_fast2_loop).
//                 Discharge stays within rtol 1e-10 of the reference (tests/test_parity_gpu.py).
// Both: branch-free snow routine, running output addressing, t = 0 peeled out of the time loop.
// Kernels: hbv_precise_kernel; hbv_fast2_kernel / hbv_fast2_wide_kernel (one or two members per thread, register cap
// or not) and hbv_rot_kernel (rotating schedule, opt-in) around the shared time loop hbv_fast2_loop; launch_hbvedu
// picks members per thread and the launch shape from the ensemble size (one CTA per SM for mid-sized ensembles).
#include "rr_common.cuh"
#include "rr_kernels.h"
#include "rr_math.cuh"
#include "rr_objective.cuh"

#include <cstdlib>
#include <type_traits>

#ifndef RRB_HBV_DEFAULT_VARIANT
#define RRB_HBV_DEFAULT_VARIANT 2
#endif
#ifndef RRB_HBV_SEASONS
#define RRB_HBV_SEASONS 1  // short A phases for groups that are warm-and-bare or cold for a whole warp (hbv_fast2_loop)
#endif
#ifndef RRB_HBV_HORNER
#define RRB_HBV_HORNER 0  // measured: 3.06 vs 2.88 ms at 65 536 members (the uniform-register constants bring BRA.DIV back)
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
        // month index outside [0, 11]: the reference would read past PE_m / T_m (numba does no bounds checking);
        // the Python layer validates host arrays, the kernel stays memory safe for any int8 a device caller passes
        const int m = min(max((int)month0[t], 0), 11);
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
        if (obj.obs_stats) { obj.obs_mean = obj.obs_stats[2 * c]; obj.obs_std = obj.obs_stats[2 * c + 1]; } \
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
    ObjAcc acc;
    acc.reset();
    int64_t t_first = slab.t_begin;
    int64_t off = i + (slab.t_begin - slab.row0) * N;  // row r of the buffers = timestep row0 + r
    if (slab_loads_state(slab)) {
        snow = slab.state[0 * N + i];
        soil = slab.state[1 * N + i];
        s1 = slab.state[2 * N + i];
        s2 = slab.state[3 * N + i];
        if (OBJ && slab.t_begin > 0) acc.load(slab.state, 4, N, i, obj);
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
        if (OBJ) acc.add(obj.qobs[0], 0.0, obj);
        off += N;
        t_first = 1;
    }
    double* q_o = WRITEQ ? out.qsim + off : nullptr;
    double* snow_o = STORAGE ? out.snow + off : nullptr;
    double* soil_o = STORAGE ? out.soil + off : nullptr;
    double* s1_o = STORAGE ? out.s1 + off : nullptr;
    double* s2_o = STORAGE ? out.s2 + off : nullptr;

    stream_forcing_regs<kHbvR, kHbvTTSmall, HbvF>(F, t_first, slab.t_end, [&](int64_t t, const HbvF& f) {
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
        if (OBJ) acc.add(obj.qobs[t], qv, obj);
    });

    if (gi < N) {
        if (slab.save_state) {
            slab.state[0 * N + i] = snow;
            slab.state[1 * N + i] = soil;
            slab.state[2 * N + i] = s1;
            slab.state[3 * N + i] = s2;
            if (OBJ && slab.save_state == 1) acc.save(slab.state, 4, N, i, obj);
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i] = acc.finish(obj);
    }
}

constexpr int kHbvGroup = 2;  // timesteps per group of the time loop

struct HbvPowK {  // polynomial coefficients held in registers (an FMA takes one constant-bank operand)
    double a1, a2, a3, a4, c1, c2, c3;
};

// Per-member constants of the FAST step, MPT members per thread (members i0 .. i0 + MPT - 1).
template <int MPT>
struct HbvPar {
    double DD[MPT], Beta[MPT], C[MPT], PWP[MPT], K_0[MPT], K_1[MPT], K_2[MPT], K_p[MPT], Lq[MPT];
    double inv_PWP[MPT], log2FC[MPT], c1[MPT], c2[MPT], Tt[MPT];
    uint32_t safe_lo[MPT], safe_span[MPT];
    // false: a member is outside the FAST contract (the constants are then meaningless, the caller flags its block)
    __device__ __forceinline__ bool load(const double* __restrict__ params, int64_t i0) {
        bool sane = true;
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            const double* p = params + 11 * (i0 + m);  // record order = HBVEdu._dtype (rrmpg/models/hbvedu.py:63-66)
            const double T_t = p[0];
            DD[m] = p[1]; const double FC = p[2]; Beta[m] = p[3]; C[m] = p[4]; PWP[m] = p[5];
            K_0[m] = p[6]; K_1[m] = p[7]; K_2[m] = p[8]; K_p[m] = p[9]; const double L = p[10];
#pragma unroll
            for (int k = 0; k < 11; ++k) sane = sane && (fabs(p[k]) <= 1e300);
            sane = sane && FC >= 0x1p-500 && FC <= 0x1p500 && PWP[m] >= 0x1p-500 && PWP[m] <= 0x1p500 && fabs(Beta[m]) < 32.0;
            inv_PWP[m] = 1.0 / PWP[m];
            log2FC[m] = log2(sane ? FC : 1.0);
            Lq[m] = L;
            // temp < T_t is read off the sign of temp - T_t; a zero threshold is taken as -0.0 so that the difference is
            // +0 for temp = +-0, like the reference's (-0.0 < 0.0) = False
            Tt[m] = (T_t == 0.0) ? -0.0 : T_t;
            c1[m] = 1.0 - K_1[m] - K_p[m];  // s1 (1 - K_1 - K_p)
            c2[m] = 1.0 - K_2[m];           // s2 (1 - K_2)
            // soil/FC within [2^-15, 2^15): one unsigned compare on the high word of soil
            safe_lo[m] = (uint32_t)__double2hiint(FC * 0x1p-15) + 1u;
            safe_span[m] = (uint32_t)__double2hiint(FC * 0x1p15) - safe_lo[m];
            pin(inv_PWP[m]); pin(log2FC[m]); pin(c1[m]); pin(c2[m]); pin(safe_lo[m]);
        }
        return sane;
    }
};

// The stores a member carries from step to step, the sticky range check and the objective sums.
template <int MPT, int OBJ>
struct HbvSt {
    double snow[MPT], soil[MPT], s1[MPT], s2[MPT];
    uint32_t worst[MPT];  // max over the steps of hi(soil) - safe_lo: >= safe_span when a soil moisture left the table range
    typename std::conditional<OBJ == 2, ObjAccKge, ObjAccSse>::type acc[MPT];
    __device__ __forceinline__ bool left_range(const HbvPar<MPT>& P) const {
        bool bad = false;
#pragma unroll
        for (int m = 0; m < MPT; ++m) bad = bad || (worst[m] >= P.safe_span[m]);
        return bad;
    }
};

// Output rows: base pointers at the row of the first step of the loop, addressed as base + row * row_bytes.
struct HbvRows {
    char *q, *snow, *soil, *s1, *s2;
    uint32_t row_bytes;  // N * 8 (N < 2^29, launch_hbvedu)
    uint32_t row;        // rows written since the first step
};

// ------------------------------------------------------------------------------------------------
// The time loop of the FAST kernels: groups of two timesteps, four straight-line bodies.
// A(t): snow routine (hbvedu_model.py:87-96) and potential evapotranspiration (:102) -- needs forcing[t], the snow pack
//       and parameters only: short chains, independent of the soil / response stores.
// B(t): soil moisture (:99-111), response routine (:114-123), discharge (:125-127) -- the loop-carried chains.
// A warp issues in order, so a basic block runs at the latency of its longest dependent chain unless it holds enough
// independent work, and with 2-4 warps per SM sub-partition nothing else hides that latency (profiles/r02_*).  The
// pow of B(t) is a ~190-cycle chain that is skipped when no member of the warp has liquid water (warp vote); a branch
// per step would cut the loop body into short blocks, each as slow as its own chain.  So the vote of BOTH steps of a
// group is taken first (in the A phase) and selects one of four bodies -- (wet|dry, wet|dry) -- each a single basic
// block holding B of the two steps.  (Software-pipelining A of the next group into those bodies raised issue
// utilisation to 55 % but cost 14 % more instructions -- register moves at the merge points -- for equal time: removed.)
//
// `run(group)` streams the forcing: it calls group(ic<2>, t0, f[2]) for aligned pairs of steps and group(ic<1>, t, f[1])
// for the ragged steps at the edges (stream_forcing_grouped for a CTA-wide ring, WarpRing::run for a warp's own).
// ABL: timing ablations for the discharge-only instantiation (development builds only, results are WRONG):
//   1 no output stores, 2 never wet, 3 always wet, 4 table loads replaced by constants
// ------------------------------------------------------------------------------------------------
template <int MPT, bool WRITEQ, bool STORAGE, int OBJ, int ABL, class Run>
__device__ __forceinline__ void hbv_fast2_loop(const HbvPar<MPT>& P, HbvSt<MPT, OBJ>& S, HbvRows& O, uint32_t tb,
                                               const HbvPowK& pk, const Objective& obj, Run&& run) {
    static_assert(kHbvGroup == 2, "the four group bodies are written for two timesteps per group");
    constexpr int GP = 2;
    auto put = [&](char* base, uint32_t r, const double* v) __attribute__((always_inline)) {
        double* p = reinterpret_cast<double*>(base + (uint64_t)r * (uint64_t)O.row_bytes);  // IMAD.WIDE.U32
        if constexpr (ABL == 1) {
            if (v[0] == 1.2345e-300) st_stream(p, v[MPT - 1]);  // keeps the value alive, (almost) never stores
        } else if (MPT == 2) st_stream_pair(p, v[0], v[MPT - 1]);
        else st_stream(p, v[0]);
    };
    struct AOut {
        double liquid[GP][MPT], pe[GP][MPT], pew[GP][MPT], snow_g[GP][MPT];
        bool need[GP];
    };
    AOut cur = {};      // A results of the group whose B phase runs next
    int64_t t_cur = 0;  // first timestep of `cur`

    auto phase_a = [&](const HbvF& f, AOut& o, int g) __attribute__((always_inline)) {
        // melt = min(snow, DD (temp - T_t)) serves both max(0, snow - m) = snow - melt and the liquid water prec + melt;
        // the cold branch (snow + prec, no liquid water) is the same two additions with -prec in place of melt:
        // snow - (-prec) and prec + (-prec) = +0 for finite precipitation
        const double nprec = __hiloint2double(__double2hiint(f.prec) ^ (int)0x80000000, __double2loint(f.prec));
        bool wet = false;
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            const double dtt = f.temp - P.Tt[m];
            const double mm = P.DD[m] * dtt;
            const double melt = (mm < S.snow[m]) ? mm : S.snow[m];
            const bool cold = __double2hiint(dtt) < 0;  // temp < T_t through the sign of the (finite) difference
            const double sel = cold ? nprec : melt;
            S.snow[m] = S.snow[m] - sel;
            o.liquid[g][m] = f.prec + sel;
            o.snow_g[g][m] = S.snow[m];
            o.pe[g][m] = fma(P.C[m], f.dT, f.PEm);  // FAST packing: dT holds dT * PEm
            o.pew[g][m] = o.pe[g][m] * P.inv_PWP[m];
            wet = wet || ((__double2hiint(o.liquid[g][m]) | __double2loint(o.liquid[g][m])) != 0);  // -0 and NaN count as water
        }
        // prec_eff = liquid * (soil/FC)^Beta (:99) is +0 whenever liquid == 0 and the power is finite, so a warp
        // evaluates the pow only if one of its members has liquid water
        o.need[g] = __any_sync(0xffffffffu, wet);
        if constexpr (ABL == 2) o.need[g] = false;
        if constexpr (ABL == 3) o.need[g] = true;
    };
    // Seasons of a warp.  The snow routine is 17 of the 22 issue slots of A(t) per member, and for most of the year it
    // does nothing that depends on the member: when the step is warm for EVERY member of the warp (temp above the largest
    // threshold) and no member has snow left, melt = min(snow, m) = +0, the pack stays empty and the liquid water is the
    // precipitation; when it is cold for every member, the precipitation joins the pack and there is no liquid water.
    // Both conditions are warp-uniform (the forcing is the same for all members; tt_hi / tt_lo are warp-reduced once; the
    // snow-free flag is voted after every full A phase), so the two steps of a group take a short A phase instead -- the
    // same values bit for bit.  Temperatures are compared through their high words (one integer compare each, an
    // unordered or borderline value simply takes the full phase).
    double tt_hi = P.Tt[0], tt_lo = P.Tt[0];
#pragma unroll
    for (int m = 1; m < MPT; ++m) { tt_hi = fmax(tt_hi, P.Tt[m]); tt_lo = fmin(tt_lo, P.Tt[m]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tt_hi = fmax(tt_hi, __shfl_xor_sync(0xffffffffu, tt_hi, o));
        tt_lo = fmin(tt_lo, __shfl_xor_sync(0xffffffffu, tt_lo, o));
    }
    // warm for all: hi(temp) > hi(max(tt_hi, +0)) as signed integers (both non-negative then: a larger high word is a larger value)
    // (a negative degree-day factor would make m = DD (temp - T_t) < 0 = snow: no short phase for such a warp)
    bool dd_ok = true;
#pragma unroll
    for (int m = 0; m < MPT; ++m) dd_ok = dd_ok && !(P.DD[m] < 0.0);
    int warm_key = __all_sync(0xffffffffu, dd_ok) ? __double2hiint(fmax(tt_hi, 0.0)) : 0x7fffffff;
    // cold for all: temp < min(tt_lo, -0) <= 0: both negative, a larger high word (unsigned) is a smaller value
    uint32_t cold_key = (uint32_t)__double2hiint(fmin(tt_lo, -0.0)) | 0x80000000u;
    pin(cold_key);
    {
        uint32_t wk = (uint32_t)warm_key;
        pin(wk);
        warm_key = (int)wk;
    }
    auto any_snow = [&]() __attribute__((always_inline)) {
        bool some = false;
#pragma unroll
        for (int m = 0; m < MPT; ++m) some = some || ((__double2hiint(S.snow[m]) | __double2loint(S.snow[m])) != 0);
        return __any_sync(0xffffffffu, some) != 0;
    };
    bool snow_free = !any_snow();
    auto a_potential_et = [&](const HbvF& f, AOut& o, int g) __attribute__((always_inline)) {
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            o.snow_g[g][m] = S.snow[m];
            o.pe[g][m] = fma(P.C[m], f.dT, f.PEm);
            o.pew[g][m] = o.pe[g][m] * P.inv_PWP[m];
        }
    };
    // warm for every member, no snow anywhere: sel = melt = +0, snow - 0 = snow, liquid = prec + 0
    auto a_warm_bare = [&](const HbvF& f, AOut& o, int g) __attribute__((always_inline)) {
        const double lq = f.prec + 0.0;  // (-0.0 + +0.0 = +0.0, as in the full phase)
#pragma unroll
        for (int m = 0; m < MPT; ++m) o.liquid[g][m] = lq;
        a_potential_et(f, o, g);
        o.need[g] = __any_sync(0xffffffffu, (__double2hiint(lq) | __double2loint(lq)) != 0);
        if constexpr (ABL == 2) o.need[g] = false;
        if constexpr (ABL == 3) o.need[g] = true;
    };
    // cold for every member: sel = -prec, snow - (-prec), liquid = prec + (-prec) = +0 (finite precipitation)
    auto a_cold = [&](const HbvF& f, AOut& o, int g) __attribute__((always_inline)) {
        const double nprec = __hiloint2double(__double2hiint(f.prec) ^ (int)0x80000000, __double2loint(f.prec));
        const double lq = f.prec + nprec;
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            S.snow[m] = S.snow[m] - nprec;
            o.liquid[g][m] = lq;
        }
        a_potential_et(f, o, g);
        o.need[g] = false;
        if constexpr (ABL == 3) o.need[g] = true;
    };
    // B of step g of group `a`; WET = the warp evaluates the pow for this step
    auto phase_b = [&](auto wet_c, const AOut& a, int g) __attribute__((always_inline)) {
        constexpr bool WET = decltype(wet_c)::value != 0;
        double qv[MPT], sp[MPT], oK[MPT], s1_new[MPT], s2_new[MPT];
        uint32_t hs[MPT];
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            hs[m] = (uint32_t)__double2hiint(S.soil[m]);
            S.worst[m] = max(S.worst[m], hs[m] - P.safe_lo[m]);  // sticky range check, judged after the time loop
            const double ea = (S.soil[m] > P.PWP[m]) ? a.pe[g][m] : a.pew[g][m] * S.soil[m];  // :105-108
            sp[m] = (S.soil[m] + a.liquid[g][m]) - ea;                                         // :111 without prec_eff
            oK[m] = max0_sane(S.s1[m] - P.Lq[m]) * P.K_0[m];
            s2_new[m] = fma(S.s1[m], P.K_p[m], S.s2[m] * P.c2[m]);                             // :121-123
            s1_new[m] = fma(S.s1[m], P.c1[m], -oK[m]);                                         // :114-118 without prec_eff
        }
        if constexpr (WET) {
            // written stage by stage ACROSS the members so that the instruction order handed to the assembler already
            // interleaves their chains
            using namespace hbvpow;
            double mant[MPT], invc[MPT], log2c[MPT], kml[MPT], r[MPT], Lg[MPT], kd[MPT], rr[MPT], scale[MPT], w[MPT];
            uint32_t ki[MPT], tlo[MPT], thi[MPT];
#pragma unroll
            for (int m = 0; m < MPT; ++m) {  // log2(soil) - log2(FC): table lookup keyed on the top 9 mantissa bits
                if constexpr (ABL == 4) { invc[m] = pk.a1 + (double)(hs[m] >> 7); log2c[m] = pk.a2; }
                else
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc[m]), "=d"(log2c[m]) : "r"(tb + ((hs[m] >> 7) & 0x1FF0u)));
                mant[m] = __hiloint2double((int)((hs[m] & 0x000FFFFFu) | 0x3FF00000u), __double2loint(S.soil[m]));
                kml[m] = (double)((int)(hs[m] >> 20) - 1023) - P.log2FC[m];
            }
#pragma unroll
            for (int m = 0; m < MPT; ++m) r[m] = fma(mant[m], invc[m], -1.0);
#if RRB_HBV_HORNER
            // Horner with ONE register-resident coefficient per polynomial: every other coefficient is a constant-bank
            // operand, so each step reads two registers instead of three (the binding resource)
            double t3[MPT], t2[MPT], t1[MPT];
#pragma unroll
            for (int m = 0; m < MPT; ++m) t3[m] = fma(r[m], A4, pk.a3);
#pragma unroll
            for (int m = 0; m < MPT; ++m) t2[m] = fma(t3[m], r[m], A2);
#pragma unroll
            for (int m = 0; m < MPT; ++m) t1[m] = fma(t2[m], r[m], A1);
#pragma unroll
            for (int m = 0; m < MPT; ++m) Lg[m] = fma(r[m], t1[m], kml[m] + log2c[m]);
#else
#pragma unroll
            for (int m = 0; m < MPT; ++m) {
                const double base = kml[m] + log2c[m];
                const double r2 = r[m] * r[m];
                const double pa = fma(r[m], pk.a2, pk.a1);
                const double pb = fma(r[m], pk.a4, pk.a3);
                const double t = fma(r2, pb, pa);
                Lg[m] = fma(r[m], t, base);
            }
#endif
#pragma unroll
            for (int m = 0; m < MPT; ++m) {  // 2^(Beta Lg) = scale (1 + rr gg)
                kd[m] = fma(P.Beta[m], Lg[m], kShift);
                ki[m] = (uint32_t)__double2loint(kd[m]);
                if constexpr (ABL == 4) { tlo[m] = ki[m] & 1023u; thi[m] = 0x3FF00000u; }
                else
                asm("ld.shared.v2.u32 {%0, %1}, [%2];"
                    : "=r"(tlo[m]), "=r"(thi[m])
                    : "r"(tb + (uint32_t)offsetof(HbvTables, exp2k) + ((ki[m] & (uint32_t)(tables::kExpKN - 1)) << 3)));
            }
#pragma unroll
            for (int m = 0; m < MPT; ++m) {
                const double kdm = kd[m] - kShift;
                rr[m] = fma(P.Beta[m], Lg[m], -kdm);
                scale[m] = __hiloint2double((int)(thi[m] + (ki[m] << 10)), (int)tlo[m]);  // bits + (ki << 42)
            }
#pragma unroll
            for (int m = 0; m < MPT; ++m) {
#if RRB_HBV_HORNER
                const double e = fma(rr[m], C3, pk.c2);
                const double gg = fma(e, rr[m], C1);
#else
                const double q2 = rr[m] * rr[m];
                const double e = fma(rr[m], pk.c2, pk.c1);
                const double gg = fma(q2, pk.c3, e);
#endif
                const double u = rr[m] * gg;
                w[m] = fma(a.liquid[g][m], u, a.liquid[g][m]);
            }
#pragma unroll
            for (int m = 0; m < MPT; ++m) {
                S.soil[m] = fma(-scale[m], w[m], sp[m]);
                s1_new[m] += scale[m] * w[m];
            }
        } else {
#pragma unroll
            for (int m = 0; m < MPT; ++m) S.soil[m] = sp[m];
        }
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            S.s1[m] = s1_new[m];
            S.s2[m] = s2_new[m];
            qv[m] = fma(s2_new[m], P.K_2[m], fma(s1_new[m], P.K_1[m], oK[m]));               // :125-127
            if (OBJ) S.acc[m].add(obj.qobs[t_cur + g], qv[m], obj);
        }
        if (WRITEQ) put(O.q, O.row, qv);
        if (STORAGE) {
            put(O.snow, O.row, a.snow_g[g]);
            put(O.soil, O.row, S.soil);
            put(O.s1, O.row, S.s1);
            put(O.s2, O.row, S.s2);
        }
        ++O.row;
    };
    // B of both steps of the group in `cur`: one of four straight-line bodies, selected by the two warp votes
    auto run_group_b = [&]() __attribute__((always_inline)) {
        if (cur.need[0]) {
            if (cur.need[1]) { phase_b(ic<1>{}, cur, 0); phase_b(ic<1>{}, cur, 1); }
            else             { phase_b(ic<1>{}, cur, 0); phase_b(ic<0>{}, cur, 1); }
        } else {
            if (cur.need[1]) { phase_b(ic<0>{}, cur, 0); phase_b(ic<1>{}, cur, 1); }
            else             { phase_b(ic<0>{}, cur, 0); phase_b(ic<0>{}, cur, 1); }
        }
    };

    run([&](auto gc, int64_t t0, const HbvF* f) __attribute__((always_inline)) {
        constexpr int G = decltype(gc)::value;
        if constexpr (G == GP) {
            const int h0 = __double2hiint(f[0].temp), h1 = __double2hiint(f[1].temp);
            const bool warm = snow_free && __all_sync(0xffffffffu, (h0 > warm_key) & (h1 > warm_key));
            const bool cold = __all_sync(0xffffffffu, ((uint32_t)h0 > cold_key) & ((uint32_t)h1 > cold_key));
            if (RRB_HBV_SEASONS && warm) {
                a_warm_bare(f[0], cur, 0);
                a_warm_bare(f[1], cur, 1);
            } else if (RRB_HBV_SEASONS && cold) {
                a_cold(f[0], cur, 0);
                a_cold(f[1], cur, 1);
                snow_free = false;
            } else {
                phase_a(f[0], cur, 0);   // A of both steps: four independent short chains per member pair
                phase_a(f[1], cur, 1);
                if (RRB_HBV_SEASONS) snow_free = !any_snow();
            }
            t_cur = t0;
            run_group_b();
        } else {  // a ragged step at the edge of a time slab
            phase_a(f[0], cur, 0);
            if (RRB_HBV_SEASONS) snow_free = !any_snow();
            t_cur = t0;
            if (cur.need[0]) phase_b(ic<1>{}, cur, 0);
            else phase_b(ic<0>{}, cur, 0);
        }
    });
}

// shared-memory address of the staged pow tables + the polynomial coefficients read back into registers
__device__ __forceinline__ HbvPowK hbv_pow_coefficients(uint32_t tb) {
    const uint32_t pa = tb + (uint32_t)offsetof(HbvTables, poly);
    return HbvPowK{lds_f64_at(pa), lds_f64_at(pa + 8), lds_f64_at(pa + 16), lds_f64_at(pa + 24),
                   lds_f64_at(pa + 32), lds_f64_at(pa + 40), lds_f64_at(pa + 48)};
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
// OBJ: 0 = no fused objective, 1 = MSE / NSE (one sum), 2 = KGE (three sums)
// ------------------------------------------------------------------------------------------------
// TT: timesteps per forcing tile (kHbvTT for large CTAs: half the CTA-wide barriers, 1-2 % faster; kHbvTTSmall otherwise)
template <int MPT, bool WRITEQ, bool STORAGE, int OBJ, int ABL, int TT>
__device__ __forceinline__ void hbv_fast2_body(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
                                               const double* __restrict__ params, int64_t N, HbvOut out, Slab slab, Objective obj,
                                               Batch batch, uint32_t* __restrict__ fflag) {
    if (*fflag != 0u) return;  // non-finite forcing: the PRECISE kernel behind takes the whole launch
    HBV_BATCH_PROLOGUE
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nthreads = (N + MPT - 1) / MPT;   // MPT = 2 is launched for even N only
    // threads past the end of the ensemble recompute the last member(s) and store the same values again
    const int64_t i0 = MPT * (gi < nthreads ? gi : nthreads - 1);
    HbvPar<MPT> P;
    bool sane = P.load(params, i0);
    sane = sane && fabs(snow0) <= 1e300 && fabs(soil0) <= 1e300 && fabs(s10) <= 1e300 && fabs(s20) <= 1e300;
    uint32_t* my_flag = hbv_cta_flags(fflag) + (blockIdx.y * gridDim.x + blockIdx.x);
    if (!__syncthreads_and(sane)) {  // CTA-uniform: a member outside the contract
        *my_flag = 1u;               // (every thread stores the same word: no divergent region in front of the warp votes)
        return;
    }

    HbvSt<MPT, OBJ> S;
#pragma unroll
    for (int m = 0; m < MPT; ++m) {  // hbvedu_model.py:78-81
        S.snow[m] = snow0; S.soil[m] = soil0; S.s1[m] = s10; S.s2[m] = s20; S.worst[m] = 0u; S.acc[m].reset();
    }
    int64_t t_first = slab.t_begin;
    int64_t off = i0 + (slab.t_begin - slab.row0) * N;  // row r of the buffers = timestep row0 + r
    if (slab_loads_state(slab)) {
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            S.snow[m] = slab.state[0 * N + i0 + m];
            S.soil[m] = slab.state[1 * N + i0 + m];
            S.s1[m] = slab.state[2 * N + i0 + m];
            S.s2[m] = slab.state[3 * N + i0 + m];
            if (OBJ && slab.t_begin > 0) S.acc[m].load(slab.state, 4, N, i0 + m, obj);
        }
    } else {
        // t = 0 is not simulated (the reference loop starts at 1, hbvedu_model.py:84): qsim[0] = 0, storages = initial states
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            if (WRITEQ) st_stream(out.qsim + off + m, 0.0);
            if (STORAGE) {
                st_stream(out.snow + off + m, S.snow[m]);
                st_stream(out.soil + off + m, S.soil[m]);
                st_stream(out.s1 + off + m, S.s1[m]);
                st_stream(out.s2 + off + m, S.s2[m]);
            }
            if (OBJ) S.acc[m].add(obj.qobs[0], 0.0, obj);
        }
        off += N;
        t_first = 1;
    }
    HbvRows O;
    O.q = reinterpret_cast<char*>(WRITEQ ? out.qsim + off : nullptr);
    O.snow = reinterpret_cast<char*>(STORAGE ? out.snow + off : nullptr);
    O.soil = reinterpret_cast<char*>(STORAGE ? out.soil + off : nullptr);
    O.s1 = reinterpret_cast<char*>(STORAGE ? out.s1 + off : nullptr);
    O.s2 = reinterpret_cast<char*>(STORAGE ? out.s2 + off : nullptr);
    O.row_bytes = (uint32_t)N * 8u;  // N < 2^29 (launch_hbvedu)
    O.row = 0u;                      // rows written since t_first
    pin(O.row_bytes);

    extern __shared__ __align__(128) unsigned char rrb_smem[];
    uint32_t tb = smem_u32(hbv_tables_to_smem(rrb_smem + forcing_smem_bytes<kHbvR, TT>()));
    pin(tb);
    __syncthreads();  // the staged tables are visible
    const HbvPowK pk = hbv_pow_coefficients(tb);

    hbv_fast2_loop<MPT, WRITEQ, STORAGE, OBJ, ABL>(P, S, O, tb, pk, obj, [&](auto&& group) __attribute__((always_inline)) {
        stream_forcing_grouped<kHbvR, TT, kHbvGroup, HbvF>(F, t_first, slab.t_end, group);
    });

    // did any soil moisture leave the range of the table-driven pow?  Then this CTA's results are void: flag it for
    // the PRECISE kernel behind and leave the carry state / objective as they were.
    const bool bad = __syncthreads_or(S.left_range(P));
    if (threadIdx.x == 0) *my_flag = bad ? 1u : 0u;
    if (bad) return;
    if (gi < nthreads) {
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            if (slab.save_state) {
                slab.state[0 * N + i0 + m] = S.snow[m];
                slab.state[1 * N + i0 + m] = S.soil[m];
                slab.state[2 * N + i0 + m] = S.s1[m];
                slab.state[3 * N + i0 + m] = S.s2[m];
                if (OBJ && slab.save_state == 1) S.acc[m].save(slab.state, 4, N, i0 + m, obj);
            }
            if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i0 + m] = S.acc[m].finish(obj);
        }
    }
}

// Two compilations of that body.  hbv_fast2_kernel is held to 128 registers (two 256-thread CTAs or one CTA of up to 16
// warps per SM); hbv_fast2_wide_kernel is not (154 registers for two members per thread): one CTA per SM of at most 12
// warps, where it is 11 % faster than the capped build (profiles/r02_seasons_ab.txt, 104 192 members).
template <int MPT, bool WRITEQ, bool STORAGE, int OBJ, int ABL = 0, int TT = kHbvTTSmall>
__global__ void __launch_bounds__(512, 1)
hbv_fast2_kernel(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
                 const double* __restrict__ params, int64_t N, HbvOut out, Slab slab, Objective obj, Batch batch,
                 uint32_t* __restrict__ fflag) {
    hbv_fast2_body<MPT, WRITEQ, STORAGE, OBJ, ABL, TT>(F, snow0, soil0, s10, s20, params, N, out, slab, obj, batch, fflag);
}
template <bool WRITEQ, int OBJ>
__global__ void hbv_fast2_wide_kernel(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
                                      const double* __restrict__ params, int64_t N, HbvOut out, Slab slab, Objective obj,
                                      Batch batch, uint32_t* __restrict__ fflag) {
    hbv_fast2_body<2, WRITEQ, false, OBJ, 0, kHbvTT>(F, snow0, soil0, s10, s20, params, N, out, slab, obj, batch, fflag);
}

// ------------------------------------------------------------------------------------------------
// FAST, rotating schedule (hbv_rot_kernel): for ensembles that leave the SM sub-partitions UNEVENLY loaded.
//
// A warp lives on the sub-partition (SMSP) it was launched on: warp w of a CTA that has an SM to itself runs on SMSP
// w % 4 (profiles/r02_smsp_layouts.txt).  65 536 members are 1024 member pairs-of-32 ("pairs": one warp, two members per
// thread) on 148 SMs = 6.92 per SM: three SMSPs carry two warps and need 385 cycles per timestep, the fourth carries
// one, needs 285, and idles for the last quarter of the launch.  No static assignment can do better -- the unit of
// work is a warp for the whole series -- so this kernel moves the pairs instead: one persistent CTA per SM owns P
// pairs and P warps, the series is cut into phases, and in phase ph warp w advances pair (w + ph) mod P.  Warps on the
// fullest SMSPs ("slow") advance their pair by a fixed number of steps; warps on a lighter SMSP ("fast") keep going
// until every slow warp of the CTA has finished its share (a shared-memory counter polled once per group of two
// steps), so the phases need no calibration.  Between phases the pair's stores, objective sums and position in time
// pass through shared memory; a CTA-wide barrier ends each phase.  Every warp streams the forcing through its OWN
// three-stage TMA ring (the pairs of a CTA are at different timesteps).  The arithmetic per member is that of
// hbv_fast2_kernel<2> (hbv_fast2_loop), so the results are bit-identical to it.
// Flags: one word per pair (64 members); the PRECISE kernel behind is launched with 64-thread CTAs to match.
// ------------------------------------------------------------------------------------------------
constexpr int kRotTT = 64;       // timesteps per tile of a warp's ring (the packed forcing is padded to kHbvTT = 2 kRotTT)
constexpr int kRotStages = 3;
constexpr uint32_t kRotTileBytes = kRotTT * kHbvR * sizeof(double);
constexpr uint32_t kRotRingBytes = kRotStages * kRotTileBytes + 32;  // + the stage barriers
constexpr int kRotMaxWarps = 16;
constexpr int kRotDone = 0x7fffffff;  // position of a pair that is out of the schedule (flagged for PRECISE)
static_assert(kHbvTT % kRotTT == 0, "the padded series holds whole tiles");

struct RotCfg {
    int rounds;    // full rotations the series is planned for (each pair visits every warp once per rotation)
    int rho_q8;    // first guess of (steps of a fast warp) / (steps of a slow warp) per phase, x 256; measured after that
    int min_steps; // shortest share of a slow warp
};

__device__ __forceinline__ void mbar_init_at(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_at(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// A warp's own forcing ring.  The barrier phases persist from one job to the next (parity bit per stage), so a job
// that stops early only has to wait for the tiles it still has in flight.
struct WarpRing {
    uint32_t tiles, bars, parity;
    int head;  // stage of the next tile to consume
    __device__ __forceinline__ void init(uint32_t base, int lane) {
        tiles = base;
        bars = base + kRotStages * kRotTileBytes;
        parity = 0u;
        head = 0;
        if (lane == 0) {
            for (int s = 0; s < kRotStages; ++s) mbar_init_at(bars + 8u * s, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    __device__ __forceinline__ void issue(const double* __restrict__ F, int64_t k, int stage) const {
        const uint32_t bar = bars + 8u * (uint32_t)stage;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kRotTileBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         tiles + (uint32_t)stage * kRotTileBytes),
                     "l"(F + (size_t)k * kRotTT * kHbvR), "r"(kRotTileBytes), "r"(bar)
                     : "memory");
    }
    // group(ic<2>, t0, f[2]) / group(ic<1>, t, f[1]) for t in [t_begin, t_end); with POLL, stop() (warp-uniform) is
    // asked after every full group and ends the job there.  Returns the first timestep NOT simulated.
    template <bool POLL, class Group, class Stop>
    __device__ __forceinline__ int64_t run(const double* __restrict__ F, int64_t t_begin, int64_t t_end, int lane,
                                           Group&& group, Stop&& stop) {
        constexpr uint32_t kRowBytes = kHbvR * sizeof(double);
        constexpr int G = kHbvGroup;
        if (t_end <= t_begin) return t_begin;
        const int64_t k_begin = t_begin / kRotTT;
        const int64_t k_end = (t_end + kRotTT - 1) / kRotTT;
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async_smem();  // the stages were last read through the generic proxy (previous job)
            for (int s = 0; s < kRotStages; ++s)
                if (k_begin + s < k_end) issue(F, k_begin + s, (head + s) % kRotStages);
        }
        int64_t t = t_begin;
        bool stopped = false;
        int64_t k = k_begin;
        for (; k < k_end; ++k) {
            mbar_wait_at(bars + 8u * (uint32_t)head, (parity >> head) & 1u);
            parity ^= 1u << head;
            const int64_t t0 = k * kRotTT;
            const int lo = (int)((t_begin > t0) ? (t_begin - t0) : 0);
            const int hi = (int)((t_end < t0 + kRotTT) ? (t_end - t0) : kRotTT);
            uint32_t addr = tiles + (uint32_t)head * kRotTileBytes + (uint32_t)lo * kRowBytes;
            int left = hi - lo;
            while (left > 0 && (t % G) != 0) {  // ragged head
                HbvF f1[1] = {HbvF::load(addr)};
                group(ic<1>{}, t, f1);
                addr += kRowBytes; ++t; --left;
            }
#pragma unroll 1
            for (int j = left / G; j > 0; --j) {
                HbvF f[G];
#pragma unroll
                for (int g = 0; g < G; ++g) f[g] = HbvF::load(addr + (uint32_t)g * kRowBytes);
                group(ic<G>{}, t, f);
                addr += G * kRowBytes;
                t += G;
                left -= G;
                if (POLL && stop()) { stopped = true; break; }
            }
            if (!stopped) {
                for (; left > 0; --left) {  // ragged tail
                    HbvF f1[1] = {HbvF::load(addr)};
                    group(ic<1>{}, t, f1);
                    addr += kRowBytes; ++t;
                }
            }
            __syncwarp();  // every lane is done reading this stage
            if (stopped) break;
            if (lane == 0 && k + kRotStages < k_end) {
                fence_proxy_async_smem();
                issue(F, k + kRotStages, head);
            }
            head = (head + 1 == kRotStages) ? 0 : head + 1;
        }
        if (stopped) {  // tiles k+1 .. are still in flight into the following stages: let them land
            head = (head + 1 == kRotStages) ? 0 : head + 1;
            const int64_t last = (k + kRotStages < k_end) ? k + kRotStages : k_end;
            for (int64_t kk = k + 1; kk < last; ++kk) {
                mbar_wait_at(bars + 8u * (uint32_t)head, (parity >> head) & 1u);
                parity ^= 1u << head;
                head = (head + 1 == kRotStages) ? 0 : head + 1;
            }
        }
        return t;
    }
};

// shared memory of hbv_rot_kernel: [W rings | pow tables | control block | pair states]
struct RotCtl {
    int pos[2][kRotMaxWarps];  // position of each pair (first timestep not yet simulated; kRotDone = flagged), double-buffered by phase
    int rho_q8;                // measured fast/slow ratio of a phase, x 256 (0 = none yet)
    uint32_t done;             // slow warps that finished their share, cumulative over the phases
    int work_slow;             // plan of the current phase (thread 0): slow warps that have work (-1: every pair is through),
    int share;                 // and the steps each of them advances its pair by
};
template <int OBJ>
__host__ __device__ constexpr int rot_state_slots() { return 4 + (OBJ == 1 ? 1 : (OBJ == 2 ? 3 : 0)); }
__host__ __device__ constexpr size_t rot_smem_bytes(int warps, int slots) {
    return (size_t)warps * kRotRingBytes + ((sizeof(HbvTables) + 15) & ~size_t(15)) + sizeof(RotCtl) +
           (size_t)warps * 64 * slots * sizeof(double);
}

template <int MAXW, bool WRITEQ, bool STORAGE, int OBJ>
__global__ void __launch_bounds__(MAXW * 32, 1)
hbv_rot_kernel(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
               const double* __restrict__ params, int64_t N, HbvOut out, Slab slab, Objective obj, Batch batch,
               uint32_t* __restrict__ fflag, RotCfg rc) {
    static_assert(!STORAGE, "the storage outputs are HBM-bound: hbv_fast2_kernel serves them");
    if (*fflag != 0u) return;  // non-finite forcing: the PRECISE kernel behind takes the whole launch
    HBV_BATCH_PROLOGUE
    constexpr int NS = rot_state_slots<OBJ>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t nthreads = N / 2;            // launched for even N only
    const int64_t npairs = (nthreads + 31) / 32;
    const int64_t per = npairs / gridDim.x, extra = npairs % gridDim.x;
    const int P = (int)(per + ((int64_t)blockIdx.x < extra ? 1 : 0));  // pairs of this CTA (<= W, launch_hbvedu)
    const int64_t pair0 = (int64_t)blockIdx.x * per + ((int64_t)blockIdx.x < extra ? (int64_t)blockIdx.x : extra);
    // thread `lane` of the warp that holds pair p owns members i0, i0 + 1 (threads past the end recompute the last two)
    auto first_member = [&](int p) -> int64_t {
        const int64_t j = (pair0 + p) * 32 + lane;
        return 2 * (j < nthreads ? j : nthreads - 1);
    };

    extern __shared__ __align__(128) unsigned char rrb_smem[];
    unsigned char* sm_tables = rrb_smem + (size_t)W * kRotRingBytes;
    RotCtl* ctl = reinterpret_cast<RotCtl*>(sm_tables + ((sizeof(HbvTables) + 15) & ~size_t(15)));
    double* xch = reinterpret_cast<double*>(ctl + 1);  // [W][NS][64]
    WarpRing ring;
    ring.init(smem_u32(rrb_smem + (size_t)warp * kRotRingBytes), lane);
    uint32_t tb = smem_u32(hbv_tables_to_smem(sm_tables));
    pin(tb);
    const uint32_t done_addr = smem_u32(&ctl->done);
    if (threadIdx.x == 0) { ctl->done = 0u; ctl->rho_q8 = 0; ctl->share = 0; ctl->work_slow = 0; }

    // ---- prologue: warp p prepares pair p (contract check, initial or carried state, row of t = 0)
    const int64_t t_end = slab.t_end;
    if (warp < P) {
        const int64_t i0 = first_member(warp);
        HbvPar<2> Pm;
        bool sane = Pm.load(params, i0);
        sane = sane && fabs(snow0) <= 1e300 && fabs(soil0) <= 1e300 && fabs(s10) <= 1e300 && fabs(s20) <= 1e300;
        sane = __all_sync(0xffffffffu, sane);
        HbvSt<2, OBJ> S;
#pragma unroll
        for (int m = 0; m < 2; ++m) { S.snow[m] = snow0; S.soil[m] = soil0; S.s1[m] = s10; S.s2[m] = s20; S.acc[m].reset(); }
        int64_t t_first = slab.t_begin;
        if (slab_loads_state(slab)) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                S.snow[m] = slab.state[0 * N + i0 + m];
                S.soil[m] = slab.state[1 * N + i0 + m];
                S.s1[m] = slab.state[2 * N + i0 + m];
                S.s2[m] = slab.state[3 * N + i0 + m];
                if (OBJ && slab.t_begin > 0) S.acc[m].load(slab.state, 4, N, i0 + m, obj);
            }
        } else {
            // t = 0 is not simulated (hbvedu_model.py:84): qsim[0] = 0
            if (WRITEQ) st_stream_pair(out.qsim + i0 + (slab.t_begin - slab.row0) * N, 0.0, 0.0);
#pragma unroll
            for (int m = 0; m < 2; ++m)
                if (OBJ) S.acc[m].add(obj.qobs[0], 0.0, obj);
            t_first = 1;
        }
        double* x = xch + (size_t)warp * NS * 64 + 2 * lane;
        reinterpret_cast<double2*>(x)[0] = make_double2(S.snow[0], S.snow[1]);
        reinterpret_cast<double2*>(x + 64)[0] = make_double2(S.soil[0], S.soil[1]);
        reinterpret_cast<double2*>(x + 128)[0] = make_double2(S.s1[0], S.s1[1]);
        reinterpret_cast<double2*>(x + 192)[0] = make_double2(S.s2[0], S.s2[1]);
        if constexpr (OBJ == 1) reinterpret_cast<double2*>(x + 256)[0] = make_double2(S.acc[0].sse, S.acc[1].sse);
        if constexpr (OBJ == 2) {
            reinterpret_cast<double2*>(x + 256)[0] = make_double2(S.acc[0].se, S.acc[1].se);
            reinterpret_cast<double2*>(x + 320)[0] = make_double2(S.acc[0].see, S.acc[1].see);
            reinterpret_cast<double2*>(x + 384)[0] = make_double2(S.acc[0].seo, S.acc[1].seo);
        }
        if (lane == 0) ctl->pos[0][warp] = sane ? (int)t_first : kRotDone;
    }
    __syncthreads();  // tables, control block and pair states are visible
    const HbvPowK pk = hbv_pow_coefficients(tb);

    // ---- the schedule.  SMSP of warp w = w % 4; its load = warps of the CTA on that SMSP.
    const int max_load = (P + 3) / 4;
    auto is_slow = [&](int w) { return (P - (w & 3) + 3) / 4 == max_load; };
    int n_slow = 0;
    for (int w = 0; w < P; ++w) n_slow += is_slow(w) ? 1 : 0;
    const int n_fast = P - n_slow;
    const bool me_slow = is_slow(warp);
    uint32_t target = 0u;  // value of ctl->done once every slow warp with work has finished the current phase
    int ph = 0;
    for (;; ++ph) {
        const int* pos = ctl->pos[ph & 1];
        int* pos_next = ctl->pos[(ph + 1) & 1];
        if (threadIdx.x == 0) {  // plan the phase
            int behind = (int)t_end, ws = 0;
            for (int w = 0; w < P; ++w) {
                int q = w + ph % P;
                q = q >= P ? q - P : q;
                const int tq = pos[q];
                if (tq < (int)t_end) {
                    behind = tq < behind ? tq : behind;
                    ws += is_slow(w) ? 1 : 0;
                }
            }
            if (behind >= (int)t_end) ws = -1;  // every pair is through (or flagged)
            else if (ph % P == 0) {  // a new rotation: the share follows from what is left and the measured ratio
                const int left_rounds = rc.rounds - ph / P > 1 ? rc.rounds - ph / P : 1;
                const int rho = ctl->rho_q8 > 0 ? ctl->rho_q8 : rc.rho_q8;
                const int64_t denom = (int64_t)left_rounds * ((int64_t)n_slow * 256 + (int64_t)n_fast * rho);
                int64_t sh = (((int64_t)t_end - behind) * 256 + denom - 1) / denom;
                sh = sh < rc.min_steps ? rc.min_steps : sh;
                ctl->share = (int)((sh + 1) & ~int64_t(1));
            }
            ctl->work_slow = ws;
        }
        __syncthreads();
        const int work_slow = ctl->work_slow, share = ctl->share;
        if (work_slow < 0) break;
        target += (uint32_t)work_slow;
        if (warp < P) {
            int p = warp + ph % P;
            p = p >= P ? p - P : p;
            const int t0 = pos[p];
            int reached = t0;
            if (t0 < (int)t_end) {
                const bool poll = !me_slow && work_slow > 0;
                const int64_t t1 = me_slow ? (t0 + share < (int)t_end ? t0 + share : (int)t_end) : t_end;
                const int64_t i0 = first_member(p);
                HbvPar<2> Pm;
                Pm.load(params, i0);
                HbvSt<2, OBJ> S;
                double* x = xch + (size_t)p * NS * 64 + 2 * lane;
                {
                    const double2 a = reinterpret_cast<const double2*>(x)[0], b = reinterpret_cast<const double2*>(x + 64)[0];
                    const double2 c = reinterpret_cast<const double2*>(x + 128)[0], d = reinterpret_cast<const double2*>(x + 192)[0];
                    S.snow[0] = a.x; S.snow[1] = a.y; S.soil[0] = b.x; S.soil[1] = b.y;
                    S.s1[0] = c.x; S.s1[1] = c.y; S.s2[0] = d.x; S.s2[1] = d.y;
                    S.worst[0] = S.worst[1] = 0u;
                    if constexpr (OBJ == 1) {
                        const double2 e = reinterpret_cast<const double2*>(x + 256)[0];
                        S.acc[0].sse = e.x; S.acc[1].sse = e.y;
                    }
                    if constexpr (OBJ == 2) {
                        const double2 e = reinterpret_cast<const double2*>(x + 256)[0], f = reinterpret_cast<const double2*>(x + 320)[0];
                        const double2 g = reinterpret_cast<const double2*>(x + 384)[0];
                        S.acc[0].se = e.x; S.acc[1].se = e.y; S.acc[0].see = f.x; S.acc[1].see = f.y;
                        S.acc[0].seo = g.x; S.acc[1].seo = g.y;
                    }
                }
                HbvRows O;
                const int64_t off = i0 + ((int64_t)t0 - slab.row0) * N;
                O.q = reinterpret_cast<char*>(WRITEQ ? out.qsim + off : nullptr);
                O.snow = O.soil = O.s1 = O.s2 = nullptr;
                O.row_bytes = (uint32_t)N * 8u;
                O.row = 0u;
                pin(O.row_bytes);
                int64_t t_reached = t0;
                if (poll) {
                    hbv_fast2_loop<2, WRITEQ, false, OBJ, 0>(Pm, S, O, tb, pk, obj, [&](auto&& group) __attribute__((always_inline)) {
                        t_reached = ring.run<true>(F, t0, t1, lane, group, [&]() __attribute__((always_inline)) {
                            return __any_sync(0xffffffffu, (int)(lds_u32_volatile(done_addr) - target) >= 0);
                        });
                    });
                } else {
                    hbv_fast2_loop<2, WRITEQ, false, OBJ, 0>(Pm, S, O, tb, pk, obj, [&](auto&& group) __attribute__((always_inline)) {
                        t_reached = ring.run<false>(F, t0, t1, lane, group, []() { return false; });
                    });
                }
                reinterpret_cast<double2*>(x)[0] = make_double2(S.snow[0], S.snow[1]);
                reinterpret_cast<double2*>(x + 64)[0] = make_double2(S.soil[0], S.soil[1]);
                reinterpret_cast<double2*>(x + 128)[0] = make_double2(S.s1[0], S.s1[1]);
                reinterpret_cast<double2*>(x + 192)[0] = make_double2(S.s2[0], S.s2[1]);
                if constexpr (OBJ == 1) reinterpret_cast<double2*>(x + 256)[0] = make_double2(S.acc[0].sse, S.acc[1].sse);
                if constexpr (OBJ == 2) {
                    reinterpret_cast<double2*>(x + 256)[0] = make_double2(S.acc[0].se, S.acc[1].se);
                    reinterpret_cast<double2*>(x + 320)[0] = make_double2(S.acc[0].see, S.acc[1].see);
                    reinterpret_cast<double2*>(x + 384)[0] = make_double2(S.acc[0].seo, S.acc[1].seo);
                }
                // a soil moisture left the range of the table-driven pow: the pair leaves the schedule
                const bool bad = __any_sync(0xffffffffu, S.left_range(Pm));
                reached = bad ? kRotDone : (int)t_reached;
                if (lane == 0) {
                    if (me_slow) atomicAdd(&ctl->done, 1u);
                    else if (poll && !bad && t_reached < t_end && work_slow == n_slow)
                        atomicExch(&ctl->rho_q8, (int)(((t_reached - t0) * 256) / share));  // any fast warp's ratio will do
                }
            }
            if (lane == 0) pos_next[p] = reached;
        }
        __syncthreads();
    }

    // ---- epilogue: warp p closes pair p (pos[ph & 1] is the buffer the last plan read: nobody writes it any more)
    if (warp < P) {
        const bool bad = ctl->pos[ph & 1][warp] == kRotDone;
        uint32_t* my_flag = hbv_cta_flags(fflag) + (blockIdx.y * npairs + pair0 + warp);
        if (lane == 0) *my_flag = bad ? 1u : 0u;
        const int64_t j = (pair0 + warp) * 32 + lane;
        if (!bad && j < nthreads) {
            const int64_t i0 = 2 * j;
            const double* x = xch + (size_t)warp * NS * 64 + 2 * lane;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (slab.save_state) {
                    slab.state[0 * N + i0 + m] = x[m];
                    slab.state[1 * N + i0 + m] = x[64 + m];
                    slab.state[2 * N + i0 + m] = x[128 + m];
                    slab.state[3 * N + i0 + m] = x[192 + m];
                }
                if constexpr (OBJ == 1) {
                    ObjAccSse a;
                    a.sse = x[256 + m];
                    if (slab.save_state == 1) a.save(slab.state, 4, N, i0 + m, obj);
                    if (obj.mse && slab.t_end >= obj.T) obj.mse[i0 + m] = a.finish(obj);
                }
                if constexpr (OBJ == 2) {
                    ObjAccKge a;
                    a.se = x[256 + m]; a.see = x[320 + m]; a.seo = x[384 + m];
                    if (slab.save_state == 1) a.save(slab.state, 4, N, i0 + m, obj);
                    if (obj.mse && slab.t_end >= obj.T) obj.mse[i0 + m] = a.finish(obj);
                }
            }
        }
    }
}

int state_slots_hbvedu() { return 4 + kObjSlots; }

// Which FAST instantiation runs: hbv_fast2_kernel with one (1) or two (2) members per thread, or hbv_rot_kernel (3, where
// it applies); rrb_opts.variant or the environment variable RRMPG_B200_HBV_VARIANT override the library's choice (A/B timing).
static int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return e ? atoi(e) : fallback;
}
static int hbv_variant() {
    static const int v = env_int("RRMPG_B200_HBV_VARIANT", -1);
    return v;
}
static RotCfg rot_cfg() {
    static const RotCfg c = {env_int("RRMPG_B200_HBV_ROT_ROUNDS", 2), env_int("RRMPG_B200_HBV_ROT_RHO", 346),
                             env_int("RRMPG_B200_HBV_ROT_MIN_STEPS", 64)};
    return c;
}

// hbv_rot_kernel (rrb_opts.variant = 3) applies to a single catchment without storage outputs whose pairs (64 members) fit
// one persistent CTA per SM: P = ceil(pairs / SMs) <= 16 warps.  Measured against hbv_fast2_kernel<2> as one CTA per SM
// (profiles/r02_layout_sweep.txt): +8.6 % at P = 5, +2.2 % at P = 6, nothing at P = 7 (65 536 members: the lone warp of a
// (2,2,2,1) SM is slowed to 342 cycles per step by the shared-memory traffic of the six others, 285 when alone;
// tools/smsp_neighbour_probe.cu), a loss from P = 8 on -- and one member per thread as one CTA per SM beats it everywhere
// (47 360 members: 2.19 ms against 2.67), so the library never picks it by itself.  Kept as the measured answer to
// "migrate the chains between sub-partitions".  Returns the warps per CTA (8 or 16), 0 = does not apply.
static int rot_warps(int64_t N, int64_t steps, int sm_count, bool pair_ok, bool storage, const Batch& batch, bool forced) {
    if (!pair_ok || storage || batch.count != 1 || sm_count <= 0) return 0;
    const int64_t pairs = (N / 2 + 31) / 32;
    const int64_t P = (pairs + sm_count - 1) / sm_count;
    if (P > kRotMaxWarps) return 0;
    (void)steps; (void)forced;
    return P <= 8 ? 8 : 16;
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
    int variant = cfg.variant > 0 ? (cfg.variant & 15) : hbv_variant();
    auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; };
    const bool pair_ok = (N % 2) == 0 && aligned16(qsim) && aligned16(snow) && aligned16(soil) && aligned16(s1) && aligned16(s2) &&
                         (!slab.state || aligned16(slab.state));
    const int64_t steps = slab.t_end - slab.t_begin;
    // (test knob: RRMPG_B200_HBV_ROT_SMS pretends a smaller GPU, so that small ensembles reach every launch shape)
    const int sms = env_int("RRMPG_B200_HBV_ROT_SMS", cfg.sm_count > 0 ? cfg.sm_count : 148);
    auto warps_per_sm = [&](int64_t threads) { return ((threads + 31) / 32 + sms - 1) / sms; };
    // ---- which kernel, how many members per thread, which launch shape (profiles/r02_layout_sweep.txt, r02_seasons_ab*.txt)
    //  * single catchment, up to 16 member-warps per SM (75 776 members on 148 SMs): ONE member per thread -- twice the warps
    //    hide the latency of the soil chain better than two chains in one warp -- launched as ONE CTA per SM: warp w of a CTA
    //    that has its SM to itself runs on sub-partition w % 4, so the warps spread evenly, which the hardware's placement
    //    of many small CTAs does not (65 536 members: 2.62 ms against 2.76 with 128-thread CTAs and 2.84 with two members
    //    per thread);
    //  * up to 16 warps per SM with TWO members per thread (151 552 members): one CTA per SM again, the uncapped build
    //    while its registers admit the CTA (12 warps);
    //  * beyond: two members per thread, 256-thread CTAs placed by the hardware.
    //  rrb_opts.variant: 1 / 2 = members per thread with the default CTA shape, 3 = hbv_rot_kernel, 5 = two members per
    //  thread as one CTA per SM.
    const bool single = batch.count == 1;
    int rotw = 0, sm_block = 0;
    bool wide = false;
    if (fast && variant == 3 && slab.t_end < (int64_t(1) << 30)) rotw = rot_warps(N, steps, sms, pair_ok, st, batch, true);
    if (fast && !rotw && single && (variant == 5 || (variant != 1 && variant != 2 && variant != 3))) {
        const int64_t w1 = warps_per_sm(N), w2 = warps_per_sm(N / 2);
        if (variant != 5 && w1 <= 16) {
            variant = 1;
            if (w1 >= 5 && cfg.block <= 0) sm_block = (int)w1 * 32;  // (a caller's CTA size keeps the member-per-thread choice)
        } else if (cfg.block <= 0 && pair_ok && w2 <= 16 && (variant == 5 || w2 >= 5)) {
            variant = 2;
            sm_block = (int)w2 * 32;
            wide = !st && w2 <= 12;
        }
    }
    if (variant != 1 && variant != 2) variant = RRB_HBV_DEFAULT_VARIANT;
    if (variant == 2 && !pair_ok) variant = 1;
    const int mpt = (fast && variant == 2) ? 2 : 1;
    const int64_t nthreads = (N + mpt - 1) / mpt;
    int block = cfg.block > 0 ? cfg.block : pick_block(nthreads * batch.count, cfg.sm_count, nthreads >= 256 ? 256 : 64);
    dim3 grid((unsigned)((nthreads + block - 1) / block), (unsigned)batch.count);
    if (rotw) {  // one flag word per pair of 32 threads
        block = 64;
        grid = dim3((unsigned)((N / 2 + 31) / 32), 1u);
    } else if (sm_block) {
        block = sm_block;
        grid = dim3((unsigned)((nthreads + block - 1) / block), 1u);
    }
    // PRECISE: one member per thread; behind a rotating / one-CTA-per-SM launch its CTAs are a whole fraction of the
    // members one flag word stands for
    const int pblock = rotw ? 64 : (sm_block ? 32 : block);
    const dim3 grid_p((unsigned)((N + pblock - 1) / pblock), (unsigned)batch.count);
    const size_t smem_ring = forcing_smem_bytes<kHbvR, kHbvTTSmall>();  // PRECISE and the small-CTA FAST launches
    const bool big_tiles = block >= 224;  // FAST: 256-step tiles for the large CTAs
    const size_t smem_ring_fast = big_tiles ? forcing_smem_bytes<kHbvR, kHbvTT>() : smem_ring;
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
    hbv_precise_kernel<Q_, S_, O_><<<grid_p, pblock, smem_ring, cfg.stream>>>(RRB_HBV_ARGS, p_flag, p_packing, (int)grid.x, p_div)
    const uint32_t* p_flag = nullptr;
    int p_packing = 0, p_div = 0;
    if (!fast) {
        RRB_HBV_DISPATCH(RRB_HBV_PRECISE);
        return cudaGetLastError();
    }
    uint32_t* wflag = const_cast<uint32_t*>(fflag);
    if (rotw) {
        const int64_t pairs = (N / 2 + 31) / 32;
        const int64_t P = (pairs + sms - 1) / sms;
        const dim3 rgrid((unsigned)((pairs + P - 1) / P), 1u);
        const RotCfg rc = rot_cfg();
        const bool kge = ob && obj.kind == RRB_OBJ_KGE_;
        const int slots = !ob ? rot_state_slots<0>() : (kge ? rot_state_slots<2>() : rot_state_slots<1>());
        // more than half of the SM's shared memory, so that a CTA has its SM (and the warp -> sub-partition map) to itself
        size_t smem = rot_smem_bytes(rotw, slots);
        if (smem < 120 * 1024) smem = 120 * 1024;
#define RRB_HBV_ROT(W_, Q_, O_)                                                                                            \
    do {                                                                                                                  \
        auto k_ = hbv_rot_kernel<W_, Q_, false, O_>;                                                                      \
        cudaError_t e_ = cudaFuncSetAttribute(k_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
        if (e_ != cudaSuccess) return e_;                                                                                 \
        k_<<<rgrid, W_ * 32, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag, rc);                                                \
    } while (0)
#define RRB_HBV_ROT_W(W_)                                          \
    do {                                                           \
        if (wq && !ob) RRB_HBV_ROT(W_, true, 0);                   \
        else if (wq && !kge) RRB_HBV_ROT(W_, true, 1);             \
        else if (wq) RRB_HBV_ROT(W_, true, 2);                     \
        else if (!kge) RRB_HBV_ROT(W_, false, 1);                  \
        else RRB_HBV_ROT(W_, false, 2);                            \
    } while (0)
        if (rotw == 8) RRB_HBV_ROT_W(8);
        else RRB_HBV_ROT_W(16);
#undef RRB_HBV_ROT_W
#undef RRB_HBV_ROT
        p_div = 1;
    } else {
        const size_t smem = (wide ? forcing_smem_bytes<kHbvR, kHbvTT>() : smem_ring_fast) + hbv_tables_smem_bytes();
        const bool kge = ob && obj.kind == RRB_OBJ_KGE_;  // three running sums instead of one
        bool wide_done = false;
#define RRB_HBV_FAST2(M_, Q_, S_, O_)                                                                                  \
    do {                                                                                                              \
        if (big_tiles) {                                                                                              \
            if (kge) hbv_fast2_kernel<M_, Q_, S_, (O_) ? 2 : 0, 0, kHbvTT><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag); \
            else hbv_fast2_kernel<M_, Q_, S_, (O_) ? 1 : 0, 0, kHbvTT><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag);     \
        } else {                                                                                                      \
            if (kge) hbv_fast2_kernel<M_, Q_, S_, (O_) ? 2 : 0><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag); \
            else hbv_fast2_kernel<M_, Q_, S_, (O_) ? 1 : 0><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag);   \
        }                                                                                                             \
    } while (0)
#define RRB_HBV_FAST2A(Q_, S_, O_) RRB_HBV_FAST2(1, Q_, S_, O_)
#define RRB_HBV_FAST2B(Q_, S_, O_) RRB_HBV_FAST2(2, Q_, S_, O_)
        if (wide) {  // the uncapped build, while its register count admits the CTA
#define RRB_HBV_WIDE(Q_, O_)                                                                      \
    do {                                                                                         \
        auto k_ = hbv_fast2_wide_kernel<Q_, O_>;                                                  \
        if (kernel_max_threads(k_) >= block) {                                                   \
            k_<<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag);                          \
            wide_done = true;                                                                    \
        }                                                                                        \
    } while (0)
            if (wq && !ob) RRB_HBV_WIDE(true, 0);
            else if (wq && !kge) RRB_HBV_WIDE(true, 1);
            else if (wq) RRB_HBV_WIDE(true, 2);
            else if (!kge) RRB_HBV_WIDE(false, 1);
            else RRB_HBV_WIDE(false, 2);
#undef RRB_HBV_WIDE
        }
        if (wide_done) {
        } else
#ifdef RRB_HBV_ABLATIONS
        const int abl = cfg.variant / 16;
        if (abl > 0 && wq && !st && !ob) {
#define RRB_HBV_ABL(M_, A_) hbv_fast2_kernel<M_, true, false, 0, A_><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag)
            if (mpt == 2) { if (abl == 1) RRB_HBV_ABL(2, 1); else if (abl == 2) RRB_HBV_ABL(2, 2); else if (abl == 3) RRB_HBV_ABL(2, 3); else RRB_HBV_ABL(2, 4); }
            else { if (abl == 1) RRB_HBV_ABL(1, 1); else if (abl == 2) RRB_HBV_ABL(1, 2); else if (abl == 3) RRB_HBV_ABL(1, 3); else RRB_HBV_ABL(1, 4); }
#undef RRB_HBV_ABL
        } else
#endif
        if (mpt == 2) RRB_HBV_DISPATCH(RRB_HBV_FAST2B);
        else RRB_HBV_DISPATCH(RRB_HBV_FAST2A);
#undef RRB_HBV_FAST2A
#undef RRB_HBV_FAST2B
#undef RRB_HBV_FAST2
        p_div = sm_block ? mpt * sm_block / 32 : mpt;  // the PRECISE launch honours the per-CTA flags
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
