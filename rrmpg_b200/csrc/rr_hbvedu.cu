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
#include "rr_objective.cuh"

#include <cstdlib>
#include <type_traits>

#ifndef RRB_HBV_DEFAULT_VARIANT
#define RRB_HBV_DEFAULT_VARIANT 2
#endif
#ifndef RRB_HBV_PIPELINE
#define RRB_HBV_PIPELINE 0
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
constexpr int kHbvGroup = 2;  // timesteps per group of the software pipeline

struct HbvPowK {  // polynomial coefficients held in registers (an FMA takes one constant-bank operand)
    double a1, a2, a3, a4, c1, c2, c3;
};

// ABL: timing ablations for the discharge-only instantiation (development builds only, results are WRONG):
//   1 no output stores, 2 never wet, 3 always wet, 4 table loads replaced by constants
// OBJ: 0 = no fused objective, 1 = MSE / NSE (one sum), 2 = KGE (four sums)
template <int MPT, bool WRITEQ, bool STORAGE, int OBJ, int ABL = 0>
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
        *my_flag = 1u;               // (every thread stores the same word: no divergent region in front of the warp votes)
        return;
    }

    double snow[MPT], soil[MPT], s1[MPT], s2[MPT];
    typename std::conditional<OBJ == 2, ObjAccKge, ObjAccSse>::type acc[MPT];
#pragma unroll
    for (int m = 0; m < MPT; ++m) { snow[m] = snow0; soil[m] = soil0; s1[m] = s10; s2[m] = s20; acc[m].reset(); }  // hbvedu_model.py:78-81
    int64_t t_first = slab.t_begin;
    int64_t off = i0 + (slab.t_begin - slab.row0) * N;  // row r of the buffers = timestep row0 + r
    if (slab_loads_state(slab)) {
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            snow[m] = slab.state[0 * N + i0 + m];
            soil[m] = slab.state[1 * N + i0 + m];
            s1[m] = slab.state[2 * N + i0 + m];
            s2[m] = slab.state[3 * N + i0 + m];
            if (OBJ && slab.t_begin > 0) acc[m].load(slab.state, 4, N, i0 + m, obj);
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
            if (OBJ) acc[m].add(obj.qobs[0], 0.0, obj);
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
    auto put = [&](char* base, uint32_t r, const double* v) __attribute__((always_inline)) {
        double* p = reinterpret_cast<double*>(base + (uint64_t)r * (uint64_t)row_bytes);  // IMAD.WIDE.U32
        if constexpr (ABL == 1) {
            if (v[0] == 1.2345e-300) st_stream(p, v[MPT - 1]);  // keeps the value alive, (almost) never stores
        } else if (MPT == 2) st_stream_pair(p, v[0], v[MPT - 1]);
        else st_stream(p, v[0]);
    };

    extern __shared__ __align__(128) unsigned char rrb_smem[];
    uint32_t tb = smem_u32(hbv_tables_to_smem(rrb_smem + forcing_smem_bytes<kHbvR, kHbvTT>()));
    pin(tb);
    __syncthreads();  // the staged tables are visible
    const uint32_t pa = tb + (uint32_t)offsetof(HbvTables, poly);
    const HbvPowK pk{lds_f64_at(pa), lds_f64_at(pa + 8), lds_f64_at(pa + 16), lds_f64_at(pa + 24),
                     lds_f64_at(pa + 32), lds_f64_at(pa + 40), lds_f64_at(pa + 48)};

    // ---- the time loop: groups of two timesteps, software-pipelined by one group, four straight-line bodies.
    // A(t): snow routine (hbvedu_model.py:87-96) and potential evapotranspiration (:102) -- needs forcing[t], the snow pack
    //       and parameters only: short chains, independent of the soil / response stores.
    // B(t): soil moisture (:99-111), response routine (:114-123), discharge (:125-127) -- the loop-carried chains.
    // A warp issues in order, so a basic block runs at the latency of its longest dependent chain unless it holds enough
    // independent work, and with 2-4 warps per SM sub-partition nothing else hides that latency (profiles/r02_*).  The
    // pow of B(t) is a ~190-cycle chain that is skipped when no member of the warp has liquid water (warp vote); a branch
    // per step would cut the loop body into short blocks, each as slow as its own chain.  So the vote of BOTH steps of a
    // group is taken first (in the A phase, one group ahead) and selects one of four bodies -- (wet|dry, wet|dry) -- each
    // a single basic block holding B of the two steps and A of the next group, which the scheduler interleaves.
    static_assert(kHbvGroup == 2, "the four group bodies are written for two timesteps per group");
    constexpr int GP = 2;
    struct AOut {
        double liquid[GP][MPT], pe[GP][MPT], pew[GP][MPT], snow_g[GP][MPT];
        bool need[GP];
    };
    AOut cur = {};        // A results of the group whose B phase runs next
    int64_t t_cur = 0;    // first timestep of `cur`
    bool pending = false; // `cur` holds A results whose B phase has not run yet

    auto phase_a = [&](const HbvF& f, AOut& o, int g) __attribute__((always_inline)) {
        // melt = min(snow, DD (temp - T_t)) serves both max(0, snow - m) = snow - melt and the liquid water prec + melt;
        // the cold branch (snow + prec, no liquid water) is the same two additions with -prec in place of melt:
        // snow - (-prec) and prec + (-prec) = +0 for finite precipitation
        const double nprec = __hiloint2double(__double2hiint(f.prec) ^ (int)0x80000000, __double2loint(f.prec));
        bool wet = false;
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            const double dtt = f.temp - Tt[m];
            const double mm = DD[m] * dtt;
            const double melt = (mm < snow[m]) ? mm : snow[m];
            const bool cold = __double2hiint(dtt) < 0;  // temp < T_t through the sign of the (finite) difference
            const double sel = cold ? nprec : melt;
            snow[m] = snow[m] - sel;
            o.liquid[g][m] = f.prec + sel;
            o.snow_g[g][m] = snow[m];
            o.pe[g][m] = fma(C[m], f.dT, f.PEm);  // FAST packing: dT holds dT * PEm
            o.pew[g][m] = o.pe[g][m] * inv_PWP[m];
            wet = wet || ((__double2hiint(o.liquid[g][m]) | __double2loint(o.liquid[g][m])) != 0);  // -0 and NaN count as water
        }
        // prec_eff = liquid * (soil/FC)^Beta (:99) is +0 whenever liquid == 0 and the power is finite, so a warp
        // evaluates the pow only if one of its members has liquid water
        o.need[g] = __any_sync(0xffffffffu, wet);
        if constexpr (ABL == 2) o.need[g] = false;
        if constexpr (ABL == 3) o.need[g] = true;
    };
    // B of step g of group `a`; WET = the warp evaluates the pow for this step
    auto phase_b = [&](auto wet_c, const AOut& a, int g) __attribute__((always_inline)) {
        constexpr bool WET = decltype(wet_c)::value != 0;
        double qv[MPT], sp[MPT], oK[MPT], s1_new[MPT], s2_new[MPT];
        uint32_t hs[MPT];
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            hs[m] = (uint32_t)__double2hiint(soil[m]);
            worst[m] = max(worst[m], hs[m] - safe_lo[m]);  // sticky range check, judged after the time loop
            const double ea = (soil[m] > PWP[m]) ? a.pe[g][m] : a.pew[g][m] * soil[m];  // :105-108
            sp[m] = (soil[m] + a.liquid[g][m]) - ea;                                     // :111 without prec_eff
            oK[m] = max0_sane(s1[m] - Lq[m]) * K_0[m];
            s2_new[m] = fma(s1[m], K_p[m], s2[m] * c2[m]);                               // :121-123
            s1_new[m] = fma(s1[m], c1[m], -oK[m]);                                       // :114-118 without prec_eff
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
                mant[m] = __hiloint2double((int)((hs[m] & 0x000FFFFFu) | 0x3FF00000u), __double2loint(soil[m]));
                kml[m] = (double)((int)(hs[m] >> 20) - 1023) - log2FC[m];
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
                kd[m] = fma(Beta[m], Lg[m], kShift);
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
                rr[m] = fma(Beta[m], Lg[m], -kdm);
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
                soil[m] = fma(-scale[m], w[m], sp[m]);
                s1_new[m] += scale[m] * w[m];
            }
        } else {
#pragma unroll
            for (int m = 0; m < MPT; ++m) soil[m] = sp[m];
        }
#pragma unroll
        for (int m = 0; m < MPT; ++m) {
            s1[m] = s1_new[m];
            s2[m] = s2_new[m];
            qv[m] = fma(s2_new[m], K_2[m], fma(s1_new[m], K_1[m], oK[m]));               // :125-127
            if (OBJ) acc[m].add(obj.qobs[t_cur + g], qv[m], obj);
        }
        if (WRITEQ) put(q_o, row, qv);
        if (STORAGE) {
            put(snow_o, row, a.snow_g[g]);
            put(soil_o, row, soil);
            put(s1_o, row, s1);
            put(s2_o, row, s2);
        }
        ++row;
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
#if RRB_HBV_PIPELINE
    // software-pipelined form: A of the next group inside every body (more overlap, more instructions: measured equal)
    auto body = [&](auto w0, auto w1, auto next_c, const HbvF* fn) __attribute__((always_inline)) {
        constexpr bool NEXT = decltype(next_c)::value != 0;
        AOut nx;
        if constexpr (NEXT) {
            phase_a(fn[0], nx, 0);
            phase_a(fn[1], nx, 1);
        }
        phase_b(w0, cur, 0);
        phase_b(w1, cur, 1);
        if constexpr (NEXT) cur = nx;
    };
    auto run_group = [&](auto next_c, const HbvF* fn) __attribute__((always_inline)) {
        if (cur.need[0]) {
            if (cur.need[1]) body(ic<1>{}, ic<1>{}, next_c, fn);
            else body(ic<1>{}, ic<0>{}, next_c, fn);
        } else {
            if (cur.need[1]) body(ic<0>{}, ic<1>{}, next_c, fn);
            else body(ic<0>{}, ic<0>{}, next_c, fn);
        }
    };
#endif

    stream_forcing_grouped<kHbvR, kHbvTT, kHbvGroup, HbvF>(
        F, t_first, slab.t_end, [&](auto gc, int64_t t0, const HbvF* f) __attribute__((always_inline)) {
            constexpr int G = decltype(gc)::value;
            if constexpr (G == GP) {
#if RRB_HBV_PIPELINE
                if (pending) {
                    run_group(ic<1>{}, f);
                } else {
                    phase_a(f[0], cur, 0);
                    phase_a(f[1], cur, 1);
                    pending = true;
                }
                t_cur = t0;
#else
                phase_a(f[0], cur, 0);   // A of both steps: four independent short chains per member pair
                phase_a(f[1], cur, 1);
                t_cur = t0;
                run_group_b();
#endif
            } else {  // a ragged step at the edge of a time slab
#if RRB_HBV_PIPELINE
                if (pending) {
                    run_group(ic<0>{}, f);
                    pending = false;
                }
#endif
                phase_a(f[0], cur, 0);
                t_cur = t0;
                if (cur.need[0]) phase_b(ic<1>{}, cur, 0);
                else phase_b(ic<0>{}, cur, 0);
            }
        });
#if RRB_HBV_PIPELINE
    if (pending) run_group(ic<0>{}, nullptr);
#else
    (void)pending;
    (void)run_group_b;
#endif

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
                if (OBJ && slab.save_state == 1) acc[m].save(slab.state, 4, N, i0 + m, obj);
            }
            if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i0 + m] = acc[m].finish(obj);
        }
    }
}

int state_slots_hbvedu() { return 4 + kObjSlots; }

// Which FAST instantiation runs: hbv_fast2_kernel with one (1) or two (2) members per thread; rrb_opts.variant or the
// environment variable RRMPG_B200_HBV_VARIANT override the default (A/B timing).
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
    int variant = cfg.variant > 0 ? (cfg.variant & 15) : hbv_variant();
    auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; };
    const bool pair_ok = (N % 2) == 0 && aligned16(qsim) && aligned16(snow) && aligned16(soil) && aligned16(s1) && aligned16(s2) &&
                         (!slab.state || aligned16(slab.state));
    if (variant != 1 && variant != 2) variant = RRB_HBV_DEFAULT_VARIANT;
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
    {
        const size_t smem = smem_ring + hbv_tables_smem_bytes();
        const bool kge = ob && obj.kind == RRB_OBJ_KGE_;  // four running sums instead of one
#define RRB_HBV_FAST2(M_, Q_, S_, O_)                                                                                  \
    do {                                                                                                              \
        if (kge) hbv_fast2_kernel<M_, Q_, S_, (O_) ? 2 : 0><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag);   \
        else hbv_fast2_kernel<M_, Q_, S_, (O_) ? 1 : 0><<<grid, block, smem, cfg.stream>>>(RRB_HBV_ARGS, wflag);       \
    } while (0)
#define RRB_HBV_FAST2A(Q_, S_, O_) RRB_HBV_FAST2(1, Q_, S_, O_)
#define RRB_HBV_FAST2B(Q_, S_, O_) RRB_HBV_FAST2(2, Q_, S_, O_)
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
