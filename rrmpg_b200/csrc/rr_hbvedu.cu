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
                                   Objective obj, Batch batch, const uint32_t* __restrict__ only_if_flag, int fast_packing) {
    // launched behind the FAST kernel: do the work only when that one declined it (see hbv_fast_kernel)
    if (only_if_flag && *only_if_flag == 0u) return;
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
    const double Tt = (T_t == T_t) ? T_t : __longlong_as_double(0x7FF8000000000000LL);
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

int state_slots_hbvedu() { return 5; }

cudaError_t launch_hbvedu(const double* F, int64_t T, const double* inits4, const double* params, int64_t N,
                          double* qsim, double* snow, double* soil, double* s1, double* s2, const Slab& slab,
                          const Objective& obj, const LaunchCfg& cfg, const uint32_t* fflag, const Batch& batch) {
    (void)T;
    if (N <= 0 || batch.count <= 0) return cudaSuccess;
    const int block = cfg.block > 0 ? cfg.block : pick_block(N * batch.count, cfg.sm_count, N >= 256 ? 256 : 64);
    const dim3 grid((unsigned)((N + block - 1) / block), (unsigned)batch.count);
    const bool fast = cfg.math == RRB_MATH_FAST_;
    const size_t smem = forcing_smem_bytes<kHbvR, kHbvTT>() + (fast ? fastmath_smem_bytes() : 0);
    const bool st = snow != nullptr, ob = obj.qobs != nullptr, wq = qsim != nullptr;
    HbvOut out{qsim, snow, soil, s1, s2};
#define RRB_HBV(M_, Q_, S_, O_)                                                                                   \
    M_<Q_, S_, O_><<<grid, block, smem, cfg.stream>>>(F, inits4[0], inits4[1], inits4[2], inits4[3], params, N, out, \
                                                      slab, obj, batch, RRB_HBV_TAIL)
#define RRB_HBV_M(M_)                                      \
    do {                                                   \
        if (wq && st && ob) RRB_HBV(M_, true, true, true);        \
        else if (wq && st) RRB_HBV(M_, true, true, false);        \
        else if (wq && ob) RRB_HBV(M_, true, false, true);        \
        else if (wq) RRB_HBV(M_, true, false, false);             \
        else if (st && ob) RRB_HBV(M_, false, true, true);        \
        else if (st) RRB_HBV(M_, false, true, false);             \
        else RRB_HBV(M_, false, false, true);                     \
    } while (0)
    if (fast) {
#define RRB_HBV_TAIL fflag
        RRB_HBV_M(hbv_fast_kernel);
#undef RRB_HBV_TAIL
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        // the fallback for a flagged forcing block: exits at once otherwise (FAST packing: dT slot = dT * PEm)
#define RRB_HBV_TAIL fflag, 1
        RRB_HBV_M(hbv_precise_kernel);
#undef RRB_HBV_TAIL
    } else {
#define RRB_HBV_TAIL nullptr, 0
        RRB_HBV_M(hbv_precise_kernel);
#undef RRB_HBV_TAIL
    }
#undef RRB_HBV_M
#undef RRB_HBV
    return cudaGetLastError();
}

}  // namespace rrb
