// rr_cemaneige.cuh -- the Cemaneige-family ensemble kernel template, shared by rr_cemaneige.cu (Cemaneige,
// CemaneigeGR4J) and rr_snowice.cu (CemaneigeGR4JIce, CemaneigeHystGR4J, CemaneigeHystGR4JIce).
//
// FAMILY bit 0 = SWE-SCA hysteresis snow routine (run_cemaneigehyst, rrmpg/models/cemaneigehyst_model.py:5-166)
//        bit 1 = degree-day ice melt added to the snow-routine outflow (run_icemelt,
//                rrmpg/models/icemelt_model.py:16-65; couplings cemaneigegr4jice_model.py:76-93,
//                cemaneigehystgr4jice_model.py:89-104)
// Everything outside GR4J is + - * / and compares, compiled without contraction: bit-identical to numba.
#pragma once
#include "rr_common.cuh"
#include "rr_gr4j.cuh"
#include "rr_kernels.h"
#include "rr_objective.cuh"

namespace rrb {

template <int LC>
struct CemaGeom {
    static constexpr int R = (3 * LC + 1 + 1) & ~1;
    static constexpr int TT = kCemaTileDoubles / R;
};

struct NoGr4j {
    static constexpr int kStateSlots = 0;
    static constexpr size_t kOrdSmemBytes = 0;
};

struct CemaOut {
    double *q, *G, *eTG, *s_store, *r_store;
    double *sca, *icemelt, *snowmelt;  // snow-ice family only
};

struct CemaArgs {
    const double* F;
    const double* g_tresh;   // [0, 16): G_tresh per layer; [16, 32): Psolannual per layer (hysteresis)
    int L;
    int64_t T;               // total series length
    double g0, e0, sca0, s_init, r_init;
    const double* params;
    int64_t pstride;
    int64_t N;
    const double* frac_ice;  // device [L], ice models only
    const uint32_t* fflag;   // forcing sanity flag written by the packer (rr_kernels.h)
    // catchment batch (blockIdx.y = catchment): count = 1 and zero strides for the ordinary call
    int count;
    int64_t forcing_stride;  // doubles between the packed forcing blocks (flag slot included)
    const double* inits_c;   // device [count][inits_stride] = (snow_pack_init, thermal_state_init, s_init, r_init[, sca_init]), nullable
    int inits_stride;        // 4, or kSnowIceInitsStride for the snow-ice batches (column 4 = sca_init)
};

template <int LC>
struct CemaF {  // forcing of one timestep: { snow[LC] | rain[LC] | mean_temp[LC] | etp | pad }
    static constexpr int R = CemaGeom<LC>::R;
    double v[R];
    static __device__ __forceinline__ CemaF load(uint32_t addr) {
        CemaF f;
#pragma unroll
        for (int k = 0; k < R; k += 2) {
            const double2 a = lds_f64x2(addr + 8u * k);
            f.v[k] = a.x;
            f.v[k + 1] = a.y;
        }
        return f;
    }
};

// numba's max(a, b) / min(a, b) on floats: (b > a) ? b : a and (b < a) ? b : a (probed against numba)
// (setp + selp like nb_min in rr_common.cuh: the C++ ternary is pattern-matched into a DSETP.MAX + NaN fix-up sequence)
__device__ __forceinline__ double nb_max(double a, double b) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %2, %1;\n\tselp.f64 %0, %2, %1, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}

// PLAIN = discharge only (no storages, no fused objective): the output flags are compile-time constants
// EXACT = the run has exactly LC layers (L == LC): the per-layer bound checks fold away
#ifndef RRB_CEMA_FROST
#define RRB_CEMA_FROST 1  // frost short cut of the snow routine (cema_kernel::snow_part)
#endif
#ifdef RRB_CEMA_MINBLOCKS
#define RRB_CEMA_BOUNDS __launch_bounds__(128, RRB_CEMA_MINBLOCKS)
#else
#define RRB_CEMA_BOUNDS
#endif
template <int LC, class Gr4j, bool FAST, bool PLAIN, bool EXACT, int FAMILY>
__global__ void RRB_CEMA_BOUNDS cema_kernel(CemaArgs a, CemaOut out, Slab slab, Objective obj) {
    constexpr bool COUPLED = Gr4j::kStateSlots > 0;
    constexpr bool HYST = (FAMILY & 1) != 0, ICE = (FAMILY & 2) != 0;
    constexpr int R = CemaGeom<LC>::R, TT = CemaGeom<LC>::TT;
    constexpr int GOFF = HYST ? 4 : 2;  // offset of x1 in the parameter record
    // blockIdx.y = catchment (0 for the ordinary call): every per-catchment pointer below is shifted by cb blocks
    const int64_t cb = (a.count > 1) ? (int64_t)blockIdx.y : 0;
    const bool WRITEQ = PLAIN || out.q != nullptr, STORAGE = !PLAIN && out.G != nullptr,
               OBJ = !PLAIN && obj.qobs != nullptr;  // CTA-uniform
    if (obj.obs_stats) { obj.obs_mean = obj.obs_stats[2 * cb]; obj.obs_std = obj.obs_stats[2 * cb + 1]; }
    const double* __restrict__ F = a.F + cb * a.forcing_stride;
    const uint32_t* __restrict__ fflag =
        reinterpret_cast<const uint32_t*>(reinterpret_cast<const double*>(a.fflag) + cb * a.forcing_stride);
    const double* __restrict__ g_tresh = a.g_tresh + cb * 2 * kCemaMaxLayers;
    double g0 = a.g0, e0 = a.e0, s_init = a.s_init, r_init = a.r_init, sca0 = a.sca0;
    if (a.inits_c) {
        const double* ic_ = a.inits_c + (int64_t)a.inits_stride * cb;
        g0 = ic_[0]; e0 = ic_[1];
        s_init = ic_[2]; r_init = ic_[3];
        if (a.inits_stride > 4) sca0 = ic_[4];
    }
    const int64_t N = a.N;
    int L = a.L;
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // threads past the end of the ensemble recompute member N-1 and store the same values to the same
    // addresses: no predicate lives in the time loop
    const int64_t i = gi < N ? gi : N - 1;
    const double* p = a.params + a.pstride * (cb * N + i);
    // record = (CTG, Kf[, Thacc, Rsp][, x1, x2, x3, x4][, DDF]) -- cemaneige.py:64-65, cemaneigegr4j.py:67-72,
    // cemaneigegr4jice.py:73-79, cemaneigehystgr4j.py:72-79, cemaneigehystgr4jice.py:78-86
    const double CTG = p[0], Kf = p[1];
    const double Thacc = HYST ? p[2] : 1.0, Rsp = HYST ? p[3] : 0.0;
    const double DDF = ICE ? p[GOFF + 4] : 0.0;
    double omCTG = 1 - CTG;  // loop invariant of cemaneige_model.py:94
    pin(omCTG);
    if (EXACT) L = LC;
    double G[LC], eTG[LC];
    double sca_prev[HYST ? LC : 1], swe_max[HYST ? LC : 1], thmelt[HYST ? LC : 1], fice[ICE ? LC : 1];
    // hysteresis contract step: reciprocal of the current ablation threshold min(swe_max, Psolannual Rsp) per layer;
    // 0 = stale (swe_max changed), < 0 = threshold outside [2^-60, 2^60] (IEEE division instead)
    double yth[HYST ? LC : 1];
#pragma unroll
    for (int l = 0; l < LC; ++l) {
        G[l] = 0.0;
        eTG[l] = 0.0;
        if (HYST) {
            // sca[t-1] at t = 0 is sca[T-1], still 0 from np.zeros -- or sca_init itself when T == 1 (:126)
            sca_prev[l] = (a.T == 1) ? sca0 : 0.0;
            swe_max[l] = 0.0;
            yth[l] = 0.0;
            thmelt[l] = ((l < L) ? g_tresh[kCemaMaxLayers + l] : 0.0) * Rsp;  // Psolannual * Rsp, :139
        }
        if (ICE) fice[l] = (l < L) ? a.frac_ice[cb * a.L + l] : 0.0;
    }
    double inv_thacc = 1.0 / Thacc;
    const uint32_t thacc_span = HYST ? div_invariant_span(Thacc) : 0u;
    extern __shared__ __align__(128) unsigned char rrb_smem[];
    // dynamic shared memory = [forcing ring | G_tresh records | FAST tables | unit hydrograph ordinates (long class)]
    constexpr size_t kOrdOff = forcing_smem_bytes<R, TT>() + (HYST ? 0 : 32 * LC) + ((COUPLED && FAST) ? fastmath_smem_bytes() : 0);
    Gr4j gr;
    if constexpr (COUPLED) gr.init(p + GOFF, s_init, r_init, smem_u32(rrb_smem + kOrdOff));
    ObjAcc acc;
    acc.reset();
    constexpr int kLayerSlots = HYST ? 4 : 2;
    constexpr int kSlots = kLayerSlots * LC + Gr4j::kStateSlots;
    if (slab.t_begin > 0) {
#pragma unroll
        for (int l = 0; l < LC; ++l) {
            G[l] = slab.state[(int64_t)l * N + i];
            eTG[l] = slab.state[(int64_t)(LC + l) * N + i];
            if (HYST) {
                sca_prev[l] = slab.state[(int64_t)(2 * LC + l) * N + i];
                swe_max[l] = slab.state[(int64_t)(3 * LC + l) * N + i];
            }
        }
        if constexpr (COUPLED) gr.load(slab.state + (int64_t)kLayerSlots * LC * N, N, i);
        if (OBJ) acc.load(slab.state, kSlots, N, i, obj);
    }
    int64_t stride = N, strideL = (int64_t)L * N;
    pin(stride); pin(strideL);
    const int64_t off = i + (slab.t_begin - slab.row0) * N + cb * a.T * N;
    const int64_t offL = i + (slab.t_begin - slab.row0) * L * N + cb * a.T * N * L;  // [rows, L, N] storages
    double* q_o = WRITEQ ? out.q + off : nullptr;
    double* s_o = (COUPLED && STORAGE) ? out.s_store + off : nullptr;
    double* r_o = (COUPLED && STORAGE) ? out.r_store + off : nullptr;
    double* G_o = STORAGE ? out.G + offL : nullptr;
    double* E_o = STORAGE ? out.eTG + offL : nullptr;
    double* S_o = (HYST && STORAGE) ? out.sca + offL : nullptr;
    double* im_o = (ICE && STORAGE) ? out.icemelt + off : nullptr;
    double* sm_o = (FAMILY == 3 && STORAGE) ? out.snowmelt + off : nullptr;
    const double* __restrict__ qobs_c = OBJ ? obj.qobs + cb * obj.T : nullptr;
    const double layers = (double)L;
    double inv_layers = 1.0 / layers;
    pin(inv_layers);

    // G_tresh is a property of the catchment, not of the member: {G_tresh, 1/G_tresh, division span} per layer
    // live in shared memory and are only read by the steps that really divide by it
    uint32_t gtrec = 0;
    if constexpr (!HYST) {
        double* rec = reinterpret_cast<double*>(rrb_smem + forcing_smem_bytes<R, TT>());
        if ((int)threadIdx.x < LC) {
            const int l = threadIdx.x;
            const double gt = (l < L) ? g_tresh[l] : 0.0;
            rec[4 * l] = gt;
            rec[4 * l + 1] = 1.0 / gt;
            reinterpret_cast<uint32_t*>(rec + 4 * l + 2)[0] = div_invariant_span(gt);
        }
        gtrec = smem_u32(rec);
        pin(gtrec);
        __syncthreads();
    }
    constexpr size_t kGtBytes = HYST ? 0 : 32 * LC;
    // ---- Snow-routine contract (plain Cemaneige routine only).  When, for every member of the CTA, Kf >= 0, the
    // parameters and initial states are finite and moderately ranged, the pack starts non-negative, the packed
    // forcing is finite with non-negative snow and rain (flag of the packer) and every G_tresh is 0 or within
    // [2^-60, 2^60], then the pack stays in [0, 1e13] for the whole series and the time step needs no special
    // cases: G / G_tresh by the Markstein sequence without range checks (for a pack below 2^-960 the quotient
    // is below 2^-900 whatever its low bits, and 0.9 ratio + 0.1 rounds to 0.1 as in the reference), evaluated
    // unconditionally like the reference does.  Bit-identical to the general step; one basic block for all
    // layers.  Any other CTA runs the general step.
    bool snow_ok = false;
    if constexpr (!HYST) {
        snow_ok = Kf >= 0.0 && Kf <= 1e6 && fabs(CTG) <= 1e6 && g0 >= 0.0 && g0 <= 1e6 && fabs(e0) <= 1e6 &&
                  *fflag == 0u;
#pragma unroll
        for (int l = 0; l < LC; ++l) {
            if (l < L) {
                const double gt = lds_f64(gtrec + 32u * l);
                uint32_t span;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(span) : "r"(gtrec + 32u * l + 16u));
                snow_ok = snow_ok && (span != 0u || gt == 0.0) && G[l] >= 0.0 && G[l] <= 1e13 && fabs(eTG[l]) <= 1e6;
            }
        }
    }
    // ---- Hysteresis-routine contract (round 2), same idea: Kf >= 0, Thacc within [2^-60, 2^60], finite non-negative
    // Rsp / initial states, finite forcing with non-negative snow and rain.  Then the pack, the snow-covered area and
    // both quotients of cemaneigehyst_model.py:135,149 are finite and non-negative for the whole series, and the step
    // needs neither the max(sca, 0) nor the min(melt, G) of :154,160 (0.9 sca + 0.1 <= 1 for sca <= 1 and RN is
    // monotonic, so melt <= pot_melt <= G), divides by Thacc with the unchecked Markstein sequence and by the ablation
    // threshold through a per-layer reciprocal that is refreshed only when swe_max changed.  Bit-identical.
    bool hyst_ok = false;
    if constexpr (HYST) {
        hyst_ok = Kf >= 0.0 && Kf <= 1e6 && fabs(CTG) <= 1e6 && thacc_span != 0u && Rsp >= 0.0 && Rsp <= 1e6 &&
                  g0 >= 0.0 && g0 <= 1e6 && fabs(e0) <= 1e6 && sca0 >= 0.0 && sca0 <= 1e6 && *fflag == 0u;
#pragma unroll
        for (int l = 0; l < LC; ++l)
            if (l < L)
                hyst_ok = hyst_ok && G[l] >= 0.0 && G[l] <= 1e13 && fabs(eTG[l]) <= 1e6 && sca_prev[l] >= 0.0 &&
                          sca_prev[l] <= 1e6 && swe_max[l] >= 0.0 && swe_max[l] <= 1e13 && thmelt[l] >= 0.0 && thmelt[l] <= 1e13;
    }
    uint32_t tb = 0;
    bool use_fast = false;
    Exp2Regs ek{};
    if constexpr (COUPLED && FAST) {
        tb = smem_u32(fastmath_tables_to_smem(rrb_smem + forcing_smem_bytes<R, TT>() + kGtBytes));
        pin(tb);
        // FAST path contract of Gr4jMember: finite, moderately ranged parameters, initial states and forcing --
        // then the snow routine feeds GR4J finite water.  CTA-uniform; the barrier also publishes the tables
        // (the peeled first step below already reads them).
        // (hysteresis: Thacc == 0 -- the lower default bound, reachable by the polish step of fit() -- makes
        // snow_balance / Thacc a NaN that the reference's max(0, NaN) = 0 absorbs; step_fast has no special-value
        // handling, so such CTAs take the reference-order step: thacc_span != 0 <=> Thacc in [2^-60, 2^60])
        bool sane = gr.sane && *fflag == 0u && fabs(g0) <= 1e6 && fabs(e0) <= 1e6 && fabs(sca0) <= 1e6 &&
                    (HYST ? thacc_span != 0u : snow_ok);
        for (int k = 0; k < (int)a.pstride; ++k) sane = sane && fabs(p[k]) <= 1e6;
        if (ICE) {
#pragma unroll
            for (int l = 0; l < LC; ++l) sane = sane && fabs(fice[l]) <= 1e6;
        }
        use_fast = __syncthreads_and(sane) != 0;
        if (use_fast) {
            ek = load_exp2_regs(tb);
            gr.enter_fast();
        }
        if constexpr (HYST) hyst_ok = __syncthreads_and(hyst_ok) != 0;
    } else if constexpr (!COUPLED && !HYST) {
        use_fast = __syncthreads_and(snow_ok) != 0;  // standalone Cemaneige: the contract step in both math modes
    }

    // one timestep; FIRST = the very first step of the series, where the stores take their initial values
    // instead of being updated (cemaneige_model.py:85-92, cemaneigehyst_model.py:107-115)
    // The step has two parts: the snow routine (all layers) -> liquid water, and the routing of that water (GR4J,
    // outputs).  The snow routine of step t+1 only needs the snow states of step t, so the coupled kernels run two
    // steps per loop trip -- snow(t), snow(t+1), route(t), route(t+1) in one basic block -- and the scheduler overlaps
    // the snow routine of the next step with the long GR4J chain of the current one.
    struct SnowOut {
        double liquid, ice_sum, snowmelt;
    };
    auto snow_part = [&](auto first_c, auto snow_c, const double* f) -> SnowOut {
        constexpr bool FIRST = decltype(first_c)::value != 0;
        constexpr bool CONTRACT = decltype(snow_c)::value != 0;
        double lw_sum = 0.0, ice_sum = 0.0;
        // Frost in every layer (a property of the forcing, hence of the whole warp): the potential melt (:99-106) is 0 for
        // Tm <= 0 whatever the thermal state, so under the contract melt = (0.9 ratio + 0.1) * 0 = +0 and the pack only
        // grows -- 3 instead of 14 fp64 instructions per layer, the same bits (lw_sum starts at +0: rain + (+0) and rain
        // agree even for a -0 rain; the ice melt max(DDF Tm, 0) is 0 as well).  One AND over the sign bits of the layer
        // temperatures and one vote decide it.
        if constexpr (RRB_CEMA_FROST && !HYST && CONTRACT && !FIRST) {
            int sign = (int)0x80000000;
#pragma unroll
            for (int l = 0; l < LC; ++l)
                if (EXACT || l < L) sign &= __double2hiint(f[2 * LC + l]);
            if (__all_sync(0xffffffffu, sign < 0 && (!ICE || DDF >= 0.0))) {  // (a negative DDF would melt ice in the frost)
#pragma unroll
                for (int l = 0; l < LC; ++l) {
                    if (EXACT || l < L) {
                        const double g = G[l] + f[l];                                      // :88
                        double e = CTG * eTG[l] + omCTG * f[2 * LC + l];                   // :94
                        e = (e > 0) ? 0.0 : e;                                             // :95-96
                        lw_sum += f[LC + l];                                               // :121, melt = +0
                        G[l] = g;
                        eTG[l] = e;
                        if (STORAGE) {
                            st_stream(G_o + (int64_t)l * stride, g);
                            st_stream(E_o + (int64_t)l * stride, e);
                        }
                    }
                }
                const double snowmelt = (L == 1) ? lw_sum : div_by_invariant(lw_sum, layers, inv_layers, kDivSpanOk);
                if (STORAGE) {
                    G_o += strideL;
                    E_o += strideL;
                }
                const double no_ice = 0.0;
                return SnowOut{ICE ? snowmelt + no_ice : snowmelt, no_ice, snowmelt};
            }
        }
        // The hysteresis routine in the frost: potential melt 0, so the snow balance (:131) is the snowfall, the layer
        // accumulates (:133-136, snowfall >= 0 under the contract) and nothing melts (:157-163).
        if constexpr (RRB_CEMA_FROST && HYST && CONTRACT && !FIRST) {
            int sign = (int)0x80000000;
#pragma unroll
            for (int l = 0; l < LC; ++l)
                if (EXACT || l < L) sign &= __double2hiint(f[2 * LC + l]);
            if (__all_sync(0xffffffffu, sign < 0 && (!ICE || DDF >= 0.0))) {
#pragma unroll
                for (int l = 0; l < LC; ++l) {
                    if (EXACT || l < L) {
                        const double snow = f[l];
                        const double g = G[l] + snow;                                      // hyst :110
                        double e = CTG * eTG[l] + omCTG * f[2 * LC + l];                   // :113
                        e = (e > 0) ? 0.0 : e;
                        double q = snow * inv_thacc;                                       // RN(bal / Thacc), bal = snow - 0
                        double r = fma(-Thacc, q, snow);
                        q = fma(r, inv_thacc, q);
                        r = fma(-Thacc, q, snow);
                        q = fma(r, inv_thacc, q);
                        double sca = sca_prev[l] + q;
                        if (g > swe_max[l]) {
                            swe_max[l] = g;
                            yth[l] = 0.0;
                        }
                        sca = nb_min(sca, 1.0);                                            // :154
                        sca_prev[l] = sca;
                        if (g == 0) {                                                      // :166-167
                            swe_max[l] = 0.0;
                            yth[l] = 0.0;
                        }
                        lw_sum += f[LC + l];                                               // rain + melt, melt = +0
                        G[l] = g;
                        eTG[l] = e;
                        if (STORAGE) {
                            st_stream(S_o + (int64_t)l * stride, sca);
                            st_stream(G_o + (int64_t)l * stride, g);
                            st_stream(E_o + (int64_t)l * stride, e);
                        }
                    }
                }
                const double snowmelt = (L == 1) ? lw_sum : div_by_invariant(lw_sum, layers, inv_layers, kDivSpanOk);
                if (STORAGE) {
                    G_o += strideL;
                    E_o += strideL;
                    S_o += strideL;
                }
                const double no_ice = 0.0;
                return SnowOut{ICE ? snowmelt + no_ice : snowmelt, no_ice, snowmelt};
            }
        }
#pragma unroll
        for (int l = 0; l < LC; ++l) {
            if (EXACT || l < L) {
                const double snow = f[l], rain = f[LC + l], Tm = f[2 * LC + l];
                double g = FIRST ? g0 : G[l] + snow;                    // :85-88
                double e = FIRST ? e0 : CTG * eTG[l] + omCTG * Tm;      // :91-94
                e = (e > 0) ? 0.0 : e;                                    // :95-96
                double melt;
                if constexpr (!HYST && CONTRACT) {
                    // contract step (see above): every operation of :99-115, no special cases
                    const double kt = Kf * Tm;
                    const double capped = (kt > g) ? g : kt;
                    const double pot = (e == 0 && Tm > 0) ? capped : 0.0;
                    const double2 gi = lds_f64x2(gtrec + 32u * l);        // {G_tresh, 1 / G_tresh}
                    double q = g * gi.y;
                    double r = fma(-gi.x, q, g);
                    q = fma(r, gi.y, q);
                    r = fma(-gi.x, q, g);
                    q = fma(r, gi.y, q);                                   // RN(G / G_tresh), div_by_invariant
                    const double ratio = (g < gi.x) ? q : 1.0;             // :109-112
                    melt = (0.9 * ratio + 0.1) * pot;                      // :115
                } else if constexpr (!HYST) {
                    // potential melt (:99-106), branch-free
                    const double kt = Kf * Tm;
                    const double capped = (kt > g) ? g : kt;
                    const double pot = (e == 0 && Tm > 0) ? capped : 0.0;
                    // snow-covered-area ratio (:109-112).  The division is only evaluated where it can change
                    // the result: with pot == 0 and a non-negative pack the product (0.9 ratio + 0.1) * pot is
                    // +0 for every ratio in [0, 1] (the sign-bit test over-approximates "G < 0": harmless).
                    double ratio = 1.0;
                    if (pot != 0.0 || __double2hiint(g) < 0) {
                        const double2 gi = lds_f64x2(gtrec + 32u * l);    // {G_tresh, 1 / G_tresh}
                        if (g < gi.x) {
                            uint32_t span;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(span) : "r"(gtrec + 32u * l + 16u));
                            ratio = div_by_invariant(g, gi.x, gi.y, span);
                        }
                    }
                    melt = (0.9 * ratio + 0.1) * pot;                     // :115
                } else if constexpr (CONTRACT) {
                    // hysteresis routine under the contract above (cemaneigehyst_model.py:117-167)
                    const double kt = Kf * Tm;
                    const double capped = (kt > g) ? g : kt;
                    const double pot = (e == 0 && Tm > 0) ? capped : 0.0;
                    const double bal = snow - pot;                        // :131
                    double sca;
                    if (bal >= 0) {                                       // accumulation, :133-136
                        double q = bal * inv_thacc;                        // RN(bal / Thacc), div_by_invariant unchecked
                        double r = fma(-Thacc, q, bal);
                        q = fma(r, inv_thacc, q);
                        r = fma(-Thacc, q, bal);
                        q = fma(r, inv_thacc, q);
                        sca = sca_prev[l] + q;
                        if (g > swe_max[l]) {                             // max(swe_max, G) = (G > swe_max) ? G : swe_max
                            swe_max[l] = g;
                            yth[l] = 0.0;
                        }
                    } else {                                              // ablation, :138-151
                        const double thmax = (swe_max[l] > thmelt[l]) ? thmelt[l] : swe_max[l];
                        double y = yth[l];
                        if (y == 0.0) {  // first ablation step since swe_max changed: one IEEE reciprocal
                            y = (thmax >= 0x1p-60 && thmax <= 0x1p60) ? 1.0 / thmax : -1.0;
                            yth[l] = y;
                        }
                        if (y > 0.0) {                                    // RN(G / Thmax), G = 0 or within [2^-452, 1e13]
                            double q = g * y;
                            double r = fma(-thmax, q, g);
                            q = fma(r, y, q);
                            r = fma(-thmax, q, g);
                            sca = fma(r, y, q);
                        } else {
                            sca = (thmax > 0) ? g / thmax : 0.0;
                        }
                    }
                    sca = nb_min(sca, 1.0);                               // :154, sca >= 0 under the contract
                    melt = (0.9 * sca + 0.1) * pot;                       // :157; <= pot <= G, so :160 is the identity
                    sca_prev[l] = sca;
                    if (STORAGE) st_stream(S_o + (int64_t)l * stride, sca);
                } else {
                    // potential melt (cemaneigehyst_model.py:117-128), branch-free
                    const double kt = Kf * Tm;
                    const double capped = (kt > g) ? g : kt;
                    const double pot = (e == 0 && Tm > 0) ? capped : 0.0;
                    // SWE-SCA hysteresis (cemaneigehyst_model.py:131-160)
                    const double bal = snow - pot;                        // :131
                    double sca;
                    if (bal >= 0) {                                       // accumulation, :133-136
                        sca = sca_prev[l] + div_by_invariant(bal, Thacc, inv_thacc, thacc_span);
                        swe_max[l] = nb_max(swe_max[l], g);
                    } else {                                              // ablation, :138-151
                        const double thmax = (swe_max[l] > thmelt[l]) ? thmelt[l] : swe_max[l];
                        sca = (thmax > 0) ? g / thmax : 0.0;
                    }
                    sca = nb_min(nb_max(sca, 0.0), 1.0);                  // :154
                    melt = (0.9 * sca + 0.1) * pot;                       // :157
                    melt = nb_min(melt, g);                               // :160
                    sca_prev[l] = sca;
                    if (STORAGE) st_stream(S_o + (int64_t)l * stride, sca);
                }
                g = g - melt;                                             // :118 / hyst :163
                if (HYST && g == 0) {                                     // hyst :166-167
                    swe_max[l] = 0.0;
                    yth[l] = 0.0;
                }
                lw_sum += rain + melt;                                    // :121, :125
                if (ICE) {                                                // icemelt_model.py:52-62
                    double im = DDF * Tm;                                 // temp - tbase, tbase = 0
                    im = (im < 0) ? 0.0 : im;
                    im = (g > 1) ? 0.0 : im;
                    ice_sum += im * fice[l];                              // np.sum(icemelt * frac_ice, axis=1)
                }
                G[l] = g;
                eTG[l] = e;
                if (STORAGE) {
                    st_stream(G_o + (int64_t)l * stride, g);
                    st_stream(E_o + (int64_t)l * stride, e);
                }
            }
        }
        // np.mean over the layers (:124-125); x / 1 == x
        const double snowmelt = (L == 1) ? lw_sum : div_by_invariant(lw_sum, layers, inv_layers, kDivSpanOk);
        const double liquid = ICE ? snowmelt + ice_sum : snowmelt;        // cemaneigegr4jice_model.py:87
        if (STORAGE) {
            G_o += strideL;
            E_o += strideL;
            if (HYST) S_o += strideL;
        }
        return SnowOut{liquid, ice_sum, snowmelt};
    };
    auto route_part = [&](auto fast_c, const SnowOut& w, double etp_t, int64_t t) {
        double qv = w.liquid;
        if constexpr (COUPLED) {                                          // cemaneigegr4j_model.py:62
            qv = gr4j_step(fast_c, gr, w.liquid, etp_t, tb, ek);
        }
        if (WRITEQ) {
            st_stream(q_o, qv);
            q_o += stride;
        }
        if (STORAGE) {
            if constexpr (COUPLED) {
                st_stream(s_o, gr.S);
                st_stream(r_o, gr.R);
                s_o += stride;
                r_o += stride;
            }
            if (ICE) {
                st_stream(im_o, w.ice_sum);
                im_o += stride;
            }
            if (FAMILY == 3) {
                st_stream(sm_o, w.snowmelt);
                sm_o += stride;
            }
        }
        if (OBJ) acc.add(qobs_c[t], qv, obj);
    };
    auto step = [&](auto first_c, auto fast_c, auto snow_c, int64_t t, const double* f) {
        const SnowOut w = snow_part(first_c, snow_c, f);
        route_part(fast_c, w, f[3 * LC], t);
    };
#ifndef RRB_CEMA_GROUP
#define RRB_CEMA_GROUP 2
#endif
    constexpr int GROUP = (COUPLED && LC <= 5) ? RRB_CEMA_GROUP : 1;

    auto run = [&](auto fast_c, auto snow_c) {
        int64_t t_first = slab.t_begin;
        if (slab.t_begin == 0 && slab.t_end > 0) {  // t = 0 peeled: its forcing row comes straight from global memory
            double f0[R];
#pragma unroll
            for (int k = 0; k < R; ++k) f0[k] = F[k];
            step(ic<1>{}, fast_c, snow_c, 0, f0);
            t_first = 1;
        }
        stream_forcing_grouped<R, TT, GROUP, CemaF<LC>>(F, t_first, slab.t_end, [&](auto gc, int64_t t, const CemaF<LC>* fp) {
            if constexpr (decltype(gc)::value == 2) {
                const SnowOut w0 = snow_part(ic<0>{}, snow_c, fp[0].v);
                const SnowOut w1 = snow_part(ic<0>{}, snow_c, fp[1].v);
                route_part(fast_c, w0, fp[0].v[3 * LC], t);
                route_part(fast_c, w1, fp[1].v[3 * LC], t + 1);
            } else {
                step(ic<0>{}, fast_c, snow_c, t, fp[0].v);
            }
        });
    };
    if constexpr (COUPLED && FAST && HYST) {
        if (use_fast && hyst_ok) run(ic<1>{}, ic<1>{});
        else if (use_fast) run(ic<1>{}, ic<0>{});
        else run(ic<0>{}, ic<0>{});
    } else if constexpr (COUPLED && FAST) {
        if (use_fast) run(ic<1>{}, ic<1>{});
        else run(ic<0>{}, ic<0>{});
    } else if constexpr (!COUPLED && !HYST) {
        if (use_fast) run(ic<0>{}, ic<1>{});
        else run(ic<0>{}, ic<0>{});
    } else {
        run(ic<0>{}, ic<0>{});
    }

    if (gi < N) {
        if (slab.save_state) {
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                slab.state[(int64_t)l * N + i] = G[l];
                slab.state[(int64_t)(LC + l) * N + i] = eTG[l];
                if (HYST) {
                    slab.state[(int64_t)(2 * LC + l) * N + i] = sca_prev[l];
                    slab.state[(int64_t)(3 * LC + l) * N + i] = swe_max[l];
                }
            }
            if constexpr (COUPLED) gr.save(slab.state + (int64_t)kLayerSlots * LC * N, N, i);
            if (OBJ && slab.save_state == 1) acc.save(slab.state, kSlots, N, i, obj);
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[cb * N + i] = acc.finish(obj);
    }
}

inline int cema_uh_class(double x4_max) {
    if (!(x4_max <= 64.0)) return -1;
    if (x4_max <= 3.0) return 0;
    if (x4_max <= 4.0) return 1;
    if (x4_max <= 10.0) return 2;
    return 3;
}
inline int cema_uh_slots(int c) {
    switch (c) {
        case 0: return 2 + 3 + 7;
        case 1: return 2 + 4 + 9;
        case 2: return 2 + 10 + 21;
        default: return Gr4jMemberDyn::kStateSlots;
    }
}

template <int LC, class Gr4j, bool FAST, int FAMILY>
static cudaError_t cema_launch_variant(const CemaArgs& a, const CemaOut& out, const Slab& slab, const Objective& obj,
                                       const LaunchCfg& cfg) {
    constexpr int R = CemaGeom<LC>::R, TT = CemaGeom<LC>::TT;
    int block = cfg.block > 0 ? cfg.block : pick_block(a.N * a.count, cfg.sm_count, a.N >= 128 ? 128 : 64);
    const bool plain = out.q && !out.G && !obj.qobs;
    if (cfg.block <= 0 && a.count == 1 && Gr4j::kOrdSmemBytes == 0) {  // 9 .. 16 warps per SM: one CTA per SM (rr_kernels.h)
        int cap;
        if (plain && a.L == LC) cap = kernel_max_threads(cema_kernel<LC, Gr4j, FAST, true, true, FAMILY>);
        else if (FAMILY == 0 && a.L == LC) cap = kernel_max_threads(cema_kernel<LC, Gr4j, FAST, false, true, FAMILY>);
        else cap = kernel_max_threads(cema_kernel<LC, Gr4j, FAST, false, false, FAMILY>);
        const int b = one_cta_block(a.N, cfg.sm_count, cap < 512 ? cap : 512, 9);
        if (b) block = b;
    }
    if (Gr4j::kOrdSmemBytes > 0 && block > kOrdThreads) block = kOrdThreads;  // ordinate columns in shared memory
    const dim3 grid((unsigned)((a.N + block - 1) / block), (unsigned)a.count);
    const size_t smem = forcing_smem_bytes<R, TT>() + ((FAMILY & 1) ? 0 : 32 * LC) +
                        ((FAST && Gr4j::kStateSlots > 0) ? fastmath_smem_bytes() : 0) + Gr4j::kOrdSmemBytes;
#define RRB_CEMA(P_, E_) cema_kernel<LC, Gr4j, FAST, P_, E_, FAMILY><<<grid, block, smem, cfg.stream>>>(a, out, slab, obj)
    if (FAMILY == 0) {
        if (plain && a.L == LC) RRB_CEMA(true, true);
        else if (a.L == LC) RRB_CEMA(false, true);
        else RRB_CEMA(false, false);
    } else {
        // the snow-ice family: the discharge-only / exact-layer-count specialisation plus one general variant
        if (plain && a.L == LC) RRB_CEMA(true, true);
        else RRB_CEMA(false, false);
    }
#undef RRB_CEMA
    return cudaGetLastError();
}

template <int LC, int FAMILY>
static cudaError_t cema_launch_coupled_lc(const CemaArgs& a, double x4_max, const CemaOut& out, const Slab& slab,
                                          const Objective& obj, const LaunchCfg& cfg) {
    const bool fast = cfg.math == RRB_MATH_FAST_;
#define RRB_GO(M_, F_) return cema_launch_variant<LC, M_, F_, FAMILY>(a, out, slab, obj, cfg)
    switch (cema_uh_class(x4_max)) {
        case 0:
            if (fast) RRB_GO(Gr4jUh3F, true);
            RRB_GO(Gr4jUh3P, false);
        case 1:
            if (fast) RRB_GO(Gr4jUh4F, true);
            RRB_GO(Gr4jUh4P, false);
        case 2:
            if (fast) RRB_GO(Gr4jUh10F, true);
            RRB_GO(Gr4jUh10P, false);
        case 3:
            RRB_GO(Gr4jMemberDyn, false);
        default:
            return cudaErrorInvalidValue;
    }
#undef RRB_GO
}

template <int FAMILY>
static cudaError_t cema_launch_coupled(const CemaArgs& a, double x4_max, const CemaOut& out, const Slab& slab,
                                       const Objective& obj, const LaunchCfg& cfg) {
    if (a.N <= 0) return cudaSuccess;
    if (a.L < 1 || a.L > kCemaMaxLayers) return cudaErrorInvalidValue;
    switch (cema_layer_class(a.L)) {
        case 1: return cema_launch_coupled_lc<1, FAMILY>(a, x4_max, out, slab, obj, cfg);
        case 5: return cema_launch_coupled_lc<5, FAMILY>(a, x4_max, out, slab, obj, cfg);
        default: return cema_launch_coupled_lc<16, FAMILY>(a, x4_max, out, slab, obj, cfg);
    }
}

}  // namespace rrb
