// rr_cemaneige.cu -- Cemaneige snow routine and the fused Cemaneige+GR4J ensemble kernels.
// Restates run_cemaneige (rrmpg/models/cemaneige_model.py:16-127) and run_cemaneigegr4j
// (rrmpg/models/cemaneigegr4j_model.py:17-64) for N members at once; replaces the member loops
// of Cemaneige.simulate (rrmpg/models/cemaneige.py:227-240) and CemaneigeGR4J.simulate
// (rrmpg/models/cemaneigegr4j.py:249-268).
//
// The reference runs the snow routine over the whole series and then feeds its outflow to
// run_gr4j; outflow[t] only depends on snow states <= t, so here both advance in the same
// timestep and the [T] intermediate never exists.
//
// Member-independent work is hoisted into the packer: snow = prec*frac, rain = prec - snow
// (cemaneige_model.py:76-77) and G_tresh = 0.9*365.25*mean(snow) (:80, sequential sum like
// numba's np.mean).  Packed forcing per timestep, LC = layer capacity (1, 5 or 16):
//   F[t] = { snow[0..LC) | rain[0..LC) | mean_temp[0..LC) | etp | pad }
// Only + - * / and compares are used and the TU is built with -fmad=false, so the snow routine
// is bit-identical to numba in both math modes.  The division G/G_tresh (:110) is skipped when
// it cannot influence the result (pot_melt == 0 and G >= 0: melt is +0 either way).
#include "rr_common.cuh"
#include "rr_gr4j.cuh"
#include "rr_kernels.h"

namespace rrb {

template <int LC>
struct CemaGeom {
    static constexpr int R = (3 * LC + 1 + 1) & ~1;
    static constexpr int TT = kCemaTileDoubles / R;
};

static int layer_class(int L) { return cema_layer_class(L); }

__global__ void cema_pack_kernel(const double* __restrict__ prec, const double* __restrict__ mean_temp,
                                 const double* __restrict__ frac, const double* __restrict__ etp, int64_t T,
                                 int64_t Tpad, int L, int LC, int R, double* __restrict__ F) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= Tpad) return;
    double* row = F + t * R;
    for (int k = 0; k < R; ++k) row[k] = 0.0;
    if (t < T) {
        for (int l = 0; l < L; ++l) {
            const double p = prec[t * L + l];
            const double snow = p * frac[t * L + l];  // cemaneige_model.py:76
            row[l] = snow;
            row[LC + l] = p - snow;                   // :77
            row[2 * LC + l] = mean_temp[t * L + l];
        }
        if (etp) row[3 * LC] = etp[t];
    }
}

// one thread per layer: sequential sum in index order = numba's np.mean (cemaneige_model.py:80)
__global__ void cema_gtresh_kernel(const double* __restrict__ F, int64_t T, int L, int R, double* __restrict__ g_tresh) {
    const int l = threadIdx.x;
    if (l >= L) return;
    double acc = 0.0;
#pragma unroll 8
    for (int64_t t = 0; t < T; ++t) acc += F[t * R + l];
    g_tresh[l] = 0.9 * 365.25 * (acc / (double)T);
}

cudaError_t pack_cemaneige(const double* prec, const double* mean_temp, const double* frac, const double* etp,
                           int64_t T, int L, double* F, double* g_tresh, cudaStream_t s) {
    if (L < 1 || L > kCemaMaxLayers) return cudaErrorInvalidValue;
    const int LC = layer_class(L);
    const int R = cema_R(LC);
    const int64_t Tpad = padded_steps(T, cema_TT(LC));
    cema_pack_kernel<<<(unsigned)((Tpad + 127) / 128), 128, 0, s>>>(prec, mean_temp, frac, etp, T, Tpad, L, LC, R, F);
    cema_gtresh_kernel<<<1, 32, 0, s>>>(F, T, L, R, g_tresh);
    return cudaGetLastError();
}

struct NoGr4j {
    static constexpr int kStateSlots = 0;
};

struct CemaOut {
    double *q, *G, *eTG, *s_store, *r_store;
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

// PLAIN = discharge only (no storages, no fused objective): the output flags are compile-time constants
// EXACT = the run has exactly LC layers (L == LC): the per-layer bound checks fold away
template <int LC, class Gr4j, bool FAST, bool PLAIN, bool EXACT>
__global__ void cema_kernel(const double* __restrict__ F, const double* __restrict__ g_tresh, int L, double g0,
                            double e0, double s_init, double r_init, const double* __restrict__ params,
                            int64_t pstride, int64_t N, CemaOut out, Slab slab, Objective obj) {
    constexpr bool COUPLED = Gr4j::kStateSlots > 0;
    constexpr int R = CemaGeom<LC>::R, TT = CemaGeom<LC>::TT;
    const bool WRITEQ = PLAIN || out.q != nullptr, STORAGE = !PLAIN && out.G != nullptr,
               OBJ = !PLAIN && obj.qobs != nullptr;  // CTA-uniform
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // threads past the end of the ensemble recompute member N-1 and store the same values to the same
    // addresses: no predicate lives in the time loop
    const int64_t i = gi < N ? gi : N - 1;
    // record = (CTG, Kf[, x1, x2, x3, x4]) -- rrmpg/models/cemaneige.py:64-65, cemaneigegr4j.py:67-72
    const double CTG = params[pstride * i + 0], Kf = params[pstride * i + 1];
    double omCTG = 1 - CTG;  // loop invariant of cemaneige_model.py:94
    pin(omCTG);
    if (EXACT) L = LC;
    double G[LC], eTG[LC], gt[LC], inv_gt[LC];
    uint32_t gt_span[LC];
#pragma unroll
    for (int l = 0; l < LC; ++l) {
        G[l] = 0.0;
        eTG[l] = 0.0;
        gt[l] = (l < L) ? g_tresh[l] : 0.0;
        inv_gt[l] = 1.0 / gt[l];
        gt_span[l] = div_invariant_span(gt[l]);
    }
    Gr4j gr;
    if constexpr (COUPLED) gr.init(params + pstride * i + 2, s_init, r_init);
    double acc = 0.0;
    constexpr int kSlots = 2 * LC + Gr4j::kStateSlots;
    if (slab.t_begin > 0) {
#pragma unroll
        for (int l = 0; l < LC; ++l) {
            G[l] = slab.state[(int64_t)l * N + i];
            eTG[l] = slab.state[(int64_t)(LC + l) * N + i];
        }
        if constexpr (COUPLED) gr.load(slab.state + (int64_t)2 * LC * N, N, i);
        if (OBJ) acc = slab.state[(int64_t)kSlots * N + i];
    }
    int64_t stride = N, strideL = (int64_t)L * N;
    pin(stride); pin(strideL);
    const int64_t off = i + (slab.t_begin - slab.row0) * N;
    double* q_o = WRITEQ ? out.q + off : nullptr;
    double* s_o = (COUPLED && STORAGE) ? out.s_store + off : nullptr;
    double* r_o = (COUPLED && STORAGE) ? out.r_store + off : nullptr;
    double* G_o = STORAGE ? out.G + i + (slab.t_begin - slab.row0) * L * N : nullptr;  // [rows, L, N]
    double* E_o = STORAGE ? out.eTG + i + (slab.t_begin - slab.row0) * L * N : nullptr;
    const double layers = (double)L;
    double inv_layers = 1.0 / layers;
    pin(inv_layers);

    extern __shared__ __align__(128) unsigned char rrb_smem[];
    uint32_t tb = 0;
    if (COUPLED && FAST) {
        tb = smem_u32(fastmath_tables_to_smem(rrb_smem + forcing_smem_bytes<R, TT>()));
        pin(tb);
    }

    // one timestep; FIRST = the very first step of the series, where the stores take their initial values
    // instead of being updated (cemaneige_model.py:85-92)
    auto step = [&](auto first_c, int64_t t, const double* f) {
        constexpr bool FIRST = decltype(first_c)::value != 0;
        double lw_sum = 0.0;
#pragma unroll
        for (int l = 0; l < LC; ++l) {
            if (EXACT || l < L) {
                const double snow = f[l], rain = f[LC + l], Tm = f[2 * LC + l];
                double g = FIRST ? g0 : G[l] + snow;                      // :85-88
                double e = FIRST ? e0 : CTG * eTG[l] + omCTG * Tm;        // :91-94
                e = (e > 0) ? 0.0 : e;                                    // :95-96
                // potential melt (:99-106), branch-free
                const double kt = Kf * Tm;
                const double capped = (kt > g) ? g : kt;
                const double pot = (e == 0 && Tm > 0) ? capped : 0.0;
                // snow-covered-area ratio (:109-112).  The division is only evaluated where it can change the
                // result: with pot == 0 and a non-negative pack the product (0.9 ratio + 0.1) * pot is +0 for
                // every ratio in [0, 1] (the sign-bit test over-approximates "G < 0", which is harmless).
                double ratio = 1.0;
                if (g < gt[l] && (pot != 0.0 || __double2hiint(g) < 0))
                    ratio = div_by_invariant(g, gt[l], inv_gt[l], gt_span[l]);
                const double melt = (0.9 * ratio + 0.1) * pot;            // :115
                g = g - melt;                                             // :118
                lw_sum += rain + melt;                                    // :121, :125
                G[l] = g;
                eTG[l] = e;
                if (STORAGE) {
                    st_stream(G_o + (int64_t)l * stride, g);
                    st_stream(E_o + (int64_t)l * stride, e);
                }
            }
        }
        // np.mean over the layers (:124-125); x / 1 == x
        const double liquid = (L == 1) ? lw_sum : div_by_invariant(lw_sum, layers, inv_layers, kDivSpanOk);
        double qv = liquid;
        if constexpr (COUPLED) qv = gr.step(liquid, f[3 * LC], tb);  // cemaneigegr4j_model.py:62
        if (WRITEQ) {
            st_stream(q_o, qv);
            q_o += stride;
        }
        if (STORAGE) {
            G_o += strideL;
            E_o += strideL;
            if constexpr (COUPLED) {
                st_stream(s_o, gr.S);
                st_stream(r_o, gr.R);
                s_o += stride;
                r_o += stride;
            }
        }
        if (OBJ) {
            const double d = obj.qobs[t] - qv;
            acc += d * d;
        }
    };

    int64_t t_first = slab.t_begin;
    if (slab.t_begin == 0 && slab.t_end > 0) {  // t = 0 peeled: its forcing row comes straight from global memory
        double f0[R];
#pragma unroll
        for (int k = 0; k < R; ++k) f0[k] = F[k];
        step(ic<1>{}, 0, f0);
        t_first = 1;
    }
    stream_forcing_grouped<R, TT, 1, CemaF<LC>>(F, t_first, slab.t_end, [&](auto, int64_t t, const CemaF<LC>* fp) {
        step(ic<0>{}, t, fp[0].v);
    });

    if (gi < N) {
        if (slab.save_state) {
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                slab.state[(int64_t)l * N + i] = G[l];
                slab.state[(int64_t)(LC + l) * N + i] = eTG[l];
            }
            if constexpr (COUPLED) gr.save(slab.state + (int64_t)2 * LC * N, N, i);
            if (OBJ) slab.state[(int64_t)kSlots * N + i] = acc;
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i] = acc / (double)obj.T;
    }
}

static int uh_class(double x4_max) {
    if (!(x4_max <= 64.0)) return -1;
    if (x4_max <= 3.0) return 0;
    if (x4_max <= 4.0) return 1;
    if (x4_max <= 10.0) return 2;
    return 3;
}
static int uh_slots(int c) {
    switch (c) {
        case 0: return 2 + 3 + 7;
        case 1: return 2 + 4 + 9;
        case 2: return 2 + 10 + 21;
        default: return Gr4jMemberDyn::kStateSlots;
    }
}

int state_slots_cemaneige(int L) { return 2 * layer_class(L) + 1; }
int state_slots_cemaneigegr4j(int L, double x4_max) { return 2 * layer_class(L) + uh_slots(uh_class(x4_max)) + 1; }

template <int LC, class Gr4j, bool FAST>
static cudaError_t launch_variant(const double* F, const double* g_tresh, int L, double g0, double e0, double s_init,
                                  double r_init, const double* params, int64_t pstride, int64_t N, const CemaOut& out,
                                  const Slab& slab, const Objective& obj, const LaunchCfg& cfg) {
    constexpr int R = CemaGeom<LC>::R, TT = CemaGeom<LC>::TT;
    const int block = cfg.block > 0 ? cfg.block : pick_block(N, cfg.sm_count, 128);
    const unsigned grid = (unsigned)((N + block - 1) / block);
    const size_t smem = forcing_smem_bytes<R, TT>() + ((FAST && Gr4j::kStateSlots > 0) ? fastmath_smem_bytes() : 0);
    const bool plain = out.q && !out.G && !obj.qobs;
#define RRB_CEMA(P_, E_)                                                                                             \
    cema_kernel<LC, Gr4j, FAST, P_, E_><<<grid, block, smem, cfg.stream>>>(F, g_tresh, L, g0, e0, s_init, r_init, params, \
                                                                           pstride, N, out, slab, obj)
    if (plain && L == LC) RRB_CEMA(true, true);
    else if (L == LC) RRB_CEMA(false, true);
    else RRB_CEMA(false, false);
#undef RRB_CEMA
    return cudaGetLastError();
}

cudaError_t launch_cemaneige(const double* F, const double* g_tresh, int64_t T, int L, double g0, double e0,
                             const double* params, int64_t pstride, int64_t N, double* outflow, double* G,
                             double* eTG, const Slab& slab, const Objective& obj, const LaunchCfg& cfg) {
    (void)T;
    if (N <= 0) return cudaSuccess;
    if (L < 1 || L > kCemaMaxLayers) return cudaErrorInvalidValue;
    CemaOut out{outflow, G, eTG, nullptr, nullptr};
#define RRB_GO(LC_) \
    return launch_variant<LC_, NoGr4j, false>(F, g_tresh, L, g0, e0, 0.0, 0.0, params, pstride, N, out, slab, obj, cfg)
    switch (layer_class(L)) {
        case 1: RRB_GO(1);
        case 5: RRB_GO(5);
        default: RRB_GO(16);
    }
#undef RRB_GO
}

template <int LC>
static cudaError_t launch_coupled_lc(const double* F, const double* g_tresh, int L, const double* in4,
                                     const double* params, int64_t N, double x4_max, const CemaOut& out,
                                     const Slab& slab, const Objective& obj, const LaunchCfg& cfg) {
    const bool fast = cfg.math == RRB_MATH_FAST_;
#define RRB_GO(M_, F_) \
    return launch_variant<LC, M_, F_>(F, g_tresh, L, in4[0], in4[1], in4[2], in4[3], params, 6, N, out, slab, obj, cfg)
    switch (uh_class(x4_max)) {
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

cudaError_t launch_cemaneigegr4j(const double* F, const double* g_tresh, int64_t T, int L, const double* inits4,
                                 const double* params, int64_t N, double x4_max, double* qsim, double* G,
                                 double* eTG, double* s_store, double* r_store, const Slab& slab,
                                 const Objective& obj, const LaunchCfg& cfg) {
    (void)T;
    if (N <= 0) return cudaSuccess;
    if (L < 1 || L > kCemaMaxLayers) return cudaErrorInvalidValue;
    CemaOut out{qsim, G, eTG, s_store, r_store};
    switch (layer_class(L)) {
        case 1: return launch_coupled_lc<1>(F, g_tresh, L, inits4, params, N, x4_max, out, slab, obj, cfg);
        case 5: return launch_coupled_lc<5>(F, g_tresh, L, inits4, params, N, x4_max, out, slab, obj, cfg);
        default: return launch_coupled_lc<16>(F, g_tresh, L, inits4, params, N, x4_max, out, slab, obj, cfg);
    }
}

}  // namespace rrb
