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
#include "rr_cemaneige.cuh"

namespace rrb {

static int layer_class(int L) { return cema_layer_class(L); }

__global__ void cema_pack_kernel(const double* __restrict__ prec, const double* __restrict__ mean_temp,
                                 const double* __restrict__ frac, const double* __restrict__ etp, int64_t T,
                                 int64_t Tpad, int L, int LC, int R, double* __restrict__ F,
                                 uint32_t* __restrict__ fflag, int64_t fstride) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= Tpad) return;
    const int64_t c = blockIdx.y;  // catchment
    prec += c * T * L; mean_temp += c * T * L; frac += c * T * L;
    if (etp) etp += c * T;
    F += c * fstride;
    fflag = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(fflag) + c * fstride);
    double* row = F + t * R;
    for (int k = 0; k < R; ++k) row[k] = 0.0;
    if (t < T) {
        bool ok = true;
        for (int l = 0; l < L; ++l) {
            const double p = prec[t * L + l];
            const double snow = p * frac[t * L + l];  // cemaneige_model.py:76
            row[l] = snow;
            row[LC + l] = p - snow;                   // :77
            row[2 * LC + l] = mean_temp[t * L + l];
            ok = ok && forcing_value_sane(snow) && forcing_value_sane(p - snow) && forcing_value_sane(row[2 * LC + l]) &&
                 snow >= 0.0 && p - snow >= 0.0;
        }
        if (etp) {
            row[3 * LC] = etp[t];
            ok = ok && forcing_value_sane(etp[t]);
        }
        if (!ok) atomicOr(fflag, 1u);
    }
}

// G_tresh needs numba's np.mean = a sequential sum in index order (cemaneige_model.py:80), one chain of T
// dependent additions per layer.  The chain itself is short (8 cycles per addition); what would dominate is the
// latency of T strided global loads, so the whole CTA stages tiles of the snow columns in shared memory and
// thread l < L walks its column there.  One CTA per catchment.
constexpr int kGtTile = 256;  // timesteps per staged tile (32 KB of static shared memory at 16 layers)
__global__ void cema_gtresh_kernel(const double* __restrict__ F, int64_t T, int L, int R, double* __restrict__ g_tresh,
                                   int64_t fstride) {
    __shared__ double tile[kGtTile * kCemaMaxLayers];
    F += blockIdx.x * fstride;  // catchment
    g_tresh += blockIdx.x * 2 * kCemaMaxLayers;
    double acc = 0.0;
    for (int64_t t0 = 0; t0 < T; t0 += kGtTile) {
        const int nt = (int)((T - t0 < kGtTile) ? (T - t0) : kGtTile);
        for (int e = threadIdx.x; e < nt * L; e += blockDim.x) {
            const int k = e / L, l = e - k * L;
            tile[e] = F[(t0 + k) * R + l];
        }
        __syncthreads();
        if ((int)threadIdx.x < L) {
            const int l = threadIdx.x;
#pragma unroll 8
            for (int k = 0; k < nt; ++k) acc += tile[k * L + l];
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < L) {
        const int l = threadIdx.x;
        const double mean = acc / (double)T;
        g_tresh[l] = 0.9 * 365.25 * mean;                 // cemaneige_model.py:80
        g_tresh[kCemaMaxLayers + l] = 365.25 * mean;      // Psolannual, cemaneigehyst_model.py:102
    }
}

cudaError_t pack_cemaneige(const double* prec, const double* mean_temp, const double* frac, const double* etp,
                           int64_t T, int L, double* F, double* g_tresh, cudaStream_t s, int count) {
    if (L < 1 || L > kCemaMaxLayers) return cudaErrorInvalidValue;
    const int LC = layer_class(L);
    const int R = cema_R(LC);
    const int64_t Tpad = padded_steps(T, cema_TT(LC));
    const int64_t fstride = forcing_stride_flagged(T, cema_TT(LC), R);
    uint32_t* fflag = forcing_flag(F, T, cema_TT(LC), R);
    cudaError_t e = cudaMemset2DAsync(fflag, sizeof(double) * (size_t)fstride, 0, kForcingFlagBytes, (size_t)count, s);
    if (e != cudaSuccess) return e;
    cema_pack_kernel<<<dim3((unsigned)((Tpad + 127) / 128), (unsigned)count), 128, 0, s>>>(prec, mean_temp, frac, etp, T, Tpad,
                                                                                        L, LC, R, F, fflag, fstride);
    cema_gtresh_kernel<<<(unsigned)count, 256, 0, s>>>(F, T, L, R, g_tresh, fstride);
    return cudaGetLastError();
}

// ---- member-independent layer preprocessing on the device ----
// Restates extrapolate_precipitation (rrmpg/models/cemaneige_utils.py:101-158), extrapolate_temperature (:161-208)
// and calculate_solid_fraction (:16-98) per (timestep, layer).  The per-layer scalars (exp() altitude gradient,
// temperature offset, which solid-fraction rule applies) are computed by the caller exactly as the reference
// does (libm exp), everything per element is + - * / and compares: bit-identical to numba.
__global__ void snow_layers_kernel(const double* __restrict__ prec, const double* __restrict__ mean_temp,
                                   const double* __restrict__ min_temp, const double* __restrict__ max_temp, int64_t T,
                                   int L, SnowLayerScalars k, double* __restrict__ layer_prec,
                                   double* __restrict__ layer_mean, double* __restrict__ frac_solid) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= T * L) return;
    const int64_t t = e / L;
    const int l = (int)(e - t * L);
    const double p = k.scale_prec[l] ? prec[t] * k.prec_factor[l] : prec[t];   // :137-147
    const double d = k.delta_temp[l];
    double mn = min_temp[t], me = mean_temp[t], mx = max_temp[t];
    if (k.shift_temp[l]) {                                                      // :201-207
        mn = mn + d;
        me = me + d;
        mx = mx + d;
    }
    double f;
    if (!k.high[l]) {                                                           // altitude < 1500 m, :58-72
        if (mx <= 0) f = 1;
        else if (mn >= 0) f = 0;
        else f = 1 - (mx / (mx - mn));
    } else {                                                                    // :76-88
        if (me >= 3) f = 0;
        else if (me <= 0) f = 1;
        else f = 1 - (me + 1) / 4;
    }
    layer_prec[e] = p;
    layer_mean[e] = me;
    frac_solid[e] = f;
}

cudaError_t launch_snow_layers(const double* prec, const double* mean_temp, const double* min_temp,
                               const double* max_temp, int64_t T, int L, const SnowLayerScalars& k, double* layer_prec,
                               double* layer_mean, double* frac_solid, cudaStream_t s) {
    if (L < 1 || L > kCemaMaxLayers) return cudaErrorInvalidValue;
    const int64_t n = T * L;
    snow_layers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(prec, mean_temp, min_temp, max_temp, T, L, k,
                                                                  layer_prec, layer_mean, frac_solid);
    return cudaGetLastError();
}

int state_slots_cemaneige(int L) { return 2 * layer_class(L) + kObjSlots; }
int state_slots_cemaneigegr4j(int L, double x4_max) {
    return 2 * layer_class(L) + cema_uh_slots(cema_uh_class(x4_max)) + kObjSlots;
}

cudaError_t launch_cemaneige(const double* F, const double* g_tresh, int64_t T, int L, double g0, double e0,
                             const double* params, int64_t pstride, int64_t N, double* outflow, double* G,
                             double* eTG, const Slab& slab, const Objective& obj, const LaunchCfg& cfg, const Batch& batch) {
    if (N <= 0 || batch.count <= 0) return cudaSuccess;
    if (L < 1 || L > kCemaMaxLayers) return cudaErrorInvalidValue;
    CemaArgs a{F, g_tresh, L, T, g0, e0, 0.0, 0.0, 0.0, params, pstride, N, nullptr,
               forcing_flag(F, T, cema_TT(layer_class(L)), cema_R(layer_class(L))), batch.count, batch.forcing_stride,
               batch.inits, 4};
    CemaOut out{outflow, G, eTG, nullptr, nullptr, nullptr, nullptr, nullptr};
    switch (layer_class(L)) {
        case 1: return cema_launch_variant<1, NoGr4j, false, 0>(a, out, slab, obj, cfg);
        case 5: return cema_launch_variant<5, NoGr4j, false, 0>(a, out, slab, obj, cfg);
        default: return cema_launch_variant<16, NoGr4j, false, 0>(a, out, slab, obj, cfg);
    }
}

cudaError_t launch_cemaneigegr4j(const double* F, const double* g_tresh, int64_t T, int L, const double* inits4,
                                 const double* params, int64_t N, double x4_max, double* qsim, double* G,
                                 double* eTG, double* s_store, double* r_store, const Slab& slab,
                                 const Objective& obj, const LaunchCfg& cfg, const Batch& batch) {
    CemaArgs a{F, g_tresh, L, T, inits4[0], inits4[1], 0.0, inits4[2], inits4[3], params, 6, N, nullptr,
               forcing_flag(F, T, cema_TT(layer_class(L)), cema_R(layer_class(L))), batch.count, batch.forcing_stride,
               batch.inits, 4};
    CemaOut out{qsim, G, eTG, s_store, r_store, nullptr, nullptr, nullptr};
    return cema_launch_coupled<0>(a, x4_max, out, slab, obj, cfg);
}

}  // namespace rrb
