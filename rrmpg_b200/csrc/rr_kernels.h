// rr_kernels.h -- host-side launch interface between the C-ABI layer (rr_api.cu) and the
// per-model kernel translation units.  Internal; the public boundary is include/rrmpg_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rrb {

// One kernel launch simulates timesteps [t_begin, t_end) of every member.  Output buffers
// are C-order [rows, N] (or [rows, L, N]) where row r holds timestep row0 + r, so a launch can
// fill a whole [T, N] array (row0 = 0) or one time slab of a ring (row0 = t_begin).
// state is a [slots, N] carry buffer: read when t_begin > 0, written when save_state != 0.
struct Slab {
    int64_t t_begin;
    int64_t t_end;
    int64_t row0;
    double* state;
    int save_state;  // 0 = no, 1 = all rows incl. the objective sums (internal carry buffer), 2 = model rows only (state_out)
    int resume;  // the series continues an earlier call (rrb_opts.state_in): the states are loaded from `state` although
                 // t_begin == 0, and timestep 0 is an ordinary step (ABC / HBV-Edu do not simulate t = 0 otherwise)
};
// true when the kernel starts from carried states instead of the initial ones
__host__ __device__ inline bool slab_loads_state(const Slab& s) { return s.t_begin > 0 || s.resume != 0; }

// Catchment batching: blockIdx.y selects one of `count` independent catchments that share T and the
// number of members N.  Element strides (in doubles) between consecutive catchments; count = 1 and zero
// strides describe the ordinary single-catchment call.
struct Batch {
    int count;
    int64_t forcing_stride;  // packed forcing: Tpad * R (GR4J / Cemaneige family: + the flag slot, forcing_stride_flagged)
    int64_t out_stride;      // every [T, N] output: T * N
    const double* inits;     // optional device array [count][4] of per-catchment initial states (nullable);
                             // GR4J: (s_init, r_init, -, -), CemaneigeGR4J: (snow_pack, thermal_state, s_init, r_init)
};

struct LaunchCfg {
    cudaStream_t stream;
    int block;  // threads per CTA (0 = auto from N and the SM count)
    int math;   // RRB_MATH_FAST / RRB_MATH_PRECISE
    int sm_count;
    int variant;  // rrb_opts.variant (0 = default)
};

// fused per-member objective (rr_objective.cuh): when qobs != nullptr the kernels accumulate the sums of the chosen
// metric in registers (carried across time slabs through kObjSlots rows of `state`, behind the model's own rows) and
// the launch whose slab ends at T writes mse[i] = the metric of member i.
enum { RRB_OBJ_MSE_ = 0, RRB_OBJ_NSE_ = 1, RRB_OBJ_KGE_ = 2 };
constexpr int kObjSlots = 4;
struct Objective {
    const double* qobs;  // [T] device, nullable
    double* mse;         // [N] device: written by the launch whose slab ends at T
    int64_t T;           // total series length (the divisor of np.mean)
    int kind;            // RRB_OBJ_*_
    double obs_mean;     // np.mean(qobs)  (NSE / KGE)
    double obs_std;      // np.std(qobs), population standard deviation  (NSE / KGE)
    const double* obs_stats;  // catchment batches: device [count][2] = (mean, std) per catchment (nullable)
};

int pick_block(int64_t N, int sm_count, int max_block);
// ONE CTA per SM for a mid-sized single-catchment launch: warp w of a CTA that has an SM to itself runs on sub-partition
// w % 4, so a CTA of ceil(warps / SMs) warps spreads the ensemble evenly over the four sub-partitions of every SM -- the
// hardware's own placement of many small CTAs does not (profiles/r02_layout_sweep.txt, r02_time_blocks.txt: HBV-Edu up to
// +40 %, GR4J +11 %, Cemaneige +4 % at 65 536 members).  Returns the CTA size, or 0 when the layout does not apply: fewer
// than min_warps or more than max_threads / 32 warps per SM (max_threads: what the kernel's register count admits).
inline int one_cta_block(int64_t threads, int sm_count, int max_threads, int min_warps) {
    if (sm_count <= 0) sm_count = 148;
    const int64_t warps = (threads + 31) / 32;
    const int64_t per_sm = (warps + sm_count - 1) / sm_count;
    if (per_sm < min_warps || per_sm * 32 > max_threads || per_sm > 32) return 0;
    return (int)per_sm * 32;
}
template <class Kernel>
inline int kernel_max_threads(Kernel k) {
    cudaFuncAttributes fa{};
    return cudaFuncGetAttributes(&fa, k) == cudaSuccess ? fa.maxThreadsPerBlock : 0;
}

// ---- forcing tile geometry (doubles per timestep R, timesteps per tile TT) ----
constexpr int kAbcR = 1, kAbcTT = 512;
constexpr int kHbvR = 4, kHbvTT = 256;  // padding unit of the packed forcing = the tile of the large-CTA FAST launches;
constexpr int kHbvTTSmall = 128;        // small CTAs (and the PRECISE kernel) stream 128-step tiles: 29 instead of 41 KB of
                                        // shared memory per CTA keeps seven of them resident per SM
constexpr int kGr4jR = 2, kGr4jTT = 256;
constexpr int kCemaTileDoubles = 1024;  // TT = kCemaTileDoubles / R
constexpr int kCemaMaxLayers = 16;
// layer capacity class LC of a run with L elevation layers; R and TT follow from LC
inline int cema_layer_class(int L) { return L <= 1 ? 1 : (L <= 5 ? 5 : 16); }
inline int cema_R(int LC) { return (3 * LC + 1 + 1) & ~1; }
inline int cema_TT(int LC) { return kCemaTileDoubles / cema_R(LC); }

inline int64_t padded_steps(int64_t T, int TT) { return ((T + TT - 1) / TT) * TT; }

// The GR4J / Cemaneige-family packers append one flag word to the packed forcing: non-zero when a forcing value is
// not finite or beyond +-1e6 (then the FAST kernels run the reference-order arithmetic, rr_gr4j.cuh).
constexpr size_t kForcingFlagBytes = 16;
inline size_t forcing_bytes(int64_t T, int TT, int R) {
    return sizeof(double) * (size_t)padded_steps(T, TT) * (size_t)R + kForcingFlagBytes;
}
template <class D>
inline uint32_t* forcing_flag(D* F, int64_t T, int TT, int R) {
    return reinterpret_cast<uint32_t*>(const_cast<double*>(F) + padded_steps(T, TT) * R);
}
// HBV-Edu scratch behind the packed forcing of `count` catchments: the forcing flag slot, then one flag word per CTA of
// the FAST launch ("leave this CTA's members to the PRECISE kernel": rr_hbvedu.cu).  A CTA covers at least 32 members.
inline size_t hbv_scratch_bytes(int64_t N, int64_t count) {
    return kForcingFlagBytes + sizeof(uint32_t) * (size_t)(((N + 31) / 32) * count) + 16;
}
// catchment batches of these models: every catchment's block is followed by its own flag slot
inline int64_t forcing_stride_flagged(int64_t T, int TT, int R) {
    return padded_steps(T, TT) * R + (int64_t)(kForcingFlagBytes / sizeof(double));
}

// ---- forcing packers (device pointers in, packed F[Tpad][R] out) ----
// count catchments: prec [count][T]; F [count][Tpad]
cudaError_t pack_abc(const double* prec, int64_t T, double* F, cudaStream_t s, int count = 1);
// count catchments: inputs are [count][T] (PE_m, T_m: [count][12]), F is [count][Tpad][R] followed by the flag word
// (non-zero: a precipitation value is not finite -- the FAST kernel then leaves the launch to the PRECISE one)
cudaError_t pack_hbvedu(const double* temp, const double* prec, const int8_t* month0, const double* PE_m,
                        const double* T_m, int64_t T, double* F, int math, int count, cudaStream_t s);
// count catchments: prec, etp [count][T]; F [count][forcing_stride_flagged]
cudaError_t pack_gr4j(const double* prec, const double* etp, int64_t T, double* F, cudaStream_t s, int count = 1);
// writes F and g_tresh[L] (sequential np.mean semantics, rrmpg/models/cemaneige_model.py:80)
// count catchments: layer arrays [count][T][L], etp [count][T]; F [count][forcing_stride_flagged], g_tresh [count][32]
cudaError_t pack_cemaneige(const double* prec, const double* mean_temp, const double* frac, const double* etp,
                           int64_t T, int L, double* F, double* g_tresh, cudaStream_t s, int count = 1);

// per-layer scalars of the Cemaneige layer preprocessing (by value: they live in the kernel parameter bank)
struct SnowLayerScalars {
    double prec_factor[kCemaMaxLayers];   // exp((z_l - z_station) * 0.0004), capped at 4000 m
    double delta_temp[kCemaMaxLayers];    // (z_l - z_station) * -0.0065
    unsigned char scale_prec[kCemaMaxLayers];  // 0: layer precipitation = station precipitation, untouched
    unsigned char shift_temp[kCemaMaxLayers];  // 0: layer temperatures = station temperatures, untouched
    unsigned char high[kCemaMaxLayers];        // altitude >= 1500 m: mean-temperature rule of the solid fraction
};
cudaError_t launch_snow_layers(const double* prec, const double* mean_temp, const double* min_temp,
                               const double* max_temp, int64_t T, int L, const SnowLayerScalars& k, double* layer_prec,
                               double* layer_mean, double* frac_solid, cudaStream_t s);

// ---- model launches ----
// batch.inits (nullable): device [count][4], column 0 = initial storage of the catchment
cudaError_t launch_abc(const double* F, int64_t T, double s0, const double* params, int64_t N, double* qsim,
                       double* storage, const Slab& slab, const Objective& obj, const LaunchCfg& cfg,
                       const Batch& batch = Batch{1, 0, 0, nullptr});

cudaError_t launch_hbvedu(const double* F, int64_t T, const double* inits4, const double* params, int64_t N,
                          double* qsim, double* snow, double* soil, double* s1, double* s2, const Slab& slab,
                          const Objective& obj, const LaunchCfg& cfg, const uint32_t* fflag,
                          const Batch& batch = Batch{1, 0, 0, nullptr});

// uh_cap: 0 = derive from x4_max
cudaError_t launch_gr4j(const double* F, int64_t T, double s_init, double r_init, const double* params,
                        int64_t N, double x4_max, double* qsim, double* s_store, double* r_store,
                        const Slab& slab, const Objective& obj, const LaunchCfg& cfg,
                        const Batch& batch = Batch{1, 0, 0, nullptr});

// batch.inits (nullable): device [count][4] = (snow_pack_init, thermal_state_init, -, -)
cudaError_t launch_cemaneige(const double* F, const double* g_tresh, int64_t T, int L, double g0, double e0,
                             const double* params, int64_t pstride, int64_t N, double* outflow, double* G,
                             double* eTG, const Slab& slab, const Objective& obj, const LaunchCfg& cfg,
                             const Batch& batch = Batch{1, 0, 0, nullptr});

cudaError_t launch_cemaneigegr4j(const double* F, const double* g_tresh, int64_t T, int L, const double* inits4,
                                 const double* params, int64_t N, double x4_max, double* qsim, double* G,
                                 double* eTG, double* s_store, double* r_store, const Slab& slab,
                                 const Objective& obj, const LaunchCfg& cfg,
                                 const Batch& batch = Batch{1, 0, 0, nullptr});

// snow-ice family: family bit 0 = hysteresis snow routine, bit 1 = ice melt.  inits5 = (snow_pack_init,
// thermal_state_init, sca_init, s_init, r_init); outputs nullable (all storages of the family or none).
struct SnowIceOut {
    double *qsim, *G, *eTG, *s_store, *r_store, *sca, *icemelt, *snowmelt;
};
// batch: frac_ice [count][L]; batch.inits (nullable): device [count][8] rows = (snow_pack_init, thermal_state_init, s_init,
// r_init, sca_init, -, -, -) -- the only launch whose inits rows are 8 wide (SnowIceBatch::kInitsStride)
constexpr int kSnowIceInitsStride = 8;
cudaError_t launch_snowice(int family, const double* F, const double* g_tresh, const double* frac_ice, int64_t T, int L,
                           const double* inits5, const double* params, int64_t N, double x4_max, const SnowIceOut& o,
                           const Slab& slab, const Objective& obj, const LaunchCfg& cfg,
                           const Batch& batch = Batch{1, 0, 0, nullptr});
int state_slots_snowice(int family, int L, double x4_max);

// number of carry slots a model needs in Slab::state
int state_slots_abc();
int state_slots_hbvedu();
int state_slots_gr4j(double x4_max);
int state_slots_cemaneige(int L);
int state_slots_cemaneigegr4j(int L, double x4_max);

}  // namespace rrb
