/*
 * rrmpg_b200.h -- C ABI of the B200 ensemble rainfall-runoff engine (librrmpg_b200.so).
 *
 * Drop-in boundary for the ONE data-parallel hot path of kratzert/RRMPG: the per-timestep
 * storage-update recurrence of ABCModel, HBVEdu, GR4J, Cemaneige and CemaneigeGR4J evaluated
 * for an ensemble of parameter sets.  In the reference this is a Python loop over members
 * around one numba kernel call per member; each entry point below replaces one such loop +
 * kernel pair with a single batched call (citations are file:line under the reference tree):
 *
 *   rrb_abc_simulate            run_abcmodel       rrmpg/models/abcmodel_model.py:16
 *                               member loop        rrmpg/models/abcmodel.py:174-181
 *   rrb_hbvedu_simulate         run_hbvedu         rrmpg/models/hbvedu_model.py:16
 *                               member loop        rrmpg/models/hbvedu.py:199-209
 *   rrb_gr4j_simulate           run_gr4j           rrmpg/models/gr4j_model.py:16
 *                               member loop        rrmpg/models/gr4j.py:169-178
 *   rrb_cemaneige_simulate      run_cemaneige      rrmpg/models/cemaneige_model.py:16
 *                               member loop        rrmpg/models/cemaneige.py:227-240
 *   rrb_cemaneigegr4j_simulate  run_cemaneigegr4j  rrmpg/models/cemaneigegr4j_model.py:17
 *                               member loop        rrmpg/models/cemaneigegr4j.py:249-268
 *   opts->qobs / opts->mse      per-member calc_mse loop of monte_carlo
 *                                                  rrmpg/tools/monte_carlo.py:66-73
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch / numpy types.  All floating point data is IEEE
 *     binary64, arrays are C-order and densely packed.
 *   - `params` is the reference's structured record array viewed as a [N, k] double matrix
 *     (field order = the model's `_dtype`, e.g. rrmpg/models/hbvedu.py:63-66).
 *   - Outputs are laid out like the wrappers' arrays: [T, N] (rrmpg/models/hbvedu.py:191) and
 *     [T, L, N] for the per-layer snow states (rrmpg/models/cemaneige.py:221-224).  Optional
 *     ("storage") outputs may be NULL; pass all of a model's storages or none.  `qsim` itself
 *     may be NULL when only the fused objective (opts->mse) is wanted.
 *   - The caller owns every buffer.  The library never returns memory it allocated, never frees
 *     caller memory and keeps no caller pointer after the call returns (RRB_MEM_HOST) or after
 *     the work enqueued on opts->stream has completed (RRB_MEM_DEVICE).
 *   - Input validation (negative precipitation, month range, ...) stays with the caller, as in
 *     the reference where the Python wrappers raise before the kernel runs.  The kernels have
 *     no numerical error path: NaN / Inf propagate exactly as in the numba code.
 *   - Every function returns RRB_OK (0) or an RRB_E* code; rrb_last_error() gives the message
 *     for the calling thread.  There is no CPU fallback: without a CUDA device the simulate
 *     calls fail with RRB_ECUDA.
 *   - Thread safety: calls are serialised per device by an internal mutex.  The per-device scratch buffers (packed
 *     forcing etc.) are reused from call to call: every call orders its stream behind the previous device-mode call's
 *     last kernel (one event wait), so RRB_MEM_DEVICE calls on different streams of one device run back to back,
 *     not concurrently.
 */
#ifndef RRMPG_B200_H
#define RRMPG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRB_VERSION 110 /* 0.1.1: rrb_opts grew (objective .. out_row_pitch); rrb_init_devices, rrb_state_rows */

#if defined(__GNUC__)
#define RRB_API __attribute__((visibility("default")))
#else
#define RRB_API
#endif

enum rrb_status {
    RRB_OK = 0,
    RRB_EINVAL = 1,       /* bad argument (NULL pointer, T < 1, N < 0, L < 1, ...) */
    RRB_ECUDA = 2,        /* CUDA runtime error (message in rrb_last_error) */
    RRB_EUNSUPPORTED = 3, /* L > RRB_MAX_LAYERS, or a GR4J x4 > RRB_MAX_X4 */
    RRB_ENOMEM = 4        /* device / pinned allocation failed */
};

enum rrb_mem {
    RRB_MEM_HOST = 0,  /* every pointer argument is host memory (pinned memory is fastest) */
    RRB_MEM_DEVICE = 1 /* every pointer argument is device memory on opts->device */
};

enum rrb_math {
    RRB_MATH_FAST = 0,   /* table-driven pow/tanh, hoisted reciprocals; qsim within rtol 1e-10 of numba */
    RRB_MATH_PRECISE = 1 /* the reference's operations one for one (CUDA libm pow/tanh, IEEE division) */
};

enum rrb_objective {     /* what opts->mse receives when opts->qobs is given (rrmpg/utils/metrics.py) */
    RRB_OBJ_MSE = 0,     /* calc_mse :110-136  mean((obs - sim)**2) */
    RRB_OBJ_NSE = 1,     /* calc_nse :29-78    1 - sum((sim - obs)**2) / sum((obs - mean(obs))**2) */
    RRB_OBJ_KGE = 2      /* calc_kge :139-188  1 - sqrt((r-1)**2 + (alpha-1)**2 + (beta-1)**2) */
};

enum rrb_model {         /* for rrb_state_rows */
    RRB_MODEL_ABC = 0,
    RRB_MODEL_HBVEDU = 1,
    RRB_MODEL_GR4J = 2
};

#define RRB_MAX_LAYERS 16 /* Cemaneige elevation layers per call */
#define RRB_MAX_X4 64.0   /* GR4J unit hydrograph time base per member */

typedef struct rrb_opts {
    int32_t struct_size; /* = sizeof(rrb_opts); guards against ABI drift */
    int32_t device;      /* CUDA device ordinal, -1 = the calling thread's current device */
    int32_t mem;         /* enum rrb_mem */
    int32_t math;        /* enum rrb_math */
    void* stream;        /* RRB_MEM_DEVICE: cudaStream_t the work is enqueued on (NULL = default stream);
                            the call returns without synchronising */
    int32_t block;       /* threads per CTA, 0 = chosen from N and the SM count */
    int32_t variant;     /* kernel variant, for A/B timing only: 0 = the library's choice.  HBV-Edu FAST: 1 = one member per
                            thread, 2 = two members per thread (needs an even N and 16-byte aligned rows, else 1), 3 = the
                            rotating schedule (hbv_rot_kernel; at most 16 member-warps per SM, no storage outputs, else 2),
                            5 = two members per thread as ONE CTA per SM (at most 16 member-warps per SM, else 2) */
    double x4_max;       /* RRB_MEM_DEVICE, GR4J family: max x4 over params if the caller knows it;
                            <= 0 lets the library reduce it on the device (one small sync) */
    const double* qobs;  /* optional [T] observed discharge: fuses the objective into the kernel */
    double* mse;         /* [N] out, required when qobs != NULL: mean((qobs - qsim[:, i])**2) */
    int64_t slab_steps;  /* RRB_MEM_HOST: timesteps per pipelined time slab, 0 = automatic */
    /* ---- RRB_VERSION >= 110 ---- */
    int32_t objective;   /* enum rrb_objective: the metric accumulated against qobs (registers, no [T, N] array needed) */
    int32_t n_devices;   /* RRB_MEM_HOST, single-catchment calls: > 1 shards the members over `devices` as contiguous
                            blocks -- one worker thread, context and stream pair per device, the forcing uploaded to each,
                            every block's rows copied straight into its column block of the caller's [T, N] arrays.
                            0 / 1 = opts->device only.  (member loop rrmpg/models/hbvedu.py:199-209) */
    const int32_t* devices;   /* n_devices ordinals, NULL = 0 .. n_devices-1 */
    const double* obs_stats;  /* HOST memory in both modes, required for NSE / KGE: (np.mean(qobs), np.std(qobs)) -- one
                                 pair, or [C][2] for the catchment batches */
    const double* state_in;   /* optional [rrb_state_rows(..)][N]: continue an earlier call -- the stores (and GR4J's UH1 /
                                 UH2 buffers, which the reference neither returns nor accepts: gr4j_model.py:82-83) are
                                 taken from here instead of the initial values, and timestep 0 of this call is an ordinary
                                 step.  ABC, HBV-Edu, GR4J single-catchment calls; memory per opts->mem. */
    double* state_out;        /* optional [rrb_state_rows(..)][N]: the stores after the last timestep (same layout) */
    int64_t out_row_pitch;    /* RRB_MEM_HOST: elements between consecutive [N] rows of the caller's output arrays
                                 (0 = N, densely packed): lets a call fill a member block of a wider [T, N_total] array */
} rrb_opts;

/* ---- library / device management ------------------------------------------------------- */
RRB_API int rrb_version(void);
RRB_API int rrb_device_count(void);            /* number of visible CUDA devices, 0 when none / no driver */
RRB_API int rrb_init(int device);              /* create the per-device context eagerly (optional) */
RRB_API int rrb_init_devices(const int32_t* devices, int n);  /* ... for a list of devices (NULL = 0 .. n-1) */
/* rows of the state_in / state_out arrays.  Layout (one [N] row each):
 *   ABC      S                                      (abcmodel_model.py:50)
 *   HBV-Edu  snow, soil, s1, s2                     (hbvedu_model.py:78-81)
 *   GR4J     S, R, uh1[0..C1), uh2[0..C2)           (gr4j_model.py:64-65, 82-83) with the unit-hydrograph capacity
 *            class (C1, C2) of the batch: (3, 7) for x4_max <= 3, (4, 9) <= 4, (10, 21) <= 10, (64, 129) <= RRB_MAX_X4;
 *            slots past a member's own ceil(x4) / ceil(2 x4 + 1) are 0.  A resumed call must use the same class.
 * Returns -1 for an unknown model or x4_max > RRB_MAX_X4. */
RRB_API int rrb_state_rows(int model, double x4_max);
RRB_API int rrb_shutdown(void);                /* release every context, stream and scratch pool */
RRB_API const char* rrb_last_error(void);      /* message of the last failing call on this thread */
RRB_API int rrb_synchronize(int device);       /* wait for all library work on the device */

/* pinned host buffers for RRB_MEM_HOST callers (pooled; freeing returns the block to the pool) */
RRB_API void* rrb_host_alloc(size_t bytes);
RRB_API void rrb_host_free(void* ptr);
RRB_API void rrb_host_pool_trim(void);         /* release pooled pinned blocks back to the OS */

/* ---- ensemble simulations --------------------------------------------------------------- */

/* ABC model.  params[N][3] = (a, b, c).  qsim, storage: [T, N].  qsim[0,:] = 0 (t = 0 is not
 * simulated, abcmodel_model.py:53). */
RRB_API int rrb_abc_simulate(const double* prec, int64_t T, double initial_state, const double* params, int64_t N,
                     double* qsim, double* storage /* nullable */, const rrb_opts* opts);

/* HBV-Edu.  month0[T] is the 0-based month index (the wrapper subtracts 1, hbvedu.py:164).
 * PE_m, T_m: [12].  inits[4] = (snow, soil, s1, s2), always host memory.  params[N][11] =
 * (T_t, DD, FC, Beta, C, PWP, K_0, K_1, K_2, K_p, L).  All outputs [T, N]; row 0 holds the
 * initial states and qsim[0,:] = 0 (hbvedu_model.py:78-84). */
RRB_API int rrb_hbvedu_simulate(const double* temp, const double* prec, const int8_t* month0, const double* PE_m,
                        const double* T_m, int64_t T, const double* inits, const double* params, int64_t N,
                        double* qsim, double* snow, double* soil, double* s1, double* s2 /* nullable x4 */,
                        const rrb_opts* opts);

/* HBV-Edu for C independent catchments that share T and the ensemble size N (the
 * reference has no batched equivalent -- a user loops HBVEdu.simulate over basins).  temp, prec, month0:
 * [C, T]; PE_m, T_m: [C, 12]; inits: host [C, 4]; params: [C, N, 11]; outputs [C, T, N]; opts->qobs [C, T] and
 * opts->mse [C, N].  One kernel launch covers every catchment (grid.y = catchment); with host buffers the
 * catchments are processed in chunks whose D2H overlaps the next chunk's kernel. */
RRB_API int rrb_hbvedu_simulate_multi(const double* temp, const double* prec, const int8_t* month0, const double* PE_m,
                                      const double* T_m, int64_t C, int64_t T, const double* inits,
                                      const double* params, int64_t N, double* qsim, double* snow, double* soil,
                                      double* s1, double* s2 /* nullable x4 */, const rrb_opts* opts);

/* GR4J.  s_init, r_init are fractions of x1 / x3 (gr4j_model.py:64-65).  params[N][4] =
 * (x1, x2, x3, x4).  Every member is simulated (the reference returns after member 0 when
 * return_storage=False, gr4j.py:178 -- a bug this API does not reproduce). */
RRB_API int rrb_gr4j_simulate(const double* prec, const double* etp, int64_t T, double s_init, double r_init,
                      const double* params, int64_t N, double* qsim, double* s_store, double* r_store /* nullable x2 */,
                      const rrb_opts* opts);

/* Cemaneige snow routine.  prec, mean_temp, frac_solid: the preprocessed [T, L] layer arrays
 * (cemaneige.py:198-219).  params[N][param_stride], fields 0,1 = (CTG, Kf) -- param_stride = 2
 * for Cemaneige records, 6 to read CemaneigeGR4J records in place.  outflow [T, N]; G, eTG
 * [T, L, N]. */
RRB_API int rrb_cemaneige_simulate(const double* prec, const double* mean_temp, const double* frac_solid, int64_t T,
                           int64_t L, double snow_pack_init, double thermal_state_init, const double* params,
                           int64_t param_stride, int64_t N, double* outflow, double* G, double* eTG /* nullable x2 */,
                           const rrb_opts* opts);

/* Cemaneige + GR4J, fused per timestep.  etp: [T].  inits[4] (always host memory) = (snow_pack_init,
 * thermal_state_init, s_init, r_init).  params[N][6] = (CTG, Kf, x1, x2, x3, x4). */
RRB_API int rrb_cemaneigegr4j_simulate(const double* prec, const double* mean_temp, const double* etp,
                               const double* frac_solid, int64_t T, int64_t L, const double* inits,
                               const double* params, int64_t N, double* qsim, double* G, double* eTG,
                               double* s_store, double* r_store /* nullable x4 */, const rrb_opts* opts);

/* ---- snow-ice family ----------------------------------------
 * Same layer arrays as rrb_cemaneigegr4j_simulate.  frac_ice: [L] glaciated fraction per layer.
 * inits (always host memory) = (snow_pack_init, thermal_state_init, s_init, r_init) for the model without
 * hysteresis and (snow_pack_init, thermal_state_init, sca_init, s_init, r_init) for the two Hyst models.
 * Storage outputs: pass all of a model's storages or none.  G, eTG, sca: [T, L, N]; the others [T, N].
 * The member-independent `rain` array the Hyst wrappers also return (cemaneigehyst_model.py:98-99) is
 * prec - prec * frac_solid and is formed by the host layer. */

/* CemaneigeGR4JIce: run_cemaneigegr4jice rrmpg/models/cemaneigegr4jice_model.py:16, member loop
 * cemaneigegr4jice.py:262-284.  params[N][7] = (CTG, Kf, x1, x2, x3, x4, DDF). */
RRB_API int rrb_cemaneigegr4jice_simulate(const double* prec, const double* mean_temp, const double* etp,
                                          const double* frac_ice, const double* frac_solid, int64_t T, int64_t L,
                                          const double* inits, const double* params, int64_t N, double* qsim,
                                          double* G, double* eTG, double* s_store, double* r_store,
                                          double* icemelt /* nullable x5 */, const rrb_opts* opts);

/* CemaneigeHystGR4J: run_cemaneigehystgr4j rrmpg/models/cemaneigehystgr4j_model.py:17, member loop
 * cemaneigehystgr4j.py:262-286.  params[N][8] = (CTG, Kf, Thacc, Rsp, x1, x2, x3, x4). */
RRB_API int rrb_cemaneigehystgr4j_simulate(const double* prec, const double* mean_temp, const double* etp,
                                           const double* frac_solid, int64_t T, int64_t L, const double* inits,
                                           const double* params, int64_t N, double* qsim, double* G, double* eTG,
                                           double* s_store, double* r_store, double* sca /* nullable x5 */,
                                           const rrb_opts* opts);

/* CemaneigeHystGR4JIce: run_cemaneigehystgr4jice rrmpg/models/cemaneigehystgr4jice_model.py:18, member loop
 * cemaneigehystgr4jice.py:276-304.  params[N][9] = (CTG, Kf, Thacc, Rsp, x1, x2, x3, x4, DDF). */
RRB_API int rrb_cemaneigehystgr4jice_simulate(const double* prec, const double* mean_temp, const double* etp,
                                              const double* frac_ice, const double* frac_solid, int64_t T, int64_t L,
                                              const double* inits, const double* params, int64_t N, double* qsim,
                                              double* G, double* eTG, double* s_store, double* r_store, double* sca,
                                              double* icemelt, double* snowmelt /* nullable x7 */,
                                              const rrb_opts* opts);

/* ---- catchment batches of GR4J and CemaneigeGR4J -----------------
 * C independent catchments with N members each in one call (grid.y = catchment); the reference has no
 * equivalent (a user loops Model.simulate over basins).  Every array gains a leading [C] axis: prec / etp
 * [C, T], layer arrays [C, T, L], params [C, N, k], outputs [C, T, N] (G, eTG: [C, T, L, N]), opts->qobs
 * [C, T], opts->mse [C, N].  inits (HOST memory in both modes): GR4J [C, 2] = (s_init, r_init) per catchment,
 * CemaneigeGR4J [C, 4] = (snow_pack_init, thermal_state_init, s_init, r_init).  Results are bit-identical
 * to C calls of the single-catchment entry points. */
RRB_API int rrb_gr4j_simulate_multi(const double* prec, const double* etp, int64_t C, int64_t T, const double* inits,
                                    const double* params, int64_t N, double* qsim, double* s_store,
                                    double* r_store /* nullable x2 */, const rrb_opts* opts);
RRB_API int rrb_cemaneigegr4j_simulate_multi(const double* prec, const double* mean_temp, const double* etp,
                                             const double* frac_solid, int64_t C, int64_t T, int64_t L,
                                             const double* inits, const double* params, int64_t N, double* qsim,
                                             double* G, double* eTG, double* s_store, double* r_store /* nullable x4 */,
                                             const rrb_opts* opts);

/* Catchment batches of the remaining models (round 2; same conventions: every array gains a leading [C] axis, results
 * bit-identical to C single-catchment calls).  inits, HOST memory in both modes: ABC [C] initial storages; Cemaneige
 * [C, 2] = (snow_pack_init, thermal_state_init); the snow-ice couplings [C, 4] = (snow_pack_init, thermal_state_init,
 * s_init, r_init) without and [C, 5] = (snow_pack_init, thermal_state_init, sca_init, s_init, r_init) with hysteresis --
 * the argument order of the single-catchment calls.  frac_ice: [C, L]. */
RRB_API int rrb_abc_simulate_multi(const double* prec, int64_t C, int64_t T, const double* inits, const double* params,
                                   int64_t N, double* qsim, double* storage /* nullable */, const rrb_opts* opts);
RRB_API int rrb_cemaneige_simulate_multi(const double* prec, const double* mean_temp, const double* frac_solid, int64_t C,
                                         int64_t T, int64_t L, const double* inits, const double* params,
                                         int64_t param_stride, int64_t N, double* outflow, double* G,
                                         double* eTG /* nullable x2 */, const rrb_opts* opts);
RRB_API int rrb_cemaneigegr4jice_simulate_multi(const double* prec, const double* mean_temp, const double* etp,
                                                const double* frac_ice, const double* frac_solid, int64_t C, int64_t T,
                                                int64_t L, const double* inits, const double* params, int64_t N,
                                                double* qsim, double* G, double* eTG, double* s_store, double* r_store,
                                                double* icemelt /* nullable x5 */, const rrb_opts* opts);
RRB_API int rrb_cemaneigehystgr4j_simulate_multi(const double* prec, const double* mean_temp, const double* etp,
                                                 const double* frac_solid, int64_t C, int64_t T, int64_t L,
                                                 const double* inits, const double* params, int64_t N, double* qsim,
                                                 double* G, double* eTG, double* s_store, double* r_store,
                                                 double* sca /* nullable x5 */, const rrb_opts* opts);
RRB_API int rrb_cemaneigehystgr4jice_simulate_multi(const double* prec, const double* mean_temp, const double* etp,
                                                    const double* frac_ice, const double* frac_solid, int64_t C,
                                                    int64_t T, int64_t L, const double* inits, const double* params,
                                                    int64_t N, double* qsim, double* G, double* eTG, double* s_store,
                                                    double* r_store, double* sca, double* icemelt,
                                                    double* snowmelt /* nullable x7 */, const rrb_opts* opts);

/* ---- member-independent layer preprocessing of the Cemaneige family, on the device --------
 * Replaces extrapolate_precipitation (rrmpg/models/cemaneige_utils.py:101-158), extrapolate_temperature
 * (:161-208) and calculate_solid_fraction (:16-98), i.e. the [T] -> [T, L] step of Cemaneige.simulate
 * (rrmpg/models/cemaneige.py:198-219) and CemaneigeGR4J.simulate (rrmpg/models/cemaneigegr4j.py:199-219).
 * Station series prec / mean_temp / min_temp / max_temp: [T] (host or device per opts->mem).  Per-layer scalars
 * are HOST memory in both modes and computed by the caller as the reference does (libm exp), so the outputs are
 * bit-identical to numba: prec_factor[L] = exp((min(z_l, 4000) - z_station) * 0.0004), delta_temp[L] =
 * (z_l - z_station) * -0.0065, flags[L] = OR of RRB_LAYER_*.  Outputs [T, L] C-order: layer precipitation,
 * layer mean temperature, fraction of solid precipitation -- the inputs of rrb_cemaneige*_simulate. */
#define RRB_LAYER_SCALE_PREC 1 /* multiply the station precipitation by prec_factor (else pass it through) */
#define RRB_LAYER_SHIFT_TEMP 2 /* add delta_temp to the station temperatures (else pass them through) */
#define RRB_LAYER_HIGH 4       /* altitude >= 1500 m: solid fraction from the mean temperature (:76-88) */
RRB_API int rrb_snow_layers(const double* prec, const double* mean_temp, const double* min_temp, const double* max_temp,
                            int64_t T, int64_t L, const double* prec_factor, const double* delta_temp,
                            const int32_t* flags, double* layer_prec, double* layer_mean_temp, double* frac_solid,
                            const rrb_opts* opts);

/* ---- host-side checks of the FAST math (no GPU needed; used by the CPU test-suite) ------- */
RRB_API void rrb_host_fast_pow(const double* x, const double* y, int64_t n, double* out);
RRB_API void rrb_host_fast_exp2m1(const double* z, int64_t n, double* out);
/* the soil-chain step of the HBV-Edu FAST kernel (rr_math.cuh, hbv_pow_step_twin): prec_eff = liquid (soil/FC)^Beta
 * (rrmpg/models/hbvedu_model.py:99) and soil_new = soil_partial - prec_eff, same operation sequence as the device code */
RRB_API void rrb_host_hbv_pow_step(const double* soil, const double* FC, const double* Beta, const double* liquid,
                                   const double* soil_partial, int64_t n, double* prec_eff, double* soil_new);

#ifdef __cplusplus
}
#endif
#endif /* RRMPG_B200_H */
