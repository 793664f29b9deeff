// rr_objective.cuh -- the fused per-member objective of the ensemble kernels.
//
// rrmpg.tools.monte_carlo (rrmpg/tools/monte_carlo.py:64-73) and every model's _loss (e.g. rrmpg/models/hbvedu.py:310-346,
// rrmpg/models/cemaneigehystgr4j.py:600-605) simulate a member, materialise its [T] discharge series and reduce it with
// calc_mse / calc_nse / calc_kge (rrmpg/utils/metrics.py:29-78, 110-136, 139-188).  Here the reduction runs in registers
// next to the recurrence: no [T, N] array exists unless the caller asks for it.
//
//   MSE  mean((obs - sim)^2)                                  sse / T
//   NSE  1 - sum((sim - obs)^2) / sum((obs - mean(obs))^2)    1 - sse / (T std_obs^2)
//   KGE  1 - sqrt((r-1)^2 + (alpha-1)^2 + (beta-1)^2), r = pearson(obs, sim), alpha = std(sim)/std(obs),
//        beta = mean(sim)/mean(obs)   from the one-pass sums of e = sim - mean_obs:  sum e, sum e^2, sum e (obs - mean_obs)
//        (shifted by the observations' mean, which is known before the launch, so the variance does not cancel)
// The observations' mean and population standard deviation come from the caller (numpy, as the reference computes
// them).  Sequential in-thread sums instead of numpy's pairwise ones: stated tolerance rtol 1e-9 (tests).
#pragma once
#include "rr_common.cuh"
#include "rr_kernels.h"

namespace rrb {

// Lean variant for kernels that template on the metric (HBV-Edu): MSE and NSE need the one sum only, and a KGE
// accumulator that is never used still costs six registers per member.
struct ObjAccSse {
    double sse;
    __device__ __forceinline__ void reset() { sse = 0.0; }
    __device__ __forceinline__ void load(const double* state, int64_t slot, int64_t N, int64_t i, const Objective&) {
        sse = state[slot * N + i];
    }
    __device__ __forceinline__ void save(double* state, int64_t slot, int64_t N, int64_t i, const Objective&) const {
        state[slot * N + i] = sse;
    }
    __device__ __forceinline__ void add(double obs, double sim, const Objective&) {
        const double d = obs - sim;
        sse += d * d;
    }
    __device__ __forceinline__ double finish(const Objective& o) const {
        const double n = (double)o.T;
        return (o.kind == RRB_OBJ_NSE_) ? 1.0 - sse / (n * o.obs_std * o.obs_std) : sse / n;
    }
};

// KGE only, for kernels that template on the metric: three sums, no branch in the time loop (a CTA-uniform branch
// would still cut the straight-line group bodies of hbv_fast2_kernel into short basic blocks)
struct ObjAccKge {
    double se, see, seo;
    __device__ __forceinline__ void reset() { se = see = seo = 0.0; }
    __device__ __forceinline__ void load(const double* state, int64_t slot, int64_t N, int64_t i, const Objective&) {
        se = state[(slot + 1) * N + i];
        see = state[(slot + 2) * N + i];
        seo = state[(slot + 3) * N + i];
    }
    __device__ __forceinline__ void save(double* state, int64_t slot, int64_t N, int64_t i, const Objective&) const {
        state[(slot + 1) * N + i] = se;
        state[(slot + 2) * N + i] = see;
        state[(slot + 3) * N + i] = seo;
    }
    __device__ __forceinline__ void add(double obs, double sim, const Objective& o) {
        const double e = sim - o.obs_mean;
        se += e;
        see = fma(e, e, see);
        seo = fma(e, obs - o.obs_mean, seo);
    }
    __device__ __forceinline__ double finish(const Objective& o) const {
        const double n = (double)o.T;
        const double me = se / n;
        const double var = see / n - me * me;
        const double sd = sqrt(var > 0.0 ? var : 0.0);
        const double r = (seo / n) / (sd * o.obs_std);
        const double alpha = sd / o.obs_std;
        const double beta = (o.obs_mean + me) / o.obs_mean;
        return 1.0 - sqrt((r - 1.0) * (r - 1.0) + (alpha - 1.0) * (alpha - 1.0) + (beta - 1.0) * (beta - 1.0));
    }
};

struct ObjAcc {
    double sse, se, see, seo;
    __device__ __forceinline__ void reset() { sse = se = see = seo = 0.0; }
    // slot = index of the first of the kObjSlots objective rows in the [slots, N] carry buffer
    __device__ __forceinline__ void load(const double* state, int64_t slot, int64_t N, int64_t i, const Objective& o) {
        sse = state[slot * N + i];
        if (o.kind == RRB_OBJ_KGE_) {
            se = state[(slot + 1) * N + i];
            see = state[(slot + 2) * N + i];
            seo = state[(slot + 3) * N + i];
        }
    }
    __device__ __forceinline__ void save(double* state, int64_t slot, int64_t N, int64_t i, const Objective& o) const {
        state[slot * N + i] = sse;
        if (o.kind == RRB_OBJ_KGE_) {
            state[(slot + 1) * N + i] = se;
            state[(slot + 2) * N + i] = see;
            state[(slot + 3) * N + i] = seo;
        }
    }
    __device__ __forceinline__ void add(double obs, double sim, const Objective& o) {
        const double d = obs - sim;
        sse += d * d;                       // (obs - sim)**2, rrmpg/utils/metrics.py:131
        if (o.kind == RRB_OBJ_KGE_) {       // CTA-uniform, a few predicated instructions
            const double e = sim - o.obs_mean;
            se += e;
            see = fma(e, e, see);
            seo = fma(e, obs - o.obs_mean, seo);
        }
    }
    __device__ __forceinline__ double finish(const Objective& o) const {
        const double n = (double)o.T;
        if (o.kind == RRB_OBJ_MSE_) return sse / n;
        if (o.kind == RRB_OBJ_NSE_) return 1.0 - sse / (n * o.obs_std * o.obs_std);
        const double me = se / n;                       // mean(sim) - mean(obs)
        const double var = see / n - me * me;           // population variance of sim (np.std)
        const double sd = sqrt(var > 0.0 ? var : 0.0);
        const double cov = seo / n;                     // sum(obs - mean_obs) = 0: no cross term
        const double r = cov / (sd * o.obs_std);
        const double alpha = sd / o.obs_std;
        const double beta = (o.obs_mean + me) / o.obs_mean;
        return 1.0 - sqrt((r - 1.0) * (r - 1.0) + (alpha - 1.0) * (alpha - 1.0) + (beta - 1.0) * (beta - 1.0));
    }
};

}  // namespace rrb
