// rr_abc.cu -- ABC model ensemble kernel.
// Restates run_abcmodel (rrmpg/models/abcmodel_model.py:16-60) for N members at once and
// replaces the member loop of ABCModel.simulate (rrmpg/models/abcmodel.py:174-181).
// Arithmetic is + and * only and the TU is built with -fmad=false, so results are
// bit-identical to the numba path.  HBM-bound: 8 B stored per member-timestep (16 B with
// return_storage), ~5 fp64 instructions.
#include "rr_common.cuh"
#include "rr_kernels.h"
#include "rr_objective.cuh"

namespace rrb {

__global__ void abc_pack_kernel(const double* __restrict__ prec, int64_t T, int64_t Tpad, double* __restrict__ F) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t c = blockIdx.y;  // catchment
    if (t < Tpad) F[c * Tpad + t] = (t < T) ? prec[c * T + t] : 0.0;
}

cudaError_t pack_abc(const double* prec, int64_t T, double* F, cudaStream_t s, int count) {
    int64_t Tpad = padded_steps(T, kAbcTT);
    abc_pack_kernel<<<dim3((unsigned)((Tpad + 255) / 256), (unsigned)count), 256, 0, s>>>(prec, T, Tpad, F);
    return cudaGetLastError();
}

// blockIdx.y = catchment of a batch (rrb_abc_simulate_multi): shift every per-catchment pointer
#define ABC_BATCH_PROLOGUE                                                            \
    if (batch.count > 1) {                                                            \
        const int64_t c = blockIdx.y;                                                 \
        F += c * batch.forcing_stride;                                                \
        params += c * N * 3;                                                          \
        if (qsim) qsim += c * batch.out_stride;                                       \
        if (storage) storage += c * batch.out_stride;                                 \
        if (obj.qobs) { obj.qobs += c * obj.T; obj.mse += c * N; }                    \
        if (obj.obs_stats) { obj.obs_mean = obj.obs_stats[2 * c]; obj.obs_std = obj.obs_stats[2 * c + 1]; } \
    }                                                                                 \
    if (batch.inits) s0 = batch.inits[4 * (batch.count > 1 ? (int64_t)blockIdx.y : 0)];

template <bool STORAGE, bool OBJ>
__global__ void abc_kernel(const double* __restrict__ F, double s0, const double* __restrict__ params, int64_t N,
                           double* __restrict__ qsim, double* __restrict__ storage, Slab slab,
                           Objective obj, Batch batch) {
    ABC_BATCH_PROLOGUE
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool active = gi < N;
    const int64_t i = active ? gi : N - 1;
    // params record = (a, b, c), rrmpg/models/abcmodel.py:53-55
    const double a = params[3 * i + 0], b = params[3 * i + 1], c = params[3 * i + 2];
    const double omab = 1 - a - b;  // loop invariants of abcmodel_model.py:56,59
    const double omc = 1 - c;
    double S = s0;
    ObjAcc acc;
    acc.reset();
    if (slab_loads_state(slab)) {
        S = slab.state[i];
        if (OBJ && slab.t_begin > 0) acc.load(slab.state, 1, N, i, obj);
    }
    const bool skip_t0 = !slab.resume;
    double* q = qsim ? qsim + i - slab.row0 * N : nullptr;
    double* st = STORAGE ? storage + i - slab.row0 * N : nullptr;

    stream_forcing<kAbcR, kAbcTT>(F, slab.t_begin, slab.t_end, [&](int64_t t, const double* f) {
        const double p = f[0];
        double qv;
        if (t == 0 && skip_t0) {
            qv = 0.0;  // abcmodel_model.py:53 -- the loop starts at t = 1, qsim[0] stays 0
        } else {
            qv = omab * p + c * S;  // :56
            S = omc * S + a * p;    // :59
        }
        if (active) {
            if (q) st_stream(q + t * N, qv);
            if (STORAGE) st_stream(st + t * N, S);
        }
        if (OBJ) acc.add(obj.qobs[t], qv, obj);
    });

    if (active) {
        if (slab.save_state) {
            slab.state[i] = S;
            if (OBJ && slab.save_state == 1) acc.save(slab.state, 1, N, i, obj);
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i] = acc.finish(obj);
    }
}


// Discharge-only fast variant: two members per thread, one 16-byte streaming store per step (the ensemble axis is
// contiguous, so members 2i and 2i+1 are neighbours in every output row), t = 0 peeled, running output pointer, no
// predicate in the time loop.  Needs an even N and 16-byte aligned rows; launch_abc falls back to abc_kernel otherwise.
// Same operations in the same order per member: bit-identical.
__device__ __forceinline__ void st_stream_v2(double* p, double x, double y) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}
__global__ void abc_pair_kernel(const double* __restrict__ F, double s0, const double* __restrict__ params, int64_t N,
                                double* __restrict__ qsim, Slab slab, Batch batch) {
    {
        double* storage = nullptr;
        Objective obj{nullptr, nullptr, 0, 0, 0.0, 1.0, nullptr};
        ABC_BATCH_PROLOGUE
    }
    const int64_t pairs = N / 2;
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = 2 * (gi < pairs ? gi : pairs - 1);  // surplus threads recompute the last pair
    const double a0 = params[3 * i + 0], b0 = params[3 * i + 1], c0 = params[3 * i + 2];
    const double a1 = params[3 * i + 3], b1 = params[3 * i + 4], c1 = params[3 * i + 5];
    const double omab0 = 1 - a0 - b0, omc0 = 1 - c0;  // loop invariants of abcmodel_model.py:56,59
    const double omab1 = 1 - a1 - b1, omc1 = 1 - c1;
    double S0 = s0, S1 = s0;
    int64_t t_first = slab.t_begin;
    if (slab_loads_state(slab)) {
        S0 = slab.state[i];
        S1 = slab.state[i + 1];
    }
    double* q = qsim + i + (slab.t_begin - slab.row0) * N;
    if (slab.t_begin == 0 && slab.t_end > 0 && !slab.resume) {  // abcmodel_model.py:53 -- the loop starts at t = 1, qsim[0] stays 0
        st_stream_v2(q, 0.0, 0.0);
        q += N;
        t_first = 1;
    }
    stream_forcing<kAbcR, kAbcTT>(F, t_first, slab.t_end, [&](int64_t, const double* f) {
        const double p = f[0];
        const double q0 = omab0 * p + c0 * S0;  // :56
        const double q1 = omab1 * p + c1 * S1;
        S0 = omc0 * S0 + a0 * p;                // :59
        S1 = omc1 * S1 + a1 * p;
        st_stream_v2(q, q0, q1);
        q += N;
    });
    if (slab.save_state && gi < pairs) {
        slab.state[i] = S0;
        slab.state[i + 1] = S1;
    }
}

int state_slots_abc() { return 1 + kObjSlots; }

cudaError_t launch_abc(const double* F, int64_t T, double s0, const double* params, int64_t N, double* qsim,
                       double* storage, const Slab& slab, const Objective& obj, const LaunchCfg& cfg, const Batch& batch) {
    (void)T;
    if (N <= 0 || batch.count <= 0) return cudaSuccess;
    const int block = cfg.block > 0 ? cfg.block : pick_block(N * batch.count, cfg.sm_count, 256);
    const dim3 grid((unsigned)((N + block - 1) / block), (unsigned)batch.count);
    const size_t smem = forcing_smem_bytes<kAbcR, kAbcTT>();
    const bool st = storage != nullptr, ob = obj.qobs != nullptr;
    if (qsim && !st && !ob && (N % 2) == 0 && (reinterpret_cast<uintptr_t>(qsim) % 16) == 0) {
        const int64_t pairs = N / 2;
        const int pblock = cfg.block > 0 ? cfg.block : pick_block(pairs * batch.count, cfg.sm_count, 128);
        abc_pair_kernel<<<dim3((unsigned)((pairs + pblock - 1) / pblock), (unsigned)batch.count), pblock, smem, cfg.stream>>>(
            F, s0, params, N, qsim, slab, batch);
        return cudaGetLastError();
    }
#define RRB_ABC(S_, O_) \
    abc_kernel<S_, O_><<<grid, block, smem, cfg.stream>>>(F, s0, params, N, qsim, storage, slab, obj, batch)
    if (st && ob) RRB_ABC(true, true);
    else if (st) RRB_ABC(true, false);
    else if (ob) RRB_ABC(false, true);
    else RRB_ABC(false, false);
#undef RRB_ABC
    return cudaGetLastError();
}

}  // namespace rrb
