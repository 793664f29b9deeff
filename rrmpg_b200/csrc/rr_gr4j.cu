// rr_gr4j.cu -- GR4J ensemble kernel.
// Restates run_gr4j (rrmpg/models/gr4j_model.py:16-157) for N members at once and replaces the
// member loop of GR4J.simulate (rrmpg/models/gr4j.py:169-178).  Unlike that loop -- which
// returns after member 0 when return_storage=False (gr4j.py:178) -- every member is simulated,
// as the docstring (gr4j.py:94-97) promises.
// Packed forcing per timestep: F[t] = { prec, etp }.  All T steps are simulated (t = 0 too).
#include "rr_common.cuh"
#include "rr_gr4j.cuh"
#include "rr_kernels.h"
#include "rr_objective.cuh"

namespace rrb {

__global__ void gr4j_pack_kernel(const double* __restrict__ prec, const double* __restrict__ etp, int64_t T,
                                 int64_t Tpad, double* __restrict__ F, uint32_t* __restrict__ fflag, int64_t fstride) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= Tpad) return;
    const int64_t c = blockIdx.y;  // catchment
    prec += c * T; etp += c * T; F += c * fstride;
    fflag = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(fflag) + c * fstride);
    double2 v = make_double2(0.0, 0.0);
    if (t < T) {
        v.x = prec[t];
        v.y = etp[t];
        if (!forcing_value_sane(v.x) || !forcing_value_sane(v.y)) atomicOr(fflag, 1u);
    }
    reinterpret_cast<double2*>(F)[t] = v;
}

cudaError_t pack_gr4j(const double* prec, const double* etp, int64_t T, double* F, cudaStream_t s, int count) {
    int64_t Tpad = padded_steps(T, kGr4jTT);
    const int64_t fstride = forcing_stride_flagged(T, kGr4jTT, kGr4jR);
    uint32_t* fflag = forcing_flag(F, T, kGr4jTT, kGr4jR);
    cudaError_t e = cudaMemset2DAsync(fflag, sizeof(double) * (size_t)fstride, 0, kForcingFlagBytes, (size_t)count, s);
    if (e != cudaSuccess) return e;
    gr4j_pack_kernel<<<dim3((unsigned)((Tpad + 255) / 256), (unsigned)count), 256, 0, s>>>(prec, etp, T, Tpad, F, fflag,
                                                                                        fstride);
    return cudaGetLastError();
}

struct Gr4jF {  // forcing of one timestep
    double prec, etp;
    static __device__ __forceinline__ Gr4jF load(uint32_t addr) {
        const double2 a = lds_f64x2(addr);
        return Gr4jF{a.x, a.y};
    }
};

constexpr int kGr4jGroup = 2;

// PLAIN = discharge only (no storages, no fused objective): the output flags are compile-time constants
template <class Member, bool FAST, bool PLAIN>
__global__ void gr4j_kernel(const double* __restrict__ F, double s_init, double r_init,
                            const double* __restrict__ params, int64_t N, double* __restrict__ qsim,
                            double* __restrict__ s_store, double* __restrict__ r_store, Slab slab,
                            Objective obj, const uint32_t* __restrict__ fflag, Batch batch) {
    if (batch.count > 1) {  // blockIdx.y = catchment: shift every per-catchment pointer
        const int64_t c = blockIdx.y;
        F += c * batch.forcing_stride;
        fflag = reinterpret_cast<const uint32_t*>(reinterpret_cast<const double*>(fflag) + c * batch.forcing_stride);
        params += c * N * 4;
        if (qsim) qsim += c * batch.out_stride;
        if (s_store) { s_store += c * batch.out_stride; r_store += c * batch.out_stride; }
        if (obj.qobs) { obj.qobs += c * obj.T; obj.mse += c * N; }
        if (obj.obs_stats) { obj.obs_mean = obj.obs_stats[2 * c]; obj.obs_std = obj.obs_stats[2 * c + 1]; }
    }
    if (batch.inits) {  // also for a batch (or a chunk of a batch) of ONE catchment
        const int64_t c = batch.count > 1 ? (int64_t)blockIdx.y : 0;
        s_init = batch.inits[4 * c];
        r_init = batch.inits[4 * c + 1];
    }
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // threads past the end of the ensemble recompute member N-1 and store the same values to the same
    // addresses: no predicate lives in the time loop
    const int64_t i = gi < N ? gi : N - 1;
    const bool WRITEQ = PLAIN || qsim != nullptr, STORAGE = !PLAIN && s_store != nullptr,
               OBJ = !PLAIN && obj.qobs != nullptr;  // CTA-uniform
    extern __shared__ __align__(128) unsigned char rrb_smem[];
    // dynamic shared memory = [forcing ring | FAST tables | unit hydrograph ordinates (long-hydrograph class only)]
    constexpr size_t kOrdOff = forcing_smem_bytes<kGr4jR, kGr4jTT>() + (FAST ? fastmath_smem_bytes() : 0);
    Member m;
    m.init(params + 4 * i, s_init, r_init, smem_u32(rrb_smem + kOrdOff));  // record = (x1, x2, x3, x4), gr4j.py:57-60
    ObjAcc acc;
    acc.reset();
    if (slab_loads_state(slab)) {
        m.load(slab.state, N, i);
        if (OBJ && slab.t_begin > 0) acc.load(slab.state, Member::kStateSlots, N, i, obj);
    }
    int64_t stride = N;
    pin(stride);
    const int64_t off = i + (slab.t_begin - slab.row0) * N;
    double* q_o = WRITEQ ? qsim + off : nullptr;
    double* s_o = STORAGE ? s_store + off : nullptr;
    double* r_o = STORAGE ? r_store + off : nullptr;

    uint32_t tb = 0;
    bool use_fast = false;
    Exp2Regs ek{};
    if constexpr (FAST) {
        tb = smem_u32(fastmath_tables_to_smem(rrb_smem + forcing_smem_bytes<kGr4jR, kGr4jTT>()));
        pin(tb);
        // CTA-uniform choice of the step (Gr4jMember: FAST path contract); the barrier also publishes the tables
        use_fast = __syncthreads_and(m.sane && *fflag == 0u) != 0;
        if (use_fast) {
            ek = load_exp2_regs(tb);
            m.enter_fast();
        }
    }

    auto run = [&](auto fast_c) {
        stream_forcing_grouped<kGr4jR, kGr4jTT, kGr4jGroup, Gr4jF>(
            F, slab.t_begin, slab.t_end, [&](auto gc, int64_t t0, const Gr4jF* f) {
                constexpr int G = decltype(gc)::value;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double qv = gr4j_step(fast_c, m, f[g].prec, f[g].etp, tb, ek);
                    if (WRITEQ) {
                        st_stream(q_o, qv);
                        q_o += stride;
                    }
                    if (STORAGE) {
                        st_stream(s_o, m.S);
                        st_stream(r_o, m.R);
                        s_o += stride;
                        r_o += stride;
                    }
                    if (OBJ) acc.add(obj.qobs[t0 + g], qv, obj);
                }
            });
    };
    if constexpr (FAST) {
        if (use_fast) run(ic<1>{});
        else run(ic<0>{});
    } else {
        run(ic<0>{});
    }

    if (gi < N) {
        if (slab.save_state) {
            m.save(slab.state, N, i);
            if (OBJ && slab.save_state == 1) acc.save(slab.state, Member::kStateSlots, N, i, obj);
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i] = acc.finish(obj);
    }
}

// unit-hydrograph capacity classes: (C1, C2) covers x4 <= C1 (C2 = 2 C1 + 1)
static int uh_class(double x4_max) {
    if (!(x4_max <= 64.0)) return -1;
    if (x4_max <= 3.0) return 0;   // default bounds (1.1, 2.9), rrmpg/models/gr4j.py:54
    if (x4_max <= 4.0) return 1;   // the reference's CemaneigeGR4J fixture (x4 = 3.098)
    if (x4_max <= 10.0) return 2;  // Hyst-family bounds go to 10
    return 3;                      // local-memory fallback
}

int state_slots_gr4j(double x4_max) {
    switch (uh_class(x4_max)) {
        case 0: return 2 + 3 + 7 + kObjSlots;
        case 1: return 2 + 4 + 9 + kObjSlots;
        case 2: return 2 + 10 + 21 + kObjSlots;
        default: return Gr4jMemberDyn::kStateSlots + kObjSlots;
    }
}

template <class Member, bool FAST>
static cudaError_t launch_variant(const double* F, int64_t T, double s_init, double r_init, const double* params, int64_t N,
                                  double* qsim, double* s_store, double* r_store, const Slab& slab,
                                  const Objective& obj, const LaunchCfg& cfg, const Batch& batch) {
    int block = cfg.block > 0 ? cfg.block : pick_block(N * batch.count, cfg.sm_count, N >= 128 ? 128 : 64);
    const bool plain = qsim && !s_store && !obj.qobs;
    if (cfg.block <= 0 && batch.count == 1 && Member::kOrdSmemBytes == 0) {  // 9 .. 16 warps per SM: one CTA per SM
        const int cap = plain ? kernel_max_threads(gr4j_kernel<Member, FAST, true>) : kernel_max_threads(gr4j_kernel<Member, FAST, false>);
        const int b = one_cta_block(N, cfg.sm_count, cap < 512 ? cap : 512, 9);
        if (b) block = b;
    }
    if (Member::kOrdSmemBytes > 0 && block > kOrdThreads) block = kOrdThreads;  // ordinate columns in shared memory
    const dim3 grid((unsigned)((N + block - 1) / block), (unsigned)batch.count);
    const size_t smem = forcing_smem_bytes<kGr4jR, kGr4jTT>() + (FAST ? fastmath_smem_bytes() : 0) + Member::kOrdSmemBytes;
    const uint32_t* fflag = forcing_flag(F, T, kGr4jTT, kGr4jR);
    if (plain)
        gr4j_kernel<Member, FAST, true><<<grid, block, smem, cfg.stream>>>(F, s_init, r_init, params, N, qsim, s_store,
                                                                           r_store, slab, obj, fflag, batch);
    else
        gr4j_kernel<Member, FAST, false><<<grid, block, smem, cfg.stream>>>(F, s_init, r_init, params, N, qsim, s_store,
                                                                            r_store, slab, obj, fflag, batch);
    return cudaGetLastError();
}

cudaError_t launch_gr4j(const double* F, int64_t T, double s_init, double r_init, const double* params,
                        int64_t N, double x4_max, double* qsim, double* s_store, double* r_store,
                        const Slab& slab, const Objective& obj, const LaunchCfg& cfg, const Batch& batch) {
    if (N <= 0 || batch.count <= 0) return cudaSuccess;
    const bool fast = cfg.math == RRB_MATH_FAST_;
#define RRB_GO(M_, F_) return launch_variant<M_, F_>(F, T, s_init, r_init, params, N, qsim, s_store, r_store, slab, obj, cfg, batch)
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

}  // namespace rrb
