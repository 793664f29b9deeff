// rr_api.cu -- the C ABI of include/rrmpg_b200.h: per-device contexts, scratch pools, the
// host-buffer path (H2D forcing + params, time-slab pipeline that overlaps the kernels with the
// D2H of finished output rows) and the device-buffer path (everything enqueued on the caller's
// stream).  No model arithmetic lives here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rrmpg_b200.h"
#include "rr_kernels.h"
#include "rr_math.cuh"

namespace rrb {

static_assert((int)RRB_MATH_FAST == (int)RRB_MATH_FAST_ && (int)RRB_MATH_PRECISE == (int)RRB_MATH_PRECISE_, "math enum mismatch");
static_assert(RRB_MAX_LAYERS == kCemaMaxLayers, "layer cap mismatch");
static_assert((int)RRB_OBJ_MSE == (int)RRB_OBJ_MSE_ && (int)RRB_OBJ_NSE == (int)RRB_OBJ_NSE_ && (int)RRB_OBJ_KGE == (int)RRB_OBJ_KGE_,
              "objective enum mismatch");

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define RRB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? RRB_ENOMEM : RRB_ECUDA;                 \
            return fail(code_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)

// CTA size: the largest of {256,128,64,32} (<= max_block) whose grid leaves the SMs evenly
// loaded.  With one thread per member a 65 536-member ensemble is only ~443 threads per SM, so
// the rounding of CTAs per SM decides the makespan (e.g. 512 CTAs of 128 on 148 SMs = 3.46 -> 4
// rounds of work on some SMs; 1024 CTAs of 64 = 6.92 -> 7, 1% imbalance).
int pick_block(int64_t N, int sm_count, int max_block) {
    if (sm_count <= 0) sm_count = 148;
    int best = 32;
    double best_eff = -1.0;
    for (int b : {256, 128, 64, 32}) {
        if (b > max_block) continue;
        const int64_t ctas = (N + b - 1) / b;
        const int64_t per_sm = (ctas + sm_count - 1) / sm_count;
        const double eff = (double)N / ((double)per_sm * sm_count * b);
        if (eff > best_eff * 1.04) {  // prefer the larger CTA unless a smaller one is >4% better
            best_eff = eff;
            best = b;
        }
    }
    return best;
}

// max over members of one parameter field (device mode, when the caller gave no x4_max hint)
__global__ void max_field_kernel(const double* __restrict__ params, int64_t N, int64_t stride, int field,
                                 double* __restrict__ out) {
    __shared__ double red[1024];
    __shared__ int any_nan;
    if (threadIdx.x == 0) any_nan = 0;
    __syncthreads();
    double m = -1.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = params[i * stride + field];
        if (v != v) any_nan = 1;
        if (v > m) m = v;
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s && red[threadIdx.x + s] > red[threadIdx.x]) red[threadIdx.x] = red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = any_nan ? NAN : red[0];
}

// ----------------------------------------------------------------------------------------
// per-device context
// ----------------------------------------------------------------------------------------
enum BufSlot {
    B_RAW0 = 0, B_RAW1, B_RAW2, B_RAW3, B_RAW4,  // uploaded raw forcing arrays
    B_F,       // packed forcing
    B_GT,      // Cemaneige G_tresh[L]
    B_PARAMS,
    B_STATE,
    B_QOBS,
    B_MSE,
    B_SCALAR,
    B_OBS,     // catchment batches: (mean, std) of every qobs series
    B_OUT0,    // output ring: B_OUT0 + 2*k + slot, k < 8
    B_COUNT = B_OUT0 + 16
};

struct Ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    cudaEvent_t ev_scratch = nullptr;  // behind the last kernel of the latest device-mode call (see HostDrain)
    bool scratch_busy = false;
    std::mutex mu;
    void* buf[B_COUNT] = {};
    size_t cap[B_COUNT] = {};

    int ensure(int slot, size_t bytes, void** out) {
        if (bytes == 0) bytes = 16;
        if (cap[slot] < bytes) {
            if (buf[slot]) {
                RRB_CUDA(cudaStreamSynchronize(compute));
                RRB_CUDA(cudaStreamSynchronize(copy));
                RRB_CUDA(cudaFree(buf[slot]));
                buf[slot] = nullptr;
                cap[slot] = 0;
            }
            size_t want = bytes + bytes / 8;  // a little slack so slowly growing sizes do not thrash
            cudaError_t e = cudaMalloc(&buf[slot], want);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                want = bytes;
                e = cudaMalloc(&buf[slot], want);
            }
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                return fail(RRB_ENOMEM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
            }
            cap[slot] = want;
        }
        *out = buf[slot];
        return RRB_OK;
    }
    void release() {
        for (int i = 0; i < B_COUNT; ++i) {
            if (buf[i]) cudaFree(buf[i]);
            buf[i] = nullptr;
            cap[i] = 0;
        }
        for (int i = 0; i < 2; ++i) {
            if (ev_done[i]) cudaEventDestroy(ev_done[i]);
            if (ev_free[i]) cudaEventDestroy(ev_free[i]);
            ev_done[i] = ev_free[i] = nullptr;
        }
        if (ev_scratch) cudaEventDestroy(ev_scratch);
        ev_scratch = nullptr;
        scratch_busy = false;
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
        compute = copy = nullptr;
    }
};

static std::mutex g_ctx_mu;
static std::map<int, std::unique_ptr<Ctx>> g_ctx;

static int get_ctx(int device, Ctx** out) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(RRB_ECUDA, "no CUDA device available (%s); rrmpg_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) RRB_CUDA(cudaGetDevice(&device));
    if (device >= ndev) return fail(RRB_EINVAL, "device %d out of range (%d visible)", device, ndev);
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it == g_ctx.end()) {
        RRB_CUDA(cudaSetDevice(device));
        auto c = std::make_unique<Ctx>();
        c->device = device;
        RRB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
        RRB_CUDA(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        RRB_CUDA(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            RRB_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
            RRB_CUDA(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        }
        RRB_CUDA(cudaEventCreateWithFlags(&c->ev_scratch, cudaEventDisableTiming));
        it = g_ctx.emplace(device, std::move(c)).first;
    }
    *out = it->second.get();
    return RRB_OK;
}

struct Opts {
    int device = -1, mem = RRB_MEM_HOST, math = RRB_MATH_FAST, block = 0, variant = 0;
    cudaStream_t stream = nullptr;
    double x4_max = 0.0;
    const double* qobs = nullptr;
    double* mse = nullptr;
    int64_t slab_steps = 0;
    int objective = RRB_OBJ_MSE;
    int n_devices = 0;
    const int32_t* devices = nullptr;
    const double* obs_stats = nullptr;  // host [C][2]
    const double* state_in = nullptr;
    double* state_out = nullptr;
    int64_t out_row_pitch = 0;
};

static int parse_opts(const rrb_opts* o, Opts* out) {
    if (!o) return RRB_OK;
    if (o->struct_size != (int32_t)sizeof(rrb_opts))
        return fail(RRB_EINVAL, "rrb_opts.struct_size = %d, expected %zu", o->struct_size, sizeof(rrb_opts));
    if (o->mem != RRB_MEM_HOST && o->mem != RRB_MEM_DEVICE) return fail(RRB_EINVAL, "rrb_opts.mem = %d", o->mem);
    if (o->math != RRB_MATH_FAST && o->math != RRB_MATH_PRECISE) return fail(RRB_EINVAL, "rrb_opts.math = %d", o->math);
    if (o->block != 0 && (o->block < 32 || o->block > 1024 || (o->block % 32)))
        return fail(RRB_EINVAL, "rrb_opts.block = %d (must be a multiple of 32 in [32, 1024])", o->block);
    if (o->qobs && !o->mse) return fail(RRB_EINVAL, "rrb_opts.qobs given without rrb_opts.mse");
    if (o->slab_steps < 0) return fail(RRB_EINVAL, "rrb_opts.slab_steps < 0");
    out->device = o->device;
    out->mem = o->mem;
    out->math = o->math;
    out->block = o->block;
    out->variant = o->variant;
    out->stream = (cudaStream_t)o->stream;
    out->x4_max = o->x4_max;
    out->qobs = o->qobs;
    out->mse = o->mse;
    out->slab_steps = o->slab_steps;
    if (o->objective != RRB_OBJ_MSE && o->objective != RRB_OBJ_NSE && o->objective != RRB_OBJ_KGE)
        return fail(RRB_EINVAL, "rrb_opts.objective = %d", o->objective);
    if (o->qobs && o->objective != RRB_OBJ_MSE && !o->obs_stats)
        return fail(RRB_EINVAL, "rrb_opts.obs_stats (mean and standard deviation of qobs) is required for NSE / KGE");
    if (o->n_devices < 0) return fail(RRB_EINVAL, "rrb_opts.n_devices = %d", o->n_devices);
    if (o->n_devices > 1 && o->mem != RRB_MEM_HOST)
        return fail(RRB_EINVAL, "rrb_opts.n_devices > 1 needs RRB_MEM_HOST (device buffers live on one device)");
    if (o->out_row_pitch < 0) return fail(RRB_EINVAL, "rrb_opts.out_row_pitch < 0");
    if (o->out_row_pitch > 0 && o->mem != RRB_MEM_HOST) return fail(RRB_EINVAL, "rrb_opts.out_row_pitch needs RRB_MEM_HOST");
    out->objective = o->objective;
    out->n_devices = o->n_devices;
    out->devices = o->devices;
    out->obs_stats = o->obs_stats;
    out->state_in = o->state_in;
    out->state_out = o->state_out;
    out->out_row_pitch = o->out_row_pitch;
    return RRB_OK;
}

// ----------------------------------------------------------------------------------------
// generic driver.  A Model supplies: how to upload + pack its forcing, its carry-state size,
// its outputs, and how to launch one slab.
// ----------------------------------------------------------------------------------------
struct OutArr {
    double* user;       // caller's buffer (host or device), nullable
    int64_t row_elems;  // doubles per timestep (N or L*N)
};

struct Job {
    int64_t T = 0, N = 0;
    std::vector<OutArr> outs;
    int state_slots = 0;   // rows of the carry buffer: the model's own + kObjSlots objective rows
    bool resumable = false;  // the entry point supports rrb_opts.state_in / state_out
    // launch(slab, out_ptrs (device, row0-relative), objective, cfg)
    std::function<cudaError_t(const Slab&, double* const*, const Objective&, const LaunchCfg&)> launch;
};

static constexpr size_t kSlabTargetBytes = size_t(128) << 20;
// device bytes of one chunk of catchments in the host-mode batch calls (two chunks are in flight); the environment
// variable is a test knob that forces many small chunks through the ring
static size_t multi_chunk_bytes() {
    const char* e = getenv("RRMPG_B200_MULTI_CHUNK_BYTES");
    const size_t v = e ? (size_t)strtoull(e, nullptr, 10) : 0;
    return v > 0 ? v : kSlabTargetBytes * 2;
}

static Objective make_objective(const Opts& o, const double* d_qobs, double* d_mse, int64_t T, int64_t catchment = 0) {
    Objective obj{d_qobs, d_mse, T, o.objective, 0.0, 1.0, nullptr};
    if (o.obs_stats) {
        obj.obs_mean = o.obs_stats[2 * catchment];
        obj.obs_std = o.obs_stats[2 * catchment + 1];
    }
    return obj;
}

// Runs job.launch over [0, T).  Device mode: one launch straight into the caller's arrays.
// Host mode: time slabs through a two-deep device ring; the D2H of slab k overlaps slab k+1.
// rrb_opts.state_in / state_out (resume): the model rows of the carry buffer are read before the first and handed
// out after the last slab; with state_in timestep 0 is an ordinary step of a continued series (Slab::resume).
static int run_job(Ctx& c, const Opts& o, const Job& job, const double* d_qobs, double* d_mse) {
    const int64_t T = job.T, N = job.N;
    const int nout = (int)job.outs.size();
    LaunchCfg cfg{};
    cfg.block = o.block;
    cfg.math = o.math;
    cfg.sm_count = c.sm_count;
    cfg.variant = o.variant;
    Objective obj = make_objective(o, d_qobs, d_mse, T);
    const bool resume = o.state_in != nullptr, want_state = o.state_out != nullptr;
    if ((resume || want_state) && !job.resumable)
        return fail(RRB_EUNSUPPORTED, "state_in / state_out are supported by the ABC, HBV-Edu and GR4J entry points only");
    const int model_rows = job.state_slots - kObjSlots;
    const size_t model_state_bytes = sizeof(double) * (size_t)model_rows * (size_t)N;

    if (o.mem == RRB_MEM_DEVICE) {
        cfg.stream = o.stream;
        std::vector<double*> ptrs(nout);
        for (int k = 0; k < nout; ++k) ptrs[k] = job.outs[k].user;
        double* state = nullptr;
        if (want_state) {
            state = o.state_out;
            if (resume && o.state_in != o.state_out)
                RRB_CUDA(cudaMemcpyAsync(o.state_out, o.state_in, model_state_bytes, cudaMemcpyDeviceToDevice, o.stream));
        } else if (resume) {
            state = const_cast<double*>(o.state_in);  // read only: save_state = 0
        }
        Slab slab{0, T, 0, state, want_state ? 2 : 0, resume ? 1 : 0};
        RRB_CUDA(job.launch(slab, ptrs.data(), obj, cfg));
        return RRB_OK;
    }

    cfg.stream = c.compute;
    const int64_t pitch = o.out_row_pitch > 0 ? o.out_row_pitch : N;  // elements between rows of the caller's arrays
    if (pitch < N) return fail(RRB_EINVAL, "rrb_opts.out_row_pitch = %lld < N = %lld", (long long)pitch, (long long)N);
    int64_t bytes_per_step = 0;
    for (auto& a : job.outs)
        if (a.user) bytes_per_step += a.row_elems * (int64_t)sizeof(double);
    int64_t steps = T;
    if (bytes_per_step > 0) {
        steps = o.slab_steps > 0 ? o.slab_steps : (int64_t)(kSlabTargetBytes / (size_t)bytes_per_step);
        if (o.slab_steps == 0 && steps >= 64) steps &= ~int64_t(63);
        steps = std::max<int64_t>(1, std::min(steps, T));
        // two slabs of (almost) everything gain nothing over one
        if (o.slab_steps == 0 && steps * 2 > T) steps = T;
    }
    const int nslabs = (int)((T + steps - 1) / steps);
    const bool ring = nslabs > 1;

    double* state = nullptr;
    if (ring || resume || want_state) {
        void* p = nullptr;
        int rc = c.ensure(B_STATE, sizeof(double) * (size_t)job.state_slots * (size_t)N, &p);
        if (rc) return rc;
        state = (double*)p;
        if (resume) RRB_CUDA(cudaMemcpyAsync(state, o.state_in, model_state_bytes, cudaMemcpyHostToDevice, c.compute));
    }
    std::vector<double*> dev[2];
    for (int s = 0; s < (ring ? 2 : 1); ++s) {
        dev[s].assign(nout, nullptr);
        for (int k = 0; k < nout; ++k) {
            if (!job.outs[k].user) continue;
            void* p = nullptr;
            int rc = c.ensure(B_OUT0 + 2 * k + s, sizeof(double) * (size_t)steps * (size_t)job.outs[k].row_elems, &p);
            if (rc) return rc;
            dev[s][k] = (double*)p;
        }
    }
    for (int k = 0; k < nslabs; ++k) {
        const int s = k & 1;
        const int64_t t0 = (int64_t)k * steps, t1 = std::min(T, t0 + steps);
        if (k >= 2) RRB_CUDA(cudaStreamWaitEvent(c.compute, c.ev_free[s], 0));
        const bool last = k == nslabs - 1;
        Slab slab{t0, t1, t0, state, (ring && !last) || (last && want_state) ? 1 : 0, (resume && k == 0) ? 1 : 0};
        RRB_CUDA(job.launch(slab, dev[s].data(), obj, cfg));
        RRB_CUDA(cudaEventRecord(c.ev_done[s], c.compute));
        RRB_CUDA(cudaStreamWaitEvent(c.copy, c.ev_done[s], 0));
        for (int j = 0; j < nout; ++j) {
            if (!job.outs[j].user) continue;
            const size_t row = (size_t)job.outs[j].row_elems;
            if (pitch == N) {
                RRB_CUDA(cudaMemcpyAsync(job.outs[j].user + (size_t)t0 * row, dev[s][j],
                                         sizeof(double) * (size_t)(t1 - t0) * row, cudaMemcpyDeviceToHost, c.copy));
            } else {
                // the caller's rows are `pitch` elements apart (a member block of a wider [T, N_total] array): every
                // [N] row -- [T, L, N] storages hold L of them per timestep -- lands at its own pitch
                const size_t rows_per_step = row / (size_t)N;
                RRB_CUDA(cudaMemcpy2DAsync(job.outs[j].user + (size_t)t0 * rows_per_step * (size_t)pitch,
                                           sizeof(double) * (size_t)pitch, dev[s][j], sizeof(double) * (size_t)N,
                                           sizeof(double) * (size_t)N, (size_t)(t1 - t0) * rows_per_step,
                                           cudaMemcpyDeviceToHost, c.copy));
            }
        }
        RRB_CUDA(cudaEventRecord(c.ev_free[s], c.copy));
    }
    if (o.qobs) {
        RRB_CUDA(cudaStreamWaitEvent(c.copy, c.ev_done[(nslabs - 1) & 1], 0));
        RRB_CUDA(cudaMemcpyAsync(o.mse, d_mse, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, c.copy));
    }
    if (want_state) {
        RRB_CUDA(cudaStreamWaitEvent(c.copy, c.ev_done[(nslabs - 1) & 1], 0));
        RRB_CUDA(cudaMemcpyAsync(o.state_out, state, model_state_bytes, cudaMemcpyDeviceToHost, c.copy));
    }
    RRB_CUDA(cudaStreamSynchronize(c.copy));
    RRB_CUDA(cudaStreamSynchronize(c.compute));
    return RRB_OK;
}

// upload helper (host mode) / pass-through (device mode)
template <class T_>
static int stage_in(Ctx& c, const Opts& o, int slot, const T_* src, size_t count, const T_** out) {
    if (o.mem == RRB_MEM_DEVICE || src == nullptr) {
        *out = src;
        return RRB_OK;
    }
    void* p = nullptr;
    int rc = c.ensure(slot, sizeof(T_) * count, &p);
    if (rc) return rc;
    RRB_CUDA(cudaMemcpyAsync(p, src, sizeof(T_) * count, cudaMemcpyHostToDevice, c.compute));
    *out = (const T_*)p;
    return RRB_OK;
}

struct Prepared {
    Ctx* c = nullptr;
    Opts o;
    cudaStream_t s = nullptr;  // stream the prologue (uploads, packing) runs on
    const double* d_params = nullptr;
    const double* d_qobs = nullptr;
    double* d_mse = nullptr;
    const double* d_obs_stats = nullptr;  // catchment batches: device copy of rrb_opts.obs_stats [C][2]
};

// Guard of one entry point, taken right after the context mutex.
//  * Host mode hands the caller's buffers to asynchronous copies.  Whatever path leaves an entry point (an error return
//    in the middle of the prologue or of the slab ring included), nothing may still be reading or writing caller memory:
//    the guard drains both context streams on scope exit (a no-op after a successful run_job, which has synchronised).
//  * The per-context scratch buffers (packed forcing, G_tresh, scalar inits, flags) are written and read asynchronously
//    on the stream of the call.  A device-mode call returns while its kernels may still be reading them, and the next
//    call -- on another user stream, or a host-mode call on the context's own stream -- would re-pack into the same
//    buffers.  Every call therefore first makes its stream wait for the event the previous device-mode call recorded
//    behind its last kernel, and a device-mode call records that event on exit.
struct HostDrain {
    const Prepared& P;
    explicit HostDrain(const Prepared& p) : P(p) {
        if (P.c && P.c->scratch_busy) (void)cudaStreamWaitEvent(P.s, P.c->ev_scratch, 0);
    }
    ~HostDrain() {
        if (!P.c) return;
        if (P.o.mem == RRB_MEM_HOST) {
            (void)cudaStreamSynchronize(P.c->compute);
            (void)cudaStreamSynchronize(P.c->copy);
            P.c->scratch_busy = false;  // the wait above and this drain: every earlier user has finished
        } else {
            P.c->scratch_busy = cudaEventRecord(P.c->ev_scratch, P.s) == cudaSuccess;
        }
    }
};

static int prepare(const rrb_opts* opts, int64_t T, int64_t N, const double* params, int64_t pwidth, Prepared* P) {
    int rc = parse_opts(opts, &P->o);
    if (rc) return rc;
    if (T < 1) return fail(RRB_EINVAL, "T = %lld (need at least one timestep)", (long long)T);
    if (N < 0) return fail(RRB_EINVAL, "N = %lld", (long long)N);
    if (N > 0 && !params) return fail(RRB_EINVAL, "params is NULL");
    rc = get_ctx(P->o.device, &P->c);
    if (rc) return rc;
    RRB_CUDA(cudaSetDevice(P->c->device));
    P->s = (P->o.mem == RRB_MEM_DEVICE) ? P->o.stream : P->c->compute;
    return RRB_OK;
}

static int stage_common(Prepared* P, int64_t T, int64_t N, const double* params, int64_t pwidth) {
    int rc = stage_in(*P->c, P->o, B_PARAMS, params, (size_t)(N * pwidth), &P->d_params);
    if (rc) return rc;
    if (P->o.qobs) {
        rc = stage_in(*P->c, P->o, B_QOBS, P->o.qobs, (size_t)T, &P->d_qobs);
        if (rc) return rc;
        if (P->o.mem == RRB_MEM_DEVICE) {
            P->d_mse = P->o.mse;
        } else {
            void* p = nullptr;
            rc = P->c->ensure(B_MSE, sizeof(double) * (size_t)std::max<int64_t>(N, 1), &p);
            if (rc) return rc;
            P->d_mse = (double*)p;
        }
    }
    return RRB_OK;
}

static int resolve_x4_max(Prepared* P, const double* host_or_dev_params, int64_t N, int64_t stride, int field,
                          double* out) {
    if (P->o.mem == RRB_MEM_HOST) {
        double m = -1.0;
        bool nan = false;
        for (int64_t i = 0; i < N; ++i) {
            const double v = host_or_dev_params[i * stride + field];
            if (v != v) nan = true;
            if (v > m) m = v;
        }
        *out = nan ? NAN : m;
        return RRB_OK;
    }
    if (P->o.x4_max > 0) {
        *out = P->o.x4_max;
        return RRB_OK;
    }
    void* d = nullptr;
    int rc = P->c->ensure(B_SCALAR, sizeof(double), &d);
    if (rc) return rc;
    max_field_kernel<<<1, 1024, 0, P->s>>>(P->d_params, N, stride, field, (double*)d);
    RRB_CUDA(cudaGetLastError());
    RRB_CUDA(cudaMemcpyAsync(out, d, sizeof(double), cudaMemcpyDeviceToHost, P->s));
    RRB_CUDA(cudaStreamSynchronize(P->s));
    return RRB_OK;
}

}  // namespace rrb

using namespace rrb;

// ------------------------------------------------------------------------------------------------
// Catchment batches (one launch, blockIdx.y = catchment) of the GR4J family.  Device mode: one launch over all
// catchments into the caller's buffers.  Host mode: chunks of catchments through a two-deep device ring, the D2H
// of chunk k overlapping the kernel of chunk k+1 (each chunk is a contiguous block of every [C, ...] output).
// ------------------------------------------------------------------------------------------------
struct MultiOut {
    double* ptr;             // caller's array (host or device per opts->mem), nullable
    int64_t per_catchment;   // elements per catchment
};
template <class Launch>
static int run_multi(Prepared& P, int64_t C, int64_t T, int64_t N, const std::vector<MultiOut>& outs, Launch&& launch) {
    Ctx& c = *P.c;
    const bool host = P.o.mem == RRB_MEM_HOST;
    LaunchCfg cfg{};
    cfg.block = P.o.block;
    cfg.math = P.o.math;
    cfg.sm_count = c.sm_count;
    cfg.variant = P.o.variant;
    cfg.stream = P.s;
    const int no = (int)outs.size();
    std::vector<double*> ptrs(no);
    if (!host) {
        for (int j = 0; j < no; ++j) ptrs[j] = outs[j].ptr;
        Objective obj = make_objective(P.o, P.d_qobs, P.d_mse, T);
        obj.obs_stats = P.d_obs_stats;
        RRB_CUDA(launch(0, C, ptrs.data(), obj, cfg));
        return RRB_OK;
    }
    int64_t bytes_per_catchment = 0;
    for (const MultiOut& o : outs)
        if (o.ptr) bytes_per_catchment += o.per_catchment * (int64_t)sizeof(double);
    int64_t cc = C;
    if (bytes_per_catchment > 0)
        cc = std::max<int64_t>(1, std::min<int64_t>(C, (int64_t)multi_chunk_bytes() / bytes_per_catchment));
    const int nchunks = (int)((C + cc - 1) / cc);
    std::vector<double*> dev[2] = {std::vector<double*>(no, nullptr), std::vector<double*>(no, nullptr)};
    int rc;
    for (int sidx = 0; sidx < (nchunks > 1 ? 2 : 1); ++sidx)
        for (int j = 0; j < no; ++j)
            if (outs[j].ptr) {
                void* p = nullptr;
                if ((rc = c.ensure(B_OUT0 + 2 * j + sidx, sizeof(double) * (size_t)(cc * outs[j].per_catchment), &p))) return rc;
                dev[sidx][j] = (double*)p;
            }
    for (int k = 0; k < nchunks; ++k) {
        const int sidx = k & 1;
        const int64_t c0 = (int64_t)k * cc, c1 = std::min(C, c0 + cc);
        if (k >= 2) RRB_CUDA(cudaStreamWaitEvent(c.compute, c.ev_free[sidx], 0));
        Objective obj = make_objective(P.o, P.d_qobs ? P.d_qobs + c0 * T : nullptr, P.d_mse ? P.d_mse + c0 * N : nullptr, T, c0);
        obj.obs_stats = P.d_obs_stats ? P.d_obs_stats + 2 * c0 : nullptr;
        RRB_CUDA(launch(c0, c1, dev[sidx].data(), obj, cfg));
        RRB_CUDA(cudaEventRecord(c.ev_done[sidx], c.compute));
        RRB_CUDA(cudaStreamWaitEvent(c.copy, c.ev_done[sidx], 0));
        for (int j = 0; j < no; ++j)
            if (outs[j].ptr)
                RRB_CUDA(cudaMemcpyAsync(outs[j].ptr + (size_t)(c0 * outs[j].per_catchment), dev[sidx][j],
                                         sizeof(double) * (size_t)((c1 - c0) * outs[j].per_catchment),
                                         cudaMemcpyDeviceToHost, c.copy));
        RRB_CUDA(cudaEventRecord(c.ev_free[sidx], c.copy));
    }
    if (P.o.qobs) {
        RRB_CUDA(cudaStreamWaitEvent(c.copy, c.ev_done[(nchunks - 1) & 1], 0));
        RRB_CUDA(cudaMemcpyAsync(P.o.mse, P.d_mse, sizeof(double) * (size_t)(C * N), cudaMemcpyDeviceToHost, c.copy));
    }
    RRB_CUDA(cudaStreamSynchronize(c.copy));
    RRB_CUDA(cudaStreamSynchronize(c.compute));
    return RRB_OK;
}

// per-catchment (mean, std) of the observations (host memory in both modes) -> device array read by the kernels
static int stage_obs_stats(Prepared& P, int64_t C) {
    if (!P.o.obs_stats) return RRB_OK;
    void* p = nullptr;
    int rc = P.c->ensure(B_OBS, sizeof(double) * 2 * (size_t)C, &p);
    if (rc) return rc;
    RRB_CUDA(cudaMemcpyAsync(p, P.o.obs_stats, sizeof(double) * 2 * (size_t)C, cudaMemcpyHostToDevice, P.s));
    P.d_obs_stats = (const double*)p;
    return RRB_OK;
}

// shared staging of a catchment batch: parameters [C][N][pwidth], per-catchment inits [C][4] (host memory in both
// modes), observations [C][T] and the per-member objective [C][N]
static int stage_multi(Prepared& P, int64_t C, int64_t T, int64_t N, const double* inits4, const double** d_inits,
                       int width = 4) {
    Ctx& c = *P.c;
    int rc;
    void* di = nullptr;
    if ((rc = c.ensure(B_SCALAR, sizeof(double) * (size_t)width * (size_t)C, &di))) return rc;
    RRB_CUDA(cudaMemcpyAsync(di, inits4, sizeof(double) * (size_t)width * (size_t)C, cudaMemcpyHostToDevice, P.s));
    *d_inits = (const double*)di;
    if (P.o.qobs) {
        if ((rc = stage_in(c, P.o, B_QOBS, P.o.qobs, (size_t)(C * T), &P.d_qobs))) return rc;
        if (P.o.mem == RRB_MEM_HOST) {
            void* p = nullptr;
            if ((rc = c.ensure(B_MSE, sizeof(double) * (size_t)(C * N), &p))) return rc;
            P.d_mse = (double*)p;
        } else {
            P.d_mse = P.o.mse;
        }
        if ((rc = stage_obs_stats(P, C))) return rc;
    }
    return RRB_OK;
}



// ----------------------------------------------------------------------------------------
// In-library multi-GPU (SURVEY.md section 8b "Threading" / 8e "Process model"): a host-mode single-catchment call
// with rrb_opts.n_devices > 1 splits the ensemble [0, N) into contiguous member blocks, one per device, and runs the
// same entry point for every block on its own worker thread (own context, streams and scratch per device).  The
// member-independent forcing is uploaded to every device (<= 0.5 MB); every block's rows are copied straight into its
// column block of the caller's [T, N] / [T, L, N] arrays (rrb_opts.out_row_pitch = N).  No collective: members are
// independent given the forcing (the loop being sharded is rrmpg/models/hbvedu.py:199-209 and siblings).
// ----------------------------------------------------------------------------------------
static bool wants_sharding(const rrb_opts* o, int64_t N) {
    return o && o->struct_size == (int32_t)sizeof(rrb_opts) && o->n_devices > 1 && o->mem == RRB_MEM_HOST && N > 0;
}
template <class T_>
static T_* shifted(T_* p, int64_t elems) { return p ? p + elems : nullptr; }

// call(sub_opts, lo, n): the entry point for members [lo, lo + n)
template <class Call>
static int shard_members(const rrb_opts* opts, int64_t N, Call&& call) {
    if (opts->state_in || opts->state_out)
        return fail(RRB_EUNSUPPORTED, "state_in / state_out are not supported together with n_devices > 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(RRB_ECUDA, "no CUDA device available; rrmpg_b200 has no CPU fallback");
    }
    std::vector<int> devs;
    for (int k = 0; k < opts->n_devices; ++k) {
        const int d = opts->devices ? opts->devices[k] : k;
        if (d < 0 || d >= ndev) return fail(RRB_EINVAL, "rrb_opts.devices[%d] = %d out of range (%d visible)", k, d, ndev);
        devs.push_back(d);  // (a device listed twice gets two blocks, serialised by its context mutex: legal, used by tests)
    }
    // contiguous blocks, multiples of 64 members (16-byte aligned rows, whole warps) except the last
    const int64_t D = (int64_t)devs.size();
    int64_t per = (N + D - 1) / D;
    per = (per + 63) & ~int64_t(63);
    struct Part { int64_t lo, n; int rc; std::string err; };
    std::vector<Part> parts;
    for (int64_t lo = 0; lo < N; lo += per) parts.push_back(Part{lo, std::min(per, N - lo), RRB_OK, ""});
    const int64_t pitch = opts->out_row_pitch > 0 ? opts->out_row_pitch : N;
    std::vector<std::thread> workers;
    for (size_t k = 0; k < parts.size(); ++k) {
        workers.emplace_back([&, k]() {
            rrb_opts so = *opts;
            so.n_devices = 0;
            so.devices = nullptr;
            so.device = devs[k];
            so.out_row_pitch = pitch;
            so.mse = shifted(opts->mse, parts[k].lo);
            parts[k].rc = call(&so, parts[k].lo, parts[k].n);
            if (parts[k].rc != RRB_OK) parts[k].err = g_err;  // thread-local: carry the message over
        });
    }
    for (auto& w : workers) w.join();
    for (auto& pt : parts)
        if (pt.rc != RRB_OK) {
            g_err = pt.err;
            return pt.rc;
        }
    return RRB_OK;
}

// ----------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------
extern "C" {

int rrb_version(void) { return RRB_VERSION; }

int rrb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int rrb_init(int device) {
    Ctx* c;
    return get_ctx(device, &c);
}

int rrb_init_devices(const int32_t* devices, int n) {
    if (n < 1) return fail(RRB_EINVAL, "rrb_init_devices: n = %d", n);
    for (int k = 0; k < n; ++k) {
        Ctx* c;
        int rc = get_ctx(devices ? devices[k] : k, &c);
        if (rc) return rc;
    }
    return RRB_OK;
}

int rrb_state_rows(int model, double x4_max) {
    switch (model) {
        case RRB_MODEL_ABC: return state_slots_abc() - kObjSlots;
        case RRB_MODEL_HBVEDU: return state_slots_hbvedu() - kObjSlots;
        case RRB_MODEL_GR4J: return (x4_max <= RRB_MAX_X4) ? state_slots_gr4j(x4_max) - kObjSlots : -1;
        default: return -1;
    }
}

int rrb_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (auto& kv : g_ctx) {
        cudaSetDevice(kv.first);
        cudaDeviceSynchronize();
        kv.second->release();
    }
    g_ctx.clear();
    rrb_host_pool_trim();
    return RRB_OK;
}

const char* rrb_last_error(void) { return g_err.c_str(); }

int rrb_synchronize(int device) {
    Ctx* c;
    int rc = get_ctx(device, &c);
    if (rc) return rc;
    RRB_CUDA(cudaSetDevice(c->device));
    RRB_CUDA(cudaDeviceSynchronize());
    return RRB_OK;
}

// ---- pinned host pool ----
static std::mutex g_pin_mu;
static std::multimap<size_t, void*> g_pin_free;
static std::map<void*, size_t> g_pin_live;

void* rrb_host_alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pin_free.lower_bound(bytes);
    if (it != g_pin_free.end() && it->first <= bytes + bytes / 4 + 4096) {
        void* p = it->second;
        g_pin_live[p] = it->first;
        g_pin_free.erase(it);
        return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        // make room and retry once
        for (auto& kv : g_pin_free) cudaFreeHost(kv.second);
        g_pin_free.clear();
        e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            fail(RRB_ENOMEM, "cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return nullptr;
        }
    }
    g_pin_live[p] = bytes;
    return p;
}

void rrb_host_free(void* ptr) {
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pin_live.find(ptr);
    if (it == g_pin_live.end()) return;
    g_pin_free.emplace(it->second, ptr);
    g_pin_live.erase(it);
}

void rrb_host_pool_trim(void) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (auto& kv : g_pin_free) cudaFreeHost(kv.second);
    g_pin_free.clear();
}

// ---- ABC ----
int rrb_abc_simulate(const double* prec, int64_t T, double initial_state, const double* params, int64_t N,
                     double* qsim, double* storage, const rrb_opts* opts) {
    if (wants_sharding(opts, N))
        return shard_members(opts, N, [&](const rrb_opts* so, int64_t lo, int64_t n) {
            return rrb_abc_simulate(prec, T, initial_state, params + lo * 3, n, shifted(qsim, lo), shifted(storage, lo), so);
        });
    Prepared P;
    int rc = prepare(opts, T, N, params, 3, &P);
    if (rc) return rc;
    if (!prec) return fail(RRB_EINVAL, "prec is NULL");
    if (!qsim && !P.o.qobs) return fail(RRB_EINVAL, "nothing to compute: qsim is NULL and no objective requested");
    if (N == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double* d_prec;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)T, &d_prec))) return rc;
    if ((rc = stage_common(&P, T, N, params, 3))) return rc;
    void* F = nullptr;
    if ((rc = P.c->ensure(B_F, sizeof(double) * (size_t)padded_steps(T, kAbcTT) * kAbcR, &F))) return rc;
    RRB_CUDA(pack_abc(d_prec, T, (double*)F, P.s));
    Job job;
    job.T = T; job.N = N;
    job.outs = {{qsim, N}, {storage, N}};
    job.state_slots = state_slots_abc();
    job.resumable = true;
    const double* dp = P.d_params;
    job.launch = [=](const Slab& sl, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        return launch_abc((const double*)F, T, initial_state, dp, N, out[0], out[1], sl, ob, cfg);
    };
    return run_job(*P.c, P.o, job, P.d_qobs, P.d_mse);
}

// ---- HBV-Edu ----
int rrb_hbvedu_simulate(const double* temp, const double* prec, const int8_t* month0, const double* PE_m,
                        const double* T_m, int64_t T, const double* inits, const double* params, int64_t N,
                        double* qsim, double* snow, double* soil, double* s1, double* s2, const rrb_opts* opts) {
    if (wants_sharding(opts, N))
        return shard_members(opts, N, [&](const rrb_opts* so, int64_t lo, int64_t n) {
            return rrb_hbvedu_simulate(temp, prec, month0, PE_m, T_m, T, inits, params + lo * 11, n, shifted(qsim, lo),
                                       shifted(snow, lo), shifted(soil, lo), shifted(s1, lo), shifted(s2, lo), so);
        });
    Prepared P;
    int rc = prepare(opts, T, N, params, 11, &P);
    if (rc) return rc;
    if (!temp || !prec || !month0 || !PE_m || !T_m || !inits) return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    const int nst = (snow != nullptr) + (soil != nullptr) + (s1 != nullptr) + (s2 != nullptr);
    if (nst != 0 && nst != 4) return fail(RRB_EINVAL, "pass all four storage outputs or none");
    if (!qsim && !P.o.qobs && nst == 0) return fail(RRB_EINVAL, "nothing to compute");
    if (N == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_temp, *d_prec, *d_pe, *d_tm;
    const int8_t* d_month;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, temp, (size_t)T, &d_temp))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, prec, (size_t)T, &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, month0, (size_t)T, &d_month))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW3, PE_m, (size_t)12, &d_pe))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW4, T_m, (size_t)12, &d_tm))) return rc;
    if ((rc = stage_common(&P, T, N, params, 11))) return rc;
    double in4[4];
    memcpy(in4, inits, sizeof(in4));  // inits is host memory in both modes
    void* F = nullptr;
    if ((rc = P.c->ensure(B_F, forcing_bytes(T, kHbvTT, kHbvR) + hbv_scratch_bytes(N, 1), &F))) return rc;
    RRB_CUDA(pack_hbvedu(d_temp, d_prec, d_month, d_pe, d_tm, T, (double*)F, P.o.math, 1, P.s));
    Job job;
    job.T = T; job.N = N;
    job.outs = {{qsim, N}, {snow, N}, {soil, N}, {s1, N}, {s2, N}};
    job.state_slots = state_slots_hbvedu();
    job.resumable = true;
    const double* dp = P.d_params;
    const double i0 = in4[0], i1 = in4[1], i2 = in4[2], i3 = in4[3];
    job.launch = [=](const Slab& sl, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        const double in[4] = {i0, i1, i2, i3};
        return launch_hbvedu((const double*)F, T, in, dp, N, out[0], out[1], out[2], out[3], out[4], sl, ob, cfg,
                             forcing_flag((const double*)F, T, kHbvTT, kHbvR));
    };
    return run_job(*P.c, P.o, job, P.d_qobs, P.d_mse);
}

// ---- HBV-Edu, C catchments in one launch ----
int rrb_hbvedu_simulate_multi(const double* temp, const double* prec, const int8_t* month0, const double* PE_m,
                              const double* T_m, int64_t C, int64_t T, const double* inits, const double* params,
                              int64_t N, double* qsim, double* snow, double* soil, double* s1, double* s2,
                              const rrb_opts* opts) {
    Prepared P;
    int rc = prepare(opts, T, N, params, 11, &P);
    if (rc) return rc;
    if (C < 0) return fail(RRB_EINVAL, "C = %lld", (long long)C);
    if (!temp || !prec || !month0 || !PE_m || !T_m || !inits) return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    const int nst = (snow != nullptr) + (soil != nullptr) + (s1 != nullptr) + (s2 != nullptr);
    if (nst != 0 && nst != 4) return fail(RRB_EINVAL, "pass all four storage outputs or none");
    if (!qsim && !P.o.qobs && nst == 0) return fail(RRB_EINVAL, "nothing to compute");
    if (C > 65535) return fail(RRB_EUNSUPPORTED, "C = %lld catchments per call (max 65535)", (long long)C);
    if (N == 0 || C == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    Ctx& c = *P.c;
    const bool host = P.o.mem == RRB_MEM_HOST;
    const double *d_temp, *d_prec, *d_pe, *d_tm;
    const int8_t* d_month;
    if ((rc = stage_in(c, P.o, B_RAW0, temp, (size_t)(C * T), &d_temp))) return rc;
    if ((rc = stage_in(c, P.o, B_RAW1, prec, (size_t)(C * T), &d_prec))) return rc;
    if ((rc = stage_in(c, P.o, B_RAW2, month0, (size_t)(C * T), &d_month))) return rc;
    if ((rc = stage_in(c, P.o, B_RAW3, PE_m, (size_t)(C * 12), &d_pe))) return rc;
    if ((rc = stage_in(c, P.o, B_RAW4, T_m, (size_t)(C * 12), &d_tm))) return rc;
    if ((rc = stage_in(c, P.o, B_PARAMS, params, (size_t)(C * N * 11), &P.d_params))) return rc;
    // per-catchment initial states: always host memory -> small device array
    void* d_inits;
    if ((rc = c.ensure(B_SCALAR, sizeof(double) * 4 * (size_t)C, &d_inits))) return rc;
    RRB_CUDA(cudaMemcpyAsync(d_inits, inits, sizeof(double) * 4 * (size_t)C, cudaMemcpyHostToDevice, P.s));
    if (P.o.qobs) {
        if ((rc = stage_in(c, P.o, B_QOBS, P.o.qobs, (size_t)(C * T), &P.d_qobs))) return rc;
        if (host) {
            void* p = nullptr;
            if ((rc = c.ensure(B_MSE, sizeof(double) * (size_t)(C * N), &p))) return rc;
            P.d_mse = (double*)p;
        } else {
            P.d_mse = P.o.mse;
        }
        if ((rc = stage_obs_stats(P, C))) return rc;
    }
    const int64_t Tpad = padded_steps(T, kHbvTT);
    void* F = nullptr;
    if ((rc = c.ensure(B_F, sizeof(double) * (size_t)(C * Tpad) * kHbvR + hbv_scratch_bytes(N, C), &F))) return rc;
    RRB_CUDA(pack_hbvedu(d_temp, d_prec, d_month, d_pe, d_tm, T, (double*)F, P.o.math, (int)C, P.s));
    const uint32_t* hbv_flag = reinterpret_cast<const uint32_t*>((const double*)F + C * Tpad * kHbvR);

    LaunchCfg cfg{};
    cfg.block = P.o.block;
    cfg.math = P.o.math;
    cfg.sm_count = c.sm_count;
    cfg.variant = P.o.variant;
    cfg.stream = P.s;
    const double zero4[4] = {0, 0, 0, 0};
    double* outs[5] = {qsim, snow, soil, s1, s2};

    if (!host) {
        Batch b{(int)C, Tpad * kHbvR, T * N, (const double*)d_inits};
        Slab slab{0, T, 0, nullptr, 0, 0};
        Objective obj = make_objective(P.o, P.d_qobs, P.d_mse, T);
        obj.obs_stats = P.d_obs_stats;
        RRB_CUDA(launch_hbvedu((const double*)F, T, zero4, P.d_params, N, qsim, snow, soil, s1, s2, slab, obj, cfg,
                               hbv_flag, b));
        return RRB_OK;
    }
    // host buffers: chunks of catchments through a two-deep device ring; the D2H of chunk k overlaps the
    // kernel of chunk k+1 (each chunk is a contiguous [cc, T, N] block of every output)
    int64_t bytes_per_catchment = 0;
    for (double* o : outs)
        if (o) bytes_per_catchment += T * N * (int64_t)sizeof(double);
    int64_t cc = C;
    if (bytes_per_catchment > 0) cc = std::max<int64_t>(1, std::min<int64_t>(C, (int64_t)multi_chunk_bytes() / bytes_per_catchment));
    const int nchunks = (int)((C + cc - 1) / cc);
    double* dev[2][5] = {};
    for (int sidx = 0; sidx < (nchunks > 1 ? 2 : 1); ++sidx)
        for (int k = 0; k < 5; ++k)
            if (outs[k]) {
                void* p = nullptr;
                if ((rc = c.ensure(B_OUT0 + 2 * k + sidx, sizeof(double) * (size_t)(cc * T * N), &p))) return rc;
                dev[sidx][k] = (double*)p;
            }
    for (int k = 0; k < nchunks; ++k) {
        const int sidx = k & 1;
        const int64_t c0 = (int64_t)k * cc, c1 = std::min(C, c0 + cc);
        if (k >= 2) RRB_CUDA(cudaStreamWaitEvent(c.compute, c.ev_free[sidx], 0));
        Batch b{(int)(c1 - c0), Tpad * kHbvR, T * N, (const double*)d_inits + 4 * c0};
        Slab slab{0, T, 0, nullptr, 0, 0};
        Objective obj = make_objective(P.o, P.d_qobs ? P.d_qobs + c0 * T : nullptr, P.d_mse ? P.d_mse + c0 * N : nullptr, T, c0);
        obj.obs_stats = P.d_obs_stats ? P.d_obs_stats + 2 * c0 : nullptr;
        RRB_CUDA(launch_hbvedu((const double*)F + c0 * Tpad * kHbvR, T, zero4, P.d_params + c0 * N * 11, N, dev[sidx][0],
                               dev[sidx][1], dev[sidx][2], dev[sidx][3], dev[sidx][4], slab, obj, cfg, hbv_flag, b));
        RRB_CUDA(cudaEventRecord(c.ev_done[sidx], c.compute));
        RRB_CUDA(cudaStreamWaitEvent(c.copy, c.ev_done[sidx], 0));
        for (int j = 0; j < 5; ++j)
            if (outs[j])
                RRB_CUDA(cudaMemcpyAsync(outs[j] + (size_t)(c0 * T * N), dev[sidx][j],
                                         sizeof(double) * (size_t)((c1 - c0) * T * N), cudaMemcpyDeviceToHost, c.copy));
        RRB_CUDA(cudaEventRecord(c.ev_free[sidx], c.copy));
    }
    if (P.o.qobs) {
        RRB_CUDA(cudaStreamWaitEvent(c.copy, c.ev_done[(nchunks - 1) & 1], 0));
        RRB_CUDA(cudaMemcpyAsync(P.o.mse, P.d_mse, sizeof(double) * (size_t)(C * N), cudaMemcpyDeviceToHost, c.copy));
    }
    RRB_CUDA(cudaStreamSynchronize(c.copy));
    RRB_CUDA(cudaStreamSynchronize(c.compute));
    return RRB_OK;
}

// ---- GR4J ----
int rrb_gr4j_simulate(const double* prec, const double* etp, int64_t T, double s_init, double r_init,
                      const double* params, int64_t N, double* qsim, double* s_store, double* r_store,
                      const rrb_opts* opts) {
    if (wants_sharding(opts, N))
        return shard_members(opts, N, [&](const rrb_opts* so, int64_t lo, int64_t n) {
            return rrb_gr4j_simulate(prec, etp, T, s_init, r_init, params + lo * 4, n, shifted(qsim, lo),
                                     shifted(s_store, lo), shifted(r_store, lo), so);
        });
    Prepared P;
    int rc = prepare(opts, T, N, params, 4, &P);
    if (rc) return rc;
    if (!prec || !etp) return fail(RRB_EINVAL, "NULL forcing pointer");
    if ((s_store != nullptr) != (r_store != nullptr)) return fail(RRB_EINVAL, "pass both storage outputs or none");
    if (!qsim && !P.o.qobs && !s_store) return fail(RRB_EINVAL, "nothing to compute");
    if (N == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_etp;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)T, &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, etp, (size_t)T, &d_etp))) return rc;
    if ((rc = stage_common(&P, T, N, params, 4))) return rc;
    double x4_max;
    if ((rc = resolve_x4_max(&P, params, N, 4, 3, &x4_max))) return rc;
    if (!(x4_max <= RRB_MAX_X4))
        return fail(RRB_EUNSUPPORTED, "GR4J x4 up to %g in this batch; the unit hydrograph buffers support x4 <= %g",
                    x4_max, RRB_MAX_X4);
    void* F = nullptr;
    if ((rc = P.c->ensure(B_F, forcing_bytes(T, kGr4jTT, kGr4jR), &F))) return rc;
    RRB_CUDA(pack_gr4j(d_prec, d_etp, T, (double*)F, P.s));
    Job job;
    job.T = T; job.N = N;
    job.outs = {{qsim, N}, {s_store, N}, {r_store, N}};
    job.state_slots = state_slots_gr4j(x4_max);
    job.resumable = true;
    const double* dp = P.d_params;
    job.launch = [=](const Slab& sl, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        return launch_gr4j((const double*)F, T, s_init, r_init, dp, N, x4_max, out[0], out[1], out[2], sl, ob, cfg);
    };
    return run_job(*P.c, P.o, job, P.d_qobs, P.d_mse);
}

// ---- Cemaneige ----
int rrb_cemaneige_simulate(const double* prec, const double* mean_temp, const double* frac_solid, int64_t T,
                           int64_t L, double snow_pack_init, double thermal_state_init, const double* params,
                           int64_t param_stride, int64_t N, double* outflow, double* G, double* eTG,
                           const rrb_opts* opts) {
    if (wants_sharding(opts, N))  // [T, L, N] storages: every [N] row of a member block starts lo elements into the row
        return shard_members(opts, N, [&](const rrb_opts* so, int64_t lo, int64_t n) {
            return rrb_cemaneige_simulate(prec, mean_temp, frac_solid, T, L, snow_pack_init, thermal_state_init,
                                          params + lo * param_stride, param_stride, n, shifted(outflow, lo), shifted(G, lo),
                                          shifted(eTG, lo), so);
        });
    Prepared P;
    int rc = prepare(opts, T, N, params, param_stride, &P);
    if (rc) return rc;
    if (!prec || !mean_temp || !frac_solid) return fail(RRB_EINVAL, "NULL forcing pointer");
    if (L < 1) return fail(RRB_EINVAL, "L = %lld", (long long)L);
    if (L > RRB_MAX_LAYERS) return fail(RRB_EUNSUPPORTED, "L = %lld elevation layers (max %d)", (long long)L, RRB_MAX_LAYERS);
    if (param_stride < 2) return fail(RRB_EINVAL, "param_stride = %lld (< 2)", (long long)param_stride);
    if ((G != nullptr) != (eTG != nullptr)) return fail(RRB_EINVAL, "pass both storage outputs or none");
    if (!outflow && !P.o.qobs && !G) return fail(RRB_EINVAL, "nothing to compute");
    if (N == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_mt, *d_fr;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(T * L), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, mean_temp, (size_t)(T * L), &d_mt))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, frac_solid, (size_t)(T * L), &d_fr))) return rc;
    if ((rc = stage_common(&P, T, N, params, param_stride))) return rc;
    const int LC = cema_layer_class((int)L);
    void *F = nullptr, *gt = nullptr;
    if ((rc = P.c->ensure(B_F, forcing_bytes(T, cema_TT(LC), cema_R(LC)), &F))) return rc;
    if ((rc = P.c->ensure(B_GT, sizeof(double) * 2 * kCemaMaxLayers, &gt))) return rc;
    RRB_CUDA(pack_cemaneige(d_prec, d_mt, d_fr, nullptr, T, (int)L, (double*)F, (double*)gt, P.s));
    Job job;
    job.T = T; job.N = N;
    job.outs = {{outflow, N}, {G, L * N}, {eTG, L * N}};
    job.state_slots = state_slots_cemaneige((int)L);
    const double* dp = P.d_params;
    job.launch = [=](const Slab& sl, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        return launch_cemaneige((const double*)F, (const double*)gt, T, (int)L, snow_pack_init, thermal_state_init, dp,
                                param_stride, N, out[0], out[1], out[2], sl, ob, cfg);
    };
    return run_job(*P.c, P.o, job, P.d_qobs, P.d_mse);
}

// ---- Cemaneige + GR4J ----
int rrb_cemaneigegr4j_simulate(const double* prec, const double* mean_temp, const double* etp,
                               const double* frac_solid, int64_t T, int64_t L, const double* inits,
                               const double* params, int64_t N, double* qsim, double* G, double* eTG,
                               double* s_store, double* r_store, const rrb_opts* opts) {
    if (wants_sharding(opts, N))
        return shard_members(opts, N, [&](const rrb_opts* so, int64_t lo, int64_t n) {
            return rrb_cemaneigegr4j_simulate(prec, mean_temp, etp, frac_solid, T, L, inits, params + lo * 6, n,
                                              shifted(qsim, lo), shifted(G, lo), shifted(eTG, lo), shifted(s_store, lo),
                                              shifted(r_store, lo), so);
        });
    Prepared P;
    int rc = prepare(opts, T, N, params, 6, &P);
    if (rc) return rc;
    if (!prec || !mean_temp || !etp || !frac_solid || !inits) return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    if (L < 1) return fail(RRB_EINVAL, "L = %lld", (long long)L);
    if (L > RRB_MAX_LAYERS) return fail(RRB_EUNSUPPORTED, "L = %lld elevation layers (max %d)", (long long)L, RRB_MAX_LAYERS);
    const int nst = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr);
    if (nst != 0 && nst != 4) return fail(RRB_EINVAL, "pass all four storage outputs or none");
    if (!qsim && !P.o.qobs && nst == 0) return fail(RRB_EINVAL, "nothing to compute");
    if (N == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_mt, *d_fr, *d_etp;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(T * L), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, mean_temp, (size_t)(T * L), &d_mt))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, frac_solid, (size_t)(T * L), &d_fr))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW3, etp, (size_t)T, &d_etp))) return rc;
    if ((rc = stage_common(&P, T, N, params, 6))) return rc;
    double in4[4];
    memcpy(in4, inits, sizeof(in4));  // inits is host memory in both modes
    double x4_max;
    if ((rc = resolve_x4_max(&P, params, N, 6, 5, &x4_max))) return rc;
    if (!(x4_max <= RRB_MAX_X4))
        return fail(RRB_EUNSUPPORTED, "GR4J x4 up to %g in this batch; the unit hydrograph buffers support x4 <= %g",
                    x4_max, RRB_MAX_X4);
    const int LC = cema_layer_class((int)L);
    void *F = nullptr, *gt = nullptr;
    if ((rc = P.c->ensure(B_F, forcing_bytes(T, cema_TT(LC), cema_R(LC)), &F))) return rc;
    if ((rc = P.c->ensure(B_GT, sizeof(double) * 2 * kCemaMaxLayers, &gt))) return rc;
    RRB_CUDA(pack_cemaneige(d_prec, d_mt, d_fr, d_etp, T, (int)L, (double*)F, (double*)gt, P.s));
    Job job;
    job.T = T; job.N = N;
    job.outs = {{qsim, N}, {G, L * N}, {eTG, L * N}, {s_store, N}, {r_store, N}};
    job.state_slots = state_slots_cemaneigegr4j((int)L, x4_max);
    const double* dp = P.d_params;
    const double i0 = in4[0], i1 = in4[1], i2 = in4[2], i3 = in4[3];
    job.launch = [=](const Slab& sl, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        const double in[4] = {i0, i1, i2, i3};
        return launch_cemaneigegr4j((const double*)F, (const double*)gt, T, (int)L, in, dp, N, x4_max, out[0], out[1],
                                    out[2], out[3], out[4], sl, ob, cfg);
    };
    return run_job(*P.c, P.o, job, P.d_qobs, P.d_mse);
}

// ---- snow-ice family ----
static int snowice_simulate(int family, const double* prec, const double* mean_temp, const double* etp,
                            const double* frac_ice, const double* frac_solid, int64_t T, int64_t L, const double* inits,
                            const double* params, int64_t N, const SnowIceOut& o, int n_storage_given,
                            int n_storage_expected, const rrb_opts* opts) {
    const bool hyst = family & 1, ice = family & 2;
    const int k = 6 + (hyst ? 2 : 0) + (ice ? 1 : 0);
    if (wants_sharding(opts, N))
        return shard_members(opts, N, [&](const rrb_opts* so, int64_t lo, int64_t n) {
            SnowIceOut so_out{shifted(o.qsim, lo), shifted(o.G, lo), shifted(o.eTG, lo), shifted(o.s_store, lo),
                              shifted(o.r_store, lo), shifted(o.sca, lo), shifted(o.icemelt, lo), shifted(o.snowmelt, lo)};
            return snowice_simulate(family, prec, mean_temp, etp, frac_ice, frac_solid, T, L, inits, params + lo * k, n,
                                    so_out, n_storage_given, n_storage_expected, so);
        });
    Prepared P;
    int rc = prepare(opts, T, N, params, k, &P);
    if (rc) return rc;
    if (!prec || !mean_temp || !etp || !frac_solid || !inits || (ice && !frac_ice))
        return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    if (L < 1) return fail(RRB_EINVAL, "L = %lld", (long long)L);
    if (L > RRB_MAX_LAYERS) return fail(RRB_EUNSUPPORTED, "L = %lld elevation layers (max %d)", (long long)L, RRB_MAX_LAYERS);
    if (n_storage_given != 0 && n_storage_given != n_storage_expected)
        return fail(RRB_EINVAL, "pass all %d storage outputs of the model or none", n_storage_expected);
    if (!o.qsim && !P.o.qobs && n_storage_given == 0) return fail(RRB_EINVAL, "nothing to compute");
    if (N == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_mt, *d_fr, *d_etp, *d_fice = nullptr;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(T * L), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, mean_temp, (size_t)(T * L), &d_mt))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, frac_solid, (size_t)(T * L), &d_fr))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW3, etp, (size_t)T, &d_etp))) return rc;
    if (ice && (rc = stage_in(*P.c, P.o, B_RAW4, frac_ice, (size_t)L, &d_fice))) return rc;
    if ((rc = stage_common(&P, T, N, params, k))) return rc;
    // inits5 = (snow_pack_init, thermal_state_init, sca_init, s_init, r_init); host memory in both modes
    double in5[5];
    if (hyst) memcpy(in5, inits, sizeof(in5));
    else { in5[0] = inits[0]; in5[1] = inits[1]; in5[2] = 0.0; in5[3] = inits[2]; in5[4] = inits[3]; }
    double x4_max;
    if ((rc = resolve_x4_max(&P, params, N, k, (hyst ? 4 : 2) + 3, &x4_max))) return rc;
    if (!(x4_max <= RRB_MAX_X4))
        return fail(RRB_EUNSUPPORTED, "GR4J x4 up to %g in this batch; the unit hydrograph buffers support x4 <= %g",
                    x4_max, RRB_MAX_X4);
    const int LC = cema_layer_class((int)L);
    void *F = nullptr, *gt = nullptr;
    if ((rc = P.c->ensure(B_F, forcing_bytes(T, cema_TT(LC), cema_R(LC)), &F))) return rc;
    if ((rc = P.c->ensure(B_GT, sizeof(double) * 2 * kCemaMaxLayers, &gt))) return rc;
    RRB_CUDA(pack_cemaneige(d_prec, d_mt, d_fr, d_etp, T, (int)L, (double*)F, (double*)gt, P.s));
    Job job;
    job.T = T; job.N = N;
    job.outs = {{o.qsim, N}, {o.G, L * N}, {o.eTG, L * N}, {o.s_store, N}, {o.r_store, N},
                {o.sca, L * N}, {o.icemelt, N}, {o.snowmelt, N}};
    job.state_slots = state_slots_snowice(family, (int)L, x4_max);
    const double* dp = P.d_params;
    const double i0 = in5[0], i1 = in5[1], i2 = in5[2], i3 = in5[3], i4 = in5[4];
    job.launch = [=](const Slab& sl, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        const double in[5] = {i0, i1, i2, i3, i4};
        SnowIceOut so{out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]};
        return launch_snowice(family, (const double*)F, (const double*)gt, d_fice, T, (int)L, in, dp, N, x4_max, so, sl,
                              ob, cfg);
    };
    return run_job(*P.c, P.o, job, P.d_qobs, P.d_mse);
}

int rrb_cemaneigegr4jice_simulate(const double* prec, const double* mean_temp, const double* etp, const double* frac_ice,
                                  const double* frac_solid, int64_t T, int64_t L, const double* inits,
                                  const double* params, int64_t N, double* qsim, double* G, double* eTG, double* s_store,
                                  double* r_store, double* icemelt, const rrb_opts* opts) {
    SnowIceOut o{qsim, G, eTG, s_store, r_store, nullptr, icemelt, nullptr};
    const int n = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr) + (icemelt != nullptr);
    return snowice_simulate(2, prec, mean_temp, etp, frac_ice, frac_solid, T, L, inits, params, N, o, n, 5, opts);
}

int rrb_cemaneigehystgr4j_simulate(const double* prec, const double* mean_temp, const double* etp,
                                   const double* frac_solid, int64_t T, int64_t L, const double* inits,
                                   const double* params, int64_t N, double* qsim, double* G, double* eTG,
                                   double* s_store, double* r_store, double* sca, const rrb_opts* opts) {
    SnowIceOut o{qsim, G, eTG, s_store, r_store, sca, nullptr, nullptr};
    const int n = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr) + (sca != nullptr);
    return snowice_simulate(1, prec, mean_temp, etp, nullptr, frac_solid, T, L, inits, params, N, o, n, 5, opts);
}

int rrb_cemaneigehystgr4jice_simulate(const double* prec, const double* mean_temp, const double* etp,
                                      const double* frac_ice, const double* frac_solid, int64_t T, int64_t L,
                                      const double* inits, const double* params, int64_t N, double* qsim, double* G,
                                      double* eTG, double* s_store, double* r_store, double* sca, double* icemelt,
                                      double* snowmelt, const rrb_opts* opts) {
    SnowIceOut o{qsim, G, eTG, s_store, r_store, sca, icemelt, snowmelt};
    const int n = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr) + (sca != nullptr) +
                  (icemelt != nullptr) + (snowmelt != nullptr);
    return snowice_simulate(3, prec, mean_temp, etp, frac_ice, frac_solid, T, L, inits, params, N, o, n, 7, opts);
}

// ---- host evaluation of the FAST math (CPU test-suite) ----
int rrb_gr4j_simulate_multi(const double* prec, const double* etp, int64_t C, int64_t T, const double* inits,
                            const double* params, int64_t N, double* qsim, double* s_store, double* r_store,
                            const rrb_opts* opts) {
    Prepared P;
    int rc = prepare(opts, T, N, params, 4, &P);
    if (rc) return rc;
    if (C < 0) return fail(RRB_EINVAL, "C = %lld", (long long)C);
    if (!prec || !etp || !inits) return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    if ((s_store != nullptr) != (r_store != nullptr)) return fail(RRB_EINVAL, "pass both storage outputs or none");
    if (!qsim && !P.o.qobs && !s_store) return fail(RRB_EINVAL, "nothing to compute");
    if (C > 65535) return fail(RRB_EUNSUPPORTED, "C = %lld catchments per call (max 65535)", (long long)C);
    if (N == 0 || C == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_etp, *d_inits;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(C * T), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, etp, (size_t)(C * T), &d_etp))) return rc;
    // the four-wide inits rows: (s_init, r_init, -, -)
    std::vector<double> in4((size_t)C * 4, 0.0);
    for (int64_t k = 0; k < C; ++k) { in4[4 * k] = inits[2 * k]; in4[4 * k + 1] = inits[2 * k + 1]; }
    if ((rc = stage_in(*P.c, P.o, B_PARAMS, params, (size_t)(C * N * 4), &P.d_params))) return rc;
    double x4_max;  // (uses the scalar scratch slot: before the inits are staged into it)
    if ((rc = resolve_x4_max(&P, params, C * N, 4, 3, &x4_max))) return rc;
    if (!(x4_max <= RRB_MAX_X4))
        return fail(RRB_EUNSUPPORTED, "GR4J x4 up to %g in this batch; the unit hydrograph buffers support x4 <= %g",
                    x4_max, RRB_MAX_X4);
    if ((rc = stage_multi(P, C, T, N, in4.data(), &d_inits))) return rc;
    RRB_CUDA(cudaStreamSynchronize(P.s));  // in4 is a temporary
    const int64_t fstride = forcing_stride_flagged(T, kGr4jTT, kGr4jR);
    void* F = nullptr;
    if ((rc = P.c->ensure(B_F, sizeof(double) * (size_t)(C * fstride), &F))) return rc;
    RRB_CUDA(pack_gr4j(d_prec, d_etp, T, (double*)F, P.s, (int)C));
    const double* dp = P.d_params;
    const std::vector<MultiOut> outs = {{qsim, T * N}, {s_store, T * N}, {r_store, T * N}};
    return run_multi(P, C, T, N, outs, [=](int64_t c0, int64_t c1, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        Batch b{(int)(c1 - c0), fstride, T * N, d_inits + 4 * c0};
        Slab slab{0, T, 0, nullptr, 0, 0};
        return launch_gr4j((const double*)F + c0 * fstride, T, 0.0, 0.0, dp + c0 * N * 4, N, x4_max, out[0], out[1], out[2],
                           slab, ob, cfg, b);
    });
}

int rrb_cemaneigegr4j_simulate_multi(const double* prec, const double* mean_temp, const double* etp,
                                     const double* frac_solid, int64_t C, int64_t T, int64_t L, const double* inits,
                                     const double* params, int64_t N, double* qsim, double* G, double* eTG,
                                     double* s_store, double* r_store, const rrb_opts* opts) {
    Prepared P;
    int rc = prepare(opts, T, N, params, 6, &P);
    if (rc) return rc;
    if (C < 0) return fail(RRB_EINVAL, "C = %lld", (long long)C);
    if (!prec || !mean_temp || !etp || !frac_solid || !inits) return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    if (L < 1) return fail(RRB_EINVAL, "L = %lld", (long long)L);
    if (L > RRB_MAX_LAYERS) return fail(RRB_EUNSUPPORTED, "L = %lld elevation layers (max %d)", (long long)L, RRB_MAX_LAYERS);
    const int nst = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr);
    if (nst != 0 && nst != 4) return fail(RRB_EINVAL, "pass all four storage outputs or none");
    if (!qsim && !P.o.qobs && nst == 0) return fail(RRB_EINVAL, "nothing to compute");
    if (C > 65535) return fail(RRB_EUNSUPPORTED, "C = %lld catchments per call (max 65535)", (long long)C);
    if (N == 0 || C == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_mt, *d_fr, *d_etp, *d_inits;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(C * T * L), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, mean_temp, (size_t)(C * T * L), &d_mt))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, frac_solid, (size_t)(C * T * L), &d_fr))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW3, etp, (size_t)(C * T), &d_etp))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_PARAMS, params, (size_t)(C * N * 6), &P.d_params))) return rc;
    double x4_max;  // (uses the scalar scratch slot: before the inits are staged into it)
    if ((rc = resolve_x4_max(&P, params, C * N, 6, 5, &x4_max))) return rc;
    if (!(x4_max <= RRB_MAX_X4))
        return fail(RRB_EUNSUPPORTED, "GR4J x4 up to %g in this batch; the unit hydrograph buffers support x4 <= %g",
                    x4_max, RRB_MAX_X4);
    if ((rc = stage_multi(P, C, T, N, inits, &d_inits))) return rc;
    RRB_CUDA(cudaStreamSynchronize(P.s));  // the caller's inits array may be a temporary
    const int LC = cema_layer_class((int)L);
    const int64_t fstride = forcing_stride_flagged(T, cema_TT(LC), cema_R(LC));
    void *F = nullptr, *gt = nullptr;
    if ((rc = P.c->ensure(B_F, sizeof(double) * (size_t)(C * fstride), &F))) return rc;
    if ((rc = P.c->ensure(B_GT, sizeof(double) * 2 * kCemaMaxLayers * (size_t)C, &gt))) return rc;
    RRB_CUDA(pack_cemaneige(d_prec, d_mt, d_fr, d_etp, T, (int)L, (double*)F, (double*)gt, P.s, (int)C));
    const double* dp = P.d_params;
    const std::vector<MultiOut> outs = {{qsim, T * N}, {G, T * L * N}, {eTG, T * L * N}, {s_store, T * N}, {r_store, T * N}};
    const double zero4[4] = {0, 0, 0, 0};
    return run_multi(P, C, T, N, outs, [=](int64_t c0, int64_t c1, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        Batch b{(int)(c1 - c0), fstride, T * N, d_inits + 4 * c0};
        Slab slab{0, T, 0, nullptr, 0, 0};
        return launch_cemaneigegr4j((const double*)F + c0 * fstride, (const double*)gt + c0 * 2 * kCemaMaxLayers, T, (int)L,
                                    zero4, dp + c0 * N * 6, N, x4_max, out[0], out[1], out[2], out[3], out[4], slab, ob,
                                    cfg, b);
    });
}


// ---- catchment batches of ABC, Cemaneige and the snow-ice couplings (completed in round 2)
int rrb_abc_simulate_multi(const double* prec, int64_t C, int64_t T, const double* inits, const double* params, int64_t N,
                           double* qsim, double* storage, const rrb_opts* opts) {
    Prepared P;
    int rc = prepare(opts, T, N, params, 3, &P);
    if (rc) return rc;
    if (C < 0) return fail(RRB_EINVAL, "C = %lld", (long long)C);
    if (!prec || !inits) return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    if (!qsim && !P.o.qobs && !storage) return fail(RRB_EINVAL, "nothing to compute");
    if (C > 65535) return fail(RRB_EUNSUPPORTED, "C = %lld catchments per call (max 65535)", (long long)C);
    if (P.o.state_in || P.o.state_out) return fail(RRB_EUNSUPPORTED, "state_in / state_out: single-catchment calls only");
    if (N == 0 || C == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_inits;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(C * T), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_PARAMS, params, (size_t)(C * N * 3), &P.d_params))) return rc;
    std::vector<double> in4((size_t)C * 4, 0.0);
    for (int64_t k = 0; k < C; ++k) in4[4 * k] = inits[k];
    if ((rc = stage_multi(P, C, T, N, in4.data(), &d_inits))) return rc;
    RRB_CUDA(cudaStreamSynchronize(P.s));  // in4 is a temporary
    const int64_t Tpad = padded_steps(T, kAbcTT);
    void* F = nullptr;
    if ((rc = P.c->ensure(B_F, sizeof(double) * (size_t)(C * Tpad) * kAbcR, &F))) return rc;
    RRB_CUDA(pack_abc(d_prec, T, (double*)F, P.s, (int)C));
    const double* dp = P.d_params;
    const std::vector<MultiOut> outs = {{qsim, T * N}, {storage, T * N}};
    return run_multi(P, C, T, N, outs, [=](int64_t c0, int64_t c1, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        Batch b{(int)(c1 - c0), Tpad * kAbcR, T * N, d_inits + 4 * c0};
        Slab slab{0, T, 0, nullptr, 0, 0};
        return launch_abc((const double*)F + c0 * Tpad * kAbcR, T, 0.0, dp + c0 * N * 3, N, out[0], out[1], slab, ob, cfg, b);
    });
}

int rrb_cemaneige_simulate_multi(const double* prec, const double* mean_temp, const double* frac_solid, int64_t C, int64_t T,
                                 int64_t L, const double* inits, const double* params, int64_t param_stride, int64_t N,
                                 double* outflow, double* G, double* eTG, const rrb_opts* opts) {
    Prepared P;
    int rc = prepare(opts, T, N, params, param_stride, &P);
    if (rc) return rc;
    if (C < 0) return fail(RRB_EINVAL, "C = %lld", (long long)C);
    if (!prec || !mean_temp || !frac_solid || !inits) return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    if (L < 1) return fail(RRB_EINVAL, "L = %lld", (long long)L);
    if (L > RRB_MAX_LAYERS) return fail(RRB_EUNSUPPORTED, "L = %lld elevation layers (max %d)", (long long)L, RRB_MAX_LAYERS);
    if (param_stride < 2) return fail(RRB_EINVAL, "param_stride = %lld (< 2)", (long long)param_stride);
    if ((G != nullptr) != (eTG != nullptr)) return fail(RRB_EINVAL, "pass both storage outputs or none");
    if (!outflow && !P.o.qobs && !G) return fail(RRB_EINVAL, "nothing to compute");
    if (C > 65535) return fail(RRB_EUNSUPPORTED, "C = %lld catchments per call (max 65535)", (long long)C);
    if (P.o.state_in || P.o.state_out) return fail(RRB_EUNSUPPORTED, "state_in / state_out: ABC, HBV-Edu, GR4J single-catchment calls only");
    if (N == 0 || C == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_mt, *d_fr, *d_inits;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(C * T * L), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, mean_temp, (size_t)(C * T * L), &d_mt))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, frac_solid, (size_t)(C * T * L), &d_fr))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_PARAMS, params, (size_t)(C * N * param_stride), &P.d_params))) return rc;
    std::vector<double> in4((size_t)C * 4, 0.0);
    for (int64_t k = 0; k < C; ++k) { in4[4 * k] = inits[2 * k]; in4[4 * k + 1] = inits[2 * k + 1]; }
    if ((rc = stage_multi(P, C, T, N, in4.data(), &d_inits))) return rc;
    RRB_CUDA(cudaStreamSynchronize(P.s));  // in4 is a temporary
    const int LC = cema_layer_class((int)L);
    const int64_t fstride = forcing_stride_flagged(T, cema_TT(LC), cema_R(LC));
    void *F = nullptr, *gt = nullptr;
    if ((rc = P.c->ensure(B_F, sizeof(double) * (size_t)(C * fstride), &F))) return rc;
    if ((rc = P.c->ensure(B_GT, sizeof(double) * 2 * kCemaMaxLayers * (size_t)C, &gt))) return rc;
    RRB_CUDA(pack_cemaneige(d_prec, d_mt, d_fr, nullptr, T, (int)L, (double*)F, (double*)gt, P.s, (int)C));
    const double* dp = P.d_params;
    const std::vector<MultiOut> outs = {{outflow, T * N}, {G, T * L * N}, {eTG, T * L * N}};
    return run_multi(P, C, T, N, outs, [=](int64_t c0, int64_t c1, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        Batch b{(int)(c1 - c0), fstride, T * N, d_inits + 4 * c0};
        Slab slab{0, T, 0, nullptr, 0, 0};
        return launch_cemaneige((const double*)F + c0 * fstride, (const double*)gt + c0 * 2 * kCemaMaxLayers, T, (int)L, 0.0, 0.0,
                                dp + c0 * N * param_stride, param_stride, N, out[0], out[1], out[2], slab, ob, cfg, b);
    });
}

static int snowice_simulate_multi(int family, const double* prec, const double* mean_temp, const double* etp,
                                  const double* frac_ice, const double* frac_solid, int64_t C, int64_t T, int64_t L,
                                  const double* inits, const double* params, int64_t N, const SnowIceOut& o,
                                  int n_storage_given, int n_storage_expected, const rrb_opts* opts) {
    const bool hyst = family & 1, ice = family & 2;
    const int k = 6 + (hyst ? 2 : 0) + (ice ? 1 : 0);
    Prepared P;
    int rc = prepare(opts, T, N, params, k, &P);
    if (rc) return rc;
    if (C < 0) return fail(RRB_EINVAL, "C = %lld", (long long)C);
    if (!prec || !mean_temp || !etp || !frac_solid || !inits || (ice && !frac_ice))
        return fail(RRB_EINVAL, "NULL forcing / inits pointer");
    if (L < 1) return fail(RRB_EINVAL, "L = %lld", (long long)L);
    if (L > RRB_MAX_LAYERS) return fail(RRB_EUNSUPPORTED, "L = %lld elevation layers (max %d)", (long long)L, RRB_MAX_LAYERS);
    if (n_storage_given != 0 && n_storage_given != n_storage_expected)
        return fail(RRB_EINVAL, "pass all %d storage outputs of the model or none", n_storage_expected);
    if (!o.qsim && !P.o.qobs && n_storage_given == 0) return fail(RRB_EINVAL, "nothing to compute");
    if (C > 65535) return fail(RRB_EUNSUPPORTED, "C = %lld catchments per call (max 65535)", (long long)C);
    if (P.o.state_in || P.o.state_out) return fail(RRB_EUNSUPPORTED, "state_in / state_out: ABC, HBV-Edu, GR4J single-catchment calls only");
    if (N == 0 || C == 0) return RRB_OK;
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_prec, *d_mt, *d_fr, *d_etp, *d_fice = nullptr, *d_inits;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)(C * T * L), &d_prec))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, mean_temp, (size_t)(C * T * L), &d_mt))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, frac_solid, (size_t)(C * T * L), &d_fr))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW3, etp, (size_t)(C * T), &d_etp))) return rc;
    if (ice && (rc = stage_in(*P.c, P.o, B_RAW4, frac_ice, (size_t)(C * L), &d_fice))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_PARAMS, params, (size_t)(C * N * k), &P.d_params))) return rc;
    double x4_max;  // (uses the scalar scratch slot: before the inits are staged into it)
    if ((rc = resolve_x4_max(&P, params, C * N, k, (hyst ? 4 : 2) + 3, &x4_max))) return rc;
    if (!(x4_max <= RRB_MAX_X4))
        return fail(RRB_EUNSUPPORTED, "GR4J x4 up to %g in this batch; the unit hydrograph buffers support x4 <= %g",
                    x4_max, RRB_MAX_X4);
    // rows of kSnowIceInitsStride: (snow_pack_init, thermal_state_init, s_init, r_init, sca_init, -, -, -)
    const int nin = hyst ? 5 : 4;
    std::vector<double> in8((size_t)C * kSnowIceInitsStride, 0.0);
    for (int64_t c = 0; c < C; ++c) {
        const double* r = inits + c * nin;
        double* w = in8.data() + c * kSnowIceInitsStride;
        w[0] = r[0]; w[1] = r[1];
        if (hyst) { w[4] = r[2]; w[2] = r[3]; w[3] = r[4]; }
        else { w[2] = r[2]; w[3] = r[3]; }
    }
    if ((rc = stage_multi(P, C, T, N, in8.data(), &d_inits, kSnowIceInitsStride))) return rc;
    RRB_CUDA(cudaStreamSynchronize(P.s));  // in8 is a temporary
    const int LC = cema_layer_class((int)L);
    const int64_t fstride = forcing_stride_flagged(T, cema_TT(LC), cema_R(LC));
    void *F = nullptr, *gt = nullptr;
    if ((rc = P.c->ensure(B_F, sizeof(double) * (size_t)(C * fstride), &F))) return rc;
    if ((rc = P.c->ensure(B_GT, sizeof(double) * 2 * kCemaMaxLayers * (size_t)C, &gt))) return rc;
    RRB_CUDA(pack_cemaneige(d_prec, d_mt, d_fr, d_etp, T, (int)L, (double*)F, (double*)gt, P.s, (int)C));
    const double* dp = P.d_params;
    const std::vector<MultiOut> outs = {{o.qsim, T * N}, {o.G, T * L * N}, {o.eTG, T * L * N}, {o.s_store, T * N},
                                        {o.r_store, T * N}, {o.sca, T * L * N}, {o.icemelt, T * N}, {o.snowmelt, T * N}};
    const double zero5[5] = {0, 0, 0, 0, 0};
    return run_multi(P, C, T, N, outs, [=](int64_t c0, int64_t c1, double* const* out, const Objective& ob, const LaunchCfg& cfg) {
        Batch b{(int)(c1 - c0), fstride, T * N, d_inits + kSnowIceInitsStride * c0};
        Slab slab{0, T, 0, nullptr, 0, 0};
        SnowIceOut so{out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]};
        return launch_snowice(family, (const double*)F + c0 * fstride, (const double*)gt + c0 * 2 * kCemaMaxLayers,
                              d_fice ? d_fice + c0 * L : nullptr, T, (int)L, zero5, dp + c0 * N * k, N, x4_max, so, slab, ob, cfg, b);
    });
}

int rrb_cemaneigegr4jice_simulate_multi(const double* prec, const double* mean_temp, const double* etp, const double* frac_ice,
                                        const double* frac_solid, int64_t C, int64_t T, int64_t L, const double* inits,
                                        const double* params, int64_t N, double* qsim, double* G, double* eTG,
                                        double* s_store, double* r_store, double* icemelt, const rrb_opts* opts) {
    SnowIceOut o{qsim, G, eTG, s_store, r_store, nullptr, icemelt, nullptr};
    const int n = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr) + (icemelt != nullptr);
    return snowice_simulate_multi(2, prec, mean_temp, etp, frac_ice, frac_solid, C, T, L, inits, params, N, o, n, 5, opts);
}

int rrb_cemaneigehystgr4j_simulate_multi(const double* prec, const double* mean_temp, const double* etp,
                                         const double* frac_solid, int64_t C, int64_t T, int64_t L, const double* inits,
                                         const double* params, int64_t N, double* qsim, double* G, double* eTG,
                                         double* s_store, double* r_store, double* sca, const rrb_opts* opts) {
    SnowIceOut o{qsim, G, eTG, s_store, r_store, sca, nullptr, nullptr};
    const int n = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr) + (sca != nullptr);
    return snowice_simulate_multi(1, prec, mean_temp, etp, nullptr, frac_solid, C, T, L, inits, params, N, o, n, 5, opts);
}

int rrb_cemaneigehystgr4jice_simulate_multi(const double* prec, const double* mean_temp, const double* etp,
                                            const double* frac_ice, const double* frac_solid, int64_t C, int64_t T, int64_t L,
                                            const double* inits, const double* params, int64_t N, double* qsim, double* G,
                                            double* eTG, double* s_store, double* r_store, double* sca, double* icemelt,
                                            double* snowmelt, const rrb_opts* opts) {
    SnowIceOut o{qsim, G, eTG, s_store, r_store, sca, icemelt, snowmelt};
    const int n = (G != nullptr) + (eTG != nullptr) + (s_store != nullptr) + (r_store != nullptr) + (sca != nullptr) +
                  (icemelt != nullptr) + (snowmelt != nullptr);
    return snowice_simulate_multi(3, prec, mean_temp, etp, frac_ice, frac_solid, C, T, L, inits, params, N, o, n, 7, opts);
}

// ---- Cemaneige-family layer preprocessing ----
int rrb_snow_layers(const double* prec, const double* mean_temp, const double* min_temp, const double* max_temp,
                    int64_t T, int64_t L, const double* prec_factor, const double* delta_temp, const int32_t* flags,
                    double* layer_prec, double* layer_mean_temp, double* frac_solid, const rrb_opts* opts) {
    Prepared P;
    int rc = prepare(opts, T, 0, nullptr, 0, &P);
    if (rc) return rc;
    if (!prec || !mean_temp || !min_temp || !max_temp) return fail(RRB_EINVAL, "NULL station series");
    if (!prec_factor || !delta_temp || !flags) return fail(RRB_EINVAL, "NULL per-layer scalars");
    if (!layer_prec || !layer_mean_temp || !frac_solid) return fail(RRB_EINVAL, "NULL output");
    if (L < 1) return fail(RRB_EINVAL, "L = %lld", (long long)L);
    if (L > RRB_MAX_LAYERS) return fail(RRB_EUNSUPPORTED, "L = %lld elevation layers (max %d)", (long long)L, RRB_MAX_LAYERS);
    SnowLayerScalars k{};
    for (int l = 0; l < (int)L; ++l) {  // the per-layer scalars are host memory in both modes
        k.prec_factor[l] = prec_factor[l];
        k.delta_temp[l] = delta_temp[l];
        k.scale_prec[l] = (flags[l] & RRB_LAYER_SCALE_PREC) != 0;
        k.shift_temp[l] = (flags[l] & RRB_LAYER_SHIFT_TEMP) != 0;
        k.high[l] = (flags[l] & RRB_LAYER_HIGH) != 0;
    }
    std::lock_guard<std::mutex> lk(P.c->mu);
    HostDrain drain(P);
    const double *d_p, *d_me, *d_mn, *d_mx;
    if ((rc = stage_in(*P.c, P.o, B_RAW0, prec, (size_t)T, &d_p))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW1, mean_temp, (size_t)T, &d_me))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW2, min_temp, (size_t)T, &d_mn))) return rc;
    if ((rc = stage_in(*P.c, P.o, B_RAW3, max_temp, (size_t)T, &d_mx))) return rc;
    const bool host = P.o.mem != RRB_MEM_DEVICE;
    const size_t bytes = sizeof(double) * (size_t)(T * L);
    double* outs_h[3] = {layer_prec, layer_mean_temp, frac_solid};
    double* outs_d[3] = {layer_prec, layer_mean_temp, frac_solid};
    if (host)
        for (int j = 0; j < 3; ++j) {
            void* p = nullptr;
            if ((rc = P.c->ensure(B_OUT0 + j, bytes, &p))) return rc;
            outs_d[j] = (double*)p;
        }
    RRB_CUDA(launch_snow_layers(d_p, d_me, d_mn, d_mx, T, (int)L, k, outs_d[0], outs_d[1], outs_d[2], P.s));
    if (host) {
        for (int j = 0; j < 3; ++j)
            RRB_CUDA(cudaMemcpyAsync(outs_h[j], outs_d[j], bytes, cudaMemcpyDeviceToHost, P.s));
        RRB_CUDA(cudaStreamSynchronize(P.s));
    }
    return RRB_OK;
}

void rrb_host_fast_pow(const double* x, const double* y, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = fast_pow(x[i], y[i], &h_fast_tables);
}
void rrb_host_hbv_pow_step(const double* soil, const double* FC, const double* Beta, const double* liquid,
                           const double* soil_partial, int64_t n, double* prec_eff, double* soil_new) {
    for (int64_t i = 0; i < n; ++i)
        prec_eff[i] = hbv_pow_step_twin(soil[i], log2(FC[i]), Beta[i], liquid[i], soil_partial[i], &h_hbv_tables, &soil_new[i]);
}
void rrb_host_fast_exp2m1(const double* z, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = fast_exp2m1_nonneg(z[i], &h_fast_tables);
}

}  // extern "C"
