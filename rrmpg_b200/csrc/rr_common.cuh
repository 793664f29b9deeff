// rr_common.cuh -- shared device machinery for the ensemble recurrence kernels (sm_100a).
//
// Execution model (all five models):
//   * one thread owns one ensemble member for the whole series; its stores (snow pack, soil
//     moisture, routing stores, unit-hydrograph buffers) live in registers,
//   * the member-independent forcing of the catchment is packed time-major as F[t][R] doubles
//     and streamed HBM -> shared memory in tiles of TT timesteps by 1-D TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx::bytes) issued by one elected thread, multi-stage
//     ring guarded by mbarriers; every thread then reads F[t][*] as a shared-memory broadcast,
//   * per step every warp stores 32 consecutive doubles of the [T, N] output row (256 B,
//     streaming st.global.cs so the multi-GB output does not evict forcing tiles from L2).
//
// The kernels replace the serial member loops of the reference wrappers
// (rrmpg/models/hbvedu.py:199-209 and siblings); each model body cites the numba kernel it
// restates.  The translation unit is compiled with -fmad=false: the reference (numba without
// fastmath) never contracts a*b+c, so neither does the PRECISE path; FAST paths ask for FMAs
// explicitly with fma().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rrb {

constexpr int kStages = 3;        // forcing ring depth
constexpr int kMaxStateSlots = 80; // upper bound of doubles a member carries across time slabs

// ----------------------------------------------------------------------------------------
// mbarrier / TMA-1D primitives (PTX; SASS: SYNCS.*, UBLKCP)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order prior generic-proxy reads of a smem buffer before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// streaming (evict-first) stores of the output rows
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
// two neighbouring members of one output row in one 16-byte streaming store (p 16-byte aligned)
__device__ __forceinline__ void st_stream_pair(double* p, double x, double y) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}

// numba's lowering of Python max(0, x) / min(a, b) on floats (oracle/rr_oracle.c header):
// max(0, x) = (x > 0) ? x : 0.0 (so max(0, NaN) = 0), min(a, b) = (b < a) ? b : a.
// Written as setp + selp so they cost one DSETP and two SELs; the C++ ternary is pattern-matched by
// the compiler into a DSETP.MAX + NaN fix-up sequence of six instructions.
__device__ __forceinline__ double nb_max0(double x) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, 0d0000000000000000;\n\t"
        "selp.f64 %0, %1, 0d0000000000000000, p;\n\t}"
        : "=d"(r)
        : "d"(x));
    return r;
}
__device__ __forceinline__ double nb_min(double a, double b) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %2, %1;\n\tselp.f64 %0, %2, %1, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}

// ----------------------------------------------------------------------------------------
// Forcing tile pipeline.
//   F      : packed forcing [Tpad][R] (Tpad = ntiles*TT, zero padded), 16-byte aligned
//   smem   : kStages * TT * R doubles + kStages mbarriers (dynamic shared memory)
//   step(t, f) is called for every t in [t_begin, t_end) with f -> the R forcing values of t.
// Every thread of the CTA must call this (it contains CTA-wide barriers).
// ----------------------------------------------------------------------------------------
template <int R, int TT, class Step>
__device__ __forceinline__ void stream_forcing(const double* __restrict__ F, int64_t t_begin, int64_t t_end,
                                               Step&& step) {
    extern __shared__ __align__(128) unsigned char rrb_smem[];
    double* tiles = reinterpret_cast<double*>(rrb_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(rrb_smem + sizeof(double) * kStages * TT * R);
    constexpr uint32_t kTileBytes = TT * R * sizeof(double);
    static_assert(kTileBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

    if (t_end <= t_begin) return;
    const int64_t k_begin = t_begin / TT;
    const int64_t k_end = (t_end + TT - 1) / TT;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            const int64_t k = k_begin + s;
            if (k < k_end) {
                mbar_arrive_expect_tx(&full[s], kTileBytes);
                tma_load_1d(tiles + (size_t)s * TT * R, F + (size_t)k * TT * R, kTileBytes, &full[s]);
            }
        }
    }
    int stage = 0;
    uint32_t parity = 0;
    for (int64_t k = k_begin; k < k_end; ++k) {
        mbar_wait(&full[stage], parity);
        const double* tile = tiles + (size_t)stage * TT * R;
        const int64_t t0 = k * TT;
        const int lo = (int)((t_begin > t0) ? (t_begin - t0) : 0);
        const int hi = (int)((t_end < t0 + TT) ? (t_end - t0) : TT);
        if (lo == 0 && hi == TT) {
#pragma unroll 4
            for (int tt = 0; tt < TT; ++tt) step(t0 + tt, tile + tt * R);
        } else {
            for (int tt = lo; tt < hi; ++tt) step(t0 + tt, tile + tt * R);
        }
        __syncthreads();  // every thread is done reading this stage
        if (threadIdx.x == 0 && k + kStages < k_end) {
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(&full[stage], kTileBytes);
            tma_load_1d(tiles + (size_t)stage * TT * R, F + (size_t)(k + kStages) * TT * R, kTileBytes,
                        &full[stage]);
        }
        if (++stage == kStages) {
            stage = 0;
            parity ^= 1u;
        }
    }
}

// explicit shared-window loads (32-bit shared address computed once; a generic pointer would make
// the compiler re-derive the window base -- S2UR SR_CgaCtaId + ULEA -- at every use)
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// ----------------------------------------------------------------------------------------
// Forcing tile pipeline, register-staged variant.  FV is a struct of the R forcing values of one
// timestep with `static FV load(uint32_t shared_addr)`; the values of step t+1 are fetched from
// shared memory while step t computes, so the ~30-cycle LDS latency never sits in front of the
// recurrence.  step(t, fv) is called for every t in [t_begin, t_end), in order.
// ----------------------------------------------------------------------------------------
template <int R, int TT, class FV, class Step>
__device__ __forceinline__ void stream_forcing_regs(const double* __restrict__ F, int64_t t_begin, int64_t t_end,
                                                    Step&& step) {
    extern __shared__ __align__(128) unsigned char rrb_smem[];
    double* tiles = reinterpret_cast<double*>(rrb_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(rrb_smem + sizeof(double) * kStages * TT * R);
    constexpr uint32_t kTileBytes = TT * R * sizeof(double);
    constexpr uint32_t kRowBytes = R * sizeof(double);
    static_assert(kTileBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

    if (t_end <= t_begin) return;
    const int64_t k_begin = t_begin / TT;
    const int64_t k_end = (t_end + TT - 1) / TT;
    const uint32_t tiles_addr = smem_u32(tiles);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            const int64_t k = k_begin + s;
            if (k < k_end) {
                mbar_arrive_expect_tx(&full[s], kTileBytes);
                tma_load_1d(tiles + (size_t)s * TT * R, F + (size_t)k * TT * R, kTileBytes, &full[s]);
            }
        }
    }
    int stage = 0;
    uint32_t parity = 0;
    for (int64_t k = k_begin; k < k_end; ++k) {
        mbar_wait(&full[stage], parity);
        const uint32_t base = tiles_addr + (uint32_t)stage * kTileBytes;
        const int64_t t0 = k * TT;
        const int lo = (int)((t_begin > t0) ? (t_begin - t0) : 0);
        const int hi = (int)((t_end < t0 + TT) ? (t_end - t0) : TT);
        FV nxt = FV::load(base + (uint32_t)lo * kRowBytes);
        if (lo == 0 && hi == TT) {
#pragma unroll 4
            for (int tt = 0; tt < TT; ++tt) {
                const FV cur = nxt;
                nxt = FV::load(base + (uint32_t)((tt + 1 < TT) ? tt + 1 : tt) * kRowBytes);
                step(t0 + tt, cur);
            }
        } else {
            for (int tt = lo; tt < hi; ++tt) {
                const FV cur = nxt;
                nxt = FV::load(base + (uint32_t)((tt + 1 < hi) ? tt + 1 : tt) * kRowBytes);
                step(t0 + tt, cur);
            }
        }
        __syncthreads();  // every thread is done reading this stage
        if (threadIdx.x == 0 && k + kStages < k_end) {
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(&full[stage], kTileBytes);
            tma_load_1d(tiles + (size_t)stage * TT * R, F + (size_t)(k + kStages) * TT * R, kTileBytes,
                        &full[stage]);
        }
        if (++stage == kStages) {
            stage = 0;
            parity ^= 1u;
        }
    }
}

// ----------------------------------------------------------------------------------------
// Forcing tile pipeline, grouped variant.  group(ic<G>, t0, fv[G]) is called for G consecutive
// timesteps at a time (tile rows aligned to G), group(ic<1>, t, fv[1]) for the ragged steps at the
// edges of a time slab.  A model can then run the part of its step that does not hang on the long
// recurrence for all G steps first (high ILP), and the dependent chains afterwards.
// ----------------------------------------------------------------------------------------
template <int V>
struct ic {
    static constexpr int value = V;
};

template <int R, int TT, int G, class FV, class Group>
__device__ __forceinline__ void stream_forcing_grouped(const double* __restrict__ F, int64_t t_begin, int64_t t_end,
                                                       Group&& group) {
    static_assert(TT % G == 0, "tiles hold whole groups");
    extern __shared__ __align__(128) unsigned char rrb_smem[];
    double* tiles = reinterpret_cast<double*>(rrb_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(rrb_smem + sizeof(double) * kStages * TT * R);
    constexpr uint32_t kTileBytes = TT * R * sizeof(double);
    constexpr uint32_t kRowBytes = R * sizeof(double);
    static_assert(kTileBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

    if (t_end <= t_begin) return;
    const int64_t k_begin = t_begin / TT;
    const int64_t k_end = (t_end + TT - 1) / TT;
    uint32_t tiles_addr = smem_u32(tiles);
    asm volatile("" : "+r"(tiles_addr));  // opaque: keep it in a register instead of re-deriving it

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            const int64_t k = k_begin + s;
            if (k < k_end) {
                mbar_arrive_expect_tx(&full[s], kTileBytes);
                tma_load_1d(tiles + (size_t)s * TT * R, F + (size_t)k * TT * R, kTileBytes, &full[s]);
            }
        }
    }
    int stage = 0;
    uint32_t parity = 0;
    for (int64_t k = k_begin; k < k_end; ++k) {
        mbar_wait(&full[stage], parity);
        const int64_t t0 = k * TT;
        const int lo = (int)((t_begin > t0) ? (t_begin - t0) : 0);
        const int hi = (int)((t_end < t0 + TT) ? (t_end - t0) : TT);
        uint32_t addr = tiles_addr + (uint32_t)stage * kTileBytes + (uint32_t)lo * kRowBytes;
        int64_t t = t0 + lo;
        int left = hi - lo;
        while (left > 0 && (t % G) != 0) {  // ragged head of a time slab
            FV f1[1] = {FV::load(addr)};
            group(ic<1>{}, t, f1);
            addr += kRowBytes; ++t; --left;
        }
#pragma unroll 1
        for (int j = left / G; j > 0; --j) {
            FV f[G];
#pragma unroll
            for (int g = 0; g < G; ++g) f[g] = FV::load(addr + (uint32_t)g * kRowBytes);
            group(ic<G>{}, t, f);
            addr += G * kRowBytes;
            t += G;
        }
        for (left %= G; left > 0; --left) {  // ragged tail
            FV f1[1] = {FV::load(addr)};
            group(ic<1>{}, t, f1);
            addr += kRowBytes; ++t;
        }
        __syncthreads();  // every thread is done reading this stage
        if (threadIdx.x == 0 && k + kStages < k_end) {
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(&full[stage], kTileBytes);
            tma_load_1d(tiles + (size_t)stage * TT * R, F + (size_t)(k + kStages) * TT * R, kTileBytes,
                        &full[stage]);
        }
        if (++stage == kStages) {
            stage = 0;
            parity ^= 1u;
        }
    }
}

// ----------------------------------------------------------------------------------------
// Correctly rounded a / b for a loop-invariant b (Markstein): with y = RN(1/b) computed once by a real
// division, q0 = RN(a y), and two residual corrections r = a - b q (exact in an FMA), q += r y, the
// result equals the IEEE quotient RN(a/b) bit for bit when nothing under- or overflows (checked against
// hardware division on random and adversarial operand pairs by the oracle-side twin, tests/test_fastmath.py).  The fast
// path is taken for a positive a with a mid-range exponent (one unsigned compare on its high word; span = 0
// disables it, e.g. when b itself is out of range) and for a == +0; everything else -- negative, tiny, huge,
// NaN -- uses the hardware division.  5 fp64 instructions instead of ~12 plus a slow-path call.
// ----------------------------------------------------------------------------------------
constexpr uint32_t kDivSpanOk = 0x78000000u;  // exponents 2^-960 .. 2^+960: with b in [2^-60, 2^60] the quotient stays
                                              // finite and normal (oracle_check_invariant_division, tests/test_fastmath.py)
__device__ __forceinline__ uint32_t div_invariant_span(double b) {
    return (b >= 0x1p-60 && b <= 0x1p60) ? kDivSpanOk : 0u;
}
static __device__ __noinline__ double div_slow_path(double a, double b) { return a / b; }
__device__ __forceinline__ double div_by_invariant(double a, double b, double inv_b, uint32_t span) {
    const uint32_t h = (uint32_t)__double2hiint(a);
    const bool fast = ((h - 0x03F00000u) < span) | (((h | (uint32_t)__double2loint(a)) == 0u) & (span != 0u));
    double q = a * inv_b;
    double r = fma(-b, q, a);
    q = fma(r, inv_b, q);
    r = fma(-b, q, a);
    q = fma(r, inv_b, q);
    if (!fast) q = div_slow_path(a, b);
    return q;
}

// keep a loop-invariant value in registers: without this the compiler prefers to recompute cheap
// expressions (1 - K_2, shared-window bases, ...) inside the time loop, where every fp64 instruction
// costs two issue slots
__device__ __forceinline__ void pin(double& v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void pin(uint32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(int64_t& v) { asm volatile("" : "+l"(v)); }

// forcing values the FAST paths accept (finite, |x| <= 1e6); everything else selects the reference-order step
__device__ __forceinline__ bool forcing_value_sane(double x) { return fabs(x) <= 1e6; }

// dynamic shared memory = [forcing ring | mbarriers | (FAST math tables)], rounded to 16 B
template <int R, int TT>
__host__ __device__ constexpr size_t forcing_smem_bytes() {
    return (sizeof(double) * kStages * TT * R + sizeof(uint64_t) * kStages + 15) & ~size_t(15);
}

}  // namespace rrb
