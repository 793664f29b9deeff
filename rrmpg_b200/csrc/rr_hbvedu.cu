// rr_hbvedu.cu -- HBV-Edu ensemble kernel (the model BASELINE.json's metric is quoted on).
// Restates run_hbvedu (rrmpg/models/hbvedu_model.py:16-129) for N members at once and replaces
// the member loop of HBVEdu.simulate (rrmpg/models/hbvedu.py:199-209).
//
// Packed forcing per timestep (member independent): F[t] = { temp, prec, dT, PEm } with
//   dT  = temp[t] - T_m[month[t]]   (the inner subtraction of hbvedu_model.py:102, bit-identical)
//   PEm = PE_m[month[t]]
// Per member: 4 stores (snow, soil, s1, s2) and 11 parameters in registers.
//
// MATH = PRECISE: IEEE divisions, CUDA libm pow (<= 2 ulp from glibc's), no contraction: every
//                 operation of the reference in the reference's order.
// MATH = FAST   : same recurrence, reciprocals of FC / PWP hoisted, table-driven pow
//                 (rr_math.cuh, ~1e-15 relative), and the pow skipped for warps whose members all
//                 have liquid_water == 0 (then prec_eff = 0 * finite = 0 exactly).  Discharge
//                 stays within rtol 1e-10 of the reference (tests/test_parity_gpu.py).
#include "rr_common.cuh"
#include "rr_kernels.h"
#include "rr_math.cuh"

namespace rrb {

__global__ void hbv_pack_kernel(const double* __restrict__ temp, const double* __restrict__ prec,
                                const int8_t* __restrict__ month0, const double* __restrict__ PE_m,
                                const double* __restrict__ T_m, int64_t T, int64_t Tpad, double* __restrict__ F) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= Tpad) return;
    double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
    if (t < T) {
        const int m = month0[t];
        v.x = temp[t];
        v.y = prec[t];
        v.z = temp[t] - T_m[m];
        v.w = PE_m[m];
    }
    reinterpret_cast<double4*>(F)[t] = v;
}

cudaError_t pack_hbvedu(const double* temp, const double* prec, const int8_t* month0, const double* PE_m,
                        const double* T_m, int64_t T, double* F, cudaStream_t s) {
    int64_t Tpad = padded_steps(T, kHbvTT);
    hbv_pack_kernel<<<(unsigned)((Tpad + 255) / 256), 256, 0, s>>>(temp, prec, month0, PE_m, T_m, T, Tpad, F);
    return cudaGetLastError();
}

struct HbvOut {
    double *qsim, *snow, *soil, *s1, *s2;
};

template <int MATH, bool STORAGE, bool OBJ>
__global__ void hbv_kernel(const double* __restrict__ F, double snow0, double soil0, double s10, double s20,
                           const double* __restrict__ params, int64_t N, HbvOut out, Slab slab,
                           Objective obj) {
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool active = gi < N;
    const int64_t i = active ? gi : N - 1;
    // record order = HBVEdu._dtype (rrmpg/models/hbvedu.py:63-66)
    const double* p = params + 11 * i;
    const double T_t = p[0], DD = p[1], FC = p[2], Beta = p[3], C = p[4], PWP = p[5];
    const double K_0 = p[6], K_1 = p[7], K_2 = p[8], K_p = p[9], L = p[10];
    const double inv_FC = 1.0 / FC, inv_PWP = 1.0 / PWP;  // FAST only

    double snow = snow0, soil = soil0, s1 = s10, s2 = s20;  // hbvedu_model.py:78-81
    double acc = 0.0;
    if (slab.t_begin > 0) {
        snow = slab.state[0 * N + i];
        soil = slab.state[1 * N + i];
        s1 = slab.state[2 * N + i];
        s2 = slab.state[3 * N + i];
        if (OBJ) acc = slab.state[4 * N + i];
    }
    const int64_t off = i - slab.row0 * N;
    double* q_o = out.qsim ? out.qsim + off : nullptr;

    extern __shared__ __align__(128) unsigned char rrb_smem[];
    const FastTables* tb = nullptr;
    if (MATH == RRB_MATH_FAST_) tb = fastmath_tables_to_smem(rrb_smem + forcing_smem_bytes<kHbvR, kHbvTT>());

    stream_forcing<kHbvR, kHbvTT>(F, slab.t_begin, slab.t_end, [&](int64_t t, const double* f) {
        double qv = 0.0;
        if (t > 0) {  // hbvedu_model.py:84 -- the loop starts at t = 1
            const double2 f01 = *reinterpret_cast<const double2*>(f);
            const double2 f23 = *reinterpret_cast<const double2*>(f + 2);
            const double temp = f01.x, prec = f01.y, dT = f23.x, PEm = f23.y;
            double snow_new, liquid;
            if (temp < T_t) {  // :87
                snow_new = snow + prec;  // :89
                liquid = 0.0;            // :91
            } else {
                const double m = DD * (temp - T_t);
                snow_new = nb_max0(snow - m);    // :94
                liquid = prec + nb_min(snow, m);  // :96
            }
            double prec_eff, ea;
            const double pe = (1 + C * dT) * PEm;  // :102
            if (MATH == RRB_MATH_FAST_) {
                const double x = soil * inv_FC;
                // liquid == 0 and a finite positive power  =>  prec_eff = +0 exactly as in the reference
                const bool need_pow = !(liquid == 0.0 && x > 0x1p-16 && x < 0x1p16 && fabs(Beta) < 32.0);
                prec_eff = 0.0;
                if (need_pow) prec_eff = liquid * fast_pow(x, Beta, tb);
                ea = (soil > PWP) ? pe : pe * (soil * inv_PWP);
            } else {
                prec_eff = liquid * pow(soil / FC, Beta);        // :99
                ea = (soil > PWP) ? pe : pe * (soil / PWP);      // :105-108
            }
            const double soil_new = soil + liquid - prec_eff - ea;  // :111
            const double over = nb_max0(s1 - L);
            const double s1_new = s1 + prec_eff - over * K_0 - s1 * K_1 - s1 * K_p;  // :114-118
            const double s2_new = s2 + s1 * K_p - s2 * K_2;                            // :121-123
            qv = over * K_0 + s1_new * K_1 + s2_new * K_2;                             // :125-127
            snow = snow_new;
            soil = soil_new;
            s1 = s1_new;
            s2 = s2_new;
        }
        if (active) {
            if (q_o) st_stream(q_o + t * N, qv);
            if (STORAGE) {
                st_stream(out.snow + off + t * N, snow);
                st_stream(out.soil + off + t * N, soil);
                st_stream(out.s1 + off + t * N, s1);
                st_stream(out.s2 + off + t * N, s2);
            }
        }
        if (OBJ) {
            const double d = obj.qobs[t] - qv;
            acc += d * d;
        }
    });

    if (active) {
        if (slab.save_state) {
            slab.state[0 * N + i] = snow;
            slab.state[1 * N + i] = soil;
            slab.state[2 * N + i] = s1;
            slab.state[3 * N + i] = s2;
            if (OBJ) slab.state[4 * N + i] = acc;
        }
        if (OBJ && obj.mse && slab.t_end >= obj.T) obj.mse[i] = acc / (double)obj.T;
    }
}

int state_slots_hbvedu() { return 5; }

cudaError_t launch_hbvedu(const double* F, int64_t T, const double* inits4, const double* params, int64_t N,
                          double* qsim, double* snow, double* soil, double* s1, double* s2, const Slab& slab,
                          const Objective& obj, const LaunchCfg& cfg) {
    (void)T;
    if (N <= 0) return cudaSuccess;
    const int block = cfg.block > 0 ? cfg.block : pick_block(N, cfg.sm_count, 256);
    const unsigned grid = (unsigned)((N + block - 1) / block);
    const bool fast = cfg.math == RRB_MATH_FAST_;
    const size_t smem = forcing_smem_bytes<kHbvR, kHbvTT>() + (fast ? fastmath_smem_bytes() : 0);
    const bool st = snow != nullptr, ob = obj.qobs != nullptr;
    HbvOut out{qsim, snow, soil, s1, s2};
#define RRB_HBV(M_, S_, O_)                                                                                  \
    hbv_kernel<M_, S_, O_><<<grid, block, smem, cfg.stream>>>(F, inits4[0], inits4[1], inits4[2], inits4[3], \
                                                              params, N, out, slab, obj)
#define RRB_HBV_M(M_)                     \
    do {                                  \
        if (st && ob) RRB_HBV(M_, true, true);   \
        else if (st) RRB_HBV(M_, true, false);   \
        else if (ob) RRB_HBV(M_, false, true);   \
        else RRB_HBV(M_, false, false);          \
    } while (0)
    if (fast) RRB_HBV_M(RRB_MATH_FAST_);
    else RRB_HBV_M(RRB_MATH_PRECISE_);
#undef RRB_HBV_M
#undef RRB_HBV
    return cudaGetLastError();
}

}  // namespace rrb
