// fp64_probe.cu -- measures the fp64 pipe of the GPU the kernels run on (the second roofline of the
// recurrence kernels): dependent-issue latency of DFMA/DADD/DMUL, and DFMA throughput as a function of
// resident warps per SM sub-partition.  Build: nvcc -arch=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, double a, double b, int iters, long long* cycles) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void dadd_chain(double* out, double a, int iters, long long* cycles) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = x + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

// Half-masked warps: only lanes [0, active) run the chain (the others leave through the early return, so the warp
// issues every DFMA with a partial active mask).  If the 16-lane fp64 pipe skips the pass of an all-inactive half
// warp, 16 active lanes cost one issue slot instead of two -- the premise of the "half-masked tail warps" remedy for
// the 65 536-member load imbalance.
template <int ILP>
__global__ void dfma_chain_masked(double* out, double a, double b, int iters, int active) {
    if ((int)(threadIdx.x & 31) >= active) return;
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// Mixed issue: every thread runs NF independent DFMA chains and NI independent integer chains (LOP3 / IADD3 on the
// ALU pipe) or NS single-precision FFMA chains (FMA pipe) in the same loop body; asm volatile keeps the instruction mix
// exactly as written.  If an fp64 warp instruction HELD the sub-partition's issue port for two cycles, 4 DFMA + 4 INT
// per trip would need 12 cycles (0.67 warp-instructions per clock); if the fp64 pipe merely accepts one instruction
// every other cycle while the port stays free, it needs 8 (1.0 per clock).  (VERDICT r1 weak #2 / next #2a.)
template <int NF, int NI, int NS>
__global__ void mixed_chain(double* out, double a, double b, int ia, int iters) {
    double x[NF > 0 ? NF : 1];
    int n[NI > 0 ? NI : 1];
    float s[NS > 0 ? NS : 1];
#pragma unroll
    for (int k = 0; k < (NF > 0 ? NF : 1); ++k) x[k] = a + k + threadIdx.x;
#pragma unroll
    for (int k = 0; k < (NI > 0 ? NI : 1); ++k) n[k] = ia + k + threadIdx.x;
#pragma unroll
    for (int k = 0; k < (NS > 0 ? NS : 1); ++k) s[k] = (float)a + k + threadIdx.x;
    const float fa = (float)a, fb = (float)b;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // four rounds per trip: amortise the loop overhead
#pragma unroll
            for (int k = 0; k < (NF > NI ? (NF > NS ? NF : NS) : (NI > NS ? NI : NS)); ++k) {
                if (k < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[k]) : "d"(a), "d"(b));
                if (k < NI) {
                    if ((k + r) & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[k]) : "r"(ia));
                    else asm volatile("add.s32 %0, %0, %1;" : "+r"(n[k]) : "r"(ia));
                }
                if (k < NS) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[k]) : "f"(fa), "f"(fb));
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int k = 0; k < (NF > 0 ? NF : 1); ++k) acc += x[k];
#pragma unroll
    for (int k = 0; k < (NI > 0 ? NI : 1); ++k) acc += n[k];
#pragma unroll
    for (int k = 0; k < (NS > 0 ? NS : 1); ++k) acc += s[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NF, int NI, int NS>
static void run_mixed(double* out, int sms, int wps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = wps * 4 * 32, iters = 1 << 14;
    mixed_chain<NF, NI, NS><<<sms, threads>>>(out, 1.0000001, 1e-9, 3, 256);
    cudaEventRecord(e0);
    mixed_chain<NF, NI, NS><<<sms, threads>>>(out, 1.0000001, 1e-9, 3, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double per_trip = 4.0 * (NF + NI + NS);
    const double winst = (double)sms * (threads / 32) * iters * per_trip;
    const double ipc = winst / (ms * 1e-3) / sms / 4 / 1.965e9;
    printf("mixed %d DFMA + %d INT + %d FFMA chains, %2d warps/SMSP: %.3f warp-instr/clk/SMSP (fp64 %.3f, int %.3f, ffma %.3f) @1.965GHz\n",
           NF, NI, NS, wps, ipc, ipc * NF / (NF + NI + NS), ipc * NI / (NF + NI + NS), ipc * NS / (NF + NI + NS));
}

// The same mix with THREE DISTINCT REGISTER operands per DFMA (x = fma(x, y_k, z_k), y_k / z_k per chain and per thread)
// and two per integer instruction -- the operand pattern of real kernel code.  mixed_chain above feeds every DFMA one
// register, one reuse-cached register and one uniform-register operand, i.e. a third of the register-file traffic.
template <int NF, int NI>
__global__ void mixed_chain_regs(double* out, double a, double b, int ia, int iters) {
    double x[NF > 0 ? NF : 1], y[NF > 0 ? NF : 1], z[NF > 0 ? NF : 1];
    int n[NI > 0 ? NI : 1], m[NI > 0 ? NI : 1];
#pragma unroll
    for (int k = 0; k < (NF > 0 ? NF : 1); ++k) {
        x[k] = a + k + threadIdx.x;
        y[k] = a + 1e-9 * (k + threadIdx.x);
        z[k] = b * (1 + k + threadIdx.x);
        asm volatile("" : "+d"(y[k]), "+d"(z[k]));
    }
#pragma unroll
    for (int k = 0; k < (NI > 0 ? NI : 1); ++k) {
        n[k] = ia + k + threadIdx.x;
        m[k] = ia * 3 + k + 7 * threadIdx.x;
        asm volatile("" : "+r"(m[k]));
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < (NF > NI ? NF : NI); ++k) {
                if (k < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[k]) : "d"(y[k]), "d"(z[k]));
                if (k < NI) {
                    if ((k + r) & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[k]) : "r"(m[k]));
                    else asm volatile("add.s32 %0, %0, %1;" : "+r"(n[k]) : "r"(m[k]));
                }
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int k = 0; k < (NF > 0 ? NF : 1); ++k) acc += x[k];
#pragma unroll
    for (int k = 0; k < (NI > 0 ? NI : 1); ++k) acc += n[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NF, int NI>
static void run_mixed_regs(double* out, int sms, int wps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = wps * 4 * 32, iters = 1 << 14;
    mixed_chain_regs<NF, NI><<<sms, threads>>>(out, 1.0000001, 1e-9, 3, 256);
    cudaEventRecord(e0);
    mixed_chain_regs<NF, NI><<<sms, threads>>>(out, 1.0000001, 1e-9, 3, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)sms * (threads / 32) * iters * 4.0 * (NF + NI);
    const double ipc = winst / (ms * 1e-3) / sms / 4 / 1.965e9;
    printf("register operands: %d DFMA + %d INT chains, %2d warps/SMSP: %.3f warp-instr/clk/SMSP (fp64 %.3f, int %.3f); "
           "2*fp64 + int issue model predicts %.3f\n", NF, NI, wps, ipc, ipc * NF / (NF + NI), ipc * NI / (NF + NI),
           (double)(NF + NI) / (2 * NF + NI));
}

int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024 * 8);
    cudaMalloc(&cyc, 8);
    const int iters = 1 << 16;
    // latency: 1 warp, 1 chain
    dfma_chain<1><<<1, 32>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency       : %.2f cycles\n", (double)h / iters);
    dadd_chain<<<1, 32>>>(out, 1e-9, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DADD dependent latency       : %.2f cycles\n", (double)h / iters);
    dfma_chain<2><<<1, 32>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 warp, ILP 2: cycles/DFMA   : %.2f\n", (double)h / iters / 2);
    dfma_chain<8><<<1, 32>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 warp, ILP 8: cycles/DFMA   : %.2f   (issue interval of one warp)\n", (double)h / iters / 8);
    // throughput: all SMs, W warps per sub-partition, ILP 4
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int wps : {1, 2, 3, 4, 6, 8, 16}) {
        const int threads = wps * 4 * 32;  // one CTA per SM
        dfma_chain<4><<<sms, threads>>>(out, 1.0000001, 1e-9, 1024, cyc);
        cudaEventRecord(e0);
        dfma_chain<4><<<sms, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double dfma = (double)sms * threads * iters * 4;
        printf("%2d warps/SMSP ILP4: %.2f TDFMA/s (%.1f TFLOP/s), %.3f DFMA/clk/SM @1.965GHz\n", wps, dfma / ms / 1e9,
               2 * dfma / ms / 1e9, dfma / (ms * 1e-3) / sms / 1.965e9);
    }
    for (int wps : {1, 2, 4, 8}) {
        const int threads = wps * 4 * 32;
        cudaEventRecord(e0);
        dfma_chain<1><<<sms, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double dfma = (double)sms * threads * iters;
        printf("%2d warps/SMSP ILP1: %.2f TDFMA/s, %.3f DFMA/clk/SM\n", wps, dfma / ms / 1e9,
               dfma / (ms * 1e-3) / sms / 1.965e9);
    }
    // issue cost of partially active warps: 4 warps per sub-partition, ILP 4 (issue bound), warp-instructions per clock
    for (int active : {32, 24, 17, 16, 8, 1}) {
        const int threads = 4 * 4 * 32;
        dfma_chain_masked<4><<<sms, threads>>>(out, 1.0000001, 1e-9, 1024, active);
        cudaEventRecord(e0);
        dfma_chain_masked<4><<<sms, threads>>>(out, 1.0000001, 1e-9, iters, active);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double winst = (double)sms * (threads / 32) * iters * 4;
        printf("%2d active lanes per warp: %.3f warp-DFMA/clk/SMSP @1.965GHz (0.5 = two passes per instruction)\n", active,
               winst / (ms * 1e-3) / sms / 4 / 1.965e9);
    }
    for (int wps : {4, 8}) {
        run_mixed<4, 0, 0>(out, sms, wps);
        run_mixed<0, 4, 0>(out, sms, wps);
        run_mixed<0, 8, 0>(out, sms, wps);
        run_mixed<0, 0, 4>(out, sms, wps);
        run_mixed<4, 4, 0>(out, sms, wps);
        run_mixed<4, 8, 0>(out, sms, wps);
        run_mixed<2, 4, 0>(out, sms, wps);
        run_mixed<4, 0, 4>(out, sms, wps);
        run_mixed<4, 4, 4>(out, sms, wps);
        run_mixed<2, 4, 4>(out, sms, wps);
    }
    for (int wps : {2, 4, 8}) {
        run_mixed_regs<4, 0>(out, sms, wps);
        run_mixed_regs<0, 4>(out, sms, wps);
        run_mixed_regs<4, 4>(out, sms, wps);
        run_mixed_regs<4, 8>(out, sms, wps);
        run_mixed_regs<2, 4>(out, sms, wps);
        run_mixed_regs<4, 2>(out, sms, wps);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
