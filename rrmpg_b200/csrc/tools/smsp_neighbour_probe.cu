// smsp_neighbour_probe.cu -- does a warp that has its SM sub-partition (SMSP) to itself slow down when the OTHER
// sub-partitions of the SM are busy?  (Premise of rotating member pairs between sub-partitions, DESIGN.md section 11:
// hbv_rot_kernel measured the lone warp of a (2,2,2,1) layout at 342 cycles per timestep, against 285 when every
// sub-partition of the SM holds one warp.)
// One CTA of 8 warps per SM (warp w -> SMSP w % 4).  The "victim" is warp 3 (alone on SMSP 3): a latency-bound fp64
// chain mix of ILP 2, timed with clock64.  The six other warps (two per SMSP 0..2) run one of: nothing, an ILP-4 DFMA
// stream, random 16-byte shared-memory loads, streaming global stores, or the victim's own mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o smsp_neighbour_probe smsp_neighbour_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) probe(int mode, int victim_kind, int iters, double a, double b, double* sink,
                                                long long* cycles, double* gbuf) {
    extern __shared__ double2 table[];  // 4096 entries of 16 bytes
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) table[i] = make_double2(1.0 + i * 1e-9, 1e-9 * i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double x0 = a + lane * 1e-3, x1 = b + lane * 1e-3, x2 = a - lane * 1e-3, x3 = b - lane * 1e-3;
    unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
    if (warp == 3) {  // victim
        const long long c0 = clock64();
        if (victim_kind == 0) {  // two dependent DFMA chains
            for (int i = 0; i < iters; ++i) {
                x0 = fma(x0, a, b);
                x1 = fma(x1, a, b);
            }
        } else {  // two chains, each: random 16-byte table load -> 4 dependent DFMAs (the shape of the HBV pow)
            for (int i = 0; i < iters / 4; ++i) {
                const double2 e0 = table[(__double2loint(x0) ^ h) & 4095];
                const double2 e1 = table[(__double2loint(x1) ^ (h >> 3)) & 4095];
                x0 = fma(x0, e0.x, e0.y); x1 = fma(x1, e1.x, e1.y);
                x0 = fma(x0, a, b); x1 = fma(x1, a, b);
                x0 = fma(x0, a, b); x1 = fma(x1, a, b);
                x0 = fma(x0, a, b); x1 = fma(x1, a, b);
            }
        }
        const long long c1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = c1 - c0;
    } else if (warp != 7 && mode != 0) {  // aggressors: warps 0,1,2,4,5,6
        const int n = iters * 2;  // keep them busy for longer than the victim runs
        if (mode == 1) {
            for (int i = 0; i < n; ++i) {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            }
        } else if (mode == 2) {
            for (int i = 0; i < n / 2; ++i) {
                const double2 e = table[h & 4095];
                h = h * 1664525u + 1013904223u + (unsigned)__double2loint(e.x);
                x0 += e.y;
            }
        } else if (mode == 3) {
            double* p = gbuf + ((size_t)blockIdx.x * 256 + threadIdx.x);
            for (int i = 0; i < n / 8; ++i) __stcs(p + (size_t)(i & 1023) * 148 * 256, x0 + i);
        } else {
            for (int i = 0; i < n / 4; ++i) {
                const double2 e0 = table[(__double2loint(x0) ^ h) & 4095];
                const double2 e1 = table[(__double2loint(x1) ^ (h >> 3)) & 4095];
                x0 = fma(x0, e0.x, e0.y); x1 = fma(x1, e1.x, e1.y);
                x0 = fma(x0, a, b); x1 = fma(x1, a, b);
                x0 = fma(x0, a, b); x1 = fma(x1, a, b);
                x0 = fma(x0, a, b); x1 = fma(x1, a, b);
            }
        }
    }
    if (x0 + x1 + x2 + x3 == 1.2345) sink[threadIdx.x] = x0;
}

int main() {
    const int iters = 1 << 16;
    long long* cycles;
    double *sink, *gbuf;
    cudaMalloc(&cycles, 148 * sizeof(long long));
    cudaMalloc(&sink, 4096);
    cudaMalloc(&gbuf, (size_t)1024 * 148 * 256 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    const char* names[] = {"idle", "ILP-4 DFMA streams", "random 16-byte shared loads", "streaming global stores",
                           "the victim's own table + DFMA mix"};
    for (int vk = 0; vk < 2; ++vk)
        for (int mode = 0; mode < 5; ++mode) {
            probe<<<148, 256, 128 * 1024>>>(mode, vk, iters, 0.999999, 1e-7, sink, cycles, gbuf);
            cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, cycles, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int i = 0; i < 148; ++i) s += (double)h[i];
            printf("victim %-28s | six neighbours on the other three SMSPs: %-34s : %.2f cycles per victim fp64 instruction pair\n",
                   vk == 0 ? "2 dependent DFMA chains" : "table load + 4 DFMA, 2 chains", names[mode], s / 148 / iters);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
