// d2h_probe.cu -- what the box's host link gives the end-to-end path: concurrent pinned device-to-host copies on
// 1..G GPUs, 1/2/4 streams per device, 128 MB slabs (the shape of run_job's slab ring, rr_api.cu).  The aggregate
// GB/s per device count is the ceiling `e2e` is reported against.
// Build: nvcc -arch=sm_100a -O3 -o d2h_probe d2h_probe.cu ; usage: d2h_probe [max_devices] [slab_MB] [slabs_per_device] [d2h-only]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const int max_dev = argc > 1 ? atoi(argv[1]) : ndev;
    if (ndev > max_dev) ndev = max_dev;
    const size_t slab = (size_t)(argc > 2 ? atoi(argv[2]) : 128) << 20;
    const int slabs = argc > 3 ? atoi(argv[3]) : 16;  // 2 GB per device per measurement
    const bool quick = argc > 4;  // "d2h-only": default pinned memory, device-to-host only, one stream per device
    const int kMaxStreams = 4;
    printf("devices %d, slab %zu MB, %d slabs per device per measurement\n", ndev, slab >> 20, slabs);
    std::vector<char*> dbuf(ndev);
    std::vector<std::vector<cudaStream_t>> st(ndev, std::vector<cudaStream_t>(kMaxStreams));
    for (int d = 0; d < ndev; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaMalloc(&dbuf[d], slab * kMaxStreams));
        CK(cudaMemset(dbuf[d], 1, slab * kMaxStreams));
        for (int s = 0; s < kMaxStreams; ++s) CK(cudaStreamCreateWithFlags(&st[d][s], cudaStreamNonBlocking));
    }
    for (unsigned flags : {(unsigned)cudaHostAllocDefault, (unsigned)cudaHostAllocPortable}) {
        if (quick && flags != cudaHostAllocDefault) break;
        // one host buffer per device, large enough that every slab lands on fresh pages (no cache-resident target)
        std::vector<char*> hbuf(ndev);
        for (int d = 0; d < ndev; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaHostAlloc((void**)&hbuf[d], slab * slabs, flags));
            for (size_t o = 0; o < slab * slabs; o += 4096) hbuf[d][o] = 0;  // touch
        }
        for (int dir = 0; dir < (quick ? 1 : 2); ++dir)
            for (int g = quick ? ndev : 1; g <= ndev; g *= 2)
                for (int ns : {1, 2, 4}) {
                    if (quick && ns != 1) break;
                    double best = 0;
                    for (int rep = 0; rep < 3; ++rep) {
                        for (int d = 0; d < g; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
                        auto t0 = std::chrono::steady_clock::now();
                        for (int k = 0; k < slabs; ++k)
                            for (int d = 0; d < g; ++d) {
                                CK(cudaSetDevice(d));
                                char* h = hbuf[d] + (size_t)k * slab;
                                char* v = dbuf[d] + (size_t)(k % ns) * slab;
                                if (dir == 0) CK(cudaMemcpyAsync(h, v, slab, cudaMemcpyDeviceToHost, st[d][k % ns]));
                                else CK(cudaMemcpyAsync(v, h, slab, cudaMemcpyHostToDevice, st[d][k % ns]));
                            }
                        for (int d = 0; d < g; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
                        auto t1 = std::chrono::steady_clock::now();
                        const double s = std::chrono::duration<double>(t1 - t0).count();
                        const double gbs = (double)slab * slabs * g / s / 1e9;
                        if (gbs > best) best = gbs;
                    }
                    printf("%s %s: %d device(s) x %d stream(s): %7.1f GB/s aggregate, %6.1f GB/s per device\n",
                           flags == cudaHostAllocPortable ? "portable" : "default ", dir == 0 ? "D2H" : "H2D", g, ns, best, best / g);
                }
        for (int d = 0; d < ndev; ++d) { CK(cudaSetDevice(d)); CK(cudaFreeHost(hbuf[d])); }
    }
    // a plain host memcpy of the same size: what one core moves (the numpy-side copy some callers add afterwards)
    if (!quick) {
        char* a = (char*)malloc(slab * 4); char* b = (char*)malloc(slab * 4);
        for (size_t o = 0; o < slab * 4; o += 4096) { a[o] = 1; b[o] = 0; }
        auto t0 = std::chrono::steady_clock::now();
        memcpy(b, a, slab * 4);
        auto t1 = std::chrono::steady_clock::now();
        printf("host memcpy, one thread: %.1f GB/s (check %d)\n", (double)slab * 4 / std::chrono::duration<double>(t1 - t0).count() / 1e9, b[4096]);
        free(a); free(b);
    }
    return 0;
}
