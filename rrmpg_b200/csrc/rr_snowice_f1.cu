// rr_snowice_f1.cu -- instantiations of the snow-ice family 1 kernels (see rr_snowice.cu); one translation unit per
// family so the three compile in parallel.
#include "rr_cemaneige.cuh"

namespace rrb {

cudaError_t launch_snowice_f1(const CemaArgs& a, double x4_max, const CemaOut& out, const Slab& slab, const Objective& obj,
                               const LaunchCfg& cfg) {
    return cema_launch_coupled<1>(a, x4_max, out, slab, obj, cfg);
}

}  // namespace rrb
