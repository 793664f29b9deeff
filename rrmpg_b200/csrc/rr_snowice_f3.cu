// rr_snowice_f3.cu -- instantiations of the snow-ice family 3 kernels (see rr_snowice.cu); one translation unit per
// family so the three compile in parallel.
#include "rr_cemaneige.cuh"

namespace rrb {

cudaError_t launch_snowice_f3(const CemaArgs& a, double x4_max, const CemaOut& out, const Slab& slab, const Objective& obj,
                               const LaunchCfg& cfg) {
    return cema_launch_coupled<3>(a, x4_max, out, slab, obj, cfg);
}

}  // namespace rrb
