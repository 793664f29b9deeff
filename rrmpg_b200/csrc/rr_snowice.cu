// rr_snowice.cu -- the snow-ice couplings of the reference:
//   CemaneigeGR4JIce      run_cemaneigegr4jice      rrmpg/models/cemaneigegr4jice_model.py:16-93     (family 2)
//   CemaneigeHystGR4J     run_cemaneigehystgr4j     rrmpg/models/cemaneigehystgr4j_model.py:17-79    (family 1)
//   CemaneigeHystGR4JIce  run_cemaneigehystgr4jice  rrmpg/models/cemaneigehystgr4jice_model.py:18-104 (family 3)
// replacing the member loops of cemaneigegr4jice.py:262-284, cemaneigehystgr4j.py:262-286 and
// cemaneigehystgr4jice.py:276-304.  The kernel is the FAMILY-templated cema_kernel of rr_cemaneige.cuh.
#include "rr_cemaneige.cuh"

namespace rrb {

cudaError_t launch_snowice_f1(const CemaArgs&, double, const CemaOut&, const Slab&, const Objective&, const LaunchCfg&);
cudaError_t launch_snowice_f2(const CemaArgs&, double, const CemaOut&, const Slab&, const Objective&, const LaunchCfg&);
cudaError_t launch_snowice_f3(const CemaArgs&, double, const CemaOut&, const Slab&, const Objective&, const LaunchCfg&);

int state_slots_snowice(int family, int L, double x4_max) {
    return ((family & 1) ? 4 : 2) * cema_layer_class(L) + cema_uh_slots(cema_uh_class(x4_max)) + kObjSlots;
}

cudaError_t launch_snowice(int family, const double* F, const double* g_tresh, const double* frac_ice, int64_t T, int L,
                           const double* inits5, const double* params, int64_t N, double x4_max, const SnowIceOut& o,
                           const Slab& slab, const Objective& obj, const LaunchCfg& cfg, const Batch& batch) {
    const int k = 6 + ((family & 1) ? 2 : 0) + ((family & 2) ? 1 : 0);
    CemaArgs a{F, g_tresh, L, T, inits5[0], inits5[1], inits5[2], inits5[3], inits5[4], params, k, N, frac_ice,
               forcing_flag(F, T, cema_TT(cema_layer_class(L)), cema_R(cema_layer_class(L))), batch.count,
               batch.forcing_stride, batch.inits, kSnowIceInitsStride};
    CemaOut out{o.qsim, o.G, o.eTG, o.s_store, o.r_store, o.sca, o.icemelt, o.snowmelt};
    switch (family) {
        case 1: return launch_snowice_f1(a, x4_max, out, slab, obj, cfg);
        case 2: return launch_snowice_f2(a, x4_max, out, slab, obj, cfg);
        case 3: return launch_snowice_f3(a, x4_max, out, slab, obj, cfg);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace rrb
