// dct_jsd_k3.cu -- register-tiled JSD instantiations for K = 3 views (one TU per K to
// keep build time parallel).  See dct_jsd_kernels.cuh.
#include "dct_jsd_kernels.cuh"

namespace dct {

int jsd_launch_k3(const JsdCall& c) {
    switch (c.C) {
        case 2: return jsd_launch_kc<3, 2>(c);
        case 3: return jsd_launch_kc<3, 3>(c);
        case 4: return jsd_launch_kc<3, 4>(c);
        case 19: return jsd_launch_kc<3, 19>(c);
        default: return DCT_ERR_UNSUPPORTED;
    }
}

}  // namespace dct
