// lambda (+)= H psi through shared-memory tiles (placeholder until the tile kernel lands: reports "not handled").
#include "common.cuh"

namespace qc {

int hpsi_tiled(qc_state* h, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc, bool accumulate,
               bool* handled) {
    (void)h, (void)T, (void)hx, (void)hz, (void)hc, (void)accumulate;
    *handled = false;
    return 0;
}

}  // namespace qc
