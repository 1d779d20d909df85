// K1 (tensor-core engine) — placeholder until the tcgen05 kernel lands.
#include "dr_common.cuh"
bool dr_tc_built() { return false; }
bool dr_tc_supported(const dr_model*, int, int) { return false; }
int dr_tc_prep_weights(dr_model*) { return DR_OK; }
int dr_launch_gru_tc(dr_model* m, const float*, int, int, float*, float*) {
    return dr_fail(m, DR_EUNSUPPORTED, "tcgen05 engine not built");
}
