// gemm_tc.cu -- tensor-core (tcgen05) engine overrides.  Placeholder: the SIMT engine runs everything.
#include "engine.cuh"
namespace oar {
bool tc_try_conv(oar_model*, int, const OpRec&, const Tensor&, Tensor&, int, int) { return false; }
bool tc_try_ctc_head(oar_model*, int, const OpRec&, const Tensor&, bool, Tensor&, CtcOut*) { return false; }
void tc_model_init(oar_model*) {}
void tc_model_free(oar_model*) {}
}  // namespace oar
