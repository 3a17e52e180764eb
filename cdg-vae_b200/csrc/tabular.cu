// Tabular CDG-VAE / CDG-TVAE step (tabular/modules/model.py:234-460, tabular/modules/train.py:173-320).
#include "latent.cuh"
#include "elementwise.cuh"

struct cdg_tabular_plan { cdg_tabular_config c; };

extern "C" int cdg_tabular_create(const cdg_tabular_config*, cdg_tabular_plan**) {
    cdg::set_error("tabular path not built yet");
    return CDG_ERR_UNSUPPORTED;
}
extern "C" void cdg_tabular_destroy(cdg_tabular_plan* p) { delete p; }
extern "C" int64_t cdg_tabular_workspace_bytes(const cdg_tabular_plan*, int64_t) { return -1; }
extern "C" int cdg_tabular_forward_backward(cdg_tabular_plan*, const cdg_tabular_io*, void*) {
    cdg::set_error("tabular path not built yet");
    return CDG_ERR_UNSUPPORTED;
}
extern "C" int cdg_tabular_forward(cdg_tabular_plan*, const cdg_tabular_io*, int32_t, void*) {
    cdg::set_error("tabular path not built yet");
    return CDG_ERR_UNSUPPORTED;
}
