// temporary: two-phase entry points not built yet
#include "lbm_internal.hpp"
namespace lbm
{
int tp_create(lbm_domain*) { set_error("two-phase models are not built yet"); return LBM_ERR_UNSUPPORTED; }
int tp_destroy(lbm_domain*) { return LBM_OK; }
int tp_step(lbm_domain*) { return LBM_ERR_UNSUPPORTED; }
int tp_commit(lbm_domain*) { return LBM_ERR_UNSUPPORTED; }
int tp_export(lbm_domain*) { return LBM_ERR_UNSUPPORTED; }
}
extern "C" {
int lbm_get_phase(lbm_domain*, double*, double*, double*) { lbm::set_error("not built yet"); return LBM_ERR_UNSUPPORTED; }
int lbm_set_u(lbm_domain*, const double*) { lbm::set_error("not built yet"); return LBM_ERR_UNSUPPORTED; }
int lbm_init_two_phase(lbm_domain*, const double*, const double*, const double*) { lbm::set_error("not built yet"); return LBM_ERR_UNSUPPORTED; }
}
