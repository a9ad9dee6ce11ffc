// temporary: multi-GPU entry points not built yet
#include "lbm_internal.hpp"
namespace lbm
{
int comm_release(lbm_domain*) { return LBM_OK; }
int comm_exchange(lbm_domain*, int) { return LBM_ERR_UNSUPPORTED; }
int comm_exchange_moments(lbm_domain*) { return LBM_OK; }
}
extern "C" {
int lbm_comm_unique_id(char*) { lbm::set_error("not built yet"); return LBM_ERR_UNSUPPORTED; }
int lbm_comm_init(lbm_domain*, const char*, int, int) { lbm::set_error("not built yet"); return LBM_ERR_UNSUPPORTED; }
int lbm_link_neighbours(lbm_domain*, lbm_domain*, lbm_domain*) { lbm::set_error("not built yet"); return LBM_ERR_UNSUPPORTED; }
}
