// Continuous PCT environment (pct_envs/PctContinuous0 in the reference) — placeholder until the kernels land.
#include "pct_kernels.h"
#include "pct_handle.h"
namespace pct {
int continuous_create(pct_env_batch *h) { h->err = "continuous domain is not built yet"; return PCT_ERR_INVALID; }
void continuous_destroy(pct_env_batch *) {}
int continuous_launch(pct_env_batch *h, int, const void *, int, const int32_t *, void *, float *, uint8_t *, pct_step_info *, cudaStream_t) { return PCT_ERR_INVALID; }
int continuous_policy_random(pct_env_batch *, int32_t *, uint64_t, int64_t, cudaStream_t) { return PCT_ERR_INVALID; }
int continuous_get_state(pct_env_batch *, int, pct_state_dump *) { return PCT_ERR_INVALID; }
int64_t continuous_state_bytes() { return 0; }
}
