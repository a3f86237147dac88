// placeholder until the tcgen05 engine lands
#include "cfm_common.cuh"
namespace cfm {
bool attention_tc_supported(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, int, int, int) { return false; }
int attention_tc(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t, int64_t, void*,
                 int, int, int, int, const uint8_t*, int64_t, int64_t, const float*, float, int, cudaStream_t) {
  set_error("tcgen05 attention not built"); return -3; }
int attention_tc_init() { return 0; }
}
