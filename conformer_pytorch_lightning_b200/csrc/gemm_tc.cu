// placeholder until the tcgen05 engine lands
#include "cfm_common.cuh"
namespace cfm {
bool gemm_tc_supported(int, int, int, int, int, int, int) { return false; }
int gemm_tc(const void*, int, const void*, const float*, void*, int, int, int, int, int, int, const float*, float,
            const uint8_t*, cudaStream_t) { set_error("tcgen05 gemm not built"); return -3; }
int gemm_tc_init() { return 0; }
}
