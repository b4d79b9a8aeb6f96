// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace ds2 {
int gemm_tc(int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int, void*, size_t,
            cudaStream_t) { return 1; }
size_t gemm_tc_workspace_bytes(int, int, int, int, int) { return 0; }
}
