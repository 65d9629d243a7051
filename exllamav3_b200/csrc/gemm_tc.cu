// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace exl3b {
bool gemm_tc_supported(const GemmArgs&) { return false; }
int launch_gemm_tc(cudaStream_t, DevCtx*, const GemmArgs&) { return fail(EXL3B_ERR_UNSUPPORTED, "tcgen05 path not built"); }
bool hgemm_tc_supported(int, int, int, int64_t) { return false; }
int launch_hgemm_tc(cudaStream_t, const half*, const half*, void*, int, int, int, bool, int64_t) { return fail(EXL3B_ERR_UNSUPPORTED, "tcgen05 hgemm not built"); }
}
