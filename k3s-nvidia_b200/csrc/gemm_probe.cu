// placeholder — replaced by the tcgen05 kernel
#include <cuda_runtime.h>
#include "common.h"
extern "C" {
int b200probe_gemm(int, const b200probe_gemm_cfg_t*, b200probe_gemm_result_t*) { b200::set_error("gemm probe not built yet"); return B200PROBE_ESTATE; }
int b200probe_gemm_launch(int, const void*, const void*, void*, int, int, int, void*) { b200::set_error("gemm probe not built yet"); return B200PROBE_ESTATE; }
int b200probe_gemm_fill(int, void*, uint64_t, uint32_t, int, void*) { b200::set_error("gemm probe not built yet"); return B200PROBE_ESTATE; }
}
