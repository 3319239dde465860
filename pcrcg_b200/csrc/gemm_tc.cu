// tcgen05 tensor-core contraction (placeholder until the UMMA kernel lands: reports "not handled"
// so gemm_dev uses the fp32 CUDA-core kernel).
#include "common.cuh"

namespace pcrcg {

int gemm_tc_dev(const float*, int, const float*, int, int, float*, int, int, int, int, const float*, cudaStream_t, bool* handled)
{
    *handled = false;
    return PCRCG_OK;
}

}  // namespace pcrcg
