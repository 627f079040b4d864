// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace coma {
bool conv_tc_supported(const coma_conv_args&) { return false; }
int conv_tc_stat_chunks(const coma_conv_args&) { return 0; }
int conv_tc_launch(const coma_conv_args&, cudaStream_t) { set_error("tcgen05 path not built"); return COMA_ERR_UNSUPPORTED; }
}
