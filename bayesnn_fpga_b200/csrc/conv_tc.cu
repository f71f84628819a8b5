// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
extern "C" int bnn_conv2d_tc(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int,
                             int, int, int, const bnn_drop_desc*, void*) {
  bnn::set_error("bnn_conv2d_tc: not built");
  return BNN_E_UNSUPPORTED;
}
