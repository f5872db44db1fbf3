// Placeholders for tcgen05 entry points that are not built yet: they fail loudly (never fall back).
#include "../../include/b200asr.h"
#include "attention.h"
#include "common.cuh"
#include "kernels.h"

namespace b200asr {
#ifndef B200ASR_HAVE_TC_ATTN
int sdpa_bwd_tc(const AttnP&, const float*, float*, float*, float*, float*, cudaStream_t) {
  set_error("tcgen05 attention backward is not available in this build");
  return B200ASR_BAD_ARG;
}
#endif

}  // namespace b200asr
