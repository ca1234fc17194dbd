// selection_binary.cu -- placeholder until the var-binary kernels land (next commit)
#include "bitmap.h"
namespace b2 {
int filter_binary(B2Context*, const B2Array*, const B2Array*, int, B2Array*, cudaStream_t) {
  return set_error(B2_NOT_IMPLEMENTED, "filter on binary types not yet implemented");
}
int take_binary(B2Context*, const B2Array*, const B2Array*, int, B2Array*, cudaStream_t) {
  return set_error(B2_NOT_IMPLEMENTED, "take on binary types not yet implemented");
}
}  // namespace b2
