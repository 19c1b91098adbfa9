#include "post.h"
namespace paa {
size_t post_workspace_bytes(int, int, int, int, int) { return 0; }
int run_postprocess(const Geometry&, const PaaPostArgs*, cudaStream_t) { set_error("not built yet"); return PAA_ERR_UNSUPPORTED; }
size_t ml_nms_workspace_bytes(int) { return 0; }
int run_ml_nms(const float*, const float*, const float*, int, float, uint8_t*, int32_t*, void*, size_t, cudaStream_t) { set_error("not built yet"); return PAA_ERR_UNSUPPORTED; }
}
