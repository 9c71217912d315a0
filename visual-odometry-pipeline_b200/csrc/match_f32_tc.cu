// placeholder until the tcgen05 kernel lands (replaced in this round)
#include "common.cuh"
namespace vo {
int match_f32_tc(vo_ctx *, const float *, const float *, int, int, int, const int32_t *, const int32_t *, int, int,
                 vo_row_partial **, int *, unsigned long long *, const float **, cudaStream_t) {
    set_error("vo_match_f32: tcgen05 path not built into this library");
    return VO_ERR_UNSUPPORTED;
}
}  // namespace vo
