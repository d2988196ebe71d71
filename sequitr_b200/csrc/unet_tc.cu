// placeholder until the tcgen05 path lands
#include "unet_plan.cuh"
int sq_tc_finalize(sq_unet_s *) { sq_set_error("bf16 tensor-core mode not built yet"); return SQ_EUNSUPPORTED; }
int sq_tc_destroy(sq_unet_s *) { return SQ_OK; }
int sq_tc_workspace_bytes(sq_unet_s *, int, int, int, int, size_t *) { return SQ_EUNSUPPORTED; }
int sq_tc_forward(sq_unet_s *, const float *, int, int, int, int, float *, uint8_t *, float *, void *, size_t, cudaStream_t) { return SQ_EUNSUPPORTED; }
