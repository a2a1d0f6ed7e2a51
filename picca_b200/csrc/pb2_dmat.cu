// Distortion matrix -- placeholder until the kernels land.
#include "pb2_common.cuh"

extern "C" {
int64_t pb2_dmat_scratch_bytes(const pb2_catalog *, const pb2_catalog *, const pb2_params *, int32_t)
{
    return 0;
}
int32_t pb2_dmat_auto(const pb2_catalog *, const pb2_catalog *, const pb2_params *,
                      const pb2_pairs *, double *, double *, double *, double *, double *,
                      double *, void *, int64_t, void *)
{
    pb2_set_error("pb2_dmat_auto: not implemented yet");
    return PB2_ECONFIG;
}
int32_t pb2_dmat_cross(const pb2_catalog *, const pb2_catalog *, const pb2_params *,
                       const pb2_pairs *, double *, double *, double *, double *, double *,
                       double *, void *, int64_t, void *)
{
    pb2_set_error("pb2_dmat_cross: not implemented yet");
    return PB2_ECONFIG;
}
}
