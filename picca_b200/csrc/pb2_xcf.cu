// Forest x object correlation (xcf) -- placeholder until the kernel lands.
#include "pb2_common.cuh"

extern "C" {
int32_t pb2_xi_cross(const pb2_catalog *, const pb2_catalog *, const pb2_params *,
                     const pb2_pairs *, const int32_t *, int64_t, double *, int32_t, void *)
{
    pb2_set_error("pb2_xi_cross: not implemented yet");
    return PB2_ECONFIG;
}
}
