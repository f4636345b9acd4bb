// DXT5-YCoCg ("DXT6") block encoder — device side.  See dxt_device.cuh for the contract.
#pragma once
#include "dxt_device.cuh"

namespace ugb {
/// dxt_encode<6>, cuda_dxt.cu:471-509  (WORK IN PROGRESS)
__device__ __forceinline__ uint4 dxt6_encode(const float (&r)[16], const float (&g)[16], const float (&b)[16])
{
        return make_uint4(0, 0, 0, 0);
}
}  // namespace ugb
