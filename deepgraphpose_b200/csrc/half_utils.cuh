// 16-bit storage helpers shared by the bandwidth-class kernels: the activation / weight storage type is bf16 (default) or
// fp16 (dgp_config.precision = 1), selected at run time by `fp16`; arithmetic is always fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dgp {
namespace h16 {

__device__ __forceinline__ void unpack8(const uint4& v, float* f, int fp16) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fp16) {
      const float2 r = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = r.x;
      f[2 * i + 1] = r.y;
    } else {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
// fp16 saturates at +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack2(float a, float b, int fp16) {
  if (fp16) {
    __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.0f), 65504.0f), fminf(fmaxf(b, -65504.0f), 65504.0f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float* f, int fp16) {
  return make_uint4(pack2(f[0], f[1], fp16), pack2(f[2], f[3], fp16), pack2(f[4], f[5], fp16), pack2(f[6], f[7], fp16));
}
__device__ __forceinline__ uint16_t cvt1(float a, int fp16) {
  if (fp16) {
    __half h = __float2half_rn(fminf(fmaxf(a, -65504.0f), 65504.0f));
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __nv_bfloat16 h = __float2bfloat16_rn(a);
  return *reinterpret_cast<uint16_t*>(&h);
}

}  // namespace h16
}  // namespace dgp
