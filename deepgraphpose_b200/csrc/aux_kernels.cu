// Bandwidth-class helper kernels around the tcgen05 conv GEMM (sm_100a).
//   prep_s2d      : u8 frames -> (pixel - mean_pixel) -> explicit conv2d_same padding (3 px) -> 2x2 space-to-depth,
//                   so that slim's 7x7 stride-2 conv1 becomes a stride-1 GEMM with K = 4 rows x 64.
//                   Reference: pose_net.py:38-40 (mean subtraction), slim resnet_utils.conv2d_same (pad 3/3 + VALID).
//   maxpool3x3s2  : slim.max_pool2d([3,3], stride=2, padding='SAME') of resnet_v1's root block.
//   deconv_col2im : scatter-free col2im of the 3x3 stride-2 transposed-conv heads (pose_net.py:18-26):
//                   out[2i+kh, 2j+kw] += x[i,j] * w[kh,kw], cropped to (2h, 2w), plus bias.
#include "kernels.cuh"

#include <cuda_fp16.h>

namespace dgp {

namespace {

__device__ __forceinline__ uint16_t to_half_bits(float f, int fp16) {
  if (fp16) { __half h = __float2half_rn(f); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 b = __float2bfloat16_rn(f);
  return *reinterpret_cast<uint16_t*>(&b);
}

__global__ void prep_s2d_kernel(const uint8_t* __restrict__ frames, int N, int H, int W, float m0, float m1, float m2,
                                uint16_t* __restrict__ out, int Hs, int Ws, int fp16) {
  const size_t total = (size_t)N * Hs * Ws;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % Ws);
    const size_t r = t / Ws;
    const int i = (int)(r % Hs);
    const int n = (int)(r / Hs);
    const float mean[3] = {m0, m1, m2};
    __align__(16) uint16_t v[16];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int y = 2 * i + u - 3;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int x = 2 * j + w - 3;
        const bool ok = (y >= 0) && (y < H) && (x >= 0) && (x < W);
        const uint8_t* px = frames + (((size_t)n * H + (ok ? y : 0)) * W + (ok ? x : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float f = ok ? (float)px[c] - mean[c] : 0.0f;
          v[(u * 2 + w) * 3 + c] = to_half_bits(f, fp16);
        }
      }
    }
#pragma unroll
    for (int c = 12; c < 16; ++c) v[c] = 0;
    uint4* dst = reinterpret_cast<uint4*>(out + t * 16);
    dst[0] = reinterpret_cast<const uint4*>(v)[0];
    dst[1] = reinterpret_cast<const uint4*>(v)[1];
  }
}

__device__ __forceinline__ uint32_t half2_max(uint32_t a, uint32_t b, int fp16) {
  if (fp16) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ in, int N, int H, int W, int C8, uint4* __restrict__ out,
                                    int Ho, int Wo, int pad_t, int pad_l, int fp16) {
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C8);
    size_t r = t / C8;
    const int q = (int)(r % Wo);
    r /= Wo;
    const int p = (int)(r % Ho);
    const int n = (int)(r / Ho);
    uint4 m;
    bool have = false;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = 2 * p - pad_t + dy;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int x = 2 * q - pad_l + dx;
        if (x < 0 || x >= W) continue;
        const uint4 v = __ldg(in + (((size_t)n * H + y) * W + x) * C8 + c);
        if (!have) {
          m = v;
          have = true;
        } else {
          m.x = half2_max(m.x, v.x, fp16); m.y = half2_max(m.y, v.y, fp16);
          m.z = half2_max(m.z, v.z, fp16); m.w = half2_max(m.w, v.w, fp16);
        }
      }
    }
    out[t] = m;
  }
}

__global__ void deconv_col2im_kernel(const float* __restrict__ contrib, int N, int h, int w, int ldn, int ctot, int nj,
                                     const float* __restrict__ bias, float* __restrict__ logits,
                                     float* __restrict__ locref) {
  const int Ho = 2 * h, Wo = 2 * w;
  const size_t total = (size_t)N * Ho * Wo * ctot;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(t % ctot);
    size_t r = t / ctot;
    const int x = (int)(r % Wo);
    r /= Wo;
    const int y = (int)(r % Ho);
    const int n = (int)(r / Ho);
    // y = 2i + kh: even y <- (i=y/2, kh=0) and (i=y/2-1, kh=2); odd y <- (i=(y-1)/2, kh=1).  Same for x.
    int is[2], khs[2], ny = 0;
    if (y & 1) { is[0] = y >> 1; khs[0] = 1; ny = 1; }
    else {
      is[0] = y >> 1; khs[0] = 0; ny = 1;
      if ((y >> 1) >= 1) { is[1] = (y >> 1) - 1; khs[1] = 2; ny = 2; }
    }
    int js[2], kws[2], nx = 0;
    if (x & 1) { js[0] = x >> 1; kws[0] = 1; nx = 1; }
    else {
      js[0] = x >> 1; kws[0] = 0; nx = 1;
      if ((x >> 1) >= 1) { js[1] = (x >> 1) - 1; kws[1] = 2; nx = 2; }
    }
    float acc = bias ? bias[co] : 0.0f;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) {
        const size_t m = ((size_t)n * h + is[a]) * w + js[b];
        acc += contrib[m * ldn + (khs[a] * 3 + kws[b]) * ctot + co];
      }
    if (co < nj) logits[(((size_t)n * Ho + y) * Wo + x) * nj + co] = acc;
    else if (locref) locref[(((size_t)n * Ho + y) * Wo + x) * (size_t)(ctot - nj) + (co - nj)] = acc;
  }
}

// evaluate_dgp's 'dgp' locref read-out (src/deepgraphpose/models/eval.py:751-785): with st the blurred spatial softmax,
//   soft = sum st * (row, col) * stride + stride/2;  offset = sum st * (locref[..., 2j], locref[..., 2j+1]) * locref_stdev
//   pose = (soft + offset)[::-1] -> (x, y, 1).  The reference adds locref's first component (DLC's dx) to the ROW
//   coordinate (no [::-1] as in argmax_pose_predict); swap_offsets = 1 applies the offsets the DLC way instead.
// One CTA per (frame, joint); fixed-order block reduction.
__global__ void __launch_bounds__(256) soft_pose_kernel(const float* __restrict__ st, const float* __restrict__ locref, int H,
                                                        int W, int nj, float stride, float locref_stdev, int swap_offsets,
                                                        float* __restrict__ pose) {
  const int b = blockIdx.x / nj, j = blockIdx.x - b * nj;
  const float* s0 = st + (size_t)b * H * W * nj + j;
  const float* l0 = locref + (size_t)b * H * W * 2 * nj + 2 * j;
  float ar = 0.0f, ac = 0.0f, o0 = 0.0f, o1 = 0.0f;
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    const float v = s0[(size_t)p * nj];
    const float2 l = *reinterpret_cast<const float2*>(l0 + (size_t)p * 2 * nj);
    ar += v * (float)r;
    ac += v * (float)c;
    o0 += v * (l.x * locref_stdev);
    o1 += v * (l.y * locref_stdev);
  }
  __shared__ float sm[4][256];
  sm[0][threadIdx.x] = ar; sm[1][threadIdx.x] = ac; sm[2][threadIdx.x] = o0; sm[3][threadIdx.x] = o1;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float row = sm[0][0] * stride + 0.5f * stride, col = sm[1][0] * stride + 0.5f * stride;
    const float orow = swap_offsets ? sm[3][0] : sm[2][0], ocol = swap_offsets ? sm[2][0] : sm[3][0];
    float* o = pose + (size_t)blockIdx.x * 3;
    o[0] = col + ocol;
    o[1] = row + orow;
    o[2] = 1.0f;
  }
}

int grid_for(size_t total, int threads) {
  size_t g = (total + threads - 1) / threads;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

cudaError_t launch_prep_s2d(const uint8_t* frames, int N, int H, int W, const float* mean3, __nv_bfloat16* out,
                            int Hs, int Ws, int fp16, cudaStream_t stream) {
  const size_t total = (size_t)N * Hs * Ws;
  prep_s2d_kernel<<<grid_for(total, 256), 256, 0, stream>>>(frames, N, H, W, mean3[0], mean3[1], mean3[2],
                                                            reinterpret_cast<uint16_t*>(out), Hs, Ws, fp16);
  return cudaGetLastError();
}

cudaError_t launch_maxpool3x3s2(const __nv_bfloat16* in, int N, int H, int W, int C, __nv_bfloat16* out, int Ho,
                                int Wo, int pad_t, int pad_l, int fp16, cudaStream_t stream) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t total = (size_t)N * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(in), N, H, W, C / 8,
                                                                reinterpret_cast<uint4*>(out), Ho, Wo, pad_t, pad_l, fp16);
  return cudaGetLastError();
}

cudaError_t launch_deconv_col2im(const float* contrib, int N, int h, int w, int ldn, int ctot, int nj,
                                 const float* bias, float* logits, float* locref, cudaStream_t stream) {
  const size_t total = (size_t)N * 4 * h * w * ctot;
  deconv_col2im_kernel<<<grid_for(total, 256), 256, 0, stream>>>(contrib, N, h, w, ldn, ctot, nj, bias, logits, locref);
  return cudaGetLastError();
}

// argmax_2d_from_cm(th=...) (src/deepgraphpose/models/fitdgp_util.py:379-399): on the blurred, renormalised softmax map
// zero every entry below th * max, renormalise by the new sum, and take E[(row, col)].  One CTA per (frame, joint), three
// fixed-order passes over the joint's H*W entries (stride nj in the NHWC map); the map is rewritten in place.
__global__ void __launch_bounds__(256) softmax_threshold_kernel(float* __restrict__ map, int H, int W, int nj, float th,
                                                                float* __restrict__ mu) {
  __shared__ float red[3][8];
  __shared__ float bc[2];
  const int b = blockIdx.x / nj, j = blockIdx.x - b * nj;
  float* p = map + (size_t)b * H * W * nj + j;
  const int n = H * W, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = 0.0f;   // entries are >= 0 (or NaN, which the reference's reduce_max also skips only by accident)
  for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, p[(size_t)i * nj]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[0][warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0][0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[0][w]);
    bc[0] = m * th;
  }
  __syncthreads();
  const float cut = bc[0];
  float s0 = 0.0f, sr = 0.0f, sc = 0.0f;
  for (int i = threadIdx.x; i < n; i += 256) {
    float v = p[(size_t)i * nj];
    if (v < cut) v = 0.0f;
    const int r = i / W, c = i - r * W;
    s0 += v; sr += v * (float)r; sc += v * (float)c;
  }
  for (int o = 16; o; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
  }
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = sr; red[2][warp] = sc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, r = 0.0f, c = 0.0f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; r += red[1][w]; c += red[2][w]; }
    bc[1] = a;
    if (mu) { mu[((size_t)b * nj + j) * 2] = r / a; mu[((size_t)b * nj + j) * 2 + 1] = c / a; }
  }
  __syncthreads();
  const float inv = 1.0f / bc[1];
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = p[(size_t)i * nj];
    p[(size_t)i * nj] = v < cut ? 0.0f : v * inv;
  }
}

cudaError_t launch_softmax_threshold(float* map, int B, int H, int W, int nj, float th, float* mu, cudaStream_t stream) {
  softmax_threshold_kernel<<<B * nj, 256, 0, stream>>>(map, H, W, nj, th, mu);
  return cudaGetLastError();
}

__global__ void cvt16_to_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, size_t n, int fp16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint16_t v = in[i];
    out[i] = fp16 ? __half2float(*reinterpret_cast<const __half*>(&v)) : __uint_as_float((uint32_t)v << 16);
  }
}

__global__ void f32_to_cvt16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, size_t n, int fp16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float f = in[i];
    if (fp16) f = fminf(fmaxf(f, -65504.0f), 65504.0f);
    out[i] = to_half_bits(f, fp16);
  }
}

cudaError_t launch_cvt16_to_f32(const void* in, float* out, size_t n, int fp16, cudaStream_t stream) {
  cvt16_to_f32_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const uint16_t*)in, out, n, fp16);
  return cudaGetLastError();
}

cudaError_t launch_f32_to_cvt16(const float* in, void* out, size_t n, int fp16, cudaStream_t stream) {
  f32_to_cvt16_kernel<<<grid_for(n, 256), 256, 0, stream>>>(in, (uint16_t*)out, n, fp16);
  return cudaGetLastError();
}

cudaError_t launch_soft_pose(const float* st, const float* locref, int B, int H, int W, int nj, float stride,
                             float locref_stdev, int swap_offsets, float* pose, cudaStream_t stream) {
  soft_pose_kernel<<<B * nj, 256, 0, stream>>>(st, locref, H, W, nj, stride, locref_stdev, swap_offsets, pose);
  return cudaGetLastError();
}

}  // namespace dgp
