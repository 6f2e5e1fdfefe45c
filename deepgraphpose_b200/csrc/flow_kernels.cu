// learn_wt on the GPU: the optical-flow magnitude field that weights the temporal clique of dgp_loss.
//
// Reference: src/deepgraphpose/models/fitdgp_util.py:454-467 -- per consecutive frame pair
//     flow = cv2.calcOpticalFlowFarneback(gray(prev), gray(next), None, 0.5, 3, 15, 3, 5, 1.2, 0);  field = |u| + |v|
// with gray = cv2.cvtColor(frame, COLOR_BGR2GRAY).  OpenCV is a third-party dependency of the reference (not vendored); this file
// restates the published algorithm of its dense Farneback flow (G. Farneback, "Two-frame motion estimation based on polynomial
// expansion", SCIA 2003; OpenCV modules/video/src/optflowgf.cpp) for exactly those arguments:
//   * pyramid of levels+1 = 4 scales (1/8 .. 1): the FULL-resolution gray image is smoothed with a Gaussian of
//     sigma = (1/scale - 1)/2 (kernel size max(round(5 sigma) | 1, 3), the fixed [1 2 1]/4 kernel at scale 1, BORDER_REFLECT_101),
//     then bilinearly resized (half-pixel centres) to round(size * scale);
//   * polynomial expansion (poly_n = 5 -> 11 separable taps of a sigma = 1.2 Gaussian with weights 1, x, x^2, replicated borders)
//     into the 5 coefficients (r3 = y, r2 = x, r5 = yy, r4 = xx, r6 = xy) through the inverse of the 6x6 moment matrix;
//   * per level, from the coarser level's flow (bilinearly upsampled, x2): UpdateMatrices (bilinear warp of the second frame's
//     expansion by the current flow, the 5 entries G11, G12, G22, h1, h2 per pixel, damped within 5 pixels of the border by
//     {0.14, 0.14, 0.4472, 0.4472, 0.4472}), then 3 iterations of { 15x15 box sum with replicated borders; flow = solve 2x2 with
//     +1e-3 on the determinant; UpdateMatrices unless last }.
// The numpy restatement of the same steps (tests/test_flow.py::farneback_numpy) agrees with cv2.calcOpticalFlowFarneback to
// 1e-5 px; this CUDA path is checked against both.  All frame pairs of a batch go through every stage in ONE launch (the
// pyramid and the polynomial expansion are computed once per FRAME, not once per pair).
#include <math.h>
#include <stdint.h>

#include "kernels.cuh"

namespace dgp {

namespace {

constexpr int kPolyN = 5;        // poly_n
constexpr int kWin = 15;         // winsize
constexpr int kIters = 3;
constexpr int kLevels = 3;       // -> 4 pyramid levels
constexpr int kMaxGauss = 19;    // largest smoothing kernel (scale 1/8: sigma 3.5)

struct PolyConsts {
  float g[2 * kPolyN + 1], xg[2 * kPolyN + 1], xxg[2 * kPolyN + 1];
  double ig11, ig03, ig33, ig55;
};
struct GaussK {
  float k[kMaxGauss];
  int size;
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i < 0 ? 0 : i;
}
__device__ __forceinline__ int clampi(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

// cv2.cvtColor(frame, COLOR_BGR2GRAY) of OpenCV 4.x on uint8: (c0 * 3735 + c1 * 19235 + c2 * 9798 + 2^14) >> 15, c0 = "B"
__global__ void gray_kernel(const uint8_t* __restrict__ frames, size_t npix, float* __restrict__ gray) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const uint8_t* p = frames + 3 * i;
    gray[i] = (float)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
  }
}

// separable Gaussian, BORDER_REFLECT_101: dir 0 = along x, 1 = along y
// (all per-pixel kernels below: grid = (ceil(W / 256), H, images) -- no 64-bit div / mod per pixel)
__global__ void gauss_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int H, int W, GaussK gk, int dir) {
  const int r = gk.size / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < W) {
    const float* img = src + (size_t)blockIdx.z * H * W;
    const size_t i = ((size_t)blockIdx.z * H + y) * W + x;
    float acc = 0.0f;
    for (int j = 0; j < gk.size; ++j) {
      const float v = dir == 0 ? img[(size_t)y * W + reflect101(x + j - r, W)] : img[(size_t)reflect101(y + j - r, H) * W + x];
      acc += gk.k[j] * v;
    }
    dst[i] = acc;
  }
}

// cv2.resize(..., INTER_LINEAR) on float data with C interleaved channels: source coordinate (d + 0.5) * src/dst - 0.5 in float,
// clamped taps; horizontal interpolation first, then vertical (both in float); the result is multiplied by `mul`.
__global__ void resize_linear_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int Hs, int Ws, int Hd, int Wd,
                                     int C, float mul) {
  const size_t total = (size_t)n * Hd * Wd * C;
  const double sx = (double)Ws / Wd, sy = (double)Hs / Hd;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int x = (int)((i / C) % Wd);
    const int y = (int)((i / ((size_t)C * Wd)) % Hd);
    const float* img = src + (i / ((size_t)C * Wd * Hd)) * (size_t)Hs * Ws * C;
    float fx = (float)((x + 0.5) * sx - 0.5);
    int x0 = (int)floorf(fx);
    fx -= (float)x0;
    if (x0 < 0) { x0 = 0; fx = 0.0f; }
    if (x0 >= Ws - 1) { x0 = Ws - 1; fx = 0.0f; }
    const int x1 = min(x0 + 1, Ws - 1);
    float fy = (float)((y + 0.5) * sy - 0.5);
    int y0 = (int)floorf(fy);
    fy -= (float)y0;
    if (y0 < 0) { y0 = 0; fy = 0.0f; }
    if (y0 >= Hs - 1) { y0 = Hs - 1; fy = 0.0f; }
    const int y1 = min(y0 + 1, Hs - 1);
    const float a = img[((size_t)y0 * Ws + x0) * C + c] * (1.0f - fx) + img[((size_t)y0 * Ws + x1) * C + c] * fx;
    const float b = img[((size_t)y1 * Ws + x0) * C + c] * (1.0f - fx) + img[((size_t)y1 * Ws + x1) * C + c] * fx;
    dst[i] = (a * (1.0f - fy) + b * fy) * mul;
  }
}

// polynomial expansion, vertical half: (sum g p, sum xg (below - above), sum xxg p) over the 11 rows, replicated borders
__global__ void polyexp_v_kernel(const float* __restrict__ src, float* __restrict__ tmp3, int n, int H, int W, PolyConsts pc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < W) {
    const float* img = src + (size_t)blockIdx.z * H * W;
    const size_t i = ((size_t)blockIdx.z * H + y) * W + x;
    float t0 = img[(size_t)y * W + x] * pc.g[kPolyN], t1 = 0.0f, t2 = 0.0f;
#pragma unroll
    for (int k = 1; k <= kPolyN; ++k) {
      const float a = img[(size_t)max(y - k, 0) * W + x], b = img[(size_t)min(y + k, H - 1) * W + x];
      const float p = a + b;
      t0 = t0 + pc.g[kPolyN + k] * p;
      t1 = t1 + pc.xg[kPolyN + k] * (b - a);
      t2 = t2 + pc.xxg[kPolyN + k] * p;
    }
    tmp3[3 * i] = t0;
    tmp3[3 * i + 1] = t1;
    tmp3[3 * i + 2] = t2;
  }
}

// ... horizontal half and the projection onto (y, x, yy, xx, xy) (double accumulators as in the library)
__global__ void polyexp_h_kernel(const float* __restrict__ tmp3, float* __restrict__ R, int n, int H, int W, PolyConsts pc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < W) {
    const size_t i = ((size_t)blockIdx.z * H + blockIdx.y) * W + x;
    const float* row = tmp3 + (i - x) * 3;
    double b1 = (double)row[3 * x] * pc.g[kPolyN], b2 = 0.0, b3 = (double)row[3 * x + 1] * pc.g[kPolyN], b4 = 0.0,
           b5 = (double)row[3 * x + 2] * pc.g[kPolyN], b6 = 0.0;
#pragma unroll
    for (int k = 1; k <= kPolyN; ++k) {
      const float* rp = row + 3 * clampi(x + k, 0, W - 1);
      const float* rm = row + 3 * clampi(x - k, 0, W - 1);
      const double tg = (double)rp[0] + (double)rm[0];
      b1 += tg * pc.g[kPolyN + k];
      b4 += tg * pc.xxg[kPolyN + k];
      b2 += ((double)rp[0] - (double)rm[0]) * pc.xg[kPolyN + k];
      b3 += ((double)rp[1] + (double)rm[1]) * pc.g[kPolyN + k];
      b6 += ((double)rp[1] - (double)rm[1]) * pc.xg[kPolyN + k];
      b5 += ((double)rp[2] + (double)rm[2]) * pc.g[kPolyN + k];
    }
    float* o = R + 5 * i;
    o[1] = (float)(b2 * pc.ig11);
    o[0] = (float)(b3 * pc.ig11);
    o[3] = (float)(b1 * pc.ig03 + b4 * pc.ig33);
    o[2] = (float)(b1 * pc.ig03 + b5 * pc.ig33);
    o[4] = (float)(b6 * pc.ig55);
  }
}

// UpdateMatrices for every pair p: R0 = R[p], R1 = R[p + 1]
__global__ void update_matrices_kernel(const float* __restrict__ R, const float* __restrict__ flow, float* __restrict__ M, int np,
                                       int H, int W) {
  const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
  const size_t plane = (size_t)H * W;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < W) {
    const size_t p = blockIdx.z;
    const size_t i = p * plane + (size_t)y * W + x;
    const float* R0 = R + (p * plane + (size_t)y * W + x) * 5;
    const float* R1 = R + (p + 1) * plane * 5;
    const float dx = flow[2 * i], dy = flow[2 * i + 1];
    float fx = (float)x + dx, fy = (float)y + dy;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1;
    fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1)) {
      const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
      const float* q = R1 + ((size_t)y1 * W + x1) * 5;
      const float* qd = q + (size_t)W * 5;
      r2 = a00 * q[0] + a01 * q[5] + a10 * qd[0] + a11 * qd[5];
      r3 = a00 * q[1] + a01 * q[6] + a10 * qd[1] + a11 * qd[6];
      r4 = a00 * q[2] + a01 * q[7] + a10 * qd[2] + a11 * qd[7];
      r5 = a00 * q[3] + a01 * q[8] + a10 * qd[3] + a11 * qd[8];
      r6 = a00 * q[4] + a01 * q[9] + a10 * qd[4] + a11 * qd[9];
      r4 = (R0[2] + r4) * 0.5f;
      r5 = (R0[3] + r5) * 0.5f;
      r6 = (R0[4] + r6) * 0.25f;
    } else {
      r2 = r3 = 0.f;
      r4 = R0[2];
      r5 = R0[3];
      r6 = R0[4] * 0.5f;
    }
    r2 = (R0[0] - r2) * 0.5f;
    r3 = (R0[1] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(W - 10) || (unsigned)(y - 5) >= (unsigned)(H - 10)) {
      const float scale = (x < 5 ? border[x] : 1.f) * (x >= W - 5 ? border[W - x - 1] : 1.f) * (y < 5 ? border[y] : 1.f) *
                          (y >= H - 5 ? border[H - y - 1] : 1.f);
      r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
    }
    float* m = M + 5 * i;
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
  }
}

// 15-row box sum of the 5 matrix entries, rows clamped to the image (double, as the library's running sums).  One thread owns
// one (x, c) column of a kBoxRows-row band and slides the window down: 2 loads per output instead of 15.
constexpr int kBoxRows = 32;
__global__ void box_v_kernel(const float* __restrict__ M, double* __restrict__ V, int np, int H, int W) {
  const int m = kWin / 2;
  const size_t rowlen = (size_t)W * 5;
  const int bands = (H + kBoxRows - 1) / kBoxRows;
  const size_t total = (size_t)np * bands * rowlen;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t col = i % rowlen;
    const int band = (int)((i / rowlen) % bands);
    const size_t p = i / (rowlen * bands);
    const float* src = M + p * (size_t)H * rowlen + col;
    double* dst = V + p * (size_t)H * rowlen + col;
    const int y0 = band * kBoxRows, y1 = min(y0 + kBoxRows, H);
    double s = 0.0;
    for (int k = -m; k <= m; ++k) s += (double)src[(size_t)clampi(y0 + k, 0, H - 1) * rowlen];
    dst[(size_t)y0 * rowlen] = s;
    for (int y = y0 + 1; y < y1; ++y) {
      s += (double)src[(size_t)min(y + m, H - 1) * rowlen] - (double)src[(size_t)max(y - m - 1, 0) * rowlen];
      dst[(size_t)y * rowlen] = s;
    }
  }
}

// 15-column box sum (columns clamped) + the 2x2 solve; `mag` != nullptr on the last iteration of the finest level: |u| + |v|.
// A block stages a 128-pixel row segment plus its 7-pixel aprons in shared memory.
constexpr int kBoxSeg = 128;
__global__ void __launch_bounds__(kBoxSeg) box_h_solve_kernel(const double* __restrict__ V, float* __restrict__ flow,
                                                              float* __restrict__ mag, int np, int H, int W) {
  const int m = kWin / 2;
  __shared__ double sh[(kBoxSeg + kWin - 1) * 5];
  const int segs = (W + kBoxSeg - 1) / kBoxSeg;
  const size_t nblocks = (size_t)np * H * segs;
  const double scale = 1.0 / (kWin * kWin);
  for (size_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int seg = (int)(blk % segs);
    const size_t rowi = blk / segs;                // (pair, y)
    const double* row = V + rowi * (size_t)W * 5;
    const int x0 = seg * kBoxSeg;
    __syncthreads();
    for (int t = threadIdx.x; t < (kBoxSeg + kWin - 1) * 5; t += kBoxSeg) {
      const int px = t / 5, c = t - px * 5;
      sh[t] = row[(size_t)clampi(x0 - m + px, 0, W - 1) * 5 + c];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x < W) {
      double s[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < kWin; ++k) {
        const double* q = sh + (threadIdx.x + k) * 5;
#pragma unroll
        for (int c = 0; c < 5; ++c) s[c] += q[c];
      }
      const double g11 = s[0] * scale, g12 = s[1] * scale, g22 = s[2] * scale, h1 = s[3] * scale, h2 = s[4] * scale;
      const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
      const float u = (float)((g11 * h2 - g12 * h1) * idet), v = (float)((g22 * h1 - g12 * h2) * idet);
      const size_t i = rowi * (size_t)W + x;
      flow[2 * i] = u;
      flow[2 * i + 1] = v;
      if (mag) mag[i] = fabsf(u) + fabsf(v);
    }
  }
}

dim3 grid3(int W, int H, int n) { return dim3((unsigned)((W + 255) / 256), (unsigned)H, (unsigned)n); }

int grid_for(size_t total) {
  size_t g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  return g < 1 ? 1 : (int)g;
}

PolyConsts make_poly_consts() {
  PolyConsts pc;
  const int n = kPolyN;
  const double sigma = 1.2;
  double s = 0.0;
  for (int x = -n; x <= n; ++x) {
    pc.g[x + n] = (float)exp(-x * x / (2 * sigma * sigma));
    s += pc.g[x + n];
  }
  s = 1. / s;
  for (int x = -n; x <= n; ++x) {
    pc.g[x + n] = (float)(pc.g[x + n] * s);
    pc.xg[x + n] = (float)(x * pc.g[x + n]);
    pc.xxg[x + n] = (float)(x * x * pc.g[x + n]);
  }
  // moment matrix of the weighted basis (1, x, y, x^2, y^2, xy); only four entries of its inverse are needed
  double G00 = 0, G11 = 0, G33 = 0, G55 = 0;
  for (int y = -n; y <= n; ++y)
    for (int x = -n; x <= n; ++x) {
      const double w = (double)pc.g[y + n] * pc.g[x + n];
      G00 += w; G11 += w * x * x; G33 += w * x * x * x * x; G55 += w * x * x * y * y;
    }
  // G = [[G00,0,0,G11,G11,0],[0,G11,..],[..G11..],[G11,0,0,G33,G55,0],[G11,0,0,G55,G33,0],[0,..,G55]]: block inverse of the
  // (1, x^2, y^2) block A = [[a,b,b],[b,c,d],[b,d,c]] with a = G00, b = G11, c = G33, d = G55
  const double a = G00, b = G11, c = G33, d = G55;
  const double det = a * (c * c - d * d) - 2.0 * b * b * (c - d);
  pc.ig11 = 1.0 / G11;
  pc.ig03 = -b * (c - d) / det;           // inv(A)[0][1]
  pc.ig33 = (a * c - b * b) / det;        // inv(A)[1][1]
  pc.ig55 = 1.0 / G55;
  return pc;
}

GaussK make_gauss(double sigma) {
  GaussK gk;
  int size = (int)nearbyint(sigma * 5) | 1;   // cvRound = round half to even
  if (size < 3) size = 3;
  gk.size = size;
  if (sigma <= 0) {
    gk.k[0] = 0.25f; gk.k[1] = 0.5f; gk.k[2] = 0.25f;
    return gk;
  }
  double sum = 0, kk[kMaxGauss];
  for (int i = 0; i < size; ++i) {
    const double x = i - (size - 1) * 0.5;
    kk[i] = exp(-0.5 / (sigma * sigma) * x * x);
    sum += kk[i];
  }
  for (int i = 0; i < size; ++i) gk.k[i] = (float)(kk[i] / sum);
  return gk;
}

}  // namespace

// bytes of workspace for T frames of H x W
size_t learn_wt_workspace_bytes(int T, int H, int W) {
  const size_t px = (size_t)H * W;
  // gray T, blur tmp 2T, pyramid images T*(1+1/4+1/16+1/64) (< 1.4 T), poly tmp 3T, R 5T, flow 2(T-1) x2 levels, M 5(T-1), V 5(T-1) doubles
  return (size_t)((T * (1 + 2 + 1.4 + 3 + 5) + (T - 1) * (4 + 5)) * px * 4 + (size_t)(T - 1) * px * 5 * 8) + (1 << 20);
}

cudaError_t launch_learn_wt(const uint8_t* frames, int T, int H, int W, float* out, void* workspace, int* launches, cudaStream_t s) {
  if (T < 2) return cudaSuccess;
  if (H > 65535 || T > 65535) return cudaErrorInvalidValue;   // grid.y = rows, grid.z = images
  static const PolyConsts pc = make_poly_consts();
  const size_t px = (size_t)H * W;
  char* w = static_cast<char*>(workspace);
  auto take = [&](size_t bytes) { void* p = w; w += (bytes + 255) & ~size_t(255); return p; };
  float* gray = (float*)take(T * px * 4);
  float* t1 = (float*)take(T * px * 4);
  float* t2 = (float*)take(T * px * 4);
  float* img = (float*)take((size_t)(T * px * 4));          // current level's resized images
  float* tmp3 = (float*)take(T * px * 3 * 4);
  float* R = (float*)take(T * px * 5 * 4);
  float* flow_a = (float*)take((size_t)(T - 1) * px * 2 * 4);
  float* flow_b = (float*)take((size_t)(T - 1) * px * 2 * 4);
  float* M = (float*)take((size_t)(T - 1) * px * 5 * 4);
  double* V = (double*)take((size_t)(T - 1) * px * 5 * 8);
  int nl = 0;
  gray_kernel<<<grid_for(T * px), 256, 0, s>>>(frames, T * px, gray);
  ++nl;
  // number of usable pyramid levels (min size 32)
  int levels = 0;
  {
    double scale = 1.0;
    for (; levels < kLevels; ++levels) {
      scale *= 0.5;
      if (W * scale < 32 || H * scale < 32) break;
    }
  }
  float* flow = flow_a;
  float* flow_prev = flow_b;
  int hp = 0, wp = 0;
  for (int k = levels; k >= 0; --k) {
    double scale = 1.0;
    for (int i = 0; i < k; ++i) scale *= 0.5;
    const double sigma = (1. / scale - 1) * 0.5;
    const GaussK gk = make_gauss(sigma);
    const int wl = (int)nearbyint(W * scale), hl = (int)nearbyint(H * scale);
    const size_t pl = (size_t)hl * wl;
    // per-frame: smooth at full resolution, resize, expand
    gauss_kernel<<<grid3(W, H, T), 256, 0, s>>>(gray, t1, T, H, W, gk, 0);
    gauss_kernel<<<grid3(W, H, T), 256, 0, s>>>(t1, t2, T, H, W, gk, 1);
    resize_linear_kernel<<<grid_for(T * pl), 256, 0, s>>>(t2, img, T, H, W, hl, wl, 1, 1.0f);
    polyexp_v_kernel<<<grid3(wl, hl, T), 256, 0, s>>>(img, tmp3, T, hl, wl, pc);
    polyexp_h_kernel<<<grid3(wl, hl, T), 256, 0, s>>>(tmp3, R, T, hl, wl, pc);
    nl += 5;
    // per pair: initial flow
    if (k == levels) {
      cudaError_t e = cudaMemsetAsync(flow, 0, (size_t)(T - 1) * pl * 2 * 4, s);
      if (e != cudaSuccess) return e;
    } else {
      float* tswap = flow; flow = flow_prev; flow_prev = tswap;
      resize_linear_kernel<<<grid_for((size_t)(T - 1) * pl * 2), 256, 0, s>>>(flow_prev, flow, T - 1, hp, wp, hl, wl, 2, 2.0f);
      ++nl;
    }
    update_matrices_kernel<<<grid3(wl, hl, T - 1), 256, 0, s>>>(R, flow, M, T - 1, hl, wl);
    ++nl;
    for (int it = 0; it < kIters; ++it) {
      box_v_kernel<<<grid_for((size_t)(T - 1) * ((hl + kBoxRows - 1) / kBoxRows) * wl * 5), 256, 0, s>>>(M, V, T - 1, hl, wl);
      {
        const size_t nb = (size_t)(T - 1) * hl * ((wl + kBoxSeg - 1) / kBoxSeg);
        box_h_solve_kernel<<<(unsigned)(nb < 148 * 64 ? nb : 148 * 64), kBoxSeg, 0, s>>>(V, flow, (k == 0 && it == kIters - 1) ? out : nullptr, T - 1, hl, wl);
      }
      nl += 2;
      if (it < kIters - 1) {
        update_matrices_kernel<<<grid3(wl, hl, T - 1), 256, 0, s>>>(R, flow, M, T - 1, hl, wl);
        ++nl;
      }
    }
    hp = hl; wp = wl;
  }
  if (launches) *launches = nl;
  return cudaGetLastError();
}

}  // namespace dgp
