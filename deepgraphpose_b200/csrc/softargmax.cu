// Fused DGP soft-argmax / peak / likelihood read-out and pairwise potentials (bandwidth-class kernels, sm_100a).
//
// Reference semantics (citations relative to /root/reference):
//   argmax_2d_from_cm        src/deepgraphpose/models/fitdgp_util.py:342-402  (softmax -> zero-padded Gaussian blur ->
//                            renormalise -> expectation of (row, col))
//   estimate_pose read-out   src/deepgraphpose/models/eval.py:331-343         (<=2x2 window around mu, literal
//                            exp(x)/(exp(x)+1) sigmoid, first-max peak, likelihood)
//   argmax_pose_predict      src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/nnet/predict.py:62-77
//   skeleton / temporal      src/deepgraphpose/models/fitdgp.py:1063-1069, 1079-1083
//
// The blur is never materialised: sum_ij blur(p)[i,j] * f(i,j) == sum_ij p[i,j] * (K^T f)[i,j], and for f in
// {1, row, col} the transposed blur of f is separable and equals f in the interior, so one pass over the logits
// with per-row / per-column border weights gives all three sums.  The logit map is read exactly once from HBM
// (128-bit loads); the softmax is computed online (running max).
#include "kernels.cuh"

#include <math_constants.h>

namespace dgp {

namespace {

constexpr int kSaThreads = 256;

struct Acc {
  float m, s0, sr, sc, bsig;
  int bidx;
};

__device__ __forceinline__ float sigmoid_tf(float x) { return 1.0f / (1.0f + expf(-x)); }
// eval.py:335-336 -- the literal formula (NaN for x > ~88.7, exactly like numpy fp32)
__device__ __forceinline__ float sigmoid_literal(float x) {
  const float e = expf(x);
  return e / (e + 1.0f);
}

// 2^x for x <= 0 on the SFU (one MUFU.EX2; results below the normal range flush to zero, which is what a softmax
// numerator that small contributes anyway)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void acc_init(Acc& a) {
  a.m = -CUDART_INF_F;
  a.s0 = a.sr = a.sc = 0.0f;
  a.bsig = -1.0f;
  a.bidx = 0x7fffffff;
}

__device__ __forceinline__ void acc_merge(Acc& a, const Acc& b) {
  if (b.m != -CUDART_INF_F) {
    if (a.m == -CUDART_INF_F) {
      a.m = b.m; a.s0 = b.s0; a.sr = b.sr; a.sc = b.sc;
    } else {
      const float m = fmaxf(a.m, b.m);
      const float fa = exp2f(a.m - m), fb = exp2f(b.m - m);  // m is kept in the log2 domain
      a.s0 = a.s0 * fa + b.s0 * fb;
      a.sr = a.sr * fa + b.sr * fb;
      a.sc = a.sc * fa + b.sc * fb;
      a.m = m;
    }
  }
  if (b.bsig > a.bsig || (b.bsig == a.bsig && b.bidx < a.bidx)) {
    a.bsig = b.bsig;
    a.bidx = b.bidx;
  }
}

// One CTA handles rows [r0, r1) of one frame for ALL joints (NHWC: the joint is the fastest axis).
// `tact` threads are active with 4*tact % nj == 0, so every thread's four float4 lanes keep a fixed joint.
// kSamePixel: nj % 4 == 0, i.e. the four lanes of a float4 belong to ONE pixel (one row/col/border test per 16 bytes).
// Softmax numerators are kept in the log2 domain (m = max of x*gamma*log2e) so that an element costs one FFMA + EX2;
// the running max is updated once per 4-float4 batch, not per element.
template <bool kSamePixel>
__global__ void __launch_bounds__(kSaThreads) softargmax_partial_kernel(
    const float* __restrict__ logits, int H, int W, int nj, float gamma, int radius, float sigma, int rows_per_split,
    int splits, int tact, SaPartial* __restrict__ part) {
  extern __shared__ float sm[];
  float* red = sm;         // [4*tact][6] (+ [nj] per-joint max)

  const int b = blockIdx.x / splits;
  const int sp = blockIdx.x - b * splits;
  const int r0 = sp * rows_per_split;
  const int r1 = min(H, r0 + rows_per_split);
  const int tid = threadIdx.x;
  (void)radius;
  (void)sigma;

  const int L = 4 * tact;
  const int n_elems = (r1 - r0) * W * nj;  // < 2^31: one frame's rows
  const float4* src = reinterpret_cast<const float4*>(logits + ((size_t)b * H + r0) * (size_t)W * nj);
  float* jmax = red + (size_t)4 * tact * 6;  // [nj] per-joint max logit of this CTA's rows
  const float g2 = gamma * 1.4426950408889634f;

  // ---- pass 1: per-joint max of the CTA's rows (HBM read; leaves the rows in L2 for pass 2)
  {
    float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    if (tid < tact) {
      for (int off = 4 * tid; off < n_elems; off += 8 * L) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // 8 independent 128-bit loads in flight per thread
          const int o = off + u * L;
          v[u] = o < n_elems ? __ldg(src + (o >> 2)) : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          mx[0] = fmaxf(mx[0], v[u].x); mx[1] = fmaxf(mx[1], v[u].y); mx[2] = fmaxf(mx[2], v[u].z); mx[3] = fmaxf(mx[3], v[u].w);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) red[4 * tid + q] = mx[q];
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < nj; j += nwarps) {
      float m = -CUDART_INF_F;
      for (int e = j + nj * lane; e < L; e += nj * 32) m = fmaxf(m, red[e]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) jmax[j] = m;
    }
    __syncthreads();
  }

  // ---- pass 2: softmax numerators against the known max -- per element one FFMA + EX2 and three accumulates with the
  // INTERIOR weights (1, row, col); the few border pixels are corrected afterwards.  The exact sigmoid is evaluated
  // only for elements that can still tie with the maximum (one branch per 16 bytes).
  if (tid < tact) {
    float m2[4], s0[4], sr[4], sc[4], thr[4], bsig[4];
    int bidx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float xm = jmax[(4 * tid + q) % nj];
      m2[q] = xm * g2;
      // DLC global peak (DESIGN.md "peak candidates"): below min(max - 2, 14) fp32 sigmoids are >= 12 ulp apart.
      thr[q] = fminf(xm - 2.0f, 14.0f);
      s0[q] = sr[q] = sc[q] = 0.0f; bsig[q] = -1.0f; bidx[q] = 0x7fffffff;
    }
    const int dP = L / nj;
    const float dPr = (float)(dP / W), dPc = (float)(dP - (dP / W) * W);
    const float Wf = (float)W;
    float frow[4], fcol[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = 4 * tid + q;
      const int pix = e / nj;
      frow[q] = (float)(r0 + pix / W);
      fcol[q] = (float)(pix - (pix / W) * W);
    }
    float4 nxt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = 4 * tid + u * L;
      if (o < n_elems) nxt[u] = __ldcs(src + (o >> 2));
    }
    for (int off = 4 * tid; off < n_elems; off += 4 * L) {
      float xs[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { xs[u][0] = nxt[u].x; xs[u][1] = nxt[u].y; xs[u][2] = nxt[u].z; xs[u][3] = nxt[u].w; }
      // the next batch's loads are issued before this batch is consumed (the rows come from L2 after pass 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int o = off + (4 + u) * L;
        if (o < n_elems) nxt[u] = __ldcs(src + (o >> 2));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (off + u * L < n_elems) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float rh = kSamePixel ? frow[0] : frow[q], rw = kSamePixel ? fcol[0] : fcol[q];
            const float e = ex2_approx(fmaf(xs[u][q], g2, -m2[q]));
            s0[q] += e;
            sr[q] = fmaf(e, rh, sr[q]);
            sc[q] = fmaf(e, rw, sc[q]);
          }
          if ((xs[u][0] >= thr[0]) | (xs[u][1] >= thr[1]) | (xs[u][2] >= thr[2]) | (xs[u][3] >= thr[3])) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (xs[u][q] >= thr[q]) {
                const float s = sigmoid_tf(xs[u][q]);
                const int idx = (int)(kSamePixel ? frow[0] : frow[q]) * W + (int)(kSamePixel ? fcol[0] : fcol[q]);
                if (s > bsig[q] || (s == bsig[q] && idx < bidx[q])) { bsig[q] = s; bidx[q] = idx; }
              }
            }
          }
        }
        if (kSamePixel) {
          fcol[0] += dPc; frow[0] += dPr;
          if (fcol[0] >= Wf) { fcol[0] -= Wf; frow[0] += 1.0f; }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            fcol[q] += dPc; frow[q] += dPr;
            if (fcol[q] >= Wf) { fcol[q] -= Wf; frow[q] += 1.0f; }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float* r = red + (size_t)(4 * tid + q) * 6;
      r[0] = m2[q]; r[1] = s0[q]; r[2] = sr[q]; r[3] = sc[q]; r[4] = bsig[q];
      r[5] = __int_as_float(bidx[q]);
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < nj; j += nwarps) {
    Acc a;
    acc_init(a);
    for (int e = j + nj * lane; e < L; e += nj * 32) {
      const float* r = red + (size_t)e * 6;
      Acc t;
      t.m = r[0]; t.s0 = r[1]; t.sr = r[2]; t.sc = r[3]; t.bsig = r[4]; t.bidx = __float_as_int(r[5]);
      acc_merge(a, t);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Acc t;
      t.m = __shfl_xor_sync(0xffffffffu, a.m, o);
      t.s0 = __shfl_xor_sync(0xffffffffu, a.s0, o);
      t.sr = __shfl_xor_sync(0xffffffffu, a.sr, o);
      t.sc = __shfl_xor_sync(0xffffffffu, a.sc, o);
      t.bsig = __shfl_xor_sync(0xffffffffu, a.bsig, o);
      t.bidx = __shfl_xor_sync(0xffffffffu, a.bidx, o);
      acc_merge(a, t);
    }
    if (lane == 0) {
      SaPartial& o = part[((size_t)b * splits + sp) * nj + j];
      o.m = a.m; o.s0 = a.s0; o.sr = a.sr; o.sc = a.sc; o.bsig = a.bsig; o.bidx = a.bidx;
    }
  }
}

// Blur border weights of source index `pos` on an axis of length n (zero padding, VALID conv, then renormalisation):
// a = sum of the taps that stay inside, r = sum of those taps times the blurred-map index they land on.
__device__ __forceinline__ void border_weights(int pos, int n, int radius, const float* kt, float& a, float& r) {
  a = 0.0f;
  r = 0.0f;
  for (int d = -radius; d <= radius; ++d) {
    const int dst = pos - d;
    if (dst >= 0 && dst < n) {
      a += kt[d + radius];
      r += kt[d + radius] * (float)dst;
    }
  }
}

// One WARP per (frame, joint): merge the row-split partials (interior weights), add the border correction
// (pixels within `radius` of an edge lose the taps that fall outside: weights (Ah*Aw, Rh*Aw, Ah*Rw) instead of
// (1, row, col); fixed lane order -> deterministic), then lane 0 does the O(1) read-outs.
__global__ void softargmax_finalize_kernel(const float* __restrict__ logits, const float* __restrict__ locref, int B,
                                           int H, int W, int nj, int splits, const SaPartial* __restrict__ part,
                                           float gamma, int radius, float sigma, float stride, float locref_stdev,
                                           float* __restrict__ mu, int* __restrict__ peak, float* __restrict__ lik,
                                           int* __restrict__ dlc_peak, float* __restrict__ dlc_pose,
                                           float* __restrict__ norm) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= B * nj) return;
  const int b = t / nj, j = t - b * nj;
  Acc a;
  acc_init(a);
  for (int sp = lane; sp < splits; sp += 32) {
    const SaPartial& p = part[((size_t)b * splits + sp) * nj + j];
    Acc q;
    q.m = p.m; q.s0 = p.s0; q.sr = p.sr; q.sc = p.sc; q.bsig = p.bsig; q.bidx = p.bidx;
    acc_merge(a, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Acc q;
    q.m = __shfl_xor_sync(0xffffffffu, a.m, o);
    q.s0 = __shfl_xor_sync(0xffffffffu, a.s0, o);
    q.sr = __shfl_xor_sync(0xffffffffu, a.sr, o);
    q.sc = __shfl_xor_sync(0xffffffffu, a.sc, o);
    q.bsig = __shfl_xor_sync(0xffffffffu, a.bsig, o);
    q.bidx = __shfl_xor_sync(0xffffffffu, a.bidx, o);
    acc_merge(a, q);
  }
  const float* fr = logits + (size_t)b * H * W * nj + j;
  {
    const float g2 = gamma * 1.4426950408889634f;
    float knorm = 0.0f;
    for (int d = -radius; d <= radius; ++d) knorm += expf(-0.5f * (d / sigma) * (d / sigma));
    float kt[9];  // normalised 1-D taps (radius <= 4)
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float d = (float)(i - radius);
      kt[i] = i <= 2 * radius ? expf(-0.5f * (d / sigma) * (d / sigma)) / knorm : 0.0f;
    }
    const bool all_border = (W <= 2 * radius) || (H <= 2 * radius);
    float d0 = 0.0f, dr = 0.0f, dc = 0.0f;
    auto add_px = [&](int r, int c) {
      const float e = ex2_approx(fmaf(__ldg(fr + ((size_t)r * W + c) * nj), g2, -a.m));
      float ah, rh, aw, rw;
      border_weights(r, H, radius, kt, ah, rh);
      border_weights(c, W, radius, kt, aw, rw);
      d0 += e * (ah * aw - 1.0f);
      dr += e * (rh * aw - (float)r);
      dc += e * (ah * rw - (float)c);
    };
    if (all_border) {
      for (int k = lane; k < H * W; k += 32) add_px(k / W, k - (k / W) * W);
    } else {
      const int nrow = 2 * radius * W;                 // full top / bottom rows
#pragma unroll 4
      for (int k = lane; k < nrow; k += 32) {
        const int rr = k / W;
        add_px(rr < radius ? rr : H - 2 * radius + rr, k - rr * W);
      }
      const int nside = (H - 2 * radius) * 2 * radius;  // left / right columns of the remaining rows
#pragma unroll 4
      for (int k = lane; k < nside; k += 32) {
        const int rr = radius + k / (2 * radius);
        const int sidx = k - (k / (2 * radius)) * (2 * radius);
        add_px(rr, sidx < radius ? sidx : W - 2 * radius + sidx);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      dc += __shfl_xor_sync(0xffffffffu, dc, o);
    }
    a.s0 += d0; a.sr += dr; a.sc += dc;
  }
  if (lane != 0) return;
  const float mur = a.sr / a.s0, muc = a.sc / a.s0;  // 0/0 -> NaN, as softmax_tensor / (sum + 1e-100) in fp32
  if (mu) { mu[2 * t] = mur; mu[2 * t + 1] = muc; }
  if (norm) { norm[2 * t] = a.m; norm[2 * t + 1] = a.s0; }

  if (peak || lik) {
    int pr = -1, pc = -1;
    float best = CUDART_NAN_F;
    if (mur == mur && muc == muc) {
      // numpy slice [floor : ceil+1] clipped to the array
      const int rlo = max((int)floorf(mur), 0), rhi = min((int)ceilf(mur) + 1, H);
      const int clo = max((int)floorf(muc), 0), chi = min((int)ceilf(muc) + 1, W);
      bool have = false, have_nan = false;
      for (int r = rlo; r < rhi && !have_nan; ++r)
        for (int c = clo; c < chi; ++c) {
          const float s = sigmoid_literal(fr[((size_t)r * W + c) * nj]);
          if (s != s) { pr = r; pc = c; best = s; have_nan = true; break; }  // np.argmax: first NaN wins
          if (!have || s > best) { best = s; pr = r; pc = c; have = true; }
        }
    }
    if (peak) { peak[2 * t] = pr; peak[2 * t + 1] = pc; }
    if (lik) lik[t] = best;
  }
  if (dlc_peak || dlc_pose) {
    const int r = a.bidx / W, c = a.bidx - r * W;
    if (dlc_peak) { dlc_peak[2 * t] = r; dlc_peak[2 * t + 1] = c; }
    if (dlc_pose) {
      float dx = 0.0f, dy = 0.0f;
      if (locref) {
        const float* lp = locref + (((size_t)b * H + r) * W + c) * (size_t)(2 * nj) + 2 * j;
        dx = lp[0] * locref_stdev;
        dy = lp[1] * locref_stdev;
      }
      dlc_pose[3 * t] = (float)c * stride + 0.5f * stride + dx;
      dlc_pose[3 * t + 1] = (float)r * stride + 0.5f * stride + dy;
      dlc_pose[3 * t + 2] = a.bsig;
    }
  }
}

// Second output of argmax_2d_from_cm (fitdgp_util.py:391): the blurred, renormalised softmax map (N,H,W,C).
// out = (K * exp(gamma*x - m)) / s0 with zero padding; m, s0 come from the soft-argmax pass.
__global__ void softmax_map_kernel(const float* __restrict__ logits, const float* __restrict__ norm, int B, int H, int W,
                                   int nj, float gamma, int radius, float sigma, float* __restrict__ out) {
  float k[9];
  float knorm = 0.0f;
  for (int d = -radius; d <= radius; ++d) knorm += expf(-0.5f * (d / sigma) * (d / sigma));
  for (int d = -radius; d <= radius && d + radius < 9; ++d) k[d + radius] = expf(-0.5f * (d / sigma) * (d / sigma)) / knorm;
  const size_t total = (size_t)B * H * W * nj;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % nj);
    size_t r = t / nj;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    const float m = norm[2 * ((size_t)b * nj + c)], s0 = norm[2 * ((size_t)b * nj + c) + 1];
    float acc = 0.0f;
    for (int dy = -radius; dy <= radius; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -radius; dx <= radius; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        acc += k[dy + radius] * k[dx + radius] *
               exp2f(fmaf(logits[(((size_t)b * H + yy) * W + xx) * nj + c], gamma * 1.4426950408889634f, -m));
      }
    }
    out[t] = acc / s0;
  }
}

// sigmoid scoremap (PoseNet.test, pose_net.py:84-90) -- only materialised when a caller asks for it.
__global__ void sigmoid_map_kernel(const float4* __restrict__ x, float4* __restrict__ y, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(x + i);
    y[i] = make_float4(sigmoid_tf(v.x), sigmoid_tf(v.y), sigmoid_tf(v.z), sigmoid_tf(v.w));
  }
}

// Skeleton distances and temporal differences on the soft-argmax coordinates.  One thread per frame, 128 frames per
// CTA; the CTA's (128 + 1 halo) x nj x 2 coordinates are staged in shared memory with coalesced 128-bit loads
// (a per-thread walk over its own 8*nj-byte row would cost one L1 wavefront per 4 bytes), and the temporal rows go
// back through shared memory so that global stores are coalesced too.
constexpr int kPotFrames = 128;
__global__ void __launch_bounds__(kPotFrames) potentials_kernel(
    const float* __restrict__ mu, const float* __restrict__ halo_next, int T, int nj, const int* __restrict__ edges,
    int nl, float stride, const float* __restrict__ ws, const float* __restrict__ ws_max, float wt_max,
    float* __restrict__ skel, float* __restrict__ temporal, float* __restrict__ e_skel, float* __restrict__ e_temp) {
  extern __shared__ float psm[];
  const int row = 2 * nj + 1;                       // +1: conflict-free column walks
  float* sm_mu = psm;                               // [kPotFrames + 1][row]
  float* sm_t = psm + (kPotFrames + 1) * row;       // [kPotFrames][nj + 1]
  const int t0 = blockIdx.x * kPotFrames;
  const int nf = min(kPotFrames, T - t0);
  const int nload = (t0 + nf < T) ? nf + 1 : nf;    // + first frame of the next CTA's range
  const float* src = mu + (size_t)t0 * nj * 2;
  {
    // flat coalesced copy with an incrementally maintained (frame, coordinate) pair: no division in the loop and all
    // of a thread's loads are independent
    const int rl = 2 * nj;
    int f = (int)threadIdx.x / rl, c = (int)threadIdx.x - f * rl;
    const int df = (int)blockDim.x / rl, dc = (int)blockDim.x - df * rl;
#pragma unroll 8
    for (int i = threadIdx.x; i < nload * rl; i += blockDim.x) {
      sm_mu[f * row + c] = src[i];
      f += df; c += dc;
      if (c >= rl) { c -= rl; f += 1; }
    }
  }
  const bool have_next_global = (t0 + nf < T);
  if (!have_next_global && halo_next != nullptr)
    for (int i = threadIdx.x; i < 2 * nj; i += blockDim.x) sm_mu[nf * row + i] = halo_next[i];
  __syncthreads();
  const int lt = threadIdx.x;
  const int t = t0 + lt;
  const bool active = lt < nf;
  const bool has_next = active && (lt + 1 < nf || have_next_global || halo_next != nullptr);
  if (active) {
    const float* m = sm_mu + lt * row;
    float es = 0.0f;
    for (int l = 0; l < nl; ++l) {
      const int a = __ldg(edges + 2 * l), b = __ldg(edges + 2 * l + 1);
      // S (mu*stride + stride/2): keep the reference's order of operations
      const float dr = (m[2 * a] * stride + 0.5f * stride) - (m[2 * b] * stride + 0.5f * stride);
      const float dc = (m[2 * a + 1] * stride + 0.5f * stride) - (m[2 * b + 1] * stride + 0.5f * stride);
      const float d = sqrtf(dr * dr + dc * dc);
      if (skel) skel[(size_t)l * T + t] = d;
      if (ws) es += __ldg(ws + l) * (fmaxf(d - __ldg(ws_max + l), 0.0f) + __ldg(ws_max + l));
    }
    if (e_skel) e_skel[t] = es;
    float et = 0.0f;
    if (has_next) {
      const float* mn = m + row;
      for (int j = 0; j < nj; ++j) {
        const float dr = (m[2 * j] * stride + 0.5f * stride) - (mn[2 * j] * stride + 0.5f * stride);
        const float dc = (m[2 * j + 1] * stride + 0.5f * stride) - (mn[2 * j + 1] * stride + 0.5f * stride);
        const float d = sqrtf(dr * dr + dc * dc);
        sm_t[lt * (nj + 1) + j] = d;
        const float dth = fmaxf(d - wt_max, 0.0f) + wt_max;
        et += dth * dth;
      }
    }
    if (e_temp) e_temp[t] = et;
  }
  __syncthreads();
  if (temporal) {
    // rows [t0, t0 + nvalid) of temporal are contiguous in global memory
    const int nvalid = (have_next_global || halo_next != nullptr) ? nf : nf - 1;
    float* dst = temporal + (size_t)t0 * nj;
    int f = (int)threadIdx.x / nj, c = (int)threadIdx.x - f * nj;
    const int df = (int)blockDim.x / nj, dc = (int)blockDim.x - df * nj;
#pragma unroll 4
    for (int i = threadIdx.x; i < nvalid * nj; i += blockDim.x) {
      dst[i] = sm_t[f * (nj + 1) + c];
      f += df; c += dc;
      if (c >= nj) { c -= nj; f += 1; }
    }
  }
}

}  // namespace

int softargmax_tact(int nj) {
  // largest thread count <= 256 with 4*t % nj == 0
  int g = nj;
  for (int a = 4, b = nj; b;) { int r = a % b; a = b; b = r; g = a; }
  const int step = nj / g;  // t must be a multiple of nj / gcd(nj, 4)
  int t = (kSaThreads / step) * step;
  return t;
}

int softargmax_splits(int B, int H, int num_sms) {
  int splits = (3 * num_sms + B - 1) / B;
  const int max_splits = (H + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

cudaError_t launch_softargmax(const float* logits, const float* locref, int B, int H, int W, int nj, float gamma,
                              float gauss_len, float stride, float locref_stdev, SaPartial* workspace, int splits,
                              float* mu, int* peak, float* lik, int* dlc_peak, float* dlc_pose, float* norm,
                              cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const int tact = softargmax_tact(nj);
  if (tact <= 0 || (H & 1) || (W & 1)) return cudaErrorInvalidValue;
  int rows_per_split = (H + splits - 1) / splits;
  rows_per_split = (rows_per_split + 1) & ~1;  // even row boundaries keep the float4 loads 16 B aligned
  const int real_splits = (H + rows_per_split - 1) / rows_per_split;
  const int radius = (int)gauss_len;
  const size_t smem = (size_t)4 * tact * 6 * 4 + (size_t)nj * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(softargmax_partial_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(softargmax_partial_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (nj % 4 == 0)
    softargmax_partial_kernel<true><<<B * real_splits, kSaThreads, smem, stream>>>(logits, H, W, nj, gamma, radius, gauss_len,
                                                                                  rows_per_split, real_splits, tact, workspace);
  else
    softargmax_partial_kernel<false><<<B * real_splits, kSaThreads, smem, stream>>>(logits, H, W, nj, gamma, radius, gauss_len,
                                                                                   rows_per_split, real_splits, tact, workspace);
  const int n = B * nj;
  softargmax_finalize_kernel<<<(n + 3) / 4, 128, 0, stream>>>(logits, locref, B, H, W, nj, real_splits, workspace, gamma,
                                                              radius, gauss_len, stride, locref_stdev, mu, peak, lik,
                                                              dlc_peak, dlc_pose, norm);
  return cudaGetLastError();
}

cudaError_t launch_softmax_map(const float* logits, const float* norm, int B, int H, int W, int nj, float gamma,
                               float gauss_len, float* out, int num_sms, cudaStream_t stream) {
  const int radius = (int)gauss_len;
  if (radius > 4) return cudaErrorInvalidValue;
  const size_t total = (size_t)B * H * W * nj;
  size_t grid = (total + 255) / 256;
  if (grid > (size_t)num_sms * 16) grid = (size_t)num_sms * 16;
  softmax_map_kernel<<<(int)grid, 256, 0, stream>>>(logits, norm, B, H, W, nj, gamma, radius, gauss_len, out);
  return cudaGetLastError();
}

cudaError_t launch_sigmoid_map(const float* x, float* y, size_t n, int num_sms, cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  const size_t n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > num_sms * 8) grid = num_sms * 8;
  if (grid < 1) grid = 1;
  sigmoid_map_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n4);
  return cudaGetLastError();
}

cudaError_t launch_potentials(const float* mu, const float* halo_next, int T, int nj, const int* edges, int nl,
                              float stride, const float* ws, const float* ws_max, float wt_max, float* skel,
                              float* temporal, float* e_skel, float* e_temp, cudaStream_t stream) {
  if (T <= 0) return cudaSuccess;
  const size_t smem = ((size_t)(kPotFrames + 1) * (2 * nj + 1) + (size_t)kPotFrames * (nj + 1)) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(potentials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  potentials_kernel<<<(T + kPotFrames - 1) / kPotFrames, kPotFrames, smem, stream>>>(mu, halo_next, T, nj, edges, nl, stride, ws,
                                                                                   ws_max, wt_max, skel, temporal, e_skel, e_temp);
  return cudaGetLastError();
}

}  // namespace dgp
