// Forward pass of the DGP loss (reference: src/deepgraphpose/models/fitdgp.py:947-1128) on the head outputs.
//
// The reference materialises ~10 tensors of shape (nt*nj, H, W) (Gaussian targets, sigmoids, scaled logits, CE maps).
// Here every marker (frame, joint) is one CTA that recomputes its Gaussian target on the fly from the soft-argmax /
// label coordinate and reduces its cross-entropy, confidence and locref-Huber sums with warp shuffles; a final
// single-warp kernel adds the per-marker partials in a fixed order (deterministic) together with the skeleton
// (spatial clique) and temporal clique terms.  Bandwidth class: 4*H*W*nj*nt bytes of logits read twice (max pass +
// loss pass, second pass from L2) + 3*8*H*W bytes per visible marker for the locref term.
#include "kernels.cuh"

#include <math_constants.h>

namespace dgp {

namespace {

constexpr int kLossThreads = 512;
constexpr size_t kPlaneSmemMax = 200 * 1024;   // a marker's logit plane is staged in shared memory when it fits

// Stage channel j of one frame's logits (stride nj floats) into shared memory with 8 independent loads in flight per thread.
// The marker kernels make two or three passes over this plane with ~40 pixels per thread: read straight from global memory
// every pass was a chain of dependent ~1 us loads (87 + 111 us per training step for 40 markers); staged, the passes run from
// shared memory.  Returns the pointer / stride the passes should use.
__device__ __forceinline__ const float* stage_plane(const float* __restrict__ x0, int HW, int nj, float* plane, int use_smem,
                                                    int* xstride) {
  if (!use_smem) { *xstride = nj; return x0; }
  for (int p0 = threadIdx.x; p0 < HW; p0 += blockDim.x * 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + u * blockDim.x;
      v[u] = p < HW ? __ldg(x0 + (size_t)p * nj) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + u * blockDim.x;
      if (p < HW) plane[p] = v[u];
    }
  }
  __syncthreads();
  *xstride = 1;
  return plane;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.0f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];  // fixed order -> deterministic
  return r;
}
__device__ float block_max(float v, float* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = -CUDART_INF_F;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, sh[i]);
  return r;
}

// tf.nn.sigmoid_cross_entropy_with_logits
__device__ __forceinline__ float bce(float z, float x) { return fmaxf(x, 0.0f) - x * z + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// combine_all_marker (fitdgp_util.py:232-272): scatter-add of hidden predictions and visible labels.
__global__ void combine_markers_kernel(const float* __restrict__ mu, const float* __restrict__ targets,
                                       const int* __restrict__ visible, int nbv, const int* __restrict__ hidden, int nbh,
                                       const int* __restrict__ vis_in_targets, float* __restrict__ all) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nbh) {
    const int m = hidden[i];
    atomicAdd(&all[2 * m], mu[2 * m]);
    atomicAdd(&all[2 * m + 1], mu[2 * m + 1]);
  } else if (i < nbh + nbv) {
    const int k = i - nbh;
    const int m = visible[k];
    const int s = vis_in_targets[k];
    const float a = targets[2 * s], b = targets[2 * s + 1];
    atomicAdd(&all[2 * m], a == a ? a : 0.0f);      // targets_nonan (fitdgp.py:898)
    atomicAdd(&all[2 * m + 1], b == b ? b : 0.0f);
  }
}

// One CTA per listed marker: blocks [0, nbv) are the visible markers, [nbv, nbv + nbh) the hidden ones.
// part[blk] = {ce_sum, weight_count, huber_sum, mask_count}
__global__ void __launch_bounds__(kLossThreads) marker_loss_kernel(
    const float* __restrict__ pred, const float* __restrict__ locref, const float* __restrict__ locref_map,
    const float* __restrict__ locref_mask, const float* __restrict__ all, const int* __restrict__ visible, int nbv,
    const int* __restrict__ hidden, int H, int W, int nj, float inv2l2, int gm2, int gm3, int locref_mse,
    float4* __restrict__ part, int use_smem) {
  extern __shared__ float plane[];
  __shared__ float sh[kLossThreads / 32];
  const bool is_vis = (int)blockIdx.x < nbv;
  const int m = is_vis ? visible[blockIdx.x] : hidden[blockIdx.x - nbv];
  const int t = m / nj, j = m - t * nj;
  const float mur = all[2 * m], muc = all[2 * m + 1];
  const int HW = H * W;
  int xs_stride;
  const float* x0 = stage_plane(pred + (size_t)t * H * W * nj + j, HW, nj, plane, use_smem, &xs_stride);

  // pass A: max of the Gaussian bump (+1e-5, fitdgp.py:973) and, for hidden markers, the confidence max sigmoid
  float gmax = 0.0f, cmax = -CUDART_INF_F;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    const float dr = (float)r - mur, dc = (float)c - muc;
    gmax = fmaxf(gmax, expf(-(dr * dr + dc * dc) * inv2l2));
    if (!is_vis && gm2 != 0) cmax = fmaxf(cmax, sigmoidf_(x0[(size_t)p * xs_stride]));
  }
  gmax = block_max(gmax, sh) + 1e-5f;
  float conf = 1.0f;
  if (!is_vis && gm2 != 0) conf = block_max(cmax, sh);

  // pass B: cross entropy against the on-the-fly target
  float ce = 0.0f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    const float dr = (float)r - mur, dc = (float)c - muc;
    float tg = expf(-(dr * dr + dc * dc) * inv2l2) / gmax;
    float x = x0[(size_t)p * xs_stride];
    if (!is_vis && gm2 != 0) {
      if (gm2 == 1) tg *= conf;                       // fitdgp.py:1000
      if (gm3 == 3) {                                 // confidence-scaled logits (fitdgp.py:1002-1004)
        const float ps = sigmoidf_(x) * conf;
        x = -logf(1.0f - ps + 1e-20f) + logf(ps + 1e-20f);
      }
    }
    ce += bce(tg, x);
  }
  ce = block_sum(ce, sh);
  float wcount = (float)HW;
  if (!is_vis && gm3 == 3) {
    const float w = 1.0f - conf;                      // SUM_BY_NONZERO_WEIGHTS over the broadcast weights
    ce *= w;
    wcount = (w != 0.0f) ? (float)HW : 0.0f;
  }

  // locref Huber (PTF/nnet/losses.py:16-45), visible markers only: channels 2j, 2j+1
  float hub = 0.0f, mcount = 0.0f;
  if (is_vis && locref != nullptr) {
    const size_t base = (size_t)t * HW * 2 * nj + 2 * j;
#pragma unroll 4
    for (int p = threadIdx.x; p < 2 * HW; p += blockDim.x) {
      const int pix = p >> 1, ch = p & 1;
      const size_t o = base + (size_t)pix * 2 * nj + ch;
      const float w = __ldg(locref_mask + o);
      const float d = __ldg(locref + o) - __ldg(locref_map + o);
      const float a = fabsf(d);
      // huber_loss (k = 1), or tf.losses.mean_squared_error when dgp_cfg.locref_huber_loss is False (fitdgp.py:1053)
      const float l = locref_mse ? d * d : (a < 1.0f ? 0.5f * d * d : a - 0.5f);
      hub += l * w;
      mcount += (w != 0.0f) ? 1.0f : 0.0f;
    }
    hub = block_sum(hub, sh);
    mcount = block_sum(mcount, sh);
  }
  if (threadIdx.x == 0) part[blockIdx.x] = make_float4(ce, wcount, hub, mcount);
}

// mean over tf.image.crop_and_resize(flow, box, [Hin, Win]) per (t, j) (fitdgp.py:1085-1110), bilinear, extrapolation 0, and
// the gradient of that mean w.r.t. the box (tf CropAndResizeGradBoxes).
// crop_and_resize samples Hin x Win points INSIDE the box (a few dozen pixels wide), so the 621 k bilinear samples of a
// 747x832 frame land on ~10^3 source pixels.  Bilinear weights are separable: with, per source row r,
//   Wy[r]  = sum over valid output rows yy of the weight yy puts on r          ((1 - ly) on floor(ys), ly on ceil(ys))
//   Dy1[r] = sum over yy of (Hin-1-yy) * ([ceil(ys) == r] - [floor(ys) == r]),  Dy2[r] likewise with yy
// (and Wx, Dx1, Dx2 per source column), the five sums are  Wy' I Wx,  Dy1' I Wx,  Wy' I Dx1,  Dy2' I Wx,  Wy' I Dx2  over the
// box's pixels.  Round 2: replaces the sampled version (36 boxes x 621 k samples, 95 us per training step).  Grid (box, slab of
// 16 source rows): a box that spans the whole frame is still a 148-SM kernel; CTAs past the box's last row exit with zeros.
// Every weight is accumulated by one thread in a fixed order, block reduction and slab sum are fixed-order: reproducible.
constexpr int kFlowRows = 16;   // source rows per CTA: grid = (box, ceil(Hin / kFlowRows)); CTAs beyond the box's rows write zeros
__global__ void __launch_bounds__(kLossThreads) flow_box_mean_kernel(const float* __restrict__ flow, const float* __restrict__ all,
                                                                   int nt, int nj, int Hin, int Win, float stride,
                                                                   int want_grad, float* __restrict__ part) {
  extern __shared__ float wts[];   // [3][kFlowRows] row weights, then [3][nc] column weights
  __shared__ float sh[kLossThreads / 32];
  const int box = blockIdx.x;
  const int t = box / nj, j = box - t * nj;
  const float* a0 = all + ((size_t)t * nj + j) * 2;
  const float* a1 = all + ((size_t)(t + 1) * nj + j) * 2;
  const float r0 = a0[0] * stride + 0.5f * stride, c0 = a0[1] * stride + 0.5f * stride;
  const float r1 = a1[0] * stride + 0.5f * stride, c1 = a1[1] * stride + 0.5f * stride;
  const float nx = (float)Hin, ny = (float)Win;
  const float rmin = fmaxf(0.0f, fminf(r0, r1) - 10.0f), rmax = fminf(nx, fmaxf(r0, r1) + 10.0f);
  const float cmin = fmaxf(0.0f, fminf(c0, c1) - 10.0f), cmax = fminf(ny, fmaxf(c0, c1) + 10.0f);
  const float y1 = rmin / nx, x1 = cmin / ny, y2 = rmax / nx, x2 = cmax / ny;
  const float sy = Hin > 1 ? (y2 - y1) * (float)(Hin - 1) / (float)(Hin - 1) : 0.0f;
  const float sx = Win > 1 ? (x2 - x1) * (float)(Win - 1) / (float)(Win - 1) : 0.0f;
  auto sample_y = [&](int yy) { return Hin > 1 ? y1 * (float)(Hin - 1) + (float)yy * sy : 0.5f * (y1 + y2) * (float)(Hin - 1); };
  auto sample_x = [&](int xx) { return Win > 1 ? x1 * (float)(Win - 1) + (float)xx * sx : 0.5f * (x1 + x2) * (float)(Win - 1); };
  // source rows / columns any sample can touch (a margin of one absorbs the rounding of the end points; NaN markers -> empty)
  const float ya = fminf(sample_y(0), sample_y(Hin - 1)), yb = fmaxf(sample_y(0), sample_y(Hin - 1));
  const float xa = fminf(sample_x(0), sample_x(Win - 1)), xb = fmaxf(sample_x(0), sample_x(Win - 1));
  int r_lo = 0, r_hi = -1, c_lo = 0, c_hi = -1;
  if (ya == ya && yb == yb && xa == xa && xb == xb) {
    r_lo = max(0, (int)floorf(fmaxf(ya, 0.0f)) - 1); r_hi = min(Hin - 1, (int)ceilf(fminf(yb, (float)(Hin - 1))) + 1);
    c_lo = max(0, (int)floorf(fmaxf(xa, 0.0f)) - 1); c_hi = min(Win - 1, (int)ceilf(fminf(xb, (float)(Win - 1))) + 1);
  }
  // this CTA's slab of the box's source rows
  r_lo += (int)blockIdx.y * kFlowRows;
  r_hi = min(r_hi, r_lo + kFlowRows - 1);
  const int nr = max(r_hi - r_lo + 1, 0), nc = nr > 0 ? max(c_hi - c_lo + 1, 0) : 0;
  float* wy = wts;               // Wy, Dy1, Dy2
  float* wx = wts + 3 * nr;      // Wx, Dx1, Dx2
  // ---- phase 1: one thread per source row / column.  The sample positions ps(k) are monotone in k, so the samples whose
  // floor (ceil) is this row form a contiguous range of k: found by bisection on the same float expression the interpolation
  // uses, then accumulated in ascending k (fixed order).
  for (int it = threadIdx.x; it < nr + nc; it += blockDim.x) {
    const bool is_row = it < nr;
    const int me = is_row ? r_lo + it : c_lo + (it - nr);
    const int n = is_row ? Hin : Win;
    auto ps_of = [&](int k) { return is_row ? sample_y(k) : sample_x(k); };
    auto first_ge = [&](float v) {   // min k in [0, n] with ps(k) >= v
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ps_of(mid) >= v) hi = mid; else lo = mid + 1; }
      return lo;
    };
    auto first_gt = [&](float v) {   // min k in [0, n] with ps(k) > v
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ps_of(mid) > v) hi = mid; else lo = mid + 1; }
      return lo;
    };
    const int kv0 = first_ge(0.0f), kv1 = first_gt((float)(n - 1));   // samples inside the image: [kv0, kv1)
    float w = 0.0f, d1 = 0.0f, d2 = 0.0f;
    const float half = 0.5f * (float)(n - 1);
    // floor(ps) == me
    for (int k = max(first_ge((float)me), kv0), ke = min(first_ge((float)(me + 1)), kv1); k < ke; ++k) {
      const float ps = ps_of(k);
      const float l = ps - floorf(ps);
      w += 1.0f - l;
      d1 -= n > 1 ? (float)(n - 1 - k) : half;
      d2 -= n > 1 ? (float)k : half;
    }
    // ceil(ps) == me  (ps <= n - 1 for every sample inside the image, so the clamp of the upper neighbour never bites)
    for (int k = max(first_gt((float)(me - 1)), kv0), ke = min(first_gt((float)me), kv1); k < ke; ++k) {
      const float ps = ps_of(k);
      const float l = ps - floorf(ps);
      w += l;
      d1 += n > 1 ? (float)(n - 1 - k) : half;
      d2 += n > 1 ? (float)k : half;
    }
    float* o = is_row ? wy + it : wx + (it - nr);
    const int ld = is_row ? nr : nc;
    o[0] = w; o[ld] = d1; o[2 * ld] = d2;
  }
  __syncthreads();
  // ---- phase 2: the five weighted sums over the touched pixels
  const float* img = flow + (size_t)t * Hin * Win;
  float acc = 0.0f, gy1 = 0.0f, gx1 = 0.0f, gy2 = 0.0f, gx2 = 0.0f;
  for (int p = threadIdx.x; p < nr * nc; p += blockDim.x) {
    const int ri = p / nc, ci = p - ri * nc;
    const float v = __ldg(img + (size_t)(r_lo + ri) * Win + (c_lo + ci));
    const float Wy = wy[ri], Dy1 = wy[nr + ri], Dy2 = wy[2 * nr + ri];
    const float Wx = wx[ci], Dx1 = wx[nc + ci], Dx2 = wx[2 * nc + ci];
    const float vy = v * Wx, vx = v * Wy;
    acc += vy * Wy;
    gy1 += vy * Dy1;
    gy2 += vy * Dy2;
    gx1 += vx * Dx1;
    gx2 += vx * Dx2;
  }
  float* out = part + ((size_t)box * gridDim.y + blockIdx.y) * 8;
  acc = block_sum(acc, sh);
  if (want_grad) {
    gy1 = block_sum(gy1, sh); gx1 = block_sum(gx1, sh); gy2 = block_sum(gy2, sh); gx2 = block_sum(gx2, sh);
  }
  if (threadIdx.x == 0) {
    out[0] = acc; out[1] = gy1; out[2] = gx1; out[3] = gy2; out[4] = gx2;
  }
}

// fixed-order sum of a box's row slabs
__global__ void flow_box_finalize_kernel(const float* __restrict__ part, int nchunks, int Hin, int Win, float* __restrict__ meanflow,
                                         float4* __restrict__ boxgrad) {
  // one warp per box: lane l adds slabs l, l + 32, ... in order, then a fixed xor tree (deterministic)
  const int box = blockIdx.x;
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int c = threadIdx.x; c < nchunks; c += 32) {
    const float* q = part + ((size_t)box * nchunks + c) * 8;
#pragma unroll
    for (int i = 0; i < 5; ++i) s[i] += q[i];
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) s[i] = warp_sum(s[i]);
  if (threadIdx.x != 0) return;
  const float k = 1.0f / (float)(Hin * Win);
  meanflow[box] = s[0] * k;
  if (boxgrad != nullptr) boxgrad[box] = make_float4(s[1] * k, s[2] * k, s[3] * k, s[4] * k);
}

// Final fixed-order reduction + clique terms.  out[6] = {visible_loss_pred, hidden_loss_pred, visible_loss_locref,
// ws_loss, wt_loss, total_loss}.  Single thread: O(nb + nl*nt + nt*nj) work.
__global__ void loss_finalize_kernel(const float4* __restrict__ part, int nbv, int nbh, const float* __restrict__ all,
                                     int nt, int nj, int H, int W, const int* __restrict__ edges, int nl,
                                     const float* __restrict__ ws, const float* __restrict__ ws_max,
                                     const float* __restrict__ meanflow, const float* __restrict__ wt_batch, float wt,
                                     float wt_max, float stride, float n_vis_total, float n_hid_total, float wn_visible,
                                     float wn_hidden, float locref_weight, float* __restrict__ out, int use_smem) {
  // The arithmetic is one thread's fixed-order walk (deterministic); its inputs are first staged into shared memory by the whole
  // block, so that the walk is not a chain of ~200 dependent global loads (20 us -> a few us per training step).
  extern __shared__ float4 fin_sm[];
  if (use_smem) {
    const int nb = nbv + nbh, nm2 = nt * nj * 2, nbox = meanflow != nullptr ? (nt - 1) * nj : 0;
    float4* s_part = fin_sm;
    float* s_all = reinterpret_cast<float*>(s_part + nb);
    float* s_mf = s_all + nm2;
    float* s_wtb = s_mf + nbox;
    float* s_ws = s_wtb + (nbox > 0 ? nt - 1 : 0);
    float* s_wsm = s_ws + nl;
    int* s_edges = reinterpret_cast<int*>(s_wsm + nl);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_part[i] = part[i];
    for (int i = threadIdx.x; i < nm2; i += blockDim.x) s_all[i] = all[i];
    for (int i = threadIdx.x; i < nbox; i += blockDim.x) s_mf[i] = meanflow[i];
    if (nbox > 0)
      for (int i = threadIdx.x; i < nt - 1; i += blockDim.x) s_wtb[i] = wt_batch[i];
    for (int i = threadIdx.x; i < nl; i += blockDim.x) { s_ws[i] = ws[i]; s_wsm[i] = ws_max[i]; }
    for (int i = threadIdx.x; i < 2 * nl; i += blockDim.x) s_edges[i] = edges[i];
    __syncthreads();
    part = s_part; all = s_all; ws = s_ws; ws_max = s_wsm; edges = s_edges;
    if (nbox > 0) { meanflow = s_mf; wt_batch = s_wtb; }
  }
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float ce_v = 0.0f, cnt_v = 0.0f, hub = 0.0f, mcnt = 0.0f, ce_h = 0.0f, cnt_h = 0.0f;
  for (int i = 0; i < nbv; ++i) { ce_v += part[i].x; cnt_v += part[i].y; hub += part[i].z; mcnt += part[i].w; }
  for (int i = nbv; i < nbv + nbh; ++i) { ce_h += part[i].x; cnt_h += part[i].y; }
  const float fnbh = (float)nbh;
  const float fnbv = nbv > 0 ? (float)nbv : fnbh;   // fitdgp.py:982-984
  const float visible_loss = cnt_v > 0.0f ? ce_v / cnt_v : 0.0f;
  float hidden_loss = cnt_h > 0.0f ? ce_h / cnt_h : 0.0f;
  hidden_loss = hidden_loss * n_vis_total / n_hid_total * fnbh / fnbv * wn_hidden / wn_visible;
  const float locref_loss = locref_weight * (mcnt > 0.0f ? hub / mcnt : 0.0f);
  float total = visible_loss + hidden_loss + locref_loss;
  float ws_loss = 0.0f, wt_loss = 0.0f;
  if (nl > 0) {
    float acc = 0.0f;
    for (int l = 0; l < nl; ++l) {
      const int a = edges[2 * l], b = edges[2 * l + 1];
      for (int t = 0; t < nt; ++t) {
        const float* m = all + (size_t)t * nj * 2;
        const float dr = (m[2 * a] * stride + 0.5f * stride) - (m[2 * b] * stride + 0.5f * stride);
        const float dc = (m[2 * a + 1] * stride + 0.5f * stride) - (m[2 * b + 1] * stride + 0.5f * stride);
        const float d = sqrtf(dr * dr + dc * dc);
        acc += (fmaxf(d - ws_max[l], 0.0f) + ws_max[l]) * ws[l];
      }
    }
    ws_loss = acc / (float)H / (float)W * n_vis_total / fnbv / (n_vis_total + n_hid_total) / wn_visible;
    total += ws_loss;
  }
  if (wt > 0.0f && meanflow != nullptr) {
    float acc = 0.0f;
    for (int t = 0; t + 1 < nt; ++t)
      for (int j = 0; j < nj; ++j) {
        const float* m0 = all + ((size_t)t * nj + j) * 2;
        const float* m1 = all + ((size_t)(t + 1) * nj + j) * 2;
        const float dr = (m0[0] * stride + 0.5f * stride) - (m1[0] * stride + 0.5f * stride);
        const float dc = (m0[1] * stride + 0.5f * stride) - (m1[1] * stride + 0.5f * stride);
        const float d = sqrtf(dr * dr + dc * dc);
        float inv = fminf(1.0f / (meanflow[t * nj + j] + 1e-10f), 1.0f);
        inv = fminf(expf(logf(inv) * 3.0f), 1.0f);
        inv = inv * wt_batch[t] / (float)H / (float)W;
        const float v = (fmaxf(d - wt_max, 0.0f) + wt_max) * inv;
        acc += v * v;
      }
    wt_loss = sqrtf(acc) * n_vis_total / fnbv / (n_vis_total + n_hid_total) / wn_visible;
    total += wt_loss;
  }
  out[0] = visible_loss; out[1] = hidden_loss; out[2] = locref_loss; out[3] = ws_loss; out[4] = wt_loss; out[5] = total;
}


// ------------------------------------------------------------------------------------------------ backward
__device__ void block_max_idx(float v, int idx, float* shv, int* shi, float& out_v, int& out_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { shv[threadIdx.x >> 5] = v; shi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  out_v = shv[0];
  out_i = shi[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i)
    if (shv[i] > out_v || (shv[i] == out_v && shi[i] < out_i)) { out_v = shv[i]; out_i = shi[i]; }
}

__device__ __forceinline__ void blur_weights(int pos, int n, int radius, const float* kt, float& a, float& r) {
  a = 0.0f;
  r = 0.0f;
  for (int d = -radius; d <= radius; ++d) {
    const int dst = pos - d;
    if (dst >= 0 && dst < n) { a += kt[d + radius]; r += kt[d + radius] * (float)dst; }
  }
}

// d total_loss / d pred and d total_loss / d locref.  One CTA per listed marker, three passes over its plane:
//   A  max of the Gaussian bump (value + pixel) and of sigmoid(x) (confidence c + its pixel)
//   B  the marker-level sums that multiply dc and d mu
//   C  per-pixel gradient = direct term + (pixel == argmax) * confidence term + soft-argmax backward of dL/dmu
// Gradients flow through the labels of the cross entropy, the confidence max and the loss weights exactly as TF's
// graph does (no stop_gradient anywhere in fitdgp.py:947-1128).
__global__ void __launch_bounds__(kLossThreads) marker_loss_bwd_kernel(
    const float* __restrict__ pred, const float* __restrict__ locref, const float* __restrict__ locref_map,
    const float* __restrict__ locref_mask, const float* __restrict__ all, const float* __restrict__ mu,
    const float* __restrict__ norm, const int* __restrict__ visible, int nbv, const int* __restrict__ hidden, int nbh,
    const float4* __restrict__ part, int nt, int H, int W, int nj, float inv2l2, int gm2, int gm3, int locref_mse, float gamma,
    int radius,
    float sigma, const int* __restrict__ edges, int nl, const float* __restrict__ ws, const float* __restrict__ ws_max,
    float stride, float n_vis_total, float n_hid_total, float wn_visible, float wn_hidden, float locref_weight,
    int visible_only, const float* __restrict__ meanflow, const float4* __restrict__ boxgrad,
    const float* __restrict__ wt_batch, float wt_max, int Hin, int Win, const float* __restrict__ losses,
    float* __restrict__ g_pred, float* __restrict__ g_locref, int use_smem) {
  extern __shared__ float plane[];
  __shared__ float shf[kLossThreads / 32];
  __shared__ int shi[kLossThreads / 32];
  __shared__ float s_cnt[3];
  const bool is_vis = (int)blockIdx.x < nbv;
  const int m = is_vis ? visible[blockIdx.x] : hidden[blockIdx.x - nbv];
  const int t = m / nj, j = m - t * nj;
  const int HW = H * W;
  int xs_stride;
  const float* x0 = stage_plane(pred + (size_t)t * HW * nj + j, HW, nj, plane, use_smem, &xs_stride);
  float* g0 = g_pred + (size_t)t * HW * nj + j;
  const float fnbv = nbv > 0 ? (float)nbv : (float)nbh;

  // global normalisers from the forward partials (fixed order; fetched by the whole block first, the sum itself stays serial)
  __shared__ float2 s_yw[1024];
  const int nball = nbv + nbh;
  const bool staged = nball <= 1024;
  if (staged)
    for (int i = threadIdx.x; i < nball; i += blockDim.x) { const float4 q = part[i]; s_yw[i] = make_float2(q.y, q.w); }
  __syncthreads();
  if (threadIdx.x == 0) {
    float cv = 0.0f, ch = 0.0f, mc = 0.0f;
    if (staged) {
      for (int i = 0; i < nbv; ++i) { cv += s_yw[i].x; mc += s_yw[i].y; }
      for (int i = nbv; i < nball; ++i) ch += s_yw[i].x;
    } else {
      for (int i = 0; i < nbv; ++i) { cv += part[i].y; mc += part[i].w; }
      for (int i = nbv; i < nball; ++i) ch += part[i].y;
    }
    s_cnt[0] = cv; s_cnt[1] = ch; s_cnt[2] = mc;
  }
  __syncthreads();
  const float cnt_v = s_cnt[0], cnt_h = s_cnt[1], mcnt = s_cnt[2];
  const float tr = all[2 * m], tc = all[2 * m + 1];   // Gaussian target centre (label or soft-argmax)

  // ---- pass A
  float gm = -1.0f, cm = -CUDART_INF_F;
  int gi = 0x7fffffff, ci = 0x7fffffff;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    const float dr = (float)r - tr, dc = (float)c - tc;
    const float g = expf(-(dr * dr + dc * dc) * inv2l2);
    if (g > gm) { gm = g; gi = p; }
    if (!is_vis && gm2 != 0) {
      const float sg = sigmoidf_(x0[(size_t)p * xs_stride]);
      if (sg > cm) { cm = sg; ci = p; }
    }
  }
  float gmax; int pg;
  block_max_idx(gm, gi, shf, shi, gmax, pg);
  const float gm5 = gmax + 1e-5f;
  float conf = 1.0f; int pstar = -1;
  if (!is_vis && gm2 != 0) block_max_idx(cm, ci, shf, shi, conf, pstar);

  if (is_vis) {
    const float kv = cnt_v > 0.0f ? 1.0f / cnt_v : 0.0f;
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
      const int r = p / W, c = p - r * W;
      const float dr = (float)r - tr, dc = (float)c - tc;
      const float tg = expf(-(dr * dr + dc * dc) * inv2l2) / gm5;
      g0[(size_t)p * nj] = kv * (sigmoidf_(x0[(size_t)p * xs_stride]) - tg);
    }
    if (locref != nullptr && g_locref != nullptr) {
      const float kl = mcnt > 0.0f ? locref_weight / mcnt : 0.0f;
      const size_t base = (size_t)t * HW * 2 * nj + 2 * j;
#pragma unroll 4
      for (int p = threadIdx.x; p < 2 * HW; p += blockDim.x) {
        const size_t o = base + (size_t)(p >> 1) * 2 * nj + (p & 1);
        const float d = __ldg(locref + o) - __ldg(locref_map + o);
        g_locref[o] = kl * __ldg(locref_mask + o) * (locref_mse ? 2.0f * d : (fabsf(d) < 1.0f ? d : (d > 0.0f ? 1.0f : -1.0f)));
      }
    }
    return;
  }
  if (visible_only) {  // fit_dgp_labeledonly optimises total_loss_visible (fitdgp.py:416): hidden markers get no gradient
    for (int p = threadIdx.x; p < HW; p += blockDim.x) g0[(size_t)p * nj] = 0.0f;
    return;
  }

  // ---- hidden marker
  const bool scaled = gm3 == 3;                 // confidence-scaled logits + (1 - c) weights
  const float cs = gm2 == 1 ? conf : 1.0f;      // target scale
  const float w = scaled ? 1.0f - conf : 1.0f;
  const float khid = cnt_h > 0.0f ? (n_vis_total / n_hid_total * (float)nbh / fnbv * wn_hidden / wn_visible) / cnt_h : 0.0f;
  const float l2inv = 2.0f * inv2l2;            // 1 / lengthscale^2
  // pass B
  float S = 0.0f, Cc = 0.0f, Ur = 0.0f, Uc = 0.0f, V = 0.0f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    const float dr = (float)r - tr, dc = (float)c - tc;
    const float G = expf(-(dr * dr + dc * dc) * inv2l2);
    const float tg = G / gm5;
    const float x = x0[(size_t)p * xs_stride];
    const float sg = sigmoidf_(x);
    float xl = x, D = 0.0f;
    if (scaled) {
      const float ps = sg * conf;
      xl = -logf(1.0f - ps + 1e-20f) + logf(ps + 1e-20f);
      D = 1.0f / (1.0f - ps + 1e-20f) + 1.0f / (ps + 1e-20f);
    }
    const float z = tg * cs;
    S += bce(z, xl);
    Cc += (gm2 == 1 ? -xl * tg : 0.0f) + (scaled ? (sigmoidf_(xl) - z) * D * sg : 0.0f);
    Ur += -xl * G * dr;   // d G / d mu_r = G * (r - mu_r) / l^2
    Uc += -xl * G * dc;
    V += -xl * G;
  }
  S = block_sum(S, shf); Cc = block_sum(Cc, shf); Ur = block_sum(Ur, shf); Uc = block_sum(Uc, shf); V = block_sum(V, shf);
  const float coef_c = (scaled ? -S : 0.0f) + w * Cc;       // multiplies dc = sigma'(x[p*]) dx[p*]
  const int rg = pg / W, cg = pg - rg * W;
  // d L_m / d mu through the Gaussian target (incl. its max normaliser)
  float Lr = w * cs * (Ur * l2inv / gm5 - V * gmax * ((float)rg - tr) * l2inv / (gm5 * gm5));
  float Lc = w * cs * (Uc * l2inv / gm5 - V * gmax * ((float)cg - tc) * l2inv / (gm5 * gm5));
  Lr *= khid; Lc *= khid;
  // + spatial clique: d ws_loss / d mu of this (hidden) marker
  if (nl > 0) {
    const float kws = 1.0f / (float)H / (float)W * n_vis_total / fnbv / (n_vis_total + n_hid_total) / wn_visible;
    const float* mf = all + (size_t)t * nj * 2;
    for (int l = 0; l < nl; ++l) {
      const int a = edges[2 * l], b = edges[2 * l + 1];
      if (a != j && b != j) continue;
      const float er = (mf[2 * a] * stride + 0.5f * stride) - (mf[2 * b] * stride + 0.5f * stride);
      const float ec = (mf[2 * a + 1] * stride + 0.5f * stride) - (mf[2 * b + 1] * stride + 0.5f * stride);
      const float d = sqrtf(er * er + ec * ec);
      if (d > ws_max[l] && d > 0.0f) {
        const float sgn = (a == j) ? 1.0f : -1.0f;
        Lr += kws * ws[l] * sgn * stride * er / d;
        Lc += kws * ws[l] * sgn * stride * ec / d;
      }
    }
  }
  // + temporal clique (fitdgp.py:1078-1124): wt_loss = c * || (relu(delta - wt_max) + wt_max) * inv ||_F.  delta and,
  // through tf.image.crop_and_resize's box gradient, the flow weight inv both depend on mu of frames k and k+1.
  if (meanflow != nullptr && losses[4] > 0.0f) {
    const float cw = n_vis_total / fnbv / (n_vis_total + n_hid_total) / wn_visible;
    const float nxi = (float)Hin, nyi = (float)Win;
    for (int e = 0; e < 2; ++e) {
      const int k = e == 0 ? t : t - 1;          // pair (k, k+1); e == 0: this marker is the first frame of the pair
      if (k < 0 || k + 1 >= nt) continue;
      const float* a0 = all + ((size_t)k * nj + j) * 2;
      const float* a1 = all + ((size_t)(k + 1) * nj + j) * 2;
      const float r0 = a0[0] * stride + 0.5f * stride, c0 = a0[1] * stride + 0.5f * stride;
      const float r1 = a1[0] * stride + 0.5f * stride, c1 = a1[1] * stride + 0.5f * stride;
      const float rme = e == 0 ? r0 : r1, cme = e == 0 ? c0 : c1, rot = e == 0 ? r1 : r0, cot = e == 0 ? c1 : c0;
      const float dlt = sqrtf((r0 - r1) * (r0 - r1) + (c0 - c1) * (c0 - c1));
      const float av = fmaxf(dlt - wt_max, 0.0f) + wt_max;
      const float mf = meanflow[k * nj + j];
      const float u = 1.0f / (mf + 1e-10f);
      const float u1 = fminf(u, 1.0f);
      const float wk = wt_batch[k] / (float)H / (float)W;
      const float inv = fminf(expf(logf(u1) * 3.0f), 1.0f) * wk;
      const float v = av * inv;
      const float gv = cw * cw * v / losses[4];  // dL/dv = c * v / ||.||_F, ||.||_F = wt_loss / c
      if (dlt > wt_max && dlt > 0.0f) {
        Lr += gv * inv * (rme - rot) / dlt * stride;
        Lc += gv * inv * (cme - cot) / dlt * stride;
      }
      if (u <= 1.0f) {
        const float dinv = -3.0f * u1 * u1 * u * u * wk;   // d(u^3)/d mf = -3 u^4
        const float4 bg = boxgrad[k * nj + j];
        // which end of the pair owns the box edges (ties split the gradient, as reduce_min / reduce_max do)
        const float wmin_r = rme < rot ? 1.0f : (rme == rot ? 0.5f : 0.0f), wmax_r = rme > rot ? 1.0f : (rme == rot ? 0.5f : 0.0f);
        const float wmin_c = cme < cot ? 1.0f : (cme == cot ? 0.5f : 0.0f), wmax_c = cme > cot ? 1.0f : (cme == cot ? 0.5f : 0.0f);
        float dr_ = 0.0f, dc_ = 0.0f;
        if (fminf(r0, r1) - 10.0f > 0.0f) dr_ += wmin_r * bg.x / nxi;
        if (fmaxf(r0, r1) + 10.0f < nxi) dr_ += wmax_r * bg.z / nxi;
        if (fminf(c0, c1) - 10.0f > 0.0f) dc_ += wmin_c * bg.y / nyi;
        if (fmaxf(c0, c1) + 10.0f < nyi) dc_ += wmax_c * bg.w / nyi;
        Lr += gv * av * dinv * dr_ * stride;
        Lc += gv * av * dinv * dc_ * stride;
      }
    }
  }
  // pass C
  const float g2 = gamma * 1.4426950408889634f;
  const float m2 = norm[2 * m], s0 = norm[2 * m + 1];
  const float mur = mu[2 * m], muc = mu[2 * m + 1];
  float knorm = 0.0f;
  for (int d = -radius; d <= radius; ++d) knorm += expf(-0.5f * (d / sigma) * (d / sigma));
  float kt[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float d = (float)(i - radius);
    kt[i] = i <= 2 * radius ? expf(-0.5f * (d / sigma) * (d / sigma)) / knorm : 0.0f;
  }
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    const float dr = (float)r - tr, dc = (float)c - tc;
    const float tg = expf(-(dr * dr + dc * dc) * inv2l2) / gm5;
    const float x = x0[(size_t)p * xs_stride];
    const float sg = sigmoidf_(x);
    float direct;
    if (scaled) {
      const float ps = sg * conf;
      const float xl = -logf(1.0f - ps + 1e-20f) + logf(ps + 1e-20f);
      const float D = 1.0f / (1.0f - ps + 1e-20f) + 1.0f / (ps + 1e-20f);
      direct = w * (sigmoidf_(xl) - tg * cs) * D * sg * (1.0f - sg) * conf;
    } else {
      direct = w * (sg - tg * cs);
    }
    float g = khid * direct;
    if (p == pstar) g += khid * coef_c * sg * (1.0f - sg);
    // soft-argmax backward: mu = (sum e * R) / (sum e * A) with border-aware blur weights
    float ah = 1.0f, rh = (float)r, aw = 1.0f, rw = (float)c;
    if (r < radius || r >= H - radius) blur_weights(r, H, radius, kt, ah, rh);
    if (c < radius || c >= W - radius) blur_weights(c, W, radius, kt, aw, rw);
    const float e = exp2f(fmaf(x, g2, -m2));
    g += gamma * e / s0 * ((rh * aw - mur * ah * aw) * Lr + (ah * rw - muc * ah * aw) * Lc);
    g0[(size_t)p * nj] = g;
  }
}

}  // namespace

cudaError_t launch_dgp_loss(const LossArgs& a, cudaStream_t stream) {
  const int nm = a.nt * a.nj;
  cudaError_t e = cudaMemsetAsync(a.all_markers, 0, (size_t)nm * 2 * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  const int nb = a.nbv + a.nbh;
  if (nb > 0) {
    combine_markers_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(a.mu, a.targets, a.visible, a.nbv, a.hidden, a.nbh,
                                                                 a.vis_in_targets, a.all_markers);
    const size_t plane_bytes = (size_t)a.H * a.W * sizeof(float);
    const int use_smem = plane_bytes <= kPlaneSmemMax;
    if (use_smem) {
      cudaError_t ea = cudaFuncSetAttribute(marker_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlaneSmemMax);
      if (ea != cudaSuccess) return ea;
    }
    marker_loss_kernel<<<nb, kLossThreads, use_smem ? plane_bytes : 0, stream>>>(
        a.pred, a.locref, a.locref_map, a.locref_mask, a.all_markers, a.visible, a.nbv, a.hidden, a.H, a.W, a.nj,
        1.0f / (2.0f * a.lengthscale * a.lengthscale), a.gm2, a.gm3, a.locref_mse, a.partials, use_smem);
  }
  const bool temporal = a.wt > 0.0f && a.flow != nullptr && a.nt > 1;
  if (temporal) {
    const int nboxes = (a.nt - 1) * a.nj, nchunks = (a.Hin + kFlowRows - 1) / kFlowRows;
    const size_t wbytes = 3 * ((size_t)kFlowRows + a.Win) * sizeof(float);
    if (wbytes > 48 * 1024) {
      if (wbytes > kPlaneSmemMax) return cudaErrorInvalidValue;
      cudaError_t ea = cudaFuncSetAttribute(flow_box_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlaneSmemMax);
      if (ea != cudaSuccess) return ea;
    }
    flow_box_mean_kernel<<<dim3(nboxes, nchunks), kLossThreads, wbytes, stream>>>(a.flow, a.all_markers, a.nt, a.nj, a.Hin, a.Win,
                                                                                  a.stride, a.boxgrad != nullptr, a.flow_part);
    flow_box_finalize_kernel<<<nboxes, 32, 0, stream>>>(a.flow_part, nchunks, a.Hin, a.Win, a.meanflow, a.boxgrad);
  }
  const size_t fin_bytes = (size_t)nb * 16 + ((size_t)nm * 2 + (temporal ? (size_t)(a.nt - 1) * (a.nj + 1) : 0) + 4 * (size_t)a.nl) * 4 + 16;
  const int fin_smem = fin_bytes <= 48 * 1024;
  loss_finalize_kernel<<<1, 128, fin_smem ? fin_bytes : 0, stream>>>(
      a.partials, a.nbv, a.nbh, a.all_markers, a.nt, a.nj, a.H, a.W, a.edges, a.nl, a.ws, a.ws_max,
      temporal ? a.meanflow : nullptr, a.wt_batch, a.wt, a.wt_max, a.stride, a.n_vis_total, a.n_hid_total, a.wn_visible,
      a.wn_hidden, a.locref_weight, a.out, fin_smem);
  return cudaGetLastError();
}

cudaError_t launch_dgp_loss_backward(const LossArgs& a, const float* norm, float gamma, float gauss_len, int visible_only,
                                     float* g_pred, float* g_locref, cudaStream_t stream) {
  const int nb = a.nbv + a.nbh;
  const bool temporal = a.wt > 0.0f && a.flow != nullptr && a.nt > 1;
  cudaError_t e = cudaMemsetAsync(g_pred, 0, (size_t)a.nt * a.H * a.W * a.nj * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  if (g_locref) {
    e = cudaMemsetAsync(g_locref, 0, (size_t)a.nt * a.H * a.W * 2 * a.nj * sizeof(float), stream);
    if (e != cudaSuccess) return e;
  }
  const size_t plane_bytes = (size_t)a.H * a.W * sizeof(float);
  const int use_smem = plane_bytes <= kPlaneSmemMax;
  if (nb > 0 && use_smem) {
    e = cudaFuncSetAttribute(marker_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlaneSmemMax);
    if (e != cudaSuccess) return e;
  }
  if (nb > 0)
    marker_loss_bwd_kernel<<<nb, kLossThreads, use_smem ? plane_bytes : 0, stream>>>(
        a.pred, a.locref, a.locref_map, a.locref_mask, a.all_markers, a.mu, norm, a.visible, a.nbv, a.hidden, a.nbh,
        a.partials, a.nt, a.H, a.W, a.nj, 1.0f / (2.0f * a.lengthscale * a.lengthscale), a.gm2, a.gm3, a.locref_mse, gamma,
        (int)gauss_len, gauss_len, a.edges, a.nl, a.ws, a.ws_max, a.stride, a.n_vis_total, a.n_hid_total, a.wn_visible,
        a.wn_hidden, a.locref_weight, visible_only, temporal ? a.meanflow : nullptr, a.boxgrad, a.wt_batch, a.wt_max, a.Hin,
        a.Win, a.out, g_pred, g_locref, use_smem);
  return cudaGetLastError();
}

}  // namespace dgp
