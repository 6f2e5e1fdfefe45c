// Host-feeder replacements of the fit_dgp training step, sm_100a.
//   locref_targets : coord2map (src/deepgraphpose/dataset.py:246-271) + PoseDataset.compute_target_part_scoremap
//                    (PTF/dataset/pose_defaultdataset.py:220-266) + the scatter of the visible frames' maps into the
//                    whole-batch tensors (src/deepgraphpose/models/fitdgp.py:781-795), written straight into device
//                    memory: the reference builds these (nt,H,W,2nj) float64 arrays in Python loops and feeds them
//                    through feed_dict every step (6 MB of H2D per step at 747x832, nt = 10).
// The arithmetic is done in double exactly as the reference's numpy float64 code, then rounded once to float32 (what the
// float32 placeholder does), so the result is bit-identical to the reference's feed.
#include "kernels.cuh"

namespace dgp {

namespace {

// one CTA per (visible frame v, joint j); threads sweep the cells that can lie within the radius
__global__ void locref_targets_kernel(const double* __restrict__ joint_loc, const int* __restrict__ frame_idx, int nj, int H,
                                      int W, double stride, double thresh, double locref_scale, float* __restrict__ lmap,
                                      float* __restrict__ lmask) {
  const int v = blockIdx.x / nj, j = blockIdx.x - v * nj;
  const double lr = joint_loc[((size_t)v * nj + j) * 2], lc = joint_loc[((size_t)v * nj + j) * 2 + 1];
  if (isnan(lr) || isnan(lc)) return;                 // missing label (dataset.py:255-257)
  const double j_x = lc * 8 + 4, j_y = lr * 8 + 4;    // hard-coded *8+4 of dataset.py:252, flipped to (x, y)
  if (j_x + j_y == 0.0) return;                       // the reference's nan_to_num(...).sum != 0 filter
  const double half = stride / 2, thr2 = thresh * thresh;
  const int reach = (int)(thresh / stride) + 2;
  const int ci = (int)floor((j_x - half) / stride), cj = (int)floor((j_y - half) / stride);
  const int side = 2 * reach + 1;
  const int t = frame_idx[v];
  for (int k = threadIdx.x; k < side * side; k += blockDim.x) {
    const int jj = cj - reach + k / side, ii = ci - reach + k % side;
    if (jj < 0 || jj >= H || ii < 0 || ii >= W) continue;
    const double pt_x = ii * stride + half, pt_y = jj * stride + half;
    const double dx = j_x - pt_x, dy = j_y - pt_y;
    if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= thr2) {  // numpy rounds each product: no FMA contraction
      const size_t o = (((size_t)t * H + jj) * W + ii) * (size_t)(2 * nj) + 2 * j;
      lmask[o] = 1.0f;
      lmask[o + 1] = 1.0f;
      lmap[o] = (float)__dmul_rn(dx, locref_scale);
      lmap[o + 1] = (float)__dmul_rn(dy, locref_scale);
    }
  }
}

}  // namespace

cudaError_t launch_locref_targets(const double* joint_loc, const int* frame_idx, int n_vis, int nt, int nj, int H, int W,
                                  double stride, double pos_dist_thresh, double locref_stdev, float* lmap, float* lmask,
                                  cudaStream_t s) {
  const size_t bytes = (size_t)nt * H * W * 2 * nj * sizeof(float);
  cudaError_t e = cudaMemsetAsync(lmap, 0, bytes, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(lmask, 0, bytes, s);
  if (e != cudaSuccess) return e;
  if (n_vis > 0)
    locref_targets_kernel<<<n_vis * nj, 128, 0, s>>>(joint_loc, frame_idx, nj, H, W, stride, pos_dist_thresh,
                                                     1.0 / locref_stdev, lmap, lmask);
  return cudaGetLastError();
}

}  // namespace dgp
