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

// gen_idx_chunk / find_marker_index (src/deepgraphpose/dataset.py:157-239): the marker index vectors of a batch.  Marker id =
// frame_position * nj + joint; a marker of a visible frame whose label is NaN counts as hidden.  One CTA: every thread classifies
// markers id, id + blockDim, ... in increasing order, an in-block exclusive scan of the per-thread counts gives each thread its
// output offsets, so the three lists come out sorted (what np.sort / np.setdiff1d / np.nonzero produce) -- integer work, bit-exact.
// counts[0] = number of visible markers, counts[1] = number of hidden markers.
__global__ void __launch_bounds__(256) marker_indices_kernel(const int* __restrict__ vis_frames, int n_vis, const int* __restrict__ hid_frames,
                                                             int n_hid, const double* __restrict__ joint_loc, int nj, int nt,
                                                             int* __restrict__ visible_marker, int* __restrict__ hidden_marker,
                                                             int* __restrict__ visible_in_targets, int* __restrict__ counts) {
  extern __shared__ int sh[];          // [nt] frame -> row of joint_loc (visible), -1 (hidden), -2 (not in the batch); then scan scratch
  int* kind = sh;
  int* scan_v = sh + nt;
  int* scan_h = scan_v + blockDim.x;
  for (int t = threadIdx.x; t < nt; t += blockDim.x) kind[t] = -2;
  __syncthreads();
  for (int i = threadIdx.x; i < n_vis; i += blockDim.x) kind[vis_frames[i]] = i;
  for (int i = threadIdx.x; i < n_hid; i += blockDim.x) kind[hid_frames[i]] = -1;
  __syncthreads();
  const int total = nt * nj;
  // contiguous range of marker ids per thread keeps the outputs sorted after the scan
  const int per = (total + blockDim.x - 1) / blockDim.x;
  const int m0 = threadIdx.x * per, m1 = min(m0 + per, total);
  int cv = 0, ch = 0;
  for (int m = m0; m < m1; ++m) {
    const int k = kind[m / nj];
    if (k == -2) continue;
    const bool vis = k >= 0 && !isnan(joint_loc[((size_t)k * nj + (m % nj)) * 2]);
    cv += vis ? 1 : 0;
    ch += vis ? 0 : 1;
  }
  scan_v[threadIdx.x] = cv;
  scan_h[threadIdx.x] = ch;
  __syncthreads();
  if (threadIdx.x == 0) {              // <= 256 entries: a serial exclusive scan is the simplest exact thing
    int av = 0, ah = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) {
      const int tv = scan_v[i], th = scan_h[i];
      scan_v[i] = av; scan_h[i] = ah;
      av += tv; ah += th;
    }
    counts[0] = av;
    counts[1] = ah;
  }
  __syncthreads();
  int ov = scan_v[threadIdx.x], oh = scan_h[threadIdx.x];
  for (int m = m0; m < m1; ++m) {
    const int t = m / nj, j = m - t * nj;
    const int k = kind[t];
    if (k == -2) continue;
    if (k >= 0 && !isnan(joint_loc[((size_t)k * nj + j) * 2])) {
      visible_marker[ov] = m;
      // position in the (n_vis, nj) targets array: rows follow the SORTED visible frame positions = joint_loc's row order
      visible_in_targets[ov] = k * nj + j;
      ++ov;
    } else {
      hidden_marker[oh++] = m;
    }
  }
}

// calculate_motion_energy (src/deepgraphpose/dataset.py:29-43): mean over the bytes of a frame of |frame - previous|, where
// the reference subtracts uint8 arrays -- the difference wraps modulo 256 and np.abs is the identity on uint8.  Integer work:
// the kernel returns the exact sum of (cur - prev) & 0xFF per frame (uint64); the host divides by the byte count in double,
// which reproduces numpy's mean bit for bit.  Grid (frame, slice): every CTA streams 16 B per thread and iteration from the two
// frames (coalesced uint4 loads), sums the 16 byte differences with packed byte arithmetic, reduces with warp shuffles and
// adds one atomic per CTA (integer atomics: order-independent, exact).
__global__ void __launch_bounds__(256) motion_energy_kernel(const uint8_t* __restrict__ frames, size_t frame_bytes,
                                                            unsigned long long* __restrict__ sums) {
  const int t = blockIdx.x + 1;                       // frame 0 has no predecessor: its energy is 0
  const uint8_t* cur = frames + (size_t)t * frame_bytes;
  const uint8_t* prev = cur - frame_bytes;
  const size_t n16 = frame_bytes >> 4;
  unsigned long long acc = 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(cur) | reinterpret_cast<uintptr_t>(prev)) & 15) == 0;
  if (aligned) {
    const uint4* c4 = reinterpret_cast<const uint4*>(cur);
    const uint4* p4 = reinterpret_cast<const uint4*>(prev);
    for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.y * blockDim.x) {
      const uint4 a = __ldg(c4 + i), b = __ldg(p4 + i);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
      uint32_t s = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = __vsub4(aw[k], bw[k]);     // per-byte wrapping difference
        s += __vsadu4(d, 0u);                         // sum of the four bytes
      }
      acc += s;
    }
    if (blockIdx.y == 0)
      for (size_t i = (n16 << 4) + threadIdx.x; i < frame_bytes; i += blockDim.x) acc += (uint8_t)(cur[i] - prev[i]);
  } else {
    for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < frame_bytes; i += (size_t)gridDim.y * blockDim.x)
      acc += (uint8_t)(cur[i] - prev[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
    for (int w = 0; w < 8; ++w) tot += sh[w];
    atomicAdd(sums + t, tot);
  }
}

}  // namespace

cudaError_t launch_marker_indices(const int* vis_frames, int n_vis, const int* hid_frames, int n_hid, const double* joint_loc, int nj,
                                  int nt, int* visible_marker, int* hidden_marker, int* visible_in_targets, int* counts,
                                  cudaStream_t s) {
  const size_t smem = ((size_t)nt + 2 * 256) * sizeof(int);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  marker_indices_kernel<<<1, 256, smem, s>>>(vis_frames, n_vis, hid_frames, n_hid, joint_loc, nj, nt, visible_marker, hidden_marker,
                                              visible_in_targets, counts);
  return cudaGetLastError();
}

cudaError_t launch_motion_energy(const uint8_t* frames, int T, size_t frame_bytes, unsigned long long* sums, int num_sms,
                                 cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)T * sizeof(unsigned long long), s);
  if (e != cudaSuccess || T < 2) return e;
  // enough slices per frame to fill the GPU a few times over, at least 64 KB per CTA
  int slices = (int)((frame_bytes + 65535) / 65536);
  const int want = (8 * num_sms + (T - 2)) / (T - 1);
  if (slices > want) slices = want;
  if (slices < 1) slices = 1;
  motion_energy_kernel<<<dim3(T - 1, slices), 256, 0, s>>>(frames, frame_bytes, sums);
  return cudaGetLastError();
}

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
