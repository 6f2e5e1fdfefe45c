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
#include <stdio.h>

#include "ptx_sm100.cuh"

namespace dgp {

namespace {

using namespace ptx;

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

// DLC global peak = first arg-max of the fp32 sigmoid map (predict.py:62-77).  sigmoid_tf is within ~1.5 ulp of the
// true sigmoid, whose relative slope is 1 - sigmoid(x) >= 1 / (1 + e^xmax) on [x, xmax]; two logits further apart than
// delta = 8 ulp_rel * (1 + e^xmax) can therefore neither tie nor swap order, so only x >= xmax - delta needs the exact
// evaluation.  Once the map saturates (xmax >= 14) everything >= 14 is a candidate: sigmoid(14) is 14 ulp below 1.
__device__ __forceinline__ float dlc_candidate_threshold(float xmax) {
  return xmax >= 14.0f ? 14.0f : xmax - 1.0e-6f * (1.0f + expf(xmax));
}

// order-preserving float <-> int map (shared-memory atomicMax on floats)
__device__ __forceinline__ int enc_ordered(float x) {
  const int b = __float_as_int(x);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float dec_ordered(int b) { return __int_as_float(b >= 0 ? b : b ^ 0x7fffffff); }

// ---------------------------------------------------------------------------------------------------------------------
// Streaming soft-argmax partials.  Persistent CTAs (one per SM); the logit maps are pulled through a 4-stage ring of
// 40 KB shared-memory buffers with 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx), so every byte crosses
// HBM -> SM exactly once and ~160 KB per SM are in flight independent of what the math warps are doing.
//   warp 12      : producer (one elected lane): waits empty[s], issues the bulk copy of the next chunk
//   warps 8..11  : scouts: per-joint max of a landed chunk (shared-memory atomicMax on order-preserving ints) -> ready[s]
//   warps 0..7   : math: one pass over the chunk against the running max (FFMA + EX2 + 3 accumulates per element,
//                  log2 domain), DLC peak candidates (x >= min(max-2, 14)), then the blur border correction for the
//                  few pixels within `radius` of an edge, all from shared memory; arrive on empty[s]
// A job is one SEGMENT (<= kSegChunks chunks) of one frame; the segmentation depends only on (H, W, nj), so a frame's
// arithmetic and its merge order never depend on the batch size or the grid (bit-exact batch invariance).
// `tact` math threads are active with 4*tact % nj == 0: every thread's four float4 lanes keep a fixed joint.
// kSamePixel: nj % 4 == 0, the four lanes of a float4 belong to ONE pixel.  kDlc: track the DLC global sigmoid peak.
#ifndef DGP_SA_MAIN
#define DGP_SA_MAIN 512
#endif
constexpr int kStMain = DGP_SA_MAIN;       // math threads
constexpr int kStScout = 128;
constexpr int kStThreads = kStMain + kStScout + 32;
constexpr int kStStages = 4;
constexpr int kStStageFloats = 10240;  // 40 KB
constexpr int kSegChunks = 4;
constexpr int kMaxJoints = 128;
constexpr int kRedFloats = 4 * kStMain * 5;
constexpr int kRedbFloats = kStMain * 3;
constexpr size_t kStSmemBytes = (size_t)kStStages * kStStageFloats * 4 + kRedFloats * 4 + kRedbFloats * 4 +
                                3 * kStStages * kMaxJoints * 4 + kStStages * kMaxJoints * 4 + kMaxJoints * 4 + 64 * 4 +
                                3 * kStStages * 8 + 128;

// blur border weights computed from scratch (degenerate maps no larger than the kernel, where every pixel is border)
__device__ __noinline__ void border_weights_slow(int pos, int n, int radius, float sigma, float& a, float& r) {
  float knorm = 0.0f;
  for (int d = -radius; d <= radius; ++d) knorm += expf(-0.5f * (d / sigma) * (d / sigma));
  a = 0.0f;
  r = 0.0f;
  for (int d = -radius; d <= radius; ++d) {
    const int dst = pos - d;
    if (dst >= 0 && dst < n) {
      const float k = expf(-0.5f * (d / sigma) * (d / sigma)) / knorm;
      a += k;
      r += k * (float)dst;
    }
  }
}

// exact sigmoid of one DLC peak candidate; out of line: the call is rare and the expf + divide would otherwise be
// replicated 20x in the unrolled hot loop
__device__ __noinline__ float dlc_sigmoid(float x) { return sigmoid_tf(x); }

// Slices per joint of the job-end reduction (each becomes its own partial): all 16 math warps get work.
__host__ __device__ __forceinline__ int reduce_slices(int nj) { return nj <= (kStMain / 32) ? (kStMain / 32) / nj : 1; }

template <bool kSamePixel, bool kDlc>
__global__ void __launch_bounds__(kStThreads, 1) softargmax_stream_kernel(
    const float* __restrict__ logits, int B, int H, int W, int nj, float gamma, int radius, float sigma, int chunk_px,
    int nseg, int tact, int stact, SaPartial* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char st_smem[];
  float* stage = reinterpret_cast<float*>(st_smem);
  float* red = stage + kStStages * kStStageFloats;      // [4*kStMain][5]: s0, sr, sc, bsig, bidx per lane
  float* redb = red + kRedFloats;                       // [kStMain][3]: border corrections
  float* pm2 = redb + kRedbFloats;                      // [stages][kMaxJoints]: running max * gamma * log2e after the chunk
  float* pf = pm2 + kStStages * kMaxJoints;             // [stages][kMaxJoints]: rescale factor 2^(old - new) of the sums
  float* pthr = pf + kStStages * kMaxJoints;            // [stages][kMaxJoints]: DLC candidate threshold
  int* cmax = reinterpret_cast<int*>(pthr + kStStages * kMaxJoints);  // [stages][kMaxJoints]: chunk max, ordered ints
  float* jm = reinterpret_cast<float*>(cmax + kStStages * kMaxJoints);  // [kMaxJoints]: final m of the job
  float* btab = jm + kMaxJoints;                        // border weights: [4][16] (Ah, Rh, Aw, Rw) x 2*radius entries
  uint64_t* full = reinterpret_cast<uint64_t*>(btab + 64);
  uint64_t* ready = full + kStStages;
  uint64_t* empty = ready + kStStages;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kStStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], kStScout);
      mbar_init(&empty[s], kStMain);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int HW = H * W;
  const int seg_px = chunk_px * kSegChunks;
  const int njobs = B * nseg;
  const float g2 = gamma * 1.4426950408889634f;

  if (tid >= kStMain + kStScout) {
    // ------------------------------------------------------------------ producer
    if (tid != kStMain + kStScout) return;
    int it = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      const int b = job / nseg, seg = job - b * nseg;
      const int p_beg = seg * seg_px, p_end = min(HW, p_beg + seg_px);
      const float* frame = logits + (size_t)b * HW * nj;
      for (int c0 = p_beg; c0 < p_end; c0 += chunk_px, ++it) {
        const int c1 = min(c0 + chunk_px, p_end);
        const int s = it % kStStages, use = it / kStStages;
        if (use > 0) mbar_wait_backoff(&empty[s], (use - 1) & 1);
        const uint32_t bytes = (uint32_t)(c1 - c0) * nj * 4u;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_load_1d(stage + s * kStStageFloats, frame + (size_t)c0 * nj, bytes, &full[s]);
      }
    }
    return;
  }

  if (tid >= kStMain) {
    // ------------------------------------------------------------------ scouts
    // Per chunk: per-joint max (float4 reads with a fixed joint per lane, REDUX across the lanes of a warp that share
    // their joints, then shared-memory atomicMax on order-preserving ints); the first nj scouts then fold it into the
    // job's running max and publish (m2, rescale factor, DLC threshold) for the math warps.
    const int stid = tid - kStMain;
    const int lane = stid & 31;
    int P = nj;
    for (int a = 4, bb = nj; bb;) { const int r = a % bb; a = bb; bb = r; P = nj / a; }
    unsigned mask = 0;
    if (32 % P == 0)
      for (int l = lane % P; l < 32; l += P) mask |= 1u << l;
    const bool full_warp = ((stid | 31) < stact);  // REDUX needs every lane of the mask to be active
    int jq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) jq[q] = (4 * stid + q) % nj;
    int it = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      const int seg = job % nseg;
      const int p_beg = seg * seg_px, p_end = min(HW, p_beg + seg_px);
      float run_x = -CUDART_INF_F, run_m2 = -CUDART_INF_F;
      for (int c0 = p_beg; c0 < p_end; c0 += chunk_px, ++it) {
        const int n4 = (min(c0 + chunk_px, p_end) - c0) * nj / 4;
        const int s = it % kStStages, ph = (it / kStStages) & 1;
        // full[s] of this use implies empty[s] of the previous one: the math warps are done with stage s's tables
        mbar_wait_backoff(&full[s], ph);
        int* cm = cmax + s * kMaxJoints;
        if (stid < nj) cm[stid] = enc_ordered(-CUDART_INF_F);
        named_bar_sync(2, kStScout);
        if (stid < stact) {
          const float4* st4 = reinterpret_cast<const float4*>(stage + s * kStStageFloats);
          float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
          int f = stid;
          for (; f + 3 * stact < n4; f += 4 * stact) {
            const float4 v0 = st4[f], v1 = st4[f + stact], v2 = st4[f + 2 * stact], v3 = st4[f + 3 * stact];
            m.x = fmaxf(fmaxf(m.x, v0.x), fmaxf(v1.x, fmaxf(v2.x, v3.x)));
            m.y = fmaxf(fmaxf(m.y, v0.y), fmaxf(v1.y, fmaxf(v2.y, v3.y)));
            m.z = fmaxf(fmaxf(m.z, v0.z), fmaxf(v1.z, fmaxf(v2.z, v3.z)));
            m.w = fmaxf(fmaxf(m.w, v0.w), fmaxf(v1.w, fmaxf(v2.w, v3.w)));
          }
          for (; f < n4; f += stact) {
            const float4 v = st4[f];
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
          }
          int e[4] = {enc_ordered(m.x), enc_ordered(m.y), enc_ordered(m.z), enc_ordered(m.w)};
          if (mask != 0 && full_warp) {
#pragma unroll
            for (int q = 0; q < 4; ++q) e[q] = __reduce_max_sync(mask, e[q]);
            if (lane < P) {
#pragma unroll
              for (int q = 0; q < 4; ++q) atomicMax(&cm[jq[q]], e[q]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) atomicMax(&cm[jq[q]], e[q]);
          }
        }
        named_bar_sync(2, kStScout);
        if (stid < nj) {
          const float nx = fmaxf(run_x, dec_ordered(cm[stid]));
          const float nm2 = nx * g2;
          pm2[s * kMaxJoints + stid] = nm2;
          pf[s * kMaxJoints + stid] = (nm2 == run_m2) ? 1.0f : ex2_approx(run_m2 - nm2);  // first chunk: 2^-inf = 0
          if (kDlc) pthr[s * kMaxJoints + stid] = dlc_candidate_threshold(nx);
          run_x = nx;
          run_m2 = nm2;
        }
        mbar_arrive(&ready[s]);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- math warps
  const int L = 4 * tact;
  const bool lane_act = tid < tact;
  const int Gb = kStMain / nj;                   // border-pass pixel classes per joint
  const bool b_act = tid < Gb * nj;
  const int jb = tid % nj, gb = tid / nj;
  const bool all_border = (W <= 2 * radius) || (H <= 2 * radius);
  const int R2 = 2 * radius;
  // border weight tables: entry i < radius is position i, entry i >= radius is position n - 2*radius + i
  if (tid < 2 * R2 && !all_border) {
    float knorm = 0.0f;
    for (int d = -radius; d <= radius; ++d) knorm += expf(-0.5f * (d / sigma) * (d / sigma));
    const int axis = tid / R2, i = tid - axis * R2;
    const int n = axis == 0 ? H : W;
    const int pos = i < radius ? i : n - R2 + i;
    float a = 0.0f, r = 0.0f;
    for (int d = -radius; d <= radius; ++d) {
      const int dst = pos - d;
      if (dst >= 0 && dst < n) {
        const float k = expf(-0.5f * (d / sigma) * (d / sigma)) / knorm;
        a += k;
        r += k * (float)dst;
      }
    }
    btab[(2 * axis) * 16 + i] = a;
    btab[(2 * axis + 1) * 16 + i] = r;
  }
  named_bar_sync(1, kStMain);
  // border side-column walk: item k = gb + i*Gb -> (row offset k / R2, side index k % R2), kept incrementally
  const int bq0 = R2 > 0 ? gb / R2 : 0, bs0 = R2 > 0 ? gb - bq0 * R2 : 0;
  const int bdq = R2 > 0 ? Gb / R2 : 0, bds = R2 > 0 ? Gb - bdq * R2 : 0;
  int jq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) jq[q] = (4 * tid + q) % nj;
  const int dP = L / nj;                         // pixels between a thread's consecutive float4s
  const float dPr = (float)(dP / W), dPc = (float)(dP - (dP / W) * W);
  const int cqW = chunk_px / W, crW = chunk_px - cqW * W;   // chunk advance in (rows, cols)
  const float Wf = (float)W;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kLanes = kSamePixel ? 1 : 4;     // (row, col) trackers per thread
  const f32x2 g2g2 = pk2(g2, g2), dpos = pk2(dPr, dPc), wrapfix = pk2(1.0f, -Wf);
  const int nslice = reduce_slices(nj);

  int it = 0;
  for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
    const int b = job / nseg, seg = job - b * nseg;
    const int p_beg = seg * seg_px, p_end = min(HW, p_beg + seg_px);
    // per-lane state; the packed variants hold (joint 0, joint 1) / (joint 2, joint 3) or (row-sum, col-sum) pairs
    float m2[4], s0[4], sr[4], sc[4], thr[4], bsig[4];
    int bidx[4];
    f32x2 s0p[2], rcp[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m2[q] = 0.0f; thr[q] = CUDART_INF_F;
      s0[q] = sr[q] = sc[q] = 0.0f; bsig[q] = -1.0f; bidx[q] = 0x7fffffff;
      rcp[q] = pk2(0.0f, 0.0f);
    }
    s0p[0] = s0p[1] = pk2(0.0f, 0.0f);
    float mb = 0.0f, d0 = 0.0f, dr = 0.0f, dc = 0.0f;
    // (row, col) of the chunk origin (ints) and of this thread's first pixel in the current chunk (floats)
    int crow_i = p_beg / W, ccol_i = p_beg - crow_i * W;
    float crow[kLanes], ccol[kLanes];
#pragma unroll
    for (int q = 0; q < kLanes; ++q) {
      const int pix = p_beg + (4 * tid + q) / nj;
      crow[q] = (float)(pix / W);
      ccol[q] = (float)(pix - (pix / W) * W);
    }

    for (int c0 = p_beg; c0 < p_end; c0 += chunk_px, ++it) {
      const int c1 = min(c0 + chunk_px, p_end);
      const int s = it % kStStages, ph = (it / kStStages) & 1;
      mbar_wait(&full[s], ph);
      mbar_wait(&ready[s], ph);
      const float* st = stage + s * kStStageFloats;
      const int n4 = (c1 - c0) * nj / 4;
      const float4* st4 = reinterpret_cast<const float4*>(st);

      if (lane_act) {
        if constexpr (kSamePixel) {
          // the thread's four joints are consecutive and 16 B aligned in the published tables
          const float4 M = *reinterpret_cast<const float4*>(pm2 + s * kMaxJoints + jq[0]);
          const float4 F = *reinterpret_cast<const float4*>(pf + s * kMaxJoints + jq[0]);
          const f32x2 nm01 = pk2(-M.x, -M.y), nm23 = pk2(-M.z, -M.w);
          s0p[0] = mul2(s0p[0], pk2(F.x, F.y));
          s0p[1] = mul2(s0p[1], pk2(F.z, F.w));
          rcp[0] = mul2(rcp[0], pk2(F.x, F.x));
          rcp[1] = mul2(rcp[1], pk2(F.y, F.y));
          rcp[2] = mul2(rcp[2], pk2(F.z, F.z));
          rcp[3] = mul2(rcp[3], pk2(F.w, F.w));
          m2[0] = M.x; m2[1] = M.y; m2[2] = M.z; m2[3] = M.w;
          if (kDlc) {
            const float4 T = *reinterpret_cast<const float4*>(pthr + s * kMaxJoints + jq[0]);
            thr[0] = T.x; thr[1] = T.y; thr[2] = T.z; thr[3] = T.w;
          }
          f32x2 pos = pk2(crow[0], ccol[0]);
          auto consume = [&](const float4& v) {
            float t0, t1, t2, t3;
            upk2(fma2(pk2(v.x, v.y), g2g2, nm01), t0, t1);
            upk2(fma2(pk2(v.z, v.w), g2g2, nm23), t2, t3);
            const float e0 = ex2_approx(t0), e1 = ex2_approx(t1), e2 = ex2_approx(t2), e3 = ex2_approx(t3);
            s0p[0] = add2(s0p[0], pk2(e0, e1));
            s0p[1] = add2(s0p[1], pk2(e2, e3));
            rcp[0] = fma2(pos, pk2(e0, e0), rcp[0]);
            rcp[1] = fma2(pos, pk2(e1, e1), rcp[1]);
            rcp[2] = fma2(pos, pk2(e2, e2), rcp[2]);
            rcp[3] = fma2(pos, pk2(e3, e3), rcp[3]);
            if (kDlc) {
              if ((v.x >= thr[0]) | (v.y >= thr[1]) | (v.z >= thr[2]) | (v.w >= thr[3])) {
                float pr, pc;
                upk2(pos, pr, pc);
                const int idx = (int)pr * W + (int)pc;
                const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (xs[q] >= thr[q]) {
                    const float sg = dlc_sigmoid(xs[q]);
                    if (sg > bsig[q] || (sg == bsig[q] && idx < bidx[q])) { bsig[q] = sg; bidx[q] = idx; }
                  }
                }
              }
            }
            pos = add2(pos, dpos);
            float pr, pc;
            upk2(pos, pr, pc);
            if (pc >= Wf) pos = add2(pos, wrapfix);
          };
          int f = tid;
          for (; f + 3 * tact < n4; f += 4 * tact) {
            const float4 v0 = st4[f], v1 = st4[f + tact], v2 = st4[f + 2 * tact], v3 = st4[f + 3 * tact];
            consume(v0); consume(v1); consume(v2); consume(v3);
          }
          for (; f < n4; f += tact) consume(st4[f]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float f = pf[s * kMaxJoints + jq[q]];
            m2[q] = pm2[s * kMaxJoints + jq[q]];
            s0[q] *= f; sr[q] *= f; sc[q] *= f;
            if (kDlc) thr[q] = pthr[s * kMaxJoints + jq[q]];
          }
          float frow[kLanes], fcol[kLanes];
#pragma unroll
          for (int q = 0; q < kLanes; ++q) { frow[q] = crow[q]; fcol[q] = ccol[q]; }
          auto consume = [&](const float4& v) {
            const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float e = ex2_approx(fmaf(xs[q], g2, -m2[q]));
              s0[q] += e;
              sr[q] = fmaf(e, frow[q % kLanes], sr[q]);
              sc[q] = fmaf(e, fcol[q % kLanes], sc[q]);
            }
            if (kDlc) {
              if ((xs[0] >= thr[0]) | (xs[1] >= thr[1]) | (xs[2] >= thr[2]) | (xs[3] >= thr[3])) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (xs[q] >= thr[q]) {
                    const float sg = dlc_sigmoid(xs[q]);
                    const int idx = (int)frow[q % kLanes] * W + (int)fcol[q % kLanes];
                    if (sg > bsig[q] || (sg == bsig[q] && idx < bidx[q])) { bsig[q] = sg; bidx[q] = idx; }
                  }
                }
              }
            }
#pragma unroll
            for (int q = 0; q < kLanes; ++q) {
              fcol[q] += dPc; frow[q] += dPr;
              if (fcol[q] >= Wf) { fcol[q] -= Wf; frow[q] += 1.0f; }
            }
          };
          int f = tid;
          for (; f + 3 * tact < n4; f += 4 * tact) {
            const float4 v0 = st4[f], v1 = st4[f + tact], v2 = st4[f + 2 * tact], v3 = st4[f + 3 * tact];
            consume(v0); consume(v1); consume(v2); consume(v3);
          }
          for (; f < n4; f += tact) consume(st4[f]);
        }
        // advance this thread's chunk origin
#pragma unroll
        for (int q = 0; q < kLanes; ++q) {
          ccol[q] += (float)crW; crow[q] += (float)cqW;
          if (ccol[q] >= Wf) { ccol[q] -= Wf; crow[q] += 1.0f; }
        }
      }

      if (b_act) {
        // blur border correction: pixels within `radius` of an edge lose the taps that fall outside, i.e. their weights
        // are (Ah*Aw, Rh*Aw, Ah*Rw) instead of the interior (1, row, col) added above
        {
          const float f = pf[s * kMaxJoints + jb];
          mb = pm2[s * kMaxJoints + jb];
          d0 *= f; dr *= f; dc *= f;
        }
        auto add_px = [&](int r, int c) {
          const int p = r * W + c;
          if (p < c0 || p >= c1) return;
          const float e = ex2_approx(fmaf(st[(p - c0) * nj + jb], g2, -mb));
          float ah = 1.0f, rh = (float)r, aw = 1.0f, rw = (float)c;
          if (all_border) {
            border_weights_slow(r, H, radius, sigma, ah, rh);
            border_weights_slow(c, W, radius, sigma, aw, rw);
          } else {
            const int ir = r < radius ? r : (r >= H - radius ? r - (H - R2) : -1);
            const int ic = c < radius ? c : (c >= W - radius ? c - (W - R2) : -1);
            if (ir >= 0) { ah = btab[ir]; rh = btab[16 + ir]; }
            if (ic >= 0) { aw = btab[32 + ic]; rw = btab[48 + ic]; }
          }
          d0 += e * (ah * aw - 1.0f);
          dr += e * (rh * aw - (float)r);
          dc += e * (ah * rw - (float)c);
        };
        const int ra = crow_i;
        int rb;
        if (c1 - c0 == chunk_px) rb = ra + cqW + ((ccol_i + crW - 1 >= W) ? 1 : 0) - (crW == 0 && ccol_i == 0 ? 1 : 0);
        else rb = (c1 - 1) / W;
        if (!all_border) {
          int rq = bq0, sx = bs0;                       // left / right columns of every row of the chunk
          for (; ra + rq <= rb; ) {
            add_px(ra + rq, sx < radius ? sx : W - R2 + sx);
            rq += bdq; sx += bds;
            if (sx >= R2) { sx -= R2; rq += 1; }
          }
          // top / bottom rows: the full span between the side columns
          for (int r = ra; r <= min(rb, radius - 1); ++r)
            for (int c = radius + gb; c < W - radius; c += Gb) add_px(r, c);
          for (int r = max(ra, H - radius); r <= rb; ++r)
            for (int c = radius + gb; c < W - radius; c += Gb) add_px(r, c);
        } else {
          for (int r = ra; r <= rb; ++r)
            for (int c = gb; c < W; c += Gb) add_px(r, c);
        }
      }
      crow_i += cqW; ccol_i += crW;
      if (ccol_i >= W) { ccol_i -= W; crow_i += 1; }
      mbar_arrive(&empty[s]);
    }

    // ---- job end: deterministic block reduction (fixed order), nslice partials per (frame, segment, joint)
    if (lane_act) {
      if constexpr (kSamePixel) {
        upk2(s0p[0], s0[0], s0[1]);
        upk2(s0p[1], s0[2], s0[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) upk2(rcp[q], sr[q], sc[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float* r = red + (4 * tid + q) * 5;
        r[0] = s0[q]; r[1] = sr[q]; r[2] = sc[q]; r[3] = bsig[q]; r[4] = __int_as_float(bidx[q]);
      }
    }
    if (b_act) {
      float* r = redb + tid * 3;
      r[0] = d0; r[1] = dr; r[2] = dc;
      if (gb == 0) jm[jb] = mb;
    }
    named_bar_sync(1, kStMain);
    {
      const int per_joint = L / nj;                       // lanes holding joint j: e = j + nj * k, k < per_joint
      const int kper = (per_joint + nslice - 1) / nslice;
      for (int w = warp; w < nj * nslice; w += kStMain / 32) {
        const int j = w % nj, sl = w / nj;
        float a0 = 0.0f, ar = 0.0f, ac = 0.0f, bs = -1.0f;
        int bi = 0x7fffffff;
        const int k1 = min(per_joint, (sl + 1) * kper);
        for (int k = sl * kper + lane; k < k1; k += 32) {
          const float* r = red + (j + nj * k) * 5;
          a0 += r[0]; ar += r[1]; ac += r[2];
          if (kDlc) {
            const int idx = __float_as_int(r[4]);
            if (r[3] > bs || (r[3] == bs && idx < bi)) { bs = r[3]; bi = idx; }
          }
        }
        if (sl == 0) {
          for (int g = lane; g < Gb; g += 32) {
            const float* r = redb + (j + nj * g) * 3;
            a0 += r[0]; ar += r[1]; ac += r[2];
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a0 += __shfl_xor_sync(0xffffffffu, a0, o);
          ar += __shfl_xor_sync(0xffffffffu, ar, o);
          ac += __shfl_xor_sync(0xffffffffu, ac, o);
          if (kDlc) {
            const float obs = __shfl_xor_sync(0xffffffffu, bs, o);
            const int obi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (obs > bs || (obs == bs && obi < bi)) { bs = obs; bi = obi; }
          }
        }
        if (lane == 0) {
          SaPartial& o = part[(((size_t)b * nseg + seg) * nslice + sl) * nj + j];
          o.m = jm[j]; o.s0 = a0; o.sr = ar; o.sc = ac; o.bsig = bs; o.bidx = bi;
        }
      }
    }
    named_bar_sync(1, kStMain);  // the next job's partials reuse red / redb / jm
  }
}

// One WARP per (frame, joint): merge the segment partials (fixed lane order -> deterministic), then lane 0 does the
// O(1) read-outs.
__global__ void softargmax_finalize_kernel(const float* __restrict__ logits, const float* __restrict__ locref, int B,
                                           int H, int W, int nj, int splits, const SaPartial* __restrict__ part,
                                           float stride, float locref_stdev,
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
  int t = (kStMain / step) * step;
  return t;
}

static int scout_tact(int nj) {
  int g = nj;
  for (int a = 4, b = nj; b;) { int r = a % b; a = b; b = r; g = a; }
  const int step = nj / g;
  return (kStScout / step) * step;
}

static int softargmax_chunk_px(int nj) { return (kStStageFloats / nj) & ~3; }

// Segments per frame: a function of the map shape only (never of the batch size) -> batch-invariant arithmetic.
static int softargmax_segments(int H, int W, int nj) {
  const int seg_px = softargmax_chunk_px(nj) * kSegChunks;
  return (H * W + seg_px - 1) / seg_px;
}

int softargmax_splits(int H, int W, int nj) {
  if (nj < 1 || nj > kMaxJoints) return 1;
  return softargmax_segments(H, W, nj) * reduce_slices(nj);
}

template <bool kSamePixel, bool kDlc>
static cudaError_t launch_stream(const float* logits, int B, int H, int W, int nj, float gamma, int radius, float sigma,
                                 int chunk_px, int nseg, int tact, int stact, SaPartial* ws, int grid, cudaStream_t stream) {
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(softargmax_stream_kernel<kSamePixel, kDlc>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  softargmax_stream_kernel<kSamePixel, kDlc><<<grid, kStThreads, kStSmemBytes, stream>>>(
      logits, B, H, W, nj, gamma, radius, sigma, chunk_px, nseg, tact, stact, ws);
  return cudaGetLastError();
}

cudaError_t launch_softargmax(const float* logits, const float* locref, int B, int H, int W, int nj, float gamma,
                              float gauss_len, float stride, float locref_stdev, SaPartial* workspace, int splits,
                              float* mu, int* peak, float* lik, int* dlc_peak, float* dlc_pose, float* norm,
                              cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const int tact = softargmax_tact(nj);
  const int radius = (int)gauss_len;
  if (tact <= 0 || nj > kMaxJoints || (H & 1) || (W & 1) || radius > 4 || ((uintptr_t)logits & 15))
    return cudaErrorInvalidValue;
  const int nseg = softargmax_segments(H, W, nj);
  if (nseg * reduce_slices(nj) != splits) return cudaErrorInvalidValue;  // workspace sized with softargmax_splits()
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  const int njobs = B * nseg;
  const int grid = njobs < num_sms ? njobs : num_sms;
  const int chunk_px = softargmax_chunk_px(nj);
  const int stact = scout_tact(nj);
  const bool dlc = dlc_peak != nullptr || dlc_pose != nullptr;
  cudaError_t e;
  if (nj % 4 == 0)
    e = dlc ? launch_stream<true, true>(logits, B, H, W, nj, gamma, radius, gauss_len, chunk_px, nseg, tact, stact, workspace, grid, stream)
            : launch_stream<true, false>(logits, B, H, W, nj, gamma, radius, gauss_len, chunk_px, nseg, tact, stact, workspace, grid, stream);
  else
    e = dlc ? launch_stream<false, true>(logits, B, H, W, nj, gamma, radius, gauss_len, chunk_px, nseg, tact, stact, workspace, grid, stream)
            : launch_stream<false, false>(logits, B, H, W, nj, gamma, radius, gauss_len, chunk_px, nseg, tact, stact, workspace, grid, stream);
  if (e != cudaSuccess) return e;
  const int n = B * nj;
  softargmax_finalize_kernel<<<(n + 3) / 4, 128, 0, stream>>>(logits, locref, B, H, W, nj, splits, workspace, stride,
                                                              locref_stdev, mu, peak, lik, dlc_peak, dlc_pose, norm);
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
