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
// Streaming soft-argmax partials.  Persistent CTAs (one per SM).  The logit maps are cut into 8 KB CHUNKS (a function
// of (H, W, nj) only, so a frame's arithmetic never depends on the batch size or the grid: bit-exact batch invariance);
// the CTA's contiguous chunk range goes through a 22-stage shared-memory ring filled by 1-D bulk async copies
// (cp.async.bulk + mbarrier complete_tx): every byte crosses HBM -> SM exactly once and ~80 KB per SM are in flight no
// matter what the math is doing.  Every warp takes the next chunk of the CTA's range from a shared counter and is otherwise
// autonomous -- no CTA-wide barrier, no producer warp (the warp that drains a stage refills it):
//   wait full[s] -> pass 1 (per-joint max, from shared memory) -> pass 2 (packed fp32x2: FFMA2 + EX2 + FADD2 + FFMA2 per
//   element pair, log2 domain; blur border correction for pixels within `radius` of an edge; DLC peak candidates)
//   -> release the stage -> warp-level reduction over the lanes that share a joint -> one partial per (chunk, joint).
// Lane l reads float4s l, l + tw, ...; `tw` <= 32 lanes are active with 4*tw % nj == 0, so a lane's four float4 slots
// keep a fixed joint.  kSamePixel: nj % 4 == 0, the four slots belong to ONE pixel.  kShfl: the lanes that share a
// joint are an xor-closed set (nj in {4, 8, ..., 128}) and reduce with shuffles; otherwise through a per-warp scratch.
// Round 2: 16 warps / 27 stages (was 12 / 22).  ncu showed the kernel issue-bound at 3 warps per scheduler (62 % issue
// utilisation, 18 % occupancy), not HBM-bound: a fourth warp per scheduler needs <= 128 registers per thread and the shared
// memory that the per-warp reduction scratch used to take -- the scratch now lives in the stage the warp has just drained
// (it is refilled right after the reduction instead of right before it).
constexpr int kWWarps = 16;
constexpr int kWThreads = kWWarps * 32;
constexpr int kWChunkFloats = 2048;  // 8 KB
constexpr int kWStages = 27;
constexpr int kMaxJoints = 128;
constexpr int kWScratchFloats = 128 * 5 + kMaxJoints;   // [4*32][5] partial entries + [nj] results, inside the drained stage
static_assert(kWScratchFloats <= kWChunkFloats, "the reduction scratch must fit into one stage");
constexpr size_t kWSmemBytes = (size_t)kWStages * kWChunkFloats * 4 + 64 * 4 + kWStages * 8 + kWStages * 4 + 4 + 128;

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
// replicated in the unrolled hot loop
__device__ __noinline__ float dlc_sigmoid(float x) { return sigmoid_tf(x); }

// Multi-value butterfly: reduces the 12 per-lane sums over the lanes that share a joint (xor offsets 16 .. kP) with
// half of the values travelling each step (24 shuffles for 12 values over 32 lanes instead of 60).  On return lane l
// holds value index `vi` (or -1) in `out0`, and for kP == 4 a second one (`vi1`, `out1`).  Fixed tree -> deterministic.
template <int kP>
__device__ __forceinline__ void butterfly12(const float (&v)[12], int lane, float (&fin)[12], int& cnt, int& base) {
  constexpr unsigned F = 0xffffffffu;
  cnt = 12; base = 0;
#pragma unroll
  for (int k = 0; k < 12; ++k) fin[k] = v[k];
  if constexpr (kP <= 16) {
    const bool up = lane & 16;
    float w[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float snd = up ? fin[k] : fin[6 + k], kp = up ? fin[6 + k] : fin[k];
      w[k] = kp + __shfl_xor_sync(F, snd, 16);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) fin[k] = w[k];
    cnt = 6; base = up ? 6 : 0;
  }
  if constexpr (kP <= 8) {
    const bool up = lane & 8;
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float snd = up ? fin[k] : fin[3 + k], kp = up ? fin[3 + k] : fin[k];
      w[k] = kp + __shfl_xor_sync(F, snd, 8);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) fin[k] = w[k];
    cnt = 3; base += up ? 3 : 0;
  }
  if constexpr (kP <= 4) {
    // 3 -> 2: the lower lanes keep {0, 1}, the upper lanes keep {2, (nothing)}
    const bool up = lane & 4;
    const float snd0 = up ? fin[0] : fin[2], kp0 = up ? fin[2] : fin[0];
    const float snd1 = up ? fin[1] : 0.0f, kp1 = up ? 0.0f : fin[1];
    const float w0 = kp0 + __shfl_xor_sync(F, snd0, 4), w1 = kp1 + __shfl_xor_sync(F, snd1, 4);
    fin[0] = w0; fin[1] = w1;
    cnt = up ? 1 : 2; base += up ? 2 : 0;
  }
  if constexpr (kP <= 2) {
    const bool up = lane & 2;
    const float snd = up ? fin[0] : fin[1], kp = up ? fin[1] : fin[0];
    fin[0] = kp + __shfl_xor_sync(F, snd, 2);
    // lanes that came in with one value: the lower one keeps it, the upper one holds nothing
    if (cnt == 1) { cnt = up ? 0 : 1; } else { base += up ? 1 : 0; cnt = 1; }
  }
  if constexpr (kP <= 1) {
    fin[0] += __shfl_xor_sync(F, fin[0], 1);
    if (lane & 1) cnt = 0;   // the even lane writes
  }
}

// kP > 0: the lanes sharing a joint are {l : l % kP == lane % kP} (nj == 4 * kP) and reduce with shuffles;
// kP == 0: arbitrary joint count, reduction through the per-warp scratch.
template <bool kSamePixel, int kP, bool kDlc>
__global__ void __launch_bounds__(kWThreads, 1) softargmax_stream_kernel(
    const float* __restrict__ logits, int B, int H, int W, int nj, float gamma, int radius, float sigma, int chunk_px,
    int cpf, int tw, SaPartial* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char st_smem[];
  float* stage = reinterpret_cast<float*>(st_smem);
  float* btab = stage + kWStages * kWChunkFloats;       // border weights: [4][16] (Ah, Rh, Aw, Rw) x 2*radius entries
  uint64_t* full = reinterpret_cast<uint64_t*>(btab + 64);
  volatile int* issued = reinterpret_cast<volatile int*>(full + kWStages);   // loads issued into each stage so far
  int* next_chunk = const_cast<int*>(issued) + kWStages;                     // next chunk of this CTA's range to hand out

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool all_border = (W <= 2 * radius) || (H <= 2 * radius);
  const int R2 = 2 * radius;
  if (tid == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(&full[s], 1);
      issued[s] = 0;
    }
    *next_chunk = 0;
    fence_mbar_init();
  }
  __syncthreads();

  const int HW = H * W;
  const long long nchunks = (long long)B * cpf;
  const int g0 = (int)(nchunks * blockIdx.x / gridDim.x), g1 = (int)(nchunks * (blockIdx.x + 1) / gridDim.x);
  const int nloc = g1 - g0;

  // There is no producer warp and no "empty" barrier: the warp that finishes chunk i refills its stage with chunk
  // i + kWStages right away (it knows the stage is free), so a slow warp never blocks the loads of the others.
  auto issue_load_at = [&](int i, int b, int c) {   // one lane
    const int c0 = c * chunk_px, npx = min(chunk_px, HW - c0);
    const int s = i % kWStages;
    const uint32_t bytes = (uint32_t)npx * nj * 4u;
    mbar_arrive_expect_tx(&full[s], bytes);
    bulk_load_1d(stage + s * kWChunkFloats, logits + ((size_t)b * HW + c0) * nj, bytes, &full[s]);
    // A parity wait cannot tell "phase u pending" from "phase u-2 pending": the consumer of use u first checks this
    // counter, i.e. that use u's transaction has been registered on the barrier.
    __threadfence_block();
    issued[s] = i / kWStages + 1;
  };
  if (lane == 0)
    for (int i = warp; i < min(nloc, kWStages); i += kWWarps) issue_load_at(i, (g0 + i) / cpf, (g0 + i) % cpf);

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
  __syncthreads();   // (the first loads are already in flight: the table set-up hides under their latency)


  // -------------------------------------------------------------------- math warps
  constexpr bool kShfl = kP > 0;
  const float g2 = gamma * 1.4426950408889634f;
  const bool act = lane < tw;
  int P = kP;                                   // lanes l, l' share their joints iff l == l' (mod P)
  if (!kShfl)
    for (int a = 4, bb = nj; bb;) { const int r = a % bb; a = bb; bb = r; P = nj / a; }
  int jq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) jq[q] = (4 * lane + q) % nj;
  // nj % 4 != 0 only: byte qp of peer[q] = the lane class (= its first lane) whose slot qp holds joint jq[q], 0xff if none
  unsigned peer[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  if constexpr (!kSamePixel) {
    for (int q = 0; q < 4; ++q)
      for (int qp = 0; qp < 4; ++qp)
        for (int c = 0; c < P; ++c)
          if ((4 * c + qp) % nj == jq[q]) {
            peer[q] = (peer[q] & ~(0xffu << (8 * qp))) | ((unsigned)c << (8 * qp));
            break;
          }
  }
  const int dP = 4 * tw / nj;                    // pixels between a lane's consecutive float4s
  const float dPr = (float)(dP / W), dPc = (float)(dP - (dP / W) * W);
  const float Wf = (float)W, invW = 1.0f / (float)W;
  const float Rf = (float)radius, HmR = (float)(H - radius), WmR = (float)(W - radius);
  constexpr int kPos = kSamePixel ? 1 : 4;       // (row, col) trackers per lane
  const f32x2 g2g2 = pk2(g2, g2), dpos = pk2(dPr, dPc), wrapfix = pk2(1.0f, -Wf);

  // (Ah*Aw - 1, Rh*Aw - r, Ah*Rw - c) of a border pixel: what its blur weights differ from the interior (1, r, c) by
  auto border_corr = [&](float pr, float pc, float& k0, float& kr, float& kc) {
    const int r = (int)pr, c = (int)pc;
    float ah = 1.0f, rh = pr, aw = 1.0f, rw = pc;
    if (all_border) {
      border_weights_slow(r, H, radius, sigma, ah, rh);
      border_weights_slow(c, W, radius, sigma, aw, rw);
    } else {
      const int ir = r < radius ? r : (r >= H - radius ? r - (H - R2) : -1);
      const int ic = c < radius ? c : (c >= W - radius ? c - (W - R2) : -1);
      if (ir >= 0) { ah = btab[ir]; rh = btab[16 + ir]; }
      if (ic >= 0) { aw = btab[32 + ic]; rw = btab[48 + ic]; }
    }
    k0 = ah * aw - 1.0f;
    kr = rh * aw - pr;
    kc = ah * rw - pc;
  };

  // lanes that share this lane's joints (REDUX masks); kShfl guarantees 32 % P == 0
  unsigned cls_mask = 0;
  if (kShfl)
    for (int l = lane % P; l < 32; l += P) cls_mask |= 1u << l;
  const int bslot = lane / P, bcls = lane - (lane / P) * P, blanes = tw / P;   // border pass: pixel slot / granule class
  const int brr0 = R2 > 0 ? bslot / R2 : 0, bsx0 = R2 > 0 ? bslot - brr0 * R2 : 0;   // side walk: (row offset, side index)
  const int bdrr = R2 > 0 ? blanes / R2 : 0, bdsx = R2 > 0 ? blanes - bdrr * R2 : 0;

  // Chunks are handed out in order to whichever warp is free (one shared counter): the oldest load is always the next one
  // consumed, so a warp that falls behind no longer delays the refill another warp is about to wait for.  A chunk's
  // arithmetic does not depend on the warp that runs it.
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(next_chunk, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= nloc) break;
    const int b = (g0 + i) / cpf, c = (g0 + i) - b * cpf;
    const int c0 = c * chunk_px, npx = min(chunk_px, HW - c0);
    const int n4 = npx * nj / 4;
    const int s = i % kWStages, ph = (i / kWStages) & 1;
    const float4* st4 = reinterpret_cast<const float4*>(stage + s * kWChunkFloats);
    if (issued[s] <= i / kWStages) {
      const uint64_t t0 = globaltimer_ns();
      while (issued[s] <= i / kWStages) {
        if (globaltimer_ns() - t0 > 4000000000ull) {
          printf("dgp_b200: soft-argmax stage %d never refilled (block %d warp %d chunk %d)\n", s, blockIdx.x, warp, i);
          __trap();
        }
      }
    }
    __threadfence_block();
    mbar_wait(&full[s], ph);

    float xm[4], m2[4], mout[4], thr[4], s0[4], sr[4], sc[4], bsig[4];
    int bidx[4];
    // Attempt 0 (no DLC peak wanted): no max pass at all -- softmax numerators 2^(x*g2) against the fixed reference 0,
    // valid while the largest exponent of the chunk stays inside [-90, 90] (any |logit*gamma| < 62).  Otherwise, and
    // whenever the DLC peak needs the true maximum, attempt 1: exact per-joint max first, then the same pass.
    for (int attempt = kDlc ? 1 : 0; attempt < 2; ++attempt) {
      if (attempt == 1) {
        // ---- pass 1: per-joint max of the chunk
        float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        if (act) {
          int f = lane;
          for (; f + 3 * tw < n4; f += 4 * tw) {
            const float4 v0 = st4[f], v1 = st4[f + tw], v2 = st4[f + 2 * tw], v3 = st4[f + 3 * tw];
            m.x = fmaxf(fmaxf(m.x, v0.x), fmaxf(v1.x, fmaxf(v2.x, v3.x)));
            m.y = fmaxf(fmaxf(m.y, v0.y), fmaxf(v1.y, fmaxf(v2.y, v3.y)));
            m.z = fmaxf(fmaxf(m.z, v0.z), fmaxf(v1.z, fmaxf(v2.z, v3.z)));
            m.w = fmaxf(fmaxf(m.w, v0.w), fmaxf(v1.w, fmaxf(v2.w, v3.w)));
          }
          for (; f < n4; f += tw) {
            const float4 v = st4[f];
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
          }
        }
        xm[0] = m.x; xm[1] = m.y; xm[2] = m.z; xm[3] = m.w;
        if constexpr (kShfl) {
#pragma unroll
          for (int q = 0; q < 4; ++q) xm[q] = dec_ordered(__reduce_max_sync(cls_mask, enc_ordered(xm[q])));
        } else {
          // lanes l, l + P, l + 2P, ... (< tw) hold the same four joints: rotate through them (the stage is still live, so
          // no shared-memory scratch here)
          float own[4] = {xm[0], xm[1], xm[2], xm[3]};
          for (int k = P; k < tw; k += P) {
            int src = lane + k;
            if (src >= tw) src -= tw;
            if (!act) src = lane;
#pragma unroll
            for (int q = 0; q < 4; ++q) xm[q] = fmaxf(xm[q], __shfl_sync(0xffffffffu, own[q], src));
          }
          if constexpr (!kSamePixel) {
            // nj % 4 != 0: a joint also sits in OTHER slots of other lane classes; fetch those classes' maxima
            const float cls[4] = {xm[0], xm[1], xm[2], xm[3]};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
              for (int qp = 0; qp < 4; ++qp) {
                const int src = (int)((peer[q] >> (8 * qp)) & 0xffu);
                const float o = __shfl_sync(0xffffffffu, cls[qp], src == 0xff ? lane : src);
                if (src != 0xff) xm[q] = fmaxf(xm[q], o);
              }
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // an all -inf chunk contributes nothing: a finite reference keeps 2^(x*g2 - m2) at 0 instead of NaN
        const bool empty_joint = attempt == 1 && xm[q] == -CUDART_INF_F;
        m2[q] = (attempt == 0 || empty_joint) ? 0.0f : xm[q] * g2;
        mout[q] = empty_joint ? -CUDART_INF_F : m2[q];   // the merge skips partials whose reference is -inf
        thr[q] = kDlc ? dlc_candidate_threshold(xm[q]) : 0.0f;
        s0[q] = sr[q] = sc[q] = 0.0f; bsig[q] = -1.0f; bidx[q] = 0x7fffffff;
      }
      float tmax = -CUDART_INF_F;               // largest exponent seen (attempt 0 only)

      // ---- pass 2
      if (act) {
        float prow[kPos], pcol[kPos];
#pragma unroll
        for (int q = 0; q < kPos; ++q) {
          const int pix = c0 + (4 * lane + q) / nj;
          int r = (int)((float)pix * invW);
          if (r * W > pix) --r;
          if ((r + 1) * W <= pix) ++r;
          prow[q] = (float)r;
          pcol[q] = (float)(pix - r * W);
        }
        if constexpr (kSamePixel) {
          const f32x2 nm01 = pk2(-m2[0], -m2[1]), nm23 = pk2(-m2[2], -m2[3]);
          const f32x2 zero2 = pk2(0.0f, 0.0f);
          // sums over joint PAIRS: (s0_0, s0_1), (s0_2, s0_3), likewise the row- and column-weighted ones
          f32x2 s0a = zero2, s0b = zero2, sra = zero2, srb = zero2, sca = zero2, scb = zero2;
          // (row, col) of the lane's current float4 as SCALARS: the packed FMAs take them as broadcast operands, and the
          // wrap at the end of a map row is a select instead of predicated 64-bit register-pair updates
          float pr = prow[0], pc = pcol[0];
          float tm0 = -CUDART_INF_F, tm1 = -CUDART_INF_F;
          auto consume = [&](const float4& v) {
            float t0, t1, t2, t3;
            const f32x2 rr = pk2(pr, pr), cc2 = pk2(pc, pc);
            upk2(fma2(pk2(v.x, v.y), g2g2, nm01), t0, t1);
            upk2(fma2(pk2(v.z, v.w), g2g2, nm23), t2, t3);
            if (!kDlc) { tm0 = fmaxf(tm0, fmaxf(t0, t1)); tm1 = fmaxf(tm1, fmaxf(t2, t3)); }
            const f32x2 ea = pk2(ex2_approx(t0), ex2_approx(t1)), eb = pk2(ex2_approx(t2), ex2_approx(t3));
            s0a = add2(s0a, ea);
            s0b = add2(s0b, eb);
            sra = fma2(ea, rr, sra);
            srb = fma2(eb, rr, srb);
            sca = fma2(ea, cc2, sca);
            scb = fma2(eb, cc2, scb);
            if (kDlc) {
              if ((v.x >= thr[0]) | (v.y >= thr[1]) | (v.z >= thr[2]) | (v.w >= thr[3])) {
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
            pr += dPr;
            pc += dPc;
            const float wrap = pc >= Wf ? 1.0f : 0.0f;   // dPc < W: at most one wrap; all values are small exact integers
            pc = fmaf(wrap, -Wf, pc);
            pr += wrap;
          };
          int f = lane;
          for (; f + 3 * tw < n4; f += 4 * tw) {
            const float4 v0 = st4[f], v1 = st4[f + tw], v2 = st4[f + 2 * tw], v3 = st4[f + 3 * tw];
            consume(v0); consume(v1); consume(v2); consume(v3);
          }
          for (; f < n4; f += tw) consume(st4[f]);
          tmax = fmaxf(tm0, tm1);

          // blur border correction: pixels within `radius` of an edge lose the taps that fall outside, i.e. their
          // weights are (Ah*Aw, Rh*Aw, Ah*Rw) instead of the interior (1, row, col) added above.  Lane (slot, class)
          // takes every blanes-th border pixel and, of its nj/4 float4s, the one holding this lane's joints.
          {
            int ra = (int)((float)c0 * invW);
            if (ra * W > c0) --ra;
            if ((ra + 1) * W <= c0) ++ra;
            const int pl = c0 + npx - 1;
            int rb = (int)((float)pl * invW);
            if (rb * W > pl) --rb;
            if ((rb + 1) * W <= pl) ++rb;
            auto fix_px = [&](int r, int cpx) {
              const int p = r * W + cpx;
              if (p < c0 || p >= c0 + npx) return;
              const float4 v = st4[(p - c0) * P + bcls];
              float t0, t1, t2, t3, k0, kr, kc;
              upk2(fma2(pk2(v.x, v.y), g2g2, nm01), t0, t1);
              upk2(fma2(pk2(v.z, v.w), g2g2, nm23), t2, t3);
              const f32x2 ea = pk2(ex2_approx(t0), ex2_approx(t1)), eb = pk2(ex2_approx(t2), ex2_approx(t3));
              border_corr((float)r, (float)cpx, k0, kr, kc);
              const f32x2 k00 = pk2(k0, k0), krr = pk2(kr, kr), kcc = pk2(kc, kc);
              s0a = fma2(ea, k00, s0a);
              s0b = fma2(eb, k00, s0b);
              sra = fma2(ea, krr, sra);
              srb = fma2(eb, krr, srb);
              sca = fma2(ea, kcc, sca);
              scb = fma2(eb, kcc, scb);
            };
            if (R2 == 0) {
              // radius 0: the blur is the identity, nothing to correct
            } else if (!all_border) {
              // left / right columns of every row of the chunk: item k = slot + n * blanes -> (row k / R2, side k % R2)
              for (int rq = brr0, sx = bsx0; ra + rq <= rb;) {
                fix_px(ra + rq, sx < radius ? sx : W - R2 + sx);
                rq += bdrr; sx += bdsx;
                if (sx >= R2) { sx -= R2; rq += 1; }
              }
              // top / bottom rows: the span between the side columns
              if (ra < radius)
                for (int r = ra; r <= min(rb, radius - 1); ++r)
                  for (int cpx = radius + bslot; cpx < W - radius; cpx += blanes) fix_px(r, cpx);
              if (rb >= H - radius)
                for (int r = max(ra, H - radius); r <= rb; ++r)
                  for (int cpx = radius + bslot; cpx < W - radius; cpx += blanes) fix_px(r, cpx);
            } else {
              for (int r = ra; r <= rb; ++r)
                for (int cpx = bslot; cpx < W; cpx += blanes) fix_px(r, cpx);
            }
          }
          upk2(s0a, s0[0], s0[1]);
          upk2(s0b, s0[2], s0[3]);
          upk2(sra, sr[0], sr[1]);
          upk2(srb, sr[2], sr[3]);
          upk2(sca, sc[0], sc[1]);
          upk2(scb, sc[2], sc[3]);
        } else {
          auto consume = [&](const float4& v) {
            const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float pr = prow[q % kPos], pc = pcol[q % kPos];
              const float t = fmaf(xs[q], g2, -m2[q]);
              if (!kDlc) tmax = fmaxf(tmax, t);
              const float e = ex2_approx(t);
              float w0 = 1.0f, wr = pr, wc = pc;
              if (all_border | (pr < Rf) | (pr >= HmR) | (pc < Rf) | (pc >= WmR)) {
                float k0, kr, kc;
                border_corr(pr, pc, k0, kr, kc);
                w0 += k0; wr += kr; wc += kc;
              }
              s0[q] = fmaf(e, w0, s0[q]);
              sr[q] = fmaf(e, wr, sr[q]);
              sc[q] = fmaf(e, wc, sc[q]);
              if (kDlc) {
                if (xs[q] >= thr[q]) {
                  const float sg = dlc_sigmoid(xs[q]);
                  const int idx = (int)pr * W + (int)pc;
                  if (sg > bsig[q] || (sg == bsig[q] && idx < bidx[q])) { bsig[q] = sg; bidx[q] = idx; }
                }
              }
            }
#pragma unroll
            for (int q = 0; q < kPos; ++q) {
              pcol[q] += dPc; prow[q] += dPr;
              if (pcol[q] >= Wf) { pcol[q] -= Wf; prow[q] += 1.0f; }
            }
          };
          int f = lane;
          for (; f + 1 * tw < n4; f += 2 * tw) {
            const float4 v0 = st4[f], v1 = st4[f + tw];
            consume(v0); consume(v1);
          }
          for (; f < n4; f += tw) consume(st4[f]);
        }
      }
      if (attempt == 0) {
        // warp-uniform decision; NaN exponents fall through to the exact pass as well
        const int tenc = __reduce_max_sync(0xffffffffu, enc_ordered(tmax));
        const float tall = dec_ordered(tenc);
        if (tall >= -90.0f && tall <= 90.0f) break;
      }
    }
    // the stage is free as soon as every lane has read its float4s: refill it (the scratch path first borrows it for its
    // reduction, see below)
    __syncwarp();
    auto refill = [&]() {
      if (lane == 0 && i + kWStages < nloc) {
        fence_proxy_async_smem();   // order the warp's generic-proxy accesses before the async-proxy write
        int nb = b, nc = c + kWStages;
        while (nc >= cpf) { nc -= cpf; ++nb; }
        issue_load_at(i + kWStages, nb, nc);
      }
    };
    if constexpr (kSamePixel) refill();   // only the nj % 4 != 0 reduction borrows the drained stage

    // ---- reduce over the lanes that share a joint; one partial per (chunk, joint), fixed order -> deterministic
    SaPartial* out = part + ((size_t)b * cpf + c) * nj;
    if constexpr (kShfl) {
      const float v12[12] = {s0[0], s0[1], s0[2], s0[3], sr[0], sr[1], sr[2], sr[3], sc[0], sc[1], sc[2], sc[3]};
      float fin[12];
      int cnt, base;
      butterfly12<kP>(v12, lane, fin, cnt, base);
      if (kDlc) {
        for (int o = 16; o >= kP; o >>= 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float obs = __shfl_xor_sync(0xffffffffu, bsig[q], o);
            const int obi = __shfl_xor_sync(0xffffffffu, bidx[q], o);
            if (obs > bsig[q] || (obs == bsig[q] && obi < bidx[q])) { bsig[q] = obs; bidx[q] = obi; }
          }
        }
      }
      // value index vi = 4 * field + q  (field 0: s0, 1: sr, 2: sc) of joint 4 * (lane % kP) + q
      const int jbase = 4 * (lane % kP);
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        if (k < cnt) {
          const int vi = base + k;
          const int fld = vi >> 2, q = vi & 3;
          reinterpret_cast<float*>(out + jbase + q)[1 + fld] = fin[k];
        }
      }
      if (lane < kP) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float* o = reinterpret_cast<float*>(out + jq[q]);
          o[0] = mout[q];
          o[4] = bsig[q];
          o[5] = __int_as_float(bidx[q]);
        }
      }
    } else if constexpr (kSamePixel) {
      // nj % 4 == 0 but the lanes of a class {c, c + P, ...} are not an xor-closed set (nj = 12, 20, 24, ...): fold the class
      // in halves while its size is even, then rotate through what is left; lane c < P ends up with the class sums in a
      // fixed order (deterministic) and stores its four joints.
      float v12[12] = {s0[0], s0[1], s0[2], s0[3], sr[0], sr[1], sr[2], sr[3], sc[0], sc[1], sc[2], sc[3]};
      int span = tw / P;
      while ((span & 1) == 0) {
        span >>= 1;
        const int src = lane + span * P;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          const float o = __shfl_sync(0xffffffffu, v12[k], src < 32 ? src : lane);
          v12[k] += o;   // meaningful on the lanes below span * P, which read a valid partner
        }
        if (kDlc) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float obs = __shfl_sync(0xffffffffu, bsig[q], src < 32 ? src : lane);
            const int obi = __shfl_sync(0xffffffffu, bidx[q], src < 32 ? src : lane);
            if (obs > bsig[q] || (obs == bsig[q] && obi < bidx[q])) { bsig[q] = obs; bidx[q] = obi; }
          }
        }
      }
      {
        float own[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) own[k] = v12[k];
        const float obs0[4] = {bsig[0], bsig[1], bsig[2], bsig[3]};
        const int obi0[4] = {bidx[0], bidx[1], bidx[2], bidx[3]};
        for (int m = 1; m < span; ++m) {
          int src = lane + m * P;
          if (src >= span * P) src -= span * P;
          if (lane >= span * P) src = lane;
#pragma unroll
          for (int k = 0; k < 12; ++k) v12[k] += __shfl_sync(0xffffffffu, own[k], src);
          if (kDlc) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float obs = __shfl_sync(0xffffffffu, obs0[q], src);
              const int obi = __shfl_sync(0xffffffffu, obi0[q], src);
              if (obs > bsig[q] || (obs == bsig[q] && obi < bidx[q])) { bsig[q] = obs; bidx[q] = obi; }
            }
          }
        }
      }
      if (lane < P) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4* o4 = reinterpret_cast<float4*>(out + jq[q]);
          o4[0] = make_float4(mout[q], v12[q], v12[4 + q], v12[8 + q]);
          o4[1] = make_float4(bsig[q], __int_as_float(bidx[q]), 0.0f, 0.0f);
        }
      }
    } else {
      float* scr = stage + s * kWChunkFloats;   // the drained stage
      if (act) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float* r = scr + (4 * lane + q) * 5;
          r[0] = s0[q]; r[1] = sr[q]; r[2] = sc[q]; r[3] = bsig[q]; r[4] = __int_as_float(bidx[q]);
          if (4 * lane + q < nj) scr[640 + 4 * lane + q] = mout[q];   // slot e < nj holds joint e
        }
      }
      __syncwarp();
      for (int j = lane; j < nj; j += 32) {
        float a0 = 0.0f, ar = 0.0f, ac = 0.0f, bs = -1.0f;
        int bi = 0x7fffffff;
        for (int e = j; e < 4 * tw; e += nj) {
          const float* r = scr + e * 5;
          a0 += r[0]; ar += r[1]; ac += r[2];
          const int idx = __float_as_int(r[4]);
          if (r[3] > bs || (r[3] == bs && idx < bi)) { bs = r[3]; bi = idx; }
        }
        float4* o4 = reinterpret_cast<float4*>(out + j);
        o4[0] = make_float4(scr[640 + j], a0, ar, ac);
        o4[1] = make_float4(bs, __int_as_float(bi), 0.0f, 0.0f);
      }
      __syncwarp();
      refill();
    }
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
    const float4* p4 = reinterpret_cast<const float4*>(part + ((size_t)b * splits + sp) * nj + j);
    const float4 p0 = __ldg(p4), p1 = __ldg(p4 + 1);
    Acc q;
    q.m = p0.x; q.s0 = p0.y; q.sr = p0.z; q.sc = p0.w; q.bsig = p1.x; q.bidx = __float_as_int(p1.y);
    if (sp < 32) a = q; else acc_merge(a, q);
  }
  {
    // one rescale per lane against the warp-wide reference, then plain sums (fixed xor tree -> deterministic)
    const float m_all = dec_ordered(__reduce_max_sync(0xffffffffu, enc_ordered(a.m)));
    const float f = (a.m == -CUDART_INF_F) ? 0.0f : exp2f(a.m - m_all);
    a.s0 *= f; a.sr *= f; a.sc *= f;
    a.m = m_all;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.s0 += __shfl_xor_sync(0xffffffffu, a.s0, o);
      a.sr += __shfl_xor_sync(0xffffffffu, a.sr, o);
      a.sc += __shfl_xor_sync(0xffffffffu, a.sc, o);
    }
    if (dlc_peak || dlc_pose) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float obs = __shfl_xor_sync(0xffffffffu, a.bsig, o);
        const int obi = __shfl_xor_sync(0xffffffffu, a.bidx, o);
        if (obs > a.bsig || (obs == a.bsig && obi < a.bidx)) { a.bsig = obs; a.bidx = obi; }
      }
    }
  }
  const float* fr = logits + (size_t)b * H * W * nj + j;
  // every lane holds the merged sums (xor trees): the <= 2 x 2 window of the estimate_pose read-out is fetched by lanes
  // 0..3 in parallel (one load latency instead of four dependent ones), lane 0 then scans it in numpy's order
  const float mur = a.sr / a.s0, muc = a.sc / a.s0;  // 0/0 -> NaN, as softmax_tensor / (sum + 1e-100) in fp32
  float win = 0.0f;
  int rlo = 0, rhi = 0, clo = 0, chi = 0;
  const bool want_win = (peak || lik) && mur == mur && muc == muc;
  if (want_win) {
    // numpy slice [floor : ceil+1] clipped to the array
    rlo = max((int)floorf(mur), 0); rhi = min((int)ceilf(mur) + 1, H);
    clo = max((int)floorf(muc), 0); chi = min((int)ceilf(muc) + 1, W);
    const int r = rlo + (lane >> 1), c = clo + (lane & 1);
    if (lane < 4 && r < rhi && c < chi) win = sigmoid_literal(fr[((size_t)r * W + c) * nj]);
  }
  const float w1 = __shfl_sync(0xffffffffu, win, 1), w2 = __shfl_sync(0xffffffffu, win, 2), w3 = __shfl_sync(0xffffffffu, win, 3);
  if (lane != 0) return;
  if (mu) { mu[2 * t] = mur; mu[2 * t + 1] = muc; }
  if (norm) { norm[2 * t] = a.m; norm[2 * t + 1] = a.s0; }

  if (peak || lik) {
    int pr = -1, pc = -1;
    float best = CUDART_NAN_F;
    if (want_win) {
      bool have = false, have_nan = false;
      for (int r = rlo; r < rhi && !have_nan; ++r)
        for (int c = clo; c < chi; ++c) {
          const int wi = (r - rlo) * 2 + (c - clo);
          const float s = wi == 0 ? win : (wi == 1 ? w1 : (wi == 2 ? w2 : w3));
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
  const int tab_off = ((kPotFrames + 1) * row + kPotFrames * (nj + 1) + 1) & ~1;   // 8-byte aligned
  int2* sm_edge = reinterpret_cast<int2*>(psm + tab_off);   // [nl] (2a, 2b)
  float2* sm_w = reinterpret_cast<float2*>(sm_edge + nl);                                                 // [nl] (ws, ws_max)
  for (int l = threadIdx.x; l < nl; l += blockDim.x) {
    sm_edge[l] = make_int2(2 * __ldg(edges + 2 * l), 2 * __ldg(edges + 2 * l + 1));
    sm_w[l] = ws ? make_float2(__ldg(ws + l), __ldg(ws_max + l)) : make_float2(0.0f, 0.0f);
  }
  const int t0 = blockIdx.x * kPotFrames;
  const int nf = min(kPotFrames, T - t0);
  const int nload = (t0 + nf < T) ? nf + 1 : nf;    // + first frame of the next CTA's range
  const float* src = mu + (size_t)t0 * nj * 2;
  {
    // flat coalesced copy, 128 bits per load and every load of a thread issued before its first use (the staging is
    // latency-bound otherwise); (frame, coordinate) of each element is maintained incrementally, no division in the loop
    const int rl = 2 * nj;
    const int total = nload * rl;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int n4 = vec_ok ? total / 4 : 0;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    constexpr int kU = 9;   // covers (128 + 1) frames x 32 floats with 128 threads in one batch
    // (frame, coordinate) of a thread's float4s advance by a fixed (df, dc) per step: one division per thread
    const int step = 4 * (int)blockDim.x;
    const int df = step / rl, dc = step - df * rl;
    int f = (4 * (int)threadIdx.x) / rl, c = 4 * (int)threadIdx.x - f * rl;
    const bool rows_whole = (rl & 3) == 0;   // a float4 never straddles two frames
    for (int base = 0; base < n4; base += kU * (int)blockDim.x) {
      float4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i4 = base + u * (int)blockDim.x + (int)threadIdx.x;
        if (i4 < n4) v[u] = __ldcs(src4 + i4);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i4 = base + u * (int)blockDim.x + (int)threadIdx.x;
        if (i4 < n4) {
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          float* d = sm_mu + f * row + c;
          if (rows_whole) {
#pragma unroll
            for (int k = 0; k < 4; ++k) d[k] = e[k] * stride + 0.5f * stride;
          } else {
            int ff = f, cc = c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              sm_mu[ff * row + cc] = e[k] * stride + 0.5f * stride;
              if (++cc == rl) { cc = 0; ++ff; }
            }
          }
        }
        f += df; c += dc;
        if (c >= rl) { c -= rl; ++f; }
      }
    }
    for (int i = 4 * n4 + (int)threadIdx.x; i < total; i += (int)blockDim.x) {
      const int f = i / rl;
      sm_mu[f * row + (i - f * rl)] = src[i] * stride + 0.5f * stride;
    }
  }
  const bool have_next_global = (t0 + nf < T);
  if (!have_next_global && halo_next != nullptr)
    for (int i = threadIdx.x; i < 2 * nj; i += blockDim.x) sm_mu[nf * row + i] = halo_next[i] * stride + 0.5f * stride;
  __syncthreads();
  const int lt = threadIdx.x;
  const int t = t0 + lt;
  const bool active = lt < nf;
  const bool has_next = active && (lt + 1 < nf || have_next_global || halo_next != nullptr);
  if (active) {
    const float* m = sm_mu + lt * row;
    // sm_mu holds S-space coordinates mu * stride + stride / 2 (the reference's order of operations)
    float es = 0.0f;
    for (int l = 0; l < nl; ++l) {
      const int2 ab = sm_edge[l];
      const float dr = m[ab.x] - m[ab.y];
      const float dc = m[ab.x + 1] - m[ab.y + 1];
      const float d = sqrtf(dr * dr + dc * dc);
      if (skel) skel[(size_t)l * T + t] = d;
      if (ws) {
        const float2 w = sm_w[l];
        es += w.x * (fmaxf(d - w.y, 0.0f) + w.y);
      }
    }
    if (e_skel) e_skel[t] = es;
    float et = 0.0f;
    if (has_next) {
      const float* mn = m + row;
      float* trow = sm_t + lt * (nj + 1);
      for (int j = 0; j < nj; ++j) {
        const float dr = m[2 * j] - mn[2 * j];
        const float dc = m[2 * j + 1] - mn[2 * j + 1];
        const float d = sqrtf(dr * dr + dc * dc);
        trow[j] = d;
        const float dth = fmaxf(d - wt_max, 0.0f) + wt_max;
        et += dth * dth;
      }
    }
    if (e_temp) e_temp[t] = et;
  }
  __syncthreads();
  if (temporal) {
    // rows [t0, t0 + nvalid) of temporal are contiguous in global memory: 128-bit coalesced stores
    const int nvalid = (have_next_global || halo_next != nullptr) ? nf : nf - 1;
    float* dst = temporal + (size_t)t0 * nj;
    const int total = nvalid * nj;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    const int n4 = vec_ok ? total / 4 : 0;
    const int step = 4 * (int)blockDim.x;
    const int df = step / nj, dc = step - df * nj;
    int f = (4 * (int)threadIdx.x) / nj, c = 4 * (int)threadIdx.x - f * nj;
    const bool rows_whole = (nj & 3) == 0;
    for (int i4 = threadIdx.x; i4 < n4; i4 += blockDim.x) {
      float e[4];
      const float* srow = sm_t + f * (nj + 1) + c;
      if (rows_whole) {
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = srow[k];
      } else {
        int ff = f, cc = c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          e[k] = sm_t[ff * (nj + 1) + cc];
          if (++cc == nj) { cc = 0; ++ff; }
        }
      }
      __stcs(reinterpret_cast<float4*>(dst) + i4, make_float4(e[0], e[1], e[2], e[3]));
      f += df; c += dc;
      if (c >= nj) { c -= nj; ++f; }
    }
    for (int i = 4 * n4 + (int)threadIdx.x; i < total; i += (int)blockDim.x) {
      const int f = i / nj;
      dst[i] = sm_t[f * (nj + 1) + (i - f * nj)];
    }
  }
}

}  // namespace

// active lanes per warp: the largest tw <= 32 with 4*tw % nj == 0 (0: unsupported joint count)
int softargmax_tact(int nj) {
  int g = nj;
  for (int a = 4, b = nj; b;) { int r = a % b; a = b; b = r; g = a; }
  const int step = nj / g;  // tw must be a multiple of nj / gcd(nj, 4)
  return (32 / step) * step;
}

static int softargmax_chunk_px(int nj) { return (kWChunkFloats / nj) & ~3; }

// Chunks per frame: a function of the map shape only (never of the batch size) -> batch-invariant arithmetic.
int softargmax_splits(int H, int W, int nj) {
  if (nj < 1 || nj > kMaxJoints) return 1;
  const int cp = softargmax_chunk_px(nj);
  return (H * W + cp - 1) / cp;
}

template <bool kSamePixel, int kP, bool kDlc>
static cudaError_t launch_stream(const float* logits, int B, int H, int W, int nj, float gamma, int radius, float sigma,
                                 int chunk_px, int cpf, int tw, SaPartial* ws, int grid, cudaStream_t stream) {
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(softargmax_stream_kernel<kSamePixel, kP, kDlc>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  softargmax_stream_kernel<kSamePixel, kP, kDlc><<<grid, kWThreads, kWSmemBytes, stream>>>(
      logits, B, H, W, nj, gamma, radius, sigma, chunk_px, cpf, tw, ws);
  return cudaGetLastError();
}

cudaError_t launch_softargmax(const float* logits, const float* locref, int B, int H, int W, int nj, float gamma,
                              float gauss_len, float stride, float locref_stdev, SaPartial* workspace, int splits,
                              float* mu, int* peak, float* lik, int* dlc_peak, float* dlc_pose, float* norm,
                              cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const int radius = (int)gauss_len;
  if (nj < 1 || nj > kMaxJoints || (H & 1) || (W & 1) || radius > 4 || ((uintptr_t)logits & 15))
    return cudaErrorInvalidValue;
  const int tw = softargmax_tact(nj);
  if (tw <= 0) return cudaErrorInvalidValue;
  const int cpf = softargmax_splits(H, W, nj);
  if (cpf != splits) return cudaErrorInvalidValue;  // the caller sized the workspace with softargmax_splits()
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  const long long nchunks = (long long)B * cpf;
  if (nchunks > 0x7fffffffLL) return cudaErrorInvalidValue;
  // small batches: still give every math warp of a CTA a chunk before spreading over more SMs than needed
  long long want = (nchunks + kWWarps - 1) / kWWarps;
  const int grid = (int)(want < num_sms ? (want < 1 ? 1 : want) : num_sms);
  const int chunk_px = softargmax_chunk_px(nj);
  const bool dlc = dlc_peak != nullptr || dlc_pose != nullptr;
  const bool same = nj % 4 == 0;
  const bool shfl = same && (128 % nj == 0);
  cudaError_t e;
#define DGP_SA_LAUNCH(SP, KP, DL) \
  launch_stream<SP, KP, DL>(logits, B, H, W, nj, gamma, radius, gauss_len, chunk_px, cpf, tw, workspace, grid, stream)
#define DGP_SA_SHFL(KP) (dlc ? DGP_SA_LAUNCH(true, KP, true) : DGP_SA_LAUNCH(true, KP, false))
  if (shfl) {
    switch (nj / 4) {
      case 1: e = DGP_SA_SHFL(1); break;
      case 2: e = DGP_SA_SHFL(2); break;
      case 4: e = DGP_SA_SHFL(4); break;
      case 8: e = DGP_SA_SHFL(8); break;
      case 16: e = DGP_SA_SHFL(16); break;
      default: e = DGP_SA_SHFL(32); break;
    }
  } else if (same) {
    e = dlc ? DGP_SA_LAUNCH(true, 0, true) : DGP_SA_LAUNCH(true, 0, false);
  } else {
    e = dlc ? DGP_SA_LAUNCH(false, 0, true) : DGP_SA_LAUNCH(false, 0, false);
  }
#undef DGP_SA_SHFL
#undef DGP_SA_LAUNCH
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
  const size_t smem = ((size_t)(kPotFrames + 1) * (2 * nj + 1) + (size_t)kPotFrames * (nj + 1) + 2) * sizeof(float) +
                      (size_t)nl * (sizeof(int2) + sizeof(float2));
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
