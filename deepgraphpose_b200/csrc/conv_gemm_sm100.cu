// Implicit-GEMM convolution on tcgen05 / TMEM / TMA (sm_100a).  See conv_gemm_sm100.cuh.
//
// CTA = 576 threads, persistent over (m_blk, n_blk) tiles:
//   warp 0       TMA producer   (one lane): A tile 128 x 64 (tiled or im2col) + B tile block_n x 64 per stage
//   warp 1       MMA issuer     (one lane): 4 x tcgen05.mma (K=16) per stage into a double-buffered TMEM accumulator
//   warps 2..17  epilogue: tcgen05.ld -> fp32 BN scale/shift (+residual)(+ReLU) -> 16 bit (four warps per TMEM quadrant)
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty pair (MMA <-> epilogue), and per epilogue warp a
// ring of 2 KB staging buffers (32 rows x 32 channels, SWIZZLE_64B): the residual tile is TMA-loaded into a buffer
// ahead of time, the warp adds it to the accumulator in place and the buffer is TMA-stored to the output, so that
// every byte of residual / output traffic is moved by the copy engine, not by the LSU, and no epilogue warp ever
// waits for another one (round 1 ran 8 warps in pairs around named barriers: latency-bound at 35 % issue activity).
#include "conv_gemm_sm100.cuh"

#include <stdio.h>

#include <type_traits>

#include "ptx_sm100.cuh"

namespace dgp {

namespace {

constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;  // TMA warp + MMA warp + 16 epilogue warps
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kMaxEpiBufs = 4;
constexpr int kEpiUnitBytes = 32 * kEpiUnitCols * 2;  // one warp's 32 rows x 32 columns

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// fp16 storage mode (same tcgen05 kind::f16 path, 11-bit mantissa): saturate instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_fp16(float a, float b) {
  __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.0f), 65504.0f), fminf(fmaxf(b, -65504.0f), 65504.0f));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float* f, int fp16) {
  if (fp16) return make_uint4(pack_fp16(f[0], f[1]), pack_fp16(f[2], f[3]), pack_fp16(f[4], f[5]), pack_fp16(f[6], f[7]));
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
// One epilogue unit = this warp's 32 accumulator rows x 32 columns: BN scale/shift (fp32, packed FFMA2) (+ residual)
// (+ ReLU, folded into the float -> 16-bit conversion) written in place into the warp's 2 KB staging buffer
// (32 rows x 64 B, SWIZZLE_64B: 16-byte unit u of row r lives at u ^ ((r >> 1) & 3), conflict-free for LDS/STS.128).
template <bool kFp16, bool kRelu, bool kRes>
__device__ __forceinline__ void epi_unit_math(const uint32_t (&v)[2][16], uint8_t* my_row, int lane, const float* ss) {
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    uint4* s0 = reinterpret_cast<uint4*>(my_row + (((2 * s2) ^ sw) << 4));
    uint4* s1 = reinterpret_cast<uint4*>(my_row + (((2 * s2 + 1) ^ sw) << 4));
    const float* scp = ss + s2 * 16;
    const float* shp = ss + 32 + s2 * 16;
    uint32_t rr[8];
    if (kRes) {
      const uint4 r0 = *s0;
      const uint4 r1 = *s1;
      rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w; rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
    }
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float4 sc = *reinterpret_cast<const float4*>(scp + 2 * i);
      const float4 sh = *reinterpret_cast<const float4*>(shp + 2 * i);
      ptx::f32x2 q0 = ptx::fma2(ptx::pk2(__uint_as_float(v[s2][2 * i]), __uint_as_float(v[s2][2 * i + 1])),
                                ptx::pk2(sc.x, sc.y), ptx::pk2(sh.x, sh.y));
      ptx::f32x2 q1 = ptx::fma2(ptx::pk2(__uint_as_float(v[s2][2 * i + 2]), __uint_as_float(v[s2][2 * i + 3])),
                                ptx::pk2(sc.z, sc.w), ptx::pk2(sh.z, sh.w));
      if (kRes) {
        if (kFp16) {
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&rr[i]));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&rr[i + 1]));
          q0 = ptx::add2(q0, ptx::pk2(a.x, a.y));
          q1 = ptx::add2(q1, ptx::pk2(b.x, b.y));
        } else {
          q0 = ptx::add2(q0, ptx::pk2(bf16lo(rr[i]), bf16hi(rr[i])));
          q1 = ptx::add2(q1, ptx::pk2(bf16lo(rr[i + 1]), bf16hi(rr[i + 1])));
        }
      }
      float a0, a1, b0, b1;
      ptx::upk2(q0, a0, a1);
      ptx::upk2(q1, b0, b1);
      o[i] = ptx::cvt_pack16<kFp16, kRelu>(a0, a1);
      o[i + 1] = ptx::cvt_pack16<kFp16, kRelu>(b0, b1);
    }
    *s0 = make_uint4(o[0], o[1], o[2], o[3]);
    *s1 = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

// kFp16 selects the 16-bit storage type at compile time (bf16 / fp16): no dtype branches in the epilogue.
template <bool kFp16>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.num_stages;
  const int msub = p.msub;                     // 128-row sub-tiles per CTA tile (BLOCK_M = 128 * msub)
  const int a_bytes = msub * kABytes;
  const int b_bytes = p.block_n * kBlockK * 2;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * a_bytes;
  uint8_t* smem_epi = smem_b + stages * b_bytes;  // [16 warps][epi_bufs][2 KiB], 1024 B aligned
  const int epi_bufs = p.epi_mode == 1 ? p.epi_bufs : 0;
  float* smem_ss = reinterpret_cast<float*>(smem_epi + kEpiWarps * epi_bufs * kEpiUnitBytes);  // [16 warps][2 sets][scale 32 | shift 32]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_ss + kEpiWarps * 128);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_bar = tmem_empty_bar + 2;  // [16 warps][kMaxEpiBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kEpiWarps * kMaxEpiBufs);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmap_a);
    ptx::prefetch_tmap(&p.tmap_b);
    if (p.epi_mode == 1) {
      ptx::prefetch_tmap(&p.tmap_out);
      if (p.residual != nullptr) ptx::prefetch_tmap(&p.tmap_res);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kEpiWarps);
    }
    for (int i = 0; i < kEpiWarps * kMaxEpiBufs; ++i) ptx::mbar_init(&res_bar[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (smem carve-up, barrier init, TMEM alloc, descriptor prefetch)
  // overlapped the previous layer's tail; from here on we touch its output, so wait for it to complete and flush.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp loops, one lane issues)
    // Kernel parameters used inside the k loop are copied to registers first: every c[0x0][..] re-read is a
    // ~10-cycle uniform load on the single-warp critical path.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = (uint32_t)(a_bytes + b_bytes);
    const int PQ = p.P * p.Q, Q = p.Q, nnb = p.num_n_blocks, nkb = p.num_k_blocks, block_n = p.block_n;
    const int a_mode = p.a_mode, cblocks = p.cblocks, S = p.S, dil = p.dil, cstride = p.conv_stride;
    const int lower_w = p.lower_w, lower_h = p.lower_h, block_m = kBlockM * msub;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / nnb;
      const int n_blk = tile - m_blk * nnb;
      const int m0 = m_blk * block_m;
      const int n0 = n_blk * block_n;
      int img = 0, cw = 0, ch = 0;
      if (a_mode == 1) {
        img = m0 / PQ;
        const int rem = m0 - img * PQ;
        const int pp = rem / Q;
        const int qq = rem - pp * Q;
        cw = qq * cstride + lower_w;
        ch = pp * cstride + lower_h;
      }
      int cb = 0, off_w = 0, off_h = 0, tap_s = 0;  // incremental (channel block, tap) counters: no divides in the loop
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          if (a_mode == 0) {
            ptx::tma_load_2d(smem_a + stage * a_bytes, &p.tmap_a, &full_bar[stage], kb * kBlockK, m0);
          } else {
            ptx::tma_load_im2col_4d(smem_a + stage * a_bytes, &p.tmap_a, &full_bar[stage], cb * kBlockK, cw, ch, img,
                                    (uint16_t)off_w, (uint16_t)off_h);
          }
          ptx::tma_load_2d(smem_b + stage * b_bytes, &p.tmap_b, &full_bar[stage], kb * kBlockK, n0);
        }
        __syncwarp();
        if (++cb == cblocks) {
          cb = 0;
          off_w += dil;
          if (++tap_s == S) {
            tap_s = 0;
            off_w = 0;
            off_h += dil;
          }
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    const int block_n = p.block_n, nkb = p.num_k_blocks;
    const uint32_t idesc = ptx::make_idesc_bf16_f32(kBlockM, block_n, kFp16 ? 1 : 0);
    const uint64_t adesc0 = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a));
    const uint64_t bdesc0 = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b));
    const uint32_t a_step = (uint32_t)a_bytes >> 4, b_step = (uint32_t)b_bytes >> 4;
    const uint32_t acc_cols = (uint32_t)(msub * block_n);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // The two tile heights get separate, branch-free issue loops (a predicate between UTCHMMAs costs issue slots).
    auto run = [&](auto msub_tag) {
      constexpr int kMsub = decltype(msub_tag)::value;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * acc_cols;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t adesc = adesc0 + (uint64_t)(stage * a_step);
            const uint64_t bdesc = bdesc0 + (uint64_t)(stage * b_step);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
              ptx::umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              if (kMsub == 2)  // rows 128..255 of a 256-row tile: A 16 KiB further, D block_n columns further
                ptx::umma_bf16(tmem_d + (uint32_t)block_n, adesc + (uint64_t)((kABytes >> 4) + k * 2),
                               bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            ptx::umma_commit(&empty_bar[stage]);
            if (kb == nkb - 1) ptx::umma_commit(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    };
    if (msub == 2) run(std::integral_constant<int, 2>{});
    else run(std::integral_constant<int, 1>{});
  } else if (p.epi_mode == 0) {
    // -------------------------------------------------------------- epilogue, direct stores (fp32 head GEMM)
    // 16 epilogue warps: four per TMEM lane quadrant, taking every fourth 16-column group.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile - m_blk * p.num_n_blocks;
      const int row = m_blk * kBlockM + quad * 32 + lane;
      const int n0 = n_blk * p.block_n;
      const bool row_ok = row < p.M;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.block_n);
      for (int c0 = cg * 16; c0 < p.block_n; c0 += 64) {
        uint32_t v[16];
        ptx::tmem_ld_x16(taddr + (uint32_t)c0, v);
        ptx::tmem_ld_wait();
        if (row_ok) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          const int n = n0 + c0;
          if (p.scale != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] *= __ldg(p.scale + n + i);
          }
          if (p.shift != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += __ldg(p.shift + n + i);
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
          if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * p.ldc + n);
#pragma unroll
            for (int i = 0; i < 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldc + n);
            op[0] = pack8(f, kFp16);
            op[1] = pack8(f + 8, kFp16);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // -------------------------------------------------------------- epilogue, TMA-staged 16-bit (+ TMA residual)
    // 16 INDEPENDENT epilogue warps, four per TMEM lane quadrant.  A tile is cut into units of 32 rows x 32 columns
    // (one warp's TMEM rows x two tcgen05.ld.x16); the warp with column-group index cg takes units cg, cg+4, ... of
    // the (sub-tile, column group) sequence.  Each warp owns a small ring of 2 KB staging buffers (32 rows x 64 B,
    // SWIZZLE_64B), TMA-loads the residual tile into a buffer ahead of time, combines in place and TMA-stores the
    // buffer: no barrier between warps anywhere in the epilogue, 16 units in flight per SM.
    const int quad = warp & 3;
    const int ew = warp - 2;
    const int cg = ew >> 2;
    const int nb = p.epi_bufs;
    uint8_t* ebuf = smem_epi + (size_t)ew * nb * kEpiUnitBytes;
    uint64_t* rbar = res_bar + ew * kMaxEpiBufs;
    const bool has_res = p.residual != nullptr;
    const int block_n = p.block_n, nnb = p.num_n_blocks;
    const int ncg = block_n >> 5;                                   // 32-column groups per tile row: 2, 4 or 8
    const int ncg_shift = ncg == 8 ? 3 : (ncg == 4 ? 2 : 1);
    const int units = msub * ncg;
    const int nsets = ncg == 8 ? 2 : 1;                             // distinct column groups this warp ever touches
    const int PQ = p.P * p.Q;
    float* ss = smem_ss + ew * 128;
    int ss_n0 = -1;

    // Residual prefetch cursor: walks this warp's (tile, unit) sequence nb-1 units ahead of the consumer.  All lanes
    // keep the (uniform) cursor; lane 0 issues.
    int pf_tile = blockIdx.x, pf_u = cg, pf_buf = 0, pf_mblk = 0, pf_n0 = 0;
    auto pf_setup_tile = [&]() {
      pf_mblk = pf_tile / nnb;
      pf_n0 = (pf_tile - pf_mblk * nnb) * block_n;
    };
    auto issue_residual = [&]() {
      if (pf_tile >= num_tiles) return;
      const int sub = pf_u >> ncg_shift;
      const int row0 = (pf_mblk * msub + sub) * kBlockM + quad * 32;
      const int col0 = pf_n0 + ((pf_u & (ncg - 1)) << 5);
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(&rbar[pf_buf], (uint32_t)kEpiUnitBytes);
        if (p.res_sub == 1) {
          ptx::tma_load_2d(ebuf + pf_buf * kEpiUnitBytes, &p.tmap_res, &rbar[pf_buf], col0, row0);
        } else {
          const int img = row0 / PQ;
          const int rem = row0 - img * PQ;
          const int pp = rem / p.Q;
          ptx::tma_load_im2col_4d(ebuf + pf_buf * kEpiUnitBytes, &p.tmap_res, &rbar[pf_buf], col0,
                                  (rem - pp * p.Q) * p.res_sub, pp * p.res_sub, img, 0, 0);
        }
      }
      if (++pf_buf == nb) pf_buf = 0;
      pf_u += 4;
      if (pf_u >= units) {
        pf_u = cg;
        pf_tile += gridDim.x;
        if (pf_tile < num_tiles) pf_setup_tile();
      }
    };

    const bool active = cg < units;  // 128-row tiles of a 64-wide layer have only two units
    if (has_res && active) {
      if (pf_tile < num_tiles) pf_setup_tile();
      for (int i = 0; i < nb - 1; ++i) issue_residual();
      __syncwarp();
    }
    int buf = 0;
    uint32_t buf_phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / nnb;
      const int n0 = (tile - m_blk * nnb) * block_n;
      if (n0 != ss_n0 && active) {
        if (lane < 16 * nsets) {
          const int set = lane >> 4, is_shift = (lane >> 3) & 1, i = (lane & 7) * 4;
          const int col = n0 + (((cg & (ncg - 1)) + 4 * set) << 5) + i;
          const float* src = is_shift ? p.shift : p.scale;
          const float fill = is_shift ? 0.f : 1.f;
          const float4 val = src ? __ldg(reinterpret_cast<const float4*>(src + col)) : make_float4(fill, fill, fill, fill);
          *reinterpret_cast<float4*>(ss + set * 64 + is_shift * 32 + i) = val;
        }
        __syncwarp();
        ss_n0 = n0;
      }
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      for (int u = cg; u < units; u += 4) {
        const int sub = u >> ncg_shift;
        const int colg = u & (ncg - 1);
        const int row0 = (m_blk * msub + sub) * kBlockM + quad * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((acc * msub + sub) * block_n + (colg << 5));
        uint8_t* my_buf = ebuf + buf * kEpiUnitBytes;
        uint32_t v[2][16];
        ptx::tmem_ld_x16(taddr, v[0]);
        ptx::tmem_ld_x16(taddr + 16, v[1]);
        if (has_res) {
          // the buffer the next prefetch lands in was last used by the unit before this one: its store must have
          // finished reading shared memory (it was committed a whole unit ago)
          if (lane == 0) ptx::bulk_wait_group_read<0>();
          issue_residual();
          ptx::mbar_wait(&rbar[buf], buf_phase);
        } else {
          // the store that last used this buffer (nb units ago) must have finished reading it
          if (lane == 0) {
            if (nb >= 3) ptx::bulk_wait_group_read<2>();
            else if (nb == 2) ptx::bulk_wait_group_read<1>();
            else ptx::bulk_wait_group_read<0>();
          }
          __syncwarp();
        }
        ptx::tmem_ld_wait();
        const float* ssu = ss + (colg >> 2) * 64;
        uint8_t* my_row = my_buf + lane * 64;
        if (has_res) {
          if (p.relu) epi_unit_math<kFp16, true, true>(v, my_row, lane, ssu);
          else epi_unit_math<kFp16, false, true>(v, my_row, lane, ssu);
        } else {
          if (p.relu) epi_unit_math<kFp16, true, false>(v, my_row, lane, ssu);
          else epi_unit_math<kFp16, false, false>(v, my_row, lane, ssu);
        }
        ptx::fence_proxy_async_smem();  // my generic-proxy smem writes -> visible to the TMA store (async proxy)
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&p.tmap_out, my_buf, n0 + (colg << 5), row0);
          ptx::bulk_commit_group();
        }
        if (++buf == nb) {
          buf = 0;
          buf_phase ^= 1;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) ptx::bulk_wait_group_read<0>();
    __syncwarp();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
char g_err[256];

}  // namespace

const char* tma_init() {
  if (g_encode_tiled && g_encode_im2col) return nullptr;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return "cuTensorMapEncodeTiled not available";
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return "cuTensorMapEncodeIm2col not available";
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  return nullptr;
}

namespace { int g_tmap_fp16 = 0; }
void tmap_set_fp16(int fp16) { g_tmap_fp16 = fp16; }

const char* make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t k, uint64_t row_stride_bytes,
                         uint32_t box_rows, uint32_t box_cols) {
  if (const char* e = tma_init()) return e;
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  const CUtensorMapSwizzle swz = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;  // 128 B / 64 B rows
  if (box_cols != 64 && box_cols != 32) return "make_tmap_2d: box_cols must be 64 or 32";
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(out, g_tmap_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed: %d (rows=%llu k=%llu stride=%llu box_rows=%u)", (int)r,
             (unsigned long long)rows, (unsigned long long)k, (unsigned long long)row_stride_bytes, box_rows);
    return g_err;
  }
  return nullptr;
}

const char* make_tmap_im2col(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N,
                             uint64_t stride_w_bytes, uint64_t stride_h_bytes, uint64_t stride_n_bytes, int lower_w,
                             int lower_h, int upper_w, int upper_h, int conv_stride, uint64_t total_bytes,
                             uint32_t pixels_per_column, uint32_t channels_per_pixel) {
  if (const char* e = tma_init()) return e;
  if (channels_per_pixel != 64 && channels_per_pixel != 32) return "make_tmap_im2col: channels_per_pixel must be 64 or 32";
  const CUtensorMapSwizzle swz = channels_per_pixel == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {stride_w_bytes, stride_h_bytes, stride_n_bytes};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)conv_stride, (cuuint32_t)conv_stride, 1};
  CUresult r = g_encode_im2col(out, g_tmap_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, (cuuint32_t)channels_per_pixel, (cuuint32_t)pixels_per_column, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err),
             "cuTensorMapEncodeIm2col failed: %d (C=%llu W=%llu H=%llu N=%llu sw=%llu sh=%llu sn=%llu lo=(%d,%d) up=(%d,%d) s=%d)",
             (int)r, (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N,
             (unsigned long long)stride_w_bytes, (unsigned long long)stride_h_bytes, (unsigned long long)stride_n_bytes,
             lower_w, lower_h, upper_w, upper_h, conv_stride);
    return g_err;
  }
  // Same driver quirk CUTLASS works around (copy_traits_sm90_im2col.hpp): for tensors < 128 KiB, drivers <= 13.1
  // set a descriptor bit that breaks im2col loads.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && total_bytes < 131072) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  return nullptr;
}

size_t conv_gemm_smem_bytes(int block_n, int num_stages, int epi_bufs, int msub) {
  return 1024 + (size_t)num_stages * ((size_t)msub * kABytes + (size_t)block_n * kBlockK * 2) +
         (size_t)kEpiWarps * epi_bufs * kEpiUnitBytes + kEpiWarps * 128 * sizeof(float) +
         (2 * kMaxStages + 4 + kEpiWarps * kMaxEpiBufs) * 8 + 16;
}

int conv_gemm_pick_stages(int block_n, int epi_bufs, int msub) {
  const size_t budget = 227 * 1024;
  int s = kMaxStages;
  while (s > 2 && conv_gemm_smem_bytes(block_n, s, epi_bufs, msub) > budget) --s;
  return s;
}

cudaError_t launch_conv_gemm(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < num_sms ? tiles : num_sms;
  const size_t smem = conv_gemm_smem_bytes(p.block_n, p.num_stages, p.epi_mode == 1 ? p.epi_bufs : 0, p.msub);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (p.fp16) return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true>, p);
  return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<false>, p);
}

}  // namespace dgp
