// Internal declarations shared by capi.cu (inference entry points) and train.cu (training step): the handle, the
// per-shape execution plan and the host helpers that build GEMM / wgrad launch parameters.
#pragma once
#include "../../include/dgp_b200.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "conv_gemm_sm100.cuh"
#include "kernels.cuh"

namespace dgp {

struct TrainState;  // train.cu
typedef __nv_bfloat16 W16;  // opaque 16-bit storage element (bf16 or fp16 bits)

struct HostVar {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct ConvLayer {
  std::string scope;
  int R = 1, S = 1, Cin = 0, Cout = 0, stride = 1, dil = 1;
  bool relu = true;
  int K = 0;        // GEMM K (multiple of 64)
  int Npad = 0;     // rows of the weight matrix (multiple of block_n)
  int block_n = 0;
  __nv_bfloat16* w = nullptr;  // [Npad][K]  (points into the handle's 16-bit arena)
  float* scale = nullptr;      // [Npad] or nullptr
  float* shift = nullptr;
  size_t w_off = 0;            // offset of the fp32 master copy [Npad][K] in the parameter arena
  int ch_off = -1;             // offset of this layer's channels in the BN blocks (gamma / beta / mean / var / scale / shift)
};

struct UnitDesc {
  std::string scope;
  int depth, base, stride, rate;
  int shortcut = -1, conv1 = -1, conv2 = -1, conv3 = -1;  // indices into layers
};

enum StepKind { STEP_PREP = 0, STEP_GEMM = 1, STEP_POOL = 2, STEP_COL2IM = 3 };

struct Step {
  StepKind kind;
  ConvGemmParams gp;      // STEP_GEMM
  std::string end_point;  // name under which the output is kept in debug mode ("" = none)
  const void* out_ptr = nullptr;
  int oN = 0, oH = 0, oW = 0, oC = 0;  // output shape (bf16 NHWC) for debug dumps
  // pool
  const __nv_bfloat16* pin = nullptr;
  int pN = 0, pH = 0, pW = 0, pC = 0, pad_t = 0, pad_l = 0;
};

struct Plan {
  int B = 0, H = 0, W = 0;
  uint64_t last_use = 0;               // plan-cache LRU clock
  int H1 = 0, W1 = 0, Hs = 0, Ws = 0;  // conv1 output / s2d dims
  int hf = 0, wf = 0;                  // feature map (stride 16)
  std::vector<DevBuf> bufs;
  __nv_bfloat16* s2d = nullptr;
  float* contrib = nullptr;
  int contrib_ld = 0;
  std::vector<Step> steps;
  int head_step = 0;                   // index of the head GEMM in `steps` (everything before it is extract_features)
  // training plans keep every activation (no buffer rotation) and own the fp32 head outputs
  bool train = false;
  struct UnitBufs {
    void *x = nullptr, *sc = nullptr, *t1 = nullptr, *t2 = nullptr, *out = nullptr;  // unit input, projection shortcut, conv1/conv2/unit outputs
    int H = 0, W = 0, Ho = 0, Wo = 0, Cin = 0;
  };
  std::vector<UnitBufs> ub;
  void *c1 = nullptr, *pool = nullptr, *feat = nullptr;  // conv1 output, pool1 output, block4 output
  int Hp = 0, Wp = 0, pool_pad_t = 0, pool_pad_l = 0;
  float *logits = nullptr, *locref = nullptr;
  std::shared_ptr<void> bwd;  // train.cu: backward step list
};

}  // namespace dgp

using namespace dgp;  // internal header: the handle's members are dgp:: types

struct dgp_handle {
  dgp::TrainState* train = nullptr;  // allocated by dgp_train_enable
  dgp_config cfg;
  int device = 0;
  int num_sms = 0;
  int fp16 = 0;  // storage precision of activations / weights: 0 = bf16, 1 = fp16
  char err[512] = "";
  std::map<std::string, HostVar> host_vars;
  bool finalized = false;
  std::vector<ConvLayer> layers;
  int conv1_layer = -1, head_layer = -1;
  std::vector<UnitDesc> units;
  float* head_bias = nullptr;  // points into the arena
  int ctot = 0;
  // Parameter arena (fp32 master, kernel layouts): [weights n_w | gamma n_ch | beta n_ch | head bias n_bias]
  float* master = nullptr;
  dgp::W16* w16 = nullptr;       // 16-bit copy of the weight part (tensor-core operands)
  float* bn_mean = nullptr;      // frozen moving statistics [n_ch]
  float* bn_var = nullptr;
  float* bn_ss = nullptr;        // [scale n_ch | shift n_ch]
  float* conv1_mask = nullptr;   // [64][256] 1 = real 7x7x3 tap of the space-to-depth conv1 matrix
  size_t n_w = 0, n_ch = 0, n_bias = 0, n_params = 0;
  std::vector<float> host_master, host_gamma, host_beta, host_mean, host_var, host_bias, host_conv1_mask;
  std::map<std::tuple<int, int, int>, std::unique_ptr<Plan>> plans, train_plans;  // inference key (-B,..) = unchunked debug plan
  uint64_t plan_clock = 0;
  bool debug_keep = false;
  struct Kept {
    void* p;
    int N, H, W, C;
  };
  std::map<std::string, Kept> kept;
  int64_t launches = 0;
  // per-kernel-family CUDA-event profiling (bench.py roofline numbers)
  bool profiling = false;
  struct ProfRec {
    int kind;
    int count = 1;   // launches bracketed by the event pair (> 1 in bracket mode)
    cudaEvent_t a, b;
  };
  bool profile_brackets = false;  // dgp_set_profiling(h, 2): one event pair around every RUN of same-kind launches
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  // softargmax workspace
  DevBuf flow_ws;   // dgp_learn_wt workspace
  SaPartial* sa_ws = nullptr;
  size_t sa_ws_bytes = 0;
  DevBuf loss_ws;
  // estimate_pose_host staging
  cudaStream_t stream = nullptr;       // compute stream of dgp_estimate_pose_host
  cudaStream_t copy_stream = nullptr;  // H2D of the next batch overlaps the current batch's kernels
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  DevBuf st_frames2[2], st_logits, st_mu, st_peak, st_lik;
  // dgp_estimate_pose_stream: pinned host ring the reader thread decodes into (stream.cu)
  void* ring_slot[3] = {nullptr, nullptr, nullptr};
  size_t ring_bytes = 0;
};

// loss forward (+ optional loss-side backward) shared by dgp_loss_forward / dgp_loss_backward / the training step
extern "C" int dgp_run_loss_impl(dgp_handle* h, const dgp_loss_cfg* cfg, const dgp_loss_batch* b, float* losses_dev,
                                   float* targets_all_dev, float* grad_pred_dev, float* grad_locref_dev, int visible_only,
                                   void* stream);

namespace dgp {

W16 cvt16(const dgp_handle* h, float f);
float cvt16_to_float(const dgp_handle* h, W16 v);
int fail(dgp_handle* h, int code, const char* fmt, ...);

#define CU_OK(h, expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(h, DGP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

int ceil_div(int a, int b);
void same_pad(int in, int k, int stride, int rate, int* beg, int* out);
int pick_block_n(int n);
int tmem_cols_for(int block_n);
void conv_geometry(int R, int S, int stride, int dil, int H, int W, int pad_mode, int* Pout, int* Qout, int* lower_h_,
                   int* lower_w_, int* upper_h_, int* upper_w_);
int setup_epilogue(dgp_handle* h, ConvGemmParams& g, const char* scope, int n_img);
int make_gemm_step(dgp_handle* h, const ConvLayer& L, const void* x, int N, int H, int W, int pad_mode, void* out,
                   bool out_f32, const __nv_bfloat16* residual, int res_sub, int res_H, int res_W, int block_n_override,
                   Step* st, int* Ho, int* Wo);
int make_wgrad_params(dgp_handle* h, const char* scope, int R, int S, int Cin, int Cout, int stride, int dil,
                      const void* x, int N, int H, int W, int pad_mode, const void* dy, WgradParams* wp);
int alloc_buf(dgp_handle* h, Plan* pl, size_t bytes, void** out);
int build_plan(dgp_handle* h, int B, int H, int W, bool train, Plan** out);
int run_forward_plan(dgp_handle* h, Plan* pl, const uint8_t* frames_dev, float* logits_dev, float* locref_dev,
                     cudaStream_t s, int first_step = 0, int last_step = -1);
int ensure(dgp_handle* h, DevBuf* b, size_t bytes);
int refresh_operands(dgp_handle* h, cudaStream_t s);
void train_destroy(dgp_handle* h);  // train.cu

inline cudaEvent_t prof_event(dgp_handle* h) {
  if (!h->ev_pool.empty()) {
    cudaEvent_t e = h->ev_pool.back();
    h->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

struct ProfScope {
  dgp_handle* h;
  cudaStream_t s;
  int idx = -1;
  ProfScope(dgp_handle* h_, int kind, cudaStream_t s_, int count = 1) : h(h_), s(s_) {
    if (!h->profiling) return;
    dgp_handle::ProfRec r;
    r.kind = kind;
    r.count = count;
    r.a = prof_event(h);
    r.b = prof_event(h);
    cudaEventRecord(r.a, s);
    idx = (int)h->prof.size();
    h->prof.push_back(r);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof[idx].b, s);
  }
};

}  // namespace dgp
