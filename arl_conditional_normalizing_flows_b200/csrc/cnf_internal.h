// Internal (non-ABI) structures shared by the host planner, the launchers and the C-ABI layer.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "cnf.h"

#define CNF_LRELU_SLOPE 0.3f  // keras LeakyReLU() default (F:345)
#define CNF_LN_EPS 1e-3       // keras LayerNormalization default epsilon (F:358)

namespace cnf {

void set_error(const char* fmt, ...);

// "this launcher does not cover the shape": never a cudaError_t value (cudaErrorInvalidValue == 1)
constexpr int CNF_NOT_ELIGIBLE = -1;
constexpr int CNF_MAX_DEVICES = 64;

// Tuning / ablation switches.  Release builds use the built-in default and never read the environment; a
// -DCNF_DEBUG build (make EXTRA=-DCNF_DEBUG) reads CNF_<NAME> so that profiling experiments need no rebuild.
int knob_int(const char* name, int dflt);
float knob_float(const char* name, float dflt);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: one cache per launcher, indexed by
// the current device, guarded by a mutex (plan.cpp).  Returns a cudaError as int.
struct SmemAttrCache {
  size_t set[CNF_MAX_DEVICES] = {};
};
int ensure_dynamic_smem(const void* kernel, size_t bytes, SmemAttrCache& cache);
int device_sm_count(int* n_sm);   // SM count of the current device (cached per device)

struct ParamEntry {
  std::string name;
  int64_t offset;  // floats from the start of the net's block
  int ndim;
  int64_t shape[4];
  int role;  // 0 kernel, 1 bias, 2 gamma, 3 beta, 4 tanh scale
};

struct Branch {
  int dil, channels, groups, gin, gout;
  int out_off;            // channel offset inside the concat (F:590)
  int64_t w_off, b_off;   // packed [group][ky][kx][gin][gout] kernels, [channels] biases
};

struct ResBlockLayout {
  int64_t ln1_g, ln1_b, pw1_w, pw1_b, ln2_g, ln2_b, ln3_g, ln3_b, pw2_w, pw2_b;
  std::vector<Branch> br;
};

}  // namespace cnf

// One coupling layer: shapes (M:355-439, M:474-498, M:1087-1104) and the flat parameter layout of
// its two s/t networks (net A at 0, net b at net_stride; identical internal layout).
struct cnf_coupling {
  int H, W, D, mask, mask_c, R, card, nk, ks, ln;
  int h, w, c1, c2, cat;
  int paths = 0;  // CNF_PATH_* bits (cnf.h): kernel families this layer must NOT use (0 = fastest path for every stage)
  std::vector<int> dil;
  int64_t stem_w, stem_b, lnf_g, lnf_b, head_w, head_b, tanh_w;
  std::vector<cnf::ResBlockLayout> rb;
  int64_t net_stride;
  std::vector<cnf::ParamEntry> entries;
  int n_ln() const { return ln ? 3 * R + 1 : 0; }
  int hw() const { return h * w; }
};

struct cnf_plan {
  int H, W, D, x_d, ks, ln;
  double lambda_y;
  std::vector<int> sq, resnext, nk, card, scale, npf;
  std::vector<cnf_block_info> blocks;
  struct LayerRef { int kind, aux; };
  std::vector<LayerRef> layers;
  std::vector<cnf_coupling*> couplings;
  std::vector<int64_t> param_off;
  std::vector<int> level;
  int64_t param_count;
  ~cnf_plan();
};

namespace cnf {

// Strided view of an "active" tensor inside the original-layout [B,H,W,D] buffer: after L
// squeeze+factor steps the active tensor is rows h = 2^L-1 (mod 2^L), each row reshaped to
// (W/2^L, 2^L*D) (tests/test_oracle_masks.py::test_active_tensor_is_a_strided_view).
struct FlowView {
  float* base;
  long long sb, sy, sx;  // element strides: sample, active row, active pixel (channels contiguous)
  int H, W, D;           // active shape
};

inline FlowView make_view(float* buf, int H0, int W0, int D0, int level) {
  FlowView v;
  const int S = 1 << level;
  v.base = buf + (long long)(S - 1) * W0 * D0;
  v.sb = (long long)H0 * W0 * D0;
  v.sy = (long long)S * W0 * D0;
  v.sx = (long long)S * D0;
  v.H = H0 / S;
  v.W = W0 / S;
  v.D = D0 * S;
  return v;
}

enum HeadMode { HEAD_FWD = 0, HEAD_INV = 1, HEAD_EMIT = 2, HEAD_EMIT_TANH = 3 };  // 3: outA = tanh(raw) (training)
enum { MASK_DENSE = 4 };  // "already compressed" input for A_wrapper/b_wrapper (M:452-472)

struct CouplingWorkspace {
  float *X, *Y1, *Y2;
  double* stats;  // [n_ln][2][B][2]
};
// Training: every stage of the s/t nets keeps its own buffer so that the backward pass can re-read it.
struct CouplingSaved {
  std::vector<float*> X;    // R+1 residual-stream states   [2][B][hw][nk]
  std::vector<float*> Y1;   // R   1x1 outputs               [2][B][hw][nk]
  std::vector<float*> Y2;   // R   grouped-conv outputs      [2][B][hw][cat]
  double* stats;            // [n_ln + 1][2][B][2] forward LayerNorm sums
  float* TH;                // tanh(raw_A) [B][hw][c2]
  float* state;             // flow buffer BEFORE this layer [B][H0][W0][D0]
};
int64_t coupling_saved_bytes(const cnf_coupling* c, int64_t B);
// carves `mem` (coupling_saved_bytes) -- `state` is assigned by the caller
CouplingSaved carve_saved(const cnf_coupling* c, int64_t B, void* mem);
// scratch of the backward pass of one layer (gradient buffers, per-sample LN-backward sums)
int64_t coupling_bwd_scratch_bytes(const cnf_coupling* c, int64_t B);

// fused_kernels.cu: one launch per coupling layer, s/t-net activations resident in shared memory (inference only).
// Returns -1 when the layer does not fit that kernel (never a cudaError value), else a cudaError as int.
int launch_fused_coupling(const cnf_coupling* c, const float* params, FlowView in_view, int in_mask, FlowView out_view,
                          int B, int mode, double* logdet_acc, void* ws, void* stream, bool dry_run = false);

int64_t coupling_ws_bytes(const cnf_coupling* c, int64_t B);
CouplingWorkspace carve_ws(const cnf_coupling* c, int64_t B, void* ws);

// launchers (kernels.cu); all enqueue on `stream`, return cudaError as int (0 ok)
int run_coupling(const cnf_coupling* c, const float* params, FlowView in_view, int in_mask,
                 FlowView out_view, int B, int mode, double* logdet_acc, float* outA, float* outB,
                 void* ws, void* stream, const CouplingSaved* sv = nullptr, int extra_excluded_paths = 0);
// backward of one coupling layer (bwd_kernels.cu): G is the gradient w.r.t. the layer OUTPUT in the flow
// buffer layout (updated in place to the gradient w.r.t. the layer INPUT); grads has the layout of params.
int run_coupling_backward(const cnf_coupling* c, const float* params, float* grads, const CouplingSaved& sv,
                          FlowView g_view, FlowView s_view, int B, float inv_batch, void* scratch, void* stream);
int dgrad_pw(const float* params, long long net_stride, long long w_off, const float* dY, float* dA, int B, int hw,
             int K, int N, void* stream, int paths = 0);
int dgrad_gconv(const cnf_coupling* c, int r, const float* params, const float* dY, float* dA, int B, void* stream,
                unsigned* leftover);
int launch_loss_grad(const float* zy, const float* xy, float* G, int64_t n, int D, int x_d, float lambda_y,
                     float inv_batch, void* stream);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps,
                float gscale, void* stream);
int read_tc3_clocks(long long* out, int n);
int run_residual_block(const cnf_coupling* c, const float* params, int r, const float* Xin, float* Xout, int B, void* ws,
                       void* stream);
int run_pw_only(const cnf_coupling* c, const float* params, int B, int which, void* ws, void* stream);
int launch_copy(const float* src, float* dst, int64_t n, void* stream);
int launch_logdet_finalize(const double* acc, float* out, int B, void* stream, int with_mean = 0);
int launch_prior_loss(const float* zy, const float* xy, const float* logdet, int B, int64_t HW, int D,
                      int x_d, double lambda_y, float* ll_z, float* ll_y, float* loss4, void* stream);
int launch_coupling_law(const float* u, const float* s, const float* t, float* v, float* logdet, int B,
                        int H, int W, int D, int mask, int inverse, void* stream);
int launch_mask(const float* uv, float* out, int B, int H, int W, int D, int mask, int compress,
                void* stream);
int launch_decompress(const float* uvc, float* out, int B, int H, int W, int D, int mask, void* stream);
int launch_space_to_depth(const float* in, float* out, int B, int H, int W, int C, int inverse,
                          void* stream);
// data_kernels.cu: input / output side helpers (F:74-318, F:635-676)
int launch_down(const float* in, float* out, int B, int H, int W, int D, int L, void* stream);
int launch_up(const float* in, float* out, int B, int H, int W, int D, int L, void* stream);
int launch_sr_preprocess(const float* hires, float* out, int B, int H, int W, int D, int lx, int ly, int residual,
                         void* stream);
int launch_logit(const float* x, float* out, long long n, double a, int inverse, void* stream);
int launch_instance_noise(const float* x, float* out, long long n, double alpha, unsigned long long seed,
                          unsigned long long offset, void* stream);
int launch_toy(const float* u, const float* params, const int* mask_idx_host, int n_layers_c, int width,
               int num_layers, int direction, float* v, float* logdet, int B, void* stream);
int launch_toy_grad(const float* xy, const float* zy, const float* params, float* grads, const int* mask_idx_host,
                    int n_layers_c, int width, int num_layers, int x_d, double lambda_y, int B, void* stream);
int launch_toy_loss(const float* zy, const float* xy, const float* logdet, int B, int x_d, double lambda_y,
                    float* ll_z, float* ll_y, float* loss4, void* stream);

}  // namespace cnf
