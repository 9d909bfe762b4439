// C-ABI layer: DLPack validation, error classes, and the per-call orchestration of the flow.
// Reference call graph being replaced: cFlow.call (M:1723-1798), cFlow.log_loss (M:1800-1848),
// coupling_layer.forward_and_Jacobian / backward (M:1258-1394).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#include "cnf_internal.h"

using namespace cnf;

namespace {

struct Ten {
  float* p = nullptr;
  int ndim = 0;
  int64_t shape[6] = {0, 0, 0, 0, 0, 0};
  int64_t numel = 0;
  int64_t bytes = 0;
};

int fail(int code, const char* fmt, ...) {
  char buf[400];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return code;
}

// Borrow a DLPack tensor: fp32 (or raw bytes when `bytes_ok`), current CUDA device, compact, 16B aligned.
int borrow(const DLManagedTensor* m, const char* name, int ndim, Ten* t, bool bytes_ok = false) {
  if (!m) return fail(CNF_ERR_ARG, "%s: null tensor", name);
  const DLTensor& d = m->dl_tensor;
  if (d.device.device_type != kDLCUDA)
    return fail(CNF_ERR_DEVICE, "%s: expected a CUDA tensor (DLPack device_type %d)", name, (int)d.device.device_type);
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(CNF_ERR_CUDA, "%s: cudaGetDevice failed", name);
  if (d.device.device_id != dev)
    return fail(CNF_ERR_DEVICE, "%s: tensor lives on cuda:%d but the current device is cuda:%d", name,
                d.device.device_id, dev);
  const bool f32 = d.dtype.code == kDLFloat && d.dtype.bits == 32 && d.dtype.lanes == 1;
  if (!f32 && !bytes_ok) return fail(CNF_ERR_DTYPE, "%s: expected float32 (got code %d, %d bits)", name, (int)d.dtype.code, (int)d.dtype.bits);
  if (ndim >= 0 && d.ndim != ndim) return fail(CNF_ERR_SHAPE, "%s: expected %d dimensions, got %d", name, ndim, d.ndim);
  if (d.ndim > 6) return fail(CNF_ERR_SHAPE, "%s: too many dimensions", name);
  t->ndim = d.ndim;
  t->numel = 1;
  for (int i = 0; i < d.ndim; ++i) {
    t->shape[i] = d.shape[i];
    t->numel *= d.shape[i];
  }
  if (d.strides && t->numel > 0) {
    int64_t expect = 1;
    for (int i = d.ndim - 1; i >= 0; --i) {
      if (d.shape[i] != 1 && d.strides[i] != expect)
        return fail(CNF_ERR_LAYOUT, "%s: tensor must be compact row-major (NHWC contiguous)", name);
      expect *= d.shape[i];
    }
  }
  char* p = (char*)d.data + d.byte_offset;
  if (t->numel && ((uintptr_t)p & 15)) return fail(CNF_ERR_LAYOUT, "%s: data pointer must be 16-byte aligned", name);
  t->p = (float*)p;
  t->bytes = t->numel * ((d.dtype.bits * d.dtype.lanes + 7) / 8);
  return CNF_OK;
}

int cuda_rc(int e, const char* what) {
  if (e == 0) return CNF_OK;
  if (e == (int)cudaErrorInvalidConfiguration)
    return fail(CNF_ERR_UNSUPPORTED, "%s: shape not supported by the built kernels (%s)", what, cudaGetErrorString((cudaError_t)e));
  return fail(CNF_ERR_CUDA, "%s: CUDA error %d (%s)", what, e, cudaGetErrorString((cudaError_t)e));
}

#define TRY(x)              \
  do {                      \
    int rc_ = (x);          \
    if (rc_ != CNF_OK) return rc_; \
  } while (0)

int check_flow_tensor(const cnf_plan* p, const Ten& t, const char* name) {
  if (t.shape[1] != p->H || t.shape[2] != p->W || t.shape[3] != p->D)
    return fail(CNF_ERR_SHAPE, "%s: expected shape [B,%d,%d,%d], got [%lld,%lld,%lld,%lld]", name, p->H, p->W, p->D,
                (long long)t.shape[0], (long long)t.shape[1], (long long)t.shape[2], (long long)t.shape[3]);
  return CNF_OK;
}

int64_t max_coupling_ws(const cnf_plan* p, int64_t B) {
  int64_t m = 0;
  for (auto* c : p->couplings) {
    const int64_t b = coupling_ws_bytes(c, B);
    if (b > m) m = b;
  }
  return m;
}

// direction +1: all couplings in order, in place on `buf`; -1: reversed with the inverse law.
int run_flow(const cnf_plan* p, const float* params, float* buf, int B, int direction, double* ldacc, void* ws,
             void* stream) {
  // whole batch per layer: every op is per-sample, and cutting the batch into L2-sized chunks under-fills the launches
  // (measured in round 1)
  const int n = (int)p->couplings.size();
  for (int s = 0; s < n; ++s) {
    const int li = direction == 1 ? s : n - 1 - s;
    const cnf_coupling* c = p->couplings[li];
    FlowView v = make_view(buf, p->H, p->W, p->D, p->level[li]);
    const int e = run_coupling(c, params + p->param_off[li], v, c->mask, v, B, direction == 1 ? HEAD_FWD : HEAD_INV, ldacc,
                               nullptr, nullptr, ws, stream);
    if (e) return cuda_rc(e, "coupling layer");
  }
  return CNF_OK;
}

}  // namespace

extern "C" {

int cnf_flow_forward(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params, DLManagedTensor* zy,
                     DLManagedTensor* logdet, DLManagedTensor* workspace, void* stream) {
  if (!p) return fail(CNF_ERR_ARG, "null plan");
  Ten X, P, Z, L, W;
  TRY(borrow(xy, "xy", 4, &X));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(zy, "zy", 4, &Z));
  TRY(borrow(logdet, "logdet", 1, &L));
  TRY(borrow(workspace, "workspace", 1, &W, true));
  TRY(check_flow_tensor(p, X, "xy"));
  TRY(check_flow_tensor(p, Z, "zy"));
  const int64_t B = X.shape[0];
  const bool with_mean = L.shape[0] == B + 1;   // [B + 1]: the last element receives the batch mean (the reference's scalar)
  if (Z.shape[0] != B || (L.shape[0] != B && !with_mean))
    return fail(CNF_ERR_SHAPE, "zy/logdet batch size differs from xy (%lld; logdet may be [B] or [B + 1])", (long long)B);
  if (P.numel < p->param_count) return fail(CNF_ERR_SHAPE, "params: need %lld floats, got %lld", (long long)p->param_count, (long long)P.numel);
  if (W.bytes < cnf_plan_workspace_bytes(p, B)) return fail(CNF_ERR_WORKSPACE, "workspace: need %lld bytes, got %lld", (long long)cnf_plan_workspace_bytes(p, B), (long long)W.bytes);
  if (B == 0) return CNF_OK;
  if (X.p != Z.p) TRY(cuda_rc(launch_copy(X.p, Z.p, X.numel, stream), "copy"));
  double* ldacc = (double*)((char*)W.p + max_coupling_ws(p, B) + 256);
  TRY(cuda_rc((int)cudaMemsetAsync(ldacc, 0, sizeof(double) * B, (cudaStream_t)stream), "memset"));
  TRY(run_flow(p, P.p, Z.p, (int)B, 1, ldacc, W.p, stream));
  return cuda_rc(launch_logdet_finalize(ldacc, L.p, (int)B, stream, with_mean ? 1 : 0), "logdet");
}

int cnf_flow_inverse(const cnf_plan* p, const DLManagedTensor* zy, const DLManagedTensor* params, DLManagedTensor* xy,
                     DLManagedTensor* workspace, void* stream) {
  if (!p) return fail(CNF_ERR_ARG, "null plan");
  Ten Z, P, X, W;
  TRY(borrow(zy, "zy", 4, &Z));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(xy, "xy", 4, &X));
  TRY(borrow(workspace, "workspace", 1, &W, true));
  TRY(check_flow_tensor(p, Z, "zy"));
  TRY(check_flow_tensor(p, X, "xy"));
  const int64_t B = Z.shape[0];
  if (X.shape[0] != B) return fail(CNF_ERR_SHAPE, "xy batch size differs from zy (%lld)", (long long)B);
  if (P.numel < p->param_count) return fail(CNF_ERR_SHAPE, "params: need %lld floats, got %lld", (long long)p->param_count, (long long)P.numel);
  if (W.bytes < cnf_plan_workspace_bytes(p, B)) return fail(CNF_ERR_WORKSPACE, "workspace: need %lld bytes, got %lld", (long long)cnf_plan_workspace_bytes(p, B), (long long)W.bytes);
  if (B == 0) return CNF_OK;
  if (X.p != Z.p) TRY(cuda_rc(launch_copy(Z.p, X.p, Z.numel, stream), "copy"));
  return run_flow(p, P.p, X.p, (int)B, -1, nullptr, W.p, stream);
}

int cnf_prior_loss(const DLManagedTensor* zy, const DLManagedTensor* xy, const DLManagedTensor* logdet, int x_d,
                   double lambda_y, DLManagedTensor* ll_z, DLManagedTensor* ll_y, DLManagedTensor* loss4, void* stream) {
  Ten Z, X, L, A, Bt, F;
  TRY(borrow(zy, "zy", 4, &Z));
  TRY(borrow(xy, "xy", 4, &X));
  TRY(borrow(logdet, "logdet", 1, &L));
  TRY(borrow(ll_z, "ll_z", 1, &A));
  TRY(borrow(ll_y, "ll_y", 1, &Bt));
  TRY(borrow(loss4, "loss4", 1, &F));
  for (int i = 0; i < 4; ++i)
    if (Z.shape[i] != X.shape[i]) return fail(CNF_ERR_SHAPE, "zy and xy must have the same shape");
  const int64_t B = Z.shape[0];
  const int D = (int)Z.shape[3];
  if (x_d < 1 || x_d > D) return fail(CNF_ERR_ARG, "x_d %d out of range for depth %d", x_d, D);
  if (L.shape[0] != B || A.shape[0] != B || Bt.shape[0] != B || F.shape[0] != 4)
    return fail(CNF_ERR_SHAPE, "logdet/ll_z/ll_y must be [B] and loss4 [4]");
  if (B == 0) return fail(CNF_ERR_ARG, "empty batch: the batch means of the loss are undefined");
  return cuda_rc(launch_prior_loss(Z.p, X.p, L.p, (int)B, Z.shape[1] * Z.shape[2], D, x_d, lambda_y, A.p, Bt.p, F.p, stream), "prior loss");
}

int cnf_flow_log_loss(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params, DLManagedTensor* zy,
                      DLManagedTensor* ll_z, DLManagedTensor* ll_y, DLManagedTensor* logdet, DLManagedTensor* loss4,
                      DLManagedTensor* workspace, void* stream) {
  if (!p) return fail(CNF_ERR_ARG, "null plan");
  Ten X;
  TRY(borrow(xy, "xy", 4, &X));
  if (X.shape[0] == 0) return fail(CNF_ERR_ARG, "empty batch: the batch means of the loss are undefined");
  if (zy && zy->dl_tensor.data && (char*)zy->dl_tensor.data + zy->dl_tensor.byte_offset == (char*)X.p)
    return fail(CNF_ERR_ARG, "zy must not alias xy: the L1 term of the loss reads xy after the flow has written zy");
  TRY(cnf_flow_forward(p, xy, params, zy, logdet, workspace, stream));
  return cnf_prior_loss(zy, xy, logdet, p->x_d, p->lambda_y, ll_z, ll_y, loss4, stream);
}

// ---- training: forward with saved activations, loss, full backward (cFlow.train_step, M:1850-1880) ----
namespace {
struct TrainLayout {
  int64_t state_bytes, states_off, G_off, ld_off, scratch_off, total;
  std::vector<int64_t> saved_off;
  int64_t ld2_off = 0, tmp_off = 0;   // recompute mode: dummy log-det accumulator, scratch flow buffer
};
// mode 0: every activation of every layer kept; 1: per-layer input states kept, activations re-computed layer by layer in
// the backward pass; 2: nothing kept but zy - each layer's input state is recovered from its output with the inverse law
// (M:1333-1394) and its activations are re-computed from that
TrainLayout train_layout(const cnf_plan* p, int64_t B, int mode = 0) {
  const bool recompute = mode != 0;
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  TrainLayout t;
  const int n = (int)p->couplings.size();
  t.state_bytes = al(B * p->H * p->W * p->D * 4);
  int64_t off = 0;
  t.states_off = off; off += t.state_bytes * (mode == 2 ? 1 : n);
  t.G_off = off; off += t.state_bytes;
  t.ld_off = off; off += al(B * 8);
  int64_t scratch = 0;
  for (auto* c : p->couplings) scratch = std::max(scratch, coupling_bwd_scratch_bytes(c, B));
  t.scratch_off = off; off += scratch;
  if (!recompute) {
    for (auto* c : p->couplings) { t.saved_off.push_back(off); off += coupling_saved_bytes(c, B); }
  } else {
    // ONE activation region, re-filled layer by layer during the backward pass (and used as the plain s/t-net workspace in
    // the forward pass)
    t.ld2_off = off; off += al(B * 8);
    t.tmp_off = off; off += t.state_bytes;
    int64_t region = 0;
    for (auto* c : p->couplings) region = std::max(region, std::max(coupling_saved_bytes(c, B), al(coupling_ws_bytes(c, B))));
    for (size_t i = 0; i < p->couplings.size(); ++i) t.saved_off.push_back(off);
    off += region;
  }
  t.total = off;
  return t;
}
}  // namespace

int64_t cnf_plan_train_workspace_bytes(const cnf_plan* p, int64_t batch) {
  if (!p || batch < 0) return -1;
  return train_layout(p, batch).total;
}

int64_t cnf_plan_train_workspace_bytes_recompute(const cnf_plan* p, int64_t batch) {
  if (!p || batch < 0) return -1;
  return train_layout(p, batch, 1).total;
}

int64_t cnf_plan_train_workspace_bytes_invert(const cnf_plan* p, int64_t batch) {
  if (!p || batch < 0) return -1;
  return train_layout(p, batch, 2).total;
}

static int flow_loss_and_grad(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                              DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                              DLManagedTensor* logdet, DLManagedTensor* loss4, DLManagedTensor* workspace, void* stream,
                              int mode, cnf_layer_grads_ready_fn ready = nullptr, void* user = nullptr) {
  if (mode < 0 || mode > 2) return fail(CNF_ERR_ARG, "training mode must be 0 (saved), 1 (recompute) or 2 (invert)");
  const bool recompute = mode != 0, invert = mode == 2;
  if (!p) return fail(CNF_ERR_ARG, "null plan");
  Ten X, P, Gd, Z, A, Bt, L, F, W;
  TRY(borrow(xy, "xy", 4, &X));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(grads, "grads", 1, &Gd));
  TRY(borrow(zy, "zy", 4, &Z));
  TRY(borrow(ll_z, "ll_z", 1, &A));
  TRY(borrow(ll_y, "ll_y", 1, &Bt));
  TRY(borrow(logdet, "logdet", 1, &L));
  TRY(borrow(loss4, "loss4", 1, &F));
  TRY(borrow(workspace, "workspace", 1, &W, true));
  TRY(check_flow_tensor(p, X, "xy"));
  TRY(check_flow_tensor(p, Z, "zy"));
  const int64_t B = X.shape[0];
  if (B == 0) return fail(CNF_ERR_ARG, "empty batch: the batch means of the loss are undefined");
  if (Z.shape[0] != B || L.shape[0] != B || A.shape[0] != B || Bt.shape[0] != B || F.shape[0] != 4)
    return fail(CNF_ERR_SHAPE, "zy must be [B,H,W,D], ll_z/ll_y/logdet [B] and loss4 [4]");
  if (P.numel < p->param_count || Gd.numel < p->param_count)
    return fail(CNF_ERR_SHAPE, "params/grads: need %lld floats", (long long)p->param_count);
  if (X.p == Z.p) return fail(CNF_ERR_ARG, "zy must not alias xy");
  const TrainLayout T = train_layout(p, B, mode);
  if (W.bytes < T.total) return fail(CNF_ERR_WORKSPACE, "workspace: need %lld bytes, got %lld", (long long)T.total, (long long)W.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)W.p;
  const int n = (int)p->couplings.size();
  const long long per = (long long)p->H * p->W * p->D;
  double* ldacc = (double*)(base + T.ld_off);
  float* G = (float*)(base + T.G_off);
  void* scratch = base + T.scratch_off;
  TRY(cuda_rc((int)cudaMemsetAsync(ldacc, 0, sizeof(double) * B, st), "memset"));
  TRY(cuda_rc((int)cudaMemsetAsync(Gd.p, 0, sizeof(float) * p->param_count, st), "memset"));
  std::vector<CouplingSaved> saved;
  // ---- forward: state[l+1] = layer_l(state[l]); the last state is zy
  if (invert) {
    // nothing is kept: the inference path (activation-resident launches where a layer fits them) in place on zy
    TRY(cuda_rc(launch_copy(X.p, Z.p, B * per, stream), "copy"));
    for (int li = 0; li < n; ++li) {
      const cnf_coupling* c = p->couplings[li];
      saved.push_back(carve_saved(c, B, base + T.saved_off[li]));
      saved.back().state = (float*)(base + T.states_off);
      FlowView v = make_view(Z.p, p->H, p->W, p->D, p->level[li]);
      TRY(cuda_rc(run_coupling(c, P.p + p->param_off[li], v, c->mask, v, (int)B, HEAD_FWD, ldacc, nullptr, nullptr,
                               base + T.saved_off[li], stream), "coupling layer"));
    }
    if (n) TRY(cuda_rc(launch_copy(Z.p, (float*)(base + T.states_off), B * per, stream), "copy"));
  }
  for (int li = 0; li < n && !invert; ++li) {
    const cnf_coupling* c = p->couplings[li];
    saved.push_back(carve_saved(c, B, base + T.saved_off[li]));
    CouplingSaved& sv = saved.back();
    sv.state = (float*)(base + T.states_off + T.state_bytes * li);
    float* dst = li == n - 1 ? Z.p : (float*)(base + T.states_off + T.state_bytes * (li + 1));
    if (li == 0) TRY(cuda_rc(launch_copy(X.p, sv.state, B * per, stream), "copy"));
    TRY(cuda_rc(launch_copy(sv.state, dst, B * per, stream), "copy"));
    FlowView v = make_view(dst, p->H, p->W, p->D, p->level[li]);
    // recompute mode: plain forward (the shared region is the s/t-net workspace); only the layer's input state is kept
    // (the layer-per-kernel path in both modes: the backward pass re-runs exactly these kernels, so the two modes give the
    // same loss and gradients)
    TRY(cuda_rc(run_coupling(c, P.p + p->param_off[li], v, c->mask, v, (int)B, HEAD_FWD, ldacc, nullptr, nullptr,
                             recompute ? base + T.saved_off[li] : nullptr, stream, recompute ? nullptr : &sv,
                             CNF_PATH_NO_RESIDENT), "coupling layer"));
  }
  if (n == 0) TRY(cuda_rc(launch_copy(X.p, Z.p, B * per, stream), "copy"));
  TRY(cuda_rc(launch_logdet_finalize(ldacc, L.p, (int)B, stream), "logdet"));
  TRY(cuda_rc(launch_prior_loss(Z.p, X.p, L.p, (int)B, (int64_t)p->H * p->W, p->D, p->x_d, p->lambda_y, A.p, Bt.p, F.p, stream), "prior loss"));
  // ---- backward
  const float invB = 1.0f / (float)B;
  TRY(cuda_rc(launch_loss_grad(Z.p, X.p, G, B * per, p->D, p->x_d, (float)p->lambda_y, invB, stream), "loss gradient"));
  for (int li = n - 1; li >= 0; --li) {
    const cnf_coupling* c = p->couplings[li];
    if (invert) {
      // the running state is this layer's OUTPUT: the inverse law turns it into the layer's input (u1 is untouched by the
      // layer, so the s/t nets see the same values; u2 = (v2 - b) / exp(A))
      FlowView cv = make_view(saved[li].state, p->H, p->W, p->D, p->level[li]);
      TRY(cuda_rc(run_coupling(c, P.p + p->param_off[li], cv, c->mask, cv, (int)B, HEAD_INV, nullptr, nullptr, nullptr,
                               base + T.saved_off[li], stream), "coupling layer (inverse)"));
    }
    if (recompute) {
      // re-run this layer's forward from its input state with every activation saved (mode 1: same kernels, same
      // reduction orders as the first pass, so the activations are the ones it produced)
      float* tmp = (float*)(base + T.tmp_off);
      double* ld2 = (double*)(base + T.ld2_off);
      TRY(cuda_rc(launch_copy(saved[li].state, tmp, B * per, stream), "copy"));
      FlowView tv = make_view(tmp, p->H, p->W, p->D, p->level[li]);
      TRY(cuda_rc(run_coupling(c, P.p + p->param_off[li], tv, c->mask, tv, (int)B, HEAD_FWD, ld2, nullptr, nullptr, nullptr, stream,
                               &saved[li]), "coupling layer (recompute)"));
    }
    FlowView gv = make_view(G, p->H, p->W, p->D, p->level[li]);
    FlowView sview = make_view(saved[li].state, p->H, p->W, p->D, p->level[li]);
    TRY(cuda_rc(run_coupling_backward(c, P.p + p->param_off[li], Gd.p + p->param_off[li], saved[li], gv, sview, (int)B,
                                      invB, scratch, stream), "coupling layer backward"));
    // everything that writes grads[param_off[li], + param_count) is enqueued: the caller may queue this slice's collective
    // behind the stream's current position while the layers below are still being differentiated
    if (ready) ready(user, li, p->param_off[li], 2 * c->net_stride);
  }
  return CNF_OK;
}

int cnf_flow_loss_and_grad_hooked(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                                  DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                                  DLManagedTensor* logdet, DLManagedTensor* loss4, DLManagedTensor* workspace, void* stream,
                                  int mode, cnf_layer_grads_ready_fn ready, void* user) {
  return flow_loss_and_grad(p, xy, params, grads, zy, ll_z, ll_y, logdet, loss4, workspace, stream, mode, ready, user);
}

int cnf_flow_loss_and_grad(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                           DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                           DLManagedTensor* logdet, DLManagedTensor* loss4, DLManagedTensor* workspace, void* stream) {
  return flow_loss_and_grad(p, xy, params, grads, zy, ll_z, ll_y, logdet, loss4, workspace, stream, 0);
}

int cnf_flow_loss_and_grad_recompute(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                                     DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                                     DLManagedTensor* logdet, DLManagedTensor* loss4, DLManagedTensor* workspace,
                                     void* stream) {
  return flow_loss_and_grad(p, xy, params, grads, zy, ll_z, ll_y, logdet, loss4, workspace, stream, 1);
}

int cnf_flow_loss_and_grad_invert(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                                  DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                                  DLManagedTensor* logdet, DLManagedTensor* loss4, DLManagedTensor* workspace, void* stream) {
  return flow_loss_and_grad(p, xy, params, grads, zy, ll_z, ll_y, logdet, loss4, workspace, stream, 2);
}

int cnf_adam_step(DLManagedTensor* params, const DLManagedTensor* grads, DLManagedTensor* m, DLManagedTensor* v,
                  int64_t step, double lr, double beta_1, double beta_2, double epsilon, double grad_scale, void* stream) {
  Ten P, Gd, M, V;
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(grads, "grads", 1, &Gd));
  TRY(borrow(m, "m", 1, &M));
  TRY(borrow(v, "v", 1, &V));
  if (Gd.numel != P.numel || M.numel != P.numel || V.numel != P.numel) return fail(CNF_ERR_SHAPE, "params, grads, m, v must have the same length");
  if (step < 1) return fail(CNF_ERR_ARG, "step counts from 1");
  const double lr_t = lr * std::sqrt(1.0 - std::pow(beta_2, (double)step)) / (1.0 - std::pow(beta_1, (double)step));
  return cuda_rc(launch_adam(P.p, Gd.p, M.p, V.p, P.numel, (float)lr_t, (float)beta_1, (float)beta_2, (float)epsilon,
                             (float)grad_scale, stream), "adam");
}

static int coupling_common(const cnf_coupling* c, const DLManagedTensor* in, const DLManagedTensor* params,
                           DLManagedTensor* out, DLManagedTensor* logdet, DLManagedTensor* workspace, void* stream,
                           int mode) {
  if (!c) return fail(CNF_ERR_ARG, "null coupling layer");
  Ten I, P, O, L, W;
  TRY(borrow(in, "input", 4, &I));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(out, "output", 4, &O));
  TRY(borrow(workspace, "workspace", 1, &W, true));
  if (I.shape[1] != c->H || I.shape[2] != c->W || I.shape[3] != c->D)  // tf.ensure_shape, M:1276 / M:1348
    return fail(CNF_ERR_SHAPE, "input: expected shape [B,%d,%d,%d], got [%lld,%lld,%lld,%lld]", c->H, c->W, c->D,
                (long long)I.shape[0], (long long)I.shape[1], (long long)I.shape[2], (long long)I.shape[3]);
  for (int i = 0; i < 4; ++i)
    if (O.shape[i] != I.shape[i]) return fail(CNF_ERR_SHAPE, "output must have the input's shape");
  if (O.p == I.p) return fail(CNF_ERR_ARG, "output must not alias input");
  const int64_t B = I.shape[0];
  if (P.numel < 2 * c->net_stride) return fail(CNF_ERR_SHAPE, "params: need %lld floats, got %lld", (long long)(2 * c->net_stride), (long long)P.numel);
  if (W.bytes < cnf_coupling_workspace_bytes(c, B)) return fail(CNF_ERR_WORKSPACE, "workspace: need %lld bytes, got %lld", (long long)cnf_coupling_workspace_bytes(c, B), (long long)W.bytes);
  double* ldacc = nullptr;
  if (mode == HEAD_FWD) {
    TRY(borrow(logdet, "logdet", 1, &L));
    if (L.shape[0] != B) return fail(CNF_ERR_SHAPE, "logdet must be [B]");
    ldacc = (double*)((char*)W.p + coupling_ws_bytes(c, B) + 256);
  }
  if (B == 0) return CNF_OK;
  TRY(cuda_rc(launch_copy(I.p, O.p, I.numel, stream), "copy"));
  if (ldacc) TRY(cuda_rc((int)cudaMemsetAsync(ldacc, 0, sizeof(double) * B, (cudaStream_t)stream), "memset"));
  FlowView v = make_view(O.p, c->H, c->W, c->D, 0);
  TRY(cuda_rc(run_coupling(c, P.p, v, c->mask, v, (int)B, mode, ldacc, nullptr, nullptr, W.p, stream), "coupling layer"));
  if (ldacc) return cuda_rc(launch_logdet_finalize(ldacc, L.p, (int)B, stream), "logdet");
  return CNF_OK;
}

int cnf_coupling_forward(const cnf_coupling* c, const DLManagedTensor* u, const DLManagedTensor* params, DLManagedTensor* v,
                         DLManagedTensor* logdet, DLManagedTensor* workspace, void* stream) {
  return coupling_common(c, u, params, v, logdet, workspace, stream, HEAD_FWD);
}

int cnf_coupling_backward(const cnf_coupling* c, const DLManagedTensor* v, const DLManagedTensor* params, DLManagedTensor* u,
                          DLManagedTensor* workspace, void* stream) {
  return coupling_common(c, v, params, u, nullptr, workspace, stream, HEAD_INV);
}

int cnf_coupling_nets(const cnf_coupling* c, const DLManagedTensor* u1c, const DLManagedTensor* params, DLManagedTensor* A,
                      DLManagedTensor* b, DLManagedTensor* workspace, void* stream) {
  if (!c) return fail(CNF_ERR_ARG, "null coupling layer");
  Ten I, P, TA, TB, W;
  TRY(borrow(u1c, "u1_compressed", 4, &I));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(A, "A", 4, &TA));
  TRY(borrow(b, "b", 4, &TB));
  TRY(borrow(workspace, "workspace", 1, &W, true));
  if (I.shape[1] != c->h || I.shape[2] != c->w || I.shape[3] != c->c1)
    return fail(CNF_ERR_SHAPE, "u1_compressed: expected shape [B,%d,%d,%d]", c->h, c->w, c->c1);
  const int64_t B = I.shape[0];
  for (const Ten* t : {&TA, &TB})
    if (t->shape[0] != B || t->shape[1] != c->h || t->shape[2] != c->w || t->shape[3] != c->c2)
      return fail(CNF_ERR_SHAPE, "A/b: expected shape [B,%d,%d,%d]", c->h, c->w, c->c2);
  if (P.numel < 2 * c->net_stride) return fail(CNF_ERR_SHAPE, "params: need %lld floats, got %lld", (long long)(2 * c->net_stride), (long long)P.numel);
  if (W.bytes < cnf_coupling_workspace_bytes(c, B)) return fail(CNF_ERR_WORKSPACE, "workspace: need %lld bytes, got %lld", (long long)cnf_coupling_workspace_bytes(c, B), (long long)W.bytes);
  if (B == 0) return CNF_OK;
  FlowView v = make_view(I.p, c->h, c->w, c->c1, 0);
  return cuda_rc(run_coupling(c, P.p, v, MASK_DENSE, v, (int)B, HEAD_EMIT, nullptr, TA.p, TB.p, W.p, stream), "s/t networks");
}

int cnf_coupling_resident_eligible(const cnf_coupling* c) {
  if (!c || (c->paths & CNF_PATH_NO_RESIDENT)) return 0;
  FlowView v = make_view(nullptr, c->H, c->W, c->D, 0);
  return launch_fused_coupling(c, nullptr, v, c->mask, v, 1, HEAD_FWD, nullptr, nullptr, nullptr, true) == 0 ? 1 : 0;
}

int cnf_residual_block(const cnf_coupling* c, int block, const DLManagedTensor* x, const DLManagedTensor* params,
                       DLManagedTensor* out, DLManagedTensor* workspace, void* stream) {
  if (!c) return fail(CNF_ERR_ARG, "null coupling layer");
  Ten X, P, O, W;
  TRY(borrow(x, "x", 5, &X));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(out, "out", 5, &O));
  TRY(borrow(workspace, "workspace", 1, &W, true));
  if (block < 0 || block >= c->R) return fail(CNF_ERR_ARG, "block %d out of range (the layer has %d residual blocks)", block, c->R);
  const int64_t B = X.shape[1];
  for (const Ten* t : {&X, &O})
    if (t->shape[0] != 2 || t->shape[1] != B || t->shape[2] != c->h || t->shape[3] != c->w || t->shape[4] != c->nk)
      return fail(CNF_ERR_SHAPE, "x/out: expected shape [2,B,%d,%d,%d] (net A, net b)", c->h, c->w, c->nk);
  if (P.numel < 2 * c->net_stride) return fail(CNF_ERR_SHAPE, "params: need %lld floats, got %lld", (long long)(2 * c->net_stride), (long long)P.numel);
  if (W.bytes < cnf_coupling_workspace_bytes(c, B)) return fail(CNF_ERR_WORKSPACE, "workspace: need %lld bytes, got %lld", (long long)cnf_coupling_workspace_bytes(c, B), (long long)W.bytes);
  if (B == 0) return CNF_OK;
  return cuda_rc(run_residual_block(c, P.p, block, X.p, O.p, (int)B, W.p, stream), "residual block");
}

int cnf_measure_stage(const cnf_coupling* c, const DLManagedTensor* params, DLManagedTensor* workspace, int64_t batch,
                      int which, void* stream) {
  if (!c) return fail(CNF_ERR_ARG, "null coupling layer");
  Ten P, W;
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(workspace, "workspace", 1, &W, true));
  if (batch < 1 || which < 0 || which > 2) return fail(CNF_ERR_ARG, "batch >= 1 and which in {0,1,2} required");
  if (P.numel < 2 * c->net_stride) return fail(CNF_ERR_SHAPE, "params too small");
  if (W.bytes < cnf_coupling_workspace_bytes(c, batch)) return fail(CNF_ERR_WORKSPACE, "workspace too small");
  return cuda_rc(run_pw_only(c, P.p, (int)batch, which, W.p, stream), "1x1 conv");
}

#ifdef CNF_DEBUG
int cnf_debug_read_clocks(long long* out, int n) {
  return cuda_rc(read_tc3_clocks(out, n), "debug clocks");
}
#endif

int cnf_coupling_law(const DLManagedTensor* u, const DLManagedTensor* s, const DLManagedTensor* t, int which_mask,
                     int inverse, DLManagedTensor* v, DLManagedTensor* logdet, void* stream) {
  Ten U, S, T, V, L;
  TRY(borrow(u, "u", 4, &U));
  TRY(borrow(s, "s", 4, &S));
  TRY(borrow(t, "t", 4, &T));
  TRY(borrow(v, "v", 4, &V));
  if (which_mask < 0 || which_mask > 3) return fail(CNF_ERR_ARG, "which_mask must be one of 0,1,2,3");
  const int64_t B = U.shape[0];
  const int H = (int)U.shape[1], Wd = (int)U.shape[2], D = (int)U.shape[3];
  if (H % 2 || Wd % 2) return fail(CNF_ERR_ARG, "u/v must have spatial dimensions divisible by 2.");
  const int mc = which_mask ^ 1;
  int h, w, c2;
  if (mc < 2) { h = H / 2; w = Wd / 2; c2 = 2 * D; }
  else { h = H; w = Wd; c2 = mc == 2 ? (D + 1) / 2 : D / 2; }
  for (const Ten* x : {&S, &T})
    if (x->shape[0] != B || x->shape[1] != h || x->shape[2] != w || x->shape[3] != c2)
      return fail(CNF_ERR_SHAPE, "s/t: expected the compressed complement shape [B,%d,%d,%d]", h, w, c2);
  for (int i = 0; i < 4; ++i)
    if (V.shape[i] != U.shape[i]) return fail(CNF_ERR_SHAPE, "v must have u's shape");
  float* lp = nullptr;
  if (logdet) {
    TRY(borrow(logdet, "logdet", 1, &L));
    if (L.shape[0] != B) return fail(CNF_ERR_SHAPE, "logdet must be [B]");
    lp = L.p;
  }
  if ((int64_t)H * Wd * D >= (1LL << 31)) return fail(CNF_ERR_UNSUPPORTED, "sample too large");
  if (B == 0) return CNF_OK;
  if (B > 65535) return fail(CNF_ERR_UNSUPPORTED, "batch > 65535 in one call");
  return cuda_rc(launch_coupling_law(U.p, S.p, T.p, V.p, lp, (int)B, H, Wd, D, which_mask, inverse, stream), "coupling law");
}

int cnf_mask(const DLManagedTensor* uv, int which_mask, int compress, DLManagedTensor* out, void* stream) {
  Ten U, O;
  TRY(borrow(uv, "uv", 4, &U));
  TRY(borrow(out, "out", 4, &O));
  if (which_mask < 0 || which_mask > 3) return fail(CNF_ERR_ARG, "which_mask must be one of 0,1,2,3");
  const int B = (int)U.shape[0], H = (int)U.shape[1], W = (int)U.shape[2], D = (int)U.shape[3];
  if (H % 2 || W % 2) return fail(CNF_ERR_ARG, "u/v must have spatial dimensions divisible by 2.");
  int64_t want[4] = {B, H, W, D};
  if (compress) {
    if (which_mask < 2) { want[1] = H / 2; want[2] = W / 2; want[3] = 2 * D; }
    else want[3] = which_mask == 2 ? (D + 1) / 2 : D / 2;
  }
  for (int i = 0; i < 4; ++i)
    if (O.shape[i] != want[i]) return fail(CNF_ERR_SHAPE, "out: expected [%lld,%lld,%lld,%lld]", (long long)want[0], (long long)want[1], (long long)want[2], (long long)want[3]);
  return cuda_rc(launch_mask(U.p, O.p, B, H, W, D, which_mask, compress, stream), "mask");
}

int cnf_decompress_mask(const DLManagedTensor* uvc, int which_mask, DLManagedTensor* out, void* stream) {
  Ten U, O;
  TRY(borrow(uvc, "uv_masked_compressed", 4, &U));
  TRY(borrow(out, "out", 4, &O));
  if (which_mask < 0 || which_mask > 3) return fail(CNF_ERR_ARG, "which_mask must be one of 0,1,2,3");
  const int B = (int)O.shape[0], H = (int)O.shape[1], W = (int)O.shape[2], D = (int)O.shape[3];
  if (H % 2 || W % 2) return fail(CNF_ERR_ARG, "u/v must have spatial dimensions divisible by 2.");
  int64_t want[4] = {B, H, W, D};
  if (which_mask < 2) { want[1] = H / 2; want[2] = W / 2; want[3] = 2 * D; }
  else want[3] = which_mask == 2 ? (D + 1) / 2 : D / 2;
  for (int i = 0; i < 4; ++i)
    if (U.shape[i] != want[i]) return fail(CNF_ERR_SHAPE, "uv_masked_compressed: expected [%lld,%lld,%lld,%lld]", (long long)want[0], (long long)want[1], (long long)want[2], (long long)want[3]);
  return cuda_rc(launch_decompress(U.p, O.p, B, H, W, D, which_mask, stream), "decompress_mask");
}

int cnf_space_to_depth(const DLManagedTensor* in, DLManagedTensor* out, void* stream) {
  Ten I, O;
  TRY(borrow(in, "in", 4, &I));
  TRY(borrow(out, "out", 4, &O));
  const int B = (int)I.shape[0], H = (int)I.shape[1], W = (int)I.shape[2], C = (int)I.shape[3];
  if (H % 2 || W % 2) return fail(CNF_ERR_ARG, "u must have spatial dimensions divisible by 2.");  // M:175-177
  if (O.shape[0] != B || O.shape[1] != H / 2 || O.shape[2] != W / 2 || O.shape[3] != 4 * C)
    return fail(CNF_ERR_SHAPE, "out: expected [%d,%d,%d,%d]", B, H / 2, W / 2, 4 * C);
  return cuda_rc(launch_space_to_depth(I.p, O.p, B, H, W, C, 0, stream), "space_to_depth");
}

int cnf_depth_to_space(const DLManagedTensor* in, DLManagedTensor* out, void* stream) {
  Ten I, O;
  TRY(borrow(in, "in", 4, &I));
  TRY(borrow(out, "out", 4, &O));
  const int B = (int)I.shape[0], h = (int)I.shape[1], w = (int)I.shape[2], C4 = (int)I.shape[3];
  if (C4 % 4) return fail(CNF_ERR_ARG, "v must have channel dimensions divisible by 4.");  // M:208-209
  if (O.shape[0] != B || O.shape[1] != 2 * h || O.shape[2] != 2 * w || O.shape[3] != C4 / 4)
    return fail(CNF_ERR_SHAPE, "out: expected [%d,%d,%d,%d]", B, 2 * h, 2 * w, C4 / 4);
  return cuda_rc(launch_space_to_depth(I.p, O.p, B, 2 * h, 2 * w, C4 / 4, 1, stream), "depth_to_space");
}

// ---- data helpers on either side of the flow --------------------------------
int cnf_down(const DLManagedTensor* img, int levels, DLManagedTensor* out, void* stream) {
  Ten I, O;
  TRY(borrow(img, "img", 4, &I));
  TRY(borrow(out, "out", 4, &O));
  if (levels < 1 || levels > 4) return fail(CNF_ERR_ARG, "down: levels must be 1..4 (got %d)", levels);
  const int B = (int)I.shape[0], H = (int)I.shape[1], W = (int)I.shape[2], D = (int)I.shape[3];
  if (O.shape[0] != B || O.shape[1] != (H >> levels) || O.shape[2] != (W >> levels) || O.shape[3] != D)
    return fail(CNF_ERR_SHAPE, "out: expected [%d,%d,%d,%d]", B, H >> levels, W >> levels, D);
  return cuda_rc(launch_down(I.p, O.p, B, H, W, D, levels, stream), "down");
}

int cnf_up(const DLManagedTensor* img, int levels, DLManagedTensor* out, void* stream) {
  Ten I, O;
  TRY(borrow(img, "img", 4, &I));
  TRY(borrow(out, "out", 4, &O));
  if (levels < 1 || levels > 4) return fail(CNF_ERR_ARG, "up: levels must be 1..4 (got %d)", levels);
  const int B = (int)I.shape[0], H = (int)I.shape[1], W = (int)I.shape[2], D = (int)I.shape[3];
  if (O.shape[0] != B || O.shape[1] != ((int64_t)H << levels) || O.shape[2] != ((int64_t)W << levels) || O.shape[3] != D)
    return fail(CNF_ERR_SHAPE, "out: expected [%d,%d,%d,%d]", B, H << levels, W << levels, D);
  return cuda_rc(launch_up(I.p, O.p, B, H, W, D, levels, stream), "up");
}

int cnf_sr_preprocess(const DLManagedTensor* hires, int levels_x, int levels_y, int residual, DLManagedTensor* xy,
                      void* stream) {
  Ten I, O;
  TRY(borrow(hires, "hires", 4, &I));
  TRY(borrow(xy, "xy", 4, &O));
  if (levels_x < 0 || levels_y <= levels_x || levels_y > 4)
    return fail(CNF_ERR_ARG, "sr_preprocess: need 0 <= levels_x < levels_y <= 4 (got %d, %d)", levels_x, levels_y);
  const int B = (int)I.shape[0], H = (int)I.shape[1], W = (int)I.shape[2], D = (int)I.shape[3];
  if ((H % (1 << levels_y)) || (W % (1 << levels_y)))
    return fail(CNF_ERR_ARG, "sr_preprocess: H and W must be divisible by %d", 1 << levels_y);
  if (O.shape[0] != B || O.shape[1] != (H >> levels_x) || O.shape[2] != (W >> levels_x) || O.shape[3] != 2 * D)
    return fail(CNF_ERR_SHAPE, "xy: expected [%d,%d,%d,%d]", B, H >> levels_x, W >> levels_x, 2 * D);
  return cuda_rc(launch_sr_preprocess(I.p, O.p, B, H, W, D, levels_x, levels_y, residual, stream), "sr_preprocess");
}

int cnf_logit_scale(const DLManagedTensor* x, double a, int inverse, DLManagedTensor* out, void* stream) {
  Ten I, O;
  TRY(borrow(x, "x", -1, &I));
  TRY(borrow(out, "out", -1, &O));
  if (!(a > 0.0 && a < 0.5)) return fail(CNF_ERR_ARG, "logit_scale: the fudge factor a must be in (0, 0.5)");
  if (O.numel != I.numel) return fail(CNF_ERR_SHAPE, "out: expected %lld elements", (long long)I.numel);
  return cuda_rc(launch_logit(I.p, O.p, I.numel, a, inverse, stream), "logit_scale");
}

int cnf_instance_noise(const DLManagedTensor* x, double alpha, uint64_t seed, uint64_t offset, DLManagedTensor* out,
                       void* stream) {
  Ten I, O;
  TRY(borrow(out, "out", -1, &O));
  if (x) {
    TRY(borrow(x, "x", -1, &I));
    if (O.numel != I.numel) return fail(CNF_ERR_SHAPE, "out: expected %lld elements", (long long)I.numel);
  }
  return cuda_rc(launch_instance_noise(x ? I.p : nullptr, O.p, O.numel, x ? alpha : 0.0, seed, offset, stream),
                 "instance_noise");
}

// ---- toy ------------------------------------------------------------------------------------
static long long toy_net_size_h(int I, int num_layers) {
  return 2LL * I + I + (long long)num_layers * ((long long)I * I + I) + 2LL * I + 4;
}

int64_t cnf_toy_layer_offset(int layer, int intermediate_dims, int num_layers) {
  return 2 * toy_net_size_h(intermediate_dims, num_layers) * layer;
}

int64_t cnf_toy_param_count(int num_coupling_layers, int intermediate_dims, int num_layers) {
  return cnf_toy_layer_offset(num_coupling_layers, intermediate_dims, num_layers);
}

static int toy_check(int n, int width, int num_layers, const int* order) {
  if (n < 1 || n > 256) return fail(CNF_ERR_UNSUPPORTED, "1..256 coupling layers are supported, got %d", n);
  if (width != 8 && width != 16 && width != 32 && width != 64)
    return fail(CNF_ERR_UNSUPPORTED, "intermediate_dims must be one of 8, 16, 32, 64 (got %d)", width);
  if (num_layers < 0) return fail(CNF_ERR_ARG, "num_layers must be >= 0");
  if (!order) return fail(CNF_ERR_ARG, "null mask_indices");
  for (int i = 0; i < n; ++i)
    if (order[i] < 0 || order[i] >= n) return fail(CNF_ERR_ARG, "mask_indices[%d] = %d out of range", i, order[i]);
  return CNF_OK;
}

int cnf_toy_call(const DLManagedTensor* u, const DLManagedTensor* params, const int* mask_indices, int n, int width,
                 int num_layers, int direction, DLManagedTensor* v, DLManagedTensor* logdet, void* stream) {
  Ten U, P, V, L;
  TRY(toy_check(n, width, num_layers, mask_indices));
  if (direction != 1 && direction != -1) return fail(CNF_ERR_ARG, "direction must be +1 or -1");
  TRY(borrow(u, "u", 2, &U));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(v, "v", 2, &V));
  TRY(borrow(logdet, "logdet", 1, &L));
  if (U.shape[1] != 3 || V.shape[1] != 3 || V.shape[0] != U.shape[0] || L.shape[0] != U.shape[0])
    return fail(CNF_ERR_SHAPE, "u, v must be [B,3] and logdet [B]");
  if (P.numel < cnf_toy_param_count(n, width, num_layers)) return fail(CNF_ERR_SHAPE, "params too small");
  return cuda_rc(launch_toy(U.p, P.p, mask_indices, n, width, num_layers, direction, V.p, L.p, (int)U.shape[0], stream), "toy flow");
}

int cnf_toy_log_loss(const DLManagedTensor* xy, const DLManagedTensor* params, const int* mask_indices, int n, int width,
                     int num_layers, int x_d, double lambda_y, DLManagedTensor* zy, DLManagedTensor* ll_z,
                     DLManagedTensor* ll_y, DLManagedTensor* logdet, DLManagedTensor* loss4, void* stream) {
  Ten X, A, Bt, L, F, Z;
  TRY(borrow(xy, "xy", 2, &X));
  TRY(borrow(zy, "zy", 2, &Z));
  TRY(borrow(ll_z, "ll_z", 1, &A));
  TRY(borrow(ll_y, "ll_y", 1, &Bt));
  TRY(borrow(logdet, "logdet", 1, &L));
  TRY(borrow(loss4, "loss4", 1, &F));
  const int64_t B = X.shape[0];
  if (B == 0) return fail(CNF_ERR_ARG, "empty batch: the batch means of the loss are undefined");
  if (x_d < 1 || x_d > 3) return fail(CNF_ERR_ARG, "x_d out of range");
  if (A.shape[0] != B || Bt.shape[0] != B || F.shape[0] != 4) return fail(CNF_ERR_SHAPE, "ll_z/ll_y must be [B] and loss4 [4]");
  TRY(cnf_toy_call(xy, params, mask_indices, n, width, num_layers, -1, zy, logdet, stream));
  return cuda_rc(launch_toy_loss(Z.p, X.p, L.p, (int)B, x_d, lambda_y, A.p, Bt.p, F.p, stream), "toy loss");
}

int cnf_toy_loss_and_grad(const DLManagedTensor* xy, const DLManagedTensor* params, const int* mask_indices, int n, int width,
                          int num_layers, int x_d, double lambda_y, DLManagedTensor* grads, DLManagedTensor* zy,
                          DLManagedTensor* ll_z, DLManagedTensor* ll_y, DLManagedTensor* logdet, DLManagedTensor* loss4,
                          void* stream) {
  Ten X, P, G, Z;
  TRY(borrow(xy, "xy", 2, &X));
  TRY(borrow(params, "params", 1, &P));
  TRY(borrow(grads, "grads", 1, &G));
  TRY(borrow(zy, "zy", 2, &Z));
  const int64_t need = cnf_toy_param_count(n, width, num_layers);
  if (need <= 0) return fail(CNF_ERR_UNSUPPORTED, "toy configuration not built (width must be 8, 16, 32 or 64)");
  if (P.numel < need || G.numel < need) return fail(CNF_ERR_SHAPE, "params/grads: need %lld floats", (long long)need);
  if (X.p == Z.p) return fail(CNF_ERR_ARG, "zy must not alias xy");
  TRY(cnf_toy_log_loss(xy, params, mask_indices, n, width, num_layers, x_d, lambda_y, zy, ll_z, ll_y, logdet, loss4, stream));
  TRY(cuda_rc((int)cudaMemsetAsync(G.p, 0, sizeof(float) * need, (cudaStream_t)stream), "memset"));
  return cuda_rc(launch_toy_grad(X.p, Z.p, P.p, G.p, mask_indices, n, width, num_layers, x_d, lambda_y, (int)X.shape[0], stream),
                 "toy gradients");
}

}  // extern "C"
