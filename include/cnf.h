/* libcnf — C-ABI of the B200-native conditional RealNVP hot path.
 *
 * The reference (USArmyResearchLab/ARL_Conditional_Normalizing_Flows) has no FFI: its boundary is
 * the Python class surface in conv_cINN_make_model.py (M), conv_cINN_base_functions.py (F) and
 * TOYcINN_make_model.py (T).  Each entry point below names the reference interface it replaces.
 * Python host code (arl_conditional_normalizing_flows_b200/) binds these with ctypes and passes
 * tensors zero-copy as DLPack `DLManagedTensor*`.
 *
 * Conventions
 *   - every function returns an int status (CNF_OK or a CNF_ERR_* class) and never throws;
 *     `cnf_last_error()` returns the thread-local message of the last failure;
 *   - tensors are fp32, on the CUDA device that is current for the calling thread, compact
 *     row-major NHWC, 16-byte aligned; libcnf borrows them for the call and never frees them;
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); no
 *     function synchronises or allocates device memory; the only process-wide state is the
 *     per-device cache of kernel attributes (mutex-protected), so one thread per device is safe;
 *   - the caller allocates every output and the workspace (`cnf_*_workspace_bytes`).
 */
#ifndef CNF_H_
#define CNF_H_

#include <stdint.h>
#include "cnf_dlpack.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CNF_VERSION 100

#define CNF_OK 0
#define CNF_ERR_ARG 1          /* a reference `assert` would have fired      -> AssertionError     */
#define CNF_ERR_SHAPE 2        /* tf.ensure_shape mismatch (M:621,1276,1348) -> ValueError         */
#define CNF_ERR_DTYPE 3        /* not float32                                -> TypeError          */
#define CNF_ERR_DEVICE 4       /* not on the current CUDA device             -> RuntimeError       */
#define CNF_ERR_LAYOUT 5       /* not compact / not 16-byte aligned          -> ValueError         */
#define CNF_ERR_CUDA 6         /* CUDA runtime error                         -> RuntimeError       */
#define CNF_ERR_UNSUPPORTED 7  /* valid in the reference, not built here     -> NotImplementedError*/
#define CNF_ERR_WORKSPACE 8    /* workspace too small                        -> ValueError         */

#define CNF_MAX_BRANCHES 8
#define CNF_MAX_BLOCKS 32
#define CNF_NAME_CAP 64

typedef struct cnf_coupling cnf_coupling; /* one coupling layer: masks + s/t nets (M:331-1394)  */
typedef struct cnf_plan cnf_plan;         /* the whole flow: cFlow.__init__ planner (M:1431-1695) */

int cnf_version(void);
const char* cnf_last_error(void);

/* ---- coupling layer descriptor ------------------------------------------------------------
 * replaces coupling_layer.__init__ (M:355-439), get_masked_compressed_shape (M:474-498) and the
 * shape logic of coupling_function (M:1087-1104).  Host ints only. */
typedef struct {
  int H, W, D;                 /* uncompressed input shape (in_shape)                             */
  int mask, mask_complement;   /* which_mask, which_mask_complement (M:426-433)                   */
  int R, cardinality, nk, ksize, layer_norm;   /* nk already halved for masks 0/1 (M:420-423)     */
  int h, w, c1, c2;            /* compressed u1 shape and the depth of u2 / of A,b (M:1093-1104)  */
  int cat;                     /* concat width after the dilated grouped convs (F:590)            */
  int n_branches;
  int dilation[CNF_MAX_BRANCHES];
  int branch_channels[CNF_MAX_BRANCHES]; /* nk // d (F:579)                                       */
  int groups[CNF_MAX_BRANCHES];          /* cardinality (1 => plain conv over all nk, F:389-395)  */
  int group_in[CNF_MAX_BRANCHES];        /* input channels per group                              */
  int group_out[CNF_MAX_BRANCHES];       /* output channels per group                             */
  int n_ln;                    /* LayerNorms per net (3R+1, or 0)                                 */
  int64_t net_stride;          /* floats from net A's parameters to net b's                       */
  int64_t param_count;         /* floats for both nets (= 2*net_stride)                           */
  int n_entries;               /* named parameter tensors per net                                 */
} cnf_coupling_info;

int cnf_coupling_create(const int in_shape[3], int which_mask, int num_res_blocks, int cardinality,
                        int num_kernels, int kernel_size, int layer_norm, const int* which_dilations,
                        int n_dilations, cnf_coupling** out);
void cnf_coupling_destroy(cnf_coupling* c);
int cnf_coupling_get_info(const cnf_coupling* c, cnf_coupling_info* info);
/* idx-th named parameter of one net, Keras-shaped (HWIO kernels, flat LN vectors); `offset` is in
 * floats from the start of that net's block.  role: 0 kernel, 1 bias, 2 gamma, 3 beta, 4 tanh scale */
int cnf_coupling_param_entry(const cnf_coupling* c, int idx, char name[CNF_NAME_CAP], int64_t* offset,
                             int* ndim, int64_t shape[4], int* role);
int64_t cnf_coupling_workspace_bytes(const cnf_coupling* c, int64_t batch);

/* ---- flow planner: replaces cFlow.__init__ (M:1431-1695) ---------------------------------- */
typedef struct {
  int n_blocks, n_coupling, n_layers; /* n_layers counts couplings + squeeze + factor layers     */
  int H, W, D, x_d, ksize, layer_norm;
  double lambda_y;
  int64_t param_count;                /* floats, all coupling layers                              */
} cnf_plan_info;

typedef struct {
  int scale, num_prev_factors;        /* scale_list[i], num_prev_factors_list[i] (M:1493-1518)    */
  int H, W, D;                        /* io_shape_list[i] (M:1521-1536)                           */
  int n_checkerboard, checkerboard[CNF_MAX_BRANCHES]; /* dilations_list[i] (M:1553-1617)          */
  int n_channelwise, channelwise[CNF_MAX_BRANCHES];
} cnf_block_info;

int cnf_plan_create(const int io_shape[3], int x_d, int n_blocks, const int* squeeze_factor_block_list,
                    const int* ResNeXt_block_list, const int* num_kernels_list,
                    const int* cardinality_list, double lambda_y, int ksize, int layer_norm,
                    int dilations, cnf_plan** out);
void cnf_plan_destroy(cnf_plan* p);
int cnf_plan_get_info(const cnf_plan* p, cnf_plan_info* info);
int cnf_plan_block_info(const cnf_plan* p, int block, cnf_block_info* info);
/* layers_list order (M:1636-1689): kind 0 coupling, 1 squeeze, 2 factor; aux = coupling index or
 * num_prev_factors */
int cnf_plan_layer(const cnf_plan* p, int idx, int* kind, int* aux);
const cnf_coupling* cnf_plan_coupling(const cnf_plan* p, int coupling_idx);
int64_t cnf_plan_coupling_param_offset(const cnf_plan* p, int coupling_idx);
int cnf_plan_coupling_level(const cnf_plan* p, int coupling_idx); /* squeezes applied before it */
int64_t cnf_plan_workspace_bytes(const cnf_plan* p, int64_t batch);
/* Inference runs a coupling layer whose per-sample s/t-net activations fit one CTA's shared memory (the 14x14 / 7x7
 * levels of config 2) as ONE activation-resident launch (csrc/fused_kernels.cu); enable = 0 forces the layer-per-kernel
 * path for every layer (A/B measurements, tests of both paths).  Default: enabled.  Host-side flag of the descriptor,
 * not thread-safe against concurrent calls that use the same plan.  Replaces nothing in the reference. */
int cnf_plan_set_fusion(cnf_plan* p, int enable);
int cnf_coupling_set_fusion(cnf_coupling* c, int enable);
/* Kernel-family selection per stage of a coupling layer.  Every stage has a fastest kernel and one or more general ones
 * that cover the shapes it does not; `excluded` is a set of CNF_PATH_* bits naming families this
 * descriptor must NOT use, so that every shipped kernel can be exercised on any shape (tests/test_gpu_parity.py) and
 * A/B-timed (bench.py).  0 (default) = the fastest eligible kernel everywhere.  All combinations give the same results to
 * fp32 rounding.  Host-side flag, not thread-safe against concurrent calls on the same descriptor. */
#define CNF_PATH_NO_RESIDENT 1  /* no activation-resident launch (fused_kernels.cu): layer-per-kernel path          */
#define CNF_PATH_NO_TCGEN05 2   /* no tcgen05 3xTF32 kernels: 1x1 convs and their data gradients (pw_tc3_kernel) -> FFMA
                                 * pw_kernel / gemm_kernel; 16 / 32-wide grouped convs and their data gradients
                                 * (gconv_tc_kernel) -> gconv_kernel / generic gradient; 1x1 weight gradients
                                 * (wgrad_tc_kernel) -> wgrad_pw_kernel                                          */
#define CNF_PATH_NO_PW_FFMA 4   /* 1x1 convs: no multi-sample FFMA kernel either -> generic gemm_kernel                */
#define CNF_PATH_NO_OCTET 8     /* grouped convs: no all-branch octet kernel (gconv_oct_kernel) -> one launch per branch */
#define CNF_PATH_NO_BRANCH 16   /* grouped convs: no per-branch gconv3_kernel -> gconv2_kernel / gconv_kernel          */
#define CNF_PATH_NO_GCONV2 32   /* grouped convs: no register-blocked gconv2_kernel -> generic gconv_kernel            */
#define CNF_PATH_NO_STEM2 64    /* stem: no per-sample stem2_kernel -> implicit-im2col gemm_kernel                     */
#define CNF_PATH_NO_HEAD2 128   /* head: no streaming head2_kernel -> halo-tile head_kernel                           */
#define CNF_PATH_ALL 255
/* 1 iff inference runs this layer as one activation-resident launch (shape covered and not excluded), else 0 */
int cnf_coupling_resident_eligible(const cnf_coupling* c);
int cnf_plan_set_kernel_paths(cnf_plan* p, int excluded);
int cnf_coupling_set_kernel_paths(cnf_coupling* c, int excluded);

/* ---- the hot path ------------------------------------------------------------------------- */
/* cFlow.call(xy, direction=+1) (M:1743-1772): zy in the ORIGINAL (H,W,D) layout plus the PER-SAMPLE
 * log-det vector [B] (superset of the reference's batch-mean scalar, Q1: scalar = mean of it).  logdet may also be
 * [B + 1]: element B then receives that batch mean (fp64 sum in a fixed order), i.e. the scalar call(+1) returns. */
int cnf_flow_forward(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                     DLManagedTensor* zy, DLManagedTensor* logdet, DLManagedTensor* workspace,
                     void* stream);
/* cFlow.call(zy, direction=-1) (M:1774-1798) */
int cnf_flow_inverse(const cnf_plan* p, const DLManagedTensor* zy, const DLManagedTensor* params,
                     DLManagedTensor* xy, DLManagedTensor* workspace, void* stream);
/* cFlow.log_loss (M:1800-1848): forward + prior/L1 reductions.  ll_z, ll_y, logdet are [B];
 * loss4 is [4] = (loss, z_loss, y_loss, detJ_loss) exactly as the reference returns them. */
int cnf_flow_log_loss(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                      DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                      DLManagedTensor* logdet, DLManagedTensor* loss4, DLManagedTensor* workspace,
                      void* stream);
/* prior / L1 part alone (M:1826-1848) given zy, xy [B,H,W,D] and a per-sample logdet [B]. */
int cnf_prior_loss(const DLManagedTensor* zy, const DLManagedTensor* xy, const DLManagedTensor* logdet,
                   int x_d, double lambda_y, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                   DLManagedTensor* loss4, void* stream);

/* ---- training step: replaces the tf.GradientTape block of cFlow.train_step (M:1863-1871) --------------
 * forward of log_loss with every s/t-net activation kept, then the hand-written backward pass.
 * grads [param_count] is OVERWRITTEN with dL/dparams of loss4[0] (same flat layout as params); the other
 * outputs are those of cnf_flow_log_loss.  workspace: cnf_plan_train_workspace_bytes(plan, B) bytes. */
int64_t cnf_plan_train_workspace_bytes(const cnf_plan* p, int64_t batch);
int cnf_flow_loss_and_grad(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                           DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z,
                           DLManagedTensor* ll_y, DLManagedTensor* logdet, DLManagedTensor* loss4,
                           DLManagedTensor* workspace, void* stream);
/* Same result with ~1/n_layers of the activation memory: the forward pass keeps only the flow state in
 * front of every coupling layer (H*W*D floats per sample and layer); during the backward pass each layer's s/t-net
 * activations are re-computed from that state into ONE shared region.  The re-computation runs the same kernels with
 * the same reduction orders, so the gradients equal those of cnf_flow_loss_and_grad; it costs one extra forward pass.
 * (Recovering the states themselves with the inverse pass, M:1333-1394, is NOT done: with trained weights the fp32
 * round trip is only ~1e-3 accurate)  workspace: cnf_plan_train_workspace_bytes_recompute bytes. */
int64_t cnf_plan_train_workspace_bytes_recompute(const cnf_plan* p, int64_t batch);
int cnf_flow_loss_and_grad_recompute(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                                     DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z,
                                     DLManagedTensor* ll_y, DLManagedTensor* logdet, DLManagedTensor* loss4,
                                     DLManagedTensor* workspace, void* stream);
/* Activation-free backward via invertibility (coupling_layer.backward, M:1333-1394): the forward pass keeps
 * nothing but zy (and runs the inference kernels); the backward pass recovers every layer's input state from its output
 * with the inverse law and re-computes the layer's activations from it.  Memory: one flow state + one layer's activations.
 * Accuracy: the recovered states carry the fp32 round-trip error of the flow (1e-6 at the reference's initial state, up to
 * 1e-3 with ill-conditioned trained-like weights), so the gradients are those of the other two modes
 * only to that accuracy.  workspace: cnf_plan_train_workspace_bytes_invert bytes. */
int64_t cnf_plan_train_workspace_bytes_invert(const cnf_plan* p, int64_t batch);
int cnf_flow_loss_and_grad_invert(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                                  DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z,
                                  DLManagedTensor* ll_y, DLManagedTensor* logdet, DLManagedTensor* loss4,
                                  DLManagedTensor* workspace, void* stream);
/* The same three training modes (mode 0 = saved activations, 1 = recompute, 2 = invert) with a host-side hook for the
 * data-parallel gradient all-reduce of cFlow.train_step (M:1850-1880 run under a distribution strategy): the
 * backward pass walks the coupling layers from the last to the first, and `ready(user, layer, param_offset, param_count)`
 * is called ON THE HOST, in that order, as soon as every kernel that writes grads[param_offset, param_offset + param_count)
 * has been enqueued on `stream`.  Work the callback queues behind the stream's current position (an event + a collective
 * on another stream) therefore overlaps the backward pass of the layers below.  The callback must not touch the other
 * arguments of the call; ready = NULL gives the plain entry points above.  Replaces nothing in the reference (keras
 * hides the bucketing inside its strategy). */
typedef void (*cnf_layer_grads_ready_fn)(void* user, int layer, int64_t param_offset, int64_t param_count);
int cnf_flow_loss_and_grad_hooked(const cnf_plan* p, const DLManagedTensor* xy, const DLManagedTensor* params,
                                  DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z,
                                  DLManagedTensor* ll_y, DLManagedTensor* logdet, DLManagedTensor* loss4,
                                  DLManagedTensor* workspace, void* stream, int mode, cnf_layer_grads_ready_fn ready,
                                  void* user);
/* optimizer.apply_gradients with keras Adam (M:1874; C:567 / P:130: lr 3e-4, beta 0.9/0.999, eps 1e-7):
 * one fused update of the flat parameter buffer; `step` counts from 1; grads are multiplied by
 * grad_scale first (1/world_size after a data-parallel sum all-reduce). */
int cnf_adam_step(DLManagedTensor* params, const DLManagedTensor* grads, DLManagedTensor* m, DLManagedTensor* v,
                  int64_t step, double lr, double beta_1, double beta_2, double epsilon, double grad_scale,
                  void* stream);

/* coupling_layer.forward_and_Jacobian (M:1258-1328) / .backward (M:1333-1394) on one [B,H,W,D]
 * tensor; `out` must not alias `in`.  logdet [B] is ACCUMULATED into (caller zeroes it). */
int cnf_coupling_forward(const cnf_coupling* c, const DLManagedTensor* u, const DLManagedTensor* params,
                         DLManagedTensor* v, DLManagedTensor* logdet, DLManagedTensor* workspace,
                         void* stream);
int cnf_coupling_backward(const cnf_coupling* c, const DLManagedTensor* v, const DLManagedTensor* params,
                          DLManagedTensor* u, DLManagedTensor* workspace, void* stream);
/* A_wrapper / b_wrapper (M:452-472): both nets on an already-compressed u1 [B,h,w,c1];
 * A = w*tanh(.) and b are [B,h,w,c2]. */
int cnf_coupling_nets(const cnf_coupling* c, const DLManagedTensor* u1_compressed,
                      const DLManagedTensor* params, DLManagedTensor* A, DLManagedTensor* b,
                      DLManagedTensor* workspace, void* stream);

/* dilated_residual_block (F:501-627, with grouped_convolution F:364-413 and add_common_layers F:330-362) number `block`
 * of the layer's two s/t networks on stand-alone activations: x, out are [2,B,h,w,nk] (net A first; out may alias x);
 * LN -> 1x1 -> LN -> grouped dilated 3x3 (concat over dilations) -> LN -> 1x1 -> + x, the launches the layer itself
 * issues for that block.  workspace: cnf_coupling_workspace_bytes(c, B). */
int cnf_residual_block(const cnf_coupling* c, int block, const DLManagedTensor* x, const DLManagedTensor* params,
                       DLManagedTensor* out, DLManagedTensor* workspace, void* stream);

/* fused standalone coupling law + mask addressing + per-sample log-det (M:1215-1253, M:1307-1326):
 * v = mask(u,m,False) + decompress(exp(s)*u2c + t, m_bar)   (inverse: (u2c - t) / exp(s)).
 * u, v are [B,H,W,D]; s, t are [B,h,w,c2] in the compressed layout of the complement mask;
 * logdet [B] is overwritten with sum(s) per sample (may be NULL). */
int cnf_coupling_law(const DLManagedTensor* u, const DLManagedTensor* s, const DLManagedTensor* t,
                     int which_mask, int inverse, DLManagedTensor* v, DLManagedTensor* logdet,
                     void* stream);

/* pure index permutations, bit-exact: coupling_layer.mask (M:500-761), decompress_mask (M:763-1073),
 * squeeze_layer (M:155-217: tf.nn.space_to_depth / depth_to_space, block 2). */
int cnf_mask(const DLManagedTensor* uv, int which_mask, int compress, DLManagedTensor* out, void* stream);
int cnf_decompress_mask(const DLManagedTensor* uv_compressed, int which_mask, DLManagedTensor* out,
                        void* stream);
int cnf_space_to_depth(const DLManagedTensor* in, DLManagedTensor* out, void* stream);
int cnf_depth_to_space(const DLManagedTensor* in, DLManagedTensor* out, void* stream);

/* ---- data helpers on either side of the flow (conv_cINN_base_functions.py), SURVEY 8f-2/8f-3 ----
 * The reference runs these as tf.data maps on the host; here each is one HBM-bound kernel on device tensors.
 * cnf_down: `levels` nested 2x2 average pools (F:74-127: down(down(..)) = mean of means; trailing odd rows /
 *   columns cropped).  img [B,H,W,D] -> out [B,H>>levels,W>>levels,D].
 * cnf_up: `levels` nested 2x2 pixel repeats (F:129-164).  img [B,H,W,D] -> out [B,H<<levels,W<<levels,D].
 * cnf_sr_preprocess: preprocess_dataset_SR (F:233-279) fused: x = down^levels_x(hires),
 *   y = up^(levels_y-levels_x)(down^levels_y(hires)), x -= y if residual, xy = concat(x, y).
 *   'SR2,1' = (0,1); 'SR4,2' = (1,2); SURVEY config 4 (8x8 condition for 64x64) = (0,3).
 *   hires [B,H,W,D] -> xy [B,H>>levels_x,W>>levels_x,2D].
 * cnf_logit_scale: inverse = 0: preprocess_dataset_class with LOGITS (F:174-231),
 *   x -> (logit(a + (1-a) b x) - logit(a)) / (logit(1-a) - logit(a)), b = (1-2a)/(1-a);
 *   inverse = 1: de_logitify (F:287-318).  Any shape; out may alias x.
 * cnf_instance_noise: out = alpha x + (1 - alpha) N(0,1) (instance_noise, F:635-653); x == NULL: out = N(0,1)
 *   (renew_noise, F:660-676).  Counter-based Philox4x32-10 + Box-Muller: elements 4q..4q+3 come from counter
 *   q + offset under key `seed`, so a (seed, offset) pair reproduces the same noise on any grid / device count. */
int cnf_down(const DLManagedTensor* img, int levels, DLManagedTensor* out, void* stream);
int cnf_up(const DLManagedTensor* img, int levels, DLManagedTensor* out, void* stream);
int cnf_sr_preprocess(const DLManagedTensor* hires, int levels_x, int levels_y, int residual, DLManagedTensor* xy,
                      void* stream);
int cnf_logit_scale(const DLManagedTensor* x, double a, int inverse, DLManagedTensor* out, void* stream);
int cnf_instance_noise(const DLManagedTensor* x, double alpha, uint64_t seed, uint64_t offset, DLManagedTensor* out,
                       void* stream);

/* ---- toy dense cINN: cINN_affine.call / log_loss (T:248-451), coupling_layer MLPs (T:29-97) --
 * u, v are [B,3]; params is the flat buffer laid out by cnf_toy_param_count(); mask_indices[n]
 * (host ints) is the layer order; direction follows the TOY convention (-1: xy->zy with log-det,
 * +1: zy->xy).  logdet [B] is overwritten (zeros for +1). */
int64_t cnf_toy_param_count(int num_coupling_layers, int intermediate_dims, int num_layers);
int64_t cnf_toy_layer_offset(int layer, int intermediate_dims, int num_layers);
int cnf_toy_call(const DLManagedTensor* u, const DLManagedTensor* params, const int* mask_indices,
                 int num_coupling_layers, int intermediate_dims, int num_layers, int direction,
                 DLManagedTensor* v, DLManagedTensor* logdet, void* stream);
int cnf_toy_log_loss(const DLManagedTensor* xy, const DLManagedTensor* params, const int* mask_indices,
                     int num_coupling_layers, int intermediate_dims, int num_layers, int x_d,
                     double lambda_y, DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                     DLManagedTensor* logdet, DLManagedTensor* loss4, void* stream);

/* the tf.GradientTape block of cINN_affine.train_step (T:453-482): cnf_toy_log_loss plus d loss4[0] / d params into
 * grads (layout of params, overwritten).  The backward pass stores no activations: every coupling layer's input is
 * recovered from its output with the inverse law (T:369-375) and its MLPs are re-evaluated. */
int cnf_toy_loss_and_grad(const DLManagedTensor* xy, const DLManagedTensor* params, const int* mask_indices,
                          int num_coupling_layers, int intermediate_dims, int num_layers, int x_d, double lambda_y,
                          DLManagedTensor* grads, DLManagedTensor* zy, DLManagedTensor* ll_z, DLManagedTensor* ll_y,
                          DLManagedTensor* logdet, DLManagedTensor* loss4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNF_H_ */
