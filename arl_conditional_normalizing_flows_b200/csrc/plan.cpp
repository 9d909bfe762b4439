// Host-side architecture planner: the integer arithmetic of cFlow.__init__ and
// coupling_layer.__init__ / coupling_function, plus the flat parameter layout the kernels read.
// Reference: conv_cINN_make_model.py M:355-439, M:474-498, M:1087-1104, M:1431-1695;
//            conv_cINN_base_functions.py F:389-411, F:577-590.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda_runtime.h>

#include "cnf_internal.h"

namespace cnf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int64_t align4(int64_t x) { return (x + 3) & ~int64_t(3); }

int knob_int(const char* name, int dflt) {
#ifdef CNF_DEBUG
  char key[96];
  snprintf(key, sizeof(key), "CNF_%s", name);
  const char* e = getenv(key);
  return e ? atoi(e) : dflt;
#else
  (void)name;
  return dflt;
#endif
}

float knob_float(const char* name, float dflt) {
#ifdef CNF_DEBUG
  char key[96];
  snprintf(key, sizeof(key), "CNF_%s", name);
  const char* e = getenv(key);
  return e ? (float)atof(e) : dflt;
#else
  (void)name;
  return dflt;
#endif
}

static std::mutex g_dev_mu;

int ensure_dynamic_smem(const void* kernel, size_t bytes, SmemAttrCache& cache) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= CNF_MAX_DEVICES) return (int)cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  const size_t want = bytes < 48 * 1024 ? (size_t)48 * 1024 : bytes;
  if (want > cache.set[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e != cudaSuccess) return (int)e;
    cache.set[dev] = want;
  }
  return 0;
}

int device_sm_count(int* n_sm) {
  static int cached[CNF_MAX_DEVICES] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= CNF_MAX_DEVICES) return (int)cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!cached[dev]) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    cached[dev] = n;
  }
  *n_sm = cached[dev];
  return 0;
}

struct LayoutBuilder {
  int64_t cur = 0;
  std::vector<ParamEntry>* entries;
  int64_t add(const std::string& name, int role, std::initializer_list<int64_t> shape) {
    cur = align4(cur);
    ParamEntry e;
    e.name = name;
    e.offset = cur;
    e.ndim = (int)shape.size();
    e.role = role;
    int64_t n = 1;
    int i = 0;
    for (int64_t s : shape) {
      e.shape[i++] = s;
      n *= s;
    }
    for (; i < 4; ++i) e.shape[i] = 1;
    entries->push_back(e);
    cur += n;
    return e.offset;
  }
  // reserve a block without a single entry (entries are added by the caller at explicit offsets)
  int64_t reserve(int64_t n) {
    cur = align4(cur);
    int64_t o = cur;
    cur += n;
    return o;
  }
  void add_at(const std::string& name, int role, int64_t offset, std::initializer_list<int64_t> shape) {
    ParamEntry e;
    e.name = name;
    e.offset = offset;
    e.ndim = (int)shape.size();
    e.role = role;
    int i = 0;
    for (int64_t s : shape) e.shape[i++] = s;
    for (; i < 4; ++i) e.shape[i] = 1;
    entries->push_back(e);
  }
};

int64_t coupling_ws_bytes(const cnf_coupling* c, int64_t B) {
  const int64_t hw = c->hw();
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  int64_t b = 0;
  b += al(2 * B * hw * c->nk * 4);                  // X
  b += al(2 * B * hw * c->nk * 4);                  // Y1
  b += al(2 * B * hw * c->cat * 4);                 // Y2
  b += al((int64_t)(c->n_ln() + 1) * 2 * B * 2 * 8);  // stats (doubles)
  return b;
}

CouplingWorkspace carve_ws(const cnf_coupling* c, int64_t B, void* ws) {
  const int64_t hw = c->hw();
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  char* p = (char*)ws;
  CouplingWorkspace w;
  w.X = (float*)p;
  p += al(2 * B * hw * c->nk * 4);
  w.Y1 = (float*)p;
  p += al(2 * B * hw * c->nk * 4);
  w.Y2 = (float*)p;
  p += al(2 * B * hw * c->cat * 4);
  w.stats = (double*)p;
  return w;
}

}  // namespace cnf

using namespace cnf;

cnf_plan::~cnf_plan() {
  for (auto* c : couplings) delete c;
}

extern "C" {

int cnf_version(void) { return CNF_VERSION; }
const char* cnf_last_error(void) { return cnf::g_err; }

int cnf_coupling_create(const int in_shape[3], int which_mask, int num_res_blocks, int cardinality,
                        int num_kernels, int kernel_size, int layer_norm, const int* which_dilations,
                        int n_dilations, cnf_coupling** out) {
  if (!in_shape || !out || !which_dilations) {
    set_error("null argument");
    return CNF_ERR_ARG;
  }
  const int H = in_shape[0], W = in_shape[1], D = in_shape[2];
  if (H <= 0 || W <= 0 || D <= 0) {
    set_error("in_shape must be positive, got [%d,%d,%d]", H, W, D);
    return CNF_ERR_ARG;
  }
  if (H % 2 || W % 2) {  // M:415-417
    set_error("u/v must have spatial dimensions divisible by 2.");
    return CNF_ERR_ARG;
  }
  if (which_mask < 0 || which_mask > 3) {
    set_error("which_mask must be one of 0,1,2,3, got %d", which_mask);
    return CNF_ERR_ARG;
  }
  if (n_dilations < 1 || n_dilations > CNF_MAX_BRANCHES) {
    set_error("between 1 and %d dilations are supported, got %d", CNF_MAX_BRANCHES, n_dilations);
    return n_dilations < 1 ? CNF_ERR_ARG : CNF_ERR_UNSUPPORTED;
  }
  if (kernel_size < 1 || kernel_size % 2 == 0) {
    set_error("only odd kernel sizes are built (symmetric 'same' padding); got %d", kernel_size);
    return CNF_ERR_UNSUPPORTED;
  }
  if (num_res_blocks < 0 || cardinality < 1 || num_kernels < 1) {
    set_error("num_res_blocks/cardinality/num_kernels out of range");
    return CNF_ERR_ARG;
  }
  auto* c = new cnf_coupling();
  c->H = H; c->W = W; c->D = D;
  c->mask = which_mask;
  static const int comp[4] = {1, 0, 3, 2};  // M:426-433
  c->mask_c = comp[which_mask];
  c->R = num_res_blocks;
  c->card = cardinality;
  c->ks = kernel_size;
  c->ln = layer_norm ? 1 : 0;
  c->nk = which_mask < 2 ? (int)(num_kernels / 2) : num_kernels;  // M:420-423
  if (which_mask < 2) {  // M:474-498
    c->h = H / 2; c->w = W / 2; c->c1 = 2 * D;
  } else {
    c->h = H; c->w = W;
    c->c1 = which_mask == 2 ? (D + 1) / 2 : D / 2;
  }
  if (D % 2 && which_mask == 2) c->c2 = c->c1 - 1;  // M:1093-1104
  else if (D % 2 && which_mask == 3) c->c2 = c->c1 + 1;
  else c->c2 = c->c1;
  if (c->nk < 1 || c->c1 < 1 || c->c2 < 1) {
    set_error("degenerate coupling layer: nk=%d c1=%d c2=%d (mask %d on depth %d)", c->nk, c->c1, c->c2,
              which_mask, D);
    delete c;
    return CNF_ERR_ARG;
  }
  c->dil.assign(which_dilations, which_dilations + n_dilations);

  const int k = c->ks, nk = c->nk;
  const int64_t hw = c->hw();
  LayoutBuilder lb;
  lb.entries = &c->entries;
  c->stem_w = lb.add("stem.kernel", 0, {k, k, c->c1, nk});
  c->stem_b = lb.add("stem.bias", 1, {nk});
  // branch geometry (F:577-588, F:389-411)
  std::vector<Branch> proto;
  int cat = 0;
  for (int d : c->dil) {
    if (d < 1) {
      set_error("dilation factors must be >= 1");
      delete c;
      return CNF_ERR_ARG;
    }
    Branch b;
    b.dil = d;
    b.channels = nk / d;
    if (c->card == 1) {
      b.groups = 1; b.gin = nk; b.gout = b.channels;
    } else {
      if (b.channels % c->card) {  // F:396
        set_error("grouped_convolution: nb_channels (%d = %d // %d) is not divisible by cardinality %d",
                  b.channels, nk, d, c->card);
        delete c;
        return CNF_ERR_ARG;
      }
      b.groups = c->card;
      b.gin = b.gout = b.channels / c->card;
    }
    if (b.channels < 1 || b.gout < 1) {
      set_error("dilation %d leaves no channels (nk=%d, cardinality=%d)", d, nk, c->card);
      delete c;
      return CNF_ERR_ARG;
    }
    b.out_off = cat;
    cat += b.channels;
    proto.push_back(b);
  }
  c->cat = cat;
  char nm[CNF_NAME_CAP];
  for (int r = 0; r < c->R; ++r) {
    ResBlockLayout L;
    auto name = [&](const char* s) { snprintf(nm, sizeof(nm), "rb%d.%s", r, s); return std::string(nm); };
    if (c->ln) {
      L.ln1_g = lb.add(name("ln1.gamma"), 2, {hw * nk});
      L.ln1_b = lb.add(name("ln1.beta"), 3, {hw * nk});
    } else L.ln1_g = L.ln1_b = -1;
    L.pw1_w = lb.add(name("pw1.kernel"), 0, {1, 1, nk, nk});
    L.pw1_b = lb.add(name("pw1.bias"), 1, {nk});
    if (c->ln) {
      L.ln2_g = lb.add(name("ln2.gamma"), 2, {hw * nk});
      L.ln2_b = lb.add(name("ln2.beta"), 3, {hw * nk});
    } else L.ln2_g = L.ln2_b = -1;
    L.br = proto;
    for (auto& b : L.br) {
      const int64_t per = (int64_t)k * k * b.gin * b.gout;
      b.w_off = lb.reserve(per * b.groups);
      b.b_off = lb.reserve(b.channels);
      for (int j = 0; j < b.groups; ++j) {
        snprintf(nm, sizeof(nm), "rb%d.gc.d%d.g%d.kernel", r, b.dil, j);
        lb.add_at(nm, 0, b.w_off + per * j, {k, k, b.gin, b.gout});
        snprintf(nm, sizeof(nm), "rb%d.gc.d%d.g%d.bias", r, b.dil, j);
        lb.add_at(nm, 1, b.b_off + (int64_t)b.gout * j, {b.gout});
      }
    }
    if (c->ln) {
      L.ln3_g = lb.add(name("ln3.gamma"), 2, {hw * cat});
      L.ln3_b = lb.add(name("ln3.beta"), 3, {hw * cat});
    } else L.ln3_g = L.ln3_b = -1;
    L.pw2_w = lb.add(name("pw2.kernel"), 0, {1, 1, cat, nk});
    L.pw2_b = lb.add(name("pw2.bias"), 1, {nk});
    c->rb.push_back(L);
  }
  if (c->ln) {
    c->lnf_g = lb.add("lnf.gamma", 2, {hw * nk});
    c->lnf_b = lb.add("lnf.beta", 3, {hw * nk});
  } else c->lnf_g = c->lnf_b = -1;
  c->head_w = lb.add("head.kernel", 0, {k, k, nk, c->c2});
  c->head_b = lb.add("head.bias", 1, {c->c2});
  c->tanh_w = lb.add("tanh_scale", 4, {1});
  c->net_stride = align4(lb.cur);
  *out = c;
  return CNF_OK;
}

void cnf_coupling_destroy(cnf_coupling* c) { delete c; }

int cnf_coupling_get_info(const cnf_coupling* c, cnf_coupling_info* o) {
  if (!c || !o) {
    set_error("null argument");
    return CNF_ERR_ARG;
  }
  memset(o, 0, sizeof(*o));
  o->H = c->H; o->W = c->W; o->D = c->D;
  o->mask = c->mask; o->mask_complement = c->mask_c;
  o->R = c->R; o->cardinality = c->card; o->nk = c->nk; o->ksize = c->ks; o->layer_norm = c->ln;
  o->h = c->h; o->w = c->w; o->c1 = c->c1; o->c2 = c->c2; o->cat = c->cat;
  o->n_branches = (int)c->dil.size();
  for (int i = 0; i < o->n_branches; ++i) {
    o->dilation[i] = c->dil[i];
    const int ch = c->nk / c->dil[i];
    o->branch_channels[i] = ch;
    o->groups[i] = c->card == 1 ? 1 : c->card;
    o->group_in[i] = c->card == 1 ? c->nk : ch / c->card;
    o->group_out[i] = c->card == 1 ? ch : ch / c->card;
  }
  o->n_ln = c->n_ln();
  o->net_stride = c->net_stride;
  o->param_count = 2 * c->net_stride;
  o->n_entries = (int)c->entries.size();
  return CNF_OK;
}

int cnf_coupling_param_entry(const cnf_coupling* c, int idx, char name[CNF_NAME_CAP], int64_t* offset,
                             int* ndim, int64_t shape[4], int* role) {
  if (!c || idx < 0 || idx >= (int)c->entries.size()) {
    set_error("parameter index %d out of range", idx);
    return CNF_ERR_ARG;
  }
  const ParamEntry& e = c->entries[idx];
  if (name) {
    strncpy(name, e.name.c_str(), CNF_NAME_CAP - 1);
    name[CNF_NAME_CAP - 1] = 0;
  }
  if (offset) *offset = e.offset;
  if (ndim) *ndim = e.ndim;
  if (shape) memcpy(shape, e.shape, sizeof(e.shape));
  if (role) *role = e.role;
  return CNF_OK;
}

int64_t cnf_coupling_workspace_bytes(const cnf_coupling* c, int64_t batch) {
  if (!c || batch < 0) return -1;
  return coupling_ws_bytes(c, batch) + 256 + ((batch * 8 + 255) & ~int64_t(255));
}

int cnf_plan_create(const int io_shape[3], int x_d, int n_blocks, const int* sq, const int* resnext,
                    const int* nkl, const int* cardl, double lambda_y, int ksize, int layer_norm,
                    int dilations, cnf_plan** out) {
  if (!io_shape || !sq || !resnext || !nkl || !cardl || !out) {
    set_error("null argument");
    return CNF_ERR_ARG;
  }
  if (n_blocks < 1 || n_blocks > CNF_MAX_BLOCKS) {
    set_error("between 1 and %d coupling blocks are supported, got %d", CNF_MAX_BLOCKS, n_blocks);
    return CNF_ERR_ARG;
  }
  const int H = io_shape[0], W = io_shape[1], D = io_shape[2];
  if (H <= 0 || W <= 0 || D <= 0 || x_d < 1 || x_d > D) {
    set_error("io_shape [%d,%d,%d] / x_d %d out of range", H, W, D, x_d);
    return CNF_ERR_ARG;
  }
  if (H % 2 || W % 2) {  // M:1464-1466
    set_error("The model input and output must have spatial dimensions divisible by 2.");
    return CNF_ERR_ARG;
  }
  for (int i = 0; i < n_blocks; ++i) {  // M:1468-1484
    if (nkl[i] % 2) {
      set_error("The number of kernels in each layer must be divisible by 2.");
      return CNF_ERR_ARG;
    }
    if (cardl[i] % 2) {
      set_error("The cardinality in each layer must be divisible by 2.");
      return CNF_ERR_ARG;
    }
    if (sq[i] != 0 && sq[i] != 1) {
      set_error("The only allowed entries in squeeze_factor_block_list are 0 and 1.");
      return CNF_ERR_ARG;
    }
  }
  if (!dilations) {
    // M:1553: dilations_list is only defined under `if self.DILATIONS`; the build loop (M:1640-1643)
    // then dereferences it, so the reference cannot construct a model with DILATIONS=False.
    set_error("'cFlow' object has no attribute 'dilations_list' (the reference only defines it when "
              "DILATIONS is True, M:1553)");
    return CNF_ERR_UNSUPPORTED;
  }
  auto* p = new cnf_plan();
  p->H = H; p->W = W; p->D = D; p->x_d = x_d; p->ks = ksize; p->ln = layer_norm ? 1 : 0;
  p->lambda_y = lambda_y;
  p->sq.assign(sq, sq + n_blocks);
  p->resnext.assign(resnext, resnext + n_blocks);
  p->nk.assign(nkl, nkl + n_blocks);
  p->card.assign(cardl, cardl + n_blocks);
  // M:1493-1518
  int npf = 0;
  for (int i = 0; i < n_blocks; ++i) {
    const int s = i == 0 ? 0 : sq[i - 1];
    p->scale.push_back(i == 0 ? 1 : (1 << s) * p->scale.back());
    npf += s;
    p->npf.push_back(npf);
  }
  // M:1521-1536 and M:1553-1617
  for (int i = 0; i < n_blocks; ++i) {
    const int scale = p->scale[i];
    if (H % (scale * 2) || W % (scale * 2)) {
      set_error("The cumulative scale (multiplied by 2 because the checkerboard-masked u/v are halved in "
                "spatial dimensions) must divide evenly into the original i/o spatial dimensions. This "
                "failed at block %d, with i/o shape = (%d, %d) and scale*2 = %d.", i, H, W, scale * 2);
      delete p;
      return CNF_ERR_ARG;
    }
    cnf_block_info b;
    memset(&b, 0, sizeof(b));
    b.scale = scale;
    b.num_prev_factors = p->npf[i];
    b.H = H / scale; b.W = W / scale; b.D = D * scale;
    const double s_ch = b.H < b.W ? b.H : b.W;
    const double s_cb = s_ch / 2.0;
    double d = 1.0, dk = ksize;
    int sanity = 0;
    if (dk > (s_ch + 1) / 2) {
      b.channelwise[b.n_channelwise++] = 1;
      b.checkerboard[b.n_checkerboard++] = 1;
    } else {
      while (dk < (s_ch + 1) / 2) {
        if (sanity >= 10 || b.n_channelwise >= CNF_MAX_BRANCHES) {  // M:1589-1590
          set_error("The dilation while loop ran unexpectedly many iterations.");
          delete p;
          return CNF_ERR_ARG;
        }
        b.channelwise[b.n_channelwise++] = (int)d;
        if (d < (s_cb + 1) / 2) b.checkerboard[b.n_checkerboard++] = (int)d;
        dk = (ksize - 1) * (dk - 1) + 1;
        d = ((dk - ksize) / (ksize - 1)) + 1;
        if (d != std::floor(d)) {
          set_error("non-integer dilation factor %g derived for ksize %d", d, ksize);
          delete p;
          return CNF_ERR_UNSUPPORTED;
        }
        ++sanity;
      }
    }
    if (b.n_checkerboard == 0) {
      // M:1640 would hand an empty list to the residual block (F:576: _which_dilations[0]).
      set_error("no dilation qualifies for the checkerboard layers of block %d (io %dx%d)", i, b.H, b.W);
      delete p;
      return CNF_ERR_ARG;
    }
    const double nkc = (double)nkl[i] / (double)cardl[i];  // M:1613-1617
    for (int j = 0; j < b.n_channelwise; ++j) {
      if (std::fmod(nkc, (double)b.channelwise[j]) != 0.0) {
        set_error("The ratio (number of kernels / cardinality) must be evenly divisible by each dilation "
                  "factor used in that coupling block. This failed in coupling block %d.", i);
        delete p;
        return CNF_ERR_ARG;
      }
    }
    p->blocks.push_back(b);
  }
  // M:1636-1689
  int level = 0;
  int64_t off = 0;
  for (int i = 0; i < n_blocks; ++i) {
    const cnf_block_info& b = p->blocks[i];
    const int shp[3] = {b.H, b.W, b.D};
    for (int m = 0; m < 4; ++m) {
      cnf_coupling* c = nullptr;
      const int* dl = m < 2 ? b.checkerboard : b.channelwise;
      const int nd = m < 2 ? b.n_checkerboard : b.n_channelwise;
      int rc = cnf_coupling_create(shp, m, resnext[i], cardl[i], nkl[i], ksize, layer_norm, dl, nd, &c);
      if (rc != CNF_OK) {
        delete p;
        return rc;
      }
      p->layers.push_back({0, (int)p->couplings.size()});
      p->couplings.push_back(c);
      p->param_off.push_back(off);
      p->level.push_back(level);
      off += 2 * c->net_stride;
    }
    if (sq[i] == 1) {
      p->layers.push_back({1, 0});
      p->layers.push_back({2, p->npf[i]});
      ++level;
    }
  }
  p->param_count = off;
  *out = p;
  return CNF_OK;
}

void cnf_plan_destroy(cnf_plan* p) { delete p; }

int cnf_plan_get_info(const cnf_plan* p, cnf_plan_info* o) {
  if (!p || !o) {
    set_error("null argument");
    return CNF_ERR_ARG;
  }
  memset(o, 0, sizeof(*o));
  o->n_blocks = (int)p->blocks.size();
  o->n_coupling = (int)p->couplings.size();
  o->n_layers = (int)p->layers.size();
  o->H = p->H; o->W = p->W; o->D = p->D; o->x_d = p->x_d; o->ksize = p->ks; o->layer_norm = p->ln;
  o->lambda_y = p->lambda_y;
  o->param_count = p->param_count;
  return CNF_OK;
}

int cnf_plan_block_info(const cnf_plan* p, int block, cnf_block_info* o) {
  if (!p || !o || block < 0 || block >= (int)p->blocks.size()) {
    set_error("block index %d out of range", block);
    return CNF_ERR_ARG;
  }
  *o = p->blocks[block];
  return CNF_OK;
}

int cnf_plan_layer(const cnf_plan* p, int idx, int* kind, int* aux) {
  if (!p || idx < 0 || idx >= (int)p->layers.size()) {
    set_error("layer index %d out of range", idx);
    return CNF_ERR_ARG;
  }
  if (kind) *kind = p->layers[idx].kind;
  if (aux) *aux = p->layers[idx].aux;
  return CNF_OK;
}

const cnf_coupling* cnf_plan_coupling(const cnf_plan* p, int i) {
  if (!p || i < 0 || i >= (int)p->couplings.size()) return nullptr;
  return p->couplings[i];
}

int64_t cnf_plan_coupling_param_offset(const cnf_plan* p, int i) {
  if (!p || i < 0 || i >= (int)p->couplings.size()) return -1;
  return p->param_off[i];
}

int cnf_plan_coupling_level(const cnf_plan* p, int i) {
  if (!p || i < 0 || i >= (int)p->couplings.size()) return -1;
  return p->level[i];
}

int cnf_coupling_set_kernel_paths(cnf_coupling* c, int excluded) {
  if (!c || excluded < 0 || excluded > CNF_PATH_ALL) {
    set_error("cnf_coupling_set_kernel_paths: null descriptor or bits outside CNF_PATH_ALL");
    return CNF_ERR_ARG;
  }
  c->paths = excluded;
  return CNF_OK;
}

int cnf_plan_set_kernel_paths(cnf_plan* p, int excluded) {
  if (!p || excluded < 0 || excluded > CNF_PATH_ALL) {
    set_error("cnf_plan_set_kernel_paths: null plan or bits outside CNF_PATH_ALL");
    return CNF_ERR_ARG;
  }
  for (auto* c : p->couplings) c->paths = excluded;
  return CNF_OK;
}

int cnf_coupling_set_fusion(cnf_coupling* c, int enable) {
  if (!c) {
    set_error("null argument");
    return CNF_ERR_ARG;
  }
  c->paths = enable ? (c->paths & ~CNF_PATH_NO_RESIDENT) : (c->paths | CNF_PATH_NO_RESIDENT);
  return CNF_OK;
}

int cnf_plan_set_fusion(cnf_plan* p, int enable) {
  if (!p) {
    set_error("null argument");
    return CNF_ERR_ARG;
  }
  for (auto* c : p->couplings) cnf_coupling_set_fusion(c, enable);
  return CNF_OK;
}

int64_t cnf_plan_workspace_bytes(const cnf_plan* p, int64_t batch) {
  if (!p || batch < 0) return -1;
  int64_t m = 0;
  for (auto* c : p->couplings) {
    int64_t b = coupling_ws_bytes(c, batch);
    if (b > m) m = b;
  }
  return m + 256 + ((batch * 8 + 255) & ~int64_t(255));  // + per-sample double log-det accumulators
}

}  // extern "C"
