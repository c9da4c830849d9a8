// Execution plan of EfficientUNet.forward + the LCM loop, and the C ABI (include/lcm_unet.h).
//
// The plan is the static op list of one forward for fixed (config, B, H, W): every tensor lives at a
// fixed offset of a caller-provided workspace, so a forward is a pure sequence of kernel launches
// on the caller's stream (no allocation, no host sync).  Op order follows
// /root/reference/src/models/efficient_unet.py:532-606; each block is cut into
//   gn_coef | expand GEMM | gn_coef(+FiLM) | depthwise(+SE pool) | se_gate | project GEMM(+residual)
// so that GroupNorm / FiLM / ReLU6 / SE-scale / residual never touch HBM on their own.
//
// Workspace regions: Z (zeroed at the start of every forward: channel statistics, SE pools, attention
// state), F (small fixed buffers: prologue coefficients, FiLM table, time embedding), A (activations,
// NHWC, offsets assigned by a plan-time first-fit allocator with lifetime-based reuse).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <string>
#include <vector>

#include "../../include/lcm_unet.h"
#include "kernels.h"

using namespace lcm;

namespace lcm {
bool pdl_enabled() {
  static const bool on = !getenv("LCM_NO_PDL");
  return on;
}

namespace {
std::mutex g_dev_mu;
std::map<std::tuple<int, const void*, int>, int> g_dev_slots;      // (device, key, sub) -> cached int
std::map<std::pair<int, const void*>, size_t> g_dev_smem;          // (device, function) -> dynamic smem opted in
}  // namespace

int ensure_dyn_smem(const void* fn, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -2;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  size_t& have = g_dev_smem[std::make_pair(dev, fn)];
  if (have >= bytes) return 0;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return -2;
  have = bytes;
  return 0;
}

int* device_cache_slot(const void* key, int sub) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  return &g_dev_slots[std::make_tuple(dev, key, sub)];   // std::map nodes are address-stable
}
}  // namespace lcm

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(x)                                                                         \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) return fail(LCM_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

struct Tensor {
  size_t off = 0;        // byte offset in region A
  size_t bytes = 0;
  int C = 0, H = 0, W = 0;
  bool f16 = false;      // stored as fp16 (hidden tensors of a block on the tcgen05 path), else the plan's type
  size_t stats_off = 0;  // byte offset of double[N][C][2] in region Z
  int refs = 0;
  // training plans: gradient buffer in region G (allocated by the first backward op that writes it)
  bool g_alloc = false, g_written = false;
  size_t g_off = 0, g_bytes = 0;
  std::string tap;
};
typedef std::shared_ptr<Tensor> TensorP;

struct RunCtx {
  char* z; char* f; char* a;   // region bases
  char* zb; char* fb; char* g; float* wg;   // training: zeroed backward scratch, backward coefficient tables, gradients, flat weight gradients
  const float* target; int loss_type; float gscale; const float* gscale_dev;   // training: loss definition
  const float* xa; int ca; long long sa;
  const float* xb; int cb; long long sb;
  const long long* t_dev; long long t_scalar;
  float* eps;
  FinalStep step;
  int* launch_err;   // set by ops whose launch wrapper rejects the shape
};

typedef std::function<void(const RunCtx&, cudaStream_t)> RunFn;
struct Op {
  std::string name, kernel;
  double bytes = 0, flops = 0;
  double ref_bytes = 0;   // SURVEY accounting of the reference op(s) this op stands for (= bytes unless a fusion moves less)
  int launches = 1;   // kernels this op enqueues
  RunFn run;
};

struct WeightSlot {
  int64_t numel = 0;
  std::vector<PackJob> jobs;
  bool set = false;
};

// first-fit free-list allocator over region A (plan-time only)
struct PoolAlloc {
  struct Blk { size_t off, size; };
  std::vector<Blk> free_list;
  size_t top = 0;
  bool reuse = true;
  size_t alloc(size_t bytes) {
    bytes = align_up(bytes, 1024);
    if (reuse) {
      for (size_t i = 0; i < free_list.size(); ++i) {
        if (free_list[i].size >= bytes) {
          size_t off = free_list[i].off;
          if (free_list[i].size == bytes) free_list.erase(free_list.begin() + i);
          else { free_list[i].off += bytes; free_list[i].size -= bytes; }
          return off;
        }
      }
      if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
        size_t off = free_list.back().off;   // grow the trailing free block
        top = off + bytes;
        free_list.pop_back();
        return off;
      }
    }
    size_t off = top;
    top += bytes;
    return off;
  }
  void release(size_t off, size_t bytes) {
    if (!reuse) return;
    bytes = align_up(bytes, 1024);
    size_t i = 0;
    while (i < free_list.size() && free_list[i].off < off) ++i;
    free_list.insert(free_list.begin() + i, Blk{off, bytes});
    for (size_t j = 0; j + 1 < free_list.size();) {
      if (free_list[j].off + free_list[j].size == free_list[j + 1].off) {
        free_list[j].size += free_list[j + 1].size;
        free_list.erase(free_list.begin() + j + 1);
      } else ++j;
    }
  }
};

}  // namespace

struct lcm_plan {
  lcm_unet_config cfg;
  int N = 0, H = 0, W = 0, prec = 0, device = 0;
  uint32_t flags = 0;
  bool bf16 = false, tc = false, taps = false;
  int num_sms = 148;
  size_t esz = 4;  // activation element size

  size_t z_bytes = 0, f_bytes = 0;
  PoolAlloc pool;
  size_t ws_bytes = 0;

  char* wbase = nullptr;   // packed weights (owned)
  size_t wbytes = 0;
  std::vector<PackJob> identity_jobs;
  std::map<std::string, WeightSlot> weights;
  std::vector<std::string> weight_order;

  int film_rows = 0;
  size_t film_f_off = 0, silu_f_off = 0;
  size_t addtmp_f_off = 0;   // condition_mode="add" (in_channels == out_channels): latents + condition features, fp32 NCHW

  std::vector<Op> ops;
  // ---- training (LCM_FLAG_TRAIN) -------------------------------------------------------------------
  bool train = false;
  int dt_act = 0, dt_hid = 0, dt_grad = 0;     // DType of the residual stream / a block's hidden tensors / gradients
  size_t gsz = 4;                              // gradient element size
  PoolAlloc gpool;                             // region G
  size_t zb_bytes = 0, fb_bytes = 0;
  size_t wg_elems = 0;                         // flat fp32 weight-gradient buffer (state_dict layouts, weight_order order)
  std::map<std::string, size_t> wgrad_off;     // element offset of every weight's gradient
  std::map<std::string, int> wgrad_last_op;    // index of the last backward op that writes it (bucketed all-reduce)
  std::vector<Op> bwd_ops;
  size_t dfilm_fb_off = 0, dst_zb_off = 0;     // d FiLM table [N][film_rows] (FB), d silu(t_emb) [N][ted] (ZB)
  size_t zero_bias_off = 0;
  size_t zb_off_ws = 0, fb_off_ws = 0, g_off_ws = 0, wg_off_ws = 0;   // region offsets inside the workspace
  std::map<std::string, TensorP> gtap_map;
  size_t zballoc(size_t bytes) { size_t o = zb_bytes; zb_bytes += align_up(bytes, 256); return o; }
  size_t fballoc(size_t bytes) { size_t o = fb_bytes; fb_bytes += align_up(bytes, 256); return o; }
  std::map<std::string, TensorP> tap_map;
  std::vector<std::string> tap_order;
  double total_bytes = 0, total_flops = 0;
  double total_bytes_ref = 0;   // the same sum with every op accounted as the reference's op sequence moves it

  size_t walloc(size_t bytes) { size_t o = wbytes; wbytes += align_up(bytes, 256); return o; }
  size_t zalloc(size_t bytes) { size_t o = z_bytes; z_bytes += align_up(bytes, 256); return o; }
  size_t falloc(size_t bytes) { size_t o = f_bytes; f_bytes += align_up(bytes, 256); return o; }

  int groups(int C) const {
    int g = C < 32 ? C : 32;
    if (C % g == 0) return g;
    return cfg.groupnorm_gcd ? gcd_i(32, C) : -1;
  }
  TensorP new_tensor(int C, int h, int w, bool stats, const std::string& tap, bool f16 = false) {
    TensorP t = std::make_shared<Tensor>();
    t->C = C; t->H = h; t->W = w; t->f16 = f16;
    t->bytes = (size_t)N * h * w * C * esz;
    t->off = pool.alloc(t->bytes);
    t->refs = 1;
    t->tap = tap;
    if (stats) t->stats_off = zalloc((size_t)N * C * 2 * sizeof(double));
    if (taps && !tap.empty()) { tap_map[tap] = t; tap_order.push_back(tap); }
    return t;
  }
  void release(const TensorP& t) { if (--t->refs == 0) pool.release(t->off, t->bytes); }

  void add_weight(const std::string& name, int64_t numel, const PackJob& job) {
    if (weights.find(name) == weights.end()) {
      weight_order.push_back(name);
      weights[name].numel = numel;
      wgrad_off[name] = wg_elems;
      wg_elems += (size_t)(numel + 3) / 4 * 4;    // 16-byte aligned slices
    }
    weights[name].jobs.push_back(job);
  }
  size_t add_copy(const std::string& name, int64_t numel) {   // verbatim fp32 copy; returns weight-arena offset
    size_t off = walloc(numel * sizeof(float));
    PackJob j{}; j.kind = PACK_COPY; j.dst = (void*)off; j.R = 1; j.Cc = (int)numel;
    add_weight(name, numel, j);
    return off;
  }
  const float* wf(size_t off) const { return (const float*)(wbase + off); }
};

namespace {

// A GEMM weight: logical [Nc][sum K_s]; in the tcgen05 image every segment is padded to 64-wide chunks.
struct GemmW {
  size_t off = 0;
  int Nc = 0, Ktot = 0, Kpad = 0, block_n = 0;
  bool expand = false;   // served by gemm_expand.cu: block_n = 128
  bool wide = false;     // served by gemm_wide.cu: block_n = 128
  bool relu6 = false;    // tcgen05 path, relu6 prologue as 6 sat(.): weights scaled by 6
  std::vector<int> seg_off, seg_pad_off;
};

struct View {
  TensorP part[2];
  int n = 0;
  int C() const { return (n > 0 ? part[0]->C : 0) + (n > 1 ? part[1]->C : 0); }
  static View of(const TensorP& t) { View v; v.part[0] = t; v.n = 1; return v; }
};

struct Builder {
  lcm_plan* p;
  int N;
  struct FilmBlock { std::string wname; int row0, rows; };
  std::vector<FilmBlock> film_blocks;
  explicit Builder(lcm_plan* plan) : p(plan), N(plan->N) {}

  // ---- training tape -------------------------------------------------------------------------------
  // Every forward builder registers a closure that, run in REVERSE registration order once the forward plan is
  // complete, emits the backward ops of its forward op and allocates / releases gradient buffers in region G with the
  // same first-fit pool as the activations (lifetimes follow the backward execution order).
  std::vector<std::function<void()>> tape;
  void on_backward(std::function<void()> fn) { if (p->train) tape.push_back(std::move(fn)); }

  void pushb(const std::string& name, const char* kernel, const std::vector<std::string>& writes, RunFn fn) {
    Op o; o.name = name; o.kernel = kernel; o.run = std::move(fn);
    for (const std::string& w : writes) p->wgrad_last_op[w] = (int)p->bwd_ops.size();
    p->bwd_ops.push_back(std::move(o));
  }
  size_t galloc(size_t bytes) { return p->gpool.alloc(bytes); }
  void gfree(size_t off, size_t bytes) { p->gpool.release(off, bytes); }
  // gradient of a forward tensor: the first writer stores, later writers accumulate
  struct GradW { size_t off; int accumulate; };
  GradW grad_w(const TensorP& t) {
    if (!t->g_alloc) {
      t->g_bytes = (size_t)N * t->H * t->W * t->C * p->gsz;
      t->g_off = galloc(t->g_bytes);
      t->g_alloc = true;
      if (p->taps && !t->tap.empty()) p->gtap_map[t->tap] = t;
    }
    GradW g{t->g_off, t->g_written ? 1 : 0};
    t->g_written = true;
    return g;
  }
  size_t grad_r(const TensorP& t) { return t->g_off; }   // valid when has_grad(t)
  bool has_grad(const TensorP& t) const { return t->g_alloc && t->g_written; }
  void grad_done(const TensorP& t) { if (t->g_alloc) gfree(t->g_off, t->g_bytes); }
  size_t wg(const std::string& name) const { return p->wgrad_off.at(name); }

  // dgrad weights of a 1x1 conv: W'[k_in][n_out] = W[n_out][k_in]  (logical [Nc = rows of W'][K = cols])
  struct DgradW { size_t off; int Nc, K, Kpad, block_n; };
  DgradW make_dgrad_w(const std::string& wname, int64_t numel, int rows_out /*Nc'*/, int K /*cols'*/, int src_ld, int src_col0) {
    DgradW d{};
    d.Nc = rows_out; d.K = K; d.Kpad = (K + 63) / 64 * 64;
    PackJob j{};
    j.kind = PACK_MAT_T;
    j.R = rows_out; j.Cc = K; j.src_ld = src_ld; j.src_col0 = src_col0;
    if (p->tc) {
      d.block_n = gemm_tc_pick_block_n(rows_out);
      d.off = p->walloc((size_t)rows_out * d.Kpad * sizeof(bf16));
      j.layout = WL_UMMA; j.bf16 = 1; j.ld = d.Kpad; j.block_n = d.block_n;
    } else {
      d.off = p->walloc((size_t)rows_out * K * p->esz);
      j.layout = WL_ROWMAJOR; j.bf16 = p->bf16 ? 1 : 0; j.ld = K;
    }
    j.dst = (void*)d.off;
    p->add_weight(wname, numel, j);
    return d;
  }
  // out[M][Nc'] = dY[M][K] . W'^T  (input gradient of a 1x1 conv); all buffers in region G
  // stats_zb (optional, tensor-core plan): offset in region ZB of a zeroed [N][Nc][2] fp64 table that receives the
  // per-(image, channel) sum / sum^2 of the output (the GEMM's statistics epilogue)
  void dgrad_gemm(const std::string& name, size_t dy_off, const DgradW& w, size_t out_off, int P, size_t stats_zb = (size_t)-1) {
    lcm_plan* pl = p; const int n = N;
    pushb(name, pl->tc ? "gemm_tc" : "gemm_simt", {}, [=](const RunCtx& c, cudaStream_t st) {
      GemmParams gp{};
      gp.nseg = 1;
      gp.seg[0].A = c.g + dy_off; gp.seg[0].K = w.K; gp.seg[0].ld = w.K; gp.seg[0].mode = XF_NONE;
      gp.Ktot = w.K; gp.W = pl->wbase + w.off; gp.out = c.g + out_off;
      gp.stats = (pl->tc && stats_zb != (size_t)-1) ? (double*)(c.zb + stats_zb) : nullptr;
      gp.P = P; gp.M = (long long)n * P; gp.Nc = w.Nc;
      if (pl->tc) { ConvGeom g{}; g.mode = -1; if (launch_gemm_tc(gp, g, w.block_n, pl->num_sms, st)) *c.launch_err = 1; }
      else launch_gemm_simt(gp, pl->bf16, st);
    });
  }
  // GroupNorm backward finalise for a view x (one or two tensors): (T1, T2) in ZB -> coef4 in FB, d gamma / d beta
  struct GnInfo { size_t coef, g_off, b_off; int G; double count; std::string wname; };
  // ascale_fb (optional): offset in region FB of a float4 table whose .z is multiplied into A (see gn_bwd_coef_kernel)
  size_t gn_backward(const std::string& name, const View& x, const GnInfo& gi, size_t t12, int film_row0, size_t ascale_fb = (size_t)-1) {
    const int C = x.C();
    const size_t coef4 = p->fballoc((size_t)N * C * sizeof(float4));
    const size_t s0 = x.part[0]->stats_off; const int C0 = x.part[0]->C;
    const size_t s1 = x.n > 1 ? x.part[1]->stats_off : 0; const int C1 = x.n > 1 ? x.part[1]->C : 0;
    lcm_plan* pl = p; const int n = N;
    const size_t dg = wg(gi.wname + ".weight"), db = wg(gi.wname + ".bias");
    pushb(name, "gn_bwd_coef", {gi.wname + ".weight", gi.wname + ".bias"}, [=](const RunCtx& c, cudaStream_t st) {
      const float* film = film_row0 >= 0 ? (const float*)(c.f + pl->film_f_off) + film_row0 : nullptr;
      float* dfilm = film_row0 >= 0 ? (float*)(c.fb + pl->dfilm_fb_off) + film_row0 : nullptr;
      launch_gn_bwd_coef((const double*)(c.zb + t12), (const double*)(c.z + s0), C0, C1 ? (const double*)(c.z + s1) : nullptr, C1,
                         gi.G, gi.count, pl->wf(gi.g_off), pl->wf(gi.b_off), film, dfilm, pl->film_rows,
                         (float4*)(c.fb + coef4), c.wg + dg, c.wg + db, n, st,
                         ascale_fb != (size_t)-1 ? (const float*)(c.fb + ascale_fb) + 2 : nullptr, 4);
    });
    return coef4;
  }

  // ref_bytes: what the reference's op sequence moves for this op where a fusion of ours moves less (default: the same)
  void push(const std::string& name, const char* kernel, double bytes, double flops, RunFn fn, double ref_bytes = -1.0) {
    Op o; o.name = name; o.kernel = kernel; o.bytes = bytes; o.flops = flops; o.run = std::move(fn);
    o.ref_bytes = ref_bytes >= 0.0 ? ref_bytes : bytes;
    p->total_bytes += bytes; p->total_flops += flops;
    p->total_bytes_ref += o.ref_bytes;
    p->ops.push_back(std::move(o));
  }

  // fused_layout: the weights of the fused expand -> depthwise kernel (n-blocks of 128 rows, the last one zero-padded)
  GemmW make_w(int Nc, const std::vector<int>& segK, int expand_P = 0, bool fused_layout = false) {
    GemmW g;
    g.Nc = Nc;
    g.expand = p->tc && expand_P > 0 && gemm_expand_supported((int)segK.size(), segK.data(), Nc, expand_P);
    g.wide = p->tc && expand_P > 0 && !g.expand && gemm_wide_supported((int)segK.size(), segK.data(), Nc, expand_P);
    g.relu6 = p->tc && expand_P > 0;
    for (int k : segK) {
      g.seg_off.push_back(g.Ktot);
      g.seg_pad_off.push_back(g.Kpad);
      g.Ktot += k;
      g.Kpad += (k + 63) / 64 * 64;
    }
    if (p->tc) {
      g.block_n = (g.expand || g.wide || fused_layout) ? 128 : gemm_tc_pick_block_n(Nc);
      g.off = p->walloc((size_t)(fused_layout ? (Nc + 127) / 128 * 128 : Nc) * g.Kpad * sizeof(bf16));
    }
    else g.off = p->walloc((size_t)Nc * g.Ktot * p->esz);
    return g;
  }
  // job that writes columns [src_col0, src_col0+Cc) of a [R][src_ld] source into segment `seg`
  PackJob mat_job(const GemmW& g, int seg, int kind, int R, int Cc, int src_ld, int src_col0, bool f16 = false) {
    PackJob j{};
    j.kind = kind;
    j.layout = p->tc ? WL_UMMA : WL_ROWMAJOR;
    j.bf16 = f16 ? 2 : (p->bf16 ? 1 : 0);   // weights of an fp16 activation segment are fp16 too
    j.dst = (void*)g.off;
    j.R = R; j.Cc = Cc; j.src_ld = src_ld; j.src_col0 = src_col0;
    j.ld = p->tc ? g.Kpad : g.Ktot;
    j.off = p->tc ? g.seg_pad_off[seg] : g.seg_off[seg];
    j.block_n = g.block_n;
    j.scale = g.relu6 ? 6.f : 0.f;
    return j;
  }

  // GroupNorm finalise -> coef buffer [N][C] in region F
  GnInfo gn_coef(const std::string& name, const View& x, const std::string& wname, int film_row0) {
    const int C = x.C(), G = p->groups(C);
    const size_t coef = p->falloc((size_t)N * C * sizeof(float2));
    const size_t g_off = p->add_copy(wname + ".weight", C), b_off = p->add_copy(wname + ".bias", C);
    const size_t s0 = x.part[0]->stats_off; const int C0 = x.part[0]->C;
    const size_t s1 = x.n > 1 ? x.part[1]->stats_off : 0; const int C1 = x.n > 1 ? x.part[1]->C : 0;
    const double count = (double)x.part[0]->H * x.part[0]->W * (C / G);
    lcm_plan* pl = p; const int n = N;
    push(name, "gn_coef", 0, 0, [=](const RunCtx& c, cudaStream_t st) {
      const float* film = film_row0 >= 0 ? (const float*)(c.f + pl->film_f_off) + film_row0 : nullptr;
      launch_gn_coef((const double*)(c.z + s0), C0, C1 ? (const double*)(c.z + s1) : nullptr, C1, G, count,
                     pl->wf(g_off), pl->wf(b_off), film, pl->film_rows, (float2*)(c.f + coef), n, st);
    });
    return GnInfo{coef, g_off, b_off, G, count, wname};
  }

  struct SegSpec { TensorP t; size_t coef; int coef_ld, coef_off, mode; };

  void gemm(const std::string& name, const std::vector<SegSpec>& segs, const GemmW& w, const TensorP& out, bool stats,
            double bytes, double flops) {
    lcm_plan* pl = p; const int n = N;
    std::vector<SegSpec> sg = segs;
    // Gram / column-sum scratch of the expand kernel: a slice of the region zeroed at the start of the forward
    const size_t xs = (pl->tc && w.expand) ? p->zalloc(gemm_expand_scratch_bytes(N)) : 0;
    push(name, pl->tc ? (w.expand ? "gemm_expand" : (w.wide ? "gemm_wide" : "gemm_tc")) : "gemm_simt", bytes, flops, [=](const RunCtx& c, cudaStream_t st) {
      GemmParams gp{};
      gp.nseg = (int)sg.size();
      for (int i = 0; i < gp.nseg; ++i) {
        gp.seg[i].A = c.a + sg[i].t->off;
        gp.seg[i].K = sg[i].t->C;
        gp.seg[i].ld = sg[i].t->C;
        gp.seg[i].coef = sg[i].mode == XF_NONE ? nullptr : (const float2*)(c.f + sg[i].coef);
        gp.seg[i].coef_ld = sg[i].coef_ld;
        gp.seg[i].coef_off = sg[i].coef_off;
        gp.seg[i].mode = sg[i].mode;
        gp.seg[i].f16 = sg[i].t->f16 ? 1 : 0;
      }
      gp.Ktot = w.Ktot;
      gp.W = pl->wbase + w.off;
      gp.out = c.a + out->off;
      gp.stats = stats ? (double*)(c.z + out->stats_off) : nullptr;
      gp.P = out->H * out->W;
      gp.M = (long long)n * gp.P;
      gp.Nc = w.Nc;
      gp.out_f16 = out->f16 ? 1 : 0;
      if (pl->tc && w.expand) { if (launch_gemm_expand(gp, c.z + xs, false, pl->num_sms, st)) *c.launch_err = 1; }
      else if (pl->tc && w.wide) { if (launch_gemm_wide(gp, pl->num_sms, st)) *c.launch_err = 1; }
      else if (pl->tc) { ConvGeom g{}; g.mode = -1; if (launch_gemm_tc(gp, g, w.block_n, pl->num_sms, st)) *c.launch_err = 1; }
      else launch_gemm_simt(gp, pl->bf16, st);
    });
  }

  // ---- InvertedResidualBlock (efficient_unet.py:203-236) -------------------------------------
  // se_ratio: the reference passes config.se_ratio to the encoder/decoder blocks only; mid_block1/2 are built with the
  // constructor default 0.25 (efficient_unet.py:467-478 vs :440,502)
  TensorP block(const std::string& name, const View& x, int Co, float se_ratio) {
    const int Ci = x.C(), Ch = Ci * p->cfg.expansion_ratio;
    int SQ = (int)(Ch * se_ratio);
    if (SQ < 1) SQ = 1;
    const int h = x.part[0]->H, w = x.part[0]->W;
    const double P = (double)h * w, es = (double)p->esz;
    const int row0 = p->film_rows;   // this block's [scale | shift] rows in the fused FiLM GEMV
    p->film_rows += 2 * Ch;
    film_blocks.push_back({name + ".time_mlp.1", row0, 2 * Ch});

    // norm1 -> ReLU6 -> expand (:207-209)
    const GnInfo gn1 = gn_coef(name + ".norm1", x, name + ".norm1", -1);
    const size_t coef1 = gn1.coef;
    // tcgen05 path: both hidden tensors are fp16 (dwconv_stream.cu explains why)
    const bool hid16 = p->tc;
    TensorP h1 = p->new_tensor(Ch, h, w, true, name + ".expand", hid16);
    bool fused = false;
    GemmW xw;
    TensorP xt;     // fused path: t = relu6(GN1(x)) / 6, written by the statistics pass
    {
      std::vector<int> segK;
      for (int i = 0; i < x.n; ++i) segK.push_back(x.part[i]->C);
      GemmW we = make_w(Ch, segK, h * w);
      std::vector<SegSpec> segs;
      int col = 0;
      for (int i = 0; i < x.n; ++i) {
        segs.push_back({x.part[i], coef1, Ci, col, XF_AFFINE_RELU6});
        p->add_weight(name + ".expand.weight", (int64_t)Ch * Ci, mat_job(we, i, PACK_MAT, Ch, x.part[i]->C, Ci, col));
        col += x.part[i]->C;
      }
      // Inference plans fuse expand -> norm2 / FiLM / ReLU6 -> depthwise into ONE kernel where the expand kernel applies
      // (xdw_fused.cu): h1 is never materialised; its statistics come from a pass over the block input alone.
      // (hidden widths of 128 m + 64 channels — the Base variant's 192 — are not the expand kernel's, but the fused kernel masks
      //  the half-empty last n-block)
      fused = p->tc && !p->train && !p->taps && (we.expand || (Ch % 128 == 64 && Ci <= 128)) &&
              xdw_fused_supported((int)segK.size(), segK.data(), Ch, h, w);
      if (fused) {
        // the weights once more, packed as ONE dense K segment (t concatenates the input parts)
        xw = make_w(Ch, std::vector<int>{Ci}, h * w, true);
        p->add_weight(name + ".expand.weight", (int64_t)Ch * Ci, mat_job(xw, 0, PACK_MAT, Ch, Ci, Ci, 0));
        xt = p->new_tensor(Ci, h, w, false, "", false);
        lcm_plan* pl = p; const int n = N;
        const size_t xs = p->zalloc(gemm_expand_scratch_bytes(N));
        std::vector<SegSpec> sg = segs;
        const GemmW wv = xw;
        const TensorP tt = xt;
        push(name + ".expand.stats", "xstats", 2.0 * Ci * N * P * es, 2.0 * N * P * Ci * Ci, [=](const RunCtx& c, cudaStream_t st) {
          GemmParams gp{};
          gp.nseg = (int)sg.size();
          for (int i = 0; i < gp.nseg; ++i) {
            gp.seg[i].A = c.a + sg[i].t->off; gp.seg[i].K = sg[i].t->C; gp.seg[i].ld = sg[i].t->C;
            gp.seg[i].coef = (const float2*)(c.f + sg[i].coef); gp.seg[i].coef_ld = sg[i].coef_ld; gp.seg[i].coef_off = sg[i].coef_off;
            gp.seg[i].mode = sg[i].mode; gp.seg[i].f16 = 0;
          }
          gp.P = h * w; gp.M = (long long)n * gp.P;
          if (launch_xstats(gp, c.a + tt->off, c.z + xs, pl->num_sms, st) ||
              launch_expand_stats_finalize(c.z + xs, pl->wbase + wv.off, (double*)(c.z + h1->stats_off), n, wv.Nc, (wv.Ktot + 63) / 64, st, wv.Ktot))
            *c.launch_err = 1;
        }, (Ci + Ch) * N * P * es + (double)Ci * Ch * es);
        p->ops.back().launches = 2;
      } else {
      gemm(name + ".expand", segs, we, h1, true, (Ci + Ch) * N * P * es + (double)Ci * Ch * es, 2.0 * N * P * Ci * Ch);
      if (we.expand) p->ops.back().launches = 2;   // GEMM + statistics finalisation
      }
    }
    // norm2 + FiLM + ReLU6 -> depthwise (:212-220), SE pool (:97)
    const GnInfo gn2 = gn_coef(name + ".norm2", View::of(h1), name + ".norm2", row0);
    const size_t coef2 = gn2.coef;
    TensorP h2 = p->new_tensor(Ch, h, w, false, name + ".depthwise", hid16);
    const size_t pool = p->zalloc((size_t)N * Ch * sizeof(double));
    const size_t dw_off = p->walloc((size_t)9 * Ch * sizeof(float));
    { PackJob j{}; j.kind = PACK_DW; j.dst = (void*)dw_off; j.R = Ch; p->add_weight(name + ".depthwise.weight", (int64_t)Ch * 9, j); }
    if (fused) {
      lcm_plan* pl = p; const int n = N;
      const GemmW wv = xw;
      const TensorP tt = xt;
      push(name + ".expand_depthwise", "xdw_fused", ((double)Ci + Ch) * N * P * es + (double)Ci * Ch * es + 36.0 * Ch,
           2.0 * N * P * Ci * Ch + 18.0 * N * P * Ch, [=](const RunCtx& c, cudaStream_t st) {
        if (launch_xdw_fused(c.a + tt->off, wv.Ktot, pl->wbase + wv.off, wv.Nc, (const float2*)(c.f + coef2), pl->wf(dw_off),
                             c.a + h2->off, (double*)(c.z + pool), n, h, w, pl->num_sms, st)) *c.launch_err = 1;
      }, 2.0 * Ch * N * P * es + 36.0 * Ch);
      p->release(xt);
    } else
    {
      lcm_plan* pl = p; const int n = N;
      push(name + ".depthwise", hid16 ? "dwconv_stream" : "dwconv", 2.0 * Ch * N * P * es + 36.0 * Ch, 18.0 * N * P * Ch,
           [=](const RunCtx& c, cudaStream_t st) {
             if (hid16) {
               if (launch_dwconv_f16(c.a + h1->off, (const float2*)(c.f + coef2), pl->wf(dw_off), c.a + h2->off,
                                     (double*)(c.z + pool), n, h, w, Ch, pl->num_sms, st)) *c.launch_err = 1;
             } else {
               launch_dwconv(c.a + h1->off, (const float2*)(c.f + coef2), pl->wf(dw_off), c.a + h2->off,
                             (double*)(c.z + pool), n, h, w, Ch, pl->bf16, 0, st);
             }
           });
    }
    p->release(h1);
    // SE gate (:98-99)
    const size_t gate = p->falloc((size_t)N * Ch * sizeof(float2));
    const size_t w1 = p->add_copy(name + ".se.fc1.weight", (int64_t)SQ * Ch), b1 = p->add_copy(name + ".se.fc1.bias", SQ);
    const size_t w2 = p->add_copy(name + ".se.fc2.weight", (int64_t)Ch * SQ), b2 = p->add_copy(name + ".se.fc2.bias", Ch);
    {
      lcm_plan* pl = p; const int n = N;
      push(name + ".se", "se_gate", 2.0 * Ch * SQ * 4 + 8.0 * N * Ch, 4.0 * N * Ch * SQ,
           [=](const RunCtx& c, cudaStream_t st) {
             if (launch_se_gate((const double*)(c.z + pool), (float)(1.0 / P), pl->wf(w1), pl->wf(b1), pl->wf(w2), pl->wf(b2),
                                (float2*)(c.f + gate), n, Ch, SQ, st)) *c.launch_err = 1;
           });
    }
    // SE scale -> project -> + (skip conv | identity)(x)  (:100,226,230-234) as ONE GEMM over [h2 | x]
    TensorP out = p->new_tensor(Co, h, w, true, name + ".out");
    bool pstream = false;
    if (p->tc && hid16 && x.n == 1) {
      const int sk[2] = {Ch, Ci}, sf[2] = {1, 0}, sm[2] = {XF_SCALE, XF_NONE};
      pstream = proj_stream_supported(2, sk, sf, sm, Co, h * w);
    }
    if (pstream) {
      // level-0 shape (K = 128 + 32 -> N = 32): the streaming kernel (proj_stream.cu), weights row-major [Co][Ch + Ci] 16-bit
      const size_t woff = p->walloc((size_t)Co * (Ch + Ci) * sizeof(bf16));
      PackJob j{};
      j.kind = PACK_MAT; j.layout = WL_ROWMAJOR; j.bf16 = 2; j.dst = (void*)woff; j.R = Co; j.Cc = Ch; j.src_ld = Ch; j.src_col0 = 0;
      j.ld = Ch + Ci; j.off = 0;
      p->add_weight(name + ".project.weight", (int64_t)Co * Ch, j);
      PackJob k = j;
      k.bf16 = 1; k.Cc = Ci; k.src_ld = Ci; k.off = Ch;
      if (Ci != Co) p->add_weight(name + ".skip.weight", (int64_t)Co * Ci, k);
      else { k.kind = PACK_IDENTITY; p->identity_jobs.push_back(k); }
      lcm_plan* pl = p; const int n = N;
      const TensorP xin = x.part[0];
      const double bytes = (Ch + Ci + Co) * N * P * es + (double)Ch * Co * es + (Ci != Co ? (double)Ci * Co * es : 0.0);
      const double flops = 2.0 * N * P * Ch * Co + (Ci != Co ? 2.0 * N * P * Ci * Co : 0.0);
      push(name + ".project", "proj_stream", bytes, flops, [=](const RunCtx& c, cudaStream_t st) {
        GemmParams gp{};
        gp.nseg = 2;
        gp.seg[0].A = c.a + h2->off; gp.seg[0].K = Ch; gp.seg[0].ld = Ch; gp.seg[0].mode = XF_SCALE; gp.seg[0].f16 = 1;
        gp.seg[0].coef = (const float2*)(c.f + gate); gp.seg[0].coef_ld = Ch; gp.seg[0].coef_off = 0;
        gp.seg[1].A = c.a + xin->off; gp.seg[1].K = Ci; gp.seg[1].ld = Ci; gp.seg[1].mode = XF_NONE; gp.seg[1].f16 = 0;
        gp.Ktot = Ch + Ci; gp.W = pl->wbase + woff; gp.out = c.a + out->off; gp.stats = (double*)(c.z + out->stats_off);
        gp.P = h * w; gp.M = (long long)n * gp.P; gp.Nc = Co; gp.out_f16 = 0;
        if (launch_proj_stream(gp, pl->num_sms, st)) *c.launch_err = 1;
      });
    } else {
      std::vector<int> pk;
      pk.push_back(Ch);
      for (int i = 0; i < x.n; ++i) pk.push_back(x.part[i]->C);
      GemmW wp = make_w(Co, pk);
      p->add_weight(name + ".project.weight", (int64_t)Co * Ch, mat_job(wp, 0, PACK_MAT, Co, Ch, Ch, 0, hid16));
      std::vector<SegSpec> segs;
      segs.push_back({h2, gate, Ch, 0, XF_SCALE});
      int col = 0;
      for (int i = 0; i < x.n; ++i) {
        segs.push_back({x.part[i], 0, 0, 0, XF_NONE});
        if (Ci != Co) p->add_weight(name + ".skip.weight", (int64_t)Co * Ci, mat_job(wp, 1 + i, PACK_MAT, Co, x.part[i]->C, Ci, col));
        else p->identity_jobs.push_back(mat_job(wp, 1 + i, PACK_IDENTITY, Co, x.part[i]->C, Ci, col));
        col += x.part[i]->C;
      }
      const double bytes = (Ch + Ci + Co) * N * P * es + (double)Ch * Co * es + (Ci != Co ? (double)Ci * Co * es : 0.0);
      const double flops = 2.0 * N * P * Ch * Co + (Ci != Co ? 2.0 * N * P * Ci * Co : 0.0);
      gemm(name + ".project", segs, wp, out, true, bytes, flops);
    }
    if (p->train) block_backward(name, x, h1, h2, out, gn1, gn2, gate, pool, dw_off, w1, b1, w2, SQ, row0, Co);
    p->release(h2);
    for (int i = 0; i < x.n; ++i) p->release(x.part[i]);
    return out;
  }

  // ---- backward of one InvertedResidualBlock (reverse of efficient_unet.py:203-236) -----------------------
  //   dq  = dY Wp                      d(gate * h2)                 [project dgrad]
  //   dWp = dY^T (gate h2), dWskip = dY^T x                         [wgrad]
  //   dgate = sum_p dq h2 -> SE FC backward -> (gate, dpool / P)    [mask_reduce, se_bwd_vec, outer_sum x2]
  //   du  = dwconv^T(gate dq + dpool / P) [0 < u < 6], dWdw, T1/T2  [dwconv_bwd]
  //   (A, B, C) of norm2 + FiLM, d gamma2 / d beta2 / d FiLM        [gn_bwd_coef]
  //   dh1 = A du + B h1 + C                                         [affine3, in place]
  //   dr  = dh1 We ; dWe = dh1^T relu6(a1 x + b1)                   [expand dgrad, wgrad]
  //   dpre = dr [0 < a1 x + b1 < 6], T1/T2                          [mask_reduce mode 1, in place]
  //   (A, B, C) of norm1                                            [gn_bwd_coef]
  //   dx (+)= A dpre + B x + C + (dY | dY Wskip)                    [affine3 per input part]
  void block_backward(const std::string& name, const View& x, const TensorP& h1, const TensorP& h2, const TensorP& out,
                      const GnInfo& gn1, const GnInfo& gn2, size_t gate, size_t pool, size_t dw_off, size_t w1, size_t b1,
                      size_t w2, int SQ, int row0, int Co) {
    const int Ci = x.C(), Ch = h1->C;
    const int h = h1->H, w = h1->W, P = h * w;
    const int n = N;
    lcm_plan* pl = p;
    // transposed weight copies for the input-gradient GEMMs (packed with the forward weights on every upload)
    const DgradW wpT = make_dgrad_w(name + ".project.weight", (int64_t)Co * Ch, Ch, Co, Ch, 0);     // [Ch][Co]
    const DgradW weT = make_dgrad_w(name + ".expand.weight", (int64_t)Ch * Ci, Ci, Ch, Ci, 0);      // [Ci][Ch]
    DgradW wsT{};
    const bool has_skip = Ci != Co;
    if (has_skip) wsT = make_dgrad_w(name + ".skip.weight", (int64_t)Co * Ci, Ci, Co, Ci, 0);       // [Ci][Co]
    View xv = x;
    // (see step 2 below) per-image project weight gradient + SE gate gradient from its products; needs Wp as plain fp32
    bool xparts16 = true;
    for (int i = 0; i < x.n; ++i) xparts16 = xparts16 && x.part[i]->C % 16 == 0;
    static int se_minp = -1;
    if (se_minp < 0) { const char* e = getenv("LCM_SE_WGRAD_MINP"); se_minp = e ? atoi(e) : 4096; }   // 8192 -> 4096: 82.4 -> 81.1 ms per step (the 64^2 level joins)
    const bool se_from_wgrad = p->tc && P % 128 == 0 && P >= se_minp && Co % 16 == 0 && Ch % 16 == 0 && xparts16 &&
                               !getenv("LCM_NO_SE_FROM_WGRAD");
    const size_t wp_f32 = se_from_wgrad ? p->add_copy(name + ".project.weight", (int64_t)Co * Ch) : 0;
    on_backward([=]() {
      const size_t gsz = pl->gsz;
      const size_t M = (size_t)n * P;
      const int dta = pl->dt_act, dth = pl->dt_hid, dtg = pl->dt_grad;
      const size_t dY = grad_r(out);
      // 1. project dgrad
      const size_t dq = galloc(M * Ch * gsz);
      // streaming packed-fp16 depthwise backward (tensor-core plan): needs sum dq^2 per (image, channel) for its scales
      const bool dwb_stream = pl->tc && Ch % 64 == 0 && !getenv("LCM_NO_DWB_STREAM");
      const size_t dq_stats = dwb_stream ? pl->zballoc((size_t)n * Ch * 2 * sizeof(double)) : (size_t)-1;
      dgrad_gemm(name + ".project.dgrad", dY, wpT, dq, P, dq_stats);
      // 2. project (+ skip) wgrad.  At the high-resolution levels of the tensor-core plan the GEMM runs in per-image mode on
      //    the UNGATED h2 and a small combine kernel produces dWp, dWskip AND the SE gate gradient sum_p dq h2 from its
      //    per-image products (no pass over dq and h2: was 11 ms of bwd_mask_reduce per step).
      const size_t t12se = pl->zballoc((size_t)n * Ch * 2 * sizeof(double));
      const size_t Rimg = se_from_wgrad ? pl->zballoc((size_t)n * (Ch + Ci) * Co * sizeof(float)) : 0;
      {
        std::vector<std::string> writes{name + ".project.weight"};
        if (has_skip) writes.push_back(name + ".skip.weight");
        const size_t gp_w = wg(name + ".project.weight"), gs_w = has_skip ? wg(name + ".skip.weight") : 0;
        pushb(name + ".project.wgrad", "wgrad_simt", se_from_wgrad ? std::vector<std::string>{} : writes, [=](const RunCtx& c, cudaStream_t st) {
          GemmParams gp{};
          int seg_dt[LCM_MAX_SEGS] = {0, 0, 0, 0};
          float* dst[LCM_MAX_SEGS] = {nullptr, nullptr, nullptr, nullptr};
          int dst_ld[LCM_MAX_SEGS] = {0, 0, 0, 0};
          gp.nseg = 1 + xv.n;
          gp.seg[0].A = c.a + h2->off; gp.seg[0].K = Ch; gp.seg[0].ld = Ch; gp.seg[0].mode = se_from_wgrad ? XF_NONE : XF_SCALE;
          gp.seg[0].coef = (const float2*)(c.f + gate); gp.seg[0].coef_ld = Ch; gp.seg[0].coef_off = 0;
          seg_dt[0] = dth; dst[0] = c.wg + gp_w; dst_ld[0] = Ch;
          int col = 0;
          for (int i = 0; i < xv.n; ++i) {
            gp.seg[1 + i].A = c.a + xv.part[i]->off; gp.seg[1 + i].K = xv.part[i]->C; gp.seg[1 + i].ld = xv.part[i]->C;
            gp.seg[1 + i].mode = XF_NONE;
            seg_dt[1 + i] = dta;
            dst[1 + i] = has_skip ? c.wg + gs_w + col : nullptr;   // identity residual: no weight
            dst_ld[1 + i] = Ci;
            col += xv.part[i]->C;
          }
          gp.Ktot = Ch + Ci; gp.P = P; gp.M = (long long)n * P; gp.Nc = Co;
          if (se_from_wgrad) {
            if (launch_wgrad_tc(gp, seg_dt, c.g + dY, dtg, dst, dst_ld, pl->num_sms, st, (float*)(c.zb + Rimg))) *c.launch_err = 1;
            return;
          }
          if (pl->tc && launch_wgrad_tc(gp, seg_dt, c.g + dY, dtg, dst, dst_ld, pl->num_sms, st) == 0) return;
          launch_wgrad_1x1(gp, seg_dt, c.g + dY, dtg, dst, dst_ld, pl->num_sms, st);
        });
        if (se_from_wgrad) {
          pushb(name + ".se.dgate", "se_project_combine", writes, [=](const RunCtx& c, cudaStream_t st) {
            launch_se_project_combine((const float*)(c.zb + Rimg), pl->wf(wp_f32), (const float2*)(c.f + gate), (double*)(c.zb + t12se),
                                      c.wg + gp_w, has_skip ? c.wg + gs_w : nullptr, n, Ch, Ci, Co, st);
          });
        }
      }
      // 3. residual path: identity (r = dY) or skip conv (r = dY Wskip)
      size_t dxres = 0;
      if (has_skip) {
        dxres = galloc(M * Ci * gsz);
        dgrad_gemm(name + ".skip.dgrad", dY, wsT, dxres, P);
      }
      // 4. SE backward
      if (!se_from_wgrad)
      pushb(name + ".se.dgate", "bwd_mask_reduce", {}, [=](const RunCtx& c, cudaStream_t st) {
        launch_bwd_mask_reduce(c.g + dq, dtg, Ch, 0, c.a + h2->off, dth, Ch, 0, nullptr, 0, (double*)(c.zb + t12se), Ch, n, P, Ch, 0, st);
      });
      const size_t coef_se = pl->fballoc((size_t)n * Ch * sizeof(float4));
      const size_t v_pm = pl->fballoc((size_t)n * Ch * 4), v_ds2 = pl->fballoc((size_t)n * Ch * 4);
      const size_t v_z = pl->fballoc((size_t)n * SQ * 4), v_dz1 = pl->fballoc((size_t)n * SQ * 4);
      pushb(name + ".se.bwd", "se_bwd_vec", {}, [=](const RunCtx& c, cudaStream_t st) {
        if (launch_se_bwd_vec((const double*)(c.z + pool), (float)(1.0 / P), pl->wf(w1), pl->wf(b1), pl->wf(w2),
                              (const float2*)(c.f + gate), (const double*)(c.zb + t12se),
                              dwb_stream ? (const double*)(c.zb + dq_stats) : nullptr, (float4*)(c.fb + coef_se),
                              (float*)(c.fb + v_pm), (float*)(c.fb + v_z), (float*)(c.fb + v_ds2), (float*)(c.fb + v_dz1), n, Ch, SQ, st))
          *c.launch_err = 1;
      });
      {
        const size_t g_w1 = wg(name + ".se.fc1.weight"), g_b1 = wg(name + ".se.fc1.bias");
        const size_t g_w2 = wg(name + ".se.fc2.weight"), g_b2 = wg(name + ".se.fc2.bias");
        pushb(name + ".se.wgrad", "outer_sum",
              {name + ".se.fc1.weight", name + ".se.fc1.bias", name + ".se.fc2.weight", name + ".se.fc2.bias"},
              [=](const RunCtx& c, cudaStream_t st) {
                launch_outer_sum((const float*)(c.fb + v_ds2), Ch, (const float*)(c.fb + v_z), SQ, c.wg + g_w2, c.wg + g_b2, n, Ch, SQ, st);
                launch_outer_sum((const float*)(c.fb + v_dz1), SQ, (const float*)(c.fb + v_pm), Ch, c.wg + g_w1, c.wg + g_b1, n, SQ, Ch, st);
              });
        p->bwd_ops.back().launches = 2;
      }
      // 5. depthwise backward (+ SE scale prologue, ReLU6 backward, norm2 reductions)
      const size_t du = galloc(M * Ch * gsz);
      const size_t t12h = pl->zballoc((size_t)n * Ch * 2 * sizeof(double));
      {
        const size_t g_dw = wg(name + ".depthwise.weight");
        pushb(name + ".depthwise.bwd", "dwconv_bwd", {name + ".depthwise.weight"}, [=](const RunCtx& c, cudaStream_t st) {
          if (dwb_stream) {
            if (launch_dwconv_bwd_stream(c.g + dq, (const float4*)(c.fb + coef_se), c.a + h1->off, (const float2*)(c.f + gn2.coef),
                                         pl->wf(dw_off), c.g + du, (double*)(c.zb + t12h), c.wg + g_dw, n, h, w, Ch, pl->num_sms, st))
              *c.launch_err = 1;
            return;
          }
          launch_dwconv_bwd(c.g + dq, dtg, (const float4*)(c.fb + coef_se), c.a + h1->off, dth, (const float2*)(c.f + gn2.coef),
                            pl->wf(dw_off), c.g + du, (double*)(c.zb + t12h), c.wg + g_dw, n, h, w, Ch, pl->num_sms, st);
        });
      }
      gfree(dq, M * Ch * gsz);
      // 6.-7. norm2 + FiLM backward, dh1 in place
      // (streaming path: du is stored as fp16 scaled by s; the 1/s rides in A, the in-place result is the plan's gradient type)
      const size_t coef4h = gn_backward(name + ".norm2.bwd", View::of(h1), gn2, t12h, row0, dwb_stream ? coef_se : (size_t)-1);
      pushb(name + ".norm2.apply", "bwd_affine3", {}, [=](const RunCtx& c, cudaStream_t st) {
        launch_bwd_affine3(c.g + du, dwb_stream ? (int)DT_F16 : dtg, Ch, 0, c.a + h1->off, dth, Ch, 0, (const float4*)(c.fb + coef4h), Ch, 0, nullptr, 0, 0, 0,
                           c.g + du, dtg, Ch, 0, 0, n, P, Ch, st);
      });
      // 8. expand dgrad
      const size_t dr = galloc(M * Ci * gsz);
      dgrad_gemm(name + ".expand.dgrad", du, weT, dr, P);
      // 9. expand wgrad
      {
        const size_t g_we = wg(name + ".expand.weight");
        pushb(name + ".expand.wgrad", "wgrad_simt", {name + ".expand.weight"}, [=](const RunCtx& c, cudaStream_t st) {
          GemmParams gp{};
          int seg_dt[LCM_MAX_SEGS] = {0, 0, 0, 0};
          float* dst[LCM_MAX_SEGS] = {nullptr, nullptr, nullptr, nullptr};
          int dst_ld[LCM_MAX_SEGS] = {0, 0, 0, 0};
          gp.nseg = xv.n;
          int col = 0;
          for (int i = 0; i < xv.n; ++i) {
            gp.seg[i].A = c.a + xv.part[i]->off; gp.seg[i].K = xv.part[i]->C; gp.seg[i].ld = xv.part[i]->C;
            gp.seg[i].mode = XF_AFFINE_RELU6;
            gp.seg[i].coef = (const float2*)(c.f + gn1.coef); gp.seg[i].coef_ld = Ci; gp.seg[i].coef_off = col;
            seg_dt[i] = dta; dst[i] = c.wg + g_we + col; dst_ld[i] = Ci;
            col += xv.part[i]->C;
          }
          gp.Ktot = Ci; gp.P = P; gp.M = (long long)n * P; gp.Nc = Ch;
          if (pl->tc && launch_wgrad_tc(gp, seg_dt, c.g + du, dtg, dst, dst_ld, pl->num_sms, st) == 0) return;
          launch_wgrad_1x1(gp, seg_dt, c.g + du, dtg, dst, dst_ld, pl->num_sms, st);
        });
      }
      gfree(du, M * Ch * gsz);
      // 10. ReLU6 backward + norm1 reductions, per input part
      const size_t t12x = pl->zballoc((size_t)n * Ci * 2 * sizeof(double));
      {
        int col = 0;
        for (int i = 0; i < xv.n; ++i) {
          const TensorP part = xv.part[i];
          const int c0 = col;
          pushb(name + ".norm1.mask" + (xv.n > 1 ? std::to_string(i) : ""), "bwd_mask_reduce", {}, [=](const RunCtx& c, cudaStream_t st) {
            launch_bwd_mask_reduce(c.g + dr, dtg, Ci, c0, c.a + part->off, dta, part->C, 0, (const float2*)(c.f + gn1.coef), Ci,
                                   (double*)(c.zb + t12x), Ci, n, P, part->C, 1, st);
          });
          col += part->C;
        }
      }
      // 11. norm1 backward coefficients
      const size_t coef4x = gn_backward(name + ".norm1.bwd", xv, gn1, t12x, -1);
      // 12. dx (+)= A dpre + B x + C + r
      {
        int col = 0;
        for (int i = 0; i < xv.n; ++i) {
          const TensorP part = xv.part[i];
          const int c0 = col;
          const GradW gx = grad_w(part);
          const size_t r_off = has_skip ? dxres : dY;
          const int r_ld = has_skip ? Ci : Co;
          pushb(name + ".norm1.apply" + (xv.n > 1 ? std::to_string(i) : ""), "bwd_affine3", {}, [=](const RunCtx& c, cudaStream_t st) {
            launch_bwd_affine3(c.g + dr, dtg, Ci, c0, c.a + part->off, dta, part->C, 0, (const float4*)(c.fb + coef4x), Ci, c0,
                               c.g + r_off, dtg, r_ld, c0, c.g + gx.off, dtg, part->C, 0, gx.accumulate, n, P, part->C, st);
          });
          col += part->C;
        }
      }
      gfree(dr, M * Ci * gsz);
      if (has_skip) gfree(dxres, M * Ci * gsz);
      grad_done(out);
    });
  }

  // ---- LinearAttention (efficient_unet.py:273-308) ---------------------------------------------
  // ---- StandardAttention (efficient_unet.py:311-357): norm -> to_qkv -> softmax(q k^T d^-0.5) v -> to_out -> + x -----------
  // state_dict: norm.{weight,bias}, to_qkv.weight, to_out.weight (a plain conv, no GroupNorm after it); the residual is the
  // identity K-segment of the to_out GEMM, like a block's project.
  TensorP standard_attention(const std::string& name, const TensorP& x) {
    const int C = x->C, heads = p->cfg.num_attention_heads, inner = heads * 32;
    const int h = x->H, w = x->W, P = h * w;
    const double es = (double)p->esz;
    const GnInfo gnn = gn_coef(name + ".norm", View::of(x), name + ".norm", -1);
    TensorP qkv = p->new_tensor(3 * inner, h, w, false, name + ".qkv");
    GemmW wq = make_w(3 * inner, {C});
    p->add_weight(name + ".to_qkv.weight", (int64_t)3 * inner * C, mat_job(wq, 0, PACK_MAT, 3 * inner, C, C, 0));
    gemm(name + ".to_qkv", {{x, gnn.coef, C, 0, XF_AFFINE}}, wq, qkv, false,
         (C + 3.0 * inner) * N * P * es + 3.0 * inner * C * es, 2.0 * N * P * C * 3 * inner);
    TensorP o = p->new_tensor(inner, h, w, false, name + ".attn");
    {
      lcm_plan* pl = p; const int n = N;
      push(name + ".softmax", "attn_softmax", 4.0 * inner * N * P * es, 4.0 * N * heads * (double)P * P * 32,
           [=](const RunCtx& c, cudaStream_t st) { launch_attn_softmax(c.a + qkv->off, c.a + o->off, n, P, heads, pl->bf16, st); });
    }
    p->release(qkv);
    TensorP y = p->new_tensor(C, h, w, true, name + ".out");
    GemmW wo = make_w(C, {inner, C});
    p->add_weight(name + ".to_out.weight", (int64_t)C * inner, mat_job(wo, 0, PACK_MAT, C, inner, inner, 0));
    p->identity_jobs.push_back(mat_job(wo, 1, PACK_IDENTITY, C, C, C, 0));
    gemm(name + ".to_out", {{o, 0, 0, 0, XF_NONE}, {x, 0, 0, 0, XF_NONE}}, wo, y, true,
         (2.0 * C + inner) * N * P * es + (double)inner * C * es, 2.0 * N * P * C * inner);
    p->release(o);
    p->release(x);
    return y;
  }

  TensorP attention(const std::string& name, const TensorP& x) {
    if (p->cfg.standard_attention) return standard_attention(name, x);
    const int C = x->C, heads = p->cfg.num_attention_heads, inner = heads * 32;
    const int h = x->H, w = x->W, P = h * w;
    const double es = (double)p->esz;
    const GnInfo gnn = gn_coef(name + ".norm", View::of(x), name + ".norm", -1);
    const size_t coefn = gnn.coef;
    TensorP qkv = p->new_tensor(3 * inner, h, w, false, name + ".qkv");
    GemmW wq = make_w(3 * inner, {C});
    p->add_weight(name + ".to_qkv.weight", (int64_t)3 * inner * C, mat_job(wq, 0, PACK_MAT, 3 * inner, C, C, 0));
    gemm(name + ".to_qkv", {{x, coefn, C, 0, XF_AFFINE}}, wq, qkv, false,
         (C + 3.0 * inner) * N * P * es + 3.0 * inner * C * es, 2.0 * N * P * C * 3 * inner);
    const size_t state = p->zalloc((size_t)N * heads * 32 * 33 * sizeof(double));
    TensorP o = p->new_tensor(inner, h, w, false, name + ".attn");
    {
      lcm_plan* pl = p; const int n = N;
      push(name + ".kv", "attn_kv", 2.0 * inner * N * P * es, 2.0 * N * heads * P * 32 * 33,
           [=](const RunCtx& c, cudaStream_t st) { launch_attn_kv(c.a + qkv->off, (double*)(c.z + state), n, P, heads, pl->bf16, st); });
      push(name + ".apply", "attn_apply", 2.0 * inner * N * P * es, 2.0 * N * heads * P * 32 * 33,
           [=](const RunCtx& c, cudaStream_t st) {
             launch_attn_apply(c.a + qkv->off, (const double*)(c.z + state), c.a + o->off, n, P, heads, pl->bf16, st);
           });
    }
    p->release(qkv);
    TensorP u = p->new_tensor(C, h, w, true, "");
    GemmW wo = make_w(C, {inner});
    p->add_weight(name + ".to_out.0.weight", (int64_t)C * inner, mat_job(wo, 0, PACK_MAT, C, inner, inner, 0));
    gemm(name + ".to_out", {{o, 0, 0, 0, XF_NONE}}, wo, u, true, (C + (double)inner) * N * P * es + (double)inner * C * es,
         2.0 * N * P * C * inner);
    p->release(o);
    const GnInfo gno = gn_coef(name + ".to_out.1", View::of(u), name + ".to_out.1", -1);
    const size_t coefo = gno.coef;
    TensorP y = p->new_tensor(C, h, w, true, name + ".out");
    {
      lcm_plan* pl = p; const int n = N;
      push(name + ".gn_residual", "affine_residual", 3.0 * C * N * P * es, 3.0 * N * P * C,
           [=](const RunCtx& c, cudaStream_t st) {
             launch_affine_residual(c.a + u->off, (const float2*)(c.f + coefo), c.a + x->off, c.a + y->off,
                                    (double*)(c.z + y->stats_off), n, P, C, pl->bf16, st);
           });
    }
    if (p->train) {
      // reverse of efficient_unet.py:273-308
      const DgradW woT = make_dgrad_w(name + ".to_out.0.weight", (int64_t)C * inner, inner, C, inner, 0);        // [inner][C]
      const DgradW wqT = make_dgrad_w(name + ".to_qkv.weight", (int64_t)3 * inner * C, C, 3 * inner, C, 0);     // [C][3 inner]
      lcm_plan* pl = p; const int n = N;
      on_backward([=]() {
        const size_t gsz = pl->gsz, M = (size_t)n * P;
        const int dta = pl->dt_act, dtg = pl->dt_grad;
        const size_t dy = grad_r(y);
        // residual edge: dx (+)= dy
        const GradW gx0 = grad_w(x);
        pushb(name + ".residual.bwd", "bwd_add", {}, [=](const RunCtx& c, cudaStream_t st) {
          launch_bwd_add(c.g + dy, dtg, C, 0, c.g + gx0.off, dtg, C, 0, gx0.accumulate, (long long)n * P, C, st);
        });
        // GroupNorm after to_out: du = A dy + B u + C
        const size_t t12u = pl->zballoc((size_t)n * C * 2 * sizeof(double));
        pushb(name + ".to_out.1.reduce", "bwd_mask_reduce", {}, [=](const RunCtx& c, cudaStream_t st) {
          launch_bwd_mask_reduce(c.g + dy, dtg, C, 0, c.a + u->off, dta, C, 0, nullptr, 0, (double*)(c.zb + t12u), C, n, P, C, 0, st);
        });
        const size_t coef4u = gn_backward(name + ".to_out.1.bwd", View::of(u), gno, t12u, -1);
        const size_t du = galloc(M * C * gsz);
        pushb(name + ".to_out.1.apply", "bwd_affine3", {}, [=](const RunCtx& c, cudaStream_t st) {
          launch_bwd_affine3(c.g + dy, dtg, C, 0, c.a + u->off, dta, C, 0, (const float4*)(c.fb + coef4u), C, 0, nullptr, 0, 0, 0,
                             c.g + du, dtg, C, 0, 0, n, P, C, st);
        });
        // to_out.0: dWo = du^T o ; do = du Wo
        {
          const size_t g_wo = wg(name + ".to_out.0.weight");
          pushb(name + ".to_out.0.wgrad", "wgrad_simt", {name + ".to_out.0.weight"}, [=](const RunCtx& c, cudaStream_t st) {
            GemmParams gp{};
            int seg_dt[LCM_MAX_SEGS] = {dta, 0, 0, 0};
            float* dst[LCM_MAX_SEGS] = {c.wg + g_wo, nullptr, nullptr, nullptr};
            int dst_ld[LCM_MAX_SEGS] = {inner, 0, 0, 0};
            gp.nseg = 1; gp.seg[0].A = c.a + o->off; gp.seg[0].K = inner; gp.seg[0].ld = inner; gp.seg[0].mode = XF_NONE;
            gp.Ktot = inner; gp.P = P; gp.M = (long long)n * P; gp.Nc = C;
            if (pl->tc && launch_wgrad_tc(gp, seg_dt, c.g + du, dtg, dst, dst_ld, pl->num_sms, st) == 0) return;
          launch_wgrad_1x1(gp, seg_dt, c.g + du, dtg, dst, dst_ld, pl->num_sms, st);
          });
        }
        const size_t dO = galloc(M * inner * gsz);
        dgrad_gemm(name + ".to_out.0.dgrad", du, woT, dO, P);
        gfree(du, M * C * gsz);
        // attention core
        const size_t dstate = pl->zballoc((size_t)n * heads * 32 * 33 * sizeof(double));
        const size_t dqkv = galloc(M * 3 * inner * gsz);
        pushb(name + ".attn.bwd", "attn_bwd", {}, [=](const RunCtx& c, cudaStream_t st) {
          launch_attn_bwd(c.a + qkv->off, dta, (const double*)(c.z + state), c.g + dO, dtg, c.g + dqkv, (double*)(c.zb + dstate), n, P,
                          heads, st);
        });
        p->bwd_ops.back().launches = 2;
        gfree(dO, M * inner * gsz);
        // to_qkv: dWqkv = dqkv^T (a x + b) ; dxn = dqkv Wqkv
        {
          const size_t g_wq = wg(name + ".to_qkv.weight");
          pushb(name + ".to_qkv.wgrad", "wgrad_simt", {name + ".to_qkv.weight"}, [=](const RunCtx& c, cudaStream_t st) {
            GemmParams gp{};
            int seg_dt[LCM_MAX_SEGS] = {dta, 0, 0, 0};
            float* dst[LCM_MAX_SEGS] = {c.wg + g_wq, nullptr, nullptr, nullptr};
            int dst_ld[LCM_MAX_SEGS] = {C, 0, 0, 0};
            gp.nseg = 1; gp.seg[0].A = c.a + x->off; gp.seg[0].K = C; gp.seg[0].ld = C; gp.seg[0].mode = XF_AFFINE;
            gp.seg[0].coef = (const float2*)(c.f + gnn.coef); gp.seg[0].coef_ld = C; gp.seg[0].coef_off = 0;
            gp.Ktot = C; gp.P = P; gp.M = (long long)n * P; gp.Nc = 3 * inner;
            if (pl->tc && launch_wgrad_tc(gp, seg_dt, c.g + dqkv, dtg, dst, dst_ld, pl->num_sms, st) == 0) return;
            launch_wgrad_1x1(gp, seg_dt, c.g + dqkv, dtg, dst, dst_ld, pl->num_sms, st);
          });
        }
        const size_t dxn = galloc(M * C * gsz);
        dgrad_gemm(name + ".to_qkv.dgrad", dqkv, wqT, dxn, P);
        gfree(dqkv, M * 3 * inner * gsz);
        // first GroupNorm: dx += A dxn + B x + C
        const size_t t12x = pl->zballoc((size_t)n * C * 2 * sizeof(double));
        pushb(name + ".norm.reduce", "bwd_mask_reduce", {}, [=](const RunCtx& c, cudaStream_t st) {
          launch_bwd_mask_reduce(c.g + dxn, dtg, C, 0, c.a + x->off, dta, C, 0, nullptr, 0, (double*)(c.zb + t12x), C, n, P, C, 0, st);
        });
        const size_t coef4x = gn_backward(name + ".norm.bwd", View::of(x), gnn, t12x, -1);
        const GradW gx1 = grad_w(x);
        pushb(name + ".norm.apply", "bwd_affine3", {}, [=](const RunCtx& c, cudaStream_t st) {
          launch_bwd_affine3(c.g + dxn, dtg, C, 0, c.a + x->off, dta, C, 0, (const float4*)(c.fb + coef4x), C, 0, nullptr, 0, 0, 0,
                             c.g + gx1.off, dtg, C, 0, gx1.accumulate, n, P, C, st);
        });
        gfree(dxn, M * C * gsz);
        grad_done(y);
      });
    }
    p->release(u);
    p->release(x);
    return y;
  }

  // ---- Downsample / Upsample (efficient_unet.py:360-384) -------------------------------------
  TensorP conv3(const std::string& name, const std::string& wname, const TensorP& x_in, int mode_in) {
    TensorP x = x_in;
    int mode = mode_in;
    TensorP up_src;   // training: the low-resolution input of an up-convolution
    if (mode_in == CONV_UP2) up_src = x_in;
    if (p->tc && mode == CONV_UP2) {
      // tensor-core path: materialise the bilinear x2 once (memory-bound), then a stride-1 conv whose operand
      // tiles are TMA boxes; the algorithmic accounting stays with the conv op below.
      TensorP up = p->new_tensor(x_in->C, x_in->H * 2, x_in->W * 2, false, "");
      lcm_plan* pl = p; const int n = N;
      TensorP src = x_in;
      push(name + ".bilinear", "upsample2x", 0, 0, [=](const RunCtx& c, cudaStream_t st) {
        launch_upsample2x(c.a + src->off, c.a + up->off, n, src->H, src->W, src->C, st);
      });
      p->release(x_in);
      x = up;
      mode = CONV_S1;
    }
    const int C = x->C;
    const int Ho = mode == CONV_S2 ? x->H / 2 : (mode == CONV_UP2 ? x->H * 2 : x->H);
    const int Wo = mode == CONV_S2 ? x->W / 2 : (mode == CONV_UP2 ? x->W * 2 : x->W);
    TensorP out = p->new_tensor(C, Ho, Wo, true, name);
    const int Cpad = (C + 63) / 64 * 64;
    PackJob j{};
    j.kind = PACK_CONV3; j.bf16 = p->bf16 ? 1 : 0; j.R = C; j.Ci = C;
    int block_n = 0;
    size_t w_off;
    if (p->tc) {
      block_n = gemm_tc_pick_block_n(C);
      w_off = p->walloc((size_t)C * 9 * Cpad * sizeof(bf16));
      j.layout = WL_UMMA; j.ld = 9 * Cpad; j.tap_stride = Cpad; j.block_n = block_n;
    } else {
      w_off = p->walloc((size_t)C * 9 * C * p->esz);
      j.layout = WL_ROWMAJOR; j.ld = 9 * C; j.tap_stride = C;
    }
    j.dst = (void*)w_off;
    p->add_weight(wname + ".weight", (int64_t)C * C * 9, j);
    const size_t b_off = p->add_copy(wname + ".bias", C);
    lcm_plan* pl = p; const int n = N;
    const int Hin = x->H, Win = x->W;
    const double es = (double)p->esz;
    push(name, pl->tc ? "conv3x3_tc" : "conv3x3_simt",
         ((double)C * (mode_in == CONV_UP2 ? Hin * Win / 4.0 : (double)Hin * Win) + (double)C * Ho * Wo) * N * es + 9.0 * C * C * es,
         18.0 * N * Ho * Wo * C * C,
         [=](const RunCtx& c, cudaStream_t st) {
           if (pl->tc) {
             GemmParams gp{};
             gp.nseg = 1; gp.seg[0].A = c.a + x->off; gp.seg[0].K = 9 * C; gp.seg[0].ld = C; gp.seg[0].mode = XF_NONE;
             gp.Ktot = 9 * C; gp.W = pl->wbase + w_off; gp.out = c.a + out->off;
             gp.stats = (double*)(c.z + out->stats_off); gp.P = Ho * Wo; gp.M = (long long)n * Ho * Wo; gp.Nc = C;
             ConvGeom cg{mode, Hin, Win, Ho, Wo, C, pl->wf(b_off)};
             if (launch_gemm_tc(gp, cg, block_n, pl->num_sms, st)) *c.launch_err = 1;
           } else {
             launch_conv3x3_simt(c.a + x->off, pl->wbase + w_off, pl->wf(b_off), c.a + out->off,
                                 (double*)(c.z + out->stats_off), n, Hin, Win, C, C, mode, pl->bf16, st);
           }
         });
    if (p->train) {
      // transposed-conv weights: W'[ci][(8 - tap) * stride + co] (flipped taps, swapped channel roles)
      PackJob jt{};
      jt.kind = PACK_CONV3_T; jt.bf16 = p->bf16 ? 1 : 0; jt.R = C; jt.Cc = C; jt.Ci = C;
      // stride-1 transposed conv = conv over dY: tcgen05 kernel.  A stride-2 conv's input gradient is the same stride-1
      // transposed conv over the zero-inserted dY (Z[2y][2x] = dY[y][x]): 4x the MMA work of a phase decomposition, but
      // the tensor core is idle otherwise and the op stays one halo-mode conv (was 10.8 ms of CUDA-core time per step)
      const bool s2_tc = p->tc && mode_in == CONV_S2 && x->H % 2 == 0 && x->W % 2 == 0;
      const bool tc_dgrad = p->tc && (mode_in == CONV_UP2 || s2_tc);
      size_t wt_off;
      int bn_t = 0;
      if (tc_dgrad) {
        bn_t = gemm_tc_pick_block_n(C);
        wt_off = p->walloc((size_t)C * 9 * Cpad * sizeof(bf16));
        jt.layout = WL_UMMA; jt.ld = 9 * Cpad; jt.tap_stride = Cpad; jt.block_n = bn_t;
      } else {
        wt_off = p->walloc((size_t)C * 9 * C * p->esz);
        jt.layout = WL_ROWMAJOR; jt.ld = 9 * C; jt.tap_stride = C;
      }
      jt.dst = (void*)wt_off;
      p->add_weight(wname + ".weight", (int64_t)C * C * 9, jt);
      on_backward([=]() {
        const size_t gsz = pl->gsz;
        const int dta = pl->dt_act, dtg = pl->dt_grad;
        const size_t dY = grad_r(out);
        const size_t g_w = wg(wname + ".weight"), g_b = wg(wname + ".bias");
        // weight + bias gradient; the fp32 plan recomputes the bilinear blend inside the loader (mode UP2 on the low-res input)
        pushb(name + ".wgrad", "wgrad_simt", {wname + ".weight", wname + ".bias"}, [=](const RunCtx& c, cudaStream_t st) {
          if (pl->tc && (mode == CONV_S1 || s2_tc) &&
              launch_wgrad_conv3_tc(c.a + x->off, c.g + dY, c.wg + g_w, n, Hin, Win, C, C, mode == CONV_S2 ? 2 : 1, pl->num_sms, st) == 0) {
            launch_colsum(c.g + dY, dtg, (long long)n * Ho * Wo, C, c.wg + g_b, st);
            return;
          }
          launch_wgrad_conv3(c.a + x->off, dta, c.g + dY, dtg, c.wg + g_w, c.wg + g_b, n, Hin, Win, C, C, mode, pl->num_sms, st);
        });
        if (mode_in == CONV_S2) {
          const GradW gx = grad_w(x);
          const size_t bytes = (size_t)n * Hin * Win * C * gsz;
          const size_t tmp = gx.accumulate ? galloc(bytes) : 0;
          if (s2_tc) {
            const size_t Z = galloc(bytes);
            pushb(name + ".dgrad.zero_insert", "zero_insert2x", {}, [=](const RunCtx& c, cudaStream_t st) {
              launch_zero_insert2x(c.g + dY, c.g + Z, n, Ho, Wo, C, st);
            });
            pushb(name + ".dgrad", "conv3x3_tc", {}, [=](const RunCtx& c, cudaStream_t st) {
              GemmParams gp{};
              gp.nseg = 1; gp.seg[0].A = c.g + Z; gp.seg[0].K = 9 * C; gp.seg[0].ld = C; gp.seg[0].mode = XF_NONE;
              gp.Ktot = 9 * C; gp.W = pl->wbase + wt_off; gp.out = c.g + (gx.accumulate ? tmp : gx.off); gp.stats = nullptr;
              gp.P = Hin * Win; gp.M = (long long)n * Hin * Win; gp.Nc = C;
              ConvGeom cg{CONV_S1, Hin, Win, Hin, Win, C, nullptr};
              if (launch_gemm_tc(gp, cg, bn_t, pl->num_sms, st)) *c.launch_err = 1;
            });
            gfree(Z, bytes);
          } else
          pushb(name + ".dgrad", "conv3x3_dgrad_simt", {}, [=](const RunCtx& c, cudaStream_t st) {
            launch_conv3x3_dgrad_simt(c.g + dY, pl->wbase + wt_off, c.g + (gx.accumulate ? tmp : gx.off), n, Hin, Win, C, C, CONV_S2,
                                      pl->bf16, st);
          });
          if (gx.accumulate) {
            pushb(name + ".dgrad.add", "bwd_add", {}, [=](const RunCtx& c, cudaStream_t st) {
              launch_bwd_add(c.g + tmp, dtg, C, 0, c.g + gx.off, dtg, C, 0, 1, (long long)n * Hin * Win, C, st);
            });
            gfree(tmp, bytes);
          }
        } else {
          // dUp = transposed stride-1 conv of dY, then the transpose of the bilinear x2
          const size_t bytes = (size_t)n * Ho * Wo * C * gsz;
          const size_t dUp = galloc(bytes);
          pushb(name + ".dgrad", tc_dgrad ? "conv3x3_tc" : "conv3x3_dgrad_simt", {}, [=](const RunCtx& c, cudaStream_t st) {
            if (tc_dgrad) {
              GemmParams gp{};
              gp.nseg = 1; gp.seg[0].A = c.g + dY; gp.seg[0].K = 9 * C; gp.seg[0].ld = C; gp.seg[0].mode = XF_NONE;
              gp.Ktot = 9 * C; gp.W = pl->wbase + wt_off; gp.out = c.g + dUp; gp.stats = nullptr;
              gp.P = Ho * Wo; gp.M = (long long)n * Ho * Wo; gp.Nc = C;
              ConvGeom cg{CONV_S1, Ho, Wo, Ho, Wo, C, nullptr};
              if (launch_gemm_tc(gp, cg, bn_t, pl->num_sms, st)) *c.launch_err = 1;
            } else {
              launch_conv3x3_dgrad_simt(c.g + dY, pl->wbase + wt_off, c.g + dUp, n, Ho, Wo, C, C, CONV_S1, pl->bf16, st);
            }
          });
          const GradW gx = grad_w(up_src);
          pushb(name + ".bilinear.bwd", "upsample2x_bwd", {}, [=](const RunCtx& c, cudaStream_t st) {
            launch_upsample2x_bwd(c.g + dUp, c.g + gx.off, dtg, n, up_src->H, up_src->W, C, gx.accumulate, st);
          });
          gfree(dUp, bytes);
        }
        grad_done(out);
      });
    }
    p->release(x);
    return out;
  }
};

int build_plan(lcm_plan* p) {
  const lcm_unet_config& c = p->cfg;
  Builder b(p);
  const int N = p->N, H = p->H, W = p->W;
  std::vector<int> widths;
  for (int i = 0; i < c.num_levels; ++i) widths.push_back(c.base_channels * c.channel_multipliers[i]);
  auto has_attn = [&](int res) {
    for (int i = 0; i < c.num_attention_resolutions; ++i)
      if (c.attention_resolutions[i] == res) return true;
    return false;
  };
  const int ted = c.time_embed_dim, base = c.base_channels;

  if (c.in_channels == c.out_channels) p->addtmp_f_off = p->falloc((size_t)N * c.in_channels * H * W * sizeof(float));
  // a2: time embedding (efficient_unet.py:550)
  {
    const size_t w1 = p->add_copy("time_mlp.1.weight", (int64_t)ted * base), b1 = p->add_copy("time_mlp.1.bias", ted);
    const size_t w3 = p->add_copy("time_mlp.3.weight", (int64_t)ted * ted), b3 = p->add_copy("time_mlp.3.bias", ted);
    p->silu_f_off = p->falloc((size_t)N * ted * sizeof(float));
    b.push("time_mlp", "time_embed", 0, 0, [=](const RunCtx& cx, cudaStream_t st) {
      launch_time_embed(cx.t_dev, cx.t_scalar, N, base, ted, p->wf(w1), p->wf(b1), p->wf(w3), p->wf(b3), nullptr,
                        (float*)(cx.f + p->silu_f_off), st);
    });
    if (p->train) {
      p->dst_zb_off = p->zballoc((size_t)N * ted * sizeof(float));
      Builder* bb = &b;
      b.on_backward([=]() {
        const size_t g1 = bb->wg("time_mlp.1.weight"), gb1 = bb->wg("time_mlp.1.bias");
        const size_t g3 = bb->wg("time_mlp.3.weight"), gb3 = bb->wg("time_mlp.3.bias");
        bb->pushb("time_mlp.bwd", "time_mlp_bwd", {"time_mlp.1.weight", "time_mlp.1.bias", "time_mlp.3.weight", "time_mlp.3.bias"},
                  [=](const RunCtx& cx, cudaStream_t st) {
                    launch_time_mlp_bwd(cx.t_dev, cx.t_scalar, N, base, ted, p->wf(w1), p->wf(b1), p->wf(w3), p->wf(b3),
                                        (const float*)(cx.zb + p->dst_zb_off), cx.wg + g1, cx.wg + gb1, cx.wg + g3, cx.wg + gb3, st);
                  });
      });
    }
  }
  const size_t film_op_index = p->ops.size();   // filled in once all blocks are enumerated
  b.push("film", "film", 0, 0, nullptr);
  const size_t film_tape_index = b.tape.size();
  b.on_backward(nullptr);                       // likewise (needs the complete row table)

  // a3: init conv fused with the conditioning concat (:553, low_light_diffusion.py:222)
  TensorP h = p->new_tensor(widths[0], H, W, true, "init_conv");
  {
    const int Cin = c.in_channels, Co = widths[0];
    const size_t w_off = p->walloc((size_t)9 * Cin * Co * sizeof(float));
    PackJob j{}; j.kind = PACK_CONV3_KN; j.dst = (void*)w_off; j.R = Co; j.Ci = Cin;
    p->add_weight("init_conv.weight", (int64_t)Co * Cin * 9, j);
    const size_t b_off = p->add_copy("init_conv.bias", Co);
    TensorP out = h;
    b.push("init_conv", "init_conv", (double)N * H * W * (Cin * 4.0 + Co * (double)p->esz), 18.0 * N * H * W * Cin * Co,
           [=](const RunCtx& cx, cudaStream_t st) {
             if (p->tc && launch_init_conv_h2(cx.xa, cx.ca, cx.sa, cx.xb, cx.cb, cx.sb, p->wf(w_off), p->wf(b_off), cx.a + out->off,
                                              (double*)(cx.z + out->stats_off), N, H, W, Co, st)) return;
             launch_init_conv(cx.xa, cx.ca, cx.sa, cx.xb, cx.cb, cx.sb, p->wf(w_off), p->wf(b_off), cx.a + out->off,
                              (double*)(cx.z + out->stats_off), N, H, W, Co, p->bf16, st);
           });
    if (p->train) {
      Builder* bb = &b;
      b.on_backward([=]() {   // the network input needs no gradient: weight + bias only
        const size_t dY = bb->grad_r(out);
        const size_t g_w = bb->wg("init_conv.weight"), g_b = bb->wg("init_conv.bias");
        bb->pushb("init_conv.wgrad", "init_conv_wgrad", {"init_conv.weight", "init_conv.bias"}, [=](const RunCtx& cx, cudaStream_t st) {
          if (launch_init_conv_wgrad(cx.xa, cx.ca, cx.sa, cx.xb, cx.cb, cx.sb, cx.g + dY, p->dt_grad, cx.wg + g_w, cx.wg + g_b, N, H, W,
                                     Co, p->num_sms, st)) *cx.launch_err = 1;
        });
        bb->grad_done(out);
      });
    }
  }

  auto run_level = [&](const std::string& prefix, int nblocks, int res, View x, int Co) -> TensorP {
    int idx = 0;
    TensorP cur;
    for (int i = 0; i < nblocks; ++i) {
      cur = b.block(prefix + "." + std::to_string(idx++), x, Co, c.se_ratio);
      if (has_attn(res)) cur = b.attention(prefix + "." + std::to_string(idx++), cur);
      x = View::of(cur);
    }
    return cur;
  };

  // encoder (:558-570); the level output is kept for the decoder (refs+1)
  std::vector<TensorP> skips;
  int res = c.image_size;
  for (int li = 0; li < c.num_levels; ++li) {
    h = run_level("encoder_blocks." + std::to_string(li), c.num_res_blocks, res, View::of(h), widths[li]);
    h->refs++;
    skips.push_back(h);
    if (li < c.num_levels - 1) {
      const std::string nm = "downsamplers." + std::to_string(li);
      h = b.conv3(nm, nm + ".down", h, CONV_S2);
      res /= 2;
    }
  }
  // middle (:573-575)
  h = b.block("mid_block1", View::of(h), widths.back(), 0.25f);
  h = b.attention("mid_attn", h);
  h = b.block("mid_block2", View::of(h), widths.back(), 0.25f);
  // decoder (:580-597)
  for (int li = 0; li < c.num_levels; ++li) {
    const int Co = widths[c.num_levels - 1 - li];
    if (li > 0) {
      const std::string nm = "upsamplers." + std::to_string(li - 1);
      h = b.conv3(nm, nm + ".conv", h, CONV_UP2);
      res *= 2;
    }
    View x;
    x.part[0] = h; x.part[1] = skips.back(); x.n = 2;
    skips.pop_back();
    h = run_level("decoder_blocks." + std::to_string(li), c.num_res_blocks + 1, res, x, Co);
  }
  // a8 + a11: final norm + SiLU + conv, LCM step fused (:600-602, lcm_scheduler.py:214-242)
  {
    const Builder::GnInfo gnf = b.gn_coef("final_norm", View::of(h), "final_norm", -1);
    const size_t coef = gnf.coef;
    const int Ci = widths[0], Co = c.out_channels;
    const size_t w_off = p->walloc((size_t)9 * Ci * Co * sizeof(float));
    PackJob j{}; j.kind = PACK_CONV3_KN; j.dst = (void*)w_off; j.R = Co; j.Ci = Ci;
    p->add_weight("final_conv.weight", (int64_t)Co * Ci * 9, j);
    const size_t b_off = p->add_copy("final_conv.bias", Co);
    TensorP in = h;
    b.push("final_conv", "final_conv", (double)N * H * W * (Ci * (double)p->esz + Co * 4.0 * 3), 18.0 * N * H * W * Ci * Co,
           [=](const RunCtx& cx, cudaStream_t st) {
             if (p->tc && launch_final_conv_h2(cx.a + in->off, (const float2*)(cx.f + coef), p->wf(w_off), p->wf(b_off), cx.eps,
                                               cx.step, N, H, W, Ci, Co, st)) return;
             launch_final_conv(cx.a + in->off, (const float2*)(cx.f + coef), p->wf(w_off), p->wf(b_off), cx.eps, cx.step,
                               N, H, W, Ci, Co, p->bf16, st);
           });
    if (p->train) {
      Builder* bb = &b;
      b.on_backward([=]() {
        // d eps = grad_scale * loss'(eps - target) / numel is formed inside the kernel; dpre = d(final_norm output)
        const size_t bytes = (size_t)N * H * W * Ci * p->gsz;
        const size_t dpre = bb->galloc(bytes);
        const size_t t12 = p->zballoc((size_t)N * Ci * 2 * sizeof(double));
        const size_t g_w = bb->wg("final_conv.weight"), g_b = bb->wg("final_conv.bias");
        bb->pushb("final_conv.bwd", "final_conv_bwd", {"final_conv.weight", "final_conv.bias"}, [=](const RunCtx& cx, cudaStream_t st) {
          if (launch_final_conv_bwd(cx.a + in->off, p->dt_act, (const float2*)(cx.f + coef), p->wf(w_off), cx.eps, cx.target,
                                    cx.loss_type, cx.gscale, cx.gscale_dev, cx.g + dpre, p->dt_grad, (double*)(cx.zb + t12),
                                    cx.wg + g_w, cx.wg + g_b, N, H, W, Ci, Co, p->num_sms, st)) *cx.launch_err = 1;
        });
        const size_t coef4 = bb->gn_backward("final_norm.bwd", View::of(in), gnf, t12, -1);
        const Builder::GradW gx = bb->grad_w(in);
        bb->pushb("final_norm.apply", "bwd_affine3", {}, [=](const RunCtx& cx, cudaStream_t st) {
          launch_bwd_affine3(cx.g + dpre, p->dt_grad, Ci, 0, cx.a + in->off, p->dt_act, Ci, 0, (const float4*)(cx.fb + coef4), Ci, 0,
                             nullptr, 0, 0, 0, cx.g + gx.off, p->dt_grad, Ci, 0, gx.accumulate, N, H * W, Ci, st);
        });
        bb->gfree(dpre, bytes);
      });
    }
  }
  // a4.3: one fused GEMV over all blocks' time_mlp.1 (:189-192,215)
  {
    const int rows = p->film_rows;
    const size_t fw = p->walloc((size_t)rows * ted * sizeof(float)), fb = p->walloc((size_t)rows * sizeof(float));
    p->film_f_off = p->falloc((size_t)N * rows * sizeof(float));
    for (auto& blk : b.film_blocks) {
      PackJob jw{}; jw.kind = PACK_COPY; jw.dst = (void*)(fw + (size_t)blk.row0 * ted * sizeof(float)); jw.R = 1; jw.Cc = blk.rows * ted;
      p->add_weight(blk.wname + ".weight", (int64_t)blk.rows * ted, jw);
      PackJob jb{}; jb.kind = PACK_COPY; jb.dst = (void*)(fb + (size_t)blk.row0 * sizeof(float)); jb.R = 1; jb.Cc = blk.rows;
      p->add_weight(blk.wname + ".bias", blk.rows, jb);
    }
    Op& f = p->ops[film_op_index];
    f.bytes = (double)rows * ted * 4; f.flops = 2.0 * N * rows * ted;
    f.ref_bytes = f.bytes;
    p->total_bytes += f.bytes; p->total_flops += f.flops; p->total_bytes_ref += f.bytes;
    f.run = [=](const RunCtx& cx, cudaStream_t st) {
      launch_film((const float*)(cx.f + p->silu_f_off), p->wf(fw), p->wf(fb), (float*)(cx.f + p->film_f_off), N, rows, ted, st);
    };
    if (p->train) {
      p->dfilm_fb_off = p->fballoc((size_t)N * rows * sizeof(float));
      Builder* bb = &b;
      std::vector<Builder::FilmBlock> blocks = b.film_blocks;
      b.tape[film_tape_index] = [=]() {
        // all blocks have written their d scale / d shift rows: weight gradients of every time_mlp.1, then d silu(t_emb)
        std::vector<std::string> writes;
        std::vector<std::pair<size_t, size_t>> offs;
        for (auto& blk : blocks) {
          writes.push_back(blk.wname + ".weight"); writes.push_back(blk.wname + ".bias");
          offs.push_back({bb->wg(blk.wname + ".weight"), bb->wg(blk.wname + ".bias")});
        }
        bb->pushb("film.wgrad", "outer_sum", writes, [=](const RunCtx& cx, cudaStream_t st) {
          for (size_t i = 0; i < blocks.size(); ++i)
            launch_outer_sum((const float*)(cx.fb + p->dfilm_fb_off) + blocks[i].row0, rows, (const float*)(cx.f + p->silu_f_off), ted,
                             cx.wg + offs[i].first, cx.wg + offs[i].second, N, blocks[i].rows, ted, st);
        });
        p->bwd_ops.back().launches = (int)blocks.size();
        bb->pushb("film.dgrad", "film_bwd_input", {}, [=](const RunCtx& cx, cudaStream_t st) {
          launch_film_bwd_input((const float*)(cx.fb + p->dfilm_fb_off), p->wf(fw), (float*)(cx.zb + p->dst_zb_off), N, rows, ted, st);
        });
      };
    }
  }
  if (p->train) {
    // replay the tape backwards: emits p->bwd_ops in execution order and sizes regions G / ZB / FB
    for (auto it = b.tape.rbegin(); it != b.tape.rend(); ++it)
      if (*it) (*it)();
  }
  return 0;
}

void finish_layout(lcm_plan* p) {
  p->ws_bytes = align_up(p->z_bytes, 1024) + align_up(p->f_bytes, 1024) + align_up(p->pool.top, 1024);
  if (p->train) {
    p->zb_off_ws = p->ws_bytes; p->ws_bytes += align_up(p->zb_bytes, 1024);
    p->fb_off_ws = p->ws_bytes; p->ws_bytes += align_up(p->fb_bytes, 1024);
    p->g_off_ws = p->ws_bytes; p->ws_bytes += align_up(p->gpool.top, 1024);
    p->wg_off_ws = p->ws_bytes; p->ws_bytes += align_up(p->wg_elems * sizeof(float), 1024);
  }
}

int check_ready(const lcm_plan* p) {
  if (!p->wbase) return fail(LCM_ERR_INVALID, "plan was created with LCM_FLAG_DRY: it describes the layout only");
  for (auto& kv : p->weights)
    if (!kv.second.set) return fail(LCM_ERR_MISSING_WEIGHT, "weight '%s' has not been set", kv.first.c_str());
  return 0;
}

RunCtx make_ctx(const lcm_plan* p, void* workspace) {
  RunCtx c{};
  c.z = (char*)workspace;
  c.f = c.z + align_up(p->z_bytes, 1024);
  c.a = c.f + align_up(p->f_bytes, 1024);
  if (p->train) {
    c.zb = c.z + p->zb_off_ws;
    c.fb = c.z + p->fb_off_ws;
    c.g = c.z + p->g_off_ws;
    c.wg = (float*)(c.z + p->wg_off_ws);
  }
  return c;
}

int run_forward(lcm_plan* p, RunCtx& c, cudaStream_t st, lcm_op_profile* rec, int cap) {
  int launch_err = 0;
  c.launch_err = &launch_err;
  CUDA_TRY(cudaMemsetAsync(c.z, 0, p->z_bytes, st));
  std::vector<cudaEvent_t> ev;
  if (rec) {
    ev.resize(p->ops.size() + 1);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(ev[0], st));
  }
  // LCM_DIAG_SKIP=<kernel label>[,<label>...]: timing experiment only (results are then garbage) — how much of the
  // step a group of launches costs inside the graph
  static const char* diag_skip = getenv("LCM_DIAG_SKIP");
  for (size_t i = 0; i < p->ops.size(); ++i) {
    if (!(diag_skip && strstr(diag_skip, p->ops[i].kernel.c_str()))) p->ops[i].run(c, st);
    if (rec) CUDA_TRY(cudaEventRecord(ev[i + 1], st));
  }
  CUDA_TRY(cudaGetLastError());
  if (launch_err) return fail(LCM_ERR_INVALID, "a tcgen05 kernel rejected its shape (unsupported configuration)");
  if (rec) {
    CUDA_TRY(cudaStreamSynchronize(st));
    for (size_t i = 0; i < p->ops.size() && (int)i < cap; ++i) {
      lcm_op_profile& r = rec[i];
      memset(&r, 0, sizeof(r));
      snprintf(r.name, sizeof(r.name), "%s", p->ops[i].name.c_str());
      snprintf(r.kernel, sizeof(r.kernel), "%s", p->ops[i].kernel.c_str());
      cudaEventElapsedTime(&r.ms, ev[i], ev[i + 1]);
      r.bytes = p->ops[i].bytes; r.flops = p->ops[i].flops; r.ref_bytes = p->ops[i].ref_bytes;
    }
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return 0;
}

int fill_inputs(lcm_plan* plan, RunCtx& c, const float* xa, int ca, int64_t sa, const float* xb, int cb, int64_t sb) {
  if (!xa || ca < 1 || cb < 0 || ca + cb != plan->cfg.in_channels || (cb > 0 && !xb))
    return fail(LCM_ERR_INVALID, "input channels %d+%d do not match in_channels=%d", ca, cb, plan->cfg.in_channels);
  c.xa = xa; c.ca = ca; c.sa = sa; c.xb = xb; c.cb = cb; c.sb = sb;
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* lcm_last_error(void) { return g_err.c_str(); }
int lcm_version(void) { return 1; }

int lcm_plan_create(const lcm_unet_config* cfg, int batch, int height, int width, int precision, uint32_t flags,
                    int device, lcm_plan** out) {
  if (!cfg || !out) return fail(LCM_ERR_INVALID, "null argument");
  *out = nullptr;
  if (batch < 1 || height < 1 || width < 1) return fail(LCM_ERR_INVALID, "bad batch/size");
  if (cfg->num_levels < 1 || cfg->num_levels > LCM_MAX_LEVELS) return fail(LCM_ERR_INVALID, "bad num_levels");
  const int div = 1 << (cfg->num_levels - 1);
  if (height % div || width % div) return fail(LCM_ERR_INVALID, "height/width must be divisible by %d", div);
  if (precision != LCM_PREC_FP32 && precision != LCM_PREC_BF16) return fail(LCM_ERR_INVALID, "Unknown precision: %d", precision);
  if (cfg->out_channels < 1 || cfg->out_channels > 4) return fail(LCM_ERR_INVALID, "out_channels must be 1..4");
  if (cfg->base_channels % 16 || cfg->base_channels < 16 || cfg->base_channels > 64)
    return fail(LCM_ERR_INVALID, "base_channels must be 16/32/48/64");
  if (cfg->in_channels < 1 || cfg->in_channels > 8) return fail(LCM_ERR_INVALID, "in_channels must be 1..8");
  if (cfg->time_embed_dim < 1 || cfg->time_embed_dim > 512) return fail(LCM_ERR_INVALID, "time_embed_dim must be 1..512");
  if (cfg->expansion_ratio < 1 || cfg->num_res_blocks < 1 || cfg->num_attention_heads < 1)
    return fail(LCM_ERR_INVALID, "bad block configuration");
  std::unique_ptr<lcm_plan> p(new lcm_plan());
  p->cfg = *cfg; p->N = batch; p->H = height; p->W = width; p->prec = precision; p->device = device; p->flags = flags;
  p->bf16 = precision == LCM_PREC_BF16;
  p->tc = p->bf16 && !(flags & LCM_FLAG_SIMT_GEMM);
  p->taps = (flags & LCM_FLAG_TAPS) != 0;
  p->train = (flags & LCM_FLAG_TRAIN) != 0;
  if (p->train && cfg->standard_attention)
    return fail(LCM_ERR_INVALID, "training plans implement the presets' linear attention; StandardAttention is inference-only");
  if (p->train && p->bf16 && !p->tc) return fail(LCM_ERR_INVALID, "training plans are fp32 or bf16 (tensor-core); LCM_FLAG_SIMT_GEMM is inference-only");
  p->pool.reuse = !p->taps && !p->train;     // the backward pass reads every forward tensor
  p->gpool.reuse = !p->taps;
  p->esz = p->bf16 ? 2 : 4;
  p->dt_act = p->bf16 ? DT_BF16 : DT_F32;
  p->dt_hid = p->tc ? DT_F16 : p->dt_act;
  p->dt_grad = p->dt_act;                    // bf16 gradients keep fp32's exponent range (d loss / d eps is O(1/numel))
  p->gsz = p->esz;
  // GroupNorm validity: the reference raises ValueError at construction when C % min(32,C) != 0 (F1)
  {
    std::vector<int> widths, cs;
    for (int i = 0; i < cfg->num_levels; ++i) widths.push_back(cfg->base_channels * cfg->channel_multipliers[i]);
    int ci = widths[0];
    for (int i = 0; i < cfg->num_levels; ++i) { cs.push_back(ci); cs.push_back(widths[i]); ci = widths[i]; }
    for (int i = 0; i < cfg->num_levels; ++i) { int co = widths[cfg->num_levels - 1 - i]; cs.push_back(ci + co); cs.push_back(co); ci = co; }
    for (int C : cs) {
      if (C % 16) return fail(LCM_ERR_INVALID, "channel count %d is not a multiple of 16", C);
      if (p->groups(C) < 0 || p->groups(C * cfg->expansion_ratio) < 0)
        return fail(LCM_ERR_INVALID, "num_channels (%d) must be divisible by num_groups (32)", p->groups(C) < 0 ? C : C * cfg->expansion_ratio);
      if ((C * cfg->expansion_ratio) % 32) return fail(LCM_ERR_INVALID, "hidden width %d is not a multiple of 32", C * cfg->expansion_ratio);
    }
  }
  if (flags & LCM_FLAG_DRY) {
    // host-side description only (op lists, workspace / gradient layout); no device is touched and nothing can run
    int rc = build_plan(p.get());
    if (rc) return rc;
    finish_layout(p.get());
    *out = p.release();
    return 0;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(LCM_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  DeviceGuard guard(device);   // the caller's current device is restored on every return path
  if (!guard.ok) return fail(LCM_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(LCM_ERR_CUDA, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  p->num_sms = prop.multiProcessorCount;

  int rc = build_plan(p.get());
  if (rc) return rc;
  finish_layout(p.get());
  CUDA_TRY(cudaMalloc((void**)&p->wbase, p->wbytes ? p->wbytes : 256));
  CUDA_TRY(cudaMemset(p->wbase, 0, p->wbytes ? p->wbytes : 256));
  for (PackJob j : p->identity_jobs) {   // residual identities of the project GEMMs
    j.dst = p->wbase + (size_t)j.dst;
    launch_pack(j, nullptr, 0);
  }
  CUDA_TRY(cudaDeviceSynchronize());
  *out = p.release();
  return 0;
}

void lcm_plan_destroy(lcm_plan* plan) {
  if (!plan) return;
  if (plan->wbase) cudaFree(plan->wbase);
  delete plan;
}

size_t lcm_plan_workspace_bytes(const lcm_plan* plan) { return plan ? plan->ws_bytes : 0; }
int lcm_plan_num_weights(const lcm_plan* plan) { return plan ? (int)plan->weight_order.size() : 0; }

int lcm_plan_weight_info(const lcm_plan* plan, int index, const char** name, int64_t* numel) {
  if (!plan || index < 0 || index >= (int)plan->weight_order.size()) return fail(LCM_ERR_INVALID, "bad weight index");
  const std::string& n = plan->weight_order[index];
  if (name) *name = n.c_str();
  if (numel) *numel = plan->weights.at(n).numel;
  return 0;
}

int lcm_plan_set_weight(lcm_plan* plan, const char* name, const float* dev_values, int64_t numel, void* stream) {
  if (!plan || !name || !dev_values) return fail(LCM_ERR_INVALID, "null argument");
  if (!plan->wbase) return fail(LCM_ERR_INVALID, "plan was created with LCM_FLAG_DRY: it describes the layout only");
  auto it = plan->weights.find(name);
  if (it == plan->weights.end()) return fail(LCM_ERR_UNKNOWN_WEIGHT, "unknown weight '%s'", name);
  if (it->second.numel != numel)
    return fail(LCM_ERR_INVALID, "weight '%s': expected %lld elements, got %lld", name, (long long)it->second.numel, (long long)numel);
  for (PackJob j : it->second.jobs) {
    j.dst = plan->wbase + (size_t)j.dst;
    launch_pack(j, dev_values, (cudaStream_t)stream);
  }
  CUDA_TRY(cudaGetLastError());
  it->second.set = true;
  return 0;
}

int lcm_unet_forward(lcm_plan* plan, const float* xa_dev, int ca, int64_t xa_batch_stride, const float* xb_dev, int cb,
                     int64_t xb_batch_stride, const int64_t* t_dev, float* eps_dev, void* workspace, void* stream) {
  if (!plan || !t_dev || !eps_dev || !workspace) return fail(LCM_ERR_INVALID, "null argument");
  int rc = check_ready(plan);
  if (rc) return rc;
  RunCtx c = make_ctx(plan, workspace);
  rc = fill_inputs(plan, c, xa_dev, ca, xa_batch_stride, xb_dev, cb, xb_batch_stride);
  if (rc) return rc;
  c.t_dev = (const long long*)t_dev;
  c.eps = eps_dev;
  c.step.enabled = 0;
  return run_forward(plan, c, (cudaStream_t)stream, nullptr, 0);
}

int lcm_plan_profile_forward(lcm_plan* plan, const float* xa_dev, int ca, int64_t xa_batch_stride, const float* xb_dev,
                             int cb, int64_t xb_batch_stride, const int64_t* t_dev, float* eps_dev, void* workspace,
                             void* stream, lcm_op_profile* records, int cap) {
  if (!plan || !t_dev || !eps_dev || !workspace || !records) return fail(LCM_ERR_INVALID, "null argument");
  int rc = check_ready(plan);
  if (rc) return rc;
  RunCtx c = make_ctx(plan, workspace);
  rc = fill_inputs(plan, c, xa_dev, ca, xa_batch_stride, xb_dev, cb, xb_batch_stride);
  if (rc) return rc;
  c.t_dev = (const long long*)t_dev;
  c.eps = eps_dev;
  c.step.enabled = 0;
  rc = run_forward(plan, c, (cudaStream_t)stream, records, cap);
  return rc ? rc : (int)plan->ops.size();
}

int lcm_enhance(lcm_plan* plan, const float* cond_dev, float* latents_dev, const float* noises_dev, int steps,
                const int64_t* timesteps, const float* coef, float* out_dev, float* trace_dev, void* workspace,
                void* stream) {
  if (!plan || !cond_dev || !latents_dev || !timesteps || !coef || !out_dev || !workspace)
    return fail(LCM_ERR_INVALID, "null argument");
  if (steps < 1) return fail(LCM_ERR_INVALID, "steps must be >= 1");
  if (steps > 1 && !noises_dev) return fail(LCM_ERR_INVALID, "noises_dev is required for steps > 1");
  if (plan->cfg.in_channels != 2 * plan->cfg.out_channels)
    return fail(LCM_ERR_INVALID, "enhance needs concat conditioning: in_channels == 2*out_channels");
  int rc = check_ready(plan);
  if (rc) return rc;
  const int C = plan->cfg.out_channels;
  const long long per = (long long)C * plan->H * plan->W, all = per * plan->N;
  RunCtx c = make_ctx(plan, workspace);
  rc = fill_inputs(plan, c, latents_dev, C, per, cond_dev, C, per);
  if (rc) return rc;
  for (int i = 0; i < steps; ++i) {
    const bool last = i == steps - 1;
    c.t_dev = nullptr;
    c.t_scalar = timesteps[i];
    c.eps = nullptr;
    c.step.enabled = 1;
    c.step.noise = last ? nullptr : noises_dev + (long long)i * all;
    c.step.latents = latents_dev;
    c.step.clamped = last ? out_dev : nullptr;
    c.step.trace = trace_dev ? trace_dev + (long long)i * all : nullptr;
    c.step.sb_t = coef[4 * i + 0]; c.step.sa_t = coef[4 * i + 1]; c.step.sa_p = coef[4 * i + 2]; c.step.sb_p = coef[4 * i + 3];
    rc = run_forward(plan, c, (cudaStream_t)stream, nullptr, 0);
    if (rc) return rc;
  }
  return 0;
}

int lcm_enhance_add(lcm_plan* plan, const float* cond_feat_dev, float* latents_dev, const float* noises_dev, int steps,
                    const int64_t* timesteps, const float* coef, float* out_dev, float* trace_dev, void* workspace, void* stream) {
  if (!plan || !cond_feat_dev || !latents_dev || !timesteps || !coef || !out_dev || !workspace)
    return fail(LCM_ERR_INVALID, "null argument");
  if (steps < 1) return fail(LCM_ERR_INVALID, "steps must be >= 1");
  if (steps > 1 && !noises_dev) return fail(LCM_ERR_INVALID, "noises_dev is required for steps > 1");
  if (plan->cfg.in_channels != plan->cfg.out_channels)
    return fail(LCM_ERR_INVALID, "enhance_add needs add conditioning: in_channels == out_channels");
  int rc = check_ready(plan);
  if (rc) return rc;
  const int C = plan->cfg.out_channels;
  const long long per = (long long)C * plan->H * plan->W, all = per * plan->N;
  RunCtx c = make_ctx(plan, workspace);
  float* tmp = (float*)(c.f + plan->addtmp_f_off);
  rc = fill_inputs(plan, c, tmp, C, per, nullptr, 0, 0);
  if (rc) return rc;
  for (int i = 0; i < steps; ++i) {
    const bool last = i == steps - 1;
    launch_add_f32(latents_dev, cond_feat_dev, tmp, all, (cudaStream_t)stream);   // model_input = latents + condition_feat (:223-225)
    c.t_dev = nullptr;
    c.t_scalar = timesteps[i];
    c.eps = nullptr;
    c.step.enabled = 1;
    c.step.noise = last ? nullptr : noises_dev + (long long)i * all;
    c.step.latents = latents_dev;
    c.step.clamped = last ? out_dev : nullptr;
    c.step.trace = trace_dev ? trace_dev + (long long)i * all : nullptr;
    c.step.sb_t = coef[4 * i + 0]; c.step.sa_t = coef[4 * i + 1]; c.step.sa_p = coef[4 * i + 2]; c.step.sb_p = coef[4 * i + 3];
    rc = run_forward(plan, c, (cudaStream_t)stream, nullptr, 0);
    if (rc) return rc;
  }
  return 0;
}

size_t lcm_condition_encode_scratch_bytes(int batch, int height, int width, int hidden) {
  const size_t px = (size_t)batch * height * width;
  return align_up(px * hidden * sizeof(float), 1024) + align_up((size_t)batch * hidden * (2 * sizeof(double) + sizeof(float2)), 1024) +
         align_up((size_t)9 * 3 * hidden * sizeof(float) * 2, 1024) + 4096;
}

int lcm_condition_encode(const float* low_dev, const float* w1_dev, const float* b1_dev, const float* w2_dev, const float* b2_dev,
                         float* out_dev, int batch, int height, int width, int hidden, void* scratch_dev, void* stream) {
  if (!low_dev || !w1_dev || !b1_dev || !w2_dev || !b2_dev || !out_dev || !scratch_dev) return fail(LCM_ERR_INVALID, "null argument");
  if (hidden % 8 || hidden < 8 || hidden > 64 || batch < 1 || height < 1 || width < 1) return fail(LCM_ERR_INVALID, "bad condition encoder shape");
  cudaStream_t st = (cudaStream_t)stream;
  char* s = (char*)scratch_dev;
  const size_t px = (size_t)batch * height * width;
  float* hid = (float*)s; s += align_up(px * hidden * sizeof(float), 1024);
  double* stats = (double*)s; s += (size_t)batch * hidden * 2 * sizeof(double);
  float2* coef = (float2*)s; s = (char*)scratch_dev + align_up(px * hidden * sizeof(float), 1024) + align_up((size_t)batch * hidden * (2 * sizeof(double) + sizeof(float2)), 1024);
  float* w1 = (float*)s; s += (size_t)9 * 3 * hidden * sizeof(float);
  float* w2 = (float*)s;
  { PackJob j{}; j.kind = PACK_CONV3_KN; j.dst = w1; j.R = hidden; j.Ci = 3; launch_pack(j, w1_dev, st); }
  { PackJob j{}; j.kind = PACK_CONV3_KN; j.dst = w2; j.R = 3; j.Ci = hidden; launch_pack(j, w2_dev, st); }
  CUDA_TRY(cudaMemsetAsync(stats, 0, (size_t)batch * hidden * 2 * sizeof(double), st));
  launch_fill_float2(coef, make_float2(1.f, 0.f), (long long)batch * hidden, st);
  // Conv2d(3, hidden, 3, padding=1) -> SiLU -> Conv2d(hidden, 3, 3, padding=1)   (low_light_diffusion.py:108-113), fp32
  launch_init_conv(low_dev, 3, (long long)3 * height * width, nullptr, 0, 0, w1, b1_dev, hid, stats, batch, height, width, hidden, 0, st);
  FinalStep none{};
  launch_final_conv(hid, coef, w2, b2_dev, out_dev, none, batch, height, width, hidden, 3, 0, st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_scheduler_step(const float* model_out_dev, const float* sample_dev, const float* noise_dev, float* prev_dev,
                       float* x0_dev, int64_t numel, int prediction, float sqrt_beta_t, float sqrt_alpha_t,
                       float sqrt_alpha_prev, float sqrt_beta_prev, void* stream) {
  if (!model_out_dev || !sample_dev || !prev_dev) return fail(LCM_ERR_INVALID, "null argument");
  if (prediction != 0 && prediction != 1) return fail(LCM_ERR_INVALID, "Unknown prediction type: %d", prediction);
  launch_lcm_step(model_out_dev, sample_dev, noise_dev, prev_dev, x0_dev, numel, prediction, sqrt_beta_t, sqrt_alpha_t,
                  sqrt_alpha_prev, sqrt_beta_prev, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_scheduler_mix(const float* a_dev, const float* b_dev, const int64_t* t_dev, const float* abar_dev, float* out_dev,
                      int batch, int64_t per_sample, int velocity, void* stream) {
  if (!a_dev || !b_dev || !t_dev || !abar_dev || !out_dev) return fail(LCM_ERR_INVALID, "null argument");
  launch_lcm_mix(a_dev, b_dev, (const long long*)t_dev, abar_dev, out_dev, batch, per_sample, velocity, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_ddim_step(const float* x_t_dev, const float* eps_dev, const int64_t* t_dev, const int64_t* t_next_dev, const float* abar_dev,
                  float* x_next_dev, int batch, int64_t per_sample, void* stream) {
  if (!x_t_dev || !eps_dev || !t_dev || !t_next_dev || !abar_dev || !x_next_dev) return fail(LCM_ERR_INVALID, "null argument");
  launch_ddim_step(x_t_dev, eps_dev, (const long long*)t_dev, (const long long*)t_next_dev, abar_dev, x_next_dev, batch, per_sample,
                   (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int lcm_consistency_loss(const float* x_t_dev, const float* eps_student_dev, const int64_t* t_dev, const float* x_next_dev,
                         const float* eps_target_dev, const int64_t* t_next_dev, const float* abar_dev, double* loss_dev,
                         float* d_eps_student_dev, int batch, int64_t per_sample, void* stream) {
  if (!x_t_dev || !eps_student_dev || !t_dev || !x_next_dev || !eps_target_dev || !t_next_dev || !abar_dev || !loss_dev ||
      !d_eps_student_dev)
    return fail(LCM_ERR_INVALID, "null argument");
  launch_consistency_loss(x_t_dev, eps_student_dev, (const long long*)t_dev, x_next_dev, eps_target_dev, (const long long*)t_next_dev,
                          abar_dev, loss_dev, d_eps_student_dev, batch, per_sample, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_image_preprocess_u8(const uint8_t* hwc_dev, float* nchw_dev, int batch, int height, int width, void* stream) {
  if (!hwc_dev || !nchw_dev || batch < 1 || height < 1 || width < 1) return fail(LCM_ERR_INVALID, "null argument or empty image");
  launch_image_pre_u8(hwc_dev, nchw_dev, batch, height, width, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int lcm_image_postprocess_u8(const float* nchw_dev, uint8_t* hwc_dev, int batch, int height, int width, void* stream) {
  if (!hwc_dev || !nchw_dev || batch < 1 || height < 1 || width < 1) return fail(LCM_ERR_INVALID, "null argument or empty image");
  launch_image_post_u8(nchw_dev, hwc_dev, batch, height, width, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_image_resize_u8(const uint8_t* src_hwc_dev, int batch, int src_h, int src_w, uint8_t* dst_hwc_dev, int dst_h, int dst_w,
                        void* stream) {
  if (!src_hwc_dev || !dst_hwc_dev || batch < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1)
    return fail(LCM_ERR_INVALID, "null argument or empty image");
  launch_image_resize_u8(src_hwc_dev, batch, src_h, src_w, dst_hwc_dev, dst_h, dst_w, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- training step ----------------------------------------------------------------------------------
int lcm_train_backward_op_info(const lcm_plan* plan, int index, const char** name, const char** kernel) {
  if (!plan || !plan->train || index < 0 || index >= (int)plan->bwd_ops.size()) return fail(LCM_ERR_INVALID, "bad op index");
  if (name) *name = plan->bwd_ops[index].name.c_str();
  if (kernel) *kernel = plan->bwd_ops[index].kernel.c_str();
  return 0;
}
int lcm_train_num_backward_ops(const lcm_plan* plan) { return plan && plan->train ? (int)plan->bwd_ops.size() : 0; }
int64_t lcm_train_grad_elems(const lcm_plan* plan) { return plan && plan->train ? (int64_t)plan->wg_elems : 0; }
size_t lcm_train_grad_offset_bytes(const lcm_plan* plan) { return plan && plan->train ? plan->wg_off_ws : 0; }

int lcm_train_grad_info(const lcm_plan* plan, int index, const char** name, int64_t* offset, int64_t* numel, int* ready_after_op) {
  if (!plan || !plan->train || index < 0 || index >= (int)plan->weight_order.size()) return fail(LCM_ERR_INVALID, "bad gradient index");
  const std::string& n = plan->weight_order[index];
  if (name) *name = n.c_str();
  if (offset) *offset = (int64_t)plan->wgrad_off.at(n);
  if (numel) *numel = plan->weights.at(n).numel;
  if (ready_after_op) {
    auto it = plan->wgrad_last_op.find(n);
    *ready_after_op = it == plan->wgrad_last_op.end() ? -1 : it->second;
  }
  return 0;
}

int lcm_train_backward(lcm_plan* plan, const float* xa_dev, int ca, int64_t xa_batch_stride, const float* xb_dev, int cb,
                       int64_t xb_batch_stride, const int64_t* t_dev, const float* eps_dev, const float* target_dev, int loss_type,
                       float grad_scale, const float* grad_scale_dev, int op_begin, int op_end, void* workspace, void* stream) {
  if (!plan || !t_dev || !eps_dev || !target_dev || !workspace) return fail(LCM_ERR_INVALID, "null argument");
  if (!plan->train) return fail(LCM_ERR_INVALID, "plan was created without LCM_FLAG_TRAIN");
  if (loss_type < 0 || loss_type > 3) return fail(LCM_ERR_INVALID, "Unknown loss type: %d", loss_type);
  const int nops = (int)plan->bwd_ops.size();
  if (op_end < 0 || op_end > nops) op_end = nops;
  if (op_begin < 0 || op_begin > op_end) return fail(LCM_ERR_INVALID, "bad backward op range");
  int rc = check_ready(plan);
  if (rc) return rc;
  RunCtx c = make_ctx(plan, workspace);
  rc = fill_inputs(plan, c, xa_dev, ca, xa_batch_stride, xb_dev, cb, xb_batch_stride);
  if (rc) return rc;
  c.t_dev = (const long long*)t_dev;
  c.eps = const_cast<float*>(eps_dev);
  c.target = target_dev; c.loss_type = loss_type; c.gscale = grad_scale; c.gscale_dev = grad_scale_dev;
  cudaStream_t st = (cudaStream_t)stream;
  int launch_err = 0;
  c.launch_err = &launch_err;
  if (op_begin == 0) {
    CUDA_TRY(cudaMemsetAsync(c.zb, 0, plan->zb_bytes, st));
    CUDA_TRY(cudaMemsetAsync(c.wg, 0, plan->wg_elems * sizeof(float), st));
  }
  for (int i = op_begin; i < op_end; ++i) plan->bwd_ops[i].run(c, st);
  CUDA_TRY(cudaGetLastError());
  if (launch_err) return fail(LCM_ERR_INVALID, "a backward kernel rejected its shape (unsupported configuration)");
  return 0;
}

int lcm_train_loss(const float* eps_dev, const float* target_dev, int64_t numel, int loss_type, double* loss_dev, void* stream) {
  if (!eps_dev || !target_dev || !loss_dev || numel < 1) return fail(LCM_ERR_INVALID, "null argument");
  if (loss_type < 0 || loss_type > 2) return fail(LCM_ERR_INVALID, "Unknown loss type: %d", loss_type);
  launch_loss(eps_dev, target_dev, numel, loss_type, loss_dev, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_plan_set_weights_flat(lcm_plan* plan, const float* flat_dev, void* stream) {
  if (!plan || !flat_dev) return fail(LCM_ERR_INVALID, "null argument");
  if (!plan->train) return fail(LCM_ERR_INVALID, "plan was created without LCM_FLAG_TRAIN");
  if (!plan->wbase) return fail(LCM_ERR_INVALID, "plan was created with LCM_FLAG_DRY: it describes the layout only");
  for (const std::string& name : plan->weight_order) {
    WeightSlot& slot = plan->weights.at(name);
    const float* src = flat_dev + plan->wgrad_off.at(name);
    for (PackJob j : slot.jobs) {
      j.dst = plan->wbase + (size_t)j.dst;
      launch_pack(j, src, (cudaStream_t)stream);
    }
    slot.set = true;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_grad_sumsq(const float* grads_dev, int64_t n, double* out_dev, void* stream) {
  if (!grads_dev || !out_dev || n < 1) return fail(LCM_ERR_INVALID, "null argument");
  launch_sumsq(grads_dev, n, out_dev, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_adamw_ema_step(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev, float* ema_dev,
                       int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float ema_decay,
                       const double* grad_sumsq_dev, float grad_div, float max_norm, void* stream) {
  if (!params_dev || !grads_dev || !exp_avg_dev || !exp_avg_sq_dev || n < 1 || step < 1 || grad_div <= 0.f)
    return fail(LCM_ERR_INVALID, "bad optimizer arguments");
  if (max_norm > 0.f && !grad_sumsq_dev) return fail(LCM_ERR_INVALID, "clipping needs the gradient sum of squares");
  launch_adamw_ema(params_dev, grads_dev, exp_avg_dev, exp_avg_sq_dev, ema_dev, n, lr, beta1, beta2, eps, weight_decay, step,
                   ema_decay, grad_sumsq_dev, grad_div, max_norm, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_train_read_grad_tap(lcm_plan* plan, const char* name, float* out_nchw_dev, void* workspace, void* stream) {
  if (!plan || !name || !out_nchw_dev || !workspace) return fail(LCM_ERR_INVALID, "null argument");
  auto it = plan->gtap_map.find(name);
  if (it == plan->gtap_map.end()) return fail(LCM_ERR_INVALID, "unknown gradient tap '%s' (needs LCM_FLAG_TRAIN | LCM_FLAG_TAPS)", name);
  RunCtx c = make_ctx(plan, workspace);
  const TensorP& t = it->second;
  launch_nhwc_to_nchw(c.g + t->g_off, out_nchw_dev, plan->N, t->H, t->W, t->C, plan->dt_grad == DT_BF16 ? 1 : 0, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_plan_num_taps(const lcm_plan* plan) { return plan ? (int)plan->tap_order.size() : 0; }
int lcm_plan_tap_info(const lcm_plan* plan, int index, const char** name, int* channels, int* height, int* width) {
  if (!plan || index < 0 || index >= (int)plan->tap_order.size()) return fail(LCM_ERR_INVALID, "bad tap index");
  const std::string& n = plan->tap_order[index];
  const TensorP& t = plan->tap_map.at(n);
  if (name) *name = n.c_str();
  if (channels) *channels = t->C;
  if (height) *height = t->H;
  if (width) *width = t->W;
  return 0;
}
int lcm_plan_read_tap(lcm_plan* plan, const char* name, float* out_nchw_dev, void* workspace, void* stream) {
  if (!plan || !name || !out_nchw_dev || !workspace) return fail(LCM_ERR_INVALID, "null argument");
  auto it = plan->tap_map.find(name);
  if (it == plan->tap_map.end()) return fail(LCM_ERR_INVALID, "unknown tap '%s' (plan created without LCM_FLAG_TAPS?)", name);
  RunCtx c = make_ctx(plan, workspace);
  const TensorP& t = it->second;
  launch_nhwc_to_nchw(c.a + t->off, out_nchw_dev, plan->N, t->H, t->W, t->C, t->f16 ? 2 : (plan->bf16 ? 1 : 0), (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lcm_plan_launches_per_forward(const lcm_plan* plan) {   // kernels of this library (the workspace memset is not counted)
  int n = 0;
  if (plan) for (const Op& o : plan->ops) n += o.launches;
  return n;
}
double lcm_plan_algorithmic_bytes(const lcm_plan* plan) { return plan ? plan->total_bytes_ref : 0; }
double lcm_plan_fused_bytes(const lcm_plan* plan) { return plan ? plan->total_bytes : 0; }
double lcm_plan_algorithmic_flops(const lcm_plan* plan) { return plan ? plan->total_flops : 0; }

}  // extern "C"
