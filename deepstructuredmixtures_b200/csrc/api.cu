// libdsmgp.so : C ABI implementation (include/dsmgp.h).  sm_100a only, no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dsmgp.h"
#include "args.h"
#include "potrf2_args.h"
#include "tree_host.h"

using namespace dsm;

static thread_local std::string g_create_error;

#define CUDA_TRY(h, expr)                                                                       \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                            \
      return e_ == cudaErrorMemoryAllocation ? DSMGP_ERR_OOM : DSMGP_ERR_CUDA;                  \
    }                                                                                           \
  } while (0)

namespace {

struct Batch {
  int s0 = 0, s1 = 0, max_nb = 0;
  std::vector<int> cnt;                  // cnt[J]: slots of the batch that own block column J (slots sorted by size)
  int64_t f_doubles = 0, w_doubles = 0, ntiles = 0, trpart_doubles = 0, gpart_doubles = 0;
  int64_t* d_tile_off = nullptr;
  int64_t* d_trpart_off = nullptr;
  int64_t* d_gpart_off = nullptr;
  int2* d_trtri_tasks = nullptr; int n_trtri = 0;
  int4* d_lauum_tasks = nullptr; int n_lauum = 0;
  int4* d_potrf2_tasks = nullptr; int n_potrf2 = 0;     // engine v2: tile tasks in look-ahead order
  int4* d_trtri3_tasks = nullptr; int n_trtri3 = 0;     // inverse: tile tasks by anti-diagonal
  int2* d_solve_tasks = nullptr; int n_solve = 0;       // back-substitution: (slot, J) by level from the bottom
  int64_t* d_flag_off = nullptr; int64_t flag_ints = 0;
  double potrf_flops = 0, gram_bytes = 0;
};

// Process-wide cache of large device buffers.  cudaMalloc / cudaFree of multi-GB arenas cost 10 ms ... 3 s each
// (measured: tools/cold_probe.py), which would dominate building a model from host arrays; freed buffers >= 32 MiB are
// kept and handed to the next handle that asks for a similar size on the same device.  dsmgp_release_cache() returns
// them to the driver.
struct BufCache {
  struct Ent { void* p; size_t bytes; int dev; };
  std::vector<Ent> ents;
  std::mutex mu;
  static constexpr size_t MIN_BYTES = size_t(32) << 20;
  size_t cached_bytes(int dev) {
    std::lock_guard<std::mutex> g(mu);
    size_t t = 0;
    for (auto& e : ents) if (e.dev == dev) t += e.bytes;
    return t;
  }
  void* take(size_t bytes, int dev, size_t* got) {
    std::lock_guard<std::mutex> g(mu);
    int best = -1;
    for (int i = 0; i < (int)ents.size(); i++)
      if (ents[i].dev == dev && ents[i].bytes >= bytes && ents[i].bytes <= bytes + bytes / 4 + (size_t(64) << 20) &&
          (best < 0 || ents[i].bytes < ents[best].bytes)) best = i;
    if (best < 0) return nullptr;
    void* p = ents[best].p;
    *got = ents[best].bytes;
    ents.erase(ents.begin() + best);
    return p;
  }
  bool give(void* p, size_t bytes, int dev) {
    if (bytes < MIN_BYTES) return false;
    std::lock_guard<std::mutex> g(mu);
    if (ents.size() >= 64) return false;
    ents.push_back({p, bytes, dev});
    return true;
  }
  void release_all() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& e : ents) { int cur = 0; cudaGetDevice(&cur); cudaSetDevice(e.dev); cudaFree(e.p); cudaSetDevice(cur); }
    ents.clear();
  }
};
static BufCache g_cache;

template <typename T>
struct DevBuf {
  T* p = nullptr; size_t n = 0; size_t cap_bytes = 0; int dev = 0;
  cudaError_t alloc(size_t count) {
    free();
    n = count;
    if (count == 0) return cudaSuccess;
    const size_t bytes = count * sizeof(T);
    cudaGetDevice(&dev);
    if (bytes >= BufCache::MIN_BYTES) {
      if (void* q = g_cache.take(bytes, dev, &cap_bytes)) { p = static_cast<T*>(q); return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {          // make room: drop the cache and retry once
      cudaGetLastError();
      g_cache.release_all();
      e = cudaMalloc(&p, bytes);
    }
    cap_bytes = bytes;
    return e;
  }
  // grow-only scratch: keeps the allocation across calls (cudaMalloc / cudaFree of GBs cost 10-100 ms per call)
  cudaError_t ensure(size_t count) {
    if (count <= n && p != nullptr) return cudaSuccess;
    return alloc(count + count / 8);
  }
  void free() {
    if (p) {
      if (!g_cache.give(p, cap_bytes, dev)) cudaFree(p);
    }
    p = nullptr; n = 0; cap_bytes = 0;
  }
};

}  // namespace

struct dsmgp_handle {
  int64_t N = 0, D = 0, L = 0;
  int nk = 0;
  std::vector<dsmgp_kernel_desc> kernels;
  std::vector<int64_t> koff;      // theta offset per kernel
  std::vector<int32_t> knp;       // nparams per kernel
  int64_t H = 0; int Hmax = 0; int row_width = 0; int pstride = 0;
  std::vector<int64_t> leaf_ptr;
  std::vector<int32_t> leaf_kid;
  std::vector<double> leaf_mean;
  HostTree tree;
  dsmgp_opts opts;
  std::vector<int32_t> owner;
  std::vector<int> slot_leaf;     // slot -> global leaf
  std::vector<int> leaf_slot;     // global leaf -> slot or -1
  std::vector<LeafMeta> meta;     // per slot
  std::vector<Batch> batches;
  std::vector<double> theta_leaf; // L x Hmax
  std::vector<double> h_prm;      // nslots x pstride
  std::vector<double> h_rows;     // L x row_width
  std::vector<double> node_lml;
  std::vector<int32_t> h_info;    // L
  std::vector<double> sum_logw;   // CSR by child_ptr (update!)
  bool have_weights = false;
  bool fitted = false, have_rows = false, have_grad = false, rows_complete = false;
  bool alpha_exact = false;       // alpha from back-substitution (fit path); the gradient path leaves X^T z
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> ev;     // 8 per batch: phase boundaries, always recorded (no extra syncs)
  bool profiling = false;
  dsmgp_timings tm = {};
  // device
  DevBuf<LeafMeta> d_meta;
  DevBuf<double> d_xg, d_y, d_z, d_alpha, d_F, d_W, d_WT, d_prm, d_trpart, d_gpart, d_rows, d_leaf_mean;
  DevBuf<LeafScal> d_scal;
  DevBuf<int> d_counter;
  DevBuf<int> d_flags;
  DevBuf<double> d_ldpart, d_zzpart;
  DevBuf<double> d_apart, d_tpart;   // per-tile partials of the tile-pipelined inverse
  DevBuf<double> p_xt, p_VT, p_mu, p_var, p_part; DevBuf<PredLeaf> p_pl; DevBuf<int2> p_tasks;   // predict scratch (grow-only)
  DevBuf<int4> p_wtasks, p_wcols; DevBuf<int> p_flags;
  DevBuf<int> d_mask; std::vector<int> h_mask; bool use_mask = false;   // per-slot gradient mask (finetune: zero-overlap experts)
  double* pin_multi = nullptr; size_t pin_multi_doubles = 0;             // rows of a multi-theta call [G][L][row_width]
  LeafScal* pin_scal_multi = nullptr; size_t pin_scal_multi_n = 0;       // per-slot scalars of a multi-theta call [G][slots]
  double* pin_rows = nullptr;
  LeafScal* pin_scal = nullptr;
  std::string err;

  ~dsmgp_handle() {
    for (auto& b : batches) {
      cudaFree(b.d_tile_off); cudaFree(b.d_trpart_off); cudaFree(b.d_gpart_off);
      cudaFree(b.d_trtri_tasks); cudaFree(b.d_lauum_tasks); cudaFree(b.d_potrf2_tasks); cudaFree(b.d_trtri3_tasks); cudaFree(b.d_solve_tasks); cudaFree(b.d_flag_off);
    }
    d_meta.free(); d_xg.free(); d_y.free(); d_z.free(); d_alpha.free(); d_F.free(); d_W.free(); d_WT.free();
    d_flags.free(); d_ldpart.free(); d_zzpart.free(); d_apart.free(); d_tpart.free();
    p_xt.free(); p_VT.free(); p_mu.free(); p_var.free(); p_pl.free(); p_tasks.free();
    p_part.free(); p_wtasks.free(); p_wcols.free(); p_flags.free();
    d_prm.free(); d_trpart.free(); d_gpart.free(); d_rows.free(); d_leaf_mean.free(); d_scal.free(); d_counter.free();
    d_mask.free();
    if (pin_multi) cudaFreeHost(pin_multi);
    if (pin_scal_multi) cudaFreeHost(pin_scal_multi);
    if (pin_rows) cudaFreeHost(pin_rows);
    if (pin_scal) cudaFreeHost(pin_scal);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }
};

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
static bool g_attr_done = false;
static cudaError_t engine_attrs() {
  if (g_attr_done) return cudaSuccess;
  cudaError_t e;
  if ((e = init_v2_kernels())) return e;
  g_attr_done = true;
  return cudaSuccess;
}

static int num_sms(int device) {
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

template <typename T>
static cudaError_t upload(T** dptr, const std::vector<T>& v) {
  *dptr = nullptr;
  if (v.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc(dptr, v.size() * sizeof(T));
  if (e) return e;
  return cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static void shard_lpt(int64_t L, const int64_t* leaf_ptr, int world, int32_t* owner) {
  std::vector<int64_t> order(L);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    return (leaf_ptr[a + 1] - leaf_ptr[a]) > (leaf_ptr[b + 1] - leaf_ptr[b]);
  });
  std::vector<double> load(world, 0.0);
  for (int64_t l : order) {
    int best = 0;
    for (int r = 1; r < world; r++) if (load[r] < load[best]) best = r;
    const double n = (double)(leaf_ptr[l + 1] - leaf_ptr[l]);
    load[best] += n * n * n;
    owner[l] = best;
  }
}

static void derive_params(const dsmgp_handle* h, int kid, const double* th, double* prm) {
  const int type = h->kernels[kid].type, np = h->kernels[kid].nparams, nl = np - 2;
  const bool se = (type == DSMGP_ISO_SE || type == DSMGP_ARD_SE);
  const double logs = th[nl], logn = th[nl + 1];
  prm[PRM_V] = se ? std::exp(2.0 * logs) : 1.0;                   // kernels.jl:68,118,181,216
  prm[PRM_S] = se ? std::exp(logs) : 1.0;                         // kernels.jl:69,119,182,217
  prm[PRM_ETA] = std::exp(2.0 * logn);                            // gaussianprocess.jl:39
  prm[PRM_C] = prm[PRM_ETA] + 1e-8;                               // gaussianprocess.jl:94, DeepStructuredMixtures.jl:27
  for (int d = 0; d < nl; d++) {
    const double l = std::exp(th[d]);
    const double l2 = l * l;                                      // kernels.jl:22,41
    prm[PRM_COEF + d] = se ? -0.5 / l2 : 1.0 / l2;                // kernels.jl:78,189
  }
}

static void upload_params(dsmgp_handle* h) {
  const int ns = (int)h->slot_leaf.size();
  for (int s = 0; s < ns; s++) {
    const int l = h->slot_leaf[s];
    derive_params(h, h->leaf_kid[l], &h->theta_leaf[(size_t)l * h->Hmax], &h->h_prm[(size_t)s * h->pstride]);
  }
  if (ns) cudaMemcpyAsync(h->d_prm.p, h->h_prm.data(), h->h_prm.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
}

// ------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------
extern "C" void dsmgp_default_opts(dsmgp_opts* o) {
  memset(o, 0, sizeof(*o));
  o->as_written_grads = 1;
  o->keep_factors = 1;
  o->rank = 0; o->world = 1;
  o->device = -1;
  o->strict_pd = 0;
  o->arena_bytes = 0;
}

extern "C" void dsmgp_release_cache(void) { g_cache.release_all(); }

extern "C" const char* dsmgp_last_error(const dsmgp_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" void dsmgp_destroy(dsmgp_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  delete h;
}

static int32_t plan_and_alloc(dsmgp_handle* h, const double* x, const int64_t* leaf_obs, const double* y_centered) {
  const int64_t L = h->L;
  // local slots sorted by size (descending), ties by leaf number
  std::vector<int> loc;
  for (int64_t l = 0; l < L; l++) if (h->owner[l] == h->opts.rank) loc.push_back((int)l);
  std::stable_sort(loc.begin(), loc.end(), [&](int a, int b) {
    return (h->leaf_ptr[a + 1] - h->leaf_ptr[a]) > (h->leaf_ptr[b + 1] - h->leaf_ptr[b]);
  });
  h->slot_leaf = loc;
  h->leaf_slot.assign(L, -1);
  const int ns = (int)loc.size();
  h->meta.resize(ns);
  h->pstride = PRM_COEF + (int)std::max<int64_t>(h->D, 1);
  int64_t voff = 0, xoff = 0;
  for (int s = 0; s < ns; s++) {
    const int l = loc[s];
    h->leaf_slot[l] = s;
    LeafMeta& m = h->meta[s];
    m.n = (int32_t)(h->leaf_ptr[l + 1] - h->leaf_ptr[l]);
    m.np = (m.n + PAD - 1) / PAD * PAD;
    m.nb = (m.np + BLK - 1) / BLK;
    m.kid = h->leaf_kid[l];
    m.ktype = h->kernels[m.kid].type;
    m.nl = h->kernels[m.kid].nparams - 2;
    m.leaf = l;
    m.nkc = m.np / KC;
    m.voff = voff; voff += m.np;
    m.xoff = xoff; xoff += (int64_t)m.np * h->D;
    m.poff = (int64_t)s * h->pstride;
    m.foff = 0; m.woff = 0;
  }
  // arena budget
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(h, cudaMemGetInfo(&free_b, &total_b));
  free_b += g_cache.cached_bytes(h->device);      // cached buffers are reusable (and released on demand)
  const int64_t fixed = (voff * 4 + xoff) * 8 + (64ll << 20);
  int64_t budget = h->opts.arena_bytes > 0 ? h->opts.arena_bytes : (int64_t)(free_b * 0.85) - fixed;
  // per leaf bytes in a batch: factor np^2 + W/WT 2*nb*BLK^2
  auto leaf_bytes = [&](const LeafMeta& m) { return (tiled_doubles(m.np) + 2ll * m.nb * WBLK_D) * 8; };
  int64_t need_all = 0;
  for (auto& m : h->meta) need_all += leaf_bytes(m);
  if (h->opts.keep_factors && need_all > budget) {
    h->err = "keep_factors=1 but the factors of the local leaves (" + std::to_string(need_all >> 20) +
             " MiB) exceed the arena budget (" + std::to_string(budget >> 20) + " MiB); use keep_factors=0";
    return DSMGP_ERR_OOM;
  }
  // streaming batches: cap the arena so that several batches pipeline well but each one still fills the GPU
  if (!h->opts.keep_factors && h->opts.arena_bytes == 0) budget = std::min<int64_t>(budget, 48ll << 30);
  h->batches.clear();
  {
    int s = 0;
    while (s < ns) {
      Batch b; b.s0 = s;
      int64_t used = 0, fo = 0, wo = 0;
      while (s < ns) {
        const int64_t lb = leaf_bytes(h->meta[s]);
        if (used > 0 && used + lb > budget) break;
        if (lb > budget) { h->err = "one expert's factor exceeds the arena budget"; return DSMGP_ERR_OOM; }
        h->meta[s].foff = fo; fo += tiled_doubles(h->meta[s].np);
        h->meta[s].woff = wo; wo += (int64_t)h->meta[s].nb * WBLK_D;
        used += lb; s++;
      }
      b.s1 = s; b.f_doubles = fo; b.w_doubles = wo;
      h->batches.push_back(b);
    }
  }
  int64_t maxF = 0, maxW = 0, maxTr = 0, maxG = 0, maxFlags = 0;
  const int sms_plan = num_sms(h->device);
  for (auto& b : h->batches) {
    const int nb_s = b.s1 - b.s0;
    b.max_nb = 0;
    for (int s = b.s0; s < b.s1; s++) b.max_nb = std::max(b.max_nb, h->meta[s].nb);
    b.cnt.assign(b.max_nb, 0);
    std::vector<int64_t> tile_off(nb_s + 1, 0), trp(nb_s, 0), gp(nb_s, 0);
    std::vector<int2> tt; std::vector<int4> lt;
    int64_t tro = 0, go = 0;
    for (int s = b.s0; s < b.s1; s++) {
      const LeafMeta& m = h->meta[s];
      for (int J = 0; J < m.nb; J++) b.cnt[J]++;
      const int64_t t64 = m.np / GT;
      tile_off[s - b.s0 + 1] = tile_off[s - b.s0] + t64 * (t64 + 1) / 2;
      trp[s - b.s0] = tro; tro += 2 * m.nb;
      gp[s - b.s0] = go; go += (int64_t)(m.nb * (m.nb + 1) / 2) * m.nl;
      for (int J = 0; J < m.nb; J++) tt.push_back(make_int2(s - b.s0, J));
      int idx = 0;
      const bool leaf_lauum = m.ktype == DSMGP_ISO_SE || m.ktype == DSMGP_ARD_LINEAR || (m.ktype == DSMGP_ARD_SE && !h->opts.as_written_grads);
      for (int I = 0; I < m.nb; I++) for (int J = 0; J <= I; J++, idx++) if (leaf_lauum) lt.push_back(make_int4(s - b.s0, I, J, idx));
      const double n = m.n;
      b.potrf_flops += n * n * n / 3.0 + n * n / 2.0 + n / 6.0;
      b.gram_bytes += 8.0 * (n * (n + 1) / 2.0) + 8.0 * n * h->D;
    }
    // cost-descending task order (dynamic LPT through the atomic task counter)
    std::stable_sort(tt.begin(), tt.end(), [&](const int2& a, const int2& c) {
      const int ra = h->meta[b.s0 + a.x].nb - a.y, rc = h->meta[b.s0 + c.x].nb - c.y; return ra > rc; });
    std::stable_sort(lt.begin(), lt.end(), [&](const int4& a, const int4& c) {
      const int ra = h->meta[b.s0 + a.x].nb - a.y, rc = h->meta[b.s0 + c.x].nb - c.y; return ra > rc; });
    const char* ord_env = getenv("DSMGP_ORDER");           // development A/B: 0 = end together, 1 = stretch
    const bool stretch = ord_env ? (ord_env[0] == '1') : (nb_s * 4 < sms_plan);
    const bool start_together = ord_env && ord_env[0] == '2';     // experiment: no shift at all
    // position of the next diagonal tile inside a level: right behind the first panel tile (1: shortest critical path) or
    // behind all panel tiles of the level (3: in a large batch the first panel tile has then finished and the diagonal
    // task does not sit on an SM waiting for it)
    const char* dg_env = getenv("DSMGP_DIAG_LATE");
    const int diag_grp = (dg_env ? dg_env[0] == '1' : false) ? 3 : 1;
    {   // engine v2 tile tasks: topological order with look-ahead
      struct TK { int s, grp, slot, I, J; };
      std::vector<TK> tk;
      std::vector<int64_t> foff(nb_s, 0);
      int64_t fo = 0;
      for (int s = b.s0; s < b.s1; s++) {
        const LeafMeta& m = h->meta[s];
        const int sl = s - b.s0, shift = b.max_nb - m.nb;
        foff[sl] = fo; fo += (int64_t)m.nb * (m.nb + 1) / 2;
        // level of block column J in the global order.  "end together" (shift) keeps the tail of a throughput-bound
        // batch parallel; "stretch" lets every expert progress proportionally through the whole launch, which spreads the
        // other experts' work evenly along the critical path of the largest one (small shards: multi-GPU strong scaling)
        auto level = [&](int J) { return start_together ? J * 1024 : stretch ? (int)(((int64_t)J * 1024 * b.max_nb) / m.nb) : (J + shift) * 1024; };
        tk.push_back({level(0) - 1, 1, sl, 0, 0});
        for (int J = 0; J + 1 < m.nb; J++) {
          tk.push_back({level(J), 0, sl, J + 1, J});
          tk.push_back({level(J), diag_grp, sl, J + 1, J + 1});
          for (int I = J + 2; I < m.nb; I++) tk.push_back({level(J), 2, sl, I, J});
        }
      }
      std::stable_sort(tk.begin(), tk.end(), [](const TK& a, const TK& c) {
        if (a.s != c.s) return a.s < c.s;
        if (a.grp != c.grp) return a.grp < c.grp;
        if (a.slot != c.slot) return a.slot < c.slot;
        return a.I < c.I; });
      std::vector<int4> pt(tk.size());
      for (size_t i = 0; i < tk.size(); i++) pt[i] = make_int4(tk[i].slot, tk[i].I, tk[i].J, 0);
      b.n_potrf2 = (int)pt.size(); b.flag_ints = fo;
      // inverse tile tasks (I > J): anti-diagonal order (a tile depends on the tiles above it in its column, all on
      // smaller anti-diagonals), experts shifted so that they end together, long tiles first inside a level
      std::vector<TK> iv;
      for (int s = b.s0; s < b.s1; s++) {
        const LeafMeta& m = h->meta[s];
        const int sl = s - b.s0, shift = b.max_nb - m.nb;
        for (int J = 0; J < m.nb; J++)
          for (int I = J + 1; I < m.nb; I++)
            iv.push_back({start_together ? (I - J) * 1024 : stretch ? (int)(((int64_t)(I - J) * 1024 * b.max_nb) / m.nb) : (I - J + shift) * 1024, -(I - J), sl, I, J});
      }
      std::stable_sort(iv.begin(), iv.end(), [](const TK& a, const TK& c) {
        if (a.s != c.s) return a.s < c.s;
        if (a.grp != c.grp) return a.grp < c.grp;
        if (a.slot != c.slot) return a.slot < c.slot;
        return a.J < c.J; });
      std::vector<int4> it(iv.size());
      for (size_t i = 0; i < iv.size(); i++) it[i] = make_int4(iv[i].slot, iv[i].I, iv[i].J, 0);
      b.n_trtri3 = (int)it.size();
      CUDA_TRY(h, upload(&b.d_trtri3_tasks, it));
      // back-substitution tasks (slot, J): level = distance from the bottom (a task depends on the tasks below it)
      std::vector<TK> sv;
      for (int s = b.s0; s < b.s1; s++) {
        const LeafMeta& m = h->meta[s];
        for (int J = 0; J < m.nb; J++) sv.push_back({m.nb - 1 - J, 0, s - b.s0, J, J});
      }
      std::stable_sort(sv.begin(), sv.end(), [](const TK& a, const TK& c) { return a.s != c.s ? a.s < c.s : a.slot < c.slot; });
      std::vector<int2> st2(sv.size());
      for (size_t i = 0; i < sv.size(); i++) st2[i] = make_int2(sv[i].slot, sv[i].J);
      b.n_solve = (int)st2.size();
      CUDA_TRY(h, upload(&b.d_solve_tasks, st2));
      CUDA_TRY(h, upload(&b.d_potrf2_tasks, pt));
      CUDA_TRY(h, upload(&b.d_flag_off, foff));
      maxFlags = std::max(maxFlags, fo);
    }
    b.ntiles = tile_off.back(); b.trpart_doubles = tro; b.gpart_doubles = go;
    b.n_trtri = (int)tt.size(); b.n_lauum = (int)lt.size();
    CUDA_TRY(h, upload(&b.d_tile_off, tile_off));
    CUDA_TRY(h, upload(&b.d_trpart_off, trp));
    CUDA_TRY(h, upload(&b.d_gpart_off, gp));
    CUDA_TRY(h, upload(&b.d_trtri_tasks, tt));
    CUDA_TRY(h, upload(&b.d_lauum_tasks, lt));
    maxF = std::max(maxF, b.f_doubles); maxW = std::max(maxW, b.w_doubles);
    maxTr = std::max(maxTr, tro); maxG = std::max(maxG, go);
  }
  CUDA_TRY(h, h->d_meta.alloc(ns));
  if (ns) CUDA_TRY(h, cudaMemcpy(h->d_meta.p, h->meta.data(), ns * sizeof(LeafMeta), cudaMemcpyHostToDevice));
  CUDA_TRY(h, h->d_xg.alloc(xoff));
  CUDA_TRY(h, h->d_y.alloc(voff));
  CUDA_TRY(h, h->d_z.alloc(voff));
  CUDA_TRY(h, h->d_alpha.alloc(voff));
  CUDA_TRY(h, h->d_F.alloc(maxF));
  CUDA_TRY(h, h->d_W.alloc(maxW));
  CUDA_TRY(h, h->d_WT.alloc(maxW));
  CUDA_TRY(h, h->d_prm.alloc((size_t)ns * h->pstride));
  CUDA_TRY(h, h->d_trpart.alloc(maxTr));
  CUDA_TRY(h, h->d_gpart.alloc(std::max<int64_t>(maxG, 1)));
  CUDA_TRY(h, h->d_rows.alloc((size_t)L * h->row_width));
  CUDA_TRY(h, h->d_scal.alloc(ns));
  CUDA_TRY(h, h->d_mask.alloc(std::max(ns, 1)));
  h->h_mask.assign(std::max(ns, 1), 1);
  CUDA_TRY(h, h->d_counter.alloc(16));
  CUDA_TRY(h, h->d_flags.alloc(std::max<int64_t>(maxFlags, 1)));
  CUDA_TRY(h, h->d_ldpart.alloc(std::max<int64_t>(maxTr / 2, 1)));
  CUDA_TRY(h, h->d_zzpart.alloc(std::max<int64_t>(maxTr / 2, 1)));
  CUDA_TRY(h, h->d_apart.alloc(std::max<int64_t>(maxFlags, 1) * BLK));
  CUDA_TRY(h, h->d_tpart.alloc(std::max<int64_t>(maxFlags, 1)));
  CUDA_TRY(h, h->d_leaf_mean.alloc(L));
  CUDA_TRY(h, cudaMemcpy(h->d_leaf_mean.p, h->leaf_mean.data(), L * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemset(h->d_rows.p, 0, (size_t)L * h->row_width * sizeof(double)));
  CUDA_TRY(h, cudaMallocHost(&h->pin_rows, (size_t)L * h->row_width * sizeof(double)));
  CUDA_TRY(h, cudaMallocHost(&h->pin_scal, std::max(ns, 1) * sizeof(LeafScal)));
  h->h_prm.assign((size_t)ns * h->pstride, 0.0);
  h->h_rows.assign((size_t)L * h->row_width, 0.0);
  h->h_info.assign(L, 0);

  // y (zero padded) and gathered inputs
  {
    std::vector<double> yv(voff, 0.0);
    std::vector<int64_t> obs, obs_off(ns + 1, 0);
    for (int s = 0; s < ns; s++) {
      const int l = loc[s];
      const int64_t b0 = h->leaf_ptr[l], n = h->meta[s].n;
      std::copy(y_centered + b0, y_centered + b0 + n, yv.begin() + h->meta[s].voff);
      obs.insert(obs.end(), leaf_obs + b0, leaf_obs + b0 + n);
      obs_off[s + 1] = obs_off[s] + n;
    }
    if (voff) CUDA_TRY(h, cudaMemcpy(h->d_y.p, yv.data(), voff * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemset(h->d_z.p, 0, voff * sizeof(double)));
    CUDA_TRY(h, cudaMemset(h->d_alpha.p, 0, voff * sizeof(double)));
    if (ns) {
      double* d_x = nullptr; int64_t* d_obs = nullptr; int64_t* d_obs_off = nullptr;
      CUDA_TRY(h, cudaMalloc(&d_x, (size_t)h->N * h->D * sizeof(double)));
      CUDA_TRY(h, cudaMemcpy(d_x, x, (size_t)h->N * h->D * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_TRY(h, upload(&d_obs, obs));
      CUDA_TRY(h, upload(&d_obs_off, obs_off));
      GatherArgs ga{h->d_meta.p, d_x, h->N, (int)h->D, d_obs, d_obs_off, h->d_xg.p};
      launch_gather(ga, h->meta[0].np, ns, h->stream);
      CUDA_TRY(h, cudaGetLastError());
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      cudaFree(d_x); cudaFree(d_obs); cudaFree(d_obs_off);
    }
  }
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_create(const double* x, int64_t N, int64_t D, int64_t L, const int64_t* leaf_ptr,
                                const int64_t* leaf_obs, const double* y_centered, const double* leaf_mean,
                                const int32_t* leaf_kernel_id, const dsmgp_kernel_desc* kernels, int32_t n_kernels,
                                const dsmgp_tree* tree, const dsmgp_opts* opts, dsmgp_handle** out) {
  if (out) *out = nullptr;
  auto fail = [&](int32_t code, const std::string& msg) { g_create_error = msg; return code; };
  if (!x || !leaf_ptr || !leaf_obs || !y_centered || !leaf_mean || !leaf_kernel_id || !kernels || !tree || !out)
    return fail(DSMGP_ERR_ARG, "null argument");
  if (N <= 0 || D <= 0 || L <= 0 || n_kernels <= 0) return fail(DSMGP_ERR_ARG, "N, D, L, n_kernels must be positive");
  if (D > 32) return fail(DSMGP_ERR_ARG, "D > 32 is not supported by the staged point tiles");
  dsmgp_handle* h = new dsmgp_handle();
  if (opts) h->opts = *opts; else dsmgp_default_opts(&h->opts);
  if (h->opts.world <= 0 || h->opts.rank < 0 || h->opts.rank >= h->opts.world) { delete h; return fail(DSMGP_ERR_ARG, "bad rank/world"); }
  h->N = N; h->D = D; h->L = L; h->nk = n_kernels;
  h->kernels.assign(kernels, kernels + n_kernels);
  h->koff.resize(n_kernels); h->knp.resize(n_kernels);
  for (int k = 0; k < n_kernels; k++) {
    const int t = kernels[k].type, np = kernels[k].nparams;
    const int nl = (t == DSMGP_ISO_SE || t == DSMGP_ISO_LINEAR) ? 1 : (int)D;
    if (t < 0 || t > 3 || np != nl + 2) { delete h; return fail(DSMGP_ERR_ARG, "kernel desc: bad type or nparams != len(logl)+2"); }
    h->koff[k] = h->H; h->knp[k] = np; h->H += np; h->Hmax = std::max(h->Hmax, np);
  }
  h->row_width = 1 + h->Hmax;
  h->leaf_ptr.assign(leaf_ptr, leaf_ptr + L + 1);
  h->leaf_kid.assign(leaf_kernel_id, leaf_kernel_id + L);
  h->leaf_mean.assign(leaf_mean, leaf_mean + L);
  if (leaf_ptr[0] != 0) { delete h; return fail(DSMGP_ERR_ARG, "leaf_ptr[0] != 0"); }
  for (int64_t l = 0; l < L; l++) {
    if (leaf_ptr[l + 1] <= leaf_ptr[l]) { delete h; return fail(DSMGP_ERR_ARG, "empty leaf"); }
    if (leaf_kernel_id[l] < 0 || leaf_kernel_id[l] >= n_kernels) { delete h; return fail(DSMGP_ERR_ARG, "leaf_kernel_id out of range"); }
    for (int64_t i = leaf_ptr[l]; i < leaf_ptr[l + 1]; i++)
      if (leaf_obs[i] < 1 || leaf_obs[i] > N) { delete h; return fail(DSMGP_ERR_ARG, "leaf_obs must be 1-based rows in 1..N"); }
  }
  std::string terr;
  if (!h->tree.load(tree, L, terr)) { delete h; return fail(DSMGP_ERR_ARG, terr); }
  h->node_lml.assign(h->tree.n_nodes, 0.0);
  h->owner.assign(L, 0);
  if (h->opts.world > 1) shard_lpt(L, leaf_ptr, h->opts.world, h->owner.data());
  // default parameters: zeros (IsoSE(0,0)-like); callers always set_params before fit
  h->theta_leaf.assign((size_t)L * h->Hmax, 0.0);

  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    delete h;
    return fail(DSMGP_ERR_CUDA, std::string("no CUDA device: libdsmgp has no CPU fallback (") + cudaGetErrorString(ce) + ")");
  }
  if (h->opts.device >= 0) h->device = h->opts.device; else cudaGetDevice(&h->device);
  if ((ce = cudaSetDevice(h->device)) != cudaSuccess) { delete h; return fail(DSMGP_ERR_CUDA, cudaGetErrorString(ce)); }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, h->device);
  if (prop.major != 10) {
    delete h;
    return fail(DSMGP_ERR_CUDA, "libdsmgp is built for sm_100a (B200) only; found compute capability " +
                                    std::to_string(prop.major) + "." + std::to_string(prop.minor));
  }
  if ((ce = engine_attrs()) != cudaSuccess) { delete h; return fail(DSMGP_ERR_CUDA, cudaGetErrorString(ce)); }
  cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  int32_t rc = plan_and_alloc(h, x, leaf_obs, y_centered);
  if (rc != DSMGP_OK) { g_create_error = h->err; delete h; return rc; }
  h->ev.assign(8 * std::max<size_t>(h->batches.size(), 1), nullptr);
  for (auto& e : h->ev) cudaEventCreate(&e);
  *out = h;
  return DSMGP_OK;
}

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
extern "C" int64_t dsmgp_nparams(const dsmgp_handle* h) { return h ? h->H : -1; }
extern "C" int64_t dsmgp_n_leaves(const dsmgp_handle* h) { return h ? h->L : -1; }
extern "C" int64_t dsmgp_n_nodes(const dsmgp_handle* h) { return h ? h->tree.n_nodes : -1; }
extern "C" int64_t dsmgp_row_width(const dsmgp_handle* h) { return h ? h->row_width : -1; }
extern "C" int64_t dsmgp_leaf_size(const dsmgp_handle* h, int64_t leaf) {
  if (!h || leaf < 0 || leaf >= h->L) return -1;
  return h->leaf_ptr[leaf + 1] - h->leaf_ptr[leaf];
}

extern "C" int32_t dsmgp_set_params(dsmgp_handle* h, const double* theta, int64_t n) {
  if (!h) return DSMGP_ERR_ARG;
  if (!theta || n != h->H) { h->err = "set_params: theta length must equal nparams"; return DSMGP_ERR_ARG; }
  for (int64_t l = 0; l < h->L; l++) {
    const int k = h->leaf_kid[l];
    std::copy(theta + h->koff[k], theta + h->koff[k] + h->knp[k], h->theta_leaf.begin() + (size_t)l * h->Hmax);
  }
  cudaSetDevice(h->device);
  upload_params(h);
  h->fitted = false; h->have_rows = false; h->have_grad = false;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_set_leaf_params(dsmgp_handle* h, int64_t leaf, const double* theta, int64_t n) {
  if (!h) return DSMGP_ERR_ARG;
  if (leaf < 0 || leaf >= h->L || !theta || n != h->knp[h->leaf_kid[leaf]]) { h->err = "set_leaf_params: bad leaf or length"; return DSMGP_ERR_ARG; }
  std::copy(theta, theta + n, h->theta_leaf.begin() + (size_t)leaf * h->Hmax);
  cudaSetDevice(h->device);
  upload_params(h);
  h->fitted = false; h->have_rows = false; h->have_grad = false;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_get_leaf_params(const dsmgp_handle* h, int64_t leaf, double* theta, int64_t n) {
  if (!h || leaf < 0 || leaf >= h->L || !theta || n != h->knp[h->leaf_kid[leaf]]) return DSMGP_ERR_ARG;
  std::copy(h->theta_leaf.begin() + (size_t)leaf * h->Hmax, h->theta_leaf.begin() + (size_t)leaf * h->Hmax + n, theta);
  return DSMGP_OK;
}

// ------------------------------------------------------------------------------------------
// the device pipeline
// ------------------------------------------------------------------------------------------
static bool needs_lauum(const dsmgp_handle* h) {
  for (int k = 0; k < h->nk; k++) {
    const int t = h->kernels[k].type;
    if (t == DSMGP_ISO_SE || t == DSMGP_ARD_LINEAR) return true;
    if (t == DSMGP_ARD_SE && !h->opts.as_written_grads) return true;
  }
  return false;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

// gram -> potrf -> solves (-> inverse -> lauum) -> rows, batch by batch
static int32_t finish_pipeline(dsmgp_handle* h, bool with_grad);
static int32_t run_pipeline(dsmgp_handle* h, bool with_grad, bool defer_sync = false) {
  cudaStream_t st = h->stream;
  const int sms = num_sms(h->device);
  const bool lau = with_grad && needs_lauum(h);
  h->tm = dsmgp_timings{};
  const int* mask_all = (with_grad && h->use_mask) ? h->d_mask.p : nullptr;
  if (h->batches.empty()) {   // a rank that owns no leaf
    h->fitted = true; h->have_rows = true; h->have_grad = with_grad; h->rows_complete = (h->opts.world == 1);
    return DSMGP_OK;
  }
  for (size_t bi = 0; bi < h->batches.size(); bi++) {
    Batch& b = h->batches[bi];
    cudaEvent_t* ev = h->ev.data() + 8 * bi;
    const int nsl = b.s1 - b.s0;
    CUDA_TRY(h, cudaEventRecord(ev[0], st));
    if (nsl == 0) { for (int k = 1; k < 8; k++) cudaEventRecord(ev[k], st); continue; }
    const LeafMeta* meta = h->d_meta.p + b.s0;
    LeafScal* scal = h->d_scal.p + b.s0;
    CUDA_TRY(h, cudaMemsetAsync(scal, 0, nsl * sizeof(LeafScal), st));
    cudaEventRecord(ev[1], st);
    GramArgs ga{meta, h->d_xg.p, h->d_prm.p, h->d_F.p, b.d_tile_off, nsl, (int)h->D};
    launch_gram_fit(ga, b.ntiles, st);
    h->tm.launches++;
    cudaEventRecord(ev[2], st);
    {
      CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
      CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p, 0, 16 * sizeof(int), st));
      Potrf2Args pa{meta, h->d_F.p, h->d_W.p, h->d_WT.p, h->d_y.p, h->d_z.p, scal, h->d_trpart.p, b.d_trpart_off,
                    h->d_ldpart.p, h->d_zzpart.p, h->d_flags.p, b.d_flag_off, b.d_potrf2_tasks, b.n_potrf2,
                    h->d_counter.p + 4, h->d_counter.p + 8, 0, nullptr};
      long long* d_trace = nullptr;
      const char* trace_file = getenv("DSMGP_TRACE_FILE");
      if (trace_file) { cudaMalloc(&d_trace, (size_t)b.n_potrf2 * 64); cudaMemsetAsync(d_trace, 0, (size_t)b.n_potrf2 * 64, st); pa.trace = d_trace; }
      launch_potrf2(pa, std::min(sms, b.n_potrf2), st);
      if (trace_file) {
        std::vector<long long> tr((size_t)b.n_potrf2 * 8);
        cudaMemcpyAsync(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        if (FILE* f = fopen(trace_file, "wb")) { fwrite(tr.data(), 8, tr.size(), f); fclose(f); }
        cudaFree(d_trace);
      }
      h->tm.launches++;
      cudaEventRecord(ev[3], st);
      if (!with_grad) {       // fit only: alpha by block back-substitution (the forward solve was fused above)
        CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
        SolveArgs sa{meta, h->d_F.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_flags.p, b.d_flag_off, b.d_solve_tasks, b.n_solve,
                     h->d_counter.p + 3, h->d_counter.p + 8};
        launch_solve(sa, std::max(1, std::min(solve_max_ctas(sms), b.n_solve)), st);
        h->tm.launches++;
      }
      cudaEventRecord(ev[4], st);
      if (with_grad) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
        Trtri3Args ta{meta, h->d_F.p, h->d_W.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_trpart.p, b.d_trpart_off,
                      h->d_flags.p, b.d_flag_off, h->d_apart.p, h->d_tpart.p, b.d_trtri3_tasks, b.n_trtri3,
                      h->d_counter.p, h->d_counter.p + 8, mask_all ? mask_all + b.s0 : nullptr};
        launch_trtri3(ta, std::max(1, std::min(sms, b.n_trtri3)), b.d_trtri_tasks, b.n_trtri, st);
        h->tm.launches += 2;
      }
    }
    cudaEventRecord(ev[5], st);
    if (lau) {
      LauumArgs la{meta, h->d_F.p, h->d_WT.p, h->d_xg.p, h->d_alpha.p, h->d_prm.p, b.d_lauum_tasks, b.n_lauum,
                   h->d_counter.p + 1, h->d_gpart.p, b.d_gpart_off, (int)h->D, h->d_counter.p + 8,
                   mask_all ? mask_all + b.s0 : nullptr};
      launch_lauum3(la, std::max(1, std::min(sms, b.n_lauum)), st);
      h->tm.launches++;
    }
    cudaEventRecord(ev[6], st);
    RowsArgs ra{meta, scal, scal, h->d_prm.p, h->d_trpart.p, b.d_trpart_off, h->d_gpart.p, b.d_gpart_off,
                h->d_rows.p, h->row_width, h->opts.as_written_grads, with_grad ? 1 : 0, lau ? 1 : 0,
                h->d_ldpart.p, h->d_zzpart.p, h->d_alpha.p, mask_all ? mask_all + b.s0 : nullptr};
    launch_rows(ra, nsl, st);
    h->tm.launches++;
    CUDA_TRY(h, cudaGetLastError());
    cudaEventRecord(ev[7], st);
    h->tm.potrf_flops += b.potrf_flops;
    h->tm.inverse_flops += with_grad ? b.potrf_flops * (lau ? 2.0 : 1.0) : 0.0;
    h->tm.gram_bytes += b.gram_bytes;
  }
  if (defer_sync) return DSMGP_OK;
  return finish_pipeline(h, with_grad);
}

static int32_t finish_pipeline(dsmgp_handle* h, bool with_grad) {
  cudaStream_t st = h->stream;
  const int ns = (int)h->slot_leaf.size();
  if (ns) CUDA_TRY(h, cudaMemcpyAsync(h->pin_scal, h->d_scal.p, ns * sizeof(LeafScal), cudaMemcpyDeviceToHost, st));
  int gerr = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&gerr, h->d_counter.p + 8, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  if (gerr != 0) { h->err = "device scheduler timeout (code " + std::to_string(gerr) + ")"; return DSMGP_ERR_STATE; }
  for (size_t bi = 0; bi < h->batches.size(); bi++) {
    cudaEvent_t* ev = h->ev.data() + 8 * bi;
    h->tm.gram_ms += ev_ms(ev[1], ev[2]);
    h->tm.potrf_ms += ev_ms(ev[2], ev[3]);
    h->tm.solve_ms += ev_ms(ev[3], ev[4]);
    h->tm.inverse_ms += ev_ms(ev[4], ev[5]);
    h->tm.grad_ms += ev_ms(ev[5], ev[7]);
  }
  h->tm.total_ms = ev_ms(h->ev[0], h->ev[8 * (h->batches.size() - 1) + 7]);
  std::fill(h->h_info.begin(), h->h_info.end(), 0);
  for (int s = 0; s < ns; s++) {
    int info = h->pin_scal[s].info;
    if (info > h->meta[s].n) info = 0;     // padding rows are identity
    h->h_info[h->slot_leaf[s]] = info;
  }
  h->fitted = true; h->have_rows = true; h->have_grad = with_grad; h->rows_complete = (h->opts.world == 1);
  h->alpha_exact = !with_grad;
  return DSMGP_OK;
}

// alpha = L^-T z by block back-substitution (gaussianprocess.jl:105).  The gradient path leaves alpha = X^T z, which is
// what tr(W) needs but carries the rounding of the explicit inverse; consumers of alpha itself (predict, accessors)
// get the back-substituted vector, exactly like the reference.
static int32_t refine_alpha(dsmgp_handle* h) {
  if (h->alpha_exact || !h->fitted || h->batches.size() != 1) return DSMGP_OK;
  Batch& b = h->batches[0];
  const int nsl = b.s1 - b.s0;
  if (nsl > 0) {
    CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p + 3, 0, sizeof(int), h->stream));
    SolveArgs sa{h->d_meta.p, h->d_F.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_flags.p, b.d_flag_off, b.d_solve_tasks, b.n_solve,
                 h->d_counter.p + 3, h->d_counter.p + 8};
    launch_solve(sa, std::max(1, std::min(solve_max_ctas(num_sms(h->device)), b.n_solve)), h->stream);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  h->alpha_exact = true;
  return DSMGP_OK;
}

static int32_t fetch_rows(dsmgp_handle* h) {
  const size_t bytes = (size_t)h->L * h->row_width * sizeof(double);
  CUDA_TRY(h, cudaMemcpyAsync(h->pin_rows, h->d_rows.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  memcpy(h->h_rows.data(), h->pin_rows, bytes);
  return DSMGP_OK;
}

static int32_t check_pd(dsmgp_handle* h) {
  if (!h->opts.strict_pd) return DSMGP_OK;
  for (int64_t l = 0; l < h->L; l++)
    if (h->h_info[l] != 0) {
      h->err = "PosDefException: leaf " + std::to_string(l) + " not positive definite at pivot " + std::to_string(h->h_info[l]);
      return DSMGP_ERR_NOT_PD;
    }
  return DSMGP_OK;
}

static void tree_grad(dsmgp_handle* h, const double* leaf_scale, double* grad) {
  std::fill(grad, grad + h->H, 0.0);
  DownCtx c{&h->tree, h->h_rows.data(), h->row_width, h->node_lml.data(), h->node_lml[h->tree.root],
            leaf_scale, h->leaf_kid.data(), h->koff.data(), h->knp.data(), grad};
  // a model with a single kernel writes at offset 0; kernel mixtures slice inside down_pass
  down_pass(c, h->tree.root, 0.0, 0.0, 0);
}

// Per-slot gradient mask from a leaf weight vector (NULL: every expert contributes).
static void set_grad_mask(dsmgp_handle* h, const double* leaf_scale) {
  h->use_mask = false;
  if (!leaf_scale) return;
  const int ns = (int)h->slot_leaf.size();
  bool any_zero = false;
  for (int s = 0; s < ns; s++) { h->h_mask[s] = leaf_scale[h->slot_leaf[s]] != 0.0 ? 1 : 0; any_zero |= !h->h_mask[s]; }
  if (!any_zero || ns == 0) return;
  cudaMemcpyAsync(h->d_mask.p, h->h_mask.data(), ns * sizeof(int), cudaMemcpyHostToDevice, h->stream);
  h->use_mask = true;
}

extern "C" int32_t dsmgp_fit(dsmgp_handle* h, int32_t* info, double* seconds) {
  if (!h) return DSMGP_ERR_ARG;
  cudaSetDevice(h->device);
  int32_t rc = run_pipeline(h, false);
  if (rc) return rc;
  if ((rc = fetch_rows(h))) return rc;
  if (info) std::copy(h->h_info.begin(), h->h_info.end(), info);
  if (seconds) *seconds = h->tm.total_ms * 1e-3;
  return check_pd(h);
}

extern "C" int32_t dsmgp_lml(dsmgp_handle* h, double* node_lml) {
  if (!h) return DSMGP_ERR_ARG;
  if (!h->have_rows) { h->err = "lml: call fit or eval first"; return DSMGP_ERR_STATE; }
  if (!h->rows_complete) { h->err = "lml: rows of other ranks missing (all-reduce the rows, then eval_finish_dev)"; return DSMGP_ERR_STATE; }
  up_pass(h->tree, h->h_rows.data(), h->row_width, h->node_lml.data());
  if (node_lml) std::copy(h->node_lml.begin(), h->node_lml.end(), node_lml);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_grad(dsmgp_handle* h, const double* leaf_scale, double* grad) {
  if (!h || !grad) return DSMGP_ERR_ARG;
  cudaSetDevice(h->device);
  int32_t rc;
  if (!h->have_grad) {
    if ((rc = run_pipeline(h, true))) return rc;
    if ((rc = fetch_rows(h))) return rc;
  }
  if (!h->rows_complete) { h->err = "grad: rows of other ranks missing"; return DSMGP_ERR_STATE; }
  up_pass(h->tree, h->h_rows.data(), h->row_width, h->node_lml.data());
  tree_grad(h, leaf_scale, grad);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_eval(dsmgp_handle* h, const double* theta, int64_t n, const double* leaf_scale,
                              double* lml, double* grad, double* node_lml) {
  if (!h) return DSMGP_ERR_ARG;
  int32_t rc;
  if (theta && (rc = dsmgp_set_params(h, theta, n))) return rc;
  cudaSetDevice(h->device);
  set_grad_mask(h, grad != nullptr ? leaf_scale : nullptr);   // finetune: experts with zero overlap weight skip the gradient kernels
  rc = run_pipeline(h, grad != nullptr);
  h->use_mask = false;
  if (rc) return rc;
  if ((rc = fetch_rows(h))) return rc;
  if (!h->rows_complete) { h->err = "eval: world > 1 needs eval_local_dev + all-reduce + eval_finish_dev"; return DSMGP_ERR_STATE; }
  up_pass(h->tree, h->h_rows.data(), h->row_width, h->node_lml.data());
  if (lml) *lml = h->node_lml[h->tree.root];
  if (node_lml) std::copy(h->node_lml.begin(), h->node_lml.end(), node_lml);
  if (grad) tree_grad(h, leaf_scale, grad);
  return check_pd(h);
}

// finetune!'s inner loop (finetuning.jl:36-58) for G anchor experts in ONE call: the G evaluations of an iteration are
// independent (theta_g is only updated from its own gradient), so they are enqueued back to back on the stream with
// no host synchronisation in between; experts with D[g, l] == 0 skip the inverse / LAUUM kernels.
extern "C" int32_t dsmgp_finetune_eval(dsmgp_handle* h, int64_t G, const int64_t* anchors, const double* thetas,
                                       const double* overlap, double* leaf_lml, double* grads, double* root_lml) {
  if (!h) return DSMGP_ERR_ARG;
  if (G <= 0 || !anchors || !thetas || !overlap || !leaf_lml || !grads) { h->err = "finetune_eval: bad argument"; return DSMGP_ERR_ARG; }
  if (h->opts.world != 1) { h->err = "finetune_eval: single-process handles only"; return DSMGP_ERR_STATE; }
  const int64_t L = h->L, H = h->H, rw = h->row_width;
  for (int64_t g = 0; g < G; g++) if (anchors[g] < 0 || anchors[g] >= L) { h->err = "finetune_eval: anchor out of range"; return DSMGP_ERR_ARG; }
  cudaSetDevice(h->device);
  const size_t need = (size_t)G * L * rw;
  if (need > h->pin_multi_doubles) {
    if (h->pin_multi) cudaFreeHost(h->pin_multi);
    h->pin_multi = nullptr; h->pin_multi_doubles = 0;
    CUDA_TRY(h, cudaMallocHost(&h->pin_multi, need * sizeof(double)));
    h->pin_multi_doubles = need;
  }
  std::vector<double> scale(L);
  std::vector<int32_t> info((size_t)G * L, 0);
  const int ns = (int)h->slot_leaf.size();
  const size_t need_scal = (size_t)G * std::max(ns, 1);
  if (need_scal > h->pin_scal_multi_n) {
    if (h->pin_scal_multi) cudaFreeHost(h->pin_scal_multi);
    h->pin_scal_multi = nullptr; h->pin_scal_multi_n = 0;
    CUDA_TRY(h, cudaMallocHost(&h->pin_scal_multi, need_scal * sizeof(LeafScal)));
    h->pin_scal_multi_n = need_scal;
  }
  LeafScal* pin_scal_multi = h->pin_scal_multi;
  int32_t rc = DSMGP_OK;
  for (int64_t g = 0; g < G && rc == DSMGP_OK; g++) {
    if ((rc = dsmgp_set_params(h, thetas + g * H, H))) break;              // setparams!(spn, hyp_)  finetuning.jl:41
    for (int64_t l = 0; l < L; l++) scale[l] = overlap[anchors[g] + l * L];   // view(D, g, :)  finetuning.jl:53
    set_grad_mask(h, scale.data());
    rc = run_pipeline(h, true, true);
    h->use_mask = false;
    if (rc) break;
    cudaMemcpyAsync(h->pin_multi + (size_t)g * L * rw, h->d_rows.p, (size_t)L * rw * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (ns) cudaMemcpyAsync(pin_scal_multi + (size_t)g * ns, h->d_scal.p, ns * sizeof(LeafScal), cudaMemcpyDeviceToHost, h->stream);
  }
  if (rc == DSMGP_OK) rc = finish_pipeline(h, true);        // one synchronisation for all G evaluations
  if (rc) return rc;
  std::vector<int64_t> node_of_leaf(L, -1);
  for (int64_t i = 0; i < h->tree.n_nodes; i++) if (h->tree.type[i] == DSMGP_NODE_LEAF) node_of_leaf[h->tree.leaf_of_node[i]] = i;
  for (int64_t g = 0; g < G; g++) {
    const double* rows = h->pin_multi + (size_t)g * L * rw;
    std::copy(rows, rows + (size_t)L * rw, h->h_rows.begin());
    for (int s = 0; s < ns; s++) {           // a non-PD expert poisons its own LML like the reference's log of a bad pivot
      const int inf = pin_scal_multi[(size_t)g * ns + s].info;
      if (inf != 0 && inf <= h->meta[s].n && h->opts.strict_pd) {
        h->err = "PosDefException: leaf " + std::to_string(h->slot_leaf[s]) + " (finetune anchor " + std::to_string(anchors[g]) + ")";
        return DSMGP_ERR_NOT_PD;
      }
    }
    up_pass(h->tree, h->h_rows.data(), rw, h->node_lml.data());             // mll!(spn, L)  finetuning.jl:47-48
    leaf_lml[g] = h->node_lml[node_of_leaf[anchors[g]]];                     // L[gp.id]      finetuning.jl:51
    if (root_lml) root_lml[g] = h->node_lml[h->tree.root];
    for (int64_t l = 0; l < L; l++) scale[l] = overlap[anchors[g] + l * L];
    tree_grad(h, scale.data(), grads + g * H);                               // finetuning.jl:53
  }
  h->rows_complete = true;
  return DSMGP_OK;
}

// train!(spn, D, gpmap, optim; iterations, lambda, earlystop) optimisers.jl:40-83 as ONE call: the loop stays inside the
// library (no per-iteration host round trip through the binding).  Optimisers = Flux.Optimise Descent / ADAM / RMSProp;
// `state_by_identity` reproduces the reference's `hyp += grad` rebinding, which gives apply! a fresh state every iteration
// (SURVEY App. B Q9).  The update is gradient ASCENT.  Returns the number of iterations executed in *n_done.
extern "C" int32_t dsmgp_train(dsmgp_handle* h, int32_t optimiser, double eta, double beta1, double beta2,
                               int32_t state_by_identity, int64_t iterations, double lambda, int64_t earlystop,
                               double* theta, double* ell, int64_t* n_done) {
  if (!h) return DSMGP_ERR_ARG;
  if (!theta || !ell || iterations <= 0 || optimiser < 0 || optimiser > 2) { h->err = "train: bad argument"; return DSMGP_ERR_ARG; }
  if (h->opts.world != 1) { h->err = "train: single-process handles only"; return DSMGP_ERR_STATE; }
  const int64_t H = h->H;
  std::vector<double> hyp(theta, theta + H), grad(H), mt(H, 0.0), vt(H, 0.0), acc(H, 0.0);
  double bp1 = beta1, bp2 = beta2;
  int64_t c = 0, it = 0;
  if (n_done) *n_done = 0;
  for (it = 0; it < iterations; it++) {
    double lml = 0.0;
    int32_t rc = dsmgp_eval(h, hyp.data(), H, nullptr, &lml, grad.data(), nullptr);     // :43-49, 68-77
    if (rc) return rc;
    ell[it] = lml;
    double delta = std::numeric_limits<double>::infinity();
    if (it >= 10) { double mean = 0.0; for (int64_t k = it - 9; k < it; k++) mean += ell[k]; delta = std::fabs(ell[it] - mean / 9.0); }   // :53
    c = (delta < lambda) ? c + 1 : 0;                                                    // :57-61
    if (c >= earlystop) { it++; break; }                                                 // :63-66 (returns before the update)
    if (state_by_identity) { std::fill(mt.begin(), mt.end(), 0.0); std::fill(vt.begin(), vt.end(), 0.0); std::fill(acc.begin(), acc.end(), 0.0); bp1 = beta1; bp2 = beta2; }
    for (int64_t k = 0; k < H; k++) {                                                    // Flux.Optimise.apply!  :78
      double d = grad[k];
      if (optimiser == 0) d *= eta;
      else if (optimiser == 1) {
        mt[k] = beta1 * mt[k] + (1.0 - beta1) * d;
        vt[k] = beta2 * vt[k] + (1.0 - beta2) * d * d;
        d = mt[k] / (1.0 - bp1) / (std::sqrt(vt[k] / (1.0 - bp2)) + 1e-8) * eta;
      } else {
        acc[k] = beta1 * acc[k] + (1.0 - beta1) * d * d;                                 // RMSProp: beta1 = rho
        d = d * (eta / (std::sqrt(acc[k]) + 1e-8));
      }
      hyp[k] = hyp[k] + d;                                                               // :79
    }
    if (optimiser == 1) { bp1 *= beta1; bp2 *= beta2; }
  }
  std::copy(hyp.begin(), hyp.end(), theta);
  if (n_done) *n_done = it;
  if (it >= iterations) {                                                                // :82-83 final setparams! + fit!
    int32_t rc = dsmgp_set_params(h, hyp.data(), H);
    if (rc) return rc;
    return dsmgp_fit(h, nullptr, nullptr);
  }
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_eval_local_dev(dsmgp_handle* h, const double* theta, int64_t n, double** rows_dev) {
  if (!h || !rows_dev) return DSMGP_ERR_ARG;
  int32_t rc;
  if (theta && (rc = dsmgp_set_params(h, theta, n))) return rc;
  cudaSetDevice(h->device);
  if ((rc = run_pipeline(h, true))) return rc;
  *rows_dev = h->d_rows.p;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_eval_finish_dev(dsmgp_handle* h, const double* leaf_scale, double* lml, double* grad, double* node_lml) {
  if (!h) return DSMGP_ERR_ARG;
  cudaSetDevice(h->device);
  int32_t rc;
  CUDA_TRY(h, cudaDeviceSynchronize());   // the caller's collective ran on its own stream
  if ((rc = fetch_rows(h))) return rc;
  h->rows_complete = true;
  up_pass(h->tree, h->h_rows.data(), h->row_width, h->node_lml.data());
  if (lml) *lml = h->node_lml[h->tree.root];
  if (node_lml) std::copy(h->node_lml.begin(), h->node_lml.end(), node_lml);
  if (grad) tree_grad(h, leaf_scale, grad);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_leaf_rows(const dsmgp_handle* h, double* rows) {
  if (!h || !rows) return DSMGP_ERR_ARG;
  if (!h->have_rows) return DSMGP_ERR_STATE;
  std::copy(h->h_rows.begin(), h->h_rows.end(), rows);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_leaf_owner(const dsmgp_handle* h, int32_t* owner) {
  if (!h || !owner) return DSMGP_ERR_ARG;
  std::copy(h->owner.begin(), h->owner.end(), owner);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_leaf_info(const dsmgp_handle* h, int32_t* info) {
  if (!h || !info) return DSMGP_ERR_ARG;
  std::copy(h->h_info.begin(), h->h_info.end(), info);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_update_weights(dsmgp_handle* h, double* sum_logweights, double* z) {
  if (!h) return DSMGP_ERR_ARG;
  if (!h->have_rows || !h->rows_complete) { h->err = "update_weights: call fit or eval first"; return DSMGP_ERR_STATE; }
  h->sum_logw.assign(h->tree.child_ptr[h->tree.n_nodes], 0.0);
  update_weights(h->tree, h->h_rows.data(), h->row_width, h->sum_logw.data(), z);
  h->have_weights = true;
  if (sum_logweights) std::copy(h->sum_logw.begin(), h->sum_logw.end(), sum_logweights);
  return DSMGP_OK;
}

// ------------------------------------------------------------------------------------------
// accessors
// ------------------------------------------------------------------------------------------
static int32_t need_resident(const dsmgp_handle* h, int64_t leaf, int* slot) {
  if (!h || leaf < 0 || leaf >= h->L) return DSMGP_ERR_ARG;
  if (!h->fitted) return DSMGP_ERR_STATE;
  if (!h->opts.keep_factors || h->batches.size() != 1) return DSMGP_ERR_STATE;
  *slot = h->leaf_slot[leaf];
  if (*slot < 0) return DSMGP_ERR_STATE;    // owned by another rank
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_leaf_alpha(const dsmgp_handle* h, int64_t leaf, double* alpha) {
  int slot; int32_t rc = need_resident(h, leaf, &slot);
  if (rc) return rc;
  cudaSetDevice(h->device);
  if (refine_alpha(const_cast<dsmgp_handle*>(h))) return DSMGP_ERR_CUDA;
  const LeafMeta& m = h->meta[slot];
  if (cudaMemcpy(alpha, h->d_alpha.p + m.voff, m.n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return DSMGP_ERR_CUDA;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_leaf_factor(const dsmgp_handle* h, int64_t leaf, double* Lfac) {
  int slot; int32_t rc = need_resident(h, leaf, &slot);
  if (rc) return rc;
  cudaSetDevice(h->device);
  const LeafMeta& m = h->meta[slot];
  double* tmp = nullptr;
  if (cudaMalloc(&tmp, (size_t)m.n * m.n * 8) != cudaSuccess) return DSMGP_ERR_OOM;
  launch_untile(h->d_F.p + m.foff, m.nkc, m.n, tmp, h->stream);
  cudaError_t ce = cudaMemcpyAsync(Lfac, tmp, (size_t)m.n * m.n * 8, cudaMemcpyDeviceToHost, h->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  if (ce != cudaSuccess) return DSMGP_ERR_CUDA;
  return DSMGP_OK;
}

// ------------------------------------------------------------------------------------------
// prediction
// ------------------------------------------------------------------------------------------
namespace {
struct Router {
  const HostTree& t; const double* x; int64_t T;
  std::vector<std::vector<int64_t>> pts;   // per leaf
  bool bad = false;
  Router(const HostTree& tt, const double* xx, int64_t TT, int64_t L) : t(tt), x(xx), T(TT), pts(L) {}
  void route(int64_t node, const std::vector<int64_t>& idx, bool poe) {
    const int ty = t.type[node];
    if (ty == DSMGP_NODE_LEAF) { auto& v = pts[t.leaf_of_node[node]]; v.insert(v.end(), idx.begin(), idx.end()); return; }
    if (ty == DSMGP_NODE_SPLIT && !poe) {
      std::vector<std::vector<int64_t>> sub(t.nchild(node));
      for (int64_t p : idx) { const int64_t k = getchild(t, node, x, T, p); if (k < 0) { bad = true; return; } sub[k].push_back(p); }
      for (int64_t k = 0; k < t.nchild(node); k++) if (!sub[k].empty()) route(t.child(node, k), sub[k], poe);
      return;
    }
    for (int64_t k = 0; k < t.nchild(node); k++) route(t.child(node, k), idx, poe);
  }
};

// common.jl mixing on the host.  Leaf predictions are stored per leaf in routing order; `cursor` replays it.
struct Mixer {
  const HostTree& t; const double* x; int64_t T;
  const std::vector<std::vector<double>>& mu; const std::vector<std::vector<double>>& var;
  const std::vector<double>& logw;
  std::vector<size_t> cursor;
  Mixer(const HostTree& tt, const double* xx, int64_t TT, const std::vector<std::vector<double>>& m,
        const std::vector<std::vector<double>>& v, const std::vector<double>& lw)
      : t(tt), x(xx), T(TT), mu(m), var(v), logw(lw), cursor(m.size(), 0) {}
  void reset() { std::fill(cursor.begin(), cursor.end(), 0); }

  // _minpredict common.jl:151-173
  void minpredict(int64_t node, const std::vector<int64_t>& idx, std::vector<double>& out) {
    const int ty = t.type[node];
    out.assign(idx.size(), 0.0);
    if (ty == DSMGP_NODE_LEAF) {
      const int64_t l = t.leaf_of_node[node];
      for (size_t i = 0; i < idx.size(); i++) out[i] = mu[l][cursor[l] + i];
      cursor[l] += idx.size();
    } else if (ty == DSMGP_NODE_SPLIT) {
      std::vector<std::vector<int64_t>> sub(t.nchild(node)); std::vector<std::vector<size_t>> pos(t.nchild(node));
      for (size_t i = 0; i < idx.size(); i++) { const int64_t k = getchild(t, node, x, T, idx[i]); sub[k].push_back(idx[i]); pos[k].push_back(i); }
      std::vector<double> o;
      for (int64_t k = 0; k < t.nchild(node); k++) {
        if (sub[k].empty()) continue;
        minpredict(t.child(node, k), sub[k], o);
        for (size_t i = 0; i < o.size(); i++) out[pos[k][i]] = o[i];
      }
    } else {
      std::fill(out.begin(), out.end(), std::numeric_limits<double>::infinity());
      std::vector<double> o;
      for (int64_t k = 0; k < t.nchild(node); k++) {
        minpredict(t.child(node, k), idx, o);
        for (size_t i = 0; i < o.size(); i++) out[i] = std::min(out[i], o[i]);
      }
    }
  }
  // _predict common.jl:134-143,181-196,275-292 : log(mu - mumin), log(mu^2), log(sigma^2)
  void predict(int64_t node, const std::vector<int64_t>& idx, const std::vector<double>& mumin,
               std::vector<double>& lm, std::vector<double>& lm2, std::vector<double>& ls) {
    const int ty = t.type[node];
    const size_t n = idx.size();
    lm.assign(n, 0.0); lm2.assign(n, 0.0); ls.assign(n, 0.0);
    if (ty == DSMGP_NODE_LEAF) {
      const int64_t l = t.leaf_of_node[node];
      for (size_t i = 0; i < n; i++) {
        const double m = mu[l][cursor[l] + i];
        double s2 = var[l][cursor[l] + i];
        if (s2 <= 0) s2 = 1e-8;                                  // common.jl:137
        lm[i] = std::log(m - mumin[i]); lm2[i] = std::log(m * m); ls[i] = std::log(s2);
      }
      cursor[l] += n;
    } else if (ty == DSMGP_NODE_SPLIT) {
      std::vector<std::vector<int64_t>> sub(t.nchild(node)); std::vector<std::vector<size_t>> pos(t.nchild(node));
      std::vector<std::vector<double>> mm(t.nchild(node));
      for (size_t i = 0; i < n; i++) {
        const int64_t k = getchild(t, node, x, T, idx[i]);
        sub[k].push_back(idx[i]); pos[k].push_back(i); mm[k].push_back(mumin[i]);
      }
      std::vector<double> a, b, c;
      for (int64_t k = 0; k < t.nchild(node); k++) {
        if (sub[k].empty()) continue;
        predict(t.child(node, k), sub[k], mm[k], a, b, c);
        for (size_t i = 0; i < a.size(); i++) { lm[pos[k][i]] = a[i]; lm2[pos[k][i]] = b[i]; ls[pos[k][i]] = c[i]; }
      }
    } else {
      const int64_t K = t.nchild(node);
      std::vector<std::vector<double>> A(K), B(K), C(K);
      for (int64_t k = 0; k < K; k++) predict(t.child(node, k), idx, mumin, A[k], B[k], C[k]);
      const double* lw = logw.data() + t.child_ptr[node];
      auto lse = [&](std::vector<std::vector<double>>& M, size_t i) {   // common.jl:309-313
        double m = -std::numeric_limits<double>::infinity();
        for (int64_t k = 0; k < K; k++) m = std::max(m, M[k][i] + lw[k]);
        double s = 0.0;
        for (int64_t k = 0; k < K; k++) s += std::exp((M[k][i] + lw[k]) - m);
        return std::log(s) + m;
      };
      for (size_t i = 0; i < n; i++) { lm[i] = lse(A, i); lm2[i] = lse(B, i); ls[i] = lse(C, i); }
    }
  }
  // _predictPoE common.jl:145-149,198-208 : (mu, precision)
  bool poe(int64_t node, const std::vector<int64_t>& idx, std::vector<double>& m, std::vector<double>& tau) {
    const int ty = t.type[node];
    const size_t n = idx.size();
    if (ty == DSMGP_NODE_LEAF) {
      const int64_t l = t.leaf_of_node[node];
      m.resize(n); tau.resize(n);
      for (size_t i = 0; i < n; i++) { m[i] = mu[l][cursor[l] + i]; tau[i] = 1.0 / var[l][cursor[l] + i]; }
      cursor[l] += n;
      return true;
    }
    if (ty != DSMGP_NODE_SPLIT) return false;    // MethodError in the reference
    m.assign(n, 0.0); tau.assign(n, 0.0);
    std::vector<double> m_, t_;
    for (int64_t k = 0; k < t.nchild(node); k++) {
      if (!poe(t.child(node, k), idx, m_, t_)) return false;
      for (size_t i = 0; i < n; i++) { tau[i] += t_[i]; m[i] += t_[i] * m_[i]; }
    }
    for (size_t i = 0; i < n; i++) m[i] = m[i] / tau[i];
    return true;
  }
};
}  // namespace

// Device prediction of every leaf on its routed points.  pts[l] = test rows routed to leaf l.
static int32_t predict_leaves(dsmgp_handle* h, const double* xtest, int64_t T, const std::vector<std::vector<int64_t>>& pts,
                              std::vector<std::vector<double>>& mu, std::vector<std::vector<double>>& var,
                              bool local_only = false) {
  if (!h->fitted) { h->err = "predict: call fit first"; return DSMGP_ERR_STATE; }
  if (!h->opts.keep_factors || h->batches.size() != 1) { h->err = "predict needs keep_factors=1"; return DSMGP_ERR_STATE; }
  { int32_t rr = refine_alpha(h); if (rr) return rr; }
  const int64_t L = h->L, D = h->D;
  mu.assign(L, {}); var.assign(L, {});
  std::vector<PredLeaf> pls; std::vector<int2> tasks; std::vector<int64_t> pl_leaf;
  int64_t xto = 0, vto = 0, oo = 0;
  for (int64_t l = 0; l < L; l++) {
    if (pts[l].empty()) continue;
    const int slot = h->leaf_slot[l];
    if (slot < 0) {
      if (local_only) continue;               // predicted by its owner (dsmgp_predict_local / _finish)
      h->err = "predict: leaf owned by another rank (use dsmgp_predict_local + all-reduce + dsmgp_predict_finish)";
      return DSMGP_ERR_STATE;
    }
    PredLeaf p; p.slot = slot; p.T = (int32_t)pts[l].size(); p.Tp = (p.T + BLK - 1) / BLK * BLK; p.pad_ = 0;
    p.xtoff = xto; xto += (int64_t)p.Tp * D;
    p.vtoff = vto; vto += (int64_t)(p.Tp / BLK) * h->meta[slot].nkc * TILE_D;
    p.ooff = oo; oo += p.Tp;
    for (int q = 0; q < p.Tp / BLK; q++) tasks.push_back(make_int2((int)pls.size(), q));
    pls.push_back(p); pl_leaf.push_back(l);
  }
  if (pls.empty()) return DSMGP_OK;
  std::stable_sort(tasks.begin(), tasks.end(), [&](const int2& a, const int2& b) {
    return h->meta[pls[a.x].slot].np > h->meta[pls[b.x].slot].np; });
  std::vector<double> xt(xto, 0.0);
  for (size_t i = 0; i < pls.size(); i++) {
    const auto& pv = pts[pl_leaf[i]];
    for (int64_t d = 0; d < D; d++)
      for (size_t q = 0; q < pv.size(); q++) xt[pls[i].xtoff + d * pls[i].Tp + q] = xtest[d * T + pv[q]];
  }
  DevBuf<double>&d_xt = h->p_xt, &d_VT = h->p_VT, &d_mu = h->p_mu, &d_var = h->p_var;
  DevBuf<PredLeaf>& d_pl = h->p_pl; DevBuf<int2>& d_tasks = h->p_tasks;
  auto cleanup = [&]() { if (d_VT.n * sizeof(double) > (size_t(16) << 30)) d_VT.free(); };   // keep the scratch unless it is huge
#define PTRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { h->err = std::string(#expr) + ": " + cudaGetErrorString(e_); cleanup(); \
    return e_ == cudaErrorMemoryAllocation ? DSMGP_ERR_OOM : DSMGP_ERR_CUDA; } } while (0)
  PTRY(d_xt.ensure(xto)); PTRY(d_VT.ensure(vto)); PTRY(d_mu.ensure(oo)); PTRY(d_var.ensure(oo));
  PTRY(d_pl.ensure(pls.size())); PTRY(d_tasks.ensure(tasks.size()));
  PTRY(cudaMemcpyAsync(d_xt.p, xt.data(), xto * 8, cudaMemcpyHostToDevice, h->stream));
  PTRY(cudaMemcpyAsync(d_pl.p, pls.data(), pls.size() * sizeof(PredLeaf), cudaMemcpyHostToDevice, h->stream));
  PTRY(cudaMemcpyAsync(d_tasks.p, tasks.data(), tasks.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
  PTRY(cudaMemsetAsync(h->d_counter.p, 0, 16 * sizeof(int), h->stream));
  PredArgs pa{h->d_meta.p, d_pl.p, d_tasks.p, (int)tasks.size(), h->d_counter.p + 2, h->d_F.p, h->d_W.p, h->d_xg.p,
              h->d_alpha.p, h->d_prm.p, h->d_leaf_mean.p, d_xt.p, d_VT.p, d_mu.p, d_var.p, (int)D, h->d_counter.p + 8,
              0, nullptr, nullptr, nullptr, nullptr, 0};
  const int sms = num_sms(h->device);
  const char* force_wave = getenv("DSMGP_PREDICT_WAVE");       // tests: "0" / "1" force the task granularity
  const bool use_wave = force_wave ? (force_wave[0] == '1') : ((int)tasks.size() < 3 * sms);
  if (use_wave) {
    // WAVE mode: too few (leaf, Q) tasks to fill the GPU -> one task per (leaf, Q, row block), ordered by row block
    // (a block depends only on the blocks above it) with the experts shifted so that they end together
    std::vector<int4> wt, wc;
    int base = 0, max_nb = 0;
    for (auto& p : pls) max_nb = std::max(max_nb, (int)h->meta[p.slot].nb);
    struct WK { int key, np, pl, Q, I, base; };
    std::vector<WK> wk;
    for (size_t i = 0; i < pls.size(); i++) {
      const LeafMeta& m = h->meta[pls[i].slot];
      for (int q = 0; q < pls[i].Tp / BLK; q++) {
        wc.push_back(make_int4((int)i, q, base, m.nb));
        for (int I = 0; I < m.nb; I++) wk.push_back({I + max_nb - m.nb, m.np, (int)i, q, I, base});
        base += m.nb;
      }
    }
    std::stable_sort(wk.begin(), wk.end(), [](const WK& a, const WK& b) { return a.key != b.key ? a.key < b.key : a.np > b.np; });
    for (auto& k : wk) wt.push_back(make_int4(k.pl, k.Q, k.I, k.base));
    PTRY(h->p_wtasks.ensure(wt.size())); PTRY(h->p_wcols.ensure(wc.size())); PTRY(h->p_flags.ensure(base));
    PTRY(h->p_part.ensure((size_t)base * 2 * BLK));
    PTRY(cudaMemcpyAsync(h->p_wtasks.p, wt.data(), wt.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    PTRY(cudaMemcpyAsync(h->p_wcols.p, wc.data(), wc.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    PTRY(cudaMemsetAsync(h->p_flags.p, 0, (size_t)base * sizeof(int), h->stream));
    PTRY(cudaStreamSynchronize(h->stream));       // wt / wc are locals
    pa.wave = 1; pa.wtasks = h->p_wtasks.p; pa.ntasks = (int)wt.size(); pa.flags = h->p_flags.p; pa.part = h->p_part.p;
    pa.wcols = h->p_wcols.p; pa.nwcols = (int)wc.size();
  }
  cudaEventRecord(h->ev[0], h->stream);
  launch_predict3(pa, std::max(1, std::min(sms, pa.ntasks)), h->stream);
  if (pa.wave) launch_predict_reduce(pa, h->stream);
  cudaEventRecord(h->ev[1], h->stream);
  h->tm.launches++;
  PTRY(cudaGetLastError());
  std::vector<double> hmu(oo), hvar(oo);
  PTRY(cudaMemcpyAsync(hmu.data(), d_mu.p, oo * 8, cudaMemcpyDeviceToHost, h->stream));
  PTRY(cudaMemcpyAsync(hvar.data(), d_var.p, oo * 8, cudaMemcpyDeviceToHost, h->stream));
  int gerr = 0;
  PTRY(cudaMemcpyAsync(&gerr, h->d_counter.p + 8, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PTRY(cudaStreamSynchronize(h->stream));
  if (gerr != 0) { h->err = "predict: device scheduler timeout (code " + std::to_string(gerr) + ")"; cleanup(); return DSMGP_ERR_STATE; }
#undef PTRY
  h->tm.predict_ms = ev_ms(h->ev[0], h->ev[1]);
  h->tm.predict_flops = 0.0; h->tm.predict_bytes = 0.0;
  for (size_t i = 0; i < pls.size(); i++) {      // TRSM n^2 T_l + 2 n T_l flop; L read once per block of 128 points
    const double n = h->meta[pls[i].slot].n, Tl = pls[i].T;
    h->tm.predict_flops += n * n * Tl + 2.0 * n * Tl;
    h->tm.predict_bytes += 8.0 * (n * (n + 1) / 2.0) * (pls[i].Tp / BLK) + 8.0 * (n + Tl) * (double)D + 16.0 * Tl;
  }
  for (size_t i = 0; i < pls.size(); i++) {
    const int64_t l = pl_leaf[i];
    mu[l].assign(hmu.begin() + pls[i].ooff, hmu.begin() + pls[i].ooff + pls[i].T);
    var[l].assign(hvar.begin() + pls[i].ooff, hvar.begin() + pls[i].ooff + pls[i].T);
  }
  cleanup();
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_leaf_predict(dsmgp_handle* h, int64_t leaf, const double* xtest, int64_t T, double* mu, double* var) {
  if (!h || leaf < 0 || leaf >= h->L || !xtest || T <= 0 || !mu || !var) return DSMGP_ERR_ARG;
  cudaSetDevice(h->device);
  std::vector<std::vector<int64_t>> pts(h->L);
  pts[leaf].resize(T);
  std::iota(pts[leaf].begin(), pts[leaf].end(), 0);
  std::vector<std::vector<double>> m, v;
  int32_t rc = predict_leaves(h, xtest, T, pts, m, v);
  if (rc) return rc;
  std::copy(m[leaf].begin(), m[leaf].end(), mu);
  std::copy(v[leaf].begin(), v[leaf].end(), var);
  return DSMGP_OK;
}

// argument checks + routing shared by the predict entry points
static int32_t predict_route(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, std::vector<std::vector<int64_t>>& pts) {
  if (!xtest || T <= 0 || mode < 0 || mode > 3) { h->err = "predict: bad argument"; return DSMGP_ERR_ARG; }
  for (int64_t i = 0; i < T * h->D; i++) if (!std::isfinite(xtest[i])) { h->err = "predict: non-finite input"; return DSMGP_ERR_ARG; }
  cudaSetDevice(h->device);
  const HostTree& t = h->tree;
  const bool poe = mode != DSMGP_PREDICT_DSMGP;
  if (poe && t.type[t.root] != DSMGP_NODE_SPLIT) { h->err = "predict: PoE/gPoE/rBCM need a split root (buildPoE/buildBCM model)"; return DSMGP_ERR_ARG; }
  if (!poe && !h->have_weights) {
    // the reference predicts with whatever logweights the sum nodes hold (uniform -log K after build)
    h->sum_logw.assign(t.child_ptr[t.n_nodes], 0.0);
    for (int64_t i = 0; i < t.n_nodes; i++)
      if (t.type[i] >= DSMGP_NODE_SUM) for (int64_t c = t.child_ptr[i]; c < t.child_ptr[i + 1]; c++) h->sum_logw[c] = -std::log((double)t.nchild(i));
  }
  std::vector<int64_t> all(T);
  std::iota(all.begin(), all.end(), 0);
  Router r(t, xtest, T, h->L);
  r.route(t.root, all, poe);
  if (r.bad) { h->err = "predict: a test point lies outside every split interval"; return DSMGP_ERR_ARG; }
  pts.swap(r.pts);
  return DSMGP_OK;
}

static int32_t predict_mix(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode,
                           const std::vector<std::vector<double>>& lmu, const std::vector<std::vector<double>>& lvar,
                           double* mu, double* var);

extern "C" int32_t dsmgp_predict(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, double* mu, double* var) {
  if (!h) return DSMGP_ERR_ARG;
  if (!mu || !var) { h->err = "predict: bad argument"; return DSMGP_ERR_ARG; }
  std::vector<std::vector<int64_t>> pts;
  int32_t rc = predict_route(h, xtest, T, mode, pts);
  if (rc) return rc;
  std::vector<std::vector<double>> lmu, lvar;
  if ((rc = predict_leaves(h, xtest, T, pts, lmu, lvar))) return rc;
  return predict_mix(h, xtest, T, mode, lmu, lvar, mu, var);
}

// Leaf-sharded prediction (one process per GPU): every rank predicts its own experts on the points routed to them and
// writes them into a buffer in (leaf, routing order) layout -- entries of other ranks' experts stay 0, so a SUM all-reduce
// assembles the buffer -- then every rank mixes (common.jl:134-307) redundantly, like the tree passes of an evaluation.
extern "C" int32_t dsmgp_predict_local(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, double* buf, int64_t* total) {
  if (!h) return DSMGP_ERR_ARG;
  if (!total) { h->err = "predict_local: bad argument"; return DSMGP_ERR_ARG; }
  std::vector<std::vector<int64_t>> pts;
  int32_t rc = predict_route(h, xtest, T, mode, pts);
  if (rc) return rc;
  int64_t tot = 0;
  for (auto& v : pts) tot += (int64_t)v.size();
  *total = tot;
  if (!buf) return DSMGP_OK;                   // size query
  std::vector<std::vector<double>> lmu, lvar;
  if ((rc = predict_leaves(h, xtest, T, pts, lmu, lvar, true))) return rc;
  std::fill(buf, buf + 2 * tot, 0.0);
  int64_t off = 0;
  for (int64_t l = 0; l < h->L; l++) {
    if (!lmu.empty() && !lmu[l].empty()) {
      std::copy(lmu[l].begin(), lmu[l].end(), buf + off);
      std::copy(lvar[l].begin(), lvar[l].end(), buf + tot + off);
    }
    off += (int64_t)pts[l].size();
  }
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_predict_finish(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, const double* buf,
                                        double* mu, double* var) {
  if (!h) return DSMGP_ERR_ARG;
  if (!buf || !mu || !var) { h->err = "predict_finish: bad argument"; return DSMGP_ERR_ARG; }
  std::vector<std::vector<int64_t>> pts;
  int32_t rc = predict_route(h, xtest, T, mode, pts);
  if (rc) return rc;
  int64_t tot = 0;
  for (auto& v : pts) tot += (int64_t)v.size();
  std::vector<std::vector<double>> lmu(h->L), lvar(h->L);
  int64_t off = 0;
  for (int64_t l = 0; l < h->L; l++) {
    lmu[l].assign(buf + off, buf + off + pts[l].size());
    lvar[l].assign(buf + tot + off, buf + tot + off + pts[l].size());
    off += (int64_t)pts[l].size();
  }
  return predict_mix(h, xtest, T, mode, lmu, lvar, mu, var);
}

static int32_t predict_mix(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode,
                           const std::vector<std::vector<double>>& lmu, const std::vector<std::vector<double>>& lvar,
                           double* mu, double* var) {
  const HostTree& t = h->tree;
  std::vector<int64_t> all(T);
  std::iota(all.begin(), all.end(), 0);
  Mixer mx(t, xtest, T, lmu, lvar, h->sum_logw);
  if (mode == DSMGP_PREDICT_DSMGP) {
    // predict(node) common.jl:175-179 (leaf), :243-254 (split root), :294-302 (sum root)
    struct Rec {
      Mixer& mx; const HostTree& t; double* mu; double* var; const double* x; int64_t T;
      void run(int64_t node, const std::vector<int64_t>& idx) {
        if (t.type[node] == DSMGP_NODE_SPLIT) {
          std::vector<std::vector<int64_t>> sub(t.nchild(node));
          for (int64_t p : idx) sub[getchild(t, node, x, T, p)].push_back(p);
          for (int64_t k = 0; k < t.nchild(node); k++) if (!sub[k].empty()) run(t.child(node, k), sub[k]);
          return;
        }
        // leaf or sum: two traversals of the subtree -> replay cursors must restart for this subtree.
        std::vector<size_t> save = mx.cursor;
        std::vector<double> mumin, lm, lm2, ls;
        mx.minpredict(node, idx, mumin);
        mx.cursor = save;
        for (auto& v : mumin) v -= 1.0;
        mx.predict(node, idx, mumin, lm, lm2, ls);
        for (size_t i = 0; i < idx.size(); i++) {
          const double m = std::exp(lm[i]) + mumin[i];
          mu[idx[i]] = m;
          var[idx[i]] = (t.type[node] == DSMGP_NODE_LEAF) ? std::exp(ls[i]) : std::exp(ls[i]) + (std::exp(lm2[i]) - m * m);
        }
      }
    } rec{mx, t, mu, var, xtest, T};
    rec.run(t.root, all);
    return DSMGP_OK;
  }
  const int64_t K = t.nchild(t.root);
  std::vector<double> m_, t_;
  if (mode == DSMGP_PREDICT_POE) {
    if (!mx.poe(t.root, all, m_, t_)) { h->err = "predictPoE: sum node below a split (MethodError in the reference)"; return DSMGP_ERR_ARG; }
    for (int64_t i = 0; i < T; i++) { mu[i] = m_[i]; var[i] = 1.0 / t_[i]; }
  } else if (mode == DSMGP_PREDICT_GPOE) {       // common.jl:211-222
    const double beta = 1.0 / (double)K;
    std::vector<double> M(T, 0.0), Tt(T, 0.0);
    for (int64_t k = 0; k < K; k++) {
      if (!mx.poe(t.child(t.root, k), all, m_, t_)) { h->err = "predictgPoE: sum node below a split"; return DSMGP_ERR_ARG; }
      for (int64_t i = 0; i < T; i++) { Tt[i] += beta * t_[i]; M[i] += beta * t_[i] * m_[i]; }
    }
    for (int64_t i = 0; i < T; i++) { mu[i] = M[i] / Tt[i]; var[i] = 1.0 / Tt[i]; }
  } else {                                       // rBCM common.jl:224-241
    int64_t nd = t.root;
    while (t.type[nd] != DSMGP_NODE_LEAF) nd = t.child(nd, 0);
    const int64_t l0 = t.leaf_of_node[nd];
    const int k0 = h->leaf_kid[l0];
    std::vector<double> prm(h->pstride);
    derive_params(h, k0, &h->theta_leaf[(size_t)l0 * h->Hmax], prm.data());
    const int type = h->kernels[k0].type;
    std::vector<double> s(T), C(T), M(T, 0.0);
    for (int64_t i = 0; i < T; i++) {
      double ktt;
      if (type == DSMGP_ISO_SE) ktt = prm[PRM_V];
      else if (type == DSMGP_ARD_SE) ktt = prm[PRM_V] * (double)h->D;
      else {
        ktt = 0.0;
        for (int64_t d = 0; d < h->D; d++) { const double xv = xtest[d * T + i]; ktt += (type == DSMGP_ISO_LINEAR ? prm[PRM_COEF] : prm[PRM_COEF + d]) * xv * xv; }
      }
      s[i] = ktt + prm[PRM_ETA];
      C[i] = 1.0 / s[i];
    }
    for (int64_t k = 0; k < K; k++) {
      if (!mx.poe(t.child(t.root, k), all, m_, t_)) { h->err = "predictrBCM: sum node below a split"; return DSMGP_ERR_ARG; }
      for (int64_t i = 0; i < T; i++) {
        const double s_ = 1.0 / t_[i];
        const double beta = 0.5 * (std::log(s[i]) - std::log(s_));
        C[i] = C[i] + (beta * t_[i]) - (beta / s[i]);
        M[i] = M[i] + m_[i] * (beta * t_[i]);
      }
    }
    for (int64_t i = 0; i < T; i++) { mu[i] = M[i] / C[i]; var[i] = 1.0 / C[i]; }
  }
  return DSMGP_OK;
}

// ------------------------------------------------------------------------------------------
// host-only helpers
// ------------------------------------------------------------------------------------------
extern "C" int32_t dsmgp_host_shard(int64_t L, const int64_t* leaf_ptr, int32_t world, int32_t* owner) {
  if (L <= 0 || !leaf_ptr || world <= 0 || !owner) return DSMGP_ERR_ARG;
  shard_lpt(L, leaf_ptr, world, owner);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_host_tree_eval(const dsmgp_tree* tree, int64_t L, const int32_t* leaf_kernel_id,
                                        const dsmgp_kernel_desc* kernels, int32_t n_kernels, const double* rows,
                                        int64_t row_width, const double* leaf_scale, double* node_lml, double* grad,
                                        double* sum_logweights, double* z) {
  if (!tree || !leaf_kernel_id || !kernels || !rows || !node_lml || L <= 0 || n_kernels <= 0) return DSMGP_ERR_ARG;
  HostTree t; std::string err;
  if (!t.load(tree, L, err)) { g_create_error = err; return DSMGP_ERR_ARG; }
  std::vector<int64_t> koff(n_kernels); std::vector<int32_t> knp(n_kernels);
  int64_t H = 0;
  for (int k = 0; k < n_kernels; k++) { koff[k] = H; knp[k] = kernels[k].nparams; H += knp[k]; }
  up_pass(t, rows, row_width, node_lml);
  if (grad) {
    std::fill(grad, grad + H, 0.0);
    DownCtx c{&t, rows, row_width, node_lml, node_lml[t.root], leaf_scale, leaf_kernel_id, koff.data(), knp.data(), grad};
    down_pass(c, t.root, 0.0, 0.0, 0);
  }
  if (sum_logweights || z) {
    std::vector<double> lw(t.child_ptr[t.n_nodes], 0.0);
    double zz = 0;
    update_weights(t, rows, row_width, lw.data(), &zz);
    if (sum_logweights) std::copy(lw.begin(), lw.end(), sum_logweights);
    if (z) *z = zz;
  }
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_get_timings(const dsmgp_handle* h, dsmgp_timings* t) {
  if (!h || !t) return DSMGP_ERR_ARG;
  *t = h->tm;
  return DSMGP_OK;
}
extern "C" int32_t dsmgp_set_profiling(dsmgp_handle* h, int32_t on) {
  if (!h) return DSMGP_ERR_ARG;
  h->profiling = on != 0;
  return DSMGP_OK;
}

// ------------------------------------------------------------------------------------------
// stand-alone operators
// ------------------------------------------------------------------------------------------
static int32_t standalone_device_check(std::string& err) {
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) { err = std::string("no CUDA device: libdsmgp has no CPU fallback (") + cudaGetErrorString(ce) + ")"; return DSMGP_ERR_CUDA; }
  if ((ce = engine_attrs()) != cudaSuccess) { err = cudaGetErrorString(ce); return DSMGP_ERR_CUDA; }
  return DSMGP_OK;
}
#define SA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_create_error = std::string(#expr) + ": " + cudaGetErrorString(e_); rc = DSMGP_ERR_CUDA; goto done; } } while (0)

extern "C" int32_t dsmgp_kernelmatrix(int32_t kernel_type, const double* theta, int64_t D, const double* x1, int64_t n1,
                                      const double* x2, int64_t n2, double* K) {
  if (kernel_type < 0 || kernel_type > 3 || !theta || !x1 || !x2 || !K || D <= 0 || n1 <= 0 || n2 <= 0) return DSMGP_ERR_ARG;
  int32_t rc = standalone_device_check(g_create_error);
  if (rc) return rc;
  const bool iso = (kernel_type == DSMGP_ISO_SE || kernel_type == DSMGP_ISO_LINEAR);
  const bool se = (kernel_type == DSMGP_ISO_SE || kernel_type == DSMGP_ARD_SE);
  const int nl = iso ? 1 : (int)D;
  std::vector<double> prm(PRM_COEF + D, 0.0);
  prm[PRM_V] = se ? std::exp(2.0 * theta[nl]) : 1.0;
  prm[PRM_S] = se ? std::exp(theta[nl]) : 1.0;
  for (int d = 0; d < nl; d++) { const double l = std::exp(theta[d]); prm[PRM_COEF + d] = se ? -0.5 / (l * l) : 1.0 / (l * l); }
  double *d1 = nullptr, *d2 = nullptr, *dk = nullptr, *dp = nullptr;
  SA_TRY(cudaMalloc(&d1, n1 * D * 8)); SA_TRY(cudaMalloc(&d2, n2 * D * 8)); SA_TRY(cudaMalloc(&dk, n1 * n2 * 8));
  SA_TRY(cudaMalloc(&dp, prm.size() * 8));
  SA_TRY(cudaMemcpy(d1, x1, n1 * D * 8, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(d2, x2, n2 * D * 8, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(dp, prm.data(), prm.size() * 8, cudaMemcpyHostToDevice));
  {
    GramRectArgs ga{kernel_type, (int)D, dp, d1, n1, (int)n1, d2, n2, (int)n2, dk, n1};
    launch_gram_rect(ga, 0);
    SA_TRY(cudaGetLastError());
    SA_TRY(cudaMemcpy(K, dk, n1 * n2 * 8, cudaMemcpyDeviceToHost));
  }
done:
  cudaFree(d1); cudaFree(d2); cudaFree(dk); cudaFree(dp);
  return rc;
}

// potrf / chol_continue on one host matrix.  k = number of leading rows/cols that already hold a valid factor.
// Runs the same persistent tile scheduler as the batched path (potrf2_kernel) on a one-expert batch.
// getOverlap(spn, D, gpmap) fit.jl:12-39 on the device.
extern "C" int32_t dsmgp_overlap(int64_t N, int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs,
                                 const int32_t* leaf_kernel_id, const dsmgp_tree* tree, double* D) {
  if (N <= 0 || L <= 0 || !leaf_ptr || !leaf_obs || !leaf_kernel_id || !tree || !D) { g_create_error = "overlap: bad argument"; return DSMGP_ERR_ARG; }
  HostTree t; std::string err;
  if (!t.load(tree, L, err)) { g_create_error = err; return DSMGP_ERR_ARG; }
  const int64_t total = leaf_ptr[L];
  for (int64_t i = 0; i < total; i++) if (leaf_obs[i] < 1 || leaf_obs[i] > N) { g_create_error = "overlap: leaf_obs must be 1-based rows in 1..N"; return DSMGP_ERR_ARG; }
  { int32_t rc = standalone_device_check(g_create_error); if (rc) return rc; }
  // ancestor chains (root first) from the child lists
  std::vector<int64_t> parent(t.n_nodes, -1);
  for (int64_t i = 0; i < t.n_nodes; i++) for (int64_t k = 0; k < t.nchild(i); k++) parent[t.child(i, k)] = i;
  std::vector<std::vector<int>> chain(L);
  int AD = 1;
  for (int64_t i = 0; i < t.n_nodes; i++) {
    if (t.type[i] != DSMGP_NODE_LEAF) continue;
    std::vector<int> c;
    for (int64_t u = parent[i]; u >= 0; u = parent[u]) c.push_back((int)u);
    std::reverse(c.begin(), c.end());
    AD = std::max<int>(AD, (int)c.size());
    chain[t.leaf_of_node[i]] = c;
  }
  std::vector<int> anc((size_t)L * AD, -1);
  for (int64_t l = 0; l < L; l++) std::copy(chain[l].begin(), chain[l].end(), anc.begin() + (size_t)l * AD);
  std::vector<int> ntype(t.type.begin(), t.type.end()), kid(leaf_kernel_id, leaf_kernel_id + L);
  int64_t* d_obs = nullptr; int64_t* d_lp = nullptr; int64_t* d_poff = nullptr;
  int *d_cnt = nullptr, *d_plist = nullptr, *d_inter = nullptr, *d_kid = nullptr, *d_anc = nullptr, *d_nt = nullptr;
  double* d_D = nullptr;
  auto cleanup = [&]() { cudaFree(d_obs); cudaFree(d_lp); cudaFree(d_poff); cudaFree(d_cnt); cudaFree(d_plist); cudaFree(d_inter);
                         cudaFree(d_kid); cudaFree(d_anc); cudaFree(d_nt); cudaFree(d_D); };
#define OTRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_create_error = std::string(#expr) + ": " + cudaGetErrorString(e_); cleanup(); \
    return e_ == cudaErrorMemoryAllocation ? DSMGP_ERR_OOM : DSMGP_ERR_CUDA; } } while (0)
  OTRY(cudaMalloc(&d_obs, total * 8)); OTRY(cudaMemcpy(d_obs, leaf_obs, total * 8, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_lp, (L + 1) * 8)); OTRY(cudaMemcpy(d_lp, leaf_ptr, (L + 1) * 8, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_cnt, N * 4)); OTRY(cudaMemset(d_cnt, 0, N * 4));
  launch_ov_count(d_obs, total, d_cnt, nullptr);
  std::vector<int> cnt(N);
  OTRY(cudaMemcpy(cnt.data(), d_cnt, N * 4, cudaMemcpyDeviceToHost));
  std::vector<int64_t> poff(N + 1, 0);
  for (int64_t p = 0; p < N; p++) poff[p + 1] = poff[p] + cnt[p];
  OTRY(cudaMalloc(&d_poff, (N + 1) * 8)); OTRY(cudaMemcpy(d_poff, poff.data(), (N + 1) * 8, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_plist, std::max<int64_t>(total, 1) * 4));
  OTRY(cudaMemset(d_cnt, 0, N * 4));
  launch_ov_fill(d_obs, d_lp, (int)L, d_poff, d_cnt, d_plist, nullptr);
  OTRY(cudaMalloc(&d_inter, (size_t)L * L * 4)); OTRY(cudaMemset(d_inter, 0, (size_t)L * L * 4));
  launch_ov_pairs(d_poff, d_plist, N, L, d_inter, nullptr);
  OTRY(cudaMalloc(&d_kid, L * 4)); OTRY(cudaMemcpy(d_kid, kid.data(), L * 4, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_anc, anc.size() * 4)); OTRY(cudaMemcpy(d_anc, anc.data(), anc.size() * 4, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_nt, ntype.size() * 4)); OTRY(cudaMemcpy(d_nt, ntype.data(), ntype.size() * 4, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_D, (size_t)L * L * 8));
  launch_ov_finish(d_inter, d_lp, d_kid, d_anc, AD, d_nt, L, d_D, nullptr);
  OTRY(cudaGetLastError());
  OTRY(cudaMemcpy(D, d_D, (size_t)L * L * 8, cudaMemcpyDeviceToHost));
#undef OTRY
  cleanup();
  return DSMGP_OK;
}

static int32_t chol_host_matrix(double* A, int64_t n, int64_t k, int32_t* info) {
  int32_t rc = standalone_device_check(g_create_error);
  if (rc) return rc;
  const int64_t kp = (k + BLK - 1) / BLK * BLK;          // leading part padded to a block boundary
  const int64_t nn = kp + (n - k);
  LeafMeta m{};
  m.n = (int32_t)nn; m.np = (int32_t)((nn + PAD - 1) / PAD * PAD); m.nb = (m.np + BLK - 1) / BLK; m.nkc = m.np / KC;
  const int np = m.np, nkc = m.nkc;
  const int64_t fd = tiled_doubles(np);
  std::vector<double> P((size_t)fd, 0.0);
  auto map = [&](int64_t i) { return i < k ? i : kp + (i - k); };
  for (int i = 0; i < np; i++) P[tidx(i, i, nkc)] = 1.0;
  for (int64_t c = 0; c < n; c++)
    for (int64_t r = c; r < n; r++) P[tidx((int)map(r), (int)map(c), nkc)] = A[c * n + r];
  std::vector<int4> tasks;
  tasks.push_back(make_int4(0, 0, 0, 0));
  for (int J = 0; J + 1 < m.nb; J++) {
    tasks.push_back(make_int4(0, J + 1, J, 0));
    tasks.push_back(make_int4(0, J + 1, J + 1, 0));
    for (int I = J + 2; I < m.nb; I++) tasks.push_back(make_int4(0, I, J, 0));
  }
  const int64_t nflags = (int64_t)m.nb * (m.nb + 1) / 2;
  double *dF = nullptr, *dW = nullptr, *dWT = nullptr, *dtr = nullptr, *dv = nullptr; LeafMeta* dm = nullptr; LeafScal* ds = nullptr;
  int64_t* doff = nullptr; int* dflags = nullptr; int* dcnt = nullptr; int4* dtasks = nullptr;
  LeafScal sc{}; int gerr = 0;
  const int64_t zero[2] = {0, 0};
  int sms = 148;
  SA_TRY(cudaMalloc(&dF, fd * 8)); SA_TRY(cudaMalloc(&dW, (size_t)m.nb * WBLK_D * 8)); SA_TRY(cudaMalloc(&dWT, (size_t)m.nb * WBLK_D * 8));
  SA_TRY(cudaMalloc(&dtr, 4 * m.nb * 8)); SA_TRY(cudaMalloc(&dv, 2 * (size_t)np * 8)); SA_TRY(cudaMalloc(&dm, sizeof(LeafMeta)));
  SA_TRY(cudaMalloc(&ds, sizeof(LeafScal))); SA_TRY(cudaMalloc(&doff, 16)); SA_TRY(cudaMalloc(&dflags, nflags * 4));
  SA_TRY(cudaMalloc(&dcnt, 64)); SA_TRY(cudaMalloc(&dtasks, tasks.size() * sizeof(int4)));
  SA_TRY(cudaMemcpy(dF, P.data(), fd * 8, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice));
  SA_TRY(cudaMemset(ds, 0, sizeof(LeafScal))); SA_TRY(cudaMemset(dflags, 0, nflags * 4)); SA_TRY(cudaMemset(dcnt, 0, 64));
  SA_TRY(cudaMemset(dv, 0, 2 * (size_t)np * 8));
  SA_TRY(cudaMemcpy(doff, zero, 16, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(dtasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  {
    Potrf2Args pa{dm, dF, dW, dWT, dv, dv + np, ds, dtr, doff, dtr + 2 * m.nb, dtr + 3 * m.nb, dflags, doff + 1, dtasks,
                  (int)tasks.size(), dcnt, dcnt + 8, (int)(kp / BLK), nullptr};
    launch_potrf2(pa, std::min(sms, (int)tasks.size()), 0);
    SA_TRY(cudaGetLastError());
    SA_TRY(cudaMemcpy(P.data(), dF, fd * 8, cudaMemcpyDeviceToHost));
    SA_TRY(cudaMemcpy(&sc, ds, sizeof(sc), cudaMemcpyDeviceToHost));
    SA_TRY(cudaMemcpy(&gerr, dcnt + 8, sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (gerr != 0) { g_create_error = "device scheduler timeout"; rc = DSMGP_ERR_STATE; goto done; }
  for (int64_t c = 0; c < n; c++)
    for (int64_t r = 0; r < n; r++) A[c * n + r] = (r >= c) ? P[tidx((int)map(r), (int)map(c), nkc)] : 0.0;    // tril!
  if (info) {
    int64_t i = sc.info;                       // 1-based pivot in padded coordinates
    if (i > 0) { i = (i - 1 >= kp) ? (i - 1 - kp) + 1 : i; if (i > n - k) i = 0; }
    *info = (int32_t)i;                        // relative to the trailing block, as LAPACK.potrf!(C) reports it
  }
done:
  cudaFree(dF); cudaFree(dW); cudaFree(dWT); cudaFree(dtr); cudaFree(dv); cudaFree(dm); cudaFree(ds); cudaFree(doff);
  cudaFree(dflags); cudaFree(dcnt); cudaFree(dtasks);
  return rc;
}

extern "C" int32_t dsmgp_potrf(double* A, int64_t n, int32_t* info) {
  if (!A || n <= 0) return DSMGP_ERR_ARG;
  return chol_host_matrix(A, n, 0, info);
}

extern "C" int32_t dsmgp_chol_continue(double* A, int64_t n, int64_t ki, int32_t* info) {
  if (!A || n <= 0 || ki < 1 || ki > n) return DSMGP_ERR_ARG;
  return chol_host_matrix(A, n, ki - 1, info);
}

extern "C" int32_t dsmgp_chol_delete_rows(const double* A, int64_t n, const int64_t* rows, int64_t nrows, double* out) {
  if (!A || n <= 0 || !rows || nrows < 0 || nrows >= n || !out) return DSMGP_ERR_ARG;
  for (int64_t q = 0; q < nrows; q++)
    if (rows[q] < 1 || rows[q] > n || (q > 0 && rows[q] <= rows[q - 1])) { g_create_error = "delete_rows: rows must be 1-based ascending"; return DSMGP_ERR_ARG; }
  int32_t rc = standalone_device_check(g_create_error);
  if (rc) return rc;
  double *dL = nullptr, *dv = nullptr; int64_t* dr = nullptr;
  std::vector<double> P((size_t)n * n);
  SA_TRY(cudaMalloc(&dL, n * n * 8)); SA_TRY(cudaMalloc(&dv, n * 8)); SA_TRY(cudaMalloc(&dr, std::max<int64_t>(nrows, 1) * 8));
  SA_TRY(cudaMemcpy(dL, A, n * n * 8, cudaMemcpyHostToDevice));
  if (nrows) SA_TRY(cudaMemcpy(dr, rows, nrows * 8, cudaMemcpyHostToDevice));
  launch_delete_rows(dL, (int)n, dr, (int)nrows, dv, 0);
  SA_TRY(cudaGetLastError());
  SA_TRY(cudaMemcpy(P.data(), dL, n * n * 8, cudaMemcpyDeviceToHost));
  {
    std::vector<int64_t> keep;
    for (int64_t i = 0, q = 0; i < n; i++) { if (q < nrows && rows[q] - 1 == i) { q++; continue; } keep.push_back(i); }
    const int64_t m = (int64_t)keep.size();
    for (int64_t c = 0; c < m; c++)
      for (int64_t r = 0; r < m; r++) out[c * m + r] = (r >= c) ? P[keep[c] * n + keep[r]] : 0.0;
  }
done:
  cudaFree(dL); cudaFree(dv); cudaFree(dr);
  return rc;
}
