// Internal definitions shared by the translation units that implement the C ABI (api*.cu, comm.cu).
#pragma once
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
#include "ozaki_args.h"
#include "tree_host.h"

namespace dsm {
std::string& create_error();                 // thread-local message of the last failed call without a handle
}

#define CUDA_TRY(h, expr)                                                                       \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                            \
      return e_ == cudaErrorMemoryAllocation ? DSMGP_ERR_OOM : DSMGP_ERR_CUDA;                  \
    }                                                                                           \
  } while (0)

namespace dsm {

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
  // sharing plan active (dsmgp_set_sharing): A = experts factored on their own, B = SHARE_PREFIX experts (second launch,
  // after their leading block rows have been copied from the source); aliased experts appear in neither
  int4* d_potrf2_A = nullptr; int n_potrf2_A = 0;
  int4* d_potrf2_B = nullptr; int n_potrf2_B = 0;
  int* d_prefix_slots = nullptr; int n_prefix = 0, max_jb = 0;   // batch-relative slots of the SHARE_PREFIX experts
  int4* d_trtri3_tasks = nullptr; int n_trtri3 = 0;     // inverse: tile tasks by anti-diagonal
  int4* d_trtri3m_tasks = nullptr;                      // inverse tiles in the order of the fused launch (fused2.cuh): by the level at
                                                        // which their block row completes in the factorisation, then row-major
  int2* d_solve_tasks = nullptr; int n_solve = 0;       // back-substitution: (slot, J) by level from the bottom
  int64_t* d_flag_off = nullptr; int64_t flag_ints = 0;
  double potrf_flops = 0, gram_bytes = 0;
  std::vector<int4> h_lauum;                            // host copy of d_lauum_tasks
  std::vector<int4> h_potrf2;                           // host copy of d_potrf2_tasks
  std::vector<int4> h_trtri3m;                          // host copy of d_trtri3m_tasks
  std::vector<int4> h_trtri3;                           // host copy of d_trtri3_tasks (filtered by the INT8 split plan)
  OzPlan oz;                                            // split inverse on the INT8 tensor cores (api_ozaki.cu), DSMGP_OZAKI=1
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
  static constexpr size_t MAX_CACHED = size_t(48) << 30;     // per device: allocations outside DevBuf (CUB, torch, small uploads) need room too
  bool give(void* p, size_t bytes, int dev) {
    if (bytes < MIN_BYTES) return false;
    std::lock_guard<std::mutex> g(mu);
    if (ents.size() >= 64) return false;
    size_t held = 0;
    for (auto& e : ents) if (e.dev == dev) held += e.bytes;
    if (held + bytes > MAX_CACHED) return false;
    ents.push_back({p, bytes, dev});
    return true;
  }
  void release_all() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& e : ents) { int cur = 0; cudaGetDevice(&cur); cudaSetDevice(e.dev); cudaFree(e.p); cudaSetDevice(cur); }
    ents.clear();
  }
};
extern BufCache g_cache;

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

}  // namespace dsm

using namespace dsm;     // internal header: only the ABI translation units include it

namespace dsm {
// comm.cu: NCCL through dlopen (the library has no link-time dependency on it)
void comm_destroy(void* comm);
int32_t comm_allreduce_sum(dsmgp_handle* h, double* dev_buf, size_t count);   // on h->stream; DSMGP_OK / DSMGP_ERR_COMM
}

struct dsmgp_handle {
  int64_t N = 0, D = 0, L = 0;
  int nk = 0;
  std::vector<dsmgp_kernel_desc> kernels;
  std::vector<int64_t> koff;      // theta offset per kernel
  std::vector<int32_t> knp;       // nparams per kernel
  int64_t H = 0; int Hmax = 0; int row_width = 0; int pstride = 0;
  std::vector<int64_t> leaf_ptr;
  std::vector<int32_t> leaf_obs32;   // observation lists (1-based), kept for the sharing plan when they fit (<= 2^26 entries)
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
  DevBuf<int> d_flags2;              // second per-tile flag array: the inverse's flags when it shares a launch with the factorisation
  DevBuf<double> d_ldpart, d_zzpart;
  DevBuf<double> d_apart, d_tpart;   // per-tile partials of the tile-pipelined inverse
  DevBuf<double> p_xt, p_VT, p_mu, p_var, p_part; DevBuf<PredLeaf> p_pl; DevBuf<int2> p_tasks;   // predict scratch (grow-only)
  DevBuf<int4> p_wtasks, p_wcols; DevBuf<int> p_flags;
  // device-side routing / mixing of large prediction batches (route.cuh)
  DevBuf<int> t_int;                    // flattened tree, integer arrays back to back
  DevBuf<double> t_split_val, p_logw, p_xtest, p_outmu, p_outvar;
  DevBuf<int> p_cnt, p_pidx, p_reach;   // p_cnt: [cnt L | fill L | ooff L | err 1]
  bool tree_on_device = false;
  bool capturing = false;               // the pipeline is being recorded into a CUDA graph: no event records, no host copies
  DevBuf<int> t_lvl;                    // level lists and per-leaf tables of the device tree passes (tree.cuh)
  int n_up = 0, n_dn = 0; size_t lvl_off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool lvl_on_device = false;
  DevBuf<double> g_dbl;                 // fused train loop: theta, out, optimiser state, tree scratch, trace, history
  DevBuf<int> g_int;
  int reach_dsmgp = 0, reach_poe = 0, tree_depth = 0, tree_maxk = 0, tree_frames = 0;
  DevBuf<int> d_mask; std::vector<int> h_mask; bool use_mask = false;   // per-slot gradient mask (finetune: zero-overlap experts)
  double* pin_multi = nullptr; size_t pin_multi_doubles = 0;             // rows of a multi-theta call [G][L][row_width]
  LeafScal* pin_scal_multi = nullptr; size_t pin_scal_multi_n = 0;       // per-slot scalars of a multi-theta call [G][slots]
  double* pin_rows = nullptr;
  LeafScal* pin_scal = nullptr;
  // sharing plan of fit! (fit.jl:71-122), built by dsmgp_set_sharing / dsmgp_fit(tau, overlap)
  struct Share {
    bool active = false; double tau = 0.05;
    std::vector<int4> slot;               // per local slot: (kind, source slot, copied block rows, 0)
    std::vector<int> leaf_kind, leaf_src; // per global leaf (diagnostics / accessors)
    int64_t n_alias = 0, n_prefix = 0, blocks_copied = 0;
    double flops_saved = 0.0;
  } share;
  DevBuf<int4> d_share;
  bool theta_global = true;               // every expert of a kernel holds the same theta (set_params); sharing needs it
  bool share_applied = false;             // the last pipeline ran with the sharing plan
  std::vector<int> exec_slot;             // global leaf -> slot that holds its results (the source's slot for an alias)
  void* comm = nullptr;                   // ncclComm_t (dsmgp_comm_init), or null
  DevBuf<double> d_comm_buf;              // device staging for the prediction all-reduce
  // split inverse on the INT8 tensor cores: slice pool + its tensor map, T^T scratch, row scales
  DevBuf<int8_t> oz_pool; DevBuf<double> oz_scratch, oz_scale; DevBuf<unsigned long long> oz_rowmax;
  alignas(64) unsigned char oz_map[256]; int oz_S = 8;     // tensor maps of the two rounds' boxes
  // measured segments of the last evaluation (CUDA events, read by dsmgp_int8_info): 0 block products, 1 slicing, 2 FP64 tile launches
  struct OzSeg { int kind; cudaEvent_t a, b; };
  std::vector<OzSeg> oz_segs; std::vector<cudaEvent_t> oz_evs; size_t oz_ev_used = 0;
  double oz_ksteps = 0.0;                 // k-steps of all block products of the last evaluation
  int64_t oz_pool_bytes = 0;
  bool oz_x_complete = false;             // the split inverse of the current batch ran (X^T complete, no masked experts)
  bool oz_inv_tiles_done = false;         // the tile-pipeline part of the inverse ran inside the factorisation launches
  bool oz_l21_ready = false;              // the L21 slices of the current batch were made by the factorisation phase
  std::string err;

  ~dsmgp_handle() {
    for (auto& b : batches) {
      cudaFree(b.d_tile_off); cudaFree(b.d_trpart_off); cudaFree(b.d_gpart_off);
      cudaFree(b.d_trtri_tasks); cudaFree(b.d_lauum_tasks); cudaFree(b.d_potrf2_tasks); cudaFree(b.d_trtri3_tasks); cudaFree(b.d_solve_tasks); cudaFree(b.d_flag_off);
      cudaFree(b.d_potrf2_A); cudaFree(b.d_potrf2_B); cudaFree(b.d_prefix_slots); cudaFree(b.d_trtri3m_tasks);
      cudaFree(b.oz.d_blob);
    }
    d_flags2.free();
    oz_pool.free(); oz_scratch.free(); oz_scale.free(); oz_rowmax.free();
    d_share.free(); d_comm_buf.free();
    dsm::comm_destroy(comm);
    d_meta.free(); d_xg.free(); d_y.free(); d_z.free(); d_alpha.free(); d_F.free(); d_W.free(); d_WT.free();
    d_flags.free(); d_ldpart.free(); d_zzpart.free(); d_apart.free(); d_tpart.free();
    p_xt.free(); p_VT.free(); p_mu.free(); p_var.free(); p_pl.free(); p_tasks.free();
    p_part.free(); p_wtasks.free(); p_wcols.free(); p_flags.free();
    t_lvl.free(); g_dbl.free(); g_int.free();
    t_int.free(); t_split_val.free(); p_logw.free(); p_xtest.free(); p_outmu.free(); p_outvar.free(); p_cnt.free(); p_pidx.free(); p_reach.free();
    d_prm.free(); d_trpart.free(); d_gpart.free(); d_rows.free(); d_leaf_mean.free(); d_scal.free(); d_counter.free();
    d_mask.free();
    if (pin_multi) cudaFreeHost(pin_multi);
    if (pin_scal_multi) cudaFreeHost(pin_scal_multi);
    if (pin_rows) cudaFreeHost(pin_rows);
    if (pin_scal) cudaFreeHost(pin_scal);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : oz_evs) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace dsm {
constexpr int GERR = 16;     // index of the scheduler error word inside d_counter ([0,16): task counters, cleared per batch)
cudaError_t engine_attrs();
int num_sms(int device);
template <typename T>
inline cudaError_t upload(T** dptr, const std::vector<T>& v) {
  *dptr = nullptr;
  if (v.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc(dptr, v.size() * sizeof(T));
  if (e == cudaErrorMemoryAllocation) {          // the buffers cached from destroyed handles may hold the memory: drop them, retry once
    cudaGetLastError();
    g_cache.release_all();
    e = cudaMalloc(dptr, v.size() * sizeof(T));
  }
  if (e) return e;
  return cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}
void shard_lpt(int64_t L, const int64_t* leaf_ptr, int world, int32_t* owner);
void derive_params(const dsmgp_handle* h, int kid, const double* th, double* prm);
inline float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
int32_t refine_alpha(dsmgp_handle* h);
int32_t ensure_dev_tree(dsmgp_handle* h);       // api_predict.cu: the flattened region graph in device memory
struct DevTree;
DevTree dev_tree(const dsmgp_handle* h);
// api.cu: the evaluation pipeline (gram -> potrf -> inverse -> rows) enqueued on the handle's stream
int32_t run_pipeline(dsmgp_handle* h, bool with_grad, const double* leaf_scale = nullptr, bool defer_sync = false,
                     bool naive = false, bool first = true);
int32_t finish_pipeline(dsmgp_handle* h, bool with_grad);
int32_t fetch_rows(dsmgp_handle* h);
int32_t check_pd(dsmgp_handle* h);
// potrf2 task list of one batch in topological look-ahead order, restricted to the slots with keep[slot - b.s0] != 0
std::vector<int4> build_potrf2_tasks(const dsmgp_handle* h, const Batch& b, const std::vector<char>& keep, int sms);
int32_t standalone_device_check(std::string& err);
// api_ozaki.cu
bool oz_enabled();
int32_t oz_plan(dsmgp_handle* h);
int32_t oz_run_inverse(dsmgp_handle* h, Batch& b, const Trtri3Args& full, int sms, cudaStream_t st);
int32_t oz_run_potrf(dsmgp_handle* h, Batch& b, const Potrf2Args& full, const Trtri3Args* inv, int sms, cudaStream_t st);
int32_t oz_run_lauum(dsmgp_handle* h, Batch& b, const LauumArgs& full, int sms, cudaStream_t st);
}  // namespace dsm
