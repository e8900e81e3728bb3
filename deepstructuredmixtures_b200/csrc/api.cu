// libdsmgp.so : C ABI implementation (include/dsmgp.h), core: lifetime, parameters, the fit / evaluation pipeline.
// sm_100a only, no CPU fallback.  Prediction lives in api_predict.cu, the stand-alone operators in api_ops.cu.
#include "handle.h"

using namespace dsm;

#ifndef DSM_LAUUM_GROUP_DEFAULT
#define DSM_LAUUM_GROUP_DEFAULT 4      // measured: cfg3 mathematical LAUUM phase 45.09 -> 44.95 ms, cfg4 120.4 -> 119.8 ms
#endif
#ifndef DSM_TRTRI_GROUP_DEFAULT
#define DSM_TRTRI_GROUP_DEFAULT 4      // measured on cfg3 (profiles/trtri3_default_r02.csv): DRAM 78.6 -> 41.3 GB per launch at equal kernel time
#endif

namespace dsm {
std::string& create_error() { static thread_local std::string e; return e; }
BufCache g_cache;
}
#define g_create_error (dsm::create_error())

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only: track the opt-in per device
// (handles may live on several GPUs of one process; create may be called from several threads).
static std::mutex g_attr_mu;
static uint64_t g_attr_done[4] = {0, 0, 0, 0};      // bitset indexed by device ordinal (< 256)
cudaError_t dsm::engine_attrs() {
  int dev = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev))) return e;
  std::lock_guard<std::mutex> g(g_attr_mu);
  if (dev >= 0 && dev < 256 && ((g_attr_done[dev >> 6] >> (dev & 63)) & 1)) return cudaSuccess;
  if ((e = init_v2_kernels())) return e;
  if ((e = oz_init_kernels())) return e;
  if (dev >= 0 && dev < 256) g_attr_done[dev >> 6] |= (uint64_t(1) << (dev & 63));
  return cudaSuccess;
}

int dsm::num_sms(int device) {
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

void dsm::shard_lpt(int64_t L, const int64_t* leaf_ptr, int world, int32_t* owner) {
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

void dsm::derive_params(const dsmgp_handle* h, int kid, const double* th, double* prm) {
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
// potrf2 tile tasks of a batch in topological order with look-ahead, restricted to the experts with keep[] != 0.
std::vector<int4> dsm::build_potrf2_tasks(const dsmgp_handle* h, const Batch& b, const std::vector<char>& keep, int sms) {
  struct TK { int s, grp, slot, I, J; };
  const int nb_s = b.s1 - b.s0;
  int nkeep = 0, max_nb = 0;
  for (int sl = 0; sl < nb_s; sl++) if (keep[sl]) { nkeep++; max_nb = std::max(max_nb, (int)h->meta[b.s0 + sl].nb); }
  const char* ord_env = getenv("DSMGP_ORDER");           // development A/B: 0 = end together, 1 = stretch
  const bool stretch = ord_env ? (ord_env[0] == '1') : (nkeep * 4 < sms);
  const bool start_together = ord_env && ord_env[0] == '2';     // experiment: no shift at all
  // position of the next diagonal tile inside a level: right behind the first panel tile (1: shortest critical path) or
  // behind all panel tiles of the level (3: in a large batch the first panel tile has then finished and the diagonal
  // task does not sit on an SM waiting for it)
  const char* dg_env = getenv("DSMGP_DIAG_LATE");
  const int diag_grp = (dg_env ? dg_env[0] == '1' : false) ? 3 : 1;
  std::vector<TK> tk;
  for (int s = b.s0; s < b.s1; s++) {
    const int sl = s - b.s0;
    if (!keep[sl]) continue;
    const LeafMeta& m = h->meta[s];
    const int shift = max_nb - m.nb;
    // level of block column J in the global order.  "end together" (shift) keeps the tail of a throughput-bound
    // batch parallel; "stretch" lets every expert progress proportionally through the whole launch, which spreads the
    // other experts' work evenly along the critical path of the largest one (small shards: multi-GPU strong scaling)
    auto level = [&](int J) { return start_together ? J * 1024 : stretch ? (int)(((int64_t)J * 1024 * max_nb) / m.nb) : (J + shift) * 1024; };
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
  return pt;
}

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
  // (24 GiB when the INT8 split path is on: its slice pool and scratch take ~3x the arena of a batch)
  if (!h->opts.keep_factors && h->opts.arena_bytes == 0) budget = std::min<int64_t>(budget, (oz_enabled() ? 24ll : 48ll) << 30);
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
    {
      // LAUUM tiles have no dependencies at all; tile (I, J) reads X^T of block rows I and J behind column I.  DSMGP_LAUUM_GROUP
      // = G orders them in G x G groups of the (I, J) plane (cost-descending by group, row-major inside) so that neighbours in the
      // task list -- which run at the same time on different SMs -- read the same tiles out of L2 (G = 1: every row I on its own).
      const char* lg = getenv("DSMGP_LAUUM_GROUP");
      const int G = lg ? std::max(1, atoi(lg)) : DSM_LAUUM_GROUP_DEFAULT;
      std::stable_sort(lt.begin(), lt.end(), [&](const int4& a, const int4& c) {
        const int ra = h->meta[b.s0 + a.x].nb - (a.y / G) * G, rc = h->meta[b.s0 + c.x].nb - (c.y / G) * G;
        if (ra != rc) return ra > rc;
        if (G == 1) return false;
        if (a.x != c.x) return a.x < c.x;
        if (a.y / G != c.y / G) return a.y / G < c.y / G;
        if (a.z / G != c.z / G) return a.z / G < c.z / G;
        if (a.y != c.y) return a.y < c.y;
        return a.z < c.z; });
    }
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
      std::vector<int64_t> foff(nb_s, 0);
      int64_t fo = 0;
      for (int s = b.s0; s < b.s1; s++) { foff[s - b.s0] = fo; fo += (int64_t)h->meta[s].nb * (h->meta[s].nb + 1) / 2; }
      std::vector<int4> pt = build_potrf2_tasks(h, b, std::vector<char>(nb_s, 1), sms_plan);
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
      // L2 reuse by task order (DSMGP_TRTRI_GROUP = G > 1): consecutive tasks are claimed by different SMs at nearly the same
      // time and stream their k-blocks in lockstep, so tasks that READ THE SAME TILES should be neighbours in the list.  Tile
      // (I, J) reads X^T of block column J and L of block row I: the tiles of a G x G group of the (I, J) plane share each
      // operand G ways.  Groups are ordered by THEIR anti-diagonal (still topological: a tile only depends on tiles above it in
      // its own column, which lie in the same group -- row-major inside a group -- or in a group of a smaller anti-diagonal).
      {
        const char* ge = getenv("DSMGP_TRTRI_GROUP");
        const int G = ge ? std::max(1, atoi(ge)) : (stretch ? 1 : DSM_TRTRI_GROUP_DEFAULT);
        if (G > 1) {
          struct GK { int level, slot, gi, gj, I, J; };
          std::vector<GK> gv;
          const int max_ng = (b.max_nb + G - 1) / G;
          for (int s = b.s0; s < b.s1; s++) {
            const LeafMeta& m = h->meta[s];
            const int sl = s - b.s0, ng = (m.nb + G - 1) / G, shift = max_ng - ng;
            for (int J = 0; J < m.nb; J++)
              for (int I = J + 1; I < m.nb; I++) {
                const int gi = I / G, gj = J / G, gd = gi - gj;
                const int level = start_together ? gd : stretch ? (int)(((int64_t)gd * 64 * max_ng) / ng) : (gd + shift) * 64;
                gv.push_back({level, sl, gi, gj, I, J});
              }
          }
          std::stable_sort(gv.begin(), gv.end(), [](const GK& a, const GK& c) {
            if (a.level != c.level) return a.level < c.level;
            if (a.slot != c.slot) return a.slot < c.slot;
            if (a.gj != c.gj) return a.gj < c.gj;
            if (a.I != c.I) return a.I < c.I;
            return a.J < c.J; });
          // Stagger: inside a group the tile (I+1, J) needs (I, J) for its LAST k-block; two rows of one group claimed back to
          // back reach that point together and the lower one would wait out the upper one's epilogue and store (measured:
          // +11 % kernel time with plain row-major groups).  So the rows of DSMGP_TRTRI_STAGGER (default 16) consecutive groups
          // are interleaved: row r of every group of the chunk, then row r+1, ... -- successors in a column are then dozens of
          // list positions (> 10 us of claim time) apart while the tiles they share are still in L2.  Measured (cfg3, ms of the
          // inverse): no grouping 31.92; G=4 without stagger 35.40; stagger 4: 33.94, 8: 32.48, 16: 32.01.
          const char* se = getenv("DSMGP_TRTRI_STAGGER");
          const int C = se ? std::max(1, atoi(se)) : 16;
          if (C > 1) {
            std::vector<GK> out; out.reserve(gv.size());
            size_t i = 0;
            while (i < gv.size()) {
              // chunk = up to C groups of the same level
              std::vector<std::pair<size_t, size_t>> grp;      // [begin, end) of each group in gv
              const int lvl = gv[i].level;
              size_t j = i;
              while (j < gv.size() && gv[j].level == lvl && (int)grp.size() < C) {
                size_t e = j;
                while (e < gv.size() && gv[e].level == lvl && gv[e].slot == gv[j].slot && gv[e].gj == gv[j].gj) e++;
                grp.push_back({j, e}); j = e;
              }
              std::vector<size_t> cur(grp.size());
              for (size_t q = 0; q < grp.size(); q++) cur[q] = grp[q].first;
              bool any = true;
              while (any) {
                any = false;
                for (size_t q = 0; q < grp.size(); q++) {
                  if (cur[q] >= grp[q].second) continue;
                  const int row = gv[cur[q]].I;
                  while (cur[q] < grp[q].second && gv[cur[q]].I == row) out.push_back(gv[cur[q]++]);
                  any = true;
                }
              }
              i = j;
            }
            gv.swap(out);
          }
          for (size_t i = 0; i < gv.size(); i++) iv[i] = {0, 0, gv[i].slot, gv[i].I, gv[i].J};
        }
      }
      std::vector<int4> it(iv.size());
      for (size_t i = 0; i < iv.size(); i++) it[i] = make_int4(iv[i].slot, iv[i].I, iv[i].J, 0);
      b.n_trtri3 = (int)it.size();
      b.h_trtri3 = it;
      CUDA_TRY(h, upload(&b.d_trtri3_tasks, it));
      {   // the same tiles in the order of the fused launch: level of block row I in the factorisation's order, then row-major
        std::vector<TK> mv;
        for (int s = b.s0; s < b.s1; s++) {
          const LeafMeta& m = h->meta[s];
          const int sl = s - b.s0, shift = b.max_nb - m.nb;
          for (int I = 1; I < m.nb; I++) {
            const int lvl = start_together ? I * 1024 : stretch ? (int)(((int64_t)I * 1024 * b.max_nb) / m.nb) : (I + shift) * 1024;
            for (int J = 0; J < I; J++) mv.push_back({lvl, I, sl, I, J});
          }
        }
        std::stable_sort(mv.begin(), mv.end(), [](const TK& a, const TK& c) {
          if (a.s != c.s) return a.s < c.s;
          if (a.slot != c.slot) return a.slot < c.slot;
          if (a.I != c.I) return a.I < c.I;
          return a.J < c.J; });
        std::vector<int4> mt(mv.size());
        for (size_t i = 0; i < mv.size(); i++) mt[i] = make_int4(mv[i].slot, mv[i].I, mv[i].J, 0);
        b.h_trtri3m = mt;
        CUDA_TRY(h, upload(&b.d_trtri3m_tasks, mt));
      }
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
      b.h_potrf2 = pt;
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
    b.h_lauum = lt;
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
  CUDA_TRY(h, h->d_counter.alloc(32));      // [0,16) task counters (cleared per batch), [16] scheduler error word (cleared per call)
  CUDA_TRY(h, cudaMemset(h->d_counter.p, 0, 32 * sizeof(int)));
  CUDA_TRY(h, h->d_share.alloc(std::max(ns, 1)));
  h->share.slot.assign(std::max(ns, 1), make_int4(0, 0, 0, 0));
  h->exec_slot = h->leaf_slot;
  CUDA_TRY(h, h->d_flags.alloc(std::max<int64_t>(maxFlags, 1)));
  CUDA_TRY(h, h->d_flags2.alloc(std::max<int64_t>(maxFlags, 1)));
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
      if (cudaMalloc(&d_x, (size_t)h->N * h->D * sizeof(double)) == cudaErrorMemoryAllocation) {     // cached buffers may hold the memory
        cudaGetLastError(); g_cache.release_all();
        CUDA_TRY(h, cudaMalloc(&d_x, (size_t)h->N * h->D * sizeof(double)));
      }
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
  return oz_plan(h);
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
  if (leaf_ptr[L] <= (int64_t(1) << 26) && N < (int64_t(1) << 31)) h->leaf_obs32.assign(leaf_obs, leaf_obs + leaf_ptr[L]);
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
  h->theta_global = true;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_set_leaf_params(dsmgp_handle* h, int64_t leaf, const double* theta, int64_t n) {
  if (!h) return DSMGP_ERR_ARG;
  if (leaf < 0 || leaf >= h->L || !theta || n != h->knp[h->leaf_kid[leaf]]) { h->err = "set_leaf_params: bad leaf or length"; return DSMGP_ERR_ARG; }
  std::copy(theta, theta + n, h->theta_leaf.begin() + (size_t)leaf * h->Hmax);
  cudaSetDevice(h->device);
  {   // only this expert's parameter block changes
    const int s = h->leaf_slot[leaf];
    if (s >= 0) {
      derive_params(h, h->leaf_kid[leaf], &h->theta_leaf[(size_t)leaf * h->Hmax], &h->h_prm[(size_t)s * h->pstride]);
      cudaMemcpyAsync(h->d_prm.p + (size_t)s * h->pstride, &h->h_prm[(size_t)s * h->pstride], h->pstride * sizeof(double),
                      cudaMemcpyHostToDevice, h->stream);
    }
  }
  h->fitted = false; h->have_rows = false; h->have_grad = false;
  h->theta_global = false;                 // per-expert theta: the sharing plan (same theta for source and dependent) is off
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


// Per-slot gradient mask from a leaf weight vector (NULL: every expert contributes); with the sharing plan active an
// aliased expert is never computed itself (mask 0) and its source is computed whenever either of them is wanted.
static void prepare_masks(dsmgp_handle* h, const double* leaf_scale, bool with_grad, bool shr) {
  h->use_mask = false;
  const int ns = (int)h->slot_leaf.size();
  if (!with_grad || ns == 0 || (!leaf_scale && !(shr && h->share.n_alias > 0))) return;
  for (int s = 0; s < ns; s++) h->h_mask[s] = (!leaf_scale || leaf_scale[h->slot_leaf[s]] != 0.0) ? 1 : 0;
  if (shr) {
    for (const Batch& b : h->batches)
      for (int s = b.s0; s < b.s1; s++)
        if (h->share.slot[s].x == SHARE_ALIAS) { h->h_mask[b.s0 + h->share.slot[s].y] |= h->h_mask[s]; }
    for (int s = 0; s < ns; s++) if (h->share.slot[s].x == SHARE_ALIAS) h->h_mask[s] = 0;
  }
  bool any_zero = false;
  for (int s = 0; s < ns; s++) any_zero |= !h->h_mask[s];
  if (!any_zero) return;
  cudaMemcpyAsync(h->d_mask.p, h->h_mask.data(), ns * sizeof(int), cudaMemcpyHostToDevice, h->stream);
  h->use_mask = true;
}

// gram -> potrf -> solves (-> inverse -> lauum) -> rows, batch by batch.
//   leaf_scale  finetune weights D[g,:] or null: experts with weight 0 skip the gradient kernels
//   naive       ignore the sharing plan (fit_naive!)
//   first       first pipeline of an API call: clears the scheduler error word (a later pipeline of the same call must not
//               erase the error of an earlier one; finish_pipeline reads it once at the end)
int32_t dsm::run_pipeline(dsmgp_handle* h, bool with_grad, const double* leaf_scale, bool defer_sync, bool naive, bool first) {
  cudaStream_t st = h->stream;
  const int sms = num_sms(h->device);
  const bool lau = with_grad && needs_lauum(h);
  const bool shr = !naive && !h->capturing && h->share.active && h->theta_global && (h->share.n_alias + h->share.n_prefix) > 0;
  #define EV_RECORD(e) do { if (!h->capturing) cudaEventRecord((e), st); } while (0)
  h->tm = dsmgp_timings{};
  h->oz_segs.clear(); h->oz_ev_used = 0; h->oz_ksteps = 0.0;
  h->share_applied = shr;
  prepare_masks(h, with_grad ? leaf_scale : nullptr, with_grad, shr);
  const int* mask_all = (with_grad && h->use_mask) ? h->d_mask.p : nullptr;
  const int4* share_all = shr ? h->d_share.p : nullptr;
  if (first) CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p + GERR, 0, sizeof(int), st));
  // multi-rank: the row table is the exchange unit (SUM all-reduce in place), so the rows of the other ranks' experts
  // must be zero again before every evaluation
  if (h->opts.world > 1) CUDA_TRY(h, cudaMemsetAsync(h->d_rows.p, 0, (size_t)h->L * h->row_width * sizeof(double), st));
  if (h->batches.empty()) {   // a rank that owns no leaf
    h->fitted = true; h->have_rows = true; h->have_grad = with_grad; h->rows_complete = (h->opts.world == 1);
    return DSMGP_OK;
  }
  for (size_t bi = 0; bi < h->batches.size(); bi++) {
    Batch& b = h->batches[bi];
    cudaEvent_t* ev = h->ev.data() + 8 * bi;
    const int nsl = b.s1 - b.s0;
    h->oz_l21_ready = false; h->oz_inv_tiles_done = false; h->oz_x_complete = false;
    EV_RECORD(ev[0]);
    if (nsl == 0) { for (int k = 1; k < 8; k++) EV_RECORD(ev[k]); continue; }
    const LeafMeta* meta = h->d_meta.p + b.s0;
    LeafScal* scal = h->d_scal.p + b.s0;
    const int4* share_b = share_all ? share_all + b.s0 : nullptr;
    CUDA_TRY(h, cudaMemsetAsync(scal, 0, nsl * sizeof(LeafScal), st));
    EV_RECORD(ev[1]);
    GramArgs ga{meta, h->d_xg.p, h->d_prm.p, h->d_F.p, b.d_tile_off, nsl, (int)h->D, share_b};
    launch_gram_fit(ga, b.ntiles, st);
    h->tm.launches++;
    EV_RECORD(ev[2]);
    {
      CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
      CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p, 0, 16 * sizeof(int), st));
      Potrf2Args pa{meta, h->d_F.p, h->d_W.p, h->d_WT.p, h->d_y.p, h->d_z.p, scal, h->d_trpart.p, b.d_trpart_off,
                    h->d_ldpart.p, h->d_zzpart.p, h->d_flags.p, b.d_flag_off, b.d_potrf2_tasks, b.n_potrf2,
                    h->d_counter.p + 4, h->d_counter.p + GERR, 0, share_b, nullptr};
      // One launch for the factorisation and the inverse (fused2.cuh) when the batch is small enough for the tails of two
      // separate launches to matter (multi-GPU shards; measured with every rank of a shard emulated on one GPU, slowest rank:
      // 8-way 10.19 -> 9.57 ms, 4-way 18.95 -> 18.41 ms; neutral on the full 144-expert batch, which keeps the two launches and
      // their per-phase timings): DSMGP_FUSED_EVAL=0|1 overrides.
      const char* fe = getenv("DSMGP_FUSED_EVAL");
      const bool fused = with_grad && !shr && b.n_trtri3 > 0 && !getenv("DSMGP_TRACE_FILE") && !(b.oz.active && b.oz.potrf) &&
                         (fe ? fe[0] == '1' : nsl * 3 < sms * 2);
      if (fused) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_flags2.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
        Trtri3Args ta{meta, h->d_F.p, h->d_W.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_trpart.p, b.d_trpart_off,
                      h->d_flags2.p, b.d_flag_off, h->d_apart.p, h->d_tpart.p, b.d_trtri3m_tasks, b.n_trtri3,
                      h->d_counter.p, h->d_counter.p + GERR, mask_all ? mask_all + b.s0 : nullptr};
        launch_eval2(pa, ta, std::min(sms, b.n_potrf2 + b.n_trtri3), b.d_trtri_tasks, b.n_trtri, st);
        h->tm.launches += 2;
        EV_RECORD(ev[3]); EV_RECORD(ev[4]);
      } else {
      if (shr) {
        // shared Cholesky (fit.jl:71-122): the experts that are factored on their own first, then the leading block rows
        // of the SHARE_PREFIX experts are copied from their sources and their factorisation continues behind them
        pa.tasks = b.d_potrf2_A; pa.ntasks = b.n_potrf2_A;
        if (pa.ntasks > 0) { launch_potrf2(pa, std::min(sms, pa.ntasks), st); h->tm.launches++; }
        if (b.n_prefix > 0) {
          launch_share_copy(meta, share_b, b.d_prefix_slots, b.n_prefix, b.max_jb, h->d_F.p, st);
          pa.tasks = b.d_potrf2_B; pa.ntasks = b.n_potrf2_B; pa.counter = h->d_counter.p + 5;
          launch_potrf2(pa, std::min(sms, pa.ntasks), st);
          h->tm.launches += 2;
        }
      } else {
        long long* d_trace = nullptr;
        const char* trace_file = h->capturing ? nullptr : getenv("DSMGP_TRACE_FILE");
        if (trace_file) { cudaMalloc(&d_trace, (size_t)b.n_potrf2 * 64); cudaMemsetAsync(d_trace, 0, (size_t)b.n_potrf2 * 64, st); pa.trace = d_trace; }
        if (b.oz.active && b.oz.potrf && !trace_file && !h->capturing) {
          // split factorisation, SYRK on the INT8 tensor cores; on a gradient evaluation the inverse's tile-pipeline tasks
          // ride behind the factorisation tiles of the two launches
          Trtri3Args tinv{meta, h->d_F.p, h->d_W.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_trpart.p, b.d_trpart_off,
                          h->d_flags2.p, b.d_flag_off, h->d_apart.p, h->d_tpart.p, nullptr, 0,
                          h->d_counter.p, h->d_counter.p + GERR, nullptr};
          const bool inv_too = with_grad && !mask_all;
          if (inv_too) CUDA_TRY(h, cudaMemsetAsync(h->d_flags2.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
          const int32_t rc = oz_run_potrf(h, b, pa, inv_too ? &tinv : nullptr, sms, st);
          if (rc != DSMGP_OK) return rc;
        } else
        launch_potrf2(pa, std::min(sms, b.n_potrf2), st);
        if (trace_file) {
          std::vector<long long> tr((size_t)b.n_potrf2 * 8);
          cudaMemcpyAsync(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost, st);
          cudaStreamSynchronize(st);
          if (FILE* f = fopen(trace_file, "wb")) { fwrite(tr.data(), 8, tr.size(), f); fclose(f); }
          cudaFree(d_trace);
        }
        h->tm.launches++;
      }
      EV_RECORD(ev[3]);
      if (!with_grad) {       // fit only: alpha by block back-substitution (the forward solve was fused above)
        CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
        SolveArgs sa{meta, h->d_F.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_flags.p, b.d_flag_off, b.d_solve_tasks, b.n_solve,
                     h->d_counter.p + 3, h->d_counter.p + GERR, share_b};
        launch_solve(sa, std::max(1, std::min(solve_max_ctas(sms), b.n_solve)), st);
        h->tm.launches++;
      }
      EV_RECORD(ev[4]);
      if (with_grad) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), st));
        Trtri3Args ta{meta, h->d_F.p, h->d_W.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_trpart.p, b.d_trpart_off,
                      h->d_flags.p, b.d_flag_off, h->d_apart.p, h->d_tpart.p, b.d_trtri3_tasks, b.n_trtri3,
                      h->d_counter.p, h->d_counter.p + GERR, mask_all ? mask_all + b.s0 : nullptr};
        // large experts: off-diagonal parts of the inverse as GEMMs on the INT8 tensor cores (api_ozaki.cu)
        if (b.oz.active && !mask_all && !shr && !h->capturing) {
          const int32_t rc = oz_run_inverse(h, b, ta, sms, st);
          if (rc != DSMGP_OK) return rc;
        } else {
          launch_trtri3(ta, std::max(1, std::min(sms, b.n_trtri3)), b.d_trtri_tasks, b.n_trtri, st);
          h->tm.launches += 2;
        }
      }
      }   // !fused
    }
    EV_RECORD(ev[5]);
    if (lau) {
      LauumArgs la{meta, h->d_F.p, h->d_WT.p, h->d_xg.p, h->d_alpha.p, h->d_prm.p, b.d_lauum_tasks, b.n_lauum,
                   h->d_counter.p + 1, h->d_gpart.p, b.d_gpart_off, (int)h->D, h->d_counter.p + GERR,
                   mask_all ? mask_all + b.s0 : nullptr, nullptr, nullptr};
      if (b.oz.active && b.oz.lauum && h->oz_x_complete) {      // contraction of the split experts on the INT8 tensor cores
        const int32_t rc = oz_run_lauum(h, b, la, sms, st);
        if (rc != DSMGP_OK) return rc;
      } else {
        launch_lauum3(la, std::max(1, std::min(sms, b.n_lauum)), st);
        h->tm.launches++;
      }
    }
    EV_RECORD(ev[6]);
    RowsArgs ra{meta, scal, scal, h->d_prm.p, h->d_trpart.p, b.d_trpart_off, h->d_gpart.p, b.d_gpart_off,
                h->d_rows.p, h->row_width, h->opts.as_written_grads, with_grad ? 1 : 0, lau ? 1 : 0,
                h->d_ldpart.p, h->d_zzpart.p, h->d_alpha.p, mask_all ? mask_all + b.s0 : nullptr, share_b};
    launch_rows(ra, nsl, st);
    h->tm.launches++;
    if (shr && h->share.n_alias > 0) { launch_rows_alias(meta, share_b, nsl, h->d_rows.p, h->row_width, scal, st); h->tm.launches++; }
    CUDA_TRY(h, cudaGetLastError());
    EV_RECORD(ev[7]);
    double pf = b.potrf_flops, gb = b.gram_bytes;
    if (shr) {                 // work that the plan removed is not counted
      for (int s = b.s0; s < b.s1; s++) {
        const int4 sh = h->share.slot[s];
        const double n = h->meta[s].n;
        if (sh.x == SHARE_ALIAS) { pf -= n * n * n / 3.0 + n * n / 2.0 + n / 6.0; gb -= 8.0 * (n * (n + 1) / 2.0) + 8.0 * n * h->D; }
        else if (sh.x == SHARE_PREFIX) { const double k = (double)sh.z * BLK; pf -= k * k * k / 3.0; gb -= 8.0 * k * (k + 1) / 2.0; }
      }
    }
    h->tm.potrf_flops += pf;
    h->tm.inverse_flops += with_grad ? pf : 0.0;      // the triangular inverse (inverse_ms); a LAUUM pass (grad_ms) is n^3/3 once more
    h->tm.gram_bytes += gb;
  }
  #undef EV_RECORD
  if (defer_sync) return DSMGP_OK;
  return finish_pipeline(h, with_grad);
}

int32_t dsm::finish_pipeline(dsmgp_handle* h, bool with_grad) {
  cudaStream_t st = h->stream;
  const int ns = (int)h->slot_leaf.size();
  if (ns) CUDA_TRY(h, cudaMemcpyAsync(h->pin_scal, h->d_scal.p, ns * sizeof(LeafScal), cudaMemcpyDeviceToHost, st));
  int gerr = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&gerr, h->d_counter.p + GERR, sizeof(int), cudaMemcpyDeviceToHost, st));
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
  if (!h->batches.empty()) h->tm.total_ms = ev_ms(h->ev[0], h->ev[8 * (h->batches.size() - 1) + 7]);
  std::fill(h->h_info.begin(), h->h_info.end(), 0);
  for (int s = 0; s < ns; s++) {
    int info = h->pin_scal[s].info;
    if (info > h->meta[s].n) info = 0;     // padding rows are identity
    h->h_info[h->slot_leaf[s]] = info;
  }
  // where the results of every expert live: its own slot, or its source's slot when it was aliased by the sharing plan
  h->exec_slot = h->leaf_slot;
  if (h->share_applied)
    for (const Batch& b : h->batches)
      for (int s = b.s0; s < b.s1; s++)
        if (h->share.slot[s].x == SHARE_ALIAS) h->exec_slot[h->slot_leaf[s]] = b.s0 + h->share.slot[s].y;
  h->fitted = true; h->have_rows = true; h->rows_complete = (h->opts.world == 1);
  h->have_grad = with_grad && !h->use_mask;     // a masked evaluation holds zero gradients for the skipped experts
  h->alpha_exact = !with_grad;
  return DSMGP_OK;
}

// alpha = L^-T z by block back-substitution (gaussianprocess.jl:105).  The gradient path leaves alpha = X^T z, which is
// what tr(W) needs but carries the rounding of the explicit inverse; consumers of alpha itself (predict, accessors)
// get the back-substituted vector, exactly like the reference.
int32_t dsm::refine_alpha(dsmgp_handle* h) {
  if (h->alpha_exact || !h->fitted || h->batches.size() != 1) return DSMGP_OK;
  Batch& b = h->batches[0];
  const int nsl = b.s1 - b.s0;
  if (nsl > 0) {
    CUDA_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, std::max<int64_t>(b.flag_ints, 1) * sizeof(int), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p + 3, 0, sizeof(int), h->stream));
    SolveArgs sa{h->d_meta.p, h->d_F.p, h->d_WT.p, h->d_z.p, h->d_alpha.p, h->d_flags.p, b.d_flag_off, b.d_solve_tasks, b.n_solve,
                 h->d_counter.p + 3, h->d_counter.p + GERR, h->share_applied ? h->d_share.p : nullptr};
    launch_solve(sa, std::max(1, std::min(solve_max_ctas(num_sms(h->device)), b.n_solve)), h->stream);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  h->alpha_exact = true;
  return DSMGP_OK;
}

// ------------------------------------------------------------------------------------------
// the sharing plan of fit! (fit.jl:71-122)
// ------------------------------------------------------------------------------------------
// Mirrors the reference's scheduling: every expert j picks the "main" expert i = argmax_i D[i,j] D[j,i] (fit.jl:78-84), the
// experts are visited by how often they are somebody's main (:86), a main is factored in full (:97-100) and j shares with
// it according to (D[i,j] == 1, D[j,i] == 1) and tau (:107-116, fitcontained! :124-292).  What "sharing" executes here:
//   (true, true)    identical experts                       -> SHARE_ALIAS  (the reference copies factors and alpha, :132-143)
//   (false, true)   j inside main, few rows to delete (tau) -> SHARE_PREFIX (the reference downdates by Givens sweeps, :145-206)
//   (true, false)   main is a leading part of j             -> SHARE_PREFIX (the reference continues the factor, :208-292)
// SHARE_PREFIX reuses the factor tiles of the leading 128-row blocks the two experts have in common and re-factors what
// follows the first differing observation (= chol_continue! from there).  That is the exact factor of update_cholesky!
// (the reference's own downdate is numerically wrong, SURVEY App. B Q7, and a Givens sweep is a serial chain of n column
// steps per deleted row -- orders of magnitude slower on this machine than re-factoring the trailing blocks on DMMA).
static int64_t sorted_diff_count(const int32_t* a, int64_t na, const int32_t* b, int64_t nb) {   // |a \ b|, both ascending
  int64_t i = 0, j = 0, c = 0;
  while (i < na) {
    while (j < nb && b[j] < a[i]) j++;
    if (j >= nb || b[j] != a[i]) c++;
    i++;
  }
  return c;
}

// The plan itself (host only): kind / source / copied block rows per leaf.
static void plan_sharing(int64_t L, const int64_t* leaf_ptr, const int32_t* obs, const int32_t* leaf_kid, const double* overlap,
                         double tau, int* kind, int* source, int* blocks) {
  auto lobs = [&](int64_t l) { return obs + leaf_ptr[l]; };
  auto ln = [&](int64_t l) { return leaf_ptr[l + 1] - leaf_ptr[l]; };
  std::fill(kind, kind + L, 0); std::fill(source, source + L, -1); std::fill(blocks, blocks + L, 0);
  // fit.jl:78-84: main of every expert, and how often an expert is a main
  std::vector<int64_t> S(L), counts(L, 0);
  for (int64_t j = 0; j < L; j++) {
    double best = -std::numeric_limits<double>::infinity(); int64_t bi = 0;
    for (int64_t i = 0; i < L; i++) {
      const double v = overlap[i + j * L] * overlap[j + i * L];          // D[:,j] .* D[j,:]
      if (v > best) { best = v; bi = i; }                                 // argmax: first maximum
    }
    S[j] = bi; counts[bi]++;
  }
  std::vector<int64_t> order(L);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t c) { return counts[a] < counts[c]; });     // :86
  std::vector<char> processed(L, 0);
  for (int64_t j : order) {
    if (processed[j]) continue;                                           // :90
    const int64_t i = S[j];
    processed[i] = 1; processed[j] = 1;                                   // :97-103
    if (i == j) continue;
    if (leaf_kid[i] != leaf_kid[j]) continue;                             // :107-109
    const int32_t *oj = lobs(j), *oi = lobs(i);
    const int64_t nj = ln(j), ni = ln(i);
    if (oj[0] < oi[0]) continue;                                          // :110-112
    const bool ione = overlap[i + j * L] == 1.0, jone = overlap[j + i * L] == 1.0;
    // resolve the source through an alias (the main may itself have been aliased to an earlier main)
    int64_t src = i;
    if (kind[src] == SHARE_ALIAS) src = source[src];
    if (ione && jone) {                                                   // :132-143 copy
      if (nj == ni && std::equal(oj, oj + nj, oi)) { kind[j] = SHARE_ALIAS; source[j] = (int)src; }
      continue;
    }
    bool share = false;
    if (!ione && jone) {                                                  // :145-206 j inside main: rows to delete behind tau
      if (oj[0] >= oi[0] && oj[nj - 1] <= oi[ni - 1]) {
        const int64_t e = std::upper_bound(oi, oi + ni, oj[nj - 1]) - oi;       // main.obs[1:e]
        const int64_t ndel = sorted_diff_count(oi, e, oj, nj);
        share = (double)ndel / (double)nj < tau;
      }
    } else if (ione && !jone) {                                           // :208-292 main is (part of) the head of j
      if (oj[0] >= oi[0] && oj[nj - 1] >= oi[ni - 1]) {
        const int64_t k1 = std::upper_bound(oj, oj + nj, oi[ni - 1]) - oj;      // s1 = j.obs[1:findfirst(== maxM)]
        const int64_t s = std::lower_bound(oi, oi + ni, oj[0]) - oi;            // idx = s:e
        const bool minsame = oj[0] == oi[0];
        if (!((k1 != ni - s) && minsame)) {                               // :246-248
          const int64_t ndel = sorted_diff_count(oi, ni, oj, k1);
          share = (double)ndel / (double)nj < tau;
        }
      }
    }
    if (!share) continue;
    if (kind[src] == SHARE_PREFIX) continue;              // the source itself waits for a copy: factor j on its own
    const int32_t* os = lobs(src);
    const int64_t nsr = ln(src);
    int64_t k = 0;
    while (k < nj && k < nsr && oj[k] == os[k]) k++;
    const int jb = (int)(k / BLK);
    if (jb < 1) continue;
    kind[j] = SHARE_PREFIX; source[j] = (int)src; blocks[j] = jb;
  }
}

// HOST-ONLY: the plan for a given structure (what dsmgp_set_sharing stores), for inspection and CPU tests.
extern "C" int32_t dsmgp_host_sharing_plan(int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs, const int32_t* leaf_kernel_id,
                                           const double* overlap, double tau, int32_t* kind, int32_t* source, int32_t* blocks) {
  if (L <= 0 || !leaf_ptr || !leaf_obs || !leaf_kernel_id || !overlap || !kind || !source || !blocks) return DSMGP_ERR_ARG;
  std::vector<int32_t> obs(leaf_obs, leaf_obs + leaf_ptr[L]);
  plan_sharing(L, leaf_ptr, obs.data(), leaf_kernel_id, overlap, tau, kind, source, blocks);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_set_sharing(dsmgp_handle* h, const double* overlap, double tau) {
  if (!h) return DSMGP_ERR_ARG;
  auto& sp = h->share;
  const int64_t L = h->L;
  const int ns = (int)h->slot_leaf.size();
  sp = dsmgp_handle::Share{};
  sp.tau = tau;
  sp.slot.assign(std::max(ns, 1), make_int4(0, 0, 0, 0));
  sp.leaf_kind.assign(L, 0); sp.leaf_src.assign(L, -1);
  for (auto& b : h->batches) {
    cudaFree(b.d_potrf2_A); cudaFree(b.d_potrf2_B); cudaFree(b.d_prefix_slots);
    b.d_potrf2_A = b.d_potrf2_B = nullptr; b.d_prefix_slots = nullptr; b.n_potrf2_A = b.n_potrf2_B = b.n_prefix = b.max_jb = 0;
  }
  if (!overlap) return DSMGP_OK;                       // fit_naive!: no sharing
  if (h->leaf_obs32.empty()) { h->err = "set_sharing: observation lists too large to keep (sharing unavailable)"; return DSMGP_ERR_STATE; }
  cudaSetDevice(h->device);
  std::vector<int> blocks(L, 0);
  plan_sharing(L, h->leaf_ptr.data(), h->leaf_obs32.data(), h->leaf_kid.data(), overlap, tau, sp.leaf_kind.data(), sp.leaf_src.data(), blocks.data());
  // map to local slots: source and dependent must live in the same batch of the same rank
  std::vector<int> slot_batch(std::max(ns, 1), -1);
  for (size_t bi = 0; bi < h->batches.size(); bi++) for (int s = h->batches[bi].s0; s < h->batches[bi].s1; s++) slot_batch[s] = (int)bi;
  // Continuing behind copied block rows costs a copy kernel and a second Cholesky launch per batch (~0.3 ms of launch and tail
  // latency, measured on the 10,000 x 1 model): the branch is only taken when the factorisation work it removes from a batch is
  // worth more than that (DSMGP_SHARE_MIN_FLOPS, default 1.5e10 flop ~ 0.5 ms at 30 TFLOP/s).  Aliases always pay.
  {
    const char* mf = getenv("DSMGP_SHARE_MIN_FLOPS");
    const double min_flops = mf ? atof(mf) : 1.5e10;
    std::vector<double> saved(h->batches.size(), 0.0);
    for (int64_t l = 0; l < L; l++) {
      const int s = h->leaf_slot[l];
      if (sp.leaf_kind[l] != SHARE_PREFIX || s < 0) continue;
      const int ss = h->leaf_slot[sp.leaf_src[l]];
      if (ss < 0 || slot_batch[s] != slot_batch[ss]) continue;
      const double k = (double)blocks[l] * BLK;
      saved[slot_batch[s]] += k * k * k / 3.0;
    }
    for (int64_t l = 0; l < L; l++) {
      const int s = h->leaf_slot[l];
      if (sp.leaf_kind[l] == SHARE_PREFIX && s >= 0 && saved[slot_batch[s]] < min_flops) { sp.leaf_kind[l] = SHARE_NONE; sp.leaf_src[l] = -1; blocks[l] = 0; }
    }
  }
  for (int64_t l = 0; l < L; l++) {
    const int s = h->leaf_slot[l];
    if (sp.leaf_kind[l] == SHARE_NONE) continue;
    const int ss = h->leaf_slot[sp.leaf_src[l]];
    if (s < 0 || ss < 0 || slot_batch[s] != slot_batch[ss]) { if (s >= 0 || ss >= 0) { /* not co-located: factor on its own */ } sp.leaf_kind[l] = SHARE_NONE; sp.leaf_src[l] = -1; continue; }
    const Batch& b = h->batches[slot_batch[s]];
    sp.slot[s] = make_int4(sp.leaf_kind[l], ss - b.s0, blocks[l], 0);
    const double n = (double)h->meta[s].n;
    if (sp.leaf_kind[l] == SHARE_ALIAS) { sp.n_alias++; sp.flops_saved += n * n * n / 3.0; }
    else { sp.n_prefix++; sp.blocks_copied += blocks[l]; const double k = (double)blocks[l] * BLK; sp.flops_saved += k * k * k / 3.0; }
  }
  // an alias whose source was demoted above keeps pointing at a valid local source or was demoted with it; a prefix expert
  // whose source is an alias was resolved before.  Task lists per batch.
  const int sms = num_sms(h->device);
  for (auto& b : h->batches) {
    const int nb_s = b.s1 - b.s0;
    std::vector<char> keepA(nb_s, 0), keepB(nb_s, 0);
    std::vector<int> pslots;
    for (int s = b.s0; s < b.s1; s++) {
      const int4 sh = sp.slot[s];
      if (sh.x == SHARE_NONE) keepA[s - b.s0] = 1;
      else if (sh.x == SHARE_PREFIX) { keepB[s - b.s0] = 1; pslots.push_back(s - b.s0); b.max_jb = std::max(b.max_jb, sh.z); }
    }
    std::vector<int4> ta = build_potrf2_tasks(h, b, keepA, sms), tb = build_potrf2_tasks(h, b, keepB, sms);
    b.n_potrf2_A = (int)ta.size(); b.n_potrf2_B = (int)tb.size(); b.n_prefix = (int)pslots.size();
    CUDA_TRY(h, upload(&b.d_potrf2_A, ta));
    CUDA_TRY(h, upload(&b.d_potrf2_B, tb));
    CUDA_TRY(h, upload(&b.d_prefix_slots, pslots));
  }
  if (ns) CUDA_TRY(h, cudaMemcpy(h->d_share.p, sp.slot.data(), ns * sizeof(int4), cudaMemcpyHostToDevice));
  sp.active = true;
  h->fitted = false; h->have_rows = false; h->have_grad = false;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_get_sharing(const dsmgp_handle* h, int32_t* kind, int32_t* source, int32_t* blocks) {
  if (!h) return DSMGP_ERR_ARG;
  for (int64_t l = 0; l < h->L; l++) {
    const bool on = h->share.active && !h->share.leaf_kind.empty();
    const int k = on ? h->share.leaf_kind[l] : 0;
    if (kind) kind[l] = k;
    if (source) source[l] = (on && k != SHARE_NONE) ? h->share.leaf_src[l] : -1;
    if (blocks) { const int s = h->leaf_slot[l]; blocks[l] = (on && s >= 0 && k == SHARE_PREFIX) ? h->share.slot[s].z : 0; }
  }
  return DSMGP_OK;
}

// After the pipeline: per-leaf rows on the host.  With a communicator (dsmgp_comm_init) the table is first assembled by a
// SUM all-reduce on the library's stream (every rank filled only its own experts' rows).
int32_t dsm::fetch_rows(dsmgp_handle* h) {
  const size_t count = (size_t)h->L * h->row_width;
  if (h->opts.world > 1 && h->comm != nullptr && !h->rows_complete) {
    int32_t rc = comm_allreduce_sum(h, h->d_rows.p, count);
    if (rc) return rc;
    h->rows_complete = true;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->pin_rows, h->d_rows.p, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  memcpy(h->h_rows.data(), h->pin_rows, count * sizeof(double));
  return DSMGP_OK;
}

int32_t dsm::check_pd(dsmgp_handle* h) {
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

static const char* kNeedComm = "rows of other ranks missing: call dsmgp_comm_init, or all-reduce the rows yourself (dsmgp_eval_local_dev / dsmgp_eval_finish_dev)";

extern "C" int32_t dsmgp_fit(dsmgp_handle* h, double tau, const double* overlap, int32_t* info, double* seconds) {
  if (!h) return DSMGP_ERR_ARG;
  cudaSetDevice(h->device);
  int32_t rc;
  if (overlap && (rc = dsmgp_set_sharing(h, overlap, tau))) return rc;      // fit!(spn, D, gpmap; tau)
  rc = run_pipeline(h, false, nullptr, false, /*naive=*/overlap == nullptr);  // overlap == NULL: fit_naive!
  if (rc) return rc;
  if ((rc = fetch_rows(h))) return rc;
  if (info) std::copy(h->h_info.begin(), h->h_info.end(), info);
  if (seconds) *seconds = h->tm.total_ms * 1e-3;
  return check_pd(h);
}

extern "C" int32_t dsmgp_lml(dsmgp_handle* h, double* node_lml) {
  if (!h) return DSMGP_ERR_ARG;
  if (!h->have_rows) { h->err = "lml: call fit or eval first"; return DSMGP_ERR_STATE; }
  if (!h->rows_complete) { h->err = std::string("lml: ") + kNeedComm; return DSMGP_ERR_STATE; }
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
  if (!h->rows_complete) { h->err = std::string("grad: ") + kNeedComm; return DSMGP_ERR_STATE; }
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
  // finetune: experts with zero overlap weight skip the gradient kernels
  rc = run_pipeline(h, grad != nullptr, grad != nullptr ? leaf_scale : nullptr);
  if (rc) return rc;
  if ((rc = fetch_rows(h))) return rc;
  if (!h->rows_complete) { h->err = std::string("eval: ") + kNeedComm; return DSMGP_ERR_STATE; }
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
    rc = run_pipeline(h, true, scale.data(), /*defer_sync=*/true, /*naive=*/false, /*first=*/g == 0);
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

extern "C" int32_t dsmgp_infer(dsmgp_handle* h, double* sum_logweights, double* z) {
  if (!h) return DSMGP_ERR_ARG;
  if (!h->have_rows || !h->rows_complete) { h->err = "infer: call fit or eval first"; return DSMGP_ERR_STATE; }
  h->sum_logw.assign(h->tree.child_ptr[h->tree.n_nodes], 0.0);
  infer_weights(h->tree, h->h_rows.data(), h->row_width, h->sum_logw.data(), z);
  h->have_weights = true;
  if (sum_logweights) std::copy(h->sum_logw.begin(), h->sum_logw.end(), sum_logweights);
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_reset_weights(dsmgp_handle* h, double* sum_logweights) {
  if (!h) return DSMGP_ERR_ARG;
  h->sum_logw.assign(h->tree.child_ptr[h->tree.n_nodes], 0.0);
  reset_weights(h->tree, h->sum_logw.data());
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
  *slot = h->exec_slot[leaf];              // an aliased expert reads its source's results (fit.jl:132-143)
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
