// Split triangular inverse with the off-diagonal products on the INT8 tensor cores (ozaki_args.h, ozaki.cuh).
//
// For an expert with nb block rows the block range [lo, hi) is split at mid (recursively, DSMGP_OZAKI_DEPTH levels):
//     X[mid:hi, lo:mid] = -X22 (L21 X11),   X11 = X[lo:mid, lo:mid],  X22 = X[mid:hi, mid:hi],  L21 = L[mid:hi, lo:mid].
// The tile pipeline (trtri3.cuh) computes only the tiles whose row and column block lie in the same leaf range; every
// split then costs two batched GEMM launches with their slicing passes, deepest level first:
//     stage 1:  T^T = X11^T L21^T      (A = rows of X^T in the upper block triangle / W^T, B = rows of L; K = [Jb, mid))
//     stage 2:  X21^T = -T^T X22^T     (A = rows of T^T, B = rows of X22 read transposed / W; K = [mid, Ib]), written
//               into the upper block triangle of the factor arena where the tile pipeline would have put it.
// The fused partials of those tiles (||X_IJ||^2, X_IJ^T z_I) come from a small kernel afterwards.
#include "handle.h"

namespace dsm {

namespace {
struct Split { int slot, lo, mid, hi; };

void collect_splits(int slot, int lo, int hi, int depth, int maxdepth, int min_nb, std::vector<std::vector<Split>>& levels,
                    std::vector<int>& range_of, int& nranges) {
  if (depth >= maxdepth || hi - lo < min_nb) {
    for (int b = lo; b < hi; b++) range_of[b] = nranges;
    nranges++;
    return;
  }
  const int mid = lo + (hi - lo) / 2;
  collect_splits(slot, lo, mid, depth + 1, maxdepth, min_nb, levels, range_of, nranges);
  collect_splits(slot, mid, hi, depth + 1, maxdepth, min_nb, levels, range_of, nranges);
  levels[depth].push_back({slot, lo, mid, hi});
}
}  // namespace

// DSMGP_OZAKI=0 keeps every evaluation on the FP64 (DMMA) tile pipelines
bool oz_enabled() {
  const char* en = getenv("DSMGP_OZAKI");
  return !(en && en[0] == '0');
}

// All device lists of one batch's plan go into ONE allocation (one cudaMalloc + one copy at create, one cudaFree at destroy:
// twenty small allocations per batch cost more than the plan itself in a create / evaluate / destroy cycle).
struct PlanBlob {
  std::vector<unsigned char> host;
  struct Fix { void** dst; size_t off; };
  std::vector<Fix> fixes;
  template <typename T>
  void add(T** dptr, const std::vector<T>& v) {
    *dptr = nullptr;
    if (v.empty()) return;
    const size_t off = (host.size() + 255) & ~size_t(255);
    host.resize(off + v.size() * sizeof(T));
    memcpy(host.data() + off, v.data(), v.size() * sizeof(T));
    fixes.push_back({reinterpret_cast<void**>(dptr), off});
  }
  cudaError_t commit(void** base) {
    *base = nullptr;
    if (host.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc(base, host.size());
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); g_cache.release_all(); e = cudaMalloc(base, host.size()); }
    if (e) return e;
    if ((e = cudaMemcpy(*base, host.data(), host.size(), cudaMemcpyHostToDevice))) return e;
    for (const Fix& f : fixes) *f.dst = static_cast<unsigned char*>(*base) + f.off;
    return cudaSuccess;
  }
};

int oz_env_slices() {
  const char* e = getenv("DSMGP_OZAKI_SLICES");
  const int s = e ? atoi(e) : 8;
  return s == 7 ? 7 : 8;
}

// Builds the plan of every batch (sizes first, then the job / tile lists with final pointers).  Called at the end of
// create; a handle without large experts, with several ranks' masks etc. simply keeps oz.active == false.
int32_t oz_plan(dsmgp_handle* h) {
  if (!oz_enabled()) return DSMGP_OK;
  const char* de = getenv("DSMGP_OZAKI_DEPTH");
  const char* me = getenv("DSMGP_OZAKI_MIN_NB");
  const int maxdepth = de ? std::max(1, std::min(4, atoi(de))) : 1;
  const int min_nb = me ? std::max(2, atoi(me)) : 8;
  const int S = oz_env_slices();
  h->oz_S = S;
  struct Tmp { std::vector<std::vector<Split>> levels; std::vector<std::vector<int>> range_of; };
  std::vector<Tmp> tmp(h->batches.size());
  int64_t max_pool = 0, max_scratch = 0, max_scale = 0, max_l21_pool = 0, max_l21_scale = 0;
  const char* pe = getenv("DSMGP_OZAKI_POTRF");
  const bool want_potrf = !(pe && pe[0] == '0');
  const char* le = getenv("DSMGP_OZAKI_LAUUM");
  bool lauum_kernels = false;
  for (int k = 0; k < h->nk; k++) {
    const int ty = h->kernels[k].type;
    lauum_kernels |= ty == DSMGP_ISO_SE || ty == DSMGP_ARD_LINEAR || (ty == DSMGP_ARD_SE && !h->opts.as_written_grads);
  }
  const bool want_lauum = lauum_kernels && !(le && le[0] == '0');
  // pass 1: splits and sizes
  for (size_t bi = 0; bi < h->batches.size(); bi++) {
    Batch& b = h->batches[bi];
    Tmp& t = tmp[bi];
    t.levels.assign(maxdepth, {});
    t.range_of.resize(b.s1 - b.s0);
    bool any = false;
    for (int s = b.s0; s < b.s1; s++) {
      const int nb = h->meta[s].nb;
      t.range_of[s - b.s0].assign(nb, 0);
      int nr = 0;
      collect_splits(s - b.s0, 0, nb, 0, maxdepth, min_nb, t.levels, t.range_of[s - b.s0], nr);
      any |= nr > 1;
    }
    if (!any) continue;
    b.oz.active = true;
    for (auto& lv : t.levels) {
      int64_t pool = 0, scratch = 0, scale = 0;
      for (const Split& sp : lv) {
        const int64_t J1 = sp.mid - sp.lo, J2 = sp.hi - sp.mid;
        pool += (J1 * J1 + 2 * J1 * J2 + J2 * J2) * OZ_KSTEPS_PER_BLK * S;     // slice tiles: X11^T, L21, T^T, X22
        scratch += J1 * J2 * WBLK_D;
        scale += 2 * (J1 + J2) * BLK;
      }
      max_pool = std::max(max_pool, pool); max_scratch = std::max(max_scratch, scratch); max_scale = std::max(max_scale, scale);
    }
    {   // LAUUM on the split path: every F^-1 tile of the split experts goes through the scratch
      int64_t sc = 0;
      for (int s = b.s0; s < b.s1; s++)
        if (t.range_of[s - b.s0].back() != t.range_of[s - b.s0].front()) sc += (int64_t)h->meta[s].nb * (h->meta[s].nb + 1) / 2 * WBLK_D;
      if (want_lauum) max_scratch = std::max(max_scratch, sc);
    }
    // the L21 slices of the root splits live in their own region: the factorisation phase makes them, the inverse reuses them
    int64_t lp = 0, ls = 0;
    for (const Split& sp : t.levels[0]) { lp += (int64_t)(sp.hi - sp.mid) * (sp.mid - sp.lo) * OZ_KSTEPS_PER_BLK * S; ls += (int64_t)(sp.hi - sp.mid) * BLK; }
    max_l21_pool = std::max(max_l21_pool, lp); max_l21_scale = std::max(max_l21_scale, ls);
  }
  if (max_pool == 0) return DSMGP_OK;
  {   // memory guard: the slice pool costs ~2.5x the factor arena of a batch; without room the FP64 pipelines stay in charge
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(h, cudaMemGetInfo(&free_b, &total_b));
    free_b += g_cache.cached_bytes(h->device);
    const double need = (double)(max_pool + max_l21_pool) * OZ_TILE_B + (double)max_scratch * 8 + (double)(max_scale + max_l21_scale) * 16;
    if (need > 0.9 * (double)free_b) {
      for (Batch& b : h->batches) b.oz.active = false;
      return DSMGP_OK;
    }
  }
  CUDA_TRY(h, h->oz_pool.alloc((size_t)(max_pool + max_l21_pool) * OZ_TILE_B));
  CUDA_TRY(h, h->oz_scratch.alloc((size_t)max_scratch));
  CUDA_TRY(h, h->oz_scale.alloc((size_t)(max_scale + max_l21_scale)));
  CUDA_TRY(h, h->oz_rowmax.alloc((size_t)(max_scale + max_l21_scale)));
  h->oz_pool_bytes = (int64_t)(max_pool + max_l21_pool) * OZ_TILE_B;
  if (oz_make_map(h->oz_map, h->oz_pool.p, (size_t)(max_pool + max_l21_pool) * OZ_TILE_B, 32 * oz_round_slices(S, 0)) != 0 ||
      oz_make_map(h->oz_map + 128, h->oz_pool.p, (size_t)(max_pool + max_l21_pool) * OZ_TILE_B, 32 * oz_round_slices(S, 1)) != 0) {
    // no tensor-map encoder in this driver: the FP64 pipelines stay in charge
    for (Batch& b : h->batches) b.oz.active = false;
    h->oz_pool.free(); h->oz_scratch.free(); h->oz_scale.free(); h->oz_rowmax.free(); h->oz_pool_bytes = 0;
    return DSMGP_OK;
  }
  // pass 2: lists
  for (size_t bi = 0; bi < h->batches.size(); bi++) {
    Batch& b = h->batches[bi];
    if (!b.oz.active) continue;
    Tmp& t = tmp[bi];
    PlanBlob blob;
    // tile-pipeline tasks restricted to the leaf ranges (order of the full list kept: still topological)
    std::vector<int4> keep;
    for (const int4& tk : b.h_trtri3) {
      const std::vector<int>& ro = t.range_of[tk.x];
      if (ro[tk.y] == ro[tk.z]) keep.push_back(tk);
    }
    for (int s = b.s0; s < b.s1; s++) {                 // rows of every diagonal range -> flops that stay on the tile pipelines
      const LeafMeta& m = h->meta[s];
      const std::vector<int>& ro = t.range_of[s - b.s0];
      double rows = 0; int cur = ro.empty() ? -1 : ro[0];
      for (int blk = 0; blk <= m.nb; blk++) {
        if (blk == m.nb || ro[blk] != cur) { b.oz.tile_flops += 2.0 / 3.0 * rows * rows * rows; rows = 0; if (blk < m.nb) cur = ro[blk]; }
        if (blk < m.nb) rows += std::min(BLK, m.n - blk * BLK > 0 ? m.n - blk * BLK : 0);
      }
    }
    b.oz.n_tasks = (int)keep.size();
    blob.add(&b.oz.d_tasks, keep);
    std::vector<OzPart> parts;
    double flops = 0.0;
    std::vector<OzJob> jL, jT; std::vector<OzTile> tS, tT;
    int64_t poolT = 0; int scaleT = 0;
    std::vector<int> kskip(b.s1 - b.s0, 0);
    int64_t l21_pool = max_pool; int l21_scale = (int)max_scale;
    for (int lv = maxdepth - 1; lv >= 0; lv--) {         // deepest level first
      if (t.levels[lv].empty()) continue;
      OzLevel& L = b.oz.levels[b.oz.n_levels++];
      std::vector<OzJob> j1, j2; std::vector<OzTile> t1, t2;
      int64_t pool = 0, scratch = 0; int scale = 0;
      for (const Split& sp : t.levels[lv]) {
        const LeafMeta& m = h->meta[b.s0 + sp.slot];
        const int J1 = sp.mid - sp.lo, J2 = sp.hi - sp.mid, KB = OZ_KSTEPS_PER_BLK;
        const double* F = h->d_F.p + m.foff;
        const double* W = h->d_W.p + m.woff; const double* WT = h->d_WT.p + m.woff;
        auto wid = [&](int blk) { const int w = m.np - blk * BLK; return w < BLK ? w : BLK; };
        auto fblk = [&](int rb, int cb) { return F + tile_off(rb, cb * (BLK / KC), m.nkc); };
        // operand pools (tile indices) and scales
        const int64_t pA1 = pool; pool += (int64_t)J1 * J1 * KB * S;      // X11^T : J1 row blocks x J1 k blocks
        int64_t pB1;                                                      // L21   : J2 x J1
        if (lv == 0) { pB1 = l21_pool; l21_pool += (int64_t)J2 * J1 * KB * S; } else { pB1 = pool; pool += (int64_t)J2 * J1 * KB * S; }
        const int64_t pA2 = pool; pool += (int64_t)J1 * J2 * KB * S;      // T^T   : J1 x J2
        const int64_t pB2 = pool; pool += (int64_t)J2 * J2 * KB * S;      // X22   : J2 x J2
        const int sA1 = scale; scale += J1 * BLK;
        int sB1;
        if (lv == 0) { sB1 = l21_scale; l21_scale += J2 * BLK; } else { sB1 = scale; scale += J2 * BLK; }
        const int sA2 = scale; scale += J1 * BLK;
        const int sB2 = scale; scale += J2 * BLK;
        double* T = h->oz_scratch.p + scratch; scratch += (int64_t)J1 * J2 * WBLK_D;
        if (lv == 0) kskip[sp.slot] = sp.mid;
        if (lv == 0 && maxdepth == 1) {
          // L21 = A21 X11^T.  Operand slices live in the (still unused) level region of the pool: A21 rows i over k < mid,
          // X11 rows j over k <= j (stored transposed above the diagonal, W_J on it)
          const int64_t qA = poolT; poolT += (int64_t)J2 * J1 * KB * S;
          const int64_t qX = poolT; poolT += (int64_t)J1 * J1 * KB * S;
          const int uA = scaleT; scaleT += J2 * BLK;
          const int uX = scaleT; scaleT += J1 * BLK;
          for (int c = 0; c < J2; c++)
            for (int k = 0; k < J1; k++)
              jT.push_back({fblk(sp.mid + c, k), 0, wid(sp.mid + c), BLK, uA + c * BLK, qA + ((int64_t)c * J1 + k) * KB * S});
          for (int a = 0; a < J1; a++) {
            for (int k = 0; k <= a; k++) {
              if (k == a) jT.push_back({W + (int64_t)a * WBLK_D, 0, BLK, BLK, uX + a * BLK, qX + ((int64_t)a * J1 + k) * KB * S});
              else jT.push_back({fblk(k, a), 1, BLK, BLK, uX + a * BLK, qX + ((int64_t)a * J1 + k) * KB * S});
            }
            for (int c = 0; c < J2; c++)
              tT.push_back({(int)(qA + (int64_t)c * J1 * KB * S), (int)(qX + (int64_t)a * J1 * KB * S), 0, (a + 1) * KB, uA + c * BLK, uX + a * BLK,
                            wid(sp.mid + c), BLK, const_cast<double*>(fblk(sp.mid + c, a)), 1.0, 0, 0});
          }
        }
        for (int a = 0; a < J1; a++) {                                       // a = Jb - lo
          const int Jb = sp.lo + a;
          for (int k = a; k < J1; k++) {                                     // X^T block (Jb, Kb), Kb >= Jb
            const int Kb = sp.lo + k;
            j1.push_back({Kb == Jb ? WT + (int64_t)Jb * WBLK_D : fblk(Jb, Kb), 0, BLK, BLK, sA1 + a * BLK, pA1 + ((int64_t)a * J1 + k) * KB * S});
          }
          for (int c = 0; c < J2; c++) {
            const int Ib = sp.mid + c;
            t1.push_back({(int)(pA1 + (int64_t)a * J1 * KB * S), (int)(pB1 + (int64_t)c * J1 * KB * S), a * KB, J1 * KB, sA1 + a * BLK, sB1 + c * BLK,
                          BLK, wid(Ib), T + ((int64_t)a * J2 + c) * WBLK_D, 1.0, 0, 0});
            j2.push_back({T + ((int64_t)a * J2 + c) * WBLK_D, 0, BLK, wid(Ib), sA2 + a * BLK, pA2 + ((int64_t)a * J2 + c) * KB * S});
            // k-steps of a half-wide last block beyond its width hold zeros: stop before them
            const int kend = c * KB + (wid(Ib) + OZ_KSTEP - 1) / OZ_KSTEP;
            t2.push_back({(int)(pA2 + (int64_t)a * J2 * KB * S), (int)(pB2 + (int64_t)c * J2 * KB * S), 0, kend, sA2 + a * BLK, sB2 + c * BLK,
                          BLK, wid(Ib), const_cast<double*>(fblk(Jb, Ib)), -1.0, 0, 0});
            parts.push_back({sp.slot, Ib, Jb});
            flops += 2.0 * BLK * BLK * OZ_KSTEP * ((double)(J1 - a) * KB + kend);
          }
        }
        for (int c = 0; c < J2; c++) {
          const int Ib = sp.mid + c;
          for (int k = 0; k < J1; k++)                                       // L block (Ib, Kb)
            (lv == 0 ? jL : j1).push_back({fblk(Ib, sp.lo + k), 0, wid(Ib), BLK, sB1 + c * BLK, pB1 + ((int64_t)c * J1 + k) * KB * S});
          if (lv == 0)                                                       // A22 -= L21 L21^T, lower block triangle (root split: lo = 0)
            for (int c2 = 0; c2 <= c; c2++)
              tS.push_back({(int)(pB1 + (int64_t)c * J1 * KB * S), (int)(pB1 + (int64_t)c2 * J1 * KB * S), 0, J1 * KB, sB1 + c * BLK, sB1 + c2 * BLK,
                            wid(Ib), wid(sp.mid + c2), const_cast<double*>(fblk(Ib, sp.mid + c2)), -1.0, 1, 0});
          for (int k = 0; k <= c; k++) {                                     // X22 block (Ib, Kb), Kb <= Ib: stored transposed above the diagonal
            const int Kb = sp.mid + k;
            if (Kb == Ib) j2.push_back({W + (int64_t)Ib * WBLK_D, 0, wid(Ib), wid(Ib), sB2 + c * BLK, pB2 + ((int64_t)c * J2 + k) * KB * S});
            else j2.push_back({fblk(Kb, Ib), 1, wid(Ib), BLK, sB2 + c * BLK, pB2 + ((int64_t)c * J2 + k) * KB * S});
          }
        }
      }
      // long blocks first: the hardware hands CTAs out in grid order
      auto by_len = [](const OzTile& x, const OzTile& y) { return (x.k1 - x.k0) > (y.k1 - y.k0); };
      std::stable_sort(t1.begin(), t1.end(), by_len);
      std::stable_sort(t2.begin(), t2.end(), by_len);
      for (const OzTile& x : t1) L.ksteps1 += x.k1 - x.k0;
      for (const OzTile& x : t2) L.ksteps2 += x.k1 - x.k0;
      L.n_jobs1 = (int)j1.size(); L.n_jobs2 = (int)j2.size(); L.n_tiles1 = (int)t1.size(); L.n_tiles2 = (int)t2.size(); L.n_scale = scale;
      blob.add(&L.d_jobs1, j1); blob.add(&L.d_jobs2, j2);
      blob.add(&L.d_tiles1, t1); blob.add(&L.d_tiles2, t2);
    }
    {   // factorisation phase of the root splits
      auto by_len = [](const OzTile& x, const OzTile& y) { return (x.k1 - x.k0) > (y.k1 - y.k0); };
      std::stable_sort(tS.begin(), tS.end(), by_len);
      for (const OzTile& x : tS) b.oz.ksteps_syrk += x.k1 - x.k0;
      for (const OzTile& x : tT) b.oz.ksteps_T += x.k1 - x.k0;
      b.oz.n_jobsL = (int)jL.size(); b.oz.n_syrk = (int)tS.size();
      b.oz.l21_scale0 = (int)max_scale; b.oz.l21_nscale = l21_scale - (int)max_scale;
      blob.add(&b.oz.d_jobsL, jL); blob.add(&b.oz.d_syrk, tS);
      std::vector<int4> pA, pB;
      for (const int4& tk : b.h_potrf2) (kskip[tk.x] > 0 && tk.z >= kskip[tk.x] ? pB : pA).push_back(tk);
      b.oz.n_potrfA = (int)pA.size(); b.oz.n_potrfB = (int)pB.size();
      blob.add(&b.oz.d_potrfA, pA); blob.add(&b.oz.d_potrfB, pB);
      blob.add(&b.oz.d_kskip, kskip);
      b.oz.potrf = want_potrf && !pB.empty();
      {
        const char* te = getenv("DSMGP_OZAKI_TRSM");
        std::vector<int4> pA11;
        for (const int4& tk : pA) if (!(kskip[tk.x] > 0 && tk.y >= kskip[tk.x])) pA11.push_back(tk);
        std::stable_sort(tT.begin(), tT.end(), by_len);
        b.oz.n_potrfA11 = (int)pA11.size(); b.oz.n_jobsT = (int)jT.size(); b.oz.n_tilesT = (int)tT.size(); b.oz.nscaleT = scaleT;
        blob.add(&b.oz.d_potrfA11, pA11); blob.add(&b.oz.d_jobsT, jT); blob.add(&b.oz.d_tilesT, tT);
        b.oz.trsm = b.oz.potrf && !tT.empty() && !(te && te[0] == '0');
      }
      std::vector<int4> iA, iB;
      for (const int4& tk : b.h_trtri3m) {
        const std::vector<int>& ro = t.range_of[tk.x];
        if (ro[tk.y] != ro[tk.z]) continue;
        (kskip[tk.x] > 0 && tk.y >= kskip[tk.x] ? iB : iA).push_back(tk);
      }
      b.oz.n_invA = (int)iA.size(); b.oz.n_invB = (int)iB.size();
      blob.add(&b.oz.d_invA, iA); blob.add(&b.oz.d_invB, iB);
    }
    if (want_lauum && !b.h_lauum.empty()) {
      // operand X^T of every split expert: row blocks Jb, k blocks Kb >= Jb (upper block triangle; W_J^T on the diagonal)
      const int KB = OZ_KSTEPS_PER_BLK;
      std::vector<OzJob> jX; std::vector<OzTile> tW;
      std::vector<int64_t> pre(b.s1 - b.s0, -1), poolX(b.s1 - b.s0, 0);
      std::vector<int> scaleX(b.s1 - b.s0, 0);
      int64_t pool = 0, wblk = 0; int scale = 0;
      for (int s = b.s0; s < b.s1; s++) {
        const int sl = s - b.s0;
        const LeafMeta& m = h->meta[s];
        if (t.range_of[sl].back() == t.range_of[sl].front()) continue;         // unsplit: contracted by lauum3 itself
        const bool leaf_lauum = m.ktype == DSMGP_ISO_SE || m.ktype == DSMGP_ARD_LINEAR || (m.ktype == DSMGP_ARD_SE && !h->opts.as_written_grads);
        if (!leaf_lauum) continue;
        const int nb = m.nb;
        const double* F = h->d_F.p + m.foff; const double* WT = h->d_WT.p + m.woff;
        auto wid = [&](int blk) { const int w = m.np - blk * BLK; return w < BLK ? w : BLK; };
        poolX[sl] = pool; pool += (int64_t)nb * nb * KB * S;
        scaleX[sl] = scale; scale += nb * BLK;
        pre[sl] = wblk; wblk += (int64_t)nb * (nb + 1) / 2;
        for (int Jb = 0; Jb < nb; Jb++)
          for (int Kb = Jb; Kb < nb; Kb++)
            jX.push_back({Kb == Jb ? WT + (int64_t)Jb * WBLK_D : F + tile_off(Jb, Kb * (BLK / KC), m.nkc), 0, wid(Jb), wid(Kb), scaleX[sl] + Jb * BLK,
                          poolX[sl] + ((int64_t)Jb * nb + Kb) * KB * S});
      }
      const int64_t need_pool = pool; const int need_scale = scale;
      if (need_pool > 0 && need_pool <= max_pool && need_scale <= (int)max_scale) {
        for (const int4& tk : b.h_lauum) {
          if (pre[tk.x] < 0) continue;
          const LeafMeta& m = h->meta[b.s0 + tk.x];
          const int nb = m.nb, I = tk.y, J = tk.z;
          auto wid = [&](int blk) { const int w = m.np - blk * BLK; return w < BLK ? w : BLK; };
          const int kend = (nb - 1) * KB + (wid(nb - 1) + OZ_KSTEP - 1) / OZ_KSTEP;
          tW.push_back({(int)(poolX[tk.x] + (int64_t)I * nb * KB * S), (int)(poolX[tk.x] + (int64_t)J * nb * KB * S), I * KB, kend,
                        scaleX[tk.x] + I * BLK, scaleX[tk.x] + J * BLK, wid(I), wid(J), h->oz_scratch.p + (pre[tk.x] + tk.w) * (int64_t)WBLK_D, -1.0, 0, 0});
        }
        auto by_len = [](const OzTile& x, const OzTile& y) { return (x.k1 - x.k0) > (y.k1 - y.k0); };
        std::stable_sort(tW.begin(), tW.end(), by_len);
        for (const OzTile& x : tW) b.oz.ksteps_W += x.k1 - x.k0;
        b.oz.n_jobsX = (int)jX.size(); b.oz.n_tilesW = (int)tW.size(); b.oz.nscaleX = need_scale;
        blob.add(&b.oz.d_jobsX, jX); blob.add(&b.oz.d_tilesW, tW); blob.add(&b.oz.d_pre_base, pre);
        b.oz.lauum = !tW.empty();
      }
    }
    b.oz.n_parts = (int)parts.size();
    blob.add(&b.oz.d_parts, parts);
    CUDA_TRY(h, blob.commit(&b.oz.d_blob));
    b.oz.gemm_flops = flops;
  }
  return DSMGP_OK;
}

// Always-on segment timing (two events per segment, no synchronisation): which part of an evaluation ran where.
struct OzSegScope {
  dsmgp_handle* h; cudaStream_t st; int idx = -1;
  OzSegScope(dsmgp_handle* h_, int kind, cudaStream_t s) : h(h_), st(s) {
    if (h->capturing) return;
    auto ev = [&]() { if (h->oz_ev_used == h->oz_evs.size()) { cudaEvent_t e; cudaEventCreate(&e); h->oz_evs.push_back(e); } return h->oz_evs[h->oz_ev_used++]; };
    cudaEvent_t a = ev(), b = ev();
    cudaEventRecord(a, st);
    h->oz_segs.push_back({kind, a, b}); idx = (int)h->oz_segs.size() - 1;
  }
  ~OzSegScope() { if (idx >= 0) cudaEventRecord(h->oz_segs[idx].b, st); }
};

// DSMGP_OZAKI_TIMING=1: CUDA-event time of every launch of the split phases, printed to stderr (synchronises: diagnostics only)
struct OzTimer {
  bool on; cudaStream_t st; std::vector<cudaEvent_t> ev; std::vector<const char*> names;
  OzTimer(cudaStream_t s, bool capturing) : on(!capturing && getenv("DSMGP_OZAKI_TIMING") != nullptr), st(s) { mark("start"); }
  void mark(const char* name) { if (!on) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); names.push_back(name); }
  void report(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    fprintf(stderr, "[ozaki %s]", what);
    for (size_t i = 1; i < ev.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); fprintf(stderr, " %s %.3f", names[i], ms); }
    float tot = 0; cudaEventElapsedTime(&tot, ev.front(), ev.back()); fprintf(stderr, " | total %.3f ms\n", tot);
    for (auto e : ev) cudaEventDestroy(e);
  }
};

// optional clock stamps of one GEMM launch (DSMGP_OZAKI_TRACE / DSMGP_OZAKI_TRACE_SYRK = file): tools/ozaki_trace.py
static int oz_ctas(const dsmgp_handle* h) {      // DSMGP_OZAKI_CTAS: experiment (per-SM operand delivery against the number of active SMs)
  const char* e = getenv("DSMGP_OZAKI_CTAS");
  const int n = num_sms(h->device);
  return e ? std::max(1, std::min(n, atoi(e))) : n;
}

static void traced_gemm(dsmgp_handle* h, int S, const OzTile* d_tiles, int n, const char* file, cudaStream_t st) {
  const int nctas = oz_ctas(h);
  if (!file) { launch_oz_gemm(S, h->oz_map, d_tiles, n, h->oz_scale.p, nctas, st); return; }
  long long* d_trace = nullptr;
  cudaMalloc(&d_trace, (size_t)n * 128); cudaMemsetAsync(d_trace, 0, (size_t)n * 128, st);
  launch_oz_gemm(S, h->oz_map, d_tiles, n, h->oz_scale.p, nctas, st, d_trace);
  std::vector<long long> tr((size_t)n * 16);
  std::vector<OzTile> tl(n);
  cudaMemcpyAsync(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(tl.data(), d_tiles, tl.size() * sizeof(OzTile), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  if (FILE* f = fopen(file, "w")) {
    for (int i = 0; i < n; i++) {
      fprintf(f, "%d", tl[i].k1 - tl[i].k0);
      for (int q = 1; q < 9; q++) fprintf(f, " %lld", tr[(size_t)i * 16 + q] - tr[(size_t)i * 16]);
      fprintf(f, "\n");
    }
    fclose(f);
  }
  cudaFree(d_trace);
}

// The factorisation of one batch split at the root: block columns < mid, then A22 -= L21 L21^T as block products on the INT8
// tensor cores, then block columns >= mid whose contractions start at mid.  The tile flags persist across the two launches.
int32_t oz_run_potrf(dsmgp_handle* h, Batch& b, const Potrf2Args& full, const Trtri3Args* inv, int sms, cudaStream_t st) {
  const int S = h->oz_S;
  OzTimer tm(st, h->capturing);
  Potrf2Args pa = full;
  pa.tasks = b.oz.d_potrfA; pa.ntasks = b.oz.n_potrfA;
  Trtri3Args ta{};
  if (inv) { ta = *inv; ta.tasks = b.oz.d_invA; ta.ntasks = b.oz.n_invA; }
  const bool gemm_panel = inv && b.oz.trsm;
  if (gemm_panel) { pa.tasks = b.oz.d_potrfA11; pa.ntasks = b.oz.n_potrfA11; }
  {
    OzSegScope seg(h, 2, st);
    if (inv) launch_eval2_only(pa, ta, std::max(1, std::min(sms, pa.ntasks + ta.ntasks)), st);
    else launch_potrf2(pa, std::max(1, std::min(sms, pa.ntasks)), st);
  }
  tm.mark("potrfA");
  if (gemm_panel) {       // L21 = A21 X11^T in place of the panel tasks; their tile flags are set for the later launches
    {
      OzSegScope seg(h, 1, st);
      CUDA_TRY(h, cudaMemsetAsync(h->oz_rowmax.p, 0, (size_t)b.oz.nscaleT * sizeof(unsigned long long), st));
      launch_oz_slice(S, b.oz.d_jobsT, b.oz.n_jobsT, 0, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
      launch_oz_slice(S, b.oz.d_jobsT, b.oz.n_jobsT, 1, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
    }
    tm.mark("sliceA21X11");
    {
      OzSegScope seg(h, 0, st);
      launch_oz_gemm(S, h->oz_map, b.oz.d_tilesT, b.oz.n_tilesT, h->oz_scale.p, num_sms(h->device), st);
      h->oz_ksteps += b.oz.ksteps_T;
    }
    launch_oz_setflags(b.oz.d_parts, b.oz.n_parts, full.flag_off, full.flags, st);
    tm.mark("gemmL21");
    h->tm.launches += 4;
  }
  {
    OzSegScope seg(h, 1, st);
    CUDA_TRY(h, cudaMemsetAsync(h->oz_rowmax.p + b.oz.l21_scale0, 0, (size_t)b.oz.l21_nscale * sizeof(unsigned long long), st));
    launch_oz_slice(S, b.oz.d_jobsL, b.oz.n_jobsL, 0, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
    launch_oz_slice(S, b.oz.d_jobsL, b.oz.n_jobsL, 1, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
  }
  tm.mark("sliceL21");
  {
    OzSegScope seg(h, 0, st);
    traced_gemm(h, S, b.oz.d_syrk, b.oz.n_syrk, getenv("DSMGP_OZAKI_TRACE_SYRK"), st);
    h->oz_ksteps += b.oz.ksteps_syrk;
  }
  tm.mark("syrk");
  pa.tasks = b.oz.d_potrfB; pa.ntasks = b.oz.n_potrfB; pa.counter = full.counter + 1; pa.kskip = b.oz.d_kskip;
  {
    OzSegScope seg(h, 2, st);
    if (inv) {
      ta.tasks = b.oz.d_invB; ta.ntasks = b.oz.n_invB;
      launch_eval2_only(pa, ta, std::max(1, std::min(sms, pa.ntasks + ta.ntasks)), st);
      h->oz_inv_tiles_done = true;
    } else launch_potrf2(pa, std::max(1, std::min(sms, pa.ntasks)), st);
  }
  tm.mark("potrfB");
  tm.report("potrf");
  h->tm.launches += 5;
  h->oz_l21_ready = true;
  return DSMGP_OK;
}

// LAUUM of one batch: X^T of the split experts is sliced once, F^-1_IJ = sum_{K >= I} X_KI^T X_KJ of all their tiles are block
// products (stored negated in the scratch), and lauum3_kernel runs its dK-trace epilogue on them -- and the whole tile, contraction
// included, for the unsplit experts (pre_base < 0).
int32_t oz_run_lauum(dsmgp_handle* h, Batch& b, const LauumArgs& full, int sms, cudaStream_t st) {
  const int S = h->oz_S;
  OzTimer tm(st, h->capturing);
  {
    OzSegScope seg(h, 1, st);
    CUDA_TRY(h, cudaMemsetAsync(h->oz_rowmax.p, 0, (size_t)b.oz.nscaleX * sizeof(unsigned long long), st));
    launch_oz_slice(S, b.oz.d_jobsX, b.oz.n_jobsX, 0, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
    launch_oz_slice(S, b.oz.d_jobsX, b.oz.n_jobsX, 1, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
  }
  tm.mark("sliceXT");
  {
    OzSegScope seg(h, 0, st);
    launch_oz_gemm(S, h->oz_map, b.oz.d_tilesW, b.oz.n_tilesW, h->oz_scale.p, num_sms(h->device), st);
    h->oz_ksteps += b.oz.ksteps_W;
  }
  tm.mark("gemmW");
  LauumArgs la = full;
  la.pre = h->oz_scratch.p; la.pre_base = b.oz.d_pre_base;
  launch_lauum3(la, std::max(1, std::min(sms, la.ntasks)), st);
  tm.mark("lauum3 traces");
  tm.report("lauum");
  h->tm.launches += 4;
  return DSMGP_OK;
}

// The inverse of one batch: restricted tile pipeline, then per level  slice -> GEMM 1 -> slice -> GEMM 2,  then the
// partials of the GEMM-written tiles and the block-column reduction (alpha, tr(F^-1)).
int32_t oz_run_inverse(dsmgp_handle* h, Batch& b, const Trtri3Args& full, int sms, cudaStream_t st) {
  Trtri3Args ta = full;
  ta.tasks = b.oz.d_tasks; ta.ntasks = b.oz.n_tasks;
  OzTimer tm(st, h->capturing);
  if (!h->oz_inv_tiles_done) { OzSegScope seg(h, 2, st); launch_trtri3_only(ta, std::max(1, std::min(sms, ta.ntasks)), st); }
  tm.mark("trtri3");
  const int S = h->oz_S;
  for (int li = 0; li < b.oz.n_levels; li++) {
    const OzLevel& L = b.oz.levels[li];
    {
      OzSegScope seg(h, 1, st);
      CUDA_TRY(h, cudaMemsetAsync(h->oz_rowmax.p, 0, (size_t)L.n_scale * sizeof(unsigned long long), st));
      launch_oz_slice(S, L.d_jobs1, L.n_jobs1, 0, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
      launch_oz_slice(S, L.d_jobs1, L.n_jobs1, 1, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
      if (li == b.oz.n_levels - 1 && !h->oz_l21_ready) {      // root level: L21 not sliced by the factorisation phase
        CUDA_TRY(h, cudaMemsetAsync(h->oz_rowmax.p + b.oz.l21_scale0, 0, (size_t)b.oz.l21_nscale * sizeof(unsigned long long), st));
        launch_oz_slice(S, b.oz.d_jobsL, b.oz.n_jobsL, 0, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
        launch_oz_slice(S, b.oz.d_jobsL, b.oz.n_jobsL, 1, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
      }
    }
    tm.mark("slice1");
    {
      OzSegScope seg(h, 0, st);
      traced_gemm(h, S, L.d_tiles1, L.n_tiles1, (li == 0 && !h->capturing) ? getenv("DSMGP_OZAKI_TRACE") : nullptr, st);
      h->oz_ksteps += L.ksteps1;
    }
    tm.mark("gemm1");
    {
      OzSegScope seg(h, 1, st);
      launch_oz_slice(S, L.d_jobs2, L.n_jobs2, 0, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
      launch_oz_slice(S, L.d_jobs2, L.n_jobs2, 1, h->oz_rowmax.p, h->oz_scale.p, h->oz_pool.p, st);
    }
    tm.mark("slice2");
    {
      OzSegScope seg(h, 0, st);
      launch_oz_gemm(S, h->oz_map, L.d_tiles2, L.n_tiles2, h->oz_scale.p, num_sms(h->device), st);
      h->oz_ksteps += L.ksteps2;
    }
    tm.mark("gemm2");
    h->tm.launches += 6;
  }
  OzPartArgs pa{full.meta, full.F, full.z, full.flag_off, full.apart, full.tpart, b.oz.d_parts};
  launch_oz_parts(pa, b.oz.n_parts, st);
  launch_alpha_reduce(full, b.d_trtri_tasks, b.n_trtri, st);
  tm.mark("parts+reduce");
  tm.report("inverse");
  h->oz_x_complete = true;
  h->tm.launches += 3;
  return DSMGP_OK;
}

}  // namespace dsm

// out[0] batches on the split path, [1] slices S, [2] INT8 operations of the block products of the last evaluation
// (2 * 128 * 128 * 32 per tcgen05.mma, S (S + 1) / 2 of them per k-step), [3] the FP64 flops they stand for, [4] / [5] / [6] CUDA-event
// ms of the block-product launches / the slicing launches / the FP64 tile-pipeline launches of the last evaluation, [7] slice pool bytes,
// [8] (n >= 9) the factorisation + inverse flops that stay on the FP64 tile pipelines (2/3 r^3 per diagonal range of r rows)
extern "C" int32_t dsmgp_host_split_plan(int64_t n, int32_t depth, int32_t min_nb, int32_t* range_of, int32_t* n_ranges, double* share_int8) {
  if (n <= 0 || !range_of || !n_ranges) return DSMGP_ERR_ARG;
  if (depth <= 0) { const char* de = getenv("DSMGP_OZAKI_DEPTH"); depth = de ? std::max(1, std::min(4, atoi(de))) : 1; }
  if (min_nb <= 0) { const char* me = getenv("DSMGP_OZAKI_MIN_NB"); min_nb = me ? std::max(2, atoi(me)) : 8; }
  const int64_t np = (n + PAD - 1) / PAD * PAD;
  const int nb = (int)((np + BLK - 1) / BLK);
  std::vector<std::vector<Split>> levels(depth);
  std::vector<int> ro(nb, 0);
  int nr = 0;
  collect_splits(0, 0, nb, 0, depth, min_nb, levels, ro, nr);
  double rest = 0.0, rows = 0.0;
  for (int b = 0; b <= nb; b++) {
    if (b == nb || (b > 0 && ro[b] != ro[b - 1])) { rest += (rows / (double)n) * (rows / (double)n) * (rows / (double)n); rows = 0.0; }
    if (b < nb) { range_of[b] = ro[b]; rows += (double)std::max<int64_t>(0, std::min<int64_t>(BLK, n - (int64_t)b * BLK)); }
  }
  *n_ranges = nr;
  if (share_int8) *share_int8 = 1.0 - rest;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_int8_info(dsmgp_handle* h, double* out, int32_t n) {
  if (!h || !out || n < 8) return DSMGP_ERR_ARG;
  cudaSetDevice(h->device);
  int act = 0;
  for (const Batch& b : h->batches) act += b.oz.active ? 1 : 0;
  const int S = h->oz_S;
  double ms[3] = {0, 0, 0};
  cudaStreamSynchronize(h->stream);
  for (const auto& sg : h->oz_segs) { float t = 0; if (cudaEventElapsedTime(&t, sg.a, sg.b) == cudaSuccess) ms[sg.kind] += t; }
  out[0] = act; out[1] = S;
  out[2] = h->oz_ksteps * (S * (S + 1) / 2) * 2.0 * BLK * BLK * OZ_KSTEP;
  out[3] = h->oz_ksteps * 2.0 * BLK * BLK * OZ_KSTEP;
  out[4] = ms[0]; out[5] = ms[1]; out[6] = ms[2]; out[7] = (double)h->oz_pool_bytes;
  if (n >= 9) { out[8] = 0; for (const Batch& b : h->batches) if (b.oz.active) out[8] += b.oz.tile_flops; }
  return DSMGP_OK;
}
