/* libdsmgp.so -- C ABI of the B200-native (sm_100a) DSMGP per-expert GP hot path.
 *
 * The reference (trappmartin/DeepStructuredMixtures, Julia) has no FFI/plugin layer: its hot
 * path is ordinary Julia methods calling LAPACK/BLAS.  This header is therefore the boundary a
 * maintainer binds with `ccall`; every entry point names the reference
 * method(s) whose body it replaces (file:line under /root/reference/src).
 *
 * Rules
 *   - plain C, `extern "C"`, no C++/torch types; every pointer is a HOST pointer unless the
 *     name ends in `_dev`;  all matrices are column-major FP64 (Julia layout);  indices that
 *     come from Julia (`leaf_obs`) are Int64 1-based, everything else is 0-based.
 *   - the caller owns every buffer; the library copies inputs at `dsmgp_create` and keeps no
 *     host pointer.  All calls are synchronous on return.  A handle is not thread-safe.
 *   - every function returns a status (0 = ok); `dsmgp_last_error` gives the message.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     DSMGP_ERR_CUDA.  The pure host helpers (tree passes, sharding) are marked HOST-ONLY.
 *
 * Hyper-parameter layout (gaussianprocess.jl:153-161, optimize.jl:188-198): per kernel
 *   theta_k = [ logl (1 for Iso*, D for Ard*) , log sigma , logNoise ]   (nparams = len(logl)+2)
 * and for a kernel mixture `KernelFunction[...]` the concatenation over kernels.  The gradient
 * uses the same layout (gaussianprocess.jl:206-217).  Linear kernels ignore the log sigma slot
 * (kernels.jl:181-183,201) and report 0 gradient there.
 */
#ifndef DSMGP_H
#define DSMGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSMGP_VERSION 201

typedef struct dsmgp_handle dsmgp_handle;

enum { DSMGP_OK = 0, DSMGP_ERR_ARG = 1, DSMGP_ERR_CUDA = 2, DSMGP_ERR_COMM = 3, DSMGP_ERR_OOM = 4,
       DSMGP_ERR_NOT_PD = 5, DSMGP_ERR_STATE = 6 };

/* kernels.jl:59-64 (IsoSE), :109-114 (ArdSE), :174-177 (IsoLinear), :209-212 (ArdLinear) */
enum { DSMGP_ISO_SE = 0, DSMGP_ARD_SE = 1, DSMGP_ISO_LINEAR = 2, DSMGP_ARD_LINEAR = 3 };

typedef struct {
  int32_t type;      /* DSMGP_ISO_SE ... */
  int32_t nparams;   /* len(logl) + 2  (gaussianprocess.jl:139-145) */
} dsmgp_kernel_desc;

/* node types: GPNode (DeepStructuredMixtures.jl:61-71), GPSplitNode (:52-59), GPSumNode{T,SPNNode}
 * (:40-45) and the kernel-mixture GPSumNode{T,GPNode} built by _buildGP (treeStructure.jl:258-286) */
enum { DSMGP_NODE_LEAF = 0, DSMGP_NODE_SPLIT = 1, DSMGP_NODE_SUM = 2, DSMGP_NODE_KSUM = 3 };

/* Flattened region graph.  Nodes are numbered so that children precede parents. */
typedef struct {
  int64_t n_nodes;
  const int32_t* node_type;    /* [n_nodes] */
  const int64_t* child_ptr;    /* [n_nodes+1] CSR into child_idx, reference child order */
  const int64_t* child_idx;
  const int64_t* leaf_of_node; /* [n_nodes] leaf number (getLeaves / gpmap order, fit.jl:9-10) or -1 */
  const int32_t* split_dim;    /* [n_nodes] 0-based split dimension of a split node, else -1 */
  const int64_t* split_ptr;    /* [n_nodes+1] CSR into split_val */
  const double*  split_val;    /* thresholds s_1..s_K of GPSplitNode.split; the last one is upperBound[d] */
  int64_t root;
} dsmgp_tree;

enum { DSMGP_GRAD_AS_WRITTEN = 1, DSMGP_GRAD_MATHEMATICAL = 0 };

typedef struct {
  int32_t as_written_grads; /* 1 (default): gradients exactly as kernels.jl:85-99,146-164,196-200 compute them
                               (extra exp(log sigma) factor, ArdSE length-scale gradient == 0);  0: true dLML/dtheta */
  int32_t keep_factors;     /* 1: keep every leaf's Cholesky factor resident after fit (needed by predict and the
                               per-leaf accessors); 0: stream (factor -> reduce -> discard), for models whose
                               factors exceed HBM (cfg5) */
  int32_t rank, world;      /* leaf sharding for one-process-per-GPU runs; (0,1) = everything local */
  int32_t device;           /* CUDA device ordinal, -1 = current */
  int32_t strict_pd;        /* 1: a non-positive pivot makes fit/eval return DSMGP_ERR_NOT_PD */
  int64_t arena_bytes;      /* cap for the factor arena; 0 = choose from free HBM */
  int32_t reserved[8];
} dsmgp_opts;

void dsmgp_default_opts(dsmgp_opts* o);

/* ---- lifetime -------------------------------------------------------------------------------
 * Replaces the per-leaf state built by GaussianProcess(x, y; ...) gaussianprocess.jl:50-80 for every
 * leaf of buildTree (treeStructure.jl:245-307): x is the GLOBAL N x D input matrix, `leaf_obs` the
 * ascending 1-based rows of each leaf (GPNode.obs), `y_centered` the leaves' mean-subtracted targets
 * concatenated in leaf order (apply_subtract!, means.jl:11-14), `leaf_mean` the ConstMean value m.
 * The distance tensor P (gaussianprocess.jl:57) is never materialised. */
int32_t dsmgp_create(const double* x, int64_t N, int64_t D,
                     int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs,
                     const double* y_centered, const double* leaf_mean,
                     const int32_t* leaf_kernel_id,
                     const dsmgp_kernel_desc* kernels, int32_t n_kernels,
                     const dsmgp_tree* tree, const dsmgp_opts* opts,
                     dsmgp_handle** out);
void dsmgp_destroy(dsmgp_handle* h);
/* Device buffers >= 32 MiB of destroyed handles are kept in a process-wide cache (cudaMalloc / cudaFree of multi-GB
 * arenas cost 10 ms ... 3 s) and reused by later handles; this returns them to the driver. */
void dsmgp_release_cache(void);
const char* dsmgp_last_error(const dsmgp_handle* h); /* h may be NULL: last create error */

/* ---- parameters ----------------------------------------------------------------------------
 * setparams!(spn, hyp) optimize.jl:188-198 -> setparams!(gp, hyper) gaussianprocess.jl:153-161:
 * ONE global theta (sum of nparams over the kernels) broadcast to every leaf of the matching kernel id. */
int32_t dsmgp_set_params(dsmgp_handle* h, const double* theta, int64_t n);
/* per-leaf theta (finetune! ends with setparams!(gp.dist, hyp[gp.id]), finetuning.jl:75-84) */
int32_t dsmgp_set_leaf_params(dsmgp_handle* h, int64_t leaf, const double* theta, int64_t n);
int32_t dsmgp_get_leaf_params(const dsmgp_handle* h, int64_t leaf, double* theta, int64_t n);
int64_t dsmgp_nparams(const dsmgp_handle* h);        /* n in train!, optimisers.jl:15-17 */
int64_t dsmgp_n_leaves(const dsmgp_handle* h);
int64_t dsmgp_n_nodes(const dsmgp_handle* h);
int64_t dsmgp_leaf_size(const dsmgp_handle* h, int64_t leaf);

/* ---- fit -----------------------------------------------------------------------------------
 * fit!(spn, D, gpmap; tau) fit.jl:71-122 / fit_naive! :294-304 -> update_cholesky! gaussianprocess.jl:82-108
 * for every (local) leaf:  F = K + (exp(2 logNoise) + 1e-8) I ; L = potrf('L', F) ; alpha = L' \ (L \ y).
 * `overlap` = the L x L matrix D of getOverlap (fit.jl:12-39, column-major) and `tau` the minimal relative overlap of
 * fit!: the SHARED CHOLESKY.  The library repeats fit!'s scheduling (main expert = argmax D[:,j] .* D[j,:], :78-86) and
 * its case split (fitcontained!, :124-292):  identical experts are factored once (their factor, alpha, LML, gradients and
 * predictions are the source's);  an expert that shares its leading observations with its main expert (j inside main with
 * fewer than tau*n_j rows to delete, or main a leading part of j) takes the factor tiles of the common leading 128-row blocks
 * from the main expert and continues the factorisation behind them (chol_continue!).  Every result is the exact factor of
 * update_cholesky! (the parity target: fit.jl:105 always runs it; the reference's own row-deletion is numerically wrong,
 * SURVEY App. B Q7).  The plan is kept in the handle and used by every later fit / eval / grad until it is replaced;
 * it is suspended while experts hold different theta (dsmgp_set_leaf_params).  overlap == NULL: fit_naive! (this call
 * factors every expert on its own; a stored plan is kept).
 * `info[L]` (may be NULL): LAPACK potrf convention per leaf (0 ok, k>0 first non-positive pivot, 1-based;
 * chol_continue! returns the same, AdvancedCholeskey.jl:171-173).  `seconds` (may be NULL): device time of
 * the call, the value fit! returns (fit.jl:88,121). */
int32_t dsmgp_fit(dsmgp_handle* h, double tau, const double* overlap, int32_t* info, double* seconds);
/* Store (overlap != NULL) or clear (NULL) the sharing plan without fitting. */
int32_t dsmgp_set_sharing(dsmgp_handle* h, const double* overlap, double tau);
/* The stored plan per leaf (any pointer may be NULL): kind 0 = factored on its own, 1 = identical to `source` (fit.jl:132-143),
 * 2 = continues behind `blocks` leading 128-row blocks copied from `source` (fit.jl:145-292); source = leaf number or -1. */
int32_t dsmgp_get_sharing(const dsmgp_handle* h, int32_t* kind, int32_t* source, int32_t* blocks);

/* mll!(spn, L) optimize.jl:27-39 (leaf: mll(gp) gaussianprocess.jl:163): fills the per-node table
 * (AxisArray keyed by node id -> node_lml[n_nodes]); returns the root value in node_lml[root]. */
int32_t dsmgp_lml(dsmgp_handle* h, double* node_lml);

/* updategradients!(spn) fit.jl:306-311 (per leaf: gaussianprocess.jl:165-178, 219-226 and the kernel's
 * updategradients! kernels.jl:85-99,146-164,196-200) followed by the down-pass
 * nabla-mll!(spn, 0, 0, L, L[root], grad) optimize.jl:42-89.  `leaf_scale` = NULL, or the row D[g,:]
 * of the overlap matrix for the finetune variant optimize.jl:92-150.  grad[dsmgp_nparams] is OVERWRITTEN
 * (the reference zero-fills it first, optimisers.jl:76). */
int32_t dsmgp_grad(dsmgp_handle* h, const double* leaf_scale, double* grad);

/* One LML+gradient evaluation = optimisers.jl:43-77 minus the Flux step:
 * setparams! -> fit! -> mll! -> updategradients! -> nabla-mll!.  The benchmarked call.
 * theta may be NULL (keep current parameters); node_lml may be NULL. */
int32_t dsmgp_eval(dsmgp_handle* h, const double* theta, int64_t n, const double* leaf_scale,
                   double* lml, double* grad, double* node_lml);

/* finetune!'s inner loop (finetuning.jl:36-58) for G anchor experts in one call.  For every g: setparams!(spn, theta_g)
 * (ALL experts get theta_g), fit!, mll!, updategradients!, and the finetune down-pass optimize.jl:92-150 with the
 * weights D[anchor_g, :] (`overlap` = the L x L matrix of getOverlap, fit.jl:12-39, column-major).  Outputs per g:
 * leaf_lml[g] = L[gp.id] (finetuning.jl:51), grads[g*H .. ] (finetuning.jl:53), root_lml[g] (may be NULL).
 * The G evaluations are independent (theta_g is only updated from its own gradient), so they run back to back on the
 * device without host synchronisation; experts whose weight D[anchor_g, l] is 0 skip the inverse / LAUUM kernels
 * (dsmgp_eval does the same whenever `leaf_scale` has zeros).  On return the handle holds theta of the last anchor. */
int32_t dsmgp_finetune_eval(dsmgp_handle* h, int64_t G, const int64_t* anchors, const double* thetas /* G x H row-major */,
                            const double* overlap /* L x L col-major */, double* leaf_lml /* G */,
                            double* grads /* G x H row-major */, double* root_lml /* G or NULL */);

/* train!(spn, D, gpmap, optim; iterations, lambda, earlystop) optimisers.jl:40-83 in one call: every
 * iteration is a dsmgp_eval followed by the Flux.Optimise step (optimiser 0 Descent(eta), 1 ADAM(eta, (beta1, beta2)),
 * 2 RMSProp(eta, rho = beta1); epsilon 1e-8) and `hyp += grad` (gradient ASCENT).  state_by_identity = 1 reproduces the
 * reference, whose rebinding of `hyp` hands apply! a fresh optimiser state every iteration.
 * theta[H]: start values in, final values out; ell[iterations]: LML trace; *n_done: iterations executed (early stopping:
 * |ell_t - mean(ell_{t-9..t-1})| < lambda for `earlystop` consecutive iterations returns BEFORE the update, as :63-66).
 * When the loop runs to the end the handle is left fitted at the final theta (:82-83). */
int32_t dsmgp_train(dsmgp_handle* h, int32_t optimiser, double eta, double beta1, double beta2, int32_t state_by_identity,
                    int64_t iterations, double lambda, int64_t earlystop, double* theta, double* ell, int64_t* n_done);

/* Per-leaf rows of the last eval: rows[l*(1+Hmax) + 0] = mll(gp_l), [1..] = nabla-mll(gp_l)
 * (gaussianprocess.jl:185-217).  Hmax = max nparams over kernels.  Also the multi-GPU exchange unit:
 * a rank fills only its own leaves, rows of other ranks are 0, so a SUM all-reduce assembles them. */
int64_t dsmgp_row_width(const dsmgp_handle* h);
int32_t dsmgp_eval_local_dev(dsmgp_handle* h, const double* theta, int64_t n, double** rows_dev);
int32_t dsmgp_eval_finish_dev(dsmgp_handle* h, const double* leaf_scale, double* lml, double* grad, double* node_lml);
int32_t dsmgp_leaf_rows(const dsmgp_handle* h, double* rows);
int32_t dsmgp_leaf_owner(const dsmgp_handle* h, int32_t* owner); /* rank owning each leaf (LPT on n^3) */

/* ---- multi-GPU inside the library -------------------------------------------------------------
 * One process (or thread) per GPU, handles created with rank/world in dsmgp_opts.  The only exchange of an evaluation is the
 * table of per-leaf rows; with a communicator attached, dsmgp_fit / dsmgp_eval / dsmgp_grad / dsmgp_lml / dsmgp_update_weights
 * and dsmgp_predict work at world > 1 exactly like at world == 1: the library all-reduces (NCCL SUM, on its own stream) and every
 * rank finishes the O(L) tree passes.  NCCL is loaded with dlopen("libnccl.so.2") at the first of these two calls; the library
 * has no link-time dependency on it.  `id` is NCCL's 128-byte ncclUniqueId: rank 0 creates it, the caller distributes it to the
 * other ranks by any means (MPI.jl / sockets / a file), every rank passes it to dsmgp_comm_init (collective call). */
#define DSMGP_COMM_ID_BYTES 128
int32_t dsmgp_comm_unique_id(void* id /* DSMGP_COMM_ID_BYTES */);
int32_t dsmgp_comm_init(dsmgp_handle* h, const void* id);

/* ---- posterior weights and prediction --------------------------------------------------------
 * update!(spn) common.jl:323-334: sum_logweights receives, for every sum node in node order, its
 * normalised child log-weights (CSR by child_ptr; pass NULL to skip), *z the root evidence. */
int32_t dsmgp_update_weights(dsmgp_handle* h, double* sum_logweights, double* z);
/* infer!(spn) common.jl:336-355: like update!, but only the kernel-mixture sum nodes (GPSumNode{T,GPNode}, :339-345) keep their
 * posterior weights; every sum node over sub-trees is reset to the uniform -log K after its evidence is computed (:347-353).
 * The weights are stored in the handle (used by dsmgp_predict) and returned like dsmgp_update_weights. */
int32_t dsmgp_infer(dsmgp_handle* h, double* sum_logweights, double* z);
/* reset_weights!(spn) common.jl:357-363: every sum node's log-weights = -log K. */
int32_t dsmgp_reset_weights(dsmgp_handle* h, double* sum_logweights);

enum { DSMGP_PREDICT_DSMGP = 0, DSMGP_PREDICT_POE = 1, DSMGP_PREDICT_GPOE = 2, DSMGP_PREDICT_RBCM = 3 };
/* predict(model, x) common.jl:294-307 -> prediction(gp, xtest) gaussianprocess.jl:110-137.
 * xtest is T x D column-major.  Requires keep_factors and a preceding fit/eval (+ update_weights for
 * the DSMGP mode, like the reference). */
int32_t dsmgp_predict(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, double* mu, double* var);

/* Leaf-sharded prediction (rank/world in dsmgp_opts, one process per GPU).  dsmgp_predict_local routes the test points,
 * predicts the LOCAL experts and writes mu into buf[0 .. total) and var into buf[total .. 2 total) in (leaf, routing order)
 * layout, 0 for experts of other ranks (buf == NULL: only *total is returned).  After a SUM all-reduce of buf over the
 * ranks, dsmgp_predict_finish mixes (common.jl:134-307) on every rank.  Needs a preceding fit/eval (+ update_weights). */
int32_t dsmgp_predict_local(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, double* buf, int64_t* total);
int32_t dsmgp_predict_finish(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, const double* buf,
                             double* mu, double* var);

/* prediction(gp, xtest) of ONE leaf: mu[T] and diag(Sigma)[T] (gaussianprocess.jl:110-137) */
int32_t dsmgp_leaf_predict(dsmgp_handle* h, int64_t leaf, const double* xtest, int64_t T, double* mu, double* var);

/* ---- per-leaf accessors (gp.alpha, gp.cK.factors stay readable from Julia) -------------------- */
int32_t dsmgp_leaf_alpha(const dsmgp_handle* h, int64_t leaf, double* alpha /* n */);
int32_t dsmgp_leaf_factor(const dsmgp_handle* h, int64_t leaf, double* Lfac /* n x n col-major, lower; upper = 0 */);
int32_t dsmgp_leaf_info(const dsmgp_handle* h, int32_t* info /* L */);

/* ---- stand-alone operators ------------------------------------------------------------------
 * kernelmatrix(kernel, x1, x2) kernels.jl:15-18.  theta = [logl..., log sigma] (noise not used).
 * x1: n1 x D, x2: n2 x D, K: n1 x n2, all column-major. */
int32_t dsmgp_kernelmatrix(int32_t kernel_type, const double* theta, int64_t D,
                           const double* x1, int64_t n1, const double* x2, int64_t n2, double* K);
/* getOverlap(spn, D, gpmap) fit.jl:12-39 (called once by build, treeStructure.jl:428-431):  D[n,m] = 1 - card(obs_n \ obs_m)
 * / card(obs_n) for experts whose lowest common ancestor is a sum node (1 when their kernel ids differ), else 0.
 * D is L x L column-major (D[n,m] at n + m*L), bit-identical to the reference's formula; computed on the device by
 * letting every point enumerate the pairs of experts that contain it. */
int32_t dsmgp_overlap(int64_t N, int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs,
                      const int32_t* leaf_kernel_id, const dsmgp_tree* tree, double* D);
/* The same matrix in CSR form by row (row n: the experts m with D[n,m] != 0, ascending), for models whose dense L x L matrix is
 * too large to move (3.4 GB at L = 20,736): only the non-zeros leave the device.  Call once with col = val = NULL to obtain
 * row_ptr[L+1] (row_ptr[L] = number of non-zeros), then again with col[nnz], val[nnz]. */
int32_t dsmgp_overlap_csr(int64_t N, int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs,
                          const int32_t* leaf_kernel_id, const dsmgp_tree* tree, int64_t* row_ptr, int32_t* col, double* val);
/* AdvancedCholesky.chol_continue!(A, ki) AdvancedCholeskey.jl:152-174 (ki 1-based): A n x n column-major in/out. */
int32_t dsmgp_chol_continue(double* A, int64_t n, int64_t ki, int32_t* info);
/* Row/column deletion from a lower Cholesky factor: the operation fit.jl:179-195 composes from
 * lowrankupdate! (AdvancedCholeskey.jl:20-59), implemented CORRECTLY.
 * A: n x n in, (n-nrows) x (n-nrows) factor written to `out` (column-major, ld = n-nrows). rows 1-based ascending. */
int32_t dsmgp_chol_delete_rows(const double* A, int64_t n, const int64_t* rows, int64_t nrows, double* out);
/* The same for `count` factors at once: one CTA per matrix, ONE column sweep per matrix for all of its deleted rows (bit-identical
 * to deleting them one after the other), one launch per 64 deleted rows.  A[m]: n[m] x n[m]; out[m]: (n[m]-nrows[m])^2. */
int32_t dsmgp_chol_delete_rows_batched(int64_t count, const double* const* A, const int64_t* n, const int64_t* const* rows,
                                       const int64_t* nrows, double* const* out);
/* plain batched-size-1 potrf('L') on a host matrix (LAPACK.potrf! as used at gaussianprocess.jl:101) */
int32_t dsmgp_potrf(double* A, int64_t n, int32_t* info);

/* ---- region-graph construction on the device ---------------------------------
 * The data passes of buildTree (treeStructure.jl:23-243) on index lists in device memory; the recursion and every random draw
 * (Beta(2,2), rand(1:2), Categorical) stay with the caller, so the partitions are bit-identical to the host builder's.
 * A partition owns a device copy of x (N x D column-major) and a list of nodes; node 0 holds all rows. */
typedef struct dsmgp_partition dsmgp_partition;
int32_t dsmgp_part_create(const double* x, int64_t N, int64_t D, dsmgp_partition** out);
void    dsmgp_part_destroy(dsmgp_partition* p);
int64_t dsmgp_part_size(const dsmgp_partition* p, int64_t node);
/* X.max - X.min per dimension over the node's rows (_buildSum, treeStructure.jl:233-235): mins[D], maxs[D] */
int32_t dsmgp_part_range(dsmgp_partition* p, int64_t node, double* mins, double* maxs);
/* column d of the node's rows in ascending order (n values): min / max / median of a sub-range / sum(. <= s) of getSplits
 * (treeStructure.jl:36-57) are binary searches on it */
int32_t dsmgp_part_sorted_column(dsmgp_partition* p, int64_t node, int64_t d, double* sorted);
/* the children of _buildSplit (treeStructure.jl:176-199): child k = rows with lower[k] < x_d <= upper[k] in the node's own
 * (ascending) order; K <= 32 new node numbers in children[], their sizes in sizes[] */
int32_t dsmgp_part_split(dsmgp_partition* p, int64_t node, int64_t d, const double* lower, const double* upper, int64_t K,
                         int64_t* children, int64_t* sizes);
/* GPNode.obs of a leaf: the node's rows, 1-based ascending (n values) */
int32_t dsmgp_part_rows(dsmgp_partition* p, int64_t node, int64_t* obs);

/* ---- HOST-ONLY helpers (no GPU needed; used by the multi-process plumbing and its CPU tests) ---
 * Tree passes over a table of per-leaf rows (layout of dsmgp_leaf_rows): up-pass mll! optimize.jl:27-39,
 * down-pass nabla-mll! :42-150, update! common.jl:323-334. */
int32_t dsmgp_host_tree_eval(const dsmgp_tree* tree, int64_t L, const int32_t* leaf_kernel_id,
                             const dsmgp_kernel_desc* kernels, int32_t n_kernels,
                             const double* rows, int64_t row_width, const double* leaf_scale,
                             double* node_lml, double* grad, double* sum_logweights, double* z);
/* The sharing plan dsmgp_set_sharing / dsmgp_fit(tau, overlap) derives from a structure (fit.jl:71-122): kind / source / blocks
 * per leaf as in dsmgp_get_sharing. */
int32_t dsmgp_host_sharing_plan(int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs, const int32_t* leaf_kernel_id,
                                const double* overlap, double tau, int32_t* kind, int32_t* source, int32_t* blocks);
/* LPT bin packing of leaves by n^3 onto `world` ranks (deterministic). */
int32_t dsmgp_host_shard(int64_t L, const int64_t* leaf_ptr, int32_t world, int32_t* owner);
/* The diagonal ranges the INT8 split path (csrc/api_ozaki.cu) gives an expert of n observations: range_of[b] for its
 * ceil(n_padded / 128) block rows (n_padded = n rounded up to 64; at least that many entries), ranges numbered from 0 in
 * row order; *share_int8 = the fraction of the factorisation + inverse flops (2/3 n^3) that runs as INT8 block products
 * (1 - sum over the ranges of (rows / n)^3).  depth / min_nb <= 0: the library defaults (environment included). */
int32_t dsmgp_host_split_plan(int64_t n, int32_t depth, int32_t min_nb, int32_t* range_of, int32_t* n_ranges, double* share_int8);

/* ---- instrumentation ------------------------------------------------------------------------ */
typedef struct {
  double gram_ms, potrf_ms, solve_ms, inverse_ms, grad_ms, tree_ms, total_ms;
  double potrf_flops, inverse_flops, gram_bytes;   /* algorithmic work of the last eval (local leaves): Cholesky (potrf_ms), triangular
                                                      inverse (inverse_ms; the LAUUM pass inside grad_ms is the same count again), Gram */
  int64_t launches;                                 /* kernels launched by the last call */
  double predict_ms, predict_flops, predict_bytes;  /* last dsmgp_predict / dsmgp_leaf_predict: device time of predict_kernel,
                                                       sum_l (n_l^2 T_l + 2 n_l T_l) flop, bytes of L + x + xt read */
} dsmgp_timings;
int32_t dsmgp_get_timings(const dsmgp_handle* h, dsmgp_timings* t);
int32_t dsmgp_set_profiling(dsmgp_handle* h, int32_t on); /* per-phase CUDA events (adds syncs) */

/* The INT8 split path of the last evaluation (csrc/api_ozaki.cu; no reference counterpart -- it replaces the BLAS-3 part of
 * LAPACK potrf!/trtri that gaussianprocess.jl:99-101,219-226 call): experts of >= 1024 observations are split at the middle
 * block row; the products L21 = A21 X11^T, A22 -= L21 L21^T, T = L21 X11, X21 = -X22 T run as error-free INT8 slice products
 * (Ozaki scheme, 8 slices of 7 bits, exact int32 accumulation in TMEM) on the tcgen05 tensor cores, the diagonal ranges stay
 * on the FP64 tile pipelines.  DSMGP_OZAKI=0 in the environment turns the path off.
 * out[0] batches on the split path, [1] slices, [2] INT8 operations of the block products, [3] the FP64 flops they stand for,
 * [4] / [5] / [6] CUDA-event ms of the block-product / slicing / FP64 tile-pipeline launches, [7] slice pool bytes, [8] (when
 * n >= 9) the factorisation + inverse flops left on the FP64 tile pipelines.  n >= 8. */
int32_t dsmgp_int8_info(dsmgp_handle* h, double* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* DSMGP_H */
