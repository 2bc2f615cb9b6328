/*
 * hkmeans.h — C ABI of libhkmeans.so, the B200-native (sm_100a) replacement for the
 * k-means Lloyd hot path of helmholtz-analytics/heat.
 *
 * Every entry point replaces one reference interface (paths relative to /root/reference):
 *
 *   hk_lloyd_step / hk_lloyd_accumulate + hk_lloyd_finalize
 *        <- KMeans.fit loop body            heat/cluster/kmeans.py:131-144
 *           = _assign_to_cluster            heat/cluster/_kcluster.py:352-370
 *           + KMeans._update_centroids      heat/cluster/kmeans.py:76-103
 *           + shift^2 / tol test            heat/cluster/kmeans.py:141-144
 *   hk_assign
 *        <- _KCluster.predict / _assign_to_cluster(eval_functional_value=True)
 *                                           heat/cluster/_kcluster.py:352-370, 398-415
 *   hk_cdist
 *        <- cdist -> _dist -> _euclidian_fast / _euclidian (X split 0|None, Y replicated)
 *                                           heat/spatial/distance.py:32-64, 136-156, 409-414
 *   hk_pairwise
 *        <- rbf / manhattan (and cdist) -> _dist -> _gaussian(_fast) / _manhattan(_fast)
 *                                           heat/spatial/distance.py:67-133, 159-207
 *   hk_comm_* / hk_allreduce_f64 (and the peer-memory exchange inside hk_lloyd_step / hk_lloyd_run)
 *        <- MPICommunication.Allreduce(MPI.IN_PLACE, t, MPI.SUM) as issued by __reduce_op
 *                                           heat/core/communication.py:1089-1110,
 *                                           heat/core/_operations.py:505-510
 *   hk_lloyd_run
 *        <- the `for epoch in range(max_iter)` loop of KMeans.fit between two convergence checks
 *                                           heat/cluster/kmeans.py:131-144
 *   hk_chunk
 *        <- MPICommunication.chunk          heat/core/communication.py:197-254
 *
 * Conventions
 *   - plain pointers and sizes only; all array pointers are DEVICE pointers owned by the caller
 *     (torch tensors held by DNDarrays).  The library never frees or retains them.
 *   - every call is asynchronous on the caller-supplied cudaStream_t (passed as void*),
 *     except hk_create/hk_destroy/hk_comm_init/hk_comm_destroy.
 *   - every export returns int: 0 ok, <0 bad argument / unsupported, >0 = 1000 + cudaError_t
 *     or 2000 + ncclResult_t.  Nothing throws or aborts.  hk_last_error() gives the text
 *     (thread-local).
 *   - dtype: HK_F32 = 0, HK_F64 = 1 (arithmetic type of X and of the centroids handed in).
 *   - partial buffers ("partials") are k*(d+1) doubles: row c = [sum_0 .. sum_{d-1}, count].
 */
#ifndef HKMEANS_H_
#define HKMEANS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HK_F32 0
#define HK_F64 1

#define HK_LABEL_NONE 0
#define HK_LABEL_U8 1
#define HK_LABEL_I32 2
#define HK_LABEL_I64 3

/* distance path selection for hk_lloyd_* / hk_assign (HK_PATH_AUTO picks by shape) */
#define HK_PATH_AUTO 0
#define HK_PATH_SIMT 1   /* exact-formula distances without the TF32 filter: FP64 tensor-core kernel for fp64
                            d=16 k<=16, warp-specialised FMA kernel for other 128-byte rows, generic tiled
                            kernel otherwise                                                  */
#define HK_PATH_TC 2     /* tcgen05 TF32 filter + exact FMA refinement (fp32 only)          */
#define HK_PATH_GENERIC 3 /* the generic tiled exact-FMA kernel, whatever the shape          */
#define HK_PATH_ROW128 4  /* the warp-specialised exact-FMA kernel for 128-byte rows (tests)  */

typedef struct hk_handle_s* hk_handle_t;

int hk_version(void);
const char* hk_last_error(void);

/* workspace + (optional) communicator; one per (process, device) */
int hk_create(hk_handle_t* out, int device);
int hk_destroy(hk_handle_t h);

/* replaces MPICommunication.chunk for split=0: rows and offset of `rank` out of `nranks` */
int hk_chunk(int64_t n_global, int nranks, int rank, int64_t* offset, int64_t* rows);

/* ---- one Lloyd pass over this rank's row shard ------------------------------------------
 * X        : n_local x d, row stride ldx elements
 * C        : k x d centroids (same dtype as X), replicated
 * labels   : optional per-row labels against C (label_kind selects the element type)
 * partials : k*(d+1) doubles, this rank's per-cluster sums and counts (overwritten)
 */
int hk_lloyd_accumulate(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype,
                        const void* C, int k, void* labels, int label_kind, double* partials,
                        void* row_ws, int64_t row_ws_bytes, int path, void* stream);

/* partials (already summed over ranks) -> new centroids, shift^2, convergence flag.
 *   C_out[c] = cast( sums[c] / double(float(max(count[c],1))) )      (quirks Q1-Q3)
 *   shift2   = sum (C_in - C_out)^2 evaluated in the centroid dtype  (kmeans.py:141)
 *   state    : int32[4] device scratch owned by the caller:
 *              [0] converged flag (sticky), [1] iterations executed, [2..3] reserved.
 *              When state[0] is already 1 the call is a no-op (lets the host enqueue
 *              iterations ahead without a sync per iteration and still get n_iter_ exact).
 *   use_tol  : 0 -> never converge (tol=None)
 *   tol_cmp  : float32(tol) as the reference compares it
 */
int hk_lloyd_finalize(hk_handle_t h, const double* partials, const void* C_in, void* C_out, int k,
                      int d, int dtype, int use_tol, double tol_cmp, void* shift2_out,
                      int32_t* state, void* stream);

/* fused: accumulate -> (sum over the ranks of the handle's communicator, if `allreduce`) -> finalize, as TWO kernel
 * launches: the pass over X and one "finish" kernel that reduces the pass's per-CTA slots, exchanges the k x (d+1)
 * partials with the other ranks through peer-mapped GPU memory (see hk_comm_peer_*; falls back to ncclAllReduce +
 * separate kernels when the peers are not mapped) and runs the finalize arithmetic.
 * C is updated IN PLACE (C_prev receives the pre-update centroids when not NULL). */
int hk_lloyd_step(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype,
                  void* C, void* C_prev, int k, void* labels, int label_kind, int use_tol,
                  double tol_cmp, void* shift2_out, int32_t* state, int allreduce, void* row_ws,
                  int64_t row_ws_bytes, int path, void* stream);

/* `iters` consecutive hk_lloyd_step calls (no labels) enqueued by one call; all but the first are replayed from a CUDA
 * graph cached in the handle while the arguments stay the same.  With `state`, steps after convergence are no-ops, so
 * the host may enqueue a chunk of iterations, read state once, and still obtain the exact n_iter_. */
int hk_lloyd_run(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, void* C,
                 void* C_prev, int k, int use_tol, double tol_cmp, void* shift2_out, int32_t* state,
                 int allreduce, void* row_ws, int64_t row_ws_bytes, int path, int iters, void* stream);

/* Optional per-matrix workspace (`row_ws`, `row_ws_bytes` of hk_lloyd_* / hk_assign): the tensor-core path keeps a
 * per-tile upper bound of |x| in it (like sklearn's x_squared_norms; it enters only the error bound that decides which
 * rows need exact re-evaluation, never a result).  The buffer is OWNED BY THE CALLER and tied to the CONTENT of X:
 * allocate hk_row_ws_bytes(n_local) bytes of device memory, ZERO them (cudaMemsetAsync) whenever the rows of X change,
 * and pass the same buffer with every pass over that X; the first pass fills it, later passes read it.  NULL is
 * always valid: the bounds are then recomputed from the rows in every pass.  The library keeps no per-matrix state. */
int64_t hk_row_ws_bytes(int64_t n_local);

/* labels (+ optional sum over rows of min_j d^2, one double) — predict */
int hk_assign(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype,
              const void* C, int k, void* labels, int label_kind, double* min_d2_sum, void* row_ws,
              int64_t row_ws_bytes, int path, void* stream);

/* out[i,j] = dist(X[i], Y[j]); quadratic_expansion != 0 -> sqrt(clamp(|x|^2+|y|^2-2xy,0)),
 * else direct sqrt(sum (x-y)^2).  sqrt_flag = 0 returns squared distances. */
int hk_cdist(hk_handle_t h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
             int64_t ldy, void* out, int64_t ldo, int dtype, int quadratic_expansion,
             int sqrt_flag, void* stream);

/* Other metrics of heat/spatial/distance.py on the same local blocks (what _dist calls per tile, distance.py:409-414
 * and inside the rings :262-359, :431-473):
 *   HK_METRIC_EUCLIDEAN  cdist      sqrt(sum (x-y)^2)                  expand = quadratic_expansion   distance.py:17-64
 *   HK_METRIC_GAUSSIAN   rbf        exp(-|x-y|^2 / (2 sigma^2))        expand = quadratic_expansion   distance.py:67-101
 *   HK_METRIC_MANHATTAN  manhattan  sum |x-y|                          expand ignored (same values)   distance.py:104-133 */
#define HK_METRIC_EUCLIDEAN 0
#define HK_METRIC_GAUSSIAN 1
#define HK_METRIC_MANHATTAN 2
int hk_pairwise(hk_handle_t h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                int64_t ldy, void* out, int64_t ldo, int dtype, int metric, int expand, double sigma,
                void* stream);

/* ---- other consumers of the assignment pattern (KMedians / KMedoids / KNeighborsClassifier) ----------------------- */

/* labels[i] = first-index argmin_j sum_f |x_if - c_jf| (+ optional sum_i of that minimum, one double)
 *   <- _KCluster._assign_to_cluster with metric = manhattan(x, y, expand=True), p = 1
 *      heat/cluster/kmedians.py:43-50, heat/cluster/kmedoids.py:45-52, heat/cluster/_kcluster.py:352-370 */
int hk_assign_l1(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* C, int k,
                 void* labels, int label_kind, double* min_d_sum, void* stream);

/* Per-cluster, per-feature medians of a row shard by radix selection; only k*d*2*256 counts travel between ranks.
 *   <- KMedians / KMedoids._update_centroids: ht.median over the rows of every cluster, all-zero rows dropped first
 *      heat/cluster/kmedians.py:60-103, heat/cluster/kmedoids.py:57-93, heat/core/statistics.py:1684-1728
 * Protocol (every array on the device; w = 0 lower middle, 1 upper middle):
 *   hk_row_keep    keep[i] = row i is not entirely zero
 *   for pass in 0 .. hk_select_passes(dtype)-1:
 *       zero hist[2][k][d][256] (int64); hk_select_hist adds this shard's counts (pass 0 fills hist[0] only: copy it
 *       to hist[1]); sum hist over the ranks;
 *       (after pass 0 the caller derives the cluster sizes from hist and sets remaining[2][k][d] to the wanted ranks)
 *       hk_select_step   picks the digit holding rank remaining[w][j][f], appends it to prefix[w][j][f] (uint64)
 *   hk_select_value   medians[j][f] = lo + (hi - lo) * frac[j]   (frac = 0.5 for even cluster sizes, else 0) */
int hk_row_keep(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, void* keep_u8,
                void* stream);
int hk_select_passes(int dtype);
int hk_select_hist(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* labels_i64,
                   const void* keep_u8, int k, const void* prefix_u64, int pass, void* hist_i64, void* stream);
int hk_select_step(hk_handle_t h, const void* hist_i64, void* remaining_i64, void* prefix_u64, int k, int d,
                   void* stream);
int hk_select_value(hk_handle_t h, const void* prefix_u64, const double* frac, int k, int d, int dtype, void* medians,
                    void* stream);

/* Centre update of the batch-parallel clusterers' local loop: clusters with rows take their mean (partials = the
 * k x (d+1) fp64 sums | counts of hk_lloyd_accumulate) or their median (medians [k][d] + counts [k], partials NULL),
 * empty clusters keep their centre; flag[0] (int32) = all |new - old| <= atol + rtol |old|
 *   <- _kmex, heat/cluster/batchparallelclustering.py:56-84 (torch.allclose(centers, centers_old, atol=tol)) */
int hk_kmex_update(hk_handle_t h, const double* partials, const void* medians, const void* counts_i64, void* C, int k,
                   int d, int dtype, double atol, double rtol, void* flag_i32, void* stream);

/* best_index[j] = global index (row_base + i) of the shard row closest in L1 to P[j], first index on ties
 *   <- KMedoids._update_centroids: dist = manhattan(x, median); idx = dist.argmin(axis=0)
 *      heat/cluster/kmedoids.py:94-110 */
int hk_nearest_rows_l1(hk_handle_t h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* P,
                       int k, int64_t row_base, double* best_dist, void* best_index_i64, void* stream);

/* The kk smallest entries of every row of D (ascending, lower index first on ties) and the class vote over them
 *   <- KNeighborsClassifier.predict: ht.topk(distances, n_neighbors, largest=False); y[indices] summed; argmax
 *      heat/classification/kneighborsclassifier.py:124-135 */
int hk_topk_rows(hk_handle_t h, const void* D, int64_t m, int64_t n, int64_t ldd, int dtype, int kk, void* values,
                 void* indices_i64, void* stream);
int hk_knn_vote(hk_handle_t h, const void* indices_i64, int64_t m, int kk, const void* Y, int64_t n, int n_classes,
                int64_t ldy, int dtype, void* classes_i64, void* stream);

/* ---- communicator (NCCL, resolved with dlopen at first use) ----------------------------- */
int hk_comm_unique_id(void* id128);                      /* 128-byte ncclUniqueId            */
int hk_comm_init(hk_handle_t h, int nranks, int rank, const void* id128);
int hk_comm_destroy(hk_handle_t h);
/* Peer-memory mailbox of the fused finish kernel (ranks = processes of ONE box, one GPU each).  After hk_comm_init:
 * every rank calls hk_comm_peer_export (allocates its mailbox for partial vectors of up to cap_doubles values and
 * writes a 64-byte cudaIpcMemHandle_t), the host gathers the handles in rank order (torch.distributed all_gather),
 * every rank calls hk_comm_peer_import with the nranks x 64 bytes, then a host barrier.  Until then (or if the import
 * fails) hk_lloyd_step uses ncclAllReduce.  hk_comm_mode: 0 single rank, 1 NCCL, 2 peer memory. */
int hk_comm_peer_export(hk_handle_t h, int64_t cap_doubles, void* handle64);
int hk_comm_peer_import(hk_handle_t h, const void* handles);
int hk_comm_mode(hk_handle_t h);
int hk_allreduce_f64(hk_handle_t h, double* buf, int64_t count, void* stream);

/* ---- introspection for tests/bench ------------------------------------------------------- */
/* CUDA-graph replay inside hk_lloyd_run: on by default; hk_graph_enable(h, 0) makes it enqueue plain launches */
int hk_graph_enable(hk_handle_t h, int enable);
int64_t hk_graph_launch_count(hk_handle_t h);
/* kernels launched by this handle since creation (graph replays count their kernel nodes) */
int64_t hk_launch_count(hk_handle_t h);
/* name of the kernel variant the last hk_lloyd_accumulate / hk_assign call selected */
const char* hk_last_variant(hk_handle_t h);
/* CUDA-event timing of the dominant kernel (the Lloyd pass / the cdist kernel) on its own stream:
 * enable != 0 brackets every such launch with an event pair; hk_profile_read synchronises, returns the
 * summed device time in ms and the number of launches measured, and clears the list. */
/* cold-path counters of the tensor-core passes of this handle since the previous call (synchronises the device, then
 * clears them): out6 = rows the TF32 filter could not decide, exact (row, centroid) evaluations, rows sent through the
 * all-centroid formula (NaN/Inf), warps that entered the cold path, rows processed, passes. */
int hk_stats_read(hk_handle_t h, int64_t* out6);
int hk_profile_enable(hk_handle_t h, int enable);
int hk_profile_read(hk_handle_t h, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* HKMEANS_H_ */
