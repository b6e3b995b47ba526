/*
 * hpf_b200.h -- C ABI of the B200-native coordinate-ascent engine for Hierarchical Poisson
 * Factorization (drop-in for the variational hot path of david-cortes/hpfrec).
 *
 * The reference has no C header: its hot path is a set of Cython `cdef ... noexcept nogil` loops
 * (L1) driven by Cython `def` procedures (L2) in hpfrec/cython_loops.pxi ("pxi" below), reached from
 * Python through the module-level callables of hpfrec.cython_loops_{float,double}.  Every entry
 * point here names the reference interface it replaces.  All pointers are plain C pointers; sizes
 * are 64-bit; no torch / C++ types cross this boundary.  Every function returns 0 on success and a
 * non-zero HPF_E* code on failure; hpf_last_error() gives the message of the calling thread's last
 * failure.  `real_bytes` selects the instantiation exactly like the reference's two modules do:
 * 4 = float (cython_float.pxi:9-10), 8 = double (cython_double.pxi:7-8).  Index arrays may be
 * 4-byte (int32) or 8-byte (the reference's `size_t ind_type`, cython_*_nonwindows.pyx:8).
 *
 * Pointer arguments marked [h|d] may be host or device pointers (resolved with
 * cudaPointerGetAttributes); copies are issued on the engine's stream.
 */
#ifndef HPF_B200_H
#define HPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPF_OK 0
#define HPF_EINVAL 1   /* bad argument                                   */
#define HPF_ECUDA 2    /* a CUDA runtime call / kernel launch failed     */
#define HPF_ESTATE 3   /* call sequence error (e.g. step before load)    */
#define HPF_ENOMEM 4   /* device allocation failed                       */

typedef struct hpf_engine hpf_engine;

/* ---- life cycle ------------------------------------------------------------------------- */

/* Version of this ABI (bumped on incompatible change). */
int hpf_abi_version(void);
const char* hpf_last_error(void);

/* Allocates device state for an (nU x k) user side and an (nI x k) item side on CUDA device
 * `device`.  Owns: Gamma_shp/Gamma_rte (nU x ld), Lambda_shp/Lambda_rte (nI x ld), k_rte (nU),
 * t_rte (nI) -- the six arrays `fit_hpf` allocates at pxi:179-181 -- plus engine-private buffers.
 * `nU` may be a *local* user-row slice when the caller shards users across processes. */
int hpf_create(hpf_engine** out, int64_t nU, int64_t nI, int32_t k, int32_t real_bytes, int32_t device);
int hpf_destroy(hpf_engine* h);

/* Hyper-parameters a, a', b', c, c', d' (arguments 1-6 of fit_hpf, pxi:147-148). */
int hpf_set_hyper(hpf_engine* h, double a, double a_prime, double b_prime,
                  double c, double c_prime, double d_prime);
/* The derived constants exactly as the Cython partial_fit receives them (pxi:429-430: add_k_rte,
 * add_t_rte, a, c, k_shp, t_shp); overrides what hpf_set_hyper derived. */
int hpf_set_constants(hpf_engine* h, double a, double c, double k_shp, double t_shp,
                      double add_k_rte, double add_t_rte);
/* CUDA stream (cudaStream_t) all engine work is issued on; default: the legacy default stream. */
int hpf_set_stream(hpf_engine* h, void* cuda_stream);
/* Tunables (defaults are the measured ones, see the block at the top of hpf_engine.cu):
 *   "panel_mb"  L2 panel size of the gathered factor side (takes effect at the next hpf_load_coo)
 *   "chunk"     nnz walked by one lane group
 *   "sweep"     0 two-pass segmented (default), 1 single-pass COO atomics, 2 fused user-major pass
 *               (gathers + REDs), 3 bulk-copy (TMA) staged gathers, 4 fused item-major pass
 *   "kernel"    two-pass kernel: 1 register gathers, 2 one-step register pipeline, 3 cp.async rings,
 *               4 cp.async rings with vector-loaded triples (no broadcast shuffles; chunk % 4 == 0)
 *   "lpg" / "minb" / "hint" / "block"   shape of the sweep kernel (lanes per row, resident CTAs per SM,
 *               load hints, CTA size); 0 (-1 for hint) = measured default of the row class;
 *               "strict"=1 makes an unknown shape an error instead of falling back to the default
 *   "use_graph" 0/1: replay one CAVI iteration as a CUDA graph
 *   "timing"    0/1: per-kernel CUDA-event timing, see hpf_phase_ms
 * Environment: HPF_OPTIONS="name=value,..." applies these to every new engine; HPF_ROW_ALIGN=32|64|128|256
 * sets the row alignment of the device matrices (default 128: whole cache lines). */
int hpf_set_option(hpf_engine* h, const char* name, double value);

/* ---- state ------------------------------------------------------------------------------ */

/* Uploads the variational state (the tuple initialize_parameters returns, pxi:143; C-contiguous
 * (n x k) matrices and length-n vectors of `real`).  [h|d] */
int hpf_load_state(hpf_engine* h, const void* Gamma_shp, const void* Gamma_rte,
                   const void* Lambda_shp, const void* Lambda_rte,
                   const void* k_rte, const void* t_rte);
/* Downloads state and/or expectations; any pointer may be NULL.  Theta = Gamma_shp/Gamma_rte
 * (pxi:251), Beta = Lambda_shp/Lambda_rte (pxi:256).  Synchronises the stream.  [h|d] */
int hpf_export_state(hpf_engine* h, void* Gamma_shp, void* Gamma_rte, void* Lambda_shp,
                     void* Lambda_rte, void* k_rte, void* t_rte, void* Theta, void* Beta);

/* ---- data ------------------------------------------------------------------------------- */

/* Uploads the observed triples (arguments Y, ix_u, ix_i of fit_hpf, pxi:149-151) and builds the
 * two device orderings the sweep streams (user-major and item-major, L2-panelled).  ix_u must
 * already be local to this engine's user slice.  [h|d] */
int hpf_load_coo(hpf_engine* h, const void* ix_u, const void* ix_i, const void* Y,
                 int64_t nnz, int32_t index_bytes);

/* ---- full-batch CAVI (replaces the loop body pxi:230-259) -------------------------------- */

/* Runs `niter` complete iterations: update_phi (pxi:551) fused with update_G_n_L_sh (pxi:613),
 * then the closed-form rate updates pxi:236, 251, 255-259, all on device. */
int hpf_step_full(hpf_engine* h, int32_t niter);

/* The same iteration split at its only cross-shard dependency, for user-sharded multi-GPU runs
 * (SURVEY §8e).  Call order per iteration:
 *   hpf_sweep            passes A+B; fills the user-side sums and the item-side PARTIAL sums
 *   hpf_update_users     Gamma_shp/Gamma_rte/k_rte of the local users + local column sums of Theta
 *   <caller all-reduces (SUM) the two buffers returned by hpf_partials over all shards>
 *   hpf_update_items     Lambda_shp/Lambda_rte/t_rte, identical on every shard
 * Device pointers: item_sums = (nI x ld) `real`; theta_colsum = k doubles. */
int hpf_sweep(hpf_engine* h);
/* One half of hpf_sweep: side 0 = item-major pass (fills item_sums, call it first so its all-reduce can
 * overlap), side 1 = user-major pass. */
int hpf_sweep_side(hpf_engine* h, int32_t side);
int hpf_update_users(hpf_engine* h);
/* materialize == 0: keep only what the next iteration needs (per-row factors, k_rte, column sums) and skip the
 * stores of Gamma_shp / Gamma_rte, exactly like the intermediate iterations of hpf_step_full; the LAST
 * iteration before an export must materialize. */
int hpf_update_users_ex(hpf_engine* h, int32_t materialize);
int hpf_update_items(hpf_engine* h);
/* hpf_sweep_side(h, 0) + hpf_update_users_ex(h, materialize) as ONE step in which the user update runs UNDER the
 * item-major pass on a second stream (it only needs the user-major pass, which the caller has already run with
 * hpf_sweep_side(h, 1)); the new user factors go into a second buffer and the two buffers swap roles on return.  A
 * caller that captures its loop into a CUDA graph must capture an EVEN number of iterations, so that a replay leaves
 * the roles as it found them.  Falls back to the two steps in sequence when engine option overlap_update is 0. */
int hpf_item_pass_with_user_update(hpf_engine* h, int32_t materialize);
int hpf_partials(hpf_engine* h, void** item_sums, int64_t* item_sums_count, void** theta_colsum,
                 int64_t* theta_colsum_count);

/* Fused alternative to <all-reduce item_sums> + hpf_update_items for one box with NVLink peer access
 * (one process per GPU).  Setup once: every rank calls hpf_peer_export (HPF_PEER_BUFFERS CUDA-IPC
 * handles of HPF_IPC_HANDLE_BYTES each: item_sums, item softmax factors, t_rte, Lambda_shp,
 * Lambda_rte), the caller all-gathers them rank-major and passes the table to hpf_peer_attach.
 * Per iteration, after hpf_sweep + hpf_update_users:
 *   <all-reduce theta_colsum>      (k doubles; also the barrier that makes every rank's item_sums final)
 *   hpf_update_items_peer          ONE kernel: each rank sums ITS slice of item rows over all ranks'
 *                                  item_sums (P2P loads), updates them, stores the results into every
 *                                  rank's replica (P2P stores); materialize != 0 also stores shp/rte
 *   <all-reduce beta_colsum>       (hpf_beta_colsum, k doubles; the barrier that publishes the stores)
 *   hpf_peer_finish                re-zeroes the local item_sums */
#define HPF_IPC_HANDLE_BYTES 64
#define HPF_PEER_BUFFERS 5
int hpf_peer_export(hpf_engine* h, void* handles);
int hpf_peer_attach(hpf_engine* h, int32_t rank, int32_t world, const void* all_handles);
/* The same exchange over caller-provided SYMMETRIC memory (every rank allocates the five item-side buffers at the
 * same offsets of one symmetric allocation, e.g. torch.distributed._symmetric_memory, which also yields an NVSwitch
 * multicast mapping of it): hpf_item_buffer_bytes gives the sizes, hpf_adopt_item_buffers makes the engine use the
 * caller's buffers (before hpf_load_state; never freed by the engine), hpf_peer_attach_ptrs takes the [world][5] table
 * of unicast peer pointers and, optionally, the five multicast pointers.  With multicast pointers the kernel of
 * hpf_update_items_peer reduces with multimem.ld_reduce (the switch sums the ranks' copies) and broadcasts with
 * multimem.st; without them it uses the peer loads / stores above. */
int hpf_item_buffer_bytes(hpf_engine* h, int64_t out[HPF_PEER_BUFFERS]);
int hpf_adopt_item_buffers(hpf_engine* h, void* const bufs[HPF_PEER_BUFFERS]);
int hpf_peer_attach_ptrs(hpf_engine* h, int32_t rank, int32_t world, void* const* peer_ptrs, void* const* mc_ptrs);
int hpf_update_items_peer(hpf_engine* h, int32_t materialize);
/* Optional split of the exchange: the reduce-scatter half alone, launched on `stream` (a cudaStream_t; NULL = the
 * engine's stream) so that it can overlap the user-major pass and hpf_update_users.  It needs every rank's item-major
 * pass (hpf_sweep_side(h, 0)) to be complete -- the caller puts a cross-rank barrier before it -- and leaves the
 * all-rank sums of this rank's slice in its own item_sums buffer; the next hpf_update_items_peer (which the caller
 * orders after it) then skips its own reduction. */
int hpf_reduce_items_peer(hpf_engine* h, void* stream);
int hpf_peer_finish(hpf_engine* h);
int hpf_beta_colsum(hpf_engine* h, void** ptr, int64_t* count);

/* ---- minibatch step (replaces Cython partial_fit pxi:423-473 and the SVI epoch bodies
 *      pxi:275-325 / 329-377) -------------------------------------------------------------- */

/* One minibatch update on explicit COO triples.  users/items = the unique ids present in the batch
 * (`users_this_batch`, `items_this_batch`), rho = step_size_batch, mult = multiplier_batch,
 * user_batch != 0 for a batch of users.  blend_all_rates != 0 reproduces partial_fit (k_rte/t_rte
 * blended over ALL rows, pxi:472-473); 0 reproduces the SVI epochs inside fit_hpf (batch rows
 * only, pxi:324-325, 376-377).  [h|d] */
int hpf_step_batch(hpf_engine* h, const void* ix_u, const void* ix_i, const void* Y, int64_t nnz,
                   const void* users, int64_t n_users, const void* items, int64_t n_items,
                   int32_t index_bytes, int32_t user_batch, double rho, double mult,
                   int32_t blend_all_rates);

/* Same update for a batch given ONLY by its row ids (users if user_batch != 0, else items): the
 * engine gathers the rows' triples from its resident orderings on the device and finds the unique
 * opposite-side ids itself (replaces get_unique_items_batch / get_i_batch_pass1/2, pxi:27-42,
 * 774-797, and the per-batch host -> device copy).  Requires hpf_load_coo to have run with a single
 * L2 panel per side (option "panel_mb" >= size of the larger factor matrix).  ids: [h|d]. */
int hpf_step_batch_ids(hpf_engine* h, const void* ids, int64_t n_ids, int32_t index_bytes,
                       int32_t user_batch, double rho, double mult, int32_t blend_all_rates);

/* One whole SVI epoch of fit_hpf (user epoch pxi:275-325 when user_batch != 0, else item epoch pxi:329-377):
 * `ids` is the shuffled id list of the batched side (pxi:277 / 329); consecutive slices of `batch_rows` ids are
 * the minibatches, each applied like hpf_step_batch_ids with multiplier n / |batch| (pxi:282 / 334) and
 * blend_all_rates = 0.  The list crosses to the device once and nothing inside waits for the device: the call
 * returns with the epoch's kernels in flight on the engine's stream.  ids: [h|d]. */
int hpf_step_epoch_ids(hpf_engine* h, const void* ids, int64_t n_ids, int32_t index_bytes,
                       int64_t batch_rows, int32_t user_batch, double rho);

/* ---- convergence metrics and scoring --------------------------- */

/* llk_plus_rmse (pxi:627) over the given triples using the engine's current Theta/Beta:
 * out[0] = sum_n Y*log(yhat) [- lgamma(Y+1) if full_llk], out[1] = sum_n (Y - yhat)^2,
 * out[2] = sum_n yhat (sum_prediction, pxi:816), out[3] = sum_j (sum_u Theta_uj)(sum_i Beta_ij)
 * (the all-pairs shortcut of pxi:78).  Accumulated in double.  [h|d] */
int hpf_llk(hpf_engine* h, const void* ix_u, const void* ix_i, const void* Y, int64_t nnz,
            int32_t index_bytes, int32_t full_llk, double out[4]);
/* Same on the training triples already resident from hpf_load_coo. */
int hpf_llk_train(hpf_engine* h, int32_t full_llk, double out[4]);
/* predict_multiple (pxi:803): out[n] = Theta[ix_u[n]] . Beta[ix_i[n]].  [h|d] */
int hpf_predict(hpf_engine* h, const void* ix_u, const void* ix_i, int64_t n, int32_t index_bytes,
                void* out);

/* ---- device-resident scoring of a fitted model -------------------------- */

/* Holds Theta (nU x k) and Beta (nI x k) -- the two arrays HPF.predict / eval_llk / topN read
 * (hpfrec/__init__.py:1198-1446) -- on the device, so repeated scoring calls do not re-upload them.  [h|d] */
typedef struct hpf_scorer hpf_scorer;
int hpf_scorer_create(hpf_scorer** out, const void* Theta, const void* Beta, int64_t nU, int64_t nI,
                      int32_t k, int32_t real_bytes, int32_t device);
int hpf_scorer_destroy(hpf_scorer* s);
/* predict_multiple (pxi:803-810): out[n] = Theta[ix_u[n]] . Beta[ix_i[n]].  [h|d] */
int hpf_scorer_predict(hpf_scorer* s, const void* ix_u, const void* ix_i, int64_t n, int32_t index_bytes,
                       void* out);
/* llk_plus_rmse + sum_prediction (pxi:627-658, 816-825): out[0] = sum Y log yhat [- lgamma(Y+1)],
 * out[1] = sum (Y - yhat)^2, out[2] = sum yhat; calc_llk (pxi:525-534) is out[0] - out[2].  [h|d] */
int hpf_scorer_llk(hpf_scorer* s, const void* ix_u, const void* ix_i, const void* Y, int64_t n,
                   int32_t index_bytes, int32_t full_llk, double out[3]);
/* HPF.topN (hpfrec/__init__.py:1296-1396): the n best item rows for user row `user` by Theta[user] . Beta[i],
 * best first, restricted to `pool` (n_pool item rows, NULL = all items) and without the rows in `seen`
 * (NULL = keep all).  out_ids (host, n entries) receives item rows, out_scores (host, `real`, may be NULL)
 * their scores, *n_out how many were written (fewer than n when the pool runs out).  pool/seen: [h|d]. */
int hpf_scorer_topn(hpf_scorer* s, int64_t user, int32_t n, const void* pool, int64_t n_pool,
                    const void* seen, int64_t n_seen, int32_t index_bytes, int64_t* out_ids,
                    void* out_scores, int32_t* n_out);

/* ---- stateless one-shot forms over caller buffers ([h|d]) of the reference's L1 loops ------ */

/* update_phi (pxi:551-553) + update_G_n_L_sh (pxi:613-615) fused: on return G_sh/L_sh hold
 * a + sum phi and c + sum phi (the state after pxi:239-249).  phi may be NULL; if not it receives the
 * (nY x k) multinomial parameters exactly as update_phi writes them. */
int hpf_update_shapes(int32_t real_bytes, int32_t index_bytes, int32_t device,
                      void* G_sh, const void* G_rt, void* L_sh, const void* L_rt, void* phi,
                      const void* Y, const void* ix_u, const void* ix_i,
                      int64_t nU, int64_t nI, int64_t nY, int32_t k, double a, double c);

/* Device digamma used by the engine, exposed for validation against scipy.special.psi. [h|d] */
int hpf_digamma(int32_t real_bytes, int32_t device, const void* x, void* out, int64_t n);

/* Releases the device blocks the library caches between engines (see the caching allocator in
 * hpf_engine.cu; cap: environment variable HPF_CACHE_MB, default 32768). */
int hpf_trim_cache(void);
/* Counters for bench/telemetry: number of kernels this engine has launched so far. */
int hpf_launch_count(hpf_engine* h, int64_t* out);
/* With option "timing"=1: accumulated device milliseconds of the four kernels of a full-batch
 * iteration since the option was set: [item-major pass, user-major pass, user update, item update]. */
int hpf_phase_ms(hpf_engine* h, double out[4], int64_t* iterations);
/* Human-readable resolved configuration (row stride, sweep mode, kernel and its shape, chunk, panels)
 * for telemetry: what bench.py prints next to its numbers.  `lpg=0` means the generic shape of a row
 * class without a measured table entry. */
int hpf_describe(hpf_engine* h, char* buf, int64_t n);
/* Leading dimension (padded k) of the engine's device matrices. */
int hpf_ld(hpf_engine* h, int32_t* out);

/* ---- ingest on the device (the host-side pandas / scipy steps of HPF._process_data and HPF._store_metadata) -------- */

/* pd.factorize of an INTEGER id column (hpfrec/__init__.py:478-479): codes_out[j] = dense code of values[j], numbered
 * in order of first appearance; uniques_out[c] = the id with code c.  values: n ids of value_bytes (4 or 8) [h|d];
 * codes_out: n integers of code_bytes (4 or 8) [h|d]; uniques_out: room for n ids of value_bytes [h|d];
 * *n_unique (host) receives the number of distinct ids.  One stable radix sort of (id, position), one of the runs'
 * first positions, three one-pass kernels.  n < 2^31. */
int hpf_factorize(int32_t device, const void* values, int64_t n, int32_t value_bytes, void* codes_out, int32_t code_bytes,
                  void* uniques_out, int64_t* n_unique);

/* The CSR arrays HPF._store_metadata keeps (hpfrec/__init__.py:587-606: coo_array(...).tocsr(), i.e. duplicates of a
 * (user, item) pair merged and item ids ascending within a user): indptr_out = nU + 1 int64 [h|d], indices_out = room
 * for n item ids of out_index_bytes (4 or 8) [h|d], *n_out (host) = number of distinct pairs.  Out-of-range indices are
 * rejected with HPF_EINVAL. */
int hpf_csr_metadata(int32_t device, const void* ix_u, const void* ix_i, int64_t n, int32_t index_bytes, int64_t nU, int64_t nI,
                     int64_t* indptr_out, void* indices_out, int32_t out_index_bytes, int64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* HPF_B200_H */
