/*
 * mfcd_b200.h -- C ABI of the B200-native (sm_100a) hot path for
 * MayeulCassier/Matrix-Factorization-With-Comparison-Data.
 *
 * The reference is pure Python and has no FFI of its own:
 * its boundary is the function surface of `structure.py` / `generation_data.py`.
 * The repo-root modules of the same names keep those signatures and call the
 * entry points below through ctypes (matrix-factorization-with-comparison-data_b200/_lib.py).
 * Each entry point names the reference lines (relative to /root/reference) whose
 * per-call ATen / Python work it replaces.
 *
 * Conventions
 *   - plain `extern "C"`, raw DEVICE pointers + sizes, no torch types;
 *   - every buffer is owned and allocated by the caller; nothing is retained
 *     after the call returns (work is enqueued on `stream`, a cudaStream_t
 *     passed as void*; NULL = the legacy default stream);
 *   - return 0 on success, MFCD_ERR_* (<0) for bad arguments, a positive
 *     cudaError_t otherwise; `mfcd_last_error()` gives a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns an error.
 */
#ifndef MFCD_B200_H
#define MFCD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFCD_ABI_VERSION 1

#define MFCD_OK 0
#define MFCD_ERR_ARG (-1)
#define MFCD_ERR_WORKSPACE (-2)
#define MFCD_ERR_UNSUPPORTED (-3)

/* One labelled comparison: user u prefers item i over item j with label z
 * (hard 0/1 or a soft label in [0,1]).  16 bytes, moved with one 128-bit load.
 * Replaces the python tuples of BTLPreferenceDataset.data (structure.py:507-519)
 * and the int64/int64/int64/float64 batches default_collate builds from them. */
typedef struct mfcd_triplet {
  int32_t u, i, j;
  float z;
} mfcd_triplet;

/* Ground-truth matrix X as the kernels see it: either a dense row-major fp32
 * matrix (X != NULL, leading dimension ldx) or a low-rank product
 * X[u][i] = scale * <A[u,:], B[i,:]> (X == NULL; A is n x dx, B is m x dx). */
typedef struct mfcd_xview {
  const float* X;
  int64_t ldx;
  const float* A;
  const float* B;
  int32_t dx;
  float scale;
} mfcd_xview;

/* ---- library / device info ------------------------------------------------ */
int mfcd_abi_version(void);
const char* mfcd_last_error(void);
int mfcd_device_sm_count(int* out);

/* Live timing of K1 (mfcd_triplet_fwd_bwd*, also inside mfcd_train_epoch): while enabled (per calling thread),
 * every K1 launch is bracketed by a CUDA event pair on its own stream.  mfcd_profile_k1_read synchronises those
 * events, returns the summed duration and the number of launches, and forgets them. */
int mfcd_profile_k1(int32_t enable);
int mfcd_profile_k1_read(double* total_ms, int64_t* launches);

/* ---- data staging ---------------------------------------------------------- */
/* int64 u,i,j + float64 z columns (what the reference's DataLoader yields,
 * structure.py:846) -> 16-byte records. */
int mfcd_pack_triplets(const int64_t* u, const int64_t* i, const int64_t* j, const double* z,
                       int64_t N, mfcd_triplet* out, void* stream);
int mfcd_unpack_triplets(const mfcd_triplet* rec, int64_t N, int64_t* u, int64_t* i, int64_t* j,
                         double* z, void* stream);
/* 8-byte wire format for HARD-labelled records, for host <-> device staging (PCIe moves half the bytes):
 * bit 0 = label, bits [1,21) = j, [21,41) = i, [41,64) = u  (n_users <= 2^23, n_items <= 2^20).
 * pack sets *bad (device int, caller zeroes it) if a record has a soft label or an index out of range. */
int mfcd_pack_triplets8(const mfcd_triplet* rec, int64_t N, uint64_t* out, int32_t* bad, void* stream);
int mfcd_unpack_triplets8(const uint64_t* packed, int64_t N, mfcd_triplet* out, void* stream);
/* The same packing on the HOST's cores, for a loader that owns raw 16-byte records and feeds the GPU over PCIe:
 * host records -> host words (both plain host pointers; `out` is normally a pinned staging buffer), split over
 * `threads` threads of a persistent pool (<= 0: all hardware threads), AVX-512 with streaming stores where the
 * CPU has it.  Bit-identical to mfcd_pack_triplets8.  *bad (host int, caller zeroes it) is set for soft labels
 * or indices out of range.  Needs no GPU.  (The reference moves 28 bytes per sample per step through
 * `x.to(device)`, structure.py:845-846.)  mfcd_host_pack_isa: 512 if the AVX-512 kernel is in use, else 0
 * (environment MFCD_HOST_PACK_ISA=scalar, read once per process, forces the portable loop). */
int mfcd_host_pack_triplets8(const mfcd_triplet* rec, int64_t N, uint64_t* out, int32_t threads, int32_t* bad);
int mfcd_host_pack_isa(void);
/* Run-length wire format for ONE user-grouped batch of B hard-labelled records with n_items <= 65536
 * (4.375 bytes per triplet + 4 per run of equal users instead of 16): u32 words
 *   [n_runs, B, 0, 0 | word_run0[nw] | zbits[nw] | nbits[nw] | ij[B] | users[n_runs]],  nw = ceil(B/32);
 *   word_run0[w] = run starts before triplet 32 w, nbits = run-start bits, ij = i | j << 16;
 *   `fixed_words` = words before users[], `capacity_words` = fixed_words + B (worst case).
 * pack: device records -> device wire (word 0 = n_runs tells how many words to ship: fixed_words + n_runs);
 * sets *bad (caller zeroes it) for soft labels or items >= 65536.  unpack: device wire -> device records,
 * the per-step side of host staging (structure.py:846 `x.to(device)` moved 32 bytes per sample). */
int mfcd_wire_layout(int64_t B, int64_t* fixed_words, int64_t* capacity_words, size_t* workspace_bytes);
int mfcd_pack_wire(const mfcd_triplet* rec, int64_t B, uint32_t* wire, int64_t capacity_words, int32_t* bad,
                   void* workspace, size_t workspace_bytes, void* stream);
int mfcd_unpack_wire(const uint32_t* wire, int64_t B, mfcd_triplet* out, void* stream);
/* out[k] = rec[perm[k]] for k in [0,N): materialise one epoch's shuffled order
 * (replaces RandomSampler + default_collate, structure.py:738, :845). */
int mfcd_gather_triplets(const mfcd_triplet* rec, const int32_t* perm, int64_t N, mfcd_triplet* out,
                         void* stream);

/* ---- epoch batching: the epoch reshuffle and the per-batch user grouping in one streaming pass -------------
 * Replaces RandomSampler + default_collate for one epoch (structure.py:738, :845) at throughput batch sizes.
 * Record r of the store has epoch position pos(r): pos[r] when `pos` is given (int32, the INVERSE of the epoch
 * permutation: see mfcd_invert_perm), else a keyed bijection of [0, N) evaluated on the fly (a 3-round mixed-radix
 * network over a domain c * 2^k that hugs N, c <= 64; mfcd_epoch_positions writes the same values out).
 * Batch b = { r : pos(r) / B == b } -- exactly the batches a loader that walks the permutation in chunks of B
 * forms (every batch has B members, the last one the remainder).  out[] receives batch 0, batch 1, ... each in
 * STORE order (stable), so a store sorted by user yields user-grouped batches (MFCD_FLAG_USER_GROUPED) for free.
 * At most mfcd_epoch_max_batches() batches per epoch and N < 2^31, else MFCD_ERR_UNSUPPORTED (callers then gather
 * through an explicit permutation, mfcd_gather_triplets).  out must not alias rec. */
int mfcd_epoch_max_batches(int32_t* out);
int mfcd_epoch_positions(int64_t N, uint64_t seed, int32_t* pos, void* stream);
int mfcd_invert_perm(const int32_t* perm, int64_t N, int32_t* pos, void* stream);
int mfcd_epoch_batches_workspace(int64_t N, int64_t B, size_t* bytes);
int mfcd_epoch_batches(const mfcd_triplet* rec, int64_t N, int64_t B, const int32_t* pos, uint64_t seed,
                       mfcd_triplet* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K1: fused forward + BCE + backward, atomic scatter --------------------
 * Replaces structure.py:848-850 (model forward :787-795, F.binary_cross_entropy,
 * loss.backward()) for the batch rec[perm[start..start+B)] (perm may be NULL =
 * identity).  U is n x d, V is m x d, row-major fp32.  Adds the batch's dense
 * gradients into gU / gV with red.global.add (they must hold zeros, or the sum
 * being accumulated) and adds  sum_b BCE_b * inv_batch  to *loss.
 * inv_batch = 1 / (global batch size): 1/B on one GPU, 1/(B*world) under
 * data-parallel training. */
int mfcd_triplet_fwd_bwd(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                         int64_t start, int64_t B, int32_t d, float inv_batch, float* gU, float* gV,
                         float* loss, void* stream);

/* Same as mfcd_triplet_fwd_bwd with hot-row privatisation for skewed item popularity: the caller names
 * up to mfcd_max_hot_items(d) item rows that receive a large share of the updates
 * (item_slot[row] = slot index in [0, n_hot) or -1, n_items int8 entries; hot_items[slot] = row).  Their
 * gradient rows are accumulated in per-warp shared-memory images and reach gV as one reduction per row
 * and CTA instead of one per triplet.  Results equal mfcd_triplet_fwd_bwd up to fp32 summation order. */
int mfcd_max_hot_items(int32_t d, int32_t* out);
int mfcd_triplet_fwd_bwd_hot(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                             int64_t start, int64_t B, int32_t d, float inv_batch, float* gU, float* gV,
                             float* loss, const int8_t* item_slot, const int32_t* hot_items, int32_t n_hot,
                             void* stream);

/* General form of the atomic-mode K1: hot rows optional (NULL / 0), plus `flags`:
 *   MFCD_FLAG_USER_GROUPED  the batch keeps each user's triplets adjacent (see mfcd_group_by_user) and perm
 *                           is NULL: U[u] is then read once and its gradient row leaves as one reduction per
 *                           run of equal u instead of one per triplet.  The flag is a hint: results are
 *                           correct for any order (same sums, different fp32 summation order).
 *   MFCD_FLAG_WIRE_RLE      `rec` is not an array of records but ONE batch in the run-length wire format
 *                           (mfcd_pack_wire; start = 0, perm = NULL, B = its size): K1 decodes it on the fly,
 *                           so a batch staged from the host needs no unpack pass.  d must be a multiple of
 *                           4 that the float4 lane groups cover exactly (4..128, 256, 384, 512). */
#define MFCD_FLAG_USER_GROUPED 1
#define MFCD_FLAG_WIRE_RLE 2
int mfcd_triplet_fwd_bwd_ex(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                            int64_t start, int64_t B, int32_t d, float inv_batch, float* gU, float* gV,
                            float* loss, const int8_t* item_slot, const int32_t* hot_items, int32_t n_hot,
                            int32_t flags, void* stream);

/* Reorders rec[0..N) in place so that inside every batch of batch_size consecutive records the triplets
 * of one user are adjacent (stable).  Batch membership is unchanged, so a training step over a batch is
 * the same sum in a different order -- the reference draws batches from a shuffled DataLoader
 * (structure.py:738) and never depends on the order inside one.  Workspace: see the query function. */
int mfcd_group_by_user_workspace(int64_t N, int64_t batch_size, size_t* bytes);
int mfcd_group_by_user(mfcd_triplet* rec, int64_t N, int64_t batch_size, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---- K1 deterministic variant ------------------------------------------------
 * Same contract, but run-to-run bit-reproducible.
 *   B <= 256 (the reference's regime): one CTA sums every row's contributions in batch order, the order of the
 *   reference's sequential CPU index_put_(accumulate);
 *   larger B, engine "fixed" (k1_fixed.cu; default up to 16384 triplets per batch): every contribution g * x is
 *   rounded once to 64-bit fixed point and added with integer atomics into a 64-bit image of the gradient tables
 *   held in the workspace (integer sums do not depend on the order the atomics land in -- nor on the order of the
 *   batch), then converted to fp32 with one rounding per element and added to gU / gV; the loss likewise.  Two
 *   launches, no sort;
 *   larger B, engine "sort" (segmented.cu; default above 16384): stable radix sort by destination row
 *   (cub::DeviceRadixSort) + segmented reduction, batch order kept inside a row; 17 launches, but vector fp32
 *   traffic: faster on big batches (64-bit atomics run at a quarter of the vector-fp32 reduction rate and
 *   serialise on hot rows).
 * mfcd_det_workspace_bytes_nm = what the default engine for this B wants; MFCD_DET_ENGINE=fixed|sort (environment)
 * forces one engine; the _fixed / _sort entry points select one explicitly. */
int mfcd_det_workspace_bytes(int64_t B, int32_t d, size_t* bytes);
int mfcd_det_workspace_bytes_nm(int64_t B, int32_t d, int64_t n_users, int64_t n_items, size_t* bytes);
int mfcd_triplet_fwd_bwd_det(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                             int64_t start, int64_t B, int32_t d, float inv_batch, int64_t n_users,
                             int64_t n_items, float* gU, float* gV, float* loss, void* workspace,
                             size_t workspace_bytes, void* stream);
/* the two engines under their own names (workspaces: mfcd_det_fixed_workspace_bytes / mfcd_det_workspace_bytes) */
int mfcd_det_fixed_workspace_bytes(int32_t d, int64_t n_users, int64_t n_items, size_t* bytes);
int mfcd_triplet_fwd_bwd_det_fixed(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                                   int64_t start, int64_t B, int32_t d, float inv_batch, int64_t n_users,
                                   int64_t n_items, float* gU, float* gV, float* loss, void* workspace,
                                   size_t workspace_bytes, void* stream);
int mfcd_triplet_fwd_bwd_det_sort(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                                  int64_t start, int64_t B, int32_t d, float inv_batch, int64_t n_users,
                                  int64_t n_items, float* gU, float* gV, float* loss, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* ---- K3: fused dense optimiser update --------------------------------------
 * torch.optim.Adam single-tensor semantics with coupled L2 (structure.py:364,
 * :851): g += wd*p; m += (g-m)(1-b1); v = b2*v + (1-b2)g*g;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps), t = step (1-based).
 * With zero_grad != 0 the gradient buffer is cleared in the same pass
 * (optimizer.zero_grad(), structure.py:847). */
int mfcd_adam_update(float* p, float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int64_t step, int32_t zero_grad,
                     void* stream);
/* torch.optim.SGD (dampening 0, no nesterov); buf may be NULL when momentum == 0. */
int mfcd_sgd_update(float* p, float* g, float* buf, int64_t numel, float lr, float momentum,
                    float weight_decay, int64_t step, int32_t zero_grad, void* stream);

/* ---- K9: fused gradient exchange + Adam over NVLink peer memory (data parallel) -------------
 * Replaces "all-reduce the flat gradient, then mfcd_adam_update on every rank" by one kernel per
 * rank: rank r owns the slice mfcd_dp_shard_range(numel, r, world) of the flat parameter vector,
 * sums that slice of the gradient over all ranks by reading the peers' gradient buffers directly
 * (peer_grads[q] = address of rank q's buffer as mapped in THIS process; or, when mc_grads /
 * mc_params are non-zero NVSwitch multicast addresses of the same buffers, one
 * multimem.ld_reduce), applies Adam (m, v: this rank's full-size moment arrays, only the owned
 * slice is touched) and writes the updated slice into every rank's parameter replica
 * (peer_params[q], or one multimem.st).  peer_* are HOST arrays of `world` device addresses.
 * The caller must (a) barrier across ranks before the call (all local gradients final) and
 * after it (all replicas complete, all gradient reads done), (b) then clear its gradients. */
int mfcd_dp_shard_range(int64_t numel, int32_t rank, int32_t world, int64_t* begin, int64_t* end);
int mfcd_dp_fused_adam(const uint64_t* peer_grads, const uint64_t* peer_params, uint64_t mc_grads,
                       uint64_t mc_params, int32_t rank, int32_t world, int64_t numel, float* m, float* v,
                       float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                       void* stream);

/* The same exchange with the synchronisation INSIDE the kernel (no host-visible barriers, no separate gradient
 * memset): peer_flags[q] = address of rank q's flag array (16 uint32, zero-initialised once, peer-mapped like the
 * other buffers): words [0,8) "gradients of rank r ready up to sequence number s", words [8,16) "rank r done with
 * my buffers up to s".  The kernel publishes ready[rank] = seq to every rank, waits for all ready[] >= seq, does
 * reduce-scatter + Adam + all-gather, writes zeros over the gradient slice it consumed in every rank's buffer,
 * then its last CTA publishes done[rank] = seq and waits for all done[] >= seq before exiting.  `seq` must grow by
 * one per call, identically on all ranks, starting at 1.  cta_counter: one zeroed device word reused across
 * calls.  *error (device) becomes 1 if a bounded spin timed out (results are then invalid).
 * zero_local != NULL: DOUBLE-BUFFERED gradients.  The caller alternates two peer-mapped gradient buffers (step k
 * accumulates into buffer k & 1 and passes that buffer's peer table); zero_local is THIS rank's other buffer
 * (numel floats padded to a multiple of 4, 16-byte aligned), which the kernel clears with local stores instead of
 * writing zeros over the consumed slices through the fabric -- every peer finished reading it before the previous
 * call returned (the done[] exchange).  Both buffers must start zeroed. */
int mfcd_dp_fused_adam_sync(const uint64_t* peer_grads, const uint64_t* peer_params, const uint64_t* peer_flags,
                            uint64_t mc_grads, uint64_t mc_params, int32_t rank, int32_t world, int64_t numel,
                            float* m, float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                            int64_t step, uint32_t seq, uint32_t* cta_counter, int32_t* error, float* zero_local,
                            void* stream);

/* ---- one training epoch, launched from C ------------------------------------
 * The inner loop of train_model (structure.py:845-852) for one epoch on one GPU:
 * for each batch k: K1 (atomic or deterministic) then K3, with the batch-mean
 * loss written to step_losses[k] (device, n_steps floats, pre-zeroed by the
 * call).  Tables, gradients and optimiser state are one flat buffer each,
 * U first then V: params[0 .. n_users*d) is U, the rest is V. */
#define MFCD_OPT_ADAM 0
#define MFCD_OPT_SGD 1
#define MFCD_MODE_ATOMIC 0
#define MFCD_MODE_DETERMINISTIC 1
typedef struct mfcd_epoch_args {
  float* params;
  float* grads;
  float* state1; /* Adam exp_avg / SGD momentum buffer */
  float* state2; /* Adam exp_avg_sq / unused */
  int64_t n_users, n_items;
  int32_t d;
  int32_t optimizer;
  int32_t mode;
  int32_t flags;             /* MFCD_FLAG_* (atomic mode) */
  const mfcd_triplet* rec;
  const int32_t* perm; /* epoch order, length n_samples, or NULL */
  int64_t n_samples;
  int64_t batch_size;
  float lr, beta1, beta2, eps, weight_decay, momentum;
  int64_t step0; /* optimiser steps already taken before this epoch */
  float* step_losses;
  void* workspace;
  size_t workspace_bytes;
  void* stream;
  const int8_t* item_slot;   /* optional hot-row privatisation (atomic mode), see mfcd_triplet_fwd_bwd_hot */
  const int32_t* hot_items;
  int32_t n_hot;
  int32_t reserved2;
} mfcd_epoch_args;
int mfcd_train_epoch(const mfcd_epoch_args* args);
/* Workspace mfcd_train_epoch wants for these arguments (0 = none).  In the reference's own regime --
 * deterministic mode, batch <= 256, Adam, tables of up to ~600k elements -- the whole epoch runs as ONE
 * persistent cooperative kernel (epoch_small.cu: per-CTA ownership of table slices, parameters and Adam
 * moments in registers, ping-pong parameter buffers, one grid barrier per step) instead of two launches
 * per step; the workspace holds the second parameter buffer.  Without it the per-step launches are used. */
int mfcd_train_epoch_workspace(const mfcd_epoch_args* args, size_t* bytes);

/* ---- K4: evaluation ----------------------------------------------------------
 * evaluate_model (structure.py:896-921) and the validation pass of train_model
 * (:858-868): batch_loss[b] = mean BCE of batch b (batches of batch_size in
 * record order, last one ragged), *correct += #{(p > 0.5) == z}.
 * batch_acc: one ZEROED 64-bit word per batch (scratch).  The per-sample losses are summed there in fixed
 * point (integer atomics: exact, and independent of launch shape and of whatever else runs on the GPU), then
 * divided by the batch's sample count and rounded once to fp32 into batch_loss. */
int mfcd_triplet_eval(const float* U, const float* V, const mfcd_triplet* rec, int64_t N, int32_t d,
                      int64_t batch_size, float* batch_loss, unsigned long long* correct, int64_t* batch_acc,
                      void* stream);
/* compute_ground_truth_metrics (structure.py:1100-1127): per-batch mean of
 * (sigmoid(X[u,i]-X[u,j]) - z)^2 (no scale s) and #{(diff > 0) == z}; batch_acc as above. */
int mfcd_ground_truth_eval(const mfcd_xview* X, const mfcd_triplet* rec, int64_t N, int64_t batch_size,
                           float* batch_mse, unsigned long long* correct, int64_t* batch_acc, void* stream);
/* MatrixFactorization.forward (structure.py:773-795): p[k] = sigmoid(<U_u, V_i - V_j>). */
int mfcd_triplet_scores(const float* U, const float* V, const int64_t* u, const int64_t* i,
                        const int64_t* j, int64_t N, int32_t d, float* p, void* stream);

/* ---- K7: triplet candidate samplers ------------------------------------------
 * Candidate c of a round uses Philox4x32-10 at (seed, counter0 + c): it is a
 * pure function of (seed, counter), so rounds and ranks draw disjoint streams.
 * Each writes keys[c] = (u*m + i)*m + j, or MFCD_KEY_NONE when the candidate
 * fails the strategy's own test (i == j, margin, ...).
 *   random     : generation_data.py:16-26   u ~ U[0,n), i,j ~ U[0,m)
 *   margin     : generation_data.py:46-84   + |X[u,i]-X[u,j]| <= margin
 *   popularity : generation_data.py:103-128 i ~ probs, j ~ probs without i (cdf = inclusive fp64 prefix sums, cdf[m-1] = total)
 *   svd block  : generation_data.py:164-174 u ~ U(top_users), i != j ~ U(top_items)
 */
#define MFCD_KEY_NONE 0xFFFFFFFFFFFFFFFFull
int mfcd_sample_random(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                       uint64_t* keys, void* stream);
int mfcd_sample_margin(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                       const mfcd_xview* X, float margin, uint64_t* keys, void* stream);
int mfcd_sample_popularity(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                           const double* cdf, uint64_t* keys, void* stream);
int mfcd_sample_block(int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                      const int32_t* top_users, int64_t n_top_users, const int32_t* top_items,
                      int64_t n_top_items, uint64_t* keys, void* stream);

/* Per-user candidate lists (next-ring strategies of SURVEY 8f): list_i is n x ki, list_j is n x kj item ids;
 * u ~ U[0,n), i ~ U(list_i[u]), j ~ U(list_j[u]).  proximity (generation_data.py:29-43): top-k and bottom-k
 * lists; top_k (generation_data.py:189-224): same_list = 1, one list, j drawn among the other k-1 entries. */
int mfcd_sample_lists(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                      const int32_t* list_i, int32_t ki, const int32_t* list_j, int32_t kj, int32_t same_list,
                      uint64_t* keys, void* stream);

/* Sequential "accept if new" over a candidate stream, done with a stable sort:
 * candidate c is accepted iff keys[c] != NONE, it equals no seen[] key and no
 * earlier candidate; the first `want` accepted keys are written, in stream
 * order, to out[0..*n_out).  Same result as the reference's python set loop
 * run over the same candidates (generation_data.py:24-25). */
int mfcd_unique_workspace_bytes(int64_t n_seen, int64_t count, size_t* bytes);
int mfcd_unique_accept(const uint64_t* seen, int64_t n_seen, const uint64_t* keys, int64_t count,
                       int64_t want, uint64_t* out, int64_t* n_out /* device */, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- K8: BTL label sampler ------------------------------------------------------
 * BTLPreferenceDataset._generate_labels (structure.py:507-519).  For triplet t
 * (key[t], decoded with m): q = sigmoid(scale*(X[u,i]-X[u,j])); draw k is
 * `uniform(t,k) < q` with uniform(t,k) the 24-bit uniform of Philox word
 * (t*K+k) (word w = component w%4 of Philox(seed, w/4)).
 * soft == 0: writes N*K records, the K copies of a triplet consecutive;
 * soft != 0: writes N records with z = mean of the K draws.
 * If `uniforms` is non-NULL (N*K floats) they are used instead of Philox:
 * replaying host uniforms gives the host's labels bit for bit. */
int mfcd_btl_labels(const mfcd_xview* X, const uint64_t* keys, int64_t N, int64_t m, int32_t K,
                    float scale, int32_t soft, uint64_t seed, const float* uniforms, mfcd_triplet* out,
                    void* stream);
/* word-exact view of the generator above, for tests: out[w] = 24-bit uniform of word w0+w */
int mfcd_philox_uniforms(uint64_t seed, uint64_t w0, int64_t count, float* out, void* stream);

/* ---- K5: reconstruction statistics ------------------------------------------------
 * One pass over W = U V^T (never materialised) and X for compute_reconstruction_error
 * (structure.py:939-955) and compute_alpha_and_norm_ratios (:980-1064).
 * row_stats[r*8 + k], fp64, for row r with x = X[r,:], w = W[r,:], a = <U_r, vbar>
 * (row mean of W), b_c = <ubar, V_c> (column mean of W):
 *   0: sum x   1: sum x^2   2: sum (w-a)   3: sum (w-a)^2   4: sum x (w-a)
 *   5: sum (w - b_c - s x)^2   6: a   7: unused
 * ubar[d] / vbar[d] are the column means of U and V (fp32, device). */
int mfcd_table_col_means(const float* T, int64_t rows, int32_t d, float* mean, void* stream);
int mfcd_recon_stats(const float* U, const float* V, int64_t n, int64_t m, int32_t d, const mfcd_xview* X,
                     float s, const float* ubar, const float* vbar, double* row_stats, void* stream);
/* Tensor-core variant of mfcd_recon_stats: 128 x 64 tiles of W from tcgen05.mma (kind::tf32 with a
 * hi/lo operand split, three MMAs per K step, fp32-grade products), accumulators in TMEM, operand and X
 * tiles fed by TMA tensor copies (cp.async.bulk.tensor.2d, 128-byte swizzle).  Eligible when X is dense
 * with 16-byte aligned rows (ldx % 4 == 0) and d <= 64; otherwise returns MFCD_ERR_UNSUPPORTED and callers
 * use mfcd_recon_stats.  `workspace` (mfcd_recon_stats_tc_workspace_bytes) holds the split staging tables.
 * *error_flag (device int) becomes non-zero if the kernel's internal pipeline timed out. */
int mfcd_recon_stats_tc_workspace_bytes(int64_t n, int64_t m, int32_t d, size_t* bytes);
int mfcd_recon_stats_tc(const float* U, const float* V, int64_t n, int64_t m, int32_t d, const mfcd_xview* X,
                        float s, const float* ubar, const float* vbar, double* row_stats, int32_t* error_flag,
                        void* workspace, size_t workspace_bytes, void* stream);
/* rows [r0, r0+nr) of W = U V^T into out (nr x m, row-major): structure.py:389-392
 * sampled rows, and the row blocks the Spearman pass ranks. */
int mfcd_reconstruct_rows(const float* U, const float* V, int64_t r0, int64_t nr, int64_t m, int32_t d,
                          float* out, void* stream);
int mfcd_xview_rows(const mfcd_xview* X, int64_t r0, int64_t nr, int64_t m, float* out, void* stream);

/* ---- K6: row ranks for Spearman -----------------------------------------------------
 * scipy.stats.spearmanr semantics (structure.py:1024-1031): ranks[r][c] = 1-based
 * rank of vals[r][c] within its row, ties sharing their average rank.
 * Then rho[r] = Pearson(rank_a[r,:], rank_b[r,:]) in fp64. */
int mfcd_rank_workspace_bytes(int64_t rows, int64_t m, size_t* bytes);
int mfcd_row_ranks(const float* vals, int64_t rows, int64_t m, float* ranks, void* workspace,
                   size_t workspace_bytes, void* stream);
int mfcd_row_pearson(const float* a, const float* b, int64_t rows, int64_t m, double* rho, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFCD_B200_H */
