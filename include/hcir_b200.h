/* hcir_b200 -- C ABI of the B200-native exact cosine top-k / kNN-vote path.
 *
 * The reference (atunnd/Hair-centric-Image-Retrieval) has no FFI of its own: its hot path is
 * a Python call surface that hands dense matrices to sklearn / numpy / torch.  Each entry
 * point below replaces one of those library calls; the reference file:line it stands in for
 * is cited on the declaration.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *   - the caller owns all buffers, including workspaces; the library allocates nothing and
 *     keeps no state except a thread-local last-error string;
 *   - kernels are enqueued on `stream` and the call returns immediately (asynchronous);
 *   - return value: HCIR_OK (0) or a negative HCIR_E* code; hcir_last_error() explains it;
 *   - there is NO CPU fallback and no other backend: on a device that is not sm_100 every
 *     compute entry point fails with HCIR_EARCH.
 *   - internal bank layout: row-major, row stride `ld` = D rounded up to a multiple of 64
 *     elements, zero padded (hcir_padded_dim()).  fp32 bank + bf16 bank of the same `ld`.
 *   - a candidate KEY (uint64) = order-preserving bits of the fp32 similarity << 32
 *     | (0xFFFFFFFF - gallery row).  Larger key = better (desc. similarity, asc. index).
 *   - gallery shards hold fewer than 2^31 rows; the global index space is < 2^32 - 1.
 */
#ifndef HCIR_B200_H_
#define HCIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCIR_OK 0
#define HCIR_EINVAL (-1)     /* bad shape / alignment / k > N / null pointer              */
#define HCIR_EARCH (-2)      /* current device is not sm_100 (B200)                        */
#define HCIR_ECUDA (-3)      /* a CUDA runtime / driver call failed                        */
#define HCIR_EWORKSPACE (-4) /* workspace too small                                        */

#define HCIR_ABI_VERSION 5

/* hcir_plan_t.flags: measurement aids, all 0 in production */
#define HCIR_FLAG_NO_EMIT 1     /* main pass emits nothing: pure contraction throughput        */
#define HCIR_FLAG_SAMPLE_ONLY 2 /* enqueue only the sample pass + threshold kernel             */
#define HCIR_FLAG_MAIN_ONLY 4   /* enqueue only the main pass (thr0 already in the workspace)  */
#define HCIR_FLAG_ROTATE 32     /* main pass: units start their gallery walk at staggered tiles */
#define HCIR_FLAG_CTA_PAIRS 16  /* main pass on CTA pairs: tcgen05.mma.cta_group::2, 256x256 tile */
#define HCIR_FLAG_K3_WIDTH_SHIFT 8 /* hcir_select_rescore CTA width: 0 = chosen from the shape,          */
#define HCIR_FLAG_K3_WIDTH_MASK 0x300 /* 1 = 128, 2 = 256, 3 = 1024 threads per query (A/B measurement) */

typedef void* hcir_stream_t; /* cudaStream_t */

int hcir_abi_version(void);
const char* hcir_last_error(void);
/* 1 if the current device can run the kernels (compute capability 10.x), else 0. */
int hcir_device_supported(void);
/* D rounded up to the internal row stride (multiple of 64). */
int hcir_padded_dim(int d);

/* K1 -- fused row L2-normalise (+ cast).  Replaces torch.nn.functional.normalize(f, dim=1)
 * at HairPretraining/src/classification_engine.py:50,62 and qualitative_test.py:57,76, the
 * sklearn `normalize` inside cosine_similarity (src/models/hair_encoder.py:193) and
 * faiss.normalize_L2 (HairPretraining/app/inference.py:75,90).
 *   x        [n, d]  fp32, row stride ldx (elements)
 *   out_f32  [n, ld] fp32 unit rows, zero padded          (nullable)
 *   out_bf16 [n, ld] bf16 unit rows, zero padded          (nullable, uint16 storage)
 *   out_delta[n]     || unit_row - float(bf16(unit_row)) ||_2  (nullable; feeds the
 *                    certification bound of hcir_select_rescore)
 * y = x / max(||x||_2, 1e-12). */
int hcir_l2norm_cast(const float* x, int64_t n, int d, int64_t ldx, float* out_f32,
                     uint16_t* out_bf16, int ld, float* out_delta, hcir_stream_t stream);

/* K2 -- similarity contraction fused with the top-kc candidate filter.  Replaces
 * torch.mm(q, G.t()) (qualitative_test.py:79, dual_view_model.py:333), the sgemm inside
 * sklearn's pairwise cosine (classification_engine.py:82; hair_encoder.py:193) and faiss
 * IndexFlatL2.search (inference.py:108) together with the N-wide partial sort that follows
 * them.  bf16 x bf16 -> fp32 on tcgen05 tensor cores (TMA-fed, accumulators in TMEM); the
 * [nq, ng] similarity matrix is never written to memory.  One call enqueues up to three
 * kernels:
 *   1. sample pass   the same contraction over a strided sample of `sample_rows` gallery rows
 *                    (addressed in place through the TMA row stride); the epilogue keeps only
 *                    the maximum of every `chunk_w` consecutive sample columns;
 *   2. thresholds    thr0[q] = thr_rank-th largest chunk maximum, thr_rank = min(kc, j):
 *                    kc is the deterministic bound (kc real gallery rows score >= it); j is the
 *                    Poisson-tail rank for which, with probability >= 1 - 1e-7 per query, at least
 *                    kc gallery rows still beat the statistic while only ~4-7 x kc rows pass it.
 *                    Exactness never depends on that probability: hcir_select_rescore certifies
 *                    against the threshold the lists really ended with; a short query is completed
 *                    like any other uncertified one.  thr_hi[q] = staging hint for K3 (= thr0 now);
 *   3. main pass     full contraction; the epilogue compares every accumulator value with the
 *                    query's threshold and appends the survivors (64-bit keys) to one list per
 *                    (query, gallery split).  A list that fills up is pruned back to its kc
 *                    best in place and its threshold raised (rare).
 * Output: for every query and split a list of at most `cap` candidate KEYS (bf16-contraction
 * scores) and the threshold the list ended with, such that every gallery row of that split
 * NOT listed scores <= that threshold.
 *
 * hcir_simtopk_plan fills the launch plan and the workspace layout (all offsets in bytes):
 *   int32 counts[nq][nlists]; float thr_out[nq][nlists]; float thr0[nq]; float thr_hi[nq];
 *   float cmax[nq][num_chunks]; uint64 keys[nq][nlists][cap]  (RAW keys: fp32 bits << 32 |
 *   0xFFFFFFFF - row; the ordered form of the header comment is applied on load). */
typedef struct {
  int32_t nsplit;        /* gallery splits of the main pass (one list per query per split) */
  int32_t cap;           /* capacity of one candidate list (>= kc + 64)                    */
  int32_t kc;            /* candidates that must survive per query                         */
  int32_t flags;         /* HCIR_FLAG_* (0 in production)                                  */
  int32_t sample_rows;   /* rows of the strided sample; 0 = no sample pass (thr0 = -inf)   */
  int32_t sample_stride; /* sample row i = gallery row i * sample_stride                   */
  int32_t chunk_w;       /* 8, 16 or 32 sample columns per chunk maximum                   */
  int32_t num_chunks;    /* sample_rows / chunk_w                                          */
  int32_t sample_nsplit; /* splits of the sample pass                                      */
  int32_t nlists;        /* candidate lists per query (= nsplit x column slices per tile)    */
  int32_t hint_rank;     /* thr_hi[q] = hint_rank-th largest chunk maximum (K3 staging hint; <= thr_rank) */
  int32_t q_rows;        /* rows the query buffer really holds (>= nq; 0 = nq).  A buffer padded
                          * to a multiple of 128 rows keeps the query TMA box in bounds, which
                          * is measurably faster than TMA out-of-bounds zero fill for small nq */
  int32_t thr_rank;      /* thr0[q] = thr_rank-th largest chunk maximum = the main-pass threshold:
                          * kc (deterministic bound) or the smaller Poisson-tail rank (simtopk.cu)  */
  int32_t reserved_;
  uint64_t counts_off, thr_out_off, thr0_off, thr_hi_off, cmax_off, keys_off;
  uint64_t bytes;        /* total workspace bytes                                          */
} hcir_plan_t;

int hcir_simtopk_plan(int64_t nq, int64_t ng, int ld, int kc, int sm_count, hcir_plan_t* plan);
int hcir_simtopk(const uint16_t* q_bf16, int64_t nq, const uint16_t* g_bf16, int64_t ng, int ld,
                 const hcir_plan_t* plan, void* workspace, hcir_stream_t stream);
/* The same launch gated by a device-side count: every CTA returns at once when *active == 0
 * (the second pass of a device-driven completion: it streams the gallery only if a query needs it). */
int hcir_simtopk_gated(const uint16_t* q_bf16, int64_t nq, const uint16_t* g_bf16, int64_t ng, int ld,
                       const hcir_plan_t* plan, void* workspace, const int32_t* active,
                       hcir_stream_t stream);
/* Debug / test entry: same kernel, additionally dumps the raw fp32 accumulator tile values
 * to scores[nq][ng] (row-major).  Only for small problems. */
int hcir_simtopk_debug(const uint16_t* q_bf16, int64_t nq, const uint16_t* g_bf16, int64_t ng,
                       int ld, const hcir_plan_t* plan, void* workspace, float* scores,
                       hcir_stream_t stream);

/* K3 -- candidate selection + fp32 re-score + exact sort + certification, and the fused TAIL of the
 * step.  Replaces torch.topk(sim, k) (qualitative_test.py:82), np.argsort(s)[::-1][:k]
 * (hair_encoder.py:194) and sklearn's argpartition+argsort (_kneighbors_reduce_func); the tail
 * replaces sklearn's `_mode(_y[neigh_ind])` vote (classification_engine.py:82) for single-GPU steps.
 * Per query (one CTA): stream the split lists, keep the kc best by bf16 score, re-score with the
 * canonical fp32 dot product only the candidates that can still reach the top-k, emit the
 * exact top-k in canonical order, and certify it:
 *     certified <=> fp32_score(k-th) > t' + eps(query),
 *     t' = max(largest list threshold, bf16 score of the kc-th best candidate when any listed
 *          candidate was left out of the kc best)
 *     eps = g_delta_max*(1+q_delta) + q_delta*(1+1e-6) + eps_acc
 * Uncertified queries are appended to uncert_list for the completion pass / hcir_exact_topk.
 *   out_sim [nq,k] fp32 descending, out_idx [nq,k] int64 (= local row + idx_offset).
 *   uncert_state  int32[4], all zero before the first launch, maintained by the kernel:
 *                 [0] running count (0 again when the launch ends), [1] RESULT: uncertified queries
 *                 of the last launch, [2] done-CTA counter (0 again when the launch ends).
 *                 No memset between launches: a captured step has no fill node.
 *   tail (nullable) what every query's CTA does with its finished top-k:
 *     labels != NULL   gather the neighbours' class indices (labels[row], row = LOCAL gallery row)
 *                      -> out_lab [nq,k] (nullable)
 *     pred != NULL     kNN vote (T <= 0 uniform, T > 0 temperature; hcir_vote's arithmetic and tie
 *                      rule) -> pred[q] = classes[best] (class index when classes is NULL)
 *     world > 0        multi-GPU: store the query's results into slot (parity of step *step + 1,
 *                      rank) of EVERY peer region (hcir_peer_* below) over NVLink --
 *                      payload 1: the packed block rows idx | sims | labels (hcir_packed_block_bytes
 *                      layout), payload 2: the int64 prediction -- and let the LAST CTA of the launch
 *                      write this rank's uncertified count as the meta word and bump
 *                      arrivals[rank] once in every region.  No staging copy, no push kernel. */
#define HCIR_PEER_MAX 16
typedef struct {
  const int32_t* labels;   /* [n_labels] class indices of the LOCAL gallery rows (nullable)         */
  int64_t n_labels;
  int32_t num_classes;
  float T;
  const int64_t* classes;  /* [num_classes] class values (nullable)                                */
  int64_t* pred;           /* [nq] (nullable)                                                      */
  int32_t* out_lab;        /* [nq,k] (nullable)                                                    */
  int32_t world, rank;     /* world == 0: single GPU, nothing below is read                        */
  int32_t payload;         /* 1 = packed block rows, 2 = predictions                               */
  int32_t reserved_;
  uint64_t slot_bytes;     /* the channel's slot size (hcir_peer_region_bytes)                      */
  void* regions[HCIR_PEER_MAX]; /* every rank's region as mapped HERE, own region included          */
  const int64_t* step;     /* the channel's completed-step counter (device)                        */
  /* completion launches (hcir_retry_setup below): */
  const int32_t* qmap;     /* launch row r answers ORIGINAL query qmap[r]; outputs go to that row   */
  const int32_t* active;   /* device count: only rows r < *active are live (the rest just count)    */
  int32_t commit_certified_only; /* write outputs only for queries this launch certifies            */
  int32_t no_signal;       /* store the peer rows but leave the arrival signal to a later launch    */
  int64_t out_rows;        /* rows of the output arrays / packed peer block (0 = nq)                */
} hcir_tail_t;

int hcir_select_rescore(const float* q_f32, const float* g_f32, int ld, int64_t nq, int64_t ng,
                        int k, int64_t idx_offset, const hcir_plan_t* plan, const void* workspace,
                        const float* q_delta, float g_delta_max, float eps_acc, float* out_sim,
                        int64_t* out_idx, int32_t* uncert_list, int32_t* uncert_state,
                        const hcir_tail_t* tail, hcir_stream_t stream);

/* Device-driven completion, step 1 of 3 (then hcir_simtopk_gated with HCIR_FLAG_MAIN_ONLY, then
 * hcir_select_rescore with tail.qmap / tail.active / tail.commit_certified_only): gather the queries a
 * first pass could not certify (uncert_list[0 .. uncert_state[1])) into a compact batch of at most
 * `capacity` rows and give each an explicit threshold -- its best-so-far fp32 k-th score minus eps:
 * every true top-k row scores above it in the bf16 contraction, so the second pass collects all that
 * can matter and certifies against that very threshold.  Replaces the host-driven gather / scatter of
 * round 1 (torch fancy indexing + two read-backs): the host looks at ONE count per step, after the
 * completion, and only near-duplicate galleries ever make it non-zero.
 *   q2_* / thr2 / qmap   the compact batch (rows >= the count: threshold +inf, nothing passes)
 *   active[0]            min(count, capacity): gates the two launches that follow
 *   final_list/final_state  queries beyond `capacity` go straight to the final uncertified list
 *                           (final_state is the uncert_state of the completion's hcir_select_rescore) */
int hcir_retry_setup(const uint16_t* q_bf16, const float* q_f32, const float* q_delta, int ld, int64_t nq,
                     int k, const float* out_sim, const int32_t* uncert_list,
                     const int32_t* uncert_state, float g_delta_max, float eps_acc, int capacity,
                     uint16_t* q2_bf16, float* q2_f32, float* q2_delta, float* thr0_2, float* thr_hi_2,
                     int32_t* qmap, int32_t* active, int32_t* final_list, int32_t* final_state,
                     hcir_stream_t stream);

/* Exact fp32 path (CUDA cores): brute-force canonical fp32 similarities + exact top-k in
 * canonical order for the queries listed in qlist[0..nlist) (qlist == NULL: queries
 * 0..nlist-1).  It is the GPU fallback for uncertified queries and the whole path when the
 * problem is too small for the tensor-core kernel.  Rows of out_* are indexed by query id.
 * Workspace: hcir_exact_workspace_bytes(nlist, ng, k). */
size_t hcir_exact_workspace_bytes(int64_t nlist, int64_t ng, int k, int sm_count);
int hcir_exact_topk(const float* q_f32, const float* g_f32, int ld, int64_t ng, int k,
                    int64_t idx_offset, const int32_t* qlist, int64_t nlist, float* out_sim,
                    int64_t* out_idx, void* workspace, size_t workspace_bytes, int sm_count,
                    hcir_stream_t stream);

/* Neighbour label gather: out[q][j] = labels[idx[q][j] - idx_offset]  (labels are class
 * indices 0..C-1, i.e. sklearn's `_y`, HairPretraining/src/classification_engine.py:81). */
int hcir_gather_labels(const int64_t* idx, int64_t count, const int32_t* labels, int64_t n_labels,
                       int64_t idx_offset, int32_t* out, hcir_stream_t stream);

/* K4 -- kNN vote.  T <= 0: uniform majority vote == sklearn `_mode(_y[neigh_ind])`
 * (classification_engine.py:82 -> neighbors/_classification.py:299-307), ties -> smallest
 * class.  T > 0: EXTENSION (not in the reference), score[c] = sum_j exp((s_j - s_0)/T) over
 * neighbours of class c in rank order, arg-max, ties -> smallest class.
 *   sims [nq,k] fp32, nbr_labels [nq,k] int32 in [0,num_classes)
 *   pred [nq] int32 class index; scores [nq,num_classes] fp32 (nullable). */
int hcir_vote(const float* sims, const int32_t* nbr_labels, int64_t nq, int k, int num_classes,
              float T, int32_t* pred, float* scores, hcir_stream_t stream);

/* K4, fused tail of the classification step: neighbour-label gather (idx are GLOBAL row indices,
 * labels the local [n_labels] class-index vector, idx_offset the first local row) -> vote -> class
 * value.  pred[q] = classes[best] (or the class index when classes is null), int64 like sklearn's
 * predict.  nbr_labels_out (nullable) receives the gathered [nq, k] neighbour labels.  Same
 * arithmetic and tie rule as hcir_gather_labels + hcir_vote. */
int hcir_vote_idx(const float* sims, const int64_t* idx, const int32_t* labels, int64_t n_labels,
                  int64_t idx_offset, int64_t nq, int k, int num_classes, float T,
                  const int64_t* classes, int64_t* pred, int32_t* nbr_labels_out,
                  hcir_stream_t stream);
/* the same vote on already gathered neighbour labels (multi-GPU: they travel with the candidates) */
int hcir_vote_classes(const float* sims, const int32_t* nbr_labels, int64_t nq, int k,
                      int num_classes, float T, const int64_t* classes, int64_t* pred,
                      hcir_stream_t stream);

/* K5 -- merge of per-shard exact top-k lists after the all-gather (new design, no reference
 * analogue).  gathered_* are [G][nq][k] with every [g][q][:] list in
 * canonical order; output the canonical top-k of the union.  Labels are optional. */
int hcir_merge_topk(const float* gathered_sim, const int64_t* gathered_idx,
                    const int32_t* gathered_lab, int G, int64_t nq, int k, float* out_sim,
                    int64_t* out_idx, int32_t* out_lab, hcir_stream_t stream);

/* K5, packed form: `gathered` is what ONE all-gather of every rank's packed result block produces,
 * G blocks of hcir_packed_block_bytes(nq, k, with_labels) bytes, each laid out as
 *   int64 idx[nq][k] | float sims[nq][k] | int32 labels[nq][k] (if with_labels)
 * and read in place (no unpacking pass).  rank_stride_bytes = distance between two ranks' blocks
 * (0 = exactly the block size; larger when every rank appends private trailer bytes, e.g. its
 * count of uncertified queries). */
size_t hcir_packed_block_bytes(int64_t nq, int k, int with_labels);
int hcir_merge_topk_packed(const void* gathered, int G, int64_t nq, int k, int with_labels,
                           size_t rank_stride_bytes, float* out_sim, int64_t* out_idx, int32_t* out_lab,
                           hcir_stream_t stream);

/* ---- candidate / result exchange over NVLink peer memory (new design, no reference analogue;
 * SURVEY.md section 8e "later fusion") ------------------------------------------------------
 * One REGION of device memory per rank, exported with CUDA IPC and mapped by every other rank of
 * the box; layout: 512-byte header (arrival counters, one int64 "meta" word per rank and parity,
 * last completed step, error word, done-CTA counters) + 2 parities x world slots of
 * round_up(slot_bytes, 256).  `step` is a device int64 per channel counting COMPLETED steps: the
 * producer of step st = *step + 1 fills slot (st & 1, rank) of every region and bumps arrivals[rank]
 * once everywhere; the consumer waits for arrivals[r] >= st for all r and writes *step = st.
 *   hcir_peer_alloc   cudaMalloc + zero + IPC export (64-byte handle) of this rank's region
 *   hcir_peer_open    map a peer's region from its handle;  hcir_peer_close / hcir_peer_free undo
 *   producer          normally the tail of hcir_select_rescore (every query's CTA stores its own
 *                     results into the peers); hcir_peer_push is the standalone form: store `bytes`
 *                     (multiple of 16) of a block that sits in memory into every region in
 *                     regions[0..world) (host array; own region included) + the int32 at meta_src
 *                     (nullable) as this rank's meta word, then signal arrival
 *   hcir_peer_wait    consumer, wait only: spin (bounded by timeout_ns; on expiry header word 49 is
 *                     set to the step) until every rank's block of the step has landed in the LOCAL
 *                     region, then complete the step
 *   hcir_peer_merge_vote   consumer of packed blocks, ONE kernel: wait as above, K5-merge the G
 *                     blocks in place, vote on the merged labels (pred nullable), complete the step */
size_t hcir_peer_region_bytes(int world, size_t slot_bytes);
size_t hcir_peer_slot_offset(int world, size_t slot_bytes, int parity, int rank);
int hcir_peer_push_ctas(size_t bytes);
int hcir_peer_alloc(size_t bytes, void** ptr, void* ipc_handle_64);
int hcir_peer_open(const void* ipc_handle_64, void** ptr);
int hcir_peer_close(void* ptr);
int hcir_peer_free(void* ptr);
int hcir_peer_push(const void* src, size_t bytes, void* const* regions, int world, int rank,
                   size_t slot_bytes, const int64_t* step, const int32_t* meta_src,
                   hcir_stream_t stream);
int hcir_peer_wait(void* region_local, int world, int64_t* step, int64_t timeout_ns,
                   hcir_stream_t stream);
int hcir_peer_merge_vote(void* region_local, int G, int64_t nq, int k, int with_labels,
                         size_t slot_bytes, int64_t* step, int64_t timeout_ns, float* out_sim,
                         int64_t* out_idx, int32_t* out_lab, int num_classes, float T,
                         const int64_t* classes, int64_t* pred, hcir_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HCIR_B200_H_ */
