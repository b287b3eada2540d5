/*
 * vfi.h — C ABI of the B200-native VeritasFi retrieval hot path.
 *
 * This is the drop-in boundary. Every entry point replaces one call the
 * reference makes into an un-vendored CPU library; citations are relative to /root/reference/.
 *
 *   vfi_index_*      replaces faiss.IndexFlatIP(d) / .add / .search        src/utils/faissRetriever.py:18,24,37
 *   vfi_normalize_l2 replaces faiss.normalize_L2(x)                         src/utils/faissRetriever.py:22,35
 *   vfi_bm25_*       replaces bm25s.BM25.load(...) / .retrieve(tokens, k)   src/utils/bm25Retriever.py:46,75-79
 *   vfi_fuse_union   replaces the shared seen_ids ordered de-dup union      src/utils/ensembleRetriever.py:58,72-74,148-150,194-196
 *   vfi_fuse_rrf     reciprocal-rank fusion (north_star; not in the reference, SURVEY.md finding 4)
 *   vfi_fuse_hybrid  the fusion step of the sharded hybrid retriever in one launch: title -> chunk mapping, per-path
 *                    de-duplication, dropping of BM25's zero-score filler, RRF (north_star configs[3])
 *   vfi_merge_topk   global top-k after the all-gather of per-shard results (no reference counterpart;
 *                    the reference only replicates workers, experiments/retriever/step3_mul.py:405-446)
 *   vfi_cosine_topk  replaces cosine_similarity + argsort top-k             experiments/retriever/step3_mul.py:255-289,
 *                                                                           experiments/retriever/continuous_retrieval.py:154-167
 *
 *   vfi_stem_english / vfi_tokenize_ascii  replace Stemmer.Stemmer('english').stemWords and the splitting step of
 *                    bm25s.tokenize (host-side text routines, no device needed)   src/utils/bm25Retriever.py:14-15,47,67
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types. `mem` says where caller buffers live.
 *   - the library never returns pointers into its own memory; the caller owns every output buffer.
 *   - no function throws; each returns a vfi_status and vfi_last_error() holds a thread-local message.
 *   - ids are row positions (the reference's row-id space, ensembleRetriever.py:45-48) plus the
 *     index's id offset (used by corpus shards to emit global ids).
 *   - result order is the total order (score descending, id ascending); rows short of k are
 *     padded with id -1 / score -FLT_MAX like faiss.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - there is NO CPU implementation behind this ABI: without a CUDA device every compute entry
 *     fails with VFI_ERR_NO_DEVICE.
 */
#ifndef VFI_H_
#define VFI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFI_ABI_VERSION 1

typedef enum vfi_status {
  VFI_OK = 0,
  VFI_ERR_INVALID = 1,     /* bad argument (dimension mismatch, k <= 0, null pointer ...) */
  VFI_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed; message has the detail */
  VFI_ERR_NOMEM = 3,       /* device or host allocation failed */
  VFI_ERR_UNSUPPORTED = 4, /* outside the supported envelope (e.g. k > VFI_MAX_K) */
  VFI_ERR_NO_DEVICE = 5,   /* no CUDA device / not an sm_100 device */
  VFI_ERR_INTERNAL = 6
} vfi_status;

/* where a caller buffer lives */
enum { VFI_MEM_HOST = 0, VFI_MEM_DEVICE = 1 };

/* how corpus rows are stored (and therefore what "the corpus" is for parity purposes)
 *   VFI_STORE_BF16: rows and queries are rounded to bf16 (round-to-nearest-even) on entry; the
 *                   canonical score is the exact dot product of those bf16 values.
 *   VFI_STORE_F32 : rows and queries keep their fp32 values (faiss.IndexFlatIP semantics); the
 *                   tensor-core pass runs on a 3-term bf16 split, the exact pass on the fp32 rows. */
enum { VFI_STORE_BF16 = 0, VFI_STORE_F32 = 1 };

/* element type of a query buffer (vfi_index_search_ex): fp32, or raw bf16 bit patterns (half the H2D bytes; exactly what a
 * VFI_STORE_BF16 index would round the queries to anyway) */
enum { VFI_DTYPE_F32 = 0, VFI_DTYPE_BF16 = 1 };

#define VFI_MAX_K 2048

typedef struct vfi_index vfi_index_t;
typedef struct vfi_bm25 vfi_bm25_t;

/* ---- library ---------------------------------------------------------------------------- */
int vfi_abi_version(void);
const char* vfi_last_error(void);
/* number of usable CUDA devices (0 when there is none); never fails */
int vfi_device_count(void);
/* kernels launched by this library since load (all threads); for bench.py's gpu_launches */
int64_t vfi_launch_count(void);

/* ---- dense flat inner-product index (faiss.IndexFlatIP) ---------------------------------- */
int vfi_index_create(int d, int store_dtype, int device, vfi_index_t** out);
int vfi_index_destroy(vfi_index_t* idx);
/* reserve device capacity for n rows in total (optional; add() grows geometrically otherwise) */
int vfi_index_reserve(vfi_index_t* idx, int64_t n);
/* append n fp32 rows [n,d] row-major; copies (the caller's array may be freed afterwards) */
int vfi_index_add(vfi_index_t* idx, const float* x, int64_t n, int mem, void* stream);
/* append n bf16 rows (raw 16-bit patterns); only for VFI_STORE_BF16 indexes */
int vfi_index_add_bf16(vfi_index_t* idx, const uint16_t* x, int64_t n, int mem, void* stream);
int64_t vfi_index_ntotal(const vfi_index_t* idx);
int vfi_index_dim(const vfi_index_t* idx);
/* ids reported by search = row position + offset (corpus shards report global ids) */
int vfi_index_set_id_offset(vfi_index_t* idx, int64_t offset);
/* copy row i (as fp32, the stored value) to out[d] */
int vfi_index_reconstruct(vfi_index_t* idx, int64_t i, float* out, int mem);
/* copy rows [first, first+n) as fp32 stored values to out[n,d] (bulk reconstruct; persisting a shard) */
int vfi_index_read_rows(vfi_index_t* idx, int64_t first, int64_t n, float* out, int mem, void* stream);
/* canonical scores between stored rows: out[i*n+j] = <row ids[i], row ids[j]> (fp64 sequential, rounded to fp32).
 * With L2-normalised rows this is the cosine matrix of ensembleRetriever.py:265-281 without re-embedding. */
int vfi_index_pairwise(vfi_index_t* idx, const int64_t* ids, int n, float* out, int mem, void* stream);
/* exact top-k of q·xᵀ. q: fp32 [nq,d]; out_scores fp32 [nq,k]; out_ids int64 [nq,k].
 * Blocks until the results are in the caller's buffers when mem == VFI_MEM_HOST; with
 * VFI_MEM_DEVICE the call returns after the device work is enqueued and verified. */
int vfi_index_search(vfi_index_t* idx, const float* q, int64_t nq, int k, float* out_scores,
                     int64_t* out_ids, int mem, void* stream);
/* the same with the query element type given (q: [nq,d] of q_dtype) */
int vfi_index_search_ex(vfi_index_t* idx, const void* q, int q_dtype, int64_t nq, int k, float* out_scores,
                        int64_t* out_ids, int mem, void* stream);

/* Threading: vfi_index_search, vfi_index_search_begin/finish, vfi_index_pairwise, vfi_index_read_rows and
 * vfi_index_reconstruct may be called from any number of threads on one index at the same time (the reference calls
 * retriever.invoke from concurrent request threads without a lock, src/utils/vllmChatService.py:85-88,404): every call
 * works in a workspace of its own taken from a pool, and a host-buffer search that passes no stream runs on a stream of
 * its own.  add / reserve / set_option / set_id_offset / destroy must not run concurrently with a search (as with faiss).
 *
 * Pipelined form for device buffers: begin() enqueues one batch (nq <= 1024) on `stream` and returns at once with a
 * ticket; finish(ticket) waits for that batch, checks its exactness certificate and repairs the rare query that failed it,
 * after which out_scores / out_ids hold the same bits vfi_index_search would have produced.  Up to 64 batches may be in
 * flight per index; q and the output buffers must stay valid until finish.  A serving loop that begins batch i+1
 * before finishing batch i never leaves the GPU idle during the host's look at the certificate flag. */
int vfi_index_search_begin(vfi_index_t* idx, const float* q, int64_t nq, int k, float* out_scores,
                           int64_t* out_ids, void* stream, int* ticket);
int vfi_index_search_begin_ex(vfi_index_t* idx, const void* q, int q_dtype, int64_t nq, int k, float* out_scores,
                              int64_t* out_ids, void* stream, int* ticket);
int vfi_index_search_finish(vfi_index_t* idx, int ticket);
/* device address of the batch's certificate counter (> 0 once its kernels ran = some query still needs the repair of
 * finish()); valid until finish(ticket).  Lets a kernel enqueued behind the batch (vfi_exchange_merge_flagged) tell its
 * peers that the rows it is about to send are not final. */
int vfi_index_ticket_flag(vfi_index_t* idx, int ticket, const int** device_flag);

/* tuning / introspection ------------------------------------------------------------------ */
enum {
  VFI_OPT_OVERFETCH = 1,     /* candidates kept per query by the tensor-core pass (0 = auto) */
  VFI_OPT_FORCE_PATH = 2,    /* 0 auto, 1 exact streaming scorer (every row scored canonically; deep k for few queries, tiny
                                shards, repairs), 2 fused tcgen05 GEMM + top-k', 3 streaming GEMV (few queries, k' <= 512) */
  VFI_OPT_PROFILE = 3,       /* 1: bracket the dominant kernel with CUDA events */
  VFI_OPT_TAU_HINT = 4,      /* 1 (default): estimate a per-query admission threshold from a row sample (chunk maxima); 0: off;
                                2: debug, admit nothing; 3: as 1 from every sampled score */
  VFI_OPT_NUM_CTAS = 5,      /* 0 = one CTA per SM */
  VFI_OPT_CTA_PAIR = 7,      /* tcgen05 cta_group::2 kernel (two SMs share every corpus tile; a batch with an odd number of
                                128-query tiles gets a padding tile): 0 auto = on, 1 off (single-CTA kernel), 2 on */
  VFI_OPT_SMALL_BATCH = 10,  /* 0 auto: batches of 9..64 queries run the swapped-operand tcgen05 kernel (corpus rows on the M side,
                                the query block resident in shared memory: the corpus stream is the only traffic); 1 off */
  VFI_OPT_TAIL_PIECE = 11,   /* bytes per lane and step of the rescoring kernel's row gather: 0 auto, 128, 256 */
  VFI_OPT_TAU_M = 9          /* admission hint = the m-th best score of a row sample: 0 auto (8 for large shards; 16 or 32 with a
                                denser sample when k'/N is large and passing rows would swamp the epilogue), or 8 / 16 / 32 */
};
int vfi_index_set_option(vfi_index_t* idx, int opt, int64_t value);

typedef struct vfi_search_stats {
  int64_t searches;            /* search calls */
  int64_t queries;             /* queries processed */
  int64_t retried_queries;     /* queries whose certificate failed and were re-run by the exact streaming scorer */
  int64_t fused_launches;      /* launches of the tcgen05 kernel */
  double fused_ms_total;       /* summed device time of those launches (VFI_OPT_PROFILE=1) */
  int64_t fused_ms_samples;    /* launches that contributed to fused_ms_total */
  int last_path;               /* path taken by the last search (VFI_OPT_FORCE_PATH values) */
  int last_overfetch;          /* k' used by the last search */
  float last_eps;              /* largest certificate epsilon of the last search */
  float max_abs_err;           /* max |tensor-core score - exact score| over rescored candidates */
  int64_t hint_retries;        /* batches redone without the admission hint */
  double tail_ms_total;        /* summed device time of the selection + rescoring + certificate tail (VFI_OPT_PROFILE=1) */
  int64_t tail_ms_samples;     /* searches that contributed to tail_ms_total */
} vfi_search_stats;
int vfi_index_get_stats(vfi_index_t* idx, vfi_search_stats* out, int reset);

/* debug/test hook: raw tensor-core scores S[nq, n] of the first n rows (device or host buffer) */
int vfi_index_debug_scores(vfi_index_t* idx, const float* q, int64_t nq, float* out, int mem,
                           void* stream);

/* in-place row-wise L2 normalisation in fp32, zero rows untouched (faiss.normalize_L2) */
int vfi_normalize_l2(float* x, int64_t n, int d, int mem, int device, void* stream);

/* exact cosine top-k between two small fp32 matrices (the experiments' select_top_chunks):
 * tie order "higher index first" as produced by np.argsort(sim)[-k:][::-1] */
int vfi_cosine_topk(const float* e, int64_t n_e, const float* c, int64_t n_c, int d, int k,
                    float* out_scores, int64_t* out_ids, int mem, int device, void* stream);

/* ---- multi-GPU merge --------------------------------------------------------------------- */
/* scores fp32 [g, nq, k_in], ids int64 [g, nq, k_in] (id -1 = padding) -> top k_out per query */
int vfi_merge_topk(const float* scores, const int64_t* ids, int g, int64_t nq, int k_in,
                   int k_out, float* out_scores, int64_t* out_ids, int mem, int device,
                   void* stream);

/* ---- multi-GPU exchange + merge over peer memory (one process per GPU) ----------------------- */
/* The fused form of "all-gather the per-shard results, then vfi_merge_topk": every rank owns a receive
 * window in its HBM that its peers map through CUDA IPC; ONE kernel per rank stores the rank's [nq,k]
 * results straight into every peer's window over NVLink, publishes per-query flags, waits for the
 * peers' flags and selects the global top-k_out.  No reference counterpart (workers are replicas,
 * experiments/retriever/step3_mul.py:405-446).  Usage, identically on every rank:
 *   vfi_exchange_create -> vfi_exchange_handle -> (host all-gathers the 64-byte handles, any transport)
 *   -> vfi_exchange_connect -> (host barrier) -> vfi_exchange_merge ... -> (host barrier) -> destroy.
 * vfi_exchange_merge is a collective: all ranks call it the same number of times with the same nq, k.
 * ids are global and must be < 2^32 - 1 (as for vfi_merge_topk); -1 = padding.  All buffers are device
 * memory.  A peer that does not arrive within the timeout (default 30 s) traps the kernel (sticky CUDA
 * error) instead of hanging the GPU.  world <= 16.  Never run two ranks on one GPU. */
typedef struct vfi_exchange vfi_exchange_t;
#define VFI_IPC_HANDLE_BYTES 64
int vfi_exchange_create(int device, int rank, int world, int64_t max_nq, int max_k, vfi_exchange_t** out);
int vfi_exchange_handle(vfi_exchange_t* ex, void* out_handle /* VFI_IPC_HANDLE_BYTES */);
int vfi_exchange_connect(vfi_exchange_t* ex, const void* handles /* world x VFI_IPC_HANDLE_BYTES, rank order */);
int vfi_exchange_set_timeout_ms(vfi_exchange_t* ex, int64_t ms);
int vfi_exchange_merge(vfi_exchange_t* ex, const float* scores, const int64_t* ids, int64_t nq, int k, int k_out,
                       float* out_scores, int64_t* out_ids, void* stream);
/* The same with the batch's state attached: fail_a / fail_b (device ints, either may be NULL; see vfi_index_ticket_flag) say
 * whether this rank's rows may still be repaired; *any_fail (device int, may be NULL; zero it before the call) is set to 1
 * on EVERY rank when any rank said so — the ranks then agree, without a host collective, to exchange the batch again
 * after their repairs. */
int vfi_exchange_merge_flagged(vfi_exchange_t* ex, const float* scores, const int64_t* ids, int64_t nq, int k, int k_out,
                               float* out_scores, int64_t* out_ids, const int* fail_a, const int* fail_b, int* any_fail,
                               void* stream);
/* Compute + collective (sharded search): vfi_index_search_begin_push is vfi_index_search_begin_ex whose rescoring kernel also
 * SENDS — the CTA that has just ordered and certified a query stores that row (global ids) into every rank's window over NVLink
 * while the other queries are still being rescored, without waiting for the round trip (a batch on a path without that kernel
 * is pushed by a kernel of its own behind the search: every rank consumes the epoch).  vfi_exchange_merge_pushed, enqueued behind
 * it on the same stream, publishes the rows (fail_flag: vfi_index_ticket_flag of the batch, may be NULL = rows are final), waits
 * for the peers' rows of that epoch and merges them; *slot names the host word that holds, once an event recorded behind the
 * call has completed, whether ANY rank's rows were not final (vfi_exchange_any_fail) — no copy operation in the stream.  Both
 * calls are collectives like vfi_exchange_merge. */
int vfi_index_search_begin_push(vfi_index_t* idx, const void* q, int q_dtype, int64_t nq, int k, float* out_scores,
                                int64_t* out_ids, vfi_exchange_t* ex, void* stream, int* ticket);
int vfi_exchange_merge_pushed(vfi_exchange_t* ex, int64_t nq, int k, int k_out, float* out_scores, int64_t* out_ids,
                              const int* fail_flag, int* slot, void* stream);
int vfi_exchange_any_fail(vfi_exchange_t* ex, int slot, int* any_fail);
int vfi_exchange_destroy(vfi_exchange_t* ex);

/* ---- BM25 over token-major postings (bm25s CSC arrays) ----------------------------------- */
/* indptr int64 [n_vocab+1], indices int32 [nnz] (doc ids ascending per token), data fp32 [nnz]
 * (precomputed impacts). doc ids are local to this shard in [0, n_docs); reported ids add
 * id_offset. Host pointers; the arrays are copied to the device. */
int vfi_bm25_create(const int64_t* indptr, const int32_t* indices, const float* data,
                    int64_t n_vocab, int64_t n_docs, int64_t id_offset, int device,
                    vfi_bm25_t** out);
/* the same with `mem` saying where the three arrays live (VFI_MEM_DEVICE: postings built on the GPU are adopted by a
 * device-to-device copy; validation runs on the device either way) */
int vfi_bm25_create_from(const int64_t* indptr, const int32_t* indices, const float* data,
                         int64_t n_vocab, int64_t n_docs, int64_t id_offset, int mem, int device,
                         vfi_bm25_t** out);
int vfi_bm25_destroy(vfi_bm25_t* b);
int64_t vfi_bm25_ndocs(const vfi_bm25_t* b);
/* q_tokens int32 [q_indptr[nq]] token ids in query order (unknown tokens already dropped,
 * repeats kept; any number of tokens per query), q_indptr int64 [nq+1]: HOST arrays (they come from the host
 * tokeniser). Scores accumulate in fp32 in query-token order. k <= VFI_MAX_K. `mem` says where out_scores / out_ids
 * live (VFI_MEM_HOST or VFI_MEM_DEVICE).  Thread-safe (per-call scratch from a pool). */
int vfi_bm25_search(vfi_bm25_t* b, const int32_t* q_tokens, const int64_t* q_indptr, int64_t nq,
                    int k, float* out_scores, int64_t* out_ids, int mem, void* stream);
/* all n_docs scores of one query (host/device fp32 [n_docs]); serves retrieve(k = N) */
int vfi_bm25_score_all(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, float* out,
                       int mem, void* stream);
/* every doc ranked: out_ids int64 [n_docs], out_scores fp32 [n_docs] in (score desc, id asc)
 * order — bm25s retrieve(k = N) as called at ensembleRetriever.py:189. Host outputs. */
int vfi_bm25_rank_all(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, float* out_scores,
                      int64_t* out_ids, void* stream);
/* ranks [first, first+count) of that order only (the tail behind an eager top-k, produced when a caller asks for it) */
int vfi_bm25_rank_range(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, int64_t first, int64_t count,
                        float* out_scores, int64_t* out_ids, void* stream);
typedef struct vfi_bm25_stats {
  int64_t launches;
  double score_ms_total;   /* device time of the scoring kernel (profile on) */
  int64_t score_ms_samples;
  int64_t postings_bytes;  /* algorithmic bytes of the last search: sum over query tokens df*8 */
} vfi_bm25_stats;
int vfi_bm25_set_profile(vfi_bm25_t* b, int on);
int vfi_bm25_get_stats(vfi_bm25_t* b, vfi_bm25_stats* out, int reset);

/* ---- rank fusion -------------------------------------------------------------------------- */
/* ids int64 [nq, n_paths, depth] ranked lists (-1 = padding, rank = position+1).
 * fused(d) = sum over paths p=0.. of 1/(k_rrf + rank_p(d)) in fp32, in path order.
 * out: top k by (fused desc, id asc). */
int vfi_fuse_rrf(const int64_t* ids, int64_t nq, int n_paths, int depth, float k_rrf, int k,
                 float* out_scores, int64_t* out_ids, int mem, int device, void* stream);
/* priority-ordered de-duplicated union (path 0 first, rank order inside a path; first
 * occurrence wins). out_ids/out_scores/out_path [nq, n_paths*depth] padded with -1;
 * out_count int32 [nq]. */
int vfi_fuse_union(const int64_t* ids, const float* scores, int64_t nq, int n_paths, int depth,
                   int64_t* out_ids, float* out_scores, int32_t* out_path, int32_t* out_count,
                   int mem, int device, void* stream);

/* The fusion step of the hybrid retriever in one launch.  Lists are read in place as
 * ids[p * path_stride + q * query_stride + r] (p < n_paths, r < depth; -1 = padding), device memory:
 *   path `title_path` (or -1): ids are rows of the title corpus, replaced by title_to_chunk[id] (device int64 [n_titles]);
 *        several titles may stand for one chunk — the first occurrence counts and the ranks behind it close up
 *        (the id-level form of src/utils/ensembleRetriever.py:143-150);
 *   path `sparse_path` (or -1): entries with score <= 0 are dropped (bm25s pads a query that matched fewer than `depth`
 *        docs with zero-score docs; they must not earn 1/(k_rrf + rank));
 *   then fused(d) = sum over paths in path order of 1/(k_rrf + rank_p(d)), top k by (fused desc, id asc).
 * n_paths <= 8, n_paths * depth <= 4096. */
int vfi_fuse_hybrid(const int64_t* ids, const float* scores, int64_t nq, int n_paths, int depth, int64_t path_stride,
                    int64_t query_stride, const int64_t* title_to_chunk, int64_t n_titles, int title_path,
                    int sparse_path, float k_rrf, int k, float* out_scores, int64_t* out_ids, int device, void* stream);

/* ---- query/document text (host only; usable without a CUDA device) --------------------------- */
/* Snowball "english" (Porter2) stemmer, UTF-8.  words: the concatenated bytes of n_words words, word i =
 * words[offsets[i] .. offsets[i+1]).  out receives the concatenated stems (never longer than the input, so
 * out_cap >= offsets[n_words] always suffices) and out_offsets [n_words+1] their boundaries.
 * Replaces Stemmer.Stemmer('english').stemWords(list) (src/utils/bm25Retriever.py:14,47). */
int vfi_stem_english(const char* words, const int64_t* offsets, int64_t n_words, char* out, int64_t out_cap,
                     int64_t* out_offsets);
/* bm25s.tokenize's splitting r"(?u)\b\w\w+\b" for ASCII text: tokens are the maximal runs of [0-9A-Za-z_] of
 * length >= 2 (byte offsets into text).  *n_tokens = tokens found; at most cap are written.  Text holding a byte
 * >= 0x80 is refused with VFI_ERR_UNSUPPORTED (the host facade then uses its Unicode-aware pattern). */
int vfi_tokenize_ascii(const char* text, int64_t len, int64_t* starts, int64_t* lens, int64_t cap, int64_t* n_tokens);

#ifdef __cplusplus
}
#endif
#endif /* VFI_H_ */
