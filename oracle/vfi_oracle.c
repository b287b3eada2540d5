/*
 * vfi_oracle.c — CPU restatement of the VeritasFi retrieval hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  Nothing under veritasfi_b200/ imports it; the product path has
 * no CPU fallback.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path, and the
 * arithmetic lives in un-vendored, un-pinned third-party libraries (faiss, bm25s, PyStemmer) that
 * are absent from /root/reference and from this image.  What is restated here is
 *   - the reference's own call sites (cited per function, paths relative to /root/reference/), and
 *   - the published algorithms of those libraries (faiss IndexFlatIP / normalize_L2; bm25s
 *     "lucene" scoring with np.add.at accumulation), marked [upstream] where taken from memory.
 * The known-answer tests in tests/test_oracle_kat.py are hand-computed in this repo.
 *
 * The canonical contract: result order is (score descending, id ascending);
 * a dense score is the sequential fp64 fused sum over j = 0..d-1 rounded once to fp32.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- helpers */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* order-preserving map used for the (score desc, id asc) total order */
static inline uint32_t orderable(float f) {
  uint32_t u = f2u(f + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
typedef struct { float s; int64_t id; } pair_t;
static int pair_cmp(const void* a, const void* b) {
  const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
  uint32_t ox = orderable(x->s), oy = orderable(y->s);
  if (ox != oy) return ox > oy ? -1 : 1;           /* higher score first */
  if (x->id != y->id) return x->id < y->id ? -1 : 1; /* lower id first */
  return 0;
}

/* fp32 -> bf16 -> fp32, round to nearest even (what VFI_STORE_BF16 defines as "the corpus") */
void vfo_bf16_round(float* x, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u = f2u(x[i]);
    if ((u & 0x7F800000u) == 0x7F800000u) { x[i] = u2f(u & 0xFFFF0000u); continue; }
    uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
    x[i] = u2f((u + r) & 0xFFFF0000u);
  }
}

/* faiss.normalize_L2(x) as called at src/utils/faissRetriever.py:22,35 [upstream]:
 * x[i,:] *= 1/sqrtf(sum_j x[i,j]^2), rows of norm 0 untouched.  The summation order is pinned
 * here (faiss leaves it to its SIMD loops): 32 strided fp64 partials, partial l summing
 * j = l, l+32, ... in order, combined by the butterfly l^16, l^8, l^4, l^2, l^1; the total is
 * rounded to fp32 before the square root. */
void vfo_normalize_l2(float* x, int64_t n, int d) {
  for (int64_t i = 0; i < n; ++i) {
    float* r = x + i * d;
    double p[32];
    for (int l = 0; l < 32; ++l) {
      double s = 0.0;
      for (int j = l; j < d; j += 32) s += (double)r[j] * (double)r[j];  /* product exact in fp64 == fma */
      p[l] = s;
    }
    for (int o = 16; o > 0; o >>= 1) {
      double t[32];
      for (int l = 0; l < 32; ++l) t[l] = p[l] + p[l ^ o];
      memcpy(p, t, sizeof(p));
    }
    float nrm2 = (float)p[0];
    if (nrm2 > 0.f) {
      float inv = 1.0f / sqrtf(nrm2);
      for (int j = 0; j < d; ++j) r[j] = r[j] * inv;
    }
  }
}

/* the canonical score */
static inline float canon_dot(const float* q, const float* x, int d) {
  double acc = 0.0;
  for (int j = 0; j < d; ++j) acc += (double)q[j] * (double)x[j];  /* product exact in fp64 == fma */
  return (float)acc;
}
float vfo_canon_dot(const float* q, const float* x, int d) { return canon_dot(q, x, d); }

/* canonical scores of listed rows: out[i] = dot(q, xb[ids[i]]) ; ids < 0 -> -FLT_MAX */
void vfo_rescore(const float* q, const float* xb, int d, const int64_t* ids, int64_t n_ids, float* out) {
  for (int64_t i = 0; i < n_ids; ++i)
    out[i] = ids[i] >= 0 ? canon_dot(q, xb + ids[i] * (int64_t)d, d) : -FLT_MAX;
}

/* top-k of (scores[i], id_base + i) under the total order; pads with -1 / -FLT_MAX */
void vfo_topk(const float* scores, int64_t n, int k, int64_t id_base, float* out_scores, int64_t* out_ids) {
  pair_t* p = (pair_t*)malloc(sizeof(pair_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) { p[i].s = scores[i]; p[i].id = id_base + i; }
  qsort(p, (size_t)n, sizeof(pair_t), pair_cmp);
  for (int i = 0; i < k; ++i) {
    if (i < n) { out_scores[i] = p[i].s + 0.0f; out_ids[i] = p[i].id; }
    else { out_scores[i] = -FLT_MAX; out_ids[i] = -1; }
  }
  free(p);
}

/* same, over explicit (score, id) pairs; id < 0 entries are padding and ignored */
void vfo_topk_pairs(const float* scores, const int64_t* ids, int64_t n, int k, float* out_scores, int64_t* out_ids) {
  pair_t* p = (pair_t*)malloc(sizeof(pair_t) * (size_t)(n > 0 ? n : 1));
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) if (ids[i] >= 0) { p[m].s = scores[i]; p[m].id = ids[i]; ++m; }
  qsort(p, (size_t)m, sizeof(pair_t), pair_cmp);
  for (int i = 0; i < k; ++i) {
    if (i < m) { out_scores[i] = p[i].s + 0.0f; out_ids[i] = p[i].id; }
    else { out_scores[i] = -FLT_MAX; out_ids[i] = -1; }
  }
  free(p);
}

/* faiss.IndexFlatIP.search(x, k) as called at src/utils/faissRetriever.py:37 [upstream], with the
 * canonical score and the fixed tie-break: exhaustive, exact.  O(nq*n*d); for small cases. */
void vfo_flat_search(const float* xq, int64_t nq, const float* xb, int64_t n, int d, int k, int64_t id_base,
                     float* out_scores, int64_t* out_ids) {
  float* s = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  for (int64_t q = 0; q < nq; ++q) {
    for (int64_t i = 0; i < n; ++i) s[i] = canon_dot(xq + q * (int64_t)d, xb + i * (int64_t)d, d);
    vfo_topk(s, n, k, id_base, out_scores + q * (int64_t)k, out_ids + q * (int64_t)k);
  }
  free(s);
}

/* bm25s.BM25.retrieve scoring as called at src/utils/bm25Retriever.py:75-79 [upstream]:
 * scores = zeros(N, fp32); for each query token in order: np.add.at(scores, indices[s:e], data[s:e]).
 * Token ids outside [0, n_vocab) are skipped (bm25s drops unknown tokens). */
void vfo_bm25_scores(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_vocab,
                     const int32_t* tokens, int64_t n_tokens, int64_t n_docs, float* scores) {
  for (int64_t i = 0; i < n_docs; ++i) scores[i] = 0.f;
  for (int64_t t = 0; t < n_tokens; ++t) {
    int32_t tok = tokens[t];
    if (tok < 0 || tok >= n_vocab) continue;
    for (int64_t p = indptr[tok]; p < indptr[tok + 1]; ++p) {
      volatile float v = scores[indices[p]] + data[p];  /* fp32 add, no contraction */
      scores[indices[p]] = v;
    }
  }
}

/* reciprocal-rank fusion (north_star; Cormack et al. 2009; not in the reference):
 * ids [n_paths][depth] (-1 padding); fused(d) = sum over entries of d, in (path, rank) order, of
 * 1/(k_rrf + rank) in fp32, rank 1-based.  Output top-k by (fused desc, id asc). */
void vfo_rrf(const int64_t* ids, int n_paths, int depth, float k_rrf, int k, float* out_scores, int64_t* out_ids) {
  int n = n_paths * depth;
  int64_t* uid = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  float* us = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int m = 0;
  for (int e = 0; e < n; ++e) {
    int64_t id = ids[e];
    if (id < 0) continue;
    volatile float denom = k_rrf + (float)(e % depth + 1);
    volatile float term = 1.0f / denom;
    int j = 0;
    while (j < m && uid[j] != id) ++j;
    if (j == m) { uid[m] = id; us[m] = 0.f; ++m; }
    volatile float acc = us[j] + term;
    us[j] = acc;
  }
  vfo_topk_pairs(us, uid, m, k, out_scores, out_ids);
  free(uid); free(us);
}

/* the reference's fusion: priority-ordered de-duplicated union through one shared seen_ids set
 * (src/utils/ensembleRetriever.py:58,72-74,148-150,194-196), at the id level.
 * out arrays have n_paths*depth slots padded with -1; returns the number kept. */
int vfo_union(const int64_t* ids, const float* scores, int n_paths, int depth, int64_t* out_ids, float* out_scores,
              int32_t* out_path) {
  int n = n_paths * depth, m = 0;
  for (int e = 0; e < n; ++e) {
    int64_t id = ids[e];
    if (id < 0) continue;
    int seen = 0;
    for (int j = 0; j < m; ++j) if (out_ids[j] == id) { seen = 1; break; }
    if (seen) continue;
    out_ids[m] = id; out_scores[m] = scores[e]; out_path[m] = e / depth; ++m;
  }
  for (int j = m; j < n; ++j) { out_ids[j] = -1; out_scores[j] = -FLT_MAX; out_path[j] = -1; }
  return m;
}
