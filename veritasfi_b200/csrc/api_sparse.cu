// api_sparse.cu — the sparse and fusion half of include/vfi.h: vfi_bm25_* (bm25s.BM25.retrieve), vfi_merge_topk,
// vfi_fuse_rrf / vfi_fuse_union / vfi_fuse_hybrid.  Host orchestration of the kernels in sparse_fuse.cuh (K3, K4, K5),
// reduce.cuh (K1c, K2b) and radix_sort.cuh.
#include <cstdlib>

#include "api_common.h"
#include "radix_sort.cuh"
#include "reduce.cuh"
#include "sparse_fuse.cuh"

using namespace vfi_host;

extern "C" {

int vfi_merge_topk(const float* scores, const int64_t* ids, int g, int64_t nq, int k_in, int k_out, float* out_scores,
                   int64_t* out_ids, int mem, int device, void* stream) {
  if (!scores || !ids || !out_scores || !out_ids || g <= 0 || nq < 0 || k_in <= 0 || k_out <= 0)
    return fail(VFI_ERR_INVALID, "bad argument to vfi_merge_topk");
  if (k_out > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k_out exceeds VFI_MAX_K");
  if (nq == 0) return VFI_OK;
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Staged sg(device);
  const size_t n_in = static_cast<size_t>(g) * nq * k_in, n_out = static_cast<size_t>(nq) * k_out;
  const float* ds;
  const int64_t* di;
  float* os;
  int64_t* oi;
  sg.in(scores, n_in, mem, &ds);
  sg.in(ids, n_in, mem, &di);
  sg.out(out_scores, n_out, mem, &os);
  sg.out(out_ids, n_out, mem, &oi);
  VFI_TRY(sg.commit(st));
  vfi::merge_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(ds, di, g, nq, k_in, k_out, os, oi);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return sg.back(st);
}

int vfi_fuse_rrf(const int64_t* ids, int64_t nq, int n_paths, int depth, float k_rrf, int k, float* out_scores,
                 int64_t* out_ids, int mem, int device, void* stream) {
  if (!ids || !out_scores || !out_ids || nq < 0 || n_paths <= 0 || depth <= 0 || k <= 0)
    return fail(VFI_ERR_INVALID, "bad argument to vfi_fuse_rrf");
  if (static_cast<int64_t>(n_paths) * depth > vfi::kSortCap) return fail(VFI_ERR_UNSUPPORTED, "n_paths*depth exceeds 4096");
  if (nq == 0) return VFI_OK;
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Staged sg(device);
  const size_t n_in = static_cast<size_t>(nq) * n_paths * depth, n_out = static_cast<size_t>(nq) * k;
  const int64_t* di;
  float* os;
  int64_t* oi;
  sg.in(ids, n_in, mem, &di);
  sg.out(out_scores, n_out, mem, &os);
  sg.out(out_ids, n_out, mem, &oi);
  VFI_TRY(sg.commit(st));
  vfi::rrf_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(di, n_paths, depth, k_rrf, k, os, oi);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return sg.back(st);
}

int vfi_fuse_union(const int64_t* ids, const float* scores, int64_t nq, int n_paths, int depth, int64_t* out_ids,
                   float* out_scores, int32_t* out_path, int32_t* out_count, int mem, int device, void* stream) {
  if (!ids || !scores || !out_ids || !out_scores || !out_path || !out_count || nq < 0 || n_paths <= 0 || depth <= 0)
    return fail(VFI_ERR_INVALID, "bad argument to vfi_fuse_union");
  if (static_cast<int64_t>(n_paths) * depth > vfi::kSortCap) return fail(VFI_ERR_UNSUPPORTED, "n_paths*depth exceeds 4096");
  if (nq == 0) return VFI_OK;
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Staged sg(device);
  const size_t n_in = static_cast<size_t>(nq) * n_paths * depth;
  const int64_t* di;
  const float* ds;
  int64_t* oi;
  float* os;
  int32_t *op, *oc;
  sg.in(ids, n_in, mem, &di);
  sg.in(scores, n_in, mem, &ds);
  sg.out(out_ids, n_in, mem, &oi);
  sg.out(out_scores, n_in, mem, &os);
  sg.out(out_path, n_in, mem, &op);
  sg.out(out_count, static_cast<size_t>(nq), mem, &oc);
  VFI_TRY(sg.commit(st));
  vfi::union_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(di, ds, n_paths, depth, oi, os, op, oc);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return sg.back(st);
}

int vfi_fuse_hybrid(const int64_t* ids, const float* scores, int64_t nq, int n_paths, int depth, int64_t path_stride,
                    int64_t query_stride, const int64_t* title_to_chunk, int64_t n_titles, int title_path, int sparse_path, float k_rrf, int k,
                    float* out_scores, int64_t* out_ids, int device, void* stream) {
  if (!ids || !scores || !out_scores || !out_ids || nq < 0 || n_paths <= 0 || n_paths > 8 || depth <= 0 || k <= 0)
    return fail(VFI_ERR_INVALID, "bad argument to vfi_fuse_hybrid");
  if (static_cast<int64_t>(n_paths) * depth > vfi::kSortCap) return fail(VFI_ERR_UNSUPPORTED, "n_paths*depth exceeds 4096");
  if (title_path >= 0 && (title_path >= n_paths || !title_to_chunk || n_titles <= 0))
    return fail(VFI_ERR_INVALID, "vfi_fuse_hybrid: title path needs a title -> chunk map");
  if (sparse_path >= n_paths) return fail(VFI_ERR_INVALID, "vfi_fuse_hybrid: sparse path out of range");
  if (nq == 0) return VFI_OK;
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vfi::hybrid_fuse_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(
      ids, scores, n_paths, depth, path_stride, query_stride, title_to_chunk, n_titles, title_path, sparse_path, k_rrf, k, out_scores, out_ids);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return VFI_OK;
}

}  // extern "C"

// =============================================================================================
// BM25
// =============================================================================================
struct BmScratch {      // per-call scratch of a BM25 search (pooled: concurrent searches on one posting set are allowed)
  DevBuf tok, qptr, cand, cand_count, keys, keys_n, bound, out_scores, out_ids, ctr, dump, sort;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t own = nullptr;
  bool busy = false;
  void destroy() {
    for (DevBuf* w : {&tok, &qptr, &cand, &cand_count, &keys, &keys_n, &bound, &out_scores, &out_ids, &ctr, &dump, &sort}) w->release();
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (own) cudaStreamDestroy(own);
  }
};

struct vfi_bm25 {
  int device = 0, num_sms = 148;
  int64_t n_vocab = 0, n_docs = 0, nnz = 0, id_offset = 0;
  int all_positive = 1;
  int64_t* indptr = nullptr;
  int32_t* indices = nullptr;     // doc ids: range-boundary searches
  uint2* packed = nullptr;        // {byte offset in the range accumulator, impact}: what the scoring loop streams
  std::vector<int64_t> h_indptr;  // host copy: df lookups for stats and validation
  int profile = 0;
  vfi_bm25_stats stats{};
  std::mutex mu;                  // scratch pool + stats
  std::vector<BmScratch*> pool;
};

namespace {

BmScratch* acquire_scratch(vfi_bm25* b) {
  {
    std::lock_guard<std::mutex> lock(b->mu);
    for (BmScratch* s : b->pool)
      if (!s->busy) {
        s->busy = true;
        return s;
      }
  }
  BmScratch* s = new BmScratch();
  s->busy = true;
  if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    s->destroy();
    delete s;
    fail(VFI_ERR_CUDA, "bm25 scratch: event/stream creation failed");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(b->mu);
  b->pool.push_back(s);
  return s;
}
void release_scratch(vfi_bm25* b, BmScratch* s) {
  std::lock_guard<std::mutex> lock(b->mu);
  s->busy = false;
}

int bm25_create_impl(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_vocab, int64_t n_docs,
                     int64_t id_offset, int mem, int device, vfi_bm25_t** out) {
  if (!out) return fail(VFI_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!indptr || n_vocab < 0 || n_docs < 0 || id_offset < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_create");
  if (mem != VFI_MEM_HOST && mem != VFI_MEM_DEVICE) return fail(VFI_ERR_INVALID, "vfi_bm25_create: mem must be VFI_MEM_HOST or VFI_MEM_DEVICE");
  if (n_docs >= 0x7FFFFFF0ll) return fail(VFI_ERR_UNSUPPORTED, "a BM25 shard holds at most 2^31-16 docs");
  if (id_offset + n_docs >= 0xFFFFFFFFll) return fail(VFI_ERR_UNSUPPORTED, "global doc ids (id offset + docs) must stay below 2^32 - 1");
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  vfi_bm25* b = new vfi_bm25();
  b->device = device;
  b->num_sms = prop.multiProcessorCount;
  b->n_vocab = n_vocab;
  b->n_docs = n_docs;
  b->id_offset = id_offset;
  auto bail = [&](int code) {
    vfi_bm25_destroy(b);
    return code;
  };
  b->h_indptr.resize(n_vocab + 1);
  if (mem == VFI_MEM_HOST) std::memcpy(b->h_indptr.data(), indptr, sizeof(int64_t) * (n_vocab + 1));
  else if (cudaMemcpy(b->h_indptr.data(), indptr, sizeof(int64_t) * (n_vocab + 1), cudaMemcpyDeviceToHost) != cudaSuccess)
    return bail(fail(VFI_ERR_CUDA, "vfi_bm25_create: reading indptr from the device failed"));
  const int64_t nnz = b->h_indptr[n_vocab];
  b->nnz = nnz;
  if (b->h_indptr[0] != 0 || nnz < 0 || (nnz > 0 && (!indices || !data))) return bail(fail(VFI_ERR_INVALID, "bad posting arrays"));
  for (int64_t t = 0; t < n_vocab; ++t)
    if (b->h_indptr[t] > b->h_indptr[t + 1]) return bail(fail(VFI_ERR_INVALID, "indptr must be non-decreasing"));
  static_assert((vfi::kBmRange & (vfi::kBmRange - 1)) == 0, "packed postings need a power-of-two range");
  float* d_data = nullptr;        // the impacts as given: validated, packed, then released
  const float* data_dev = data;
  if (cudaMalloc(&b->indptr, sizeof(int64_t) * (n_vocab + 1)) != cudaSuccess ||
      cudaMalloc(&b->indices, std::max<size_t>(16, sizeof(int32_t) * nnz)) != cudaSuccess ||
      cudaMalloc(&b->packed, std::max<size_t>(16, sizeof(uint2) * nnz)) != cudaSuccess ||
      (mem == VFI_MEM_HOST && cudaMalloc(&d_data, std::max<size_t>(16, sizeof(float) * nnz)) != cudaSuccess))
    return bail(fail(VFI_ERR_NOMEM, "cudaMalloc postings failed"));
  struct FreeData {
    float* p;
    ~FreeData() { if (p) cudaFree(p); }
  } free_data{d_data};
  const cudaMemcpyKind kind = mem == VFI_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  cudaMemcpy(b->indptr, b->h_indptr.data(), sizeof(int64_t) * (n_vocab + 1), cudaMemcpyHostToDevice);
  if (nnz > 0) {
    cudaMemcpy(b->indices, indices, sizeof(int32_t) * nnz, kind);
    if (mem == VFI_MEM_HOST) {
      cudaMemcpy(d_data, data, sizeof(float) * nnz, kind);
      data_dev = d_data;
    }
  }
  // The kernel indexes its shared accumulator with the doc ids: every posting list must hold in-range, strictly
  // ascending doc ids (the bm25s layout).  Checked on the device, which also finds whether every impact is positive.
  uint32_t* d_flags = nullptr;
  if (cudaMalloc(&d_flags, 16) != cudaSuccess) return bail(fail(VFI_ERR_NOMEM, "cudaMalloc failed"));
  cudaMemset(d_flags, 0, 16);
  if (nnz > 0) {
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(nnz, 256), 148 * 32));
    vfi::bm25_validate_kernel<<<blocks, 256>>>(b->indptr, b->indices, data_dev, n_vocab, nnz, n_docs, d_flags);
    vfi::bm25_pack_kernel<<<blocks, 256>>>(b->indices, data_dev, nnz, b->packed);
    LAUNCHED();
    LAUNCHED();
  }
  uint32_t h_flags[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpy(h_flags, d_flags, 16, cudaMemcpyDeviceToHost);
  cudaFree(d_flags);
  if (e != cudaSuccess) return bail(fail(VFI_ERR_CUDA, std::string("bm25 setup: ") + cudaGetErrorString(e)));
  if (h_flags[0]) return bail(fail(VFI_ERR_INVALID, "posting doc id out of range"));
  if (h_flags[1]) return bail(fail(VFI_ERR_INVALID, "postings of a token must have strictly ascending doc ids"));
  b->all_positive = h_flags[2] ? 0 : 1;
  cudaFuncSetAttribute(vfi::bm25_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  e = cudaGetLastError();
  if (e != cudaSuccess) return bail(fail(VFI_ERR_CUDA, std::string("bm25 setup: ") + cudaGetErrorString(e)));
  *out = b;
  return VFI_OK;
}

}  // namespace

extern "C" {

int vfi_bm25_create(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_vocab, int64_t n_docs,
                    int64_t id_offset, int device, vfi_bm25_t** out) {
  return bm25_create_impl(indptr, indices, data, n_vocab, n_docs, id_offset, VFI_MEM_HOST, device, out);
}
int vfi_bm25_create_from(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_vocab, int64_t n_docs,
                         int64_t id_offset, int mem, int device, vfi_bm25_t** out) {
  return bm25_create_impl(indptr, indices, data, n_vocab, n_docs, id_offset, mem, device, out);
}

int vfi_bm25_destroy(vfi_bm25_t* b) {
  if (!b) return VFI_OK;
  DeviceGuard guard(b->device);
  cudaDeviceSynchronize();
  if (b->indptr) cudaFree(b->indptr);
  if (b->indices) cudaFree(b->indices);
  if (b->packed) cudaFree(b->packed);
  for (BmScratch* s : b->pool) {
    s->destroy();
    delete s;
  }
  cudaGetLastError();
  delete b;
  return VFI_OK;
}

int64_t vfi_bm25_ndocs(const vfi_bm25_t* b) { return b ? b->n_docs : 0; }

int vfi_bm25_set_profile(vfi_bm25_t* b, int on) {
  if (!b) return fail(VFI_ERR_INVALID, "null argument");
  b->profile = on;
  return VFI_OK;
}

int vfi_bm25_get_stats(vfi_bm25_t* b, vfi_bm25_stats* out, int reset) {
  if (!b || !out) return fail(VFI_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(b->mu);
  *out = b->stats;
  if (reset) b->stats = vfi_bm25_stats{};
  return VFI_OK;
}

}  // extern "C"

static int bm25_validate_tokens(vfi_bm25* b, const int32_t* q_tokens, const int64_t* q_indptr, int64_t nq, int64_t* bytes) {
  int64_t total = 0;
  for (int64_t q = 0; q < nq; ++q) {
    const int64_t t0 = q_indptr[q], t1 = q_indptr[q + 1];
    if (t1 < t0) return fail(VFI_ERR_INVALID, "q_indptr must be non-decreasing");
    if (t1 - t0 > 0x7FFFFFFF) return fail(VFI_ERR_INVALID, "a query has more than 2^31 tokens");
    for (int64_t i = t0; i < t1; ++i) {
      const int32_t t = q_tokens[i];
      if (t < 0 || t >= b->n_vocab) return fail(VFI_ERR_INVALID, "token id out of range (drop unknown tokens before the call)");
      total += (b->h_indptr[t + 1] - b->h_indptr[t]) * 8;
    }
  }
  *bytes = total;
  return VFI_OK;
}

// segmenting of the doc range: work item = (query, segment of <= kBmMaxRanges ranges)
static void bm25_segments(const vfi_bm25* b, int64_t nq, int n_ctas, int64_t* o_n_seg, int64_t* o_seg_docs) {
  const int64_t n_ranges = std::max<int64_t>(1, ceil_div(b->n_docs, vfi::kBmRange));
  // about 8 work items per resident CTA (the atomic work queue balances heavy and light queries)
  int64_t n_seg = std::max<int64_t>(ceil_div(n_ranges, vfi::kBmMaxRanges),
                                    std::min<int64_t>(n_ranges, ceil_div(static_cast<int64_t>(8) * n_ctas, std::max<int64_t>(nq, 1))));
  n_seg = std::min<int64_t>(n_seg, n_ranges);
  const int64_t seg_docs = ceil_div(n_ranges, n_seg) * vfi::kBmRange;
  *o_n_seg = std::max<int64_t>(1, ceil_div(b->n_docs, seg_docs));
  *o_seg_docs = seg_docs;
}

extern "C" {

int vfi_bm25_search(vfi_bm25_t* b, const int32_t* q_tokens, const int64_t* q_indptr, int64_t nq, int k, float* out_scores,
                    int64_t* out_ids, int mem, void* stream) {
  if (!b || !q_indptr || nq < 0 || (nq > 0 && (!out_scores || !out_ids))) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_search");
  if (k <= 0) return fail(VFI_ERR_INVALID, "k must be positive");
  if (k > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "bm25 top-k supports k <= 2048 (use vfi_bm25_rank_all for k = N)");
  if (nq == 0) return VFI_OK;
  if (mem != VFI_MEM_HOST && mem != VFI_MEM_DEVICE) return fail(VFI_ERR_INVALID, "vfi_bm25_search: mem must be VFI_MEM_HOST or VFI_MEM_DEVICE");
  const bool dev_out = mem == VFI_MEM_DEVICE;   // token buffers are host memory either way; `mem` says where the outputs live
  int64_t bytes = 0;
  VFI_TRY(bm25_validate_tokens(b, q_tokens, q_indptr, nq, &bytes));
  DeviceGuard guard(b->device);
  BmScratch* w = acquire_scratch(b);
  if (!w) return VFI_ERR_CUDA;
  cudaStream_t st = (stream == nullptr && !dev_out) ? w->own : static_cast<cudaStream_t>(stream);
  auto body = [&]() -> int {
    const int64_t n_tok = q_indptr[nq];
    const int keep = static_cast<int>(round_up(k, 32));
    // Key buffer of k' + 256 keys rounded to a power of two (512 keys for a depth-200 list): with the 32 KB range
    // accumulator that is 42 KB per CTA, five CTAs (40 warps) per SM.  Round 2, C4 shard: 2.05 ms with a 2048-key buffer
    // compacted by a bitonic sort and four CTAs per SM, 1.78 ms with the radix-walk compaction, 1.69 ms with five CTAs.
    int cap = 1;
    while (cap < keep + vfi::kBmScan) cap <<= 1;
    const size_t smem = ((sizeof(vfi::Bm25Smem) + 15) & ~size_t(15)) + static_cast<size_t>(cap) * 8;
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vfi::bm25_kernel, vfi::kBmThreads, smem) != cudaSuccess) {
      cudaGetLastError();
      occ = 1;
    }
    const int n_ctas = b->num_sms * std::max(1, occ);
    int64_t n_seg = 1, seg_docs = vfi::kBmRange;
    bm25_segments(b, nq, n_ctas, &n_seg, &seg_docs);
    if (n_seg > 1024) return fail(VFI_ERR_UNSUPPORTED, "bm25 shard too large: more than 1024 doc segments (shard the postings)");
    VFI_TRY(w->tok.ensure(std::max<size_t>(16, static_cast<size_t>(n_tok) * 4)));
    VFI_TRY(w->qptr.ensure(static_cast<size_t>(nq + 1) * 8));
    VFI_TRY(w->cand.ensure(static_cast<size_t>(n_seg) * nq * keep * 8));
    VFI_TRY(w->cand_count.ensure(static_cast<size_t>(n_seg) * nq * 4));
    VFI_TRY(w->keys.ensure(static_cast<size_t>(nq) * keep * 8));
    VFI_TRY(w->keys_n.ensure(static_cast<size_t>(nq) * 4));
    VFI_TRY(w->bound.ensure(static_cast<size_t>(nq) * 4));
    if (!dev_out) {
      VFI_TRY(w->out_scores.ensure(static_cast<size_t>(nq) * k * 4));
      VFI_TRY(w->out_ids.ensure(static_cast<size_t>(nq) * k * 8));
    }
    float* d_scores = dev_out ? out_scores : w->out_scores.as<float>();
    int64_t* d_ids = dev_out ? out_ids : w->out_ids.as<int64_t>();
    VFI_TRY(w->ctr.ensure(16 + static_cast<size_t>(nq) * 8));
    if (n_tok > 0) VFI_CUDA(cudaMemcpyAsync(w->tok.p, q_tokens, static_cast<size_t>(n_tok) * 4, cudaMemcpyHostToDevice, st));
    VFI_CUDA(cudaMemcpyAsync(w->qptr.p, q_indptr, static_cast<size_t>(nq + 1) * 8, cudaMemcpyHostToDevice, st));
    VFI_CUDA(cudaMemsetAsync(w->ctr.p, 0, 16 + static_cast<size_t>(nq) * 8, st));
    VFI_CUDA(cudaMemsetAsync(w->cand_count.p, 0, static_cast<size_t>(n_seg) * nq * 4, st));
    vfi::Bm25Params p{};
    p.indptr = b->indptr;
    p.indices = b->indices;
    p.packed = b->packed;
    p.n_docs = b->n_docs;
    p.n_seg = static_cast<int>(n_seg);
    p.seg_docs = seg_docs;
    p.q_tokens = w->tok.as<int32_t>();
    p.q_indptr = w->qptr.as<int64_t>();
    p.nq = static_cast<int>(nq);
    p.nq_pad = static_cast<int>(nq);
    p.keep = keep;
    p.cap = cap;
    p.all_positive = b->all_positive;
    p.cand = w->cand.as<uint64_t>();
    p.cand_count = w->cand_count.as<uint32_t>();
    p.work_counter = w->ctr.as<uint32_t>();
    p.qtau = reinterpret_cast<unsigned long long*>(w->ctr.as<uint8_t>() + 16);
    p.dump = nullptr;
    const int grid = static_cast<int>(std::min<int64_t>(n_ctas, nq * n_seg));
    const bool prof = b->profile != 0;
    if (prof) cudaEventRecord(w->ev0, st);
    vfi::bm25_kernel<<<grid, vfi::kBmThreads, smem, st>>>(p);
    LAUNCHED();
    if (prof) cudaEventRecord(w->ev1, st);
    VFI_CUDA(cudaGetLastError());
    vfi::cand_reduce_kernel<vfi::SelectSmem><<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(
        w->cand.as<uint64_t>(), w->cand_count.as<uint32_t>(), static_cast<int>(n_seg), static_cast<int>(nq), keep, keep, nullptr,
        w->keys.as<uint64_t>(), w->keys_n.as<uint32_t>(), w->bound.as<float>());
    LAUNCHED();
    vfi::finalize_kernel<0><<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(
        w->keys.as<uint64_t>(), keep, keep, nullptr, w->keys_n.as<uint32_t>(), k, b->id_offset, nullptr, nullptr, d_scores, d_ids,
        nullptr, nullptr, nullptr, nullptr);
    LAUNCHED();
    if (b->all_positive) {
      vfi::bm25_zero_fill_kernel<<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(d_scores, d_ids, static_cast<int>(nq), k,
                                                                                        b->n_docs, b->id_offset);
      LAUNCHED();
    }
    VFI_CUDA(cudaGetLastError());
    if (!dev_out) {
      VFI_CUDA(cudaMemcpyAsync(out_scores, d_scores, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, st));
      VFI_CUDA(cudaMemcpyAsync(out_ids, d_ids, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, st));
    }
    // the token staging buffers are reused by the next call and the host token arrays may be freed by the caller: wait
    VFI_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    const bool have_ms = prof && cudaEventElapsedTime(&ms, w->ev0, w->ev1) == cudaSuccess;
    cudaGetLastError();
    std::lock_guard<std::mutex> lock(b->mu);
    b->stats.postings_bytes = bytes;
    b->stats.launches++;
    if (have_ms) {
      b->stats.score_ms_total += ms;
      b->stats.score_ms_samples++;
    }
    return VFI_OK;
  };
  const int rc = body();
  if (rc != VFI_OK) cudaStreamSynchronize(st);
  release_scratch(b, w);
  return rc;
}

}  // extern "C"

// all scores of one query into `dump` (device fp32 [n_docs]); enqueued on st, tokens staged in w
static int bm25_dump_scores(vfi_bm25* b, BmScratch* w, const int32_t* q_tokens, int64_t n_tokens, float* dump, cudaStream_t st) {
  const int64_t qptr[2] = {0, n_tokens};
  const int cap = 32;       // the key buffer is not used when every score is written out
  const size_t smem = ((sizeof(vfi::Bm25Smem) + 15) & ~size_t(15)) + static_cast<size_t>(cap) * 8;
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vfi::bm25_kernel, vfi::kBmThreads, smem) != cudaSuccess) {
    cudaGetLastError();
    occ = 1;
  }
  const int n_ctas = b->num_sms * std::max(1, occ);
  const int64_t n_ranges = std::max<int64_t>(1, ceil_div(b->n_docs, vfi::kBmRange));
  int64_t n_seg = std::max<int64_t>(ceil_div(n_ranges, vfi::kBmMaxRanges), std::min<int64_t>(n_ranges, n_ctas));
  const int64_t seg_docs = ceil_div(n_ranges, n_seg) * vfi::kBmRange;
  n_seg = std::max<int64_t>(1, ceil_div(b->n_docs, seg_docs));
  VFI_TRY(w->tok.ensure(std::max<size_t>(16, static_cast<size_t>(n_tokens) * 4)));
  VFI_TRY(w->qptr.ensure(16));
  VFI_TRY(w->ctr.ensure(16));
  if (n_tokens > 0) VFI_CUDA(cudaMemcpyAsync(w->tok.p, q_tokens, static_cast<size_t>(n_tokens) * 4, cudaMemcpyHostToDevice, st));
  VFI_CUDA(cudaMemcpyAsync(w->qptr.p, qptr, 16, cudaMemcpyHostToDevice, st));
  VFI_CUDA(cudaMemsetAsync(w->ctr.p, 0, 16, st));
  vfi::Bm25Params p{};
  p.indptr = b->indptr;
  p.indices = b->indices;
  p.packed = b->packed;
  p.n_docs = b->n_docs;
  p.n_seg = static_cast<int>(n_seg);
  p.seg_docs = seg_docs;
  p.q_tokens = w->tok.as<int32_t>();
  p.q_indptr = w->qptr.as<int64_t>();
  p.nq = 1;
  p.nq_pad = 1;
  p.keep = 32;
  p.cap = cap;
  p.all_positive = b->all_positive;
  p.cand = nullptr;
  p.cand_count = nullptr;
  p.work_counter = w->ctr.as<uint32_t>();
  p.qtau = nullptr;
  p.dump = dump;
  vfi::bm25_kernel<<<static_cast<int>(std::min<int64_t>(n_ctas, n_seg)), vfi::kBmThreads, smem, st>>>(p);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  VFI_CUDA(cudaStreamSynchronize(st));   // qptr lives on this stack frame
  return VFI_OK;
}

extern "C" {

int vfi_bm25_score_all(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, float* out, int mem, void* stream) {
  if (!b || !out || n_tokens < 0 || (n_tokens > 0 && !q_tokens)) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_score_all");
  const int64_t qptr[2] = {0, n_tokens};
  int64_t bytes = 0;
  VFI_TRY(bm25_validate_tokens(b, q_tokens, qptr, 1, &bytes));
  if (b->n_docs == 0) return VFI_OK;
  DeviceGuard guard(b->device);
  BmScratch* w = acquire_scratch(b);
  if (!w) return VFI_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto body = [&]() -> int {
    float* dump = out;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(w->dump.ensure(static_cast<size_t>(b->n_docs) * 4));
      dump = w->dump.as<float>();
    }
    VFI_TRY(bm25_dump_scores(b, w, q_tokens, n_tokens, dump, st));
    if (mem == VFI_MEM_HOST) {
      VFI_CUDA(cudaMemcpyAsync(out, dump, static_cast<size_t>(b->n_docs) * 4, cudaMemcpyDeviceToHost, st));
      VFI_CUDA(cudaStreamSynchronize(st));
    }
    return VFI_OK;
  };
  const int rc = body();
  if (rc != VFI_OK) cudaStreamSynchronize(st);
  release_scratch(b, w);
  return rc;
}

// Every doc ranked: scores of all docs, then a stable LSD radix sort of (~orderable(score), id) pairs that start in id
// order (radix_sort.cuh) — ties keep ascending ids, which is the total order.  Outputs start at rank `first`.
int vfi_bm25_rank_range(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, int64_t first, int64_t count,
                        float* out_scores, int64_t* out_ids, void* stream) {
  if (!b || n_tokens < 0 || (n_tokens > 0 && !q_tokens) || first < 0 || count < 0 || (count > 0 && (!out_scores || !out_ids)))
    return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_rank_range");
  const int64_t n = b->n_docs;
  if (first + count > n) return fail(VFI_ERR_INVALID, "vfi_bm25_rank_range: ranks out of range");
  if (count == 0) return VFI_OK;
  const int64_t qptr[2] = {0, n_tokens};
  int64_t bytes = 0;
  VFI_TRY(bm25_validate_tokens(b, q_tokens, qptr, 1, &bytes));
  DeviceGuard guard(b->device);
  BmScratch* w = acquire_scratch(b);
  if (!w) return VFI_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto body = [&]() -> int {
    const int n_tiles = static_cast<int>(ceil_div(n, vfi::kRsTile));
    const size_t n_pad = static_cast<size_t>(round_up(n, 64));
    // layout: scores f32 | keys A | vals A | keys B | vals B | tile counters | out ids
    const size_t bytes_needed = n_pad * 4 * 5 + static_cast<size_t>(256) * n_tiles * 4 + n_pad * 8 + 1024;
    VFI_TRY(w->sort.ensure(bytes_needed));
    uint8_t* base = w->sort.as<uint8_t>();
    float* scores = reinterpret_cast<float*>(base);
    uint32_t* ka = reinterpret_cast<uint32_t*>(base + n_pad * 4);
    uint32_t* va = ka + n_pad;
    uint32_t* kb = va + n_pad;
    uint32_t* vb = kb + n_pad;
    uint32_t* hist = vb + n_pad;
    int64_t* ids64 = reinterpret_cast<int64_t*>(reinterpret_cast<uint8_t*>(hist) + ((static_cast<size_t>(256) * n_tiles * 4 + 255) & ~size_t(255)));
    VFI_TRY(bm25_dump_scores(b, w, q_tokens, n_tokens, scores, st));
    const unsigned blocks = static_cast<unsigned>(ceil_div(n, 256));
    vfi::rank_keys_kernel<<<blocks, 256, 0, st>>>(scores, n, ka, va);
    LAUNCHED();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 8 * pass;
      vfi::rs_hist_kernel<<<n_tiles, vfi::kRsThreads, 0, st>>>(ka, n, shift, hist, n_tiles);
      vfi::rs_scan_kernel<<<1, 1024, 0, st>>>(hist, static_cast<int64_t>(256) * n_tiles);
      vfi::rs_scatter_kernel<<<n_tiles, vfi::kRsThreads, 0, st>>>(ka, va, n, shift, hist, n_tiles, kb, vb);
      LAUNCHED(); LAUNCHED(); LAUNCHED();
      std::swap(ka, kb);
      std::swap(va, vb);
    }
    vfi::ranked_out_kernel<<<static_cast<unsigned>(ceil_div(count, 256)), 256, 0, st>>>(ka + first, va + first, count, b->id_offset, scores, ids64);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    VFI_CUDA(cudaMemcpyAsync(out_scores, scores, static_cast<size_t>(count) * 4, cudaMemcpyDeviceToHost, st));
    VFI_CUDA(cudaMemcpyAsync(out_ids, ids64, static_cast<size_t>(count) * 8, cudaMemcpyDeviceToHost, st));
    VFI_CUDA(cudaStreamSynchronize(st));
    return VFI_OK;
  };
  const int rc = body();
  if (rc != VFI_OK) cudaStreamSynchronize(st);
  release_scratch(b, w);
  return rc;
}

int vfi_bm25_rank_all(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, float* out_scores, int64_t* out_ids, void* stream) {
  if (!b) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_rank_all");
  return vfi_bm25_rank_range(b, q_tokens, n_tokens, 0, b->n_docs, out_scores, out_ids, stream);
}

}  // extern "C"
