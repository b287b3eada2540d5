// api_dense.cu — the dense half of include/vfi.h: vfi_index_* (faiss.IndexFlatIP), vfi_normalize_l2, vfi_cosine_topk.
// Host orchestration of the kernels in dense_fused.cuh (K1), dense_support.cuh (K1b, K2, K6), reduce.cuh (K1c) and
// exact_stream.cuh (exact streaming scorer).
//
// Re-entrancy (SURVEY.md §8b "Threading": the reference calls retriever.invoke from concurrent request threads with no
// lock, /root/reference/src/utils/vllmChatService.py:85-88,404): a search holds the index's reader lock only and works in a
// Workspace of its own taken from a pool (device scratch, prepared queries, certificate flag, events, a private stream
// for host-buffer calls), so any number of threads may search one index at once; add/reserve/set_option take the writer
// lock (not concurrent with search, as with faiss).
//
// Sharded search (vfi_index_search_begin_push, with api_exchange.h): the rescoring kernel of a batch is also the sender of
// the multi-GPU exchange; the batch's certificate count reaches the host through mapped memory (no copy operation).
// Debugging aids: VFI_TRACE_HOST (wall-clock phases of a host-buffer call), VFI_TRACE_STEPS (device time per stream operation).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <shared_mutex>
#include <unordered_map>

#include "api_common.h"
#include "api_exchange.h"
#include "dense_fused.cuh"
#include "dense_small.cuh"
#include "dense_support.cuh"
#include "exact_stream.cuh"

using namespace vfi_host;

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// bf16 row-major [rows][cols] with 128B swizzle, box = [box_rows][64]
int make_tmap(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(VFI_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_elems) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VFI_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  return VFI_OK;
}

// the streaming scorers are instantiated per query count (1..8) so the accumulators stay in registers
template <template <typename, int> class Fn, typename RowT, typename... Args>
void dispatch_nq(int nq, Args&&... args) {
  switch (nq) {
    case 1: Fn<RowT, 1>::run(args...); break;
    case 2: Fn<RowT, 2>::run(args...); break;
    case 3: Fn<RowT, 3>::run(args...); break;
    case 4: Fn<RowT, 4>::run(args...); break;
    case 5: Fn<RowT, 5>::run(args...); break;
    case 6: Fn<RowT, 6>::run(args...); break;
    case 7: Fn<RowT, 7>::run(args...); break;
    default: Fn<RowT, 8>::run(args...); break;
  }
}
template <typename RowT, int NQ>
struct GemvLaunch {
  static void run(int grid, size_t smem, cudaStream_t st, const RowT* rows, int64_t pitch, int dp, int64_t n, const float* q,
                  int keep, int cap_s, uint64_t* cand, uint32_t* cnt, int nq_pad, int cand_cap) {
    vfi::gemv_topk_kernel<RowT, NQ><<<grid, vfi::kGemvThreads, smem, st>>>(rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap);
  }
};
template <typename RowT, int NQ>
struct GemvOcc {
  static void run(size_t smem, int* out) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vfi::gemv_topk_kernel<RowT, NQ>, vfi::kGemvThreads, smem) != cudaSuccess) {
      cudaGetLastError();
      occ = 1;
    }
    *out = occ < 1 ? 1 : occ;
  }
};
template <typename RowT, int NQ>
struct GemvAttr {
  static void run() { cudaFuncSetAttribute(vfi::gemv_topk_kernel<RowT, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); }
};
template <typename RowT, int NQ>
struct ExactAttr {
  static void run() {
    cudaFuncSetAttribute(vfi::exact_scores_kernel<RowT, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kExSmemBudget);
    constexpr int kPref = vfi::exact_pref_threads(NQ);
    if (kPref != vfi::kExThreads)
      cudaFuncSetAttribute(vfi::exact_scores_kernel<RowT, NQ, kPref>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kExSmemBudget);
  }
};
template <typename RowT, int NQ>
struct ExactLaunch {
  static void run(int num_sms, cudaStream_t st, const RowT* rows, int64_t pitch, int dp, int64_t n, const float* qcanon,
                  const int* qsel, int q0, float* scores, int64_t ld) {
    constexpr int kPref = vfi::exact_pref_threads(NQ);
    // more warps per SM for small groups when the queries and the row tiles still fit shared memory
    if (kPref != vfi::kExThreads && vfi::exact_smem_bytes(dp, NQ, kPref) <= static_cast<size_t>(vfi::kExSmemBudget)) {
      const int grid = static_cast<int>(std::min<int64_t>(num_sms, ceil_div(n, kPref)));
      vfi::exact_scores_kernel<RowT, NQ, kPref><<<grid, kPref, vfi::exact_smem_bytes(dp, NQ, kPref), st>>>(rows, pitch, dp, n, qcanon, qsel,
                                                                                                          q0, scores, ld);
    } else {
      const int grid = static_cast<int>(std::min<int64_t>(num_sms, ceil_div(n, vfi::kExThreads)));
      vfi::exact_scores_kernel<RowT, NQ><<<grid, vfi::kExThreads, vfi::exact_smem_bytes(dp, NQ), st>>>(rows, pitch, dp, n, qcanon, qsel, q0,
                                                                                                      scores, ld);
    }
  }
};

constexpr int kMaxQueriesPerLaunch = 1024;
constexpr int kExhaustiveRows = 4096;   // shards this small skip the tensor-core pass
constexpr int kFusedMaxKeep = 2560;     // k' the candidate buffers of K1 take (k = 2048 -> k' = 2560)
constexpr int kGemvMaxKeep = 512;
constexpr int kExactSmallRows = 32768;  // one CTA per query selects straight from the score array up to here
constexpr int kMaxWorkspaces = 64;      // batches in flight + concurrent synchronous callers per index

// Everything one batch in flight needs.  Pooled per index; a synchronous search borrows one for the call, a pipelined
// batch keeps it from vfi_index_search_begin to vfi_index_search_finish.
struct Workspace {
  DevBuf qin, qcanon, qg, eps, cand, cand_count, keys, keys_n, bound, keys2, flag, out_scores, out_ids, dbg, sel, tau,
      ex_scores, ex_state, ex_keys;
  int* h_flag = nullptr;         // pinned: number of queries of the batch whose certificate failed
  int* h_flag_dev = nullptr;     // the same word as the device sees it (written by the last CTA of the tail)
  cudaEvent_t done = nullptr;    // recorded behind the flag copy
  cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};   // VFI_OPT_PROFILE: around the dominant kernel [0,1] and the tail [2,3]
  cudaStream_t own = nullptr;    // host-buffer calls that pass no stream run here, so concurrent callers overlap
  cudaEvent_t tev[10] = {};      // VFI_TRACE_STEPS=1 (debugging aid): one event behind every stream operation of a batch
  // batch state
  bool needs_check = false, used_tau = false, profiled = false;
  const void* q = nullptr;
  int q_dtype = VFI_DTYPE_F32;
  int nq = 0, k = 0;
  float* o_scores = nullptr;
  int64_t* o_ids = nullptr;
  cudaStream_t st = nullptr;
  int ticket = -1;

  int init() {
    VFI_CUDA(cudaHostAlloc(&h_flag, sizeof(int) * 4, cudaHostAllocMapped | cudaHostAllocPortable));
    VFI_CUDA(cudaHostGetDevicePointer(&h_flag_dev, h_flag, 0));
    VFI_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    for (cudaEvent_t& e : pev) VFI_CUDA(cudaEventCreate(&e));
    VFI_CUDA(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
    return VFI_OK;
  }
  void destroy() {
    for (DevBuf* b : {&qin, &qcanon, &qg, &eps, &cand, &cand_count, &keys, &keys_n, &bound, &keys2, &flag, &out_scores,
                      &out_ids, &dbg, &sel, &tau, &ex_scores, &ex_state, &ex_keys})
      b->release();
    if (h_flag) cudaFreeHost(h_flag);
    if (done) cudaEventDestroy(done);
    for (cudaEvent_t e : pev)
      if (e) cudaEventDestroy(e);
    if (own) cudaStreamDestroy(own);
    for (cudaEvent_t e : tev)
      if (e) cudaEventDestroy(e);
  }
};

// VFI_TRACE_STEPS=1: where the device time of one batch goes, operation by operation (gaps included), on stderr
const bool g_trace_steps = std::getenv("VFI_TRACE_STEPS") != nullptr;
cudaEvent_t g_trace_epoch = nullptr;
inline void trace_mark(Workspace* ws, int i, cudaStream_t st) {
  if (!g_trace_steps) return;
  if (!g_trace_epoch) { cudaEventCreate(&g_trace_epoch); cudaEventRecord(g_trace_epoch, st); }
  if (!ws->tev[i]) cudaEventCreate(&ws->tev[i]);
  cudaEventRecord(ws->tev[i], st);
}

}  // namespace

// =============================================================================================
struct vfi_index {
  int d = 0, dp = 0, store = VFI_STORE_BF16, device = 0;
  int64_t kp = 0;          // gemm operand row length (dp or 3*dp)
  int64_t n = 0, cap_rows = 0;
  int64_t id_offset = 0;
  int num_sms = 148;
  int max_pairs = 0;       // co-resident 2-CTA clusters of the pair kernel (queried at create)
  uint16_t* g = nullptr;   // [cap_rows][kp] bf16 gemm operand rows
  float* master = nullptr; // [cap_rows][dp] fp32 rows (F32 store only)
  uint32_t* xnorm_bits = nullptr;
  uint32_t* d_max_err = nullptr;
  DevBuf stage;            // add()/read_rows staging (writer lock / own lock)
  std::mutex stage_mu;
  // options (written under the writer lock)
  int64_t opt_overfetch = 0, opt_force_path = 0, opt_profile = 0, opt_tau_hint = 1, opt_num_ctas = 0, opt_cta_pair = 0, opt_tau_m = 0, opt_small = 0, opt_tail_piece = 0;
  std::shared_mutex rw;    // searches share it, add/reserve/options own it
  std::mutex pool_mu;      // workspace pool, tickets, stats
  std::vector<std::unique_ptr<Workspace>> pool;
  std::vector<Workspace*> free_ws;
  std::unordered_map<int, Workspace*> tickets;   // batches between search_begin and search_finish
  int next_ticket = 1;
  vfi_search_stats stats{};
};

namespace {

Workspace* acquire_ws(vfi_index* idx) {
  {
    std::lock_guard<std::mutex> lock(idx->pool_mu);
    if (!idx->free_ws.empty()) {
      Workspace* w = idx->free_ws.back();
      idx->free_ws.pop_back();
      return w;
    }
    if (static_cast<int>(idx->pool.size()) >= kMaxWorkspaces) {
      fail(VFI_ERR_UNSUPPORTED, "too many batches in flight on this index (64): finish a ticket of vfi_index_search_begin first");
      return nullptr;
    }
  }
  std::unique_ptr<Workspace> w(new Workspace());
  if (w->init() != VFI_OK) {
    w->destroy();
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(idx->pool_mu);
  idx->pool.push_back(std::move(w));
  return idx->pool.back().get();
}
void release_ws(vfi_index* idx, Workspace* w) {
  std::lock_guard<std::mutex> lock(idx->pool_mu);
  w->ticket = -1;
  idx->free_ws.push_back(w);
}

}  // namespace

extern "C" {

int vfi_index_create(int d, int store_dtype, int device, vfi_index_t** out) {
  if (!out) return fail(VFI_ERR_INVALID, "out is null");
  *out = nullptr;
  if (d <= 0 || d > 8192) return fail(VFI_ERR_INVALID, "dimension must be in [1, 8192]");
  if (store_dtype != VFI_STORE_BF16 && store_dtype != VFI_STORE_F32) return fail(VFI_ERR_INVALID, "unknown store_dtype");
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  vfi_index* idx = new vfi_index();
  idx->d = d;
  idx->dp = static_cast<int>(round_up(d, 64));
  idx->store = store_dtype;
  idx->kp = (store_dtype == VFI_STORE_F32) ? 3 * static_cast<int64_t>(idx->dp) : idx->dp;
  idx->device = device;
  idx->num_sms = prop.multiProcessorCount;
  auto bail = [&](int code) {
    vfi_index_destroy(idx);
    return code;
  };
  if (cudaMalloc(&idx->xnorm_bits, 8) != cudaSuccess) return bail(fail(VFI_ERR_NOMEM, "cudaMalloc failed"));
  cudaMemset(idx->xnorm_bits, 0, 8);
  idx->d_max_err = idx->xnorm_bits + 1;
  static std::once_flag attrs_once[16];
  std::call_once(attrs_once[device & 15], [] {
    cudaFuncSetAttribute(vfi::dense_fused_kernel<vfi::MODE_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kDenseSmemBytes);
    cudaFuncSetAttribute(vfi::dense_fused_kernel<vfi::MODE_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kDenseSmemBytes);
    cudaFuncSetAttribute(vfi::dense_fused_kernel<vfi::MODE_CHUNKMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kDenseSmemBytes);
    cudaFuncSetAttribute(vfi::dense_fused_pair_kernel<vfi::MODE_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kPairSmemBytes);
    cudaFuncSetAttribute(vfi::dense_fused_pair_kernel<vfi::MODE_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kPairSmemBytes);
    cudaFuncSetAttribute(vfi::dense_fused_pair_kernel<vfi::MODE_CHUNKMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kPairSmemBytes);
    cudaFuncSetAttribute(vfi::dense_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    GemvAttr<uint16_t, 1>::run(); GemvAttr<uint16_t, 2>::run(); GemvAttr<uint16_t, 3>::run(); GemvAttr<uint16_t, 4>::run();
    GemvAttr<uint16_t, 5>::run(); GemvAttr<uint16_t, 6>::run(); GemvAttr<uint16_t, 7>::run(); GemvAttr<uint16_t, 8>::run();
    GemvAttr<float, 1>::run(); GemvAttr<float, 2>::run(); GemvAttr<float, 3>::run(); GemvAttr<float, 4>::run();
    GemvAttr<float, 5>::run(); GemvAttr<float, 6>::run(); GemvAttr<float, 7>::run(); GemvAttr<float, 8>::run();
    ExactAttr<uint16_t, 1>::run(); ExactAttr<uint16_t, 2>::run(); ExactAttr<uint16_t, 3>::run(); ExactAttr<uint16_t, 4>::run();
    ExactAttr<uint16_t, 5>::run(); ExactAttr<uint16_t, 6>::run(); ExactAttr<uint16_t, 7>::run(); ExactAttr<uint16_t, 8>::run();
    ExactAttr<float, 1>::run(); ExactAttr<float, 2>::run(); ExactAttr<float, 3>::run(); ExactAttr<float, 4>::run();
    ExactAttr<float, 5>::run(); ExactAttr<float, 6>::run(); ExactAttr<float, 7>::run(); ExactAttr<float, 8>::run();
    cudaFuncSetAttribute(vfi::rescore_finalize_kernel<uint16_t, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(vfi::rescore_bulk_smem<256>(vfi::kRfMaxDp, 256, 1)));
    cudaFuncSetAttribute(vfi::rescore_finalize_kernel<float, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(vfi::rescore_bulk_smem<256>(vfi::kRfMaxDp, 256, 1)));
    cudaFuncSetAttribute(vfi::rescore_finalize_kernel<uint16_t, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(vfi::rescore_bulk_smem<128>(vfi::kRfMaxDp, 256, 1)));
    cudaFuncSetAttribute(vfi::rescore_finalize_kernel<float, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(vfi::rescore_bulk_smem<128>(vfi::kRfMaxDp, 256, 1)));
  });
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bail(fail(VFI_ERR_CUDA, std::string("kernel attribute setup: ") + cudaGetErrorString(e)));
  {   // how many 2-CTA clusters of the pair kernel the device holds at once (74 on a full B200)
    cudaLaunchConfig_t oc{};
    oc.gridDim = dim3(static_cast<unsigned>(idx->num_sms));
    oc.blockDim = dim3(vfi::kDenseThreads);
    oc.dynamicSmemBytes = vfi::kPairSmemBytes;
    cudaLaunchAttribute oa[1];
    oa[0].id = cudaLaunchAttributeClusterDimension;
    oa[0].val.clusterDim.x = 2;
    oa[0].val.clusterDim.y = 1;
    oa[0].val.clusterDim.z = 1;
    oc.attrs = oa;
    oc.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, vfi::dense_fused_pair_kernel<vfi::MODE_TOPK>, &oc) != cudaSuccess || nc <= 0) {
      cudaGetLastError();
      nc = idx->num_sms / 2;
    }
    idx->max_pairs = nc;
  }
  *out = idx;
  return VFI_OK;
}

int vfi_index_destroy(vfi_index_t* idx) {
  if (!idx) return VFI_OK;
  DeviceGuard guard(idx->device);
  cudaDeviceSynchronize();
  if (idx->g) cudaFree(idx->g);
  if (idx->master) cudaFree(idx->master);
  if (idx->xnorm_bits) cudaFree(idx->xnorm_bits);
  idx->stage.release();
  for (auto& w : idx->pool) w->destroy();
  cudaGetLastError();
  delete idx;
  return VFI_OK;
}

}  // extern "C"

static int grow_rows(vfi_index* idx, int64_t need_rows, cudaStream_t st) {
  if (need_rows <= idx->cap_rows) return VFI_OK;
  int64_t new_cap = std::max<int64_t>(need_rows, idx->cap_rows + idx->cap_rows / 2);
  new_cap = std::max<int64_t>(new_cap, 1024);
  uint16_t* ng = nullptr;
  float* nm = nullptr;
  cudaError_t e = cudaMalloc(&ng, static_cast<size_t>(new_cap) * idx->kp * 2);
  if (e != cudaSuccess) return fail(VFI_ERR_NOMEM, std::string("cudaMalloc corpus: ") + cudaGetErrorString(e));
  if (idx->store == VFI_STORE_F32) {
    e = cudaMalloc(&nm, static_cast<size_t>(new_cap) * idx->dp * 4);
    if (e != cudaSuccess) {
      cudaFree(ng);
      return fail(VFI_ERR_NOMEM, std::string("cudaMalloc corpus master: ") + cudaGetErrorString(e));
    }
  }
  if (idx->n > 0) {
    VFI_CUDA(cudaMemcpyAsync(ng, idx->g, static_cast<size_t>(idx->n) * idx->kp * 2, cudaMemcpyDeviceToDevice, st));
    if (nm) VFI_CUDA(cudaMemcpyAsync(nm, idx->master, static_cast<size_t>(idx->n) * idx->dp * 4, cudaMemcpyDeviceToDevice, st));
    VFI_CUDA(cudaStreamSynchronize(st));
  }
  if (idx->g) cudaFree(idx->g);
  if (idx->master) cudaFree(idx->master);
  idx->g = ng;
  idx->master = nm;
  idx->cap_rows = new_cap;
  return VFI_OK;
}

extern "C" int vfi_index_reserve(vfi_index_t* idx, int64_t n) {
  if (!idx) return fail(VFI_ERR_INVALID, "index is null");
  std::unique_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  return grow_rows(idx, n, nullptr);
}

template <typename InT>
static int add_rows(vfi_index* idx, const InT* x, int64_t n, int mem, cudaStream_t st) {
  if (n == 0) return VFI_OK;
  if (idx->n + n >= 0x7FFFFF00ll) return fail(VFI_ERR_UNSUPPORTED, "a shard holds at most 2^31-256 rows");
  if (idx->id_offset + idx->n + n >= 0xFFFFFFFFll)
    return fail(VFI_ERR_UNSUPPORTED, "global row ids (id offset + rows) must stay below 2^32 - 1");
  VFI_TRY(grow_rows(idx, idx->n + n, st));
  const int64_t chunk = std::max<int64_t>(1, (static_cast<int64_t>(256) << 20) / (static_cast<int64_t>(idx->d) * sizeof(InT)));
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t rows = std::min(chunk, n - r0);
    const InT* src = x + r0 * idx->d;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(idx->stage.ensure(static_cast<size_t>(rows) * idx->d * sizeof(InT)));
      VFI_CUDA(cudaMemcpyAsync(idx->stage.p, src, static_cast<size_t>(rows) * idx->d * sizeof(InT), cudaMemcpyHostToDevice, st));
      src = idx->stage.as<InT>();
    }
    const int threads = 256;
    const int64_t blocks = ceil_div(rows * 32, threads);
    uint16_t* gdst = idx->g + (idx->n + r0) * idx->kp;
    if (idx->store == VFI_STORE_F32) {
      vfi::prep_rows_kernel<true, InT><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
          src, rows, idx->d, idx->dp, gdst, idx->kp, idx->master + (idx->n + r0) * idx->dp, idx->xnorm_bits);
    } else {
      vfi::prep_rows_kernel<false, InT><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
          src, rows, idx->d, idx->dp, gdst, idx->kp, nullptr, idx->xnorm_bits);
    }
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    if (mem == VFI_MEM_HOST) VFI_CUDA(cudaStreamSynchronize(st));  // staging buffer is reused
  }
  VFI_CUDA(cudaStreamSynchronize(st));
  idx->n += n;
  return VFI_OK;
}

extern "C" {

int vfi_index_add(vfi_index_t* idx, const float* x, int64_t n, int mem, void* stream) {
  if (!idx || (!x && n > 0) || n < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_add");
  std::unique_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  return add_rows<float>(idx, x, n, mem, static_cast<cudaStream_t>(stream));
}

int vfi_index_add_bf16(vfi_index_t* idx, const uint16_t* x, int64_t n, int mem, void* stream) {
  if (!idx || (!x && n > 0) || n < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_add_bf16");
  if (idx->store != VFI_STORE_BF16) return fail(VFI_ERR_INVALID, "add_bf16 needs a VFI_STORE_BF16 index");
  std::unique_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  return add_rows<uint16_t>(idx, x, n, mem, static_cast<cudaStream_t>(stream));
}

int64_t vfi_index_ntotal(const vfi_index_t* idx) { return idx ? idx->n : 0; }
int vfi_index_dim(const vfi_index_t* idx) { return idx ? idx->d : 0; }

int vfi_index_set_id_offset(vfi_index_t* idx, int64_t offset) {
  if (!idx || offset < 0) return fail(VFI_ERR_INVALID, "bad id offset");
  if (offset + idx->n >= 0xFFFFFFFFll) return fail(VFI_ERR_UNSUPPORTED, "global row ids (id offset + rows) must stay below 2^32 - 1");
  std::unique_lock<std::shared_mutex> lock(idx->rw);
  idx->id_offset = offset;
  return VFI_OK;
}

int vfi_index_set_option(vfi_index_t* idx, int opt, int64_t value) {
  if (!idx) return fail(VFI_ERR_INVALID, "index is null");
  std::unique_lock<std::shared_mutex> lock(idx->rw);
  switch (opt) {
    case VFI_OPT_OVERFETCH: idx->opt_overfetch = value; break;
    case VFI_OPT_FORCE_PATH:
      if (value < 0 || value > 3) return fail(VFI_ERR_INVALID, "VFI_OPT_FORCE_PATH: 0 auto, 1 exact streaming, 2 fused tcgen05, 3 streaming GEMV");
      idx->opt_force_path = value;
      break;
    case VFI_OPT_PROFILE: idx->opt_profile = value; break;
    case VFI_OPT_TAU_HINT: idx->opt_tau_hint = value; break;
    case VFI_OPT_NUM_CTAS: idx->opt_num_ctas = value; break;
    case VFI_OPT_TAU_M:
      if (value != 0 && value != 8 && value != 16 && value != 32) return fail(VFI_ERR_INVALID, "VFI_OPT_TAU_M: 0 auto, 8, 16 or 32");
      idx->opt_tau_m = value;
      break;
    case VFI_OPT_SMALL_BATCH:
      if (value < 0 || value > 1) return fail(VFI_ERR_INVALID, "VFI_OPT_SMALL_BATCH: 0 auto (swapped-operand kernel for 9..64 queries), 1 off");
      idx->opt_small = value;
      break;
    case VFI_OPT_TAIL_PIECE:
      if (value != 0 && value != 128 && value != 256) return fail(VFI_ERR_INVALID, "VFI_OPT_TAIL_PIECE: 0 auto, 128 or 256");
      idx->opt_tail_piece = value;
      break;
    case VFI_OPT_CTA_PAIR:
      if (value < 0 || value > 2) return fail(VFI_ERR_INVALID, "VFI_OPT_CTA_PAIR: 0 auto, 1 off, 2 on");
      idx->opt_cta_pair = value;
      break;
    default: return fail(VFI_ERR_INVALID, "unknown option");
  }
  return VFI_OK;
}

int vfi_index_get_stats(vfi_index_t* idx, vfi_search_stats* out, int reset) {
  if (!idx || !out) return fail(VFI_ERR_INVALID, "null argument");
  DeviceGuard guard(idx->device);
  uint32_t bits = 0;
  VFI_CUDA(cudaMemcpy(&bits, idx->d_max_err, 4, cudaMemcpyDeviceToHost));
  float f;
  std::memcpy(&f, &bits, 4);
  std::lock_guard<std::mutex> lock(idx->pool_mu);
  idx->stats.max_abs_err = f;
  *out = idx->stats;
  if (reset) {
    idx->stats = vfi_search_stats{};
    VFI_CUDA(cudaMemset(idx->d_max_err, 0, 4));
  }
  return VFI_OK;
}

int vfi_index_reconstruct(vfi_index_t* idx, int64_t i, float* out, int mem) {
  if (!idx || !out) return fail(VFI_ERR_INVALID, "null argument");
  if (i < 0 || i >= idx->n) return fail(VFI_ERR_INVALID, "row out of range");
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  std::vector<float> row(idx->d);
  if (idx->store == VFI_STORE_F32) {
    VFI_CUDA(cudaMemcpy(row.data(), idx->master + i * idx->dp, sizeof(float) * idx->d, cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint16_t> h(idx->d);
    VFI_CUDA(cudaMemcpy(h.data(), idx->g + i * idx->kp, 2 * static_cast<size_t>(idx->d), cudaMemcpyDeviceToHost));
    for (int j = 0; j < idx->d; ++j) {
      uint32_t u = static_cast<uint32_t>(h[j]) << 16;
      std::memcpy(&row[j], &u, 4);
    }
  }
  VFI_CUDA(cudaMemcpy(out, row.data(), sizeof(float) * idx->d, mem == VFI_MEM_DEVICE ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
  return VFI_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// search internals
// ---------------------------------------------------------------------------------------------
namespace {

// rows must fit the register-resident pieces of the streaming scorer
bool gemv_ok(const vfi_index* idx) {
  const int bytes = idx->dp * (idx->store == VFI_STORE_F32 ? 4 : 2);
  return bytes <= vfi::kGemvMaxVec * 32 * 16;
}

int prep_queries(vfi_index* idx, Workspace* ws, const void* q_dev, int q_dtype, int nq, cudaStream_t st, int* zero_a = nullptr,
                 int* zero_b = nullptr) {
  VFI_TRY(ws->qcanon.ensure(static_cast<size_t>(nq) * idx->dp * 4));
  VFI_TRY(ws->qg.ensure(static_cast<size_t>(nq) * idx->kp * 2));
  VFI_TRY(ws->eps.ensure(static_cast<size_t>(nq) * 4));
  const int threads = 256;
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, threads));
  float* canon = ws->qcanon.as<float>();
  uint16_t* qg = ws->qg.as<uint16_t>();
  float* eps = ws->eps.as<float>();
  const bool split = idx->store == VFI_STORE_F32;
  if (q_dtype == VFI_DTYPE_BF16) {
    const uint16_t* q = static_cast<const uint16_t*>(q_dev);
    if (split) vfi::prep_queries_kernel<true, uint16_t><<<blocks, threads, 0, st>>>(q, nq, idx->d, idx->dp, canon, qg, idx->kp, idx->xnorm_bits, eps, zero_a, zero_b);
    else vfi::prep_queries_kernel<false, uint16_t><<<blocks, threads, 0, st>>>(q, nq, idx->d, idx->dp, canon, qg, idx->kp, idx->xnorm_bits, eps, zero_a, zero_b);
  } else {
    const float* q = static_cast<const float*>(q_dev);
    if (split) vfi::prep_queries_kernel<true, float><<<blocks, threads, 0, st>>>(q, nq, idx->d, idx->dp, canon, qg, idx->kp, idx->xnorm_bits, eps, zero_a, zero_b);
    else vfi::prep_queries_kernel<false, float><<<blocks, threads, 0, st>>>(q, nq, idx->d, idx->dp, canon, qg, idx->kp, idx->xnorm_bits, eps, zero_a, zero_b);
  }
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return VFI_OK;
}

// Exact streaming pass: canonical scores of ALL rows for the selected queries (ws->qcanon rows qsel_dev[0..nsel), or
// 0..nsel-1 when qsel_dev is null), then the exact top-k.  Groups of up to 8 queries share one pass over the corpus.
int exact_pass(vfi_index* idx, Workspace* ws, const int* qsel_dev, int nsel, int k, float* out_scores, int64_t* out_ids,
               cudaStream_t st, bool profile) {
  const int64_t n = idx->n;
  const int group = vfi::exact_max_group(idx->dp);
  if (group < 1) return fail(VFI_ERR_UNSUPPORTED, "exact streaming scorer: the query does not fit shared memory");
  const int64_t ld = round_up(std::max<int64_t>(n, 1), 4);
  VFI_TRY(ws->ex_scores.ensure(static_cast<size_t>(group) * ld * 4));
  const int k_eff = static_cast<int>(std::min<int64_t>(k, n));
  const bool big = n > kExactSmallRows;
  const size_t hist_bytes = static_cast<size_t>(group) * vfi::kRxBins * 4;
  const size_t state_bytes = static_cast<size_t>(group) * sizeof(vfi::RadixState);
  if (big) {
    VFI_TRY(ws->ex_state.ensure(state_bytes + hist_bytes));
    VFI_TRY(ws->ex_keys.ensure(static_cast<size_t>(group) * std::max(k_eff, 1) * 8));
  }
  const bool prof = idx->opt_profile != 0 && profile;
  if (prof) cudaEventRecord(ws->pev[0], st);
  for (int s0 = 0; s0 < nsel; s0 += group) {
    const int ng = std::min(group, nsel - s0);
    float* scores = ws->ex_scores.as<float>();
    if (n > 0) {
      if (idx->store == VFI_STORE_F32)
        dispatch_nq<ExactLaunch, float>(ng, idx->num_sms, st, idx->master, static_cast<int64_t>(idx->dp), idx->dp, n, ws->qcanon.as<float>(),
                                        qsel_dev, s0, scores, ld);
      else
        dispatch_nq<ExactLaunch, uint16_t>(ng, idx->num_sms, st, idx->g, idx->kp, idx->dp, n, ws->qcanon.as<float>(), qsel_dev, s0, scores, ld);
      LAUNCHED();
      VFI_CUDA(cudaGetLastError());
    }
    if (!big) {
      vfi::exact_small_kernel<<<ng, 256, sizeof(vfi::SelectSmem), st>>>(scores, ld, n, k, qsel_dev, s0, idx->id_offset, out_scores, out_ids);
      LAUNCHED();
    } else {
      vfi::RadixState* state = ws->ex_state.as<vfi::RadixState>();
      uint32_t* hist = reinterpret_cast<uint32_t*>(ws->ex_state.as<uint8_t>() + state_bytes);
      VFI_CUDA(cudaMemsetAsync(ws->ex_state.p, 0, state_bytes + hist_bytes, st));
      const int cpq = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 8192), (4 * idx->num_sms) / ng)));
      dim3 grid(static_cast<unsigned>(cpq), static_cast<unsigned>(ng));
      for (int pass = 0; pass < vfi::kRxPasses; ++pass) {
        vfi::radix_hist_kernel<<<grid, 256, 0, st>>>(scores, ld, n, k_eff, pass, state, hist);
        LAUNCHED();
      }
      uint32_t np = 2;
      while (np < static_cast<uint32_t>(k_eff)) np <<= 1;
      vfi::radix_gather_kernel<<<grid, 256, np * 8, st>>>(scores, ld, n, k, k_eff, state, ws->ex_keys.as<uint64_t>(), qsel_dev, s0,
                                                         idx->id_offset, out_scores, out_ids);
      LAUNCHED();
    }
    VFI_CUDA(cudaGetLastError());
  }
  if (prof) cudaEventRecord(ws->pev[1], st);
  return VFI_OK;
}

// n_rows/row_stride: the corpus view scanned — (idx->n, 1) for the full shard, (R, s) for a strided row sample
int launch_fused(vfi_index* idx, Workspace* ws, int nq, int keep, int mode, float* scores_out, int64_t ld_scores, const float* tau,
                 int* o_groups, int* o_nq_pad, int* o_cap, cudaStream_t st, int64_t n_rows = -1, int64_t row_stride = 1,
                 bool profile = true) {
  if (n_rows < 0) n_rows = idx->n;
  const int n_mtiles_real = static_cast<int>(ceil_div(nq, vfi::kBM));
  // CTA pairs (tcgen05 cta_group::2) work on two 128-query tiles at once.  A batch with an odd number of tiles gets one
  // padding tile (TMA fills the missing query rows with zeros, their results are never read): every batch size runs on
  // the pair kernel, which moves a third less data per FLOP into the SMs than the single-CTA kernel.
  const bool pair = idx->opt_cta_pair != 1;
  const int n_mtiles = pair ? static_cast<int>(round_up(n_mtiles_real, 2)) : n_mtiles_real;
  int n_ctas = idx->opt_num_ctas > 0 ? static_cast<int>(idx->opt_num_ctas) : idx->num_sms;
  n_ctas = std::min(std::max(n_ctas, n_mtiles), 512);   // 2*groups key buffers per query must stay <= 1024
  const int n_tiles = static_cast<int>(ceil_div(n_rows, vfi::kBN));
  int n_groups = std::max(1, n_ctas / n_mtiles);
  n_groups = std::min(n_groups, std::max(1, n_tiles));
  // pair kernel: every SM pair owns corpus tiles (n_groups = number of pairs) and walks all query tile pairs itself
  if (pair) n_groups = std::max(1, std::min(std::min(n_ctas / 2, idx->max_pairs), std::max(1, n_tiles)));
  const int n_bufs = pair ? n_groups * vfi::pair_sets_per_query(n_mtiles) : 2 * n_groups;   // key buffers per query
  const int nq_pad = n_mtiles * vfi::kBM;
  const int cap = 2 * keep + 32;
  CUtensorMap tq, td;
  VFI_TRY(make_tmap(&tq, ws->qg.p, nq, idx->kp, idx->kp, vfi::kBM));
  VFI_TRY(make_tmap(&td, idx->g, n_rows, idx->kp, idx->kp * row_stride, pair ? vfi::kBN / 2 : vfi::kBN));
  vfi::DenseParams p{};
  p.nq = nq;
  p.nq_pad = nq_pad;
  p.n_rows = static_cast<int>(n_rows);
  p.n_kblocks = static_cast<int>(idx->kp / vfi::kBK);
  p.n_mtiles = n_mtiles;
  p.n_groups = n_groups;
  p.n_tiles = n_tiles;
  p.keep = keep;
  p.cap = cap;
  p.tau_init = tau;
  p.scores_out = scores_out;
  p.ld_scores = ld_scores;
  if (mode == vfi::MODE_TOPK) {
    VFI_TRY(ws->cand.ensure(static_cast<size_t>(n_bufs) * nq_pad * cap * 8));
    VFI_TRY(ws->cand_count.ensure(static_cast<size_t>(n_bufs) * nq_pad * 4));
    p.cand = ws->cand.as<uint64_t>();
    p.cand_count = ws->cand_count.as<uint32_t>();
  }
  const bool prof = idx->opt_profile != 0 && profile;
  if (prof) cudaEventRecord(ws->pev[0], st);
  const int grid = pair ? 2 * n_groups : n_groups * n_mtiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(vfi::kDenseThreads);
  cfg.dynamicSmemBytes = pair ? vfi::kPairSmemBytes : vfi::kDenseSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2u : 1u;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pair) {
    if (mode == vfi::MODE_TOPK) VFI_CUDA(cudaLaunchKernelEx(&cfg, vfi::dense_fused_pair_kernel<vfi::MODE_TOPK>, tq, td, p));
    else if (mode == vfi::MODE_CHUNKMAX) VFI_CUDA(cudaLaunchKernelEx(&cfg, vfi::dense_fused_pair_kernel<vfi::MODE_CHUNKMAX>, tq, td, p));
    else VFI_CUDA(cudaLaunchKernelEx(&cfg, vfi::dense_fused_pair_kernel<vfi::MODE_STORE>, tq, td, p));
  } else {
    if (mode == vfi::MODE_TOPK) VFI_CUDA(cudaLaunchKernelEx(&cfg, vfi::dense_fused_kernel<vfi::MODE_TOPK>, tq, td, p));
    else if (mode == vfi::MODE_CHUNKMAX) VFI_CUDA(cudaLaunchKernelEx(&cfg, vfi::dense_fused_kernel<vfi::MODE_CHUNKMAX>, tq, td, p));
    else VFI_CUDA(cudaLaunchKernelEx(&cfg, vfi::dense_fused_kernel<vfi::MODE_STORE>, tq, td, p));
  }
  LAUNCHED();
  if (prof) cudaEventRecord(ws->pev[1], st);
  VFI_CUDA(cudaGetLastError());
  if (o_groups) *o_groups = n_bufs;
  if (o_nq_pad) *o_nq_pad = nq_pad;
  if (o_cap) *o_cap = cap;
  return VFI_OK;
}

// K1s: batches of up to 64 queries whose bf16 operand block fits shared memory (dense_small.cuh)
bool small_ok(const vfi_index* idx, int nq) {
  const int64_t nq_pad = round_up(nq, 16);
  return idx->opt_small == 0 && nq_pad <= vfi::kSmMaxQ && nq_pad * idx->kp * 2 <= vfi::kSmQBudget;
}

int launch_small(vfi_index* idx, Workspace* ws, int nq, int keep, const float* tau, int* o_groups, int* o_nq_pad, int* o_cap,
                 cudaStream_t st) {
  const int nq_pad = static_cast<int>(round_up(nq, 16));
  const int n_tiles = static_cast<int>(ceil_div(idx->n, vfi::kSmTileRows));
  const int n_ctas = std::max(1, std::min(idx->num_sms, n_tiles));
  const int cap = 2 * keep + 160;
  const int nkb = static_cast<int>(idx->kp / 64);
  CUtensorMap tq, td;
  VFI_TRY(make_tmap(&tq, ws->qg.p, nq, idx->kp, idx->kp, nq_pad));
  VFI_TRY(make_tmap(&td, idx->g, idx->n, idx->kp, idx->kp, vfi::kSmTileRows));
  VFI_TRY(ws->cand.ensure(static_cast<size_t>(n_ctas) * nq_pad * cap * 8));
  VFI_TRY(ws->cand_count.ensure(static_cast<size_t>(n_ctas) * nq_pad * 4));
  vfi::SmallParams p{};
  p.nq = nq;
  p.nq_pad = nq_pad;
  p.n_rows = static_cast<int>(idx->n);
  p.n_kblocks = nkb;
  p.n_tiles = n_tiles;
  p.keep = keep;
  p.cap = cap;
  p.cand = ws->cand.as<uint64_t>();
  p.cand_count = ws->cand_count.as<uint32_t>();
  p.tau_init = tau;
  const bool prof = idx->opt_profile != 0;
  if (prof) cudaEventRecord(ws->pev[0], st);
  vfi::dense_small_kernel<<<n_ctas, vfi::kSmThreads, vfi::dense_small_smem(nq_pad, nkb), st>>>(tq, td, p);
  LAUNCHED();
  if (prof) cudaEventRecord(ws->pev[1], st);
  VFI_CUDA(cudaGetLastError());
  *o_groups = n_ctas;
  *o_nq_pad = nq_pad;
  *o_cap = cap;
  return VFI_OK;
}

struct LaunchInfo {
  int path = 0, keep = 0;
  bool fused = false;
  bool pushed = false;   // the rescoring kernel pushed the batch's rows into the peers' exchange windows
};

// Enqueue one batch of <= kMaxQueriesPerLaunch queries already on the device (results to device buffers) in workspace
// `ws` without waiting: everything up to the copy of the certificate flag.  search_finish() waits and repairs.
int search_launch(vfi_index* idx, Workspace* ws, const void* q_dev, int q_dtype, int nq, int k, float* out_scores, int64_t* out_ids,
                  cudaStream_t st, bool no_hint, LaunchInfo* info, const vfi::PushTarget* push = nullptr) {
  ws->needs_check = false;
  ws->used_tau = false;
  ws->profiled = false;
  ws->q = q_dev;
  ws->q_dtype = q_dtype;
  ws->nq = nq;
  ws->k = k;
  ws->o_scores = out_scores;
  ws->o_ids = out_ids;
  ws->st = st;
  trace_mark(ws, 0, st);
  VFI_TRY(ws->flag.ensure(static_cast<size_t>(kMaxQueriesPerLaunch + 2) * 4));
  // [0] number of queries whose certificate failed, [1..1024] those queries, [1025] finished CTAs of the last tail kernel;
  // both counters are zeroed by the query preparation kernel, and the count reaches the host without a copy operation
  // (publish_flag_count)
  int* d_flag = ws->flag.as<int>();
  int* d_done = d_flag + kMaxQueriesPerLaunch + 1;
  VFI_TRY(prep_queries(idx, ws, q_dev, q_dtype, nq, st, d_flag, d_done));
  trace_mark(ws, 1, st);
  const int64_t n = idx->n;
  int keep = idx->opt_overfetch > 0 ? static_cast<int>(idx->opt_overfetch)
                                    : static_cast<int>(round_up(k + std::max(16, k / 4), 32));
  keep = std::max(keep, static_cast<int>(round_up(k, 32)));
  int path = static_cast<int>(idx->opt_force_path);
  if (path == 0) {
    if (n <= kExhaustiveRows || keep >= n) path = 1;
    else if (nq <= vfi::kGemvMaxQ && keep <= kGemvMaxKeep && gemv_ok(idx)) path = 3;
    else if (nq <= vfi::kExMaxQ) path = 1;    // few queries, deep k (the reference's online call: k = 2048 for 1-4 strings)
    else if (keep <= kFusedMaxKeep) path = 2;
    else path = 1;
  }
  if (path == 2 && (keep > kFusedMaxKeep || n == 0)) path = 1;
  if (path == 3 && (nq > vfi::kGemvMaxQ || keep > kGemvMaxKeep || n == 0 || !gemv_ok(idx))) path = (keep <= kFusedMaxKeep && n > 0) ? 2 : 1;
  info->path = path;
  info->keep = keep;
  info->fused = path != 1;
  ws->profiled = idx->opt_profile != 0;
  if (path == 1) {      // every row scored canonically: nothing to certify
    VFI_TRY(exact_pass(idx, ws, nullptr, nq, k, out_scores, out_ids, st, true));
    VFI_CUDA(cudaEventRecord(ws->done, st));
    return VFI_OK;
  }

  int n_groups = 0, nq_pad = 0, cap = 0;
  const float* tau = nullptr;
  if (path == 2 && idx->opt_tau_hint != 0 && !no_hint) {
    // Admission hint: tau = the m-th best score of a strided row sample (every s-th row).  The rows of the shard above tau
    // number m sampled ones plus a NegBinomial(m, 1/s) count of unsampled ones; (m, s) pairs are chosen so that fewer than
    // k' rows pass with probability <= 1e-7 per query (then the batch is simply redone without the hint).  The hint only
    // prunes work; exactness is re-established below.
    //   m = 8,  s = 2k'+1   : ~16 k' rows pass, the sample is 1/(2k') of the shard        (C3 on 1-2 GPUs)
    //   m = 16, s = 0.34 k' : ~5.5 k' rows pass, sample 3/k'                               (C2; C3 shards on 4-8 GPUs; the
    //                                                                                        chunk path of the hybrid)
    //   m = 32, s = k'/10   : ~3.2 k' rows pass, sample 10/k'                              (small shards, deep lists: the
    //                         title path of the hybrid retriever, 125k rows at k' = 256)
    // A passing row costs epilogue instructions (32 queries share a warp, so a warp takes the slow branch of a 4-column
    // group with probability ~128 f, f = passing fraction).  Measured on B200: K1 gets 5.5 % faster when f falls from
    // 0.16 % to 0.06 % (a 1/8 shard of C3: 2.05 -> 1.93 ms, 0.93 -> 0.98 of peak) and 5 % at C2; at f = 3 % the epilogue,
    // not the tensor pipe, bounds K1 (0.43 of peak on the title path with m = 8, 0.79 with m = 32).  The denser sample costs
    // 1/s of a K1 pass, so it pays when f8 = 16k'/N exceeds ~0.06 % (m = 16) / ~0.75 % (m = 32).
    int m = 8;
    int64_t rs = 2 * static_cast<int64_t>(keep) + 1;
    {
      const double f8 = 16.0 * keep / static_cast<double>(n);
      const int64_t s16 = std::max<int64_t>(2, static_cast<int64_t>(0.34 * keep));
      const int64_t s32 = std::max<int64_t>(2, keep / 10);
      int want = 8;
      if (f8 > 0.0075 && s32 >= 12) want = 32;
      else if (f8 > 0.0006) want = 16;
      if (idx->opt_tau_m == 8 || idx->opt_tau_m == 16 || idx->opt_tau_m == 32) want = static_cast<int>(idx->opt_tau_m);
      if (want == 16) { m = 16; rs = s16; }
      if (want == 32) { m = 32; rs = s32; }
    }
    const int64_t rr = n / rs;
    if (rr >= 4 * m || idx->opt_tau_hint == 2) {
      // The sample pass keeps one value per (query, 32 sampled rows) — the chunk's largest tensor-core score — instead of
      // every score: the m-th largest chunk maximum is at most the m-th largest sampled score (equal unless two of the
      // best m samples share a chunk), so it is a valid, marginally looser hint, and the threshold kernel reads 32 x less.
      // VFI_OPT_TAU_HINT = 3 keeps the full sample (every score stored) for comparison.
      // (with few sampled rows per query the chunk maxima collide: every score is kept instead)
      const bool full = idx->opt_tau_hint == 3 || ceil_div(rr, 32) < 8 * m;
      const int64_t ld = full ? ceil_div(rr, vfi::kBN) * vfi::kBN : ceil_div(rr, vfi::kBN) * (vfi::kBN / 32);
      const int64_t n_vals = full ? rr : ceil_div(rr, 32);
      const int64_t nq_pad_s = round_up(ceil_div(nq, vfi::kBM), 2) * vfi::kBM;
      VFI_TRY(ws->dbg.ensure(static_cast<size_t>(nq_pad_s) * ld * 4));
      VFI_TRY(ws->tau.ensure(static_cast<size_t>(nq) * 4));
      VFI_TRY(launch_fused(idx, ws, nq, 32, full ? vfi::MODE_STORE : vfi::MODE_CHUNKMAX, ws->dbg.as<float>(), ld, nullptr, nullptr,
                           nullptr, nullptr, st, rr, rs, false));
      trace_mark(ws, 3, st);
      vfi::tau_from_scores_kernel<<<static_cast<unsigned>(ceil_div(nq, vfi::kTauWarps)), vfi::kTauWarps * 32, 0, st>>>(
          ws->dbg.as<float>(), ld, static_cast<int>(n_vals), nq, m, ws->tau.as<float>(), idx->opt_tau_hint == 2 ? 1 : 0);
      LAUNCHED();
      VFI_CUDA(cudaGetLastError());
      tau = ws->tau.as<float>();
      trace_mark(ws, 4, st);
    }
  }
  if (path == 2 && small_ok(idx, nq)) {
    VFI_TRY(launch_small(idx, ws, nq, keep, tau, &n_groups, &nq_pad, &cap, st));
  } else if (path == 2) {
    VFI_TRY(launch_fused(idx, ws, nq, keep, vfi::MODE_TOPK, nullptr, 0, tau, &n_groups, &nq_pad, &cap, st));
  } else {
    // streaming scorer: per-CTA shared key buffers
    int cap_s = 1;
    while (cap_s < keep + vfi::kGemvRowsPerRound) cap_s <<= 1;
    const size_t smem = ((static_cast<size_t>(nq) * idx->dp * 4 + 15) & ~size_t(15)) + static_cast<size_t>(nq) * cap_s * 8 + 256;
    if (smem > 160 * 1024) return fail(VFI_ERR_UNSUPPORTED, "streaming scorer: query block does not fit shared memory");
    int ctas_per_sm = 1;
    if (idx->store == VFI_STORE_F32) dispatch_nq<GemvOcc, float>(nq, smem, &ctas_per_sm);
    else dispatch_nq<GemvOcc, uint16_t>(nq, smem, &ctas_per_sm);
    n_groups = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(idx->num_sms) * ctas_per_sm,
                                                  ceil_div(n, vfi::kGemvRowsPerRound)));
    nq_pad = nq;
    cap = keep;
    VFI_TRY(ws->cand.ensure(static_cast<size_t>(n_groups) * nq_pad * cap * 8));
    VFI_TRY(ws->cand_count.ensure(static_cast<size_t>(n_groups) * nq_pad * 4));
    const bool prof = idx->opt_profile != 0;
    if (prof) cudaEventRecord(ws->pev[0], st);
    if (idx->store == VFI_STORE_F32)
      dispatch_nq<GemvLaunch, float>(nq, n_groups, smem, st, idx->master, static_cast<int64_t>(idx->dp), idx->dp, n, ws->qcanon.as<float>(),
                                     keep, cap_s, ws->cand.as<uint64_t>(), ws->cand_count.as<uint32_t>(), nq_pad, cap);
    else
      dispatch_nq<GemvLaunch, uint16_t>(nq, n_groups, smem, st, idx->g, idx->kp, idx->dp, n, ws->qcanon.as<float>(), keep, cap_s,
                                        ws->cand.as<uint64_t>(), ws->cand_count.as<uint32_t>(), nq_pad, cap);
    LAUNCHED();
    if (prof) cudaEventRecord(ws->pev[1], st);
    VFI_CUDA(cudaGetLastError());
  }
  trace_mark(ws, 5, st);
  const bool tail_prof = idx->opt_profile != 0;
  if (tail_prof) cudaEventRecord(ws->pev[2], st);
  VFI_TRY(ws->keys.ensure(static_cast<size_t>(nq) * keep * 8));
  VFI_TRY(ws->keys_n.ensure(static_cast<size_t>(nq) * 4));
  VFI_TRY(ws->bound.ensure(static_cast<size_t>(nq) * 4));
  // K1c: per-query union of the group buffers -> k' best by tensor-core score.  About 16 k' keys reach this kernel per
  // query: up to k' = 128 they fit the light 1024-key selection buffer's fast paths (seven CTAs per SM); above that the
  // 4096-key buffer avoids the radix walk over global memory
  if (keep <= 128)
    vfi::cand_reduce_kernel<vfi::CandSmem><<<nq, 256, sizeof(vfi::CandSmem), st>>>(
        ws->cand.as<uint64_t>(), ws->cand_count.as<uint32_t>(), n_groups, nq_pad, cap, keep, tau, ws->keys.as<uint64_t>(),
        ws->keys_n.as<uint32_t>(), ws->bound.as<float>());
  else
    vfi::cand_reduce_kernel<vfi::SelectSmem><<<nq, 256, sizeof(vfi::SelectSmem), st>>>(
        ws->cand.as<uint64_t>(), ws->cand_count.as<uint32_t>(), n_groups, nq_pad, cap, keep, tau, ws->keys.as<uint64_t>(),
        ws->keys_n.as<uint32_t>(), ws->bound.as<float>());
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  trace_mark(ws, 6, st);
  if (keep <= 256 && idx->dp <= vfi::kRfMaxDp) {
    // K2: one thread per candidate: rescoring + final order + certificate (+ the push of a sharded batch's rows)
    vfi::PushTarget pt{};
    if (push != nullptr) {
      pt = *push;
      info->pushed = true;
    }
    const int threads = static_cast<int>(round_up(keep, 32));
    // Bytes per lane and step of the row gather.  256-byte pieces reach the higher gather bandwidth, 128-byte pieces halve
    // the shared memory per CTA (four CTAs of 256 threads per SM instead of two).  Measured (round 2, A/B in one process,
    // profiles/r2_exp_rescore_piece.jsonl): for depth-200 lists (k' = 256) and 1024 queries the tail is 5-8 % faster with
    // 128-byte pieces; at k' = 128 the two are within 1-5 % of each other live (and the 256-byte form is the faster one when
    // the kernel is timed alone under ncu); a grid that is a fraction of a wave (256 queries) is 25 % slower with 128.
    int piece = static_cast<int>(idx->opt_tail_piece);
    if (piece == 0) {
      const size_t smem256 = vfi::rescore_bulk_smem<256>(static_cast<int>(idx->dp), threads, 1);
      const int occ256 = static_cast<int>(std::max<size_t>(1, std::min<size_t>(228 * 1024 / (smem256 + 1024), static_cast<size_t>(2048 / threads))));
      piece = (threads > 128 && nq > idx->num_sms * occ256) ? 128 : 256;
    }
    const size_t smem = piece == 128 ? vfi::rescore_bulk_smem<128>(static_cast<int>(idx->dp), threads, 1)
                                     : vfi::rescore_bulk_smem<256>(static_cast<int>(idx->dp), threads, 1);
#define VFI_RESCORE(RowT, PIECE, ROWS, PITCH)                                                                                 \
    vfi::rescore_finalize_kernel<RowT, PIECE, 1><<<nq, threads, smem, st>>>(                                                   \
        ws->keys.as<uint64_t>(), ws->keys_n.as<uint32_t>(), ws->bound.as<float>(), keep, ROWS, PITCH,                          \
        static_cast<int>(idx->dp), ws->qcanon.as<float>(), k, idx->id_offset, ws->eps.as<float>(), out_scores, out_ids,        \
        d_flag + 1, d_flag, idx->d_max_err, d_done, ws->h_flag_dev, pt)
    if (idx->store == VFI_STORE_F32) {
      if (piece == 128) VFI_RESCORE(float, 128, idx->master, static_cast<int64_t>(idx->dp));
      else VFI_RESCORE(float, 256, idx->master, static_cast<int64_t>(idx->dp));
    } else {
      if (piece == 128) VFI_RESCORE(uint16_t, 128, idx->g, idx->kp);
      else VFI_RESCORE(uint16_t, 256, idx->g, idx->kp);
    }
#undef VFI_RESCORE
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
  } else {
    // K2a: canonical rescoring of the candidates, K2b: final order + certificate
    VFI_TRY(ws->keys2.ensure(static_cast<size_t>(nq) * keep * 8));
    dim3 grid(static_cast<unsigned>(ceil_div(keep, 128)), static_cast<unsigned>(nq));
    if (idx->store == VFI_STORE_F32)
      vfi::canon_score_kernel<float><<<grid, 128, 0, st>>>(idx->master, idx->dp, idx->dp, ws->qcanon.as<float>(), nullptr,
                                                           ws->keys.as<uint64_t>(), ws->keys_n.as<uint32_t>(), keep, 0,
                                                           ws->keys2.as<uint64_t>(), keep, idx->d_max_err);
    else
      vfi::canon_score_kernel<uint16_t><<<grid, 128, 0, st>>>(idx->g, idx->kp, idx->dp, ws->qcanon.as<float>(), nullptr,
                                                              ws->keys.as<uint64_t>(), ws->keys_n.as<uint32_t>(), keep, 0,
                                                              ws->keys2.as<uint64_t>(), keep, idx->d_max_err);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    vfi::finalize_kernel<1><<<nq, 256, sizeof(vfi::SelectSmem), st>>>(ws->keys2.as<uint64_t>(), keep, keep, nullptr,
                                                                     ws->keys_n.as<uint32_t>(), k, idx->id_offset, ws->bound.as<float>(),
                                                                     ws->eps.as<float>(), out_scores, out_ids, d_flag + 1, d_flag,
                                                                     d_done, ws->h_flag_dev);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
  }
  trace_mark(ws, 7, st);
  if (tail_prof) cudaEventRecord(ws->pev[3], st);
  trace_mark(ws, 8, st);
  VFI_CUDA(cudaEventRecord(ws->done, st));
  ws->needs_check = true;
  ws->used_tau = tau != nullptr;
  return VFI_OK;
}

void note_launch(vfi_index* idx, const LaunchInfo& info) {
  std::lock_guard<std::mutex> lock(idx->pool_mu);
  idx->stats.last_path = info.path;
  idx->stats.last_overfetch = info.keep;
  idx->stats.fused_launches++;
}

// Wait for the batch in `ws`, read its certificate flag and repair what failed: a batch pruned too hard by the admission
// hint is redone without it, queries whose candidates tie across the cut are re-run by the exact streaming pass.
// whole_stream: wait for everything enqueued on the batch's stream (a host-buffer call has its result copies queued behind the
// batch: one host wake-up instead of two); *repaired: the results were rewritten after the first wait.
int search_finish(vfi_index* idx, Workspace* ws, bool whole_stream = false, bool* repaired = nullptr) {
  if (repaired) *repaired = false;
  if (whole_stream) VFI_CUDA(cudaStreamSynchronize(ws->st));
  else VFI_CUDA(cudaEventSynchronize(ws->done));
  if (g_trace_steps && ws->needs_check && ws->tev[8]) {
    static const char* names[9] = {"start", "prep", "-", "sample", "tau", "K1", "cand_reduce", "rescore", "-"};
    float t0 = 0.f;
    cudaEventElapsedTime(&t0, g_trace_epoch, ws->tev[0]);
    std::string line = "[vfi steps] start at " + std::to_string(t0) + " ms:";
    cudaEvent_t prev = ws->tev[0];
    for (int i = 1; i < 9; ++i) {
      if (!ws->tev[i]) continue;
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, prev, ws->tev[i]) != cudaSuccess) { cudaGetLastError(); continue; }
      char buf[64];
      std::snprintf(buf, sizeof buf, " %s %.1f us;", names[i], ms * 1e3f);
      line += buf;
      prev = ws->tev[i];
    }
    float tot = 0.f;
    cudaEventElapsedTime(&tot, ws->tev[0], ws->tev[8]);
    std::fprintf(stderr, "%s total %.1f us\n", line.c_str(), tot * 1e3f);
  }
  if (ws->profiled) {
    float ms_k = 0.f, ms_t = 0.f;
    const bool ok_k = cudaEventElapsedTime(&ms_k, ws->pev[0], ws->pev[1]) == cudaSuccess;
    const bool ok_t = ws->needs_check && cudaEventElapsedTime(&ms_t, ws->pev[2], ws->pev[3]) == cudaSuccess;
    cudaGetLastError();
    std::lock_guard<std::mutex> lock(idx->pool_mu);
    if (ok_k) { idx->stats.fused_ms_total += ms_k; idx->stats.fused_ms_samples++; }
    if (ok_t) { idx->stats.tail_ms_total += ms_t; idx->stats.tail_ms_samples++; }
  }
  if (!ws->needs_check) return VFI_OK;
  const int n_flagged = ws->h_flag[0];
  if (n_flagged <= 0) return VFI_OK;
  if (ws->used_tau && (idx->opt_tau_hint == 1 || idx->opt_tau_hint == 3)) {
    // the hint pruned too much for some query: redo the batch without it (same kernels, no pruning)
    {
      std::lock_guard<std::mutex> lock(idx->pool_mu);
      idx->stats.hint_retries++;
    }
    LaunchInfo info;
    if (repaired) *repaired = true;
    VFI_TRY(search_launch(idx, ws, ws->q, ws->q_dtype, ws->nq, ws->k, ws->o_scores, ws->o_ids, ws->st, true, &info));
    const int rc = search_finish(idx, ws);
    if (repaired) *repaired = true;
    return rc;
  }
  {
    std::lock_guard<std::mutex> lock(idx->pool_mu);
    idx->stats.retried_queries += n_flagged;
  }
  // the prepared queries of this batch are still in its workspace
  if (repaired) *repaired = true;
  VFI_TRY(exact_pass(idx, ws, ws->flag.as<int>() + 1, n_flagged, ws->k, ws->o_scores, ws->o_ids, ws->st, false));
  VFI_CUDA(cudaStreamSynchronize(ws->st));
  return VFI_OK;
}

}  // namespace

extern "C" {

int vfi_index_search(vfi_index_t* idx, const float* q, int64_t nq, int k, float* out_scores, int64_t* out_ids, int mem,
                     void* stream) {
  return vfi_index_search_ex(idx, q, VFI_DTYPE_F32, nq, k, out_scores, out_ids, mem, stream);
}

int vfi_index_search_ex(vfi_index_t* idx, const void* q_any, int q_dtype, int64_t nq, int k, float* out_scores, int64_t* out_ids,
                        int mem, void* stream) {
  if (!idx || (nq > 0 && (!q_any || !out_scores || !out_ids)) || nq < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_search");
  if (q_dtype != VFI_DTYPE_F32 && q_dtype != VFI_DTYPE_BF16) return fail(VFI_ERR_INVALID, "q_dtype must be VFI_DTYPE_F32 or VFI_DTYPE_BF16");
  const size_t qsz = q_dtype == VFI_DTYPE_BF16 ? 2 : 4;
  const uint8_t* q = static_cast<const uint8_t*>(q_any);
  if (k <= 0) return fail(VFI_ERR_INVALID, "k must be positive");
  if (k > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k exceeds VFI_MAX_K (2048)");
  if (nq == 0) return VFI_OK;
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  Workspace* ws = acquire_ws(idx);
  if (!ws) return VFI_ERR_NOMEM;
  // host buffers and no stream given: the workspace's own stream, so concurrent callers do not queue up on the
  // legacy default stream
  cudaStream_t st = (stream == nullptr && mem == VFI_MEM_HOST) ? ws->own : static_cast<cudaStream_t>(stream);
  {
    std::lock_guard<std::mutex> slock(idx->pool_mu);
    idx->stats.searches++;
    idx->stats.queries += nq;
  }
  // VFI_TRACE_HOST=1: wall-clock phases of the host-buffer call on stderr (debugging aid for the e2e number)
  static const bool trace = std::getenv("VFI_TRACE_HOST") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_in = trace ? now() : 0.0, t_launch = 0.0, t_finish = 0.0;
  int rc = VFI_OK;
  for (int64_t q0 = 0; q0 < nq && rc == VFI_OK; q0 += kMaxQueriesPerLaunch) {
    const int nb = static_cast<int>(std::min<int64_t>(kMaxQueriesPerLaunch, nq - q0));
    const void* qd = q + static_cast<size_t>(q0) * idx->d * qsz;
    float* os = out_scores + q0 * k;
    int64_t* oi = out_ids + q0 * k;
    auto body = [&]() -> int {
      if (mem == VFI_MEM_HOST) {
        VFI_TRY(ws->qin.ensure(static_cast<size_t>(nb) * idx->d * qsz));
        VFI_TRY(ws->out_scores.ensure(static_cast<size_t>(nb) * k * 4));
        VFI_TRY(ws->out_ids.ensure(static_cast<size_t>(nb) * k * 8));
        VFI_CUDA(cudaMemcpyAsync(ws->qin.p, qd, static_cast<size_t>(nb) * idx->d * qsz, cudaMemcpyHostToDevice, st));
        qd = ws->qin.p;
        os = ws->out_scores.as<float>();
        oi = ws->out_ids.as<int64_t>();
      }
      LaunchInfo info;
      VFI_TRY(search_launch(idx, ws, qd, q_dtype, nb, k, os, oi, st, false, &info));
      note_launch(idx, info);
      if (trace) t_launch = now();
      if (mem == VFI_MEM_HOST) {
        // The result copies are queued right behind the batch and the host waits once for all of it; the certificate almost
        // never fails, and when it does the repaired results are copied again.
        auto copy_out = [&]() -> int {
          VFI_CUDA(cudaMemcpyAsync(out_scores + q0 * k, os, static_cast<size_t>(nb) * k * 4, cudaMemcpyDeviceToHost, st));
          VFI_CUDA(cudaMemcpyAsync(out_ids + q0 * k, oi, static_cast<size_t>(nb) * k * 8, cudaMemcpyDeviceToHost, st));
          return VFI_OK;
        };
        VFI_TRY(copy_out());
        bool repaired = false;
        VFI_TRY(search_finish(idx, ws, true, &repaired));
        if (trace) t_finish = now();
        if (repaired) {
          VFI_TRY(copy_out());
          VFI_CUDA(cudaStreamSynchronize(st));
        }
      } else {
        VFI_TRY(search_finish(idx, ws));
        if (trace) t_finish = now();
      }
      return VFI_OK;
    };
    rc = body();
  }
  if (rc != VFI_OK) cudaStreamSynchronize(st);   // nothing of this call may still be using the workspace
  if (trace)
    std::fprintf(stderr, "[vfi trace] search nq=%lld: enqueue %.3f ms, wait for the batch %.3f ms, results out %.3f ms\n",
                 static_cast<long long>(nq), t_launch - t_in, t_finish - t_launch, now() - t_finish);
  release_ws(idx, ws);
  return rc;
}

int vfi_index_search_begin(vfi_index_t* idx, const float* q, int64_t nq, int k, float* out_scores, int64_t* out_ids, void* stream,
                           int* ticket) {
  return vfi_index_search_begin_ex(idx, q, VFI_DTYPE_F32, nq, k, out_scores, out_ids, stream, ticket);
}

int vfi_index_search_begin_ex(vfi_index_t* idx, const void* q, int q_dtype, int64_t nq, int k, float* out_scores, int64_t* out_ids,
                              void* stream, int* ticket) {
  if (!idx || !ticket || !q || !out_scores || !out_ids || nq <= 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_search_begin");
  if (q_dtype != VFI_DTYPE_F32 && q_dtype != VFI_DTYPE_BF16) return fail(VFI_ERR_INVALID, "q_dtype must be VFI_DTYPE_F32 or VFI_DTYPE_BF16");
  if (k <= 0) return fail(VFI_ERR_INVALID, "k must be positive");
  if (k > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k exceeds VFI_MAX_K (2048)");
  if (nq > kMaxQueriesPerLaunch) return fail(VFI_ERR_UNSUPPORTED, "vfi_index_search_begin takes at most 1024 queries per batch");
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  Workspace* ws = acquire_ws(idx);
  if (!ws) return VFI_ERR_NOMEM;
  LaunchInfo info;
  const int rc = search_launch(idx, ws, q, q_dtype, static_cast<int>(nq), k, out_scores, out_ids, static_cast<cudaStream_t>(stream), false, &info);
  if (rc != VFI_OK) {
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    release_ws(idx, ws);
    return rc;
  }
  note_launch(idx, info);
  std::lock_guard<std::mutex> slock(idx->pool_mu);
  idx->stats.searches++;
  idx->stats.queries += nq;
  const int t = idx->next_ticket;
  idx->next_ticket = (idx->next_ticket == 0x7FFFFFFF) ? 1 : idx->next_ticket + 1;
  idx->tickets[t] = ws;
  ws->ticket = t;
  *ticket = t;
  return VFI_OK;
}

int vfi_index_search_begin_push(vfi_index_t* idx, const void* q, int q_dtype, int64_t nq, int k, float* out_scores, int64_t* out_ids,
                                vfi_exchange_t* ex, void* stream, int* ticket) {
  if (!idx || !ticket || !q || !out_scores || !out_ids || !ex || nq <= 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_search_begin_push");
  if (q_dtype != VFI_DTYPE_F32 && q_dtype != VFI_DTYPE_BF16) return fail(VFI_ERR_INVALID, "q_dtype must be VFI_DTYPE_F32 or VFI_DTYPE_BF16");
  if (k <= 0) return fail(VFI_ERR_INVALID, "k must be positive");
  if (k > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k exceeds VFI_MAX_K (2048)");
  if (nq > kMaxQueriesPerLaunch) return fail(VFI_ERR_UNSUPPORTED, "vfi_index_search_begin takes at most 1024 queries per batch");
  if (ex->device != idx->device) return fail(VFI_ERR_INVALID, "the exchange and the index live on different devices");
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  Workspace* ws = acquire_ws(idx);
  if (!ws) return VFI_ERR_NOMEM;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the epoch is consumed on every rank whatever path the local search takes: rows that the rescoring kernel does not push
  // (exact streaming path, k' > 256) are pushed by a kernel of their own right behind the search
  vfi::PushTarget target{};
  int rc = exchange_reserve_push(ex, nq, k, &target);
  if (rc != VFI_OK) {
    release_ws(idx, ws);
    return rc;
  }
  LaunchInfo info;
  rc = search_launch(idx, ws, q, q_dtype, static_cast<int>(nq), k, out_scores, out_ids, st, false, &info, &target);
  if (rc == VFI_OK && !info.pushed)
    rc = exchange_push_rows(ex, target, out_scores, out_ids, nq, k, ws->needs_check ? ws->flag.as<int>() : nullptr, st);
  if (rc != VFI_OK) {
    cudaStreamSynchronize(st);
    release_ws(idx, ws);
    return rc;     // NOTE: the epoch stays reserved; the peers' merge of it will time out — a failed collective
  }
  note_launch(idx, info);
  std::lock_guard<std::mutex> slock(idx->pool_mu);
  idx->stats.searches++;
  idx->stats.queries += nq;
  const int t = idx->next_ticket;
  idx->next_ticket = (idx->next_ticket == 0x7FFFFFFF) ? 1 : idx->next_ticket + 1;
  idx->tickets[t] = ws;
  ws->ticket = t;
  *ticket = t;
  return VFI_OK;
}

int vfi_index_ticket_flag(vfi_index_t* idx, int ticket, const int** device_flag) {
  if (!idx || ticket < 0 || !device_flag) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_ticket_flag");
  std::lock_guard<std::mutex> slock(idx->pool_mu);
  auto it = idx->tickets.find(ticket);
  if (it == idx->tickets.end()) return fail(VFI_ERR_INVALID, "no batch in flight for this ticket");
  *device_flag = it->second->flag.as<int>();
  return VFI_OK;
}

int vfi_index_search_finish(vfi_index_t* idx, int ticket) {
  if (!idx || ticket < 0) return fail(VFI_ERR_INVALID, "bad ticket");
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  Workspace* ws = nullptr;
  {
    std::lock_guard<std::mutex> slock(idx->pool_mu);
    auto it = idx->tickets.find(ticket);
    if (it != idx->tickets.end()) {
      ws = it->second;
      idx->tickets.erase(it);
    }
  }
  if (!ws) return fail(VFI_ERR_INVALID, "no batch in flight for this ticket");
  const int rc = search_finish(idx, ws);
  if (rc != VFI_OK) cudaStreamSynchronize(ws->st);
  release_ws(idx, ws);
  return rc;
}

int vfi_index_read_rows(vfi_index_t* idx, int64_t first, int64_t n, float* out, int mem, void* stream) {
  if (!idx || (n > 0 && !out) || first < 0 || n < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_read_rows");
  if (first + n > idx->n) return fail(VFI_ERR_INVALID, "rows out of range");
  if (n == 0) return VFI_OK;
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  std::lock_guard<std::mutex> slock(idx->stage_mu);
  DeviceGuard guard(idx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunk = std::max<int64_t>(1, (static_cast<int64_t>(256) << 20) / (static_cast<int64_t>(idx->d) * 4));
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t rows = std::min(chunk, n - r0);
    float* dst = out + r0 * idx->d;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(idx->stage.ensure(static_cast<size_t>(rows) * idx->d * 4));
      dst = idx->stage.as<float>();
    }
    const unsigned blocks = static_cast<unsigned>(ceil_div(rows * idx->d, 256));
    if (idx->store == VFI_STORE_F32)
      vfi::rows_to_f32_kernel<float><<<blocks, 256, 0, st>>>(idx->master, idx->dp, first + r0, rows, idx->d, dst);
    else
      vfi::rows_to_f32_kernel<uint16_t><<<blocks, 256, 0, st>>>(idx->g, idx->kp, first + r0, rows, idx->d, dst);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    if (mem == VFI_MEM_HOST) {
      VFI_CUDA(cudaMemcpyAsync(out + r0 * idx->d, dst, static_cast<size_t>(rows) * idx->d * 4, cudaMemcpyDeviceToHost, st));
      VFI_CUDA(cudaStreamSynchronize(st));
    }
  }
  VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
}

int vfi_index_pairwise(vfi_index_t* idx, const int64_t* ids, int n, float* out, int mem, void* stream) {
  if (!idx || n < 0 || (n > 0 && (!ids || !out))) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_pairwise");
  if (n > 1024) return fail(VFI_ERR_UNSUPPORTED, "pairwise supports up to 1024 rows");
  if (n == 0) return VFI_OK;
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  Workspace* ws = acquire_ws(idx);
  if (!ws) return VFI_ERR_NOMEM;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto body = [&]() -> int {
    std::vector<int64_t> hid(n);
    const int64_t* dids = ids;
    if (mem == VFI_MEM_DEVICE) VFI_CUDA(cudaMemcpyAsync(hid.data(), ids, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
    else std::memcpy(hid.data(), ids, sizeof(int64_t) * n);
    VFI_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i)
      if (hid[i] < 0 || hid[i] >= idx->n) return fail(VFI_ERR_INVALID, "pairwise: row id out of range");
    VFI_TRY(ws->sel.ensure(static_cast<size_t>(n) * 8));
    if (mem == VFI_MEM_HOST) {
      VFI_CUDA(cudaMemcpyAsync(ws->sel.p, hid.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
      dids = ws->sel.as<int64_t>();
    }
    VFI_TRY(ws->qcanon.ensure(static_cast<size_t>(n) * idx->dp * 4));
    VFI_TRY(ws->keys.ensure(static_cast<size_t>(n) * n * 8));
    VFI_TRY(ws->keys2.ensure(static_cast<size_t>(n) * n * 8));
    VFI_TRY(ws->keys_n.ensure(static_cast<size_t>(n) * 4));
    const int64_t work = std::max<int64_t>(static_cast<int64_t>(n) * idx->dp, static_cast<int64_t>(n) * n);
    const unsigned blocks = static_cast<unsigned>(ceil_div(work, 256));
    dim3 grid(static_cast<unsigned>(ceil_div(n, 128)), static_cast<unsigned>(n));
    if (idx->store == VFI_STORE_F32) {
      vfi::gather_rows_kernel<float><<<blocks, 256, 0, st>>>(idx->master, idx->dp, dids, n, idx->dp, ws->qcanon.as<float>(),
                                                             ws->keys.as<uint64_t>(), ws->keys_n.as<uint32_t>());
      vfi::canon_score_kernel<float><<<grid, 128, 0, st>>>(idx->master, idx->dp, idx->dp, ws->qcanon.as<float>(), nullptr,
                                                           ws->keys.as<uint64_t>(), ws->keys_n.as<uint32_t>(), n, 0,
                                                           ws->keys2.as<uint64_t>(), n, nullptr);
    } else {
      vfi::gather_rows_kernel<uint16_t><<<blocks, 256, 0, st>>>(idx->g, idx->kp, dids, n, idx->dp, ws->qcanon.as<float>(),
                                                                ws->keys.as<uint64_t>(), ws->keys_n.as<uint32_t>());
      vfi::canon_score_kernel<uint16_t><<<grid, 128, 0, st>>>(idx->g, idx->kp, idx->dp, ws->qcanon.as<float>(), nullptr,
                                                              ws->keys.as<uint64_t>(), ws->keys_n.as<uint32_t>(), n, 0,
                                                              ws->keys2.as<uint64_t>(), n, nullptr);
    }
    LAUNCHED();
    LAUNCHED();
    float* dout = out;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(ws->dbg.ensure(static_cast<size_t>(n) * n * 4));
      dout = ws->dbg.as<float>();
    }
    vfi::keys_to_scores_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(n) * n, 256)), 256, 0, st>>>(
        ws->keys2.as<uint64_t>(), static_cast<int64_t>(n) * n, dout);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    if (mem == VFI_MEM_HOST) VFI_CUDA(cudaMemcpyAsync(out, dout, static_cast<size_t>(n) * n * 4, cudaMemcpyDeviceToHost, st));
    VFI_CUDA(cudaStreamSynchronize(st));
    return VFI_OK;
  };
  const int rc = body();
  if (rc != VFI_OK) cudaStreamSynchronize(st);
  release_ws(idx, ws);
  return rc;
}

int vfi_index_debug_scores(vfi_index_t* idx, const float* q, int64_t nq, float* out, int mem, void* stream) {
  if (!idx || !q || !out || nq <= 0 || nq > kMaxQueriesPerLaunch) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_debug_scores");
  if (idx->n == 0) return VFI_OK;
  std::shared_lock<std::shared_mutex> lock(idx->rw);
  DeviceGuard guard(idx->device);
  Workspace* ws = acquire_ws(idx);
  if (!ws) return VFI_ERR_NOMEM;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto body = [&]() -> int {
    const float* qd = q;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(ws->qin.ensure(static_cast<size_t>(nq) * idx->d * 4));
      VFI_CUDA(cudaMemcpyAsync(ws->qin.p, q, static_cast<size_t>(nq) * idx->d * 4, cudaMemcpyHostToDevice, st));
      qd = ws->qin.as<float>();
    }
    VFI_TRY(prep_queries(idx, ws, qd, VFI_DTYPE_F32, static_cast<int>(nq), st));
    const int64_t ld = ceil_div(idx->n, vfi::kBN) * vfi::kBN;
    const int64_t nq_pad = round_up(ceil_div(nq, vfi::kBM), 2) * vfi::kBM;
    VFI_TRY(ws->dbg.ensure(static_cast<size_t>(nq_pad) * ld * 4));
    VFI_TRY(launch_fused(idx, ws, static_cast<int>(nq), 32, vfi::MODE_STORE, ws->dbg.as<float>(), ld, nullptr, nullptr, nullptr, nullptr,
                         st, -1, 1, false));
    VFI_CUDA(cudaMemcpy2DAsync(out, static_cast<size_t>(idx->n) * 4, ws->dbg.p, static_cast<size_t>(ld) * 4,
                               static_cast<size_t>(idx->n) * 4, static_cast<size_t>(nq),
                               mem == VFI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    VFI_CUDA(cudaStreamSynchronize(st));
    return VFI_OK;
  };
  const int rc = body();
  if (rc != VFI_OK) cudaStreamSynchronize(st);
  release_ws(idx, ws);
  return rc;
}

int vfi_normalize_l2(float* x, int64_t n, int d, int mem, int device, void* stream) {
  if ((!x && n > 0) || n < 0 || d <= 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_normalize_l2");
  if (n == 0) return VFI_OK;
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunk = std::max<int64_t>(1, (static_cast<int64_t>(256) << 20) / (static_cast<int64_t>(d) * 4));
  StageArena* arena = nullptr;
  float* stage = nullptr;
  if (mem == VFI_MEM_HOST) {
    arena = borrow_arena(device);
    const int rc = arena->buf.ensure(static_cast<size_t>(std::min(chunk, n)) * d * 4);
    if (rc != VFI_OK) {
      return_arena(arena);
      return rc;
    }
    stage = arena->buf.as<float>();
  }
  int rc = VFI_OK;
  for (int64_t r0 = 0; r0 < n && rc == VFI_OK; r0 += chunk) {
    const int64_t rows = std::min(chunk, n - r0);
    float* p = x + r0 * d;
    float* dp = p;
    if (mem == VFI_MEM_HOST) {
      if (cudaMemcpyAsync(stage, p, static_cast<size_t>(rows) * d * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = fail(VFI_ERR_CUDA, "H2D copy failed"); break; }
      dp = stage;
    }
    vfi::normalize_l2_kernel<<<static_cast<unsigned>(ceil_div(rows * 32, 256)), 256, 0, st>>>(dp, rows, d);
    LAUNCHED();
    if (mem == VFI_MEM_HOST) {
      if (cudaMemcpyAsync(p, stage, static_cast<size_t>(rows) * d * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = fail(VFI_ERR_CUDA, "D2H copy failed"); break; }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(VFI_ERR_CUDA, std::string("normalize_l2: ") + cudaGetErrorString(e));
  }
  if (arena) return_arena(arena);
  return rc;
}

// cosine top-k of the experiment scripts: normalise both sides, exact scores of every row, then the
// argsort()[-k:][::-1] order = (score desc, index DESC).  Implemented on top of a temporary
// F32 index by mapping index i -> n_c-1-i so that "higher index first" becomes "lower id first".
int vfi_cosine_topk(const float* e, int64_t n_e, const float* c, int64_t n_c, int d, int k, float* out_scores,
                    int64_t* out_ids, int mem, int device, void* stream) {
  if (!e || !c || !out_scores || !out_ids || n_e < 0 || n_c < 0 || d <= 0 || k <= 0)
    return fail(VFI_ERR_INVALID, "bad argument to vfi_cosine_topk");
  if (mem != VFI_MEM_HOST) return fail(VFI_ERR_UNSUPPORTED, "vfi_cosine_topk takes host buffers");
  if (n_e == 0) return VFI_OK;
  std::vector<float> cn(static_cast<size_t>(n_c) * d), en(e, e + static_cast<size_t>(n_e) * d);
  for (int64_t i = 0; i < n_c; ++i)
    std::memcpy(&cn[static_cast<size_t>(n_c - 1 - i) * d], c + static_cast<size_t>(i) * d, sizeof(float) * d);
  VFI_TRY(vfi_normalize_l2(cn.data(), n_c, d, VFI_MEM_HOST, device, stream));
  VFI_TRY(vfi_normalize_l2(en.data(), n_e, d, VFI_MEM_HOST, device, stream));
  // one scratch index per (device, d), kept between calls (the experiment scripts call this once per question): emptied and
  // refilled instead of created and destroyed
  static std::mutex cache_mu;
  static std::vector<std::pair<std::pair<int, int>, vfi_index_t*>> cache;
  std::lock_guard<std::mutex> cache_lock(cache_mu);
  vfi_index_t* idx = nullptr;
  for (auto& ent : cache)
    if (ent.first.first == device && ent.first.second == d) idx = ent.second;
  if (idx == nullptr) {
    VFI_TRY(vfi_index_create(d, VFI_STORE_F32, device, &idx));
    cache.push_back({{device, d}, idx});
  }
  {
    std::unique_lock<std::shared_mutex> lock(idx->rw);
    DeviceGuard guard(device);
    idx->n = 0;                                        // rows are overwritten by the add below
    cudaMemset(idx->xnorm_bits, 0, 4);
  }
  int rc = vfi_index_add(idx, cn.data(), n_c, VFI_MEM_HOST, stream);
  if (rc == VFI_OK) rc = vfi_index_search(idx, en.data(), n_e, k, out_scores, out_ids, VFI_MEM_HOST, stream);
  if (rc != VFI_OK) return rc;
  for (int64_t i = 0; i < n_e * k; ++i)
    if (out_ids[i] >= 0) out_ids[i] = n_c - 1 - out_ids[i];
  return VFI_OK;
}

}  // extern "C"
