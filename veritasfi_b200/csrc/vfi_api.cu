// vfi_api.cu — the C ABI declared in include/vfi.h: host orchestration of the CUDA kernels.
// One translation unit builds libvfi.so:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC vfi_api.cu
// There is no CPU implementation here; every compute entry needs an sm_100 device.
#include <cub/device/device_radix_sort.cuh>
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vfi.h"
#include "dense_fused.cuh"
#include "dense_support.cuh"
#include "sparse_fuse.cuh"
#include "peer_exchange.cuh"
#include "text_host.h"

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define VFI_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      return fail(e__ == cudaErrorMemoryAllocation ? VFI_ERR_NOMEM : VFI_ERR_CUDA,           \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                      \
    }                                                                                        \
  } while (0)
#define VFI_TRY(expr)              \
  do {                             \
    int s__ = (expr);              \
    if (s__ != VFI_OK) return s__; \
  } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return VFI_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = std::max(need, static_cast<size_t>(256));
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(VFI_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
    }
    bytes = want;
    return VFI_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    ok = (cudaSetDevice(dev) == cudaSuccess);
    if (!ok) cudaGetLastError();
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

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

// cudaGetDeviceProperties costs milliseconds: query each device once
int device_props(int device, cudaDeviceProp* prop) {
  static std::mutex mu;
  static std::vector<cudaDeviceProp> cache;
  static std::vector<char> have;
  static int n_dev = -1;
  std::lock_guard<std::mutex> lock(mu);
  if (n_dev < 0) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      return fail(VFI_ERR_NO_DEVICE, "no CUDA device: this library has no CPU implementation");
    }
    n_dev = n;
    cache.resize(n);
    have.assign(n, 0);
  }
  if (device < 0 || device >= n_dev) return fail(VFI_ERR_INVALID, "device index out of range");
  if (!have[device]) {
    VFI_CUDA(cudaGetDeviceProperties(&cache[device], device));
    have[device] = 1;
  }
  *prop = cache[device];
  if (prop->major != 10)
    return fail(VFI_ERR_NO_DEVICE, std::string("device '") + prop->name + "' is not sm_100 (kernels are built for sm_100a only)");
  return VFI_OK;
}

// the streaming scorer is instantiated per query count (1..8) so the accumulators stay in registers
template <typename RowT, int NQ>
void gemv_launch_one(int grid, size_t smem, cudaStream_t st, const RowT* rows, int64_t pitch, int dp, int64_t n, const float* q,
                     int keep, int cap_s, uint64_t* cand, uint32_t* cnt, int nq_pad, int cand_cap) {
  vfi::gemv_topk_kernel<RowT, NQ><<<grid, vfi::kGemvThreads, smem, st>>>(rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap);
}
template <typename RowT>
void gemv_launch(int nq, int grid, size_t smem, cudaStream_t st, const RowT* rows, int64_t pitch, int dp, int64_t n, const float* q,
                 int keep, int cap_s, uint64_t* cand, uint32_t* cnt, int nq_pad, int cand_cap) {
  switch (nq) {
    case 1: gemv_launch_one<RowT, 1>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
    case 2: gemv_launch_one<RowT, 2>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
    case 3: gemv_launch_one<RowT, 3>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
    case 4: gemv_launch_one<RowT, 4>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
    case 5: gemv_launch_one<RowT, 5>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
    case 6: gemv_launch_one<RowT, 6>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
    case 7: gemv_launch_one<RowT, 7>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
    default: gemv_launch_one<RowT, 8>(grid, smem, st, rows, pitch, dp, n, q, keep, cap_s, cand, cnt, nq_pad, cand_cap); break;
  }
}
template <typename RowT, int NQ>
int gemv_occ_one(size_t smem) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vfi::gemv_topk_kernel<RowT, NQ>, vfi::kGemvThreads, smem) != cudaSuccess) {
    cudaGetLastError();
    occ = 1;
  }
  return occ < 1 ? 1 : occ;
}
template <typename RowT>
int gemv_occupancy(int nq, size_t smem) {
  switch (nq) {
    case 1: return gemv_occ_one<RowT, 1>(smem);
    case 2: return gemv_occ_one<RowT, 2>(smem);
    case 3: return gemv_occ_one<RowT, 3>(smem);
    case 4: return gemv_occ_one<RowT, 4>(smem);
    case 5: return gemv_occ_one<RowT, 5>(smem);
    case 6: return gemv_occ_one<RowT, 6>(smem);
    case 7: return gemv_occ_one<RowT, 7>(smem);
    default: return gemv_occ_one<RowT, 8>(smem);
  }
}
template <typename RowT, int NQ>
void gemv_attr_one() {
  cudaFuncSetAttribute(vfi::gemv_topk_kernel<RowT, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}
void gemv_set_smem_attr() {
  gemv_attr_one<uint16_t, 1>(); gemv_attr_one<uint16_t, 2>(); gemv_attr_one<uint16_t, 3>(); gemv_attr_one<uint16_t, 4>();
  gemv_attr_one<uint16_t, 5>(); gemv_attr_one<uint16_t, 6>(); gemv_attr_one<uint16_t, 7>(); gemv_attr_one<uint16_t, 8>();
  gemv_attr_one<float, 1>(); gemv_attr_one<float, 2>(); gemv_attr_one<float, 3>(); gemv_attr_one<float, 4>();
  gemv_attr_one<float, 5>(); gemv_attr_one<float, 6>(); gemv_attr_one<float, 7>(); gemv_attr_one<float, 8>();
}

constexpr int kMaxQueriesPerLaunch = 1024;
constexpr int kExhaustiveRows = 4096;   // shards this small skip the tensor-core pass
constexpr int kFusedMaxKeep = 1024;
constexpr int kGemvMaxKeep = 512;

}  // namespace

// One batch between vfi_index_search_begin and vfi_index_search_finish (the synchronous search uses a slot too).
constexpr int kSearchSlots = 4;
struct SearchSlot {
  bool busy = false;         // launched, not finished yet
  bool needs_check = false;  // a certificate flag is on its way to h_flag[slot]
  bool used_tau = false;     // the admission hint was on (a failed certificate is first retried without it)
  const float* q = nullptr;
  int nq = 0, k = 0;
  float* out_scores = nullptr;
  int64_t* out_ids = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t done = nullptr;   // recorded behind the flag copy
  cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};   // VFI_OPT_PROFILE: around the dominant kernel [0,1] and the tail [2,3]
  uint64_t seq = 0;             // launch number (see vfi_index::prep_owner)
};

// =============================================================================================
struct vfi_index {
  int d = 0, dp = 0, store = VFI_STORE_BF16, device = 0;
  int64_t kp = 0;          // gemm operand row length (dp or 3*dp)
  int64_t n = 0, cap_rows = 0;
  int64_t id_offset = 0;
  int num_sms = 148;
  int max_pairs = 0;       // co-resident 2-CTA clusters of the pair kernel (queried once)
  uint16_t* g = nullptr;   // [cap_rows][kp] bf16 gemm operand rows
  float* master = nullptr; // [cap_rows][dp] fp32 rows (F32 store only)
  uint32_t* xnorm_bits = nullptr;
  // options
  int64_t opt_overfetch = 0, opt_force_path = 0, opt_profile = 0, opt_tau_hint = 1, opt_num_ctas = 0, opt_cluster = 0, opt_cta_pair = 0, opt_tail = 0;
  // workspace
  DevBuf w_qin, w_qcanon, w_qg, w_eps, w_cand, w_cand_count, w_keys, w_keys_n, w_bound, w_keys2, w_flag,
      w_out_scores, w_out_ids, w_stage, w_dbg, w_sel, w_tau;
  int* h_flag = nullptr;   // pinned: [slot] = number of queries of that batch whose certificate failed
  SearchSlot slots[kSearchSlots];
  uint64_t launch_seq = 0;
  uint64_t prep_owner = 0;   // seq of the batch whose prepared queries (w_qcanon, w_qg, w_eps) are in the workspace
  int cur_slot = 0;        // slot of the batch being launched / finished (under mu)
  vfi_search_stats stats{};
  uint32_t* d_max_err = nullptr;
  std::mutex mu;
};

extern "C" {

int vfi_abi_version(void) { return VFI_ABI_VERSION; }
const char* vfi_last_error(void) { return g_err.c_str(); }
int vfi_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int64_t vfi_launch_count(void) { return g_launches.load(); }

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
  if (cudaMallocHost(&idx->h_flag, sizeof(int) * (kMaxQueriesPerLaunch + 1)) != cudaSuccess)
    return bail(fail(VFI_ERR_NOMEM, "cudaMallocHost failed"));
  for (SearchSlot& sl : idx->slots) {
    cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
    for (cudaEvent_t& e : sl.pev) cudaEventCreate(&e);
  }
  cudaFuncSetAttribute(vfi::dense_fused_kernel<vfi::MODE_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kDenseSmemBytes);
  cudaFuncSetAttribute(vfi::dense_fused_kernel<vfi::MODE_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kDenseSmemBytes);
  cudaFuncSetAttribute(vfi::dense_fused_kernel<vfi::MODE_CHUNKMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kDenseSmemBytes);
  cudaFuncSetAttribute(vfi::dense_fused_kernel<vfi::MODE_TOPK>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
  cudaFuncSetAttribute(vfi::dense_fused_pair_kernel<vfi::MODE_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kPairSmemBytes);
  cudaFuncSetAttribute(vfi::dense_fused_pair_kernel<vfi::MODE_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kPairSmemBytes);
  cudaFuncSetAttribute(vfi::dense_fused_pair_kernel<vfi::MODE_CHUNKMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, vfi::kPairSmemBytes);
  gemv_set_smem_attr();
  cudaFuncSetAttribute(vfi::select_rescore_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(vfi::TailSmem)));
  cudaFuncSetAttribute(vfi::select_rescore_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(vfi::TailSmem)));
#define VFI_RF_ATTR(P, S)                                                                                                  \
  cudaFuncSetAttribute(vfi::rescore_finalize_kernel<uint16_t, P, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                       static_cast<int>(vfi::rescore_bulk_smem<P>(vfi::kRfMaxDp, 256, S)));                                 \
  cudaFuncSetAttribute(vfi::rescore_finalize_kernel<float, P, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                       static_cast<int>(vfi::rescore_bulk_smem<P>(vfi::kRfMaxDp, 256, S)))
  VFI_RF_ATTR(256, 1);
  VFI_RF_ATTR(128, 1);
  VFI_RF_ATTR(256, 2);
#undef VFI_RF_ATTR
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bail(fail(VFI_ERR_CUDA, std::string("kernel attribute setup: ") + cudaGetErrorString(e)));
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
  if (idx->h_flag) cudaFreeHost(idx->h_flag);
  for (SearchSlot& sl : idx->slots) {
    if (sl.done) cudaEventDestroy(sl.done);
    for (cudaEvent_t e : sl.pev)
      if (e) cudaEventDestroy(e);
  }
  for (DevBuf* b : {&idx->w_qin, &idx->w_qcanon, &idx->w_qg, &idx->w_eps, &idx->w_cand, &idx->w_cand_count, &idx->w_keys,
                    &idx->w_keys_n, &idx->w_bound, &idx->w_keys2, &idx->w_flag, &idx->w_out_scores, &idx->w_out_ids,
                    &idx->w_stage, &idx->w_dbg, &idx->w_sel, &idx->w_tau})
    b->release();
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
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  return grow_rows(idx, n, nullptr);
}

template <typename InT>
static int add_rows(vfi_index* idx, const InT* x, int64_t n, int mem, cudaStream_t st) {
  if (n == 0) return VFI_OK;
  if (idx->n + n >= 0x7FFFFF00ll) return fail(VFI_ERR_UNSUPPORTED, "a shard holds at most 2^31-256 rows");
  VFI_TRY(grow_rows(idx, idx->n + n, st));
  const int64_t chunk = std::max<int64_t>(1, (static_cast<int64_t>(256) << 20) / (static_cast<int64_t>(idx->d) * sizeof(InT)));
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t rows = std::min(chunk, n - r0);
    const InT* src = x + r0 * idx->d;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(idx->w_stage.ensure(static_cast<size_t>(rows) * idx->d * sizeof(InT)));
      VFI_CUDA(cudaMemcpyAsync(idx->w_stage.p, src, static_cast<size_t>(rows) * idx->d * sizeof(InT), cudaMemcpyHostToDevice, st));
      src = idx->w_stage.as<InT>();
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
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  return add_rows<float>(idx, x, n, mem, static_cast<cudaStream_t>(stream));
}

int vfi_index_add_bf16(vfi_index_t* idx, const uint16_t* x, int64_t n, int mem, void* stream) {
  if (!idx || (!x && n > 0) || n < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_add_bf16");
  if (idx->store != VFI_STORE_BF16) return fail(VFI_ERR_INVALID, "add_bf16 needs a VFI_STORE_BF16 index");
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  return add_rows<uint16_t>(idx, x, n, mem, static_cast<cudaStream_t>(stream));
}

int64_t vfi_index_ntotal(const vfi_index_t* idx) { return idx ? idx->n : 0; }
int vfi_index_dim(const vfi_index_t* idx) { return idx ? idx->d : 0; }

int vfi_index_set_id_offset(vfi_index_t* idx, int64_t offset) {
  if (!idx || offset < 0) return fail(VFI_ERR_INVALID, "bad id offset");
  idx->id_offset = offset;
  return VFI_OK;
}

int vfi_index_set_option(vfi_index_t* idx, int opt, int64_t value) {
  if (!idx) return fail(VFI_ERR_INVALID, "index is null");
  std::lock_guard<std::mutex> lock(idx->mu);
  switch (opt) {
    case VFI_OPT_OVERFETCH: idx->opt_overfetch = value; break;
    case VFI_OPT_FORCE_PATH: idx->opt_force_path = value; break;
    case VFI_OPT_PROFILE: idx->opt_profile = value; break;
    case VFI_OPT_TAU_HINT: idx->opt_tau_hint = value; break;
    case VFI_OPT_NUM_CTAS: idx->opt_num_ctas = value; break;
    case VFI_OPT_CTA_PAIR:
      if (value < 0 || value > 2) return fail(VFI_ERR_INVALID, "VFI_OPT_CTA_PAIR: 0 auto, 1 off, 2 on");
      idx->opt_cta_pair = value;
      break;
    case VFI_OPT_TAIL:
      if (value < 0 || value > 3) return fail(VFI_ERR_INVALID, "VFI_OPT_TAIL: 0 auto, 1 single-launch tail, 2 128-byte pieces, 3 two-stage ring");
      idx->opt_tail = value;
      break;
    case VFI_OPT_CLUSTER:
      if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return fail(VFI_ERR_INVALID, "cluster must be 1, 2, 4 or 8");
      idx->opt_cluster = value;
      break;
    default: return fail(VFI_ERR_INVALID, "unknown option");
  }
  return VFI_OK;
}

int vfi_index_get_stats(vfi_index_t* idx, vfi_search_stats* out, int reset) {
  if (!idx || !out) return fail(VFI_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  uint32_t bits = 0;
  VFI_CUDA(cudaMemcpy(&bits, idx->d_max_err, 4, cudaMemcpyDeviceToHost));
  float f;
  std::memcpy(&f, &bits, 4);
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
  std::lock_guard<std::mutex> lock(idx->mu);
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

int prep_queries(vfi_index* idx, const float* q_dev, int nq, cudaStream_t st) {
  VFI_TRY(idx->w_qcanon.ensure(static_cast<size_t>(nq) * idx->dp * 4));
  VFI_TRY(idx->w_qg.ensure(static_cast<size_t>(nq) * idx->kp * 2));
  VFI_TRY(idx->w_eps.ensure(static_cast<size_t>(nq) * 4));
  const int threads = 256;
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, threads));
  if (idx->store == VFI_STORE_F32)
    vfi::prep_queries_kernel<true><<<blocks, threads, 0, st>>>(q_dev, nq, idx->d, idx->dp, idx->w_qcanon.as<float>(),
                                                                idx->w_qg.as<uint16_t>(), idx->kp, idx->xnorm_bits,
                                                                idx->w_eps.as<float>());
  else
    vfi::prep_queries_kernel<false><<<blocks, threads, 0, st>>>(q_dev, nq, idx->d, idx->dp, idx->w_qcanon.as<float>(),
                                                                 idx->w_qg.as<uint16_t>(), idx->kp, idx->xnorm_bits,
                                                                 idx->w_eps.as<float>());
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return VFI_OK;
}

// exact canonical scores of ALL rows for the selected queries, then exact top-k (no certificate)
int exhaustive_pass(vfi_index* idx, const int* qsel_dev, int nsel, int k, float* out_scores, int64_t* out_ids,
                    cudaStream_t st) {
  const int64_t n = idx->n;
  const int64_t per_query = std::max<int64_t>(n, 1) * 8;
  const int chunk = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nsel, (static_cast<int64_t>(1) << 30) / per_query)));
  VFI_TRY(idx->w_keys2.ensure(static_cast<size_t>(chunk) * per_query));
  if (qsel_dev == nullptr) {  // all queries of the batch: identity list
    std::vector<int> iota(nsel);
    for (int i = 0; i < nsel; ++i) iota[i] = i;
    VFI_TRY(idx->w_sel.ensure(static_cast<size_t>(nsel) * 4));
    VFI_CUDA(cudaMemcpyAsync(idx->w_sel.p, iota.data(), static_cast<size_t>(nsel) * 4, cudaMemcpyHostToDevice, st));
    VFI_CUDA(cudaStreamSynchronize(st));
    qsel_dev = idx->w_sel.as<int>();
  }
  for (int s0 = 0; s0 < nsel; s0 += chunk) {
    const int ns = std::min(chunk, nsel - s0);
    const int* sel = qsel_dev + s0;
    if (n > 0) {
      dim3 grid(static_cast<unsigned>(ceil_div(n, 128)), static_cast<unsigned>(ns));
      if (idx->store == VFI_STORE_F32)
        vfi::canon_score_kernel<float><<<grid, 128, 0, st>>>(idx->master, idx->dp, idx->dp, idx->w_qcanon.as<float>(), sel,
                                                             nullptr, nullptr, 0, n, idx->w_keys2.as<uint64_t>(), n, nullptr);
      else
        vfi::canon_score_kernel<uint16_t><<<grid, 128, 0, st>>>(idx->g, idx->kp, idx->dp, idx->w_qcanon.as<float>(), sel,
                                                                nullptr, nullptr, 0, n, idx->w_keys2.as<uint64_t>(), n, nullptr);
      LAUNCHED();
      VFI_CUDA(cudaGetLastError());
    }
    vfi::finalize_kernel<<<ns, 256, sizeof(vfi::SelectSmem), st>>>(idx->w_keys2.as<uint64_t>(), n, n, sel, nullptr, k,
                                                                  idx->id_offset, nullptr, nullptr, 0, out_scores, out_ids,
                                                                  nullptr, nullptr);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
  }
  return VFI_OK;
}

// n_rows/row_stride: the corpus view scanned — (idx->n, 1) for the full shard, (R, s) for a strided row sample
int launch_fused(vfi_index* idx, int nq, int keep, int mode, float* scores_out, int64_t ld_scores, const float* tau,
                 int* o_groups, int* o_nq_pad, int* o_cap, cudaStream_t st, int64_t n_rows = -1, int64_t row_stride = 1,
                 bool profile = true) {
  if (n_rows < 0) n_rows = idx->n;
  const int n_mtiles = static_cast<int>(ceil_div(nq, vfi::kBM));
  // CTA pairs (tcgen05 cta_group::2) need an even number of query tiles; they are the default where possible
  // because they move a third less data per FLOP into the SMs (the single-CTA kernel is power-capped first).
  const bool pair = (n_mtiles % 2 == 0) && idx->opt_cta_pair != 1 && idx->opt_cluster <= 1;
  int n_ctas = idx->opt_num_ctas > 0 ? static_cast<int>(idx->opt_num_ctas) : idx->num_sms;
  n_ctas = std::min(std::max(n_ctas, n_mtiles), 512);   // 2*groups key buffers per query must stay <= 1024
  const int n_tiles = static_cast<int>(ceil_div(n_rows, vfi::kBN));
  int n_groups = std::max(1, n_ctas / n_mtiles);
  n_groups = std::min(n_groups, std::max(1, n_tiles));
  // pair kernel: every SM pair owns corpus tiles (n_groups = number of pairs) and walks all query tile pairs itself
  if (pair) {
    if (idx->max_pairs <= 0) {   // how many 2-CTA clusters of this kernel the device holds at once (74 on a full B200)
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
    n_groups = std::max(1, std::min(std::min(n_ctas / 2, idx->max_pairs), std::max(1, n_tiles)));
  }
  const int n_bufs = pair ? n_groups * vfi::pair_sets_per_query(n_mtiles) : 2 * n_groups;   // key buffers per query
  const int nq_pad = n_mtiles * vfi::kBM;
  const int cap = 2 * keep + 32;
  CUtensorMap tq, td;
  VFI_TRY(make_tmap(&tq, idx->w_qg.p, nq, idx->kp, idx->kp, vfi::kBM));
  // clusters of CTAs that work on the same corpus tile (consecutive query tiles of one group) fetch it once
  int cluster = idx->opt_cluster > 0 ? static_cast<int>(idx->opt_cluster) : 1;
  while (cluster > 1 && (n_mtiles % cluster) != 0) cluster >>= 1;
  if (pair) cluster = 2;
  VFI_TRY(make_tmap(&td, idx->g, n_rows, idx->kp, idx->kp * row_stride, vfi::kBN / cluster));
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
  p.cluster = cluster;
  if (mode == vfi::MODE_TOPK) {
    VFI_TRY(idx->w_cand.ensure(static_cast<size_t>(n_bufs) * nq_pad * cap * 8));
    VFI_TRY(idx->w_cand_count.ensure(static_cast<size_t>(n_bufs) * nq_pad * 4));
    p.cand = idx->w_cand.as<uint64_t>();
    p.cand_count = idx->w_cand_count.as<uint32_t>();
  }
  const bool prof = idx->opt_profile != 0 && profile;
  if (prof) cudaEventRecord(idx->slots[idx->cur_slot].pev[0], st);
  const int grid = pair ? 2 * n_groups : n_groups * n_mtiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(vfi::kDenseThreads);
  cfg.dynamicSmemBytes = pair ? vfi::kPairSmemBytes : vfi::kDenseSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
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
  if (prof) cudaEventRecord(idx->slots[idx->cur_slot].pev[1], st);
  VFI_CUDA(cudaGetLastError());
  if (profile) idx->stats.fused_launches++;
  if (o_groups) *o_groups = n_bufs;
  if (o_nq_pad) *o_nq_pad = nq_pad;
  if (o_cap) *o_cap = cap;
  return VFI_OK;
}

// Enqueue one batch of <= kMaxQueriesPerLaunch queries already on the device (results to device buffers) into `slot`
// without waiting: everything up to the copy of the certificate flag.  search_finish() waits and repairs.
int search_launch(vfi_index* idx, int slot_id, const float* q_dev, int nq, int k, float* out_scores, int64_t* out_ids,
                  cudaStream_t st, bool no_hint) {
  SearchSlot& sl = idx->slots[slot_id];
  sl.busy = true;
  sl.needs_check = false;
  sl.used_tau = false;
  sl.q = q_dev;
  sl.nq = nq;
  sl.k = k;
  sl.out_scores = out_scores;
  sl.out_ids = out_ids;
  sl.st = st;
  sl.seq = ++idx->launch_seq;
  idx->cur_slot = slot_id;
  VFI_TRY(prep_queries(idx, q_dev, nq, st));
  idx->prep_owner = sl.seq;
  const int64_t n = idx->n;
  int keep = idx->opt_overfetch > 0 ? static_cast<int>(idx->opt_overfetch)
                                    : static_cast<int>(round_up(k + std::max(16, k / 4), 32));
  keep = std::max(keep, static_cast<int>(round_up(k, 32)));
  int path = static_cast<int>(idx->opt_force_path);
  if (path == 0) {
    if (n <= kExhaustiveRows || keep >= n) path = 1;
    else if (nq <= vfi::kGemvMaxQ && keep <= kGemvMaxKeep && gemv_ok(idx)) path = 3;
    else if (keep <= kFusedMaxKeep) path = 2;
    else path = 1;
  }
  if (path == 2 && (keep > kFusedMaxKeep || n == 0)) path = 1;
  if (path == 3 && (nq > vfi::kGemvMaxQ || keep > kGemvMaxKeep || n == 0 || !gemv_ok(idx))) path = (keep <= kFusedMaxKeep && n > 0) ? 2 : 1;
  idx->stats.last_path = path;
  idx->stats.last_overfetch = keep;
  if (path == 1) {      // every row scored canonically: nothing to certify
    VFI_TRY(exhaustive_pass(idx, nullptr, nq, k, out_scores, out_ids, st));
    VFI_CUDA(cudaEventRecord(sl.done, st));
    return VFI_OK;
  }

  int n_groups = 0, nq_pad = 0, cap = 0;
  const float* tau = nullptr;
  if (path == 2 && idx->opt_tau_hint != 0 && !no_hint) {
    // admission hint: the m-th best score of a strided row sample estimates a threshold that about
    // 4k' rows of the shard exceed.  It only prunes work; exactness is re-established below.
    // Sample every s-th row with s = 2k'+1.  tau = the m-th best sampled score with m = 8: about
    // m*s = 16k' rows of the shard are expected above it, and fewer than k' with probability
    // P(Gamma(8) < 1/2) ~ 1e-7 per query (then the batch is simply redone without the hint).
    const int m = 8;
    const int64_t rs = 2 * static_cast<int64_t>(keep) + 1;
    const int64_t rr = n / rs;
    if (rr >= 4 * m || idx->opt_tau_hint == 2) {
      // The sample pass keeps one value per (query, 32 sampled rows) — the chunk's largest tensor-core score — instead of
      // every score: the m-th largest chunk maximum is at most the m-th largest sampled score (equal unless two of the
      // best m samples share a chunk), so it is a valid, marginally looser hint, and the threshold kernel reads 32 x less.
      // VFI_OPT_TAU_HINT = 3 keeps the full sample (every score stored) for comparison.
      const bool full = idx->opt_tau_hint == 3;
      const int64_t ld = full ? ceil_div(rr, vfi::kBN) * vfi::kBN : ceil_div(rr, vfi::kBN) * (vfi::kBN / 32);
      const int64_t n_vals = full ? rr : ceil_div(rr, 32);
      const int64_t nq_pad_s = ceil_div(nq, vfi::kBM) * vfi::kBM;
      VFI_TRY(idx->w_dbg.ensure(static_cast<size_t>(nq_pad_s) * ld * 4));
      VFI_TRY(idx->w_tau.ensure(static_cast<size_t>(nq) * 4));
      VFI_TRY(launch_fused(idx, nq, 32, full ? vfi::MODE_STORE : vfi::MODE_CHUNKMAX, idx->w_dbg.as<float>(), ld, nullptr, nullptr,
                           nullptr, nullptr, st, rr, rs, false));
      vfi::tau_from_scores_kernel<<<nq, 256, sizeof(vfi::SelectSmem), st>>>(idx->w_dbg.as<float>(), ld, static_cast<int>(n_vals), m,
                                                                           idx->w_tau.as<float>(), idx->opt_tau_hint == 2 ? 1 : 0);
      LAUNCHED();
      VFI_CUDA(cudaGetLastError());
      tau = idx->w_tau.as<float>();
    }
  }
  if (path == 2) {
    VFI_TRY(launch_fused(idx, nq, keep, vfi::MODE_TOPK, nullptr, 0, tau, &n_groups, &nq_pad, &cap, st));
  } else {
    // streaming scorer: per-CTA shared key buffers
    int cap_s = 1;
    while (cap_s < keep + vfi::kGemvRowsPerRound) cap_s <<= 1;
    const size_t smem = ((static_cast<size_t>(nq) * idx->dp * 4 + 15) & ~size_t(15)) + static_cast<size_t>(nq) * cap_s * 8 + 256;
    if (smem > 160 * 1024) return fail(VFI_ERR_UNSUPPORTED, "streaming scorer: query block does not fit shared memory");
    const int ctas_per_sm = (idx->store == VFI_STORE_F32) ? gemv_occupancy<float>(nq, smem) : gemv_occupancy<uint16_t>(nq, smem);
    n_groups = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(idx->num_sms) * ctas_per_sm,
                                                  ceil_div(n, vfi::kGemvRowsPerRound)));
    nq_pad = nq;
    cap = keep;
    VFI_TRY(idx->w_cand.ensure(static_cast<size_t>(n_groups) * nq_pad * cap * 8));
    VFI_TRY(idx->w_cand_count.ensure(static_cast<size_t>(n_groups) * nq_pad * 4));
    const bool prof = idx->opt_profile != 0;
    if (prof) cudaEventRecord(idx->slots[idx->cur_slot].pev[0], st);
    if (idx->store == VFI_STORE_F32)
      gemv_launch<float>(nq, n_groups, smem, st, idx->master, idx->dp, idx->dp, n, idx->w_qcanon.as<float>(), keep, cap_s,
                         idx->w_cand.as<uint64_t>(), idx->w_cand_count.as<uint32_t>(), nq_pad, cap);
    else
      gemv_launch<uint16_t>(nq, n_groups, smem, st, idx->g, idx->kp, idx->dp, n, idx->w_qcanon.as<float>(), keep, cap_s,
                            idx->w_cand.as<uint64_t>(), idx->w_cand_count.as<uint32_t>(), nq_pad, cap);
    LAUNCHED();
    if (prof) cudaEventRecord(idx->slots[idx->cur_slot].pev[1], st);
    VFI_CUDA(cudaGetLastError());
    idx->stats.fused_launches++;
  }
  VFI_TRY(idx->w_flag.ensure(static_cast<size_t>(kSearchSlots) * (kMaxQueriesPerLaunch + 1) * 4));
  int* d_flag = idx->w_flag.as<int>() + static_cast<size_t>(slot_id) * (kMaxQueriesPerLaunch + 1);   // [0] count, [1..] flagged queries
  VFI_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
  const bool tail_prof = idx->opt_profile != 0;
  if (tail_prof) cudaEventRecord(idx->slots[idx->cur_slot].pev[2], st);
  if (keep <= 256 && idx->dp <= vfi::kRfMaxDp && idx->opt_tail != 1) {
    // K1c: per-query union of the group buffers -> k' best by tensor-core score; then K2: one thread per candidate
    // rescoring + final order + certificate
    VFI_TRY(idx->w_keys.ensure(static_cast<size_t>(nq) * keep * 8));
    VFI_TRY(idx->w_keys_n.ensure(static_cast<size_t>(nq) * 4));
    VFI_TRY(idx->w_bound.ensure(static_cast<size_t>(nq) * 4));
    // about 16 k' keys reach this kernel per query: up to k' = 128 they fit the light 1024-key selection buffer's fast
    // paths (seven CTAs per SM); above that the 4096-key buffer avoids the radix walk over global memory
    if (keep <= 128)
      vfi::cand_reduce_kernel<vfi::CandSmem><<<nq, 256, sizeof(vfi::CandSmem), st>>>(
          idx->w_cand.as<uint64_t>(), idx->w_cand_count.as<uint32_t>(), n_groups, nq_pad, cap, keep, tau, idx->w_keys.as<uint64_t>(),
          idx->w_keys_n.as<uint32_t>(), idx->w_bound.as<float>());
    else
      vfi::cand_reduce_kernel<vfi::SelectSmem><<<nq, 256, sizeof(vfi::SelectSmem), st>>>(
          idx->w_cand.as<uint64_t>(), idx->w_cand_count.as<uint32_t>(), n_groups, nq_pad, cap, keep, tau, idx->w_keys.as<uint64_t>(),
          idx->w_keys_n.as<uint32_t>(), idx->w_bound.as<float>());
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    const int threads = static_cast<int>(round_up(keep, 32));
#define VFI_RF_LAUNCH(T, P, S, ROWS, PITCH)                                                                              \
  vfi::rescore_finalize_kernel<T, P, S><<<nq, threads, vfi::rescore_bulk_smem<P>(static_cast<int>(idx->dp), threads, S), st>>>( \
      idx->w_keys.as<uint64_t>(), idx->w_keys_n.as<uint32_t>(), idx->w_bound.as<float>(), keep, ROWS, PITCH,             \
      static_cast<int>(idx->dp), idx->w_qcanon.as<float>(), k, idx->id_offset, idx->w_eps.as<float>(), out_scores, out_ids, \
      d_flag + 1, d_flag, idx->d_max_err)
    const bool f32 = idx->store == VFI_STORE_F32;
    if (idx->opt_tail == 2) { if (f32) VFI_RF_LAUNCH(float, 128, 1, idx->master, idx->dp); else VFI_RF_LAUNCH(uint16_t, 128, 1, idx->g, idx->kp); }
    else if (idx->opt_tail == 3) { if (f32) VFI_RF_LAUNCH(float, 256, 2, idx->master, idx->dp); else VFI_RF_LAUNCH(uint16_t, 256, 2, idx->g, idx->kp); }
    else { if (f32) VFI_RF_LAUNCH(float, 256, 1, idx->master, idx->dp); else VFI_RF_LAUNCH(uint16_t, 256, 1, idx->g, idx->kp); }
#undef VFI_RF_LAUNCH
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
  } else if (keep <= 256) {
    // single-launch tail: union of the group buffers + k' selection + canonical rescoring + certificate
    if (idx->store == VFI_STORE_F32)
      vfi::select_rescore_kernel<float><<<nq, 256, sizeof(vfi::TailSmem), st>>>(
          idx->w_cand.as<uint64_t>(), idx->w_cand_count.as<uint32_t>(), n_groups, nq_pad, cap, keep, tau, idx->master, idx->dp,
          idx->dp, idx->w_qcanon.as<float>(), k, idx->id_offset, idx->w_eps.as<float>(), out_scores, out_ids, d_flag + 1, d_flag,
          idx->d_max_err);
    else
      vfi::select_rescore_kernel<uint16_t><<<nq, 256, sizeof(vfi::TailSmem), st>>>(
          idx->w_cand.as<uint64_t>(), idx->w_cand_count.as<uint32_t>(), n_groups, nq_pad, cap, keep, tau, idx->g, idx->kp,
          idx->dp, idx->w_qcanon.as<float>(), k, idx->id_offset, idx->w_eps.as<float>(), out_scores, out_ids, d_flag + 1, d_flag,
          idx->d_max_err);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
  } else {
    // K1c: per-query union of the group buffers
    VFI_TRY(idx->w_keys.ensure(static_cast<size_t>(nq) * keep * 8));
    VFI_TRY(idx->w_keys_n.ensure(static_cast<size_t>(nq) * 4));
    VFI_TRY(idx->w_bound.ensure(static_cast<size_t>(nq) * 4));
    vfi::cand_reduce_kernel<vfi::SelectSmem><<<nq, 256, sizeof(vfi::SelectSmem), st>>>(idx->w_cand.as<uint64_t>(), idx->w_cand_count.as<uint32_t>(),
                                                                     n_groups, nq_pad, cap, keep, tau, idx->w_keys.as<uint64_t>(),
                                                                     idx->w_keys_n.as<uint32_t>(), idx->w_bound.as<float>());
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    // K2a: canonical rescoring of the candidates
    VFI_TRY(idx->w_keys2.ensure(static_cast<size_t>(nq) * keep * 8));
    dim3 grid(static_cast<unsigned>(ceil_div(keep, 128)), static_cast<unsigned>(nq));
    if (idx->store == VFI_STORE_F32)
      vfi::canon_score_kernel<float><<<grid, 128, 0, st>>>(idx->master, idx->dp, idx->dp, idx->w_qcanon.as<float>(), nullptr,
                                                           idx->w_keys.as<uint64_t>(), idx->w_keys_n.as<uint32_t>(), keep, 0,
                                                           idx->w_keys2.as<uint64_t>(), keep, idx->d_max_err);
    else
      vfi::canon_score_kernel<uint16_t><<<grid, 128, 0, st>>>(idx->g, idx->kp, idx->dp, idx->w_qcanon.as<float>(), nullptr,
                                                              idx->w_keys.as<uint64_t>(), idx->w_keys_n.as<uint32_t>(), keep, 0,
                                                              idx->w_keys2.as<uint64_t>(), keep, idx->d_max_err);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
    // K2b: final order + certificate
    vfi::finalize_kernel<<<nq, 256, sizeof(vfi::SelectSmem), st>>>(idx->w_keys2.as<uint64_t>(), keep, keep, nullptr,
                                                                  idx->w_keys_n.as<uint32_t>(), k, idx->id_offset,
                                                                  idx->w_bound.as<float>(), idx->w_eps.as<float>(), 1, out_scores,
                                                                  out_ids, d_flag + 1, d_flag);
    LAUNCHED();
    VFI_CUDA(cudaGetLastError());
  }
  if (tail_prof) cudaEventRecord(idx->slots[idx->cur_slot].pev[3], st);
  VFI_CUDA(cudaMemcpyAsync(idx->h_flag + slot_id, d_flag, 4, cudaMemcpyDeviceToHost, st));
  VFI_CUDA(cudaEventRecord(sl.done, st));
  sl.needs_check = true;
  sl.used_tau = tau != nullptr;
  return VFI_OK;
}

// Wait for the batch in `slot`, read its certificate flag and repair what failed: a batch pruned too hard by the admission
// hint is redone without it, queries whose candidates tie across the cut are re-run by the exhaustive canonical pass.
int search_finish(vfi_index* idx, int slot_id) {
  SearchSlot& sl = idx->slots[slot_id];
  if (!sl.busy) return fail(VFI_ERR_INVALID, "no batch in flight for this ticket");
  sl.busy = false;
  VFI_CUDA(cudaEventSynchronize(sl.done));
  idx->cur_slot = slot_id;
  if (idx->opt_profile && sl.needs_check) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, idx->slots[idx->cur_slot].pev[2], idx->slots[idx->cur_slot].pev[3]) == cudaSuccess) {
      idx->stats.tail_ms_total += ms;
      idx->stats.tail_ms_samples++;
    } else {
      cudaGetLastError();
    }
    if (cudaEventElapsedTime(&ms, idx->slots[idx->cur_slot].pev[0], idx->slots[idx->cur_slot].pev[1]) == cudaSuccess) {
      idx->stats.fused_ms_total += ms;
      idx->stats.fused_ms_samples++;
    } else {
      cudaGetLastError();
    }
  }
  if (!sl.needs_check) return VFI_OK;
  const int n_flagged = idx->h_flag[slot_id];
  if (n_flagged <= 0) return VFI_OK;
  if (sl.used_tau && (idx->opt_tau_hint == 1 || idx->opt_tau_hint == 3)) {
    // the hint pruned too much for some query: redo the batch without it (same kernels, no pruning)
    idx->stats.hint_retries++;
    VFI_TRY(search_launch(idx, slot_id, sl.q, sl.nq, sl.k, sl.out_scores, sl.out_ids, sl.st, true));
    return search_finish(idx, slot_id);
  }
  idx->stats.retried_queries += n_flagged;
  if (idx->prep_owner != sl.seq) {   // a later batch (or the repair of another one) has replaced the prepared queries
    VFI_TRY(prep_queries(idx, sl.q, sl.nq, sl.st));
    idx->prep_owner = sl.seq;
  }
  const int* flagged = idx->w_flag.as<int>() + static_cast<size_t>(slot_id) * (kMaxQueriesPerLaunch + 1) + 1;
  VFI_TRY(exhaustive_pass(idx, flagged, n_flagged, sl.k, sl.out_scores, sl.out_ids, sl.st));
  VFI_CUDA(cudaStreamSynchronize(sl.st));
  return VFI_OK;
}

int free_slot(const vfi_index* idx) {
  for (int i = 0; i < kSearchSlots; ++i)
    if (!idx->slots[i].busy) return i;
  return -1;
}

// the synchronous form: launch + finish
int search_batch(vfi_index* idx, const float* q_dev, int nq, int k, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  const int slot = free_slot(idx);
  if (slot < 0) return fail(VFI_ERR_UNSUPPORTED, "too many batches in flight: finish a ticket of vfi_index_search_begin first");
  const int rc = search_launch(idx, slot, q_dev, nq, k, out_scores, out_ids, st, false);
  if (rc != VFI_OK) {
    idx->slots[slot].busy = false;
    return rc;
  }
  return search_finish(idx, slot);
}

}  // namespace

extern "C" {

int vfi_index_search(vfi_index_t* idx, const float* q, int64_t nq, int k, float* out_scores, int64_t* out_ids, int mem,
                     void* stream) {
  if (!idx || (nq > 0 && (!q || !out_scores || !out_ids)) || nq < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_search");
  if (k <= 0) return fail(VFI_ERR_INVALID, "k must be positive");
  if (k > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k exceeds VFI_MAX_K (2048)");
  if (nq == 0) return VFI_OK;
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  idx->stats.searches++;
  idx->stats.queries += nq;
  for (int64_t q0 = 0; q0 < nq; q0 += kMaxQueriesPerLaunch) {
    const int nb = static_cast<int>(std::min<int64_t>(kMaxQueriesPerLaunch, nq - q0));
    const float* qd = q + q0 * idx->d;
    float* os = out_scores + q0 * k;
    int64_t* oi = out_ids + q0 * k;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(idx->w_qin.ensure(static_cast<size_t>(nb) * idx->d * 4));
      VFI_TRY(idx->w_out_scores.ensure(static_cast<size_t>(nb) * k * 4));
      VFI_TRY(idx->w_out_ids.ensure(static_cast<size_t>(nb) * k * 8));
      VFI_CUDA(cudaMemcpyAsync(idx->w_qin.p, qd, static_cast<size_t>(nb) * idx->d * 4, cudaMemcpyHostToDevice, st));
      qd = idx->w_qin.as<float>();
      os = idx->w_out_scores.as<float>();
      oi = idx->w_out_ids.as<int64_t>();
    }
    VFI_TRY(search_batch(idx, qd, nb, k, os, oi, st));
    if (mem == VFI_MEM_HOST) {
      VFI_CUDA(cudaMemcpyAsync(out_scores + q0 * k, os, static_cast<size_t>(nb) * k * 4, cudaMemcpyDeviceToHost, st));
      VFI_CUDA(cudaMemcpyAsync(out_ids + q0 * k, oi, static_cast<size_t>(nb) * k * 8, cudaMemcpyDeviceToHost, st));
      VFI_CUDA(cudaStreamSynchronize(st));
    }
  }
  return VFI_OK;
}

int vfi_index_search_begin(vfi_index_t* idx, const float* q, int64_t nq, int k, float* out_scores, int64_t* out_ids, void* stream,
                           int* ticket) {
  if (!idx || !ticket || !q || !out_scores || !out_ids || nq <= 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_search_begin");
  if (k <= 0) return fail(VFI_ERR_INVALID, "k must be positive");
  if (k > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k exceeds VFI_MAX_K (2048)");
  if (nq > kMaxQueriesPerLaunch) return fail(VFI_ERR_UNSUPPORTED, "vfi_index_search_begin takes at most 1024 queries per batch");
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  const int slot = free_slot(idx);
  if (slot < 0) return fail(VFI_ERR_UNSUPPORTED, "too many batches in flight (4): finish a ticket first");
  idx->stats.searches++;
  idx->stats.queries += nq;
  const int rc = search_launch(idx, slot, q, static_cast<int>(nq), k, out_scores, out_ids, static_cast<cudaStream_t>(stream), false);
  if (rc != VFI_OK) {
    idx->slots[slot].busy = false;
    return rc;
  }
  *ticket = slot;
  return VFI_OK;
}

int vfi_index_search_finish(vfi_index_t* idx, int ticket) {
  if (!idx || ticket < 0 || ticket >= kSearchSlots) return fail(VFI_ERR_INVALID, "bad ticket");
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  return search_finish(idx, ticket);
}

int vfi_index_read_rows(vfi_index_t* idx, int64_t first, int64_t n, float* out, int mem, void* stream) {
  if (!idx || (n > 0 && !out) || first < 0 || n < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_read_rows");
  if (first + n > idx->n) return fail(VFI_ERR_INVALID, "rows out of range");
  if (n == 0) return VFI_OK;
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunk = std::max<int64_t>(1, (static_cast<int64_t>(256) << 20) / (static_cast<int64_t>(idx->d) * 4));
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t rows = std::min(chunk, n - r0);
    float* dst = out + r0 * idx->d;
    if (mem == VFI_MEM_HOST) {
      VFI_TRY(idx->w_stage.ensure(static_cast<size_t>(rows) * idx->d * 4));
      dst = idx->w_stage.as<float>();
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
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<int64_t> hid(n);
  const int64_t* dids = ids;
  if (mem == VFI_MEM_DEVICE) VFI_CUDA(cudaMemcpyAsync(hid.data(), ids, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
  else std::memcpy(hid.data(), ids, sizeof(int64_t) * n);
  VFI_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i)
    if (hid[i] < 0 || hid[i] >= idx->n) return fail(VFI_ERR_INVALID, "pairwise: row id out of range");
  VFI_TRY(idx->w_sel.ensure(static_cast<size_t>(n) * 8));
  if (mem == VFI_MEM_HOST) {
    VFI_CUDA(cudaMemcpyAsync(idx->w_sel.p, hid.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    dids = idx->w_sel.as<int64_t>();
  }
  VFI_TRY(idx->w_qcanon.ensure(static_cast<size_t>(n) * idx->dp * 4));
  VFI_TRY(idx->w_keys.ensure(static_cast<size_t>(n) * n * 8));
  VFI_TRY(idx->w_keys2.ensure(static_cast<size_t>(n) * n * 8));
  VFI_TRY(idx->w_keys_n.ensure(static_cast<size_t>(n) * 4));
  const int64_t work = std::max<int64_t>(static_cast<int64_t>(n) * idx->dp, static_cast<int64_t>(n) * n);
  const unsigned blocks = static_cast<unsigned>(ceil_div(work, 256));
  dim3 grid(static_cast<unsigned>(ceil_div(n, 128)), static_cast<unsigned>(n));
  if (idx->store == VFI_STORE_F32) {
    vfi::gather_rows_kernel<float><<<blocks, 256, 0, st>>>(idx->master, idx->dp, dids, n, idx->dp, idx->w_qcanon.as<float>(),
                                                           idx->w_keys.as<uint64_t>(), idx->w_keys_n.as<uint32_t>());
    vfi::canon_score_kernel<float><<<grid, 128, 0, st>>>(idx->master, idx->dp, idx->dp, idx->w_qcanon.as<float>(), nullptr,
                                                         idx->w_keys.as<uint64_t>(), idx->w_keys_n.as<uint32_t>(), n, 0,
                                                         idx->w_keys2.as<uint64_t>(), n, nullptr);
  } else {
    vfi::gather_rows_kernel<uint16_t><<<blocks, 256, 0, st>>>(idx->g, idx->kp, dids, n, idx->dp, idx->w_qcanon.as<float>(),
                                                              idx->w_keys.as<uint64_t>(), idx->w_keys_n.as<uint32_t>());
    vfi::canon_score_kernel<uint16_t><<<grid, 128, 0, st>>>(idx->g, idx->kp, idx->dp, idx->w_qcanon.as<float>(), nullptr,
                                                            idx->w_keys.as<uint64_t>(), idx->w_keys_n.as<uint32_t>(), n, 0,
                                                            idx->w_keys2.as<uint64_t>(), n, nullptr);
  }
  LAUNCHED();
  LAUNCHED();
  float* dout = out;
  if (mem == VFI_MEM_HOST) {
    VFI_TRY(idx->w_dbg.ensure(static_cast<size_t>(n) * n * 4));
    dout = idx->w_dbg.as<float>();
  }
  vfi::keys_to_scores_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(n) * n, 256)), 256, 0, st>>>(
      idx->w_keys2.as<uint64_t>(), static_cast<int64_t>(n) * n, dout);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  if (mem == VFI_MEM_HOST) VFI_CUDA(cudaMemcpyAsync(out, dout, static_cast<size_t>(n) * n * 4, cudaMemcpyDeviceToHost, st));
  VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
}

int vfi_index_debug_scores(vfi_index_t* idx, const float* q, int64_t nq, float* out, int mem, void* stream) {
  if (!idx || !q || !out || nq <= 0 || nq > kMaxQueriesPerLaunch) return fail(VFI_ERR_INVALID, "bad argument to vfi_index_debug_scores");
  if (idx->n == 0) return VFI_OK;
  std::lock_guard<std::mutex> lock(idx->mu);
  DeviceGuard guard(idx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* qd = q;
  if (mem == VFI_MEM_HOST) {
    VFI_TRY(idx->w_qin.ensure(static_cast<size_t>(nq) * idx->d * 4));
    VFI_CUDA(cudaMemcpyAsync(idx->w_qin.p, q, static_cast<size_t>(nq) * idx->d * 4, cudaMemcpyHostToDevice, st));
    qd = idx->w_qin.as<float>();
  }
  VFI_TRY(prep_queries(idx, qd, static_cast<int>(nq), st));
  const int64_t ld = ceil_div(idx->n, vfi::kBN) * vfi::kBN;
  const int64_t nq_pad = ceil_div(nq, vfi::kBM) * vfi::kBM;
  VFI_TRY(idx->w_dbg.ensure(static_cast<size_t>(nq_pad) * ld * 4));
  VFI_TRY(launch_fused(idx, static_cast<int>(nq), 32, vfi::MODE_STORE, idx->w_dbg.as<float>(), ld, nullptr, nullptr, nullptr, nullptr, st));
  VFI_CUDA(cudaMemcpy2DAsync(out, static_cast<size_t>(idx->n) * 4, idx->w_dbg.p, static_cast<size_t>(ld) * 4,
                             static_cast<size_t>(idx->n) * 4, static_cast<size_t>(nq),
                             mem == VFI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
}

int vfi_normalize_l2(float* x, int64_t n, int d, int mem, int device, void* stream) {
  if ((!x && n > 0) || n < 0 || d <= 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_normalize_l2");
  if (n == 0) return VFI_OK;
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunk = std::max<int64_t>(1, (static_cast<int64_t>(256) << 20) / (static_cast<int64_t>(d) * 4));
  float* stage = nullptr;
  if (mem == VFI_MEM_HOST) VFI_CUDA(cudaMalloc(&stage, static_cast<size_t>(std::min(chunk, n)) * d * 4));
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
  if (stage) cudaFree(stage);
  return rc;
}

}  // extern "C"

// =============================================================================================
// small helpers shared by merge / fusion / cosine entry points: stage host buffers on the device
// =============================================================================================
namespace {

struct Staged {
  std::vector<void*> to_free;
  ~Staged() {
    for (void* p : to_free) cudaFree(p);
  }
  template <class T>
  int in(const T* src, size_t count, int mem, cudaStream_t st, const T** out) {
    if (mem == VFI_MEM_DEVICE) { *out = src; return VFI_OK; }
    void* p = nullptr;
    VFI_CUDA(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)));
    to_free.push_back(p);
    VFI_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *out = reinterpret_cast<const T*>(p);
    return VFI_OK;
  }
  template <class T>
  int out(T* dst, size_t count, int mem, T** dev) {
    if (mem == VFI_MEM_DEVICE) { *dev = dst; return VFI_OK; }
    void* p = nullptr;
    VFI_CUDA(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)));
    to_free.push_back(p);
    *dev = reinterpret_cast<T*>(p);
    return VFI_OK;
  }
  template <class T>
  int back(T* dst, const T* dev, size_t count, int mem, cudaStream_t st) {
    if (mem == VFI_MEM_DEVICE) return VFI_OK;
    VFI_CUDA(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    return VFI_OK;
  }
};

}  // namespace

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
  Staged sg;
  const size_t n_in = static_cast<size_t>(g) * nq * k_in, n_out = static_cast<size_t>(nq) * k_out;
  const float* ds;
  const int64_t* di;
  float* os;
  int64_t* oi;
  VFI_TRY(sg.in(scores, n_in, mem, st, &ds));
  VFI_TRY(sg.in(ids, n_in, mem, st, &di));
  VFI_TRY(sg.out(out_scores, n_out, mem, &os));
  VFI_TRY(sg.out(out_ids, n_out, mem, &oi));
  vfi::merge_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(ds, di, g, nq, k_in, k_out, os, oi);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  VFI_TRY(sg.back(out_scores, os, n_out, mem, st));
  VFI_TRY(sg.back(out_ids, oi, n_out, mem, st));
  if (mem == VFI_MEM_HOST) VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
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
  Staged sg;
  const size_t n_in = static_cast<size_t>(nq) * n_paths * depth, n_out = static_cast<size_t>(nq) * k;
  const int64_t* di;
  float* os;
  int64_t* oi;
  VFI_TRY(sg.in(ids, n_in, mem, st, &di));
  VFI_TRY(sg.out(out_scores, n_out, mem, &os));
  VFI_TRY(sg.out(out_ids, n_out, mem, &oi));
  vfi::rrf_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(di, n_paths, depth, k_rrf, k, os, oi);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  VFI_TRY(sg.back(out_scores, os, n_out, mem, st));
  VFI_TRY(sg.back(out_ids, oi, n_out, mem, st));
  if (mem == VFI_MEM_HOST) VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
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
  Staged sg;
  const size_t n_in = static_cast<size_t>(nq) * n_paths * depth;
  const int64_t* di;
  const float* ds;
  int64_t* oi;
  float* os;
  int32_t *op, *oc;
  VFI_TRY(sg.in(ids, n_in, mem, st, &di));
  VFI_TRY(sg.in(scores, n_in, mem, st, &ds));
  VFI_TRY(sg.out(out_ids, n_in, mem, &oi));
  VFI_TRY(sg.out(out_scores, n_in, mem, &os));
  VFI_TRY(sg.out(out_path, n_in, mem, &op));
  VFI_TRY(sg.out(out_count, static_cast<size_t>(nq), mem, &oc));
  vfi::union_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(di, ds, n_paths, depth, oi, os, op, oc);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  VFI_TRY(sg.back(out_ids, oi, n_in, mem, st));
  VFI_TRY(sg.back(out_scores, os, n_in, mem, st));
  VFI_TRY(sg.back(out_path, op, n_in, mem, st));
  VFI_TRY(sg.back(out_count, oc, static_cast<size_t>(nq), mem, st));
  if (mem == VFI_MEM_HOST) VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
}

// cosine top-k of the experiment scripts: normalise both sides, exhaustive exact scores, then the
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
  vfi_index_t* idx = nullptr;
  VFI_TRY(vfi_index_create(d, VFI_STORE_F32, device, &idx));
  int rc = vfi_index_add(idx, cn.data(), n_c, VFI_MEM_HOST, stream);
  if (rc == VFI_OK) rc = vfi_index_search(idx, en.data(), n_e, k, out_scores, out_ids, VFI_MEM_HOST, stream);
  vfi_index_destroy(idx);
  if (rc != VFI_OK) return rc;
  for (int64_t i = 0; i < n_e * k; ++i)
    if (out_ids[i] >= 0) out_ids[i] = n_c - 1 - out_ids[i];
  return VFI_OK;
}

}  // extern "C"

// =============================================================================================
// K3p: peer-memory exchange + merge (multi-GPU, one process per GPU)
// =============================================================================================
struct vfi_exchange {
  int device = 0, rank = 0, world = 1, max_k = 0, num_sms = 148, max_resident = 148;
  int64_t max_nq = 0;
  size_t win_bytes = 0, total_bytes = 0;
  uint8_t* local = nullptr;                 // own window + flags (cudaMalloc, exported by CUDA IPC)
  uint8_t* peer[vfi::kMaxPeers] = {};       // every rank's window as mapped here (peer[rank] == local)
  bool connected = false;
  uint32_t epoch = 0;
  uint64_t timeout_ms = 30000;
  std::mutex mu;
};

extern "C" {

int vfi_exchange_create(int device, int rank, int world, int64_t max_nq, int max_k, vfi_exchange_t** out) {
  if (!out || world < 1 || world > vfi::kMaxPeers || rank < 0 || rank >= world || max_nq <= 0 || max_k <= 0 || max_k > VFI_MAX_K)
    return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_create (world <= 16, 0 < max_k <= VFI_MAX_K)");
  *out = nullptr;
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  auto* ex = new vfi_exchange();
  ex->device = device;
  ex->rank = rank;
  ex->world = world;
  ex->max_nq = max_nq;
  ex->max_k = max_k;
  ex->num_sms = prop.multiProcessorCount;
  ex->win_bytes = static_cast<size_t>(round_up(2ll * world * max_nq * max_k * 8, 256));
  ex->total_bytes = ex->win_bytes + static_cast<size_t>(round_up(2ll * world * max_nq * 4, 256));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, ex->total_bytes);
  if (e != cudaSuccess) {
    delete ex;
    return fail(VFI_ERR_NOMEM, std::string("cudaMalloc(exchange window): ") + cudaGetErrorString(e));
  }
  ex->local = static_cast<uint8_t*>(p);
  ex->peer[rank] = ex->local;
  e = cudaMemset(p, 0, ex->total_bytes);          // flags start at epoch 0; synchronous w.r.t. the host
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  int occ = 1;
  if (e == cudaSuccess)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vfi::exchange_merge_kernel, 256, sizeof(vfi::SelectSmem));
  if (e != cudaSuccess) {
    cudaFree(p);
    delete ex;
    return fail(VFI_ERR_CUDA, std::string("exchange window setup: ") + cudaGetErrorString(e));
  }
  ex->max_resident = std::max(1, occ) * ex->num_sms;   // the grid never exceeds what is resident at once
  ex->connected = (world == 1);
  *out = ex;
  return VFI_OK;
}

int vfi_exchange_handle(vfi_exchange_t* ex, void* out_handle) {
  if (!ex || !out_handle) return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == VFI_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  DeviceGuard guard(ex->device);
  cudaIpcMemHandle_t h;
  VFI_CUDA(cudaIpcGetMemHandle(&h, ex->local));
  std::memcpy(out_handle, &h, sizeof(h));
  return VFI_OK;
}

int vfi_exchange_connect(vfi_exchange_t* ex, const void* handles) {
  if (!ex || !handles) return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_connect");
  std::lock_guard<std::mutex> lock(ex->mu);
  if (ex->connected) return VFI_OK;
  DeviceGuard guard(ex->device);
  const uint8_t* hb = static_cast<const uint8_t*>(handles);
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hb + static_cast<size_t>(r) * VFI_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int j = 0; j < r; ++j)
        if (j != ex->rank && ex->peer[j]) { cudaIpcCloseMemHandle(ex->peer[j]); ex->peer[j] = nullptr; }
      return fail(VFI_ERR_CUDA, "cudaIpcOpenMemHandle(rank " + std::to_string(r) + "): " + cudaGetErrorString(e) +
                                    " (peer windows need one process per GPU on one NVLink/PCIe-P2P node)");
    }
    ex->peer[r] = static_cast<uint8_t*>(p);
  }
  ex->connected = true;
  return VFI_OK;
}

int vfi_exchange_set_timeout_ms(vfi_exchange_t* ex, int64_t ms) {
  if (!ex || ms <= 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_set_timeout_ms");
  ex->timeout_ms = static_cast<uint64_t>(ms);
  return VFI_OK;
}

int vfi_exchange_merge(vfi_exchange_t* ex, const float* scores, const int64_t* ids, int64_t nq, int k, int k_out,
                       float* out_scores, int64_t* out_ids, void* stream) {
  if (!ex || nq < 0 || k <= 0 || k_out <= 0 || (nq > 0 && (!scores || !ids || !out_scores || !out_ids)))
    return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_merge");
  if (!ex->connected) return fail(VFI_ERR_INVALID, "vfi_exchange_merge before vfi_exchange_connect");
  if (nq > ex->max_nq || k > ex->max_k) return fail(VFI_ERR_INVALID, "nq or k exceeds the window geometry given at create");
  if (k_out > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k_out exceeds VFI_MAX_K");
  std::lock_guard<std::mutex> lock(ex->mu);
  DeviceGuard guard(ex->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  ex->epoch++;                       // every rank must make the same sequence of calls (a collective)
  if (nq == 0) return VFI_OK;
  vfi::ExchangeParams p{};
  for (int r = 0; r < ex->world; ++r) {
    p.win[r] = reinterpret_cast<uint64_t*>(ex->peer[r]);
    p.flags[r] = reinterpret_cast<uint32_t*>(ex->peer[r] + ex->win_bytes);
  }
  p.rank = ex->rank;
  p.world = ex->world;
  p.nq = static_cast<int>(nq);
  p.k = k;
  p.k_out = k_out;
  p.max_nq = ex->max_nq;
  p.max_k = ex->max_k;
  p.epoch = ex->epoch;
  p.scores = scores;
  p.ids = ids;
  p.out_scores = out_scores;
  p.out_ids = out_ids;
  p.timeout_ns = ex->timeout_ms * 1000000ull;
  const int grid = static_cast<int>(std::min<int64_t>(nq, ex->max_resident));
  vfi::exchange_merge_kernel<<<grid, 256, sizeof(vfi::SelectSmem), static_cast<cudaStream_t>(stream)>>>(p);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return VFI_OK;
}

int vfi_exchange_destroy(vfi_exchange_t* ex) {
  if (!ex) return VFI_OK;
  DeviceGuard guard(ex->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < ex->world; ++r)
    if (r != ex->rank && ex->peer[r]) cudaIpcCloseMemHandle(ex->peer[r]);
  if (ex->local) cudaFree(ex->local);
  cudaGetLastError();
  delete ex;
  return VFI_OK;
}

}  // extern "C"

// =============================================================================================
// BM25
// =============================================================================================
struct vfi_bm25 {
  int device = 0, num_sms = 148;
  int64_t n_vocab = 0, n_docs = 0, nnz = 0, id_offset = 0;
  int all_positive = 1;
  int64_t* indptr = nullptr;
  int32_t* indices = nullptr;
  float* data = nullptr;
  std::vector<int64_t> h_indptr;  // host copy: df lookups for stats and validation
  DevBuf w_tok, w_qptr, w_cand, w_cand_count, w_keys, w_keys_n, w_bound, w_out_scores, w_out_ids, w_ctr, w_dump, w_sort;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int profile = 0;
  vfi_bm25_stats stats{};
  std::mutex mu;
};

extern "C" {

int vfi_bm25_create(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_vocab, int64_t n_docs,
                    int64_t id_offset, int device, vfi_bm25_t** out) {
  if (!out) return fail(VFI_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!indptr || n_vocab < 0 || n_docs < 0 || id_offset < 0) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_create");
  const int64_t nnz = indptr[n_vocab];
  if (nnz < 0 || (nnz > 0 && (!indices || !data))) return fail(VFI_ERR_INVALID, "bad posting arrays");
  if (n_docs >= 0x7FFFFFF0ll) return fail(VFI_ERR_UNSUPPORTED, "a BM25 shard holds at most 2^31-16 docs");
  for (int64_t t = 0; t < n_vocab; ++t) {
    if (indptr[t] > indptr[t + 1]) return fail(VFI_ERR_INVALID, "indptr must be non-decreasing");
    // the kernel indexes its shared accumulator with these ids: every posting list must hold in-range,
    // strictly ascending doc ids (the bm25s layout)
    for (int64_t i = indptr[t]; i < indptr[t + 1]; ++i) {
      if (indices[i] < 0 || indices[i] >= n_docs) return fail(VFI_ERR_INVALID, "posting doc id out of range");
      if (i > indptr[t] && indices[i] <= indices[i - 1]) return fail(VFI_ERR_INVALID, "postings of a token must have strictly ascending doc ids");
    }
  }
  cudaDeviceProp prop;
  VFI_TRY(device_props(device, &prop));
  DeviceGuard guard(device);
  vfi_bm25* b = new vfi_bm25();
  b->device = device;
  b->num_sms = prop.multiProcessorCount;
  b->n_vocab = n_vocab;
  b->n_docs = n_docs;
  b->nnz = nnz;
  b->id_offset = id_offset;
  b->h_indptr.assign(indptr, indptr + n_vocab + 1);
  float mn = 1.f;
  for (int64_t i = 0; i < nnz; ++i) mn = std::min(mn, data[i]);
  b->all_positive = (mn > 0.f) ? 1 : 0;
  auto bail = [&](int code) {
    vfi_bm25_destroy(b);
    return code;
  };
  if (cudaMalloc(&b->indptr, sizeof(int64_t) * (n_vocab + 1)) != cudaSuccess ||
      cudaMalloc(&b->indices, std::max<size_t>(16, sizeof(int32_t) * nnz)) != cudaSuccess ||
      cudaMalloc(&b->data, std::max<size_t>(16, sizeof(float) * nnz)) != cudaSuccess)
    return bail(fail(VFI_ERR_NOMEM, "cudaMalloc postings failed"));
  cudaMemcpy(b->indptr, indptr, sizeof(int64_t) * (n_vocab + 1), cudaMemcpyHostToDevice);
  if (nnz > 0) {
    cudaMemcpy(b->indices, indices, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice);
    cudaMemcpy(b->data, data, sizeof(float) * nnz, cudaMemcpyHostToDevice);
  }
  cudaEventCreate(&b->ev0);
  cudaEventCreate(&b->ev1);
  cudaFuncSetAttribute(vfi::bm25_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bail(fail(VFI_ERR_CUDA, std::string("bm25 setup: ") + cudaGetErrorString(e)));
  *out = b;
  return VFI_OK;
}

int vfi_bm25_destroy(vfi_bm25_t* b) {
  if (!b) return VFI_OK;
  DeviceGuard guard(b->device);
  cudaDeviceSynchronize();
  if (b->indptr) cudaFree(b->indptr);
  if (b->indices) cudaFree(b->indices);
  if (b->data) cudaFree(b->data);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  for (DevBuf* w : {&b->w_tok, &b->w_qptr, &b->w_cand, &b->w_cand_count, &b->w_keys, &b->w_keys_n, &b->w_bound,
                    &b->w_out_scores, &b->w_out_ids, &b->w_ctr, &b->w_dump, &b->w_sort})
    w->release();
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

static int bm25_validate_tokens(vfi_bm25* b, const int32_t* q_tokens, const int64_t* q_indptr, int64_t nq, int64_t* bytes) {
  int64_t total = 0;
  for (int64_t q = 0; q < nq; ++q) {
    const int64_t t0 = q_indptr[q], t1 = q_indptr[q + 1];
    if (t1 < t0) return fail(VFI_ERR_INVALID, "q_indptr must be non-decreasing");
    if (t1 - t0 > vfi::kBmMaxTok) return fail(VFI_ERR_UNSUPPORTED, "a query has more than 64 tokens");
    for (int64_t i = t0; i < t1; ++i) {
      const int32_t t = q_tokens[i];
      if (t < 0 || t >= b->n_vocab) return fail(VFI_ERR_INVALID, "token id out of range (drop unknown tokens before the call)");
      total += (b->h_indptr[t + 1] - b->h_indptr[t]) * 8;
    }
  }
  *bytes = total;
  return VFI_OK;
}

int vfi_bm25_search(vfi_bm25_t* b, const int32_t* q_tokens, const int64_t* q_indptr, int64_t nq, int k, float* out_scores,
                    int64_t* out_ids, int mem, void* stream) {
  if (!b || !q_indptr || nq < 0 || (nq > 0 && (!out_scores || !out_ids))) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_search");
  if (k <= 0) return fail(VFI_ERR_INVALID, "k must be positive");
  if (k > VFI_MAX_K / 2) return fail(VFI_ERR_UNSUPPORTED, "bm25 top-k supports k <= 1024 (use vfi_bm25_score_all for k = N)");
  if (nq == 0) return VFI_OK;
  if (mem != VFI_MEM_HOST && mem != VFI_MEM_DEVICE) return fail(VFI_ERR_INVALID, "vfi_bm25_search: mem must be VFI_MEM_HOST or VFI_MEM_DEVICE");
  const bool dev_out = mem == VFI_MEM_DEVICE;   // token buffers are host memory either way; `mem` says where the outputs live
  std::lock_guard<std::mutex> lock(b->mu);
  DeviceGuard guard(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t bytes = 0;
  VFI_TRY(bm25_validate_tokens(b, q_tokens, q_indptr, nq, &bytes));
  b->stats.postings_bytes = bytes;
  const int64_t n_tok = q_indptr[nq];
  const int keep = static_cast<int>(round_up(k, 32));
  int cap = 1;
  while (cap < keep + vfi::kBmScan) cap <<= 1;
  const size_t smem = ((sizeof(vfi::Bm25Smem) + 15) & ~size_t(15)) + static_cast<size_t>(cap) * 8;
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vfi::bm25_kernel, vfi::kBmThreads, smem) != cudaSuccess) {
    cudaGetLastError();
    occ = 1;
  }
  const int n_ctas = b->num_sms * std::max(1, occ);
  const int64_t n_ranges = std::max<int64_t>(1, ceil_div(b->n_docs, vfi::kBmRange));
  // about 8 work items per resident CTA (the atomic work queue balances heavy and light queries),
  // segments of at most kBmMaxRanges ranges
  int64_t n_seg = std::max<int64_t>(ceil_div(n_ranges, vfi::kBmMaxRanges),
                                    std::min<int64_t>(n_ranges, ceil_div(static_cast<int64_t>(8) * n_ctas, nq)));
  n_seg = std::min<int64_t>(n_seg, n_ranges);
  const int64_t seg_docs = ceil_div(n_ranges, n_seg) * vfi::kBmRange;
  n_seg = std::max<int64_t>(1, ceil_div(b->n_docs, seg_docs));
  if (n_seg > 1024) return fail(VFI_ERR_UNSUPPORTED, "bm25 shard too large: more than 1024 doc segments (shard the postings)");
  VFI_TRY(b->w_tok.ensure(std::max<size_t>(16, static_cast<size_t>(n_tok) * 4)));
  VFI_TRY(b->w_qptr.ensure(static_cast<size_t>(nq + 1) * 8));
  VFI_TRY(b->w_cand.ensure(static_cast<size_t>(n_seg) * nq * keep * 8));
  VFI_TRY(b->w_cand_count.ensure(static_cast<size_t>(n_seg) * nq * 4));
  VFI_TRY(b->w_keys.ensure(static_cast<size_t>(nq) * keep * 8));
  VFI_TRY(b->w_keys_n.ensure(static_cast<size_t>(nq) * 4));
  VFI_TRY(b->w_bound.ensure(static_cast<size_t>(nq) * 4));
  if (!dev_out) {
    VFI_TRY(b->w_out_scores.ensure(static_cast<size_t>(nq) * k * 4));
    VFI_TRY(b->w_out_ids.ensure(static_cast<size_t>(nq) * k * 8));
  }
  float* d_scores = dev_out ? out_scores : b->w_out_scores.as<float>();
  int64_t* d_ids = dev_out ? out_ids : b->w_out_ids.as<int64_t>();
  VFI_TRY(b->w_ctr.ensure(16 + static_cast<size_t>(nq) * 8));
  if (n_tok > 0) VFI_CUDA(cudaMemcpyAsync(b->w_tok.p, q_tokens, static_cast<size_t>(n_tok) * 4, cudaMemcpyHostToDevice, st));
  VFI_CUDA(cudaMemcpyAsync(b->w_qptr.p, q_indptr, static_cast<size_t>(nq + 1) * 8, cudaMemcpyHostToDevice, st));
  VFI_CUDA(cudaMemsetAsync(b->w_ctr.p, 0, 16 + static_cast<size_t>(nq) * 8, st));
  VFI_CUDA(cudaMemsetAsync(b->w_cand_count.p, 0, static_cast<size_t>(n_seg) * nq * 4, st));
  vfi::Bm25Params p{};
  p.indptr = b->indptr;
  p.indices = b->indices;
  p.data = b->data;
  p.n_docs = b->n_docs;
  p.n_seg = static_cast<int>(n_seg);
  p.seg_docs = seg_docs;
  p.q_tokens = b->w_tok.as<int32_t>();
  p.q_indptr = b->w_qptr.as<int64_t>();
  p.nq = static_cast<int>(nq);
  p.nq_pad = static_cast<int>(nq);
  p.keep = keep;
  p.cap = cap;
  p.all_positive = b->all_positive;
  p.cand = b->w_cand.as<uint64_t>();
  p.cand_count = b->w_cand_count.as<uint32_t>();
  p.work_counter = b->w_ctr.as<uint32_t>();
  p.qtau = reinterpret_cast<unsigned long long*>(b->w_ctr.as<uint8_t>() + 16);
  p.dump = nullptr;
  const int grid = static_cast<int>(std::min<int64_t>(n_ctas, nq * n_seg));
  if (b->profile) cudaEventRecord(b->ev0, st);
  vfi::bm25_kernel<<<grid, vfi::kBmThreads, smem, st>>>(p);
  LAUNCHED();
  if (b->profile) cudaEventRecord(b->ev1, st);
  VFI_CUDA(cudaGetLastError());
  b->stats.launches++;
  vfi::cand_reduce_kernel<vfi::SelectSmem><<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(
      b->w_cand.as<uint64_t>(), b->w_cand_count.as<uint32_t>(), static_cast<int>(n_seg), static_cast<int>(nq), keep, keep, nullptr,
      b->w_keys.as<uint64_t>(), b->w_keys_n.as<uint32_t>(), b->w_bound.as<float>());
  LAUNCHED();
  vfi::finalize_kernel<<<static_cast<unsigned>(nq), 256, sizeof(vfi::SelectSmem), st>>>(
      b->w_keys.as<uint64_t>(), keep, keep, nullptr, b->w_keys_n.as<uint32_t>(), k, b->id_offset, nullptr, nullptr, 0,
      d_scores, d_ids, nullptr, nullptr);
  LAUNCHED();
  if (b->all_positive) {
    vfi::bm25_zero_fill_kernel<<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
        d_scores, d_ids, static_cast<int>(nq), k, b->n_docs, b->id_offset);
    LAUNCHED();
  }
  VFI_CUDA(cudaGetLastError());
  if (!dev_out) {
    VFI_CUDA(cudaMemcpyAsync(out_scores, d_scores, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, st));
    VFI_CUDA(cudaMemcpyAsync(out_ids, d_ids, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, st));
  }
  // the token staging buffers are reused by the next call and the host token arrays may be freed by the caller: wait
  VFI_CUDA(cudaStreamSynchronize(st));
  if (b->profile) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) {
      b->stats.score_ms_total += ms;
      b->stats.score_ms_samples++;
    } else {
      cudaGetLastError();
    }
  }
  return VFI_OK;
}

int vfi_bm25_score_all(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, float* out, int mem, void* stream) {
  if (!b || !out || n_tokens < 0 || (n_tokens > 0 && !q_tokens)) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_score_all");
  std::lock_guard<std::mutex> lock(b->mu);
  DeviceGuard guard(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t qptr[2] = {0, n_tokens};
  int64_t bytes = 0;
  VFI_TRY(bm25_validate_tokens(b, q_tokens, qptr, 1, &bytes));
  if (b->n_docs == 0) return VFI_OK;
  const int cap = 2048;
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
  VFI_TRY(b->w_tok.ensure(std::max<size_t>(16, static_cast<size_t>(n_tokens) * 4)));
  VFI_TRY(b->w_qptr.ensure(16));
  VFI_TRY(b->w_ctr.ensure(16));
  float* dump = out;
  if (mem == VFI_MEM_HOST) {
    VFI_TRY(b->w_dump.ensure(static_cast<size_t>(b->n_docs) * 4));
    dump = b->w_dump.as<float>();
  }
  if (n_tokens > 0) VFI_CUDA(cudaMemcpyAsync(b->w_tok.p, q_tokens, static_cast<size_t>(n_tokens) * 4, cudaMemcpyHostToDevice, st));
  VFI_CUDA(cudaMemcpyAsync(b->w_qptr.p, qptr, 16, cudaMemcpyHostToDevice, st));
  VFI_CUDA(cudaMemsetAsync(b->w_ctr.p, 0, 16, st));
  vfi::Bm25Params p{};
  p.indptr = b->indptr;
  p.indices = b->indices;
  p.data = b->data;
  p.n_docs = b->n_docs;
  p.n_seg = static_cast<int>(n_seg);
  p.seg_docs = seg_docs;
  p.q_tokens = b->w_tok.as<int32_t>();
  p.q_indptr = b->w_qptr.as<int64_t>();
  p.nq = 1;
  p.nq_pad = 1;
  p.keep = 32;
  p.cap = cap;
  p.all_positive = b->all_positive;
  p.cand = nullptr;
  p.cand_count = nullptr;
  p.work_counter = b->w_ctr.as<uint32_t>();
  p.qtau = nullptr;
  p.dump = dump;
  vfi::bm25_kernel<<<static_cast<int>(std::min<int64_t>(n_ctas, n_seg)), vfi::kBmThreads, smem, st>>>(p);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  if (mem == VFI_MEM_HOST) VFI_CUDA(cudaMemcpyAsync(out, dump, static_cast<size_t>(b->n_docs) * 4, cudaMemcpyDeviceToHost, st));
  VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
}


}  // extern "C"

namespace vfi {
__global__ void scores_to_keys_kernel(const float* __restrict__ s, int64_t n, uint64_t* __restrict__ keys) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = make_key(s[i], static_cast<uint32_t>(i));
}
__global__ void keys_to_ranked_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t id_offset,
                                      float* __restrict__ scores, int64_t* __restrict__ ids) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    scores[i] = key_score(keys[i]);
    ids[i] = static_cast<int64_t>(key_id(keys[i])) + id_offset;
  }
}
}  // namespace vfi

extern "C" int vfi_bm25_rank_all(vfi_bm25_t* b, const int32_t* q_tokens, int64_t n_tokens, float* out_scores,
                                 int64_t* out_ids, void* stream) {
  if (!b || !out_scores || !out_ids) return fail(VFI_ERR_INVALID, "bad argument to vfi_bm25_rank_all");
  const int64_t n = b->n_docs;
  if (n == 0) return VFI_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevBuf scores, keys_in, keys_out, tmp, ids;
  struct Cleanup {
    std::vector<DevBuf*> v;
    ~Cleanup() { for (DevBuf* d : v) d->release(); }
  } cleanup{{&scores, &keys_in, &keys_out, &tmp, &ids}};
  {
    DeviceGuard guard(b->device);
    VFI_TRY(scores.ensure(static_cast<size_t>(n) * 4));
  }
  VFI_TRY(vfi_bm25_score_all(b, q_tokens, n_tokens, scores.as<float>(), VFI_MEM_DEVICE, stream));
  std::lock_guard<std::mutex> lock(b->mu);
  DeviceGuard guard(b->device);
  VFI_TRY(keys_in.ensure(static_cast<size_t>(n) * 8));
  VFI_TRY(keys_out.ensure(static_cast<size_t>(n) * 8));
  VFI_TRY(ids.ensure(static_cast<size_t>(n) * 8));
  const unsigned blocks = static_cast<unsigned>(ceil_div(n, 256));
  vfi::scores_to_keys_kernel<<<blocks, 256, 0, st>>>(scores.as<float>(), n, keys_in.as<uint64_t>());
  LAUNCHED();
  size_t tmp_bytes = 0;
  VFI_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp_bytes, keys_in.as<uint64_t>(), keys_out.as<uint64_t>(), n, 0, 64, st));
  VFI_TRY(tmp.ensure(tmp_bytes));
  VFI_CUDA(cub::DeviceRadixSort::SortKeysDescending(tmp.p, tmp_bytes, keys_in.as<uint64_t>(), keys_out.as<uint64_t>(), n, 0, 64, st));
  vfi::keys_to_ranked_kernel<<<blocks, 256, 0, st>>>(keys_out.as<uint64_t>(), n, b->id_offset, scores.as<float>(), ids.as<int64_t>());
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  VFI_CUDA(cudaMemcpyAsync(out_scores, scores.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, st));
  VFI_CUDA(cudaMemcpyAsync(out_ids, ids.p, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, st));
  VFI_CUDA(cudaStreamSynchronize(st));
  return VFI_OK;
}

extern "C" {
// ---- host-side text routines (no device involved) -------------------------------------------------
int vfi_stem_english(const char* words, const int64_t* offsets, int64_t n_words, char* out, int64_t out_cap,
                     int64_t* out_offsets) {
  if (n_words < 0 || !offsets || !out_offsets || (n_words > 0 && (!words || !out)))
    return fail(VFI_ERR_INVALID, "bad argument to vfi_stem_english");
  vfi_text::EnglishStemmer st;
  int64_t pos = 0;
  out_offsets[0] = 0;
  for (int64_t i = 0; i < n_words; ++i) {
    const int64_t a = offsets[i], b = offsets[i + 1];
    if (a < 0 || b < a) return fail(VFI_ERR_INVALID, "vfi_stem_english: offsets must be non-decreasing");
    const std::string& r = st.stem(words + a, static_cast<size_t>(b - a));
    if (pos + static_cast<int64_t>(r.size()) > out_cap) return fail(VFI_ERR_INVALID, "vfi_stem_english: out_cap too small");
    std::memcpy(out + pos, r.data(), r.size());
    pos += static_cast<int64_t>(r.size());
    out_offsets[i + 1] = pos;
  }
  return VFI_OK;
}

int vfi_tokenize_ascii(const char* text, int64_t len, int64_t* starts, int64_t* lens, int64_t cap, int64_t* n_tokens) {
  if (len < 0 || cap < 0 || !n_tokens || (len > 0 && !text) || (cap > 0 && (!starts || !lens)))
    return fail(VFI_ERR_INVALID, "bad argument to vfi_tokenize_ascii");
  const int64_t n = vfi_text::tokenize_ascii(text, len, starts, lens, cap);
  if (n < 0) return fail(VFI_ERR_UNSUPPORTED, "vfi_tokenize_ascii: non-ASCII text");
  *n_tokens = n;
  return VFI_OK;
}

}
