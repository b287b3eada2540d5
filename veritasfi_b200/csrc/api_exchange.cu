// api_exchange.cu — K3p: the multi-GPU exchange + merge over NVLink peer memory (vfi_exchange_* of include/vfi.h).
// One process per GPU; host side of csrc/peer_exchange.cuh.
#include "api_exchange.h"

using namespace vfi_host;

namespace vfi {
// push + wait + merge in one launch (vfi_exchange_merge / _flagged)
__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeParams p) {
  extern __shared__ uint8_t smem_raw[];
  exchange_push_phase(p);
  exchange_wait_merge_phase(p, reinterpret_cast<SelectSmem*>(smem_raw));
}
// the halves: a rank whose rows were not pushed by its rescoring kernel pushes them here; everyone waits and merges
__global__ void __launch_bounds__(256) exchange_push_kernel(const ExchangeParams p) { exchange_push_phase(p); }
__global__ void __launch_bounds__(256) exchange_wait_merge_kernel(const ExchangeParams p) {
  extern __shared__ uint8_t smem_raw[];
  exchange_publish_phase(p);     // the rows were stored by the kernel(s) before this one in the stream
  exchange_wait_merge_phase(p, reinterpret_cast<SelectSmem*>(smem_raw));
}

}  // namespace vfi

namespace {
void fill_params(vfi_exchange* ex, vfi::ExchangeParams* p) {
  for (int r = 0; r < ex->world; ++r) {
    p->win[r] = reinterpret_cast<uint64_t*>(ex->peer[r]);
    p->flags[r] = reinterpret_cast<uint32_t*>(ex->peer[r] + ex->win_bytes);
  }
  p->rank = ex->rank;
  p->world = ex->world;
  p->max_nq = ex->max_nq;
  p->max_k = ex->max_k;
  p->timeout_ns = ex->timeout_ms * 1000000ull;
}
uint32_t next_epoch(vfi_exchange* ex) {   // under ex->mu; every rank makes the same sequence of calls (a collective)
  ex->epoch++;
  if (ex->epoch >= 0x7FFFFFFFu) ex->epoch = 1;   // 31 bits travel in the flag word (never reached in practice)
  return ex->epoch;
}
}  // namespace

namespace vfi_host {

int exchange_reserve_push(vfi_exchange* ex, int64_t nq, int k, vfi::PushTarget* t) {
  if (!ex || !t || nq <= 0 || k <= 0) return fail(VFI_ERR_INVALID, "bad argument to the exchange push");
  if (!ex->connected) return fail(VFI_ERR_INVALID, "exchange used before vfi_exchange_connect");
  if (nq > ex->max_nq || k > ex->max_k) return fail(VFI_ERR_INVALID, "nq or k exceeds the window geometry given at create");
  std::lock_guard<std::mutex> lock(ex->mu);
  for (int r = 0; r < vfi::kMaxPeers; ++r) { t->win[r] = nullptr; t->flags[r] = nullptr; }
  for (int r = 0; r < ex->world; ++r) {
    t->win[r] = reinterpret_cast<uint64_t*>(ex->peer[r]);
    t->flags[r] = reinterpret_cast<uint32_t*>(ex->peer[r] + ex->win_bytes);
  }
  t->rank = ex->rank;
  t->world = ex->world;
  t->max_nq = ex->max_nq;
  t->max_k = ex->max_k;
  t->epoch = next_epoch(ex);
  return VFI_OK;
}

int exchange_push_rows(vfi_exchange* ex, const vfi::PushTarget& t, const float* scores, const int64_t* ids, int64_t nq, int k,
                       const int* fail_flag, cudaStream_t st) {
  vfi::ExchangeParams p{};
  fill_params(ex, &p);
  p.nq = static_cast<int>(nq);
  p.k = k;
  p.k_out = k;
  p.epoch = t.epoch;
  p.scores = scores;
  p.ids = ids;
  p.fail_a = fail_flag;
  const int grid = static_cast<int>(std::min<int64_t>(nq, ex->max_resident));
  vfi::exchange_push_kernel<<<grid, 256, 0, st>>>(p);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return VFI_OK;
}

}  // namespace vfi_host

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
  if (cudaMalloc(&ex->d_state, sizeof(int) * 2 * vfi_exchange::kSlots) != cudaSuccess ||
      cudaMemset(ex->d_state, 0, sizeof(int) * 2 * vfi_exchange::kSlots) != cudaSuccess ||
      cudaHostAlloc(&ex->h_any, sizeof(int) * vfi_exchange::kSlots, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
      cudaHostGetDevicePointer(&ex->h_any_dev, ex->h_any, 0) != cudaSuccess) {
    cudaGetLastError();
    if (ex->d_state) cudaFree(ex->d_state);
    if (ex->h_any) cudaFreeHost(ex->h_any);
    cudaFree(p);
    delete ex;
    return fail(VFI_ERR_NOMEM, "exchange state allocation failed");
  }
  std::memset(ex->h_any, 0, sizeof(int) * vfi_exchange::kSlots);
  cudaDeviceSynchronize();
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

int vfi_exchange_merge_flagged(vfi_exchange_t* ex, const float* scores, const int64_t* ids, int64_t nq, int k, int k_out,
                               float* out_scores, int64_t* out_ids, const int* fail_a, const int* fail_b, int* any_fail,
                               void* stream) {
  if (!ex || nq < 0 || k <= 0 || k_out <= 0 || (nq > 0 && (!scores || !ids || !out_scores || !out_ids)))
    return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_merge");
  if (!ex->connected) return fail(VFI_ERR_INVALID, "vfi_exchange_merge before vfi_exchange_connect");
  if (nq > ex->max_nq || k > ex->max_k) return fail(VFI_ERR_INVALID, "nq or k exceeds the window geometry given at create");
  if (k_out > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k_out exceeds VFI_MAX_K");
  // An empty batch launches nothing and must not consume an epoch: the double-buffer argument (peer_exchange.cuh)
  // needs every epoch's kernel to have waited for the peers' flags.
  if (nq == 0) return VFI_OK;
  std::lock_guard<std::mutex> lock(ex->mu);
  DeviceGuard guard(ex->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  vfi::ExchangeParams p{};
  fill_params(ex, &p);
  p.epoch = next_epoch(ex);
  p.nq = static_cast<int>(nq);
  p.k = k;
  p.k_out = k_out;
  p.scores = scores;
  p.ids = ids;
  p.out_scores = out_scores;
  p.out_ids = out_ids;
  p.fail_a = fail_a;
  p.fail_b = fail_b;
  p.any_fail = any_fail;
  const int grid = static_cast<int>(std::min<int64_t>(nq, ex->max_resident));
  vfi::exchange_merge_kernel<<<grid, 256, sizeof(vfi::SelectSmem), static_cast<cudaStream_t>(stream)>>>(p);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  return VFI_OK;
}

int vfi_exchange_merge(vfi_exchange_t* ex, const float* scores, const int64_t* ids, int64_t nq, int k, int k_out,
                       float* out_scores, int64_t* out_ids, void* stream) {
  return vfi_exchange_merge_flagged(ex, scores, ids, nq, k, k_out, out_scores, out_ids, nullptr, nullptr, nullptr, stream);
}

int vfi_exchange_merge_pushed(vfi_exchange_t* ex, int64_t nq, int k, int k_out, float* out_scores, int64_t* out_ids,
                              const int* fail_flag, int* slot, void* stream) {
  if (!ex || nq <= 0 || k <= 0 || k_out <= 0 || !out_scores || !out_ids || !slot)
    return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_merge_pushed");
  if (!ex->connected) return fail(VFI_ERR_INVALID, "vfi_exchange_merge_pushed before vfi_exchange_connect");
  if (nq > ex->max_nq || k > ex->max_k) return fail(VFI_ERR_INVALID, "nq or k exceeds the window geometry given at create");
  if (k_out > VFI_MAX_K) return fail(VFI_ERR_UNSUPPORTED, "k_out exceeds VFI_MAX_K");
  std::lock_guard<std::mutex> lock(ex->mu);
  DeviceGuard guard(ex->device);
  if (!guard.ok) return fail(VFI_ERR_CUDA, "cudaSetDevice failed");
  if (ex->epoch == 0) return fail(VFI_ERR_INVALID, "vfi_exchange_merge_pushed: no epoch has been pushed");
  vfi::ExchangeParams p{};
  fill_params(ex, &p);
  p.epoch = ex->epoch;               // the epoch reserved by the last vfi_index_search_begin_push
  p.nq = static_cast<int>(nq);
  p.k = k;
  p.k_out = k_out;
  p.out_scores = out_scores;
  p.out_ids = out_ids;
  p.fail_a = fail_flag;
  const int s = ex->next_slot;
  ex->next_slot = (s + 1) % vfi_exchange::kSlots;
  p.fail_acc = ex->d_state + 2 * s;
  p.done_ctas = ex->d_state + 2 * s + 1;
  p.host_any_fail = ex->h_any_dev + s;
  const int grid = static_cast<int>(std::min<int64_t>(nq, ex->max_resident));
  vfi::exchange_wait_merge_kernel<<<grid, 256, sizeof(vfi::SelectSmem), static_cast<cudaStream_t>(stream)>>>(p);
  LAUNCHED();
  VFI_CUDA(cudaGetLastError());
  *slot = s;
  return VFI_OK;
}

int vfi_exchange_any_fail(vfi_exchange_t* ex, int slot, int* any_fail) {
  if (!ex || !any_fail || slot < 0 || slot >= vfi_exchange::kSlots) return fail(VFI_ERR_INVALID, "bad argument to vfi_exchange_any_fail");
  *any_fail = *reinterpret_cast<volatile int*>(ex->h_any + slot);
  return VFI_OK;
}

int vfi_exchange_destroy(vfi_exchange_t* ex) {
  if (!ex) return VFI_OK;
  DeviceGuard guard(ex->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < ex->world; ++r)
    if (r != ex->rank && ex->peer[r]) cudaIpcCloseMemHandle(ex->peer[r]);
  if (ex->local) cudaFree(ex->local);
  if (ex->d_state) cudaFree(ex->d_state);
  if (ex->h_any) cudaFreeHost(ex->h_any);
  cudaGetLastError();
  delete ex;
  return VFI_OK;
}

}  // extern "C"
