// api_exchange.cu — K3p: the multi-GPU exchange + merge over NVLink peer memory (vfi_exchange_* of include/vfi.h).
// One process per GPU; host side of csrc/peer_exchange.cuh.
#include "api_common.h"
#include "peer_exchange.cuh"

using namespace vfi_host;

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
  ex->epoch++;                       // every rank must make the same sequence of calls (a collective)
  if (ex->epoch >= 0x7FFFFFFFu) ex->epoch = 1;   // 31 bits travel in the flag word (never reached in practice)
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
  p.fail_a = fail_a;
  p.fail_b = fail_b;
  p.any_fail = any_fail;
  p.timeout_ns = ex->timeout_ms * 1000000ull;
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
