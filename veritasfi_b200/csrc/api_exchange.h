// api_exchange.h — the exchange object shared by api_exchange.cu (vfi_exchange_*) and api_dense.cu (the rescoring kernel pushes
// a sharded batch's rows itself: vfi_index_search_begin_push).
#pragma once
#include "api_common.h"
#include "peer_exchange.cuh"

struct vfi_exchange {
  int device = 0, rank = 0, world = 1, max_k = 0, num_sms = 148, max_resident = 148;
  int64_t max_nq = 0;
  size_t win_bytes = 0, total_bytes = 0;
  uint8_t* local = nullptr;                 // own window + flags (cudaMalloc, exported by CUDA IPC)
  uint8_t* peer[vfi::kMaxPeers] = {};       // every rank's window as mapped here (peer[rank] == local)
  bool connected = false;
  uint32_t epoch = 0;
  uint64_t timeout_ms = 30000;
  // "any rank's rows were not final": one slot per launch in flight, device accumulators + a mapped host word
  static constexpr int kSlots = 64;
  int* d_state = nullptr;                   // [kSlots][2]: OR of the fail bits, finished CTAs (zero between launches)
  int* h_any = nullptr;                     // [kSlots] mapped pinned host memory
  int* h_any_dev = nullptr;
  int next_slot = 0;
  std::mutex mu;
};

namespace vfi_host {
// Reserve the next epoch of `ex` for rows that the caller's own kernel will push (nq rows of k keys): fills `t`.
int exchange_reserve_push(vfi_exchange* ex, int64_t nq, int k, vfi::PushTarget* t);
// Push rows [nq][k] (global ids, -1 = padding) for the epoch in `t` with a standalone kernel (the route of a batch whose
// rescoring kernel could not do it); fail: device int, > 0 = rows not final (may be null).
int exchange_push_rows(vfi_exchange* ex, const vfi::PushTarget& t, const float* scores, const int64_t* ids, int64_t nq, int k,
                       const int* fail, cudaStream_t st);
}  // namespace vfi_host
