// api_common.h — host-side plumbing shared by the translation units of libvfi.so (api_*.cu): status/error
// reporting, device guards, grow-only device buffers, staged host<->device arguments.  No kernels here.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vfi.h"

namespace vfi_host {

// defined in api_core.cu
int fail(int code, const std::string& msg);
void count_launch();
int device_props(int device, cudaDeviceProp* prop);

#define VFI_CUDA(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      return ::vfi_host::fail(e__ == cudaErrorMemoryAllocation ? VFI_ERR_NOMEM : VFI_ERR_CUDA,         \
                              std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
    }                                                                                                  \
  } while (0)
#define VFI_TRY(expr)              \
  do {                             \
    int s__ = (expr);              \
    if (s__ != VFI_OK) return s__; \
  } while (0)
#define LAUNCHED() (::vfi_host::count_launch())

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

// Host arguments of the small entry points (merge / fusion) are staged through a per-device pool of grow-only
// buffers instead of a cudaMalloc/cudaFree pair per call.  A Staged object borrows one arena for the duration of
// the call (arenas are handed out under a mutex, so concurrent callers get different ones).
struct StageArena {
  DevBuf buf;
  bool busy = false;
  int device = -1;
};
StageArena* borrow_arena(int device);   // api_core.cu
void return_arena(StageArena* a);

struct Staged {
  StageArena* arena = nullptr;
  size_t used = 0;
  size_t need = 0;
  int device;
  struct Item { const void* src; void* host_dst; size_t bytes; size_t off; bool in; };
  std::vector<Item> items;
  explicit Staged(int dev) : device(dev) {}
  ~Staged() {
    if (arena) return_arena(arena);
  }
  // declare phase: every host buffer reserves a slice; commit() allocates once and copies the inputs
  template <class T>
  void in(const T* src, size_t count, int mem, const T** out) {
    if (mem == VFI_MEM_DEVICE) { *out = src; return; }
    items.push_back({src, nullptr, count * sizeof(T), need, true});
    slots.push_back(reinterpret_cast<void**>(const_cast<T**>(out)));
    need += (count * sizeof(T) + 255) & ~size_t(255);
  }
  template <class T>
  void out(T* dst, size_t count, int mem, T** dev) {
    if (mem == VFI_MEM_DEVICE) { *dev = dst; return; }
    items.push_back({nullptr, dst, count * sizeof(T), need, false});
    slots.push_back(reinterpret_cast<void**>(dev));
    need += (count * sizeof(T) + 255) & ~size_t(255);
  }
  int commit(cudaStream_t st) {
    if (items.empty()) return VFI_OK;
    arena = borrow_arena(device);
    VFI_TRY(arena->buf.ensure(std::max<size_t>(need, 256)));
    uint8_t* base = arena->buf.as<uint8_t>();
    for (size_t i = 0; i < items.size(); ++i) {
      *slots[i] = base + items[i].off;
      if (items[i].in && items[i].bytes) VFI_CUDA(cudaMemcpyAsync(base + items[i].off, items[i].src, items[i].bytes, cudaMemcpyHostToDevice, st));
    }
    return VFI_OK;
  }
  int back(cudaStream_t st) {
    if (items.empty()) return VFI_OK;
    uint8_t* base = arena->buf.as<uint8_t>();
    for (const Item& it : items)
      if (!it.in && it.bytes) VFI_CUDA(cudaMemcpyAsync(it.host_dst, base + it.off, it.bytes, cudaMemcpyDeviceToHost, st));
    VFI_CUDA(cudaStreamSynchronize(st));
    return VFI_OK;
  }

 private:
  std::vector<void**> slots;
};

}  // namespace vfi_host
