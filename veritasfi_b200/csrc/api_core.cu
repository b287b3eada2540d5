// api_core.cu — library-level entries of include/vfi.h (version, errors, device probing, launch counter) and the
// host-only text routines.  libvfi.so is built from the api_*.cu translation units (veritasfi_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC -c api_X.cu ; nvcc -shared *.o
// There is no CPU implementation in this library; every compute entry needs an sm_100 device.
#include <memory>

#include "api_common.h"
#include "text_host.h"

namespace vfi_host {

namespace {
thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
std::mutex g_arena_mu;
std::vector<std::unique_ptr<StageArena>> g_arenas;
}  // namespace

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

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

StageArena* borrow_arena(int device) {
  std::lock_guard<std::mutex> lock(g_arena_mu);
  for (auto& a : g_arenas)
    if (!a->busy && a->device == device) {
      a->busy = true;
      return a.get();
    }
  g_arenas.emplace_back(new StageArena());
  g_arenas.back()->busy = true;
  g_arenas.back()->device = device;
  return g_arenas.back().get();
}
void return_arena(StageArena* a) {
  std::lock_guard<std::mutex> lock(g_arena_mu);
  a->busy = false;
}

}  // namespace vfi_host

using vfi_host::fail;

extern "C" {

int vfi_abi_version(void) { return VFI_ABI_VERSION; }
const char* vfi_last_error(void) { return vfi_host::g_err.c_str(); }
int vfi_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int64_t vfi_launch_count(void) { return vfi_host::g_launches.load(); }

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

}  // extern "C"
