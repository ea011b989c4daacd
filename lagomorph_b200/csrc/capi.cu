// capi.cu -- error state, debug mode, launch counter of the C ABI.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace lgm {

static thread_local char g_err[512] = "";
static std::atomic<int> g_debug{0};
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool debug_mode() { return g_debug.load(std::memory_order_relaxed) != 0; }

// Launch-error check after enqueueing. In debug mode (the reference's
// set_debug_mode, include/defs.h:15-23) also synchronise the stream, and unlike
// the reference RETURN the error instead of printing it.
int finish(cudaStream_t s, const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess && debug_mode()) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &st);
    if (st == cudaStreamCaptureStatusNone) e = cudaStreamSynchronize(s);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear
    return set_error((int)e, "%s: CUDA error: %s", what, cudaGetErrorString(e));
  }
  return LGM_OK;
}

}  // namespace lgm

extern "C" int lgm_version(void) { return 1; }
extern "C" const char* lgm_last_error(void) { return lgm::g_err; }
extern "C" void lgm_set_debug_mode(int on) { lgm::g_debug.store(on ? 1 : 0); }
extern "C" int lgm_get_debug_mode(void) { return lgm::g_debug.load(); }
extern "C" int64_t lgm_launch_count(void) { return (int64_t)lgm::g_launches.load(); }
