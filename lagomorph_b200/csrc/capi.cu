// capi.cu -- error state, debug mode, launch counter of the C ABI.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
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


// Profile mode (bench only): an event is recorded after every launch; the time of a launch is
// the gap to the previous event on the stream, i.e. kernel time plus its launch gap.
struct ProfEntry { const char* name; cudaEvent_t ev; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;

void count_launch(const char* name, cudaStream_t s) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfEntry e;
    e.name = name;
    if (cudaEventCreate(&e.ev) == cudaSuccess) {
      cudaEventRecord(e.ev, s);
      g_prof.push_back(e);
    }
  }
}
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

extern "C" int lgm_profile_begin(void* stream) {
  std::lock_guard<std::mutex> lk(lgm::g_prof_mu);
  for (auto& e : lgm::g_prof) cudaEventDestroy(e.ev);
  lgm::g_prof.clear();
  lgm::ProfEntry e;
  e.name = "";
  cudaError_t rc = cudaEventCreate(&e.ev);
  if (rc != cudaSuccess) return lgm::set_error((int)rc, "lgm_profile_begin: %s", cudaGetErrorString(rc));
  cudaEventRecord(e.ev, (cudaStream_t)stream);
  lgm::g_prof.push_back(e);
  lgm::g_prof_on = true;
  return LGM_OK;
}

extern "C" int lgm_profile_end(char* json, int64_t json_bytes) {
  std::lock_guard<std::mutex> lk(lgm::g_prof_mu);
  lgm::g_prof_on = false;
  std::map<std::string, std::pair<long long, double>> agg;
  cudaError_t rc = cudaSuccess;
  if (!lgm::g_prof.empty()) rc = cudaEventSynchronize(lgm::g_prof.back().ev);
  for (size_t i = 1; rc == cudaSuccess && i < lgm::g_prof.size(); ++i) {
    float ms = 0.f;
    rc = cudaEventElapsedTime(&ms, lgm::g_prof[i - 1].ev, lgm::g_prof[i].ev);
    auto& a = agg[lgm::g_prof[i].name];
    a.first += 1;
    a.second += ms;
  }
  for (auto& e : lgm::g_prof) cudaEventDestroy(e.ev);
  lgm::g_prof.clear();
  if (rc != cudaSuccess) return lgm::set_error((int)rc, "lgm_profile_end: %s", cudaGetErrorString(rc));
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ",
             kv.first.c_str(), kv.second.first, kv.second.second);
    out += buf;
    first = false;
  }
  out += "}";
  if ((int64_t)out.size() + 1 > json_bytes) return lgm::set_error(LGM_ENOSPC, "lgm_profile_end: buffer too small");
  memcpy(json, out.c_str(), out.size() + 1);
  return LGM_OK;
}
