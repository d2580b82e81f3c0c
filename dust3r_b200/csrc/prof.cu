#include "prof.h"
#include "d3r_common.cuh"
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace d3r {
namespace prof {

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_enabled{0};
struct Rec { const char* tag; cudaEvent_t a, b; double flops, bytes; int launches; std::string detail; };
static std::vector<Rec> g_recs;
static std::mutex g_mu;

Scope::Scope(const char* tag, cudaStream_t s, double flops, double bytes, int launches, const char* detail) : idx(-1), st(s) {
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  if (!g_enabled.load(std::memory_order_relaxed)) return;
  Rec r{tag, nullptr, nullptr, flops, bytes, launches, detail ? detail : ""};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_mu);
  idx = (int)g_recs.size();
  g_recs.push_back(r);
}
Scope::~Scope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[idx].b, st);
}

}  // namespace prof
}  // namespace d3r

using namespace d3r::prof;

extern "C" long long d3r_launch_count(void) { return g_launches.load(); }
extern "C" void d3r_launch_count_reset(void) { g_launches.store(0); }
extern "C" void d3r_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_recs.clear();
  g_enabled.store(on);
}
extern "C" int d3r_prof_report(char* buf, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  struct Agg { long long count = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Agg& a = agg[r.tag];
    a.count += r.launches; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  std::string s = "{";
  bool first = true;
  for (auto& kv : agg) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"count\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.count, kv.second.ms, kv.second.flops, kv.second.bytes);
    s += tmp;
    first = false;
  }
  s += "}";
  if ((int)s.size() + 1 > cap) return -1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}
extern "C" int d3r_prof_dump(char* buf, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::string s = "[";
  bool first = true;
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    char tmp[384];
    snprintf(tmp, sizeof(tmp), "%s{\"tag\": \"%s\", \"detail\": \"%s\", \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             r.tag, r.detail.c_str(), ms, r.flops, r.bytes);
    s += tmp;
    first = false;
  }
  s += "]";
  if ((int)s.size() + 1 > cap) return -1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}
