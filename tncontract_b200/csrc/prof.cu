// Optional per-kernel-class device timing (CUDA events on the launching stream).
// Off by default; bench.py switches it on for ONE extra pass after the timed
// steps to attribute device time to kernel classes and to compute the roofline
// figures of the dominant kernel from live measurements.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace tnb {

bool g_prof_on = false;

struct ProfRec { int cls; cudaEvent_t e0, e1; double work; long long launches; };
static std::mutex g_mu;  // the record / event pools are shared by every host thread driving the library
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static double g_ms[KC_COUNT], g_work[KC_COUNT];
static long long g_launch[KC_COUNT], g_scopes[KC_COUNT];

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int c, cudaStream_t s, double w) : cls(c), st(s), work(w), l0(g_launches.load()), idx(-1) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_mu);
  ProfRec r;
  r.cls = c; r.work = w; r.launches = 0;
  r.e0 = get_event(); r.e1 = get_event();
  cudaEventRecord(r.e0, s);
  idx = (long long)g_recs.size();
  g_recs.push_back(r);
}

ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lock(g_mu);
  if ((size_t)idx >= g_recs.size()) return;  // profile was drained while this scope was open
  ProfRec& r = g_recs[(size_t)idx];
  r.launches = g_launches.load() - l0;
  r.work = work;
  cudaEventRecord(r.e1, st);
}

static void drain() {
  std::lock_guard<std::mutex> lock(g_mu);
  for (ProfRec& r : g_recs) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      g_ms[r.cls] += ms; g_work[r.cls] += r.work; g_launch[r.cls] += r.launches; g_scopes[r.cls] += 1;
    }
    g_pool.push_back(r.e0);
    g_pool.push_back(r.e1);
  }
  g_recs.clear();
}

}  // namespace tnb

using namespace tnb;

extern "C" int tnb_profile_enable(int on) {
  drain();
  if (on) {
    for (int i = 0; i < KC_COUNT; ++i) { g_ms[i] = 0; g_work[i] = 0; g_launch[i] = 0; g_scopes[i] = 0; }
  }
  g_prof_on = (on != 0);
  return 0;
}

extern "C" int tnb_profile_get(int cls, double* ms, double* work, long long* launches, long long* scopes) {
  if (cls < 0 || cls >= KC_COUNT) return TNB_E_ARG;
  drain();
  if (ms) *ms = g_ms[cls];
  if (work) *work = g_work[cls];
  if (launches) *launches = g_launch[cls];
  if (scopes) *scopes = g_scopes[cls];
  return 0;
}
