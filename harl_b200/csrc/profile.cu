// Launch counter and optional per-kernel timing (CUDA events on the launching stream).
//
// Every kernel launch in the library ends in HB_LAUNCH_DONE -> note_launch().  With profiling
// off this is one relaxed atomic increment.  With profiling on (bench.py's roofline pass) an
// event is recorded after every launch; a kernel's duration is the gap to the previous event on
// the stream, i.e. it includes any launch gap in front of it (conservative).
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace hb {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_profiling{0};
static std::mutex g_mu;
struct Mark { cudaEvent_t ev; const char* what; };
static std::vector<Mark> g_marks;

void note_launch(const char* what, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_profiling.load(std::memory_order_relaxed)) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  std::lock_guard<std::mutex> lk(g_mu);
  g_marks.push_back({ev, what});
}

// Shape-qualified label, interned for the process lifetime; plain `base` when profiling is off.
const char* shape_label(const char* base, int64_t m, int n, int k) {
  if (!g_profiling.load(std::memory_order_relaxed)) return base;
  static std::map<std::string, const char*> interned;
  char buf[96];
  snprintf(buf, sizeof buf, "%s[M%lld,N%d,K%d]", base, (long long)m, n, k);
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = interned.find(buf);
  if (it != interned.end()) return it->second;
  char* keep = strdup(buf);
  interned[buf] = keep;
  return keep;
}

}  // namespace hb

extern "C" {

uint64_t hb_kernel_launch_count(void) { return hb::g_launches.load(); }

int hb_profile_begin(void* stream) {
  using namespace hb;
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
  cudaEvent_t ev;
  cudaError_t e = cudaEventCreate(&ev);
  if (e != cudaSuccess) return cuda_fail(e, "hb_profile_begin");
  cudaEventRecord(ev, (cudaStream_t)stream);
  g_marks.push_back({ev, "<begin>"});
  g_profiling.store(1);
  return HB_OK;
}

// Stops profiling, waits for the recorded events and writes one line per kernel label:
// "label count total_ms\n", sorted by total time.  Returns the number of bytes written (<0 on error).
int hb_profile_end(char* out, int out_size) {
  using namespace hb;
  g_profiling.store(0);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_marks.empty()) return 0;
  cudaError_t e = cudaEventSynchronize(g_marks.back().ev);
  if (e != cudaSuccess) return cuda_fail(e, "hb_profile_end");
  std::map<std::string, std::pair<long, double>> agg;
  for (size_t i = 1; i < g_marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev);
    auto& a = agg[g_marks[i].what];
    a.first += 1;
    a.second += ms;
  }
  for (auto& m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
  std::vector<std::pair<std::string, std::pair<long, double>>> rows(agg.begin(), agg.end());
  std::sort(rows.begin(), rows.end(), [](auto& a, auto& b) { return a.second.second > b.second.second; });
  int n = 0;
  for (auto& r : rows) {
    int w = snprintf(out + n, out_size > n ? out_size - n : 0, "%s %ld %.6f\n", r.first.c_str(), r.second.first, r.second.second);
    if (w < 0 || n + w >= out_size) break;
    n += w;
  }
  return n;
}
}
