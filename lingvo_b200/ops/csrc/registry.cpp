// Collects per-file registration hooks so each .cu can export its own ops
// without editing bindings.cpp.
#include <torch/extension.h>

#include <atomic>
#include <vector>

namespace lb {
using RegFn = void (*)(pybind11::module&);
std::vector<RegFn>& Registry() {
  static std::vector<RegFn> r;
  return r;
}
int AddRegistration(RegFn fn) {
  Registry().push_back(fn);
  return 0;
}
std::atomic<long long> g_launch_count{0};
void CountLaunch(int n) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
void RegisterAll(pybind11::module& m) {
  m.def("launch_count", []() { return g_launch_count.load(); },
        "Number of lingvo_b200 CUDA kernels launched by this process.");
  m.def("reset_launch_count", []() { g_launch_count.store(0); });
  for (auto fn : Registry()) fn(m);
}
}  // namespace lb
