// Collects per-file registration hooks so each .cu can export its own ops
// without editing bindings.cpp.
#include <torch/extension.h>

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
void RegisterAll(pybind11::module& m) {
  for (auto fn : Registry()) fn(m);
}
}  // namespace lb
