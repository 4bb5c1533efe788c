#pragma once
#include <torch/extension.h>
namespace lb {
using RegFn = void (*)(pybind11::module&);
int AddRegistration(RegFn fn);
void CountLaunch(int n = 1);
}  // namespace lb
#define LB_REGISTER(name)                                         \
  static void lb_reg_##name(pybind11::module& m);                 \
  static int lb_reg_token_##name = ::lb::AddRegistration(&lb_reg_##name); \
  static void lb_reg_##name(pybind11::module& m)
