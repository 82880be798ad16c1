// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// No-op stand-in for pybind11 so that the reference's cfar.cpp
// (bruce_slam/src/bruce_slam/cpp/cfar.cpp:1-3,194-204) compiles unmodified into
// a plain shared object: the PYBIND11_MODULE block becomes an unused static
// function.  The functions themselves (ca, soca, goca, os, *2) keep external
// linkage and are called through oracle/cfar_refshim.cpp's extern "C" wrappers.
#pragma once
namespace pybind11 {
struct module_ {
  template <class F>
  module_ &def(const char *, F) { return *this; }
};
}  // namespace pybind11
#define PYBIND11_MODULE(name, var) \
  static void __attribute__((unused)) sfe_shim_module_##name(pybind11::module_ &var)
