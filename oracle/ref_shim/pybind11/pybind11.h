// Stand-in for <pybind11/pybind11.h> when the reference's pybind_*.cpp files are compiled into
// oracle/_ref without Python: PYBIND11_MODULE bodies compile to unused functions.
#ifndef ORACLE_PYBIND_STUB_H_
#define ORACLE_PYBIND_STUB_H_
namespace pybind11 {
struct arg {
  const char* name;
  explicit arg(const char* n) : name(n) {}
  template <typename T>
  arg& operator=(const T&) { return *this; }
};
struct module_ {
  template <typename... A>
  module_& def(A&&...) { return *this; }
  template <typename... A>
  module_& attr(A&&...) { return *this; }
};
}  // namespace pybind11
#define PYBIND11_MODULE(name, var) static void oracle_unused_module_##name(pybind11::module_& var)
#endif
