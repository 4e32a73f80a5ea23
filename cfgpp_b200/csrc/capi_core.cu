// cfgpp_b200 — C ABI core: version, last error.
#include "capi_util.h"

namespace cfgpp {
std::string& last_error_storage() {
  static thread_local std::string s;
  return s;
}
}  // namespace cfgpp

extern "C" {

CFGPP_API int cfgpp_version(void) { return 100; }

CFGPP_API const char* cfgpp_last_error(void) { return cfgpp::last_error_storage().c_str(); }

}  // extern "C"
