// cfgpp_b200 — C ABI plumbing: export macro, exception -> status translation, last-error storage.
#pragma once
#include <string>

#include "host.h"

#define CFGPP_API __attribute__((visibility("default")))

namespace cfgpp {

std::string& last_error_storage();

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const Error& e) {
    last_error_storage() = e.what();
    return e.code;
  } catch (const std::exception& e) {
    last_error_storage() = e.what();
    return -100;
  } catch (...) {
    last_error_storage() = "unknown C++ exception";
    return -101;
  }
}

}  // namespace cfgpp
