// Minimal stand-in for <gtest/gtest.h> (googletest is not installable offline): just
// enough to compile the reference's own test/gtest/device/spmv_test.cpp UNMODIFIED against
// the B200 backend.  TEST registers a function; EXPECT_NEAR / EXPECT_EQ count failures;
// main() runs everything and returns non-zero on any failure.
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace mini_gtest {
struct registry {
  static std::vector<std::pair<std::string, std::function<void()>>>& tests() {
    static std::vector<std::pair<std::string, std::function<void()>>> t;
    return t;
  }
  static int& failures() {
    static int f = 0;
    return f;
  }
  static int& checks() {
    static int c = 0;
    return c;
  }
};
struct registrar {
  registrar(const char* name, std::function<void()> fn) {
    registry::tests().emplace_back(name, std::move(fn));
  }
};
inline void report(const char* file, int line, const char* what) {
  if (++registry::failures() <= 20)
    std::printf("%s:%d: failure: %s\n", file, line, what);
}
} // namespace mini_gtest

#define TEST(suite, name)                                                            \
  static void suite##_##name##_body();                                               \
  static mini_gtest::registrar suite##_##name##_reg(#suite "." #name,                \
                                                    suite##_##name##_body);          \
  static void suite##_##name##_body()

#define EXPECT_NEAR(a, b, tol)                                                       \
  do {                                                                               \
    ++mini_gtest::registry::checks();                                                \
    if (!(std::abs((a) - (b)) <= (tol)))                                             \
      mini_gtest::report(__FILE__, __LINE__, "EXPECT_NEAR(" #a ", " #b ")");         \
  } while (0)

#define EXPECT_EQ(a, b)                                                              \
  do {                                                                               \
    ++mini_gtest::registry::checks();                                                \
    if (!((a) == (b)))                                                               \
      mini_gtest::report(__FILE__, __LINE__, "EXPECT_EQ(" #a ", " #b ")");           \
  } while (0)

#ifndef MINI_GTEST_NO_MAIN
int main() {
  for (auto& [name, fn] : mini_gtest::registry::tests()) {
    const int before = mini_gtest::registry::failures();
    fn();
    std::printf("[%s] %s\n", mini_gtest::registry::failures() == before ? "  OK  " : "FAILED",
                name.c_str());
  }
  std::printf("%d checks, %d failures\n", mini_gtest::registry::checks(),
              mini_gtest::registry::failures());
  return mini_gtest::registry::failures() == 0 ? 0 : 1;
}
#endif
