// oracle/shim/gtest/gtest.h — a ~150-line stand-in for the googletest subset the reference's
// *_test.cc files use (TEST, EXPECT_*/ASSERT_* with streamed messages, FAIL, EXPECT_THROW).
// TEST INFRASTRUCTURE ONLY: googletest 1.17.0 (subprojects/gtest.wrap) is not available here.
// It lets the reference's own known-answer tests run UNMODIFIED against the reference sources
// built with the Eigen/absl shims, which is how the shims (and therefore oracle/_ref) are pinned.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <optional>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace testing {

struct TestInfo {
  const char* suite;
  const char* name;
  std::function<void()> fn;
};
inline std::vector<TestInfo>& registry() {
  static std::vector<TestInfo> r;
  return r;
}
inline int& failures_in_current() {
  static int f = 0;
  return f;
}
struct Registrar {
  Registrar(const char* s, const char* n, std::function<void()> f) { registry().push_back({s, n, std::move(f)}); }
};

// Collects a streamed message; reports on destruction if the check failed.
class Reporter {
 public:
  Reporter(bool failed, const char* file, int line, std::string what)
      : failed_(failed), file_(file), line_(line), what_(std::move(what)) {}
  Reporter(const Reporter&) = delete;
  ~Reporter() {
    if (failed_) {
      ++failures_in_current();
      std::cerr << file_ << ":" << line_ << ": Failure: " << what_ << " " << ss_.str() << std::endl;
    }
  }
  template <typename T>
  Reporter& operator<<(const T& v) {
    if (failed_) ss_ << v;
    return *this;
  }
  Reporter& operator<<(std::ostream& (*m)(std::ostream&)) {
    if (failed_) ss_ << m;
    return *this;
  }
  bool failed() const { return failed_; }

 private:
  bool failed_;
  const char* file_;
  int line_;
  std::string what_;
  std::ostringstream ss_;
};

// `return Voidify() = Reporter(...) << msg;` lets ASSERT_* return from a void test body.
struct Voidify {
  void operator=(const Reporter&) const {}
};

// googletest's 4-ULP float comparison.
inline bool almost_equal_float(float a, float b) {
  if (std::isnan(a) || std::isnan(b)) return false;
  auto biased = [](float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u + 1 : u | 0x80000000u;
  };
  uint32_t x = biased(a), y = biased(b);
  return (x > y ? x - y : y - x) <= 4;
}

template <typename A, typename B>
bool eq(const A& a, const B& b) {
  if constexpr (std::is_arithmetic_v<A> && std::is_arithmetic_v<B>) {
    // mimic usual arithmetic conversions without sign-compare warnings
    using CT = std::common_type_t<A, B>;
    return static_cast<CT>(a) == static_cast<CT>(b);
  } else {
    return a == b;
  }
}

}  // namespace testing

#define SHIM_TEST_NAME_(s, n) s##_##n##_Test
#define TEST(s, n)                                                        \
  static void SHIM_TEST_NAME_(s, n)();                                     \
  static ::testing::Registrar s##_##n##_registrar(#s, #n, SHIM_TEST_NAME_(s, n)); \
  static void SHIM_TEST_NAME_(s, n)()

#define SHIM_CHECK_(cond, text) ::testing::Reporter(!(cond), __FILE__, __LINE__, text)
// ASSERT_*: `if (ok) ; else return Voidify() = Reporter(...) << "msg";`
#define SHIM_ASSERT_(cond, text) \
  if (cond)                      \
    ;                            \
  else                           \
    return ::testing::Voidify() = ::testing::Reporter(true, __FILE__, __LINE__, text)

#define EXPECT_TRUE(c) SHIM_CHECK_(static_cast<bool>(c), "EXPECT_TRUE(" #c ")")
#define EXPECT_FALSE(c) SHIM_CHECK_(!static_cast<bool>(c), "EXPECT_FALSE(" #c ")")
#define EXPECT_EQ(a, b) SHIM_CHECK_(::testing::eq((a), (b)), "EXPECT_EQ(" #a ", " #b ")")
#define EXPECT_NE(a, b) SHIM_CHECK_(!::testing::eq((a), (b)), "EXPECT_NE(" #a ", " #b ")")
#define EXPECT_LT(a, b) SHIM_CHECK_((a) < (b), "EXPECT_LT(" #a ", " #b ")")
#define EXPECT_LE(a, b) SHIM_CHECK_((a) <= (b), "EXPECT_LE(" #a ", " #b ")")
#define EXPECT_GT(a, b) SHIM_CHECK_((a) > (b), "EXPECT_GT(" #a ", " #b ")")
#define EXPECT_GE(a, b) SHIM_CHECK_((a) >= (b), "EXPECT_GE(" #a ", " #b ")")
#define EXPECT_FLOAT_EQ(a, b) \
  SHIM_CHECK_(::testing::almost_equal_float(static_cast<float>(a), static_cast<float>(b)), "EXPECT_FLOAT_EQ(" #a ", " #b ")")
#define EXPECT_NEAR(a, b, tol) \
  SHIM_CHECK_(std::fabs(static_cast<double>(a) - static_cast<double>(b)) <= static_cast<double>(tol), "EXPECT_NEAR(" #a ", " #b ", " #tol ")")
#define EXPECT_STREQ(a, b) SHIM_CHECK_(std::strcmp((a), (b)) == 0, "EXPECT_STREQ(" #a ", " #b ")")

#define ASSERT_TRUE(c) SHIM_ASSERT_(static_cast<bool>(c), "ASSERT_TRUE(" #c ")")
#define ASSERT_FALSE(c) SHIM_ASSERT_(!static_cast<bool>(c), "ASSERT_FALSE(" #c ")")
#define ASSERT_EQ(a, b) SHIM_ASSERT_(::testing::eq((a), (b)), "ASSERT_EQ(" #a ", " #b ")")
#define ASSERT_NE(a, b) SHIM_ASSERT_(!::testing::eq((a), (b)), "ASSERT_NE(" #a ", " #b ")")
#define ASSERT_GE(a, b) SHIM_ASSERT_((a) >= (b), "ASSERT_GE(" #a ", " #b ")")
#define ASSERT_GT(a, b) SHIM_ASSERT_((a) > (b), "ASSERT_GT(" #a ", " #b ")")
#define ASSERT_LE(a, b) SHIM_ASSERT_((a) <= (b), "ASSERT_LE(" #a ", " #b ")")
#define ASSERT_LT(a, b) SHIM_ASSERT_((a) < (b), "ASSERT_LT(" #a ", " #b ")")

#define FAIL() return ::testing::Voidify() = ::testing::Reporter(true, __FILE__, __LINE__, "FAIL()")
#define ADD_FAILURE() ::testing::Reporter(true, __FILE__, __LINE__, "ADD_FAILURE()")
#define SUCCEED() ::testing::Reporter(false, __FILE__, __LINE__, "")

#define EXPECT_THROW(stmt, ex)                                                          \
  do {                                                                                  \
    bool shim_caught = false;                                                           \
    try {                                                                               \
      stmt;                                                                             \
    } catch (const ex&) {                                                               \
      shim_caught = true;                                                               \
    } catch (...) {                                                                     \
    }                                                                                   \
    SHIM_CHECK_(shim_caught, "EXPECT_THROW(" #stmt ", " #ex ")");                       \
  } while (0)
#define EXPECT_NO_THROW(stmt)                                      \
  do {                                                             \
    bool shim_threw = false;                                       \
    try {                                                          \
      stmt;                                                        \
    } catch (...) {                                                \
      shim_threw = true;                                           \
    }                                                              \
    SHIM_CHECK_(!shim_threw, "EXPECT_NO_THROW(" #stmt ")");        \
  } while (0)

inline int RUN_ALL_TESTS(const char* filter = nullptr) {
  int failed = 0, ran = 0;
  for (auto& t : ::testing::registry()) {
    std::string full = std::string(t.suite) + "." + t.name;
    if (filter && full.find(filter) == std::string::npos) continue;
    ::testing::failures_in_current() = 0;
    try {
      t.fn();
    } catch (const std::exception& e) {
      ++::testing::failures_in_current();
      std::cerr << "uncaught exception in " << full << ": " << e.what() << std::endl;
    }
    ++ran;
    bool ok = ::testing::failures_in_current() == 0;
    std::cout << (ok ? "[  OK  ] " : "[ FAIL ] ") << full << std::endl;
    if (!ok) ++failed;
  }
  std::cout << "ran " << ran << " tests, " << failed << " failed" << std::endl;
  return failed == 0 ? 0 : 1;
}

#ifdef SHIM_GTEST_MAIN
int main(int argc, char** argv) { return RUN_ALL_TESTS(argc > 1 ? argv[1] : nullptr); }
#endif
