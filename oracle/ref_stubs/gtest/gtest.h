// Stand-in for gtest/gtest.h, just enough to run the REFERENCE'S OWN unit tests
// (/root/reference/test) against oracle/_ref's build of the reference: TEST, typed tests,
// EXPECT_EQ / TRUE / FALSE / FLOAT_EQ / DOUBLE_EQ with a streamed message.  FLOAT_EQ and
// DOUBLE_EQ are googletest's: equal within 4 units in the last place, NaN never equal.
// TEST INFRASTRUCTURE (oracle/Makefile, target ref_gtests).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

class Test {
 public:
  virtual ~Test() {}
  virtual void TestBody() = 0;
};

template <class... Ts> struct Types {};

namespace internal {

struct Case {
  std::string name;
  std::function<void()> run;
};
inline std::vector<Case> &cases() {
  static std::vector<Case> v;
  return v;
}
inline int &failures() {
  static int n = 0;
  return n;
}
inline int add(const std::string &name, std::function<void()> run) {
  cases().push_back(Case{name, run});
  return 0;
}

template <template <class> class T, class U> int add_one(const char *fixture, const char *name, int index) {
  return add(std::string(fixture) + "/" + std::to_string(index) + "." + name, [] {
    T<U> t;
    t.TestBody();
  });
}
template <template <class> class T, class... Ts>
int add_typed(const char *fixture, const char *name, Types<Ts...>) {
  int index = 0;
  int dummy[] = {add_one<T, Ts>(fixture, name, index++)...};
  (void)dummy;
  return 0;
}

// what EXPECT_* evaluates to: reports a failure, with whatever was streamed into it, when it dies
class Report {
 public:
  Report(bool ok, const char *what, const char *file, int line) : ok_(ok), what_(what), file_(file), line_(line) {}
  Report(Report &&o) : ok_(o.ok_), what_(o.what_), file_(o.file_), line_(o.line_) {
    msg_ << o.msg_.str();
    o.ok_ = true;
  }
  ~Report() {
    if (!ok_) {
      ++failures();
      std::printf("%s:%d: Failure\n  %s\n  %s\n", file_, line_, what_, msg_.str().c_str());
    }
  }
  template <class V> Report &operator<<(const V &v) {
    msg_ << v;
    return *this;
  }

 private:
  bool ok_;
  const char *what_, *file_;
  int line_;
  std::ostringstream msg_;
};

template <class A, class B> bool equal(const A &a, const B &b) { return a == b; }

// googletest's FloatingPoint<T>::AlmostEquals: sign-and-magnitude -> biased, distance <= 4 ULP
template <class F, class U> bool almost_equal(F a, F b) {
  if (a != a || b != b) return false;
  U ua, ub;
  std::memcpy(&ua, &a, sizeof(F));
  std::memcpy(&ub, &b, sizeof(F));
  const U sign = U(1) << (8 * sizeof(U) - 1);
  const U ba = (ua & sign) ? ~ua + 1 : ua | sign;
  const U bb = (ub & sign) ? ~ub + 1 : ub | sign;
  return (ba >= bb ? ba - bb : bb - ba) <= 4;
}
inline bool float_eq(float a, float b) { return almost_equal<float, uint32_t>(a, b); }
inline bool double_eq(double a, double b) { return almost_equal<double, uint64_t>(a, b); }

}  // namespace internal

inline void InitGoogleTest(int *, char **) {}

}  // namespace testing

inline int RUN_ALL_TESTS() {
  int failed_cases = 0;
  for (const auto &c : ::testing::internal::cases()) {
    const int before = ::testing::internal::failures();
    c.run();
    const bool ok = ::testing::internal::failures() == before;
    std::printf("[ %s ] %s\n", ok ? "      OK" : "  FAILED", c.name.c_str());
    failed_cases += !ok;
  }
  std::printf("[==========] %zu tests ran, %d failed.\n", ::testing::internal::cases().size(), failed_cases);
  return failed_cases ? 1 : 0;
}

#define QMCL_GTEST_REPORT(ok, what) ::testing::internal::Report((ok), what, __FILE__, __LINE__)

#define TEST(suite, name)                                                                     \
  static void suite##_##name##_body();                                                        \
  static int suite##_##name##_registered = ::testing::internal::add(#suite "." #name, suite##_##name##_body); \
  static void suite##_##name##_body()

#define TYPED_TEST_CASE(fixture, types) typedef types gtest_types_##fixture

#define TYPED_TEST(fixture, name)                                                              \
  template <class gtest_T> class fixture##_##name##_Test : public fixture<gtest_T> {           \
   public:                                                                                     \
    typedef fixture<gtest_T> TestFixture;                                                      \
    typedef gtest_T TypeParam;                                                                 \
    void TestBody() override;                                                                  \
  };                                                                                           \
  static int fixture##_##name##_registered =                                                   \
      ::testing::internal::add_typed<fixture##_##name##_Test>(#fixture, #name, gtest_types_##fixture()); \
  template <class gtest_T> void fixture##_##name##_Test<gtest_T>::TestBody()

#define EXPECT_TRUE(c) QMCL_GTEST_REPORT(static_cast<bool>(c), "EXPECT_TRUE(" #c ")")
#define EXPECT_FALSE(c) QMCL_GTEST_REPORT(!static_cast<bool>(c), "EXPECT_FALSE(" #c ")")
#define EXPECT_EQ(a, b) QMCL_GTEST_REPORT(::testing::internal::equal((a), (b)), "EXPECT_EQ(" #a ", " #b ")")
#define EXPECT_FLOAT_EQ(a, b) \
  QMCL_GTEST_REPORT(::testing::internal::float_eq(static_cast<float>(a), static_cast<float>(b)), "EXPECT_FLOAT_EQ(" #a ", " #b ")")
#define EXPECT_DOUBLE_EQ(a, b) \
  QMCL_GTEST_REPORT(::testing::internal::double_eq(static_cast<double>(a), static_cast<double>(b)), "EXPECT_DOUBLE_EQ(" #a ", " #b ")")
