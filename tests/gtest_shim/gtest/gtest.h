// gtest/gtest.h -- a minimal GoogleTest-compatible shim (GoogleTest is not installed in this image and there is no
// network).  It provides exactly what the reference's test sources use (test/src/*.cpp of bassoy/ttv): TEST, TEST_F,
// ::testing::Test with SetUp/TearDown, EXPECT_/ASSERT_{TRUE,FALSE,EQ,NE,LT,LE,GT,GE,FLOAT_EQ,DOUBLE_EQ,NEAR},
// ::testing::InitGoogleTest and RUN_ALL_TESTS.  TEST INFRASTRUCTURE ONLY: it lets tests/ref_gtests.py compile the
// reference's own, unmodified test sources against THIS repo's include/tlib headers.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace testing {

class Test {
public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};

namespace internal {

struct Registry {
  struct Entry { std::string suite, name; std::function<Test*()> make; };
  std::vector<Entry> tests;
  int failures_in_current = 0;
  long checks = 0;
  bool fatal = false;
  std::string filter;
  static Registry& get() { static Registry r; return r; }
};

struct Registrar {
  Registrar(const char* suite, const char* name, std::function<Test*()> make) { Registry::get().tests.push_back({suite, name, std::move(make)}); }
};

// message sink so that `EXPECT_x(...) << "text"` compiles
struct Message {
  bool failed;
  std::ostringstream os;
  explicit Message(bool f) : failed(f) {}
  Message(Message&& o) : failed(o.failed), os(std::move(o.os)) {}
  template<class T> Message& operator<<(T const& v) { if (failed) os << v; return *this; }
  ~Message() { if (failed) { auto s = os.str(); if (!s.empty()) std::cerr << "    " << s << "\n"; } }
};

inline Message report(bool ok, const char* file, int line, const char* what, std::string const& detail = std::string())
{
  auto& r = Registry::get();
  ++r.checks;
  if (!ok) {
    if (r.failures_in_current < 10)
      std::cerr << file << ":" << line << ": Failure: " << what << (detail.empty() ? "" : "  [" + detail + "]") << "\n";
    ++r.failures_in_current;
  }
  return Message(!ok);
}

template<class A, class B>
std::string describe(A const& a, B const& b)
{
  std::ostringstream os;
  os << a << " vs " << b;
  return os.str();
}

// 4-ULP comparison like GoogleTest's EXPECT_FLOAT_EQ / EXPECT_DOUBLE_EQ
template<class F, class I>
bool almost_equal_ulps(F a, F b)
{
  if (std::isnan(a) || std::isnan(b)) return false;
  if (a == b) return true;
  I ia, ib;
  std::memcpy(&ia, &a, sizeof a);
  std::memcpy(&ib, &b, sizeof b);
  auto biased = [](I v) { const I sign = I(1) << (sizeof(I) * 8 - 1); return (v & sign) ? ~v + 1 : (v | sign); };
  const I x = biased(ia), y = biased(ib);
  return (x > y ? x - y : y - x) <= 4;
}

} // namespace internal

inline void InitGoogleTest(int* argc, char** argv)
{
  for (int i = 1; argc && i < *argc; ++i) {
    std::string a = argv[i];
    const std::string key = "--gtest_filter=";
    if (a.compare(0, key.size(), key) == 0) internal::Registry::get().filter = a.substr(key.size());
  }
}

} // namespace testing

inline int RUN_ALL_TESTS()
{
  auto& r = ::testing::internal::Registry::get();
  int failed = 0, ran = 0;
  for (auto& t : r.tests) {
    const std::string full = t.suite + "." + t.name;
    if (!r.filter.empty() && r.filter != "*" && full.find(r.filter) == std::string::npos) continue;
    r.failures_in_current = 0;
    r.fatal = false;
    std::cout << "[ RUN      ] " << full << std::endl;
    ::testing::Test* obj = t.make();
    try {
      obj->SetUp();
      obj->TestBody();
      obj->TearDown();
    } catch (std::exception const& e) {
      std::cerr << "  uncaught exception: " << e.what() << "\n";
      ++r.failures_in_current;
    }
    delete obj;
    ++ran;
    if (r.failures_in_current) { ++failed; std::cout << "[  FAILED  ] " << full << " (" << r.failures_in_current << " failed checks)" << std::endl; }
    else std::cout << "[       OK ] " << full << std::endl;
  }
  std::cout << "[==========] " << ran << " tests ran, " << r.checks << " checks." << std::endl;
  std::cout << (failed ? "[  FAILED  ] " : "[  PASSED  ] ") << (failed ? failed : ran) << " tests." << std::endl;
  return failed ? 1 : 0;
}

#define GTEST_SHIM_CLASS_(suite, name) suite##_##name##_Test

#define GTEST_SHIM_TEST_(suite, name, parent)                                                                       \
  class GTEST_SHIM_CLASS_(suite, name) : public parent {                                                            \
  public:                                                                                                           \
    void TestBody() override;                                                                                       \
  };                                                                                                                \
  static ::testing::internal::Registrar gtest_shim_registrar_##suite##_##name(                                      \
      #suite, #name, []() -> ::testing::Test* { return new GTEST_SHIM_CLASS_(suite, name)(); });                    \
  void GTEST_SHIM_CLASS_(suite, name)::TestBody()

#define TEST(suite, name)      GTEST_SHIM_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name)  GTEST_SHIM_TEST_(fixture, name, fixture)

#define GTEST_SHIM_CHECK_(cond, text, detail) ::testing::internal::report((cond), __FILE__, __LINE__, text, detail)
#define GTEST_SHIM_FATAL_(cond, text, detail)                                                                       \
  if (bool gtest_shim_ok = (cond); gtest_shim_ok) ::testing::internal::report(true, __FILE__, __LINE__, text);      \
  else return (void)::testing::internal::report(false, __FILE__, __LINE__, text, detail)

#define EXPECT_TRUE(c)   GTEST_SHIM_CHECK_(static_cast<bool>(c), "EXPECT_TRUE(" #c ")", "")
#define EXPECT_FALSE(c)  GTEST_SHIM_CHECK_(!static_cast<bool>(c), "EXPECT_FALSE(" #c ")", "")
#define EXPECT_EQ(a, b)  GTEST_SHIM_CHECK_((a) == (b), "EXPECT_EQ(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_NE(a, b)  GTEST_SHIM_CHECK_((a) != (b), "EXPECT_NE(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_LT(a, b)  GTEST_SHIM_CHECK_((a) <  (b), "EXPECT_LT(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_LE(a, b)  GTEST_SHIM_CHECK_((a) <= (b), "EXPECT_LE(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_GT(a, b)  GTEST_SHIM_CHECK_((a) >  (b), "EXPECT_GT(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_GE(a, b)  GTEST_SHIM_CHECK_((a) >= (b), "EXPECT_GE(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_FLOAT_EQ(a, b)  GTEST_SHIM_CHECK_((::testing::internal::almost_equal_ulps<float, std::uint32_t>(static_cast<float>(a), static_cast<float>(b))), "EXPECT_FLOAT_EQ(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_DOUBLE_EQ(a, b) GTEST_SHIM_CHECK_((::testing::internal::almost_equal_ulps<double, std::uint64_t>(static_cast<double>(a), static_cast<double>(b))), "EXPECT_DOUBLE_EQ(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define EXPECT_NEAR(a, b, tol) GTEST_SHIM_CHECK_(std::fabs(static_cast<double>(a) - static_cast<double>(b)) <= static_cast<double>(tol), "EXPECT_NEAR(" #a ", " #b ")", ::testing::internal::describe((a), (b)))

#define ASSERT_TRUE(c)   GTEST_SHIM_FATAL_(static_cast<bool>(c), "ASSERT_TRUE(" #c ")", "")
#define ASSERT_FALSE(c)  GTEST_SHIM_FATAL_(!static_cast<bool>(c), "ASSERT_FALSE(" #c ")", "")
#define ASSERT_EQ(a, b)  GTEST_SHIM_FATAL_((a) == (b), "ASSERT_EQ(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
#define ASSERT_NE(a, b)  GTEST_SHIM_FATAL_((a) != (b), "ASSERT_NE(" #a ", " #b ")", ::testing::internal::describe((a), (b)))
