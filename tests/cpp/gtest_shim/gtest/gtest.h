// gtest.h -- a minimal stand-in for GoogleTest 1.7 (not installed in this image), just enough to compile and run the
// reference's own test headers (test/t*.h) UNMODIFIED: TEST, EXPECT_/ASSERT_ {TRUE, FALSE, EQ, NE, GT, LT, NEAR, FLOAT_EQ,
// STREQ}, a trailing `<< message`, and a runner that prints one line per test.  Test infrastructure, not product code.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
struct TestInfo { const char* suite; const char* name; void (*fn)(); };
inline std::vector<TestInfo>& registry() { static std::vector<TestInfo> r; return r; }
struct Registrar { Registrar(const char* s, const char* n, void (*f)()) { registry().push_back({s, n, f}); } };
inline int& failures_in_test() { static int n = 0; return n; }
struct Message {                                       // swallows `<< anything` after an assertion
    bool failed; std::ostringstream os;
    explicit Message(bool f) : failed(f) {}
    Message(Message&& o) : failed(o.failed) { os << o.os.str(); }
    ~Message() { if (failed && !os.str().empty()) std::printf("    %s\n", os.str().c_str()); }
    template <typename T> Message& operator<<(const T& v) { if (failed) os << v; return *this; }
};
struct AbortTest {};
template <typename A, typename B> std::string show(const A& a, const B& b) { std::ostringstream o; o.precision(17); o << a << " vs " << b; return o.str(); }
inline Message report(bool ok, bool fatal, const char* file, int line, const char* expr, const std::string& detail)
{
    if (!ok) {
        failures_in_test()++;
        std::printf("  %s:%d: %s  (%s)\n", file, line, expr, detail.c_str());
        if (fatal) throw AbortTest();
    }
    return Message(!ok);
}
inline bool float_eq(float a, float b)                 // gtest's AlmostEquals: within 4 units in the last place
{
    if (std::isnan(a) || std::isnan(b)) return false;
    auto biased = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return (u & 0x80000000u) ? ~u + 1 : u | 0x80000000u; };
    const uint32_t x = biased(a), y = biased(b);
    return (x > y ? x - y : y - x) <= 4;
}
inline void InitGoogleTest(int*, char**) {}
inline int run_all(const char* filter = nullptr)
{
    int failed = 0, ran = 0;
    for (const TestInfo& t : registry()) {
        const std::string full = std::string(t.suite) + "." + t.name;
        if (filter && full.find(filter) == std::string::npos) continue;
        failures_in_test() = 0; ran++;
        try { t.fn(); } catch (const AbortTest&) {}
        std::printf("[%s] %s\n", failures_in_test() ? "FAILED" : "  OK  ", full.c_str());
        std::fflush(stdout);
        if (failures_in_test()) failed++;
    }
    std::printf("%d tests, %d failures\n", ran, failed);
    return failed;
}
} // namespace testing

#define TEST(suite, name) \
    static void suite##_##name##_body(); \
    static ::testing::Registrar suite##_##name##_reg(#suite, #name, &suite##_##name##_body); \
    static void suite##_##name##_body()
#define VXT_CHECK(ok, fatal, expr, detail) ::testing::report((ok), fatal, __FILE__, __LINE__, expr, detail)
#define VXT_BIN(a, b, op, fatal, name) VXT_CHECK(((a) op (b)), fatal, name "(" #a ", " #b ")", ::testing::show((a), (b)))
#define EXPECT_TRUE(c) VXT_CHECK(!!(c), false, "EXPECT_TRUE(" #c ")", "false")
#define EXPECT_FALSE(c) VXT_CHECK(!(c), false, "EXPECT_FALSE(" #c ")", "true")
#define ASSERT_TRUE(c) VXT_CHECK(!!(c), true, "ASSERT_TRUE(" #c ")", "false")
#define ASSERT_FALSE(c) VXT_CHECK(!(c), true, "ASSERT_FALSE(" #c ")", "true")
#define EXPECT_EQ(a, b) VXT_BIN(a, b, ==, false, "EXPECT_EQ")
#define EXPECT_NE(a, b) VXT_BIN(a, b, !=, false, "EXPECT_NE")
#define EXPECT_GT(a, b) VXT_BIN(a, b, >, false, "EXPECT_GT")
#define EXPECT_LT(a, b) VXT_BIN(a, b, <, false, "EXPECT_LT")
#define EXPECT_NEAR(a, b, tol) VXT_CHECK(std::fabs((double)(a) - (double)(b)) <= (double)(tol), false, "EXPECT_NEAR(" #a ", " #b ", " #tol ")", ::testing::show((a), (b)))
#define ASSERT_NEAR(a, b, tol) VXT_CHECK(std::fabs((double)(a) - (double)(b)) <= (double)(tol), true, "ASSERT_NEAR(" #a ", " #b ", " #tol ")", ::testing::show((a), (b)))
#define EXPECT_FLOAT_EQ(a, b) VXT_CHECK(::testing::float_eq((float)(a), (float)(b)), false, "EXPECT_FLOAT_EQ(" #a ", " #b ")", ::testing::show((a), (b)))
#define ASSERT_FLOAT_EQ(a, b) VXT_CHECK(::testing::float_eq((float)(a), (float)(b)), true, "ASSERT_FLOAT_EQ(" #a ", " #b ")", ::testing::show((a), (b)))
#define EXPECT_STREQ(a, b) VXT_CHECK(std::strcmp((a), (b)) == 0, false, "EXPECT_STREQ(" #a ", " #b ")", ::testing::show((a), (b)))
#define RUN_ALL_TESTS() ::testing::run_all()
