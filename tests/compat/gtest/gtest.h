/* Minimal stand-in for GoogleTest (which cannot be fetched here, cmake/setup_GTest.cmake needs the network): just what
 * the reference's test/unit_cuda sources use - TEST, EXPECT_EQ / NE / TRUE / FALSE / NEAR / LT / LE / GT / GE, ASSERT_*,
 * InitGoogleTest, RUN_ALL_TESTS.  TEST INFRASTRUCTURE ONLY. */
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace testing
{

struct Registry
{
    struct Case
    {
        std::string name;
        std::function<void()> body;
    };
    std::vector<Case> cases;
    int failures{0}; // failed expectations of the test that is running
    static Registry& get()
    {
        static Registry r;
        return r;
    }
};

struct Registrar
{
    Registrar(const char* suite, const char* name, std::function<void()> body)
    {
        Registry::get().cases.push_back({std::string(suite) + "." + name, std::move(body)});
    }
};

inline void InitGoogleTest(int*, char**) {}

inline void fail(const char* file, int line, const char* what)
{
    std::printf("%s:%d: Failure: %s\n", file, line, what);
    ++Registry::get().failures;
}

} // namespace testing

inline int RUN_ALL_TESTS()
{
    auto& r    = testing::Registry::get();
    int failed = 0;
    for (auto& c : r.cases)
    {
        r.failures = 0;
        std::printf("[ RUN      ] %s\n", c.name.c_str());
        try
        {
            c.body();
        }
        catch (std::exception& e)
        {
            std::printf("exception: %s\n", e.what());
            ++r.failures;
        }
        std::printf(r.failures ? "[  FAILED  ] %s\n" : "[       OK ] %s\n", c.name.c_str());
        failed += r.failures != 0;
    }
    std::printf("[==========] %zu tests ran, %d failed\n", r.cases.size(), failed);
    return failed ? 1 : 0;
}

#define TEST(suite, name)                                                                                              \
    static void suite##_##name##_body();                                                                               \
    static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);                            \
    static void suite##_##name##_body()

#define CS_GT_CHECK(cond, text, fatal)                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            ::testing::fail(__FILE__, __LINE__, text);                                                                 \
            if (fatal) { return; }                                                                                     \
        }                                                                                                              \
    } while (0)

#define EXPECT_EQ(a, b) CS_GT_CHECK((a) == (b), #a " == " #b, false)
#define EXPECT_NE(a, b) CS_GT_CHECK(!((a) == (b)), #a " != " #b, false)
#define EXPECT_LT(a, b) CS_GT_CHECK((a) < (b), #a " < " #b, false)
#define EXPECT_LE(a, b) CS_GT_CHECK((a) <= (b), #a " <= " #b, false)
#define EXPECT_GT(a, b) CS_GT_CHECK((a) > (b), #a " > " #b, false)
#define EXPECT_GE(a, b) CS_GT_CHECK((a) >= (b), #a " >= " #b, false)
#define EXPECT_TRUE(a) CS_GT_CHECK((a), #a, false)
#define EXPECT_FALSE(a) CS_GT_CHECK(!(a), "!" #a, false)
#define EXPECT_NEAR(a, b, tol) CS_GT_CHECK(std::abs((a) - (b)) <= (tol), #a " ~ " #b, false)
#define ASSERT_EQ(a, b) CS_GT_CHECK((a) == (b), #a " == " #b, true)
#define ASSERT_TRUE(a) CS_GT_CHECK((a), #a, true)
#define ASSERT_FALSE(a) CS_GT_CHECK(!(a), "!" #a, true)
