// libepseon.hpp -- debug/release switches and assertion macros.
// Mirrors the only header of the reference's cpp/include (cpp/include/epseon/libepseon.hpp:10-44):
// same macro names and meaning, without the fmt dependency.
#pragma once
#include <cassert>

#if defined(DEBUG) || defined(_DEBUG) || !defined(NDEBUG)
    #define LIB_EPSEON_DEBUG 1
    #define LIB_EPSEON_RELEASE 0
#else
    #define LIB_EPSEON_DEBUG 0
    #define LIB_EPSEON_RELEASE 1
#endif

static_assert(LIB_EPSEON_DEBUG + LIB_EPSEON_RELEASE == 1, "exactly one of DEBUG / RELEASE must be set");

#if LIB_EPSEON_DEBUG
    #define LIB_EPSEON_ASSERT_TRUE(EXPRESSION) assert(EXPRESSION)
    #define LIB_EPSEON_ASSERT_FALSE(EXPRESSION) assert(!(static_cast<bool>(EXPRESSION)))
#else
    #define LIB_EPSEON_ASSERT_TRUE(EXPRESSION)
    #define LIB_EPSEON_ASSERT_FALSE(EXPRESSION)
#endif
