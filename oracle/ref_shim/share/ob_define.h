// Minimal stand-ins for the macros ob_bit_stream.h expects from share/ob_define.h, so that the
// REFERENCE's own encoding/ob_bit_stream.{h,cpp} compile unmodified, from where they lie under
// /root/reference, into oracle/_ref/libref_bitstream.so (checker only; see oracle/Makefile).
// Nothing here restates reference code: it only defines logging / attribute macros as no-ops.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#define OB_INLINE inline __attribute__((always_inline))
#define OB_LIKELY(x) __builtin_expect(!!(x), 1)
#define OB_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define STORAGE_LOG(...) ((void)0)
#define K(x) 0
#define KP(x) 0
#define K_(x) 0
#define KP_(x) 0
#define TO_STRING_KV(...)
#define DISALLOW_COPY_AND_ASSIGN(T) T(const T &) = delete; T &operator=(const T &) = delete
#define RLOCAL_INLINE(TYPE, NAME) static thread_local TYPE NAME
#define MEMSET memset
#define MEMCPY memcpy
#define UNUSED(x) ((void)(x))
namespace oceanbase { namespace common {
constexpr int OB_SUCCESS = 0;
constexpr int OB_INVALID_ARGUMENT = -4002;
constexpr int OB_NOT_INIT = -4006;
constexpr int OB_INDEX_OUT_OF_RANGE = -4003 - 9;  /* value irrelevant to the checker */
} }
