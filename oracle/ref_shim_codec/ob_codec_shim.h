// Minimal stand-ins for the logging / attribute / error-code macros the REFERENCE's integer codec library
// (deps/oblib/src/lib/codec/*.h, *.cpp) expects from the rest of oblib, so that those files compile unmodified,
// from where they lie under /root/reference, into oracle/_ref/libref_codec.so (checker only; oracle/Makefile).
// Nothing here restates reference code: logging becomes a no-op, error codes keep the reference's values
// (deps/oblib/src/lib/ob_errno.h), the allocator interface is the two virtuals the codecs call.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <limits.h>
#include <algorithm>
#include <type_traits>
#define OB_INLINE inline __attribute__((always_inline))
#define OB_NOINLINE __attribute__((noinline))
#define OB_LIKELY(x) __builtin_expect(!!(x), 1)
#define OB_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define OB_ISNULL(p) (OB_UNLIKELY(nullptr == (p)))
#define OB_SUCC(x) (OB_LIKELY(::oceanbase::common::OB_SUCCESS == (ret = (x))))
#define OB_FAIL(x) (OB_UNLIKELY(::oceanbase::common::OB_SUCCESS != (ret = (x))))
#define FAILEDx(x) (OB_SUCC(ret) && OB_FAIL(x))
#define OB_TMP_FAIL(x) (OB_UNLIKELY(::oceanbase::common::OB_SUCCESS != (tmp_ret = (x))))
#define LIB_LOG(...) ((void)0)
#define STORAGE_LOG(...) ((void)0)
#define COMMON_LOG(...) ((void)0)
#define OB_LOG(...) ((void)0)
#define K(x) 0
#define KP(x) 0
#define K_(x) 0
#define KP_(x) 0
#define KR(x) 0
#define KPC(x) 0
#define KPC_(x) 0
#define TO_STRING_KV(...)
#define VIRTUAL_TO_STRING_KV(...)
#define INHERIT_TO_STRING_KV(...)
#define DISALLOW_COPY_AND_ASSIGN(T) T(const T &) = delete; T &operator=(const T &) = delete
#define MEMSET memset
#define MEMCPY memcpy
#define MEMCMP memcmp
#define UNUSED(x) ((void)(x))
#define UNUSEDx(...)
#define OB_ASSERT(x) ((void)0)
#define IS_INIT (is_inited_)
#define IS_NOT_INIT (!is_inited_)
#define KPHEX_(x, y) 0
#define KPHEX(x, y) 0
#define OB_NOT_NULL(p) (OB_LIKELY(nullptr != (p)))
#define ob_abort() abort()
#ifndef CACHE_ALIGN_SIZE
#define CACHE_ALIGN_SIZE 64
#endif
#define CACHE_ALIGNED __attribute__((aligned(CACHE_ALIGN_SIZE)))
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
// clang builtins the reference (a clang-only build) uses; g++ has no equivalents
static inline uint8_t ob_shim_bitrev8(uint8_t v) { v = (uint8_t)((v >> 4) | (v << 4)); v = (uint8_t)(((v & 0xcc) >> 2) | ((v & 0x33) << 2)); return (uint8_t)(((v & 0xaa) >> 1) | ((v & 0x55) << 1)); }
static inline uint16_t ob_shim_bitrev16(uint16_t v) { return (uint16_t)((ob_shim_bitrev8((uint8_t)v) << 8) | ob_shim_bitrev8((uint8_t)(v >> 8))); }
static inline uint32_t ob_shim_bitrev32(uint32_t v) { return ((uint32_t)ob_shim_bitrev16((uint16_t)v) << 16) | ob_shim_bitrev16((uint16_t)(v >> 16)); }
static inline uint64_t ob_shim_bitrev64(uint64_t v) { return ((uint64_t)ob_shim_bitrev32((uint32_t)v) << 32) | ob_shim_bitrev32((uint32_t)(v >> 32)); }
#define __builtin_bitreverse8(x) ob_shim_bitrev8(x)
#define __builtin_bitreverse16(x) ob_shim_bitrev16(x)
#define __builtin_bitreverse32(x) ob_shim_bitrev32(x)
#define __builtin_bitreverse64(x) ob_shim_bitrev64(x)
namespace oceanbase {
namespace common {
constexpr int OB_SUCCESS = 0;
constexpr int OB_INVALID_ARGUMENT = -4002;
constexpr int OB_ERROR_OUT_OF_RANGE = -4175;
constexpr int OB_INIT_TWICE = -4005;
constexpr int OB_ERROR = -4000;
constexpr int OB_ITER_END = -4008;
constexpr int OB_NOT_INIT = -4006;
constexpr int OB_NOT_SUPPORTED = -4007;
constexpr int OB_ALLOCATE_MEMORY_FAILED = -4013;
constexpr int OB_INNER_STAT_ERROR = -4014;
constexpr int OB_ERR_UNEXPECTED = -4016;
constexpr int OB_SIZE_OVERFLOW = -4019;
constexpr int OB_BUF_NOT_ENOUGH = -4024;
constexpr int OB_INVALID_DATA = -4070;
class ObIAllocator {
public:
  virtual ~ObIAllocator() {}
  virtual void *alloc(const int64_t size) = 0;
  virtual void free(void *ptr) = 0;
};
enum ObCompressorType : uint8_t { INVALID_COMPRESSOR = 0, NONE_COMPRESSOR = 1, MAX_COMPRESSOR = 16 };
}  // namespace common
}  // namespace oceanbase
