// Stand-in for oblib's serialization macros: only the declarations ObBatchChecksum (lib/checksum/ob_crc64.h) needs so
// that ob_crc64.cpp compiles unmodified. The checker never serializes anything.
#pragma once
#include "ob_codec_shim.h"
#define OB_UNIS_VERSION(v)                                                      \
public:                                                                         \
  int serialize(char *buf, const int64_t buf_len, int64_t &pos) const;          \
  int deserialize(const char *buf, const int64_t data_len, int64_t &pos);       \
  int64_t get_serialize_size() const
#define OB_DEF_SERIALIZE(cls) int cls::serialize(char *buf, const int64_t buf_len, int64_t &pos) const
#define OB_DEF_DESERIALIZE(cls) int cls::deserialize(const char *buf, const int64_t data_len, int64_t &pos)
#define OB_DEF_SERIALIZE_SIZE(cls) int64_t cls::get_serialize_size() const
#define CLOG_LOG(...) ((void)0)
#define _OB_LOG_RET(...) ((void)0)
#define _OB_LOG(...) ((void)0)
namespace oceanbase { namespace common {
constexpr int OB_SERIALIZE_ERROR = -4010;
constexpr int OB_DESERIALIZE_ERROR = -4011;
namespace serialization {
inline int encode_i64(char *, const int64_t, int64_t &, int64_t) { return OB_NOT_SUPPORTED; }
inline int decode_i64(const char *, const int64_t, int64_t &, int64_t *) { return OB_NOT_SUPPORTED; }
inline int64_t encoded_length_i64(int64_t) { return 8; }
}  // namespace serialization
} }
