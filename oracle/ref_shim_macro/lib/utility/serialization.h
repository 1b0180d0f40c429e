#pragma once
#include "ob_macro_shim.h"
// declarations only: ObRecordHeader (common/ob_record_header.h) is never serialized by the checker
#define NEED_SERIALIZE_AND_DESERIALIZE                                  \
  int serialize(char *buf, const int64_t buf_len, int64_t &pos) const;  \
  int deserialize(const char *buf, const int64_t data_len, int64_t &pos); \
  int64_t get_serialize_size() const
