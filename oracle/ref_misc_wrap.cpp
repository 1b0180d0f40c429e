#include "ob_codec_shim.h"
// extern "C" doors onto two more pieces of the REFERENCE that compile from their own sources (from /root/reference, not
// copied), checker only:
//   * the payload checksum of a micro block: ob_crc64_sse42 = crc32c, seed 0, no final xor, in its three implementations
//     (hardware instruction, byte table, slicing-by-8 table; deps/oblib/src/lib/checksum/ob_crc64.cpp:423-1103)
//   * ObIntegerArray / ObIntArrayFuncTable (storage/blocksstable/encoding/ob_integer_array.h): the row-index / RLE
//     row-id arrays with their lower_bound / upper_bound
// tests/test_checksum_ref_kat.py pins the oracle's and the writer's checksums and the oracle's run lookups to them.
#include "lib/checksum/ob_crc64.h"
#include "ob_integer_array.h"

using namespace oceanbase::common;
using namespace oceanbase::blocksstable;

extern "C" {
// never reached: the checker calls the implementations directly, not the vendor dispatch
unsigned int crc32_iscsi(unsigned char *, int, unsigned int) { abort(); }

uint64_t ref_crc64_sse42(uint64_t crc, const char *buf, int64_t len) { return crc64_sse42(crc, buf, len); }
uint64_t ref_crc64_sse42_manually(uint64_t crc, const char *buf, int64_t len) { return crc64_sse42_manually(crc, buf, len); }
uint64_t ref_fast_crc64_sse42_manually(uint64_t crc, const char *buf, int64_t len) { return fast_crc64_sse42_manually(crc, buf, len); }

int64_t ref_int_array_at(const void *array, int64_t byte, int64_t idx) {
  ObIntegerArrayGenerator gen;
  if (gen.init(static_cast<const char *>(array), byte) != OB_SUCCESS) return INT64_MIN;
  return gen.get_array().at(idx);
}
int64_t ref_int_array_lower_bound(const void *array, int64_t byte, int64_t begin, int64_t end, int64_t key) {
  return ObIntArrayFuncTable::instance(byte).lower_bound_(array, begin, end, key);
}
int64_t ref_int_array_upper_bound(const void *array, int64_t byte, int64_t begin, int64_t end, int64_t key) {
  return ObIntArrayFuncTable::instance(byte).upper_bound_(array, begin, end, key);
}
}
