// extern "C" doors onto the REFERENCE's ObBitStream (compiled from /root/reference, not copied):
// used by tests/test_ref_pin.py to pin oracle/ob_oracle.c's bit-stream restatement and the
// writer's bit packing against the real thing. Built only where /root/reference exists.
#include "ob_bit_stream.h"
using oceanbase::blocksstable::ObBitStream;
extern "C" {
uint64_t ref_bs_get(const unsigned char *buf, int64_t offset, int64_t cnt) {
  int64_t v = 0;
  ObBitStream::get(buf, offset, cnt, v);
  return (uint64_t)v;
}
uint64_t ref_bs_get_unpack(const unsigned char *buf, int64_t offset, int64_t cnt, int64_t bs_len) {
  int64_t v = 0;
  ObBitStream::get_unpack_func(cnt)(buf, offset, cnt, bs_len, v);
  return (uint64_t)v;
}
void ref_bs_memory_safe_set(unsigned char *buf, int64_t pos, int64_t len, uint64_t v) {
  ObBitStream::memory_safe_set(buf, pos, len, v);
}
int ref_bs_set(unsigned char *buf, int64_t buf_len, int64_t offset, int64_t cnt, int64_t value) {
  ObBitStream bs;
  int ret = bs.init(buf, buf_len);
  if (ret == 0) ret = bs.set(offset, cnt, value);
  return ret;
}
uint64_t ref_bs_get_mask(int64_t len) { return ObBitStream::get_mask(len); }
}
