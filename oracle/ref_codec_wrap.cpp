// extern "C" doors onto the REFERENCE's integer stream codecs (deps/oblib/src/lib/codec, compiled from
// /root/reference, not copied): the codec each ObIntegerStream::EncodingType maps to is the one
// ObIntegerStreamEncoder::do_codec_encode / ObIntegerStreamDecoder::decode_body instantiate
// (cs_encoding/ob_integer_stream_encoder.h:116-188, ob_integer_stream_decoder.h:157-229).
// tests/test_stream_codec_kat.py pins the writer's stream encoders, the oracle's and the device's stream
// decoders against these. Built only where /root/reference exists.
#include "lib/codec/ob_composite_codec.h"
#include "lib/codec/ob_simd_fixed_pfor.h"
#include "lib/codec/ob_double_delta_zigzag_rle.h"
#include "lib/codec/ob_double_delta_zigzag_pfor.h"
#include "lib/codec/ob_delta_zigzag_rle.h"
#include "lib/codec/ob_delta_zigzag_pfor.h"
#include "lib/codec/ob_xor_fixed_pfor.h"

namespace oceanbase { namespace common {
// common/ob_target_specific.h declares these two; the reference defines them in ob_target_specific.cpp from its
// CpuFlagSet. Here: the compiler's own CPU probe.
uint32_t arches = 0;
void init_arches() {
  arches = 0;
  if (__builtin_cpu_supports("sse4.2")) arches |= ObTargetArch::SSE42;
  if (__builtin_cpu_supports("avx")) arches |= ObTargetArch::AVX;
  if (__builtin_cpu_supports("avx2")) arches |= ObTargetArch::AVX2;
  if (__builtin_cpu_supports("avx512bw")) arches |= ObTargetArch::AVX512;
}
} }

using namespace oceanbase::common;

namespace {
struct MallocAllocator : public ObIAllocator {
  void *alloc(const int64_t size) override { return malloc((size_t)size); }
  void free(void *p) override { ::free(p); }
};
// ObIntegerStream::EncodingType (cs_encoding/ob_stream_encoding_struct.h:64-76)
enum { RAW = 1, DOUBLE_DELTA_ZIGZAG_RLE = 2, DOUBLE_DELTA_ZIGZAG_PFOR = 3, DELTA_ZIGZAG_RLE = 4, DELTA_ZIGZAG_PFOR = 5,
       SIMD_FIXEDPFOR = 6, UNIVERSAL_COMPRESS = 7, XOR_FIXED_PFOR = 8 };

template <typename F>
int with_codec(int type, int uint_bytes, F f) {
  static const bool arch_inited = (init_arches(), true);
  (void)arch_inited;
  MallocAllocator alloc;
  auto run = [&](ObCodec &c) {
    c.set_uint_bytes((uint8_t)uint_bytes);
    c.set_pfor_packing_type(ObCodec::CPU_ARCH_INDEPENDANT_SCALAR);
    c.set_allocator(alloc);
    return f(c);
  };
  switch (type) {
    case RAW: { ObSimpleBitPacking c; c.set_uint_packing_bits((uint8_t)(uint_bytes * 8)); return run(c); }
    case 100: { ObSimpleBitPacking c; return run(c); }   // [b][n x b bits]: the remainder coder of the PFOR family
    case SIMD_FIXEDPFOR: { ObCompositeCodec<ObSIMDFixedPFor, ObSimpleBitPacking> c; return run(c); }
    case DOUBLE_DELTA_ZIGZAG_RLE: { ObDoubleDeltaZigzagRle c; return run(c); }
    case DOUBLE_DELTA_ZIGZAG_PFOR: { ObDoubleDeltaZigzagPFor c; return run(c); }
    case DELTA_ZIGZAG_RLE: { ObDeltaZigzagRle c; return run(c); }
    case DELTA_ZIGZAG_PFOR: { ObDeltaZigzagPFor c; return run(c); }
    case XOR_FIXED_PFOR: { ObXorFixedPfor c; return run(c); }
    default: return OB_NOT_SUPPORTED;
  }
}
}  // namespace

extern "C" {
int ref_codec_encode(int type, int uint_bytes, const void *in, uint64_t count, void *out, uint64_t out_cap, uint64_t *out_len) {
  uint64_t pos = 0;
  const int ret = with_codec(type, uint_bytes, [&](ObCodec &c) {
    return c.encode((const char *)in, count * (uint64_t)uint_bytes, (char *)out, out_cap, pos);
  });
  *out_len = pos;
  return ret;
}
int ref_codec_decode(int type, int uint_bytes, const void *in, uint64_t in_len, uint64_t count, void *out, uint64_t *consumed) {
  uint64_t in_pos = 0, out_pos = 0;
  const int ret = with_codec(type, uint_bytes, [&](ObCodec &c) {
    return c.decode((const char *)in, in_len, in_pos, count, (char *)out, count * (uint64_t)uint_bytes, out_pos);
  });
  if (consumed) *consumed = in_pos;
  return ret;
}
}
