/* CPU oracle, part 2: the integer stream codecs of CS_ENCODING_ROW_STORE blocks.
 *
 * TEST INFRASTRUCTURE ONLY (see ob_oracle.h). Plain-C restatement of the DECODE side of the reference's codec
 * library, deps/oblib/src/lib/codec/, as instantiated per stream type by ObIntegerStreamDecoder::decode_body
 * (cs_encoding/ob_integer_stream_decoder.h:157-229):
 *   RAW                       ObSimpleBitPacking at the stream's byte width            ob_codecs.h:236-404
 *   SIMD_FIXEDPFOR            ObCompositeCodec<ObSIMDFixedPFor, ObSimpleBitPacking>   ob_simd_fixed_pfor.h:253-330, ob_composite_codec.h:112-143
 *   DELTA_ZIGZAG_PFOR         ObDeltaZigzagFixedPfor::decode                            ob_delta_zigzag_pfor.h:96-139
 *   DOUBLE_DELTA_ZIGZAG_PFOR  ObDoubleDeltaZigzagFixedPfor::decode                      ob_double_delta_zigzag_pfor.h:101-150
 *   XOR_FIXED_PFOR            ObXorFixedPforInner::decode                               ob_xor_fixed_pfor.h:130-190
 *   DELTA_ZIGZAG_RLE          ObDeltaZigzagRleInner::decode                             ob_delta_zigzag_rle.h:186-320
 *   DOUBLE_DELTA_ZIGZAG_RLE   ObDoubleDeltaZigzagRleInner::decode                       ob_double_delta_zigzag_rle.h:184-310
 * with the CPU_ARCH_INDEPENDANT_SCALAR packing (plain LSB-first bit streams, ob_bp_helpers.h:1385-1459).
 * Pinning: tests/test_stream_codec_kat.py decodes what the REAL reference encoders produce (oracle/_ref/
 * libref_codec.so, compiled from /root/reference) and compares with the real decoders' output.
 */
#include "ob_oracle.h"

#include <string.h>

/* ---- plain LSB-first bit reader over [p, end) ------------------------------------------------------------ */
typedef struct bitrd { const uint8_t *p; int64_t len; int64_t bit; } bitrd;

static uint64_t rd_bits(const bitrd *r, int64_t bit, uint32_t w) {  /* w <= 64 */
  uint64_t v = 0;
  /* gather up to 9 bytes (bytes past the end read as zero, like ObBitUtils::d_slide) */
  const int64_t b0 = bit >> 3;
  const uint32_t sh = (uint32_t)(bit & 7);
  unsigned __int128 acc = 0;
  for (int k = 0; k < 9; ++k) {
    const int64_t at = b0 + k;
    const uint8_t byte = (at >= 0 && at < r->len) ? r->p[at] : 0;
    acc |= (unsigned __int128)byte << (8 * k);
  }
  acc >>= sh;
  v = (uint64_t)acc;
  if (w < 64) v &= (w == 0 ? 0ull : ((1ull << w) - 1ull));
  return v;
}

static uint64_t mask_w(uint32_t bytes) { return bytes >= 8 ? ~0ull : ((1ull << (bytes * 8)) - 1ull); }

static void put_val(void *out, uint32_t wb, int64_t i, uint64_t v) {
  switch (wb) {
    case 1: ((uint8_t *)out)[i] = (uint8_t)v; break;
    case 2: ((uint16_t *)out)[i] = (uint16_t)v; break;
    case 4: ((uint32_t *)out)[i] = (uint32_t)v; break;
    default: ((uint64_t *)out)[i] = v; break;
  }
}

/* ObSimpleBitPacking with uint_packing_bits == 0: [b][n x b bits] (the remainder coder of the PFOR family) */
static int dec_simple_bp(const uint8_t *in, int64_t in_len, int64_t *pos, int64_t n, uint64_t *vals) {
  if (*pos >= in_len) return ORA_INVALID_DATA;
  const uint32_t b = in[(*pos)++];
  if (b > 64) return ORA_INVALID_DATA;
  bitrd r = {in + *pos, in_len - *pos, 0};
  for (int64_t i = 0; i < n; ++i) vals[i] = rd_bits(&r, i * (int64_t)b, b);
  *pos += (n * (int64_t)b + 7) / 8;
  return *pos <= in_len ? ORA_SUCCESS : ORA_INVALID_DATA;
}

/* one ObSIMDFixedPFor block of 128 values: [b | 0x80][bx][exception bitmap 16 B][exceptions][128 x b bits] */
static int dec_pfor_block(const uint8_t *in, int64_t in_len, int64_t *pos, uint32_t wb, uint64_t *vals) {
  if (*pos >= in_len) return ORA_INVALID_DATA;
  uint32_t b = in[(*pos)++];
  if (!(b & 0x80)) {
    if (b > wb * 8) return ORA_INVALID_DATA;
    bitrd r = {in + *pos, in_len - *pos, 0};
    for (int i = 0; i < 128; ++i) vals[i] = rd_bits(&r, (int64_t)i * b, b);
    *pos += 16 * (int64_t)b;
    return *pos <= in_len ? ORA_SUCCESS : ORA_INVALID_DATA;
  }
  b &= 0x7f;
  if (*pos + 17 > in_len) return ORA_INVALID_DATA;
  const uint32_t bx = in[(*pos)++];
  if (b > wb * 8 || bx > wb * 8) return ORA_INVALID_DATA;
  uint64_t xm[2];
  memcpy(xm, in + *pos, 16);
  *pos += 16;
  const int xn = __builtin_popcountll(xm[0]) + __builtin_popcountll(xm[1]);
  uint64_t ex[128];
  {
    bitrd r = {in + *pos, in_len - *pos, 0};
    for (int i = 0; i < xn; ++i) ex[i] = rd_bits(&r, (int64_t)i * bx, bx);
    *pos += ((int64_t)xn * bx + 7) / 8;
  }
  if (*pos + 16 * (int64_t)b > in_len) return ORA_INVALID_DATA;
  {
    bitrd r = {in + *pos, in_len - *pos, 0};
    for (int i = 0; i < 128; ++i) vals[i] = rd_bits(&r, (int64_t)i * b, b);
    *pos += 16 * (int64_t)b;
  }
  int k = 0;
  for (int i = 0; i < 128; ++i)
    if ((xm[i >> 6] >> (i & 63)) & 1ull) vals[i] |= (b >= 64 ? 0ull : (ex[k++] << b));
  return ORA_SUCCESS;
}

static uint64_t zigzag_dec(uint64_t v, uint64_t m) { return ((v >> 1) ^ (0ull - (v & 1ull))) & m; }

static uint64_t bitrev_w(uint64_t v, uint32_t wb) {   /* bit_reverse<UIntT> */
  uint64_t r = 0;
  const uint32_t bits = wb * 8;
  for (uint32_t i = 0; i < bits; ++i)
    if ((v >> i) & 1ull) r |= 1ull << (bits - 1 - i);
  return r;
}

/* PFOR family: kind 0 plain, 1 delta zigzag, 2 double delta zigzag, 3 xor */
static int dec_pfor_family(int kind, const uint8_t *in, int64_t in_len, uint32_t wb, int64_t count, void *out, int64_t *consumed) {
  const uint64_t m = mask_w(wb);
  int64_t pos = 0, done = 0;
  uint64_t start = 0, pd = 0;
  uint64_t vals[128];
  while (done < count) {
    const int64_t n = count - done >= 128 ? 128 : count - done;
    uint32_t xb = 0;
    if (kind == 3) {
      if (pos >= in_len) return ORA_INVALID_DATA;
      xb = in[pos++];
      if (xb > wb * 8) return ORA_INVALID_DATA;
    }
    int ret = n == 128 ? dec_pfor_block(in, in_len, &pos, wb, vals) : dec_simple_bp(in, in_len, &pos, n, vals);
    if (ret) return ret;
    for (int64_t i = 0; i < n; ++i) {
      uint64_t v = vals[i] & m;
      switch (kind) {
        case 1: v = (zigzag_dec(v, m) + start) & m; start = v; break;
        case 2: v = (zigzag_dec(v, m) + start + pd) & m; pd = (v - start) & m; start = v; break;
        case 3: v = ((xb >= 64 ? 0ull : (bitrev_w(v, wb) >> xb)) ^ start) & m; start = v; break;
        default: break;
      }
      put_val(out, wb, done + i, v);
    }
    done += n;
  }
  if (consumed) *consumed = pos;
  return ORA_SUCCESS;
}

/* RLE family: one LSB-first stream of variable-length codes (tiers N2 / N3 / N4 per width: ob_bp_util.h:336-388) */
static int dec_rle_family(int dbl, const uint8_t *in, int64_t in_len, uint32_t wb, int64_t count, void *out, int64_t *consumed) {
  static const uint32_t N2[9] = {0, 3, 6, 0, 6, 0, 0, 0, 6}, N3[9] = {0, 5, 12, 0, 10, 0, 0, 0, 12}, N4[9] = {0, 9, 17, 0, 17, 0, 0, 0, 20};
  const uint64_t m = mask_w(wb);
  const uint32_t n2 = N2[wb], n3 = N3[wb], n4 = N4[wb];
  bitrd r = {in, in_len, 0};
  int64_t bit = 0, done = 0;
  uint64_t start = 0, pd = 0;
  while (done < count) {
    const uint64_t peek = rd_bits(&r, bit, 7);
    uint64_t delta;
    if (peek & 1) { bit += 1; delta = 0; }
    else if (peek & 2) { delta = rd_bits(&r, bit + 2, n2); bit += n2 + 2; }
    else if (peek & 4) { delta = rd_bits(&r, bit + 3, n3); bit += n3 + 3; }
    else if (peek & 8) { delta = rd_bits(&r, bit + 4, n4); bit += n4 + 4; }
    else {
      const uint32_t b = (uint32_t)(peek >> 4);
      bit += 7;
      if (b == 1) return ORA_ERR_UNEXPECTED;          /* "can not be overflow" */
      if (b == 0) {                                    /* repeat count record: 3 bits (bytes - 1), then the count */
        const uint32_t nb = (uint32_t)rd_bits(&r, bit, 3) + 1;
        bit += 3;
        uint64_t rep = rd_bits(&r, bit, nb * 8 > 57 ? 57 : nb * 8);   /* get<uint32_t>: bitget57 */
        bit += nb * 8;
        rep += 18;                                     /* BASE_REPEAT_CNT */
        if (rep > (uint64_t)(count - done)) return ORA_INVALID_DATA;
        for (uint64_t k = 0; k < rep; ++k) {
          if (dbl) start = (start + pd) & m;
          put_val(out, wb, done++, start);
        }
        continue;
      }
      const uint32_t w = (b + 1) * 8;                  /* delta of b + 1 bytes; uint64: high part first when > 45 bits */
      if (wb == 8 && w > 45) {
        const uint64_t hi = rd_bits(&r, bit, w - 32);
        const uint64_t lo = rd_bits(&r, bit + (w - 32), 32);
        delta = (hi << 32) | lo;
      } else {
        delta = rd_bits(&r, bit, w);
      }
      bit += w;
    }
    if (bit > in_len * 8 + 64) return ORA_INVALID_DATA;
    const uint64_t d = zigzag_dec(delta & m, m);
    if (dbl) { pd = (pd + d) & m; start = (start + pd) & m; }
    else start = (start + d) & m;
    put_val(out, wb, done++, start);
  }
  if (consumed) *consumed = (bit + 7) / 8;
  return ORA_SUCCESS;
}

int ora_int_stream_decode(int32_t type, uint32_t width_bytes, const void *in, int64_t in_len, int64_t count, void *out,
                          int64_t *consumed) {
  if (!in || !out || count < 0 || (width_bytes != 1 && width_bytes != 2 && width_bytes != 4 && width_bytes != 8))
    return ORA_INVALID_ARGUMENT;
  if (count == 0) { if (consumed) *consumed = 0; return ORA_SUCCESS; }
  const uint8_t *p = (const uint8_t *)in;
  switch (type) {
    case 1:  /* RAW */
      if ((int64_t)width_bytes * count > in_len) return ORA_INVALID_DATA;
      memcpy(out, in, (size_t)width_bytes * (size_t)count);
      if (consumed) *consumed = (int64_t)width_bytes * count;
      return ORA_SUCCESS;
    case 2: return dec_rle_family(1, p, in_len, width_bytes, count, out, consumed);
    case 3: return dec_pfor_family(2, p, in_len, width_bytes, count, out, consumed);
    case 4: return dec_rle_family(0, p, in_len, width_bytes, count, out, consumed);
    case 5: return dec_pfor_family(1, p, in_len, width_bytes, count, out, consumed);
    case 6: return dec_pfor_family(0, p, in_len, width_bytes, count, out, consumed);
    case 8: return dec_pfor_family(3, p, in_len, width_bytes, count, out, consumed);
    default: return ORA_NOT_SUPPORTED;   /* 7 UNIVERSAL_COMPRESS needs a general compressor (deps/3rd, absent) */
  }
}
