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

#include <stdlib.h>
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

/* =============================================================================================
 * CS block -> the same block with every integer stream restated as RAW.
 *
 * The reference decodes non-RAW integer streams once, when a block enters the block cache
 * (ObCSMicroBlockTransformer::full_transform, cs_encoding/ob_cs_micro_block_transformer.cpp:721-898 ->
 * ObIntegerStreamDecoder::transform_to_raw_array, ob_integer_stream_decoder.cpp:398-431), into an in-memory
 * image that also carries C++ structs (ObMicroBlockTransformDesc, decoder ctxs). Here the result is restated
 * in the ON-DISK layout instead -- what ObMicroBlockCSEncoder::build_block would have written had every stream
 * chosen RAW: [header][ObAllColumnHeader][ObCSColumnHeader x ncol][per column: meta + streams (ObIntegerStreamMeta
 * with type RAW + width-byte array)][all string data][stream END offsets, RAW] -- so the RAW-only readers (this
 * oracle's block decoder, the device kernels) run on it unchanged. The micro header is copied as is (its length /
 * checksum fields keep describing the on-disk block, as in the reference's deep copy).
 * ============================================================================================= */
typedef struct xmeta { uint8_t version, attr, type, wtag; uint32_t width; int64_t meta_len; } xmeta;

static uint32_t x_rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t x_rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

static int x_parse_meta(const uint8_t *p, int64_t len, xmeta *m) {   /* ObIntegerStreamMeta::deserialize */
  if (len < 4) return ORA_INVALID_DATA;
  int64_t pos = 4;
  m->version = p[0]; m->attr = p[1]; m->type = p[2]; m->wtag = p[3];
  for (int k = 0; k < 2; ++k) {
    if (!(m->attr & (1 << k))) continue;
    for (;;) {   /* vi64 */
      if (pos >= len) return ORA_INVALID_DATA;
      if (!(p[pos++] & 0x80)) break;
    }
  }
  if (m->attr & 0x4) return ORA_NOT_SUPPORTED;   /* decimal int: precision fields follow */
  if (m->version > 0) { if (pos >= len) return ORA_INVALID_DATA; ++pos; }
  if (m->wtag > 3) return ORA_NOT_SUPPORTED;
  m->width = 1u << m->wtag;
  m->meta_len = pos;
  return ORA_SUCCESS;
}

static uint32_t x_bytes_for(uint64_t v) { return v <= 0xff ? 1u : (v <= 0xffff ? 2u : (v <= 0xffffffffull ? 4u : 8u)); }

int ora_cs_transform(const void *block, int64_t size, void *out, int64_t out_cap, int64_t *out_size) {
  if (!block || size < 64 || !out_size) return ORA_INVALID_ARGUMENT;
  const uint8_t *p = (const uint8_t *)block;
  const uint32_t header_size = x_rd32(p + 4), ncol = x_rd16(p + 10), rows = x_rd32(p + 16);
  if (p[20] != 3 /* CS_ENCODING_ROW_STORE */ || header_size < 64 || (int64_t)header_size + 12 + 4ll * ncol > size) return ORA_INVALID_ARGUMENT;
  const uint8_t *ah = p + header_size;
  if (ah[0] != 0 || (ah[1] & 0x3)) return ORA_NOT_SUPPORTED;
  const uint32_t all_string_len = x_rd32(ah + 2), offsets_len = x_rd32(ah + 6), n_streams_total = x_rd16(ah + 10);
  if ((int64_t)offsets_len + all_string_len > size - header_size) return ORA_INVALID_DATA;
  const uint32_t str_begin = (uint32_t)(size - offsets_len - all_string_len);
  /* stream END offsets (an integer stream of its own, any codec) */
  uint64_t *ends = (uint64_t *)calloc(n_streams_total + 1u, sizeof(uint64_t));
  uint64_t *new_ends = (uint64_t *)calloc(n_streams_total + 1u, sizeof(uint64_t));
  if (!ends || !new_ends) { free(ends); free(new_ends); return ORA_ERR_UNEXPECTED; }
  int ret = ORA_SUCCESS;
  xmeta om;
  memset(&om, 0, sizeof(om));
  if (n_streams_total > 0) {
    const uint8_t *so = p + size - offsets_len;
    if ((ret = x_parse_meta(so, offsets_len, &om))) { free(ends); free(new_ends); return ret; }
    uint8_t tmp[8 * 4096];
    if ((om.attr & 0x3) || n_streams_total > 4096) { free(ends); free(new_ends); return ORA_NOT_SUPPORTED; }
    if ((ret = ora_int_stream_decode(om.type, om.width, so + om.meta_len, offsets_len - om.meta_len, n_streams_total, tmp, 0))) { free(ends); free(new_ends); return ret; }
    for (uint32_t k = 0; k < n_streams_total; ++k) {
      uint64_t v = 0;
      memcpy(&v, tmp + (size_t)k * om.width, om.width);
      ends[k] = v;
    }
  }
  /* output: worst case every stream grows to 8 bytes per value */
  const int64_t cap_need = size + 16;   /* rechecked as we go */
  (void)cap_need;
  uint8_t *o = (uint8_t *)out;
  int64_t opos = 0;
#define X_PUT(src, n) do { const int64_t n__ = (int64_t)(n); if (o) { if (opos + n__ > out_cap) { ret = ORA_BUF_NOT_ENOUGH; goto done; } memcpy(o + opos, (src), (size_t)n__); } opos += n__; } while (0)
  X_PUT(p, header_size + 12u + 4u * ncol);
  {
    const int64_t bitmap_bytes = ((int64_t)rows + 7) / 8;
    uint32_t pos = header_size + 12u + 4u * ncol;
    int32_t si = 0;   /* next stream index */
    for (uint32_t c = 0; c < ncol && ret == ORA_SUCCESS; ++c) {
      const uint8_t *h = ah + 12 + 4 * c;
      const uint8_t type = h[1], attrs = h[2];
      int64_t meta_len;
      /* streams of this column: (is_string, count) in order */
      int n_s = 0, is_str[3] = {0, 0, 0};
      int64_t cnt[3] = {0, 0, 0};
      if (type == 0) { meta_len = ((attrs & 0x02) ? bitmap_bytes : 0) + ((attrs & 0x08) ? bitmap_bytes : 0); n_s = 1; cnt[0] = rows; }
      else if (type == 1) {
        meta_len = ((attrs & 0x02) ? bitmap_bytes : 0) + ((attrs & 0x08) ? bitmap_bytes : 0);
        is_str[0] = 1; n_s = 1;
        if (!(attrs & 0x01)) { cnt[1] = rows; n_s = 2; }
      } else if (type == 2 || type == 3) {
        if ((int64_t)pos + 10 > size) { ret = ORA_INVALID_DATA; break; }
        const uint8_t *dm = p + pos;
        const uint32_t distinct = x_rd32(dm + 2);
        const int64_t ref_cnt = (dm[1] & 0x4) ? (int64_t)x_rd32(dm + 6) : (int64_t)rows;
        meta_len = 10 + ((attrs & 0x08) ? bitmap_bytes : 0);
        if (distinct > 0) {
          if (type == 2) { cnt[0] = distinct; cnt[1] = ref_cnt; n_s = 2; }
          else if (attrs & 0x01) { is_str[0] = 1; cnt[1] = ref_cnt; n_s = 2; }
          else { is_str[0] = 1; cnt[1] = distinct; cnt[2] = ref_cnt; n_s = 3; }
        }
      } else { ret = ORA_NOT_SUPPORTED; break; }
      if ((int64_t)pos + meta_len > size) { ret = ORA_INVALID_DATA; break; }
      X_PUT(p + pos, meta_len);
      uint32_t at = pos + (uint32_t)meta_len;
      for (int k = 0; k < n_s; ++k, ++si) {
        if (si >= (int32_t)n_streams_total) { ret = ORA_INVALID_DATA; break; }
        const uint32_t end = (uint32_t)ends[si];
        if (end < at || end > (uint32_t)size) { ret = ORA_INVALID_DATA; break; }
        if (is_str[k]) {
          X_PUT(p + at, end - at);                 /* ObStringStreamMeta only: the bytes live in the all-string-data area */
        } else {
          xmeta m;
          if ((ret = x_parse_meta(p + at, end - at, &m))) break;
          if (o) {
            if (opos + m.meta_len + cnt[k] * (int64_t)m.width > out_cap) { ret = ORA_BUF_NOT_ENOUGH; break; }
            memcpy(o + opos, p + at, (size_t)m.meta_len);
            o[opos + 2] = 1;                        /* ObIntegerStream::EncodingType::RAW */
            int64_t used = 0;
            ret = ora_int_stream_decode(m.type, m.width, p + at + m.meta_len, (int64_t)end - at - m.meta_len, cnt[k],
                                        o + opos + m.meta_len, &used);
            if (ret) break;
            if (used != (int64_t)end - at - m.meta_len) { ret = ORA_INVALID_DATA; break; }
          }
          opos += m.meta_len + cnt[k] * (int64_t)m.width;
        }
        new_ends[si] = (uint64_t)opos;
        at = end;
      }
      pos = n_s == 0 ? pos + (uint32_t)meta_len : (uint32_t)ends[si - 1];
    }
    if (ret == ORA_SUCCESS && si != (int32_t)n_streams_total) ret = ORA_INVALID_DATA;
    if (ret == ORA_SUCCESS && pos != str_begin) ret = ORA_INVALID_DATA;
  }
  if (ret) goto done;
  X_PUT(p + str_begin, all_string_len);
  {
    const int64_t off_at = opos;
    if (n_streams_total > 0) {
      const uint32_t w = x_bytes_for(new_ends[n_streams_total - 1]);
      if (w > 4) { ret = ORA_NOT_SUPPORTED; goto done; }
      uint8_t meta[8];
      memcpy(meta, p + size - offsets_len, (size_t)om.meta_len);
      meta[2] = 1;
      meta[3] = (uint8_t)(w == 1 ? 0 : (w == 2 ? 1 : 2));
      X_PUT(meta, om.meta_len);
      for (uint32_t k = 0; k < n_streams_total; ++k) X_PUT(&new_ends[k], w);
    }
    if (o) { const uint32_t nl = (uint32_t)(opos - off_at); memcpy(o + header_size + 6, &nl, 4); }
  }
done:
#undef X_PUT
  free(ends);
  free(new_ends);
  *out_size = opos;
  return ret;
}
