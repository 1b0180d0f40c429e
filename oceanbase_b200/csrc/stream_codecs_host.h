// Integer stream ENCODERS of CS_ENCODING_ROW_STORE blocks (host side, part of libobgpu_writer.so).
//
// Restates the encode side of the reference's codec library (deps/oblib/src/lib/codec/) as ObIntegerStreamEncoder
// instantiates it per ObIntegerStream::EncodingType (cs_encoding/ob_integer_stream_encoder.h:116-188), with the
// CPU_ARCH_INDEPENDANT_SCALAR packing (plain LSB-first bit streams):
//   RAW                       ObSimpleBitPacking at the byte width                ob_codecs.h:236-318
//   SIMD_FIXEDPFOR            ObSIMDFixedPFor blocks of 128 + ObSimpleBitPacking   ob_simd_fixed_pfor.h:35-250, ob_composite_codec.h:84-110
//   DELTA_ZIGZAG_PFOR         ObDeltaZigzagFixedPfor::encode                       ob_delta_zigzag_pfor.h:34-93
//   DOUBLE_DELTA_ZIGZAG_PFOR  ObDoubleDeltaZigzagFixedPfor::encode                 ob_double_delta_zigzag_pfor.h:34-98
//   XOR_FIXED_PFOR            ObXorFixedPforInner::encode                          ob_xor_fixed_pfor.h:44-127
//   DELTA_ZIGZAG_RLE          ObDeltaZigzagRleInner::encode                        ob_delta_zigzag_rle.h:33-183
//   DOUBLE_DELTA_ZIGZAG_RLE   ObDoubleDeltaZigzagRleInner::encode                  ob_double_delta_zigzag_rle.h:31-182
// and the codec choice of ObIntegerStreamEncoder::dectect_candidate_codec (:307-400): the candidate that encodes a
// sample smallest wins, RAW when nothing beats it. Byte-exact against the real encoders: tests/test_stream_codec_kat.py.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace obstream {

enum : int { T_RAW = 1, T_DD_RLE = 2, T_DD_PFOR = 3, T_D_RLE = 4, T_D_PFOR = 5, T_FIXEDPFOR = 6, T_UNIVERSAL = 7, T_XOR_PFOR = 8 };

inline uint32_t gccbits(uint64_t v) { return v == 0 ? 0u : 64u - (uint32_t)__builtin_clzll(v); }
inline uint32_t pad8(uint32_t x) { return (x + 7u) / 8u; }
inline uint64_t mask_of(uint32_t wb) { return wb >= 8 ? ~0ull : ((1ull << (wb * 8)) - 1ull); }

// plain LSB-first bit writer
struct BitWriter {
  std::vector<uint8_t> &out;
  uint64_t acc = 0;
  uint32_t n = 0;
  explicit BitWriter(std::vector<uint8_t> &o) : out(o) {}
  void put(uint64_t v, uint32_t w) {  // w <= 64
    while (w > 0) {
      const uint32_t take = std::min<uint32_t>(w, 64u - n);
      const uint64_t part = take == 64 ? v : (v & ((1ull << take) - 1ull));
      acc |= part << n;
      n += take;
      v = take == 64 ? 0 : v >> take;
      w -= take;
      while (n >= 8) { out.push_back((uint8_t)acc); acc >>= 8; n -= 8; }
    }
  }
  void flush() { if (n > 0) { out.push_back((uint8_t)acc); acc = 0; n = 0; } }
};

inline void pack_bits(const uint64_t *v, size_t n, uint32_t b, std::vector<uint8_t> &out) {   // scalar_bit_packing
  BitWriter w(out);
  for (size_t i = 0; i < n; ++i) w.put(b >= 64 ? v[i] : (v[i] & ((1ull << b) - 1ull)), b);
  w.flush();
}

inline void enc_simple_bp(const uint64_t *v, size_t n, std::vector<uint8_t> &out) {   // ObSimpleBitPacking, packing bits 0
  uint64_t acc = 0;
  for (size_t i = 0; i < n; ++i) acc |= v[i];
  const uint32_t b = gccbits(acc);
  out.push_back((uint8_t)b);
  pack_bits(v, n, b, out);
}

// one ObSIMDFixedPFor block: 128 values
inline void enc_pfor_block(const uint64_t *in, uint32_t wb, std::vector<uint8_t> &out) {
  const uint32_t n = 128;
  uint32_t cnt[65] = {0};
  uint64_t u = 0;
  for (uint32_t i = 0; i < n; ++i) { ++cnt[gccbits(in[i])]; u |= in[i]; }
  int32_t b = (int32_t)gccbits(u);
  const uint32_t bx0 = (uint32_t)b, bmp8 = pad8(n);
  int32_t ml = (int32_t)pad8(n * (uint32_t)b) + 1;
  uint32_t x = cnt[b];
  for (int32_t i = b - 1; i >= 0; --i) {   // find_most_fit_bx
    const int32_t l = (int32_t)(2 + bmp8 + pad8(x * (bx0 - (uint32_t)i)) + pad8(n * (uint32_t)i));
    x += cnt[i];
    if (l < ml) { ml = l; b = i; }
  }
  const uint32_t bx = bx0 - (uint32_t)b;
  (void)wb;
  if (bx == 0) {
    out.push_back((uint8_t)b);
    pack_bits(in, n, (uint32_t)b, out);
    return;
  }
  out.push_back((uint8_t)(0x80 | b));
  out.push_back((uint8_t)bx);
  const uint64_t msk = (1ull << b) - 1ull;   // b < 64 here (bx > 0)
  uint64_t xmap[2] = {0, 0}, inx[128], low[128];
  uint32_t xn = 0;
  for (uint32_t i = 0; i < n; ++i) {
    low[i] = in[i] & msk;
    if (in[i] > msk) { xmap[i >> 6] |= 1ull << (i & 63); inx[xn++] = in[i] >> b; }
  }
  const size_t at = out.size();
  out.resize(at + 16);
  memcpy(out.data() + at, xmap, 16);
  pack_bits(inx, xn, bx, out);
  pack_bits(low, n, (uint32_t)b, out);
}

inline uint64_t zigzag_enc(uint64_t d, uint32_t wb) {
  const uint32_t bits = wb * 8;
  const uint64_t m = mask_of(wb), sign = (d >> (bits - 1)) & 1ull;
  return ((d << 1) ^ (0ull - sign)) & m;
}
inline uint64_t bitrev_w(uint64_t v, uint32_t wb) {
  uint64_t r = 0;
  const uint32_t bits = wb * 8;
  for (uint32_t i = 0; i < bits; ++i)
    if ((v >> i) & 1ull) r |= 1ull << (bits - 1 - i);
  return r;
}

// kind 0 plain, 1 delta zigzag, 2 double delta zigzag, 3 xor
inline void enc_pfor_family(int kind, const uint64_t *vals, size_t count, uint32_t wb, std::vector<uint8_t> &out) {
  const uint64_t m = mask_of(wb);
  uint64_t start = 0, pd = 0, t[128];
  for (size_t done = 0; done < count;) {
    const size_t n = std::min<size_t>(128, count - done);
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
      const uint64_t u = vals[done + i] & m;
      switch (kind) {
        case 1: t[i] = zigzag_enc((u - start) & m, wb); start = u; break;
        case 2: { const uint64_t d1 = (u - start) & m; t[i] = zigzag_enc((d1 - pd) & m, wb); pd = d1; start = u; break; }
        case 3: t[i] = (u ^ start) & m; acc |= t[i]; start = u; break;
        default: t[i] = u; break;
      }
    }
    if (kind == 3) {
      const uint32_t b = wb * 8 - gccbits(acc);
      for (size_t i = 0; i < n; ++i) t[i] = bitrev_w(b >= 64 ? 0ull : ((t[i] << b) & m), wb);
      out.push_back((uint8_t)b);
    }
    if (n == 128) enc_pfor_block(t, wb, out);
    else enc_simple_bp(t, n, out);
    done += n;
  }
}

inline void rle_put_wide(BitWriter &w, uint64_t v, uint32_t width, bool split64) {
  if (split64 && width > 45) {   // ObBitUtils::put<uint64_t>: high part first, then the low 32 bits
    w.put(v >> 32, width - 32);
    w.put(v & 0xffffffffull, 32);
  } else {
    w.put(v, width);
  }
}

inline void enc_rle_family(bool dbl, const uint64_t *vals, size_t count, uint32_t wb, std::vector<uint8_t> &out) {
  static const uint32_t N2[9] = {0, 3, 6, 0, 6, 0, 0, 0, 6}, N3[9] = {0, 5, 12, 0, 10, 0, 0, 0, 12}, N4[9] = {0, 9, 17, 0, 17, 0, 0, 0, 20};
  const uint64_t m = mask_of(wb);
  const uint32_t n2 = N2[wb], n3 = N3[wb], n4 = N4[wb];
  BitWriter w(out);
  uint64_t start = 0, pd = 0, run = 0;
  auto emit = [&](uint64_t r, uint64_t delta) {   // encode_delta_and_repeat_cnt
    if (r > 18) {
      r -= 18;
      const uint32_t b = (gccbits(r) + 7) >> 3;
      w.put((uint64_t)(b - 1) << 7, 10);
      rle_put_wide(w, r, b << 3, true);
    } else {
      for (uint64_t k = 0; k < r; ++k) w.put(1, 1);
    }
    const uint64_t zz = zigzag_enc(delta, wb);
    if (zz == 0) return;
    if (zz < (1ull << (n2 - 1))) w.put((zz << 2) | 2, n2 + 2);
    else if (zz < (1ull << (n3 - 1))) w.put((zz << 3) | 4, n3 + 3);
    else if (zz < (1ull << (n4 - 1))) w.put((zz << 4) | 8, n4 + 4);
    else {
      const uint32_t b = (gccbits(zz) + 7) >> 3;
      w.put((uint64_t)(b - 1) << 4, 7);
      rle_put_wide(w, zz, b << 3, wb == 8);
    }
  };
  uint64_t last_delta = 0;
  for (size_t i = 0; i < count; ++i) {
    const uint64_t u = vals[i] & m;
    uint64_t delta;
    if (dbl) { const uint64_t d1 = (u - start) & m; delta = (d1 - pd) & m; pd = d1; }
    else delta = (u - start) & m;
    start = u;
    last_delta = delta;
    if (delta != 0) { emit(run, delta); run = 0; }
    else ++run;
  }
  if (run > 0) emit(run, last_delta);   // trailing zero deltas: last_delta == 0, no delta record follows
  w.flush();
}

// Encodes `count` values (low width_bytes bytes of each) with codec `type`; false: unknown type.
inline bool encode(int type, uint32_t wb, const uint64_t *vals, size_t count, std::vector<uint8_t> &out) {
  switch (type) {
    case T_RAW:
      for (size_t i = 0; i < count; ++i) {
        const size_t at = out.size();
        out.resize(at + wb);
        memcpy(out.data() + at, &vals[i], wb);
      }
      return true;
    case T_DD_RLE: enc_rle_family(true, vals, count, wb, out); return true;
    case T_DD_PFOR: enc_pfor_family(2, vals, count, wb, out); return true;
    case T_D_RLE: enc_rle_family(false, vals, count, wb, out); return true;
    case T_D_PFOR: enc_pfor_family(1, vals, count, wb, out); return true;
    case T_FIXEDPFOR: enc_pfor_family(0, vals, count, wb, out); return true;
    case T_XOR_PFOR: enc_pfor_family(3, vals, count, wb, out); return true;
    default: return false;
  }
}

// ObIntegerStreamEncoder::choose_stream_codec + dectect_candidate_codec: RAW below 4 values; otherwise every enabled
// candidate encodes a sample (all values below 1024, else max(25 %, 1024)) and the smallest wins (RAW = sample bytes).
inline int detect(uint32_t wb, const uint64_t *vals, size_t count, bool monotonic_inc) {
  if (count < 4) return T_RAW;
  const size_t sample = count < 1024 ? count : std::max<size_t>(count * 25 / 100, 1024);
  static const int kCandidates[] = {T_FIXEDPFOR, T_DD_RLE, T_DD_PFOR, T_D_RLE, T_D_PFOR, T_XOR_PFOR};
  int best = T_RAW;
  size_t best_len = sample * wb;
  // lib::ob_sort over the cost array is not stable: on equal cost the earlier candidate is kept here
  size_t cand_best_len = (size_t)-1;
  int cand_best = T_RAW;
  std::vector<uint8_t> tmp;
  for (int t : kCandidates) {
    if (t == T_XOR_PFOR && monotonic_inc) continue;
    tmp.clear();
    encode(t, wb, vals, sample, tmp);
    if (tmp.size() < cand_best_len) { cand_best_len = tmp.size(); cand_best = t; }
  }
  if (cand_best_len < best_len) best = cand_best;
  return best;
}

}  // namespace obstream
