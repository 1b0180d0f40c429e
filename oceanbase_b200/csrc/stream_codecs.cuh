// CS_ENCODING_ROW_STORE blocks whose integer streams carry the non-RAW codecs: decoded once,
// at batch open, into a RAW restatement of the block that every scan kernel then reads unchanged.
//
// The reference does the same thing when a block enters its block cache: ObCSMicroBlockTransformer::full_transform
// (cs_encoding/ob_cs_micro_block_transformer.cpp:721-898) -> ObIntegerStreamDecoder::transform_to_raw_array
// (ob_integer_stream_decoder.cpp:398-431) turns every non-RAW stream into a width-byte array. Its output image also
// carries in-memory C++ structs (ObMicroBlockTransformDesc, decoder ctxs); here the result keeps the ON-DISK layout --
// what ObMicroBlockCSEncoder::build_block would have written had every stream chosen RAW:
//   [header][ObAllColumnHeader][ObCSColumnHeader x ncol][per column: meta + streams (ObIntegerStreamMeta type RAW +
//   width-byte array)][all string data][stream END offsets, RAW]
// Three kernels over a page batch:
//   cs_survey_kernel   one thread per block: walks the column headers / stream offsets, sizes the restated block
//   cs_rewrite_kernel  one thread per block: writes everything but the decoded arrays, emits one job per non-RAW stream
//   cs_decode_kernel   one thread per job: the codec's decoder (deps/oblib/src/lib/codec, CPU_ARCH_INDEPENDANT_SCALAR
//                      packing): SIMD_FIXEDPFOR, DELTA / DOUBLE_DELTA_ZIGZAG_PFOR, XOR_FIXED_PFOR (blocks of 128 +
//                      simple bit packing), DELTA / DOUBLE_DELTA_ZIGZAG_RLE (one variable-length bit code)
// UNIVERSAL_COMPRESS streams and compressed string areas need a general-purpose decompressor: OB_NOT_SUPPORTED.
#pragma once

namespace obcs {

enum : int { XF_OK = 0, XF_NONRAW = 1, XF_UNSUPPORTED = 2, XF_CORRUPT = 4 };

struct StreamJob {
  uint64_t src;       // byte offset of the codec bytes in the source image
  uint64_t dst;       // byte offset of the raw array in the restated image
  uint32_t enc_len;
  uint32_t count;
  uint8_t type, width;
  uint8_t pad[6];
};
static_assert(sizeof(StreamJob) == 32, "StreamJob");

struct XMeta { uint32_t width, meta_len; uint8_t type, attr, ok; };

// Where a block of a restated batch came from: VEC_DISCRETE pointers must keep addressing the CALLER's block buffer.
// The all-string-data area is copied verbatim, so a string cell of the restated block sits str_delta bytes further
// on in the original one.
struct XformRec { uint64_t orig_off; int64_t str_delta; };

__device__ __forceinline__ uint32_t g_rd(const uint8_t *s, uint32_t off, uint32_t n) {   // n <= 4 bytes, little endian
  uint32_t v = 0;
  for (uint32_t k = 0; k < n; ++k) v |= (uint32_t)s[off + k] << (8u * k);
  return v;
}

__device__ __forceinline__ void x_parse_meta(const uint8_t *s, uint32_t at, uint32_t end, XMeta &m) {   // ObIntegerStreamMeta::deserialize
  m = XMeta{};
  if (at + 4u > end) return;
  const uint8_t version = s[at], attr = s[at + 1], wtag = s[at + 3];
  uint32_t pos = at + 4u;
  for (int k = 0; k < 2; ++k) {
    if (!(attr & (1 << k))) continue;
    for (;;) {
      if (pos >= end) return;
      if (!(s[pos++] & 0x80)) break;
    }
  }
  if ((attr & 0x4) || wtag > 3) return;   // decimal int / wide integers: not handled
  if (version > 0) { if (pos >= end) return; ++pos; }
  m.type = s[at + 2];
  m.attr = attr;
  m.width = 1u << wtag;
  m.meta_len = pos - at;
  m.ok = 1;
}

// ---- bit reader over global memory (bytes past the end read as zero, like ObBitUtils::d_slide) -----------------
struct BitRd {
  const uint8_t *p;
  int64_t len;
  __device__ __forceinline__ uint64_t get(int64_t bit, uint32_t w) const {   // w <= 64
    const int64_t b0 = bit >> 3;
    const uint32_t sh = (uint32_t)(bit & 7);
    uint64_t lo = 0;
    uint32_t hi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t at = b0 + k;
      if (at < len) lo |= (uint64_t)p[at] << (8 * k);
    }
    if (b0 + 8 < len) hi = p[b0 + 8];
    uint64_t v = sh ? ((lo >> sh) | ((uint64_t)hi << (64u - sh))) : lo;
    if (w < 64) v &= (w == 0 ? 0ull : ((1ull << w) - 1ull));
    return v;
  }
};

__device__ __forceinline__ void put_val(uint8_t *out, uint32_t wb, int64_t i, uint64_t v) {
  uint8_t *o = out + i * (int64_t)wb;
  for (uint32_t k = 0; k < wb; ++k) o[k] = (uint8_t)(v >> (8u * k));
}
__device__ __forceinline__ uint64_t mask_w(uint32_t wb) { return wb >= 8 ? ~0ull : ((1ull << (wb * 8u)) - 1ull); }
__device__ __forceinline__ uint64_t zigzag_dec(uint64_t v, uint64_t m) { return ((v >> 1) ^ (0ull - (v & 1ull))) & m; }
__device__ __forceinline__ uint64_t bitrev_w(uint64_t v, uint32_t wb) { return __brevll(v) >> (64u - wb * 8u); }

// PFOR family (ob_simd_fixed_pfor.h:253-330, ob_delta_zigzag_pfor.h:96-139, ob_double_delta_zigzag_pfor.h:101-150,
// ob_xor_fixed_pfor.h:130-190): kind 0 plain, 1 delta zigzag, 2 double delta zigzag, 3 xor. Returns bytes consumed or -1.
__device__ int64_t dec_pfor_family(int kind, const uint8_t *in, int64_t in_len, uint32_t wb, int64_t count, uint8_t *out) {
  const uint64_t m = mask_w(wb);
  int64_t pos = 0, done = 0;
  uint64_t start = 0, pd = 0;
  while (done < count) {
    const int64_t n = count - done >= 128 ? 128 : count - done;
    uint32_t xb = 0;
    if (kind == 3) {
      if (pos >= in_len) return -1;
      xb = in[pos++];
    }
    if (pos >= in_len) return -1;
    uint32_t b = in[pos++], bx = 0;
    uint64_t xm0 = 0, xm1 = 0;
    int64_t ex_pos = 0;
    if (n == 128 && (b & 0x80)) {   // exceptions: [bx][bitmap 16 B][xn x bx bits]
      b &= 0x7f;
      if (pos + 17 > in_len) return -1;
      bx = in[pos++];
      for (int k = 0; k < 8; ++k) { xm0 |= (uint64_t)in[pos + k] << (8 * k); xm1 |= (uint64_t)in[pos + 8 + k] << (8 * k); }
      pos += 16;
      ex_pos = pos;
      pos += ((int64_t)(__popcll(xm0) + __popcll(xm1)) * bx + 7) / 8;
    }
    if (b > 64 || bx > 64) return -1;
    const int64_t data_bytes = (n * (int64_t)b + 7) / 8;
    if (pos + data_bytes > in_len) return -1;
    const BitRd data{in + pos, in_len - pos}, exr{in + ex_pos, in_len - ex_pos};
    int xk = 0;
    for (int64_t i = 0; i < n; ++i) {
      uint64_t v = data.get(i * (int64_t)b, b);
      if (bx) {
        const uint64_t word = i < 64 ? xm0 : xm1;
        if ((word >> (i & 63)) & 1ull) { v |= b >= 64 ? 0ull : (exr.get((int64_t)xk * bx, bx) << b); ++xk; }
      }
      v &= m;
      if (kind == 1) { v = (zigzag_dec(v, m) + start) & m; start = v; }
      else if (kind == 2) { v = (zigzag_dec(v, m) + start + pd) & m; pd = (v - start) & m; start = v; }
      else if (kind == 3) { v = ((xb >= 64 ? 0ull : (bitrev_w(v, wb) >> xb)) ^ start) & m; start = v; }
      put_val(out, wb, done + i, v);
    }
    pos += data_bytes;
    done += n;
  }
  return pos;
}

// RLE family (ob_delta_zigzag_rle.h:186-320, ob_double_delta_zigzag_rle.h:184-310): one LSB-first stream of
// variable-length codes; tiers N2 / N3 / N4 per width (ob_bp_util.h:336-388).
__device__ int64_t dec_rle_family(bool dbl, const uint8_t *in, int64_t in_len, uint32_t wb, int64_t count, uint8_t *out) {
  const uint32_t n2 = wb == 1 ? 3u : 6u, n3 = wb == 1 ? 5u : (wb == 4 ? 10u : 12u), n4 = wb == 1 ? 9u : (wb == 8 ? 20u : 17u);
  const uint64_t m = mask_w(wb);
  const BitRd r{in, in_len};
  int64_t bit = 0, done = 0;
  uint64_t start = 0, pd = 0;
  while (done < count) {
    const uint64_t peek = r.get(bit, 7);
    uint64_t delta;
    if (peek & 1) { bit += 1; delta = 0; }
    else if (peek & 2) { delta = r.get(bit + 2, n2); bit += n2 + 2; }
    else if (peek & 4) { delta = r.get(bit + 3, n3); bit += n3 + 3; }
    else if (peek & 8) { delta = r.get(bit + 4, n4); bit += n4 + 4; }
    else {
      const uint32_t b = (uint32_t)(peek >> 4);
      bit += 7;
      if (b == 1) return -1;
      if (b == 0) {   // repeat record: 3 bits (bytes - 1), then the count - 18
        const uint32_t nb = (uint32_t)r.get(bit, 3) + 1u;
        bit += 3;
        uint64_t rep = r.get(bit, nb * 8u > 57u ? 57u : nb * 8u) + 18ull;
        bit += nb * 8u;
        if (rep > (uint64_t)(count - done)) return -1;
        for (uint64_t k = 0; k < rep; ++k) {
          if (dbl) start = (start + pd) & m;
          put_val(out, wb, done++, start);
        }
        continue;
      }
      const uint32_t w = (b + 1u) * 8u;
      if (wb == 8 && w > 45) delta = (r.get(bit, w - 32) << 32) | r.get(bit + (w - 32), 32);   // high part first
      else delta = r.get(bit, w);
      bit += w;
    }
    if (bit > in_len * 8 + 64) return -1;
    const uint64_t d = zigzag_dec(delta & m, m);
    if (dbl) { pd = (pd + d) & m; start = (start + pd) & m; }
    else start = (start + d) & m;
    put_val(out, wb, done++, start);
  }
  return (bit + 7) / 8;
}

__device__ __forceinline__ int64_t dec_stream(int type, uint32_t wb, const uint8_t *in, int64_t in_len, int64_t count, uint8_t *out) {
  switch (type) {
    case 2: return dec_rle_family(true, in, in_len, wb, count, out);
    case 3: return dec_pfor_family(2, in, in_len, wb, count, out);
    case 4: return dec_rle_family(false, in, in_len, wb, count, out);
    case 5: return dec_pfor_family(1, in, in_len, wb, count, out);
    case 6: return dec_pfor_family(0, in, in_len, wb, count, out);
    case 8: return dec_pfor_family(3, in, in_len, wb, count, out);
    default: return -1;
  }
}

// ---- the block walk shared by the survey and the rewrite ---------------------------------------------------------
// V::bytes(src_off, len): verbatim bytes; V::int_stream(at, end, meta, count): one integer stream; both in block order.
template <typename V>
__device__ int cs_walk(const uint8_t *s, uint32_t size, V &v) {
  const uint32_t header_size = g_rd(s, 4, 4), ncol = g_rd(s, 10, 2), rows = g_rd(s, 16, 4);
  if (header_size < 64u || (uint64_t)header_size + 12ull + 4ull * ncol > size || rows == 0) return XF_CORRUPT;
  const uint32_t ah = header_size;
  if (s[ah] != 0) return XF_CORRUPT;
  if (s[ah + 1] & 0x3) return XF_UNSUPPORTED;   // already transformed / compressed string data
  const uint32_t all_string_len = g_rd(s, ah + 2, 4), offsets_len = g_rd(s, ah + 6, 4), n_streams = g_rd(s, ah + 10, 2);
  if ((uint64_t)offsets_len + all_string_len > size - header_size) return XF_CORRUPT;
  const uint32_t str_begin = size - offsets_len - all_string_len, off_at = size - offsets_len;
  XMeta om{};
  if (n_streams > 0) {
    x_parse_meta(s, off_at, size, om);
    if (!om.ok || (om.attr & 0x3) || om.width > 4) return XF_UNSUPPORTED;
    if (om.type != 1) return XF_UNSUPPORTED;   // the stream-offsets stream itself: RAW (4+ streams could pick a codec: not seen from this writer)
    if (om.meta_len + om.width * n_streams != offsets_len) return XF_CORRUPT;
  }
  auto stream_end = [&](uint32_t k) { return g_rd(s, off_at + om.meta_len + k * om.width, om.width); };
  v.bytes(0, header_size + 12u + 4u * ncol);
  const uint32_t bitmap_bytes = (rows + 7u) >> 3;
  uint32_t pos = header_size + 12u + 4u * ncol, si = 0;
  int flags = XF_OK;
  for (uint32_t c = 0; c < ncol; ++c) {
    const uint32_t h = ah + 12u + 4u * c;
    const uint32_t type = s[h + 1], attrs = s[h + 2];
    uint32_t meta_len, n_s = 0, cnt[3] = {0, 0, 0};
    bool is_str[3] = {false, false, false};
    if (type == 0) { meta_len = ((attrs & 0x02) ? bitmap_bytes : 0u) + ((attrs & 0x08) ? bitmap_bytes : 0u); n_s = 1; cnt[0] = rows; }
    else if (type == 1) {
      meta_len = ((attrs & 0x02) ? bitmap_bytes : 0u) + ((attrs & 0x08) ? bitmap_bytes : 0u);
      is_str[0] = true; n_s = 1;
      if (!(attrs & 0x01)) { cnt[1] = rows; n_s = 2; }
    } else if (type == 2 || type == 3) {
      if ((uint64_t)pos + 10u > size) return XF_CORRUPT;
      const uint32_t distinct = g_rd(s, pos + 2, 4);
      const uint32_t ref_cnt = (s[pos + 1] & 0x4) ? g_rd(s, pos + 6, 4) : rows;
      meta_len = 10u + ((attrs & 0x08) ? bitmap_bytes : 0u);
      if (distinct > 0) {
        if (type == 2) { cnt[0] = distinct; cnt[1] = ref_cnt; n_s = 2; }
        else if (attrs & 0x01) { is_str[0] = true; cnt[1] = ref_cnt; n_s = 2; }
        else { is_str[0] = true; cnt[1] = distinct; cnt[2] = ref_cnt; n_s = 3; }
      }
    } else {
      return XF_UNSUPPORTED;   // semistruct columns
    }
    if ((uint64_t)pos + meta_len > size) return XF_CORRUPT;
    v.bytes(pos, meta_len);
    uint32_t at = pos + meta_len;
    for (uint32_t k = 0; k < n_s; ++k, ++si) {
      if (si >= n_streams) return XF_CORRUPT;
      const uint32_t end = stream_end(si);
      if (end < at || end > size) return XF_CORRUPT;
      if (is_str[k]) v.bytes(at, end - at);
      else {
        XMeta m;
        x_parse_meta(s, at, end, m);
        if (!m.ok || m.type == 7 || m.type == 0 || m.type > 8) return XF_UNSUPPORTED;
        if (m.type != 1) flags |= XF_NONRAW;
        else if (at + m.meta_len + (uint64_t)cnt[k] * m.width != end) return XF_CORRUPT;
        v.int_stream(at, end, m, cnt[k]);
      }
      v.stream_done(si);
      at = end;
    }
    pos = n_s == 0 ? pos + meta_len : at;
  }
  if (si != n_streams || pos != str_begin) return XF_CORRUPT;
  v.bytes(str_begin, all_string_len);
  v.offsets(off_at, om, n_streams);
  return flags;
}

struct SizeVisitor {
  uint64_t size = 0, last_end = 0;
  uint32_t jobs = 0, streams = 0;
  __device__ void bytes(uint32_t, uint32_t len) { size += len; }
  __device__ void int_stream(uint32_t, uint32_t, const XMeta &m, uint32_t count) {
    size += m.meta_len + (uint64_t)count * m.width;
    if (m.type != 1) ++jobs;
  }
  __device__ void stream_done(uint32_t) { last_end = size; ++streams; }
  __device__ void offsets(uint32_t, const XMeta &om, uint32_t n_streams) {
    if (n_streams == 0) return;
    // END offsets of the restated block: width by the last one (ObMicroBlockCSEncoder::store_stream_offsets_)
    const uint64_t last = last_end;
    const uint32_t w = last <= 0xffull ? 1u : (last <= 0xffffull ? 2u : 4u);
    size += om.meta_len + (uint64_t)w * n_streams;
  }
};

__global__ void __launch_bounds__(128) cs_survey_kernel(const uint8_t *image, const uint64_t *blk_off, const uint32_t *blk_size, int n_blocks,
                                                        uint32_t *out /* [n][4]: new size, jobs, flags, streams */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_blocks) return;
  const uint8_t *s = image + blk_off[i];
  const uint32_t size = blk_size[i];
  uint32_t new_size = size, jobs = 0, flags = XF_OK, streams = 0;
  if ((g_rd(s, 20, 1) & 0xffu) == CS_ENCODING_ROW_STORE) {
    SizeVisitor v;
    flags = (uint32_t)cs_walk(s, size, v);
    if (!(flags & (XF_UNSUPPORTED | XF_CORRUPT))) {
      if (v.size > 0x7fffffffull) flags |= XF_UNSUPPORTED;
      new_size = (uint32_t)v.size;
      jobs = v.jobs;
      streams = v.streams;
    }
  }
  out[4 * i] = new_size;
  out[4 * i + 1] = jobs;
  out[4 * i + 2] = flags;
  out[4 * i + 3] = streams;
}

struct WriteVisitor {
  const uint8_t *s;
  uint8_t *o;
  uint64_t src_base, dst_base;   // image offsets of the two blocks (jobs carry image offsets)
  StreamJob *jobs;
  uint32_t pos = 0, n_jobs = 0, last_end = 0;
  uint8_t *end_scratch;          // where stream k's new END offset is parked (4 bytes each) until offsets()
  __device__ void copy(uint32_t from, uint32_t len) {
    for (uint32_t k = 0; k < len; ++k) o[pos + k] = s[from + k];
    pos += len;
  }
  __device__ void bytes(uint32_t from, uint32_t len) { copy(from, len); }
  __device__ void int_stream(uint32_t at, uint32_t end, const XMeta &m, uint32_t count) {
    if (m.type == 1) { copy(at, end - at); return; }
    copy(at, m.meta_len);
    o[pos - m.meta_len + 2] = 1;   // ObIntegerStream::EncodingType::RAW
    StreamJob j{};
    j.src = src_base + at + m.meta_len;
    j.dst = dst_base + pos;
    j.enc_len = end - at - m.meta_len;
    j.count = count;
    j.type = m.type;
    j.width = (uint8_t)m.width;
    jobs[n_jobs++] = j;
    pos += count * m.width;
  }
  __device__ void stream_done(uint32_t k) {
    uint8_t *e = end_scratch + 4u * k;
    e[0] = (uint8_t)pos; e[1] = (uint8_t)(pos >> 8); e[2] = (uint8_t)(pos >> 16); e[3] = (uint8_t)(pos >> 24);
    last_end = pos;
  }
  __device__ void offsets(uint32_t off_at, const XMeta &om, uint32_t n_streams) {
    if (n_streams == 0) return;
    const uint32_t last = last_end, w = last <= 0xffu ? 1u : (last <= 0xffffu ? 2u : 4u);
    const uint32_t start = pos;
    copy(off_at, om.meta_len);
    o[start + 2] = 1;
    o[start + 3] = (uint8_t)(w == 1 ? 0 : (w == 2 ? 1 : 2));
    for (uint32_t k = 0; k < n_streams; ++k)
      for (uint32_t b = 0; b < w; ++b) o[pos++] = end_scratch[4u * k + b];
    const uint32_t header_size = g_rd(s, 4, 4), nl = pos - start;
    uint8_t *al = o + header_size + 6;   // ObAllColumnHeader::stream_offsets_length_
    al[0] = (uint8_t)nl; al[1] = (uint8_t)(nl >> 8); al[2] = (uint8_t)(nl >> 16); al[3] = (uint8_t)(nl >> 24);
  }
};

// end_scratch: 4 bytes per stream of the block, at scratch + 4 * stream_base[block] (stream counts prefix)
__global__ void __launch_bounds__(128) cs_rewrite_kernel(const uint8_t *image, const uint64_t *blk_off, const uint32_t *blk_size,
                                                         int n_blocks, uint8_t *new_image, const uint64_t *new_off, const uint32_t *new_size,
                                                         const uint64_t *job_base, StreamJob *jobs, uint8_t *scratch,
                                                         const uint64_t *scratch_base, XformRec *xf, int *status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_blocks) return;
  const uint8_t *s = image + blk_off[i];
  const uint32_t size = blk_size[i];
  uint8_t *o = new_image + new_off[i];
  xf[i] = XformRec{blk_off[i], 0};
  if ((g_rd(s, 20, 1) & 0xffu) != CS_ENCODING_ROW_STORE) {
    for (uint32_t k = 0; k < size; ++k) o[k] = s[k];
    return;
  }
  WriteVisitor v;
  v.s = s;
  v.o = o;
  v.src_base = blk_off[i];
  v.dst_base = new_off[i];
  v.jobs = jobs + job_base[i];
  v.end_scratch = scratch + scratch_base[i];
  const int flags = cs_walk(s, size, v);
  if ((flags & (XF_UNSUPPORTED | XF_CORRUPT)) || v.pos != new_size[i]) {
    atomicOr(status, flags | ((v.pos != new_size[i]) ? XF_CORRUPT : 0));
    return;
  }
  // string area: [size - offsets_len - all_string_len, ...) in both blocks
  const uint32_t hs = g_rd(s, 4, 4), all_string_len = g_rd(s, hs + 2, 4);
  const int64_t old_begin = (int64_t)size - g_rd(s, hs + 6, 4) - all_string_len;
  const int64_t new_begin = (int64_t)v.pos - g_rd(o, hs + 6, 4) - all_string_len;
  xf[i].str_delta = old_begin - new_begin;
}

__global__ void __launch_bounds__(128) cs_decode_kernel(const uint8_t *image, uint8_t *new_image, const StreamJob *jobs, int64_t n_jobs, int *status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_jobs) return;
  const StreamJob j = jobs[i];
  const int64_t used = dec_stream(j.type, j.width, image + j.src, j.enc_len, j.count, new_image + j.dst);
  if (used != (int64_t)j.enc_len) atomicOr(status, XF_CORRUPT);
}

}  // namespace obcs
