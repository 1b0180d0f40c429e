// PAX string codecs whose values do not exist in the block: HEX_PACKING, STRING_DIFF, STRING_PREFIX.
//
// Reference: ObHexStringDecoder / ObStringDiffDecoder / ObStringPrefixDecoder rebuild a cell into memory of the decoder's allocator
// every time it is read (encoding/ob_hex_string_decoder.cpp:33-127, ob_string_diff_decoder.cpp:34-120, ob_string_prefix_decoder.cpp:
// 30-110); white filters on such columns take the retro path (decode each row, compare). The device rebuilds them ONCE per page
// batch, at obgpu_batch_open, the way CS stream codecs are restated (stream_codecs.cuh): every block that has such a column is
// copied into the batch's own image with a MATERIALISED AREA appended per column,
//     [NULL bits, 1 per row, LSB first][END offset u32 x rows][the strings]
// and the column's decode plan becomes the plan of a CS STRING column (K_CSSTR: END offset per row) over that area, so count /
// project / the per-block entry points need nothing new. The codec's own header in the COPY is patched to say where the area is
// (version byte -> kMatMarker, offset_ field -> area offset); the decode reads the untouched original.
// Rebuilt strings are not part of the caller's image: a scan hands them out through obgpu_result_fetch_strings (a dense heap of the
// selected rows' bytes), the per-block entry point through obgpu_project_strings.
#pragma once

namespace obmat {

using namespace obdev;

constexpr uint8_t kMatMarker = 0xA5;
enum : uint32_t { MF_ANY = 1, MF_UNSUPPORTED = 2, MF_CORRUPT = 4 };

struct MatJob {
  uint64_t old_off;    // block in the image the batch was opened on
  uint64_t new_off;    // its copy
  uint32_t old_size;
  uint32_t col;
  uint32_t area_off;   // materialised area, relative to the copy's start
  uint32_t area_cap;
};

struct CodecHdr {
  uint32_t meta, length;      // block offset of the codec meta, its length (column header offset_ / length_)
  uint32_t pos_off, pos_len;  // the codec header's offset_ / length_ (row position of var cells, or the fixed cell length)
  uint32_t max_len;           // max_string_size / string_size
  uint32_t pos_field;         // block offset of the offset_ field inside the codec header
  uint8_t type, attr, ok;
};

// the three codec headers (ob_hex_string_encoder.h:139-152, ob_string_diff_encoder.h:27-104, ob_string_prefix_encoder.h:72-107)
__device__ __forceinline__ void read_codec_hdr(const uint8_t *s, const BlockView &b, int col, CodecHdr &h) {
  h = CodecHdr{};
  const uint32_t ch = b.header_size + 16u * (uint32_t)col;
  const uint32_t w0 = ld32(s, ch);
  h.type = (uint8_t)((w0 >> 8) & 0xff);
  h.attr = (uint8_t)((w0 >> 16) & 0xff);
  const uint8_t obj_type = (uint8_t)(w0 >> 24);
  if (h.type != COL_STRING_DIFF && h.type != COL_HEX_PACKING && h.type != COL_STRING_PREFIX) return;
  if (store_class_of(obj_type) != 5) return;
  const uint32_t offset = ld32(s, ch + 8), length = ld32(s, ch + 12);
  if (offset > b.size || b.meta_off > b.size - offset || length < 13u || b.meta_off + offset + length > b.size) return;
  h.meta = b.meta_off + offset;
  h.length = length;
  const uint8_t ver = s[h.meta];
  if (ver != 0 && ver != kMatMarker) return;
  if (h.type == COL_HEX_PACKING) {
    h.pos_field = h.meta + 1u;
    h.max_len = (uint32_t)ld_bytes(s, h.meta + 9u, 4);
  } else if (h.type == COL_STRING_DIFF) {
    h.pos_field = h.meta + 4u;
    h.max_len = (uint32_t)ld_bytes(s, h.meta + 2u, 2);
  } else {
    if (length < 15u) return;
    h.pos_field = h.meta + 2u;
    h.max_len = (uint32_t)ld_bytes(s, h.meta + 10u, 4);
  }
  h.pos_off = (uint32_t)ld_bytes(s, h.pos_field, 4);
  h.pos_len = (uint32_t)ld_bytes(s, h.pos_field + 4u, 4);
  h.ok = 1;
}

__device__ __forceinline__ uint32_t area_bytes(uint32_t rows, uint32_t max_len) {
  return (((rows + 31u) / 32u) * 4u + rows * 4u + rows * max_len + 16u + 15u) & ~15u;
}

// Does any block of the batch carry such a column? (one flag: the detailed survey and its copy back only run when it is set)
__global__ void __launch_bounds__(256) mat_probe_kernel(const uint8_t *image, const uint64_t *blk_off, const uint32_t *blk_size, int n_blocks,
                                                        uint32_t *flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_blocks) return;
  const uint8_t *s = image + blk_off[i];
  BlockView b;
  parse_block(s, blk_size[i], b);
  if (!b.ok || b.is_cs) return;
  for (uint32_t c = 0; c < b.column_count; ++c) {
    const uint32_t t = (ld32(s, b.header_size + 16u * c) >> 8) & 0xffu;
    if (t == COL_STRING_DIFF || t == COL_HEX_PACKING || t == COL_STRING_PREFIX) { atomicOr(flag, MF_ANY); return; }
  }
}

// out[4 i ..]: size of the copy (the block, 16-byte aligned, + the areas), jobs, flags, 0
__global__ void __launch_bounds__(128) mat_survey_kernel(const uint8_t *image, const uint64_t *blk_off, const uint32_t *blk_size, int n_blocks,
                                                         uint32_t *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_blocks) return;
  const uint8_t *s = image + blk_off[i];
  const uint32_t size = blk_size[i];
  BlockView b;
  parse_block(s, size, b);
  uint32_t total = (size + 15u) & ~15u, jobs = 0, flags = 0;
  if (b.ok && !b.is_cs) {
    for (uint32_t c = 0; c < b.column_count; ++c) {
      const uint32_t t = (ld32(s, b.header_size + 16u * c) >> 8) & 0xffu;
      if (t != COL_STRING_DIFF && t != COL_HEX_PACKING && t != COL_STRING_PREFIX) continue;
      CodecHdr h;
      read_codec_hdr(s, b, (int)c, h);
      if (!h.ok || s[h.meta] == kMatMarker) { flags |= MF_UNSUPPORTED; continue; }   // the index kernel leaves the column unsupported
      const uint64_t need = (uint64_t)area_bytes(b.row_count, h.max_len);
      if ((uint64_t)total + need > 0x7fffff00ull) { flags |= MF_UNSUPPORTED; continue; }
      total += (uint32_t)need;
      ++jobs;
      flags |= MF_ANY;
    }
  }
  out[4 * i] = jobs ? total : size;   // an untouched block keeps its exact size (a CS block finds its stream offsets from its end)
  out[4 * i + 1] = jobs;
  out[4 * i + 2] = flags;
  out[4 * i + 3] = 0;
}

// One warp per block: copy it, then lane 0 lays the areas out behind it, patches the codec headers of the copy and lists the jobs.
__global__ void __launch_bounds__(128) mat_rewrite_kernel(const uint8_t *image, const uint64_t *blk_off, const uint32_t *blk_size, int n_blocks,
                                                          uint8_t *new_image, const uint64_t *new_off, const uint32_t *new_size,
                                                          const uint64_t *job_base, MatJob *jobs) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n_blocks) return;
  const uint8_t *s = image + blk_off[i];
  uint8_t *d = new_image + new_off[i];
  const uint32_t size = blk_size[i], padded = (size + 15u) & ~15u;
  const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
  uint4 *d4 = reinterpret_cast<uint4 *>(d);
  for (uint32_t k = (uint32_t)lane; k < padded / 16u; k += 32u) d4[k] = s4[k];   // blocks are 16-byte aligned and padded in both images
  __syncwarp();
  if (lane != 0 || new_size[i] <= padded) return;   // nothing materialised in this block
  BlockView b;
  parse_block(s, size, b);
  uint32_t at = padded;
  uint64_t j = job_base[i];
  for (uint32_t c = 0; c < b.column_count; ++c) {
    const uint32_t t = (ld32(s, b.header_size + 16u * c) >> 8) & 0xffu;
    if (t != COL_STRING_DIFF && t != COL_HEX_PACKING && t != COL_STRING_PREFIX) continue;
    CodecHdr h;
    read_codec_hdr(s, b, (int)c, h);
    if (!h.ok) continue;
    const uint32_t need = area_bytes(b.row_count, h.max_len);
    if ((uint64_t)at + need > (uint64_t)new_size[i]) break;   // cannot happen: the survey walked the same columns
    d[h.meta] = kMatMarker;
    d[h.pos_field] = (uint8_t)at; d[h.pos_field + 1] = (uint8_t)(at >> 8); d[h.pos_field + 2] = (uint8_t)(at >> 16); d[h.pos_field + 3] = (uint8_t)(at >> 24);
    jobs[j++] = MatJob{blk_off[i], new_off[i], size, c, at, need};
    at += need;
  }
}

__device__ __forceinline__ uint8_t hex_at(const uint8_t *s, uint32_t map, uint32_t data, uint32_t pos) {
  return s[map + ((s[data + pos / 2u] >> (((pos + 1u) & 1u) * 4u)) & 0xfu)];
}

// encoded cell of a row: fixed store right behind the codec meta, var store through the row index like a RAW var cell
struct MatCol {
  CodecHdr h;
  ColDesc var;          // K_VARSTR plan of the var-stored cell (codec header's row position)
  uint32_t fix_data;    // fixed store: block offset of cell 0
  uint32_t ext_bit_off; // fixed store: bit offset of the ext values
  uint32_t hex_map;     // block offset of the alphabet (0: none)
  uint32_t descs, n_descs, common;       // STRING_DIFF
  uint32_t pfx_index, pfx_ib, pfx_count; // STRING_PREFIX
  bool fixed, has_ext;
};

__device__ __forceinline__ bool mat_col_init(const uint8_t *s, const BlockView &b, int col, MatCol &m) {
  read_codec_hdr(s, b, col, m.h);
  if (!m.h.ok) return false;
  m.fixed = (m.h.attr & ATTR_FIX_LENGTH) != 0;
  m.has_ext = (m.h.attr & ATTR_HAS_EXTEND_VALUE) != 0;
  const uint32_t col_data = m.h.meta + m.h.length;
  m.ext_bit_off = col_data * 8u;
  m.fix_data = col_data + (m.has_ext ? ((uint32_t)b.ext_bit * b.row_count + 7u) / 8u : 0u);
  m.var = ColDesc{};
  m.var.kind = K_VARSTR;
  m.var.sc = 5;
  m.var.ok = 1;
  m.var.var_ext_in_row = m.has_ext;
  m.var.ext_bit = m.has_ext ? b.ext_bit : 0;
  m.var.ext_index = ld32(s, b.header_size + 16u * (uint32_t)col + 4u);
  m.var.var_header_off = m.h.pos_off;
  m.var.var_k = m.h.pos_len;
  m.var.var_is_last = (m.h.attr & ATTR_LAST_VAR_FIELD) != 0;
  m.hex_map = m.descs = m.n_descs = m.common = m.pfx_index = m.pfx_ib = m.pfx_count = 0;
  if (m.h.type == COL_HEX_PACKING) {
    m.hex_map = m.h.meta + 13u;
  } else if (m.h.type == COL_STRING_DIFF) {
    const uint32_t hex_size = s[m.h.meta + 1u];
    m.n_descs = s[m.h.meta + 12u];
    m.descs = m.h.meta + 13u;
    m.hex_map = hex_size ? m.descs + m.n_descs : 0u;
    m.common = m.descs + m.n_descs + hex_size;
    if (m.common > m.h.meta + m.h.length) return false;
  } else {
    m.pfx_count = s[m.h.meta + 1u];
    m.pfx_ib = s[m.h.meta + 14u] & 3u;
    const uint32_t hex_size = (s[m.h.meta + 14u] >> 2) & 0x1fu;
    m.hex_map = hex_size ? m.h.meta + 15u : 0u;
    m.pfx_index = m.h.meta + 15u + hex_size;
    m.common = m.pfx_index + (m.pfx_count ? m.pfx_count - 1u : 0u) * m.pfx_ib;
    if (m.pfx_count == 0 || (m.pfx_count > 1 && m.pfx_ib != 1 && m.pfx_ib != 2) || m.common > m.h.meta + m.h.length) return false;
  }
  if (!m.fixed && b.row_index_byte == 0) return false;
  return true;
}

// (cell, cell length, NULL) of a row, then the length of the rebuilt string; false: the block is corrupt
__device__ __forceinline__ bool mat_row(const uint8_t *s, const BlockView &b, const MatCol &m, uint32_t row, uint32_t &cell, uint32_t &clen,
                                        bool &is_null, uint32_t &len) {
  is_null = false;
  len = 0;
  if (m.fixed) {
    if (m.has_ext && ld_bits32(s, m.ext_bit_off + row * b.ext_bit, b.ext_bit) != STORED_NOT_EXT) { is_null = true; return true; }
    clen = m.h.pos_len;
    cell = m.fix_data + row * clen;
  } else {
    str_cell(b, m.var, nullptr, row, cell, clen, is_null);
    if (is_null) return true;
  }
  if ((uint64_t)cell + clen > b.size) return false;
  if (m.h.type == COL_HEX_PACKING) {
    if (m.fixed) len = m.h.max_len;
    else {
      if (clen < 1u) return false;
      len = (clen - 1u) * 2u - s[cell];
    }
  } else if (m.h.type == COL_STRING_DIFF) {
    len = m.h.max_len;
  } else {
    if (clen < 3u) return false;
    const uint32_t odd = s[cell] >> 4, common = (uint32_t)s[cell + 1] | ((uint32_t)s[cell + 2] << 8);
    len = common + (m.hex_map ? (clen - 3u) * 2u - odd : clen - 3u);
  }
  return len <= m.h.max_len;
}

__device__ __forceinline__ bool mat_write(const uint8_t *s, const BlockView &b, const MatCol &m, uint32_t cell, uint32_t clen, uint32_t len,
                                          uint8_t *dst) {
  if (m.h.type == COL_HEX_PACKING) {
    const uint32_t data = m.fixed ? cell : cell + 1u;
    for (uint32_t i = 0; i < len; ++i) dst[i] = hex_at(s, m.hex_map, data, i);
    return true;
  }
  if (m.h.type == COL_STRING_DIFF) {
    uint32_t fpos = 0, cpos = 0, ppos = 0;   // position in the string, in the common bytes, in the row's part
    for (uint32_t i = 0; i < m.n_descs; ++i) {
      const uint32_t diff = s[m.descs + i] & 1u, cnt = s[m.descs + i] >> 1;
      if (fpos + cnt > len) return false;
      for (uint32_t k = 0; k < cnt; ++k, ++fpos) {
        if (!diff) dst[fpos] = s[m.common + cpos++];
        else { dst[fpos] = m.hex_map ? hex_at(s, m.hex_map, cell, ppos) : s[cell + ppos]; ++ppos; }
      }
    }
    return fpos == len && (m.hex_map ? (ppos + 1u) / 2u : ppos) <= clen;
  }
  const uint32_t ref = s[cell] & 0xfu, common = (uint32_t)s[cell + 1] | ((uint32_t)s[cell + 2] << 8);
  if (ref >= m.pfx_count) return false;
  const uint32_t poff = ref ? (uint32_t)ld_bytes(s, m.pfx_index + (ref - 1u) * m.pfx_ib, m.pfx_ib) : 0u;
  if ((uint64_t)m.common + poff + common > (uint64_t)m.h.meta + m.h.length) return false;
  for (uint32_t i = 0; i < common; ++i) dst[i] = s[m.common + poff + i];
  const uint32_t rest = len - common;
  if (m.hex_map) for (uint32_t i = 0; i < rest; ++i) dst[common + i] = hex_at(s, m.hex_map, cell + 3u, i);
  else for (uint32_t i = 0; i < rest; ++i) dst[common + i] = s[cell + 3u + i];
  return true;
}

// One warp per (block, column): lengths -> END offsets (a warp scan per 32 rows, the running total carried), NULL bits by ballot,
// then every lane rebuilds its rows' strings at their offsets.
__global__ void __launch_bounds__(128) mat_decode_kernel(const uint8_t *image, uint8_t *new_image, const MatJob *jobs, int64_t n_jobs, int *status) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n_jobs) return;
  const MatJob job = jobs[j];
  const uint8_t *s = image + job.old_off;
  BlockView b;
  parse_block(s, job.old_size, b);
  MatCol m;
  if (!b.ok || b.is_cs || !mat_col_init(s, b, (int)job.col, m)) {
    if (lane == 0) atomicOr(status, (int)MF_CORRUPT);
    return;
  }
  uint8_t *area = new_image + job.new_off + job.area_off;
  const uint32_t rows = b.row_count, nwords = (rows + 31u) / 32u;
  uint32_t *null_words = reinterpret_cast<uint32_t *>(area);
  uint32_t *ends = reinterpret_cast<uint32_t *>(area + nwords * 4u);
  uint8_t *bytes = area + nwords * 4u + rows * 4u;
  const uint32_t cap = job.area_cap - nwords * 4u - rows * 4u;
  uint32_t run = 0;
  bool bad = false;
  for (uint32_t r0 = 0; r0 < rows; r0 += 32u) {
    const uint32_t row = r0 + (uint32_t)lane;
    uint32_t cell = 0, clen = 0, len = 0;
    bool is_null = false;
    if (row < rows && !mat_row(s, b, m, row, cell, clen, is_null, len)) { bad = true; len = 0; }
    uint32_t inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    const uint32_t end = run + inc;
    const uint32_t nb = __ballot_sync(0xffffffffu, row < rows && is_null);
    if (lane == 0) null_words[r0 >> 5] = nb;
    if (row < rows) {
      if (end > cap) bad = true;
      else {
        ends[row] = end;
        if (len && !is_null && !bad && !mat_write(s, b, m, cell, clen, len, bytes + (end - len))) bad = true;
      }
    }
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, (int)MF_CORRUPT);
}

}  // namespace obmat
