// Device-side micro-block decode primitives (sm_100a).
//
// A micro-block (one "page", ~16 KiB) is staged into shared memory by one TMA bulk copy; every
// primitive below reads the block image out of shared memory with 32-bit aligned loads and
// funnel shifts (the on-disk layout is byte/bit granular, see ob_format.h).
//
// Reference loops these replace (file:line in /root/reference/src/storage/blocksstable):
//   K1 bit unpack        encoding/ob_bit_stream.h:169-283
//   K2 RAW fixed load    encoding/ob_raw_decoder.cpp:128-173,530-591
//   K3 RAW var locate    encoding/ob_raw_decoder.cpp:29-125, ob_icolumn_decoder.h:463-527
//   K5 DICT gather       encoding/ob_dict_decoder.cpp:26-120,243-314
//   K7 RLE lookup        encoding/ob_rle_decoder.cpp:25-49,528-583
//   K8 base-diff         encoding/ob_integer_base_diff_decoder.cpp:25-82, .h:140-170
//   K9 ext (NULL) bits   encoding/ob_icolumn_decoder.h:259-318
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ob_format.h"

namespace obdev {

using namespace obf;

constexpr int kMaxUsedCols = 24;   // distinct columns referenced by one scan (filter U projection)
constexpr int kMaxNodes = 16;      // filter tree nodes
constexpr int kMaxParams = 48;     // filter constants
constexpr int kMaxProj = 24;       // projected columns
constexpr int kThreads = 256;      // threads per CTA (8 warps)
constexpr int kWarps = kThreads / 32;

// status bits written by kernels
enum : int { ST_UNSUPPORTED = 1, ST_OVERFLOW = 2, ST_CORRUPT = 4 };

// white-filter ops (sql::ObWhiteFilterOperatorType) + two host-resolved constants
enum : int { OP_EQ = 0, OP_LE, OP_LT, OP_GE, OP_GT, OP_NE, OP_BT, OP_IN, OP_NU, OP_NN,
             OP_FALSE = 100, OP_TRUE = 101 };
enum : int { NODE_WHITE = 0, NODE_AND = 1, NODE_OR = 2 };

// ---- unaligned loads from the shared-memory block image ---------------------------------------
// `s` is 16-byte aligned and has >= 16 readable bytes of slack after the block.
__device__ __forceinline__ uint32_t ld32(const uint8_t *s, uint32_t word_byte_off) {
  return *reinterpret_cast<const uint32_t *>(s + word_byte_off);
}

// n bytes (1..8) at byte offset off, zero extended
__device__ __forceinline__ uint64_t ld_bytes(const uint8_t *s, uint32_t off, uint32_t n) {
  const uint32_t a = off & ~3u;
  const uint32_t sh = (off & 3u) * 8u;
  const uint32_t w0 = ld32(s, a), w1 = ld32(s, a + 4);
  const uint32_t lo = __funnelshift_r(w0, w1, sh);
  if (n <= 4) return (uint64_t)(n == 4 ? lo : (lo & ((1u << (n * 8u)) - 1u)));
  const uint32_t w2 = ld32(s, a + 8);
  const uint32_t hi = __funnelshift_r(w1, w2, sh);
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return n == 8 ? v : (v & ((1ull << (n * 8u)) - 1ull));
}

// w bits (1..64) at absolute bit offset bit_off (LSB-first stream, ObBitStream::get)
__device__ __forceinline__ uint64_t ld_bits(const uint8_t *s, uint32_t bit_off, uint32_t w) {
  const uint32_t byte = bit_off >> 3;
  const uint32_t a = byte & ~3u;
  const uint32_t sh = ((byte & 3u) << 3) + (bit_off & 7u);  // 0..31
  const uint32_t w0 = ld32(s, a), w1 = ld32(s, a + 4);
  const uint32_t lo = __funnelshift_r(w0, w1, sh);
  if (w <= 32) return (uint64_t)(w == 32 ? lo : (lo & ((1u << w) - 1u)));
  const uint32_t w2 = ld32(s, a + 8);
  const uint32_t hi = __funnelshift_r(w1, w2, sh);
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return w == 64 ? v : (v & ((1ull << w) - 1ull));
}

// ---- per-column decode descriptor, built once per block per referenced column -----------------
struct ColDesc {
  uint8_t type;        // ColType
  uint8_t attr;        // ColAttr
  uint8_t obj_type;
  uint8_t sc;          // 1 signed int class, 2 unsigned int class, 5 string
  uint8_t elem_len;    // datum length of integer classes (8 / 4 / 1)
  uint8_t width;       // RAW / BASE_DIFF: bits (bit-packed) or bytes; DICT: row_ref_size
  uint8_t ext_bit;     // extend_value_bit if the column stores ext bits, else 0
  uint8_t ok;          // 0 => unsupported encoding / type for the device path
  uint32_t data_off;   // block offset of the column data (ext bits start here)
  uint32_t val_off;    // bit offset (bit-packed) or byte offset (fixed bytes) of value 0
  uint64_t base;       // BASE_DIFF base
  uint64_t int_mask;   // sign-extension mask (~INTEGER_MASK_TABLE[type_store_size] for ObIntTC)
  // dictionary (DICT / RLE)
  uint32_t dict_payload;   // block offset of dict payload (after the 9-byte meta header)
  uint32_t dict_var;       // block offset of var data (var dict)
  uint32_t dict_end;       // block offset one past the dict meta (last var cell ends here)
  uint32_t dict_count;
  uint16_t dict_data_size; // fixed: bytes per entry; var: index_byte
  uint8_t dict_fixed;
  uint8_t var_is_last;     // RAW var: LAST_VAR_FIELD
  // RLE
  uint32_t rle_count;
  uint32_t rle_row_ids;    // block offset
  uint32_t rle_refs;       // block offset
  uint8_t rle_row_id_byte, rle_ref_byte;
  // RAW var-length cells in the row data
  uint8_t var_in_row;      // 1 => cell lives in the row data
  uint8_t var_ext_in_row;  // ext bits inside each row at bit ext_index
  uint32_t var_header_off; // bytes of per-row ext bits (row_offset_)
  uint32_t var_k;          // index among the var columns
  uint32_t ext_index;
};

struct BlockView {
  const uint8_t *s;        // shared-memory image
  uint32_t size;
  uint32_t row_count;
  uint32_t header_size;
  uint32_t column_count;
  uint32_t meta_off;       // header + column headers
  uint32_t row_data_off;
  uint32_t row_index_off;  // start of the row index array (var row index)
  uint8_t row_index_byte, ext_bit;
  uint16_t var_col_cnt;
  uint8_t ok;
};

__device__ __forceinline__ void parse_block(const uint8_t *s, uint32_t size, BlockView &b) {
  b.s = s;
  b.size = size;
  const uint32_t w0 = ld32(s, 0);
  const int16_t magic = (int16_t)(w0 & 0xffff), version = (int16_t)(w0 >> 16);
  b.header_size = ld32(s, 4);
  b.column_count = ld32(s, 8) >> 16;
  b.row_count = ld32(s, 16);
  const uint32_t w5 = ld32(s, 20);
  const uint32_t row_store_type = w5 & 0xff, opt = (w5 >> 8) & 0xff;
  b.var_col_cnt = (uint16_t)(w5 >> 16);
  b.row_index_byte = opt & 7;
  b.ext_bit = (opt >> 3) & 7;
  b.row_data_off = ld32(s, 24);
  b.meta_off = b.header_size + 16u * b.column_count;
  b.ok = magic == MICRO_BLOCK_HEADER_MAGIC && version >= 1 && version <= 3 &&
         (row_store_type == ENCODING_ROW_STORE || row_store_type == SELECTIVE_ENCODING_ROW_STORE) &&
         b.meta_off <= size && b.row_data_off <= size && b.header_size >= 64;
  b.row_index_off = 0;
  if (b.ok && b.row_index_byte > 0) {
    const uint32_t need = (uint32_t)b.row_index_byte * (b.row_count + 1);
    if (need > size - b.row_data_off) b.ok = 0;
    else b.row_index_off = size - need;
  }
}

// Builds the descriptor of column `col`. Mirrors the decoder init of each codec.
__device__ __forceinline__ void build_col_desc(const BlockView &b, int col, ColDesc &d) {
  const uint8_t *s = b.s;
  d = ColDesc{};
  if (col < 0 || (uint32_t)col >= b.column_count) return;
  const uint32_t ch = b.header_size + 16u * (uint32_t)col;
  const uint32_t w0 = ld32(s, ch);
  if ((w0 & 0xff) != 0) return;  // version
  d.type = (uint8_t)((w0 >> 8) & 0xff);
  d.attr = (uint8_t)((w0 >> 16) & 0xff);
  d.obj_type = (uint8_t)(w0 >> 24);
  d.ext_index = ld32(s, ch + 4);
  const uint32_t offset = ld32(s, ch + 8), length = ld32(s, ch + 12);
  const int sc = store_class_of(d.obj_type);
  if (sc == 0) return;
  d.sc = (uint8_t)sc;
  d.elem_len = (uint8_t)datum_len_of(d.obj_type);
  d.int_mask = integer_mask_of(d.obj_type);
  const uint32_t meta = b.meta_off + offset;
  const bool has_ext = d.attr & ATTR_HAS_EXTEND_VALUE;
  const bool fixed = d.attr & ATTR_FIX_LENGTH, bp = d.attr & ATTR_BIT_PACKING;
  switch (d.type) {
    case COL_RAW: {
      if (fixed || bp) {
        if (meta > b.size) return;
        d.data_off = meta;
        d.ext_bit = has_ext ? b.ext_bit : 0;
        const uint32_t ext_bits = (uint32_t)d.ext_bit * b.row_count;
        d.width = (uint8_t)length;
        if (bp) {
          if (length == 0 || length > 64) return;
          d.val_off = meta * 8u + ext_bits;
        } else {
          if (length == 0 || (sc != 5 && length > 8)) return;
          d.val_off = meta + (ext_bits + 7u) / 8u;
          if (sc == 5) d.width = 0;  // fixed-length string: byte length kept in dict_data_size
          d.dict_data_size = (uint16_t)length;
          if (sc == 5 && length > 0xffff) return;
        }
      } else {
        if (sc != 5) return;  // integer var store is not produced for the supported shapes
        if (b.row_index_byte == 0) return;
        d.var_in_row = 1;
        d.var_ext_in_row = has_ext;
        d.ext_bit = has_ext ? b.ext_bit : 0;
        d.var_header_off = offset;
        d.var_k = length;
        d.var_is_last = (d.attr & ATTR_LAST_VAR_FIELD) != 0;
      }
      d.ok = 1;
      return;
    }
    case COL_INTEGER_BASE_DIFF: {
      if (sc == 5 || meta + length > b.size) return;
      const int ts = type_store_size(d.obj_type);
      d.width = s[meta + 1];
      if (d.width == 0 || d.width > 64) return;
      uint64_t base = ld_bytes(s, meta + 2, (uint32_t)ts);
      const uint64_t mask = ~low_mask((uint32_t)ts * 8u);
      if (sc == 1 && mask != 0 && (base & (mask >> 1))) base |= mask;
      d.base = base;
      d.data_off = meta + length;
      d.ext_bit = has_ext ? b.ext_bit : 0;
      const uint32_t ext_bits = (uint32_t)d.ext_bit * b.row_count;
      if (bp) d.val_off = d.data_off * 8u + ext_bits;
      else {
        if (d.width > 8) return;
        d.val_off = d.data_off + (ext_bits + 7u) / 8u;
      }
      d.ok = 1;
      return;
    }
    case COL_DICT:
    case COL_RLE: {
      uint32_t dm = meta;
      uint32_t dict_len = length;
      if (d.type == COL_RLE) {
        if (meta + 10 > b.size) return;
        const uint8_t a = s[meta + 1];
        d.rle_row_id_byte = a & 7;
        d.rle_ref_byte = (a >> 3) & 7;
        d.rle_count = (uint32_t)ld_bytes(s, meta + 2, 4);
        const uint32_t doff = (uint32_t)ld_bytes(s, meta + 6, 4);
        d.rle_row_ids = meta + 10;
        d.rle_refs = d.rle_row_ids + d.rle_count * d.rle_row_id_byte;
        if (d.rle_count == 0 || d.rle_row_id_byte == 0 || d.rle_ref_byte == 0 || doff > length) return;
        // the reference keeps count*row_id_byte in an int16 (ob_rle_decoder.h:193)
        if (d.rle_count * d.rle_row_id_byte > 32767u) return;
        dm = meta + doff;
        dict_len = length - doff;
      }
      if (dm + 9 > b.size || dm + dict_len > b.size) return;
      d.width = s[dm + 1];  // row_ref_size
      d.dict_count = (uint32_t)ld_bytes(s, dm + 2, 4);
      d.dict_data_size = (uint16_t)ld_bytes(s, dm + 6, 2);
      const uint8_t dattr = s[dm + 8];
      d.dict_fixed = dattr & DICT_FIX_LENGTH;
      d.dict_payload = dm + 9;
      d.dict_end = dm + dict_len;
      if (!d.dict_fixed) {
        if (d.dict_data_size != 1 && d.dict_data_size != 2 && d.dict_data_size != 4) return;
        d.dict_var = d.dict_payload + (d.dict_count ? d.dict_count - 1 : 0) * d.dict_data_size;
        if (sc != 5) return;  // var dict of integers does not occur
      } else if (sc != 5 && (d.dict_data_size == 0 || d.dict_data_size > 8)) {
        return;
      }
      if (d.type == COL_DICT) {
        d.data_off = meta + length;  // refs follow the dict meta
        if (bp) {
          if (d.width == 0 || d.width > 32) return;
          d.val_off = d.data_off * 8u;
        } else {
          if (d.width == 0 || d.width > 4) return;
          d.val_off = d.data_off;
        }
      }
      d.ok = 1;
      return;
    }
    default:
      return;  // CONST / STRING_DIFF / HEX / PREFIX / span columns: caller falls back
  }
}

// ---- row -> dictionary reference -------------------------------------------------------------
__device__ __forceinline__ uint32_t rle_ref_of(const uint8_t *s, const ColDesc &d, uint32_t row) {
  // upper_bound over the run starts, then refs[pos - 1]
  uint32_t lo = 0, hi = d.rle_count;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t v = (uint32_t)ld_bytes(s, d.rle_row_ids + mid * d.rle_row_id_byte, d.rle_row_id_byte);
    if (v <= row) lo = mid + 1; else hi = mid;
  }
  const uint32_t pos = lo > 0 ? lo - 1 : 0;
  return (uint32_t)ld_bytes(s, d.rle_refs + pos * d.rle_ref_byte, d.rle_ref_byte);
}

__device__ __forceinline__ uint32_t ref_of(const uint8_t *s, const ColDesc &d, uint32_t row) {
  if (d.type == COL_RLE) return rle_ref_of(s, d, row);
  if (d.attr & ATTR_BIT_PACKING) return (uint32_t)ld_bits(s, d.val_off + row * d.width, d.width);
  return (uint32_t)ld_bytes(s, d.val_off + row * d.width, d.width);
}

__device__ __forceinline__ uint64_t sign_fix(const ColDesc &d, uint64_t v) {
  // load_data_to_datum: if (mask && (v & (mask >> 1))) v |= mask   (ob_encoding_util.h:505-509)
  if (d.int_mask != 0 && (v & (d.int_mask >> 1))) v |= d.int_mask;
  return v;
}

__device__ __forceinline__ uint64_t dict_int(const uint8_t *s, const ColDesc &d, uint32_t ref) {
  return sign_fix(d, ld_bytes(s, d.dict_payload + ref * d.dict_data_size, d.dict_data_size));
}

// dictionary string cell -> (block offset, length)
__device__ __forceinline__ void dict_str(const uint8_t *s, const ColDesc &d, uint32_t ref,
                                         uint32_t &cell, uint32_t &len) {
  if (d.dict_fixed) {
    cell = d.dict_payload + ref * d.dict_data_size;
    len = d.dict_data_size;
    return;
  }
  const uint32_t ib = d.dict_data_size;
  const uint32_t off = ref == 0 ? 0u : (uint32_t)ld_bytes(s, d.dict_payload + (ref - 1) * ib, ib);
  cell = d.dict_var + off;
  len = ref == d.dict_count - 1 ? d.dict_end - cell
                                : (uint32_t)ld_bytes(s, d.dict_payload + ref * ib, ib) - off;
}

// ---- integer-class cell ------------------------------------------------------------------------
// Returns the 64-bit value image the reference would MEMCPY into the datum (low elem_len bytes
// significant); is_null set for NULL (and NOP) cells.
__device__ __forceinline__ uint64_t int_cell(const BlockView &b, const ColDesc &d, uint32_t row,
                                             bool &is_null) {
  const uint8_t *s = b.s;
  is_null = false;
  if (d.type == COL_DICT || d.type == COL_RLE) {
    const uint32_t ref = ref_of(s, d, row);
    if (ref >= d.dict_count) { is_null = true; return 0; }
    return dict_int(s, d, ref);
  }
  if (d.ext_bit) {
    if (ld_bits(s, d.data_off * 8u + row * d.ext_bit, d.ext_bit) != STORED_NOT_EXT) {
      is_null = true;
      return 0;
    }
  }
  uint64_t v;
  if (d.attr & ATTR_BIT_PACKING) v = ld_bits(s, d.val_off + row * d.width, d.width);
  else v = ld_bytes(s, d.val_off + row * d.width, d.width);
  if (d.type == COL_INTEGER_BASE_DIFF) return v + d.base;
  return (d.attr & ATTR_BIT_PACKING) ? v : sign_fix(d, v);
}

// value used for comparisons: sign-extended from the datum length for signed classes
__device__ __forceinline__ int64_t cmp_image(const ColDesc &d, uint64_t v) {
  if (d.elem_len == 4) return d.sc == 1 ? (int64_t)(int32_t)(uint32_t)v : (int64_t)(uint32_t)v;
  if (d.elem_len == 1) return (int64_t)(uint8_t)v;
  return (int64_t)v;
}

// ---- string-class cell -------------------------------------------------------------------------
__device__ __forceinline__ void str_cell(const BlockView &b, const ColDesc &d, uint32_t row,
                                         uint32_t &cell, uint32_t &len, bool &is_null) {
  const uint8_t *s = b.s;
  is_null = false;
  cell = 0;
  len = 0;
  if (d.type == COL_DICT || d.type == COL_RLE) {
    const uint32_t ref = ref_of(s, d, row);
    if (ref >= d.dict_count) { is_null = true; return; }
    dict_str(s, d, ref, cell, len);
    return;
  }
  if (!d.var_in_row) {  // RAW fixed-length string
    if (d.ext_bit && ld_bits(s, d.data_off * 8u + row * d.ext_bit, d.ext_bit) != STORED_NOT_EXT) {
      is_null = true;
      return;
    }
    len = d.dict_data_size;
    cell = d.val_off + row * len;
    return;
  }
  // RAW var-length: row = [ext bits][col_idx_byte][idx x (nvar-1)][cells]
  const uint32_t rib = b.row_index_byte;
  const uint32_t ro = (uint32_t)ld_bytes(s, b.row_index_off + row * rib, rib);
  const uint32_t re = (uint32_t)ld_bytes(s, b.row_index_off + (row + 1) * rib, rib);
  const uint32_t rowp = b.row_data_off + ro;
  const uint32_t row_len = re - ro;
  if (d.var_ext_in_row && ld_bits(s, rowp * 8u + d.ext_index, d.ext_bit) != STORED_NOT_EXT) {
    is_null = true;
    return;
  }
  if (b.var_col_cnt == 1) {
    cell = rowp + d.var_header_off;
    len = row_len - d.var_header_off;
    return;
  }
  const uint32_t ib = s[rowp + d.var_header_off];
  const uint32_t idx = rowp + d.var_header_off + 1;
  const uint32_t var = idx + ib * (b.var_col_cnt - 1u);
  const uint32_t col_off = d.var_k == 0 ? 0u : (uint32_t)ld_bytes(s, idx + (d.var_k - 1) * ib, ib);
  len = d.var_is_last ? row_len - col_off - (var - rowp)
                      : (uint32_t)ld_bytes(s, idx + d.var_k * ib, ib) - col_off;
  cell = var + col_off;
}

// memcmp-then-length order of a shared-memory cell against a global-memory constant
__device__ __forceinline__ int str_cmp(const uint8_t *s, uint32_t cell, uint32_t len,
                                       const uint8_t *c, uint32_t clen) {
  const uint32_t m = len < clen ? len : clen;
  for (uint32_t i = 0; i < m; ++i) {
    const int a = s[cell + i], bb = c[i];
    if (a != bb) return a < bb ? -1 : 1;
  }
  return len < clen ? -1 : (len > clen ? 1 : 0);
}

__device__ __forceinline__ bool cmp_to_bool(int op, int c) {
  switch (op) {
    case OP_EQ: return c == 0;
    case OP_LE: return c <= 0;
    case OP_LT: return c < 0;
    case OP_GE: return c >= 0;
    case OP_GT: return c > 0;
    case OP_NE: return c != 0;
    default: return false;
  }
}

}  // namespace obdev
