/*
 * ob_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see ob_oracle.h).
 *
 * Restates, in plain C, the reference's decode/filter/projection algorithm for PAX
 * ("ENCODING_ROW_STORE") and CS ("CS_ENCODING_ROW_STORE", integer columns) micro-blocks and the
 * major-compaction merge.  Reference = /root/reference/src/storage/blocksstable
 * unless another root is given.  Nothing here is shared with the product code under
 * oceanbase_b200/: the layout constants are restated independently from the same spec.
 */
#include "ob_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- layout constants (ob_micro_block_header.h:97-153, ob_block_sstable_struct.h:201-264) ---- */
enum { T_RAW = 0, T_DICT = 1, T_RLE = 2, T_CONST = 3, T_BASE_DIFF = 4, T_STRING_DIFF = 5, T_HEX = 6, T_STRING_PREFIX = 7, T_COLUMN_EQUAL = 8, T_COLUMN_SUBSTR = 9, T_CS_INTEGER = 100, T_CS_STRING = 101, T_CS_INT_DICT = 102, T_CS_STR_DICT = 103 /* CS block: 100 + ObCSColumnHeader::Type */ };
enum { A_FIX = 0x1, A_EXT = 0x2, A_BITPACK = 0x4, A_LASTVAR = 0x8 };
enum { EXT_NOT = 0, EXT_NULL = 1, EXT_NOPE = 2 };
#define MAGIC 1005

static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint64_t rd_len(const uint8_t *p, int len) {
  uint64_t v = 0;
  memcpy(&v, p, (size_t)len);
  return v;
}

/* INTEGER_MASK_TABLE, encoding/ob_encoding_util.cpp:32-35 */
static const uint64_t INT_MASK[9] = {0x0ull, 0xffull, 0xffffull, 0xffffffull, 0xffffffffull,
                                     0xffffffffffull, 0xffffffffffffull, 0xffffffffffffffull,
                                     0xffffffffffffffffull};

/* store class / sizes, encoding/ob_encoding_util.h:59-195; datum map share/datum/ob_datum.cpp:22-92 */
static int obj_store_class(uint8_t t) { /* 1 ObIntSC, 2 ObUIntSC, 5 ObStringSC, 0 other */
  if ((t >= 1 && t <= 5) || t == 17 || t == 18 || t == 19 || t == 20) return 1;
  if ((t >= 6 && t <= 10) || t == 21) return 2;
  if (t == 22 || t == 23) return 5;
  return 0;
}
static int obj_is_int_tc(uint8_t t) { return t >= 1 && t <= 5; } /* ObIntTC only */
static int obj_signed_cmp(uint8_t t) { return obj_store_class(t) == 1; }
static int obj_type_size(uint8_t t) {
  switch (t) {
    case 1: case 6: case 21: return 1;
    case 2: case 7: return 2;
    case 3: case 4: case 8: case 9: case 19: return 4;
    case 5: case 10: case 17: case 18: case 20: return 8;
    default: return -1;
  }
}
static int obj_datum_len(uint8_t t) { return t == 21 ? 1 : (t == 19 ? 4 : 8); }
static uint64_t obj_integer_mask(uint8_t t) {
  /* ~INTEGER_MASK_TABLE[type_store_size] for ObIntTC, else 0 (ob_dict_decoder.cpp:198-203) */
  return obj_is_int_tc(t) ? ~INT_MASK[obj_type_size(t)] : 0;
}

/* =============================================================================================
 * Bit stream: LSB-first, byte granular (encoding/ob_bit_stream.h:169-187 get; :215-283 fast paths)
 * ============================================================================================= */
uint64_t ora_bs_get(const uint8_t *buf, int64_t offset, int64_t cnt) {
  int64_t index = offset, done = 0;
  uint64_t v = 0;
  while (done < cnt) {
    const int64_t word_off = index % 8;
    const int64_t bit = (cnt - done) < (8 - word_off) ? (cnt - done) : (8 - word_off);
    const uint64_t wv = (uint8_t)(((1u << bit) - 1u) & (buf[index / 8] >> word_off));
    v |= wv << done;
    done += bit;
    index += bit;
  }
  return v;
}

/* get<PACKED_LEN_LESS_THAN_10 / _26 / DEFAULT> selected by width (get_unpack_func, :285-296) */
uint64_t ora_bs_get_fast(const uint8_t *buf, int64_t offset, int64_t cnt, int64_t bs_len) {
  if (cnt < 10) {
    const int64_t byte_offset = offset >> 3, bit_offset = offset & 7;
    if (offset >= bs_len - 8 && bit_offset + cnt <= 8) {
      const uint8_t mask = (uint8_t)((1 << cnt) - 1);
      return (uint64_t)((uint8_t)(buf[byte_offset] >> bit_offset) & mask);
    } else {
      const uint16_t mask = (uint16_t)((1 << cnt) - 1);
      uint16_t v = rd16(buf + byte_offset);
      v = (uint16_t)(v >> bit_offset);
      return (uint64_t)(v & mask);
    }
  } else if (cnt < 26) {
    if (offset >= bs_len - 24) return ora_bs_get(buf, offset, cnt);
    const int64_t byte_offset = offset >> 3, bit_offset = offset & 7;
    const uint32_t mask = (uint32_t)((1u << cnt) - 1u);
    uint32_t v = rd32(buf + byte_offset);
    v >>= bit_offset;
    return (uint64_t)(v & mask);
  }
  return ora_bs_get(buf, offset, cnt);
}

/* ObBitStream::set (:147-167): OR bits into a zeroed buffer */
void ora_bs_set(uint8_t *buf, int64_t offset, int64_t cnt, uint64_t value) {
  int64_t index = offset, done = 0;
  while (done < cnt) {
    const int64_t word_off = index % 8;
    const int64_t bit = (cnt - done) < (8 - word_off) ? (cnt - done) : (8 - word_off);
    buf[index / 8] |= (uint8_t)((((1u << bit) - 1u) & (value >> done)) << word_off);
    done += bit;
    index += bit;
  }
}

/* =============================================================================================
 * Block init: ObIEncodeBlockReader::get_micro_metas (encoding/ob_micro_block_decoder.cpp:363-388),
 * ObMicroBlockDecoder::do_init (:1268-1322), ObMicroBlockHeader::is_valid (header.cpp:53-61)
 * ============================================================================================= */
/* ---- CS blocks: ObIntegerStreamMeta (de)serialization (ob_stream_encoding_struct.cpp:27-77) ------ */
typedef struct int_stream_meta {
  uint8_t version, attr, type, width_tag;
  uint64_t base, null_replaced;
  int width;       /* bytes */
  int64_t meta_len;
} int_stream_meta;

static int rd_vi64(const uint8_t *p, int64_t len, int64_t *pos, uint64_t *v) { /* serialization::decode_vi64 */
  uint64_t r = 0;
  int shift = 0;
  while (*pos < len) {
    const uint8_t c = p[(*pos)++];
    r |= (uint64_t)(c & 0x7f) << shift;
    if (!(c & 0x80)) { *v = r; return ORA_SUCCESS; }
    shift += 7;
    if (shift > 63) return ORA_INVALID_DATA;
  }
  return ORA_INVALID_DATA;
}

static int parse_int_stream_meta(const uint8_t *p, int64_t len, int_stream_meta *m) {
  static const int widths[4] = {1, 2, 4, 8};
  if (len < 4) return ORA_INVALID_DATA;
  int64_t pos = 4;
  m->version = p[0]; m->attr = p[1]; m->type = p[2]; m->width_tag = p[3];
  m->base = 0; m->null_replaced = 0;
  int ret = ORA_SUCCESS;
  if ((m->attr & 0x1) && (ret = rd_vi64(p, len, &pos, &m->base))) return ret;
  if ((m->attr & 0x2) && (ret = rd_vi64(p, len, &pos, &m->null_replaced))) return ret;
  if (m->attr & 0x4) return ORA_NOT_SUPPORTED;                 /* decimal int */
  if (m->version > 0) { if (pos >= len) return ORA_INVALID_DATA; ++pos; } /* pfor_packing_type_ */
  if (m->width_tag > 3) return ORA_NOT_SUPPORTED;
  if (m->type != 1) return ORA_NOT_SUPPORTED;                  /* only RAW streams: the other codecs need the transformer */
  m->width = widths[m->width_tag];
  m->meta_len = pos;
  return ORA_SUCCESS;
}

/* ObCSMicroBlockTransformer::init + decode_stream_offsets_ (ob_cs_micro_block_transformer.cpp:106-202) */
static int cs_block_init(ora_block *b) {
  const uint8_t *p = b->buf;
  const int64_t payload_len = b->size - b->header_size;
  if (payload_len < 12 + 4ll * b->column_count) return ORA_INVALID_DATA;
  const uint8_t *ah = p + b->header_size;          /* ObAllColumnHeader: version, attrs, u32 string len, u32 offsets len, u16 streams */
  if (ah[0] != 0 || (ah[1] & 0x1)) return ORA_INVALID_DATA;
  if (ah[1] & 0x2) return ORA_NOT_SUPPORTED;       /* compressed string data */
  const uint32_t all_string_len = rd32(ah + 2), offsets_len = rd32(ah + 6);
  b->cs_stream_count = rd16(ah + 10);
  b->cs_col_headers = ah + 12;
  b->cs_first_stream_begin = b->header_size + 12u + 4u * b->column_count;
  if ((int64_t)offsets_len + all_string_len > payload_len) return ORA_INVALID_DATA;
  b->cs_all_string_offset = (uint32_t)(b->size - offsets_len - all_string_len);
  if (b->cs_stream_count > 0) {
    const uint8_t *so = p + b->size - offsets_len;
    int_stream_meta m;
    int ret = parse_int_stream_meta(so, offsets_len, &m);
    if (ret) return ret;
    if ((m.attr & 0x1) || m.width > 4) return ORA_ERR_UNEXPECTED; /* "stream offsets encoding must has no base" */
    if (m.meta_len + (int64_t)m.width * b->cs_stream_count != offsets_len) return ORA_INVALID_DATA;
    b->cs_off_data = so + m.meta_len;
    b->cs_off_width = (uint8_t)m.width;
  }
  return ORA_SUCCESS;
}

int ora_block_init(ora_block *b, const void *buf, int64_t size) {
  if (!b || !buf || size < 64) return ORA_INVALID_ARGUMENT;
  const uint8_t *p = (const uint8_t *)buf;
  memset(b, 0, sizeof(*b));
  const int16_t magic = (int16_t)rd16(p), version = (int16_t)rd16(p + 2);
  b->buf = p;
  b->size = size;
  b->header_size = rd32(p + 4);
  b->column_count = rd16(p + 10);
  b->rowkey_column_count = rd16(p + 12);
  b->row_count = rd32(p + 16);
  const uint8_t row_store_type = p[20], opt = p[21];
  b->row_index_byte = opt & 7;
  b->extend_value_bit = (opt >> 3) & 7;
  b->var_column_count = rd16(p + 22);
  b->row_data_offset = rd32(p + 24);
  if (magic != MAGIC || version < 1 || version > 3 || b->column_count < b->rowkey_column_count)
    return ORA_INVALID_DATA;
  b->row_store_type = row_store_type;
  if (row_store_type == 3) return cs_block_init(b);
  if (row_store_type != 1 && row_store_type != 2) return ORA_NOT_SUPPORTED;
  if ((int64_t)b->header_size + 16ll * b->column_count > size || b->row_data_offset > size)
    return ORA_INVALID_ARGUMENT;
  b->col_headers = p + b->header_size;
  b->meta = b->col_headers + 16u * b->column_count;
  b->row_data = p + b->row_data_offset;
  b->row_data_len = size - b->row_data_offset;
  return ORA_SUCCESS;
}

/* crc32c, seed 0, no final xor == ob_crc64_sse42 (lib/checksum/ob_crc64.cpp:423-449) */
static uint32_t crc32c_tab[256];
static int crc32c_ready = 0;
static uint64_t crc64_sse42(uint64_t crc, const uint8_t *p, int64_t len) {
  if (!crc32c_ready) {
    for (uint32_t n = 0; n < 256; n++) {
      uint32_t c = n;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0x82f63b78u ^ (c >> 1) : c >> 1;
      crc32c_tab[n] = c;
    }
    crc32c_ready = 1;
  }
  uint32_t c = (uint32_t)crc;
  for (int64_t i = 0; i < len; i++) c = crc32c_tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
  return c;
}

static void fmt_i64(int64_t v, int16_t *cs) { for (int i = 0; i < 4; i++) *cs = (int16_t)(*cs ^ ((v >> (i * 16)) & 0xFFFF)); }
static void fmt_i32(int32_t v, int16_t *cs) { for (int i = 0; i < 2; i++) *cs = (int16_t)(*cs ^ ((v >> (i * 16)) & 0xFFFF)); }

uint64_t ora_crc64_sse42(uint64_t crc, const void *p, int64_t len) { return crc64_sse42(crc, (const uint8_t *)p, len); }

/* Column checksum of a run of integer-class cells: ObMicroBlockChecksumHelper::cal_column_checksum
 * (blocksstable/ob_micro_block_checksum_helper.cpp:127-146; the portable loop cal_column_checksum_normal and the sse4.2
 * loop :149-257 add the same value per cell) = wrapping int64 sum over the rows of ObDatum::checksum(0)
 * (share/datum/ob_datum.h:849-856): crc of the 4 pack_ bytes {len_:29, flag_:2, null_:1} (ob_datum.h:142-158), then of the
 * len_ value bytes; a NULL datum has len_ 0 and null_ 1. */
int64_t ora_column_checksum(const int64_t *vals, const uint8_t *nulls, int64_t n, int32_t datum_len) {
  int64_t sum = 0;
  for (int64_t r = 0; r < n; r++) {
    const int is_null = nulls && nulls[r] != 0;
    const uint32_t pack = is_null ? 0x80000000u : (uint32_t)datum_len;
    uint64_t c = crc64_sse42(0, (const uint8_t *)&pack, 4);
    if (!is_null && datum_len > 0) c = crc64_sse42(c, (const uint8_t *)&vals[r], datum_len);
    sum = (int64_t)((uint64_t)sum + c);
  }
  return sum;
}

/* Macro block: [ObMacroBlockCommonHeader 24 B][ObSSTableMacroBlockHeader][micro blocks back to back] (ob_macro_block.cpp:455-520).
 * ObMacroBlockCommonHeader::deserialize + check_integrity (ob_macro_block_common_header.cpp:54-69,118-143),
 * ObSSTableMacroBlockHeader::deserialize + FixedHeader::is_valid (ob_sstable_macro_block_header.cpp:118-140,216-273).
 * fields[] as the reference door lists them (oracle/ref_macro_wrap.cpp). The payload checksum ((int32_t)ob_crc64(payload) = crc32c)
 * and the running data checksum over the micro headers (ob_macro_block.cpp:301-303) are verified when verify != 0. */
int ora_macro_block_parse(const uint8_t *buf, int64_t len, int64_t *f, int32_t verify) {
  if (!buf || len < 24 + 128 || !f) return ORA_INVALID_ARGUMENT;
  for (int i = 0; i < 6; i++) f[i] = (int32_t)rd32(buf + 4 * i);
  if (!(f[0] == 24 && f[1] == 1 && f[2] == 1001)) return ORA_INVALID_DATA;          /* is_integrity */
  if (!(f[3] >= 0 && f[3] < 10)) return ORA_INVALID_DATA;                              /* is_valid: attr in [None, MaxMacroType) */
  const uint8_t *h = buf + 24;
  f[6] = rd32(h); f[7] = rd16(h + 4); f[8] = rd16(h + 6); f[9] = (int64_t)rd64(h + 8); f[10] = (int64_t)rd64(h + 16); f[11] = (int64_t)rd64(h + 24);
  for (int i = 0; i < 12; i++) f[12 + i] = (int32_t)rd32(h + 32 + 4 * i);
  f[24] = (int64_t)rd64(h + 80);
  const int64_t encrypt_id = (int64_t)rd64(h + 88), master_key_id = (int64_t)rd64(h + 96);
  f[25] = h[104];
  const int64_t type_cols = f[7] == 2 ? f[13] : f[12];
  const int64_t ck_off = 24 + 128 + type_cols * 8;
  if (ck_off + f[12] * 8 + 1 > len) return ORA_INVALID_DATA;
  f[27] = ck_off;
  f[26] = buf[ck_off + f[12] * 8];
  f[6] = 128 + type_cols * 8 + f[12] * 8 + 1;                                           /* header_size_ = get_serialize_size() */
  if (!(f[7] >= 1 && f[7] <= 2 && f[8] == 1007 && f[9] != 0 && f[10] >= 0 && f[13] >= 0 && f[14] >= 0 && f[15] > 0 && f[16] > 0 &&
        f[17] > 0 && f[18] > 0 && f[19] > 0 && f[24] >= 0 && encrypt_id >= 0 && master_key_id >= -1 && f[25] > 0))
    return ORA_INVALID_DATA;
  if (verify) {
    if (24 + f[4] > len || f[4] < 0) return ORA_INVALID_DATA;
    if ((int32_t)crc64_sse42(0, buf + 24, f[4]) != (int32_t)f[5]) return ORA_INVALID_DATA;
  }
  return ORA_SUCCESS;
}

/* The micro blocks of a macro block: micro_block_count_ headers walked from micro_block_data_offset_, each block
 * header_size_ + data_zlength_ bytes long (ob_micro_block_header.h:95-153); the walk must end exactly at
 * micro_block_data_offset_ + micro_block_data_size_. */
int ora_macro_block_micro_blocks(const uint8_t *buf, int64_t len, int64_t *offs, int64_t *sizes, int32_t cap, int32_t *n_out, int32_t verify) {
  int64_t f[28];
  int ret = ora_macro_block_parse(buf, len, f, verify);
  if (ret != ORA_SUCCESS) return ret;
  if (f[17] > cap) return ORA_BUF_NOT_ENOUGH;
  int64_t at = f[18];
  const int64_t end = f[18] + f[19];
  if (end > len) return ORA_INVALID_DATA;
  uint64_t ck = 0;
  int64_t rows = 0;
  for (int64_t i = 0; i < f[17]; i++) {
    if (at + 64 > end) return ORA_INVALID_DATA;
    const uint8_t *m = buf + at;
    if ((int16_t)rd16(m) != 1005) return ORA_INVALID_DATA;
    const int64_t sz = (int64_t)rd32(m + 4) + (int32_t)rd32(m + 44);
    if (sz < 64 || at + sz > end) return ORA_INVALID_DATA;
    offs[i] = at;
    sizes[i] = sz;
    rows += rd32(m + 16);
    ck = crc64_sse42(ck, m + 48, 8);
    at += sz;
  }
  if (at != end || rows != f[15]) return ORA_INVALID_DATA;
  if (verify && (int64_t)ck != f[24]) return ORA_INVALID_DATA;
  *n_out = (int32_t)f[17];
  return ORA_SUCCESS;
}

/* check_header_checksum / check_payload_checksum (ob_micro_block_header.cpp:236-285) */
int ora_block_verify_checksums(const ora_block *b) {
  const uint8_t *p = b->buf;
  int16_t cs = 0;
  cs = (int16_t)(cs ^ (int16_t)rd16(p));      /* magic */
  cs = (int16_t)(cs ^ (int16_t)rd16(p + 2));  /* version */
  cs = (int16_t)(cs ^ (int16_t)rd16(p + 8));  /* header_checksum */
  cs = (int16_t)(cs ^ (int16_t)p[20]);        /* row_store_type */
  cs = (int16_t)(cs ^ (int16_t)p[21]);        /* opt */
  fmt_i32(b->column_count, &cs);
  fmt_i32(b->rowkey_column_count, &cs);
  fmt_i32(rd16(p + 14) & 1, &cs);
  fmt_i32(rd16(p + 22), &cs);
  fmt_i64(b->header_size, &cs);
  fmt_i64(b->row_count, &cs);
  fmt_i64(b->row_data_offset, &cs);
  fmt_i64((int32_t)rd32(p + 28), &cs);
  fmt_i64((int64_t)rd64(p + 32), &cs);
  fmt_i64((int32_t)rd32(p + 40), &cs);
  fmt_i64((int32_t)rd32(p + 44), &cs);
  fmt_i64((int64_t)rd64(p + 48), &cs);
  if (cs != 0) return ORA_INVALID_DATA;
  const int32_t zlen = (int32_t)rd32(p + 44);
  if (zlen != b->size - b->header_size) return ORA_INVALID_DATA;
  if (crc64_sse42(0, p + b->header_size, zlen) != rd64(p + 48)) return ORA_INVALID_DATA;
  return ORA_SUCCESS;
}

/* ---- column header view --------------------------------------------------------------------- */
typedef struct col_hdr {
  int8_t type, attr;
  uint8_t obj_type;
  uint32_t ext_index, offset, length;
} col_hdr;

static int get_col(const ora_block *b, int32_t col, col_hdr *h) {
  if (col < 0 || col >= b->column_count) return ORA_INVALID_ARGUMENT;
  if (b->row_store_type == 3) { /* ObCSColumnHeader: version, type, attrs, obj_type */
    const uint8_t *c = b->cs_col_headers + 4 * col;
    if (c[0] != 0) return ORA_INVALID_DATA;
    h->type = (int8_t)(c[1] <= 3 ? 100 + c[1] : 127);
    h->attr = (int8_t)c[2];
    h->obj_type = c[3];
    h->ext_index = h->offset = h->length = 0;
    return ORA_SUCCESS;
  }
  const uint8_t *p = b->col_headers + 16 * col;
  if (p[0] != 0) return ORA_INVALID_DATA;
  h->type = (int8_t)p[1];
  h->attr = (int8_t)p[2];
  h->obj_type = p[3];
  h->ext_index = rd32(p + 4);
  h->offset = rd32(p + 8);
  h->length = rd32(p + 12);
  return ORA_SUCCESS;
}

/* load_data_to_datum for ObIntSC / ObUIntSC (encoding/ob_encoding_util.h:491-519) */
static void load_int(uint8_t obj_type, const uint8_t *cell, int64_t cell_len, ora_datum *d) {
  uint64_t value = rd_len(cell, (int)cell_len);
  const uint64_t mask = obj_integer_mask(obj_type);
  if (mask != 0 && (value & (mask >> 1))) value |= mask;
  d->len = (uint32_t)obj_datum_len(obj_type);
  d->ival = d->len >= 8 ? value : (value & INT_MASK[d->len]);
  d->ptr = 0;
  d->is_null = 0;
}
static void set_int(uint8_t obj_type, uint64_t value, ora_datum *d) { /* MEMCPY(datum.ptr_, &v, datum_len) */
  d->len = (uint32_t)obj_datum_len(obj_type);
  d->ival = d->len >= 8 ? value : (value & INT_MASK[d->len]);
  d->ptr = 0;
  d->is_null = 0;
}
static void set_null(ora_datum *d) { d->is_null = 1; d->len = 0; d->ptr = 0; d->ival = 0; }

/* ---- row index (encoding/ob_row_index.h:100-140) -------------------------------------------- */
static int locate_row(const ora_block *b, int64_t row, const uint8_t **data, int64_t *len) {
  if (b->row_index_byte > 0) {
    const int ib = b->row_index_byte;
    const uint8_t *idx = b->row_data + b->row_data_len - (int64_t)ib * (b->row_count + 1);
    const uint64_t off = rd_len(idx + row * ib, ib);
    *data = b->row_data + off;
    *len = (int64_t)(rd_len(idx + (row + 1) * ib, ib) - off);
  } else {
    const int64_t row_size = b->row_count ? b->row_data_len / b->row_count : 0;
    *data = b->row_data + row * row_size;
    *len = row_size;
  }
  return ORA_SUCCESS;
}

/* locate_cell_data for a var-stored column (encoding/ob_icolumn_decoder.h:463-527,
 * ob_raw_decoder.cpp:29-125): row = [ext bits][col_idx_byte][idx x (nvar-1)][cells] */
static void locate_var_cell(const ora_block *b, const col_hdr *h, const uint8_t *row, int64_t row_len,
                            const uint8_t **cell, int64_t *cell_len) {
  const int64_t header_off = h->offset; /* row_offset_: bytes of ext bits */
  const int64_t k = h->length;          /* index among var columns */
  const int last = (h->attr & A_LASTVAR) != 0;
  if (b->var_column_count == 1) {
    *cell = row + header_off;
    *cell_len = row_len - header_off;
    return;
  }
  const uint8_t *var = row + header_off;
  const int ib = (int8_t)*var;
  var += 1;
  const uint8_t *idx = var;
  var += (int64_t)ib * (b->var_column_count - 1);
  const int64_t col_off = k == 0 ? 0 : (int64_t)rd_len(idx + (k - 1) * ib, ib);
  if (last) *cell_len = row_len - col_off - (var - row);
  else *cell_len = (int64_t)rd_len(idx + k * ib, ib) - col_off;
  *cell = var + col_off;
}

/* =============================================================================================
 * Dict (encoding/ob_dict_decoder.cpp:191-314)
 * ============================================================================================= */
typedef struct dict_view {
  const uint8_t *hdr; /* ObDictMetaHeader */
  uint32_t count;
  uint8_t row_ref_size, attr;
  uint16_t data_size; /* or index_byte */
  const uint8_t *payload, *var_data;
  int64_t meta_length; /* bytes from hdr to end of dict payload */
} dict_view;

static void dict_init(dict_view *d, const uint8_t *hdr, int64_t meta_length) {
  d->hdr = hdr;
  d->row_ref_size = hdr[1];
  d->count = rd32(hdr + 2);
  d->data_size = rd16(hdr + 6);
  d->attr = hdr[8];
  d->payload = hdr + 9;
  d->meta_length = meta_length;
  d->var_data = (d->attr & 1) ? 0 : d->payload + (int64_t)(d->count ? d->count - 1 : 0) * d->data_size;
}

/* ObDictDecoder::decode(obj_type, datum, ref, meta_length) (:243-314) */
static int dict_decode(const dict_view *d, uint8_t obj_type, int64_t ref, ora_datum *out) {
  const int64_t count = d->count;
  if (ref >= count || count == 0) {
    if (ref == count || count == 0) { set_null(out); return ORA_SUCCESS; }
    if (ref == count + 1) { set_null(out); out->is_null = 2; return ORA_SUCCESS; } /* NOP */
    return ORA_ERR_UNEXPECTED;
  }
  const uint8_t *cell;
  int64_t cell_len;
  if (d->attr & 1) {
    cell = d->payload + ref * d->data_size;
    cell_len = d->data_size;
  } else {
    const int ib = d->data_size;
    const int64_t offset = ref == 0 ? 0 : (int64_t)rd_len(d->payload + (ref - 1) * ib, ib);
    cell = d->var_data + offset;
    if (ref == count - 1) cell_len = (d->hdr + d->meta_length) - cell;
    else cell_len = (int64_t)rd_len(d->payload + ref * ib, ib) - offset;
  }
  if (obj_store_class(obj_type) == 5) {
    out->ptr = cell;
    out->len = (uint32_t)cell_len;
    out->is_null = 0;
    out->ival = 0;
  } else {
    load_int(obj_type, cell, cell_len, out);
  }
  return ORA_SUCCESS;
}

/* =============================================================================================
 * Per-column decoder state
 * ============================================================================================= */
typedef struct col_dec {
  col_hdr h;
  int sc;
  const uint8_t *meta; /* block meta + h.offset */
  dict_view dict;      /* DICT / RLE */
  /* RLE (ob_rle_decoder.h:179-207) */
  uint32_t rle_count;
  int rle_row_id_byte, rle_ref_byte;
  const uint8_t *rle_row_ids, *rle_refs;
  /* BASE_DIFF (ob_integer_base_diff_decoder.h:140-170) */
  uint64_t base;
  uint8_t diff_len;
  /* CONST (ob_const_encoder.h:28-50, ob_const_decoder.cpp:60-137) */
  uint32_t const_count;       /* exception rows */
  uint8_t const_ref;          /* dict ref of the constant (no dict: 0 value, 1 NULL, 2 NOP) */
  int const_row_id_byte;
  const uint8_t *const_refs;  /* count x uint8 */
  const uint8_t *const_row_ids;
  const uint8_t *const_value; /* count == 0 && ref == 0: value image after the header */
  int64_t const_value_len;
  /* CS INTEGER column (cs_encoding/ob_integer_column_decoder.cpp:26-110) */
  const uint8_t *cs_null_bitmap, *cs_nop_bitmap; /* MSB-first per byte, or NULL */
  const uint8_t *cs_data;
  int cs_width;
  int cs_replace_null;
  uint64_t cs_base, cs_null_raw;                 /* null_replaced_value - base */
  /* CS INT_DICT column (cs_encoding/ob_int_dict_column_decoder.cpp:25-60): cs_data / cs_width / cs_base describe
   * the dictionary value stream */
  uint32_t cs_distinct;
  const uint8_t *cs_ref_data;
  int cs_ref_width;
  int cs_const_refs;            /* ObDictEncodingMeta::CONST_ENCODING_REF: ref stream = [exception cnt][const ref][row ids][refs] */
  uint32_t cs_ref_cnt;          /* elements of the ref stream (ref_row_cnt_) */
  /* CS STRING / STR_DICT (cs_encoding/ob_string_stream_decoder.cpp:79-82, ob_dict_column_decoder.cpp): bytes in the
   * block's all-string-data area, END offsets (cs_off_*) unless fixed length */
  const uint8_t *cs_str;           /* first byte of this column's string stream */
  int64_t cs_fixed_len;            /* >= 0: fixed length strings */
  const uint8_t *cs_off;           /* END offsets: one per row (STRING) or per dictionary entry (STR_DICT) */
  int cs_off_w;
  int cs_zero_len_null;
  /* HEX_PACKING / STRING_DIFF / STRING_PREFIX: the codec's own header (position of var cells in the row: offset / length) */
  col_hdr mat_pos;              /* {offset_, length_} of the codec header + the column's attr (LAST_VAR_FIELD) */
  const uint8_t *mat_hex;       /* alphabet (NULL: no hex packing) */
  uint32_t mat_max;             /* max_string_size / string_size */
  const uint8_t *mat_descs;     /* STRING_DIFF: DiffDesc[mat_desc_cnt] */
  int mat_desc_cnt;
  const uint8_t *mat_common;    /* STRING_DIFF: common bytes; STRING_PREFIX: prefix bytes */
  const uint8_t *mat_index;     /* STRING_PREFIX: (count - 1) start offsets */
  int mat_index_byte, mat_count;
  /* span columns (COLUMN_EQUAL / COLUMN_SUBSTR): the decoder of the referenced column, the exception meta (NULL: none) */
  const struct col_dec *span_ref;
  const uint8_t *span_exc;
  int64_t span_exc_len;
  const uint8_t *sub_rows;      /* COLUMN_SUBSTR: per-row [start_pos][length] behind the meta */
  int sub_pos_byte, sub_len_byte, sub_same_pos, sub_fix_len;
  uint32_t sub_start, sub_length;
} col_dec;

/* Strings that HEX_PACKING / STRING_DIFF / STRING_PREFIX rebuild do not exist in the block: the reference decodes them into memory
 * of the decoder's allocator (ctx.allocator_->alloc, ob_hex_string_decoder.cpp:87-93). The oracle's equivalent: a thread-local bump
 * arena that lives until ora_arena_reset(). */
typedef struct arena_chunk { struct arena_chunk *next; size_t used, cap; } arena_chunk;
static __thread arena_chunk *g_arena = 0;
static uint8_t *arena_alloc(size_t n) {
  n = (n + 15) & ~(size_t)15;
  if (!g_arena || g_arena->used + n > g_arena->cap) {
    const size_t cap = n > (1u << 20) ? n : (1u << 20);
    arena_chunk *c = (arena_chunk *)malloc(sizeof(arena_chunk) + cap);
    if (!c) return 0;
    c->next = g_arena; c->used = 0; c->cap = cap;
    g_arena = c;
  }
  uint8_t *p = (uint8_t *)(g_arena + 1) + g_arena->used;
  g_arena->used += n;
  return p;
}
void ora_arena_reset(void) {
  while (g_arena) { arena_chunk *n = g_arena->next; free(g_arena); g_arena = n; }
}
/* ObHexStringUnpacker::unpack (ob_hex_string_encoder.h:86-91): nibble `pos` of the packed bytes, even positions in the high half */
static inline uint8_t hex_at(const uint8_t *map, const uint8_t *data, int64_t pos) {
  return map[(data[pos / 2] >> (((pos + 1) % 2) * 4)) & 0xf];
}

/* ObStringStreamMeta, serialized (ob_stream_encoding_struct.cpp:255-283) */
typedef struct str_stream_meta { uint8_t attr; uint32_t uncompressed_len, fixed_len; int64_t meta_len; } str_stream_meta;
static int parse_str_stream_meta(const uint8_t *p, int64_t len, str_stream_meta *m) {
  if (len < 3 || p[0] != 0) return ORA_INVALID_DATA;
  int64_t pos = 2;
  uint64_t v = 0;
  m->attr = p[1];
  int ret = rd_vi64(p, len, &pos, &v);
  if (ret) return ret;
  m->uncompressed_len = (uint32_t)v;
  m->fixed_len = 0;
  if (m->attr & 0x2) { if ((ret = rd_vi64(p, len, &pos, &v))) return ret; m->fixed_len = (uint32_t)v; }
  m->meta_len = pos;
  return ORA_SUCCESS;
}

static uint32_t cs_stream_end(const ora_block *b, int32_t idx) {
  return (uint32_t)rd_len(b->cs_off_data + (int64_t)idx * b->cs_off_width, b->cs_off_width);
}

/* Const-encoded refs (ObConstEncodingRefDesc, cs_encoding/ob_dict_column_decoder.h:73-95; the per-row lookup of
 * extract_ref_and_null_count_, ob_dict_column_decoder.cpp:199-261): the exception row ids are ascending, a row
 * that is not listed has the const ref. */
static int cs_check_const_refs(const ora_block *b, const col_dec *c) {
  if (!c->cs_const_refs) return ORA_SUCCESS;
  if (c->cs_ref_cnt < 2) return ORA_INVALID_DATA;
  const uint64_t exc = rd_len(c->cs_ref_data, c->cs_ref_width);
  if (c->cs_ref_cnt != 2 + 2 * exc || exc > b->row_count) return ORA_INVALID_DATA;
  return ORA_SUCCESS;
}
static uint64_t cs_dict_ref(const col_dec *c, int64_t row) {
  const int w = c->cs_ref_width;
  if (!c->cs_const_refs) return rd_len(c->cs_ref_data + row * w, w);
  const int64_t exc = (int64_t)rd_len(c->cs_ref_data, w);
  const uint8_t *ids = c->cs_ref_data + 2 * w, *refs = ids + exc * w;
  int64_t lo = 0, hi = exc;              /* lower_bound over the exception row ids */
  while (lo < hi) {
    const int64_t mid = (lo + hi) / 2;
    if ((int64_t)rd_len(ids + mid * w, w) < row) lo = mid + 1; else hi = mid;
  }
  if (lo < exc && (int64_t)rd_len(ids + lo * w, w) == row) return rd_len(refs + lo * w, w);
  return rd_len(c->cs_ref_data + w, w);
}

/* Walks the column headers like ObCSMicroBlockTransformer::build_original_transform_desc_
 * (ob_cs_micro_block_transformer.cpp:216-380) to find column `col`'s meta position and first stream. */
static int cs_int_col_init(const ora_block *b, int32_t col, col_dec *c) {
  const int64_t bitmap_bytes = ((int64_t)b->row_count + 7) / 8;
  uint32_t pos = b->cs_first_stream_begin;   /* absolute offset where the current column's meta starts */
  int32_t stream_idx = -1;
  uint32_t str_at = b->cs_all_string_offset; /* running position inside the all-string-data area (stream order) */
  for (int32_t i = 0; i <= col; ++i) {
    const uint8_t *h = b->cs_col_headers + 4 * i;
    const uint8_t type = h[1], attrs = h[2];
    int32_t n_streams = 0;
    int64_t meta_len = 0;
    if (type == 0) {                         /* INTEGER */
      n_streams = 1;
      meta_len = ((attrs & 0x02) ? bitmap_bytes : 0) + ((attrs & 0x08) ? bitmap_bytes : 0);
    } else if (type == 1) {                  /* STRING: bytes stream (+ offsets stream when not fixed length) */
      n_streams = (attrs & 0x01) ? 1 : 2;
      meta_len = ((attrs & 0x02) ? bitmap_bytes : 0) + ((attrs & 0x08) ? bitmap_bytes : 0);
    } else if (type == 2 || type == 3) {     /* INT_DICT / STR_DICT */
      if ((int64_t)pos + 10 > b->size) return ORA_INVALID_DATA;
      const uint32_t distinct = rd32(b->buf + pos + 2);
      meta_len = 10 + ((attrs & 0x08) ? bitmap_bytes : 0);
      if (distinct == 0) n_streams = 0;
      else if (type == 2) n_streams = 2;
      else n_streams = (attrs & 0x01) ? 2 : 3;
    } else {
      return ORA_NOT_SUPPORTED;
    }
    if ((type == 1 || type == 3) && n_streams > 0) { /* this column owns one string stream: first stream after the meta */
      if (stream_idx + 1 >= b->cs_stream_count) return ORA_INVALID_DATA;
      const uint32_t send = cs_stream_end(b, stream_idx + 1);
      if ((int64_t)pos + meta_len > send || send > b->size) return ORA_INVALID_DATA;
      str_stream_meta sm;
      const int ret = parse_str_stream_meta(b->buf + pos + meta_len, (int64_t)send - pos - meta_len, &sm);
      if (ret) return ret;
      if (i == col) {
        c->cs_str = b->buf + str_at;
        c->cs_fixed_len = (sm.attr & 0x2) ? (int64_t)sm.fixed_len : -1;
        c->cs_zero_len_null = (sm.attr & 0x1) != 0;
        if ((int64_t)str_at + sm.uncompressed_len > b->size) return ORA_INVALID_DATA;
        const int32_t s_off = stream_idx + 2;   /* END offsets stream (variable length only) */
        if (type == 1) {
          if (attrs & 0x08) return ORA_NOT_SUPPORTED;
          c->cs_null_bitmap = (attrs & 0x02) ? b->buf + pos : 0;
          if (c->cs_fixed_len < 0) {
            int_stream_meta m;
            const uint32_t oend = cs_stream_end(b, s_off);
            int r2 = parse_int_stream_meta(b->buf + send, (int64_t)oend - send, &m);
            if (r2) return r2;
            if ((m.attr & 0x3) || send + m.meta_len + (int64_t)m.width * b->row_count != oend) return ORA_INVALID_DATA;
            c->cs_off = b->buf + send + m.meta_len;
            c->cs_off_w = m.width;
          }
          return ORA_SUCCESS;
        }
        /* STR_DICT: [dict meta][string stream][END offsets x distinct (variable)][refs x rows] */
        const uint8_t *dm = b->buf + pos;
        if (dm[0] != 0 || (attrs & 0x08)) return ORA_NOT_SUPPORTED;
        c->cs_distinct = rd32(dm + 2);
        c->cs_const_refs = (dm[1] & 0x4) != 0;
        c->cs_ref_cnt = c->cs_const_refs ? rd32(dm + 6) : b->row_count;
        uint32_t at = send;
        int32_t si = s_off;
        int_stream_meta m;
        if (c->cs_fixed_len < 0) {
          const uint32_t oend = cs_stream_end(b, si);
          int r2 = parse_int_stream_meta(b->buf + at, (int64_t)oend - at, &m);
          if (r2) return r2;
          if ((m.attr & 0x3) || at + m.meta_len + (int64_t)m.width * c->cs_distinct != oend) return ORA_INVALID_DATA;
          c->cs_off = b->buf + at + m.meta_len;
          c->cs_off_w = m.width;
          at = oend;
          ++si;
        }
        const uint32_t rend = cs_stream_end(b, si);
        int r3 = parse_int_stream_meta(b->buf + at, (int64_t)rend - at, &m);
        if (r3) return r3;
        if ((m.attr & 0x3) || at + m.meta_len + (int64_t)m.width * c->cs_ref_cnt != rend) return ORA_INVALID_DATA;
        c->cs_ref_data = b->buf + at + m.meta_len;
        c->cs_ref_width = m.width;
        return cs_check_const_refs(b, c);
      }
      str_at += sm.uncompressed_len;
    }
    if (i == col) {
      if (type == 3) { c->cs_distinct = 0; return ORA_SUCCESS; }   /* STR_DICT without streams: every row NULL */
      if (type == 2) { /* INT_DICT: [ObDictEncodingMeta 10 B][dict value stream][ref stream] */
        const uint8_t *dm = b->buf + pos;
        if (dm[0] != 0 || (attrs & 0x08)) return ORA_NOT_SUPPORTED; /* nop bitmap */
        c->cs_distinct = rd32(dm + 2);
        if (c->cs_distinct == 0) return ORA_SUCCESS;        /* every row NULL */
        c->cs_const_refs = (dm[1] & 0x4) != 0;
        c->cs_ref_cnt = c->cs_const_refs ? rd32(dm + 6) : b->row_count;
        if (stream_idx + 2 >= b->cs_stream_count) return ORA_INVALID_DATA;
        const uint32_t end0 = cs_stream_end(b, stream_idx + 1), end1 = cs_stream_end(b, stream_idx + 2);
        if ((int64_t)pos + meta_len > end0 || end0 > end1 || end1 > b->size) return ORA_INVALID_DATA;
        int_stream_meta m;
        int ret = parse_int_stream_meta(dm + meta_len, (int64_t)end0 - pos - meta_len, &m);
        if (ret) return ret;
        if (pos + meta_len + m.meta_len + (int64_t)m.width * c->cs_distinct != end0) return ORA_INVALID_DATA;
        c->cs_data = dm + meta_len + m.meta_len;
        c->cs_width = m.width;
        c->cs_base = (m.attr & 0x1) ? m.base : 0;
        if ((ret = parse_int_stream_meta(b->buf + end0, (int64_t)end1 - end0, &m))) return ret;
        if ((m.attr & 0x3) || end0 + m.meta_len + (int64_t)m.width * c->cs_ref_cnt != end1) return ORA_INVALID_DATA;
        c->cs_ref_data = b->buf + end0 + m.meta_len;
        c->cs_ref_width = m.width;
        return cs_check_const_refs(b, c);
      }
      if (type != 0) return ORA_NOT_SUPPORTED;
      if (stream_idx + 1 >= b->cs_stream_count) return ORA_INVALID_DATA;
      const uint32_t end = cs_stream_end(b, stream_idx + 1);
      const uint8_t *meta = b->buf + pos;
      if ((int64_t)pos + meta_len > end || end > b->size) return ORA_INVALID_DATA;
      c->cs_null_bitmap = (attrs & 0x02) ? meta : 0;
      c->cs_nop_bitmap = (attrs & 0x08) ? meta + ((attrs & 0x02) ? bitmap_bytes : 0) : 0;
      int_stream_meta m;
      const int ret = parse_int_stream_meta(meta + meta_len, (int64_t)end - pos - meta_len, &m);
      if (ret) return ret;
      if (pos + meta_len + m.meta_len + (int64_t)m.width * b->row_count != end) return ORA_INVALID_DATA;
      c->cs_data = meta + meta_len + m.meta_len;
      c->cs_width = m.width;
      c->cs_base = (m.attr & 0x1) ? m.base : 0;
      c->cs_replace_null = (m.attr & 0x2) != 0 && !c->cs_null_bitmap;
      c->cs_null_raw = m.null_replaced - c->cs_base;
      if (c->cs_width < 8) c->cs_null_raw &= INT_MASK[c->cs_width];
      return ORA_SUCCESS;
    }
    if (n_streams == 0) pos += (uint32_t)meta_len;
    else {
      stream_idx += n_streams;
      if (stream_idx >= b->cs_stream_count) return ORA_INVALID_DATA;
      pos = cs_stream_end(b, stream_idx);
    }
  }
  return ORA_ERR_UNEXPECTED;
}

static int col_dec_init(const ora_block *b, int32_t col, col_dec *c) {
  int ret = get_col(b, col, &c->h);
  if (ret) return ret;
  c->sc = obj_store_class(c->h.obj_type);
  if (c->sc == 0) return ORA_NOT_SUPPORTED;
  c->meta = b->meta + c->h.offset;
  switch (c->h.type) {
    case T_RAW: break;
    case T_DICT: dict_init(&c->dict, c->meta, c->h.length); break;
    case T_RLE: {
      const uint8_t *m = c->meta;
      c->rle_row_id_byte = m[1] & 7;
      c->rle_ref_byte = (m[1] >> 3) & 7;
      c->rle_count = rd32(m + 2);
      const uint32_t dict_off = rd32(m + 6);
      c->rle_row_ids = m + 10;
      /* ref_offset_ is an int16 in the reference (ob_rle_decoder.h:193) */
      c->rle_refs = c->rle_row_ids + (int16_t)(c->rle_count * c->rle_row_id_byte);
      dict_init(&c->dict, m + dict_off, (int64_t)c->h.length - dict_off);
      break;
    }
    case T_BASE_DIFF: {
      if (c->sc != 1 && c->sc != 2) return ORA_ERR_UNEXPECTED;
      const int store_size = obj_type_size(c->h.obj_type);
      c->diff_len = c->meta[1];
      c->base = rd_len(c->meta + 2, store_size);
      const uint64_t mask = ~INT_MASK[store_size];
      if (c->sc == 1 && mask != 0 && (c->base & (mask >> 1))) c->base |= mask;
      break;
    }
    case T_HEX: { /* ObHexStringHeader {version, offset u32, length u32, max_string_size u32} + alphabet (ob_hex_string_encoder.h:139-152) */
      if (c->sc != 5 || c->h.length < 13 || c->meta[0] != 0) return ORA_ERR_UNEXPECTED;
      c->mat_pos.offset = rd32(c->meta + 1); c->mat_pos.length = rd32(c->meta + 5); c->mat_pos.attr = c->h.attr;
      c->mat_max = rd32(c->meta + 9);
      c->mat_hex = c->meta + 13;
      break;
    }
    case T_STRING_DIFF: { /* ObStringDiffHeader (ob_string_diff_encoder.h:27-104) */
      if (c->sc != 5 || c->h.length < 13 || c->meta[0] != 0) return ORA_ERR_UNEXPECTED;
      const int hex_size = c->meta[1];
      c->mat_max = rd16(c->meta + 2);
      c->mat_pos.offset = rd32(c->meta + 4); c->mat_pos.length = rd32(c->meta + 8); c->mat_pos.attr = c->h.attr;
      c->mat_desc_cnt = c->meta[12];
      c->mat_descs = c->meta + 13;
      c->mat_hex = hex_size ? c->mat_descs + c->mat_desc_cnt : 0;
      c->mat_common = c->mat_descs + c->mat_desc_cnt + hex_size;
      break;
    }
    case T_STRING_PREFIX: { /* ObStringPrefixMetaHeader (ob_string_prefix_encoder.h:72-107) */
      if (c->sc != 5 || c->h.length < 15 || c->meta[0] != 0) return ORA_ERR_UNEXPECTED;
      c->mat_count = c->meta[1];
      c->mat_pos.offset = rd32(c->meta + 2); c->mat_pos.length = rd32(c->meta + 6); c->mat_pos.attr = c->h.attr;
      c->mat_max = rd32(c->meta + 10);
      c->mat_index_byte = c->meta[14] & 3;
      const int hex_size = (c->meta[14] >> 2) & 0x1f;
      c->mat_hex = hex_size ? c->meta + 15 : 0;
      c->mat_index = c->meta + 15 + hex_size;
      c->mat_common = c->mat_index + (int64_t)(c->mat_count > 0 ? c->mat_count - 1 : 0) * c->mat_index_byte;
      break;
    }
    case T_COLUMN_EQUAL:     /* ObColumnEqualMetaHeader {version u8, ref_col_idx u16} (ob_column_equal_encoder.h:25-39) */
    case T_COLUMN_SUBSTR: {  /* ObInterColSubStrMetaHeader {version, attr, start_pos u16, length u16, ref_col_idx u16} (ob_inter_column_substring_encoder.h:26-60) */
      const int hdr = c->h.type == T_COLUMN_EQUAL ? 3 : 8;
      if ((int64_t)c->h.length < hdr || c->meta[0] != 0) return ORA_ERR_UNEXPECTED;
      if (c->h.type == T_COLUMN_SUBSTR && c->sc != 5) return ORA_ERR_UNEXPECTED;
      const int32_t ref = rd16(c->meta + (c->h.type == T_COLUMN_EQUAL ? 1 : 6));
      col_hdr rh;
      if (ref == col || get_col(b, ref, &rh) || rh.type == T_COLUMN_EQUAL || rh.type == T_COLUMN_SUBSTR || rh.obj_type != c->h.obj_type)
        return ORA_ERR_UNEXPECTED;   /* the referenced column is an ordinary column of the same type */
      col_dec *rc = (col_dec *)arena_alloc(sizeof(col_dec));
      if (!rc) return ORA_ERR_UNEXPECTED;
      memset(rc, 0, sizeof(*rc));
      ret = col_dec_init(b, ref, rc);
      if (ret) return ret;
      c->span_ref = rc;
      c->span_exc = (int64_t)c->h.length > hdr ? c->meta + hdr : 0;   /* has_exc (ob_column_equal_decoder.h:60-61) */
      c->span_exc_len = (int64_t)c->h.length - hdr;
      if (c->h.type == T_COLUMN_SUBSTR) {
        const uint8_t attr = c->meta[1];
        c->sub_pos_byte = attr & 3; c->sub_len_byte = (attr >> 2) & 3;
        c->sub_same_pos = (attr >> 4) & 1; c->sub_fix_len = (attr >> 5) & 1;
        c->sub_start = rd16(c->meta + 2); c->sub_length = rd16(c->meta + 4);
        c->sub_rows = c->meta + c->h.length;
      }
      break;
    }
    case T_CS_INTEGER:
    case T_CS_STRING:
    case T_CS_INT_DICT:
    case T_CS_STR_DICT: return cs_int_col_init(b, col, c);
    case T_CONST: {
      const uint8_t *m = c->meta; /* version, count, const_ref, attr(row_id_byte:3), offset u16 */
      if (m[0] != 0) return ORA_ERR_UNEXPECTED;
      c->const_count = m[1];
      c->const_ref = m[2];
      c->const_row_id_byte = m[3] & 7;
      const uint32_t dict_off = rd16(m + 4);
      c->const_refs = m + 6;
      c->const_row_ids = m + 6 + c->const_count;
      c->const_value = m + 6;
      c->const_value_len = (int64_t)c->h.length - 6;
      if (c->const_count > 0) dict_init(&c->dict, m + dict_off, (int64_t)c->h.length - dict_off);
      break;
    }
    default: return ORA_NOT_SUPPORTED;
  }
  return ORA_SUCCESS;
}

/* ObIntegerArray<T>::lower_bound / upper_bound (encoding/ob_integer_array.h:38-51, ObIntArrayFuncTable :117-139) over an
 * ascending array of `byte`-wide unsigned integers: first index in [begin, end) whose element is >= key / > key */
int64_t ora_int_array_lower_bound(const void *array, int64_t byte, int64_t begin, int64_t end, int64_t key) {
  const uint8_t *a = (const uint8_t *)array;
  int64_t lo = begin, hi = end;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (rd_len(a + mid * byte, (int)byte) < (uint64_t)key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
int64_t ora_int_array_upper_bound(const void *array, int64_t byte, int64_t begin, int64_t end, int64_t key) {
  const uint8_t *a = (const uint8_t *)array;
  int64_t lo = begin, hi = end;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (rd_len(a + mid * byte, (int)byte) <= (uint64_t)key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* CONST: ref of a row = exception ref if the row is in the sorted exception list, else the
 * constant's ref (ObConstDecoder::decode, ob_const_decoder.cpp:93-121: lower_bound over row ids) */
static int64_t const_row_ref(const col_dec *c, int64_t row) {
  const int64_t lo = ora_int_array_lower_bound(c->const_row_ids, c->const_row_id_byte, 0, c->const_count, row);
  if (lo < c->const_count &&
      (int64_t)rd_len(c->const_row_ids + lo * c->const_row_id_byte, c->const_row_id_byte) == row)
    return c->const_refs[lo];
  return c->const_ref;
}

/* upper_bound over the RLE run-start array (ObIntArrayFuncTable::upper_bound_) */
static int64_t rle_upper_bound(const col_dec *c, int64_t row) {
  return ora_int_array_upper_bound(c->rle_row_ids, c->rle_row_id_byte, 0, c->rle_count, row);
}
static int64_t rle_ref_at(const col_dec *c, int64_t pos) {
  return (int64_t)rd_len(c->rle_refs + pos * c->rle_ref_byte, c->rle_ref_byte);
}

/* dict column: ref of a row (ob_dict_decoder.cpp:229-241) */
static int64_t dict_row_ref(const ora_block *b, const col_dec *c, int64_t row) {
  const uint8_t *col_data = c->meta + c->h.length;
  const int rs = c->dict.row_ref_size;
  if (c->h.attr & A_BITPACK) return (int64_t)ora_bs_get(col_data, row * rs, rs);
  (void)b;
  return (int64_t)rd_len(col_data + row * rs, rs);
}

/* ext value of a fixed/bit-packed column row; column data starts at col_data */
static uint64_t fixed_ext(const ora_block *b, const col_dec *c, const uint8_t *col_data, int64_t row) {
  if (!(c->h.attr & A_EXT)) return EXT_NOT;
  return ora_bs_get(col_data, row * b->extend_value_bit, b->extend_value_bit);
}

/* BitSet::get_ref (ob_encoding_bitset.h:68-71, count_before :131-149): -1 when bit `pos` of the 64-bit words is clear, else the
 * number of set bits below it */
int64_t ora_bitset_get_ref(const void *words, int64_t pos) {
  const uint8_t *buf = (const uint8_t *)words;
  if (!((rd64(buf + pos / 64 * 8) >> (pos % 64)) & 1)) return -1;
  int64_t r = 0;
  for (int64_t w = 0; w < pos / 64; ++w) r += __builtin_popcountll(rd64(buf + w * 8));
  return r + __builtin_popcountll(rd64(buf + pos / 64 * 8) & ((1ull << (pos % 64)) - 1));
}

/* Exception rows of a span column: ObBitMapMetaReader<StoreClass>::read / read_exc_cell (encoding/ob_encoding_bitset.h:574-760).
 *   ObBitMapMetaHeader {ext_offset u8, index_offset u8, data_offset u8, {bit_packing_len | fix_data_cnt | index_byte} u8}, then the
 *   BitSet over the block's rows (64-bit words); *ref = -1 when `row` is no exception, else its rank among the exception rows
 *   (BitSet::get_ref, :68-71) and *out = its value. `len` = bytes of the whole exception meta. */
static int bitmap_meta_read(const uint8_t *buf, int bit_packing, int64_t row, int64_t len, int sc, uint8_t obj_type, int64_t *ref,
                            ora_datum *out) {
  if (len <= 4) return ORA_INVALID_ARGUMENT;
  const uint8_t ext_offset = buf[0], index_offset = buf[1], data_offset = buf[2], u = buf[3];
  buf += 4;
  const int64_t r = ora_bitset_get_ref(buf, row);
  *ref = r;
  if (r < 0) return ORA_SUCCESS;
  uint64_t ext = EXT_NOT;
  if (index_offset > ext_offset) ext = ora_bs_get(buf + ext_offset, r * 2, 2);   /* has_ext_val */
  if (ext != EXT_NOT) { set_null(out); if (ext == EXT_NOPE) out->is_null = 2; return ORA_SUCCESS; }
  const int64_t data_len = len - 4 - data_offset;
  const uint8_t *data = buf + data_offset;
  if (sc != 5) {
    if (bit_packing) {   /* MEMCPY(datum.ptr_, &v, datum_len): no sign extension, a bit-packed image never has its top bit set */
      set_int(obj_type, ora_bs_get(data, r * u, u), out);
    } else {
      if (u == 0) return ORA_ERR_UNEXPECTED;
      const int64_t cell_len = data_len / u;   /* get_fix_data_size: fix_data_cnt_ = number of exceptions */
      if (cell_len < 0 || cell_len > 8) return ORA_ERR_UNEXPECTED;
      load_int(obj_type, data + r * cell_len, cell_len, out);
    }
    return ORA_SUCCESS;
  }
  int64_t offset = 0, cell_len = 0;
  if (data_offset == index_offset) {   /* !is_var_exc */
    if (u == 0) return ORA_ERR_UNEXPECTED;
    cell_len = data_len / u;
    offset = r * cell_len;
  } else {
    const int64_t exc_cnt = (data_offset - index_offset) / u + 1;   /* get_var_cnt */
    if (r != 0) offset = (int64_t)rd_len(buf + index_offset + (r - 1) * u, u);
    cell_len = (r == exc_cnt - 1 ? data_len : (int64_t)rd_len(buf + index_offset + r * u, u)) - offset;
  }
  out->ptr = data + offset; out->len = (uint32_t)cell_len; out->is_null = 0; out->ival = 0;
  return ORA_SUCCESS;
}

/* One cell: ObRawDecoder::decode (ob_raw_decoder.cpp:243-326), ObDictDecoder::decode (:214-241),
 * ObRLEDecoder::decode (ob_rle_decoder.cpp:25-49), ObIntegerBaseDiffDecoder::decode (:25-82) */
static int decode_cell(const ora_block *b, const col_dec *c, int64_t row, ora_datum *out) {
  switch (c->h.type) {
    case T_RAW: {
      const uint8_t *col_data = c->meta;
      const int fixed = (c->h.attr & A_FIX) != 0, bp = (c->h.attr & A_BITPACK) != 0;
      uint64_t ext = EXT_NOT;
      int64_t data_offset = 0;
      const uint8_t *row_data = 0;
      int64_t row_len = 0;
      if (c->h.attr & A_EXT) {
        if (fixed || bp) {
          data_offset = (int64_t)b->row_count * b->extend_value_bit;
          ext = ora_bs_get(col_data, row * b->extend_value_bit, b->extend_value_bit);
        } else {
          locate_row(b, row, &row_data, &row_len);
          ext = ora_bs_get(row_data, c->h.ext_index, b->extend_value_bit);
        }
      }
      if (ext != EXT_NOT) { set_null(out); if (ext == EXT_NOPE) out->is_null = 2; return ORA_SUCCESS; }
      if (bp) {
        set_int(c->h.obj_type, ora_bs_get(col_data, data_offset + row * c->h.length, c->h.length), out);
        return ORA_SUCCESS;
      }
      const uint8_t *cell;
      int64_t cell_len;
      if (fixed) {
        data_offset = (data_offset + 7) / 8;
        cell = col_data + data_offset + row * (int64_t)c->h.length;
        cell_len = c->h.length;
      } else {
        if (!row_data) locate_row(b, row, &row_data, &row_len);
        locate_var_cell(b, &c->h, row_data, row_len, &cell, &cell_len);
      }
      if (c->sc == 5) { out->ptr = cell; out->len = (uint32_t)cell_len; out->is_null = 0; out->ival = 0; }
      else load_int(c->h.obj_type, cell, cell_len, out);
      return ORA_SUCCESS;
    }
    case T_HEX:
    case T_STRING_DIFF:
    case T_STRING_PREFIX: {
      /* ext bits and cell location are shared by the three decoders (ob_hex_string_decoder.cpp:33-90, ob_string_diff_decoder.cpp:
       * 34-95, ob_string_prefix_decoder.cpp:30-70): fixed store = [ext bits][cells] after the meta, var store = a cell of the row */
      const int fixed = (c->h.attr & A_FIX) != 0;
      const uint8_t *col_data = c->meta + c->h.length;
      const uint8_t *row_data = 0, *cell = 0;
      int64_t row_len = 0, cell_len = 0, data_offset = 0;
      uint64_t ext = EXT_NOT;
      if (!fixed) locate_row(b, row, &row_data, &row_len);
      if (c->h.attr & A_EXT) {
        if (fixed) {
          data_offset = ((int64_t)b->row_count * b->extend_value_bit + 7) / 8;
          ext = ora_bs_get(col_data, row * b->extend_value_bit, b->extend_value_bit);
        } else {
          ext = ora_bs_get(row_data, c->h.ext_index, b->extend_value_bit);
        }
      }
      if (ext != EXT_NOT) { set_null(out); if (ext == EXT_NOPE) out->is_null = 2; return ORA_SUCCESS; }
      if (fixed) { cell = col_data + data_offset + row * (int64_t)c->mat_pos.length; cell_len = c->mat_pos.length; }
      else locate_var_cell(b, &c->mat_pos, row_data, row_len, &cell, &cell_len);
      uint8_t *buf = arena_alloc(c->mat_max > 128 ? c->mat_max : 128);
      if (!buf) return ORA_ERR_UNEXPECTED;
      int64_t len = 0;
      if (c->h.type == T_HEX) {
        len = c->mat_max;
        if (!fixed) { len = (cell_len - 1) * 2 - cell[0]; cell += 1; }   /* ObVarHexCellHeader::odd_ */
        for (int64_t i = 0; i < len; ++i) buf[i] = hex_at(c->mat_hex, cell, i);
      } else if (c->h.type == T_STRING_DIFF) {
        len = c->mat_max;
        int64_t fpos = 0, cpos = 0, ppos = 0;   /* position in the string, in the common bytes, in the row's part */
        for (int i = 0; i < c->mat_desc_cnt; ++i) {
          const int diff = c->mat_descs[i] & 1, cnt = c->mat_descs[i] >> 1;
          for (int k = 0; k < cnt; ++k, ++fpos) {
            if (!diff) buf[fpos] = c->mat_common[cpos++];
            else { buf[fpos] = c->mat_hex ? hex_at(c->mat_hex, cell, ppos) : cell[ppos]; ++ppos; }
          }
        }
      } else {
        if (cell_len < 3) return ORA_ERR_UNEXPECTED;
        const int ref = cell[0] & 0xf, odd = cell[0] >> 4;
        const int64_t common = rd16(cell + 1);
        const int64_t poff = ref ? (int64_t)rd_len(c->mat_index + (int64_t)(ref - 1) * c->mat_index_byte, c->mat_index_byte) : 0;
        if (ref >= c->mat_count || common > c->mat_max) return ORA_ERR_UNEXPECTED;
        memcpy(buf, c->mat_common + poff, (size_t)common);
        cell += 3; cell_len -= 3;
        if (c->mat_hex) {
          const int64_t rest = cell_len * 2 - odd;
          for (int64_t i = 0; i < rest; ++i) buf[common + i] = hex_at(c->mat_hex, cell, i);
          len = common + rest;
        } else {
          memcpy(buf + common, cell, (size_t)cell_len);
          len = common + cell_len;
        }
      }
      out->ptr = buf; out->len = (uint32_t)len; out->is_null = 0; out->ival = 0;
      return ORA_SUCCESS;
    }
    case T_COLUMN_EQUAL:
    case T_COLUMN_SUBSTR: {
      /* ObColumnEqualDecoder::decode (ob_column_equal_decoder.cpp:32-133), ObInterColSubStrDecoder::decode
       * (ob_inter_column_substring_decoder.cpp:32-93): an exception row reads its value from the exception meta, every other
       * row from the referenced column (whole value / the bytes [start_pos, start_pos + length) of it) */
      int64_t ref = -1;
      if (c->span_exc) {
        const int ret = bitmap_meta_read(c->span_exc, (c->h.attr & A_BITPACK) != 0, row, c->span_exc_len,
                                         c->h.type == T_COLUMN_SUBSTR ? 5 : c->sc, c->h.obj_type, &ref, out);
        if (ret) return ret;
      }
      if (ref != -1) return ORA_SUCCESS;
      const int ret = decode_cell(b, c->span_ref, row, out);
      if (ret || c->h.type == T_COLUMN_EQUAL || out->is_null) return ret;
      const uint8_t *cell = c->sub_rows + row * (int64_t)(c->sub_pos_byte + c->sub_len_byte);
      const int64_t start = c->sub_same_pos ? c->sub_start : (int64_t)rd_len(cell, c->sub_pos_byte);
      const int64_t len = c->sub_fix_len ? c->sub_length : (int64_t)rd_len(cell + c->sub_pos_byte, c->sub_len_byte);
      if (start + len > out->len) return ORA_ERR_UNEXPECTED;
      out->ptr = (const uint8_t *)out->ptr + start;
      out->len = (uint32_t)len;
      return ORA_SUCCESS;
    }
    case T_DICT:
      return dict_decode(&c->dict, c->h.obj_type, dict_row_ref(b, c, row), out);
    case T_RLE: {
      const int64_t pos = rle_upper_bound(c, row);
      return dict_decode(&c->dict, c->h.obj_type, rle_ref_at(c, pos - 1), out);
    }
    case T_CS_INTEGER: { /* ConvertUintToDatum_T (ob_integer_stream_decoder.cpp:37-350): value = raw + base */
      if (c->cs_null_bitmap && ((c->cs_null_bitmap[row / 8] >> (7 - row % 8)) & 1)) {
        set_null(out);
        if (c->cs_nop_bitmap && ((c->cs_nop_bitmap[row / 8] >> (7 - row % 8)) & 1)) out->is_null = 2;
        return ORA_SUCCESS;
      }
      const uint64_t raw = rd_len(c->cs_data + row * c->cs_width, c->cs_width);
      if (c->cs_replace_null && raw == c->cs_null_raw) { set_null(out); return ORA_SUCCESS; }
      set_int(c->h.obj_type, raw + c->cs_base, out);
      return ORA_SUCCESS;
    }
    case T_CS_STRING:
    case T_CS_STR_DICT: {
      int64_t idx = row;
      if (c->h.type == T_CS_STR_DICT) {
        if (c->cs_distinct == 0) { set_null(out); return ORA_SUCCESS; }
        idx = (int64_t)cs_dict_ref(c, row);
        if (idx == c->cs_distinct) { set_null(out); return ORA_SUCCESS; }
        if (idx > c->cs_distinct) return ORA_ERR_UNEXPECTED;
      } else if (c->cs_null_bitmap && ((c->cs_null_bitmap[row / 8] >> (7 - row % 8)) & 1)) {
        set_null(out);
        return ORA_SUCCESS;
      }
      int64_t start, len;
      if (c->cs_fixed_len >= 0) { start = idx * c->cs_fixed_len; len = c->cs_fixed_len; }
      else {
        start = idx ? (int64_t)rd_len(c->cs_off + (idx - 1) * c->cs_off_w, c->cs_off_w) : 0;
        len = (int64_t)rd_len(c->cs_off + idx * c->cs_off_w, c->cs_off_w) - start;
      }
      if (c->h.type == T_CS_STRING && c->cs_zero_len_null && len == 0) { set_null(out); return ORA_SUCCESS; }
      out->ptr = c->cs_str + start; out->len = (uint32_t)len; out->is_null = 0; out->ival = 0;
      return ORA_SUCCESS;
    }
    case T_CS_INT_DICT: { /* ObIntDictColumnDecoder::decode: ref == distinct_val_cnt is NULL, value = dict[ref] + base */
      if (c->cs_distinct == 0) { set_null(out); return ORA_SUCCESS; }
      const uint64_t ref = cs_dict_ref(c, row);
      if (ref == c->cs_distinct) { set_null(out); return ORA_SUCCESS; }
      if (ref > c->cs_distinct) return ORA_ERR_UNEXPECTED;
      set_int(c->h.obj_type, rd_len(c->cs_data + ref * c->cs_width, c->cs_width) + c->cs_base, out);
      return ORA_SUCCESS;
    }
    case T_CONST: {
      if (c->const_count == 0) { /* decode_without_dict (ob_const_decoder.cpp:25-58) */
        if (c->const_ref == 0) {
          if (c->sc == 5) { out->ptr = c->const_value; out->len = (uint32_t)c->const_value_len; out->is_null = 0; out->ival = 0; }
          else load_int(c->h.obj_type, c->const_value, c->const_value_len, out);
          return ORA_SUCCESS;
        }
        if (c->const_ref > 2) return ORA_ERR_UNEXPECTED;
        set_null(out);
        if (c->const_ref == 2) out->is_null = 2;
        return ORA_SUCCESS;
      }
      return dict_decode(&c->dict, c->h.obj_type, const_row_ref(c, row), out);
    }
    case T_BASE_DIFF: {
      const uint8_t *col_data = c->meta + c->h.length;
      int64_t data_offset = 0;
      uint64_t ext = EXT_NOT;
      if (c->h.attr & A_EXT) {
        data_offset = (int64_t)b->row_count * b->extend_value_bit;
        ext = ora_bs_get(col_data, row * b->extend_value_bit, b->extend_value_bit);
      }
      if (ext != EXT_NOT) { set_null(out); if (ext == EXT_NOPE) out->is_null = 2; return ORA_SUCCESS; }
      uint64_t v;
      if (c->h.attr & A_BITPACK) {
        v = ora_bs_get(col_data, data_offset + row * c->diff_len, c->diff_len);
      } else {
        data_offset = (data_offset + 7) / 8;
        v = rd_len(col_data + data_offset + row * c->diff_len, c->diff_len);
      }
      set_int(c->h.obj_type, v + c->base, out);
      return ORA_SUCCESS;
    }
    default: return ORA_NOT_SUPPORTED;
  }
}

int ora_decode_cell(const ora_block *b, int32_t col, int64_t row, ora_datum *out) {
  if (!b || !out || row < 0 || row >= b->row_count) return ORA_INVALID_ARGUMENT;
  col_dec c;
  const int ret = col_dec_init(b, col, &c);
  if (ret) return ret;
  return decode_cell(b, &c, row, out);
}

/* =============================================================================================
 * Dictionary surface (pushdown GROUP BY / black filter on a dictionary column):
 * ObDictDecoder::get_distinct_count / read_distinct / read_reference (encoding/ob_dict_decoder.cpp:1681-1830),
 * ObDictColumnDecoder (cs_encoding/ob_dict_column_decoder.cpp: get_distinct_count, read_distinct, read_reference).
 * A CONST column without exceptions (ObConstDecoder::get_distinct_count / read_distinct / read_reference,
 * ob_const_decoder.cpp:966-1019): one distinct value = the stored constant, every ref 0; an all-NULL / all-NOP one has
 * only the NULL group, which this surface numbers like everywhere else (ref == distinct count, here 0).
 * ============================================================================================= */
static int dict_col(const col_dec *c, int64_t *count) {
  switch (c->h.type) {
    case T_DICT: case T_RLE: *count = c->dict.count; return ORA_SUCCESS;
    case T_CONST: *count = c->const_count == 0 ? (c->const_ref == 0 ? 1 : 0) : c->dict.count; return ORA_SUCCESS;
    case T_CS_INT_DICT: case T_CS_STR_DICT: *count = c->cs_distinct; return ORA_SUCCESS;
    default: return ORA_NOT_SUPPORTED;
  }
}

int ora_dict_count(const ora_block *b, int32_t col, int64_t *count) {
  if (!b || !count) return ORA_INVALID_ARGUMENT;
  col_dec c;
  const int ret = col_dec_init(b, col, &c);
  return ret ? ret : dict_col(&c, count);
}

/* entry `ref` of the dictionary, decoded like a cell */
int ora_dict_entry(const ora_block *b, int32_t col, int64_t ref, ora_datum *out) {
  if (!b || !out) return ORA_INVALID_ARGUMENT;
  col_dec c;
  int64_t count;
  int ret = col_dec_init(b, col, &c);
  if (!ret) ret = dict_col(&c, &count);
  if (ret) return ret;
  if (ref < 0 || ref >= count) return ORA_INVALID_ARGUMENT;
  if (c.h.type == T_CONST && c.const_count == 0) return decode_cell(b, &c, 0, out);
  if (c.h.type == T_CS_INT_DICT) {
    set_int(c.h.obj_type, rd_len(c.cs_data + ref * c.cs_width, c.cs_width) + c.cs_base, out);
    return ORA_SUCCESS;
  }
  if (c.h.type == T_CS_STR_DICT) {
    int64_t start, len;
    if (c.cs_fixed_len >= 0) { start = ref * c.cs_fixed_len; len = c.cs_fixed_len; }
    else {
      start = ref ? (int64_t)rd_len(c.cs_off + (ref - 1) * c.cs_off_w, c.cs_off_w) : 0;
      len = (int64_t)rd_len(c.cs_off + ref * c.cs_off_w, c.cs_off_w) - start;
    }
    out->ptr = c.cs_str + start; out->len = (uint32_t)len; out->is_null = 0; out->ival = 0;
    return ORA_SUCCESS;
  }
  return dict_decode(&c.dict, c.h.obj_type, ref, out);
}

/* ref of every listed row; NULL and NOP rows report the distinct count (read_reference's NULL group) */
int ora_dict_refs(const ora_block *b, int32_t col, const int32_t *row_ids, int64_t row_cap, uint32_t *refs) {
  if (!b || !row_ids || !refs) return ORA_INVALID_ARGUMENT;
  col_dec c;
  int64_t count;
  int ret = col_dec_init(b, col, &c);
  if (!ret) ret = dict_col(&c, &count);
  if (ret) return ret;
  for (int64_t i = 0; i < row_cap; ++i) {
    const int64_t row = row_ids[i];
    if (row < 0 || row >= b->row_count) return ORA_INVALID_ARGUMENT;
    int64_t ref;
    switch (c.h.type) {
      case T_DICT: ref = dict_row_ref(b, &c, row); break;
      case T_RLE: ref = rle_ref_at(&c, rle_upper_bound(&c, row) - 1); break;
      case T_CONST: ref = c.const_count == 0 ? 0 : const_row_ref(&c, row); break;
      default: ref = count ? (int64_t)cs_dict_ref(&c, row) : 0; break;
    }
    if (ref > count + 1) return ORA_ERR_UNEXPECTED;
    refs[i] = (uint32_t)(ref < count ? ref : count);
  }
  return ORA_SUCCESS;
}

/* =============================================================================================
 * Batch projection: ObMicroBlockDecoder::get_rows -> decode_vector
 * (encoding/ob_micro_block_decoder.cpp:2473-2544; ob_raw_decoder.cpp:530-701;
 *  ob_dict_decoder.cpp:473-541; ob_rle_decoder.cpp:528-583 monotone cursor)
 * ============================================================================================= */
static inline void bitvec_set(uint64_t *words, int64_t idx) { words[idx / 64] |= 1ull << (idx % 64); }

/* RLE monotone cursor (extract_ref_and_null_count): refs for ascending row ids with a minimum of
 * binary searches */
typedef struct rle_cursor { int64_t pos, next_row; int64_t cur_ref; } rle_cursor;
static void rle_cursor_init(const col_dec *c, rle_cursor *k) {
  k->pos = 0;
  k->next_row = (int64_t)rd_len(c->rle_row_ids, c->rle_row_id_byte);
  k->cur_ref = rle_ref_at(c, 0);
}
static int64_t rle_cursor_ref(const col_dec *c, rle_cursor *k, int64_t row) {
  const int64_t n = c->rle_count;
  if (k->pos == n || row < k->next_row) {
  } else if (row == k->next_row) {
    ++k->pos;
    if (k->pos < n) k->next_row = (int64_t)rd_len(c->rle_row_ids + k->pos * c->rle_row_id_byte, c->rle_row_id_byte);
    k->cur_ref = rle_ref_at(c, k->pos - 1);
  } else {
    k->pos = rle_upper_bound(c, row);
    if (k->pos < n) k->next_row = (int64_t)rd_len(c->rle_row_ids + k->pos * c->rle_row_id_byte, c->rle_row_id_byte);
    k->cur_ref = rle_ref_at(c, k->pos - 1);
  }
  return k->cur_ref;
}

static int batch_decode(const ora_block *b, const col_dec *c, const int32_t *row_ids, int64_t row_cap,
                        ora_datum *datums) {
  if (c->h.type == T_RLE && row_cap > 0) {
    /* rows ascending (forward scan) */
    rle_cursor k;
    rle_cursor_init(c, &k);
    for (int64_t i = 0; i < row_cap; ++i) {
      const int ret = dict_decode(&c->dict, c->h.obj_type, rle_cursor_ref(c, &k, row_ids[i]), &datums[i]);
      if (ret) return ret;
    }
    return ORA_SUCCESS;
  }
  if (c->h.type == T_RAW && (c->h.attr & A_BITPACK)) {
    /* decode_vector_bitpacked: fast unpack paths by width */
    const uint8_t *col_data = c->meta;
    const int64_t bs_len = (int64_t)c->h.length * b->row_count;
    const int64_t data_offset = (c->h.attr & A_EXT) ? (int64_t)b->row_count * b->extend_value_bit : 0;
    for (int64_t i = 0; i < row_cap; ++i) {
      const int64_t row = row_ids[i];
      if ((c->h.attr & A_EXT) && fixed_ext(b, c, col_data, row) != EXT_NOT) { set_null(&datums[i]); continue; }
      set_int(c->h.obj_type, ora_bs_get_fast(col_data, data_offset + row * c->h.length, c->h.length, bs_len), &datums[i]);
    }
    return ORA_SUCCESS;
  }
  for (int64_t i = 0; i < row_cap; ++i) {
    const int ret = decode_cell(b, c, row_ids[i], &datums[i]);
    if (ret) return ret;
  }
  return ORA_SUCCESS;
}

#define ORA_MAX_BATCH 4096

int ora_get_rows_fixed(const ora_block *b, int32_t col, const int32_t *row_ids, int64_t row_cap,
                       int64_t vec_offset, void *data, int32_t elem_len, uint64_t *nulls,
                       int32_t *has_null) {
  if (!b || !row_ids || !data || row_cap < 0) return ORA_INVALID_ARGUMENT;
  col_dec c;
  int ret = col_dec_init(b, col, &c);
  if (ret) return ret;
  if (c.sc == 5 || elem_len != obj_datum_len(c.h.obj_type)) return ORA_INVALID_ARGUMENT;
  ora_datum tmp[256];
  for (int64_t done = 0; done < row_cap; done += 256) {
    const int64_t n = row_cap - done < 256 ? row_cap - done : 256;
    if ((ret = batch_decode(b, &c, row_ids + done, n, tmp))) return ret;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t at = vec_offset + done + i;
      if (tmp[i].is_null) { /* payload slot left unwritten (ob_vector_decode_util.h:726-734) */
        if (nulls) bitvec_set(nulls, at);
        if (has_null) *has_null = 1;
      } else {
        memcpy((uint8_t *)data + at * elem_len, &tmp[i].ival, (size_t)elem_len);
      }
    }
  }
  return ORA_SUCCESS;
}

int ora_get_rows_discrete(const ora_block *b, int32_t col, const int32_t *row_ids, int64_t row_cap,
                          int64_t vec_offset, const uint8_t **ptrs, int32_t *lens, uint64_t *nulls,
                          int32_t *has_null) {
  if (!b || !row_ids || !ptrs || !lens || row_cap < 0) return ORA_INVALID_ARGUMENT;
  col_dec c;
  int ret = col_dec_init(b, col, &c);
  if (ret) return ret;
  if (c.sc != 5) return ORA_INVALID_ARGUMENT;
  ora_datum tmp[256];
  for (int64_t done = 0; done < row_cap; done += 256) {
    const int64_t n = row_cap - done < 256 ? row_cap - done : 256;
    if ((ret = batch_decode(b, &c, row_ids + done, n, tmp))) return ret;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t at = vec_offset + done + i;
      if (tmp[i].is_null) {
        if (nulls) bitvec_set(nulls, at);
        if (has_null) *has_null = 1;
      } else {
        ptrs[at] = tmp[i].ptr;
        lens[at] = (int32_t)tmp[i].len;
      }
    }
  }
  return ORA_SUCCESS;
}

/* =============================================================================================
 * White filters.  NULL never matches a comparison; NE excludes NULL; only NU / NN see NULL
 * (ob_raw_decoder.cpp:703-706, ob_dict_decoder.cpp:1005-1012).  A NULL constant yields all-false
 * unless the op is NU / NN (ob_micro_block_decoder.cpp:1713-1715).
 * ============================================================================================= */
static int cmp_datum_param(const ora_datum *d, uint8_t obj_type, int sc, const ora_param *p) {
  if (sc == 5) {
    const uint32_t m = d->len < p->len ? d->len : p->len;
    const int c = m ? memcmp(d->ptr, p->ptr, m) : 0;
    if (c) return c < 0 ? -1 : 1;
    return d->len < p->len ? -1 : (d->len > p->len ? 1 : 0);
  }
  if (obj_signed_cmp(obj_type)) {
    /* datum value of a 4-byte map type (date) is a signed 32-bit image */
    int64_t a = (int64_t)d->ival;
    if (d->len == 4) a = (int32_t)(uint32_t)d->ival;
    const int64_t c = p->i64;
    return a < c ? -1 : (a > c ? 1 : 0);
  }
  const uint64_t a = d->ival, c = (uint64_t)p->i64;
  return a < c ? -1 : (a > c ? 1 : 0);
}

/* get_filter_cmp_ret_func: result of a compare -> predicate truth */
static int cmp_to_bool(int op, int c) {
  switch (op) {
    case ORA_OP_EQ: return c == 0;
    case ORA_OP_LE: return c <= 0;
    case ORA_OP_LT: return c < 0;
    case ORA_OP_GE: return c >= 0;
    case ORA_OP_GT: return c > 0;
    case ORA_OP_NE: return c != 0;
    default: return 0;
  }
}

/* predicate over one non-null datum */
static int eval_pred(const ora_datum *d, uint8_t obj_type, int sc, int op, const ora_param *params, int32_t n) {
  switch (op) {
    case ORA_OP_BT:
      return cmp_datum_param(d, obj_type, sc, &params[0]) >= 0 && cmp_datum_param(d, obj_type, sc, &params[1]) <= 0;
    case ORA_OP_IN:
      for (int32_t i = 0; i < n; ++i)
        if (!params[i].is_null && cmp_datum_param(d, obj_type, sc, &params[i]) == 0) return 1;
      return 0;
    default:
      return cmp_to_bool(op, cmp_datum_param(d, obj_type, sc, &params[0]));
  }
}

static int params_contain_null(int op, const ora_param *params, int32_t n) {
  if (op == ORA_OP_IN) return 0; /* IN list NULLs are skipped by the IN set build */
  for (int32_t i = 0; i < n; ++i) if (params[i].is_null) return 1;
  return 0;
}

int ora_filter_white(const ora_block *b, int32_t col, int32_t op, const ora_param *params,
                     int32_t n_params, int64_t start, int64_t count, uint8_t *bitmap) {
  if (!b || !bitmap || start < 0 || count < 0 || start + count > b->row_count || op < 0 || op >= ORA_OP_MAX)
    return ORA_INVALID_ARGUMENT;
  if ((op <= ORA_OP_NE && n_params != 1) || (op == ORA_OP_BT && n_params != 2) || (op == ORA_OP_IN && n_params < 1))
    return ORA_INVALID_ARGUMENT;
  col_dec c;
  int ret = col_dec_init(b, col, &c);
  if (ret) return ret;
  memset(bitmap, 0, (size_t)count);
  if (op != ORA_OP_NU && op != ORA_OP_NN && params_contain_null(op, params, n_params)) return ORA_SUCCESS;
  ora_datum d;
  if (c.h.type == T_DICT || c.h.type == T_RLE) {
    /* predicate over the dictionary -> ref bitset -> scan refs
     * (ob_dict_decoder.cpp:931-1017 eq/ne, :1061-1222 cmp, :1387-1487 set_res_with_bitset;
     *  RLE: ob_rle_decoder.cpp:233-520 walks the runs) */
    const uint32_t cnt = c.dict.count;
    uint8_t *hit = (uint8_t *)calloc((size_t)cnt + 2, 1);
    if (!hit) return ORA_ERR_UNEXPECTED;
    if (op == ORA_OP_NU) hit[cnt] = 1;
    else if (op == ORA_OP_NN) { memset(hit, 1, cnt); }
    else {
      for (uint32_t r = 0; r < cnt; ++r) {
        if ((ret = dict_decode(&c.dict, c.h.obj_type, r, &d))) { free(hit); return ret; }
        hit[r] = (uint8_t)eval_pred(&d, c.h.obj_type, c.sc, op, params, n_params);
      }
    }
    if (c.h.type == T_DICT) {
      for (int64_t i = 0; i < count; ++i) {
        int64_t ref = dict_row_ref(b, &c, start + i);
        if (ref > cnt + 1) { free(hit); return ORA_ERR_UNEXPECTED; }
        bitmap[i] = hit[ref];
      }
    } else {
      const int64_t n = c.rle_count;
      for (int64_t k = rle_upper_bound(&c, start) - 1; k < n; ++k) {
        const int64_t rs = (int64_t)rd_len(c.rle_row_ids + k * c.rle_row_id_byte, c.rle_row_id_byte);
        const int64_t re = k + 1 < n ? (int64_t)rd_len(c.rle_row_ids + (k + 1) * c.rle_row_id_byte, c.rle_row_id_byte) : b->row_count;
        if (rs >= start + count) break;
        const int64_t ref = rle_ref_at(&c, k);
        if (ref > cnt + 1) { free(hit); return ORA_ERR_UNEXPECTED; }
        if (!hit[ref]) continue;
        const int64_t lo = rs < start ? start : rs, hi = re > start + count ? start + count : re;
        for (int64_t r = lo; r < hi; ++r) bitmap[r - start] = 1;
      }
    }
    free(hit);
    return ORA_SUCCESS;
  }
  /* RAW / INTEGER_BASE_DIFF: null bitmap first, then the operator (ob_raw_decoder.cpp:707-794,
   * traverse_all_data :1141-1253; fast typed compare loops :1338-1382 give the same bits) */
  if (c.h.type == T_RAW && (c.h.attr & A_FIX) && !(c.h.attr & (A_EXT | A_BITPACK)) && c.sc != 5 &&
      op <= ORA_OP_NE && (c.h.length == 1 || c.h.length == 2 || c.h.length == 4 || c.h.length == 8) &&
      obj_type_size(c.h.obj_type) == (int)c.h.length) {
    /* fast_binary_comparison_operator shape: typed loop over the raw array */
    const uint8_t *col_data = c.meta;
    const int sgn = obj_signed_cmp(c.h.obj_type);
    for (int64_t i = 0; i < count; ++i) {
      const uint64_t raw = rd_len(col_data + (start + i) * c.h.length, (int)c.h.length);
      int cr;
      if (sgn) {
        int64_t a;
        switch (c.h.length) { case 1: a = (int8_t)raw; break; case 2: a = (int16_t)raw; break;
                              case 4: a = (int32_t)raw; break; default: a = (int64_t)raw; }
        cr = a < params[0].i64 ? -1 : (a > params[0].i64 ? 1 : 0);
      } else {
        const uint64_t cc = (uint64_t)params[0].i64;
        cr = raw < cc ? -1 : (raw > cc ? 1 : 0);
      }
      bitmap[i] = (uint8_t)cmp_to_bool(op, cr);
    }
    return ORA_SUCCESS;
  }
  for (int64_t i = 0; i < count; ++i) {
    if ((ret = decode_cell(b, &c, start + i, &d))) return ret;
    if (op == ORA_OP_NU) bitmap[i] = d.is_null == 1;
    else if (op == ORA_OP_NN) bitmap[i] = d.is_null != 1;
    else bitmap[i] = d.is_null ? 0 : (uint8_t)eval_pred(&d, c.h.obj_type, c.sc, op, params, n_params);
  }
  return ORA_SUCCESS;
}

/* ObPushdownFilterExecutor::execute (sql/engine/basic/ob_pushdown_filter.cpp:1551-1624): logic
 * nodes start all-true (AND) / all-false (OR), bit_and / bit_or each child, early-out when the
 * accumulated bitmap is all-false / all-true. Post-order node array, root last. */
static int exec_node(const ora_block *b, const ora_filter *f, int32_t idx, int64_t start, int64_t count,
                     uint8_t *result, int32_t *first_node) {
  const ora_node *nd = &f->nodes[idx];
  if (nd->kind == ORA_NODE_WHITE) {
    *first_node = idx;
    if (nd->param_begin < 0 || nd->param_begin + nd->n_params > f->n_params) return ORA_INVALID_ARGUMENT;
    return ora_filter_white(b, nd->col, nd->op, f->params + nd->param_begin, nd->n_params, start, count, result);
  }
  if (nd->n_children < 2) return ORA_ERR_UNEXPECTED;
  /* children are the n_children subtrees ending right before idx, in order */
  int32_t roots[64];
  if (nd->n_children > 64) return ORA_NOT_SUPPORTED;
  int32_t cur = idx - 1;
  uint8_t *tmp = (uint8_t *)malloc((size_t)(count > 0 ? count : 1));
  if (!tmp) return ORA_ERR_UNEXPECTED;
  /* find subtree roots right-to-left */
  for (int32_t k = nd->n_children - 1; k >= 0; --k) {
    if (cur < 0) { free(tmp); return ORA_INVALID_ARGUMENT; }
    roots[k] = cur;
    /* skip the subtree rooted at cur */
    int32_t need = 1;
    while (need > 0) {
      if (cur < 0) { free(tmp); return ORA_INVALID_ARGUMENT; }
      const ora_node *x = &f->nodes[cur];
      need += (x->kind == ORA_NODE_WHITE ? 0 : x->n_children) - 1;
      --cur;
    }
  }
  *first_node = cur + 1;
  const int is_and = nd->kind == ORA_NODE_AND;
  memset(result, is_and ? 1 : 0, (size_t)count);
  int ret = ORA_SUCCESS;
  for (int32_t k = 0; k < nd->n_children && ret == ORA_SUCCESS; ++k) {
    int32_t dummy;
    ret = exec_node(b, f, roots[k], start, count, tmp, &dummy);
    if (ret) break;
    int64_t ones = 0;
    if (is_and) { for (int64_t i = 0; i < count; ++i) { result[i] &= tmp[i]; ones += result[i]; } if (ones == 0) break; }
    else { for (int64_t i = 0; i < count; ++i) { result[i] |= tmp[i]; ones += result[i]; } if (ones == count) break; }
  }
  free(tmp);
  return ret;
}

int ora_filter_tree(const ora_block *b, const ora_filter *f, int64_t start, int64_t count, uint8_t *bitmap) {
  if (!b || !f || !f->nodes || f->n_nodes <= 0 || !bitmap) return ORA_INVALID_ARGUMENT;
  int32_t first = 0;
  const int ret = exec_node(b, f, f->n_nodes - 1, start, count, bitmap, &first);
  if (ret) return ret;
  return first == 0 ? ORA_SUCCESS : ORA_INVALID_ARGUMENT;
}

/* =============================================================================================
 * ObBitmap::get_row_ids (lib/container/ob_bitmap.cpp:300-333, :540-561), popcnt (:475)
 * ============================================================================================= */
int ora_bitmap_get_row_ids(const uint8_t *data, int64_t valid_bytes, int32_t *row_ids, int64_t *row_count,
                           int64_t *from, int64_t to, int64_t limit, int64_t id_offset) {
  if (!data || !row_ids || !row_count || !from) return ORA_INVALID_ARGUMENT;
  if (*from < 0 || to > valid_bytes || to < *from || limit <= 0 || *from < id_offset) return ORA_INVALID_ARGUMENT;
  const uint8_t *pos = data + *from, *end_pos = data + to;
  int64_t n = 0;
  while (n < limit && pos < end_pos) {
    if (*pos) row_ids[n++] = (int32_t)(pos - data - id_offset);
    ++pos;
  }
  if (n >= limit) {
    *from = row_ids[limit - 1] + id_offset + 1;
    n = limit;
  } else {
    *from = to;
  }
  *row_count = n;
  return ORA_SUCCESS;
}

int64_t ora_bitmap_popcnt(const uint8_t *bitmap, int64_t n) {
  int64_t c = 0;
  for (int64_t i = 0; i < n; ++i) c += bitmap[i] != 0;
  return c;
}

/* =============================================================================================
 * Whole path: per block  apply_filter -> ObBitmap ; then batches of batch_size:
 * ObBlockBatchedRowStore::get_row_ids -> ObMicroBlockDecoder::get_rows per projected column.
 * ============================================================================================= */
int ora_scan_blocks(const void *image, const int64_t *offsets, const int64_t *sizes, int32_t block_begin,
                    int32_t block_end, const ora_filter *filter, const int32_t *proj_cols, int32_t n_proj,
                    int32_t batch_size, int64_t out_row_begin, ora_scan_out *out, int64_t *total_rows,
                    int64_t *selected) {
  if (!image || !offsets || !sizes || !out || batch_size <= 0 || batch_size > ORA_MAX_BATCH) return ORA_INVALID_ARGUMENT;
  int64_t nsel = out_row_begin, nrows = 0;
  uint8_t *bitmap = 0;
  int64_t bitmap_cap = 0;
  int32_t row_ids[ORA_MAX_BATCH];
  int ret = ORA_SUCCESS;
  if (out->sel_offset) out->sel_offset[block_begin] = out_row_begin;
  for (int32_t bi = block_begin; bi < block_end && ret == ORA_SUCCESS; ++bi) {
    ora_block b;
    const uint8_t *bbuf = (const uint8_t *)image + offsets[bi];
    if ((ret = ora_block_init(&b, bbuf, sizes[bi]))) break;
    const int64_t rc = b.row_count;
    nrows += rc;
    if (rc > bitmap_cap) {
      free(bitmap);
      bitmap_cap = rc + 1024;
      bitmap = (uint8_t *)malloc((size_t)bitmap_cap);
      if (!bitmap) { ret = ORA_ERR_UNEXPECTED; break; }
    }
    if (filter && filter->n_nodes > 0) {
      if ((ret = ora_filter_tree(&b, filter, 0, rc, bitmap))) break;
    } else {
      memset(bitmap, 1, (size_t)rc);
    }
    int64_t from = 0;
    while (from < rc && ret == ORA_SUCCESS) {
      int64_t cnt = 0;
      if ((ret = ora_bitmap_get_row_ids(bitmap, rc, row_ids, &cnt, &from, rc, batch_size, 0))) break;
      if (cnt == 0) continue;
      if (nsel + cnt > out->cap_rows) { ret = ORA_BUF_NOT_ENOUGH; break; }
      if (out->row_ids) memcpy(out->row_ids + nsel, row_ids, (size_t)cnt * 4);
      for (int32_t p = 0; p < n_proj && ret == ORA_SUCCESS; ++p) {
        col_hdr h;
        if ((ret = get_col(&b, proj_cols[p], &h))) break;
        if (obj_store_class(h.obj_type) == 5) {
          if (!out->data || !out->data[p]) continue;
          const uint8_t *ptrs[ORA_MAX_BATCH];
          int32_t lens[ORA_MAX_BATCH];
          uint64_t nulls[ORA_MAX_BATCH / 64 + 1];
          memset(nulls, 0, sizeof(nulls));
          memset(ptrs, 0, sizeof(ptrs[0]) * (size_t)cnt);
          memset(lens, 0, sizeof(lens[0]) * (size_t)cnt);
          int32_t hn = 0;
          if ((ret = ora_get_rows_discrete(&b, proj_cols[p], row_ids, cnt, 0, ptrs, lens, nulls, &hn))) break;
          uint64_t *dp = (uint64_t *)out->data[p];
          for (int64_t i = 0; i < cnt; ++i) {
            const int isnull = (nulls[i / 64] >> (i % 64)) & 1;
            /* a string a codec rebuilt (HEX / STRING_DIFF / STRING_PREFIX) lives in the oracle's arena, not in the image: its
             * absolute address is reported (the reference's datum points into allocator memory there, too) */
            const int in_block = !isnull && ptrs[i] >= (const uint8_t *)b.buf && ptrs[i] < (const uint8_t *)b.buf + b.size;
            dp[nsel + i] = isnull ? 0 : (in_block ? out->string_base + (uint64_t)(ptrs[i] - (const uint8_t *)image) : (uint64_t)(uintptr_t)ptrs[i]);
            if (out->lens && out->lens[p]) out->lens[p][nsel + i] = isnull ? 0 : lens[i];
            if (isnull && out->nulls && out->nulls[p]) bitvec_set(out->nulls[p], nsel + i);
          }
          if (hn && out->has_null) out->has_null[p] = 1;
        } else {
          if (!out->data || !out->data[p]) continue;
          int32_t hn = 0;
          if ((ret = ora_get_rows_fixed(&b, proj_cols[p], row_ids, cnt, nsel, out->data[p], obj_datum_len(h.obj_type),
                                        out->nulls ? out->nulls[p] : 0, &hn))) break;
          if (hn && out->has_null) out->has_null[p] = 1;
        }
      }
      nsel += cnt;
    }
    if (out->sel_offset) out->sel_offset[bi + 1] = nsel;
  }
  free(bitmap);
  if (total_rows) *total_rows = nrows;
  if (selected) *selected = nsel - out_row_begin;
  return ret;
}

/* ---- multi-threaded timing harness ----------------------------------------------------------- */
#define MT_CHUNK 32
typedef struct mt_arg {
  const void *image; const int64_t *offsets, *sizes; int32_t b0, b1;
  int32_t *next_block; int32_t n_blocks; /* shared work queue: chunks of MT_CHUNK blocks, claimed atomically */
  const ora_filter *filter; const int32_t *proj; int32_t n_proj, batch;
  int64_t rows, sel; uint64_t checksum; int ret;
} mt_arg;

static void *mt_worker(void *vp) {
  mt_arg *a = (mt_arg *)vp;
  a->rows = a->sel = 0; a->checksum = 0; a->ret = ORA_SUCCESS;
  uint8_t *bitmap = 0; int64_t bitmap_cap = 0;
  int32_t row_ids[ORA_MAX_BATCH];
  uint64_t vals[ORA_MAX_BATCH];
  const uint8_t *ptrs[ORA_MAX_BATCH];
  int32_t lens[ORA_MAX_BATCH];
  uint64_t nulls[ORA_MAX_BATCH / 64 + 1];
  for (;;) {
    if (a->ret != ORA_SUCCESS) break;
    if (a->b0 >= a->b1) { /* claim the next chunk (blocks of one thread stay contiguous, like a PX granule) */
      a->b0 = __atomic_fetch_add(a->next_block, MT_CHUNK, __ATOMIC_RELAXED);
      if (a->b0 >= a->n_blocks || a->ret != ORA_SUCCESS) break;
      a->b1 = a->b0 + MT_CHUNK < a->n_blocks ? a->b0 + MT_CHUNK : a->n_blocks;
    }
    const int32_t bi = a->b0++;
    ora_block b;
    if ((a->ret = ora_block_init(&b, (const uint8_t *)a->image + a->offsets[bi], a->sizes[bi]))) break;
    const int64_t rc = b.row_count;
    a->rows += rc;
    if (rc > bitmap_cap) { free(bitmap); bitmap_cap = rc + 1024; bitmap = (uint8_t *)malloc((size_t)bitmap_cap); }
    if (a->filter && a->filter->n_nodes > 0) { if ((a->ret = ora_filter_tree(&b, a->filter, 0, rc, bitmap))) break; }
    else memset(bitmap, 1, (size_t)rc);
    int64_t from = 0;
    while (from < rc && a->ret == ORA_SUCCESS) {
      int64_t cnt = 0;
      if ((a->ret = ora_bitmap_get_row_ids(bitmap, rc, row_ids, &cnt, &from, rc, a->batch, 0))) break;
      if (!cnt) continue;
      for (int32_t p = 0; p < a->n_proj && a->ret == ORA_SUCCESS; ++p) {
        col_hdr h;
        if ((a->ret = get_col(&b, a->proj[p], &h))) break;
        int32_t hn = 0;
        memset(nulls, 0, sizeof(uint64_t) * (size_t)(cnt / 64 + 1));
        if (obj_store_class(h.obj_type) == 5) {
          if ((a->ret = ora_get_rows_discrete(&b, a->proj[p], row_ids, cnt, 0, ptrs, lens, nulls, &hn))) break;
          for (int64_t i = 0; i < cnt; ++i) if (!((nulls[i / 64] >> (i % 64)) & 1)) a->checksum += (uint64_t)lens[i] + ptrs[i][0];
        } else {
          const int el = obj_datum_len(h.obj_type);
          memset(vals, 0, sizeof(uint64_t) * (size_t)cnt);
          if ((a->ret = ora_get_rows_fixed(&b, a->proj[p], row_ids, cnt, 0, vals, el, nulls, &hn))) break;
          if (el == 8) for (int64_t i = 0; i < cnt; ++i) a->checksum += vals[i];
          else for (int64_t i = 0; i < cnt; ++i) a->checksum += rd_len((const uint8_t *)vals + i * el, el);
        }
      }
      a->sel += cnt;
    }
  }
  free(bitmap);
  return 0;
}

int ora_scan_blocks_mt(const void *image, const int64_t *offsets, const int64_t *sizes, int32_t n_blocks,
                       const ora_filter *filter, const int32_t *proj_cols, int32_t n_proj, int32_t batch_size,
                       int32_t n_threads, int64_t *total_rows, int64_t *selected, uint64_t *checksum) {
  if (!image || n_blocks <= 0 || n_threads <= 0 || batch_size <= 0 || batch_size > ORA_MAX_BATCH) return ORA_INVALID_ARGUMENT;
  if (n_threads > n_blocks) n_threads = n_blocks;
  mt_arg *args = (mt_arg *)calloc((size_t)n_threads, sizeof(mt_arg));
  pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
  if (!args || !th) { free(args); free(th); return ORA_ERR_UNEXPECTED; }
  int32_t next_block = 0;
  for (int32_t t = 0; t < n_threads; ++t) {
    args[t].image = image; args[t].offsets = offsets; args[t].sizes = sizes;
    args[t].b0 = args[t].b1 = 0; args[t].next_block = &next_block; args[t].n_blocks = n_blocks;
    args[t].filter = filter; args[t].proj = proj_cols; args[t].n_proj = n_proj; args[t].batch = batch_size;
    if (t > 0) pthread_create(&th[t], 0, mt_worker, &args[t]);
  }
  mt_worker(&args[0]);
  int ret = args[0].ret;
  int64_t rows = args[0].rows, sel = args[0].sel;
  uint64_t cs = args[0].checksum;
  for (int32_t t = 1; t < n_threads; ++t) {
    pthread_join(th[t], 0);
    if (args[t].ret && !ret) ret = args[t].ret;
    rows += args[t].rows; sel += args[t].sel; cs += args[t].checksum;
  }
  free(args); free(th);
  if (total_rows) *total_rows = rows;
  if (selected) *selected = sel;
  if (checksum) *checksum = cs;
  return ret;
}

/* =============================================================================================
 * Major compaction merge
 * ============================================================================================= */
int ora_decode_column_ext(const void *image, const int64_t *offsets, const int64_t *sizes, int32_t n_blocks,
                          int32_t col, int64_t *vals, uint8_t *ext, int64_t cap, int64_t *rows) {
  if (!image || !offsets || !sizes || !vals || !ext) return ORA_INVALID_ARGUMENT;
  int64_t n = 0;
  for (int32_t bi = 0; bi < n_blocks; ++bi) {
    ora_block b;
    int ret = ora_block_init(&b, (const uint8_t *)image + offsets[bi], sizes[bi]);
    if (ret) return ret;
    col_dec c;
    if ((ret = col_dec_init(&b, col, &c))) return ret;
    if (c.sc == 5) return ORA_NOT_SUPPORTED;
    if (n + b.row_count > cap) return ORA_BUF_NOT_ENOUGH;
    for (int64_t r = 0; r < b.row_count; ++r) {
      ora_datum d;
      if ((ret = decode_cell(&b, &c, r, &d))) return ret;
      vals[n] = d.is_null ? 0 : (int64_t)d.ival;
      ext[n] = (uint8_t)d.is_null;
      ++n;
    }
  }
  if (rows) *rows = n;
  return ORA_SUCCESS;
}

/* The reference keeps the run heads in a loser tree (ObPartitionMajorRowsMerger); any priority
 * structure that pops (rowkey ascending, newer table first among equal rowkeys) gives the same
 * sequence. A binary heap over (key, -run) is used here. */
typedef struct mrg_head { int64_t key; int32_t run; int64_t at; } mrg_head;
static const ora_merge_run *g_mrg_runs;   /* the runs of the merge in progress (single-threaded test infrastructure) */
/* rowkey order: column by column (ObStorageDatumUtils / ObPartitionMergeLoserTreeCmp::compare_rowkey) */
static int mrg_key_cmp(const mrg_head *a, const mrg_head *b) {
  if (a->key != b->key) return a->key < b->key ? -1 : 1;
  const ora_merge_run *ra = &g_mrg_runs[a->run], *rb = &g_mrg_runs[b->run];
  for (int32_t c = 0; c < ra->n_more_keys; ++c) {
    const int64_t x = ra->more_keys[c][a->at], y = rb->more_keys[c][b->at];
    if (x != y) return x < y ? -1 : 1;
  }
  return 0;
}
static int mrg_less(const mrg_head *a, const mrg_head *b) {
  const int c = mrg_key_cmp(a, b);
  if (c) return c < 0;
  return a->run > b->run; /* newer table first */
}
static void mrg_sift_down(mrg_head *h, int32_t n, int32_t i) {
  for (;;) {
    int32_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && mrg_less(&h[l], &h[m])) m = l;
    if (r < n && mrg_less(&h[r], &h[m])) m = r;
    if (m == i) return;
    const mrg_head t = h[i]; h[i] = h[m]; h[m] = t;
    i = m;
  }
}

int ora_major_merge(const ora_merge_run *runs, int32_t n_runs, int32_t n_cols, const int64_t *default_vals,
                    const uint8_t *default_null, int64_t out_cap, int64_t *out_key, int64_t *const *out_vals,
                    uint8_t *const *out_null, int64_t *out_rows, int64_t *stats) {
  return ora_major_merge_keys(runs, n_runs, n_cols, default_vals, default_null, out_cap, out_key, 0, out_vals, out_null,
                              out_rows, stats);
}

int ora_major_merge_keys(const ora_merge_run *runs, int32_t n_runs, int32_t n_cols, const int64_t *default_vals,
                         const uint8_t *default_null, int64_t out_cap, int64_t *out_key, int64_t *const *out_more_keys,
                         int64_t *const *out_vals, uint8_t *const *out_null, int64_t *out_rows, int64_t *stats) {
  if (!runs || n_runs <= 0 || n_runs > 64 || n_cols < 0 || n_cols > 64 || !out_key || !out_rows) return ORA_INVALID_ARGUMENT;
  const int32_t n_more = runs[0].n_more_keys;
  for (int32_t r = 0; r < n_runs; ++r)
    if (runs[r].n_more_keys != n_more || (n_more > 0 && (!runs[r].more_keys || !out_more_keys))) return ORA_INVALID_ARGUMENT;
  g_mrg_runs = runs;
  mrg_head heap[64];
  int64_t pos[64];
  int32_t hn = 0;
  for (int32_t r = 0; r < n_runs; ++r) {
    pos[r] = 0;
    if (runs[r].n > 0) { heap[hn].key = runs[r].key[0]; heap[hn].run = r; heap[hn].at = 0; ++hn; }
  }
  for (int32_t i = hn / 2 - 1; i >= 0; --i) mrg_sift_down(heap, hn, i);
  int64_t nout = 0, dropped = 0, fused = 0;
  while (hn > 0) {
    /* find_rowkey_minimum_iters: every iter whose current rowkey equals the minimum, newest first */
    const int64_t key = heap[0].key;
    const mrg_head min_head = heap[0];
    int32_t iters[64];
    int32_t ni = 0;
    while (hn > 0 && mrg_key_cmp(&heap[0], &min_head) == 0) {
      iters[ni++] = heap[0].run;
      const int32_t r = heap[0].run;
      ++pos[r];
      if (pos[r] < runs[r].n) {
        mrg_head nxt = {runs[r].key[pos[r]], r, pos[r]};
        const mrg_head cur = {runs[r].key[pos[r] - 1], r, pos[r] - 1};
        if (mrg_key_cmp(&nxt, &cur) <= 0) return ORA_INVALID_DATA; /* run not strictly ascending */
        heap[0] = nxt;
      } else {
        heap[0] = heap[--hn];
      }
      mrg_sift_down(heap, hn, 0);
    }
    /* fuse_row: newest first */
    int64_t rv[64];
    uint8_t rs[64];     /* 0 value, 1 NULL, 2 NOP (still open) */
    int result_delete = 0, first = 1, left = 0;
    for (int32_t k = 0; k < ni; ++k) {
      const ora_merge_run *run = &runs[iters[k]];
      const int64_t at = pos[iters[k]] - 1;
      const int flag = run->flag ? run->flag[at] : ORA_DF_INSERT;
      if (flag == ORA_DF_NOT_EXIST) continue;                 /* former.row_flag_.is_not_exist(): nothing */
      if (flag != ORA_DF_DELETE && flag != ORA_DF_INSERT && flag != ORA_DF_UPDATE) return ORA_INVALID_ARGUMENT;
      if (flag == ORA_DF_DELETE) {
        if (first) result_delete = 1;                         /* result flag = delete */
        break;                                                /* final_result = true */
      }
      left = 0;
      for (int32_t c = 0; c < n_cols; ++c) {
        if (!first && rs[c] != 2) continue;                   /* only the open NOP positions */
        const uint8_t e = run->ext[c][at];
        if (first || e != 2) { rv[c] = e ? 0 : run->vals[c][at]; rs[c] = e; }
        if (e == 2) ++left;
      }
      first = 0;
      if (left == 0) break;                                   /* final_result */
    }
    if (first && !result_delete) continue;                    /* only not-exist rows: no output */
    if (result_delete) { ++dropped; continue; }               /* inner_process drops delete rows */
    /* end_fuse_row: open NOPs take the default row */
    for (int32_t c = 0; c < n_cols; ++c) {
      if (rs[c] == 2) {
        const int dn = default_null ? default_null[c] : 1;
        rs[c] = dn ? 1 : 0;
        rv[c] = dn ? 0 : (default_vals ? default_vals[c] : 0);
      }
    }
    if (nout >= out_cap) return ORA_BUF_NOT_ENOUGH;
    out_key[nout] = key;
    for (int32_t c = 0; c < n_more; ++c) out_more_keys[c][nout] = runs[min_head.run].more_keys[c][min_head.at];
    for (int32_t c = 0; c < n_cols; ++c) { out_vals[c][nout] = rv[c]; out_null[c][nout] = rs[c]; }
    if (ni > 1) ++fused;
    ++nout;
  }
  *out_rows = nout;
  if (stats) { stats[0] = dropped; stats[1] = fused; }
  return ORA_SUCCESS;
}

/* =============================================================================================
 * Skip index: aggregate row reader + min / max / null-count filter verdicts.
 * ============================================================================================= */
static uint64_t rd_le(const uint8_t *p, int bytes) {
  uint64_t v = 0;
  memcpy(&v, p, (size_t)bytes);
  return v;
}

/* ObAggRowReader::init / binary_search_col / find_col / read_cell (ob_agg_row_struct.cpp:303-482) */
int ora_agg_row_read(const void *buf_, int64_t size, uint32_t col_idx, int32_t col_type, const uint8_t **data,
                     int32_t *len, int32_t *is_prefix) {
  const uint8_t *buf = (const uint8_t *)buf_;
  if (!buf || size < 8 || !data || !len || !is_prefix || col_type < 0 || col_type >= 6) return ORA_INVALID_ARGUMENT;
  *data = 0;
  *len = 0;
  *is_prefix = 0;
  const int16_t version = (int16_t)rd_le(buf, 2), cnt = (int16_t)rd_le(buf + 4, 2);
  const uint16_t pack = (uint16_t)rd_le(buf + 6, 2);
  const int idx_size = pack & 0x3f, idx_off_size = (pack >> 6) & 7, cell_off_size = (pack >> 9) & 7, bitmap_size = (pack >> 12) & 0xf;
  if (version < 1 || version > 3 || cnt <= 0 || idx_size <= 0 || idx_off_size <= 0 || bitmap_size != 1) return ORA_INVALID_DATA;
  if (idx_size > 4 || (idx_off_size != 1 && idx_off_size != 2) || (cell_off_size != 1 && cell_off_size != 2)) return ORA_INVALID_DATA;
  const int64_t header_size = 8 + (int64_t)cnt * idx_size + (int64_t)cnt * idx_off_size;
  if (size < header_size) return ORA_INVALID_DATA;
  const uint8_t *idx_arr = buf + 8, *off_arr = idx_arr + (int64_t)cnt * idx_size;
  int lo = 0, hi = cnt;                    /* lower_bound over the sorted column indexes */
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (rd_le(idx_arr + (int64_t)mid * idx_size, idx_size) < col_idx) lo = mid + 1; else hi = mid;
  }
  if (lo >= cnt || rd_le(idx_arr + (int64_t)lo * idx_size, idx_size) != col_idx) return ORA_SUCCESS; /* not aggregated */
  const int64_t pos = (int64_t)rd_le(off_arr + (int64_t)lo * idx_off_size, idx_off_size);
  if (pos == 0) return ORA_SUCCESS;
  const int bitmaps = version >= 2 ? 2 : 1;
  if (pos + bitmaps > size) return ORA_INVALID_DATA;
  const uint8_t *cell = buf + pos;
  const uint32_t types = cell[0], mask = 1u << col_type;
  if (!(types & mask)) return ORA_SUCCESS;
  if (pos + bitmaps + cell_off_size > size) return ORA_INVALID_DATA;
  int pre = 0;
  for (uint32_t n = types & (mask - 1); n; n &= n - 1) ++pre;
  const uint8_t *offs = cell + bitmaps;
  if (pos + bitmaps + (int64_t)(pre + 2) * cell_off_size > size) return ORA_INVALID_DATA;
  const int64_t a = (int64_t)rd_le(offs + (int64_t)pre * cell_off_size, cell_off_size);
  const int64_t b = (int64_t)rd_le(offs + (int64_t)(pre + 1) * cell_off_size, cell_off_size);
  if (b < a || pos + b > size) return ORA_INVALID_DATA;
  *data = cell + a;
  *len = (int32_t)(b - a);
  if (version >= 2) *is_prefix = (cell[1] & mask) != 0;
  return ORA_SUCCESS;
}

/* ObSkipIndexCmpRes */
typedef struct sk_cmp { int cmp, certain; } sk_cmp;
static int sk_gt(sk_cmp r) { return r.certain && r.cmp > 0; }
static int sk_lt(sk_cmp r) { return r.certain && r.cmp < 0; }
static int sk_le(sk_cmp r) { return r.certain && r.cmp <= 0; }
static int sk_ge(sk_cmp r) { return r.certain && r.cmp >= 0; }
static int sk_eq(sk_cmp r) { return r.certain && r.cmp == 0; }

/* ObSkipIndexFilterExecutor::compare (:498-528) with compare_for_non_pad_charset (:476-496) and
 * compare_with_prefix (:415-452) for a binary collation (one character = one byte). skip == NULL: the
 * aggregate is missing, i.e. -infinity for a min and +infinity for a max. */
static sk_cmp sk_compare(const ora_datum *skip, int compare_min, int is_prefix, uint8_t obj_type, int sc, const ora_param *f) {
  sk_cmp r = {0, 0};
  if (!skip) { r.certain = 1; r.cmp = compare_min ? -1 : 1; return r; }
  r.cmp = cmp_datum_param(skip, obj_type, sc, f);
  if (!is_prefix) { r.certain = 1; return r; }
  if (r.cmp >= 0) { r.cmp = 1; r.certain = 1; return r; }
  if (skip->len >= f->len) { r.certain = 1; return r; }
  ora_param fp = *f;                       /* the constant cut to the prefix's length */
  fp.len = skip->len;
  const int c2 = cmp_datum_param(skip, obj_type, sc, &fp);
  r.certain = c2 == r.cmp;                 /* c2 == 0: the prefix is a prefix of the constant -> uncertain */
  return r;
}

static int sk_param_less(const ora_param *a, const ora_param *b, uint8_t obj_type, int sc) {
  ora_datum d;
  memset(&d, 0, sizeof(d));
  d.ptr = (const uint8_t *)a->ptr;
  d.len = sc == 5 ? a->len : 8;
  d.ival = (uint64_t)a->i64;
  return cmp_datum_param(&d, obj_type, sc, b) < 0;
}

int ora_skip_index_leaf(const void *agg, int64_t agg_size, int64_t row_count, int32_t col, uint8_t obj_type, int32_t op,
                        const ora_param *params, int32_t n_params, int32_t *mask) {
  if (!mask || op < 0 || op >= ORA_OP_MAX || col < 0) return ORA_INVALID_ARGUMENT;
  *mask = ORA_MASK_UNCERTAIN;
  const int sc = obj_store_class(obj_type);
  if (sc != 1 && sc != 2 && sc != 5) return ORA_NOT_SUPPORTED;
  /* is_cmp_op_with_null_ref_value (:163-164): a NULL constant (or an IN list left empty) never matches */
  if (op != ORA_OP_NU && op != ORA_OP_NN) {
    int null_ref = params_contain_null(op, params, n_params);
    if (op == ORA_OP_IN) {
      null_ref = 1;
      for (int32_t i = 0; i < n_params; ++i) if (!params[i].is_null) null_ref = 0;
    }
    if (null_ref) { *mask = ORA_MASK_ALWAYS_FALSE; return ORA_SUCCESS; }
  }
  if (!agg || agg_size <= 0) return ORA_SUCCESS;    /* !has_agg_data(): nothing to decide with */
  const uint8_t *p_nc = 0, *p_min = 0, *p_max = 0;
  int32_t l_nc = 0, l_min = 0, l_max = 0, pre_nc = 0, min_prefix = 0, max_prefix = 0;
  int ret = ora_agg_row_read(agg, agg_size, (uint32_t)col, 2, &p_nc, &l_nc, &pre_nc);
  if (!ret) ret = ora_agg_row_read(agg, agg_size, (uint32_t)col, 0, &p_min, &l_min, &min_prefix);
  if (!ret) ret = ora_agg_row_read(agg, agg_size, (uint32_t)col, 1, &p_max, &l_max, &max_prefix);
  if (ret) return ret;
  if (!p_nc && !p_min && !p_max) return ORA_SUCCESS;  /* ObMinMaxFilterParam::is_uncertain */
  int64_t null_count = 0;
  if (p_nc) {
    if (l_nc != 8) return ORA_INVALID_DATA;
    memcpy(&null_count, p_nc, 8);
    if (null_count < 0 || null_count > row_count) return ORA_ERR_UNEXPECTED;
  }
  const int all_null = p_nc && null_count == row_count;
  const int all_not_null = p_nc && null_count == 0;
  const int has_null = p_nc ? (null_count > 0 && null_count < row_count) : 1;
  const int min_max_null = !p_min && !p_max;
  ora_datum dmin, dmax;
  memset(&dmin, 0, sizeof(dmin));
  memset(&dmax, 0, sizeof(dmax));
  if (p_min) { dmin.ptr = p_min; dmin.len = (uint32_t)l_min; if (sc != 5) { if (l_min > 8) return ORA_INVALID_DATA; memcpy(&dmin.ival, p_min, (size_t)l_min); } }
  if (p_max) { dmax.ptr = p_max; dmax.len = (uint32_t)l_max; if (sc != 5) { if (l_max > 8) return ORA_INVALID_DATA; memcpy(&dmax.ival, p_max, (size_t)l_max); } }
  const ora_datum *mn = p_min ? &dmin : 0, *mx = p_max ? &dmax : 0;
  int m = ORA_MASK_UNCERTAIN;
#define CMP_MIN(f) sk_compare(mn, 1, min_prefix, obj_type, sc, (f))
#define CMP_MAX(f) sk_compare(mx, 0, max_prefix, obj_type, sc, (f))
  if (op == ORA_OP_NU) {
    m = all_not_null ? ORA_MASK_ALWAYS_FALSE : (all_null ? ORA_MASK_ALWAYS_TRUE : ORA_MASK_UNCERTAIN);
  } else if (op == ORA_OP_NN) {
    m = all_null ? ORA_MASK_ALWAYS_FALSE : (all_not_null ? ORA_MASK_ALWAYS_TRUE : ORA_MASK_UNCERTAIN);
  } else if (all_null) {
    m = ORA_MASK_ALWAYS_FALSE;
  } else if (min_max_null) {
    m = ORA_MASK_UNCERTAIN;                /* the reference leaves the (reset) mask untouched here (:296-297) */
  } else {
    const ora_param *ref = &params[0];
    sk_cmp a, b;
    switch (op) {
      case ORA_OP_EQ:
        a = CMP_MIN(ref);
        if (sk_gt(a)) { m = ORA_MASK_ALWAYS_FALSE; break; }
        b = CMP_MAX(ref);
        if (sk_lt(b)) m = ORA_MASK_ALWAYS_FALSE;
        else if (sk_eq(a) && sk_eq(b)) m = ORA_MASK_ALWAYS_TRUE;
        break;
      case ORA_OP_NE:
        a = CMP_MIN(ref);
        if (sk_gt(a)) { m = ORA_MASK_ALWAYS_TRUE; break; }
        b = CMP_MAX(ref);
        if (sk_lt(b)) m = ORA_MASK_ALWAYS_TRUE;
        else if (sk_eq(a) && sk_eq(b)) m = ORA_MASK_ALWAYS_FALSE;
        break;
      case ORA_OP_GT:
        a = CMP_MIN(ref);
        if (sk_gt(a)) { m = ORA_MASK_ALWAYS_TRUE; break; }
        b = CMP_MAX(ref);
        if (sk_le(b)) m = ORA_MASK_ALWAYS_FALSE;
        break;
      case ORA_OP_GE:
        a = CMP_MIN(ref);
        if (sk_ge(a)) { m = ORA_MASK_ALWAYS_TRUE; break; }
        b = CMP_MAX(ref);
        if (sk_lt(b)) m = ORA_MASK_ALWAYS_FALSE;
        break;
      case ORA_OP_LT:
        a = CMP_MIN(ref);
        if (sk_ge(a)) { m = ORA_MASK_ALWAYS_FALSE; break; }
        b = CMP_MAX(ref);
        if (sk_lt(b)) m = ORA_MASK_ALWAYS_TRUE;
        break;
      case ORA_OP_LE:
        a = CMP_MIN(ref);
        if (sk_gt(a)) { m = ORA_MASK_ALWAYS_FALSE; break; }
        b = CMP_MAX(ref);
        if (sk_le(b)) m = ORA_MASK_ALWAYS_TRUE;
        break;
      case ORA_OP_BT: {
        const ora_param *left = &params[0], *right = &params[1];
        a = CMP_MIN(right);
        if (sk_gt(a)) { m = ORA_MASK_ALWAYS_FALSE; break; }
        b = CMP_MAX(left);
        if (sk_lt(b)) { m = ORA_MASK_ALWAYS_FALSE; break; }
        if (sk_ge(CMP_MIN(left)) && sk_le(CMP_MAX(right))) m = ORA_MASK_ALWAYS_TRUE;
        break;
      }
      case ORA_OP_IN: {
        /* the executor keeps the non-NULL constants sorted and distinct (init_in_eval_datums) */
        ora_param *sorted = (ora_param *)malloc(sizeof(ora_param) * (size_t)(n_params > 0 ? n_params : 1));
        if (!sorted) return ORA_ERR_UNEXPECTED;
        int32_t n = 0;
        for (int32_t i = 0; i < n_params; ++i) if (!params[i].is_null) sorted[n++] = params[i];
        for (int32_t i = 1; i < n; ++i) {            /* insertion sort: IN lists are short */
          ora_param t = sorted[i];
          int32_t j = i;
          while (j > 0 && sk_param_less(&t, &sorted[j - 1], obj_type, sc)) { sorted[j] = sorted[j - 1]; --j; }
          sorted[j] = t;
        }
        int32_t pos = 0, equal = 0;
        if (mn) {
          /* min prefix: upper_bound (first constant > min), else lower_bound (first constant >= min) */
          while (pos < n) {
            const int c = cmp_datum_param(mn, obj_type, sc, &sorted[pos]);   /* min vs constant */
            if (min_prefix ? c >= 0 : c > 0) ++pos; else { equal = !min_prefix && c == 0; break; }
          }
        }
        if (pos == n) {
          m = ORA_MASK_ALWAYS_FALSE;
        } else {
          b = CMP_MAX(&sorted[pos]);
          if (sk_gt(b)) m = ORA_MASK_UNCERTAIN;
          else if (sk_lt(b)) m = ORA_MASK_ALWAYS_FALSE;
          else if (equal) m = max_prefix ? ORA_MASK_UNCERTAIN : ORA_MASK_ALWAYS_TRUE;
        }
        free(sorted);
        break;
      }
      default:
        return ORA_NOT_SUPPORTED;
    }
  }
#undef CMP_MIN
#undef CMP_MAX
  if (has_null && m == ORA_MASK_ALWAYS_TRUE) m = ORA_MASK_UNCERTAIN;
  *mask = m;
  return ORA_SUCCESS;
}

int ora_skip_index_filter(const void *agg, int64_t agg_size, int64_t row_count, const uint8_t *col_types, int32_t n_cols,
                          const ora_filter *f, int32_t *mask) {
  if (!f || !mask || !col_types || f->n_nodes <= 0 || f->n_nodes > 256) return ORA_INVALID_ARGUMENT;
  int32_t stack[256];
  int32_t sp = 0;
  for (int32_t i = 0; i < f->n_nodes; ++i) {
    const ora_node *nd = &f->nodes[i];
    if (nd->kind == ORA_NODE_WHITE) {
      if (nd->col < 0 || nd->col >= n_cols) return ORA_INVALID_ARGUMENT;
      int32_t m = ORA_MASK_UNCERTAIN;
      const int ret = ora_skip_index_leaf(agg, agg_size, row_count, nd->col, col_types[nd->col], nd->op,
                                          f->params + nd->param_begin, nd->n_params, &m);
      if (ret == ORA_NOT_SUPPORTED) m = ORA_MASK_UNCERTAIN;   /* no skip index for this column type */
      else if (ret) return ret;
      stack[sp++] = m;
    } else {
      if (nd->n_children < 1 || nd->n_children > sp) return ORA_INVALID_ARGUMENT;
      /* ObBoolMask operator& / operator| (ob_pushdown_filter.h:133-158), left to right */
      int32_t bm = stack[sp - nd->n_children];
      for (int32_t k = 1; k < nd->n_children; ++k) {
        const int32_t c = stack[sp - nd->n_children + k];
        if (nd->kind == ORA_NODE_AND) {
          if (c == ORA_MASK_ALWAYS_TRUE) { /* bm stays */ }
          else if (bm == ORA_MASK_ALWAYS_TRUE) bm = c;
          else if (c == ORA_MASK_ALWAYS_FALSE || bm == ORA_MASK_ALWAYS_FALSE) bm = ORA_MASK_ALWAYS_FALSE;
          else bm = ORA_MASK_UNCERTAIN;
        } else {
          if (c == ORA_MASK_ALWAYS_FALSE) { /* bm stays */ }
          else if (bm == ORA_MASK_ALWAYS_FALSE) bm = c;
          else if (c == ORA_MASK_ALWAYS_TRUE || bm == ORA_MASK_ALWAYS_TRUE) bm = ORA_MASK_ALWAYS_TRUE;
          else bm = ORA_MASK_UNCERTAIN;
        }
      }
      sp -= nd->n_children;
      stack[sp++] = bm;
    }
  }
  if (sp != 1) return ORA_INVALID_ARGUMENT;
  *mask = stack[0];
  return ORA_SUCCESS;
}
