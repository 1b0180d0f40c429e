/*
 * ob_oracle.h -- CPU oracle for the columnar-scan hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's (OceanBase 4.6.0.0)
 * CPU algorithm for PAX and CS micro-block decode, pushed-down white filters, ObBitmap row-id
 * extraction, batch projection and the major-compaction merge (row fuse).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it; the product (libobgpu_scan.so) never does.
 *
 * Pinning status: the bit-stream primitives are checked against the reference's own
 * ObBitStream compiled from /root/reference (oracle/_ref, see Makefile + tests/test_bitstream_kat.py)
 * and against the reference's known-answer tests (unittest/.../test_bit_stream.cpp:151-200);
 * filter semantics are checked against the popcount expectations of
 * unittest/.../test_raw_decoder.cpp:774-1200 (tests/test_oracle_filter_kat.py) and
 * test_const_decoder.cpp:111-770 (tests/test_const_kat.py); the row fuse of the compaction merge
 * against unittest/storage/test_row_fuse.cpp:118-227 (tests/test_major_merge_kat.py); the aggregate
 * row and the skip-index verdicts against test_agg_row_struct.cpp:209-293 and
 * test_skip_index_filter.cpp:405-1420 (tests/test_skip_index_kat.py). CS blocks: filter results are
 * checked against the datasets and expected counts of the reference's CS pd-filter unit tests
 * (cs_encoding/test_integer_pd_filter.cpp, test_int_dict_pd_filter.cpp, test_string_pd_filter.cpp;
 * tests/test_cs_reference_filter_kat.py); the block framing itself has no reference vectors and is a
 * restatement only -- byte-level CS parity is unpinned.
 * The reference ships no byte-level golden micro-blocks, so whole-block decode is
 * pinned by construction (layout restated from the encoder) and by the independent GPU decoder.
 *
 * Each function cites the reference file:line it follows.
 */
#ifndef OB_ORACLE_H_
#define OB_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_SUCCESS 0
#define ORA_INVALID_ARGUMENT (-4002)
#define ORA_NOT_SUPPORTED (-4007)
#define ORA_ERR_UNEXPECTED (-4016)
#define ORA_BUF_NOT_ENOUGH (-4024)
#define ORA_INVALID_DATA (-4070)

enum { ORA_OP_EQ = 0, ORA_OP_LE, ORA_OP_LT, ORA_OP_GE, ORA_OP_GT, ORA_OP_NE, ORA_OP_BT, ORA_OP_IN,
       ORA_OP_NU, ORA_OP_NN, ORA_OP_MAX };
enum { ORA_NODE_WHITE = 0, ORA_NODE_AND = 1, ORA_NODE_OR = 2 };

typedef struct ora_param {
  int64_t i64;
  const char *ptr;
  uint32_t len;
  int32_t is_null;
} ora_param;

typedef struct ora_node {
  int32_t kind, op, col, param_begin, n_params, n_children;
} ora_node;

typedef struct ora_filter {
  const ora_node *nodes;
  int32_t n_nodes;
  const ora_param *params;
  int32_t n_params;
} ora_filter;

/* One decoded cell (ObDatum restated: ptr_/len_/null_). Integers are materialised in ival. */
typedef struct ora_datum {
  const uint8_t *ptr; /* strings: points into the block buffer */
  uint32_t len;       /* datum length: 8/4/1 for integer classes, byte length for strings */
  int32_t is_null;
  uint64_t ival;      /* integer classes: value image (len low bytes are significant) */
} ora_datum;

/* Parsed view of one micro-block (ObMicroBlockDecoder::do_init). */
typedef struct ora_block {
  const uint8_t *buf;
  int64_t size;
  uint32_t header_size, row_count, row_data_offset;
  uint16_t column_count, rowkey_column_count, var_column_count;
  uint8_t row_index_byte, extend_value_bit;
  const uint8_t *col_headers; /* column_count x 16 bytes */
  const uint8_t *meta;        /* encoding meta start: column offsets are relative to it */
  const uint8_t *row_data;
  int64_t row_data_len;
  /* CS_ENCODING_ROW_STORE blocks (cs_encoding/ob_cs_micro_block_transformer.cpp:106-143) */
  uint8_t row_store_type;          /* 1 / 2 PAX, 3 CS */
  uint8_t cs_off_width;            /* bytes per stream end offset */
  uint16_t cs_stream_count;
  const uint8_t *cs_col_headers;   /* column_count x 4 bytes (ObCSColumnHeader) */
  const uint8_t *cs_off_data;      /* stream end offsets, relative to the block start */
  uint32_t cs_first_stream_begin;  /* header + ObAllColumnHeader + column headers */
  uint32_t cs_all_string_offset;
} ora_block;

/* ---- bit stream (encoding/ob_bit_stream.h:169-283) -------------------------------------------- */
uint64_t ora_bs_get(const uint8_t *buf, int64_t offset, int64_t cnt);
uint64_t ora_bs_get_fast(const uint8_t *buf, int64_t offset, int64_t cnt, int64_t bs_len);
void ora_bs_set(uint8_t *buf, int64_t offset, int64_t cnt, uint64_t value);

/* ---- block ------------------------------------------------------------------------------------ */
int ora_block_init(ora_block *blk, const void *buf, int64_t size);
/* data checksum / header checksum verification (ob_micro_block_header.cpp:236-285) */
int ora_block_verify_checksums(const ora_block *blk);
/* building blocks pinned one by one to the compiled reference (tests/test_checksum_ref_kat.py): the payload checksum
 * (ob_crc64_sse42: crc32c, seed as given, no final xor) and the integer-array searches of the RLE / CONST / row-index lookups */
uint64_t ora_crc64_sse42(uint64_t crc, const void *p, int64_t len);
/* macro block headers (fields[28]: see oracle/ref_macro_wrap.cpp) and the walk over its micro blocks */
int ora_macro_block_parse(const uint8_t *buf, int64_t len, int64_t *fields, int32_t verify);
int ora_macro_block_micro_blocks(const uint8_t *buf, int64_t len, int64_t *offs, int64_t *sizes, int32_t cap, int32_t *n_out, int32_t verify);
/* column checksum (K16) of n integer-class cells: sum over the rows of ObDatum::checksum(0) */
int64_t ora_column_checksum(const int64_t *vals, const uint8_t *nulls, int64_t n, int32_t datum_len);
/* BitSet::get_ref (encoding/ob_encoding_bitset.h:68-71): rank of bit `pos` among the set bits, -1 when clear */
int64_t ora_bitset_get_ref(const void *words, int64_t pos);
int64_t ora_int_array_lower_bound(const void *array, int64_t byte, int64_t begin, int64_t end, int64_t key);
int64_t ora_int_array_upper_bound(const void *array, int64_t byte, int64_t begin, int64_t end, int64_t key);
int ora_decode_cell(const ora_block *blk, int32_t col, int64_t row, ora_datum *out);
/* Strings rebuilt by HEX_PACKING / STRING_DIFF / STRING_PREFIX live in a thread-local arena of the oracle until this call
 * (the reference: the decoder's allocator). Pointers the oracle returned for such cells are invalid afterwards. */
void ora_arena_reset(void);
/* dictionary surface of a dictionary-coded column (DICT / RLE / CONST with exceptions / CS INT_DICT / STR_DICT):
 * distinct count, entry `ref` decoded like a cell, refs of rows (NULL / NOP rows: the distinct count) */
int ora_dict_count(const ora_block *blk, int32_t col, int64_t *count);
int ora_dict_entry(const ora_block *blk, int32_t col, int64_t ref, ora_datum *out);
int ora_dict_refs(const ora_block *blk, int32_t col, const int32_t *row_ids, int64_t row_cap, uint32_t *refs);

/* ObMicroBlockDecoder::get_rows into VEC_FIXED (data, nulls as ObBitVector words). */
int ora_get_rows_fixed(const ora_block *blk, int32_t col, const int32_t *row_ids, int64_t row_cap,
                       int64_t vec_offset, void *data, int32_t elem_len, uint64_t *nulls,
                       int32_t *has_null);
/* ... into VEC_DISCRETE (ptrs into the block buffer, lens). */
int ora_get_rows_discrete(const ora_block *blk, int32_t col, const int32_t *row_ids,
                          int64_t row_cap, int64_t vec_offset, const uint8_t **ptrs, int32_t *lens,
                          uint64_t *nulls, int32_t *has_null);

/* ---- filters ---------------------------------------------------------------------------------- */
int ora_filter_white(const ora_block *blk, int32_t col, int32_t op, const ora_param *params,
                     int32_t n_params, int64_t start, int64_t count, uint8_t *bitmap);
int ora_filter_tree(const ora_block *blk, const ora_filter *filter, int64_t start, int64_t count,
                    uint8_t *bitmap);

/* ---- ObBitmap (lib/container/ob_bitmap.cpp) ---------------------------------------------------- */
int ora_bitmap_get_row_ids(const uint8_t *bitmap, int64_t valid_bytes, int32_t *row_ids,
                           int64_t *row_count, int64_t *from, int64_t to, int64_t limit,
                           int64_t id_offset);
int64_t ora_bitmap_popcnt(const uint8_t *bitmap, int64_t n);

/* ---- whole-path scan (reference control structure) ----------------------------- */
typedef struct ora_scan_out {
  /* per projected column, dense over selected rows; any pointer may be NULL to skip storing */
  void **data;          /* n_proj: int columns elem_len-byte values; string columns uint64 ptrs */
  int32_t **lens;       /* n_proj: string columns lens (NULL entries for int columns) */
  uint64_t **nulls;     /* n_proj: ObBitVector words */
  int32_t *has_null;    /* n_proj */
  int32_t *row_ids;     /* block-relative row ids of selected rows (optional) */
  int64_t *sel_offset;  /* n_blocks + 1 (optional) */
  int64_t cap_rows;     /* capacity of the dense arrays in rows */
  uint64_t string_base; /* ptr value = string_base + (cell address - image) */
} ora_scan_out;

/* Scans blocks [block_begin, block_end) of the image with the reference's control structure:
 * per block filter -> ObBitmap -> get_row_ids(limit = batch_size) -> per column get_rows.
 * Dense output starts at out_row_begin. Returns selected row count in *selected. */
int ora_scan_blocks(const void *image, const int64_t *offsets, const int64_t *sizes,
                    int32_t block_begin, int32_t block_end, const ora_filter *filter,
                    const int32_t *proj_cols, int32_t n_proj, int32_t batch_size,
                    int64_t out_row_begin, ora_scan_out *out, int64_t *total_rows,
                    int64_t *selected);

/* Multi-threaded timing harness for the CPU baseline: blocks sharded contiguously over n_threads
 * threads, each thread projecting into its own scratch (sized batch_size) -- i.e. the work of the
 * scan without materialising one dense result. Returns rows scanned / selected and a checksum of
 * every projected value so the work cannot be optimised away. */
int ora_scan_blocks_mt(const void *image, const int64_t *offsets, const int64_t *sizes,
                       int32_t n_blocks, const ora_filter *filter, const int32_t *proj_cols,
                       int32_t n_proj, int32_t batch_size, int32_t n_threads, int64_t *total_rows,
                       int64_t *selected, uint64_t *checksum);

/* ---- major compaction merge (compaction/ob_partition_merger.cpp:678-829 merge_partition ->
 * ObPartitionMergeHelper::find_rowkey_minimum_iters (ob_partition_rows_merger.cpp:815-857) ->
 * ObMergeFuser::fuse_row (ob_partition_merge_fuser.cpp:106-142) -> ObRowFuse::fuse_row
 * (storage/ob_row_fuse.cpp:191-275) -> ObMajorPartitionMergeFuser::end_fuse_row (:284-322) ->
 * ObPartitionMajorMerger::inner_process (:648-676, delete rows dropped)). ---------------------- */
/* all cells of column `col` of a table, in row order: value image + ext (0 value, 1 NULL, 2 NOP) */
int ora_decode_column_ext(const void *image, const int64_t *offsets, const int64_t *sizes, int32_t n_blocks,
                          int32_t col, int64_t *vals, uint8_t *ext, int64_t cap, int64_t *rows);
#define ORA_DF_NOT_EXIST 0
#define ORA_DF_LOCK 1
#define ORA_DF_UPDATE 2
#define ORA_DF_INSERT 3
#define ORA_DF_DELETE 4
typedef struct ora_merge_run {       /* one sorted run (rowkey ascending, unique inside the run) */
  int64_t n;
  const int64_t *key;                /* INT64 rowkey */
  const uint8_t *flag;               /* ObDmlFlag per row; NULL: every row DF_INSERT */
  const int64_t *const *vals;        /* [n_cols][n] */
  const uint8_t *const *ext;         /* [n_cols][n]: 0 value, 1 NULL, 2 NOP */
  const int64_t *const *more_keys;   /* composite rowkey: the columns after `key`, [n_more_keys][n] */
  int32_t n_more_keys;
} ora_merge_run;
/* runs[0] is the OLDEST table, runs[n_runs - 1] the newest (iters are fused newest first).
 * out_null[c][i]: 1 => NULL. stats[0] = keys dropped because the fused row is a delete,
 * stats[1] = output rows fused from more than one run. */
int ora_major_merge(const ora_merge_run *runs, int32_t n_runs, int32_t n_cols, const int64_t *default_vals,
                    const uint8_t *default_null, int64_t out_cap, int64_t *out_key, int64_t *const *out_vals,
                    uint8_t *const *out_null, int64_t *out_rows, int64_t *stats);
/* Same with the extra rowkey columns of a composite rowkey written to out_more_keys[n_more_keys][..]. */
int ora_major_merge_keys(const ora_merge_run *runs, int32_t n_runs, int32_t n_cols, const int64_t *default_vals,
                         const uint8_t *default_null, int64_t out_cap, int64_t *out_key, int64_t *const *out_more_keys,
                         int64_t *const *out_vals, uint8_t *const *out_null, int64_t *out_rows, int64_t *stats);

/* ---- skip index (pre-aggregated min / max / null count per micro-block) ------------------------------------
 * ObAggRowReader::read (index_block/ob_agg_row_struct.cpp:339-482): aggregate `col_type` (ObSkipIndexColType)
 * of column `col_idx` inside a serialized aggregate row; *data == NULL when it is not stored (NULL datum). */
int ora_agg_row_read(const void *buf, int64_t size, uint32_t col_idx, int32_t col_type, const uint8_t **data,
                     int32_t *len, int32_t *is_prefix);
/* ObBoolMaskType: verdict of a filter over a block */
enum { ORA_MASK_UNCERTAIN = 0, ORA_MASK_ALWAYS_TRUE = 1, ORA_MASK_ALWAYS_FALSE = 2 };
/* ObSkipIndexFilterExecutor::falsifiable_pushdown_filter -> filter_on_min_max for one white leaf
 * (ob_skip_index_filter_executor.cpp:114-189, 250-396, 498-822); binary collation for strings. */
int ora_skip_index_leaf(const void *agg, int64_t agg_size, int64_t row_count, int32_t col, uint8_t obj_type, int32_t op,
                        const ora_param *params, int32_t n_params, int32_t *mask);
/* ObSSTableIndexFilter::check_range: every leaf, then ObPushdownFilterExecutor::execute_skipping_filter
 * (sql/engine/basic/ob_pushdown_filter.cpp:1707-1740). col_types[c] = obj type of column store index c. */
int ora_skip_index_filter(const void *agg, int64_t agg_size, int64_t row_count, const uint8_t *col_types, int32_t n_cols,
                          const ora_filter *filter, int32_t *mask);

#ifdef __cplusplus
}
#endif
/* ---- integer stream codecs of CS blocks (ob_stream_codecs.c): ObIntegerStreamDecoder::decode_body -------------------
 * type = ObIntegerStream::EncodingType (1 RAW, 2 DOUBLE_DELTA_ZIGZAG_RLE, 3 DOUBLE_DELTA_ZIGZAG_PFOR, 4 DELTA_ZIGZAG_RLE,
 * 5 DELTA_ZIGZAG_PFOR, 6 SIMD_FIXEDPFOR, 8 XOR_FIXED_PFOR); decodes `count` values of width_bytes each from the
 * codec's bytes (stream data WITHOUT the ObIntegerStreamMeta) into out[count]. */
int ora_int_stream_decode(int32_t type, uint32_t width_bytes, const void *in, int64_t in_len, int64_t count, void *out,
                          int64_t *consumed);

/* CS block -> the same block with every integer stream restated as RAW (what ObCSMicroBlockTransformer::full_transform
 * achieves at cache fill, kept in the on-disk layout; see ob_stream_codecs.c). out == NULL: only *out_size. */
int ora_cs_transform(const void *block, int64_t size, void *out, int64_t out_cap, int64_t *out_size);

#endif
