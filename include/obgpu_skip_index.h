/* obgpu_skip_index.h -- C-ABI of the B200-native skip-index (min / max / null-count) block pruning.
 *
 * Drop-in boundary for the reference's pre-aggregated index filter
 *   storage/access/ob_sstable_index_filter.cpp:56-108                ObSSTableIndexFilter::check_range
 *   storage/blocksstable/index_block/ob_skip_index_filter_executor.cpp
 *       :114-189  falsifiable_pushdown_filter   :250-396  filter_on_min_max
 *       :498-822  compare / eq .. bt / in operators (prefix rules :398-496)
 *   sql/engine/basic/ob_pushdown_filter.cpp:1707-1740                execute_skipping_filter (AND / OR of ObBoolMask)
 *   storage/blocksstable/index_block/ob_agg_row_struct.{h,cpp}       ObAggRowHeader / ObAggRowWriter / ObAggRowReader
 * The reference walks the index tree on the CPU and asks, per micro-block index row, whether the pushed-down
 * filter is always false (skip the block), always true (every row passes, no filter evaluation) or uncertain.
 * Here the serialized aggregate rows of a page batch's micro-blocks (ObMicroIndexInfo::agg_row_buf_ /
 * agg_buf_size_, exactly the bytes the index block stores) travel with the batch; one kernel parses them and
 * evaluates the whole filter tree for every block, and obgpu_scan honours the verdicts: an always-false block
 * is never read, an always-true block is not filtered.
 *
 * Same conventions as obgpu_scan.h: int OB codes, no exceptions, caller-owned outputs. */
#ifndef OBGPU_SKIP_INDEX_H_
#define OBGPU_SKIP_INDEX_H_

#include "obgpu_scan.h"

#ifdef __cplusplus
extern "C" {
#endif

/* blocksstable::ObSkipIndexColType (index_block/ob_index_block_util.h:40-50) */
enum {
  OBGPU_SK_IDX_MIN = 0,
  OBGPU_SK_IDX_MAX = 1,
  OBGPU_SK_IDX_NULL_COUNT = 2,
  OBGPU_SK_IDX_SUM = 3,
  OBGPU_SK_IDX_BM25_MAX_SCORE_TOKEN_FREQ = 4,
  OBGPU_SK_IDX_BM25_MAX_SCORE_DOC_LEN = 5,
  OBGPU_SK_IDX_MAX_COL_TYPE = 6
};

/* sql::ObBoolMaskType (sql/engine/basic/ob_pushdown_filter.h): verdict of a filter over one block */
enum {
  OBGPU_BOOL_MASK_UNCERTAIN = 0,    /* PROBABILISTIC */
  OBGPU_BOOL_MASK_ALWAYS_TRUE = 1,
  OBGPU_BOOL_MASK_ALWAYS_FALSE = 2
};

#define OBGPU_SKIP_INDEX_MAX_COL_LENGTH 40 /* ObSkipIndexColMeta::MAX_SKIP_INDEX_COL_LENGTH */

/* One aggregate of an aggregate row: (ObSkipIndexColMeta, ObStorageDatum, is_min_max_prefix). */
typedef struct obgpu_agg_cell {
  uint32_t col_idx;  /* column store index the aggregate refers to                      */
  uint8_t col_type;  /* OBGPU_SK_IDX_*                                                  */
  uint8_t is_null;   /* NULL / NOP datum: the aggregate is not stored                   */
  uint8_t is_prefix; /* MIN / MAX of a string longer than 40 bytes: only a prefix kept  */
  uint8_t reserved;
  int32_t len;       /* datum length in bytes                                           */
  const void *data;  /* datum bytes                                                     */
} obgpu_agg_cell;

/* ObAggRowWriter::init + write_agg_data (ob_agg_row_struct.cpp:49-300): serializes the cells (any order)
 * as one aggregate row. version: 1, 2 (prefix bitmap) or 3 (revised max prefix; the current one).
 * out == NULL: only *out_size is computed. */
int obgpu_agg_row_write(const obgpu_agg_cell *cells, int32_t n_cells, int32_t version, void *out,
                        int64_t out_cap, int64_t *out_size);

/* ObSkipIndexAggregator over rows [row_begin, row_begin + nrows) of the writer's column inputs
 * (index_block/ob_index_block_aggregator.cpp): MIN / MAX / NULL_COUNT of every column listed in agg_cols,
 * serialized as one version-3 aggregate row. Strings longer than 40 bytes keep a 40-byte prefix. */
int obgpu_writer_block_agg_row(const obgpu_col_input *cols, int32_t n_cols, const int32_t *agg_cols,
                               int32_t n_agg_cols, int64_t row_begin, int64_t nrows, void *out,
                               int64_t out_cap, int64_t *out_size);
/* One aggregate row per block of obgpu_writer_encode_table's blocking: row b occupies
 * [offsets[b], offsets[b + 1]) of `out` (n_blocks + 1 offsets). out == NULL: only *out_size. */
int obgpu_writer_table_agg_rows(const obgpu_col_input *cols, int32_t n_cols, const int32_t *agg_cols,
                                int32_t n_agg_cols, int64_t total_rows, int64_t rows_per_block, void *out,
                                int64_t out_cap, int64_t *offsets, int64_t *out_size);

/* Attaches the aggregate rows of the batch's micro-blocks (host buffers; copied to the device on the ctx
 * stream): block b's row is agg_rows[agg_off[b] .. agg_off[b + 1]); an empty range means "no aggregate
 * data" (ObMicroIndexInfo::has_agg_data() false: every filter is uncertain on that block). From then on
 * obgpu_scan prunes with them. Passing NULL detaches. */
int obgpu_batch_set_agg_rows(obgpu_batch *batch, const void *agg_rows, const int64_t *agg_off);

/* ObSSTableIndexFilter::check_range for every block of the batch: block_mask[b] = OBGPU_BOOL_MASK_* of the
 * whole filter tree on block b (ObMicroIndexInfo::set_filter_constant_type). Synchronises the ctx stream. */
int obgpu_batch_skip_index_filter(obgpu_batch *batch, const obgpu_filter *filter, uint8_t *block_mask);

/* Blocks of the scan the skip index decided (valid after obgpu_result_info_get). */
int obgpu_result_skip_info(obgpu_result *res, int64_t *always_false_blocks, int64_t *always_true_blocks);

#ifdef __cplusplus
}
#endif
#endif /* OBGPU_SKIP_INDEX_H_ */
