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
