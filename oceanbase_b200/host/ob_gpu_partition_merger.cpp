#include "ob_gpu_partition_merger.h"

#include <algorithm>

namespace oceanbase {
namespace compaction {
using namespace common;

ObGpuPartitionMajorMerger::~ObGpuPartitionMajorMerger() { reset(); }

void ObGpuPartitionMajorMerger::reset() {
  if (result_) obgpu_merge_result_free(result_);
  result_ = nullptr;
  for (obgpu_batch *b : batches_) obgpu_batch_close(b);
  batches_.clear();
  if (ctx_) obgpu_ctx_destroy(ctx_);
  ctx_ = nullptr;
  merged_ = false;
  cursor_ = 0;
  info_ = obgpu_merge_info{};
}

int ObGpuPartitionMajorMerger::init(int device, const std::vector<ObGpuMergeTable> &tables, const ObGpuMergeSchema &schema) {
  int ret = OB_SUCCESS;
  if (ctx_ != nullptr) {
    ret = OB_INIT_TWICE;
  } else if (tables.empty() || (!schema.default_vals_.empty() && schema.default_vals_.size() != schema.payload_cols_.size()) ||
             (!schema.default_null_.empty() && schema.default_null_.size() != schema.payload_cols_.size()) ||
             (!schema.payload_is_string_.empty() && schema.payload_is_string_.size() != schema.payload_cols_.size()) ||
             schema.more_rowkey_cols_.size() + 1 > OBGPU_MERGE_MAX_KEY_COLS) {
    ret = OB_INVALID_ARGUMENT;
  } else if (OB_SUCCESS != obgpu_ctx_create(device, &ctx_)) {
    ctx_ = nullptr;
    ret = OB_NOT_SUPPORTED;  // no device: the caller keeps ObPartitionMajorMerger
  } else {
    schema_ = schema;
    for (size_t i = 0; OB_SUCCESS == ret && i < tables.size(); ++i) {
      const ObGpuMergeTable &t = tables[i];
      obgpu_batch *b = nullptr;
      ret = obgpu_batch_open(ctx_, t.image_, t.image_size_, t.offsets_, t.sizes_, t.block_count_, 0, nullptr, &b);
      if (OB_SUCCESS == ret) batches_.push_back(b);
    }
    if (OB_SUCCESS != ret) reset();
  }
  return ret;
}

int ObGpuPartitionMajorMerger::merge_partition() {
  int ret = OB_SUCCESS;
  if (ctx_ == nullptr) {
    ret = OB_NOT_INIT;
  } else if (merged_) {
    ret = OB_ERR_UNEXPECTED;
  } else {
    std::vector<int32_t> keys(1, schema_.rowkey_col_);
    keys.insert(keys.end(), schema_.more_rowkey_cols_.begin(), schema_.more_rowkey_cols_.end());
    ret = obgpu_merge_runs_keys(ctx_, batches_.data(), (int32_t)batches_.size(), keys.data(), (int32_t)keys.size(), schema_.flag_col_,
                                schema_.payload_cols_.data(), (int32_t)schema_.payload_cols_.size(),
                                schema_.default_vals_.empty() ? nullptr : schema_.default_vals_.data(),
                                schema_.default_null_.empty() ? nullptr : schema_.default_null_.data(), &result_);
    if (OB_SUCCESS == ret) ret = obgpu_merge_result_info(result_, &info_);
    if (OB_SUCCESS == ret) merged_ = true;
  }
  return ret;
}

int ObGpuPartitionMajorMerger::get_next_rows(int64_t max_rows, ObGpuMergedRows &rows) {
  int ret = OB_SUCCESS;
  if (!merged_) {
    ret = OB_NOT_INIT;
  } else if (max_rows <= 0) {
    ret = OB_INVALID_ARGUMENT;
  } else if (cursor_ >= info_.out_rows) {
    rows.row_count_ = 0;
    ret = OB_ITER_END;
  } else {
    const int64_t n = std::min(max_rows, info_.out_rows - cursor_);
    const size_t nc = schema_.payload_cols_.size();
    rows.row_count_ = n;
    rows.rowkeys_.resize((size_t)n);
    rows.values_.resize(nc);
    rows.nulls_.resize(nc);
    rows.heap_.assign(nc, {});
    rows.offsets_.assign(nc, {});
    rows.more_rowkeys_.resize(schema_.more_rowkey_cols_.size());
    ret = obgpu_merge_result_fetch(result_, -1, cursor_, n, rows.rowkeys_.data(), nullptr);
    for (size_t k = 0; OB_SUCCESS == ret && k < rows.more_rowkeys_.size(); ++k) {
      rows.more_rowkeys_[k].resize((size_t)n);
      ret = obgpu_merge_result_fetch(result_, -2 - (int32_t)k, cursor_, n, rows.more_rowkeys_[k].data(), nullptr);
    }
    for (size_t c = 0; OB_SUCCESS == ret && c < nc; ++c) {
      rows.values_[c].assign((size_t)n, 0);
      rows.nulls_[c].resize((size_t)n);
      if (!schema_.payload_is_string_.empty() && schema_.payload_is_string_[c]) {
        // references -> bytes for this window only (the writer consumes ObString cells)
        rows.offsets_[c].resize((size_t)n + 1);
        int64_t need = 0;
        ret = obgpu_merge_result_fetch_strings(result_, (int32_t)c, cursor_, n, nullptr, 0, rows.offsets_[c].data(),
                                               rows.nulls_[c].data(), &need);
        if (OBGPU_BUF_NOT_ENOUGH == ret || (OB_SUCCESS == ret && need > 0)) {
          rows.heap_[c].resize((size_t)need);
          ret = obgpu_merge_result_fetch_strings(result_, (int32_t)c, cursor_, n, rows.heap_[c].data(), need,
                                                 rows.offsets_[c].data(), rows.nulls_[c].data(), &need);
        }
      } else {
        ret = obgpu_merge_result_fetch(result_, (int32_t)c, cursor_, n, rows.values_[c].data(), rows.nulls_[c].data());
      }
    }
    if (OB_SUCCESS == ret) cursor_ += n;
  }
  return ret;
}

}  // namespace compaction
}  // namespace oceanbase
