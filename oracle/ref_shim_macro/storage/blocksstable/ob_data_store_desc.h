// Stand-in for ObDataStoreDesc: exactly the getters ObSSTableMacroBlockHeader::init reads (ob_sstable_macro_block_header.cpp:299-360).
#pragma once
#include "ob_macro_shim.h"
#include "share/ob_encryption_util.h"
namespace oceanbase { namespace blocksstable {
class ObDataStoreDesc {
public:
  bool is_valid() const { return true; }
  uint16_t get_fixed_header_version() const { return version_; }
  int64_t get_row_column_count() const { return column_count_; }
  int64_t get_rowkey_column_count() const { return rowkey_column_count_; }
  common::ObTabletID get_tablet_id() const { return common::ObTabletID(tablet_id_); }
  int64_t get_logical_version() const { return logical_version_; }
  int32_t get_row_store_type() const { return row_store_type_; }
  int64_t get_encrypt_id() const { return 0; }
  int64_t get_master_key_id() const { return 0; }
  const char *get_encrypt_key() const { return key_; }
  int64_t get_encrypt_key_size() const { return share::OB_MAX_TABLESPACE_ENCRYPT_KEY_LENGTH; }
  common::ObCompressorType get_compressor_type() const { return common::NONE_COMPRESSOR; }
  bool is_major_merge_type() const { return major_; }
  const common::ObIArray<share::schema::ObColDesc> &get_full_stored_col_descs() const { return cols_; }
  const common::ObIArray<share::schema::ObColDesc> &get_rowkey_col_descs() const { return cols_; }
  bool is_cg() const { return is_cg_; }
  uint16_t version_ = 1;
  int64_t column_count_ = 0, rowkey_column_count_ = 0, logical_version_ = 0;
  uint64_t tablet_id_ = 0;
  int32_t row_store_type_ = 0;
  bool major_ = true, is_cg_ = false;
  char key_[16] = {0};
  common::ObIArray<share::schema::ObColDesc> cols_;
};
} }
