// Stand-in for storage/ob_i_store.h: what ObMicroBlockHeader (blocksstable/ob_micro_block_header.{h,cpp}) needs from the storage layer --
// the row store type enum with the reference's values (common/ob_store_format.h:31-40), the header magic
// (ob_block_sstable_struct.h:48), the extra rowkey column count of multi-version rows (ob_i_store.h:186: trans version + sql
// sequence) -- and the REAL format_i32 / format_i64 folding helpers (common/ob_record_header.h, included from the reference).
#pragma once
#include "ob_macro_shim.h"
#include "lib/checksum/ob_crc64.h"
#include "common/ob_record_header.h"
#define STATIC_ASSERT(cond, msg) static_assert(cond, msg)
#define LOG_DBA_ERROR(...) ((void)0)
#define LOG_DBA_ERROR_V2(...) ((void)0)
namespace oceanbase {
namespace common {
enum ObRowStoreType : uint8_t {
  FLAT_ROW_STORE = 0, ENCODING_ROW_STORE = 1, SELECTIVE_ENCODING_ROW_STORE = 2, CS_ENCODING_ROW_STORE = 3, FLAT_OPT_ROW_STORE = 4,
  MAX_ROW_STORE, DUMMY_ROW_STORE = UINT8_MAX
};
constexpr int OB_PHYSIC_CHECKSUM_ERROR = -4108;   // lib/ob_errno.h:80
}  // namespace common
namespace storage {
struct ObMultiVersionRowkeyHelpper {
  constexpr static int get_extra_rowkey_col_cnt() { return 2; }
};
}  // namespace storage
namespace blocksstable {
const int16_t MICRO_BLOCK_HEADER_MAGIC = 1005;   // ob_block_sstable_struct.h:48
}
}  // namespace oceanbase
