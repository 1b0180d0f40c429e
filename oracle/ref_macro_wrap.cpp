#include "ob_macro_shim.h"
// extern "C" doors onto the REFERENCE's macro-block headers, compiled from their own sources under /root/reference (not copied),
// checker only:
//   * ObMacroBlockCommonHeader  (storage/blocksstable/ob_macro_block_common_header.{h,cpp}): 24 bytes, serialize / deserialize /
//     check_integrity
//   * ObSSTableMacroBlockHeader (storage/blocksstable/ob_sstable_macro_block_header.{h,cpp}): FixedHeader + column type / order /
//     checksum arrays + is_normal_cg_, init / serialize / deserialize / is_valid
//   * ObMicroBlockHeader (storage/blocksstable/ob_micro_block_header.{h,cpp}): the 64-byte micro-block header -- init / set_header_checksum /
//     serialize, and deserialize_and_check_record = is_valid + header checksum (format_i32 / format_i64 of common/ob_record_header.h,
//     the real ones) + payload checksum over a whole micro-block
//   * ob_crc64 (deps/oblib/src/lib/checksum/ob_crc64.cpp), the payload checksum of the common header (ob_macro_block.cpp:516):
//     crc32c again (ob_crc64 -> ob_crc64_sse42)
// tests/test_macro_block_kat.py pins the writer's macro blocks and the oracle's parser to them.
#include "lib/checksum/ob_crc64.h"
#include "storage/blocksstable/ob_macro_block_common_header.h"
#include "storage/blocksstable/ob_data_store_desc.h"
#include "storage/blocksstable/ob_micro_block_header.h"
#define private public
#include "storage/blocksstable/ob_sstable_macro_block_header.h"
#undef private

using namespace oceanbase;
using namespace oceanbase::common;
using namespace oceanbase::blocksstable;

extern "C" {
// ob_crc64_sse42's vendor dispatch (ob_crc64.cpp:1113-1140) selects ISA-L's crc32_iscsi on Intel CPUs. ISA-L is an external assembly
// library that is not in the tree; its crc32_iscsi(buf, len, init) is crc32c over buf starting from init without inversions, i.e. the
// function the in-tree crc64_sse42 (crc32 instruction) computes -- tests/test_checksum_ref_kat.py holds the three in-tree versions equal.
unsigned int crc32_iscsi(unsigned char *buf, int len, unsigned int init) { return (unsigned int)crc64_sse42(init, (const char *)buf, len); }

// ob_crc64(pv, cb) (ob_crc64.cpp:348-364) is ob_crc64_sse42(0, pv, cb): the vendor dispatch picks ISA-L's crc32_iscsi, the crc32
// instruction or the table version -- one function, crc32c with seed 0 and no final xor (tests/test_checksum_ref_kat.py holds the
// three in-tree implementations equal). The door calls the instruction version directly instead of going through cpuid.
uint64_t ref_ob_crc64(const void *buf, int64_t len) { return crc64_sse42(0, static_cast<const char *>(buf), len); }

int64_t ref_macro_common_header_size() { return ObMacroBlockCommonHeader::get_serialize_size(); }
int64_t ref_macro_fixed_header_size() { return ObSSTableMacroBlockHeader::get_fixed_header_size(); }

// Serialized [common header][sstable macro block header] of a data macro block, the way ObMacroBlock::reserve_header /
// write_micro_block / write_macro_header fill them (ob_macro_block.cpp:455-520,264-303). col_types: n_type_cols x 4 bytes
// (ObObjMeta), col_orders: n_type_cols ints. Returns the bytes written, < 0 on error.
int64_t ref_macro_headers_build(uint16_t version, uint64_t tablet_id, int64_t logical_version, int64_t data_seq, int32_t column_count,
                                int32_t rowkey_column_count, int32_t row_store_type, int32_t row_count, int32_t micro_block_count,
                                int32_t micro_block_data_size, int32_t occupy_size, int64_t data_checksum, int32_t is_cg,
                                const uint8_t *col_types, const int32_t *col_orders, const int64_t *col_checksums,
                                int32_t payload_size, int32_t payload_checksum, char *out, int64_t cap) {
  const int64_t n_type_cols = version == 2 ? rowkey_column_count : column_count;
  share::schema::ObColDesc descs[256];
  if (n_type_cols > 256) return -1;
  for (int64_t i = 0; i < n_type_cols; ++i) {
    memcpy(&descs[i].col_type_, col_types + 4 * i, 4);
    descs[i].col_order_ = (ObOrderType)col_orders[i];
  }
  ObDataStoreDesc desc;
  desc.version_ = version;
  desc.column_count_ = column_count;
  desc.rowkey_column_count_ = rowkey_column_count;
  desc.tablet_id_ = tablet_id;
  desc.logical_version_ = logical_version;
  desc.row_store_type_ = row_store_type;
  desc.is_cg_ = is_cg != 0;
  desc.cols_ = ObIArray<share::schema::ObColDesc>(descs, n_type_cols);
  ObMacroBlockCommonHeader common;
  if (common.set_attr(ObMacroBlockCommonHeader::SSTableData) != OB_SUCCESS) return -2;
  const int64_t chs = common.get_serialize_size();
  ObSSTableMacroBlockHeader mh;
  char *var = out + chs + mh.get_fixed_header_size();
  ObObjMeta *types = reinterpret_cast<ObObjMeta *>(var);
  ObOrderType *orders = reinterpret_cast<ObOrderType *>(var + sizeof(ObObjMeta) * n_type_cols);
  int64_t *cks = reinterpret_cast<int64_t *>(var + (sizeof(ObObjMeta) + sizeof(ObOrderType)) * n_type_cols);
  if (chs + mh.get_fixed_header_size() + (int64_t)(sizeof(ObObjMeta) + sizeof(ObOrderType)) * n_type_cols + 8 * (int64_t)column_count + 1 > cap) return -3;
  memset(out + chs, 0, (size_t)mh.get_fixed_header_size());
  if (mh.init(desc, types, orders, cks) != OB_SUCCESS) return -4;
  for (int32_t i = 0; i < column_count; ++i) cks[i] = col_checksums ? col_checksums[i] : 0;
  mh.fixed_header_.data_seq_ = data_seq;
  mh.fixed_header_.row_count_ = row_count;
  mh.fixed_header_.micro_block_count_ = micro_block_count;
  mh.fixed_header_.micro_block_data_size_ = micro_block_data_size;
  mh.fixed_header_.occupy_size_ = occupy_size;
  mh.fixed_header_.data_checksum_ = data_checksum;
  int64_t pos = 0;
  if (mh.serialize(out + chs, cap - chs, pos) != OB_SUCCESS) return -5;
  common.set_payload_size(payload_size);
  common.set_payload_checksum(payload_checksum);
  if (common.build_serialized_header(out, cap) != OB_SUCCESS) return -6;
  return chs + pos;
}

// The reference's own integrity check of one whole micro-block (header + payload): ObMicroBlockHeader::deserialize_and_check_record
// (ob_micro_block_header.cpp:279-369: deserialize, is_valid, check_header_checksum, check_payload_checksum). Returns an OB code.
int ref_micro_block_check(const char *block, int64_t size) {
  return ObMicroBlockHeader::deserialize_and_check_record(block, size, MICRO_BLOCK_HEADER_MAGIC);
}

// A micro-block header built by the reference: init (version 3, magic, header size), the layout facts of the block, then
// set_header_checksum and serialize. flag16 carries all_lob_in_row_ etc. as the writer sets them. Returns bytes written (< 0: error).
int64_t ref_micro_header_build(int32_t column_count, int32_t rowkey_column_count, int32_t row_store_type, uint16_t flag16, uint32_t row_count,
                               uint8_t opt, uint16_t opt2, uint32_t row_offset, int32_t original_length, int64_t max_merged_trans_version,
                               int32_t data_length, int32_t data_zlength, int64_t data_checksum, char *out, int64_t cap) {
  ObMicroBlockHeader h;
  if (h.init(0, column_count, rowkey_column_count, (ObRowStoreType)row_store_type, false) != OB_SUCCESS) return -1;
  h.flag16_ = flag16;
  h.row_count_ = row_count;
  h.opt_ = opt;
  h.opt2_ = opt2;
  h.row_offset_ = row_offset;
  h.original_length_ = original_length;
  h.max_merged_trans_version_ = max_merged_trans_version;
  h.data_length_ = data_length;
  h.data_zlength_ = data_zlength;
  h.data_checksum_ = data_checksum;
  h.set_header_checksum();
  int64_t pos = 0;
  if (h.serialize(out, cap, pos) != OB_SUCCESS) return -2;
  return pos;
}

// Deserialises both headers with the reference's code; fields[]: 0 common header_size, 1 version, 2 magic, 3 attr, 4 payload_size,
// 5 payload_checksum, 6 macro header_size, 7 version, 8 magic, 9 tablet_id, 10 logical_version, 11 data_seq, 12 column_count,
// 13 rowkey_column_count, 14 row_store_type, 15 row_count, 16 occupy_size, 17 micro_block_count, 18 micro_block_data_offset,
// 19 micro_block_data_size, 20 idx_block_offset, 21 idx_block_size, 22 meta_block_offset, 23 meta_block_size, 24 data_checksum,
// 25 compressor_type, 26 is_normal_cg, 27 offset of the column checksum array from the block start. Returns an OB code.
int ref_macro_headers_parse(const char *buf, int64_t len, int64_t *fields) {
  ObMacroBlockCommonHeader common;
  int64_t pos = 0;
  int ret = common.deserialize(buf, len, pos);
  if (ret != OB_SUCCESS) return ret;
  if ((ret = common.check_integrity()) != OB_SUCCESS) return ret;
  fields[0] = common.get_header_size(); fields[1] = common.get_version(); fields[2] = common.get_magic(); fields[3] = common.get_attr();
  fields[4] = common.get_payload_size(); fields[5] = common.get_payload_checksum();
  ObSSTableMacroBlockHeader mh;
  if ((ret = mh.deserialize(buf, len, pos)) != OB_SUCCESS) return ret;
  const auto &f = mh.fixed_header_;
  fields[6] = f.header_size_; fields[7] = f.version_; fields[8] = f.magic_; fields[9] = (int64_t)f.tablet_id_; fields[10] = f.logical_version_;
  fields[11] = f.data_seq_; fields[12] = f.column_count_; fields[13] = f.rowkey_column_count_; fields[14] = f.row_store_type_;
  fields[15] = f.row_count_; fields[16] = f.occupy_size_; fields[17] = f.micro_block_count_; fields[18] = f.micro_block_data_offset_;
  fields[19] = f.micro_block_data_size_; fields[20] = f.idx_block_offset_; fields[21] = f.idx_block_size_; fields[22] = f.meta_block_offset_;
  fields[23] = f.meta_block_size_; fields[24] = f.data_checksum_; fields[25] = (int64_t)f.compressor_type_; fields[26] = mh.is_normal_cg_;
  fields[27] = reinterpret_cast<const char *>(mh.column_checksum_) - buf;
  return mh.is_valid() ? OB_SUCCESS : OB_INVALID_DATA;
}
}
