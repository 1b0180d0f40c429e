// C++ parity test of the host adapter (oceanbase_b200/host), written in the shape of the reference's
// own decoder tests: unittest/storage/blocksstable/encoding/test_raw_decoder.cpp:774-1200 (filter
// popcounts over [seedA .. | seedB x 10 | NULL x 10] blocks, whole block and pd_filter_info windows)
// and test_micro_block_decoder.cpp:155-189 (every cell decodes back). The oracle
// (oracle/libob_oracle.so) is the checker; the adapter runs on the GPU through the C-ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../oceanbase_b200/host/ob_gpu_micro_block_decoder.h"
extern "C" {
#include "../../include/obgpu_writer.h"
#include "../../oracle/ob_oracle.h"
}

using namespace oceanbase;
using namespace oceanbase::common;
using namespace oceanbase::blocksstable;

static int g_fail = 0;
#define ASSERT_EQ(a, b)                                                                           \
  do {                                                                                            \
    const long long va__ = (long long)(a), vb__ = (long long)(b);                                 \
    if (va__ != vb__) {                                                                           \
      printf("FAIL %s:%d  %s = %lld, expected %lld\n", __FILE__, __LINE__, #a, va__, vb__);       \
      ++g_fail;                                                                                   \
    }                                                                                             \
  } while (0)

static const int64_t ROW_CNT = 64;

struct Rows {
  std::vector<int64_t> ints;
  std::vector<std::string> strs;
  std::vector<uint8_t> nulls;
};

static int64_t seed_int(int64_t seed) { return seed * 1000 + 7; }
static std::string seed_str(int64_t seed) { char b[32]; snprintf(b, sizeof(b), "seed-%04lld", (long long)seed); return b; }

static std::vector<uint8_t> build_block(const Rows &r, int enc_int, int enc_str) {
  const int64_t n = (int64_t)r.ints.size();
  std::vector<int64_t> pad(n);
  std::iota(pad.begin(), pad.end(), 0);
  std::string heap;
  std::vector<int64_t> off(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) { heap += r.strs[i]; off[i + 1] = (int64_t)heap.size(); }
  heap.push_back('\0');
  obgpu_col_input cols[3];
  memset(cols, 0, sizeof(cols));
  cols[0].obj_type = OBGPU_OBJ_INT; cols[0].encoding = OBGPU_ENC_RAW; cols[0].i64 = pad.data();
  cols[1].obj_type = OBGPU_OBJ_INT; cols[1].encoding = enc_int; cols[1].i64 = r.ints.data(); cols[1].is_null = r.nulls.data();
  cols[2].obj_type = OBGPU_OBJ_VARCHAR; cols[2].encoding = enc_str; cols[2].str_heap = heap.data(); cols[2].str_off = off.data();
  cols[2].is_null = r.nulls.data();
  int64_t size = 0;
  if (obgpu_writer_encode_block(cols, 3, 1, 0, n, nullptr, 0, &size) != 0) { printf("encode failed\n"); exit(2); }
  std::vector<uint8_t> blk((size_t)size);
  obgpu_writer_encode_block(cols, 3, 1, 0, n, blk.data(), size, &size);
  return blk;
}

static Rows layout(std::initializer_list<std::pair<int64_t, int64_t>> parts) {  // (seed or -1 = NULL, count)
  Rows r;
  for (auto &p : parts)
    for (int64_t i = 0; i < p.second; ++i) {
      r.nulls.push_back(p.first < 0);
      r.ints.push_back(seed_int(p.first < 0 ? 0 : p.first));
      r.strs.push_back(seed_str(p.first < 0 ? 0 : p.first));
    }
  return r;
}

static int64_t pushdown_popcnt(ObGpuMicroBlockDecoder &dec, int col, sql::ObWhiteFilterOperatorType op,
                               std::vector<int64_t> seeds, bool str, int64_t start, int64_t count) {
  sql::ObWhiteFilterExecutor filter(col, op);
  std::vector<std::string> keep;
  keep.reserve(seeds.size());
  for (int64_t s : seeds) {
    ObStorageDatum d;
    if (str) { keep.push_back(seed_str(s)); d.set_string(keep.back().data(), (uint32_t)keep.back().size()); }
    else d.set_int(seed_int(s));
    filter.get_datums().push_back(d);
  }
  sql::PushdownFilterInfo pd;
  pd.start_ = start;
  pd.count_ = count;
  ObBitmap bm;
  bm.init(count);
  ASSERT_EQ(OB_SUCCESS, dec.filter_pushdown_filter(nullptr, filter, pd, bm));
  return (int64_t)bm.popcnt();
}

static void test_filter_pushdown(ObGpuScanRuntime &rt) {
  const int encs[][2] = {{OBGPU_ENC_RAW, OBGPU_ENC_RAW}, {OBGPU_ENC_DICT, OBGPU_ENC_DICT},
                         {OBGPU_ENC_RLE, OBGPU_ENC_RLE}, {OBGPU_ENC_INTEGER_BASE_DIFF, OBGPU_ENC_DICT}};
  for (auto &e : encs) {
    for (int str = 0; str < 2; ++str) {
      const int col = str ? 2 : 1;
      {  // filter_pushdown_all_eq_ne: [seed1 x N-20 | seed2 x 10 | NULL x 10]
        std::vector<uint8_t> blk = build_block(layout({{0xF, ROW_CNT - 20}, {0x0, 10}, {-1, 10}}), e[0], e[1]);
        ObGpuMicroBlockDecoder dec(rt);
        ObMicroBlockData data{(const char *)blk.data(), (int64_t)blk.size()};
        ASSERT_EQ(OB_SUCCESS, dec.init(data));
        int64_t rc = 0;
        ASSERT_EQ(OB_SUCCESS, dec.get_row_count(rc));
        ASSERT_EQ(ROW_CNT, rc);
        ASSERT_EQ(ROW_CNT - 20, pushdown_popcnt(dec, col, sql::WHITE_OP_EQ, {0xF}, str, 0, ROW_CNT));
        ASSERT_EQ(25, pushdown_popcnt(dec, col, sql::WHITE_OP_EQ, {0xF}, str, ROW_CNT - 45, 30));
        ASSERT_EQ(10, pushdown_popcnt(dec, col, sql::WHITE_OP_NE, {0xF}, str, 0, ROW_CNT));
        ASSERT_EQ(5, pushdown_popcnt(dec, col, sql::WHITE_OP_NE, {0xF}, str, ROW_CNT - 45, 30));
        ASSERT_EQ(OB_SUCCESS, dec.init(data));  // can be init twice
      }
      {  // filter_push_down_gt_lt_ge_le: [seed0 x N-30 | seed1 x 10 | seed2 x 10 | NULL x 10]
        std::vector<uint8_t> blk = build_block(layout({{0, ROW_CNT - 30}, {1, 10}, {2, 10}, {-1, 10}}), e[0], e[1]);
        ObGpuMicroBlockDecoder dec(rt);
        ObMicroBlockData data{(const char *)blk.data(), (int64_t)blk.size()};
        ASSERT_EQ(OB_SUCCESS, dec.init(data));
        ASSERT_EQ(10, pushdown_popcnt(dec, col, sql::WHITE_OP_GT, {1}, str, 0, ROW_CNT));
        ASSERT_EQ(5, pushdown_popcnt(dec, col, sql::WHITE_OP_GT, {1}, str, ROW_CNT - 45, 30));
        ASSERT_EQ(ROW_CNT - 30, pushdown_popcnt(dec, col, sql::WHITE_OP_LT, {1}, str, 0, ROW_CNT));
        ASSERT_EQ(15, pushdown_popcnt(dec, col, sql::WHITE_OP_LT, {1}, str, ROW_CNT - 45, 30));
        ASSERT_EQ(20, pushdown_popcnt(dec, col, sql::WHITE_OP_GE, {1}, str, 0, ROW_CNT));
        ASSERT_EQ(15, pushdown_popcnt(dec, col, sql::WHITE_OP_GE, {1}, str, ROW_CNT - 45, 30));
        ASSERT_EQ(ROW_CNT - 20, pushdown_popcnt(dec, col, sql::WHITE_OP_LE, {1}, str, 0, ROW_CNT));
        ASSERT_EQ(25, pushdown_popcnt(dec, col, sql::WHITE_OP_LE, {1}, str, ROW_CNT - 45, 30));
      }
      {  // filter_push_down_bt / in / nu / nn
        std::vector<uint8_t> blk = build_block(layout({{0, ROW_CNT - 40}, {1, 10}, {2, 10}, {3, 10}, {-1, 10}}), e[0], e[1]);
        ObGpuMicroBlockDecoder dec(rt);
        ObMicroBlockData data{(const char *)blk.data(), (int64_t)blk.size()};
        ASSERT_EQ(OB_SUCCESS, dec.init(data));
        ASSERT_EQ(ROW_CNT - 20, pushdown_popcnt(dec, col, sql::WHITE_OP_BT, {0, 2}, str, 0, ROW_CNT));
        ASSERT_EQ(0, pushdown_popcnt(dec, col, sql::WHITE_OP_BT, {2, 0}, str, 0, ROW_CNT));
        ASSERT_EQ(20, pushdown_popcnt(dec, col, sql::WHITE_OP_IN, {1, 2, 5}, str, 0, ROW_CNT));
        ASSERT_EQ(15, pushdown_popcnt(dec, col, sql::WHITE_OP_IN, {1, 2, 5}, str, ROW_CNT - 35, 30));
        ASSERT_EQ(0, pushdown_popcnt(dec, col, sql::WHITE_OP_IN, {5, 5, 5}, str, 0, ROW_CNT));
        ASSERT_EQ(10, pushdown_popcnt(dec, col, sql::WHITE_OP_NU, {}, str, 0, ROW_CNT));
        ASSERT_EQ(5, pushdown_popcnt(dec, col, sql::WHITE_OP_NU, {}, str, ROW_CNT - 35, 30));
        ASSERT_EQ(ROW_CNT - 10, pushdown_popcnt(dec, col, sql::WHITE_OP_NN, {}, str, 0, ROW_CNT));
        ASSERT_EQ(25, pushdown_popcnt(dec, col, sql::WHITE_OP_NN, {}, str, ROW_CNT - 35, 30));
      }
    }
  }
}

static void test_get_rows_vs_oracle(ObGpuScanRuntime &rt) {
  // batch decode into VEC_FIXED / VEC_DISCRETE == oracle's get_rows, incl. vec_offset and NULLs
  Rows r = layout({{3, 20}, {-1, 5}, {9, 30}, {-1, 2}, {1, 7}});
  for (size_t i = 0; i < r.ints.size(); ++i) { r.ints[i] += (int64_t)i * 13; r.strs[i] += std::to_string(i % 7); }
  std::vector<uint8_t> blk = build_block(r, OBGPU_ENC_RAW, OBGPU_ENC_DICT);
  ObGpuMicroBlockDecoder dec(rt);
  ObMicroBlockData data{(const char *)blk.data(), (int64_t)blk.size()};
  ASSERT_EQ(OB_SUCCESS, dec.init(data));
  ora_block ob;
  ASSERT_EQ(0, ora_block_init(&ob, blk.data(), (int64_t)blk.size()));
  std::vector<int32_t> row_ids;
  for (int32_t i = 1; i < (int32_t)r.ints.size(); i += 2) row_ids.push_back(i);
  const int64_t cap = (int64_t)row_ids.size(), voff = 3;
  ObFixedLengthVector fv;
  fv.len_ = 8;
  fv.reserve_rows(voff + cap);
  ASSERT_EQ(OB_SUCCESS, dec.get_rows(1, row_ids.data(), cap, voff, fv));
  std::vector<uint64_t> ev((size_t)(voff + cap), 0), en((size_t)(voff + cap + 63) / 64, 0);
  int32_t ehn = 0;
  ASSERT_EQ(0, ora_get_rows_fixed(&ob, 1, row_ids.data(), cap, voff, ev.data(), 8, en.data(), &ehn));
  ASSERT_EQ(ehn, (int)fv.has_null_);
  for (int64_t i = 0; i < voff + cap; ++i) {
    ASSERT_EQ((en[(size_t)i / 64] >> (i % 64)) & 1, (int)fv.is_null(i));
    if (!fv.is_null(i)) ASSERT_EQ((int64_t)ev[(size_t)i], fv.get_int(i));
  }
  ObDiscreteVector dv;
  dv.reserve_rows(voff + cap);
  ASSERT_EQ(OB_SUCCESS, dec.get_rows(2, row_ids.data(), cap, voff, dv));
  for (int64_t i = 0; i < cap; ++i) {
    const int32_t row = row_ids[(size_t)i];
    ASSERT_EQ((int)r.nulls[(size_t)row], (int)dv.is_null(voff + i));
    if (!r.nulls[(size_t)row]) {
      ASSERT_EQ((int64_t)r.strs[(size_t)row].size(), dv.lens_[(size_t)(voff + i)]);
      // zero-copy: the pointer lands inside the caller's block buffer
      const char *p = dv.ptrs_[(size_t)(voff + i)];
      ASSERT_EQ(1, p >= (const char *)blk.data() && p < (const char *)blk.data() + blk.size());
      ASSERT_EQ(0, memcmp(p, r.strs[(size_t)row].data(), r.strs[(size_t)row].size()));
    }
  }
  // datum format (ObMicroBlockDecoder::get_rows into ObDatum[], ob_micro_block_decoder.cpp:2100-2140): integers are
  // written through the datums' own pointers (the expression's reserved slots), strings point into the block
  const int64_t doff = 2;
  std::vector<int64_t> slots((size_t)(doff + cap), 0x5a5a5a5a5a5a5a5aLL);
  std::vector<common::ObDatum> datums((size_t)(doff + cap));
  for (size_t i = 0; i < datums.size(); ++i) datums[i].ptr_ = reinterpret_cast<const char *>(&slots[i]);
  ASSERT_EQ(OB_SUCCESS, dec.get_rows(1, row_ids.data(), cap, doff, datums.data()));
  for (int64_t i = 0; i < cap; ++i) {
    const common::ObDatum &d = datums[(size_t)(doff + i)];
    const bool is_null = (en[(size_t)(voff + i) / 64] >> ((voff + i) % 64)) & 1;
    ASSERT_EQ((int)is_null, (int)d.is_null());
    if (is_null) { ASSERT_EQ(0x5a5a5a5a5a5a5a5aLL, slots[(size_t)(doff + i)]); continue; }
    ASSERT_EQ(8, (int)d.len_);
    ASSERT_EQ(1, d.ptr_ == reinterpret_cast<const char *>(&slots[(size_t)(doff + i)]));
    ASSERT_EQ((int64_t)ev[(size_t)(voff + i)], d.get_int());
  }
  ASSERT_EQ(0, (int)datums[0].pack_);
  std::vector<common::ObDatum> sd((size_t)cap);
  ASSERT_EQ(OB_SUCCESS, dec.get_rows(2, row_ids.data(), cap, 0, sd.data()));
  for (int64_t i = 0; i < cap; ++i) {
    const int32_t row = row_ids[(size_t)i];
    ASSERT_EQ((int)r.nulls[(size_t)row], (int)sd[(size_t)i].is_null());
    if (!r.nulls[(size_t)row]) {
      ASSERT_EQ((int64_t)r.strs[(size_t)row].size(), (int64_t)sd[(size_t)i].len_);
      ASSERT_EQ(1, sd[(size_t)i].ptr_ >= (const char *)blk.data() && sd[(size_t)i].ptr_ < (const char *)blk.data() + blk.size());
      ASSERT_EQ(0, memcmp(sd[(size_t)i].ptr_, r.strs[(size_t)row].data(), r.strs[(size_t)row].size()));
    }
  }
}

static void test_filter_tree_and_batch_scanner(ObGpuScanRuntime &rt) {
  // a small SSTable: 20 000 rows in blocks of 777; AND(OR(a < 100, a >= 900), s = 'k3') ; project a, s
  const int64_t n = 20000, rpb = 777;
  std::vector<int64_t> a(n), b(n);
  std::string heap;
  std::vector<int64_t> off(n + 1, 0);
  std::vector<uint8_t> nulls(n, 0);
  uint64_t x = 88172645463325252ull;
  for (int64_t i = 0; i < n; ++i) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    a[i] = (int64_t)(x % 1000);
    b[i] = i;
    nulls[i] = (x >> 20) % 17 == 0;
    heap += "k" + std::to_string((x >> 32) % 8);
    off[i + 1] = (int64_t)heap.size();
  }
  heap.push_back('\0');
  obgpu_col_input cols[3];
  memset(cols, 0, sizeof(cols));
  cols[0].obj_type = OBGPU_OBJ_INT; cols[0].encoding = OBGPU_ENC_INTEGER_BASE_DIFF; cols[0].i64 = b.data();
  cols[1].obj_type = OBGPU_OBJ_INT; cols[1].encoding = OBGPU_ENC_RAW; cols[1].i64 = a.data(); cols[1].is_null = nulls.data();
  cols[2].obj_type = OBGPU_OBJ_VARCHAR; cols[2].encoding = OBGPU_ENC_DICT; cols[2].str_heap = heap.data(); cols[2].str_off = off.data();
  obgpu_table_image *img = nullptr;
  ASSERT_EQ(0, obgpu_writer_encode_table(cols, 3, 1, n, rpb, 128, 2, &img));
  int64_t image_size = 0;
  int32_t nb = 0;
  obgpu_table_image_info(img, &image_size, &nb);
  std::vector<uint8_t> image((size_t)image_size + 64, 0);
  std::vector<int64_t> offs((size_t)nb), sizes((size_t)nb);
  ASSERT_EQ(0, obgpu_table_image_export(img, image.data(), image_size, offs.data(), sizes.data(), nb));
  obgpu_table_image_free(img);

  sql::ObWhiteFilterExecutor lt(1, sql::WHITE_OP_LT), ge(1, sql::WHITE_OP_GE), eq(2, sql::WHITE_OP_EQ);
  ObStorageDatum d;
  d.set_int(100); lt.get_datums().push_back(d);
  d.set_int(900); ge.get_datums().push_back(d);
  ObStorageDatum ds; ds.set_string("k3", 2); eq.get_datums().push_back(ds);
  sql::ObOrFilterExecutor orf; orf.add_child(&lt); orf.add_child(&ge);
  sql::ObAndFilterExecutor andf; andf.add_child(&orf); andf.add_child(&eq);

  // (1) per-block tree execution through ObPushdownFilterExecutor::execute semantics
  int64_t expect_total = 0;
  for (int32_t blk = 0; blk < nb; blk += 7) {
    ObGpuMicroBlockDecoder dec(rt);
    ObMicroBlockData data{(const char *)image.data() + offs[(size_t)blk], sizes[(size_t)blk]};
    ASSERT_EQ(OB_SUCCESS, dec.init(data));
    int64_t rc = 0;
    dec.get_row_count(rc);
    sql::PushdownFilterInfo pd;
    pd.start_ = 0; pd.count_ = rc;
    ASSERT_EQ(OB_SUCCESS, execute_pushdown_filter(&andf, nullptr, pd, dec));
    int64_t exp = 0;
    for (int64_t i = 0; i < rc; ++i) {
      const int64_t g = (int64_t)blk * rpb + i;
      const bool m = !nulls[(size_t)g] && (a[(size_t)g] < 100 || a[(size_t)g] >= 900) &&
                     heap.compare((size_t)off[(size_t)g], (size_t)(off[(size_t)g + 1] - off[(size_t)g]), "k3") == 0;
      exp += m;
      ASSERT_EQ((int)m, (int)andf.get_result()->test(i));
    }
    ASSERT_EQ(exp, (int64_t)andf.get_result()->popcnt());
    expect_total += exp;
  }
  // (2) page-batch scanner: get_next_rows until OB_ITER_END == brute force over the generator
  ObGpuSSTableBatchScanner scanner(rt);
  ASSERT_EQ(OB_SUCCESS, scanner.init(image.data(), image_size, offs.data(), sizes.data(), nb, &andf, {1, 2, 0}, 256));
  ObGpuSSTableBatchScanner::Batch batch;
  int64_t seen = 0, g_expect = 0, last_block = -1, last_row = -1;
  int ret;
  while ((ret = scanner.get_next_rows(batch)) == OB_SUCCESS) {
    ASSERT_EQ(1, batch.count > 0 && batch.count <= 256);
    for (int64_t i = 0; i < batch.count; ++i) {
      const int64_t row = batch.row_ids[(size_t)i];
      // order: blocks ascending, rows ascending inside a block
      ASSERT_EQ(1, batch.block_idx > last_block || (batch.block_idx == last_block && row > last_row));
      last_block = batch.block_idx; last_row = row;
      const int64_t g = (int64_t)batch.block_idx * rpb + row;
      // advance the brute-force cursor to the next matching row: must be exactly g
      while (g_expect < n && !( !nulls[(size_t)g_expect] && (a[(size_t)g_expect] < 100 || a[(size_t)g_expect] >= 900) &&
             heap.compare((size_t)off[(size_t)g_expect], (size_t)(off[(size_t)g_expect + 1] - off[(size_t)g_expect]), "k3") == 0)) ++g_expect;
      ASSERT_EQ(g_expect, g);
      ++g_expect;
      ASSERT_EQ(0, (int)batch.is_null[0][(size_t)i]);
      ASSERT_EQ(a[(size_t)g], batch.ints[0][(size_t)i]);
      ASSERT_EQ(2, batch.str_lens[1][(size_t)i]);
      ASSERT_EQ(0, memcmp(batch.str_ptrs[1][(size_t)i], "k3", 2));
      ASSERT_EQ(b[(size_t)g], batch.ints[2][(size_t)i]);
    }
    seen += batch.count;
  }
  ASSERT_EQ(OB_ITER_END, ret);
  ASSERT_EQ(scanner.total_selected(), seen);
  ASSERT_EQ(OB_ITER_END, scanner.get_next_rows(batch));
  (void)expect_total;

  // (2b) the row-at-a-time contract (ObIStoreRowIterator::get_next_row until OB_ITER_END, reuse() rescans): same rows, same order
  {
    ObGpuStoreRowIterator iter(rt);
    ASSERT_EQ(OB_SUCCESS, iter.init(image.data(), image_size, offs.data(), sizes.data(), nb, &andf, {1, 2, 0}, 256));
    for (int pass = 0; pass < 2; ++pass) {
      const ObDatumRow *row = nullptr;
      int64_t cursor = 0, rows = 0;
      int r2;
      while ((r2 = iter.get_next_row(row)) == OB_SUCCESS) {
        while (cursor < n && !(!nulls[(size_t)cursor] && (a[(size_t)cursor] < 100 || a[(size_t)cursor] >= 900) &&
               heap.compare((size_t)off[(size_t)cursor], (size_t)(off[(size_t)cursor + 1] - off[(size_t)cursor]), "k3") == 0)) ++cursor;
        ASSERT_EQ(3, row->get_column_count());
        ASSERT_EQ(a[(size_t)cursor], row->storage_datums_[0].get_int());
        ASSERT_EQ(2, row->storage_datums_[1].len_);
        ASSERT_EQ(0, memcmp(row->storage_datums_[1].ptr_, "k3", 2));
        ASSERT_EQ(b[(size_t)cursor], row->storage_datums_[2].get_int());
        ++cursor;
        ++rows;
      }
      ASSERT_EQ(OB_ITER_END, r2);
      ASSERT_EQ(seen, rows);
      if (pass == 0) ASSERT_EQ(OB_SUCCESS, iter.reuse());
    }
    iter.reset();
    const ObDatumRow *row = nullptr;
    ASSERT_EQ(OB_NOT_INIT, iter.get_next_row(row));
  }

  // (3) skip index: b (column 0) is the row number, so BETWEEN on it is decided by min / max for all but the two
  // boundary blocks; same rows come back, the index infos carry the verdicts, pruned blocks are counted
  int32_t agg_cols[1] = {0};
  int64_t agg_size = 0;
  std::vector<int64_t> agg_off((size_t)nb + 1);
  ASSERT_EQ(0, obgpu_writer_table_agg_rows(cols, 3, agg_cols, 1, n, rpb, nullptr, 0, nullptr, &agg_size));
  std::vector<char> agg((size_t)agg_size);
  ASSERT_EQ(0, obgpu_writer_table_agg_rows(cols, 3, agg_cols, 1, n, rpb, agg.data(), agg_size, agg_off.data(), &agg_size));
  std::vector<ObMicroIndexInfo> infos((size_t)nb);
  for (int32_t i = 0; i < nb; ++i) {
    infos[(size_t)i].agg_row_buf_ = agg.data() + agg_off[(size_t)i];
    infos[(size_t)i].agg_buf_size_ = agg_off[(size_t)i + 1] - agg_off[(size_t)i];
  }
  const int64_t lo = 3 * rpb + 5, hi = 9 * rpb - 1;   // blocks 4..8 are inside, 3 partly, the rest outside
  sql::ObWhiteFilterExecutor bt(0, sql::WHITE_OP_BT);
  d.set_int(lo); bt.get_datums().push_back(d);
  d.set_int(hi); bt.get_datums().push_back(d);
  ObGpuSSTableBatchScanner pruned(rt);
  ASSERT_EQ(OB_SUCCESS, pruned.set_index_infos(infos.data(), nb));
  ASSERT_EQ(OB_SUCCESS, pruned.init(image.data(), image_size, offs.data(), sizes.data(), nb, &bt, {0, 1}, 256));
  ASSERT_EQ(hi - lo + 1, pruned.total_selected());
  ASSERT_EQ((int64_t)nb - 6, pruned.skipped_blocks());
  ASSERT_EQ(5, pruned.unfiltered_blocks());
  for (int32_t i = 0; i < nb; ++i) {
    const bool inside = i >= 4 && i <= 8;
    ASSERT_EQ(inside, infos[(size_t)i].is_filter_always_true());
    ASSERT_EQ(i != 3 && !inside, infos[(size_t)i].is_filter_always_false());
  }
  int64_t next = lo;
  while ((ret = pruned.get_next_rows(batch)) == OB_SUCCESS) {
    for (int64_t i = 0; i < batch.count; ++i) {
      ASSERT_EQ(next, batch.ints[0][(size_t)i]);
      ASSERT_EQ(next, (int64_t)batch.block_idx * rpb + batch.row_ids[(size_t)i]);
      if (!batch.is_null[1][(size_t)i]) ASSERT_EQ(a[(size_t)next], batch.ints[1][(size_t)i]);
      ASSERT_EQ((int)nulls[(size_t)next], (int)batch.is_null[1][(size_t)i]);
      ++next;
    }
  }
  ASSERT_EQ(OB_ITER_END, ret);
  ASSERT_EQ(hi + 1, next);
  ASSERT_EQ(OB_INIT_TWICE, pruned.set_index_infos(infos.data(), nb));

  // (4) reverse scan: same rows, blocks last to first, rows descending, batches of at most 100
  ObGpuSSTableBatchScanner rev(rt);
  rev.set_reverse_scan(true);
  ASSERT_EQ(OB_SUCCESS, rev.init(image.data(), image_size, offs.data(), sizes.data(), nb, &bt, {0, 2}, 100));
  next = hi;
  while ((ret = rev.get_next_rows(batch)) == OB_SUCCESS) {
    ASSERT_EQ(1, batch.count > 0 && batch.count <= 100);
    for (int64_t i = 0; i < batch.count; ++i) {
      ASSERT_EQ(next, batch.ints[0][(size_t)i]);
      ASSERT_EQ(next, (int64_t)batch.block_idx * rpb + batch.row_ids[(size_t)i]);
      const std::string want = heap.substr((size_t)off[(size_t)next], (size_t)(off[(size_t)next + 1] - off[(size_t)next]));
      ASSERT_EQ((int64_t)want.size(), (int64_t)batch.str_lens[1][(size_t)i]);
      ASSERT_EQ(0, memcmp(batch.str_ptrs[1][(size_t)i], want.data(), want.size()));
      --next;
    }
  }
  ASSERT_EQ(OB_ITER_END, ret);
  ASSERT_EQ(lo - 1, next);

  // (4b) pipelined open (obgpu_pipeline_scan underneath): same rows served from host memory, forward and reverse
  for (int rev_mode = 0; rev_mode < 2; ++rev_mode) {
    ObGpuSSTableBatchScanner pl(rt);
    pl.set_pipelined(3, 4);
    pl.set_reverse_scan(rev_mode == 1);
    ASSERT_EQ(OB_SUCCESS, pl.init(image.data(), image_size, offs.data(), sizes.data(), nb, &bt, {0, 1, 2}, 100));
    ASSERT_EQ(hi - lo + 1, pl.total_selected());
    next = rev_mode ? hi : lo;
    while ((ret = pl.get_next_rows(batch)) == OB_SUCCESS) {
      for (int64_t i = 0; i < batch.count; ++i) {
        ASSERT_EQ(next, batch.ints[0][(size_t)i]);
        ASSERT_EQ(next, (int64_t)batch.block_idx * rpb + batch.row_ids[(size_t)i]);
        ASSERT_EQ((int)nulls[(size_t)next], (int)batch.is_null[1][(size_t)i]);
        if (!batch.is_null[1][(size_t)i]) ASSERT_EQ(a[(size_t)next], batch.ints[1][(size_t)i]);
        const std::string want = heap.substr((size_t)off[(size_t)next], (size_t)(off[(size_t)next + 1] - off[(size_t)next]));
        ASSERT_EQ((int64_t)want.size(), (int64_t)batch.str_lens[2][(size_t)i]);
        ASSERT_EQ(0, memcmp(batch.str_ptrs[2][(size_t)i], want.data(), want.size()));
        next += rev_mode ? -1 : 1;
      }
    }
    ASSERT_EQ(OB_ITER_END, ret);
    ASSERT_EQ(rev_mode ? lo - 1 : hi + 1, next);
  }

  // (5) LIMIT / OFFSET (ObBlockBatchedRowStore::get_row_ids, ob_block_batched_row_store.cpp:163-186): the first `offset`
  // selected rows are dropped batch by batch, the scan ends with the batch that reaches `limit`
  const int64_t cases[][2] = {{0, 10}, {7, 300}, {255, 2}, {256, 256}, {1000, -1}, {hi - lo + 5, 10}, {3, 0}, {0, hi - lo + 100}};
  for (const auto &cs : cases) {
    const int64_t offset = cs[0], limit = cs[1];
    ObGpuSSTableBatchScanner lim(rt);
    lim.set_limit(offset, limit);
    ASSERT_EQ(OB_SUCCESS, lim.init(image.data(), image_size, offs.data(), sizes.data(), nb, &bt, {0}, 256));
    const int64_t total = hi - lo + 1;
    const int64_t first = std::min(offset, total), want_n = limit < 0 ? total - first : std::min(limit, total - first);
    int64_t got = 0;
    next = lo + first;
    while ((ret = lim.get_next_rows(batch)) == OB_SUCCESS) {
      ASSERT_EQ(1, batch.count > 0);
      for (int64_t i = 0; i < batch.count; ++i) ASSERT_EQ(next++, batch.ints[0][(size_t)i]);
      got += batch.count;
    }
    ASSERT_EQ(OB_ITER_END, ret);
    ASSERT_EQ(want_n, got);
  }
}

// HEX_PACKING / STRING_PREFIX columns through the adapter: the white filter popcounts of the reference layout and VEC_DISCRETE
// get_rows whose pointers land in the decoder's own arena (the values do not exist in the block).
static void test_rebuilt_string_codecs(ObGpuScanRuntime &rt) {
  const int encs[] = {OBGPU_ENC_HEX_PACKING, OBGPU_ENC_STRING_PREFIX};
  for (int enc : encs) {
    Rows r = layout({{0, ROW_CNT - 40}, {1, 10}, {2, 10}, {3, 10}, {-1, 10}});
    std::vector<uint8_t> blk = build_block(r, OBGPU_ENC_DICT, enc);
    ObGpuMicroBlockDecoder dec(rt);
    ObMicroBlockData data{(const char *)blk.data(), (int64_t)blk.size()};
    ASSERT_EQ(OB_SUCCESS, dec.init(data));
    ASSERT_EQ(ROW_CNT - 40, pushdown_popcnt(dec, 2, sql::WHITE_OP_EQ, {0}, 1, 0, ROW_CNT));
    ASSERT_EQ(20, pushdown_popcnt(dec, 2, sql::WHITE_OP_IN, {1, 2, 5}, 1, 0, ROW_CNT));
    ASSERT_EQ(10, pushdown_popcnt(dec, 2, sql::WHITE_OP_GT, {2}, 1, 0, ROW_CNT));
    ASSERT_EQ(10, pushdown_popcnt(dec, 2, sql::WHITE_OP_NU, {}, 1, 0, ROW_CNT));
    ASSERT_EQ(5, pushdown_popcnt(dec, 2, sql::WHITE_OP_NU, {}, 1, ROW_CNT - 35, 30));
    std::vector<int32_t> rid;
    for (int32_t i = 1; i < ROW_CNT; i += 3) rid.push_back(i);
    ObDiscreteVector vec;
    vec.reserve_rows((int64_t)rid.size() + 2);
    ASSERT_EQ(OB_SUCCESS, dec.get_rows(2, rid.data(), (int64_t)rid.size(), 2, vec));
    for (size_t i = 0; i < rid.size(); ++i) {
      const int32_t row = rid[i];
      ASSERT_EQ(r.nulls[(size_t)row] != 0, vec.is_null((int64_t)i + 2));
      if (!r.nulls[(size_t)row]) {
        ASSERT_EQ((long long)r.strs[(size_t)row].size(), vec.lens_[i + 2]);
        ASSERT_EQ(0, memcmp(vec.ptrs_[i + 2], r.strs[(size_t)row].data(), r.strs[(size_t)row].size()));
        const bool inside = vec.ptrs_[i + 2] >= (const char *)blk.data() && vec.ptrs_[i + 2] < (const char *)blk.data() + blk.size();
        ASSERT_EQ(0, inside);   // not a pointer into the block: there is nothing to point at
      }
    }
  }
}

// Black filter on one dictionary column + the group-by surface (test_dict_decoder.cpp's batch black-filter / group-by
// cases in shape): the expression is only known to the caller, here "value % 2000 == 7 or NULL" / "string ends in an odd digit".
struct OddSeedFilter : public sql::ObBlackFilterExecutor {
  bool str_;
  int calls_ = 0;
  OddSeedFilter(int32_t col, bool str) : sql::ObBlackFilterExecutor({col}), str_(str) {}
  int filter(const common::ObDatum &d, bool &filtered) override {
    ++calls_;
    if (d.is_null()) { filtered = false; return OB_SUCCESS; }          // NULL rows pass (IS NULL OR ...)
    if (str_) filtered = ((d.ptr_[d.len_ - 1] - '0') & 1) == 0;
    else filtered = ((d.get_int() / 1000) & 1) == 0;
    return OB_SUCCESS;
  }
};

static void test_black_filter_and_group_by_surface(ObGpuScanRuntime &rt) {
  const int encs[] = {OBGPU_ENC_DICT, OBGPU_ENC_RLE};
  for (int enc : encs) {
    // [seed0 x 24 | seed1 x 10 | seed2 x 10 | seed3 x 10 | NULL x 10]
    std::vector<uint8_t> blk = build_block(layout({{0, ROW_CNT - 40}, {1, 10}, {2, 10}, {3, 10}, {-1, 10}}), enc, enc);
    ObGpuMicroBlockDecoder dec(rt);
    ObMicroBlockData data{(const char *)blk.data(), (int64_t)blk.size()};
    ASSERT_EQ(OB_SUCCESS, dec.init(data));
    for (int str = 0; str < 2; ++str) {
      const int col = str ? 2 : 1;
      int64_t cnt = 0;
      ASSERT_EQ(OB_SUCCESS, dec.get_distinct_count(col, cnt));
      ASSERT_EQ(4, cnt);
      std::vector<uint64_t> slots(4);
      std::vector<ObDatum> dist(4);
      for (int i = 0; i < 4; ++i) dist[i].ptr_ = (const char *)&slots[i];
      ASSERT_EQ(OB_SUCCESS, dec.read_distinct(col, dist.data(), 4, cnt));
      for (int i = 0; i < 4; ++i) {   // the writer's dictionaries are sorted: entry i is seed i
        if (str) ASSERT_EQ(0, memcmp(dist[i].ptr_, seed_str(i).data(), dist[i].len_) + (int)(dist[i].len_ != seed_str(i).size()));
        else ASSERT_EQ(seed_int(i), dist[i].get_int());
      }
      ASSERT_EQ(OB_BUF_NOT_ENOUGH, dec.read_distinct(col, dist.data(), 3, cnt));
      std::vector<int32_t> rid = {0, 23, 24, 33, 34, 44, 53, 54, 63};
      std::vector<uint32_t> refs(rid.size());
      ASSERT_EQ(OB_SUCCESS, dec.read_reference(col, rid.data(), (int64_t)rid.size(), refs.data()));
      const uint32_t want_refs[] = {0, 0, 1, 1, 2, 3, 3, 4, 4};
      for (size_t i = 0; i < rid.size(); ++i) ASSERT_EQ(want_refs[i], refs[i]);
      // black filter: seeds 1 and 3 pass (10 + 10 rows) + the 10 NULL rows
      OddSeedFilter f(col, str != 0);
      sql::PushdownFilterInfo pd;
      pd.start_ = 0; pd.count_ = ROW_CNT;
      ObBitmap bm;
      bm.init(ROW_CNT);
      bool applied = false;
      ASSERT_EQ(OB_SUCCESS, dec.filter_black_filter_batch(nullptr, f, pd, bm, applied));
      ASSERT_EQ(1, applied);
      ASSERT_EQ(30, bm.popcnt());
      ASSERT_EQ(5, f.calls_);   // once per distinct value + once for NULL, not once per row
      for (int64_t r = 0; r < ROW_CNT; ++r) ASSERT_EQ((r >= 24 && r < 34) || r >= 44, bm.test(r));
      pd.start_ = ROW_CNT - 35; pd.count_ = 30;   // rows 29..58
      ObBitmap win;
      win.init(30);
      ASSERT_EQ(OB_SUCCESS, dec.filter_black_filter_batch(nullptr, f, pd, win, applied));
      ASSERT_EQ(5 + 10 + 5, win.popcnt());
    }
    // column 0 is RAW: not applied, bitmap untouched, no error (the caller keeps its row-wise path)
    OddSeedFilter f0(0, false);
    sql::PushdownFilterInfo pd;
    pd.start_ = 0; pd.count_ = ROW_CNT;
    ObBitmap bm;
    bm.init(ROW_CNT);
    bool applied = true;
    ASSERT_EQ(OB_SUCCESS, dec.filter_black_filter_batch(nullptr, f0, pd, bm, applied));
    ASSERT_EQ(0, applied);
    ASSERT_EQ(0, bm.popcnt());
    int64_t cnt = 0;
    ASSERT_EQ(OB_NOT_SUPPORTED, dec.get_distinct_count(0, cnt));
  }
}

int main() {
  ObGpuScanRuntime rt(0);
  if (!rt.is_valid()) {
    printf("no CUDA device: the adapter has no CPU fallback (expected on a CPU-only box)\n");
    return 77;
  }
  test_filter_pushdown(rt);
  test_get_rows_vs_oracle(rt);
  test_filter_tree_and_batch_scanner(rt);
  test_black_filter_and_group_by_surface(rt);
  test_rebuilt_string_codecs(rt);
  if (g_fail) { printf("%d assertion(s) failed\n", g_fail); return 1; }
  printf("host adapter tests passed\n");
  return 0;
}
