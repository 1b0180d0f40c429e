// C++ parity test of the compaction adapter (oceanbase_b200/host/ob_gpu_partition_merger.h): K encoded runs
// with overlapping rowkey ranges, NOP cells in incremental rows and delete rows are merged on the device
// through ObGpuPartitionMajorMerger and compared row by row with the oracle's restatement of
// ObPartitionMajorMerger::merge_partition (oracle/ob_oracle.c: ora_major_merge). Also replays the reference's
// row-fuse expectation test_fuse_nomal (unittest/storage/test_row_fuse.cpp:118-131) as five single-row tables.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../oceanbase_b200/host/ob_gpu_partition_merger.h"
extern "C" {
#include "../../include/obgpu_writer.h"
#include "../../oracle/ob_oracle.h"
}

using namespace oceanbase;
using namespace oceanbase::common;
using namespace oceanbase::compaction;

static int g_fail = 0;
#define ASSERT_EQ(a, b)                                                                           \
  do {                                                                                            \
    const long long va__ = (long long)(a), vb__ = (long long)(b);                                 \
    if (va__ != vb__) {                                                                           \
      printf("FAIL %s:%d  %s = %lld, expected %lld\n", __FILE__, __LINE__, #a, va__, vb__);       \
      ++g_fail;                                                                                   \
    }                                                                                             \
  } while (0)

struct Run {
  std::vector<int64_t> key, flag;
  std::vector<std::vector<int64_t>> vals;   // [3]
  std::vector<std::vector<uint8_t>> ext;    // [3]: 0 value, 1 NULL, 2 NOP
  std::vector<uint8_t> image;
  std::vector<int64_t> offsets, sizes;
};

static uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

static void encode(Run &r, int64_t rows_per_block) {
  const int64_t n = (int64_t)r.key.size();
  obgpu_col_input cols[5];
  memset(cols, 0, sizeof(cols));
  cols[0].obj_type = OBGPU_OBJ_INT; cols[0].encoding = OBGPU_ENC_INTEGER_BASE_DIFF; cols[0].i64 = r.key.data();
  cols[1].obj_type = OBGPU_OBJ_TINYINT; cols[1].encoding = OBGPU_ENC_RAW; cols[1].i64 = r.flag.data();
  for (int c = 0; c < 3; ++c) {
    cols[2 + c].obj_type = OBGPU_OBJ_INT; cols[2 + c].encoding = OBGPU_ENC_RAW;
    cols[2 + c].i64 = r.vals[c].data(); cols[2 + c].is_null = r.ext[c].data();
  }
  obgpu_table_image *img = nullptr;
  if (obgpu_writer_encode_table(cols, 5, 1, n, rows_per_block, 128, 2, &img) != 0) { printf("encode failed\n"); exit(2); }
  int64_t size = 0; int32_t nb = 0;
  obgpu_table_image_info(img, &size, &nb);
  r.image.assign((size_t)size + 64, 0);
  r.offsets.resize((size_t)nb); r.sizes.resize((size_t)nb);
  obgpu_table_image_export(img, r.image.data(), size, r.offsets.data(), r.sizes.data(), nb);
  obgpu_table_image_free(img);
}

static void check_against_oracle(std::vector<Run> &runs, const std::vector<int64_t> &dv, const std::vector<uint8_t> &dn) {
  std::vector<ObGpuMergeTable> tables;
  for (Run &r : runs) {
    ObGpuMergeTable t;
    t.image_ = r.image.data(); t.image_size_ = (int64_t)r.image.size() - 64;
    t.offsets_ = r.offsets.data(); t.sizes_ = r.sizes.data(); t.block_count_ = (int32_t)r.offsets.size();
    tables.push_back(t);
  }
  ObGpuMergeSchema schema;
  schema.rowkey_col_ = 0; schema.flag_col_ = 1; schema.payload_cols_ = {2, 3, 4};
  schema.default_vals_ = dv; schema.default_null_ = dn;
  ObGpuPartitionMajorMerger merger;
  ASSERT_EQ(merger.init(0, tables, schema), OB_SUCCESS);
  ASSERT_EQ(merger.merge_partition(), OB_SUCCESS);
  // oracle
  std::vector<ora_merge_run> oruns(runs.size());
  std::vector<std::vector<uint8_t>> flag8(runs.size());
  std::vector<std::vector<const int64_t *>> vp(runs.size());
  std::vector<std::vector<const uint8_t *>> ep(runs.size());
  int64_t total = 0;
  for (size_t i = 0; i < runs.size(); ++i) {
    flag8[i].assign(runs[i].flag.begin(), runs[i].flag.end());
    for (int c = 0; c < 3; ++c) { vp[i].push_back(runs[i].vals[c].data()); ep[i].push_back(runs[i].ext[c].data()); }
    oruns[i].n = (int64_t)runs[i].key.size(); oruns[i].key = runs[i].key.data(); oruns[i].flag = flag8[i].data();
    oruns[i].vals = vp[i].data(); oruns[i].ext = ep[i].data();
    total += oruns[i].n;
  }
  std::vector<int64_t> okey((size_t)total + 1);
  std::vector<std::vector<int64_t>> ov(3, std::vector<int64_t>((size_t)total + 1));
  std::vector<std::vector<uint8_t>> on(3, std::vector<uint8_t>((size_t)total + 1));
  int64_t *ovp[3] = {ov[0].data(), ov[1].data(), ov[2].data()};
  uint8_t *onp[3] = {on[0].data(), on[1].data(), on[2].data()};
  int64_t orows = 0, stats[2] = {0, 0};
  ASSERT_EQ(ora_major_merge(oruns.data(), (int32_t)runs.size(), 3, dv.empty() ? nullptr : dv.data(), dn.empty() ? nullptr : dn.data(),
                            total, okey.data(), ovp, onp, &orows, stats), 0);
  ASSERT_EQ(merger.get_output_row_count(), orows);
  ASSERT_EQ(merger.get_dropped_delete_count(), stats[0]);
  ASSERT_EQ(merger.get_fused_row_count(), stats[1]);
  int64_t at = 0;
  ObGpuMergedRows rows;
  int ret;
  while (OB_SUCCESS == (ret = merger.get_next_rows(1000, rows))) {
    for (int64_t i = 0; i < rows.row_count_ && g_fail < 10; ++i, ++at) {
      ASSERT_EQ(rows.rowkeys_[(size_t)i], okey[(size_t)at]);
      for (int c = 0; c < 3; ++c) {
        ASSERT_EQ(rows.nulls_[c][(size_t)i], on[c][(size_t)at]);
        ASSERT_EQ(rows.values_[c][(size_t)i], ov[c][(size_t)at]);
      }
    }
  }
  ASSERT_EQ(ret, OB_ITER_END);
  ASSERT_EQ(at, orows);
  // ---- column-oriented merge, writer side: the same merged stream replayed into three column groups on the device
  // (ObCOMergeLogReplayer / ObWriteHelper::project): every group's blocks must be the host writer's blocks over the ORACLE's
  // merged rows, pass the oracle's checksum checks, and carry the oracle's column checksums
  {
    std::vector<ObGpuColumnGroup> groups(3);
    groups[0].cols_ = {-1, 0, 1, 2}; groups[0].obj_types_.assign(4, OBGPU_OBJ_INT); groups[0].rowkey_col_cnt_ = 1;   // all-column group
    groups[1].cols_ = {1}; groups[1].obj_types_ = {OBGPU_OBJ_INT};
    groups[2].cols_ = {2, 0}; groups[2].obj_types_.assign(2, OBGPU_OBJ_INT);
    std::vector<ObGpuEncodedColumnGroup> enc;
    const int64_t rpb = 700;
    ASSERT_EQ(merger.write_column_groups(groups, rpb, 128, enc), OB_SUCCESS);
    ASSERT_EQ((long long)enc.size(), 3);
    for (size_t g = 0; g < enc.size() && g_fail < 10; ++g) {
      const ObGpuColumnGroup &cg = groups[g];
      std::vector<obgpu_col_input> in(cg.cols_.size());
      for (size_t c = 0; c < cg.cols_.size(); ++c) {
        memset(&in[c], 0, sizeof(in[c]));
        in[c].obj_type = OBGPU_OBJ_INT; in[c].encoding = OBGPU_ENC_RAW;
        in[c].i64 = cg.cols_[c] < 0 ? okey.data() : ov[(size_t)cg.cols_[c]].data();
        in[c].is_null = cg.cols_[c] < 0 ? nullptr : on[(size_t)cg.cols_[c]].data();
        const int64_t want = ora_column_checksum(in[c].i64, in[c].is_null, orows, 8);
        ASSERT_EQ(enc[g].column_checksums_[c], want);
      }
      obgpu_table_image *img = nullptr;
      ASSERT_EQ(obgpu_writer_encode_table(in.data(), (int32_t)in.size(), cg.rowkey_col_cnt_, orows, rpb, 128, 2, &img), 0);
      int64_t size = 0; int32_t nb = 0;
      obgpu_table_image_info(img, &size, &nb);
      std::vector<uint8_t> want((size_t)size);
      std::vector<int64_t> woff((size_t)nb), wsz((size_t)nb);
      obgpu_table_image_export(img, want.data(), size, woff.data(), wsz.data(), nb);
      obgpu_table_image_free(img);
      ASSERT_EQ(enc[g].row_count_, orows);
      ASSERT_EQ((long long)enc[g].offsets_.size(), nb);
      ASSERT_EQ((long long)enc[g].image_.size(), size);
      if (enc[g].image_.size() == want.size()) ASSERT_EQ(memcmp(enc[g].image_.data(), want.data(), want.size()), 0);
      for (int32_t b = 0; b < nb && g_fail < 10; ++b) {
        ASSERT_EQ(enc[g].offsets_[(size_t)b], woff[(size_t)b]);
        ASSERT_EQ(enc[g].sizes_[(size_t)b], wsz[(size_t)b]);
        ora_block blk;
        ASSERT_EQ(ora_block_init(&blk, enc[g].image_.data() + enc[g].offsets_[(size_t)b], enc[g].sizes_[(size_t)b]), 0);
        ASSERT_EQ(ora_block_verify_checksums(&blk), 0);
      }
    }
  }
}

int main() {
  {  // device available?
    obgpu_ctx *probe = nullptr;
    if (obgpu_ctx_create(0, &probe) != 0) { printf("no CUDA device: the adapter refuses (no CPU fallback)\n"); return 77; }
    obgpu_ctx_destroy(probe);
  }
  // ---- K = 5 overlapping runs, duplicates with NOP cells, deletes ------------------------------------
  {
    const int K = 5;
    const int64_t window = 30000;
    std::vector<Run> runs(K);
    for (int r = 0; r < K; ++r) {
      Run &run = runs[r];
      run.vals.assign(3, {}); run.ext.assign(3, {});
      for (int64_t i = r * window / 2; i < r * window / 2 + window; ++i) {
        const uint64_t h = mix((uint64_t)i * 7919u);
        const bool dup = (h >> 20) % 100 < 12;
        const int64_t first_cov = std::max<int64_t>((i - window) / (window / 2) + 1, 0);
        const int64_t last_cov = std::min<int64_t>(i / (window / 2), K - 1);
        const int64_t home = first_cov + (int64_t)(h % (uint64_t)(last_cov - first_cov + 1));
        if (home != r && !dup) continue;
        const bool newer = dup && first_cov < r;
        const uint64_t hr = mix(h ^ (uint64_t)(r + 1) * 0x1234567ull);
        int64_t flag = newer ? OBGPU_DF_UPDATE : OBGPU_DF_INSERT;
        if (r >= 1 && (hr >> 8) % 100 < 3) flag = OBGPU_DF_DELETE;
        run.key.push_back(1000003 + i * 5 + (int64_t)(h % 5));
        run.flag.push_back(flag);
        for (int c = 0; c < 3; ++c) {
          const uint64_t hv = mix(hr + (uint64_t)c * 977u);
          uint8_t e = hv % 100 < 6 ? 1 : 0;
          if (newer && (hv >> 7) % 100 < 55) e = 2;
          if (flag == OBGPU_DF_DELETE) e = 2;
          run.vals[c].push_back(e ? 0 : (int64_t)(hv >> 22));
          run.ext[c].push_back(e);
        }
      }
      encode(run, 1200);
    }
    check_against_oracle(runs, {}, {});
    check_against_oracle(runs, {11, 22, 33}, {0, 1, 0});
  }
  // ---- two runs whose second payload column is NULL in 60 % of the rows with 60-bit values: ObRawEncoder stores such a
  // column as var-length cells, the device leaves those blocks to the host writer (write_column_groups splices them in) ----
  {
    std::vector<Run> runs(2);
    for (int r = 0; r < 2; ++r) {
      Run &run = runs[r];
      run.vals.assign(3, {}); run.ext.assign(3, {});
      for (int64_t i = 0; i < 4000; ++i) {
        const uint64_t h = mix((uint64_t)i * 131u + (uint64_t)r);
        if (h % 3 == 0) continue;
        run.key.push_back(500 + i * 2 + r);   // disjoint rowkeys: nothing fuses
        run.flag.push_back(OBGPU_DF_INSERT);
        for (int c = 0; c < 3; ++c) {
          const uint64_t hv = mix(h + (uint64_t)c * 977u);
          const uint8_t e = (c == 1 && hv % 100 < 60) ? 1 : 0;
          run.vals[c].push_back(e ? 0 : (int64_t)(hv >> 4));
          run.ext[c].push_back(e);
        }
      }
      encode(run, 900);
    }
    check_against_oracle(runs, {}, {});
  }
  // ---- the reference's test_fuse_nomal as five single-row tables (rows listed newest first there) ----------
  {
    const int64_t NOPV = INT64_MIN, NULLV = INT64_MIN + 1, MAXV = (int64_t)1 << 62;
    const int64_t rows_newest_first[5][3] = {{NOPV, NOPV, NOPV}, {22, NOPV, NULLV}, {33, 333, 5555}, {44, NOPV, NOPV}, {99, 999, NULLV}};
    (void)MAXV;
    std::vector<Run> runs(5);
    for (int r = 0; r < 5; ++r) {
      Run &run = runs[r];
      const int64_t *src = rows_newest_first[4 - r];  // run 0 = oldest = last listed
      run.key = {7}; run.flag = {OBGPU_DF_INSERT};
      run.vals.assign(3, {}); run.ext.assign(3, {});
      for (int c = 0; c < 3; ++c) {
        const uint8_t e = src[c] == NOPV ? 2 : (src[c] == NULLV ? 1 : 0);
        run.vals[c].push_back(e ? 0 : src[c]);
        run.ext[c].push_back(e);
      }
      encode(run, 10);
    }
    check_against_oracle(runs, {}, {});
    // expected by the reference: var2 (22), 3.33 (333), NULL
    std::vector<ObGpuMergeTable> tables;
    for (Run &r : runs) {
      ObGpuMergeTable t;
      t.image_ = r.image.data(); t.image_size_ = (int64_t)r.image.size() - 64;
      t.offsets_ = r.offsets.data(); t.sizes_ = r.sizes.data(); t.block_count_ = 1;
      tables.push_back(t);
    }
    ObGpuMergeSchema schema;
    schema.flag_col_ = 1; schema.payload_cols_ = {2, 3, 4};
    ObGpuPartitionMajorMerger merger;
    ASSERT_EQ(merger.init(0, tables, schema), OB_SUCCESS);
    ASSERT_EQ(merger.merge_partition(), OB_SUCCESS);
    ObGpuMergedRows rows;
    ASSERT_EQ(merger.get_next_rows(16, rows), OB_SUCCESS);
    ASSERT_EQ(rows.row_count_, 1);
    ASSERT_EQ(rows.values_[0][0], 22);
    ASSERT_EQ(rows.values_[1][0], 333);
    ASSERT_EQ(rows.nulls_[2][0], 1);
    ASSERT_EQ(merger.get_next_rows(16, rows), OB_ITER_END);
  }
  // ---- composite rowkey (tenant, order) + a VARCHAR payload column: three tables, every (tenant, order) in one or two of
  // them; the newest table's string wins ---------------------------------------------------------------------------------
  {
    const int K = 3;
    const int64_t tenants = 40, orders = 60;
    struct SRun { std::vector<int64_t> k0, k1, flag; std::string heap; std::vector<int64_t> off; std::vector<uint8_t> image;
                  std::vector<int64_t> offsets, sizes; };
    std::vector<SRun> runs(K);
    std::vector<std::string> expect((size_t)(tenants * orders));     // by (tenant, order): winning string ("" = absent)
    for (int r = 0; r < K; ++r) {
      SRun &run = runs[r];
      run.off.push_back(0);
      for (int64_t t = 0; t < tenants; ++t)
        for (int64_t o = 0; o < orders; ++o) {
          if (mix((uint64_t)(t * 1000 + o) * 31u + (uint64_t)r) % 3 == 0) continue;   // about two thirds of the keys per table
          run.k0.push_back(t * 7 - 100);
          run.k1.push_back(o - 30);
          run.flag.push_back(r == 0 ? OBGPU_DF_INSERT : OBGPU_DF_UPDATE);
          const std::string v = "t" + std::to_string(t) + "/o" + std::to_string(o) + "@" + std::to_string(r);
          run.heap += v;
          run.off.push_back((int64_t)run.heap.size());
          expect[(size_t)(t * orders + o)] = v;                                       // later (newer) tables overwrite
        }
      run.heap.push_back('\0');
      obgpu_col_input cols[4];
      memset(cols, 0, sizeof(cols));
      cols[0].obj_type = OBGPU_OBJ_INT; cols[0].encoding = OBGPU_ENC_RLE; cols[0].i64 = run.k0.data();
      cols[1].obj_type = OBGPU_OBJ_INT; cols[1].encoding = OBGPU_ENC_RAW; cols[1].i64 = run.k1.data();
      cols[2].obj_type = OBGPU_OBJ_TINYINT; cols[2].encoding = OBGPU_ENC_RAW; cols[2].i64 = run.flag.data();
      cols[3].obj_type = OBGPU_OBJ_VARCHAR; cols[3].encoding = OBGPU_ENC_RAW; cols[3].str_heap = run.heap.data(); cols[3].str_off = run.off.data();
      obgpu_table_image *img = nullptr;
      ASSERT_EQ(obgpu_writer_encode_table(cols, 4, 2, (int64_t)run.k0.size(), 500, 128, 2, &img), 0);
      int64_t size = 0; int32_t nb = 0;
      obgpu_table_image_info(img, &size, &nb);
      run.image.assign((size_t)size + 64, 0);
      run.offsets.resize((size_t)nb); run.sizes.resize((size_t)nb);
      obgpu_table_image_export(img, run.image.data(), size, run.offsets.data(), run.sizes.data(), nb);
      obgpu_table_image_free(img);
    }
    std::vector<ObGpuMergeTable> tables;
    for (SRun &r : runs) {
      ObGpuMergeTable t;
      t.image_ = r.image.data(); t.image_size_ = (int64_t)r.image.size() - 64;
      t.offsets_ = r.offsets.data(); t.sizes_ = r.sizes.data(); t.block_count_ = (int32_t)r.offsets.size();
      tables.push_back(t);
    }
    ObGpuMergeSchema schema;
    schema.rowkey_col_ = 0; schema.more_rowkey_cols_ = {1}; schema.flag_col_ = 2;
    schema.payload_cols_ = {3}; schema.payload_is_string_ = {1};
    ObGpuPartitionMajorMerger merger;
    ASSERT_EQ(merger.init(0, tables, schema), OB_SUCCESS);
    ASSERT_EQ(merger.merge_partition(), OB_SUCCESS);
    int64_t present = 0;
    for (const std::string &e : expect) present += !e.empty();
    ASSERT_EQ(merger.get_output_row_count(), present);
    ObGpuMergedRows rows;
    int64_t t = 0, o = -1, seen = 0;
    int ret;
    while ((ret = merger.get_next_rows(700, rows)) == OB_SUCCESS) {
      for (int64_t i = 0; i < rows.row_count_; ++i) {
        do { if (++o == orders) { o = 0; ++t; } } while (t < tenants && expect[(size_t)(t * orders + o)].empty());   // next present key
        ASSERT_EQ(rows.rowkeys_[(size_t)i], t * 7 - 100);
        ASSERT_EQ(rows.more_rowkeys_[0][(size_t)i], o - 30);
        const std::string &want = expect[(size_t)(t * orders + o)];
        const int64_t a = rows.offsets_[0][(size_t)i], b = rows.offsets_[0][(size_t)i + 1];
        ASSERT_EQ(b - a, (int64_t)want.size());
        ASSERT_EQ(memcmp(rows.heap_[0].data() + a, want.data(), want.size()), 0);
        ASSERT_EQ((int)rows.nulls_[0][(size_t)i], 0);
        ++seen;
      }
    }
    ASSERT_EQ(ret, OB_ITER_END);
    ASSERT_EQ(seen, present);
  }
  if (g_fail) { printf("%d failures\n", g_fail); return 1; }
  printf("partition merger tests passed\n");
  return 0;
}
