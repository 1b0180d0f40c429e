// skip_index.cuh -- min / max / null-count block pruning on the device (included by obgpu_scan.cu).
//
// One thread per micro-block parses the block's serialized aggregate row (ObAggRowReader,
// index_block/ob_agg_row_struct.cpp:303-482), evaluates every white leaf of the pushed-down filter against
// (min, max, null count) the way ObSkipIndexFilterExecutor::filter_on_min_max does
// (index_block/ob_skip_index_filter_executor.cpp:250-396 with the operators at :530-822) and folds the leaf
// verdicts through the tree with ObBoolMask's & / | (sql/engine/basic/ob_pushdown_filter.h:133-158,
// execute_skipping_filter ob_pushdown_filter.cpp:1707-1740). The count kernel then skips always-false and
// always-true blocks and, inside an uncertain block, the leaves that are constant on it.
#pragma once

namespace skipidx {

using namespace obdev;

enum : uint8_t { MASK_UNCERTAIN = 0, MASK_TRUE = 1, MASK_FALSE = 2 };

struct AggDatum {
  const uint8_t *p;   // nullptr: aggregate not stored (NULL datum)
  uint32_t len;
  bool prefix;
};

__device__ __forceinline__ uint64_t agg_le(const uint8_t *p, uint32_t bytes) {
  uint64_t v = 0;
  for (uint32_t i = 0; i < bytes; ++i) v |= (uint64_t)p[i] << (8u * i);
  return v;
}

struct AggRow {
  const uint8_t *buf;
  uint32_t size;
  uint32_t cnt, idx_size, idx_off_size, cell_off_size, bitmaps;
  bool ok;
};

__device__ __forceinline__ void agg_row_init(const uint8_t *buf, uint32_t size, AggRow &r) {
  r.buf = buf;
  r.size = size;
  r.ok = false;
  if (size < 8u) return;
  const int16_t version = (int16_t)agg_le(buf, 2), cnt = (int16_t)agg_le(buf + 4, 2);
  const uint32_t pack = (uint32_t)agg_le(buf + 6, 2);
  r.idx_size = pack & 0x3fu;
  r.idx_off_size = (pack >> 6) & 7u;
  r.cell_off_size = (pack >> 9) & 7u;
  if (version < 1 || version > 3 || cnt <= 0 || ((pack >> 12) & 0xfu) != 1u) return;
  if (r.idx_size == 0 || r.idx_size > 4u || (r.idx_off_size != 1u && r.idx_off_size != 2u) ||
      (r.cell_off_size != 1u && r.cell_off_size != 2u))
    return;
  r.cnt = (uint32_t)cnt;
  r.bitmaps = version >= 2 ? 2u : 1u;
  if (8u + r.cnt * (r.idx_size + r.idx_off_size) > size) return;
  r.ok = true;
}

// ObAggRowReader::read: binary search of the column, then the cell's type bitmap / offsets
__device__ __forceinline__ bool agg_row_read(const AggRow &r, uint32_t col_idx, uint32_t type, AggDatum &d) {
  d.p = nullptr;
  d.len = 0;
  d.prefix = false;
  const uint8_t *idx_arr = r.buf + 8, *off_arr = idx_arr + r.cnt * r.idx_size;
  uint32_t lo = 0, hi = r.cnt;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if ((uint32_t)agg_le(idx_arr + mid * r.idx_size, r.idx_size) < col_idx) lo = mid + 1; else hi = mid;
  }
  if (lo >= r.cnt || (uint32_t)agg_le(idx_arr + lo * r.idx_size, r.idx_size) != col_idx) return true;
  const uint32_t pos = (uint32_t)agg_le(off_arr + lo * r.idx_off_size, r.idx_off_size);
  if (pos == 0) return true;
  if (pos + r.bitmaps > r.size) return false;
  const uint8_t *cell = r.buf + pos;
  const uint32_t types = cell[0], mask = 1u << type;
  if (!(types & mask)) return true;
  const uint32_t pre = (uint32_t)__popc(types & (mask - 1u));
  if (pos + r.bitmaps + (pre + 2u) * r.cell_off_size > r.size) return false;
  const uint8_t *offs = cell + r.bitmaps;
  const uint32_t a = (uint32_t)agg_le(offs + pre * r.cell_off_size, r.cell_off_size);
  const uint32_t b = (uint32_t)agg_le(offs + (pre + 1u) * r.cell_off_size, r.cell_off_size);
  if (b < a || pos + b > r.size) return false;
  d.p = cell + a;
  d.len = b - a;
  d.prefix = r.bitmaps == 2u && (cell[1] & mask) != 0;
  return true;
}

// ObSkipIndexCmpRes
struct Cmp {
  int cmp;
  bool certain;
  __device__ bool gt() const { return certain && cmp > 0; }
  __device__ bool lt() const { return certain && cmp < 0; }
  __device__ bool le() const { return certain && cmp <= 0; }
  __device__ bool ge() const { return certain && cmp >= 0; }
  __device__ bool eq() const { return certain && cmp == 0; }
};

__device__ __forceinline__ int bytes_cmp(const uint8_t *a, uint32_t alen, const uint8_t *b, uint32_t blen) {
  const uint32_t m = alen < blen ? alen : blen;
  for (uint32_t i = 0; i < m; ++i)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return alen < blen ? -1 : (alen > blen ? 1 : 0);
}

struct LeafCtx {
  const ScanParams *p;
  int sc;            // 1 signed / 2 unsigned integer class, 5 string
  uint32_t elem_len; // datum length of integer classes
};

__device__ __forceinline__ int64_t agg_image(const LeafCtx &lc, const AggDatum &d) {
  const uint64_t v = agg_le(d.p, d.len < 8u ? d.len : 8u);
  if (lc.elem_len == 4) return lc.sc == 1 ? (int64_t)(int32_t)(uint32_t)v : (int64_t)(uint32_t)v;
  if (lc.elem_len == 1) return (int64_t)(uint8_t)v;
  return (int64_t)v;
}
__device__ __forceinline__ bool img_less(const LeafCtx &lc, int64_t a, int64_t b) {
  return lc.sc == 1 ? a < b : (uint64_t)a < (uint64_t)b;   // images of narrow unsigned datums are zero-extended
}

// ObSkipIndexFilterExecutor::compare (+ compare_for_non_pad_charset / compare_with_prefix, binary collation):
// skip datum (min when is_min) against constant k of the leaf
__device__ __forceinline__ Cmp skip_compare(const LeafCtx &lc, const AggDatum &d, bool is_min, const ParamDev &pp) {
  Cmp r{0, false};
  if (!d.p) { r.certain = true; r.cmp = is_min ? -1 : 1; return r; }
  if (lc.sc != 5) {
    const int64_t a = agg_image(lc, d), c = pp.i64;
    r.cmp = img_less(lc, a, c) ? -1 : (img_less(lc, c, a) ? 1 : 0);
    r.certain = true;
    return r;
  }
  const uint8_t *c = lc.p->param_heap + pp.heap_off;
  r.cmp = bytes_cmp(d.p, d.len, c, pp.len);
  if (!d.prefix) { r.certain = true; return r; }
  if (r.cmp >= 0) { r.cmp = 1; r.certain = true; return r; }
  if (d.len >= pp.len) { r.certain = true; return r; }
  r.certain = bytes_cmp(d.p, d.len, c, d.len) == r.cmp;   // 0: the stored prefix is a prefix of the constant
  return r;
}

__device__ __forceinline__ uint8_t leaf_mask(const ScanParams &p, const FilterNodeDev &nd, const ColDesc &d, const AggRow &row,
                                             uint32_t row_count) {
  const int op = nd.op;
  if (op == OP_FALSE) return MASK_FALSE;   // NULL constant / empty IN list / empty range (host-resolved)
  if (op == OP_TRUE) return MASK_TRUE;
  if (!row.ok || !d.ok) return MASK_UNCERTAIN;
  LeafCtx lc{&p, (int)d.sc, (uint32_t)d.elem_len};
  if (lc.sc != 1 && lc.sc != 2 && lc.sc != 5) return MASK_UNCERTAIN;
  const uint32_t col = (uint32_t)p.used_col[nd.used_idx];
  AggDatum nc, mn, mx;
  if (!agg_row_read(row, col, 2u, nc) || !agg_row_read(row, col, 0u, mn) || !agg_row_read(row, col, 1u, mx)) return MASK_UNCERTAIN;
  if (!nc.p && !mn.p && !mx.p) return MASK_UNCERTAIN;
  int64_t null_count = 0;
  if (nc.p) {
    if (nc.len != 8u) return MASK_UNCERTAIN;
    null_count = (int64_t)agg_le(nc.p, 8);
    if (null_count < 0 || null_count > (int64_t)row_count) return MASK_UNCERTAIN;  // the reference raises OB_ERR_UNEXPECTED
  }
  if (lc.sc != 5 && ((mn.p && mn.len > 8u) || (mx.p && mx.len > 8u))) return MASK_UNCERTAIN;
  const bool all_null = nc.p && null_count == (int64_t)row_count;
  const bool all_not_null = nc.p && null_count == 0;
  const bool has_null = nc.p ? (null_count > 0 && null_count < (int64_t)row_count) : true;
  uint8_t m = MASK_UNCERTAIN;
  if (op == OP_NU) {
    m = all_not_null ? MASK_FALSE : (all_null ? MASK_TRUE : MASK_UNCERTAIN);
  } else if (op == OP_NN) {
    m = all_null ? MASK_FALSE : (all_not_null ? MASK_TRUE : MASK_UNCERTAIN);
  } else if (all_null) {
    m = MASK_FALSE;
  } else if (!mn.p && !mx.p) {
    m = MASK_UNCERTAIN;
  } else if (nd.range_ok) {
    // integer compare already reduced to "image in [lo, lo + span]" (NE: outside [c, c]); several range leaves of
    // one column under an AND arrive merged, which gives the same verdict as the AND of their verdicts
    const int64_t lo = (int64_t)nd.lo, hi = (int64_t)(nd.lo + nd.span);
    const int64_t a = mn.p ? agg_image(lc, mn) : 0, b = mx.p ? agg_image(lc, mx) : 0;
    const bool min_gt_hi = mn.p && img_less(lc, hi, a), max_lt_lo = mx.p && img_less(lc, b, lo);
    const bool inside = mn.p && mx.p && !img_less(lc, a, lo) && !img_less(lc, hi, b);
    if (nd.negate) m = (min_gt_hi || max_lt_lo) ? MASK_TRUE : (inside ? MASK_FALSE : MASK_UNCERTAIN);
    else m = (min_gt_hi || max_lt_lo) ? MASK_FALSE : (inside ? MASK_TRUE : MASK_UNCERTAIN);
  } else {
    const ParamDev *pr = p.params + nd.param_begin;
    Cmp a, b;
    switch (op) {
      case OP_EQ:
      case OP_NE: {
        const uint8_t hit = op == OP_EQ ? MASK_TRUE : MASK_FALSE, miss = op == OP_EQ ? MASK_FALSE : MASK_TRUE;
        a = skip_compare(lc, mn, true, pr[0]);
        if (a.gt()) { m = miss; break; }
        b = skip_compare(lc, mx, false, pr[0]);
        if (b.lt()) m = miss;
        else if (a.eq() && b.eq()) m = hit;
        break;
      }
      case OP_GT:
        a = skip_compare(lc, mn, true, pr[0]);
        if (a.gt()) { m = MASK_TRUE; break; }
        if (skip_compare(lc, mx, false, pr[0]).le()) m = MASK_FALSE;
        break;
      case OP_GE:
        a = skip_compare(lc, mn, true, pr[0]);
        if (a.ge()) { m = MASK_TRUE; break; }
        if (skip_compare(lc, mx, false, pr[0]).lt()) m = MASK_FALSE;
        break;
      case OP_LT:
        a = skip_compare(lc, mn, true, pr[0]);
        if (a.ge()) { m = MASK_FALSE; break; }
        if (skip_compare(lc, mx, false, pr[0]).lt()) m = MASK_TRUE;
        break;
      case OP_LE:
        a = skip_compare(lc, mn, true, pr[0]);
        if (a.gt()) { m = MASK_FALSE; break; }
        if (skip_compare(lc, mx, false, pr[0]).le()) m = MASK_TRUE;
        break;
      case OP_BT:
        if (skip_compare(lc, mn, true, pr[1]).gt()) { m = MASK_FALSE; break; }
        if (skip_compare(lc, mx, false, pr[0]).lt()) { m = MASK_FALSE; break; }
        if (skip_compare(lc, mn, true, pr[0]).ge() && skip_compare(lc, mx, false, pr[1]).le()) m = MASK_TRUE;
        break;
      case OP_IN: {
        // in_operator (:720-781) over the sorted constants; restated without the sort: the constant the
        // reference's lower / upper bound lands on is the smallest one that is >= min (> min for a min prefix)
        int best = -1;
        bool equal = false;
        for (int k = 0; k < nd.n_params; ++k) {
          int c = -1;   // min vs constant k (min missing: -infinity)
          if (mn.p) {
            if (lc.sc == 5) c = bytes_cmp(mn.p, mn.len, p.param_heap + pr[k].heap_off, pr[k].len);
            else { const int64_t x = agg_image(lc, mn); c = img_less(lc, x, pr[k].i64) ? -1 : (img_less(lc, pr[k].i64, x) ? 1 : 0); }
          }
          if (mn.prefix ? c >= 0 : c > 0) continue;   // constant below the bound
          bool smaller = best < 0;
          if (!smaller) {
            if (lc.sc == 5) smaller = bytes_cmp(p.param_heap + pr[k].heap_off, pr[k].len, p.param_heap + pr[best].heap_off, pr[best].len) < 0;
            else smaller = img_less(lc, pr[k].i64, pr[best].i64);
          }
          if (smaller) { best = k; equal = !mn.prefix && mn.p && c == 0; }
        }
        if (best < 0) { m = MASK_FALSE; break; }
        b = skip_compare(lc, mx, false, pr[best]);
        if (b.gt()) m = MASK_UNCERTAIN;
        else if (b.lt()) m = MASK_FALSE;
        else if (equal) m = mx.prefix ? MASK_UNCERTAIN : MASK_TRUE;
        break;
      }
      default:
        break;
    }
  }
  if (has_null && m == MASK_TRUE) m = MASK_UNCERTAIN;
  return m;
}

// one thread per block: leaf verdicts -> tree verdict
__global__ void __launch_bounds__(128) skip_index_kernel(const __grid_constant__ ScanParams p, const uint8_t *agg,
                                                         const int64_t *agg_off, uint8_t *blk_const, uint8_t *leaf_const,
                                                         unsigned long long *counters) {
  const int block = blockIdx.x * blockDim.x + threadIdx.x;
  if (block >= p.n_blocks) return;
  const int64_t a0 = agg_off[block], a1 = agg_off[block + 1];
  AggRow row;
  agg_row_init(agg + a0, (uint32_t)(a1 - a0), row);
  const uint32_t rows = p.rows[block];
  uint32_t stack_t = 0, stack_f = 0;   // bit i: entry i of the evaluation stack is always-true / always-false
  for (int i = 0; i < p.n_nodes; ++i) {
    const FilterNodeDev &nd = p.nodes[i];
    uint8_t m;
    if (nd.kind == NODE_WHITE) {
      m = leaf_mask(p, nd, p.plans[(int64_t)block * p.max_cols + p.used_col[nd.used_idx]], row, rows);
      if (rows == 0) m = MASK_UNCERTAIN;
    } else {
      const uint32_t k = (uint32_t)nd.n_children, full = (1u << k) - 1u;
      const uint32_t t = stack_t & full, f = stack_f & full;
      if (nd.kind == NODE_AND) m = f ? MASK_FALSE : (t == full ? MASK_TRUE : MASK_UNCERTAIN);
      else m = t ? MASK_TRUE : (f == full ? MASK_FALSE : MASK_UNCERTAIN);
      stack_t >>= k;
      stack_f >>= k;
    }
    leaf_const[(int64_t)block * p.n_nodes + i] = m;
    stack_t = (stack_t << 1) | (m == MASK_TRUE ? 1u : 0u);
    stack_f = (stack_f << 1) | (m == MASK_FALSE ? 1u : 0u);
  }
  const uint8_t verdict = (stack_t & 1u) ? MASK_TRUE : ((stack_f & 1u) ? MASK_FALSE : MASK_UNCERTAIN);
  blk_const[block] = verdict;
  if (counters && verdict != MASK_UNCERTAIN) atomicAdd(&counters[verdict == MASK_FALSE ? 0 : 1], 1ull);
}

}  // namespace skipidx
