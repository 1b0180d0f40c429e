// Small micro-blocks (a 16 KiB block of a wide dictionary-coded table holds ~130 rows): per block there are only a
// few hundred bytes to filter and a dozen rows to project, so a scan is bound by the per-block chain of dependent
// global round trips (record -> plans -> column bytes), not by bandwidth. These two kernels keep ONE WARP per block
// but run it as a software pipeline over the blocks it owns (persistent grid, blocks strided over the warps):
//
//     iteration b:   wait   regions(b), meta(b + 1)          (cp.async groups, issued one / two iterations ago)
//                    issue  regions(b + 1)   <- needs meta(b + 1): block offset, decode plans -> column byte ranges
//                    issue  meta(b + 2)      <- block record, decode plans (and the two prefix entries)
//                    work   on block b from shared memory
//
// so a block's three round trips overlap the work on the two blocks before it. All copies are 16-byte cp.async
// (LDGSTS: no registers, no mbarrier), completion is cp.async.wait_group + __syncwarp.
//   obgpu_count_pipe_kernel   : stages every filter column's region of the block at once, then the same leaf loops
//                               as obgpu_count_kernel (K4 / K6 / K9 / K14)
//   obgpu_project_pipe_kernel : stages the block's bitmap words and the projected columns' byte ranges; a projected
//                               VARCHAR dictionary column needs only its refs and its offset array (VEC_DISCRETE
//                               output is pointers into the caller's block: the dictionary's bytes are never read)
#pragma once

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int kPipeMaxFilterCols = 8;
constexpr uint32_t kCountHdrBytes = 64u;   // count region slot: deltas + flags | regions

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- lean leaf pieces for blocks of at most 1024 rows (bitmap in registers, lane g owns word g) ----------------------
// Range leaf over the fixed-width integer dictionary of a K_DICT column -> predicate bitset (bit count = NULL: never set).
__device__ __forceinline__ void lean_bitset_int_range(const ColDesc &d, const FilterNodeDev &nd, uint32_t sbit, uint32_t *bits, int lane) {
  const uint32_t n = d.dict_count, dbits = d.dict_data_size * 8u, dpay = sbit + d.dict_payload * 8u;
  const uint64_t lo = nd.lo, span = nd.span, base = d.base, mask = d.int_mask;
  const bool fix = d.sign_fix != 0, neg = nd.negate != 0;
  const uint32_t el = d.elem_len, sc = d.sc;
  for (uint32_t b0 = 0; b0 < n + 2u; b0 += 32u) {
    const uint32_t idx = b0 + (uint32_t)lane;
    bool r = false;
    if (idx < n) {
      uint64_t v = (dbits <= 32u ? (uint64_t)sbits32(dpay + idx * dbits, dbits) : sbits(dpay + idx * dbits, dbits)) + base;
      if (fix) v = sign_fix(mask, v);
      if (el == 4) v = sc == 1 ? (uint64_t)(int64_t)(int32_t)(uint32_t)v : (uint64_t)(uint32_t)v;
      else if (el == 1) v = (uint64_t)(uint8_t)v;
      r = ((v - lo) <= span) != neg;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, r);
    if (lane == 0) bits[b0 >> 5] = word;
  }
}

// Rows of a K_DICT column against a predicate over refs -- a bitset in shared memory (BITSET) or a ref interval
// [a, e), complemented inside [0, count) when neg (sorted dictionary). Lane g collects word g; returns the lane's
// updated bitmap word.
template <bool BITSET>
__device__ __forceinline__ uint32_t lean_rows(const ColDesc &d, uint32_t sbit, const uint32_t *bits, uint32_t a, uint32_t e, bool neg,
                                              uint32_t mybm, uint32_t myvalid, uint32_t nwords, bool and_mode, int lane) {
  const uint32_t step = 32u * d.stride, width = d.width, dcount = d.dict_count, cntp1 = dcount + 1u, span = e - a;
  uint32_t bit = sbit + d.val_bit + (uint32_t)lane * d.stride, acc = 0;
  for (uint32_t g = 0; g < nwords; ++g, bit += step) {
    // lanes past the last row read a few refs beyond the column (inside the warp's shared memory): masked by myvalid
    const uint32_t ref = sbits32(bit, width);
    bool hit;
    if (BITSET) {
      const uint32_t rr = min(ref, cntp1);
      hit = (bits[rr >> 5] >> (rr & 31u)) & 1u;
    } else {
      hit = ((ref - a) < span) != (neg && ref < dcount);
    }
    const uint32_t w = __ballot_sync(0xffffffffu, hit);
    if ((uint32_t)lane == g) acc = w;
  }
  acc &= myvalid;
  return and_mode ? (mybm & acc) : (mybm | acc);
}

// Sorted fixed-width integer dictionary and a range leaf: the matching refs are [a, e) = [#entries below the range,
// #entries not above it) (the reference binary-searches the bounds, ob_dict_decoder.cpp:967-988,1085-1176).
__device__ __forceinline__ void lean_interval_sorted_int(const ColDesc &d, const FilterNodeDev &nd, uint32_t sbit, int lane,
                                                         uint32_t &a, uint32_t &e) {
  const uint32_t n = d.dict_count, dbits = d.dict_data_size * 8u, dpay = sbit + d.dict_payload * 8u;
  const uint64_t lo = nd.lo, hi = nd.lo + nd.span, base = d.base, mask = d.int_mask;
  const bool fix = d.sign_fix != 0, sg = d.sc == 1;
  const uint32_t el = d.elem_len;
  uint32_t below = 0, not_above = 0;
  for (uint32_t b0 = 0; b0 < n; b0 += 32u) {
    const uint32_t idx = b0 + (uint32_t)lane;
    bool lt = false, le = false;
    if (idx < n) {
      uint64_t v = (dbits <= 32u ? (uint64_t)sbits32(dpay + idx * dbits, dbits) : sbits(dpay + idx * dbits, dbits)) + base;
      if (fix) v = sign_fix(mask, v);
      if (el == 4) v = sg ? (uint64_t)(int64_t)(int32_t)(uint32_t)v : (uint64_t)(uint32_t)v;
      else if (el == 1) v = (uint64_t)(uint8_t)v;
      lt = sg ? (int64_t)v < (int64_t)lo : v < lo;
      le = sg ? (int64_t)v <= (int64_t)hi : v <= hi;
    }
    below += __popc(__ballot_sync(0xffffffffu, lt));
    not_above += __popc(__ballot_sync(0xffffffffu, le));
  }
  a = below;
  e = not_above < below ? below : not_above;
}

// EQ / NE / IN leaf over the string dictionary of a K_DICT column -> predicate bitset. Every entry is screened by
// (length, first 8 bytes) against the constants; only the (few) entries that pass compare their tails, 8 bytes at a
// time, out of shared memory.
__device__ __forceinline__ void lean_bitset_str_eq(const ScanParams &p, const FilterNodeDev &nd, const ColDesc &d, uint32_t sbit,
                                                   uint32_t *bits, int lane) {
  const uint32_t n = d.dict_count, ib8 = d.dict_data_size * 8u, ibit = sbit + d.dict_payload * 8u, heap_len = d.dict_end - d.dict_var;
  const bool fixed = d.dict_fixed != 0, ne = nd.op == OP_NE;
  const int hm = nd.pad;   // 0: first-byte screen + constant list; m + 1: hash slots under multiplier m
  for (uint32_t b0 = 0; b0 < n + 2u; b0 += 32u) {
    const uint32_t idx = b0 + (uint32_t)lane;
    bool r = false;
    if (idx < n) {
      uint32_t cell, len;
      if (fixed) {
        len = d.dict_data_size;
        cell = d.dict_payload + idx * len;
      } else {
        const uint32_t off = idx == 0 ? 0u : sbits32(ibit + (idx - 1u) * ib8, ib8);
        const uint32_t end = idx == n - 1u ? heap_len : sbits32(ibit + idx * ib8, ib8);
        cell = d.dict_var + off;
        len = end - off;
      }
      const uint32_t pl = len < 8u ? len : 8u;
      const uint64_t pre = pl ? sbits(sbit + cell * 8u, pl * 8u) : 0ull;
      // screens built on the host: does any constant have this length (most entries stop here), then either the hash slot of
      // (length, first 8 bytes) -- it names the only constant the entry can equal -- or the first-byte screen + the list
      const bool len_ok = (nd.lo >> (len & 63u)) & 1ull;
      int k0 = 0, k1 = 0;
      if (len_ok && hm) {
        const uint32_t h = str_eq_slot(pre, len, hm - 1);
        if ((nd.span >> h) & 1ull) { k0 = __popcll(nd.span & ((1ull << h) - 1ull)); k1 = k0 + 1; }
      } else if (len_ok && (len == 0u || ((nd.span >> (pre & 63ull)) & 1ull))) {
        k1 = nd.n_params;
      }
      for (int k = k0; k < k1 && !r; ++k) {
        const ParamDev &pp = p.params[nd.param_begin + k];
        if (pp.len != len || (uint64_t)pp.i64 != pre) continue;
        bool same = true;
        for (uint32_t i = 8u; i < len && same; i += 8u) {
          const uint32_t nb = len - i < 8u ? len - i : 8u;
          const uint64_t a = sbits(sbit + (cell + i) * 8u, nb * 8u);
          const uint64_t c = *reinterpret_cast<const uint64_t *>(p.param_heap + pp.heap_off + i) & (~0ull >> (64u - nb * 8u));
          same = a == c;
        }
        r = same;
      }
      r = r != ne;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, r);
    if (lane == 0) bits[b0 >> 5] = word;
  }
}

// AND leaf on a string K_DICT column when few rows are still alive: evaluate the leaf on the survivors' own
// dictionary entries (one pass over <= alive rows) instead of on every dictionary entry.
__device__ __forceinline__ uint32_t lean_survivor_str(const ScanParams &p, const FilterNodeDev &nd, const ColDesc &d, uint32_t sbit,
                                                      const uint8_t *rs_generic, uint32_t mybm, uint32_t nwords, uint32_t alive,
                                                      uint32_t *bm, int lane) {
  // survivor list (ascending rows) in the spilled-bitmap scratch: uint16 rows after the 32 bitmap words
  uint16_t *list = reinterpret_cast<uint16_t *>(bm + 32);
  const uint32_t local = __popc(mybm);
  uint32_t inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  const uint32_t excl = inc - local;
  for (uint32_t g = 0; g < nwords; ++g) {
    const uint32_t wg = __shfl_sync(0xffffffffu, mybm, g), og = __shfl_sync(0xffffffffu, excl, g);
    if ((wg >> lane) & 1u) list[og + __popc(wg & ((1u << lane) - 1u))] = (uint16_t)(g * 32u + (uint32_t)lane);
  }
  bm[lane] = mybm;
  __syncwarp();
  const uint32_t vbit = sbit + d.val_bit, stride = d.stride, width = d.width, dcount = d.dict_count;
  const uint32_t ib8 = d.dict_data_size * 8u, ibit = sbit + d.dict_payload * 8u, heap_len = d.dict_end - d.dict_var;
  const int op = nd.op;
  const uint8_t *s = rs_generic;   // generic pointer for the (rare) full string compare
  for (uint32_t j = (uint32_t)lane; j < alive; j += 32u) {
    const uint32_t row = list[j];
    const uint32_t ref = sbits32(vbit + row * stride, width);
    bool pass;
    if (ref >= dcount) pass = op == OP_NU;
    else if (op == OP_NU) pass = false;
    else if (op == OP_NN) pass = true;
    else {
      uint32_t cell, len;
      if (d.dict_fixed) {
        len = d.dict_data_size;
        cell = d.dict_payload + ref * len;
      } else {
        const uint32_t off = ref == 0 ? 0u : sbits32(ibit + (ref - 1u) * ib8, ib8);
        const uint32_t end = ref == dcount - 1u ? heap_len : sbits32(ibit + ref * ib8, ib8);
        cell = d.dict_var + off;
        len = end - off;
      }
      if (op == OP_EQ || op == OP_NE || op == OP_IN) {
        const uint32_t pl = len < 8u ? len : 8u;
        const uint64_t pre = pl ? sbits(sbit + cell * 8u, pl * 8u) : 0ull;
        bool hit = false;
        const bool len_ok = (nd.lo >> (len & 63u)) & 1ull;
        int k0 = 0, k1 = 0;
        if (len_ok && nd.pad) {
          const uint32_t h = str_eq_slot(pre, len, nd.pad - 1);
          if ((nd.span >> h) & 1ull) { k0 = __popcll(nd.span & ((1ull << h) - 1ull)); k1 = k0 + 1; }
        } else if (len_ok && (len == 0u || ((nd.span >> (pre & 63ull)) & 1ull))) {
          k1 = nd.n_params;
        }
        for (int k = k0; k < k1 && !hit; ++k) {
          const ParamDev &pp = p.params[nd.param_begin + k];
          hit = pp.len == len && (uint64_t)pp.i64 == pre && (len <= 8u || str_cmp(s, cell, len, p.param_heap + pp.heap_off, pp.len) == 0);
        }
        pass = hit != (op == OP_NE);
      } else {
        pass = str_pred(p, nd, s, cell, len);
      }
    }
    if (!pass) atomicAnd(&bm[row >> 5], ~(1u << (row & 31u)));
  }
  __syncwarp();
  return (uint32_t)lane < nwords ? bm[lane] : 0u;
}

// =================================================================================================
// Count, pipelined. Per warp: meta ring (3 slots: block record + the filter columns' plans), region ring (2 slots:
// header with the per-column deltas + the filter columns' regions), bitmap words, predicate bitsets.
// =================================================================================================
__global__ void __launch_bounds__(kThreads) obgpu_count_pipe_kernel(const __grid_constant__ ScanParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps_total = (int)gridDim.x * kWarps;
  int blk = (int)blockIdx.x * kWarps + warp;
  if (blk >= p.n_blocks) return;
  uint8_t *wr = g_smem + (uint32_t)warp * p.pc_bytes;
  uint8_t *meta0 = wr + p.pc_meta, *reg0 = wr + p.pc_region;
  uint32_t *bm = reinterpret_cast<uint32_t *>(wr + p.pc_bm);
  uint32_t *bitsets = reinterpret_cast<uint32_t *>(wr + p.pc_bitset);
  uint64_t *bars = reinterpret_cast<uint64_t *>(wr + p.pc_bar);   // one mbarrier per region slot
  const int nf = p.pf_n;
  Team t;
  t.tid = lane; t.nthreads = 32; t.warp = 0; t.nwarps = 1; t.lane = lane; t.bar_id = -1;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_barrier_init();
  }
  __syncwarp();

  auto issue_meta = [&](int b, int slot) {
    if (b >= p.n_blocks) return;
    const uint32_t sa = smem_u32(meta0 + (uint32_t)slot * p.pc_meta_bytes);
    const uint8_t *rec = reinterpret_cast<const uint8_t *>(p.recs + b);
    const uint8_t *plans = reinterpret_cast<const uint8_t *>(p.plans + (int64_t)b * p.max_cols);
    for (int q = lane; q < 3 + nf * 6; q += 32) {
      if (q < 3) cp_async16(sa + (uint32_t)q * 16u, rec + q * 16);
      else {
        const int i = (q - 3) / 6, piece = (q - 3) % 6;
        cp_async16(sa + 64u + (uint32_t)i * 96u + (uint32_t)piece * 16u, plans + (size_t)p.used_col[i] * sizeof(ColDesc) + piece * 16);
      }
    }
  };
  // region slot: [int32 delta[kPipeMaxFilterCols] | uint32 flags] (64 bytes) then the regions at p.pf_off[i]
  auto issue_regions = [&](int b, int mslot, int rslot, uint32_t verdict) {
    if (b >= p.n_blocks) return;
    const uint8_t *m = meta0 + (uint32_t)mslot * p.pc_meta_bytes;
    uint8_t *rs = reg0 + (uint32_t)rslot * p.pc_region_bytes;
    const BlockRec &rec = *reinterpret_cast<const BlockRec *>(m);
    const ColDesc *descs = reinterpret_cast<const ColDesc *>(m + 64);
    int32_t *hdr = reinterpret_cast<int32_t *>(rs);
    uint32_t lo = 0, hi = 0;
    bool bad = false;
    if (lane < nf && rec.rows != 0 && verdict == 0) {
      BlockView bv;
      view_from_rec(rec, nullptr, bv);
      const ColDesc &d = descs[lane];
      if (!d.ok || !col_region(d, bv, lo, hi) || hi - lo > p.pf_span[lane] || hi > ((rec.size + 15u) & ~15u) + 32u) { bad = true; lo = hi = 0; }
      hdr[lane] = (int32_t)(kCountHdrBytes + p.pf_off[lane]) - (int32_t)lo;
    }
    const uint32_t badmask = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) hdr[kPipeMaxFilterCols] = (int32_t)badmask;
    // one bulk copy (TMA) per filter column, all completing on the slot's mbarrier
    const uint32_t total = warp_sum_u32(hi - lo);
    uint64_t *bar = bars + rslot;
    if (lane == 0) mbar_expect_tx(bar, total);
    __syncwarp();
    if (hi > lo) tma_bulk_g2s(rs + kCountHdrBytes + p.pf_off[lane], p.image + rec.off + lo, hi - lo, bar);
  };

  // prologue: meta(b0), then regions(b0) + meta(b1)
  uint32_t v_cur = 0, v_next = 0, v_next2 = 0;   // skip-index verdicts, fetched two blocks ahead
  if (p.blk_const != nullptr) {
    v_cur = p.blk_const[blk];
    if (blk + nwarps_total < p.n_blocks) v_next = p.blk_const[blk + nwarps_total];
  }
  issue_meta(blk, 0);
  cp_async_commit();
  cp_async_wait_all();
  __syncwarp();
  issue_regions(blk, 0, 0, v_cur);
  cp_async_commit();
  issue_meta(blk + nwarps_total, 1);
  cp_async_commit();
  int it = 0;
  for (; blk < p.n_blocks; blk += nwarps_total, ++it) {
    const int ms = it % 3, rsl = it & 1;
    cp_async_wait_all();
    mbar_wait(bars + rsl, (uint32_t)(it >> 1) & 1u);
    __syncwarp();
    const int b2 = blk + 2 * nwarps_total;
    if (p.blk_const != nullptr && b2 < p.n_blocks) v_next2 = p.blk_const[b2];
    issue_regions(blk + nwarps_total, (it + 1) % 3, rsl ^ 1, v_next);
    cp_async_commit();
    issue_meta(b2, (it + 2) % 3);
    cp_async_commit();

    // ---- block `blk` from shared memory -------------------------------------------------------------------------
    const uint8_t *m = meta0 + (uint32_t)ms * p.pc_meta_bytes;
    uint8_t *rs = reg0 + (uint32_t)rsl * p.pc_region_bytes;
    const BlockRec rec = *reinterpret_cast<const BlockRec *>(m);
    const ColDesc *descs = reinterpret_cast<const ColDesc *>(m + 64);
    const int32_t *hdr = reinterpret_cast<const int32_t *>(rs);
    const uint32_t rows = rec.rows;
    uint32_t *gbm = p.bitmap_words + rec.bm_word_off;
    const uint32_t nwords = (rows + 31u) >> 5;
    const uint32_t verdict = v_cur;
    v_cur = v_next;
    v_next = v_next2;
    if (verdict != 0 && rows != 0) {   // decided by the skip index: the block was not read
      for (uint32_t g = (uint32_t)lane; g < nwords; g += 32u) gbm[g] = verdict == 1 ? valid_mask_of(rows, g) : 0u;
      if (lane == 0) p.counts[blk] = verdict == 1 ? rows : 0u;
      __syncwarp();
      continue;
    }
    if (rows == 0 || hdr[kPipeMaxFilterCols] != 0) {
      if (lane == 0) {
        atomicOr(p.status, rows == 0 ? ST_CORRUPT : ST_UNSUPPORTED);
        p.counts[blk] = 0;
      }
      __syncwarp();
      continue;
    }
    BlockCtx c;
    view_from_rec(rec, nullptr, c.b);
    c.descs = descs;
    c.bitsets = bitsets;
    c.rle_base = nullptr;
    c.rle_slot_bytes = c.rle_starts_bytes = 0;
    const bool and_mode = p.simple_shape == 1;
    const int n_leaves = p.n_nodes == 1 ? 1 : p.n_nodes - 1;
    uint32_t cnt = 0;
    if (nwords <= 32u) {
      // ---- lean path: the block's bitmap lives in registers (lane g owns word g); dictionary-coded leaves run
      // through explicit shared-memory loads, everything else through the generic leaf code on a spilled bitmap
      const uint32_t myvalid = (uint32_t)lane < nwords ? valid_mask_of(rows, (uint32_t)lane) : 0u;
      uint32_t mybm = and_mode ? myvalid : 0u;
      for (int i = 0; i < n_leaves; ++i) {
        const FilterNodeDev &nd = p.nodes[i];
        if (p.leaf_const != nullptr && p.leaf_const[(int64_t)blk * p.n_nodes + i] != 0) continue;
        const ColDesc &d = descs[nd.used_idx];
        const uint32_t sbit = (smem_u32(rs) + (uint32_t)hdr[nd.used_idx]) * 8u;
        if (d.kind == K_DICT && nd.slot >= 0 && nd.op != OP_FALSE && nd.op != OP_TRUE && d.width <= 32u) {
          uint32_t *bits = bitsets + nd.slot * p.bitset_words;
          const bool is_str = d.sc == 5;
          if (is_str && and_mode) {
            // few surviving rows and a larger dictionary: test the survivors' own entries instead of every entry
            const uint32_t alive = warp_sum_u32(__popc(mybm));
            if (alive * 4u <= d.dict_count) {
              mybm = lean_survivor_str(p, nd, d, sbit, rs + hdr[nd.used_idx], mybm, nwords, alive, bm, lane);
              continue;
            }
          }
          if (!is_str && nd.range_ok && d.dict_sorted) {
            uint32_t a, e;
            lean_interval_sorted_int(d, nd, sbit, lane, a, e);
            mybm = lean_rows<false>(d, sbit, nullptr, a, e, nd.negate != 0, mybm, myvalid, nwords, and_mode, lane);
            goto leaf_done;
          }
          if (!is_str && nd.range_ok) lean_bitset_int_range(d, nd, sbit, bits, lane);
          else if (is_str && (nd.op == OP_EQ || nd.op == OP_NE || nd.op == OP_IN)) lean_bitset_str_eq(p, nd, d, sbit, bits, lane);
          else {
            c.b.s = rs + hdr[nd.used_idx];
            c.sbit = sbit;
            build_dict_bitset(p, c.b, d, nd, bits, t);
          }
          __syncwarp();
          mybm = lean_rows<true>(d, sbit, bits, 0u, 0u, false, mybm, myvalid, nwords, and_mode, lane);
        } else {
          bm[lane] = mybm;   // words_cap >= 32 words are reserved for the spilled bitmap
          __syncwarp();
          if (nd.op != OP_FALSE && nd.op != OP_TRUE) {
            c.b.s = rs + hdr[nd.used_idx];
            c.sbit = sbit;
          }
          if (nd.slot >= 0 && is_dict_kind(d)) {
            build_dict_bitset(p, c.b, d, nd, bitsets + nd.slot * p.bitset_words, t);
            __syncwarp();
          }
          leaf_over_words<false>(p, c, nd, bm, rows, nwords, and_mode, t);
          __syncwarp();
          mybm = (uint32_t)lane < nwords ? bm[lane] : 0u;
        }
      leaf_done:
        if (i + 1 < n_leaves && !__any_sync(0xffffffffu, and_mode ? mybm != 0u : mybm != myvalid)) break;   // early-out
      }
      if ((uint32_t)lane < nwords) gbm[lane] = mybm;
      cnt = __popc(mybm);
    } else {
      bool inited = false;
      for (int i = 0; i < n_leaves; ++i) {
        const FilterNodeDev &nd = p.nodes[i];
        if (p.leaf_const != nullptr && p.leaf_const[(int64_t)blk * p.n_nodes + i] != 0) continue;
        if (nd.op != OP_FALSE && nd.op != OP_TRUE) {   // block-relative offsets of this leaf's column resolve into its staged region
          c.b.s = rs + hdr[nd.used_idx];
          c.sbit = (smem_u32(rs) + (uint32_t)hdr[nd.used_idx]) * 8u;
        }
        const ColDesc &d = descs[nd.used_idx];
        if (nd.slot >= 0 && is_dict_kind(d)) {
          build_dict_bitset(p, c.b, d, nd, bitsets + nd.slot * p.bitset_words, t);
          __syncwarp();
        }
        if (i == 0 && leaf_first_fast<false>(p, c, nd, bm, rows, nwords, t)) {
          inited = true;
          __syncwarp();
          continue;
        }
        if (!inited) {
          for (uint32_t g = (uint32_t)lane; g < nwords; g += 32u) bm[g] = and_mode ? valid_mask_of(rows, g) : 0u;
          inited = true;
          __syncwarp();
        }
        leaf_over_words<false>(p, c, nd, bm, rows, nwords, and_mode, t);
        __syncwarp();
        if (i + 1 < n_leaves) {   // early-out of the AND / OR (ob_pushdown_filter.cpp:1603-1615)
          bool undecided = false;
          for (uint32_t g = (uint32_t)lane; g < nwords; g += 32u)
            undecided = undecided || (and_mode ? bm[g] != 0u : bm[g] != valid_mask_of(rows, g));
          if (!__any_sync(0xffffffffu, undecided)) break;
        }
      }
      if (!inited) {
        for (uint32_t g = (uint32_t)lane; g < nwords; g += 32u) bm[g] = and_mode ? valid_mask_of(rows, g) : 0u;
        __syncwarp();
      }
      for (uint32_t g = (uint32_t)lane; g < nwords; g += 32u) {
        const uint32_t w = bm[g];
        gbm[g] = w;
        cnt += __popc(w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) p.counts[blk] = cnt;
    __syncwarp();   // every lane is done with this iteration's slots before the next iteration refills them
  }
  cp_async_wait_all();
}

// =================================================================================================
// Projection, pipelined. Meta slot: [record 48 | sel_offset pair 16 | plans n_proj x 96]; region slot:
// [header: per column {delta0, delta1} + flags | bitmap words | column byte ranges at p.pp_off[pc]].
// =================================================================================================
#define ROW(j) (IDENT ? (uint32_t)(j) : (uint32_t)sel[j])
template <bool IDENT>
__device__ __forceinline__ void project_str_dict_shallow(const ScanParams &p, const ColDesc &d, int pc, const uint16_t *sel,
                                                         uint32_t cnt, int64_t base_row, uint64_t blk_addr, uint32_t ref_sbit,
                                                         uint32_t idx_sbit, int lane) {
  uint64_t *optr = reinterpret_cast<uint64_t *>(p.out_data[pc]) + base_row;
  int32_t *olen = p.out_lens[pc] + base_row;
  const uint32_t vbit = ref_sbit + d.val_bit, stride = d.stride, width = d.width, dcount = d.dict_count;
  const uint32_t ib8 = d.dict_data_size * 8u, ibit = idx_sbit + d.dict_payload * 8u, heap_len = d.dict_end - d.dict_var;
  const bool fixed = d.dict_fixed != 0;
  bool saw_null = false;
  for (uint32_t j = (uint32_t)lane; j < cnt; j += 32u) {
    const uint32_t ref = sbits32(vbit + ROW(j) * stride, width);
    uint32_t cell = 0, len = 0;
    const bool is_null = ref >= dcount;
    if (!is_null) {
      if (fixed) {
        len = d.dict_data_size;
        cell = d.dict_payload + ref * len;
      } else {
        const uint32_t off = ref == 0 ? 0u : sbits32(ibit + (ref - 1u) * ib8, ib8);
        const uint32_t end = ref == dcount - 1u ? heap_len : sbits32(ibit + ref * ib8, ib8);
        cell = d.dict_var + off;
        len = end - off;
      }
    }
    __stcs(&optr[j], is_null ? 0ull : blk_addr + cell);
    __stcs(&olen[j], is_null ? 0 : (int32_t)len);
    if (is_null) {
      const int64_t o = base_row + (int64_t)j;
      atomicOr(&p.out_nulls[pc][o >> 5], 1u << (o & 31));
      saw_null = true;
    }
  }
  if (saw_null) p.has_null[pc] = 1;
}
template <bool IDENT>
__device__ __forceinline__ void lean_project_dict_i64(const ScanParams &p, const ColDesc &d, int pc, const uint16_t *sel, uint32_t cnt,
                                                      int64_t base_row, uint32_t sbit, int lane) {
  uint64_t *out = reinterpret_cast<uint64_t *>(p.out_data[pc]) + base_row;
  const uint32_t vbit = sbit + d.val_bit, stride = d.stride, width = d.width, dcount = d.dict_count;
  const uint32_t dbits = d.dict_data_size * 8u, dpay = sbit + d.dict_payload * 8u;
  const uint64_t dbase = d.base, mask = d.int_mask;
  const bool fix = d.sign_fix != 0;
  bool saw_null = false;
  for (uint32_t j = (uint32_t)lane; j < cnt; j += 32u) {
    const uint32_t ref = sbits32(vbit + ROW(j) * stride, width);
    uint64_t v = 0;
    if (ref >= dcount) {
      const int64_t o = base_row + (int64_t)j;
      atomicOr(&p.out_nulls[pc][o >> 5], 1u << (o & 31));
      saw_null = true;
    } else {
      v = (dbits <= 32u ? (uint64_t)sbits32(dpay + ref * dbits, dbits) : sbits(dpay + ref * dbits, dbits)) + dbase;
      if (fix) v = sign_fix(mask, v);
    }
    __stcs(&out[j], v);
  }
  if (saw_null) p.has_null[pc] = 1;
}
#undef ROW

__global__ void __launch_bounds__(kThreads) obgpu_project_pipe_kernel(const __grid_constant__ ScanParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps_total = (int)gridDim.x * kWarps;
  int blk = (int)blockIdx.x * kWarps + warp;
  if (blk >= p.n_blocks) return;
  uint8_t *wr = g_smem + (uint32_t)warp * p.pp_bytes;
  uint8_t *meta0 = wr + p.pp_meta, *reg0 = wr + p.pp_region;
  uint16_t *sel = reinterpret_cast<uint16_t *>(wr + p.pp_sel);
  uint8_t *wscr = wr + p.pp_wscr;
  uint64_t *bars = reinterpret_cast<uint64_t *>(wr + p.pp_bar);   // one mbarrier per region slot
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const int np = p.n_proj;
  const uint32_t hdr_bytes = p.pp_hdr_bytes, bm_bytes = p.pp_bm_bytes;
  Team t;
  t.tid = lane; t.nthreads = 32; t.warp = 0; t.nwarps = 1; t.lane = lane; t.bar_id = -1;

  auto issue_meta = [&](int b, int slot) {
    if (b >= p.n_blocks) return;
    const uint32_t sa = smem_u32(meta0 + (uint32_t)slot * p.pp_meta_bytes);
    const uint8_t *rec = reinterpret_cast<const uint8_t *>(p.recs + b);
    const uint8_t *plans = reinterpret_cast<const uint8_t *>(p.plans + (int64_t)b * p.max_cols);
    for (int q = lane; q < 5 + np * 6; q += 32) {
      if (q < 3) cp_async16(sa + (uint32_t)q * 16u, rec + q * 16);
      else if (q < 5) cp_async8(sa + 48u + (uint32_t)(q - 3) * 8u, p.sel_offset + b + (q - 3));
      else {
        const int i = (q - 5) / 6, piece = (q - 5) % 6;
        cp_async16(sa + 64u + (uint32_t)i * 96u + (uint32_t)piece * 16u,
                   plans + (size_t)p.used_col[p.proj_used[i]] * sizeof(ColDesc) + piece * 16);
      }
    }
  };
  // what a block needs beyond its meta: nothing (no selected row / overflow / corrupt), or bitmap words + column ranges
  auto issue_regions = [&](int b, int mslot, int rslot) {
    if (b >= p.n_blocks) return;
    const uint8_t *m = meta0 + (uint32_t)mslot * p.pp_meta_bytes;
    uint8_t *rs = reg0 + (uint32_t)rslot * p.pp_region_bytes;
    const BlockRec &rec = *reinterpret_cast<const BlockRec *>(m);
    const int64_t base = *reinterpret_cast<const int64_t *>(m + 48);
    const uint32_t cnt = (uint32_t)(*reinterpret_cast<const int64_t *>(m + 56) - base);
    const uint32_t rows = rec.rows;
    if (rows == 0 || cnt == 0 || base + (int64_t)cnt > p.out_cap) {
      if (lane == 0) mbar_expect_tx(bars + rslot, 0u);   // nothing to stage: the slot's phase still completes
      return;
    }
    if (cnt != rows) {
      const uint32_t nwords = (rows + 31u) >> 5;
      const uint32_t *gbm = p.bitmap_words + rec.bm_word_off;
      for (uint32_t g = (uint32_t)lane; g < nwords; g += 32u) cp_async4(smem_u32(rs) + hdr_bytes + g * 4u, gbm + g);
    }
    const ColDesc *plans = reinterpret_cast<const ColDesc *>(m + 64);
    int32_t *hdr = reinterpret_cast<int32_t *>(rs);
    uint32_t r[4] = {0, 0, 0, 0};
    int nr = 0;
    if (lane < np) {
      BlockView bv;
      view_from_rec(rec, nullptr, bv);
      const ColDesc &d = plans[lane];
      nr = d.ok ? proj_ranges(d, bv, r) : 0;
      const uint32_t lim = ((rec.size + 15u) & ~15u) + 32u;
      if (nr > 0 && ((r[1] - r[0]) + (nr == 2 ? r[3] - r[2] : 0u) > p.pp_span[lane] || r[1] > lim || (nr == 2 && r[3] > lim))) nr = 0;
      if (nr == 0) r[0] = r[1] = r[2] = r[3] = 0;
      const uint32_t o0 = hdr_bytes + bm_bytes + p.pp_off[lane], o1 = o0 + (r[1] - r[0]);
      hdr[2 * lane] = (int32_t)o0 - (int32_t)r[0];
      hdr[2 * lane + 1] = nr == 2 ? (int32_t)o1 - (int32_t)r[2] : (int32_t)o0 - (int32_t)r[0];
    }
    const uint32_t badmask = __ballot_sync(0xffffffffu, lane < np && nr == 0);
    if (lane == 0) hdr[2 * kMaxProj] = (int32_t)badmask;
    // one bulk copy (TMA) per byte range, all completing on the slot's mbarrier
    const uint32_t len0 = r[1] - r[0], len1 = nr == 2 ? r[3] - r[2] : 0u;
    const uint32_t total = warp_sum_u32(len0 + len1);
    uint64_t *bar = bars + rslot;
    if (lane == 0) mbar_expect_tx(bar, total);
    __syncwarp();
    if (lane < np && nr >= 1) {
      uint8_t *dst0 = rs + hdr_bytes + bm_bytes + p.pp_off[lane];
      const uint8_t *gblk = p.image + rec.off;
      tma_bulk_g2s(dst0, gblk + r[0], len0, bar);
      if (nr == 2) tma_bulk_g2s(dst0 + len0, gblk + r[2], len1, bar);
    }
  };

  issue_meta(blk, 0);
  cp_async_commit();
  cp_async_wait_all();
  __syncwarp();
  issue_regions(blk, 0, 0);
  cp_async_commit();
  issue_meta(blk + nwarps_total, 1);
  cp_async_commit();
  int it = 0;
  for (; blk < p.n_blocks; blk += nwarps_total, ++it) {
    const int ms = it % 3, rsl = it & 1;
    cp_async_wait_all();
    mbar_wait(bars + rsl, (uint32_t)(it >> 1) & 1u);
    __syncwarp();
    issue_regions(blk + nwarps_total, (it + 1) % 3, rsl ^ 1);
    cp_async_commit();
    issue_meta(blk + 2 * nwarps_total, (it + 2) % 3);
    cp_async_commit();

    uint8_t *m = meta0 + (uint32_t)ms * p.pp_meta_bytes;
    uint8_t *rs = reg0 + (uint32_t)rsl * p.pp_region_bytes;
    const BlockRec rec = *reinterpret_cast<const BlockRec *>(m);
    const int64_t base = *reinterpret_cast<const int64_t *>(m + 48);
    const uint32_t cnt = (uint32_t)(*reinterpret_cast<const int64_t *>(m + 56) - base);
    const uint32_t rows = rec.rows;
    if (rows == 0 || cnt == 0 || base + (int64_t)cnt > p.out_cap) {
      if (lane == 0 && rows == 0) atomicOr(p.status, ST_CORRUPT);
      if (lane == 0 && rows != 0 && cnt != 0) atomicOr(p.status, ST_OVERFLOW);
      __syncwarp();
      continue;
    }
    ColDesc *plans = reinterpret_cast<ColDesc *>(m + 64);
    const int32_t *hdr = reinterpret_cast<const int32_t *>(rs);
    const uint32_t badmask = (uint32_t)hdr[2 * kMaxProj];
    const bool all_rows = cnt == rows;
    if (!all_rows) {
      // bitmap words -> ascending selected-row list: lane g owns word g of each group of 32 words
      const uint32_t *bmw = reinterpret_cast<const uint32_t *>(rs + hdr_bytes);
      const uint32_t nwords = (rows + 31u) >> 5;
      uint32_t running = 0;
      for (uint32_t base_w = 0; base_w < nwords; base_w += 32u) {
        const uint32_t w = base_w + (uint32_t)lane;
        const uint32_t word = w < nwords ? bmw[w] : 0u;
        const uint32_t local = __popc(word);
        uint32_t inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += u;
        }
        const uint32_t excl = running + inc - local;
        const uint32_t ng = min(32u, nwords - base_w);
        for (uint32_t g = 0; g < ng; ++g) {   // every lane tests its own bit of word g: coalesced, divergence-free
          const uint32_t wg = __shfl_sync(0xffffffffu, word, g), og = __shfl_sync(0xffffffffu, excl, g);
          if ((wg >> lane) & 1u) sel[og + __popc(wg & ((1u << lane) - 1u))] = (uint16_t)((base_w + g) * 32u + (uint32_t)lane);
        }
        running += __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
    }
    if (p.want_row_ids) {
      int32_t *rid = p.row_ids + base;
      if (all_rows) for (uint32_t j = (uint32_t)lane; j < cnt; j += 32u) rid[j] = (int32_t)j;
      else for (uint32_t j = (uint32_t)lane; j < cnt; j += 32u) rid[j] = (int32_t)sel[j];
    }
    BlockCtx c;
    view_from_rec(rec, nullptr, c.b);
    c.bitsets = nullptr;
    c.descs = plans;
    c.rle_base = wscr + p.pw_rle;
    c.rle_slot_bytes = 0;
    c.rle_starts_bytes = p.words_cap * 4u;
    const uint64_t blk_addr = block_string_addr(p, blk, rec.off);
    for (int pc = 0; pc < np; ++pc) {
      ColDesc *wdesc = plans + pc;
      if (!wdesc->ok || ((badmask >> pc) & 1u)) {
        if (lane == 0) atomicOr(p.status, ST_UNSUPPORTED);
        continue;
      }
      const int32_t d0 = hdr[2 * pc], d1 = hdr[2 * pc + 1];
      if (wdesc->kind == K_DICT && wdesc->sc == 5) {
        const uint32_t idx_sbit = (smem_u32(rs) + (uint32_t)d0) * 8u, ref_sbit = (smem_u32(rs) + (uint32_t)d1) * 8u;
        // which delta belongs to the refs: with two ranges they are ordered by block offset (proj_ranges)
        uint32_t rbit = ref_sbit, ibit = idx_sbit;
        if (d0 == d1) rbit = ibit = idx_sbit;
        else if ((wdesc->val_bit >> 3) < wdesc->dict_payload) { rbit = idx_sbit; ibit = ref_sbit; }
        if (all_rows) project_str_dict_shallow<true>(p, *wdesc, pc, sel, cnt, base, blk_addr, rbit, ibit, lane);
        else project_str_dict_shallow<false>(p, *wdesc, pc, sel, cnt, base, blk_addr, rbit, ibit, lane);
        __syncwarp();
        continue;
      }
      if (wdesc->kind == K_DICT && wdesc->elem_len == 8) {   // fixed-width integer dictionary -> 8-byte datums
        const uint32_t sb = (smem_u32(rs) + (uint32_t)d0) * 8u;
        if (all_rows) lean_project_dict_i64<true>(p, *wdesc, pc, sel, cnt, base, sb, lane);
        else lean_project_dict_i64<false>(p, *wdesc, pc, sel, cnt, base, sb, lane);
        continue;
      }
      c.b.s = rs + d0;
      c.sbit = (smem_u32(rs) + (uint32_t)d0) * 8u;
      project_column_staged(p, c, wdesc, pc, sel, cnt, base, blk_addr, all_rows, rows, wscr, t);
      __syncwarp();
    }
    __syncwarp();
  }
  cp_async_wait_all();
}
