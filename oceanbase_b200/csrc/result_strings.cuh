// Bytes of projected string cells as a dense heap: obgpu_result_fetch_strings (the selected rows of a scan) and
// obgpu_project_strings (one block, a row list). Needed for the columns whose values exist only on the device -- HEX_PACKING /
// STRING_DIFF / STRING_PREFIX, rebuilt at batch open (mat_codecs.cuh) -- and usable for every string column (the caller then gets
// bytes instead of pointers into its own image). Same pattern as the merge's string materialisation: lengths -> exclusive scan ->
// one warp per row copies the cell.
#pragma once

namespace resstr {

// where the cell of every row starts inside the batch's device image: the block of a dense output row by binary search in the
// per-block prefix, the cell offset from the pointer the projection reported (string_base + the block's place in the caller's image)
__global__ void __launch_bounds__(256) src_off_kernel(const uint64_t *__restrict__ ptrs, const int32_t *__restrict__ lens, int64_t row_begin, int64_t n,
                                                      const int64_t *__restrict__ sel_offset, int n_blocks, const BlockRec *__restrict__ recs,
                                                      const obcs::XformRec *__restrict__ xf, uint64_t string_base, uint64_t *__restrict__ src_off,
                                                      int *__restrict__ status) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t row = row_begin + k;
  src_off[k] = 0;
  if (lens[row] <= 0) return;
  int lo = 0, hi = n_blocks;   // last block whose first output row is <= row
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (sel_offset[mid] <= row) lo = mid; else hi = mid;
  }
  const BlockRec rec = recs[lo];
  const uint64_t base = string_base + (xf ? xf[lo].orig_off + (uint64_t)xf[lo].str_delta : rec.off);
  const uint64_t cell = ptrs[row] - base;
  if (cell + (uint64_t)lens[row] > (uint64_t)rec.size) { atomicOr(status, ST_CORRUPT); return; }
  src_off[k] = rec.off + cell;
}

__global__ void __launch_bounds__(256) gather_kernel(const uint8_t *__restrict__ image, const uint64_t *__restrict__ src_off,
                                                     const int32_t *__restrict__ lens, int64_t n, const int64_t *__restrict__ off,
                                                     uint8_t *__restrict__ heap) {
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= n) return;
  const int32_t len = lens[k];
  if (len <= 0) return;
  const uint8_t *src = image + src_off[k];
  uint8_t *dst = heap + off[k];
  for (int32_t i = lane; i < len; i += 32) dst[i] = src[i];
}

}  // namespace resstr

namespace {

// lens (device, int32, <= 0 for NULL rows) + src offsets (device) -> host offsets [n + 1] and heap
int gather_to_host(obgpu_ctx *ctx, const uint8_t *d_image, const uint64_t *d_src_off, const int32_t *d_lens, int64_t n, void *host_heap,
                   int64_t heap_cap, int64_t *host_off, int64_t *heap_bytes, uint8_t *scratch, size_t o_off, size_t o_chunk) {
  const int n_chunks = (int)((n + kPrefixChunk - 1) / kPrefixChunk);
  int64_t *d_off = (int64_t *)(scratch + o_off);
  obgpu_prefix_local_kernel<<<n_chunks, 256, 0, ctx->stream>>>((const uint32_t *)d_lens, (int)n, d_off, (unsigned long long *)(scratch + o_chunk));
  obgpu_prefix_fix_kernel<<<n_chunks + 1, 256, 0, ctx->stream>>>((int)n, n_chunks, d_off, (const unsigned long long *)(scratch + o_chunk));
  ctx->launches += 2;
  CUDA_TRY(ctx, cudaMemcpyAsync(host_off, d_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t total = host_off[n];
  *heap_bytes = total;
  if (total > heap_cap || (total > 0 && !host_heap)) return OBGPU_BUF_NOT_ENOUGH;
  if (total == 0) return OBGPU_SUCCESS;
  uint8_t *d_heap = nullptr;
  CUDA_TRY(ctx, cudaMallocAsync((void **)&d_heap, (size_t)total + 16, ctx->stream));
  resstr::gather_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_image, d_src_off, d_lens, n, d_off, d_heap);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_heap, d_heap, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFreeAsync(d_heap, ctx->stream);
  if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return OBGPU_ERR_SYS; }
  return OBGPU_SUCCESS;
}

}  // namespace

extern "C" {

int obgpu_batch_column_materialised(const obgpu_batch *b, int32_t col, int32_t *materialised) {
  if (!b || !materialised || col < 0 || (size_t)col >= b->col_mat.size()) return OBGPU_INVALID_ARGUMENT;
  *materialised = b->col_mat[(size_t)col];
  return OBGPU_SUCCESS;
}

int obgpu_result_fetch_strings(obgpu_result *r, int32_t i, int64_t row_begin, int64_t row_count, void *host_heap, int64_t heap_cap,
                               int64_t *host_off, int64_t *heap_bytes) {
  if (!r || i < 0 || i >= r->n_proj || row_begin < 0 || row_count < 0 || row_begin + row_count > r->cap || !host_off || !heap_bytes)
    return OBGPU_INVALID_ARGUMENT;
  const ResultCol &c = r->cols[i];
  obgpu_ctx *ctx = r->ctx;
  if (!c.is_string) { ctx->err = "not a string column"; return OBGPU_INVALID_ARGUMENT; }
  {
    obgpu_result_info info;
    const int ret = obgpu_result_info_get(r, &info);
    if (ret != OBGPU_SUCCESS) return ret;
    if (row_begin + row_count > info.selected_rows) return OBGPU_INVALID_ARGUMENT;
  }
  cudaSetDevice(ctx->device);
  host_off[0] = 0;
  *heap_bytes = 0;
  if (row_count == 0) return OBGPU_SUCCESS;
  obgpu_batch *b = r->batch;
  const int64_t n = row_count;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_src = 256, o_off = o_src + al((size_t)n * 8), o_chunk = o_off + al(((size_t)n + 1) * 8);
  const size_t total = o_chunk + al(((size_t)(n / kPrefixChunk) + 3) * 8);
  TempDev tmp(ctx);
  CUDA_TRY(ctx, tmp.alloc(total));
  uint8_t *t = (uint8_t *)tmp.p;
  CUDA_TRY(ctx, cudaMemsetAsync(t, 0, 256, ctx->stream));
  resstr::src_off_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      (const uint64_t *)c.data, c.lens, row_begin, n, r->d_sel_offset, b->n_blocks, b->d_recs, b->d_xf, r->string_base, (uint64_t *)(t + o_src), (int *)t);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  int ret = gather_to_host(ctx, b->d_image, (const uint64_t *)(t + o_src), c.lens + row_begin, n, host_heap, heap_cap, host_off, heap_bytes, t, o_off,
                           o_chunk);
  if (ret != OBGPU_SUCCESS) return ret;
  int st = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&st, t, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return check_status(ctx, st);
}

int obgpu_project_strings(obgpu_batch *b, int32_t block, int32_t col, const int32_t *row_ids, int64_t row_cap, void *host_heap, int64_t heap_cap,
                          int64_t *host_off, uint64_t *host_nulls, int32_t *has_null, int64_t *heap_bytes) {
  if (!b || !row_ids || row_cap < 0 || !host_off || !heap_bytes || block < 0 || block >= b->n_blocks) return OBGPU_INVALID_ARGUMENT;
  obgpu_ctx *ctx = b->ctx;
  host_off[0] = 0;
  *heap_bytes = 0;
  if (row_cap == 0) return OBGPU_SUCCESS;
  // the ordinary discrete projection with string_base 0 reports every cell at (the block's place in the caller's image) + cell
  std::vector<uint64_t> ptrs((size_t)row_cap, 0), nulls((size_t)(row_cap + 63) / 64, 0);
  std::vector<int32_t> lens((size_t)row_cap, 0);
  int32_t hn = 0;
  int ret = obgpu_project_discrete(b, block, col, row_ids, row_cap, 0, 0, ptrs.data(), lens.data(), nulls.data(), &hn);
  if (ret != OBGPU_SUCCESS) return ret;
  if (host_nulls) memcpy(host_nulls, nulls.data(), nulls.size() * 8);
  if (has_null) *has_null = hn;
  uint64_t base = (uint64_t)b->offsets[(size_t)block];
  if (b->d_xf) {
    obcs::XformRec x;
    CUDA_TRY(ctx, cudaMemcpyAsync(&x, b->d_xf + block, sizeof(x), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    base = x.orig_off + (uint64_t)x.str_delta;
  }
  for (int64_t k = 0; k < row_cap; ++k) {
    const bool is_null = (nulls[(size_t)k / 64] >> (k % 64)) & 1ull;
    if (is_null) { lens[(size_t)k] = 0; ptrs[(size_t)k] = 0; continue; }
    const uint64_t cell = ptrs[(size_t)k] - base;
    if (cell + (uint64_t)lens[(size_t)k] > (uint64_t)b->sizes[(size_t)block]) { ctx->err = "string cell outside its micro block"; return OBGPU_INVALID_DATA; }
    ptrs[(size_t)k] = (uint64_t)b->offsets[(size_t)block] + cell;   // offset inside the batch's device image
  }
  const int64_t n = row_cap;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_src = 0, o_len = al((size_t)n * 8), o_off = o_len + al((size_t)n * 4), o_chunk = o_off + al(((size_t)n + 1) * 8);
  const size_t total = o_chunk + al(((size_t)(n / kPrefixChunk) + 3) * 8);
  TempDev tmp(ctx);
  CUDA_TRY(ctx, tmp.alloc(total));
  uint8_t *t = (uint8_t *)tmp.p;
  CUDA_TRY(ctx, cudaMemcpyAsync(t + o_src, ptrs.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(t + o_len, lens.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  return gather_to_host(ctx, b->d_image, (const uint64_t *)(t + o_src), (const int32_t *)(t + o_len), n, host_heap, heap_cap, host_off, heap_bytes, t,
                        o_off, o_chunk);
}

}  // extern "C"
