// Disk-format bytes straight into the block cache: a run of fixed-size macro blocks (ObMacroBlock, ob_macro_block.cpp:455-520) is
// parsed ON THE DEVICE and re-laid as a page batch. Inside a macro block the micro-blocks lie back to back at arbitrary byte
// offsets; the scan kernels stage blocks with bulk copies (TMA), which need 16-byte aligned sources, so the open path is
//   survey  : one thread per macro block -- ObMacroBlockCommonHeader::check_integrity, FixedHeader::is_valid
//             (ob_macro_block_common_header.cpp:54-69, ob_sstable_macro_block_header.cpp:118-140) -> micro_block_count_
//   walk    : one thread per macro block follows the micro headers (header_size_ + data_zlength_ each) from
//             micro_block_data_offset_; the walk must end at micro_block_data_offset_ + micro_block_data_size_ with row_count_ rows
//   realign : one CTA per micro-block copies it to a 128-byte aligned slot of a new image (unaligned source words through
//             funnel shifts, 16-byte stores, zero padding) -- one read + one write of the data, at HBM speed, once per cache fill
// 16 bytes per micro-block (offset, size) and 4 per macro block come back to the host for obgpu_batch_open's tables; the block
// bytes never touch the CPU. Compressed payloads (compressor_type_ != NONE) and encrypted blocks are refused.
#pragma once

namespace mb {

constexpr int32_t kStOk = 0, kStBadCommon = 1, kStBadFixed = 2, kStBadWalk = 3, kStCompressed = 4;

__device__ __forceinline__ uint32_t ld32u(const uint8_t *p) {   // unaligned little-endian loads
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint64_t ld64u(const uint8_t *p) { return (uint64_t)ld32u(p) | ((uint64_t)ld32u(p + 4) << 32); }

struct Fixed {
  int32_t micro_count, data_off, data_size, row_count;
};

__device__ __forceinline__ int32_t parse_headers(const uint8_t *m, int64_t macro_size, Fixed &f) {
  // ObMacroBlockCommonHeader: header_size_, version_, magic_, attr_, payload_size_, payload_checksum_ (aligned: macro blocks start
  // at multiples of the macro block size)
  const int32_t *c = reinterpret_cast<const int32_t *>(m);
  if (c[0] != 24 || c[1] != 1 || c[2] != 1001 || c[3] != 1 /*SSTableData*/) return kStBadCommon;
  if (c[4] <= 0 || 24 + (int64_t)c[4] > macro_size) return kStBadCommon;
  const uint8_t *h = m + 24;
  const uint32_t version = (uint32_t)h[4] | ((uint32_t)h[5] << 8), magic = (uint32_t)h[6] | ((uint32_t)h[7] << 8);
  const uint64_t tablet = ld64u(h + 8);
  const int64_t logical = (int64_t)ld64u(h + 16);
  const int32_t column_count = (int32_t)ld32u(h + 32), rowkey_cnt = (int32_t)ld32u(h + 36), row_store_type = (int32_t)ld32u(h + 40);
  f.row_count = (int32_t)ld32u(h + 44);
  const int32_t occupy = (int32_t)ld32u(h + 48);
  f.micro_count = (int32_t)ld32u(h + 52);
  f.data_off = (int32_t)ld32u(h + 56);
  f.data_size = (int32_t)ld32u(h + 60);
  const int64_t data_checksum = (int64_t)ld64u(h + 80), encrypt_id = (int64_t)ld64u(h + 88), master_key = (int64_t)ld64u(h + 96);
  const uint32_t compressor = h[104];
  if (!(version >= 1 && version <= 2 && magic == 1007 && tablet != 0 && logical >= 0 && rowkey_cnt >= 0 && row_store_type >= 0 &&
        f.row_count > 0 && occupy > 0 && f.micro_count > 0 && f.data_off > 0 && f.data_size > 0 && data_checksum >= 0 && encrypt_id >= 0 &&
        master_key >= -1 && compressor > 0))
    return kStBadFixed;
  const int64_t type_cols = version == 2 ? rowkey_cnt : column_count;
  if (f.data_off != 24 + 128 + type_cols * 8 + (int64_t)column_count * 8 + 1) return kStBadFixed;
  if ((int64_t)f.data_off + f.data_size > macro_size || occupy != f.data_off + f.data_size) return kStBadFixed;
  if ((int64_t)f.micro_count * 64 > f.data_size) return kStBadFixed;   // a micro-block is at least its 64-byte header
  if (compressor != 1 /*NONE_COMPRESSOR*/ || encrypt_id != 0) return kStCompressed;
  return kStOk;
}

__global__ void obgpu_macro_survey_kernel(const uint8_t *image, int64_t macro_size, int32_t n_macro, int32_t *counts, int32_t *status) {
  const int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n_macro) return;
  Fixed f;
  const int32_t st = parse_headers(image + (int64_t)i * macro_size, macro_size, f);
  counts[i] = st == kStOk ? f.micro_count : 0;
  if (st != kStOk) atomicMax(status, st);
}

// first[i]: index of macro block i's first micro-block in the output tables
__global__ void obgpu_macro_walk_kernel(const uint8_t *image, int64_t macro_size, int32_t n_macro, const int64_t *first, int64_t *src_off,
                                        int64_t *sizes, int32_t *status) {
  const int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n_macro) return;
  const uint8_t *m = image + (int64_t)i * macro_size;
  Fixed f;
  if (parse_headers(m, macro_size, f) != kStOk) return;
  int64_t at = f.data_off, rows = 0;
  const int64_t end = (int64_t)f.data_off + f.data_size;
  bool ok = true;
  for (int32_t k = 0; k < f.micro_count; ++k) {
    if (at + 64 > end) { ok = false; break; }
    const uint8_t *h = m + at;
    const uint32_t magic = (uint32_t)h[0] | ((uint32_t)h[1] << 8);
    const int64_t sz = (int64_t)ld32u(h + 4) + (int32_t)ld32u(h + 44);   // header_size_ + data_zlength_
    if (magic != 1005u || sz < 64 || at + sz > end) { ok = false; break; }
    src_off[first[i] + k] = (int64_t)i * macro_size + at;
    sizes[first[i] + k] = sz;
    rows += ld32u(h + 16);
    at += sz;
  }
  if (!ok || at != end || rows != f.row_count) atomicMax(status, kStBadWalk);
}

constexpr int kCopyThreads = 128;
__global__ void __launch_bounds__(kCopyThreads) obgpu_macro_realign_kernel(const uint8_t *image, int64_t image_size, const int64_t *src_off,
                                                                            const int64_t *sizes, const int64_t *dst_off, uint8_t *out) {
  const int64_t blk = blockIdx.x;
  const int64_t src = src_off[blk], sz = sizes[blk];
  const int64_t slot = (sz + 127) & ~127ll;
  uint4 *dst = reinterpret_cast<uint4 *>(out + dst_off[blk]);
  const uint32_t sh = (uint32_t)(src & 3) * 8u;
  const uint32_t *w = reinterpret_cast<const uint32_t *>(image + (src & ~3ll));
  const int64_t w_cap = (image_size - (src & ~3ll)) >> 2;   // whole words readable from w
  for (int64_t j = threadIdx.x; j < slot / 16; j += kCopyThreads) {
    uint32_t v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int64_t idx = j * 4 + k;
      v[k] = idx < w_cap ? __ldg(w + idx) : 0u;
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = __funnelshift_r(v[k], v[k + 1], sh);
    const int64_t left = sz - j * 16;   // bytes of this chunk that belong to the block
    if (left < 16) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t lb = left - 4 * k;
        o[k] = lb >= 4 ? o[k] : (lb <= 0 ? 0u : (o[k] & (0xffffffffu >> (32 - 8 * (int)lb))));
      }
    }
    dst[j] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace mb

extern "C" {

int obgpu_batch_open_macro_blocks(obgpu_ctx *ctx, const void *macro_image, int64_t image_size, int64_t macro_block_size, int32_t n_macro_blocks,
                                  int32_t image_on_device, obgpu_batch **out, int32_t *n_micro_out) {
  if (!ctx || !macro_image || !out || n_macro_blocks <= 0 || macro_block_size < 4096 || (macro_block_size & 15) != 0 ||
      image_size < macro_block_size * (int64_t)n_macro_blocks)
    return OBGPU_INVALID_ARGUMENT;
  cudaSetDevice(ctx->device);
  const uint8_t *d_macro = (const uint8_t *)macro_image;
  void *tmp_image = nullptr;
  if (!image_on_device) {
    CUDA_TRY(ctx, cudaMallocAsync(&tmp_image, (size_t)image_size, ctx->stream));
    cudaError_t e = cudaMemcpyAsync(tmp_image, macro_image, (size_t)image_size, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); cudaFreeAsync(tmp_image, ctx->stream); return OBGPU_ERR_SYS; }
    d_macro = (const uint8_t *)tmp_image;
  } else if (((uintptr_t)macro_image & 15u) != 0) {
    ctx->err = "device-resident macro blocks must be 16-byte aligned";
    return OBGPU_INVALID_ARGUMENT;
  }
  int ret = OBGPU_SUCCESS;
  void *d_small = nullptr, *d_tab = nullptr, *d_out = nullptr;
  std::vector<int32_t> counts((size_t)n_macro_blocks + 1);
  std::vector<int64_t> first((size_t)n_macro_blocks + 1), src, sizes, dst;
  int64_t n_micro = 0, out_bytes = 0;
  auto fail = [&](int code, const char *what) { if (what) ctx->err = what; ret = code; };
  do {
    if (cudaMallocAsync(&d_small, ((size_t)n_macro_blocks + 1) * 12 + 64, ctx->stream) != cudaSuccess) { fail(OBGPU_ALLOCATE_MEMORY_FAILED, "macro survey tables"); break; }
    int32_t *d_counts = (int32_t *)d_small, *d_status = d_counts + n_macro_blocks;
    int64_t *d_first = (int64_t *)((uint8_t *)d_small + (((size_t)n_macro_blocks + 1) * 4 + 15) / 16 * 16);
    cudaMemsetAsync(d_status, 0, 4, ctx->stream);
    mb::obgpu_macro_survey_kernel<<<(unsigned)((n_macro_blocks + 127) / 128), 128, 0, ctx->stream>>>(d_macro, macro_block_size, n_macro_blocks, d_counts, d_status);
    ctx->launches++;
    if (cudaMemcpyAsync(counts.data(), d_counts, ((size_t)n_macro_blocks + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { fail(OBGPU_ERR_SYS, "macro survey"); break; }
    if (counts[(size_t)n_macro_blocks] != mb::kStOk) {
      fail(counts[(size_t)n_macro_blocks] == mb::kStCompressed ? OBGPU_NOT_SUPPORTED : OBGPU_INVALID_DATA,
           counts[(size_t)n_macro_blocks] == mb::kStCompressed ? "compressed or encrypted macro block" : "macro block header is invalid");
      break;
    }
    for (int32_t i = 0; i < n_macro_blocks; ++i) { first[(size_t)i] = n_micro; n_micro += counts[(size_t)i]; }
    first[(size_t)n_macro_blocks] = n_micro;
    if (n_micro <= 0 || n_micro > 0x7fffffff) { fail(OBGPU_INVALID_DATA, "macro blocks hold no micro-block"); break; }
    if (cudaMallocAsync(&d_tab, (size_t)n_micro * 24 + 64, ctx->stream) != cudaSuccess) { fail(OBGPU_ALLOCATE_MEMORY_FAILED, "micro-block tables"); break; }
    int64_t *d_src = (int64_t *)d_tab, *d_sizes = d_src + n_micro, *d_dst = d_sizes + n_micro;
    cudaMemcpyAsync(d_first, first.data(), ((size_t)n_macro_blocks + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
    mb::obgpu_macro_walk_kernel<<<(unsigned)((n_macro_blocks + 63) / 64), 64, 0, ctx->stream>>>(d_macro, macro_block_size, n_macro_blocks, d_first, d_src, d_sizes, d_status);
    ctx->launches++;
    src.resize((size_t)n_micro); sizes.resize((size_t)n_micro); dst.resize((size_t)n_micro);
    int32_t st = 0;
    if (cudaMemcpyAsync(sizes.data(), d_sizes, (size_t)n_micro * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { fail(OBGPU_ERR_SYS, "macro walk"); break; }
    if (st != mb::kStOk) { fail(OBGPU_INVALID_DATA, "micro-block chain of a macro block is inconsistent"); break; }
    for (int64_t k = 0; k < n_micro; ++k) { dst[(size_t)k] = out_bytes; out_bytes += (sizes[(size_t)k] + 127) & ~127ll; }
    if (cudaMallocAsync(&d_out, (size_t)out_bytes + 64, ctx->stream) != cudaSuccess) { fail(OBGPU_ALLOCATE_MEMORY_FAILED, "realigned image"); break; }
    cudaMemsetAsync((uint8_t *)d_out + out_bytes, 0, 64, ctx->stream);
    cudaMemcpyAsync(d_dst, dst.data(), (size_t)n_micro * 8, cudaMemcpyHostToDevice, ctx->stream);
    mb::obgpu_macro_realign_kernel<<<(unsigned)n_micro, mb::kCopyThreads, 0, ctx->stream>>>(d_macro, image_size, d_src, d_sizes, d_dst, (uint8_t *)d_out);
    ctx->launches++;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { fail(OBGPU_ERR_SYS, "macro realign"); break; }   // dst (host vector) was the copy source
    obgpu_batch *b = nullptr;
    ret = obgpu_batch_open(ctx, d_out, out_bytes, dst.data(), sizes.data(), (int32_t)n_micro, 1, nullptr, &b);
    if (ret != OBGPU_SUCCESS) break;
    b->own_image = true;   // the realigned image lives and dies with the batch
    d_out = nullptr;
    *out = b;
    if (n_micro_out) *n_micro_out = (int32_t)n_micro;
  } while (0);
  if (d_small) cudaFreeAsync(d_small, ctx->stream);
  if (d_tab) cudaFreeAsync(d_tab, ctx->stream);
  if (d_out) cudaFreeAsync(d_out, ctx->stream);
  if (tmp_image) cudaFreeAsync(tmp_image, ctx->stream);
  return ret;
}

}  // extern "C"
