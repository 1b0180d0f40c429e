// extern "C" doors onto the REFERENCE's common::ObBitmap (deps/oblib/src/lib/container/ob_bitmap.{h,cpp}, compiled from
// /root/reference, not copied): the byte-per-row selection vector of the scan path. tests/test_bitmap_kat.py pins the
// oracle's restatement (ora_bitmap_get_row_ids, and / or / not / popcnt) and the device's obgpu_bitmap_to_row_ids to it.
#include "lib/container/ob_bitmap.h"
#include "common/ob_target_specific.h"

using namespace oceanbase::common;

namespace oceanbase { namespace common {
uint32_t arches = 0;
void init_arches() {
  arches = 0;
  if (__builtin_cpu_supports("sse4.2")) arches |= ObTargetArch::SSE42;
  if (__builtin_cpu_supports("avx")) arches |= ObTargetArch::AVX;
  if (__builtin_cpu_supports("avx2")) arches |= ObTargetArch::AVX2;
  if (__builtin_cpu_supports("avx512bw")) arches |= ObTargetArch::AVX512;
}
} }

namespace {
struct MallocAllocator : public ObIAllocator {
  void *alloc(const int64_t size) override { return malloc((size_t)size); }
  void free(void *p) override { ::free(p); }
};
struct Holder {
  MallocAllocator alloc;
  ObBitmap bm;
  Holder() : bm(alloc) {}
};
}  // namespace

extern "C" {
void *ref_bitmap_create(const uint8_t *bytes, int64_t n) {
  static const bool inited = (init_arches(), true);
  (void)inited;
  Holder *h = new Holder();
  if (h->bm.init((uint64_t)n, false) != 0) { delete h; return nullptr; }
  for (int64_t i = 0; i < n; ++i) if (bytes[i]) h->bm.set((uint64_t)i, true);
  return h;
}
void ref_bitmap_destroy(void *p) { delete (Holder *)p; }
int ref_bitmap_get_row_ids(void *p, int32_t *row_ids, int64_t *row_count, int64_t *from, int64_t to, int64_t limit, int64_t id_offset) {
  return ((Holder *)p)->bm.get_row_ids(row_ids, *row_count, *from, to, limit, id_offset);
}
int ref_bitmap_and(void *a, void *b) { return ((Holder *)a)->bm.bit_and(((Holder *)b)->bm); }
int ref_bitmap_or(void *a, void *b) { return ((Holder *)a)->bm.bit_or(((Holder *)b)->bm); }
int ref_bitmap_not(void *a) { return ((Holder *)a)->bm.bit_not(); }
uint64_t ref_bitmap_popcnt(void *a) { return ((Holder *)a)->bm.popcnt(); }
int ref_bitmap_all_false(void *a) { return ((Holder *)a)->bm.is_all_false(); }
int ref_bitmap_all_true(void *a) { return ((Holder *)a)->bm.is_all_true(); }
int64_t ref_bitmap_next_valid_idx(void *a, int64_t start, int64_t count, int is_reverse) {
  int64_t off = -1;
  ((Holder *)a)->bm.next_valid_idx(start, count, is_reverse != 0, off);
  return off;
}
void ref_bitmap_bytes(void *a, uint8_t *out, int64_t n) { memcpy(out, ((Holder *)a)->bm.get_data(), (size_t)n); }
int ref_bitmap_to_bits_mask(void *a, int64_t from, int64_t to, int need_flip, uint8_t *bits) {
  return ((Holder *)a)->bm.to_bits_mask(from, to, need_flip != 0, bits);
}
}
