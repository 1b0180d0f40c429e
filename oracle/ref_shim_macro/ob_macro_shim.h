// Stand-ins that let the REFERENCE's macro-block headers compile unmodified, from where they lie under /root/reference
// (storage/blocksstable/ob_macro_block_common_header.{h,cpp}, ob_sstable_macro_block_header.{h,cpp}), into
// oracle/_ref/libref_macro.so (checker only; oracle/Makefile). On top of ob_codec_shim.h (logging no-ops, error codes with
// the reference's values): the few types whose SIZE enters the header layout, with the reference's layout
// (ObObjMeta: four bytes, common/object/ob_object.h:576-587; ObOrderType: a plain enum, lib/ob_define.h:2458-2462;
// ObTabletID::INVALID_TABLET_ID = 0, common/ob_tablet_id.h:32), and the JSON-printing macros of to_string as no-ops.
#pragma once
#include "ob_codec_shim.h"
#include "lib/utility/ob_unify_serialize.h"   // (shim) OB_SERIALIZE_ERROR / OB_DESERIALIZE_ERROR with the reference's values
#define LOG_WARN(...) ((void)0)
#define LOG_ERROR(...) ((void)0)
#define J_OBJ_START() ((void)0)
#define J_OBJ_END() ((void)0)
#define J_KV(...) ((void)0)
#define J_COMMA() ((void)0)
#define J_NAME(x) ((void)0)
#define J_COLON() ((void)0)
#define J_ARRAY_START() ((void)0)
#define J_ARRAY_END() ((void)0)
#define BUF_PRINTO(x) ((void)0)
#define FALSE_IT(stmt) ({ (stmt); false; })
namespace oceanbase {
namespace common {
const int64_t OB_MAX_MACRO_BLOCK_TYPE = 16;   // lib/ob_define.h:1990
constexpr int OB_INVALID_MACRO_BLOCK_TYPE = -4189;   // share/ob_errno.h:100
class ObString {
public:
  ObString() : p_(nullptr) {}
  explicit ObString(const char *p) : p_(p) {}
  const char *ptr() const { return p_; }
private:
  const char *p_;
};
struct ObObjMeta {
  uint8_t type_, cs_level_, cs_type_;
  int8_t scale_;
};
static_assert(sizeof(ObObjMeta) == 4, "ObObjMeta is four bytes");
enum ObOrderType { ASC = 0, DESC = -1 };
class ObTabletID {
public:
  static const uint64_t INVALID_TABLET_ID = 0;
  explicit ObTabletID(const uint64_t id = INVALID_TABLET_ID) : id_(id) {}
  uint64_t id() const { return id_; }
private:
  uint64_t id_;
};
template <typename T>
class ObIArray {
public:
  ObIArray() : d_(nullptr), n_(0) {}
  ObIArray(const T *d, int64_t n) : d_(d), n_(n) {}
  int64_t count() const { return n_; }
  const T &at(int64_t i) const { return d_[i]; }
private:
  const T *d_;
  int64_t n_;
};
}  // namespace common
namespace blocksstable { using namespace common; }   // the reference's headers pull common:: in the same way
namespace share {
namespace schema {
struct ObColDesc {
  common::ObObjMeta col_type_;
  common::ObOrderType col_order_;
};
}  // namespace schema
}  // namespace share
}  // namespace oceanbase
