#pragma once
#include "ob_macro_shim.h"
namespace oceanbase { namespace share {
const int64_t OB_MAX_TABLESPACE_ENCRYPT_KEY_LENGTH = 16;   // share/ob_encryption_util.h:121
} }
