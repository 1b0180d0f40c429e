#include "ob_codec_shim.h"
