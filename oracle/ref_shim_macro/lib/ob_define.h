#pragma once
#include "ob_macro_shim.h"
