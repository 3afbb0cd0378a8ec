// Stand-in for csrc/ga_common.cuh when the host-only sources of the library (csrc/ga_traverse.cu, csrc/ga_parse.cu) are
// built as plain C++ with a sanitizer (tests/test_sanitizers.py): the C ABI header and the error hook, nothing of CUDA.
#pragma once
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "ga_b200.h"

static inline void ga_set_error(const char* fmt, ...) { (void)fmt; }
