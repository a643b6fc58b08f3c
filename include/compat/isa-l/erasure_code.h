/* Stand-in for <isa-l/erasure_code.h>: the reference includes that header when
 * LIZARDFS_HAVE_ISA_L_ERASURE_CODE_H is set (src/common/reed_solomon.h:27-28).  With this directory on the include
 * path the same five names resolve to liblzgpu.so. */
#pragma once
#include "../../lzgpu.h"
