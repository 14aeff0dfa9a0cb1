// vgb200_check.hpp -- maps C-ABI status codes onto the reference's error convention:
// "[func::time] message" on stderr, then exit (include/cuda_error_handling.hpp:10-16 and the
// 51 `exit(1)` call sites of the reference; no exceptions, no return codes).
#pragma once
#include <cstdlib>
#include <iostream>

#include "get_time.hpp"  // reference header: getTime()
#include "vgb200.h"

#define VGB200_CHECK(call)                                                                           \
    do {                                                                                             \
        int vg_rc__ = (call);                                                                        \
        if (vg_rc__ != VG_OK) {                                                                      \
            std::cerr << "[" << __func__ << "::" << getTime() << "] " << vg_last_error() << std::endl; \
            std::exit(vg_rc__ == VG_E_CUDA ? 2 : 1);                                                 \
        }                                                                                            \
    } while (0)
