/* Stand-in for the CMake-generated sampleConfig.h (src/sampleConfig.h.in) so the
 * reference's own sources can be compiled in place from /root/reference by
 * oracle/Makefile without running the reference's build system.  Glue only:
 * no reference code lives here.  TEST INFRASTRUCTURE — never shipped. */
#pragma once
#define SAMPLES_DIR ""
#define SAMPLES_PTX_DIR ""
#define SAMPLES_CUDA_DIR ""
#define SAMPLES_RELATIVE_INCLUDE_DIRS
#define SAMPLES_ABSOLUTE_INCLUDE_DIRS
#define CUDA_NVRTC_ENABLED 0
#define CUDA_NVRTC_OPTIONS
