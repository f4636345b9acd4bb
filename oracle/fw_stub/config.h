/* TEST INFRASTRUCTURE: stands in for the autoconf-generated config.h that src/lib_common.cpp includes */
#define BUILD_LIBRARIES 1
#define LIB_DIR "/nonexistent"
