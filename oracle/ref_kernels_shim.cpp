/* TEST INFRASTRUCTURE - the one global the unmodified src/cuda_wrapper/kernels.cu needs from the rest of UltraGrid (src/debug.cpp). */
volatile int log_level = 0;
