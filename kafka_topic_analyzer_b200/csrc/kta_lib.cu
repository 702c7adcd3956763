// kta_lib.cu — single translation unit of libkta_gpu.so (the kernels header defines __global__
// functions, so the host files are compiled together).
#include "kta_api.cu"
#include "kta_synth.cu"
