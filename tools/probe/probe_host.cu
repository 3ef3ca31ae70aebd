// host-side error plumbing for the stand-alone probe library (the product library's version lives in capi.cu)
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
namespace sat_host {
int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr);
    return code;
}
int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -2; }
    return 0;
}
}
