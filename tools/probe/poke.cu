// Probe (not part of libsatrans_b200.so): one thread touches *p in a given way -- used to check what peer-mapped (CUDA IPC /
// NVLink) memory supports: mode 0 store, 1 atomicOr, 2 atomicAdd f32, 3 atomicAdd float4, 4 load -> *out.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void k_debug_poke(void* p, int mode, float* out) {
    if (mode == 0) *reinterpret_cast<volatile uint32_t*>(p) = 0x12345678u;
    else if (mode == 1) atomicOr(reinterpret_cast<uint32_t*>(p), 0xF0u);
    else if (mode == 2) atomicAdd(reinterpret_cast<float*>(p), 1.0f);
    else if (mode == 3) atomicAdd(reinterpret_cast<float4*>(p), make_float4(1.f, 2.f, 3.f, 4.f));
    else if (mode == 4) *out = *reinterpret_cast<volatile float*>(p);
}
extern "C" int sat_debug_poke(void* p, int32_t mode, float* out, void* stream) {
    k_debug_poke<<<1, 1, 0, (cudaStream_t)stream>>>(p, mode, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
