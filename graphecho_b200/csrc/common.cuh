// Shared device/host helpers for the graphecho_b200 sm_100a kernels.
// Everything here is header-only; the C-ABI entry points live in the *.cu files
// and are declared in include/graphecho_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>

#define GE_OK 0
#define GE_ERR_ARG (-1)       // null pointer / non-positive dimension
#define GE_ERR_SHAPE (-2)     // shape or alignment the kernel does not support
#define GE_ERR_CAPACITY (-3)  // problem does not fit on-chip for the chosen kernel
#define GE_ERR_DTYPE (-4)

// thread-local last-error text, exported through ge_last_error()
void ge_set_error(const char* fmt, ...);
extern "C" void ge_count_launches(unsigned long long n);

#define GE_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            ge_set_error(__VA_ARGS__);       \
            return (code);                   \
        }                                    \
    } while (0)

// Launch check: positive return = cudaError_t, as the header documents.
#define GE_CHECK_LAUNCH(name)                                                   \
    do {                                                                        \
        cudaError_t e__ = cudaGetLastError();                                   \
        if (e__ != cudaSuccess) {                                               \
            ge_set_error("%s: %s", name, cudaGetErrorString(e__));              \
            return (int)e__;                                                    \
        }                                                                       \
        ge_count_launches(1);                                                   \
    } while (0)

#define GE_CUDA(call, name)                                                     \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) {                                               \
            ge_set_error("%s: %s", name, cudaGetErrorString(e__));              \
            return (int)e__;                                                    \
        }                                                                       \
    } while (0)

namespace ge {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ __forceinline__ int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Block-wide reductions through a caller-provided scratch of >= 32 floats.
// All threads of the block must call; result is broadcast to every thread.
template <typename Op>
__device__ __forceinline__ float block_reduce(float v, float* scratch, Op op, float identity) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(kFull, v, o));
    __syncthreads();  // protect scratch from a previous use
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(kFull, r, o));
    return r;
}
struct OpSum { __device__ __forceinline__ float operator()(float a, float b) const { return a + b; } };
struct OpMax { __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMin { __device__ __forceinline__ float operator()(float a, float b) const { return fminf(a, b); } };

__device__ __forceinline__ float block_sum(float v, float* s) { return block_reduce(v, s, OpSum(), 0.f); }
__device__ __forceinline__ float block_max(float v, float* s) { return block_reduce(v, s, OpMax(), -INFINITY); }
__device__ __forceinline__ float block_min(float v, float* s) { return block_reduce(v, s, OpMin(), INFINITY); }

// ---- activation-dtype helpers (fp32 or bf16 storage, fp32 math) ----
template <typename T> struct Vec4;  // 4 consecutive elements
template <> struct Vec4<float> {
    float4 v;
    __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void get(float (&f)[4]) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    __device__ __forceinline__ void set(const float (&f)[4]) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec4<__nv_bfloat16> {
    uint2 v;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint2*>(p); }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint2*>(p) = v; }
    __device__ __forceinline__ void get(float (&f)[4]) const {
        __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
        __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        f[0] = fa.x; f[1] = fa.y; f[2] = fb.x; f[3] = fb.y;
    }
    __device__ __forceinline__ void set(const float (&f)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
        v.x = *reinterpret_cast<unsigned*>(&a);
        v.y = *reinterpret_cast<unsigned*>(&b);
    }
};

// 8 consecutive elements (one 128-bit access for bf16, two for fp32)
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<const unsigned*>(&t);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// SM count of the current device, cached per process (grids are sized in multiples of it).
int sm_count();

}  // namespace ge
