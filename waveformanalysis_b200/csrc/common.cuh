// Shared device/host helpers for libwfb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/wfb200.h"

namespace wfb {

// ---- error plumbing (thread-local last error, C-ABI status codes) ---------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define WFB_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return wfb::cuda_fail(e__, #call); \
    } while (0)

#define WFB_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            wfb::set_error(__VA_ARGS__);  \
            return WFB_ERR_INVALID;       \
        }                                 \
    } while (0)

int sm_count();

constexpr int kRecordsRowBytes = 102;  // RECORDS_DTYPE
constexpr int kFeatRowBytes = 36;      // BASIC_FEATURES_DTYPE
constexpr int kHitRowBytes = 60;       // THRESHOLD_HIT_DTYPE
constexpr int kWidthRowBytes = 56;     // WAVEFORM_WIDTH_DTYPE
constexpr int kWidthIntRowBytes = 52;  // WAVEFORM_WIDTH_INTEGRAL_DTYPE

constexpr unsigned kFull = 0xffffffffu;

// ---- warp helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(kFull, lo, m);
    hi = __shfl_xor_sync(kFull, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(kFull, lo, src);
    hi = __shfl_sync(kFull, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_f64(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(kFull, lo, d);
    hi = __shfl_up_sync(kFull, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}
__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor_f64(v, m));
    return v;
}
__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor_f64(v, m));
    return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, m));
    return v;
}
__device__ __forceinline__ float warp_min_f32(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, m));
    return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        int lo = __shfl_xor_sync(kFull, (int)(v & 0xffffffffll), m);
        int hi = __shfl_xor_sync(kFull, (int)(v >> 32), m);
        v += ((long long)hi << 32) | (unsigned)lo;
    }
    return v;
}

// python seq[start:end] bounds for a sequence of length len (step 1)
__device__ __forceinline__ void resolve_slice(long long start, long long end, int len, int& lo, int& hi) {
    long long s = start < 0 ? (start + len < 0 ? 0 : start + len) : (start > len ? len : start);
    long long e = end < 0 ? (end + len < 0 ? 0 : end + len) : (end > len ? len : end);
    lo = (int)s;
    hi = (int)(e > s ? e : s);
}

// streaming (read-once) 16-byte load: keep it out of L1
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace wfb
