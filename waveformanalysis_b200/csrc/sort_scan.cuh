// Device-wide primitives used by the records builder (K1) and the event grouping (K4):
// stable LSD radix sort of 64-bit keys with 64-bit payloads, inclusive scans.
#pragma once
#include "common.cuh"

namespace wfb {

enum KeyKind { kKeyUnsigned = 0, kKeySigned = 1, kKeyFloat64 = 2 };

// bytes of scratch needed by radix_sort_pairs for n elements
size_t radix_sort_workspace_bytes(long long n);

// Stable ascending sort of (key, value) pairs; keys are int64 / uint64 / float64 bit patterns.
// keys_out / vals_out receive the result; the inputs are left untouched.  All device pointers.
int radix_sort_pairs(const unsigned long long* keys_in, const long long* vals_in, unsigned long long* keys_out,
                     long long* vals_out, long long n, KeyKind kind, void* workspace, size_t workspace_bytes,
                     cudaStream_t st);

// scans (n elements, device pointers; `partials` = scratch of scan_workspace_bytes(n))
size_t scan_workspace_bytes(long long n);
int inclusive_scan_sum_i64(const long long* in, long long* out, long long n, void* workspace, cudaStream_t st);
int inclusive_scan_max_f64(const double* in, double* out, long long n, void* workspace, cudaStream_t st);

// running maximum that restarts with every segment (seg ascending); needs 2 * scan_workspace_bytes(n)
struct SegMax {
    double v;
    long long seg;
};
int inclusive_scan_segmax(const SegMax* in, SegMax* out, long long n, void* workspace, cudaStream_t st);

}  // namespace wfb
