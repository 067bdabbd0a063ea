// Stable LSD radix sort (8 bits per pass, 64-bit keys + 64-bit payloads) and inclusive scans.
// Replaces the numpy lexsort / argsort calls of the reference's time ordering
// (core/processing/records_builder.py:115-120, core/processing/event_grouping.py:142, 418).
#include <algorithm>

#include "sort_scan.cuh"

namespace wfb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;  // keys per thread, processed warp-striped: 32 consecutive keys per round
constexpr int kSortTile = kSortThreads * kSortItems;

__device__ __forceinline__ unsigned long long sortable(unsigned long long k, int kind) {
    if (kind == kKeySigned) return k ^ 0x8000000000000000ull;
    if (kind == kKeyFloat64) return (k >> 63) ? ~k : (k ^ 0x8000000000000000ull);
    return k;
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const unsigned long long* __restrict__ keys, long long n,
                                                                  int shift, int kind, unsigned* __restrict__ block_hist,
                                                                  int nblocks) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        long long i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(sortable(keys[i], kind) >> shift) & 0xff], 1u);
    }
    __syncthreads();
    block_hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `total` unsigned counters in place: per-tile sums, one block over the (<= 1024)
// tile sums, then each tile scans itself from its offset; all loads and stores are coalesced
__device__ __forceinline__ unsigned block_exclusive_scan_1024(unsigned v, unsigned* warp_sums /*[32]*/, unsigned& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = warp_sums[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_up_sync(kFull, wi, d);
            if (lane >= d) wi += t;
        }
        warp_sums[lane] = wi - w;  // exclusive prefix of the warp
        if (lane == 31) warp_sums[32] = wi;
    }
    __syncthreads();
    const unsigned excl = warp_sums[warp] + incl - v;
    block_total = warp_sums[32];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(1024) radix_tile_sums_kernel(const unsigned* __restrict__ v, long long total, long long tile,
                                                              unsigned* __restrict__ sums) {
    __shared__ unsigned ws[33];
    const long long lo = (long long)blockIdx.x * tile, hi = min(total, lo + tile);
    unsigned s = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += 1024) s += v[i];
    unsigned tot;
    block_exclusive_scan_1024(s, ws, tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) radix_scan_sums_kernel(unsigned* __restrict__ sums, int ntiles) {
    __shared__ unsigned ws[33];
    unsigned tot;
    const unsigned v = (int)threadIdx.x < ntiles ? sums[threadIdx.x] : 0u;
    const unsigned e = block_exclusive_scan_1024(v, ws, tot);
    if ((int)threadIdx.x < ntiles) sums[threadIdx.x] = e;
}

__global__ void __launch_bounds__(1024) radix_scan_tiles_kernel(unsigned* __restrict__ v, long long total, long long tile,
                                                               const unsigned* __restrict__ sums) {
    __shared__ unsigned ws[33];
    const long long lo = (long long)blockIdx.x * tile, hi = min(total, lo + tile);
    unsigned run = sums[blockIdx.x];
    for (long long base = lo; base < hi; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned c = i < hi ? v[i] : 0u;
        unsigned tot;
        const unsigned e = block_exclusive_scan_1024(c, ws, tot);
        if (i < hi) v[i] = run + e;
        run += tot;
    }
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const unsigned long long* __restrict__ keys_in,
                                                                     const long long* __restrict__ vals_in,
                                                                     unsigned long long* __restrict__ keys_out,
                                                                     long long* __restrict__ vals_out, long long n, int shift,
                                                                     int kind, const unsigned* __restrict__ block_hist,
                                                                     int nblocks) {
    __shared__ unsigned wcount[kSortThreads / 32][256];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const long long wbase = (long long)blockIdx.x * kSortTile + (long long)warp * (32 * kSortItems);
    unsigned long long key[kSortItems];
    unsigned dig[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        long long i = wbase + r * 32 + lane;
        bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0;
        dig[r] = valid ? (unsigned)((sortable(key[r], kind) >> shift) & 0xff) : 0x100u;
        unsigned peers = __match_any_sync(kFull, dig[r]);
        if (valid && lane == __ffs(peers) - 1) wcount[warp][dig[r]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive offsets: global start of this block's digit run + counts of the warps before
        const int d = threadIdx.x;
        unsigned run = block_hist[(size_t)d * nblocks + blockIdx.x];
        for (int w = 0; w < kSortThreads / 32; ++w) {
            unsigned c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        long long i = wbase + r * 32 + lane;
        bool valid = i < n;
        unsigned peers = __match_any_sync(kFull, dig[r]);
        if (valid) {
            unsigned rank = __popc(peers & ((1u << lane) - 1u));
            unsigned dst = wcount[warp][dig[r]] + rank;
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in[i];
        }
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) wcount[warp][dig[r]] += __popc(peers);
        __syncwarp();
    }
}

// OR and AND of all sortable keys: a digit whose bits are the same in every key needs no pass
__global__ void __launch_bounds__(256) radix_key_span_kernel(const unsigned long long* __restrict__ keys, long long n, int kind,
                                                            unsigned long long* __restrict__ span /* [0] OR, [1] AND */) {
    unsigned long long o = 0ull, a = ~0ull;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = sortable(keys[i], kind);
        o |= k;
        a &= k;
    }
    const unsigned olo = __reduce_or_sync(kFull, (unsigned)o), ohi = __reduce_or_sync(kFull, (unsigned)(o >> 32));
    const unsigned alo = __reduce_and_sync(kFull, (unsigned)a), ahi = __reduce_and_sync(kFull, (unsigned)(a >> 32));
    if (lane_id() == 0) {
        atomicOr(span, ((unsigned long long)ohi << 32) | olo);
        atomicAnd(span + 1, ((unsigned long long)ahi << 32) | alo);
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t radix_sort_workspace_bytes(long long n) {
    long long nblocks = (std::max<long long>(n, 1) + kSortTile - 1) / kSortTile;
    return align256((size_t)std::max<long long>(n, 1) * 8) * 2 + align256((size_t)nblocks * 256 * 4 + 1024 * 4) + 256;
}

int radix_sort_pairs(const unsigned long long* keys_in, const long long* vals_in, unsigned long long* keys_out,
                     long long* vals_out, long long n, KeyKind kind, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
    if (n <= 0) return WFB_OK;
    WFB_REQUIRE(n < (1ll << 32), "radix sort: more than 2^32 elements");
    WFB_REQUIRE(workspace_bytes >= radix_sort_workspace_bytes(n), "radix sort: workspace too small");
    const int nblocks = (int)((n + kSortTile - 1) / kSortTile);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    unsigned long long* tmp_k = reinterpret_cast<unsigned long long*>(ws);
    long long* tmp_v = reinterpret_cast<long long*>(ws + align256((size_t)n * 8));
    unsigned* hist = reinterpret_cast<unsigned*>(ws + 2 * align256((size_t)n * 8));
    // the counter scan runs over <= 1024 tiles
    const long long total = (long long)nblocks * 256;
    const long long tile = std::max<long long>(4096, (total + 1023) / 1024);
    const int ntiles = (int)((total + tile - 1) / tile);
    unsigned* tile_sums = hist + total;
    // which of the eight digits differ between keys at all (time stamps of one run share their top bytes, board /
    // channel keys all but the lowest): one reduction + a 16-byte read-back saves every pass over a constant digit
    unsigned long long* span = reinterpret_cast<unsigned long long*>(ws + radix_sort_workspace_bytes(n) - 256);
    WFB_CUDA(cudaMemsetAsync(span, 0, 8, st));
    WFB_CUDA(cudaMemsetAsync(span + 1, 0xff, 8, st));
    radix_key_span_kernel<<<(unsigned)std::min<long long>(1184, (n + 255) / 256), 256, 0, st>>>(keys_in, n, (int)kind, span);
    unsigned long long h_span[2] = {0ull, 0ull};
    WFB_CUDA(cudaMemcpyAsync(h_span, span, 16, cudaMemcpyDeviceToHost, st));
    WFB_CUDA(cudaStreamSynchronize(st));
    const unsigned long long differ = h_span[0] ^ h_span[1];
    int passes[8], np = 0;
    for (int pass = 0; pass < 8; ++pass)
        if ((differ >> (8 * pass)) & 0xffull) passes[np++] = pass;
    if (np == 0) {  // all keys equal: the stable order is the input order
        WFB_CUDA(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
        WFB_CUDA(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
        return WFB_OK;
    }
    const unsigned long long* src_k = keys_in;
    const long long* src_v = vals_in;
    for (int j = 0; j < np; ++j) {
        const int pass = passes[j];
        const bool to_out = ((np - 1 - j) & 1) == 0;  // the last pass lands in keys_out
        unsigned long long* dst_k = to_out ? keys_out : tmp_k;
        long long* dst_v = to_out ? vals_out : tmp_v;
        radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(src_k, n, pass * 8, (int)kind, hist, nblocks);
        radix_tile_sums_kernel<<<ntiles, 1024, 0, st>>>(hist, total, tile, tile_sums);
        radix_scan_sums_kernel<<<1, 1024, 0, st>>>(tile_sums, ntiles);
        radix_scan_tiles_kernel<<<ntiles, 1024, 0, st>>>(hist, total, tile, tile_sums);
        radix_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(src_k, src_v, dst_k, dst_v, n, pass * 8, (int)kind, hist, nblocks);
        src_k = dst_k;
        src_v = dst_v;
    }
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

// ---- scans -------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct OpSumI64 {
    typedef long long T;
    __device__ static T id() { return 0; }
    __device__ static T op(T a, T b) { return a + b; }
};
struct OpMaxF64 {
    typedef double T;
    __device__ static T id() { return __longlong_as_double((long long)0xfff0000000000000ull); }  // -inf
    __device__ static T op(T a, T b) { return fmax(a, b); }
};

struct OpSegMax {
    typedef SegMax T;
    __device__ static T id() { return SegMax{__longlong_as_double((long long)0xfff0000000000000ull), -1}; }
    __device__ static T op(T a, T b) {
        if (b.seg < 0) return a;
        if (a.seg == b.seg) b.v = fmax(a.v, b.v);
        return b;
    }
};

template <typename Op>
__device__ typename Op::T block_inclusive_scan(typename Op::T v, typename Op::T* smem /*[kScanThreads]*/) {
    typedef typename Op::T T;
    smem[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < kScanThreads; d <<= 1) {
        T t = threadIdx.x >= d ? smem[threadIdx.x - d] : Op::id();
        __syncthreads();
        if (threadIdx.x >= d) smem[threadIdx.x] = Op::op(t, smem[threadIdx.x]);
        __syncthreads();
    }
    return smem[threadIdx.x];
}

template <typename Op>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const typename Op::T* __restrict__ in, long long n,
                                                                   typename Op::T* __restrict__ partials) {
    typedef typename Op::T T;
    __shared__ T smem[kScanThreads];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    T acc = Op::id();
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) acc = Op::op(acc, in[base + k]);
    T incl = block_inclusive_scan<Op>(acc, smem);
    if (threadIdx.x == kScanThreads - 1) partials[blockIdx.x] = incl;
}
template <typename Op>
__global__ void __launch_bounds__(1024) scan_partials_kernel(typename Op::T* __restrict__ partials, long long m) {
    typedef typename Op::T T;
    __shared__ T sums[1024];
    const long long chunk = (m + 1023) / 1024;
    const long long lo = min(m, chunk * (long long)threadIdx.x), hi = min(m, lo + chunk);
    T s = Op::id();
    for (long long i = lo; i < hi; ++i) s = Op::op(s, partials[i]);
    sums[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        T t = threadIdx.x >= d ? sums[threadIdx.x - d] : Op::id();
        __syncthreads();
        if (threadIdx.x >= d) sums[threadIdx.x] = Op::op(t, sums[threadIdx.x]);
        __syncthreads();
    }
    T run = threadIdx.x ? sums[threadIdx.x - 1] : Op::id();  // exclusive carry for this chunk
    for (long long i = lo; i < hi; ++i) {
        T c = partials[i];
        partials[i] = run;  // exclusive prefix of the tiles
        run = Op::op(run, c);
    }
}
template <typename Op>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const typename Op::T* __restrict__ in, long long n,
                                                                  const typename Op::T* __restrict__ partials,
                                                                  typename Op::T* __restrict__ out) {
    typedef typename Op::T T;
    __shared__ T smem[kScanThreads];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    T v[kScanItems];
    T acc = Op::id();
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : Op::id();
        acc = Op::op(acc, v[k]);
        v[k] = acc;
    }
    T incl = block_inclusive_scan<Op>(acc, smem);
    __syncthreads();
    T before = threadIdx.x ? smem[threadIdx.x - 1] : Op::id();
    T carry = Op::op(partials[blockIdx.x], before);
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) out[base + k] = Op::op(carry, v[k]);
    (void)incl;
}

size_t scan_workspace_bytes(long long n) {
    long long tiles = (std::max<long long>(n, 1) + kScanTile - 1) / kScanTile;
    return align256((size_t)tiles * 8) + 256;
}

template <typename Op>
static int run_scan(const typename Op::T* in, typename Op::T* out, long long n, void* workspace, cudaStream_t st) {
    if (n <= 0) return WFB_OK;
    const int tiles = (int)((n + kScanTile - 1) / kScanTile);
    typename Op::T* partials = static_cast<typename Op::T*>(workspace);
    scan_reduce_kernel<Op><<<tiles, kScanThreads, 0, st>>>(in, n, partials);
    scan_partials_kernel<Op><<<1, 1024, 0, st>>>(partials, tiles);
    scan_apply_kernel<Op><<<tiles, kScanThreads, 0, st>>>(in, n, partials, out);
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

int inclusive_scan_sum_i64(const long long* in, long long* out, long long n, void* workspace, cudaStream_t st) {
    return run_scan<OpSumI64>(in, out, n, workspace, st);
}
int inclusive_scan_max_f64(const double* in, double* out, long long n, void* workspace, cudaStream_t st) {
    return run_scan<OpMaxF64>(in, out, n, workspace, st);
}
int inclusive_scan_segmax(const SegMax* in, SegMax* out, long long n, void* workspace, cudaStream_t st) {
    return run_scan<OpSegMax>(in, out, n, workspace, st);
}

}  // namespace wfb

using namespace wfb;

extern "C" size_t wfb_sort_workspace_bytes(int64_t n) { return radix_sort_workspace_bytes(n); }

extern "C" int wfb_sort_pairs_i64(const int64_t* keys_in_dev, const int64_t* vals_in_dev, int64_t* keys_out_dev,
                                  int64_t* vals_out_dev, int64_t n, void* workspace_dev, size_t workspace_bytes,
                                  void* stream) {
    WFB_REQUIRE(n >= 0, "wfb_sort_pairs_i64: negative n");
    if (n == 0) return WFB_OK;
    WFB_REQUIRE(keys_in_dev && vals_in_dev && keys_out_dev && vals_out_dev && workspace_dev, "wfb_sort_pairs_i64: NULL pointer");
    return radix_sort_pairs(reinterpret_cast<const unsigned long long*>(keys_in_dev), reinterpret_cast<const long long*>(vals_in_dev),
                            reinterpret_cast<unsigned long long*>(keys_out_dev), reinterpret_cast<long long*>(vals_out_dev), n,
                            kKeySigned, workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}
