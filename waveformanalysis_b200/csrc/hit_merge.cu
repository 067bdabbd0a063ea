// hit_merge on the device: per hardware channel, hits ordered by absolute window start are chained
// while merge_gap_ns > 0, the sampling interval matches, the gap to the running cluster end is
// <= merge_gap and the total width stays <= max_total_width; every cluster becomes one HIT_MERGED
// row (anchor = highest member, earliest timestamp among equals).
//
// Reference: core/plugins/builtin/cpu/hit_merge.py:115-181 (_build_merged_clusters),
// :256-322 (_emit_cluster), :60-84 (abs_start / abs_end in float64).
//
// The chain is greedy (the width cap refers to the start of the running cluster), so it is not a
// plain scan.  It is cut where a break is certain - channel or dt changes, or the window start lies
// more than merge_gap behind the running maximum of ALL earlier window ends of the channel (the
// cluster end can only be smaller) - and one thread replays the reference loop inside each piece.
#include <algorithm>

#include "common.cuh"
#include "np_sum.cuh"
#include "sort_scan.cuh"

namespace wfb {

constexpr int kMergedRowBytes = 72;  // HIT_MERGED_DTYPE

static size_t hm_al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct HitCols {  // a THRESHOLD_HIT row (60 B packed, hit_finder.py:33-49) as 15 words
    const unsigned* w;
    __device__ long long position() const { return (long long)(((unsigned long long)w[1] << 32) | w[0]); }
    __device__ float height() const { return __uint_as_float(w[2]); }
    __device__ float integral() const { return __uint_as_float(w[3]); }
    __device__ int edge_start() const { return (int)w[4]; }
    __device__ int edge_end() const { return (int)w[5]; }
    __device__ int dt() const { return (int)w[7]; }
    __device__ long long timestamp() const { return (long long)(((unsigned long long)w[11] << 32) | w[10]); }
    __device__ unsigned board_channel() const { return w[12]; }
    __device__ long long record_id() const { return (long long)(((unsigned long long)w[14] << 32) | w[13]); }
};
__device__ __forceinline__ HitCols hit_at(const uint8_t* rows, long long i) {
    return HitCols{reinterpret_cast<const unsigned*>(rows + i * kHitRowBytes)};
}

// abs_start (sortable float64 key), abs_end, channel key, identity payload
__global__ void hm_columns_kernel(const uint8_t* __restrict__ rows, long long n, double* __restrict__ a0, double* __restrict__ a1,
                                  long long* __restrict__ idx) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const HitCols h = hit_at(rows, i);
    const double dt_ps = __dmul_rn((double)h.dt(), 1e3);
    const double t = (double)h.timestamp(), p = (double)h.position();
    a0[i] = __dadd_rn(t, __dmul_rn(__dsub_rn((double)h.edge_start(), p), dt_ps));
    a1[i] = __dadd_rn(t, __dmul_rn(__dsub_rn((double)h.edge_end(), p), dt_ps));
    idx[i] = i;
}
// (board, channel) as one unsigned key in the reference's order: ascending board, then channel (signed)
__global__ void hm_chan_keys_kernel(const uint8_t* __restrict__ rows, const long long* __restrict__ idx, long long n,
                                    unsigned long long* __restrict__ key) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned bc = hit_at(rows, idx[i]).board_channel();
    const unsigned b = (unsigned)((int)(short)(bc & 0xffffu) + 32768), c = (unsigned)((int)(short)(bc >> 16) + 32768);
    key[i] = ((unsigned long long)b << 16) | c;
}
// pieces that certainly start a cluster: first hit, other channel, other dt
__global__ void hm_key_flags_kernel(const uint8_t* __restrict__ rows, const long long* __restrict__ order, long long n,
                                    long long* __restrict__ flag) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    bool f = k == 0;
    if (!f) {
        const HitCols a = hit_at(rows, order[k - 1]), b = hit_at(rows, order[k]);
        f = a.board_channel() != b.board_channel() || a.dt() != b.dt();
    }
    flag[k] = f ? 1 : 0;
}
__global__ void hm_pack_kernel(const double* __restrict__ a1, const long long* __restrict__ order, const long long* __restrict__ seg,
                               long long n, SegMax* __restrict__ out) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k].v = a1[order[k]];
    out[k].seg = seg[k];
}
__global__ void hm_hard_breaks_kernel(const double* __restrict__ a0, const long long* __restrict__ order, const long long* __restrict__ kflag,
                                      const SegMax* __restrict__ pm, long long n, double gap_ps, int chain, long long* __restrict__ hb) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    bool f = kflag[k] != 0 || !chain;
    if (!f) f = __dsub_rn(a0[order[k]], pm[k - 1].v) > gap_ps;  // even the largest possible cluster end is too far back
    hb[k] = f ? 1 : 0;
}
// one thread per piece replays hit_merge.py:150-177
__global__ void hm_walk_kernel(const double* __restrict__ a0, const double* __restrict__ a1, const long long* __restrict__ order,
                               const long long* __restrict__ hb, long long n, double gap_ps, double maxw_ps,
                               long long* __restrict__ start_flag) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n || !hb[k]) return;
    start_flag[k] = 1;
    double c_start = a0[order[k]], c_end = a1[order[k]];
    for (long long j = k + 1; j < n && !hb[j]; ++j) {
        const long long i = order[j];
        const double s = a0[i], e = a1[i];
        const double next_end = fmax(c_end, e);
        const bool join = __dsub_rn(s, c_end) <= gap_ps && __dsub_rn(next_end, c_start) <= maxw_ps;
        if (join) {
            c_end = next_end;
            start_flag[j] = 0;
        } else {
            start_flag[j] = 1;
            c_start = s;
            c_end = e;
        }
    }
}
__global__ void hm_cluster_starts_kernel(const long long* __restrict__ start_flag, const long long* __restrict__ incl, long long n,
                                         long long* __restrict__ cluster_index, long long* __restrict__ cstart,
                                         long long* __restrict__ n_clusters) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long c = incl[k] - 1;
    cluster_index[k] = c;
    if (start_flag[k]) cstart[c] = k;
    if (k == n - 1) *n_clusters = c + 1;
}

struct IntegralSrc {
    const uint8_t* rows;
    const long long* order;
    long long base;
    __device__ double operator()(int i) const { return (double)hit_at(rows, order[base + i]).integral(); }
};

// one thread per cluster: the HIT_MERGED row (hit_merge.py:256-322)
__global__ void hm_rows_kernel(const uint8_t* __restrict__ rows, const long long* __restrict__ order, const long long* __restrict__ cstart,
                               const long long* __restrict__ n_clusters, long long n, uint8_t* __restrict__ merged) {
    long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long ncl = *n_clusters;
    if (c >= ncl) return;
    const long long s = cstart[c], e = (c + 1 < ncl) ? cstart[c + 1] : n;
    const int cnt = (int)(e - s);
    long long anchor = order[s];
    float height = hit_at(rows, anchor).height(), integral = hit_at(rows, anchor).integral();
    int smp0 = hit_at(rows, anchor).edge_start(), smp1 = hit_at(rows, anchor).edge_end();
    float width = __uint_as_float(hit_at(rows, anchor).w[6]);
    if (cnt > 1) {
        // highest member; among equal heights the earliest timestamp, then the first in cluster order
        double max_h = (double)height;
        long long best_ts = hit_at(rows, anchor).timestamp();
        const long long rid0 = hit_at(rows, anchor).record_id();
        bool one_record = true;
        for (long long k = s + 1; k < e; ++k) {
            const HitCols h = hit_at(rows, order[k]);
            const double hh = (double)h.height();
            if (hh > max_h || (hh == max_h && h.timestamp() < best_ts)) {
                max_h = hh;
                best_ts = h.timestamp();
                anchor = order[k];
            }
            one_record = one_record && h.record_id() == rid0;
            smp0 = min(smp0, h.edge_start());
            smp1 = max(smp1, h.edge_end());
        }
        height = (float)max_h;
        IntegralSrc src{rows, order, s};
        integral = (float)numpy_pairwise_sum(src, cnt);
        if (!one_record) { smp0 = -1; smp1 = -1; }
        width = (smp0 < 0 || smp1 < 0) ? -1.0f : (float)fmax((double)(smp1 - smp0), 0.0);
    }
    const HitCols a = hit_at(rows, anchor);
    unsigned* dst = reinterpret_cast<unsigned*>(merged + c * kMergedRowBytes);
    dst[0] = a.w[0]; dst[1] = a.w[1];                 // position
    dst[2] = __float_as_uint(height);
    dst[3] = __float_as_uint(integral);
    dst[4] = (unsigned)smp0;
    dst[5] = (unsigned)smp1;
    dst[6] = __float_as_uint(width);
    dst[7] = a.w[7];                                  // dt
    dst[8] = a.w[8]; dst[9] = a.w[9];                 // rise_time, fall_time
    dst[10] = a.w[10]; dst[11] = a.w[11];             // timestamp
    dst[12] = a.w[12];                                // board, channel
    dst[13] = a.w[13]; dst[14] = a.w[14];             // record_id
    dst[15] = (unsigned)(s & 0xffffffffll);           // component_offset
    dst[16] = (unsigned)((unsigned long long)s >> 32);
    dst[17] = (unsigned)cnt;                          // component_count
}

// grouping columns of packed hit rows (THRESHOLD_HIT 60 B or HIT_MERGED 72 B: the first 15 words are laid out alike)
__global__ void hit_columns_kernel(const uint8_t* __restrict__ rows, long long n, int row_bytes, long long* __restrict__ ts,
                                   int* __restrict__ dt, long long* __restrict__ rid, double* __restrict__ a0, double* __restrict__ a1) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const HitCols h{reinterpret_cast<const unsigned*>(rows + i * row_bytes)};
    ts[i] = h.timestamp();
    dt[i] = h.dt();
    rid[i] = h.record_id();
    if (a0 != nullptr) {
        const double dt_ps = __dmul_rn((double)h.dt(), 1e3);
        const double t = (double)h.timestamp(), p = (double)h.position();
        a0[i] = __dadd_rn(t, __dmul_rn(__dsub_rn((double)h.edge_start(), p), dt_ps));
        a1[i] = __dadd_rn(t, __dmul_rn(__dsub_rn((double)h.edge_end(), p), dt_ps));
    }
}

// absolute window of every merged row = min / max over the windows of its component hits (event_grouping.py:369-414)
__global__ void merged_windows_kernel(const uint8_t* __restrict__ hits, const long long* __restrict__ order, const uint8_t* __restrict__ merged,
                                      long long n_clusters, double* __restrict__ a0, double* __restrict__ a1) {
    long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= n_clusters) return;
    const unsigned* m = reinterpret_cast<const unsigned*>(merged + c * kMergedRowBytes);
    const long long s = (long long)(((unsigned long long)m[16] << 32) | m[15]);
    const int cnt = (int)m[17];
    double lo = __longlong_as_double(0x7ff0000000000000ll), hi = -lo;
    for (int k = 0; k < cnt; ++k) {
        const HitCols h = hit_at(hits, order[s + k]);
        const double dt_ps = __dmul_rn((double)h.dt(), 1e3);
        const double t = (double)h.timestamp(), p = (double)h.position();
        lo = fmin(lo, __dadd_rn(t, __dmul_rn(__dsub_rn((double)h.edge_start(), p), dt_ps)));
        hi = fmax(hi, __dadd_rn(t, __dmul_rn(__dsub_rn((double)h.edge_end(), p), dt_ps)));
    }
    a0[c] = lo;
    a1[c] = hi;
}

}  // namespace wfb

using namespace wfb;

static unsigned hm_nb(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

extern "C" size_t wfb_hit_merge_workspace_bytes(int64_t n) {
    const size_t m = hm_al256((size_t)std::max<int64_t>(n, 1) * 8);
    return 8 * m + 2 * hm_al256((size_t)std::max<int64_t>(n, 1) * 16) + radix_sort_workspace_bytes(n) + 2 * scan_workspace_bytes(n) + 512;
}

extern "C" int wfb_hit_merge(const void* hits_dev, int64_t n, double merge_gap_ns, double max_total_width_ns, int64_t* order_dev,
                             int64_t* cluster_index_dev, void* merged_dev, int64_t* n_clusters_dev, void* workspace_dev,
                             size_t workspace_bytes, void* stream) {
    WFB_REQUIRE(n >= 0, "wfb_hit_merge: negative n");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (n_clusters_dev) WFB_CUDA(cudaMemsetAsync(n_clusters_dev, 0, 8, st));
        return WFB_OK;
    }
    WFB_REQUIRE(hits_dev && order_dev && cluster_index_dev && merged_dev && n_clusters_dev && workspace_dev, "wfb_hit_merge: NULL pointer");
    WFB_REQUIRE(((uintptr_t)hits_dev & 3) == 0 && ((uintptr_t)merged_dev & 3) == 0, "wfb_hit_merge: rows must be 4-byte aligned");
    WFB_REQUIRE(workspace_bytes >= wfb_hit_merge_workspace_bytes(n), "wfb_hit_merge: workspace too small");
    const uint8_t* rows = static_cast<const uint8_t*>(hits_dev);
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    const size_t m = hm_al256((size_t)n * 8), m16 = hm_al256((size_t)n * 16);
    double* a0 = reinterpret_cast<double*>(ws);
    double* a1 = reinterpret_cast<double*>(ws + m);
    unsigned long long* kA = reinterpret_cast<unsigned long long*>(ws + 2 * m);
    long long* vA = reinterpret_cast<long long*>(ws + 3 * m);
    unsigned long long* kB = reinterpret_cast<unsigned long long*>(ws + 4 * m);
    long long* vB = reinterpret_cast<long long*>(ws + 5 * m);
    long long* f0 = reinterpret_cast<long long*>(ws + 6 * m);
    long long* f1 = reinterpret_cast<long long*>(ws + 7 * m);
    SegMax* sm_in = reinterpret_cast<SegMax*>(ws + 8 * m);
    SegMax* sm_out = reinterpret_cast<SegMax*>(ws + 8 * m + m16);
    uint8_t* sws = ws + 8 * m + 2 * m16;
    const size_t sort_bytes = radix_sort_workspace_bytes(n);
    void* scan_ws = sws + sort_bytes;
    long long* order = reinterpret_cast<long long*>(order_dev);
    long long* cidx = reinterpret_cast<long long*>(cluster_index_dev);
    long long* ncl = reinterpret_cast<long long*>(n_clusters_dev);

    // order: stable by abs_start, then stable by (board, channel)
    hm_columns_kernel<<<hm_nb(n), 256, 0, st>>>(rows, n, a0, a1, vA);
    WFB_CUDA(cudaMemcpyAsync(kA, a0, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    int rc = radix_sort_pairs(kA, vA, kB, vB, n, kKeyFloat64, sws, sort_bytes, st);
    if (rc != WFB_OK) return rc;
    hm_chan_keys_kernel<<<hm_nb(n), 256, 0, st>>>(rows, vB, n, kA);
    rc = radix_sort_pairs(kA, vB, kB, order, n, kKeyUnsigned, sws, sort_bytes, st);
    if (rc != WFB_OK) return rc;
    // certain breaks
    const int chain = merge_gap_ns > 0.0 ? 1 : 0;
    hm_key_flags_kernel<<<hm_nb(n), 256, 0, st>>>(rows, order, n, f0);
    rc = inclusive_scan_sum_i64(f0, f1, n, scan_ws, st);
    if (rc != WFB_OK) return rc;
    hm_pack_kernel<<<hm_nb(n), 256, 0, st>>>(a1, order, f1, n, sm_in);
    rc = inclusive_scan_segmax(sm_in, sm_out, n, scan_ws, st);
    if (rc != WFB_OK) return rc;
    hm_hard_breaks_kernel<<<hm_nb(n), 256, 0, st>>>(a0, order, f0, sm_out, n, merge_gap_ns * 1e3, chain, f1);
    // greedy replay inside each piece, cluster numbering, rows
    hm_walk_kernel<<<hm_nb(n), 256, 0, st>>>(a0, a1, order, f1, n, merge_gap_ns * 1e3, max_total_width_ns * 1e3, f0);
    rc = inclusive_scan_sum_i64(f0, vA, n, scan_ws, st);
    if (rc != WFB_OK) return rc;
    hm_cluster_starts_kernel<<<hm_nb(n), 256, 0, st>>>(f0, vA, n, cidx, vB, ncl);
    hm_rows_kernel<<<hm_nb(n), 256, 0, st>>>(rows, order, vB, ncl, n, static_cast<uint8_t*>(merged_dev));
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

extern "C" int wfb_hit_columns(const void* rows_dev, int64_t n, int32_t row_bytes, int64_t* timestamp_dev, int32_t* dt_dev,
                               int64_t* record_id_dev, double* abs_start_dev, double* abs_end_dev, void* stream) {
    WFB_REQUIRE(n >= 0 && (row_bytes == kHitRowBytes || row_bytes == kMergedRowBytes), "wfb_hit_columns: bad arguments");
    if (n == 0) return WFB_OK;
    WFB_REQUIRE(rows_dev && timestamp_dev && dt_dev && record_id_dev, "wfb_hit_columns: NULL pointer");
    WFB_REQUIRE((abs_start_dev == nullptr) == (abs_end_dev == nullptr), "wfb_hit_columns: abs_start / abs_end go together");
    hit_columns_kernel<<<hm_nb(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(rows_dev), n, row_bytes,
                                                                               reinterpret_cast<long long*>(timestamp_dev), dt_dev,
                                                                               reinterpret_cast<long long*>(record_id_dev), abs_start_dev, abs_end_dev);
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

extern "C" int wfb_merged_abs_windows(const void* hits_dev, const int64_t* order_dev, const void* merged_dev, int64_t n_clusters,
                                      double* abs_start_dev, double* abs_end_dev, void* stream) {
    WFB_REQUIRE(n_clusters >= 0, "wfb_merged_abs_windows: negative count");
    if (n_clusters == 0) return WFB_OK;
    WFB_REQUIRE(hits_dev && order_dev && merged_dev && abs_start_dev && abs_end_dev, "wfb_merged_abs_windows: NULL pointer");
    merged_windows_kernel<<<hm_nb(n_clusters), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t*>(hits_dev), reinterpret_cast<const long long*>(order_dev), static_cast<const uint8_t*>(merged_dev), n_clusters,
        abs_start_dev, abs_end_dev);
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}
