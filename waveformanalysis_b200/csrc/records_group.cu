// K1: raw int16 rows -> time-sorted records + wave_pool (+ baseline);
// K4: event grouping (chain clustering of absolute hit windows, anchored fixed windows).
//
// Reference: core/processing/records_builder.py:115-120, 212-302, 341-426 (K1);
//            core/processing/event_grouping.py:365-367, 418, 453-470, 477-510 (K4).
#include <algorithm>

#include "sort_scan.cuh"

namespace wfb {

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// ================================================================================================
// K1
// ================================================================================================
__global__ void k1_keys_kernel(const short* __restrict__ board, const short* __restrict__ channel, long long n,
                               unsigned long long* __restrict__ key, long long* __restrict__ val) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // (board, channel) ascending as signed int16 pairs
    key[i] = ((unsigned long long)(unsigned)((int)board[i] + 32768) << 16) | (unsigned long long)(unsigned)((int)channel[i] + 32768);
    val[i] = i;
}
__global__ void gather_i64_kernel(const long long* __restrict__ src, const long long* __restrict__ idx, long long n,
                                  unsigned long long* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned long long)src[idx[i]];
}

// halfword k (0..50) of the packed RECORDS_DTYPE row (core/processing/dtypes.py:80-100)
__device__ __forceinline__ uint16_t rec_halfword(int k, const wfb_rec_meta& m, long long time_ns, unsigned flags) {
    auto h64 = [](unsigned long long v, int q) { return (uint16_t)(v >> (16 * q)); };
    if (k < 4) return h64((unsigned long long)m.timestamp, k);
    if (k < 6) return 0;  // pid
    if (k == 6) return (uint16_t)m.board;
    if (k == 7) return (uint16_t)m.channel;
    if (k < 12) return h64((unsigned long long)__double_as_longlong(m.baseline), k - 8);
    if (k < 16) return h64(0x7ff8000000000000ull, k - 12);  // baseline_upstream = NaN
    if (k < 32) {
        const char* s = "unknown";
        int ci = (k - 16) >> 1;
        return ((k & 1) == 0 && ci < 7) ? (uint16_t)s[ci] : (uint16_t)0;
    }
    if (k < 36) return h64((unsigned long long)m.record_id, k - 32);
    if (k < 38) return (uint16_t)((unsigned)m.dt >> (16 * (k - 36)));
    if (k == 38) return 0;  // trigger_type
    if (k < 41) return (uint16_t)(flags >> (16 * (k - 39)));
    if (k < 45) return h64((unsigned long long)m.wave_offset, k - 41);
    if (k < 47) return (uint16_t)((unsigned)m.event_length >> (16 * (k - 45)));
    return h64((unsigned long long)time_ns, k - 47);
}

// one warp per OUTPUT record: gather the row, mean of the baseline window, pack the outputs
__global__ void __launch_bounds__(256) k1_gather_kernel(const short* __restrict__ samples, const long long* __restrict__ ts,
                                                       const short* __restrict__ board, const short* __restrict__ channel,
                                                       const double* __restrict__ baselines_in,
                                                       const long long* __restrict__ order, long long n, int L, int bl_start,
                                                       int bl_end, int dt_ns, long long epoch_ns,
                                                       uint16_t* __restrict__ rows, uint16_t* __restrict__ pool,
                                                       wfb_rec_meta* __restrict__ meta) {
    const int lane = lane_id();
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const long long src = order[r];
    const short* in = samples + src * (long long)L;
    uint16_t* out = pool + r * (long long)L;
    long long bsum = 0;
    const int be = min(bl_end, L);
    const bool vec = ((L & 7) == 0);
    if (vec) {
        for (int j0 = lane * 8; j0 < L; j0 += 256) {
            uint4 q = *reinterpret_cast<const uint4*>(in + j0);
            *reinterpret_cast<uint4*>(out + j0) = q;
            const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int j = j0 + k;
                int v = (int)(short)((w[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
                if (j >= bl_start && j < be) bsum += v;
            }
        }
    } else {
        for (int j = lane; j < L; j += 32) {
            short v = in[j];
            out[j] = (uint16_t)v;
            if (j >= bl_start && j < be) bsum += v;
        }
    }
    bsum = warp_sum_i64(bsum);
    wfb_rec_meta m;
    m.timestamp = ts[src];
    if (baselines_in != nullptr) m.baseline = baselines_in[src];
    else m.baseline = (be > bl_start) ? (double)bsum / (double)(be - bl_start) : __longlong_as_double(0x7ff8000000000000ll);
    m.wave_offset = r * (long long)L;
    m.event_length = L;
    m.dt = dt_ns;
    m.board = board[src];
    m.channel = channel[src];
    m.polarity = WFB_POL_UNKNOWN;
    m.pad_[0] = m.pad_[1] = m.pad_[2] = 0;
    m.record_id = r;
    if (meta != nullptr && lane == 0) meta[r] = m;
    if (rows != nullptr) {
        // floor division like numpy's // for negative timestamps
        long long t = m.timestamp / 1000;
        if ((m.timestamp % 1000) != 0 && m.timestamp < 0) --t;
        const long long time_ns = t + epoch_ns;
        uint16_t* row = rows + r * (kRecordsRowBytes / 2);
        row[lane] = rec_halfword(lane, m, time_ns, 0u);
        if (lane + 32 < 51) row[lane + 32] = rec_halfword(lane + 32, m, time_ns, 0u);
    }
}

// ================================================================================================
// K4
// ================================================================================================
__global__ void k4_abs_windows_kernel(const long long* __restrict__ ts, const long long* __restrict__ pos,
                                      const int* __restrict__ start, const int* __restrict__ end, const int* __restrict__ dt,
                                      long long n, double* __restrict__ abs0, double* __restrict__ abs1) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dt_ps = __dmul_rn((double)dt[i], 1e3);
    const double t = (double)ts[i], p = (double)pos[i];
    abs0[i] = __dadd_rn(t, __dmul_rn(__dsub_rn((double)start[i], p), dt_ps));
    abs1[i] = __dadd_rn(t, __dmul_rn(__dsub_rn((double)end[i], p), dt_ps));
}
__global__ void iota_kernel(long long* v, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
__global__ void gather_i32_as_i64_kernel(const int* __restrict__ src, const long long* __restrict__ idx, long long n,
                                         unsigned long long* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned long long)(long long)src[idx[i]];
}
__global__ void gather_f64_kernel(const double* __restrict__ src, const long long* __restrict__ idx, long long n,
                                  double* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
// new-cluster flags from the running maximum of the window ends (event_grouping.py:462)
__global__ void k4_chain_flags_kernel(const double* __restrict__ abs0_sorted, const double* __restrict__ pm, long long n,
                                      double gap_ps, long long* __restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || abs0_sorted[i] > __dadd_rn(pm[i - 1], gap_ps)) ? 1 : 0;
}
__global__ void k4_scatter_ids_kernel(const long long* __restrict__ incl, const long long* __restrict__ order, long long n,
                                      long long* __restrict__ event_id, long long* __restrict__ n_events) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long ev = incl[i] - 1;
    if (order != nullptr) event_id[order[i]] = ev;
    else event_id[i] = ev;
    if (i == n - 1) *n_events = ev + 1;
}

// anchored fixed windows (event_grouping.py:496-508): a gap > W always starts a cluster, inside a
// gap-free segment the anchors are found by one thread walking the segment
__global__ void k4_segment_flags_kernel(const long long* __restrict__ ts, long long n, double w_ps, long long* __restrict__ seg) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    seg[i] = (i == 0 || (double)ts[i] > __dadd_rn((double)ts[i - 1], w_ps)) ? 1 : 0;
}
__global__ void k4_anchor_walk_kernel(const long long* __restrict__ ts, long long n, double w_ps,
                                      const long long* __restrict__ seg, long long* __restrict__ anchor) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !seg[i]) return;
    long long cur = i;
    for (;;) {
        anchor[cur] = 1;
        const double target = __dadd_rn((double)ts[cur], w_ps);
        long long nxt = cur + 1;
        while (nxt < n && !seg[nxt] && (double)ts[nxt] <= target) ++nxt;
        if (nxt >= n || seg[nxt]) break;
        cur = nxt;
    }
}


// ---- st_waveforms: structured rows (core/processing/dtypes.py:36-64, 76 header bytes + wave_length int16 samples) ----
// One warp per raw row, in input order (channel by channel, waveforms.py:644-799, 906-913): the mean over the baseline
// window is an integer sum divided once (exact, any summation order), the upstream baseline is carried along, the
// samples are copied (truncated to wave_length, the rest of the row stays 0) and event_length records how many.
__global__ void __launch_bounds__(256) st_structure_kernel(const short* __restrict__ samples, long long n, int l_src, int wave_len,
                                                           const long long* __restrict__ ts, const short* __restrict__ board,
                                                           const short* __restrict__ chan, const double* __restrict__ bl_in,
                                                           const double* __restrict__ bl_up, int bl_lo, int bl_hi, int dt_ns,
                                                           uint8_t* __restrict__ out, long long record_base) {
    const int lane = lane_id();
    const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const short* src = samples + i * (long long)l_src;
    const long long row_bytes = 76 + 2ll * wave_len;
    unsigned short* dst = reinterpret_cast<unsigned short*>(out + i * row_bytes);  // rows are 2-byte aligned
    const int n_copy = min(l_src, wave_len);
    double baseline;
    if (bl_in != nullptr) {
        baseline = bl_in[i];
    } else if (bl_hi > bl_lo) {
        long long acc = 0;
        for (int j = bl_lo + lane; j < bl_hi; j += 32) acc += src[j];
        acc = warp_sum_i64(acc);
        baseline = (double)acc / (double)(bl_hi - bl_lo);
    } else {
        baseline = __longlong_as_double(0x7ff8000000000000ll);
    }
    const double upstream = bl_up != nullptr ? bl_up[i] : __longlong_as_double(0x7ff8000000000000ll);
    // the 38 half-words of the header
    auto q = [](unsigned long long v, int k) { return (unsigned short)(v >> (16 * k)); };
    for (int k = lane; k < 38; k += 32) {
        unsigned short h = 0;
        if (k < 4) h = q((unsigned long long)__double_as_longlong(baseline), k);
        else if (k < 8) h = q((unsigned long long)__double_as_longlong(upstream), k - 4);
        else if (k < 24) {  // polarity 'unknown' as UTF-32
            const char* u = "unknown";
            const int ci = (k - 8) >> 1;
            h = ((k & 1) == 0 && ci < 7) ? (unsigned short)u[ci] : (unsigned short)0;
        } else if (k < 28) h = q((unsigned long long)ts[i], k - 24);
        else if (k < 32) h = q((unsigned long long)(record_base + i), k - 28);
        else if (k < 34) h = (unsigned short)((unsigned)dt_ns >> (16 * (k - 32)));
        else if (k < 36) h = (unsigned short)((unsigned)n_copy >> (16 * (k - 34)));
        else if (k == 36) h = (unsigned short)board[i];
        else h = (unsigned short)chan[i];
        dst[k] = h;
    }
    unsigned short* w = dst + 38;
    for (int j = lane; j < wave_len; j += 32) w[j] = j < n_copy ? (unsigned short)src[j] : (unsigned short)0;
}

}  // namespace wfb

using namespace wfb;

static unsigned nb(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

extern "C" size_t wfb_build_records_workspace_bytes(int64_t n) {
    size_t m = al256((size_t)std::max<int64_t>(n, 1) * 8);
    return 4 * m + radix_sort_workspace_bytes(n) + 256;
}

extern "C" int wfb_build_records(const int16_t* samples_dev, const int64_t* ts_dev, const int16_t* board_dev,
                                 const int16_t* channel_dev, const double* baselines_in_dev, int64_t n, int32_t n_samples,
                                 int32_t bl_start, int32_t bl_end, int32_t dt_ns, int64_t epoch_ns, void* records_aos_dev,
                                 uint16_t* pool_dev, wfb_rec_meta* meta_dev, void* workspace_dev, size_t workspace_bytes,
                                 void* stream) {
    WFB_REQUIRE(n >= 0 && n_samples >= 0, "wfb_build_records: negative size");
    if (n == 0) return WFB_OK;
    WFB_REQUIRE(ts_dev && board_dev && channel_dev && workspace_dev, "wfb_build_records: NULL pointer");
    WFB_REQUIRE(n_samples == 0 || (samples_dev && pool_dev), "wfb_build_records: NULL sample buffers");
    WFB_REQUIRE(workspace_bytes >= wfb_build_records_workspace_bytes(n), "wfb_build_records: workspace too small");
    WFB_REQUIRE(((uintptr_t)samples_dev & 15) == 0 && ((uintptr_t)pool_dev & 15) == 0, "wfb_build_records: buffers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    const size_t m = al256((size_t)n * 8);
    unsigned long long* kA = reinterpret_cast<unsigned long long*>(ws);
    long long* vA = reinterpret_cast<long long*>(ws + m);
    unsigned long long* kB = reinterpret_cast<unsigned long long*>(ws + 2 * m);
    long long* vB = reinterpret_cast<long long*>(ws + 3 * m);
    void* sws = ws + 4 * m;
    const size_t sws_bytes = workspace_bytes - 4 * m;
    // lexsort((seq, channel, board, pid, timestamp)): stable sort by (board, channel), then by timestamp
    k1_keys_kernel<<<nb(n), 256, 0, st>>>(board_dev, channel_dev, n, kA, vA);
    int rc = radix_sort_pairs(kA, vA, kB, vB, n, kKeyUnsigned, sws, sws_bytes, st);
    if (rc != WFB_OK) return rc;
    gather_i64_kernel<<<nb(n), 256, 0, st>>>(reinterpret_cast<const long long*>(ts_dev), vB, n, kA);
    rc = radix_sort_pairs(kA, vB, kB, vA, n, kKeySigned, sws, sws_bytes, st);
    if (rc != WFB_OK) return rc;
    k1_gather_kernel<<<nb(n * 32), 256, 0, st>>>(samples_dev, reinterpret_cast<const long long*>(ts_dev), board_dev, channel_dev,
                                                 baselines_in_dev, vA, n, n_samples, bl_start, bl_end, dt_ns, epoch_ns,
                                                 static_cast<uint16_t*>(records_aos_dev), pool_dev, meta_dev);
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

extern "C" size_t wfb_group_workspace_bytes(int64_t n) {
    size_t m = al256((size_t)std::max<int64_t>(n, 1) * 8);
    return 6 * m + radix_sort_workspace_bytes(n) + scan_workspace_bytes(n) + 512;
}

// sort by (abs_start, dt, timestamp, record_id), running maximum of the window ends, boundary flags, event ids
static int group_from_abs_windows(const int64_t* timestamp_dev, const int32_t* dt_dev, const int64_t* record_id_dev, int64_t n,
                                  double time_window_ns, int64_t* order_dev, int64_t* event_id_dev, const double* abs_start_dev,
                                  const double* abs_end_dev, int64_t* n_events_dev, void* workspace_dev, cudaStream_t st) {
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    const size_t m = al256((size_t)n * 8);
    unsigned long long* kA = reinterpret_cast<unsigned long long*>(ws);
    long long* vA = reinterpret_cast<long long*>(ws + m);
    unsigned long long* kB = reinterpret_cast<unsigned long long*>(ws + 2 * m);
    long long* vB = reinterpret_cast<long long*>(ws + 3 * m);
    double* t0 = reinterpret_cast<double*>(ws + 4 * m);
    double* t1 = reinterpret_cast<double*>(ws + 5 * m);
    uint8_t* sws = ws + 6 * m;
    const size_t sort_bytes = radix_sort_workspace_bytes(n);
    void* scan_ws = sws + sort_bytes;
    const long long* ts = reinterpret_cast<const long long*>(timestamp_dev);
    const long long* rid = reinterpret_cast<const long long*>(record_id_dev);
    long long* order = reinterpret_cast<long long*>(order_dev);
    // order = lexsort((record_id, timestamp, dt, abs_start)): four stable passes, least significant first
    iota_kernel<<<nb(n), 256, 0, st>>>(vA, n);
    gather_i64_kernel<<<nb(n), 256, 0, st>>>(rid, vA, n, kA);
    int rc = radix_sort_pairs(kA, vA, kB, vB, n, kKeySigned, sws, sort_bytes, st);
    if (rc != WFB_OK) return rc;
    gather_i64_kernel<<<nb(n), 256, 0, st>>>(ts, vB, n, kA);
    rc = radix_sort_pairs(kA, vB, kB, vA, n, kKeySigned, sws, sort_bytes, st);
    if (rc != WFB_OK) return rc;
    gather_i32_as_i64_kernel<<<nb(n), 256, 0, st>>>(dt_dev, vA, n, kA);
    rc = radix_sort_pairs(kA, vA, kB, vB, n, kKeySigned, sws, sort_bytes, st);
    if (rc != WFB_OK) return rc;
    gather_f64_kernel<<<nb(n), 256, 0, st>>>(abs_start_dev, vB, n, reinterpret_cast<double*>(kA));
    rc = radix_sort_pairs(kA, vB, kB, order, n, kKeyFloat64, sws, sort_bytes, st);
    if (rc != WFB_OK) return rc;
    // running max of the window ends in sorted order, boundary flags, event ids
    gather_f64_kernel<<<nb(n), 256, 0, st>>>(abs_end_dev, order, n, t0);
    rc = inclusive_scan_max_f64(t0, t1, n, scan_ws, st);
    if (rc != WFB_OK) return rc;
    k4_chain_flags_kernel<<<nb(n), 256, 0, st>>>(reinterpret_cast<const double*>(kB), t1, n, time_window_ns * 1e3, vA);
    rc = inclusive_scan_sum_i64(vA, vB, n, scan_ws, st);
    if (rc != WFB_OK) return rc;
    k4_scatter_ids_kernel<<<nb(n), 256, 0, st>>>(vB, order, n, reinterpret_cast<long long*>(event_id_dev),
                                                 reinterpret_cast<long long*>(n_events_dev));
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

extern "C" int wfb_group_hit_windows(const int64_t* timestamp_dev, const int64_t* position_dev, const int32_t* start_dev,
                                     const int32_t* end_dev, const int32_t* dt_dev, const int64_t* record_id_dev,
                                     int64_t n, double time_window_ns, int64_t* order_dev, int64_t* event_id_dev,
                                     double* abs_start_dev, double* abs_end_dev, int64_t* n_events_dev, void* workspace_dev,
                                     size_t workspace_bytes, void* stream) {
    WFB_REQUIRE(n >= 0, "wfb_group_hit_windows: negative n");
    WFB_REQUIRE(time_window_ns >= 0, "time_window_ns must be >= 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (n_events_dev) WFB_CUDA(cudaMemsetAsync(n_events_dev, 0, 8, st));
        return WFB_OK;
    }
    WFB_REQUIRE(timestamp_dev && position_dev && start_dev && end_dev && dt_dev && record_id_dev && order_dev && event_id_dev &&
                    abs_start_dev && abs_end_dev && n_events_dev && workspace_dev,
                "wfb_group_hit_windows: NULL pointer");
    WFB_REQUIRE(workspace_bytes >= wfb_group_workspace_bytes(n), "wfb_group_hit_windows: workspace too small");
    k4_abs_windows_kernel<<<nb(n), 256, 0, st>>>(reinterpret_cast<const long long*>(timestamp_dev), reinterpret_cast<const long long*>(position_dev),
                                                 start_dev, end_dev, dt_dev, n, abs_start_dev, abs_end_dev);
    return group_from_abs_windows(timestamp_dev, dt_dev, record_id_dev, n, time_window_ns, order_dev, event_id_dev, abs_start_dev, abs_end_dev,
                                  n_events_dev, workspace_dev, st);
}

extern "C" int wfb_group_abs_windows(const int64_t* timestamp_dev, const double* abs_start_dev, const double* abs_end_dev,
                                     const int32_t* dt_dev, const int64_t* record_id_dev, int64_t n, double time_window_ns,
                                     int64_t* order_dev, int64_t* event_id_dev, int64_t* n_events_dev, void* workspace_dev,
                                     size_t workspace_bytes, void* stream) {
    WFB_REQUIRE(n >= 0, "wfb_group_abs_windows: negative n");
    WFB_REQUIRE(time_window_ns >= 0, "time_window_ns must be >= 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (n_events_dev) WFB_CUDA(cudaMemsetAsync(n_events_dev, 0, 8, st));
        return WFB_OK;
    }
    WFB_REQUIRE(timestamp_dev && abs_start_dev && abs_end_dev && dt_dev && record_id_dev && order_dev && event_id_dev && n_events_dev && workspace_dev,
                "wfb_group_abs_windows: NULL pointer");
    WFB_REQUIRE(workspace_bytes >= wfb_group_workspace_bytes(n), "wfb_group_abs_windows: workspace too small");
    return group_from_abs_windows(timestamp_dev, dt_dev, record_id_dev, n, time_window_ns, order_dev, event_id_dev, abs_start_dev, abs_end_dev,
                                  n_events_dev, workspace_dev, st);
}

extern "C" int wfb_group_time_window(const int64_t* ts_sorted_dev, int64_t n, double time_window_ns, int64_t* event_id_dev,
                                     int64_t* n_events_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    WFB_REQUIRE(n >= 0, "wfb_group_time_window: negative n");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (n_events_dev) WFB_CUDA(cudaMemsetAsync(n_events_dev, 0, 8, st));
        return WFB_OK;
    }
    WFB_REQUIRE(ts_sorted_dev && event_id_dev && n_events_dev && workspace_dev, "wfb_group_time_window: NULL pointer");
    WFB_REQUIRE(workspace_bytes >= wfb_group_workspace_bytes(n), "wfb_group_time_window: workspace too small");
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    const size_t m = al256((size_t)n * 8);
    long long* seg = reinterpret_cast<long long*>(ws);
    long long* anchor = reinterpret_cast<long long*>(ws + m);
    long long* incl = reinterpret_cast<long long*>(ws + 2 * m);
    void* scan_ws = ws + 6 * m + radix_sort_workspace_bytes(n);
    const long long* ts = reinterpret_cast<const long long*>(ts_sorted_dev);
    const double w_ps = time_window_ns * 1e3;
    k4_segment_flags_kernel<<<nb(n), 256, 0, st>>>(ts, n, w_ps, seg);
    WFB_CUDA(cudaMemsetAsync(anchor, 0, (size_t)n * 8, st));
    k4_anchor_walk_kernel<<<nb(n), 256, 0, st>>>(ts, n, w_ps, seg, anchor);
    int rc = inclusive_scan_sum_i64(anchor, incl, n, scan_ws, st);
    if (rc != WFB_OK) return rc;
    k4_scatter_ids_kernel<<<nb(n), 256, 0, st>>>(incl, nullptr, n, reinterpret_cast<long long*>(event_id_dev),
                                                 reinterpret_cast<long long*>(n_events_dev));
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

extern "C" int wfb_structure_waveforms(const int16_t* samples_dev, const int64_t* ts_dev, const int16_t* board_dev, const int16_t* channel_dev,
                                       const double* baselines_in_dev, const double* baseline_upstream_dev, int64_t n, int32_t n_samples,
                                       int32_t wave_length, int32_t bl_start, int32_t bl_end, int32_t dt_ns, int64_t record_base,
                                       void* rows_out_dev, void* stream) {
    WFB_REQUIRE(n >= 0 && n_samples >= 0 && wave_length >= 0, "wfb_structure_waveforms: negative size");
    if (n == 0) return WFB_OK;
    WFB_REQUIRE(ts_dev && board_dev && channel_dev && rows_out_dev && (samples_dev || n_samples == 0), "wfb_structure_waveforms: NULL pointer");
    bl_start = std::max(bl_start, 0);
    bl_end = std::min(bl_end, n_samples);
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    st_structure_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(samples_dev, n, n_samples, wave_length,
                                                                              reinterpret_cast<const long long*>(ts_dev), board_dev, channel_dev,
                                                                              baselines_in_dev, baseline_upstream_dev, bl_start, bl_end, dt_ns,
                                                                              static_cast<uint8_t*>(rows_out_dev), record_base);
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}
