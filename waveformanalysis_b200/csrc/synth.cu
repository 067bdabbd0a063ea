// Device-side seeded synthetic run generator (bench / large parity tests): SURVEY.md 8(d).
// One warp per record; counter-based hash RNG so any record can be regenerated independently.
#include "common.cuh"

namespace wfb {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

// halfword k (0..50) of the packed RECORDS_DTYPE row
__device__ __forceinline__ uint16_t record_halfword(int k, const wfb_rec_meta& m, long long time_ns) {
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
    if (k == 38) return 0;                 // trigger_type
    if (k < 41) return 0;                  // flags
    if (k < 45) return h64((unsigned long long)m.wave_offset, k - 41);
    if (k < 47) return (uint16_t)((unsigned)m.event_length >> (16 * (k - 45)));
    return h64((unsigned long long)time_ns, k - 47);
}

__global__ void __launch_bounds__(256) synth_kernel(uint16_t* __restrict__ pool, wfb_rec_meta* __restrict__ meta,
                                                   uint16_t* __restrict__ rows, long long n, int L, int n_channels,
                                                   int dt_ns, uint64_t seed, long long record_base) {
    const int lane = lane_id();
    const long long rec = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (rec >= n) return;
    const long long gid = record_base + rec;
    const uint64_t h0 = mix64(seed ^ (uint64_t)gid * 0x9e3779b97f4a7c15ull);
    const int ch = (int)(h0 % (uint64_t)n_channels);
    const long long tick = 1000ll * dt_ns;
    const long long ts = (gid * 1500 + (long long)((h0 >> 20) % 1000)) * tick + ch;
    const float base = 8000.f + 10.f * ch;
    // up to three pulses
    const int np = 1 + (int)((h0 >> 8) % 3);
    float amp[3], st[3];
    bool flat[3];
    const int lo = L > 400 ? 100 : max(L / 8, 1);
    const int hi = L > 400 ? L - 200 : max(L / 2, lo + 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint64_t hk = mix64(h0 + 0x1234567ull * (k + 1));
        amp[k] = (k < np) ? 20.f + 480.f * u01(hk) : 0.f;
        st[k] = (float)(lo + (int)((hk >> 3) % (uint64_t)(hi - lo)));
        flat[k] = ((hk >> 50) % 10) == 0;
    }
    uint16_t* w = pool + rec * (long long)L;
    long long bsum = 0;
    const int nb = min(40, L);
    for (int j0 = lane * 8; j0 < L; j0 += 256) {
        unsigned v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int j = j0 + k;
            uint64_t hs = mix64(h0 ^ ((uint64_t)(j + 1) * 0xd6e8feb86659fd93ull));
            int s4 = (int)(hs & 255) + (int)((hs >> 8) & 255) + (int)((hs >> 16) & 255) + (int)((hs >> 24) & 255);
            float x = base + (float)(s4 - 510) * (3.0f / 147.8f);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float t = (float)j - st[q];
                if (t >= 0.f && amp[q] > 0.f) {
                    float shape = flat[q] ? (t < 200.f ? 1.f : 0.f) : 1.35f * (1.f - __expf(-t * 0.25f)) * __expf(-t * (1.f / 30.f));
                    x -= amp[q] * shape;
                }
            }
            int xi = __float2int_rn(x);
            xi = min(max(xi, 0), 16383);
            v[k] = (unsigned)xi;
            if (j < nb) bsum += xi;
        }
        if (j0 + 8 <= L && ((L & 7) == 0)) {
            uint4 q = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
            *reinterpret_cast<uint4*>(w + j0) = q;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (j0 + k < L) w[j0 + k] = (uint16_t)v[k];
        }
    }
    bsum = warp_sum_i64(bsum);
    wfb_rec_meta m;
    m.timestamp = ts;
    m.baseline = nb > 0 ? (double)bsum / (double)nb : __longlong_as_double(0x7ff8000000000000ll);
    m.wave_offset = gid * (long long)L;
    m.event_length = L;
    m.dt = dt_ns;
    m.board = 0;
    m.channel = (short)ch;
    m.polarity = WFB_POL_UNKNOWN;
    m.pad_[0] = m.pad_[1] = m.pad_[2] = 0;
    m.record_id = gid;
    if (meta != nullptr && lane == 0) meta[rec] = m;
    if (rows != nullptr) {
        uint16_t* row = rows + rec * (kRecordsRowBytes / 2);
        row[lane] = record_halfword(lane, m, ts / 1000);
        if (lane + 32 < 51) row[lane + 32] = record_halfword(lane + 32, m, ts / 1000);
    }
}

}  // namespace wfb

using namespace wfb;

extern "C" int wfb_synth_fill(uint16_t* pool_dev, wfb_rec_meta* meta_dev, void* records_aos_dev, int64_t n,
                              int32_t n_samples, int32_t n_channels, int32_t dt_ns, uint64_t seed, int64_t record_base,
                              void* stream) {
    WFB_REQUIRE(n >= 0 && n_samples > 0 && n_channels > 0 && dt_ns > 0, "wfb_synth_fill: bad sizes");
    if (n == 0) return WFB_OK;
    WFB_REQUIRE(pool_dev != nullptr, "wfb_synth_fill: pool_dev is NULL");
    WFB_REQUIRE(((uintptr_t)pool_dev & 15) == 0, "wfb_synth_fill: pool_dev must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long threads = n * 32;
    synth_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(pool_dev, meta_dev, static_cast<uint16_t*>(records_aos_dev), n,
                                                                    n_samples, n_channels, dt_ns, seed, record_base);
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}
