// wave_pool_filtered on the device: Savitzky-Golay (mode="interp") and zero-phase Butterworth
// (sosfiltfilt) per record, float32 output aligned with the input wave_offsets.
//
// Reference: core/plugins/builtin/cpu/filtering.py:181-241 (_apply_filter_core),
// :377-407 (filter_wave_pool_batch), records.py:368-438 (WavePoolFilteredPlugin.compute).
// Third-party arithmetic restated (scipy 1.18.1, un-vendored dependency of the reference):
//   savgol_filter(mode="interp"): interior = correlation with the savgol taps accumulated in
//     float64 and stored float32; first/last halflen samples = degree-`poly` least-squares fit over
//     the first/last `window` samples evaluated at those positions (projection rows from the host).
//   sosfiltfilt: odd extension by padlen, DF2T cascade forward from zi*x0, reversed pass from
//     zi*y_last, extension dropped; float64 with separate multiply/add (no FMA) = bit-exact.
#include <algorithm>

#include <vector>

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace wfb {

template <typename T>
__device__ __forceinline__ double sample_f64(const T* p, long long i) {
    return (double)(float)p[i];  // reference casts the pool to float32 first (filtering.py:169)
}

// Interior of the Savitzky-Golay correlation for 8 consecutive outputs of one lane (uint16 pool, record start
// 16-byte aligned): the 24 samples around the block are loaded as three 16-byte chunks and converted once,
// the taps sit in registers; the accumulation order (j ascending, multiply then add, no contraction) is the
// generic loop's, so the bits are the same.
template <int H>
__device__ __forceinline__ void sg_block8(const uint16_t* __restrict__ x, int k0, const double* __restrict__ taps, float* __restrict__ y) {
    constexpr int W = 2 * H + 1;
    const uint4* c = reinterpret_cast<const uint4*>(x + k0);
    const uint4 q0 = __ldg(c - 1), q1 = __ldg(c), q2 = __ldg(c + 1);
    const unsigned raw[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    double v[24];
#pragma unroll
    for (int i = 8 - H; i < 16 + H; ++i) v[i] = (double)(float)((raw[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
    double t[W];
#pragma unroll
    for (int j = 0; j < W; ++j) t[j] = taps[j];
    float out[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) acc = __dadd_rn(acc, __dmul_rn(t[j], v[8 + o - H + j]));
        out[o] = (float)acc;
    }
    *reinterpret_cast<float4*>(y + k0) = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4*>(y + k0 + 4) = make_float4(out[4], out[5], out[6], out[7]);
}

// ---- Savitzky-Golay: one warp per record -----------------------------------------------------
// table layout (doubles): [0, w) taps (y[k] = sum_j taps[j] * x[k-h+j]); [w, w + h*w) rows of the
// projector for outputs 0..h-1 over x[0..w); [w + h*w, w + 2*h*w) rows for outputs L-h..L-1 over
// x[L-w..L).  The effective window w is stored in front: table[-1] is not used; w comes from
// sg_window_eff[rec].
template <typename T>
__global__ void __launch_bounds__(256) sg_filter_kernel(const T* __restrict__ pool, long long pool_len,
                                                       const wfb_rec_meta* __restrict__ meta, long long n,
                                                       const double* __restrict__ tables,
                                                       const int* __restrict__ table_offset,
                                                       const int* __restrict__ cfg_index,
                                                       const wfb_filter_cfg* __restrict__ cfgs, float* __restrict__ out,
                                                       long long pool_base) {
    const int lane = lane_id();
    const long long rec = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (rec >= n) return;
    const wfb_filter_cfg& cfg = cfgs[cfg_index[rec]];
    if (cfg.type != WFB_FILTER_SG) return;
    const long long off = meta[rec].wave_offset - pool_base;
    const int L = meta[rec].event_length;
    if (L <= 0 || off < 0 || off + L > pool_len) return;
    const T* x = pool + off;
    float* y = out + off;
    const int toff = table_offset[rec];
    if (toff < 0) {  // window <= poly: identity (filtering.py:231-232)
        for (int k = lane; k < L; k += 32) y[k] = (float)x[k];
        return;
    }
    int w = min(cfg.sg_window, L);
    if ((w & 1) == 0) --w;
    const int h = w >> 1;
    const double* taps = tables + toff;
    const double* first = taps + w;
    const double* last = first + (size_t)h * w;
    // fast interior: blocks of 8 outputs that do not touch the polynomial-fit edges
    int fast_lo = 0, fast_hi = 0;  // outputs [fast_lo, fast_hi) are written by sg_block8
    if (sizeof(T) == 2 && h >= 1 && h <= 8 && (off & 7) == 0 && L >= 32) {
        fast_lo = 8 * ((h + 7) / 8);
        fast_hi = 8 * ((L - h) / 8);
        if (fast_hi <= fast_lo) fast_lo = fast_hi = 0;
        const uint16_t* xs = reinterpret_cast<const uint16_t*>(x);
        for (int k0 = fast_lo + 8 * lane; k0 < fast_hi; k0 += 256) {
            switch (h) {
                case 1: sg_block8<1>(xs, k0, taps, y); break;
                case 2: sg_block8<2>(xs, k0, taps, y); break;
                case 3: sg_block8<3>(xs, k0, taps, y); break;
                case 4: sg_block8<4>(xs, k0, taps, y); break;
                case 5: sg_block8<5>(xs, k0, taps, y); break;
                case 6: sg_block8<6>(xs, k0, taps, y); break;
                case 7: sg_block8<7>(xs, k0, taps, y); break;
                default: sg_block8<8>(xs, k0, taps, y); break;
            }
        }
    }
    for (int k = lane; k < L; k += 32) {
        if (k >= fast_lo && k < fast_hi) continue;
        double acc = 0.0;
        if (k >= h && k < L - h) {
            for (int j = 0; j < w; ++j) acc = __dadd_rn(acc, __dmul_rn(taps[j], sample_f64(x, k - h + j)));
        } else if (k < h) {
            const double* row = first + (size_t)k * w;
            for (int j = 0; j < w; ++j) acc = __dadd_rn(acc, __dmul_rn(row[j], sample_f64(x, j)));
        } else {
            const double* row = last + (size_t)(k - (L - h)) * w;
            for (int j = 0; j < w; ++j) acc = __dadd_rn(acc, __dmul_rn(row[j], sample_f64(x, L - w + j)));
        }
        y[k] = (float)acc;
    }
}

// ---- Butterworth sosfiltfilt: one thread per record, time-serial recursion --------------------
// scratch: float64 forward-pass output, time-major and interleaved over the threads of the grid
// (scratch[t * n_threads + tid]) so that the 32 lanes of a warp store consecutive doubles.
// FAST (the default): fused multiply-adds (5 instead of 9 float64 instructions per section and sample, half the
// dependent chain), the forward pass parked as float32, uint16 samples converted with the 2^52 trick instead of
// I2F.F64 - within 1e-9 relative of scipy's sosfiltfilt (the parity bar for floats is rel 1e-5), not bit-exact.
// WFB_BW_EXACT=1 selects the exact variant (separate multiplies and adds in scipy's order: bit-exact rows).
__device__ __forceinline__ double bw_sample(const uint16_t* p, long long i, bool fast) {
    if (fast) return __dsub_rn(__hiloint2double(0x43300000, (int)p[i]), 4503599627370496.0);
    return (double)(float)p[i];
}
__device__ __forceinline__ double bw_sample(const float* p, long long i, bool) { return (double)p[i]; }

// Fast Butterworth kernel: one thread per record like the exact one, but the time loop is split into its phases
// (odd-extension head, the record, odd-extension tail; backward: tail, record - the backward steps over the head only
// produce padding and are skipped), the samples come in 16-byte vectors, the forward pass is parked as float32 with a
// pointer that advances by the grid stride, and the cascade uses fused multiply-adds: ~28 instructions per sample and
// pass instead of ~70.
// UNI: every record uses ONE configuration with exactly MAXS sections (the normal case: no per-channel filter
// overrides).  The coefficients then are kernel parameters - the float64 instructions read them straight from the
// constant bank, which frees ~40 registers: six blocks per SM instead of three hide the latency of the serial cascade.
struct BwCoef {
    double b0[4], b1[4], b2[4], na1[4], na2[4], zi0[4], zi1[4];
    int edge;
};
template <typename T, int MAXS, bool UNI>
__global__ void __launch_bounds__(128, UNI ? 6 : 1) bw_fast_kernel(const T* __restrict__ pool, long long pool_len,
                                                     const wfb_rec_meta* __restrict__ meta, long long n,
                                                     const int* __restrict__ cfg_index,
                                                     const wfb_filter_cfg* __restrict__ cfgs, float* __restrict__ out,
                                                     long long pool_base, double* __restrict__ scratch_f64, int scratch_len,
                                                     const __grid_constant__ BwCoef cc) {
    float* __restrict__ scratch = reinterpret_cast<float*>(scratch_f64);
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long rec = tid; rec < n; rec += n_threads) {
        const wfb_filter_cfg& cfg = cfgs[UNI ? 0 : cfg_index[rec]];
        if (!UNI && (cfg.type != WFB_FILTER_BW || cfg.n_sections > MAXS)) continue;
        const long long off = meta[rec].wave_offset - pool_base;
        const int L = meta[rec].event_length;
        if (L <= 0 || off < 0 || off + L > pool_len) continue;
        const T* x = pool + off;
        float* y = out + off;
        const int ns = UNI ? MAXS : cfg.n_sections;
        int edge = cc.edge;
        if (!UNI) {
            int z_b = 0, z_a = 0;
            for (int s = 0; s < ns; ++s) {
                z_b += cfg.sos[s][2] == 0.0;
                z_a += cfg.sos[s][5] == 0.0;
            }
            edge = 3 * (2 * ns + 1 - min(z_b, z_a));  // filtering.py:198-203
        }
        if (L <= edge || L + 2 * edge > scratch_len) {       // unfiltered copy (filtering.py:221-222)
            for (int k = 0; k < L; ++k) y[k] = (float)x[k];
            continue;
        }
        double cb0[MAXS], cb1[MAXS], cb2[MAXS], na1[MAXS], na2[MAXS], z0[MAXS], z1[MAXS];
        if (!UNI) {
#pragma unroll
            for (int s = 0; s < MAXS; ++s) {
                const bool on = s < ns;
                cb0[s] = on ? cfg.sos[s][0] : 1.0; cb1[s] = on ? cfg.sos[s][1] : 0.0; cb2[s] = on ? cfg.sos[s][2] : 0.0;
                na1[s] = on ? -cfg.sos[s][4] : 0.0; na2[s] = on ? -cfg.sos[s][5] : 0.0;
            }
        }
        auto cascade = [&](double v) -> double {
#pragma unroll
            for (int s = 0; s < MAXS; ++s) {
                if constexpr (UNI) {
                    const double o = fma(cc.b0[s], v, z0[s]);
                    z0[s] = fma(cc.b1[s], v, fma(cc.na1[s], o, z1[s]));
                    z1[s] = fma(cc.b2[s], v, cc.na2[s] * o);
                    v = o;
                } else if (s < ns) {
                    const double o = fma(cb0[s], v, z0[s]);
                    z0[s] = fma(cb1[s], v, fma(na1[s], o, z1[s]));
                    z1[s] = fma(cb2[s], v, na2[s] * o);
                    v = o;
                }
            }
            return v;
        };
        auto smp = [&](int j) -> double { return bw_sample(x, j, true); };
        const double x_first = smp(0), x_last = smp(L - 1);
        const double x0 = fma(2.0, x_first, -smp(edge));
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            z0[s] = UNI ? cc.zi0[s] * x0 : (s < ns ? cfg.zi[s][0] * x0 : 0.0);
            z1[s] = UNI ? cc.zi1[s] * x0 : (s < ns ? cfg.zi[s][1] * x0 : 0.0);
        }
        float* sp = scratch + tid;
        double v = 0.0;
        // ---- forward: head (odd extension about the first sample), the record, tail
        for (int j = edge; j >= 1; --j) {
            v = cascade(fma(2.0, x_first, -smp(j)));
            *sp = (float)v;
            sp += n_threads;
        }
        int j = 0;
        if (sizeof(T) == 2 && (((uintptr_t)x) & 15) == 0 && L >= 16) {
            // 16 samples per pass from two 16-byte loads that were issued one pass earlier: the global-memory latency
            // overlaps the (serial) cascade of the previous 16 samples
            const uint4* xv = reinterpret_cast<const uint4*>(x);
            uint4 qa = __ldg(xv), qb = __ldg(xv + 1);
            for (; j + 16 <= L; j += 16) {
                const unsigned w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                if (j + 32 <= L) {
                    qa = __ldg(xv + ((j + 16) >> 3));
                    qb = __ldg(xv + ((j + 16) >> 3) + 1);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int s16 = (int)((w[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
                    v = cascade(__dsub_rn(__hiloint2double(0x43300000, s16), 4503599627370496.0));
                    *sp = (float)v;
                    sp += n_threads;
                }
            }
        }
        for (; j < L; ++j) {
            v = cascade(smp(j));
            *sp = (float)v;
            sp += n_threads;
        }
        for (int k = 0; k < edge; ++k) {
            v = cascade(fma(2.0, x_last, -smp(L - 2 - k)));
            *sp = (float)v;
            sp += n_threads;
        }
        // ---- backward over the parked forward pass: tail, then the record (the head would only produce padding)
        const double y0 = v;
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            z0[s] = UNI ? cc.zi0[s] * y0 : (s < ns ? cfg.zi[s][0] * y0 : 0.0);
            z1[s] = UNI ? cc.zi1[s] * y0 : (s < ns ? cfg.zi[s][1] * y0 : 0.0);
        }
        for (int k = 0; k < edge; ++k) {
            sp -= n_threads;
            v = cascade((double)*sp);
        }
        int o = L - 1;
        if (L >= 16) {  // 16 parked values per pass, loaded one pass ahead (see the forward loop)
            float nx[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) nx[k] = *(sp - (long long)(k + 1) * n_threads);
            for (; o >= 15; o -= 16) {
                float cu[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) cu[k] = nx[k];
                sp -= 16 * n_threads;
                if (o - 16 >= 15) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) nx[k] = *(sp - (long long)(k + 1) * n_threads);
                }
                float r[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) r[k] = (float)(v = cascade((double)cu[k]));
                if ((((uintptr_t)(y + o - 15)) & 15) == 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        *reinterpret_cast<float4*>(y + o - 15 + 4 * g) = make_float4(r[15 - 4 * g], r[14 - 4 * g], r[13 - 4 * g], r[12 - 4 * g]);
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) y[o - k] = r[k];
                }
            }
        }
        for (; o >= 0; --o) {
            sp -= n_threads;
            v = cascade((double)*sp);
            y[o] = (float)v;
        }
    }
}

template <typename T, int MAXS, bool FAST>
__global__ void __launch_bounds__(128) bw_filter_kernel(const T* __restrict__ pool, long long pool_len,
                                                       const wfb_rec_meta* __restrict__ meta, long long n,
                                                       const int* __restrict__ cfg_index,
                                                       const wfb_filter_cfg* __restrict__ cfgs, float* __restrict__ out,
                                                       long long pool_base, double* __restrict__ scratch_f64,
                                                       int scratch_len) {
    typedef typename std::conditional<FAST, float, double>::type scr_t;
    scr_t* __restrict__ scratch = reinterpret_cast<scr_t*>(scratch_f64);
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long rec = tid; rec < n; rec += n_threads) {
        const wfb_filter_cfg& cfg = cfgs[cfg_index[rec]];
        if (cfg.type != WFB_FILTER_BW || cfg.n_sections > MAXS) continue;
        const long long off = meta[rec].wave_offset - pool_base;
        const int L = meta[rec].event_length;
        if (L <= 0 || off < 0 || off + L > pool_len) continue;
        const T* x = pool + off;
        float* y = out + off;
        const int ns = cfg.n_sections;
        int z_b = 0, z_a = 0;
        for (int s = 0; s < ns; ++s) {
            z_b += cfg.sos[s][2] == 0.0;
            z_a += cfg.sos[s][5] == 0.0;
        }
        const int edge = 3 * (2 * ns + 1 - min(z_b, z_a));  // filtering.py:198-203
        if (L <= edge || L + 2 * edge > scratch_len) {       // unfiltered copy (filtering.py:221-222)
            for (int k = 0; k < L; ++k) y[k] = (float)x[k];
            continue;
        }
        const int n_ext = L + 2 * edge;
        const double x_first = bw_sample(x, 0, FAST), x_last = bw_sample(x, L - 1, FAST);
        auto ext = [&](int t) -> double {
            if (t < edge) return __dsub_rn(__dmul_rn(2.0, x_first), bw_sample(x, edge - t, FAST));
            if (t < edge + L) return bw_sample(x, t - edge, FAST);
            return __dsub_rn(__dmul_rn(2.0, x_last), bw_sample(x, L - 2 - (t - edge - L), FAST));
        };
        // coefficients and delay lines live in registers (sections unrolled to the compile-time maximum)
        double cb0[MAXS], cb1[MAXS], cb2[MAXS], ca1[MAXS], ca2[MAXS];
        double z0[MAXS], z1[MAXS];
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            const bool on = s < ns;
            cb0[s] = on ? cfg.sos[s][0] : 1.0; cb1[s] = on ? cfg.sos[s][1] : 0.0; cb2[s] = on ? cfg.sos[s][2] : 0.0;
            ca1[s] = on ? cfg.sos[s][4] : 0.0; ca2[s] = on ? cfg.sos[s][5] : 0.0;
        }
        auto cascade = [&](double v) -> double {
#pragma unroll
            for (int s = 0; s < MAXS; ++s) {
                if (s < ns) {
                    if (FAST) {
                        const double o = fma(cb0[s], v, z0[s]);
                        z0[s] = fma(cb1[s], v, fma(-ca1[s], o, z1[s]));
                        z1[s] = fma(cb2[s], v, -(ca2[s] * o));
                        v = o;
                    } else {
                        const double o = __dadd_rn(__dmul_rn(cb0[s], v), z0[s]);
                        z0[s] = __dadd_rn(__dsub_rn(__dmul_rn(cb1[s], v), __dmul_rn(ca1[s], o)), z1[s]);
                        z1[s] = __dsub_rn(__dmul_rn(cb2[s], v), __dmul_rn(ca2[s], o));
                        v = o;
                    }
                }
            }
            return v;
        };
        const double x0 = ext(0);
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            z0[s] = s < ns ? __dmul_rn(cfg.zi[s][0], x0) : 0.0;
            z1[s] = s < ns ? __dmul_rn(cfg.zi[s][1], x0) : 0.0;
        }
        // inputs do not depend on the filter state: they are fetched kPre steps ahead so that the global-memory
        // latency overlaps the (serial) cascade instead of stalling every step
        constexpr int kPre = 4;
        double pre[kPre];
#pragma unroll
        for (int k = 0; k < kPre; ++k) pre[k] = ext(min(k, n_ext - 1));
        double v = 0.0;
        for (int t0 = 0; t0 < n_ext; t0 += kPre) {
#pragma unroll
            for (int k = 0; k < kPre; ++k) {
                const int t = t0 + k;
                const double in = pre[k];
                pre[k] = ext(min(t + kPre, n_ext - 1));
                if (t < n_ext) {
                    v = cascade(in);
                    scratch[(size_t)t * n_threads + tid] = (scr_t)v;
                }
            }
        }
        const double y0 = v;  // last forward output
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            z0[s] = s < ns ? __dmul_rn(cfg.zi[s][0], y0) : 0.0;
            z1[s] = s < ns ? __dmul_rn(cfg.zi[s][1], y0) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < kPre; ++k) pre[k] = (double)scratch[(size_t)max(n_ext - 1 - k, 0) * n_threads + tid];
        for (int t0 = n_ext - 1; t0 >= 0; t0 -= kPre) {
#pragma unroll
            for (int k = 0; k < kPre; ++k) {
                const int t = t0 - k;
                const double in = pre[k];
                pre[k] = (double)scratch[(size_t)max(t - kPre, 0) * n_threads + tid];
                if (t >= 0) {
                    v = cascade(in);
                    if (t >= edge && t < edge + L) y[t - edge] = (float)v;
                }
            }
        }
    }
}

}  // namespace wfb

using namespace wfb;

extern "C" int wfb_filter_pool(const void* pool_dev, int32_t pool_is_f32, int64_t pool_len, const wfb_rec_meta* meta_dev,
                               int64_t n, const wfb_filter_cfg* cfgs_dev, int32_t n_cfg, const int32_t* cfg_index_dev,
                               const double* sg_tables_dev, const int32_t* sg_table_offset_dev, float* out_dev,
                               int64_t pool_base, void* workspace_dev, size_t workspace_bytes, int32_t lmax,
                               void* stream) {
    WFB_REQUIRE(n >= 0 && pool_len >= 0 && n_cfg >= 0, "wfb_filter_pool: negative size");
    if (n == 0) return WFB_OK;
    WFB_REQUIRE(pool_dev && meta_dev && cfgs_dev && cfg_index_dev && out_dev, "wfb_filter_pool: NULL pointer");
    WFB_REQUIRE(n_cfg > 0, "wfb_filter_pool: no filter configuration");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the output pool is zero where no record points (records.py:382)
    WFB_CUDA(cudaMemsetAsync(out_dev, 0, (size_t)pool_len * sizeof(float), st));
    const unsigned sg_blocks = (unsigned)((n * 32 + 255) / 256);
    // BW: as many threads as the scratch allows, at most one per record
    const long long scratch_len = (long long)lmax + 2 * 3 * (2 * WFB_MAX_SOS_SECTIONS + 1);
    long long bw_threads = 0;
    if (workspace_dev && workspace_bytes >= (size_t)scratch_len * 8 * 128) {
        bw_threads = (long long)(workspace_bytes / ((size_t)scratch_len * 8));
        bw_threads = std::min<long long>(bw_threads, (long long)sm_count() * 2048);
        bw_threads = std::min<long long>(bw_threads, ((n + 127) / 128) * 128);
        bw_threads = (bw_threads / 128) * 128;
    }
    // the Butterworth kernel keeps the section coefficients in registers: pick the smallest compiled size
    int max_sections = 0;
    wfb_filter_cfg uni_cfg;
    memset(&uni_cfg, 0, sizeof(uni_cfg));
    if (bw_threads > 0) {
        std::vector<wfb_filter_cfg> host_cfg((size_t)n_cfg);
        WFB_CUDA(cudaMemcpyAsync(host_cfg.data(), cfgs_dev, sizeof(wfb_filter_cfg) * (size_t)n_cfg, cudaMemcpyDeviceToHost, st));
        WFB_CUDA(cudaStreamSynchronize(st));
        for (const auto& c : host_cfg)
            if (c.type == WFB_FILTER_BW) max_sections = std::max(max_sections, (int)c.n_sections);
        uni_cfg = host_cfg[0];
        WFB_REQUIRE(max_sections <= WFB_MAX_SOS_SECTIONS, "wfb_filter_pool: too many second-order sections");
    }
    auto launch_bw = [&](auto tag) {
        typedef decltype(tag) T;
        const T* p = static_cast<const T*>(pool_dev);
        double* scr = static_cast<double*>(workspace_dev);
        const unsigned grid = (unsigned)(bw_threads / 128);
        const char* ex = getenv("WFB_BW_EXACT");
        const bool exact = ex && ex[0] == '1';
        BwCoef cc;
        memset(&cc, 0, sizeof(cc));
        // one configuration for every record (no per-channel overrides): coefficients as kernel parameters
        const char* nu = getenv("WFB_BW_UNIFORM");
        const bool uni = !exact && n_cfg == 1 && uni_cfg.type == WFB_FILTER_BW && (uni_cfg.n_sections == 2 || uni_cfg.n_sections == 4) &&
                         !(nu && nu[0] == '0');
        if (uni) {
            const int ns = uni_cfg.n_sections;
            int z_b = 0, z_a = 0;
            for (int s = 0; s < ns; ++s) {
                cc.b0[s] = uni_cfg.sos[s][0]; cc.b1[s] = uni_cfg.sos[s][1]; cc.b2[s] = uni_cfg.sos[s][2];
                cc.na1[s] = -uni_cfg.sos[s][4]; cc.na2[s] = -uni_cfg.sos[s][5];
                cc.zi0[s] = uni_cfg.zi[s][0]; cc.zi1[s] = uni_cfg.zi[s][1];
                z_b += uni_cfg.sos[s][2] == 0.0;
                z_a += uni_cfg.sos[s][5] == 0.0;
            }
            cc.edge = 3 * (2 * ns + 1 - std::min(z_b, z_a));  // filtering.py:198-203
            if (ns == 2)
                bw_fast_kernel<T, 2, true><<<grid, 128, 0, st>>>(p, pool_len, meta_dev, n, cfg_index_dev, cfgs_dev, out_dev, pool_base, scr, (int)scratch_len, cc);
            else
                bw_fast_kernel<T, 4, true><<<grid, 128, 0, st>>>(p, pool_len, meta_dev, n, cfg_index_dev, cfgs_dev, out_dev, pool_base, scr, (int)scratch_len, cc);
            return;
        }
#define WFB_BW_LAUNCH(MS)                                                                                                          \
    do {                                                                                                                           \
        if (exact)                                                                                                                 \
            bw_filter_kernel<T, MS, false><<<grid, 128, 0, st>>>(p, pool_len, meta_dev, n, cfg_index_dev, cfgs_dev, out_dev, pool_base, scr, \
                                                                (int)scratch_len);                                                 \
        else                                                                                                                       \
            bw_fast_kernel<T, MS, false><<<grid, 128, 0, st>>>(p, pool_len, meta_dev, n, cfg_index_dev, cfgs_dev, out_dev, pool_base, scr,       \
                                                       (int)scratch_len, cc);                                                      \
    } while (0)
        if (max_sections <= 4) WFB_BW_LAUNCH(4);
        else if (max_sections <= 8) WFB_BW_LAUNCH(8);
        else WFB_BW_LAUNCH(WFB_MAX_SOS_SECTIONS);
#undef WFB_BW_LAUNCH
    };
    if (pool_is_f32) {
        if (sg_tables_dev && sg_table_offset_dev)
            sg_filter_kernel<float><<<sg_blocks, 256, 0, st>>>(static_cast<const float*>(pool_dev), pool_len, meta_dev, n, sg_tables_dev,
                                                             sg_table_offset_dev, cfg_index_dev, cfgs_dev, out_dev, pool_base);
        if (bw_threads > 0 && max_sections > 0) launch_bw(float());
    } else {
        if (sg_tables_dev && sg_table_offset_dev)
            sg_filter_kernel<uint16_t><<<sg_blocks, 256, 0, st>>>(static_cast<const uint16_t*>(pool_dev), pool_len, meta_dev, n, sg_tables_dev,
                                                                sg_table_offset_dev, cfg_index_dev, cfgs_dev, out_dev, pool_base);
        if (bw_threads > 0 && max_sections > 0) launch_bw(uint16_t());
    }
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

extern "C" size_t wfb_filter_workspace_bytes(int64_t n, int32_t lmax) {
    // float64 forward-pass scratch of the Butterworth kernel, one row per resident thread
    const long long scratch_len = (long long)lmax + 2 * 3 * (2 * WFB_MAX_SOS_SECTIONS + 1);
    long long threads = std::min<long long>((long long)sm_count() * 2048, ((std::max<int64_t>(n, 1) + 127) / 128) * 128);
    return (size_t)scratch_len * 8 * (size_t)threads;
}
