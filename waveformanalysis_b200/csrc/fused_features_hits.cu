// K2 + K3: fused baseline-subtract -> threshold hit finding -> basic_features with single-pass
// stream compaction of the variable-length hit rows.
//
// Reference semantics restated here (paths relative to waveform_analysis/):
//   basic_features  core/plugins/builtin/cpu/basic_features.py:108-195 (records branch)
//   signals()       core/data/records_view.py:87-100
//   hit_threshold   core/plugins/builtin/cpu/hit_finder.py:122-255, 329-413
//
// Work decomposition
//   * block = 8 warps = one tile of 256 consecutive records; warp w owns 32 of them, and LANE j
//     of the warp keeps the bookkeeping of record j (metadata, per-channel rule, integer
//     threshold, python-slice bounds, feature results, hit count) so all per-record scalar work
//     is done 32-wide.  The samples of one record are then scanned by the whole warp.
//   * samples reach the SM through a per-warp ring of shared-memory slots filled by 1-D TMA bulk
//     copies (cp.async.bulk + one mbarrier per slot), issued kSlots-1 records ahead, so the scan
//     reads LDS.128 chunks and the per-hit segment reductions never go back to L2/HBM.  Records
//     too long for a slot run the same code over a global-memory accessor.
//   * hits are staged per warp in shared memory; one decoupled look-back per 256-record tile
//     gives the tile's first output row, so rows come out in the reference's order
//     (record-major, then start sample); rows are assembled lane-parallel (one hit per lane),
//     transposed through shared memory and written coalesced.
//
// uint16 pool: every per-sample operation is integer (packed 16x2 min / max, IDP.2A sums,
// |diff| via max-min, compare against a per-record integer threshold that is exactly equivalent
// to the reference's float64 `baseline - wave >= thr`); float64 only per record / per hit.
// float32 pool (wave_pool_filtered): per-sample float64, as the reference does.
#include <float.h>
#include <limits.h>
#include <stdlib.h>

#include <algorithm>

#include "fused_common.cuh"

namespace wfb {

constexpr int kWarps = 8;                 // warps per block
constexpr int kTile = kWarps * 32;        // records per tile / look-back unit
constexpr int kEntPerWarp = 256;          // staged hits per warp per tile (8 per record on average)
constexpr int kSlots = 3;                 // TMA slot ring depth per warp
constexpr int kRowBufBytes = 32 * 17 * 4; // row transposition buffer per warp (32 rows x (15 + 2) words)

// ---- sample sources ---------------------------------------------------------------------------
// Both expose the record on a grid of 16-byte chunks: virtual index v = mis + i, where i is the
// sample index in the record and mis the number of samples between the 16-byte boundary below
// the record start and the record start.
template <typename T>
struct GlobalSrc {
    const T* base;     // pool + (off - mis): 16-byte aligned
    long long avail;   // samples readable from base
    __device__ __forceinline__ uint4 load16(int v) const {
        constexpr int PER = 16 / sizeof(T);
        if (v + PER <= avail) return ldg_stream16(base + v);
        T tmp[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) tmp[j] = (v + j < avail) ? base[v + j] : T(0);
        return *reinterpret_cast<uint4*>(tmp);
    }
    __device__ __forceinline__ T at(int v) const { return base[v]; }
};
template <typename T>
struct SmemSrc {
    const T* base;  // slot start (16-byte aligned shared memory)
    __device__ __forceinline__ uint4 load16(int v) const { return *reinterpret_cast<const uint4*>(base + v); }
    __device__ __forceinline__ T at(int v) const { return base[v]; }
};

// ---- hit sinks --------------------------------------------------------------------------------
struct StageSink {  // phase A: count every hit, keep the rows that fit the warp's staging area
    HitEnt* buf;    // next free entry of the warp
    int room;
    int n;
    unsigned owner;  // lane that owns the record
    __device__ __forceinline__ void store(int p, int s, int e, float height, float integral) {
        if (n < room && lane_id() == 0) {
            HitEnt h;
            h.p = p;
            h.s = s;
            h.e_rec = (unsigned)e | (owner << 27);
            h.height = height;
            h.integral = integral;
            buf[n] = h;
        }
        ++n;
    }
};
struct RowSink {  // re-scan of a record whose hits did not fit: write rows straight to the output
    uint8_t* out;
    long long row, cap;
    RowRec r;
    int left, right, lmax;
    int n;
    __device__ __forceinline__ void store(int p, int s, int e, float height, float integral) {
        if (lane_id() == 0 && row < cap) {
            unsigned w[15];
            hit_row_words(w, p, s, e, height, integral, r, left, right, lmax);
            unsigned* dst = reinterpret_cast<unsigned*>(out + row * kHitRowBytes);
#pragma unroll
            for (int k = 0; k < 15; ++k) dst[k] = w[k];
        }
        ++row;
        ++n;
    }
};

// ---- per-hit segment reduction (hit_finder.py:369-381) --------------------------------------
template <typename T, typename Src, typename Sink>
__device__ __forceinline__ void emit_hit(const Src& src, const ScanRec& r, const FHArgs& a, int s, int e, Sink& sink) {
    const int lane = lane_id();
    const bool positive = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_RAW_POSITIVE;
    const int a0 = max(0, s - a.p.left_extension);
    const int a1 = min(a.lmax, e + a.p.right_extension);
    if (a1 <= a0) return;
    const double b = r.b_rec;
    int hp;
    float hheight, hint;
    if constexpr (sizeof(T) == 2) {
        // integral = sum(max(sig, 0)) = +-(cnt*b - sum(w)) over the samples on the signal side of b
        const int wlim = r.wlim + r.bias;  // bound in the offset domain
        const int sx = r.bias ? 0x8000 : 0;
        const int seg = a1 - a0;
        int kmin, cnt_i;
        long long swt;
        if (seg <= 32) {
            // one sample per lane: argmin-first and (count, sum) each in a single warp reduction
            const int i = a0 + lane;
            const bool act = lane < seg;
            // padding samples are 0 (records_view.py:189), i.e. `bias` in the offset domain
            const int w = (act && i < r.len) ? ((int)src.at(r.mis + i) ^ sx) : r.bias;
            const int kv = positive ? 65535 - w : w;
            const unsigned key = act ? (((unsigned)kv << 16) | (unsigned)lane) : 0xffffffffu;
            const unsigned kred = __reduce_min_sync(kFull, key);
            kmin = (int)(kred >> 16);
            hp = a0 + (int)(kred & 0xffffu);
            const bool in = act && (positive ? (w >= wlim) : (w <= wlim));
            const unsigned packed = in ? ((1u << 26) | (unsigned)w) : 0u;
            const unsigned pred = __reduce_add_sync(kFull, packed);
            cnt_i = (int)(pred >> 26);
            swt = (long long)(pred & 0x3ffffffu);
        } else {
            int kbest = INT_MAX, ibest = INT_MAX;
            unsigned cnt = 0;
            unsigned long long sw = 0;
            for (int i = a0 + lane; i < a1; i += 32) {
                int w = (i < r.len) ? ((int)src.at(r.mis + i) ^ sx) : r.bias;
                int kv = positive ? 65535 - w : w;
                if (kv < kbest) { kbest = kv; ibest = i; }
                bool in = positive ? (w >= wlim) : (w <= wlim);
                cnt += in ? 1u : 0u;
                sw += in ? (unsigned)w : 0u;
            }
            kmin = __reduce_min_sync(kFull, kbest);
            hp = __reduce_min_sync(kFull, kbest == kmin ? ibest : INT_MAX);
            cnt_i = (int)__reduce_add_sync(kFull, cnt);
            swt = (seg <= 32768) ? (long long)__reduce_add_sync(kFull, (unsigned)sw) : warp_sum_i64((long long)sw);
        }
        const int wp = (positive ? 65535 - kmin : kmin) - r.bias;
        hheight = (float)(positive ? __dsub_rn((double)wp, b) : __dsub_rn(b, (double)wp));
        const long long c = cnt_i;
        swt -= c * r.bias;
        double integ;
        if (r.b_small) {
            long long ipart = positive ? (swt - c * (long long)r.bi) : (c * (long long)r.bi - swt);
            double fpart = __dmul_rn((double)c, r.bf);
            integ = positive ? __dsub_rn((double)ipart, fpart) : __dadd_rn((double)ipart, fpart);
        } else {
            integ = positive ? __dsub_rn((double)swt, __dmul_rn((double)c, b)) : __dsub_rn(__dmul_rn((double)c, b), (double)swt);
        }
        hint = (float)integ;
    } else {
        if (a1 - a0 <= 32) {
            // one sample per lane: the best sample through an integer min-reduction of the ordered float keys
            // (sig is monotone in x), its position through a ballot on sig itself (first of equal sig values)
            const int i = a0 + lane;
            const bool act = lane < a1 - a0;
            const float x = (act && i < r.len) ? (float)src.at(r.mis + i) : 0.f;
            const double sig = positive ? __dsub_rn((double)x, b) : __dsub_rn(b, (double)x);
            const unsigned key = (!act || x != x) ? 0xffffffffu : f32_order_key(positive ? -x : x);
            const unsigned kmin = __reduce_min_sync(kFull, key);
            double smax = -DBL_MAX;
            hp = INT_MAX;
            if (kmin != 0xffffffffu) {
                const float xm = f32_from_key(kmin);
                smax = positive ? __dsub_rn((double)(-xm), b) : __dsub_rn(b, (double)xm);
                hp = a0 + __ffs(__ballot_sync(kFull, act && sig == smax)) - 1;
            }
            hheight = (float)smax;
            hint = (float)warp_sum_f64(act ? fmax(sig, 0.0) : 0.0);
        } else {
            double sbest = -DBL_MAX, acc = 0.0;
            int ibest = INT_MAX;
            for (int i = a0 + lane; i < a1; i += 32) {
                double x = (i < r.len) ? (double)src.at(r.mis + i) : 0.0;
                double sig = positive ? __dsub_rn(x, b) : __dsub_rn(b, x);
                if (sig > sbest) { sbest = sig; ibest = i; }
                acc += fmax(sig, 0.0);
            }
            double smax = warp_max_f64(sbest);
            hp = __reduce_min_sync(kFull, sbest == smax ? ibest : INT_MAX);
            hheight = (float)smax;
            hint = (float)warp_sum_f64(acc);
        }
    }
    sink.store(hp, s, e, hheight, hint);
}

struct FeatAcc {
    float height, amp, area, mad;
};

__device__ __forceinline__ unsigned umin16x2(unsigned a, unsigned b) { return __vminu2(a, b); }
__device__ __forceinline__ unsigned umax16x2(unsigned a, unsigned b) { return __vmaxu2(a, b); }

// ---- one record: features + threshold-run detection -----------------------------------------
template <typename T, bool FEAT, bool HITS, typename Src, typename Sink>
__device__ __forceinline__ void scan_record(const Src& src, const ScanRec& r, const FHArgs& a, Sink& sink, FeatAcc& fa) {
    constexpr bool U16 = sizeof(T) == 2;
    const int lane = lane_id();
    const int len = r.len;
    const int mis = r.mis;
    const int vtotal = mis + len;
    const bool rawpos = r.pol == WFB_POL_RAW_POSITIVE;  // positive pulses, raw float64 arithmetic
    const bool positive = r.pol == WFB_POL_POSITIVE || rawpos;
    const bool known = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_NEGATIVE;  // float32 signal path
    const unsigned sx32 = (U16 && r.bias) ? 0x80008000u : 0u;
    const int p0 = r.p0, p1 = r.p1, c0 = r.c0, c1 = r.c1;
    const int kmax = r.kmax;
    const unsigned xm = (HITS && U16 && positive) ? 0xffffffffu : 0u;
    const float b32 = (float)r.b_feat;
    // u16 fast path accumulators (packed 16x2)
    unsigned pmin = 0xffffffffu, pmax = 0u, pdiff = 0u, isum32 = 0;
    // generic accumulators
    int imin = INT_MAX, imax = INT_MIN;
    float fmin_ = FLT_MAX, fmax_ = -FLT_MAX;
    unsigned long long isum = 0;
    double dsum = 0.0;
    int idiff = 0;
    double ddiff = 0.0;
    // float32 fast path: |diff| in float32 (what np.diff of a float32 array gives), raw sum of the samples in the area range
    float fdiff = 0.f;
    double xsum = 0.0;
    int nfast = 0;
    const float xb = r.xb;
    unsigned carry_w = 0;  // last 32-bit word of the previous window (u16: 2 samples, f32: 1 sample)
    bool open = false;
    int run_start = 0;

    for (int vb = 0; vb < vtotal; vb += 256) {
        const int v0 = vb + lane * 8;
        const int lo = min(max(mis - v0, 0), 8);
        const int hi = min(max(vtotal - v0, 0), 8);
        const int i0 = v0 - mis;  // record index of sample 0 of this lane's chunk
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
        if (hi > lo) {
            q0 = src.load16(v0);
            if (!U16) q1 = src.load16(v0 + 4);
            if (U16) { q0.x ^= sx32; q0.y ^= sx32; q0.z ^= sx32; q0.w ^= sx32; }  // int16 -> offset binary
        }
        // warp-uniform: every lane holds 8 valid samples that all lie inside the area range and
        // outside the height range (the common case away from the record edges)
        const int wi0 = vb - mis;
        const bool ultra = U16 && FEAT && !known && wi0 > 0 && vb + 256 <= vtotal && c0 <= wi0 && c1 >= wi0 + 256 &&
                           (p1 <= p0 || p1 <= wi0 || p0 >= wi0 + 256);
        const unsigned lastw = U16 ? q0.w : q1.w;
        unsigned prevw = __shfl_up_sync(kFull, lastw, 1);
        if (lane == 0) prevw = carry_w;
        carry_w = __shfl_sync(kFull, lastw, 31);
        unsigned m8 = 0;

        if (U16 && hi - lo == 8) {
            // ---------------- fast path: 8 valid uint16 samples in q0 ------------------------
            if (FEAT) {
                // |diff| against the previous sample, packed: max - min per halfword; the first
                // sample of the record has no predecessor (pair it with itself)
                unsigned pw = (i0 > 0) ? prevw : (q0.x << 16);
                unsigned f0 = __funnelshift_r(pw, q0.x, 16), f1 = __funnelshift_r(q0.x, q0.y, 16);
                unsigned f2 = __funnelshift_r(q0.y, q0.z, 16), f3 = __funnelshift_r(q0.z, q0.w, 16);
                unsigned d0 = umax16x2(q0.x, f0) - umin16x2(q0.x, f0), d1 = umax16x2(q0.y, f1) - umin16x2(q0.y, f1);
                unsigned d2 = umax16x2(q0.z, f2) - umin16x2(q0.z, f2), d3 = umax16x2(q0.w, f3) - umin16x2(q0.w, f3);
                pdiff = umax16x2(pdiff, umax16x2(umax16x2(d0, d1), umax16x2(d2, d3)));
                if (ultra) {
                    // whole window inside the area range and outside the height range: sum only
                    unsigned s = __dp2a_lo(q0.x, 0x0101u, isum32);
                    s = __dp2a_lo(q0.y, 0x0101u, s);
                    s = __dp2a_lo(q0.z, 0x0101u, s);
                    isum32 = __dp2a_lo(q0.w, 0x0101u, s);
                } else {
                // height range
                    const int jlo = max(0, p0 - i0), jhi = min(8, p1 - i0);
                    if (jhi > jlo) {
                        if (jlo == 0 && jhi == 8) {
                            pmin = umin16x2(pmin, umin16x2(umin16x2(q0.x, q0.y), umin16x2(q0.z, q0.w)));
                            pmax = umax16x2(pmax, umax16x2(umax16x2(q0.x, q0.y), umax16x2(q0.z, q0.w)));
                        } else {
                            const unsigned ww[4] = {q0.x, q0.y, q0.z, q0.w};
    #pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                int w = (int)((ww[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
                                if (j >= jlo && j < jhi) { imin = min(imin, w); imax = max(imax, w); }
                            }
                        }
                    }
                    // area range
                    const int klo = max(0, c0 - i0), khi = min(8, c1 - i0);
                    if (khi > klo) {
                        if (klo == 0 && khi == 8 && !known) {
                            unsigned s = __dp2a_lo(q0.x, 0x0101u, isum32);
                            s = __dp2a_lo(q0.y, 0x0101u, s);
                            s = __dp2a_lo(q0.z, 0x0101u, s);
                            isum32 = __dp2a_lo(q0.w, 0x0101u, s);
                        } else {
                            const unsigned ww[4] = {q0.x, q0.y, q0.z, q0.w};
    #pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                unsigned w = (ww[j >> 1] >> ((j & 1) * 16)) & 0xffffu;
                                if (j >= klo && j < khi) {
                                    if (!known) {
                                        isum += w;
                                    } else {
                                        float sv = positive ? __fsub_rn((float)w, b32) : __fsub_rn(b32, (float)w);
                                        dsum += (double)sv;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            if (HITS) {
                unsigned k0 = q0.x ^ xm, k1 = q0.y ^ xm, k2 = q0.z ^ xm, k3 = q0.w ^ xm;
                unsigned mn = umin16x2(umin16x2(k0, k1), umin16x2(k2, k3));
                int lmin = (int)min(mn & 0xffffu, mn >> 16);
                if (lmin <= kmax) {
                    const unsigned kk[4] = {k0, k1, k2, k3};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int kv = (int)((kk[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
                        m8 |= (kv <= kmax ? 1u : 0u) << j;
                    }
                }
            }
        } else if (!U16 && hi - lo == 8 && (!FEAT || (((p0 <= i0 && p1 >= i0 + 8) || p1 <= p0 || p1 <= i0 || p0 >= i0 + 8) &&
                                                      ((c0 <= i0 && c1 >= i0 + 8) || c1 <= c0 || c1 <= i0 || c0 >= i0 + 8)))) {
            // ---------------- float32 fast path: 8 valid samples, each range holds all or none of them ---------
            const float c[8] = {__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), __uint_as_float(q0.w),
                                __uint_as_float(q1.x), __uint_as_float(q1.y), __uint_as_float(q1.z), __uint_as_float(q1.w)};
            if (FEAT) {
                float d = (i0 > 0) ? fabsf(__fsub_rn(c[0], __uint_as_float(prevw))) : 0.f;
#pragma unroll
                for (int j = 1; j < 8; ++j) d = fmaxf(d, fabsf(__fsub_rn(c[j], c[j - 1])));
                fdiff = fmaxf(fdiff, d);
                if (p1 > p0 && p0 <= i0 && p1 >= i0 + 8) {
                    fmin_ = fminf(fmin_, fminf(fminf(fminf(c[0], c[1]), fminf(c[2], c[3])), fminf(fminf(c[4], c[5]), fminf(c[6], c[7]))));
                    fmax_ = fmaxf(fmax_, fmaxf(fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3])), fmaxf(fmaxf(c[4], c[5]), fmaxf(c[6], c[7]))));
                }
                if (c1 > c0 && c0 <= i0 && c1 >= i0 + 8) {
                    if (!known) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) xsum += (double)c[j];
                        nfast += 8;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) dsum += (double)(positive ? __fsub_rn(c[j], b32) : __fsub_rn(b32, c[j]));
                    }
                }
            }
            if (HITS) {
#pragma unroll
                for (int j = 0; j < 8; ++j) m8 |= ((positive ? (c[j] >= xb) : (c[j] <= xb)) ? 1u : 0u) << j;
            }
        } else if (hi > lo) {
            // ---------------- generic path: partial chunks (and float32 chunks on a range boundary) ---------------
            T c[8];
            if (U16) {
                const unsigned ww[4] = {q0.x, q0.y, q0.z, q0.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) c[j] = (T)((ww[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
            } else {
                const unsigned ww[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) c[j] = (T)__uint_as_float(ww[j]);
            }
            const T prev = U16 ? (T)(prevw >> 16) : (T)__uint_as_float(prevw);
            if (FEAT) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    T pj = (j == 0) ? prev : c[j - 1];
                    bool ok = (j >= lo) && (j < hi) && (i0 + j > 0);
                    if (U16) {
                        int d = abs((int)c[j] - (int)pj);
                        if (ok) idiff = max(idiff, d);
                    } else {
                        double d = fabs((double)c[j] - (double)pj);
                        if (ok) ddiff = fmax(ddiff, d);
                    }
                }
                int jlo = max(lo, p0 - i0), jhi = min(hi, p1 - i0);
                if (jhi > jlo) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j >= jlo && j < jhi) {
                            if (U16) {
                                imin = min(imin, (int)c[j]);
                                imax = max(imax, (int)c[j]);
                            } else {
                                // raw extremes; the float32 signal b32 - x / x - b32 is monotone in x and applied at the end
                                fmin_ = fminf(fmin_, (float)c[j]);
                                fmax_ = fmaxf(fmax_, (float)c[j]);
                            }
                        }
                    }
                }
                jlo = max(lo, c0 - i0);
                jhi = min(hi, c1 - i0);
                if (jhi > jlo) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j >= jlo && j < jhi) {
                            if (U16 && !known) {
                                isum += (unsigned)c[j];
                            } else if (!known) {
                                dsum += rawpos ? __dsub_rn((double)c[j], r.b_feat) : __dsub_rn(r.b_feat, (double)c[j]);
                            } else {
                                float sv = positive ? __fsub_rn((float)c[j], b32) : __fsub_rn(b32, (float)c[j]);
                                dsum += (double)sv;
                            }
                        }
                    }
                }
            }
            if (HITS) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    bool bit;
                    if (U16) {
                        bit = (int)(((unsigned)c[j]) ^ (xm & 0xffffu)) <= kmax;
                    } else {
                        double x = (double)c[j];
                        double sig = positive ? __dsub_rn(x, r.b_rec) : __dsub_rn(r.b_rec, x);
                        bit = sig >= r.thr;
                    }
                    bit = bit && (j >= lo) && (j < hi);
                    m8 |= (bit ? 1u : 0u) << j;
                }
            }
        }

        if (HITS) {
            unsigned anyb = __ballot_sync(kFull, m8 != 0);
            if (anyb == 0 && !open) continue;
            // 32-bit mask word shared by each group of 4 lanes: samples [vb + 32q, vb + 32q + 32)
            const int g = lane & ~3;
            unsigned m32 = __shfl_sync(kFull, m8, g) | (__shfl_sync(kFull, m8, g + 1) << 8) |
                           (__shfl_sync(kFull, m8, g + 2) << 16) | (__shfl_sync(kFull, m8, g + 3) << 24);
            unsigned prevm = __shfl_up_sync(kFull, m32, 4);
            unsigned pbit = (lane < 4) ? (open ? 1u : 0u) : (prevm >> 31);
            unsigned tr = m32 ^ ((m32 << 1) | pbit);  // set where the mask flips
            unsigned tb = __ballot_sync(kFull, tr != 0 && (lane & 3) == 0);
            while (tb) {
                int ql = __ffs(tb) - 1;
                unsigned t = __shfl_sync(kFull, tr, ql);
                int k = __ffs(t) - 1;
                int pos = vb + (ql >> 2) * 32 + k - mis;
                if (!open) {
                    open = true;
                    run_start = pos;
                } else {
                    open = false;
                    emit_hit<T>(src, r, a, run_start, pos, sink);
                }
                if ((lane >> 2) == (ql >> 2)) tr &= tr - 1;
                tb = __ballot_sync(kFull, tr != 0 && (lane & 3) == 0);
            }
        }
    }
    if (HITS && open) emit_hit<T>(src, r, a, run_start, len, sink);

    if (FEAT) {
        fa.height = fa.amp = fa.area = fa.mad = 0.f;
        if (U16) {
            int d = max(idiff, (int)max(pdiff & 0xffffu, pdiff >> 16));
            fa.mad = (float)__reduce_max_sync(kFull, d);
        } else {
            fa.mad = fmaxf((float)warp_max_f64(ddiff), warp_max_f32(fdiff));
        }
        if (p1 > p0) {
            if (U16) {
                imin = min(imin, (int)min(pmin & 0xffffu, pmin >> 16));
                imax = max(imax, (int)max(pmax & 0xffffu, pmax >> 16));
                int wmin = __reduce_min_sync(kFull, imin);
                int wmax = __reduce_max_sync(kFull, imax);
                if (!known) {
                    // baseline - min(wave), or max(wave) - baseline for raw positive pulses
                    fa.height = rawpos ? (float)__dsub_rn((double)(wmax - r.bias), r.b_feat)
                                       : (float)__dsub_rn(r.b_feat, (double)(wmin - r.bias));
                    fa.amp = (float)(wmax - wmin);
                } else {
                    float smax = positive ? __fsub_rn((float)wmax, b32) : __fsub_rn(b32, (float)wmin);
                    float smin = positive ? __fsub_rn((float)wmin, b32) : __fsub_rn(b32, (float)wmax);
                    fa.height = smax;
                    fa.amp = (float)__dsub_rn((double)smax, (double)smin);
                }
            } else {
                float vmin = warp_min_f32(fmin_), vmax = warp_max_f32(fmax_);
                if (!known) {
                    fa.height = rawpos ? (float)__dsub_rn((double)vmax, r.b_feat) : (float)__dsub_rn(r.b_feat, (double)vmin);
                    fa.amp = (float)__dsub_rn((double)vmax, (double)vmin);
                } else {
                    // -signals(): negative -> b32 - x, positive -> x - b32 (float32)
                    const float smax = positive ? __fsub_rn(vmax, b32) : __fsub_rn(b32, vmin);
                    const float smin = positive ? __fsub_rn(vmin, b32) : __fsub_rn(b32, vmax);
                    fa.height = smax;
                    fa.amp = (float)__dsub_rn((double)smax, (double)smin);
                }
            }
        }
        if (c1 > c0) {
            if (U16 && !known) {
                // sum(b - w) = n*floor(b) - sum(w) (exact integer) + n*frac(b)
                long long nC = c1 - c0;
                long long sw = (long long)__reduce_add_sync(kFull, isum32) + warp_sum_i64((long long)isum) - nC * r.bias;
                double b = r.b_feat;
                double area;
                if (fabs(b) < 1e12) {
                    double bi = floor(b);
                    double bf = __dsub_rn(b, bi);
                    double fpart = __dmul_rn((double)nC, bf);
                    if (rawpos) area = __dsub_rn((double)(sw - nC * (long long)bi), fpart);  // sum(w - b)
                    else area = __dadd_rn((double)(nC * (long long)bi - sw), fpart);        // sum(b - w)
                } else {
                    area = rawpos ? __dsub_rn((double)sw, __dmul_rn((double)nC, b)) : __dsub_rn(__dmul_rn((double)nC, b), (double)sw);
                }
                fa.area = (float)area;
            } else {
                if (!U16 && nfast) dsum += rawpos ? __dsub_rn(xsum, __dmul_rn((double)nfast, r.b_feat)) : __dsub_rn(__dmul_rn((double)nfast, r.b_feat), xsum);
                fa.area = (float)warp_sum_f64(dsum);
            }
        }
    }
}

__device__ __forceinline__ void hit_constants(ScanRec& r) {
    const double b = r.b_rec;
    const bool positive = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_RAW_POSITIVE;
    r.b_small = fabs(b) < 2e9;
    r.bi = floor(b);
    r.bf = __dsub_rn(b, r.bi);
    const int ib = r.b_small ? (int)r.bi : (b > 0 ? INT_MAX : INT_MIN);
    // negative: w < b  <=>  w <= ceil(b)-1 ; positive: w > b  <=>  w >= floor(b)+1
    r.wlim = positive ? ib + 1 : ((r.bi == b) ? ib - 1 : ib);
}

__device__ __forceinline__ double bcast_f64(double v, int src) { return shfl_f64(v, src); }

// ---- the kernel --------------------------------------------------------------------------------
template <typename T, bool STAGED, bool FEAT, bool HITS>
__global__ void __launch_bounds__(kWarps * 32) fused_features_hits_kernel(const FHArgs a) {
    constexpr bool U16 = sizeof(T) == 2;
    constexpr int AL = U16 ? 8 : 4;
    extern __shared__ __align__(16) uint8_t dyn_smem[];  // per warp: ring_bytes (slot ring / row buffer)
    __shared__ HitEnt s_ent[HITS ? kWarps : 1][HITS ? kEntPerWarp : 1];
    __shared__ long long s_wtot[kWarps];
    __shared__ long long s_wbase[kWarps];
    __shared__ int s_tile;
    __shared__ __align__(8) unsigned long long s_bar[kWarps * kSlots];

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const T* pool = static_cast<const T*>(a.pool);
    uint8_t* ring = dyn_smem + (size_t)warp * a.ring_bytes;
    unsigned long long* bars = &s_bar[warp * kSlots];
    const int nslots = a.n_slots;  // ring depth in use: 2 or 3 (warp-uniform)
    if (STAGED) {
        if (lane < kSlots) mbar_init(&bars[lane], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned phase_bits = 0;  // parity to wait for next, per slot

    for (;;) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= a.n_tiles) break;

        // ---------------- per-lane bookkeeping of record `lane` of this warp
        const long long rec = (long long)tile * kTile + warp * 32 + lane;
        const bool have = rec < a.n;
        long long off = 0, ts = 0, rid = 0;
        int len = 0, dt = 1, pol = 0, clen = -1;
        unsigned bc = 0;
        double b_rec = 0.0, b_feat = 0.0, thr = a.p.threshold;
        if (have) {
            const uint4* q = reinterpret_cast<const uint4*>(a.meta + rec);
            uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            ts = (long long)(((unsigned long long)q0.y << 32) | q0.x);
            b_rec = __hiloint2double((int)q0.w, (int)q0.z);
            off = (long long)(((unsigned long long)q1.y << 32) | q1.x) - a.p.pool_base;
            len = (int)q1.z;
            dt = (int)q1.w;
            bc = q2.x;
            pol = (int)(q2.y & 0xff);
            clen = (int)(q2.y >> 8) - 1;  // edge clamp of the hit rows (wfb_meta_set_clamp), -1: event_length
            rid = (long long)(((unsigned long long)q2.w << 32) | q2.z);
            b_feat = b_rec;
            const int board = (int)(short)(bc & 0xffff), channel = (int)(short)(bc >> 16);
            for (int i = 0; i < a.p.n_rules; ++i) {
                const wfb_chan_rule rule = a.p.rules_dev[i];
                if (rule.board == board && rule.channel == channel) {
                    if (rule.has_threshold) thr = rule.threshold;
                    if (rule.has_fixed_baseline) b_feat = rule.fixed_baseline;
                }
            }
            if (len < 0) len = 0;
            if (len > 0 && (off < 0 || off + len > a.pool_len)) {
                atomicExch(a.err_flag, 1);  // records reference samples outside the pool
                len = 0;
            }
        }
        const int mis = (int)(off & (AL - 1));
        int p0 = 0, p1 = 0, c0 = 0, c1 = 0;
        if (FEAT) {
            resolve_slice(a.p.height_start, a.p.height_end, len, p0, p1);
            resolve_slice(a.p.area_start, a.p.area_end, len, c0, c1);
        }
        int kmax = -1;
        const int bias = (U16 && a.p.signed_samples) ? 32768 : 0;
        if (HITS && U16 && len > 0)
            kmax = integer_threshold_u16(b_rec, thr, pol == WFB_POL_POSITIVE || pol == WFB_POL_RAW_POSITIVE, bias);
        float xb = 0.f;
        if (HITS && !U16 && len > 0) xb = f32_threshold_bound(b_rec, thr, pol == WFB_POL_POSITIVE || pol == WFB_POL_RAW_POSITIVE);
        unsigned copy_bytes = 0;
        if (STAGED && len > 0) {
            copy_bytes = (unsigned)((((long long)mis + len + AL - 1) & ~(long long)(AL - 1)) * (long long)sizeof(T));
            if ((int)copy_bytes > a.slot_bytes) {
                atomicExch(a.err_flag, 2);  // lmax passed by the caller is too small
                copy_bytes = 0;
                len = 0;
            }
        }
        auto issue = [&](int slot) {  // called by the lane that owns the record
            if (copy_bytes) {
                mbar_arrive_expect_tx(&bars[slot], copy_bytes);
                tma_bulk_g2s(ring + (size_t)slot * a.slot_bytes, pool + (off - mis), copy_bytes, &bars[slot]);
            } else {
                mbar_arrive(&bars[slot]);
            }
        };
        if (STAGED) {
            fence_proxy_async();  // generic-proxy accesses of the ring (previous tile) before the refill
            if (lane < nslots - 1) issue(lane);
        }

        // ---------------- phase A: the warp scans its 32 records one after the other
        FeatAcc my_fa = {0.f, 0.f, 0.f, 0.f};
        int my_cnt = 0, my_ent0 = 0;
        int used = 0;
        unsigned ovf_mask = 0;
        const int nrec = (int)max((long long)0, min((long long)32, a.n - ((long long)tile * kTile + warp * 32)));
        for (int j = 0; j < nrec; ++j) {
            ScanRec r;
            r.mis = __shfl_sync(kFull, mis, j);
            r.len = __shfl_sync(kFull, len, j);
            r.pol = __shfl_sync(kFull, pol, j);
            r.b_rec = bcast_f64(b_rec, j);
            r.b_feat = FEAT ? bcast_f64(b_feat, j) : 0.0;
            r.thr = (HITS && !U16) ? bcast_f64(thr, j) : 0.0;
            r.kmax = (HITS && U16) ? __shfl_sync(kFull, kmax, j) : -1;
            r.xb = (HITS && !U16) ? __shfl_sync(kFull, xb, j) : 0.f;
            r.p0 = FEAT ? __shfl_sync(kFull, p0, j) : 0;
            r.p1 = FEAT ? __shfl_sync(kFull, p1, j) : 0;
            r.c0 = FEAT ? __shfl_sync(kFull, c0, j) : 0;
            r.c1 = FEAT ? __shfl_sync(kFull, c1, j) : 0;
            r.bias = bias;
            if (HITS && U16) hit_constants(r);
            StageSink sink{HITS ? &s_ent[warp][used] : nullptr, kEntPerWarp - used, 0, (unsigned)j};
            FeatAcc fa = {0.f, 0.f, 0.f, 0.f};
            if (STAGED) {
                const int slot = j % nslots;
                // prefetch record j + nslots - 1 into the slot record j - 1 just left
                const int jn = j + nslots - 1;
                if (jn < 32) {
                    __syncwarp();
                    if (lane == jn) {
                        fence_proxy_async();
                        issue(jn % nslots);
                    }
                }
                mbar_wait(&bars[slot], (phase_bits >> slot) & 1u);
                phase_bits ^= 1u << slot;
                SmemSrc<T> src{reinterpret_cast<const T*>(ring + (size_t)slot * a.slot_bytes)};
                scan_record<T, FEAT, HITS>(src, r, a, sink, fa);
            } else {
                const long long offj = bcast_i64(off, j);
                GlobalSrc<T> src{pool + (offj - r.mis), a.pool_len - (offj - r.mis)};
                scan_record<T, FEAT, HITS>(src, r, a, sink, fa);
            }
            if (lane == j) {
                my_fa = fa;
                my_cnt = sink.n;
                my_ent0 = used;
            }
            if (HITS) {
                if (sink.n <= kEntPerWarp - used) used += sink.n;
                else ovf_mask |= 1u << j;
            }
        }
        if (STAGED) {
            // drain the barriers of slots armed for records >= nrec (tail tile) so phases stay in step
            for (int j = nrec; j < min(32, nrec + nslots - 1); ++j) {
                const int slot = j % nslots;
                mbar_wait(&bars[slot], (phase_bits >> slot) & 1u);
                phase_bits ^= 1u << slot;
            }
        }
        if (FEAT && have) {
            unsigned* dst = reinterpret_cast<unsigned*>(a.feat_out + rec * kFeatRowBytes);
            const long long ev = a.p.row_base + rec;
            dst[0] = __float_as_uint(my_fa.height);
            dst[1] = __float_as_uint(my_fa.amp);
            dst[2] = __float_as_uint(my_fa.area);
            dst[3] = __float_as_uint(my_fa.mad);
            dst[4] = (unsigned)(ts & 0xffffffffll);
            dst[5] = (unsigned)((unsigned long long)ts >> 32);
            dst[6] = bc;
            dst[7] = (unsigned)(ev & 0xffffffffll);
            dst[8] = (unsigned)((unsigned long long)ev >> 32);
        }
        if (!HITS) {
            __syncthreads();  // s_tile reuse
            continue;
        }
        if (a.hit_counts != nullptr && have) a.hit_counts[rec] = my_cnt;

        // ---------------- warp scan of the per-record counts, tile scan, decoupled look-back
        int incl = my_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long c = (lane < kWarps) ? s_wtot[lane] : 0;
            long long wincl = c;
#pragma unroll
            for (int d = 1; d < kWarps; d <<= 1) {
                long long t = bcast_i64(wincl, max(lane - d, 0));
                if (lane >= d) wincl += t;
            }
            const long long total = bcast_i64(wincl, kWarps - 1);
            const long long excl = tile_lookback(a, tile, total);
            if (lane < kWarps) s_wbase[lane] = excl + (wincl - c);
        }
        __syncthreads();

        // ---------------- phase B: one staged hit per lane -> packed rows -> coalesced stores
        const long long my_row0 = s_wbase[warp] + (incl - my_cnt);  // first output row of record `lane`
        unsigned* rowbuf = reinterpret_cast<unsigned*>(ring);
        for (int e0 = 0; e0 < used; e0 += 32) {
            const int e = e0 + lane;
            const bool act = e < used;
            HitEnt h = s_ent[warp][act ? e : 0];
            const int owner = (int)(h.e_rec >> 27);
            RowRec rr;
            rr.ts = bcast_i64(ts, owner);
            rr.rid = bcast_i64(rid, owner);
            rr.len = __shfl_sync(kFull, clen >= 0 ? clen : len, owner);
            rr.dt = __shfl_sync(kFull, dt, owner);
            rr.bc = __shfl_sync(kFull, bc, owner);
            const long long row = bcast_i64(my_row0, owner) + (e - __shfl_sync(kFull, my_ent0, owner));
            unsigned w[15];
            hit_row_words(w, h.p, h.s, (int)(h.e_rec & 0x7ffffffu), h.height, h.integral, rr, a.p.left_extension,
                          a.p.right_extension, a.lmax);
            __syncwarp();
            if (act) {
#pragma unroll
                for (int k = 0; k < 15; ++k) rowbuf[lane * 17 + k] = w[k];
                rowbuf[lane * 17 + 15] = (unsigned)(row & 0xffffffffll);
                rowbuf[lane * 17 + 16] = (unsigned)((unsigned long long)row >> 32);
            }
            __syncwarp();
            const int nrow = min(32, used - e0);
            for (int k = lane; k < nrow * 15; k += 32) {
                const int hr = k / 15, word = k - hr * 15;
                const long long grow = (long long)(((unsigned long long)rowbuf[hr * 17 + 16] << 32) | rowbuf[hr * 17 + 15]);
                if (grow < a.hit_cap) reinterpret_cast<unsigned*>(a.hit_out + grow * kHitRowBytes)[word] = rowbuf[hr * 17 + word];
            }
        }
        // records whose hits did not fit the staging area: scan them again, rows go straight out
        while (ovf_mask) {
            const int j = __ffs(ovf_mask) - 1;
            ovf_mask &= ovf_mask - 1;
            ScanRec r;
            r.mis = __shfl_sync(kFull, mis, j);
            r.len = __shfl_sync(kFull, len, j);
            r.pol = __shfl_sync(kFull, pol, j);
            r.b_rec = bcast_f64(b_rec, j);
            r.b_feat = 0.0;
            r.thr = bcast_f64(thr, j);
            r.kmax = __shfl_sync(kFull, kmax, j);
            r.xb = __shfl_sync(kFull, xb, j);
            r.p0 = r.p1 = r.c0 = r.c1 = 0;
            r.bias = bias;
            hit_constants(r);
            RowSink sink;
            sink.out = a.hit_out;
            sink.row = bcast_i64(my_row0, j);
            sink.cap = a.hit_cap;
            sink.r.ts = bcast_i64(ts, j);
            sink.r.rid = bcast_i64(rid, j);
            sink.r.len = __shfl_sync(kFull, clen >= 0 ? clen : len, j);
            sink.r.dt = __shfl_sync(kFull, dt, j);
            sink.r.bc = __shfl_sync(kFull, bc, j);
            sink.left = a.p.left_extension;
            sink.right = a.p.right_extension;
            sink.lmax = a.lmax;
            sink.n = 0;
            const long long offj = bcast_i64(off, j);
            GlobalSrc<T> src{pool + (offj - r.mis), a.pool_len - (offj - r.mis)};
            FeatAcc fa;
            scan_record<T, false, true>(src, r, a, sink, fa);
        }
        __syncthreads();  // staging buffers, ring and s_tile are reused by the next tile
    }
}

// ---- max event_length reduction (lmax = 0 in params) ----------------------------------------
__global__ void max_len_kernel(const wfb_rec_meta* meta, long long n, int* out) {
    int m = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = max(m, meta[i].event_length);
    m = __reduce_max_sync(kFull, m);
    if (lane_id() == 0) atomicMax(out, m);
}

struct WsLayout {
    size_t ticket, err, lmax, state, gdesc, gpref, total;  // total: end of the part that is cleared per call
    size_t gpool, end;
    int gpool_warps;
};
constexpr int kPoolCap = 512;        // hits per warp per tile kept for the deferred row pass (16 B each)
constexpr int kPoolWarpsMax = 2560;  // resident warps of the lane-per-record kernel the pool is sized for (16 per SM)
static WsLayout ws_layout(long long n) {
    long long n_tiles = (n + 31) / 32;  // smallest tile of the kernel variants (one warp in the lane-per-record kernel)
    WsLayout w;
    w.ticket = 0;
    w.err = 4;
    w.lmax = 8;
    const long long n_groups = (n_tiles + 31) / 32 + 1;
    w.state = 64;
    w.gdesc = w.state + (size_t)n_tiles * 8;
    w.gpref = w.gdesc + (size_t)n_groups * 8;
    w.total = w.gpref + (size_t)n_groups * 8;
    // hit pool of the lane-per-record kernel: [resident warp][2 halves][kPoolCap] entries, never cleared
    w.gpool = (w.total + 255) & ~(size_t)255;
    w.gpool_warps = (int)std::min<long long>(kPoolWarpsMax, (n + 31) / 32 + 8);
    w.end = w.gpool + (size_t)w.gpool_warps * 2 * kPoolCap * 16;
    return w;
}

template <typename T, bool STAGED>
static int launch_variant(FHArgs a, int flags, cudaStream_t st) {
    const bool f = flags & WFB_DO_FEATURES, h = flags & WFB_DO_HITS;
    const size_t dyn = (size_t)kWarps * a.ring_bytes;
    a.n_tiles = (int)((a.n + kTile - 1) / kTile);
    auto go = [&](auto kern) -> int {
        // static (staged hits) + dynamic (slot rings) shared memory exceeds the 48 KB default: opt in
        WFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        int per_sm = 0;
        WFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, dyn));
        if (per_sm < 1) per_sm = 1;
        int grid = (int)std::min<long long>((long long)sm_count() * per_sm, a.n_tiles);
        kern<<<grid, kWarps * 32, dyn, st>>>(a);
        WFB_CUDA(cudaGetLastError());
        return WFB_OK;
    };
    if (f && h) return go(fused_features_hits_kernel<T, STAGED, true, true>);
    if (f) return go(fused_features_hits_kernel<T, STAGED, true, false>);
    return go(fused_features_hits_kernel<T, STAGED, false, true>);
}

template <typename T>
static int launch_fused(FHArgs a, int flags, cudaStream_t st) {
    // shared-memory slot: the record rounded out to 16-byte boundaries on both sides
    const long long slot = (((long long)a.lmax * (long long)sizeof(T) + 15) & ~15ll) + 16;
    const char* force = getenv("WFB_FUSED_VARIANT");  // diagnostics: "global" | "staged"
    const bool want_global = force && !strcmp(force, "global");
    // three resident blocks per SM need <= ~64 KB each (ring + 20 KB of staged hits)
    if (!want_global && slot * 2 * kWarps <= 150 * 1024 && a.lmax < (1 << 27)) {
        // ring depth: three slots per warp unless two let more blocks share the SM (the staged hit pool of the
        // hit variants is ~41 KB of static shared memory; float32 records of 800 samples: 1 block with 3 slots, 2 with 2)
        const long long fixed = ((flags & WFB_DO_HITS) ? 42 : 1) * 1024 + 1024;
        auto blocks = [&](int ns) { return (227 * 1024) / (std::max<long long>(slot * ns, kRowBufBytes) * kWarps + fixed); };
        a.n_slots = (slot * kSlots * kWarps <= 150 * 1024 && blocks(kSlots) >= blocks(2)) ? kSlots : 2;
        if (const char* e = getenv("WFB_FUSED_SLOTS")) a.n_slots = (atoi(e) == 2) ? 2 : ((slot * kSlots * kWarps <= 150 * 1024) ? 3 : 2);
        a.slot_bytes = (int)slot;
        a.ring_bytes = (int)std::max<long long>(slot * a.n_slots, kRowBufBytes);
        return launch_variant<T, true>(a, flags, st);
    }
    a.slot_bytes = 0;
    a.n_slots = kSlots;
    a.ring_bytes = kRowBufBytes;
    return launch_variant<T, false>(a, flags, st);
}

}  // namespace wfb

using namespace wfb;

extern "C" size_t wfb_features_hits_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return ws_layout(n).end + 64;
}

extern "C" int wfb_features_hits(const void* pool_dev, int64_t pool_len, const wfb_rec_meta* meta_dev, int64_t n,
                                 const wfb_fh_params* params, void* feat_out_dev, void* hit_out_dev,
                                 int64_t hit_cap, int32_t* hit_counts_dev, const int64_t* hit_base_dev,
                                 int64_t* total_hits_dev, void* workspace_dev, size_t workspace_bytes,
                                 void* stream) {
    WFB_REQUIRE(params != nullptr, "wfb_features_hits: params is NULL");
    WFB_REQUIRE(n >= 0 && pool_len >= 0, "wfb_features_hits: negative size");
    const int flags = params->flags;
    WFB_REQUIRE((flags & (WFB_DO_FEATURES | WFB_DO_HITS)) != 0, "wfb_features_hits: flags select nothing");
    WFB_REQUIRE(!(flags & WFB_DO_FEATURES) || feat_out_dev != nullptr || n == 0, "wfb_features_hits: feat_out_dev is NULL");
    WFB_REQUIRE(!(flags & WFB_DO_HITS) || total_hits_dev != nullptr, "wfb_features_hits: total_hits_dev is NULL");
    WFB_REQUIRE(!(flags & WFB_DO_HITS) || hit_cap == 0 || hit_out_dev != nullptr, "wfb_features_hits: hit_out_dev is NULL");
    WFB_REQUIRE(params->left_extension >= 0 && params->right_extension >= 0, "wfb_features_hits: negative extension");
    WFB_REQUIRE(((uintptr_t)pool_dev & 15) == 0, "wfb_features_hits: pool_dev must be 16-byte aligned");
    WFB_REQUIRE(((uintptr_t)meta_dev & 15) == 0, "wfb_features_hits: meta_dev must be 16-byte aligned");
    WFB_REQUIRE(workspace_bytes >= wfb_features_hits_workspace_bytes(n), "wfb_features_hits: workspace too small");
    WFB_REQUIRE(((uintptr_t)workspace_dev & 15) == 0, "wfb_features_hits: workspace_dev must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (flags & WFB_DO_HITS) {
            if (hit_base_dev) WFB_CUDA(cudaMemcpyAsync(total_hits_dev, hit_base_dev, 8, cudaMemcpyDeviceToDevice, st));
            else WFB_CUDA(cudaMemsetAsync(total_hits_dev, 0, 8, st));
        }
        return WFB_OK;
    }
    WsLayout w = ws_layout(n);
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    WFB_CUDA(cudaMemsetAsync(ws, 0, w.total, st));
    FHArgs a;
    a.pool = pool_dev;
    a.pool_len = pool_len;
    a.meta = meta_dev;
    a.n = n;
    a.p = *params;
    a.feat_out = static_cast<uint8_t*>(feat_out_dev);
    a.hit_out = static_cast<uint8_t*>(hit_out_dev);
    a.hit_cap = hit_cap;
    a.hit_counts = hit_counts_dev;
    a.hit_base = reinterpret_cast<const long long*>(hit_base_dev);
    a.total_out = reinterpret_cast<long long*>(total_hits_dev);
    a.ticket = reinterpret_cast<unsigned*>(ws + w.ticket);
    a.err_flag = reinterpret_cast<int*>(ws + w.err);
    a.tile_state = reinterpret_cast<unsigned long long*>(ws + w.state);
    a.group_desc = reinterpret_cast<unsigned long long*>(ws + w.gdesc);
    a.group_pref = reinterpret_cast<unsigned long long*>(ws + w.gpref);
    a.gpool = reinterpret_cast<uint4*>(ws + w.gpool);
    a.gpool_cap = kPoolCap;
    a.gpool_warps = w.gpool_warps;
    a.n_tiles = 0;
    a.slot_bytes = 0;
    a.ring_bytes = 0;
    a.n_slots = kSlots;
    a.lmax = params->lmax;
    if (a.lmax <= 0) {
        // padded width = max event_length of the records passed (hit_finder.py:364); costs a sync
        int* d_l = reinterpret_cast<int*>(ws + w.lmax);
        max_len_kernel<<<(unsigned)std::min<long long>(1024, (n + 255) / 256), 256, 0, st>>>(meta_dev, n, d_l);
        int h_l = 0;
        WFB_CUDA(cudaMemcpyAsync(&h_l, d_l, 4, cudaMemcpyDeviceToHost, st));
        WFB_CUDA(cudaStreamSynchronize(st));
        a.lmax = h_l;
    }
    // lane-per-record kernel (16-bit pools: block items; float32 pools: register-resident runs) when the records fit
    // its shared-memory slots and the extensions its hit state
    const int rc = launch_lpr(a, flags, st);
    if (rc != 1) return rc;
    if (params->pool_is_f32) return launch_fused<float>(a, flags, st);
    return launch_fused<uint16_t>(a, flags, st);
}

// error flag of the last wfb_features_hits call on this workspace.  Synchronises the stream.
extern "C" int wfb_features_hits_check(void* workspace_dev, void* stream) {
    int flag = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WFB_CUDA(cudaMemcpyAsync(&flag, static_cast<uint8_t*>(workspace_dev) + 4, 4, cudaMemcpyDeviceToHost, st));
    WFB_CUDA(cudaStreamSynchronize(st));
    if (flag == 1) {
        set_error("records reference samples outside wave_pool bounds");
        return WFB_ERR_LAYOUT;
    }
    if (flag == 2) {
        set_error("lmax is smaller than the longest record passed");
        return WFB_ERR_INVALID;
    }
    return WFB_OK;
}
