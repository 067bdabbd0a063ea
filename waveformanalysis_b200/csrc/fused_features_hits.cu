// K2 + K3: fused baseline-subtract -> threshold hit finding -> basic_features, one warp per
// record, with single-pass stream compaction of the variable-length hit rows
// (decoupled look-back over tiles of 32 records, so rows come out in the reference's order:
// record-major, then start sample).
//
// Reference semantics restated here (paths relative to waveform_analysis/):
//   basic_features  core/plugins/builtin/cpu/basic_features.py:108-195 (records branch)
//   signals()       core/data/records_view.py:87-100
//   hit_threshold   core/plugins/builtin/cpu/hit_finder.py:122-255, 329-413
//
// uint16 pool: every per-sample operation is integer (min / max / sum / |diff| / compare with
// a per-record integer threshold that is exactly equivalent to the reference's float64
// `baseline - wave >= thr`); float64 is only used per record / per hit.
// float32 pool (wave_pool_filtered): per-sample float64, as the reference does.
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace wfb {

constexpr int kWarps = 8;                     // warps per block
constexpr int kRecPerWarp = 4;                // consecutive records handled by one warp per tile
constexpr int kTile = kWarps * kRecPerWarp;   // records per tile (one look-back unit)
constexpr int kEntPerWarp = 64;               // staged hits per warp per tile before re-scan

struct HitEnt {  // a found hit, staged in shared memory until the tile's output offset is known
    int p, s, e;
    float height, integral;
};

struct FHArgs {
    const void* pool;
    long long pool_len;
    const wfb_rec_meta* meta;
    long long n;
    wfb_fh_params p;
    int lmax;
    uint8_t* feat_out;
    uint8_t* hit_out;
    long long hit_cap;
    int* hit_counts;
    const long long* hit_base;
    long long* total_out;
    unsigned long long* tile_state;  // [n_tiles] decoupled look-back descriptors
    unsigned* ticket;                // dynamic tile counter
    int* err_flag;
    int n_tiles;
};

struct RecInfo {
    long long off, ts, rid;
    int len, dt, board, channel, pol;
    double b_rec, b_feat, thr;
};

// ---- tile descriptor: status in the top 2 bits, value below ---------------------------------
constexpr unsigned long long kStAgg = 1ull << 62, kStPrefix = 2ull << 62, kStMask = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- sample chunks: 8 consecutive samples per lane -------------------------------------------
template <typename T>
struct Chunk {
    T v[8];
};

__device__ __forceinline__ void load_chunk(const uint16_t* pool, long long pool_len, long long a, int lo, int hi,
                                           Chunk<uint16_t>& c) {
    if (hi <= lo) {
#pragma unroll
        for (int j = 0; j < 8; ++j) c.v[j] = 0;
        return;
    }
    if (a + 8 <= pool_len) {
        uint4 q = ldg_stream16(pool + a);
        c.v[0] = q.x & 0xffff; c.v[1] = q.x >> 16;
        c.v[2] = q.y & 0xffff; c.v[3] = q.y >> 16;
        c.v[4] = q.z & 0xffff; c.v[5] = q.z >> 16;
        c.v[6] = q.w & 0xffff; c.v[7] = q.w >> 16;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) c.v[j] = (a + j < pool_len) ? pool[a + j] : (uint16_t)0;
    }
}
__device__ __forceinline__ void load_chunk(const float* pool, long long pool_len, long long a, int lo, int hi,
                                           Chunk<float>& c) {
    if (hi <= lo) {
#pragma unroll
        for (int j = 0; j < 8; ++j) c.v[j] = 0.f;
        return;
    }
    if (a + 8 <= pool_len) {
        uint4 q0 = ldg_stream16(pool + a), q1 = ldg_stream16(pool + a + 4);
        c.v[0] = __uint_as_float(q0.x); c.v[1] = __uint_as_float(q0.y);
        c.v[2] = __uint_as_float(q0.z); c.v[3] = __uint_as_float(q0.w);
        c.v[4] = __uint_as_float(q1.x); c.v[5] = __uint_as_float(q1.y);
        c.v[6] = __uint_as_float(q1.z); c.v[7] = __uint_as_float(q1.w);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) c.v[j] = (a + j < pool_len) ? pool[a + j] : 0.f;
    }
}

// ---- the reference's float64 threshold test as an exact integer bound (uint16 pool) ---------
// negative/unknown polarity: hit iff fl(b - w) >= thr  <=>  w <= kmax   (monotone in w)
// positive polarity:         hit iff fl(w - b) >= thr  <=>  (65535 - w) <= kmax
__device__ __forceinline__ bool hit_test_u16(double b, double thr, bool positive, int w) {
    double sig = positive ? __dsub_rn((double)w, b) : __dsub_rn(b, (double)w);
    return sig >= thr;
}
__device__ int integer_threshold_u16(double b, double thr, bool positive) {
    // largest k in [-1, 65535] such that every key kv <= k passes, kv = positive ? 65535 - w : w
    if (!(b == b) || !(thr == thr)) return -1;
    double guess = positive ? (65535.0 - (b + thr)) : (b - thr);
    int k = guess >= 65535.0 ? 65535 : (guess < -1.0 ? -1 : (int)floor(guess));
    auto pass = [&](int kv) { return hit_test_u16(b, thr, positive, positive ? 65535 - kv : kv); };
    for (int it = 0; it < 4 && k < 65535 && pass(k + 1); ++it) ++k;
    for (int it = 0; it < 4 && k >= 0 && !pass(k); ++it) --k;
    // the guess is within one step of the true bound; verify and fall back to a bisection if not
    if ((k >= 0 && !pass(k)) || (k < 65535 && pass(k + 1))) {
        int lo = -1, hi = 65535;  // invariant: pass(lo) (or lo == -1), !pass(hi + 1)
        while (lo < hi) {
            int mid = lo + (hi - lo + 1) / 2;
            if (pass(mid)) lo = mid; else hi = mid - 1;
        }
        k = lo;
    }
    return k;
}

// ---- output rows ----------------------------------------------------------------------------
__device__ __forceinline__ void write_feature_row(uint8_t* out, long long row, const RecInfo& r, float height,
                                                  float amp, float area, float mad, long long event_index) {
    int lane = lane_id();
    if (lane < 9) {
        unsigned w;
        switch (lane) {
            case 0: w = __float_as_uint(height); break;
            case 1: w = __float_as_uint(amp); break;
            case 2: w = __float_as_uint(area); break;
            case 3: w = __float_as_uint(mad); break;
            case 4: w = (unsigned)(r.ts & 0xffffffffll); break;
            case 5: w = (unsigned)((unsigned long long)r.ts >> 32); break;
            case 6: w = ((unsigned)r.board & 0xffffu) | ((unsigned)r.channel << 16); break;
            case 7: w = (unsigned)(event_index & 0xffffffffll); break;
            default: w = (unsigned)((unsigned long long)event_index >> 32); break;
        }
        reinterpret_cast<unsigned*>(out + row * kFeatRowBytes)[lane] = w;
    }
}

// word `k` (0..14) of the packed 60-byte THRESHOLD_HIT row (hit_finder.py:33-49, 382-409)
__device__ __forceinline__ unsigned hit_row_word(int k, const HitEnt& h, const RecInfo& r, int left, int right,
                                                 int lmax) {
    int a0 = max(0, h.s - left);
    int a1 = min(lmax, h.e + right);
    int rl = max(r.len, 0);
    int es = min(max(a0, 0), rl);
    int ee = max(min(max(a1, 0), rl), es);
    switch (k) {
        case 0: return (unsigned)h.p;
        case 1: return 0u;  // position < 2^31
        case 2: return __float_as_uint(h.height);
        case 3: return __float_as_uint(h.integral);
        case 4: return (unsigned)es;
        case 5: return (unsigned)ee;
        case 6: return __float_as_uint((float)(ee - es));
        case 7: return (unsigned)r.dt;
        case 8: return __float_as_uint((float)((long long)max(h.p - h.s, 0) * r.dt));
        case 9: return __float_as_uint((float)((long long)max((h.e - 1) - h.p, 0) * r.dt));
        case 10:
        case 11: {
            // int(timestamp + pos * (dt * 1e3)) evaluated in float64, no FMA contraction
            double step = __dmul_rn((double)r.dt, 1e3);
            double t = __dadd_rn((double)r.ts, __dmul_rn((double)h.p, step));
            long long ti = (long long)t;
            return k == 10 ? (unsigned)(ti & 0xffffffffll) : (unsigned)((unsigned long long)ti >> 32);
        }
        case 12: return ((unsigned)r.board & 0xffffu) | ((unsigned)r.channel << 16);
        case 13: return (unsigned)(r.rid & 0xffffffffll);
        default: return (unsigned)((unsigned long long)r.rid >> 32);
    }
}

// ---- hit sinks --------------------------------------------------------------------------------
struct StageSink {  // phase A: count every hit, keep the rows that fit the warp's staging area
    HitEnt* buf;
    int room;   // entries still free in the warp buffer
    int n;      // hits of this record
    __device__ __forceinline__ void store(const HitEnt& h, const RecInfo&, const FHArgs&) {
        if (n < room && lane_id() == 0) buf[n] = h;
        ++n;
    }
};
struct RowSink {  // re-scan of a record whose hits did not fit: write rows straight to the output
    long long row;
    int n;
    __device__ __forceinline__ void store(const HitEnt& h, const RecInfo& r, const FHArgs& a) {
        int lane = lane_id();
        if (row < a.hit_cap && lane < 15)
            reinterpret_cast<unsigned*>(a.hit_out + row * kHitRowBytes)[lane] =
                hit_row_word(lane, h, r, a.p.left_extension, a.p.right_extension, a.lmax);
        ++row;
        ++n;
    }
};

// ---- per-hit segment reduction (hit_finder.py:369-381) --------------------------------------
template <typename T, typename Sink>
__device__ void emit_hit(const T* __restrict__ pool, const RecInfo& r, const FHArgs& a, int s, int e, Sink& sink) {
    const int lane = lane_id();
    const bool positive = r.pol == WFB_POL_POSITIVE;
    const int a0 = max(0, s - a.p.left_extension);
    const int a1 = min(a.lmax, e + a.p.right_extension);
    if (a1 <= a0) return;
    const double b = r.b_rec;
    double acc = 0.0;
    HitEnt h;
    h.s = s;
    h.e = e;
    if constexpr (sizeof(T) == 2) {
        int kbest = INT_MAX, ibest = INT_MAX;
        for (int i = a0 + lane; i < a1; i += 32) {
            int w = (i < r.len) ? (int)pool[r.off + i] : 0;  // padding samples are 0 (records_view.py:189)
            int kv = positive ? 65535 - w : w;
            if (kv < kbest) { kbest = kv; ibest = i; }
            double sig = positive ? __dsub_rn((double)w, b) : __dsub_rn(b, (double)w);
            acc += fmax(sig, 0.0);
        }
        int kmin = __reduce_min_sync(kFull, kbest);
        h.p = __reduce_min_sync(kFull, kbest == kmin ? ibest : INT_MAX);
        int wp = positive ? 65535 - kmin : kmin;
        h.height = (float)(positive ? __dsub_rn((double)wp, b) : __dsub_rn(b, (double)wp));
    } else {
        double sbest = -DBL_MAX;
        int ibest = INT_MAX;
        for (int i = a0 + lane; i < a1; i += 32) {
            double x = (i < r.len) ? (double)pool[r.off + i] : 0.0;
            double sig = positive ? __dsub_rn(x, b) : __dsub_rn(b, x);
            if (sig > sbest) { sbest = sig; ibest = i; }
            acc += fmax(sig, 0.0);
        }
        double smax = warp_max_f64(sbest);
        h.p = __reduce_min_sync(kFull, sbest == smax ? ibest : INT_MAX);
        h.height = (float)smax;
    }
    h.integral = (float)warp_sum_f64(acc);
    sink.store(h, r, a);
}

// ---- one record: features + threshold-run detection -----------------------------------------
struct FeatAcc {
    float height, amp, area, mad;
};

template <typename T, bool FEAT, bool HITS, typename Sink>
__device__ void scan_record(const T* __restrict__ pool, const RecInfo& r, const FHArgs& a, Sink& sink, FeatAcc& fa) {
    constexpr bool U16 = sizeof(T) == 2;
    constexpr int AL = U16 ? 8 : 4;  // samples per 16 bytes
    const int lane = lane_id();
    const int len = r.len;
    const int mis = (int)(r.off & (AL - 1));
    const long long abase = r.off - mis;
    const int vtotal = mis + max(len, 0);
    const bool positive = r.pol == WFB_POL_POSITIVE;
    const bool known = r.pol != WFB_POL_UNKNOWN;

    int p0 = 0, p1 = 0, c0 = 0, c1 = 0;
    if (FEAT) {
        resolve_slice(a.p.height_start, a.p.height_end, max(len, 0), p0, p1);
        resolve_slice(a.p.area_start, a.p.area_end, max(len, 0), c0, c1);
    }
    // hit threshold
    int kmax = -1;
    int xormask = 0;
    if (HITS && U16) {
        kmax = integer_threshold_u16(r.b_rec, r.thr, positive);
        xormask = positive ? 0xffff : 0;
    }
    const float b32 = (float)r.b_feat;

    // accumulators
    int imin = INT_MAX, imax = INT_MIN;            // u16: min/max over height range
    float fmin_ = FLT_MAX, fmax_ = -FLT_MAX;       // f32 pool (raw) or known-polarity signal
    unsigned long long isum = 0;                   // u16 unknown polarity: sum over area range
    double dsum = 0.0;                             // everything else
    int idiff = 0;
    double ddiff = 0.0;
    T carry_last = T(0);
    bool open = false;
    int run_start = 0;

    for (int vb = 0; vb < vtotal; vb += 256) {
        const int v0 = vb + lane * 8;
        const int lo = min(max(mis - v0, 0), 8);
        const int hi = min(max(vtotal - v0, 0), 8);
        const int i0 = v0 - mis;  // record index of c.v[0]
        Chunk<T> c;
        load_chunk(pool, a.pool_len, abase + v0, lo, hi, c);

        if (FEAT) {
            // ---- max |diff| over the whole record (basic_features.py:187-189)
            T prev = __shfl_up_sync(kFull, c.v[7], 1);
            if (lane == 0) prev = carry_last;
            carry_last = __shfl_sync(kFull, c.v[7], 31);
            if (hi > lo) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    T pj = (j == 0) ? prev : c.v[j - 1];
                    bool ok = (j >= lo) && (j < hi) && (i0 + j > 0);
                    if (U16) {
                        int d = abs((int)c.v[j] - (int)pj);
                        if (ok) idiff = max(idiff, d);
                    } else {
                        double d = fabs((double)c.v[j] - (double)pj);
                        if (ok) ddiff = fmax(ddiff, d);
                    }
                }
                // ---- height range [p0, p1)
                int jlo = max(lo, p0 - i0), jhi = min(hi, p1 - i0);
                if (jhi > jlo) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j >= jlo && j < jhi) {
                            if (U16) {
                                imin = min(imin, (int)c.v[j]);
                                imax = max(imax, (int)c.v[j]);
                            } else if (!known) {
                                fmin_ = fminf(fmin_, (float)c.v[j]);
                                fmax_ = fmaxf(fmax_, (float)c.v[j]);
                            } else {
                                // -signals(): negative -> b32 - x, positive -> x - b32 (float32)
                                float sv = positive ? __fsub_rn((float)c.v[j], b32) : __fsub_rn(b32, (float)c.v[j]);
                                fmin_ = fminf(fmin_, sv);
                                fmax_ = fmaxf(fmax_, sv);
                            }
                        }
                    }
                }
                // ---- area range [c0, c1)
                jlo = max(lo, c0 - i0);
                jhi = min(hi, c1 - i0);
                if (jhi > jlo) {
                    if (U16 && !known) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j >= jlo && j < jhi) isum += (unsigned)c.v[j];
                    } else if (!known) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j >= jlo && j < jhi) dsum += __dsub_rn(r.b_feat, (double)c.v[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j >= jlo && j < jhi) {
                                float sv = positive ? __fsub_rn((float)c.v[j], b32) : __fsub_rn(b32, (float)c.v[j]);
                                dsum += (double)sv;
                            }
                    }
                }
            }
        }

        if (HITS) {
            unsigned m8 = 0;
            if (hi > lo) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    bool bit;
                    if (U16) {
                        bit = (((int)c.v[j]) ^ xormask) <= kmax;
                    } else {
                        double x = (double)c.v[j];
                        double sig = positive ? __dsub_rn(x, r.b_rec) : __dsub_rn(r.b_rec, x);
                        bit = sig >= r.thr;
                    }
                    bit = bit && (j >= lo) && (j < hi);
                    m8 |= (bit ? 1u : 0u) << j;
                }
            }
            unsigned anyb = __ballot_sync(kFull, m8 != 0);
            if (anyb == 0 && !open) continue;
            // 32-bit mask word shared by each group of 4 lanes: samples [vb + 32q, vb + 32q + 32)
            const int g = lane & ~3;
            unsigned m32 = __shfl_sync(kFull, m8, g) | (__shfl_sync(kFull, m8, g + 1) << 8) |
                           (__shfl_sync(kFull, m8, g + 2) << 16) | (__shfl_sync(kFull, m8, g + 3) << 24);
            unsigned prevw = __shfl_up_sync(kFull, m32, 4);
            unsigned pbit = (lane < 4) ? (open ? 1u : 0u) : (prevw >> 31);
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
                    emit_hit<T>(pool, r, a, run_start, pos, sink);
                }
                if ((lane >> 2) == (ql >> 2)) tr &= tr - 1;
                tb = __ballot_sync(kFull, tr != 0 && (lane & 3) == 0);
            }
        }
    }
    if (HITS && open) emit_hit<T>(pool, r, a, run_start, max(len, 0), sink);

    if (FEAT) {
        fa.height = fa.amp = fa.area = fa.mad = 0.f;
        if (U16) {
            fa.mad = (float)__reduce_max_sync(kFull, idiff);
        } else {
            fa.mad = (float)warp_max_f64(ddiff);
        }
        if (p1 > p0) {
            if (U16) {
                int wmin = __reduce_min_sync(kFull, imin);
                int wmax = __reduce_max_sync(kFull, imax);
                if (!known) {
                    fa.height = (float)__dsub_rn(r.b_feat, (double)wmin);  // baseline - min(wave)
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
                    fa.height = (float)__dsub_rn(r.b_feat, (double)vmin);
                    fa.amp = (float)__dsub_rn((double)vmax, (double)vmin);
                } else {
                    fa.height = vmax;
                    fa.amp = (float)__dsub_rn((double)vmax, (double)vmin);
                }
            }
        }
        if (c1 > c0) {
            if (U16 && !known) {
                // sum(b - w) = n*floor(b) - sum(w) (exact integer) + n*frac(b)
                long long sw = warp_sum_i64((long long)isum);
                double b = r.b_feat;
                long long nC = c1 - c0;
                double area;
                if (fabs(b) < 1e12) {
                    double bi = floor(b);
                    double bf = __dsub_rn(b, bi);
                    long long ipart = nC * (long long)bi - sw;
                    area = __dadd_rn((double)ipart, __dmul_rn((double)nC, bf));
                } else {
                    area = __dsub_rn(__dmul_rn((double)nC, b), (double)sw);
                }
                fa.area = (float)area;
            } else {
                fa.area = (float)warp_sum_f64(dsum);
            }
        }
    }
}

// ---- per-channel rule lookup ------------------------------------------------------------------
__device__ __forceinline__ void apply_rules(const FHArgs& a, RecInfo& r) {
    r.thr = a.p.threshold;
    r.b_feat = r.b_rec;
    const wfb_chan_rule* rules = a.p.rules_dev;
    for (int i = 0; i < a.p.n_rules; ++i) {
        if (rules[i].board == r.board && rules[i].channel == r.channel) {
            if (rules[i].has_threshold) r.thr = rules[i].threshold;
            if (rules[i].has_fixed_baseline) r.b_feat = rules[i].fixed_baseline;
        }
    }
}

__device__ __forceinline__ bool load_rec(const FHArgs& a, long long rec, RecInfo& r) {
    const wfb_rec_meta* m = a.meta + rec;
    const uint4* q = reinterpret_cast<const uint4*>(m);
    uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    r.ts = (long long)(((unsigned long long)q0.y << 32) | q0.x);
    r.b_rec = __hiloint2double((int)q0.w, (int)q0.z);
    r.off = (long long)(((unsigned long long)q1.y << 32) | q1.x) - a.p.pool_base;
    r.len = (int)q1.z;
    r.dt = (int)q1.w;
    r.board = (int)(short)(q2.x & 0xffff);
    r.channel = (int)(short)(q2.x >> 16);
    r.pol = (int)(q2.y & 0xff);
    r.rid = (long long)(((unsigned long long)q2.w << 32) | q2.z);
    apply_rules(a, r);
    bool ok = r.len <= 0 || (r.off >= 0 && r.off + r.len <= a.pool_len);
    if (!ok) {
        if (lane_id() == 0) atomicExch(a.err_flag, 1);
        r.len = 0;
    }
    return ok;
}

// ---- the kernel --------------------------------------------------------------------------------
template <typename T, bool FEAT, bool HITS>
__global__ void __launch_bounds__(kWarps * 32) fused_features_hits_kernel(const FHArgs a) {
    __shared__ HitEnt s_ent[kWarps][kEntPerWarp];
    __shared__ int s_cnt[kTile];
    __shared__ long long s_off[kTile];  // absolute output row of each record's first hit
    __shared__ int s_tile;

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const T* pool = static_cast<const T*>(a.pool);

    for (;;) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= a.n_tiles) break;
        const long long rec0 = (long long)tile * kTile + warp * kRecPerWarp;

        // ---------------- phase A: scan my records, stage hits, write features
        int used = 0;              // staged entries of this warp
        int cnt[kRecPerWarp];
        int ent0[kRecPerWarp];     // first staged entry of record k, or -1 if it did not fit
#pragma unroll
        for (int k = 0; k < kRecPerWarp; ++k) {
            cnt[k] = 0;
            ent0[k] = 0;
            long long rec = rec0 + k;
            if (rec < a.n) {
                RecInfo r;
                load_rec(a, rec, r);
                StageSink sink{&s_ent[warp][used], kEntPerWarp - used, 0};
                FeatAcc fa;
                scan_record<T, FEAT, HITS, StageSink>(pool, r, a, sink, fa);
                if (FEAT) write_feature_row(a.feat_out, rec, r, fa.height, fa.amp, fa.area, fa.mad, a.p.row_base + rec);
                if (HITS) {
                    cnt[k] = sink.n;
                    if (sink.n <= kEntPerWarp - used) {
                        ent0[k] = used;
                        used += sink.n;
                    } else {
                        ent0[k] = -1;
                    }
                    if (a.hit_counts != nullptr && lane == 0) a.hit_counts[rec] = sink.n;
                }
            }
            if (HITS && lane == 0) s_cnt[warp * kRecPerWarp + k] = cnt[k];
        }
        if (!HITS) {
            __syncthreads();  // s_tile reuse
            continue;
        }
        __syncthreads();

        // ---------------- tile scan + decoupled look-back (warp 0)
        if (warp == 0) {
            int c = s_cnt[lane];  // kTile == 32
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            const long long total = __shfl_sync(kFull, incl, 31);
            if (lane == 0) st_state(a.tile_state + tile, kStAgg | (unsigned long long)total);
            long long excl = 0;
            int look = tile - 1;
            for (;;) {
                int idx = look - lane;
                unsigned long long v;
                if (idx >= 0) {
                    do { v = ld_state(a.tile_state + idx); } while ((v & kStMask) == 0);
                } else {
                    v = (idx == -1) ? (kStPrefix | (unsigned long long)(a.hit_base ? *a.hit_base : 0)) : kStPrefix;
                }
                unsigned isp = __ballot_sync(kFull, (v & kStMask) == kStPrefix);
                int first = __ffs(isp) - 1;  // nearest predecessor holding an inclusive prefix
                long long val = (first < 0 || lane <= first) ? (long long)(v & ~kStMask) : 0;
                excl += warp_sum_i64(val);
                if (first >= 0) break;
                look -= 32;
            }
            if (lane == 0) {
                st_state(a.tile_state + tile, kStPrefix | (unsigned long long)(excl + total));
                if (tile == a.n_tiles - 1) *a.total_out = excl + total;
            }
            s_off[lane] = excl + (incl - c);
        }
        __syncthreads();

        // ---------------- phase B: write my records' rows at their final position
#pragma unroll
        for (int k = 0; k < kRecPerWarp; ++k) {
            long long rec = rec0 + k;
            if (rec >= a.n || cnt[k] == 0) continue;
            const long long row0 = s_off[warp * kRecPerWarp + k];
            RecInfo r;
            load_rec(a, rec, r);
            if (ent0[k] >= 0) {
                const HitEnt* ents = &s_ent[warp][ent0[k]];
                for (int idx = lane; idx < cnt[k] * 15; idx += 32) {
                    int hrow = idx / 15, word = idx - hrow * 15;
                    long long row = row0 + hrow;
                    if (row < a.hit_cap)
                        reinterpret_cast<unsigned*>(a.hit_out + row * kHitRowBytes)[word] =
                            hit_row_word(word, ents[hrow], r, a.p.left_extension, a.p.right_extension, a.lmax);
                }
            } else {
                RowSink sink{row0, 0};
                FeatAcc fa;
                scan_record<T, false, true, RowSink>(pool, r, a, sink, fa);
            }
        }
        __syncthreads();  // staging buffers and s_tile are reused by the next tile
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
    size_t ticket, err, lmax, state, total;
};
static WsLayout ws_layout(long long n) {
    long long n_tiles = (n + kTile - 1) / kTile;
    WsLayout w;
    w.ticket = 0;
    w.err = 4;
    w.lmax = 8;
    w.state = 64;
    w.total = 64 + (size_t)n_tiles * 8;
    return w;
}

}  // namespace wfb

using namespace wfb;

extern "C" size_t wfb_features_hits_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return ws_layout(n).total + 64;
}

template <typename T>
static void launch_fused(const FHArgs& a, int flags, int grid, cudaStream_t st) {
    const bool f = flags & WFB_DO_FEATURES, h = flags & WFB_DO_HITS;
    if (f && h) fused_features_hits_kernel<T, true, true><<<grid, kWarps * 32, 0, st>>>(a);
    else if (f) fused_features_hits_kernel<T, true, false><<<grid, kWarps * 32, 0, st>>>(a);
    else fused_features_hits_kernel<T, false, true><<<grid, kWarps * 32, 0, st>>>(a);
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
    a.n_tiles = (int)((n + kTile - 1) / kTile);
    a.lmax = params->lmax;
    if ((flags & WFB_DO_HITS) && a.lmax <= 0) {
        // padded width = max event_length of the records passed (hit_finder.py:364)
        int* d_l = reinterpret_cast<int*>(ws + w.lmax);
        max_len_kernel<<<std::min<long long>(1024, (n + 255) / 256), 256, 0, st>>>(meta_dev, n, d_l);
        int h_l = 0;
        WFB_CUDA(cudaMemcpyAsync(&h_l, d_l, 4, cudaMemcpyDeviceToHost, st));
        WFB_CUDA(cudaStreamSynchronize(st));
        a.lmax = h_l;
    }
    int blocks_per_sm = 0;
    if (params->pool_is_f32)
        WFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fused_features_hits_kernel<float, true, true>, kWarps * 32, 0));
    else
        WFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fused_features_hits_kernel<uint16_t, true, true>, kWarps * 32, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int grid = (int)std::min<long long>((long long)sm_count() * blocks_per_sm, a.n_tiles);
    if (params->pool_is_f32) launch_fused<float>(a, flags, grid, st);
    else launch_fused<uint16_t>(a, flags, grid, st);
    WFB_CUDA(cudaGetLastError());
    return WFB_OK;
}

// error flag of the last wfb_features_hits call on this workspace (1 = a record pointed outside
// the pool).  Synchronises the stream.
extern "C" int wfb_features_hits_check(void* workspace_dev, void* stream) {
    int flag = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WFB_CUDA(cudaMemcpyAsync(&flag, static_cast<uint8_t*>(workspace_dev) + 4, 4, cudaMemcpyDeviceToHost, st));
    WFB_CUDA(cudaStreamSynchronize(st));
    if (flag) {
        set_error("records reference samples outside wave_pool bounds");
        return WFB_ERR_LAYOUT;
    }
    return WFB_OK;
}
