// Shared pieces of the fused baseline / threshold-hit / basic_features kernels.
#pragma once
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace wfb {

struct HitEnt {  // a found hit, staged in shared memory until the tile's output offset is known
    int p, s;
    unsigned e_rec;  // end sample | owning lane << 27
    float height, integral;
};

struct FHArgs {
    const void* pool;
    long long pool_len;
    const wfb_rec_meta* meta;
    long long n;
    wfb_fh_params p;
    int lmax;
    int slot_bytes;  // shared-memory bytes per staged record (0: read samples from global memory)
    int ring_bytes;  // shared-memory bytes per warp (slot ring, reused as the row transposition buffer)
    uint8_t* feat_out;
    uint8_t* hit_out;
    long long hit_cap;
    int* hit_counts;
    const long long* hit_base;
    long long* total_out;
    unsigned long long* tile_state;  // [n_tiles] decoupled look-back descriptors
    unsigned long long* group_desc;  // [n_groups] per 32 tiles: finished tiles << 48 | sum of their hit counts
    unsigned long long* group_pref;  // [n_groups] status | exclusive prefix of the group's first tile
    uint4* gpool;                    // hit entries of the lane-per-record kernel: [resident warp][2][gpool_cap] (L2 resident)
    int gpool_cap;
    int n_slots;                     // warp-per-record kernel: slots of the per-warp TMA ring in use (2 or 3)
    int gpool_warps;                 // warps the pool area was sized for
    unsigned* ticket;                // dynamic tile counter
    int* err_flag;
    int n_tiles;
};

// what the scan of one record needs (warp-uniform, broadcast from the owning lane)
struct ScanRec {
    int mis, len, pol;
    double b_rec, b_feat, thr;
    int kmax;
    int p0, p1, c0, c1;
    // uint16 hits: floor(b_rec) / its int value / frac(b_rec), and the integer bound of the samples
    // on the signal side of the baseline (negative: w <= wlim, positive: w >= wlim)
    double bi, bf;
    int wlim;
    bool b_small;  // |b_rec| < 2e9: the integer split is usable
    int bias;      // 32768 when the 16-bit samples are int16: the kernel works on w' = w + 32768
    float xb;      // float32 pools: a sample is above threshold iff x <= xb (negative pulses) / x >= xb (positive); NaN: never
};
// what a hit row needs beyond the staged entry
struct RowRec {
    long long ts, rid;
    int len, dt;
    unsigned bc;  // board | channel << 16
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

// ---- mbarrier + TMA bulk copy (global -> shared) ---------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WFB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WFB_DONE;\n"
        "bra WFB_WAIT;\n"
        "WFB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the reference's float64 threshold test as an exact integer bound (uint16 pool) ---------
// negative/unknown polarity: hit iff fl(b - w) >= thr  <=>  w <= kmax   (monotone in w)
// positive polarity:         hit iff fl(w - b) >= thr  <=>  (65535 - w) <= kmax
__device__ __forceinline__ bool hit_test_u16(double b, double thr, bool positive, int w) {
    double sig = positive ? __dsub_rn((double)w, b) : __dsub_rn(b, (double)w);
    return sig >= thr;
}
static __device__ __noinline__ int integer_threshold_u16(double b, double thr, bool positive, int bias) {
    // largest k in [-1, 65535] such that every key kv <= k passes, kv = positive ? 65535 - w' : w',
    // w' = w + bias the stored (offset) sample and w the true sample value
    if (!(b == b) || !(thr == thr)) return -1;
    double guess = positive ? (65535.0 - (b + thr + bias)) : (b - thr + bias);
    int k = !(guess < 65535.0) ? 65535 : (guess < -1.0 ? -1 : (int)floor(guess));
    auto pass = [&](int kv) { return hit_test_u16(b, thr, positive, (positive ? 65535 - kv : kv) - bias); };
    for (int it = 0; it < 4 && k < 65535 && pass(k + 1); ++it) ++k;
    for (int it = 0; it < 4 && k >= 0 && !pass(k); ++it) --k;
    // the guess is within one step of the true bound; verify and fall back to a bisection if not
    if ((k >= 0 && !pass(k)) || (k < 65535 && pass(k + 1))) {
        int lo = -1, hi = 65535;
        while (lo < hi) {
            int mid = lo + (hi - lo + 1) / 2;
            if (pass(mid)) lo = mid; else hi = mid - 1;
        }
        k = lo;
    }
    return k;
}

// ---- packed THRESHOLD_HIT row (hit_finder.py:33-49, 382-409): 15 little-endian words ----------
__device__ __forceinline__ void hit_row_words(unsigned w[15], int p, int s, int e, float height, float integral,
                                              const RowRec& r, int left, int right, int lmax) {
    int a0 = max(0, s - left);
    int a1 = min(lmax, e + right);
    int rl = max(r.len, 0);
    int es = min(max(a0, 0), rl);
    int ee = max(min(max(a1, 0), rl), es);
    // int(timestamp + pos * (dt * 1e3)) evaluated in float64, no FMA contraction
    double step = __dmul_rn((double)r.dt, 1e3);
    long long ti = (long long)__dadd_rn((double)r.ts, __dmul_rn((double)p, step));
    w[0] = (unsigned)p;
    w[1] = 0u;  // position < 2^31
    w[2] = __float_as_uint(height);
    w[3] = __float_as_uint(integral);
    w[4] = (unsigned)es;
    w[5] = (unsigned)ee;
    w[6] = __float_as_uint((float)(ee - es));
    w[7] = (unsigned)r.dt;
    w[8] = __float_as_uint((float)((long long)max(p - s, 0) * r.dt));
    w[9] = __float_as_uint((float)((long long)max((e - 1) - p, 0) * r.dt));
    w[10] = (unsigned)(ti & 0xffffffffll);
    w[11] = (unsigned)((unsigned long long)ti >> 32);
    w[12] = r.bc;
    w[13] = (unsigned)(r.rid & 0xffffffffll);
    w[14] = (unsigned)((unsigned long long)r.rid >> 32);
}


// Two-level decoupled look-back: returns the exclusive prefix (first output row) of `tile`; called by one
// full warp.  `total` is the tile's own hit count.  With a persistent grid every resident block reaches this
// point at about the same time, so a plain look-back walks over all ~SMs x blocks unresolved predecessors,
// 32 per L2 round trip.  Here the 32 tiles of a GROUP also add their counts into one group descriptor: a tile
// first looks at its in-group predecessors (one probe); if none of them holds a prefix yet it takes the
// group's exclusive prefix from the 32 preceding group descriptors (one more probe reaches 1024 tiles back).
// publish the tile's own hit count (aggregate): never waits
__device__ __forceinline__ void tile_publish(const FHArgs& a, int tile, long long total) {
    if (lane_id() == 0) {
        st_state(a.tile_state + tile, kStAgg | (unsigned long long)total);
        atomicAdd(a.group_desc + (tile >> 5), (1ull << 48) | (unsigned long long)total);
    }
}

// resolve the exclusive prefix of a tile that has been published; publishes its inclusive prefix
__device__ __forceinline__ long long tile_resolve(const FHArgs& a, int tile, long long total) {
    const int lane = lane_id();
    const int g = tile >> 5, r = tile & 31;
    // a tile without hits has no rows to place: it does not wait for its prefix (the last tile still resolves
    // it, it reports the grand total)
    if (total == 0 && tile != a.n_tiles - 1) return 0;
    // ---- in-group predecessors r-1 .. 0 (lane L looks at tile - 1 - L)
    unsigned long long v = kStPrefix;  // lanes past the group start contribute nothing
    if (lane < r) {
        while (((v = ld_state(a.tile_state + tile - 1 - lane)) & kStMask) == 0) __nanosleep(64);  // do not steal issue slots
    }
    const unsigned isp = __ballot_sync(kFull, lane < r && (v & kStMask) == kStPrefix);
    const int first = __ffs(isp) - 1;  // nearest in-group predecessor holding an inclusive prefix
    long long excl = warp_sum_i64((lane < r && (first < 0 || lane <= first)) ? (long long)(v & ~kStMask) : 0);
    if (first < 0) {
        // ---- the group's exclusive prefix: nearest preceding group with a published prefix + complete groups between
        long long gex = a.hit_base ? *a.hit_base : 0;
        if (g > 0) {
            gex = 0;
            int look = g - 1;
            for (;;) {
                const int h = look - lane;
                unsigned long long pv = 0, dv = 0;
                bool have_p = false;
                if (h >= 0) {
                    const unsigned long long want = 32;  // every group in front of `g` is full
                    while (((dv = ld_state(a.group_desc + h)) >> 48) < want) {
                        pv = ld_state(a.group_pref + h);
                        __nanosleep(64);
                    }
                    pv = ld_state(a.group_pref + h);
                    have_p = (pv & kStMask) == kStPrefix;
                } else if (h == -1) {
                    have_p = true;  // in front of group 0: the caller's base
                    pv = kStPrefix | (unsigned long long)(a.hit_base ? *a.hit_base : 0);
                }
                const unsigned gp = __ballot_sync(kFull, have_p);
                const int gfirst = __ffs(gp) - 1;
                long long val = 0;
                if (h >= -1 && (gfirst < 0 || lane <= gfirst)) {
                    val = (long long)(dv & 0xffffffffffffull);                       // the group's own sum
                    if (lane == gfirst) val += (long long)(pv & ~kStMask);           // plus everything in front of it
                }
                gex += warp_sum_i64(val);
                if (gfirst >= 0) break;
                look -= 32;
            }
        }
        if (lane == 0) st_state(a.group_pref + g, kStPrefix | (unsigned long long)gex);
        excl += gex;
    }
    if (lane == 0) {
        st_state(a.tile_state + tile, kStPrefix | (unsigned long long)(excl + total));
        if (r == 0) st_state(a.group_pref + g, kStPrefix | (unsigned long long)excl);
        if (tile == a.n_tiles - 1) *a.total_out = excl + total;
    }
    return excl;
}

__device__ __forceinline__ long long tile_lookback(const FHArgs& a, int tile, long long total) {
    tile_publish(a, tile, total);
    return tile_resolve(a, tile, total);
}

// ---- float32 pools: the threshold as a bound on the raw sample ------------------------------
// hit_finder.py:329-340 compares sig = b - x (negative pulses) or x - b (positive), evaluated in float64, with the
// threshold.  Rounding is monotone, so the samples that pass are exactly those on one side of a float32 bound: the
// largest x with fl64(b - x) >= thr, or the smallest x with fl64(x - b) >= thr.  The bound is found once per record
// from the rounded estimate b -+ thr by stepping over neighbouring floats with the reference's own expression.
__device__ __forceinline__ float f32_step(float x, bool up) {
    if (x != x) return x;
    if (x == 0.f) return __uint_as_float(up ? 1u : 0x80000001u);
    unsigned u = __float_as_uint(x);
    const bool pos = (u >> 31) == 0u;
    if (pos == up) {
        if ((u & 0x7fffffffu) == 0x7f800000u) return x;  // +-inf stays
        ++u;
    } else {
        --u;
    }
    return __uint_as_float(u);
}
__device__ __forceinline__ float f32_threshold_bound(double b, double thr, bool positive) {
    const float qnan = __uint_as_float(0x7fc00000u);
    if (b != b || thr != thr) return qnan;
    auto ok = [&](float x) { return (positive ? __dsub_rn((double)x, b) : __dsub_rn(b, (double)x)) >= thr; };
    float cand = positive ? (float)__dadd_rn(b, thr) : (float)__dsub_rn(b, thr);
    if (cand != cand) return qnan;
    // towards the passing side until the candidate passes, then back while the neighbour still passes
    for (int it = 0; it < 4 && !ok(cand); ++it) cand = f32_step(cand, positive);
    if (!ok(cand)) return qnan;  // (infinite baselines and the like: no sample is reported)
    for (int it = 0; it < 4; ++it) {
        const float nb = f32_step(cand, !positive);
        if (nb == cand || !ok(nb)) break;
        cand = nb;
    }
    return cand;
}
// unsigned key with the order of the floats (-inf lowest); NaN is handled by the callers
__device__ __forceinline__ unsigned f32_order_key(float x) {
    const unsigned u = __float_as_uint(x);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_key(unsigned k) { return __uint_as_float((k >> 31) ? (k ^ 0x80000000u) : ~k); }

__device__ __forceinline__ long long bcast_i64(long long v, int src) {
    int lo = __shfl_sync(kFull, (int)(v & 0xffffffffll), src);
    int hi = __shfl_sync(kFull, (int)(v >> 32), src);
    return ((long long)hi << 32) | (unsigned)lo;
}

// lane-per-record variant (fused_lpr.cu): returns WFB_OK after launching, or 1 if the problem does
// not fit that kernel (caller falls back to the warp-per-record kernel)
int launch_lpr(FHArgs a, int flags, cudaStream_t st);

}  // namespace wfb
