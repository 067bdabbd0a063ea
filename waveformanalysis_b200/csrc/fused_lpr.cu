// K2 + K3, lane-per-record streaming variant (uint16 / int16 pools).
//
// Each LANE owns one record and streams it through a small double-buffered shared-memory slot:
// segments of kSC 16-byte chunks (8 samples each) arrive by 1-D TMA bulk copies (one copy per
// lane per segment, one mbarrier per warp buffer, issued one segment ahead), the slot stride is
// an odd multiple of 16 bytes so the 32 lanes' LDS.128 hit distinct bank groups.  The sample
// loop has no cross-lane traffic at all - no shuffles, no warp reductions - only packed 16x2
// integer ops on the lane's own registers.
//
// Hits: while scanning, a lane notes in a bitmask which chunks contain samples above threshold;
// after each segment a lane-parallel pass visits only those chunks (still in the slot), walks the
// threshold runs with bit tricks and accumulates each hit's argmax / integral on the fly; runs
// may stay open across segments.  The pass lags the scan by kOV chunks and each slot carries
// 2*kOV chunks of history, so the left / right extensions (<= 8*kOV samples) of any run it
// closes are always inside the slot.  Hits go to a per-warp shared-memory pool (claimed with a
// shared-memory atomic); one decoupled look-back per 128-record tile gives the first output
// row; rows are assembled one hit per lane.
//
// Reference semantics: see fused_features_hits.cu (same arithmetic, same results).
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "fused_common.cuh"

namespace wfb {

constexpr int kLprWarps = 4;
constexpr int kLprTile = kLprWarps * 32;  // records per tile / look-back unit
constexpr int kLprEnt = 320;              // staged hits per warp per tile (10 per record on average)
constexpr int kOV = 2;                    // chunks of lag / history: extensions up to 16 samples
constexpr int kNBuf = 2;                  // slot buffers per lane

struct LprEnt {  // 16 bytes
    unsigned ps;   // p | s << 16
    unsigned eo;   // e | owner lane << 16 | ordinal << 21
    float height, integral;
};

struct LaneRec {  // everything a lane knows about its record
    long long off, ts, rid;
    int len, dt, pol, mis, bias;
    unsigned bc;
    double b_rec, b_feat, thr;
    int kmax;
    double bi, bf;  // floor / frac of b_rec, integer bound of the signal side (hit integral)
    int wlim;
    bool b_small, positive;
};

struct FeatState {  // per-lane feature accumulators
    unsigned pmin, pmax, pdiff, isum32, prev_w;
    int imin, imax, idiff;
    double dsum;
};

// running aggregates of the open run of a lane
struct RunAgg {
    int kbest, ibest;
    unsigned cnt;
    unsigned long long sw;
    __device__ __forceinline__ void reset() { kbest = INT_MAX; ibest = INT_MAX; cnt = 0; sw = 0; }
    __device__ __forceinline__ void add(int i, int w, const LaneRec& r) {  // w in the offset domain
        int kv = r.positive ? 65535 - w : w;
        if (kv < kbest) { kbest = kv; ibest = i; }
        bool in = r.positive ? (w >= r.wlim + r.bias) : (w <= r.wlim + r.bias);
        cnt += in ? 1u : 0u;
        sw += in ? (unsigned)w : 0u;
    }
};

struct HitState {  // per-lane state of the run walker, lives across segments
    bool open;
    int run_s, prev_c, nh;
    unsigned carry;  // interesting-chunk bits of the previous segment not yet visited
    RunAgg g;
};

__device__ __forceinline__ int u16_at(const uint4& q, int j) {
    const unsigned w = (j < 4) ? ((j < 2) ? q.x : q.y) : ((j < 6) ? q.z : q.w);
    return (int)((w >> ((j & 1) * 16)) & 0xffffu);
}

// ---- hit sinks (per lane) ---------------------------------------------------------------------
struct PoolSink {
    LprEnt* pool;
    int* counter;
    bool overflow;
    __device__ __forceinline__ void store(int p, int s, int e, float height, float integral, int ord, const LaneRec&, const FHArgs&) {
        int idx = atomicAdd(counter, 1);
        if (idx < kLprEnt) {
            LprEnt h;
            h.ps = (unsigned)p | ((unsigned)s << 16);
            h.eo = (unsigned)e | ((unsigned)lane_id() << 16) | ((unsigned)ord << 21);
            h.height = height;
            h.integral = integral;
            *reinterpret_cast<uint4*>(&pool[idx]) = *reinterpret_cast<uint4*>(&h);
        } else {
            overflow = true;
        }
    }
};
struct DirectSink {  // rows straight to the output (records whose hits did not fit the pool)
    long long row0;
    bool active;
    __device__ __forceinline__ void store(int p, int s, int e, float height, float integral, int ord, const LaneRec& r, const FHArgs& a) {
        long long row = row0 + ord;
        if (active && row < a.hit_cap) {
            RowRec rr{r.ts, r.rid, r.len, r.dt, r.bc};
            unsigned w[15];
            hit_row_words(w, p, s, e, height, integral, rr, a.p.left_extension, a.p.right_extension, a.lmax);
            unsigned* dst = reinterpret_cast<unsigned*>(a.hit_out + row * kHitRowBytes);
#pragma unroll
            for (int k = 0; k < 15; ++k) dst[k] = w[k];
        }
    }
};

// ---- lane-parallel hit pass over one window of chunks ------------------------------------------
// `buf` is the lane's slot for the current segment; slot chunk k holds global chunk c = cbase + k.
// `hw` bit t marks global chunk c = c0 + t as interesting.
template <typename Sink>
__device__ __forceinline__ void lpr_hit_window(const uint8_t* buf, int cbase, int c0, int c_end, unsigned hw, const LaneRec& r,
                                               const FHArgs& a, HitState& hs, Sink& sink) {
    const unsigned sx = r.bias ? 0x80008000u : 0u;
    const unsigned xm = r.positive ? 0xffffffffu : 0u;
    const unsigned short* s16 = reinterpret_cast<const unsigned short*>(buf);
    const int left = a.p.left_extension, right = a.p.right_extension;
    const int vtotal = r.mis + r.len;
    auto sample = [&](int i) -> int {  // stored (offset-domain) value of record sample i; padding = true 0
        return (i < r.len) ? ((int)s16[r.mis + i - cbase * 8] ^ (int)(sx & 0xffffu)) : r.bias;
    };
    auto close_run = [&](int e) {
        const int a1 = min(a.lmax, e + right);
        for (int i = e; i < a1; ++i) hs.g.add(i, sample(i), r);
        const int wp = (r.positive ? 65535 - hs.g.kbest : hs.g.kbest) - r.bias;
        const float height = (float)(r.positive ? __dsub_rn((double)wp, r.b_rec) : __dsub_rn(r.b_rec, (double)wp));
        const long long c = hs.g.cnt;
        const long long swt = (long long)hs.g.sw - c * r.bias;
        double integ;
        if (r.b_small) {
            long long ipart = r.positive ? (swt - c * (long long)r.bi) : (c * (long long)r.bi - swt);
            double fpart = __dmul_rn((double)c, r.bf);
            integ = r.positive ? __dsub_rn((double)ipart, fpart) : __dadd_rn((double)ipart, fpart);
        } else {
            integ = r.positive ? __dsub_rn((double)swt, __dmul_rn((double)c, r.b_rec)) : __dsub_rn(__dmul_rn((double)c, r.b_rec), (double)swt);
        }
        sink.store(hs.g.ibest, hs.run_s, e, height, (float)integ, hs.nh, r, a);
        ++hs.nh;
        hs.open = false;
    };
    auto open_run = [&](int s) {
        hs.open = true;
        hs.run_s = s;
        hs.g.reset();
        for (int i = max(0, s - left); i < s; ++i) hs.g.add(i, sample(i), r);
    };
    // a run left open by the previous window ends at the window start if the next chunk is quiet
    if (hs.open && hs.prev_c + 1 == c0 && !(hw & 1u)) close_run(c0 * 8 - r.mis);
    unsigned m = hw;
    while (m) {
        const int t = __ffs(m) - 1;
        m &= m - 1;
        const int c = c0 + t;
        if (hs.open && c != hs.prev_c + 1) close_run((hs.prev_c + 1) * 8 - r.mis);  // a quiet chunk in between
        hs.prev_c = c;
        uint4 q = *reinterpret_cast<const uint4*>(buf + (c - cbase) * 16);
        q.x ^= sx; q.y ^= sx; q.z ^= sx; q.w ^= sx;
        const int v0 = c * 8;
        const int lo = min(max(r.mis - v0, 0), 8), hi = min(max(vtotal - v0, 0), 8);
        int wv[8];
        unsigned m8 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            wv[j] = u16_at(q, j);
            int kv = (int)((unsigned)wv[j] ^ (xm & 0xffffu));
            m8 |= ((kv <= r.kmax && j >= lo && j < hi) ? 1u : 0u) << j;
        }
        const int i0 = v0 - r.mis;
        int j = 0;
        while (j < 8) {
            if (hs.open) {
                const unsigned tz = (~(m8 >> j)) | 0x100u;  // first zero at or after j
                const int ones = min(__ffs(tz) - 1, 8 - j);
                const unsigned rm = ((1u << (j + ones)) - 1u) & ~((1u << j) - 1u);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    if ((rm >> jj) & 1u) hs.g.add(i0 + jj, wv[jj], r);
                j += ones;
                if (j < 8) close_run(i0 + j);
            } else {
                const unsigned tt = m8 >> j;
                if (!tt) break;
                j += __ffs(tt) - 1;
                open_run(i0 + j);
            }
        }
    }
    // still open after the last interesting chunk of the window and the next chunk (quiet, or past
    // the record end) belongs to this window: the run ends there, its extension is in the slot
    if (hs.open && hs.prev_c + 1 < c_end) close_run(min(r.len, (hs.prev_c + 1) * 8 - r.mis));
}

// ---- stream one record per lane through the slot ring -------------------------------------------
struct Ring {
    uint8_t* slot;  // this lane's slots: kNBuf buffers at stride buf_stride
    int buf_stride;
    unsigned long long* bars;  // kNBuf mbarriers of this warp
    unsigned* phase_bits;
    // 2-D tensor-map path: the 32 records of the warp are rows of one box (fixed-length contiguous
    // records), fetched by ONE cp.async.bulk.tensor per segment instead of 32 bulk copies
    const CUtensorMap* tmap;
    bool use2d;
    int row0;  // pool row (= record index in the pool) of lane 0's record
};

__device__ __forceinline__ void tma_tensor2d_g2s(void* dst_smem, const CUtensorMap* tmap, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

template <bool FEAT, bool HITS, typename Sink>
__device__ __forceinline__ void lpr_stream(const FHArgs& a, const uint16_t* pool, const LaneRec& r, int sc, const Ring& ring,
                                           int p0, int p1, int c0, int c1, FeatState& fs, HitState& hs, Sink& sink) {
    const int lane = lane_id();
    const int mis = r.mis, vtotal = r.mis + r.len;
    const int nch = (r.len > 0) ? ((vtotal + 7) >> 3) : 0;
    const int nch_max = __reduce_max_sync(kFull, nch);
    const int nseg = (nch_max + sc - 1) / sc;
    const bool known = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_NEGATIVE;
    const unsigned sx32 = r.bias ? 0x80008000u : 0u;
    const unsigned xm = r.positive ? 0xffffffffu : 0u;
    const float b32 = (float)r.b_feat;
    // chunk ranges of the "plain" fast path: [plainA, plainB) minus [plainHa, plainHb)
    int plainA = 0, plainB = 0, plainHa = 0, plainHb = 0;
    if (r.len > 0 && (!FEAT || !known)) {
        const int vlo = FEAT ? mis + max(c0, 1) : mis + 1;     // first sample that has a predecessor, inside the area range
        const int vhi = FEAT ? mis + min(c1, r.len) : vtotal;  // end of the area range
        plainA = (vlo + 7) >> 3;
        plainB = vhi >> 3;
        if (FEAT && p1 > p0) { plainHa = (mis + p0) >> 3; plainHb = (mis + p1 + 7) >> 3; }
    }

    auto issue = [&](int s) {
        const int b = s % kNBuf;
        const int cb = s * sc - 2 * kOV;  // global chunk held by slot chunk 0
        const int clo = max(0, cb), chi = min(nch, (s + 1) * sc);
        const unsigned bytes = chi > clo ? (unsigned)(chi - clo) * 16u : 0u;
        fence_proxy_async();  // generic reads of this buffer (two segments ago) before the refill
        if (ring.use2d) {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(&ring.bars[b], (unsigned)ring.buf_stride);
                tma_tensor2d_g2s(ring.slot + b * ring.buf_stride, ring.tmap, cb * 8, ring.row0, &ring.bars[b]);
            }
            return;
        }
        const unsigned total = __reduce_add_sync(kFull, bytes);
        if (lane == 0) {
            if (total) mbar_arrive_expect_tx(&ring.bars[b], total);
            else mbar_arrive(&ring.bars[b]);
        }
        __syncwarp();
        if (bytes) tma_bulk_g2s(ring.slot + b * ring.buf_stride + (clo - cb) * 16, pool + (r.off - mis) + (long long)clo * 8, bytes, &ring.bars[b]);
    };

    if (nseg > 0) issue(0);
    for (int s = 0; s < nseg; ++s) {
        const int b = s % kNBuf;
        if (s + 1 < nseg) issue(s + 1);
        mbar_wait(&ring.bars[b], (*ring.phase_bits >> b) & 1u);
        *ring.phase_bits ^= 1u << b;
        const uint8_t* buf = ring.slot + b * ring.buf_stride;
        const int cbase = s * sc - 2 * kOV;
        unsigned mw = 0;
        const int cend = min(sc, nch_max - s * sc);
        for (int cb = 0; cb < cend; ++cb) {
            const int vc = s * sc + cb;
            if (vc >= nch) continue;
            uint4 q = *reinterpret_cast<const uint4*>(buf + (2 * kOV + cb) * 16);
            if (sx32) { q.x ^= sx32; q.y ^= sx32; q.z ^= sx32; q.w ^= sx32; }  // int16 -> offset binary
            const int v0 = vc * 8;
            const int lo = max(mis - v0, 0), hi = min(vtotal - v0, 8);
            const int i0 = v0 - mis;
            if (vc >= plainA && vc < plainB && (vc < plainHa || vc >= plainHb)) {
                // plain interior chunk: 8 valid samples inside the area range, outside the height range
                if (FEAT) {
                    unsigned f0 = __funnelshift_r(fs.prev_w, q.x, 16), f1 = __funnelshift_r(q.x, q.y, 16);
                    unsigned f2 = __funnelshift_r(q.y, q.z, 16), f3 = __funnelshift_r(q.z, q.w, 16);
                    unsigned d0 = __vmaxu2(q.x, f0) - __vminu2(q.x, f0), d1 = __vmaxu2(q.y, f1) - __vminu2(q.y, f1);
                    unsigned d2 = __vmaxu2(q.z, f2) - __vminu2(q.z, f2), d3 = __vmaxu2(q.w, f3) - __vminu2(q.w, f3);
                    fs.pdiff = __vmaxu2(fs.pdiff, __vmaxu2(__vmaxu2(d0, d1), __vmaxu2(d2, d3)));
                    unsigned sacc = __dp2a_lo(q.x, 0x0101u, fs.isum32);
                    sacc = __dp2a_lo(q.y, 0x0101u, sacc);
                    sacc = __dp2a_lo(q.z, 0x0101u, sacc);
                    fs.isum32 = __dp2a_lo(q.w, 0x0101u, sacc);
                }
                if (HITS) {
                    unsigned mn = __vminu2(__vminu2(q.x ^ xm, q.y ^ xm), __vminu2(q.z ^ xm, q.w ^ xm));
                    int lmin = (int)min(mn & 0xffffu, mn >> 16);
                    mw |= (lmin <= r.kmax ? 1u : 0u) << cb;
                }
            } else if (lo == 0 && hi == 8) {
                if (FEAT) {
                    const unsigned pw = (i0 > 0) ? fs.prev_w : (q.x << 16);
                    unsigned f0 = __funnelshift_r(pw, q.x, 16), f1 = __funnelshift_r(q.x, q.y, 16);
                    unsigned f2 = __funnelshift_r(q.y, q.z, 16), f3 = __funnelshift_r(q.z, q.w, 16);
                    unsigned d0 = __vmaxu2(q.x, f0) - __vminu2(q.x, f0), d1 = __vmaxu2(q.y, f1) - __vminu2(q.y, f1);
                    unsigned d2 = __vmaxu2(q.z, f2) - __vminu2(q.z, f2), d3 = __vmaxu2(q.w, f3) - __vminu2(q.w, f3);
                    fs.pdiff = __vmaxu2(fs.pdiff, __vmaxu2(__vmaxu2(d0, d1), __vmaxu2(d2, d3)));
                    const int jlo = max(0, p0 - i0), jhi = min(8, p1 - i0);
                    if (jhi > jlo) {
                        if (jlo == 0 && jhi == 8) {
                            fs.pmin = __vminu2(fs.pmin, __vminu2(__vminu2(q.x, q.y), __vminu2(q.z, q.w)));
                            fs.pmax = __vmaxu2(fs.pmax, __vmaxu2(__vmaxu2(q.x, q.y), __vmaxu2(q.z, q.w)));
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                int w = u16_at(q, j);
                                if (j >= jlo && j < jhi) { fs.imin = min(fs.imin, w); fs.imax = max(fs.imax, w); }
                            }
                        }
                    }
                    const int klo = max(0, c0 - i0), khi = min(8, c1 - i0);
                    if (khi > klo) {
                        if (klo == 0 && khi == 8 && !known) {
                            unsigned sacc = __dp2a_lo(q.x, 0x0101u, fs.isum32);
                            sacc = __dp2a_lo(q.y, 0x0101u, sacc);
                            sacc = __dp2a_lo(q.z, 0x0101u, sacc);
                            fs.isum32 = __dp2a_lo(q.w, 0x0101u, sacc);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                unsigned w = (unsigned)u16_at(q, j);
                                if (j >= klo && j < khi) {
                                    if (!known) fs.isum32 += w;
                                    else fs.dsum += (double)(r.positive ? __fsub_rn((float)w, b32) : __fsub_rn(b32, (float)w));
                                }
                            }
                        }
                    }
                }
                if (HITS) {
                    unsigned mn = __vminu2(__vminu2(q.x ^ xm, q.y ^ xm), __vminu2(q.z ^ xm, q.w ^ xm));
                    int lmin = (int)min(mn & 0xffffu, mn >> 16);
                    mw |= (lmin <= r.kmax ? 1u : 0u) << cb;
                }
            } else {
                // partial chunk (record start / end): per-sample
                const int prev_s = (int)(fs.prev_w >> 16);
                bool any = false;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int w = u16_at(q, j);
                    const bool ok = (j >= lo) && (j < hi);
                    if (FEAT && ok) {
                        const int i = i0 + j;
                        if (i > 0) fs.idiff = max(fs.idiff, abs(w - ((j == 0) ? prev_s : u16_at(q, j - 1))));
                        if (i >= p0 && i < p1) { fs.imin = min(fs.imin, w); fs.imax = max(fs.imax, w); }
                        if (i >= c0 && i < c1) {
                            if (!known) fs.isum32 += (unsigned)w;
                            else fs.dsum += (double)(r.positive ? __fsub_rn((float)w, b32) : __fsub_rn(b32, (float)w));
                        }
                    }
                    if (HITS && ok) any = any || ((int)((unsigned)w ^ (xm & 0xffffu)) <= r.kmax);
                }
                if (HITS) mw |= (any ? 1u : 0u) << cb;
            }
            fs.prev_w = q.w;
        }
        if (HITS) {
            // window of this pass: chunks [s*sc - kOV, (s+1)*sc - kOV), the last segment runs to the end
            const bool last = (s == nseg - 1);
            unsigned hw = hs.carry | (mw << kOV);
            if (!last) {
                hs.carry = hw >> sc;                 // chunks (s+1)*sc - kOV .. : next pass
                hw &= (sc >= 32) ? 0xffffffffu : ((1u << sc) - 1u);
            }
            // chunks past the end of the window: the last pass covers everything that is left
            const int c_end = last ? (1 << 28) : (s + 1) * sc - kOV;
            lpr_hit_window(buf, cbase, s * sc - kOV, c_end, hw, r, a, hs, sink);
        }
        __syncwarp();  // every lane is done with buffer b before it is refilled
    }
}

// ---- the kernel --------------------------------------------------------------------------------
template <bool FEAT, bool HITS>
__global__ void __launch_bounds__(kLprWarps * 32) lpr_kernel(const FHArgs a, const int sc, const __grid_constant__ CUtensorMap tmap,
                                                             const int have_tmap) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];  // [warp][kNBuf][lane] slots of a.slot_bytes
    __shared__ __align__(16) LprEnt s_ent[HITS ? kLprWarps : 1][HITS ? kLprEnt : 1];
    __shared__ int s_pool[kLprWarps];
    __shared__ long long s_wtot[kLprWarps];
    __shared__ long long s_wbase[kLprWarps];
    __shared__ int s_tile;
    __shared__ __align__(8) unsigned long long s_bar[kLprWarps * kNBuf];

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint16_t* pool = static_cast<const uint16_t*>(a.pool);
    unsigned phase_bits = 0;
    Ring ring;
    // lane stride = slot_bytes (odd multiple of 16): conflict-free LDS.128; buffers kNBuf apart by 32 lanes
    ring.buf_stride = 32 * a.slot_bytes;
    ring.slot = dyn_smem + (size_t)warp * kNBuf * ring.buf_stride + (size_t)lane * a.slot_bytes;
    ring.bars = &s_bar[warp * kNBuf];
    ring.phase_bits = &phase_bits;
    ring.tmap = &tmap;
    ring.use2d = false;
    ring.row0 = 0;
    if (lane < kNBuf) mbar_init(&ring.bars[lane], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    for (;;) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= a.n_tiles) break;

        // ---------------- per-lane bookkeeping
        const long long rec = (long long)tile * kLprTile + warp * 32 + lane;
        const bool have = rec < a.n;
        LaneRec r;
        r.off = 0; r.ts = 0; r.rid = 0; r.len = 0; r.dt = 1; r.pol = 0; r.bc = 0;
        r.b_rec = 0.0; r.b_feat = 0.0; r.thr = a.p.threshold;
        if (have) {
            const uint4* q = reinterpret_cast<const uint4*>(a.meta + rec);
            uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            r.ts = (long long)(((unsigned long long)q0.y << 32) | q0.x);
            r.b_rec = __hiloint2double((int)q0.w, (int)q0.z);
            r.off = (long long)(((unsigned long long)q1.y << 32) | q1.x) - a.p.pool_base;
            r.len = (int)q1.z;
            r.dt = (int)q1.w;
            r.bc = q2.x;
            r.pol = (int)(q2.y & 0xff);
            r.rid = (long long)(((unsigned long long)q2.w << 32) | q2.z);
            r.b_feat = r.b_rec;
            const int board = (int)(short)(r.bc & 0xffff), channel = (int)(short)(r.bc >> 16);
            for (int i = 0; i < a.p.n_rules; ++i) {
                const wfb_chan_rule rule = a.p.rules_dev[i];
                if (rule.board == board && rule.channel == channel) {
                    if (rule.has_threshold) r.thr = rule.threshold;
                    if (rule.has_fixed_baseline) r.b_feat = rule.fixed_baseline;
                }
            }
            if (r.len < 0) r.len = 0;
            if (r.len > 0 && (r.off < 0 || r.off + r.len > a.pool_len)) {
                atomicExch(a.err_flag, 1);
                r.len = 0;
            }
            if (r.len > a.lmax) {
                atomicExch(a.err_flag, 2);  // lmax passed by the caller is too small
                r.len = 0;
            }
        }
        r.mis = (int)(r.off & 7);
        r.bias = a.p.signed_samples ? 32768 : 0;
        r.positive = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_RAW_POSITIVE;
        const bool rawpos = r.pol == WFB_POL_RAW_POSITIVE;
        const bool known = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_NEGATIVE;
        int p0 = 0, p1 = 0, c0 = 0, c1 = 0;
        if (FEAT) {
            resolve_slice(a.p.height_start, a.p.height_end, r.len, p0, p1);
            resolve_slice(a.p.area_start, a.p.area_end, r.len, c0, c1);
        }
        r.kmax = -1;
        r.b_small = true; r.bi = 0; r.bf = 0; r.wlim = 0;
        if (HITS && r.len > 0) {
            r.kmax = integer_threshold_u16(r.b_rec, r.thr, r.positive, r.bias);
            r.b_small = fabs(r.b_rec) < 2e9;
            r.bi = floor(r.b_rec);
            r.bf = __dsub_rn(r.b_rec, r.bi);
            const int ib = r.b_small ? (int)r.bi : (r.b_rec > 0 ? INT_MAX : INT_MIN);
            r.wlim = r.positive ? ib + 1 : ((r.bi == r.b_rec) ? ib - 1 : ib);
        }

        // one tensor copy per segment when the warp's records are rows of the fixed-length pool
        {
            const long long row = (long long)tile * kLprTile + warp * 32 + lane;  // record index == pool row
            const bool fits = !have || (r.len == a.lmax && r.off == row * (long long)a.lmax);
            ring.use2d = have_tmap && __all_sync(kFull, fits);
            ring.row0 = (int)((long long)tile * kLprTile + warp * 32);
        }
        // ---------------- stream the record: features + hits into the warp's pool
        FeatState fs;
        fs.pmin = 0xffffffffu; fs.pmax = 0u; fs.pdiff = 0u; fs.isum32 = 0u; fs.prev_w = 0u;
        fs.imin = INT_MAX; fs.imax = INT_MIN; fs.idiff = 0; fs.dsum = 0.0;
        HitState hs;
        hs.open = false; hs.run_s = 0; hs.prev_c = -4; hs.nh = 0; hs.carry = 0u; hs.g.reset();
        if (HITS && lane == 0) s_pool[warp] = 0;
        __syncwarp();
        PoolSink psink{HITS ? &s_ent[warp][0] : nullptr, &s_pool[warp], false};
        lpr_stream<FEAT, HITS>(a, pool, r, sc, ring, p0, p1, c0, c1, fs, hs, psink);

        // ---------------- features of my record
        if (FEAT && have) {
            const float b32 = (float)r.b_feat;
            float height = 0.f, amp = 0.f, area = 0.f;
            const float mad = (float)max(fs.idiff, (int)max(fs.pdiff & 0xffffu, fs.pdiff >> 16));
            if (p1 > p0) {
                const int wmin = min(fs.imin, (int)min(fs.pmin & 0xffffu, fs.pmin >> 16));
                const int wmax = max(fs.imax, (int)max(fs.pmax & 0xffffu, fs.pmax >> 16));
                if (!known) {
                    height = rawpos ? (float)__dsub_rn((double)(wmax - r.bias), r.b_feat) : (float)__dsub_rn(r.b_feat, (double)(wmin - r.bias));
                    amp = (float)(wmax - wmin);
                } else {
                    float smax = r.positive ? __fsub_rn((float)wmax, b32) : __fsub_rn(b32, (float)wmin);
                    float smin = r.positive ? __fsub_rn((float)wmin, b32) : __fsub_rn(b32, (float)wmax);
                    height = smax;
                    amp = (float)__dsub_rn((double)smax, (double)smin);
                }
            }
            if (c1 > c0) {
                if (!known) {
                    const long long nC = c1 - c0;
                    const long long sw = (long long)fs.isum32 - nC * r.bias;
                    const double b = r.b_feat;
                    double ar;
                    if (fabs(b) < 1e12) {
                        double bi = floor(b), bf = __dsub_rn(b, bi);
                        double fpart = __dmul_rn((double)nC, bf);
                        if (rawpos) ar = __dsub_rn((double)(sw - nC * (long long)bi), fpart);
                        else ar = __dadd_rn((double)(nC * (long long)bi - sw), fpart);
                    } else {
                        ar = rawpos ? __dsub_rn((double)sw, __dmul_rn((double)nC, b)) : __dsub_rn(__dmul_rn((double)nC, b), (double)sw);
                    }
                    area = (float)ar;
                } else {
                    area = (float)fs.dsum;
                }
            }
            unsigned* dst = reinterpret_cast<unsigned*>(a.feat_out + rec * kFeatRowBytes);
            const long long ev = a.p.row_base + rec;
            dst[0] = __float_as_uint(height);
            dst[1] = __float_as_uint(amp);
            dst[2] = __float_as_uint(area);
            dst[3] = __float_as_uint(mad);
            dst[4] = (unsigned)(r.ts & 0xffffffffll);
            dst[5] = (unsigned)((unsigned long long)r.ts >> 32);
            dst[6] = r.bc;
            dst[7] = (unsigned)(ev & 0xffffffffll);
            dst[8] = (unsigned)((unsigned long long)ev >> 32);
        }
        if (!HITS) {
            __syncthreads();  // s_tile reuse
            continue;
        }
        const int my_cnt = hs.nh;
        if (a.hit_counts != nullptr && have) a.hit_counts[rec] = my_cnt;

        int incl = my_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long c = (lane < kLprWarps) ? s_wtot[lane] : 0;
            long long wincl = c;
#pragma unroll
            for (int d = 1; d < kLprWarps; d <<= 1) {
                long long t = bcast_i64(wincl, max(lane - d, 0));
                if (lane >= d) wincl += t;
            }
            const long long total = bcast_i64(wincl, kLprWarps - 1);
            const long long excl = tile_lookback(a, tile, total);
            if (lane < kLprWarps) s_wbase[lane] = excl + (wincl - c);
        }
        __syncthreads();

        // ---------------- phase B: one pooled hit per lane -> packed row
        const long long my_row0 = s_wbase[warp] + (incl - my_cnt);
        const unsigned ovf = __ballot_sync(kFull, psink.overflow);
        const int used = min(s_pool[warp], kLprEnt);
        for (int e0 = 0; e0 < used; e0 += 32) {
            const int e = e0 + lane;
            const bool act = e < used;
            LprEnt h;
            *reinterpret_cast<uint4*>(&h) = *reinterpret_cast<const uint4*>(&s_ent[warp][act ? e : 0]);
            const int owner = (int)((h.eo >> 16) & 31u);
            RowRec rr;
            rr.ts = bcast_i64(r.ts, owner);
            rr.rid = bcast_i64(r.rid, owner);
            rr.len = __shfl_sync(kFull, r.len, owner);
            rr.dt = __shfl_sync(kFull, r.dt, owner);
            rr.bc = __shfl_sync(kFull, r.bc, owner);
            const long long row = bcast_i64(my_row0, owner) + (long long)(h.eo >> 21);
            if (act && !((ovf >> owner) & 1u) && row < a.hit_cap) {
                unsigned w[15];
                hit_row_words(w, (int)(h.ps & 0xffffu), (int)(h.ps >> 16), (int)(h.eo & 0xffffu), h.height, h.integral, rr,
                              a.p.left_extension, a.p.right_extension, a.lmax);
                unsigned* dst = reinterpret_cast<unsigned*>(a.hit_out + row * kHitRowBytes);
#pragma unroll
                for (int k = 0; k < 15; ++k) dst[k] = w[k];
            }
        }
        if (ovf) {  // some records' hits did not all fit the pool: stream those again, rows go straight out
            FeatState fs2 = fs;
            HitState hs2;
            hs2.open = false; hs2.run_s = 0; hs2.prev_c = -4; hs2.nh = 0; hs2.carry = 0u; hs2.g.reset();
            DirectSink dsink{my_row0, psink.overflow};
            lpr_stream<false, true>(a, pool, r, sc, ring, 0, 0, 0, 0, fs2, hs2, dsink);
        }
        __syncthreads();  // pool, slots and s_tile are reused by the next tile
    }
}

int launch_lpr(FHArgs a, int flags, cudaStream_t st) {
    if (a.p.pool_is_f32) return 1;
    const char* force = getenv("WFB_FUSED_VARIANT");
    if (force && (!strcmp(force, "global") || !strcmp(force, "staged"))) return 1;
    if (a.lmax >= 65536 - 16) return 1;  // positions are packed in 16 bits
    if (a.p.left_extension > 8 * kOV || a.p.right_extension > 8 * kOV) return 1;
    // segment length in chunks: 2 * kOV history chunks + sc new chunks per slot, sc + kOV <= 32 mask bits
    int sc = 12;
    if (const char* e = getenv("WFB_LPR_SC")) sc = std::max(4, std::min(30, atoi(e)));
    int slot_chunks = sc + 2 * kOV;
    if ((slot_chunks & 1) == 0) ++slot_chunks;  // odd multiple of 16 bytes: conflict-free LDS.128
    a.slot_bytes = slot_chunks * 16;
    const size_t dyn = (size_t)kLprTile * kNBuf * a.slot_bytes;
    a.n_tiles = (int)((a.n + kLprTile - 1) / kLprTile);
    const bool f = flags & WFB_DO_FEATURES, h = flags & WFB_DO_HITS;
    // tensor map over the pool seen as rows of lmax samples: one box = 32 records x one slot
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int have_tmap = 0;
    const char* no2d = getenv("WFB_LPR_NO_TMAP");
    if (!(no2d && no2d[0] == '1') && a.lmax > 0 && a.lmax % 8 == 0 && a.pool_len >= a.lmax && a.n < (1ll << 31)) {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static encode_fn encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
                encode = reinterpret_cast<encode_fn>(fn);
        }
        if (encode) {
            cuuint64_t dims[2] = {(cuuint64_t)a.lmax, (cuuint64_t)(a.pool_len / a.lmax)};
            cuuint64_t strides[1] = {(cuuint64_t)a.lmax * 2};
            cuuint32_t box[2] = {(cuuint32_t)(a.slot_bytes / 2), 32};
            cuuint32_t estr[2] = {1, 1};
            CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(a.pool), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            have_tmap = (cr == CUDA_SUCCESS) ? 1 : 0;
        }
    }
    auto go = [&](auto kern) -> int {
        WFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        int per_sm = 0;
        WFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLprWarps * 32, dyn));
        if (per_sm < 1) {
            set_error("lane-per-record kernel does not fit the SM");
            return WFB_ERR_CUDA;
        }
        int grid = (int)std::min<long long>((long long)sm_count() * per_sm, a.n_tiles);
        kern<<<grid, kLprWarps * 32, dyn, st>>>(a, sc, tmap, have_tmap);
        WFB_CUDA(cudaGetLastError());
        return WFB_OK;
    };
    if (f && h) return go(lpr_kernel<true, true>);
    if (f) return go(lpr_kernel<true, false>);
    return go(lpr_kernel<false, true>);
}

}  // namespace wfb
