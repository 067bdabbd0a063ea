// K2 + K3, lane-per-record streaming variant (uint16 / int16 pools).
//
// Each LANE owns one record and streams it through a small double-buffered shared-memory slot:
// segments of `sc` 16-byte chunks (8 samples each) arrive by TMA - one 2-D tensor copy per warp
// and segment when the 32 records are rows of a fixed-length pool, else one 1-D bulk copy per
// lane - issued one segment ahead; the slot stride is an odd multiple of 16 bytes so the 32 lanes'
// LDS.128 hit distinct bank groups.  The sample loop has no cross-lane traffic: packed 16x2
// integer min / max / |diff| / dot-product-sum on the lane's own registers.
//
// Hits are found in two layers so that the per-sample run walking is done densely:
//   * the converged scan classifies each chunk from two packed reductions (min and max of the
//     threshold keys): QUIET (no sample above threshold), FULL (all eight above) or MIXED.  FULL
//     chunks only feed a per-lane running aggregate (first arg-min key, count, sum) - they are
//     interior to a run.  Every run start and end lies in a non-FULL chunk that is MIXED, follows
//     an above-threshold sample or precedes a FULL chunk; those chunks become ITEMS: the lane
//     queues (owner, chunk index, snapshot of its FULL-chunk aggregate) in a 64-slot per-warp ring;
//   * whenever 32 items wait, the warp runs a DENSE ROUND: lane t takes item t whoever owns it,
//     fetches the owner's record constants by shuffle, re-reads the chunk and its two neighbours
//     from L2 (the extensions, <= 8 samples, live there), walks the runs of that chunk, and stitches
//     "run still open at the chunk end" fragments to the next item of the same record through a
//     per-owner carry in shared memory (the items of one record are consecutive in chunk order
//     and only FULL chunks can lie between the two ends of a run).  Finished hits carry their
//     owner and per-record ordinal into a per-warp pool in global memory (L2 resident, double
//     buffered); one decoupled look-back per 128-record tile gives the first output row - resolved
//     one tile LATER, after the block has streamed its next tile, so nobody waits for a
//     predecessor - and the rows are assembled one hit per lane.
//
// Reference semantics: see fused_features_hits.cu (same arithmetic, same results).
#include <cuda.h>
#include <float.h>
#include <stdlib.h>

#include <algorithm>

#include "fused_common.cuh"

namespace wfb {

#ifndef WFB_LPR_WARPS
#define WFB_LPR_WARPS 4
#endif
constexpr int kLprWarps = WFB_LPR_WARPS;
constexpr int kLprTile = kLprWarps * 32;  // records per tile / look-back unit
#ifndef WFB_LPR_MINBLOCKS
#define WFB_LPR_MINBLOCKS 3  // blocks per SM the register allocation must allow
#endif
static_assert(WFB_LPR_WARPS <= 8, "a block has at most eight warps (workspace hit pool, block scan)");
constexpr int kHist = 0;                  // chunks of history in front of each segment (none: items re-read their chunks from L2)
constexpr int kMaxExt = 8;                // extensions must fit the neighbouring chunk
constexpr int kNBuf = 2;                  // slot buffers per lane
constexpr int kQCap = 32;                 // items per dense round
constexpr int kQRing = 64;                // queue slots: a round always takes 32 items while more are waiting

struct LprEnt {  // 16 bytes, raw aggregates: height / integral are finished one hit per lane in phase B
    unsigned ps;  // p | s << 16
    unsigned eo;  // e | owner lane << 16 | ordinal << 21
    unsigned kc;  // best threshold key | signal-side sample count << 16
    unsigned sw;  // sum of the signal-side samples (offset domain)
};

struct LaneRec {  // everything a lane knows about its record
    long long off, ts, rid;
    int len, dt, pol, mis, bias;
    int clen;  // length the edges of the hit rows are clamped to (event_length unless wfb_meta_set_clamp set another)
    unsigned bc;
    double b_rec, b_feat, thr;
    int kmax;
    int wlim;  // integer bound of the samples on the signal side of the baseline (hit integral)
    float xb;  // float32 pools: above threshold <=> x <= xb (negative pulses) / x >= xb (positive); NaN: never
    bool positive, degen;
};

struct FeatState {  // per-lane feature accumulators
    unsigned pmin, pmax, pdiff, isum32, prev_w;
    int imin, imax, idiff;
    double dsum;
};

struct DefLane {  // what phase B needs of a record once its registers are gone
    long long ts, rid;
    double b;
    int len, dt;
    unsigned bc;
    unsigned rel_pos;  // first row of the record relative to the tile's first row << 1 | positive
};

struct ChunkQ {  // chunk-item queue of the chunk-granular variant (per warp)
    uint4 q_nb[2][32];     // the round's scratch: chunks P-1 and P+1 of lane t's item
    uint4 q_hdr[kQRing];   // ring of queued items: owner | (P + 1) << 5 ; aggregate key, samples, sum of the FULL chunks in front
};

struct WarpHits {  // per-warp shared memory of the hit machinery
    uint4 stage[32];       // this round's "open at chunk end" fragments: start ; key, count, key sum
    uint4 carry[32];       // the same, per owner, across rounds
    int stage_n[32];       // runs started by this round's items
    int carry_n[32];       // runs started so far, per owner
    DefLane def[32];       // row data of the tile whose rows are still to be written (deferred phase B)
    int pool_cnt;
    unsigned ovf;  // owners whose hits did not all fit the pool
    int def_used;          // pooled hits / overflow owners of the deferred tile
    unsigned def_ovf;
};

__device__ __forceinline__ int u16_at(const uint4& q, int j) {
    const unsigned w = (j < 4) ? ((j < 2) ? q.x : q.y) : ((j < 6) ? q.z : q.w);
    return (int)((w >> ((j & 1) * 16)) & 0xffffu);
}

// height / integral of a hit from its raw aggregates (hit_finder.py:372-381; float64, no contraction)
__device__ __forceinline__ void hit_values(int kbest, unsigned cnt, unsigned sw, bool pos, double b, int bias, float& height, float& integral) {
    const int wp = (pos ? 65535 - kbest : kbest) - bias;
    height = (float)(pos ? __dsub_rn((double)wp, b) : __dsub_rn(b, (double)wp));
    const long long c = cnt;
    const long long swt = (long long)sw - c * bias;
    double integ;
    if (fabs(b) < 2e9) {
        const double bi = floor(b), bf = __dsub_rn(b, bi);
        const long long ipart = pos ? (swt - c * (long long)bi) : (c * (long long)bi - swt);
        const double fpart = __dmul_rn((double)c, bf);
        integ = pos ? __dsub_rn((double)ipart, fpart) : __dadd_rn((double)ipart, fpart);
    } else {
        integ = pos ? __dsub_rn((double)swt, __dmul_rn((double)c, b)) : __dsub_rn(__dmul_rn((double)c, b), (double)swt);
    }
    integral = (float)integ;
}

// L2 policy of the hit pool: it is written and read back one tile later, so it should stay in L2 (evict
// last) instead of being pushed out to HBM by the streaming samples and fetched again
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_pool(uint4* dst, const uint4 v, unsigned long long policy) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
}

// ---- hit sinks ---------------------------------------------------------------------------------
struct PoolSink {  // hits of the tile being streamed: the warp's half of its L2-resident pool
    WarpHits* ws;
    uint4* ent;
    int cap;
    unsigned long long policy;  // L2 evict-last descriptor, created once per tile
    __device__ __forceinline__ void prepare(int, const LaneRec&) {}
    __device__ __forceinline__ void store(int p, int s, int e, int kbest, unsigned cnt, unsigned sw, int ord, int owner, const FHArgs&) {
        int idx = atomicAdd(&ws->pool_cnt, 1);
        if (idx < cap) {
            st_pool(ent + idx,
                    make_uint4((unsigned)p | ((unsigned)s << 16), (unsigned)e | ((unsigned)owner << 16) | ((unsigned)ord << 21),
                               (unsigned)kbest | (cnt << 16), sw),
                    policy);
        } else {
            atomicOr(&ws->ovf, 1u << owner);
        }
    }
    // float32 pools: height and integral are final already
    __device__ __forceinline__ void store_f32(int p, int s, int e, float height, float integral, int ord, int owner, const FHArgs&) {
        int idx = atomicAdd(&ws->pool_cnt, 1);
        if (idx < cap) {
            st_pool(ent + idx,
                    make_uint4((unsigned)p | ((unsigned)s << 16), (unsigned)e | ((unsigned)owner << 16) | ((unsigned)ord << 21),
                               __float_as_uint(height), __float_as_uint(integral)),
                    policy);
        } else {
            atomicOr(&ws->ovf, 1u << owner);
        }
    }
};
struct DirectSink {  // rows straight to the output (records whose hits did not fit the pool)
    long long my_row0;
    bool my_active;
    long long row0;
    bool active, pos;
    double b;
    int bias;
    RowRec rr;
    __device__ __forceinline__ void prepare(int src, const LaneRec& r) {  // converged: fetch the owner's row data
        row0 = bcast_i64(my_row0, src);
        active = __shfl_sync(kFull, (int)my_active, src) != 0;
        pos = __shfl_sync(kFull, (int)r.positive, src) != 0;
        b = shfl_f64(r.b_rec, src);
        bias = r.bias;
        rr.ts = bcast_i64(r.ts, src);
        rr.rid = bcast_i64(r.rid, src);
        rr.len = __shfl_sync(kFull, r.clen, src);
        rr.dt = __shfl_sync(kFull, r.dt, src);
        rr.bc = __shfl_sync(kFull, r.bc, src);
    }
    __device__ __forceinline__ void store(int p, int s, int e, int kbest, unsigned cnt, unsigned sw, int ord, int, const FHArgs& a) {
        long long row = row0 + ord;
        if (active && row < a.hit_cap) {
            float height, integral;
            hit_values(kbest, cnt, sw, pos, b, bias, height, integral);
            unsigned w[15];
            hit_row_words(w, p, s, e, height, integral, rr, a.p.left_extension, a.p.right_extension, a.lmax);
            unsigned* dst = reinterpret_cast<unsigned*>(a.hit_out + row * kHitRowBytes);
#pragma unroll
            for (int k = 0; k < 15; ++k) dst[k] = w[k];
        }
    }
    __device__ __forceinline__ void store_f32(int p, int s, int e, float height, float integral, int ord, int, const FHArgs& a) {
        long long row = row0 + ord;
        if (active && row < a.hit_cap) {
            unsigned w[15];
            hit_row_words(w, p, s, e, height, integral, rr, a.p.left_extension, a.p.right_extension, a.lmax);
            unsigned* dst = reinterpret_cast<unsigned*>(a.hit_out + row * kHitRowBytes);
#pragma unroll
            for (int k = 0; k < 15; ++k) dst[k] = w[k];
        }
    }
};

// ---- dense round: lane t walks the runs of queued item t ---------------------------------------
// Aggregates are kept in the key domain (kv = w for negative pulses, 65535 - w for positive ones;
// above threshold <=> kv <= kmax, signal side <=> kv <= kin): a 32-bit key kv << 16 | sample index
// whose minimum is the first arg-max of the signal, and count << 21 | sum(kv) of the signal-side
// samples (at most 24 samples per fragment, so the packed sum cannot carry into the count).
// EXT22: both extensions are two samples (the defaults): the four neighbour samples are decoded once per
// round and every fragment is one unrolled masked pass over the 12-sample window, all in registers.
template <bool EXT22, typename Sink>
__device__ __forceinline__ void lpr_round(WarpHits& ws, ChunkQ& cq, int qh, int qn, const LaneRec& r, const FHArgs& a, Sink& sink) {
    __syncwarp();  // the pushes are visible
    const int lane = lane_id();
    const bool act = lane < qn;
    const uint4 hd = cq.q_hdr[(qh + (act ? lane : 0)) & (kQRing - 1)];
    const int src = act ? (int)(hd.x & 31u) : lane;  // owner lane
    const int P = (int)((hd.x >> 5) & 0x3fffu) - 1;  // -1: the virtual chunk in front of a record that starts FULL
    // the owner's record constants
    const int o_kmax = __shfl_sync(kFull, r.kmax, src);
    const int o_mis = __shfl_sync(kFull, r.mis, src);
    const int o_len = __shfl_sync(kFull, r.len, src);
    const int o_wlim = __shfl_sync(kFull, r.wlim, src);
    const bool o_pos = __shfl_sync(kFull, (int)r.positive, src) != 0;
    const long long o_off = bcast_i64(r.off, src);
    sink.prepare(src, r);
    const int bias = r.bias;  // uniform
    const unsigned cx = (bias ? 0x8000u : 0u) ^ (o_pos ? 0xffffu : 0u);  // raw sample -> key
    const int padkv = o_pos ? 65535 - bias : bias;                       // key of a padding sample (true 0)
    const int kin = o_pos ? 65535 - (o_wlim + bias) : o_wlim + bias;     // signal side <=> kv <= kin
    const int left = a.p.left_extension, right = a.p.right_extension;
    const int i0 = 8 * P - o_mis;
    constexpr unsigned kOne = 1u << 21;
    // the item's chunk and its two neighbours come back from L2 (they were streamed a few microseconds
    // ago); chunks outside the record are never dereferenced.  They are parked in shared memory for the
    // dynamically indexed extension samples.
    const int o_nch = (o_len > 0) ? ((o_mis + o_len + 7) >> 3) : 0;
    const uint4* gsrc = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(a.pool) + (o_off - o_mis)) + (P - 1);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const uint4 c0q = (act && P - 1 >= 0 && P - 1 < o_nch) ? __ldg(gsrc) : zero4;
    const uint4 c1 = (act && P >= 0 && P < o_nch) ? __ldg(gsrc + 1) : zero4;
    const uint4 c2q = (act && P + 1 < o_nch) ? __ldg(gsrc + 2) : zero4;
    if (!EXT22) {
        cq.q_nb[0][lane] = c0q;
        cq.q_nb[1][lane] = c2q;
    }
    // keys of the item's own chunk, packed contributions, above-threshold mask
    unsigned ckey[8], cval[8];
    unsigned m8 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i = i0 + j;
        const int kv = (i < o_len) ? (int)((unsigned)u16_at(c1, j) ^ cx) : padkv;
        ckey[j] = ((unsigned)kv << 16) + (unsigned)i;
        cval[j] = (kv <= kin) ? kOne + (unsigned)kv : 0u;
        m8 |= ((act && i >= 0 && i < o_len && kv <= o_kmax) ? 1u : 0u) << j;
    }
    // does a run enter the chunk (last sample of chunk P-1 above threshold), is chunk P+1 FULL?
    const bool in_open = act && i0 >= 1 && i0 - 1 < o_len && (int)((c0q.w >> 16) ^ cx) <= o_kmax;
    const unsigned cx32 = cx | (cx << 16);
    const unsigned nmx = __vmaxu2(__vimax3_u16x2(c2q.x ^ cx32, c2q.y ^ cx32, c2q.z ^ cx32), c2q.w ^ cx32);
    const bool next_full = act && i0 + 8 >= 0 && i0 + 16 <= o_len && (int)max(nmx & 0xffffu, nmx >> 16) <= o_kmax;
    const bool nfs = next_full && !(m8 & 0x80u);  // a run starts with the FULL chunk that follows
    const bool has_trail = (m8 & 0x80u) || nfs;
    const int nstarts = __popc(m8 & ~((m8 << 1) | (in_open ? 1u : 0u))) + (nfs ? 1 : 0);
    ws.stage_n[lane] = nstarts;
    const unsigned peers = __match_any_sync(kFull, act ? src : 32 + lane);
    const unsigned lt = peers & ((1u << lane) - 1u);
    __syncwarp();
    int ord_base = act ? ws.carry_n[src] : 0;
    for (unsigned m = lt; m; m &= m - 1) ord_base += ws.stage_n[__ffs(m) - 1];

    // EXT22: samples i0-2, i0-1 (end of chunk P-1) and i0+8, i0+9 (start of chunk P+1)
    unsigned nkey[4], nval[4];
    if (EXT22) {
        const int kl0 = (int)((c0q.w & 0xffffu) ^ cx), kl1 = (int)((c0q.w >> 16) ^ cx);
        const int kr0 = (i0 + 8 < o_len) ? (int)((c2q.x & 0xffffu) ^ cx) : padkv;
        const int kr1 = (i0 + 9 < o_len) ? (int)((c2q.x >> 16) ^ cx) : padkv;
        nkey[0] = ((unsigned)kl0 << 16) + (unsigned)(i0 - 2);
        nkey[1] = ((unsigned)kl1 << 16) + (unsigned)(i0 - 1);
        nkey[2] = ((unsigned)kr0 << 16) + (unsigned)(i0 + 8);
        nkey[3] = ((unsigned)kr1 << 16) + (unsigned)(i0 + 9);
        nval[0] = (kl0 <= kin) ? kOne + (unsigned)kl0 : 0u;
        nval[1] = (kl1 <= kin) ? kOne + (unsigned)kl1 : 0u;
        nval[2] = (kr0 <= kin) ? kOne + (unsigned)kr0 : 0u;
        nval[3] = (kr1 <= kin) ? kOne + (unsigned)kr1 : 0u;
    }
    // aggregate of samples [a0, a1) (always inside [i0 - left, i0 + 8 + right)): own chunk from registers, the
    // neighbours' few samples from registers (EXT22) or from the round's scratch rows
    auto frag = [&](int a0, int a1, unsigned& key, unsigned& acc) {
        if (EXT22) {
            const int wlo = min(max(a0 - (i0 - 2), 0), 12), whi = min(max(a1 - (i0 - 2), 0), 12);
            const unsigned msk = ((1u << whi) - 1u) & ~((1u << wlo) - 1u);  // bit w <-> sample i0 - 2 + w
            if (msk & 1u) { key = min(key, nkey[0]); acc += nval[0]; }
            if (msk & 2u) { key = min(key, nkey[1]); acc += nval[1]; }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (msk & (4u << j)) { key = min(key, ckey[j]); acc += cval[j]; }
            if (msk & 0x400u) { key = min(key, nkey[2]); acc += nval[2]; }
            if (msk & 0x800u) { key = min(key, nkey[3]); acc += nval[3]; }
            return;
        }
        {
            const int jlo = min(max(a0 - i0, 0), 8), jhi = min(max(a1 - i0, 0), 8);
            const unsigned msk = ((1u << jhi) - 1u) & ~((1u << jlo) - 1u);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (msk & (1u << j)) { key = min(key, ckey[j]); acc += cval[j]; }
        }
        // the neighbours' samples: warp-uniform trip counts (the extensions), predicated bodies
        const unsigned short* lo_row = reinterpret_cast<const unsigned short*>(&cq.q_nb[0][lane]);
        for (int e = 1; e <= left; ++e) {  // left neighbour (never padding: i < i0 <= len)
            const int i = i0 - e;
            const int kv = (int)((unsigned)lo_row[8 - e] ^ cx);
            if (i >= a0 && i < a1) {
                key = min(key, ((unsigned)kv << 16) + (unsigned)i);
                acc += (kv <= kin) ? kOne + (unsigned)kv : 0u;
            }
        }
        const unsigned short* hi_row = reinterpret_cast<const unsigned short*>(&cq.q_nb[1][lane]);
        for (int e = 0; e < right; ++e) {  // right neighbour
            const int i = i0 + 8 + e;
            const int kv = (i < o_len) ? (int)((unsigned)hi_row[e] ^ cx) : padkv;
            if (i >= a0 && i < a1) {
                key = min(key, ((unsigned)kv << 16) + (unsigned)i);
                acc += (kv <= kin) ? kOne + (unsigned)kv : 0u;
            }
        }
    };
    // the fragment still open at the chunk end: from its start (minus the left extension) to the chunk end
    int js_tr = 8;
    if (has_trail) {
        if (!nfs) js_tr = 8 - __clz(~(m8 << 24));
        const int s_tr = i0 + js_tr;
        unsigned key = 0xffffffffu, acc = 0;
        frag(max(0, s_tr - left), i0 + 8, key, acc);
        ws.stage[lane] = make_uint4((unsigned)s_tr, key, acc >> 21, acc & (kOne - 1u));
    }
    __syncwarp();
    // runs that end in this chunk
    uint4 prev = make_uint4(0u, 0xffffffffu, 0u, 0u);
    if (in_open) {
        prev = lt ? ws.stage[31 - __clz(lt)] : ws.carry[src];
        if (hd.y != 0xffffffffu) {  // FULL chunks between the two items: re-read the one that holds their minimum (L2)
            const uint4 bq = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(a.pool) + (o_off - o_mis) + 8ll * (long long)(hd.y & 0xffffu)));
            const unsigned bkv = hd.y >> 16;
            int jb = 7;
#pragma unroll
            for (int j = 6; j >= 0; --j)
                if (((unsigned)u16_at(bq, j) ^ cx) == bkv) jb = j;
            prev.y = min(prev.y, (bkv << 16) + (unsigned)(8 * (int)(hd.y & 0xffffu) - o_mis + jb));
            prev.z += hd.z;
            prev.w += o_pos ? hd.z * 65535u - hd.w : hd.w;
        }
    }
    unsigned mm = (has_trail && !nfs) ? (m8 & ((1u << js_tr) - 1u)) : m8;
    bool lead = in_open;
    int k = 0;
    while (lead || mm) {
        int s, e, a0, ord;
        unsigned key = 0xffffffffu, cnt = 0, skv = 0;
        if (lead) {
            const int t0 = __ffs(~m8 & 0x1ffu) - 1;  // the incoming run ends at the first sample below threshold
            e = i0 + t0;
            s = (int)prev.x;
            a0 = i0;
            key = prev.y;
            cnt = prev.z;
            skv = prev.w;
            ord = ord_base - 1;
            mm &= ~((1u << t0) - 1u);
            lead = false;
        } else {
            const int js = __ffs(mm) - 1;
            const int je = js + __ffs(~(mm >> js)) - 1;
            s = i0 + js;
            e = i0 + je;
            a0 = max(0, s - left);
            ord = ord_base + k;
            ++k;
            mm &= ~((1u << je) - 1u);
        }
        unsigned acc = 0;
        frag(a0, min(a.lmax, e + right), key, acc);
        cnt += acc >> 21;
        skv += acc & (kOne - 1u);
        sink.store((int)(key & 0xffffu), s, e, (int)(key >> 16), cnt, o_pos ? cnt * 65535u - skv : skv, ord, src, a);
    }
    __syncwarp();  // every read of stage / carry is done
    if (act && !(peers >> lane >> 1)) {  // the last item of this owner in the round
        if (has_trail) ws.carry[src] = ws.stage[lane];
        ws.carry_n[src] = ord_base + nstarts;
    }
    __syncwarp();
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
    // (no evict-first hint here: with the 256-byte L2 promotion it made HBM re-fetch the promoted sectors, 2.6 x the reads)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

// chunk summary handed from the sample code to the block-level hit bookkeeping
struct ChunkSum {
    unsigned f, i, l;   // FULL / has samples above threshold / last sample above (0 or 1)
    unsigned cand;      // min key << 16 | chunk index (FULL chunks)
    unsigned n, sw;     // signal-side samples and their sum (FULL chunks)
};

template <bool FEAT, bool HITS, bool SGN, bool EXT22, typename Sink>
__device__ __forceinline__ void lpr_stream(const FHArgs& a, const uint16_t* pool, const LaneRec& r, int sc, const Ring& ring,
                                           int p0, int p1, int c0, int c1, FeatState& fs, WarpHits& ws, ChunkQ& cq, Sink& sink) {
    const int lane = lane_id();
    const int mis = r.mis, vtotal = r.mis + r.len;
    const int nch = (r.len > 0) ? ((vtotal + 7) >> 3) : 0;
    const int nch_max = __reduce_max_sync(kFull, nch);
    // the scan advances in blocks of four chunks; the hit bookkeeping of a block covers the chunks
    // one behind it (it needs each chunk's successor), runs are closed at a virtual chunk behind the
    // record, so the scan runs two chunks past the longest record of the warp
    const int nsteps = (nch_max > 0) ? (HITS ? nch_max + 2 : nch_max) : 0;
    const int nblk = (nsteps + 3) >> 2;
    const int bps = sc >> 2;  // blocks per segment
    const int nseg = (nblk + bps - 1) / bps;
    const bool known = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_NEGATIVE;
    const float b32 = (float)r.b_feat;
    // chunk ranges of the "plain" fast path: [plainA, plainB) minus [plainHa, plainHb)
    int plainA = 0, plainB = 0, plainHa = 0, plainHb = 0;
    if (FEAT && r.len > 0 && !known) {
        const int vlo = mis + max(c0, 1);     // first sample that has a predecessor, inside the area range
        const int vhi = mis + min(c1, r.len);  // end of the area range
        plainA = (vlo + 7) >> 3;
        plainB = vhi >> 3;
        if (p1 > p0) { plainHa = (mis + p0) >> 3; plainHb = (mis + p1 + 7) >> 3; }
    }
    const int wholeA = (mis + 7) >> 3, wholeB = vtotal >> 3;  // chunks [wholeA, wholeB) hold 8 samples of the record
    // threshold test on the offset-domain sample w: negative pulses w <= kmax, positive ones w >= 65535 - kmax
    const int wthr = r.positive ? 65535 - r.kmax : r.kmax;
    const unsigned xm16 = r.positive ? 0xffffu : 0u;
    // block-level class masks: bit 0 chunk 4b-2, bit 1 chunk 4b-1, bits 2..5 the block's own chunks
    unsigned Fm = 0u, Im = 0u, Lm = 0u;
    // aggregate of the FULL chunks in front of chunk 4b-1 and of chunk 4b
    unsigned am_key = 0xffffffffu, am_n = 0u, am_sw = 0u, a0_key = 0xffffffffu, a0_n = 0u, a0_sw = 0u;
    int qn = 0, qh = 0;  // queued items and the ring position of the first one (warp-uniform)
    bool quiet_hint = false;  // the previous block produced no item in any lane (warp-uniform)

    auto issue = [&](int s) {
        const int b = s % kNBuf;
        const int cb = s * sc - kHist;  // global chunk held by slot chunk 0
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

    // features of a plain chunk: 8 samples inside the area range, outside the height range
    auto feat_plain = [&](const uint4& q, unsigned csum) {
        unsigned f0 = __funnelshift_r(fs.prev_w, q.x, 16), f1 = __funnelshift_r(q.x, q.y, 16);
        unsigned f2 = __funnelshift_r(q.y, q.z, 16), f3 = __funnelshift_r(q.z, q.w, 16);
        unsigned d0 = __vmaxu2(q.x, f0) - __vminu2(q.x, f0), d1 = __vmaxu2(q.y, f1) - __vminu2(q.y, f1);
        unsigned d2 = __vmaxu2(q.z, f2) - __vminu2(q.z, f2), d3 = __vmaxu2(q.w, f3) - __vminu2(q.w, f3);
        fs.pdiff = __vimax3_u16x2(fs.pdiff, __vmaxu2(d0, d1), __vmaxu2(d2, d3));
        fs.isum32 += csum;
        fs.prev_w = q.w;
    };
    // class of a chunk that holds 8 samples of the record
    auto class_whole = [&](const uint4& q, int vc, unsigned csum, ChunkSum& c) {
        const unsigned mn2 = __vminu2(__vimin3_u16x2(q.x, q.y, q.z), q.w);
        const unsigned mx2 = __vmaxu2(__vimax3_u16x2(q.x, q.y, q.z), q.w);
        const int wmin = (int)min(mn2 & 0xffffu, mn2 >> 16), wmax = (int)max(mx2 & 0xffffu, mx2 >> 16);
        const int kvmin = (int)((unsigned)(r.positive ? wmax : wmin) ^ xm16), kvmax = (int)((unsigned)(r.positive ? wmin : wmax) ^ xm16);
        const int kvl = (int)((q.w >> 16) ^ xm16);
        c.i = kvmin <= r.kmax;
        c.f = kvmax <= r.kmax;
        c.l = kvl <= r.kmax;
        c.cand = ((unsigned)kvmin << 16) | (unsigned)vc;
        c.n = 8u;
        c.sw = csum;
    };
    // any chunk (record start / end, height range, known polarity, behind the record)
    auto chunk_generic = [&](const uint8_t* buf, int pos, int vc, ChunkSum& c) {
        c.f = 0u; c.i = 0u; c.l = 0u; c.cand = 0xffffffffu; c.n = 0u; c.sw = 0u;
        if (vc >= nch) return;
        uint4 q = *reinterpret_cast<const uint4*>(buf + pos * 16);
        if (SGN) { q.x ^= 0x80008000u; q.y ^= 0x80008000u; q.z ^= 0x80008000u; q.w ^= 0x80008000u; }  // int16 -> offset binary
        const int v0 = vc * 8;
        const int lo = max(mis - v0, 0), hi = min(vtotal - v0, 8);
        const int i0 = v0 - mis;
        const bool whole = lo == 0 && hi == 8;
        unsigned csum = 0u;
        if (whole && (HITS || !known)) {
            csum = __dp2a_lo(q.x, 0x0101u, 0u);
            csum = __dp2a_lo(q.y, 0x0101u, csum);
            csum = __dp2a_lo(q.z, 0x0101u, csum);
            csum = __dp2a_lo(q.w, 0x0101u, csum);
        }
        if (FEAT) {
            if (vc >= plainA && vc < plainB && (vc < plainHa || vc >= plainHb)) {
                feat_plain(q, csum);
            } else if (whole) {
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
                        fs.isum32 += csum;
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
                fs.prev_w = q.w;
            } else {
                // partial chunk (record start / end): per-sample
                const int prev_s = (int)(fs.prev_w >> 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int w = u16_at(q, j);
                    if (j >= lo && j < hi) {
                        const int i = i0 + j;
                        if (i > 0) fs.idiff = max(fs.idiff, abs(w - ((j == 0) ? prev_s : u16_at(q, j - 1))));
                        if (i >= p0 && i < p1) { fs.imin = min(fs.imin, w); fs.imax = max(fs.imax, w); }
                        if (i >= c0 && i < c1) {
                            if (!known) fs.isum32 += (unsigned)w;
                            else fs.dsum += (double)(r.positive ? __fsub_rn((float)w, b32) : __fsub_rn(b32, (float)w));
                        }
                    }
                }
                fs.prev_w = q.w;
            }
        }
        if (HITS) {
            if (whole) {
                class_whole(q, vc, csum, c);
                if (c.f && r.degen) {  // negative threshold: samples above it may lie on the far side of the baseline
                    c.n = 0u; c.sw = 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int w = u16_at(q, j);
                        const bool in = r.positive ? (w >= r.wlim + r.bias) : (w <= r.wlim + r.bias);
                        c.n += in ? 1u : 0u;
                        c.sw += in ? (unsigned)w : 0u;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int w = u16_at(q, j);
                    const bool ab = (j >= lo) && (j < hi) && (r.positive ? (w >= wthr) : (w <= wthr));
                    c.i |= ab ? 1u : 0u;
                    if (j == 7) c.l = ab ? 1u : 0u;
                }
            }
        }
    };
    // FULL-chunk aggregate in front of the next chunk
    auto fa_next = [&](const ChunkSum& c, unsigned& key, unsigned& n, unsigned& sw) {
        const unsigned k2 = min(key, c.cand), n2 = n + c.n, s2 = sw + c.sw;
        key = c.f ? k2 : 0xffffffffu;
        n = c.f ? n2 : 0u;
        sw = c.f ? s2 : 0u;
    };

    if (nseg > 0) issue(0);
    for (int s = 0; s < nseg; ++s) {
        const int b = s % kNBuf;
        if (s + 1 < nseg) issue(s + 1);
        mbar_wait(&ring.bars[b], (*ring.phase_bits >> b) & 1u);
        *ring.phase_bits ^= 1u << b;
        const uint8_t* buf = ring.slot + b * ring.buf_stride;
        const int tend = min(bps, nblk - s * bps);
        for (int t = 0; t < tend; ++t) {
            const int vc0 = s * sc + 4 * t;        // first chunk of the block
            const int pos0 = kHist + 4 * t;        // its slot position
            ChunkSum c0s, c1s, c2s, c3s;
            bool fast = vc0 >= wholeA && vc0 + 4 <= wholeB && !(HITS && r.degen);
            if (FEAT) fast = fast && vc0 >= plainA && vc0 + 4 <= plainB && (vc0 + 4 <= plainHa || vc0 >= plainHb);
            const bool warp_fast = HITS && __all_sync(kFull, fast);
            bool step_skipped = false;  // warp-uniform
            if (fast) {
                uint4 q0 = *reinterpret_cast<const uint4*>(buf + pos0 * 16);
                uint4 q1 = *reinterpret_cast<const uint4*>(buf + pos0 * 16 + 16);
                uint4 q2 = *reinterpret_cast<const uint4*>(buf + pos0 * 16 + 32);
                uint4 q3 = *reinterpret_cast<const uint4*>(buf + pos0 * 16 + 48);
                if (SGN) {
                    q0.x ^= 0x80008000u; q0.y ^= 0x80008000u; q0.z ^= 0x80008000u; q0.w ^= 0x80008000u;
                    q1.x ^= 0x80008000u; q1.y ^= 0x80008000u; q1.z ^= 0x80008000u; q1.w ^= 0x80008000u;
                    q2.x ^= 0x80008000u; q2.y ^= 0x80008000u; q2.z ^= 0x80008000u; q2.w ^= 0x80008000u;
                    q3.x ^= 0x80008000u; q3.y ^= 0x80008000u; q3.z ^= 0x80008000u; q3.w ^= 0x80008000u;
                }
                auto csum4 = [](const uint4& q) {
                    unsigned cs = __dp2a_lo(q.x, 0x0101u, 0u);
                    cs = __dp2a_lo(q.y, 0x0101u, cs);
                    cs = __dp2a_lo(q.z, 0x0101u, cs);
                    return __dp2a_lo(q.w, 0x0101u, cs);
                };
                const unsigned s0 = csum4(q0), s1 = csum4(q1), s2 = csum4(q2), s3 = csum4(q3);
                if (FEAT) { feat_plain(q0, s0); feat_plain(q1, s1); feat_plain(q2, s2); feat_plain(q3, s3); }
                if (HITS) {
                    // after a block without items, first ask whether ANY lane of the warp has a sample above threshold
                    // in these 32 samples or a run coming in: if not, the whole hit bookkeeping of the block is skipped
                    if (warp_fast && quiet_hint) {
                        const unsigned mn2 = __vimin3_u16x2(
                            __vimin3_u16x2(__vimin3_u16x2(q0.x, q0.y, q0.z), __vimin3_u16x2(q0.w, q1.x, q1.y), __vimin3_u16x2(q1.z, q1.w, q2.x)),
                            __vimin3_u16x2(__vimin3_u16x2(q2.y, q2.z, q2.w), __vimin3_u16x2(q3.x, q3.y, q3.z), q3.w), 0xffffffffu);
                        const unsigned mx2 = __vimax3_u16x2(
                            __vimax3_u16x2(__vimax3_u16x2(q0.x, q0.y, q0.z), __vimax3_u16x2(q0.w, q1.x, q1.y), __vimax3_u16x2(q1.z, q1.w, q2.x)),
                            __vimax3_u16x2(__vimax3_u16x2(q2.y, q2.z, q2.w), __vimax3_u16x2(q3.x, q3.y, q3.z), q3.w), 0u);
                        const int wmin = (int)min(mn2 & 0xffffu, mn2 >> 16), wmax = (int)max(mx2 & 0xffffu, mx2 >> 16);
                        const int kvmin = (int)((unsigned)(r.positive ? wmax : wmin) ^ xm16);
                        step_skipped = !__any_sync(kFull, kvmin <= r.kmax || (Fm | Im | Lm) != 0u);
                    }
                    if (!step_skipped) {
                        class_whole(q0, vc0, s0, c0s);
                        class_whole(q1, vc0 + 1, s1, c1s);
                        class_whole(q2, vc0 + 2, s2, c2s);
                        class_whole(q3, vc0 + 3, s3, c3s);
                    }
                }
            } else {
                chunk_generic(buf, pos0, vc0, c0s);
                chunk_generic(buf, pos0 + 1, vc0 + 1, c1s);
                chunk_generic(buf, pos0 + 2, vc0 + 2, c2s);
                chunk_generic(buf, pos0 + 3, vc0 + 3, c3s);
            }
            if (HITS && step_skipped) {
                // four QUIET chunks and nothing carried in: no items, the carries stay empty
                am_key = 0xffffffffu; am_n = 0u; am_sw = 0u;
                a0_key = 0xffffffffu; a0_n = 0u; a0_sw = 0u;
            } else if (HITS) {
                Fm |= (c0s.f << 2) | (c1s.f << 3) | (c2s.f << 4) | (c3s.f << 5);
                Im |= (c0s.i << 2) | (c1s.i << 3) | (c2s.i << 4) | (c3s.i << 5);
                Lm |= (c0s.l << 2) | (c1s.l << 3) | (c2s.l << 4) | (c3s.l << 5);
                // aggregates in front of chunks 1, 2 (items) and 3, 4 (carried to the next block)
                unsigned a1_key = a0_key, a1_n = a0_n, a1_sw = a0_sw;
                fa_next(c0s, a1_key, a1_n, a1_sw);
                unsigned a2_key = a1_key, a2_n = a1_n, a2_sw = a1_sw;
                fa_next(c1s, a2_key, a2_n, a2_sw);
                unsigned a3_key = a2_key, a3_n = a2_n, a3_sw = a2_sw;
                fa_next(c2s, a3_key, a3_n, a3_sw);
                unsigned a4_key = a3_key, a4_n = a3_n, a4_sw = a3_sw;
                fa_next(c3s, a4_key, a4_n, a4_sw);
                // chunks 4b-1 .. 4b+2 (bits 1..4): an item is a non-FULL chunk that holds, follows or precedes samples above threshold
                unsigned pend = ~Fm & (Im | (Lm << 1) | (Fm >> 1)) & 0x1eu;
                const bool last_step = (s == nseg - 1) && (t == tend - 1);
                bool first_iter = true;
                for (;;) {  // every pass queues the next pending chunk of each lane; a round runs as soon as 32 items wait
                    const bool want = pend != 0u;
                    const unsigned bal = __ballot_sync(kFull, want);
                    if (first_iter) { quiet_hint = bal == 0u; first_iter = false; }
                    if (want) {
                        const int j = __ffs(pend) - 1;  // bit j <-> chunk vc0 + j - 2
                        pend &= pend - 1u;
                        const int slot = (qh + qn + __popc(bal & ((1u << lane) - 1u))) & (kQRing - 1);
                        const unsigned sk = j == 1 ? am_key : (j == 2 ? a0_key : (j == 3 ? a1_key : a2_key));
                        const unsigned sn = j == 1 ? am_n : (j == 2 ? a0_n : (j == 3 ? a1_n : a2_n));
                        const unsigned ss = j == 1 ? am_sw : (j == 2 ? a0_sw : (j == 3 ? a1_sw : a2_sw));
                        cq.q_hdr[slot] = make_uint4((unsigned)lane | ((unsigned)(vc0 + j - 1) << 5), sk, sn, ss);
                    }
                    qn += __popc(bal);
                    if (qn >= kQCap || (bal == 0u && last_step && qn > 0)) {
                        const int take = min(qn, kQCap);
                        lpr_round<EXT22>(ws, cq, qh, take, r, a, sink);
                        qh = (qh + take) & (kQRing - 1);
                        qn -= take;
                    }
                    if (bal == 0u) break;
                }
                // carry to the next block
                Fm = (Fm >> 4) & 0x2u;
                Im = (Im >> 4) & 0x2u;
                Lm = (Lm >> 4) & 0x3u;
                am_key = a3_key; am_n = a3_n; am_sw = a3_sw;
                a0_key = a4_key; a0_n = a4_n; a0_sw = a4_sw;
            }
        }
        __syncwarp();  // every lane is done with buffer b before it is refilled
    }
}

}  // namespace wfb

#include "fused_blk.cuh"
#include "fused_f32.cuh"

namespace wfb {

// ---- the kernel --------------------------------------------------------------------------------
// BLK: block-granular items copied to a shared-memory ring (fused_blk.cuh); else chunk-granular items re-read from L2
// F32: the pool holds float32 samples (fused_f32.cuh); SGN / EXT22 / BLK do not apply then
#ifndef WFB_F32_MINBLOCKS
#define WFB_F32_MINBLOCKS 4  // float32 pools: the register-resident hit state is small, more warps hide its latencies
#endif
template <bool FEAT, bool HITS, bool SGN, bool EXT22, bool BLK, bool F32 = false>
__global__ void __launch_bounds__(kLprWarps * 32, F32 ? WFB_F32_MINBLOCKS : ((HITS && !FEAT) ? WFB_LPR_MINBLOCKS + 1 : WFB_LPR_MINBLOCKS)) lpr_kernel(const FHArgs a, const int sc, const __grid_constant__ CUtensorMap tmap,
                                                             const int have_tmap, const int ent_cap) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];  // [warp][kNBuf][lane] slots of a.slot_bytes, then the block-item rings
    __shared__ __align__(16) WarpHits s_hits[HITS ? kLprWarps : 1];
    __shared__ __align__(16) ChunkQ s_cq[(HITS && !BLK) ? kLprWarps : 1];
    __shared__ long long s_wtot[kLprWarps];
    __shared__ long long s_base[2];  // first row of the deferred tile / of this tile when it is finished at once
    __shared__ long long s_total;
    __shared__ int s_woff[kLprWarps];  // the warps' first rows relative to the tile's
    __shared__ int s_tile;
    __shared__ unsigned s_ovf_any;
    __shared__ __align__(8) unsigned long long s_bar[kLprWarps * kNBuf];

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint16_t* pool = static_cast<const uint16_t*>(a.pool);
    WarpHits& ws = s_hits[HITS ? warp : 0];
    ChunkQ& cq = s_cq[(HITS && !BLK) ? warp : 0];
    unsigned* const bq = reinterpret_cast<unsigned*>(dyn_smem + (size_t)kLprTile * kNBuf * a.slot_bytes) + (size_t)warp * kBQRing * kBQWords;
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

    // Rows are written one tile late: a tile publishes its hit count as soon as it has been streamed, the
    // block goes on to stream its next tile, and only then resolves the first tile's output offset (by then
    // the predecessors have published) and writes its rows from the pool.  The pool is double buffered and
    // lives in global memory (it never leaves L2); what phase B needs of the records stays in ws.def.
    uint4* const gp = HITS ? a.gpool + (size_t)(blockIdx.x * kLprWarps + warp) * 2 * (size_t)a.gpool_cap : nullptr;
    int cur = 0;
    bool pend = false;  // block-uniform: the previous tile's rows are still to be written
    int p_tile = 0;
    long long p_total = 0;
    const int c_bias = a.p.signed_samples ? 32768 : 0;
    auto phase_b = [&](const uint4* ent, const long long base) {  // one pooled hit per lane -> packed row
        const int used = ws.def_used;
        const unsigned ovf = ws.def_ovf;
        for (int e0 = 0; e0 < used; e0 += 32) {
            const int e = e0 + lane;
            if (e < used) {
                const uint4 h = __ldcg(ent + e);
                const int owner = (int)((h.y >> 16) & 31u);
                const DefLane d = ws.def[owner];
                const long long row = base + (long long)(d.rel_pos >> 1) + (long long)(h.y >> 21);
                if (!((ovf >> owner) & 1u) && row < a.hit_cap) {
                    RowRec rr;
                    rr.ts = d.ts; rr.rid = d.rid; rr.len = d.len; rr.dt = d.dt; rr.bc = d.bc;
                    float height, integral;
                    if constexpr (F32) {
                        height = __uint_as_float(h.z);
                        integral = __uint_as_float(h.w);
                    } else {
                        hit_values((int)(h.z & 0xffffu), h.z >> 16, h.w, (d.rel_pos & 1u) != 0, d.b, c_bias, height, integral);
                    }
                    unsigned w[15];
                    hit_row_words(w, (int)(h.x & 0xffffu), (int)(h.x >> 16), (int)(h.y & 0xffffu), height, integral, rr,
                                  a.p.left_extension, a.p.right_extension, a.lmax);
                    unsigned* dst = reinterpret_cast<unsigned*>(a.hit_out + row * kHitRowBytes);
#pragma unroll
                    for (int k = 0; k < 15; ++k) dst[k] = w[k];
                }
            }
        }
    };

    for (;;) {
        if (threadIdx.x == 0) {
            s_tile = (int)atomicAdd(a.ticket, 1u);
            s_ovf_any = 0u;
        }
        __syncthreads();
        const int tile = s_tile;
        const bool valid = tile < a.n_tiles;
        if (!valid && !(HITS && pend)) break;  // a last pass without records writes the deferred rows

        // ---------------- per-lane bookkeeping
        const long long rec = (long long)tile * kLprTile + warp * 32 + lane;
        const bool have = valid && rec < a.n;
        LaneRec r;
        r.off = 0; r.ts = 0; r.rid = 0; r.len = 0; r.dt = 1; r.pol = 0; r.bc = 0; r.clen = -1;
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
            r.clen = (int)(q2.y >> 8) - 1;
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
        if (r.clen < 0) r.clen = r.len;
        r.mis = (int)(r.off & (F32 ? 3 : 7));
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
        r.wlim = 0;
        r.xb = 0.f;
        r.degen = r.thr < 0.0;
        if (HITS && F32) {
            if (r.len > 0) r.xb = f32_threshold_bound(r.b_rec, r.thr, r.positive);
        } else if (HITS && r.len > 0) {
            r.kmax = integer_threshold_u16(r.b_rec, r.thr, r.positive, r.bias);
            const double bi = floor(r.b_rec);
            const int ib = (fabs(r.b_rec) < 2e9) ? (int)bi : (r.b_rec > 0 ? INT_MAX - 65536 : INT_MIN + 65536);
            r.wlim = r.positive ? ib + 1 : ((bi == r.b_rec) ? ib - 1 : ib);
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
        if (HITS) {
            ws.carry_n[lane] = 0;
            if (lane == 0) { ws.pool_cnt = 0; ws.ovf = 0u; }
        }
        __syncwarp();
        PoolSink psink{&ws, gp + (size_t)cur * a.gpool_cap, ent_cap, l2_policy_evict_last()};
        FeatF32 ff;
        ff.fmin = FLT_MAX; ff.fmax = -FLT_MAX; ff.fdiff = 0.f; ff.prev = 0.f; ff.xsum = 0.0; ff.dsum = 0.0; ff.nraw = 0;
        if constexpr (F32) f32_stream<FEAT, HITS>(a, static_cast<const float*>(a.pool), r, sc, ring, p0, p1, c0, c1, ff, ws, psink);
        else if constexpr (HITS && BLK) blk_stream<FEAT, SGN>(a, pool, r, sc, ring, p0, p1, c0, c1, fs, ws, bq, psink);
        else lpr_stream<FEAT, HITS, SGN, EXT22>(a, pool, r, sc, ring, p0, p1, c0, c1, fs, ws, cq, psink);

        // ---------------- features of my record
        if (FEAT && have) {
            const float b32 = (float)r.b_feat;
            float height = 0.f, amp = 0.f, area = 0.f;
            float mad = 0.f;
            if constexpr (F32) {  // as fused_features_hits.cu, float32 samples (basic_features.py:142-278)
                mad = ff.fdiff;
                if (p1 > p0) {
                    if (!known) {
                        height = rawpos ? (float)__dsub_rn((double)ff.fmax, r.b_feat) : (float)__dsub_rn(r.b_feat, (double)ff.fmin);
                        amp = (float)__dsub_rn((double)ff.fmax, (double)ff.fmin);
                    } else {
                        const float smax = r.positive ? __fsub_rn(ff.fmax, b32) : __fsub_rn(b32, ff.fmin);
                        const float smin = r.positive ? __fsub_rn(ff.fmin, b32) : __fsub_rn(b32, ff.fmax);
                        height = smax;
                        amp = (float)__dsub_rn((double)smax, (double)smin);
                    }
                }
                if (c1 > c0) {
                    double ds = ff.dsum;
                    if (ff.nraw) ds += rawpos ? __dsub_rn(ff.xsum, __dmul_rn((double)ff.nraw, r.b_feat)) : __dsub_rn(__dmul_rn((double)ff.nraw, r.b_feat), ff.xsum);
                    area = (float)ds;
                }
            } else {
            mad = (float)max(fs.idiff, (int)max(fs.pdiff & 0xffffu, fs.pdiff >> 16));
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
        __syncwarp();
        const int my_cnt = ws.carry_n[lane];
        if (a.hit_counts != nullptr && have) a.hit_counts[rec] = my_cnt;

        int incl = my_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_wtot[warp] = incl;
        if (lane == 0 && ws.ovf != 0u) s_ovf_any = 1u;
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
            if (lane < kLprWarps) s_woff[lane] = (int)(wincl - c);
            if (valid) tile_publish(a, tile, total);  // never waits
            if (pend) {
                const long long excl = tile_resolve(a, p_tile, p_total);
                if (lane == 0) s_base[0] = excl;
            }
            if (valid && s_ovf_any != 0u) {  // a pool overflowed: this tile is finished before the next one starts
                const long long excl = tile_resolve(a, tile, total);
                if (lane == 0) s_base[1] = excl;
            }
            if (lane == 0) s_total = total;
        }
        __syncthreads();

        // ---------------- phase B of the PREVIOUS tile, from the other half of the pool
        if (pend) phase_b(gp + (size_t)(cur ^ 1) * a.gpool_cap, s_base[0]);
        __syncwarp();
        // ---------------- this tile's row data replaces it
        const int rel = s_woff[warp] + (incl - my_cnt);
        {
            DefLane d;
            d.ts = r.ts; d.rid = r.rid; d.b = r.b_rec; d.len = r.clen; d.dt = r.dt; d.bc = r.bc;
            d.rel_pos = ((unsigned)rel << 1) | (r.positive ? 1u : 0u);
            ws.def[lane] = d;
            if (lane == 0) {
                ws.def_used = min(ws.pool_cnt, ent_cap);
                ws.def_ovf = ws.ovf;
            }
        }
        __syncwarp();
        if (valid && s_ovf_any != 0u) {
            const long long base = s_base[1];
            phase_b(gp + (size_t)cur * a.gpool_cap, base);
            const unsigned ovf = ws.def_ovf;
            if (ovf) {  // some records' hits did not all fit the pool: stream those again, rows go straight out
                __syncwarp();
                ws.carry_n[lane] = 0;
                __syncwarp();
                FeatState fs2 = fs;
                DirectSink dsink;
                dsink.my_row0 = base + rel;
                dsink.my_active = (ovf >> lane) & 1u;
                if constexpr (F32) f32_stream<false, true>(a, static_cast<const float*>(a.pool), r, sc, ring, 0, 0, 0, 0, ff, ws, dsink);
                else if constexpr (BLK) blk_stream<false, SGN>(a, pool, r, sc, ring, 0, 0, 0, 0, fs2, ws, bq, dsink);
                else lpr_stream<false, true, SGN, EXT22>(a, pool, r, sc, ring, 0, 0, 0, 0, fs2, ws, cq, dsink);
            }
            pend = false;
        } else {
            pend = valid;
            p_tile = tile;
            p_total = s_total;
        }
        cur ^= 1;
        __syncthreads();  // slots, s_tile and the block totals are reused by the next tile
        if (!valid) break;
    }
}

int launch_lpr(FHArgs a, int flags, cudaStream_t st) {
    const char* force = getenv("WFB_FUSED_VARIANT");
    if (force && (!strcmp(force, "global") || !strcmp(force, "staged"))) return 1;
    if (a.lmax >= 65536 - 16) return 1;  // positions are packed in 16 bits
    if (a.p.left_extension > kMaxExt || a.p.right_extension > kMaxExt) return 1;
    const bool f32 = a.p.pool_is_f32 != 0;
    if (f32) {
        // float32 pools: the register-resident run / tail state of fused_f32.cuh needs extensions of at most two samples;
        // WFB_F32_IMPL=warp keeps the warp-per-record kernel
        const char* fi = getenv("WFB_F32_IMPL");
        if (fi && !strcmp(fi, "warp")) return 1;
        if ((flags & WFB_DO_HITS) && (a.p.left_extension > kBQMaxExt || a.p.right_extension > kBQMaxExt)) return 1;
    }
    // segment length in chunks: kHist history chunks + sc new chunks per slot
    // a multiple of the 4-chunk scan block.  Measured per mode (profiles/README.md): features + hits 12 chunks at three
    // blocks per SM; hits only 8 chunks, whose smaller slots and 128 registers let a fourth block in; features only 16
    const bool want_f = flags & WFB_DO_FEATURES, want_h = flags & WFB_DO_HITS;
    // block-granular items (fused_blk.cuh): extensions of at most two samples; WFB_LPR_IMPL=chunk keeps the chunk-granular variant
    const char* impl = getenv("WFB_LPR_IMPL");
    const bool blk = !f32 && want_h && a.p.left_extension <= kBQMaxExt && a.p.right_extension <= kBQMaxExt && !(impl && !strcmp(impl, "chunk"));
    int sc = blk ? 8 : ((want_f && want_h) ? 12 : (want_h ? 8 : 16));
    if (f32) sc = 8;  // 32 samples per segment: four blocks (16 warps) per SM fit
    if (const char* e = getenv("WFB_LPR_SC")) sc = std::max(4, std::min(28, atoi(e) & ~3));
    int ent_cap = a.gpool_cap;  // WFB_LPR_POOL shrinks the per-warp hit pool (exercises the overflow path in tests)
    if (const char* e = getenv("WFB_LPR_POOL")) ent_cap = std::max(0, std::min(a.gpool_cap, atoi(e)));
    int slot_chunks = sc + kHist;
    if ((slot_chunks & 1) == 0) ++slot_chunks;  // odd multiple of 16 bytes: conflict-free LDS.128
    a.slot_bytes = slot_chunks * 16;
    const size_t dyn = (size_t)kLprTile * kNBuf * a.slot_bytes + (blk ? (size_t)kLprWarps * kBQRing * kBQWords * 4 : 0);
    a.n_tiles = (int)((a.n + kLprTile - 1) / kLprTile);
    const bool f = flags & WFB_DO_FEATURES, h = flags & WFB_DO_HITS;
    // tensor map over the pool seen as rows of lmax samples: one box = 32 records x one slot
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int have_tmap = 0;
    const char* no2d = getenv("WFB_LPR_NO_TMAP");
    const int per16 = f32 ? 4 : 8;  // samples per 16-byte chunk
    if (!(no2d && no2d[0] == '1') && a.lmax > 0 && a.lmax % per16 == 0 && a.pool_len >= a.lmax && a.n < (1ll << 31)) {
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
            cuuint64_t strides[1] = {(cuuint64_t)a.lmax * (f32 ? 4 : 2)};
            cuuint32_t box[2] = {(cuuint32_t)(a.slot_bytes / (f32 ? 4 : 2)), 32};
            cuuint32_t estr[2] = {1, 1};
            CUresult cr = encode(&tmap, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(a.pool), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
        if (h) grid = std::min(grid, a.gpool_warps / kLprWarps);
        kern<<<grid, kLprWarps * 32, dyn, st>>>(a, sc, tmap, have_tmap, ent_cap);
        WFB_CUDA(cudaGetLastError());
        return WFB_OK;
    };
    const bool e22 = h && a.p.left_extension == 2 && a.p.right_extension == 2;  // the defaults (hit_finder.py:104-105)
    if (f32) {
        if (f && h) return go(lpr_kernel<true, true, false, false, false, true>);
        if (f) return go(lpr_kernel<true, false, false, false, false, true>);
        return go(lpr_kernel<false, true, false, false, false, true>);
    }
    if (blk) {
        if (a.p.signed_samples) return f ? go(lpr_kernel<true, true, true, false, true>) : go(lpr_kernel<false, true, true, false, true>);
        return f ? go(lpr_kernel<true, true, false, false, true>) : go(lpr_kernel<false, true, false, false, true>);
    }
    if (a.p.signed_samples) {
        if (f && h) return e22 ? go(lpr_kernel<true, true, true, true, false>) : go(lpr_kernel<true, true, true, false, false>);
        if (f) return go(lpr_kernel<true, false, true, false, false>);
        return e22 ? go(lpr_kernel<false, true, true, true, false>) : go(lpr_kernel<false, true, true, false, false>);
    }
    if (f && h) return e22 ? go(lpr_kernel<true, true, false, true, false>) : go(lpr_kernel<true, true, false, false, false>);
    if (f) return go(lpr_kernel<true, false, false, false, false>);
    return e22 ? go(lpr_kernel<false, true, false, true, false>) : go(lpr_kernel<false, true, false, false, false>);
}

}  // namespace wfb
