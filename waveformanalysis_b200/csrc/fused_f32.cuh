// float32 pools (wave_pool_filtered) in the lane-per-record kernel (included by fused_lpr.cu after the shared
// structs; uses LaneRec, Ring, WarpHits and the sinks defined there).
//
// Every lane streams ITS record through the TMA slot ring, four samples (one 16-byte chunk) at a time, and keeps the
// whole hit state in registers - no shuffles, no warp reductions, no item queue:
//
//   * above threshold <=> x <= xb (negative pulses) / x >= xb (positive): the per-record float32 bound that is exactly
//     equivalent to the reference's float64 test b - x >= thr (fused_common.cuh: f32_threshold_bound);
//   * ONE run state {start, end, best sample + position, sum of max(sig, 0)}: at the run's first sample the one or two
//     samples in front of it (left extension) are folded in from registers; after its end it keeps collecting the
//     `right` samples behind it and is then emitted.  Hit windows of neighbouring runs overlap (both see the samples):
//     with extensions of at most two samples a new run can only start on the LAST tail sample of the run before, which
//     then takes that sample and is emitted on the spot - so no second state is needed;
//   * samples between the record end and the padded width (hit_finder.py:364; true zeros, records_view.py:189) are
//     handled after the loop.
//
// First maximum of sig = first minimum (negative pulses) / maximum (positive) of the raw sample: sig is monotone in x and
// the samples are visited in index order with strict comparisons.  height = float64 sig of that sample; integral =
// sum(max(sig, 0)) computed as count * b - sum(x) over the samples on the signal side (hit_finder.py:369-381; float64,
// compared with rel 1e-5).  Features as in the warp-per-record kernel
// (fused_features_hits.cu, float32 fast path): |diff| in float32, raw extremes with the monotone float32 signal transform
// applied once per record, area as one float64 add per sample.
#pragma once

namespace wfb {

struct FeatF32 {  // per-lane feature accumulators of a float32 record
    float fmin, fmax, fdiff, prev;
    double xsum, dsum;
    int nraw;
};

template <bool FEAT, bool HITS, typename Sink>
__device__ __forceinline__ void f32_stream(const FHArgs& a, const float* pool, const LaneRec& r, int sc, const Ring& ring,
                                           int p0, int p1, int c0, int c1, FeatF32& fs, WarpHits& ws, Sink& sink) {
    const int lane = lane_id();
    const int mis = r.mis, vtotal = r.mis + r.len;
    const int nch = (r.len > 0) ? ((vtotal + 3) >> 2) : 0;
    const int nch_max = __reduce_max_sync(kFull, nch);
    const int nseg = (nch_max + sc - 1) / sc;
    const bool known = r.pol == WFB_POL_POSITIVE || r.pol == WFB_POL_NEGATIVE;
    const bool positive = r.positive;
    const float b32 = (float)r.b_feat;
    const double b = r.b_rec;
    const float xb = r.xb;
    const int left = a.p.left_extension, right = a.p.right_extension;
    // Polarity is folded into the sample once: y = -x for positive pulses, so that "above threshold" is y <= yb, the first
    // maximum of the signal the first minimum of y, and sig = x - b = (-b) - (-x) = bn - y (negation is exact).
    const unsigned sgn = positive ? 0x80000000u : 0u;
    const float yb = __uint_as_float(__float_as_uint(xb) ^ sgn);
    const double bn = positive ? -b : b;

    // hit state of the lane: ONE run.  While it is open its samples are folded in; after its end `trem` more samples (the
    // right extension) are, then it is emitted.  A run that starts inside that tail (only possible at the tail's last
    // sample, extensions <= 2) first gives the sample to the old run, emits it, and starts over.
    // integral = sum of max(sig, 0) over the window = cnt * bn - sum(y) over the samples on the signal side of the baseline,
    // sig > 0 <=> (double)y < bn <=> y <= ybn (the largest float below bn): one float compare, one float64 add per sample
    float ybn = (float)bn;
    if ((double)ybn >= bn) ybn = f32_step(ybn, false);
    // st: -1 the run is open, k > 0 the run has ended and k samples of its right extension are still to come, 0 idle.
    // best sample / count / sum are folded on EVERY sample: while idle they collect garbage that the next run start
    // resets, so the per-sample code needs no "inside a window" predicate.
    int st = 0;
    int rs = 0, re = 0, bpos = 0, nh = 0, cnt = 0;
    float by = INFINITY;
    double acc = 0.0;  // sum of the signal-side samples y
    float y1 = 0.f, y2 = 0.f;  // the two samples in front of the current one
    if (HITS) sink.prepare(lane, r);

    auto fold = [&](float y, int i) {
        const bool lt = y < by;
        by = lt ? y : by;
        bpos = lt ? i : bpos;
        const bool sg = y <= ybn;
        cnt += sg ? 1 : 0;
        acc += (double)(sg ? y : 0.f);
    };
    auto emit = [&]() {
        const double integ = __dsub_rn(__dmul_rn((double)cnt, bn), acc);
        sink.store_f32(bpos, rs, re, (float)__dsub_rn(bn, (double)by), (float)integ, nh, lane, a);
        ++nh;
    };
    // one sample of the record (0 <= i < len) through the hit state
    auto hit_step = [&](float x, int i) {
        const float y = __uint_as_float(__float_as_uint(x) ^ sgn);
        const bool in = y <= yb;
        if (in != (st < 0)) {  // an edge (rare per lane)
            if (in) {          // a run starts
                if (st > 0) {  // ... on the last tail sample of the run before: that run takes the sample and is done
                    fold(y, i);
                    emit();
                }
                st = -1;
                rs = i;
                by = INFINITY;
                acc = 0.0;
                cnt = 0;
                // the left extension, in index order (strict comparisons keep the FIRST maximum)
                if (left >= 2 && i >= 2) fold(y2, i - 2);
                if (left >= 1 && i >= 1) fold(y1, i - 1);
            } else {           // the run ends in front of this sample, which is the first of the right extension
                re = i;
                st = right;
                if (right == 0) emit();
            }
        }
        fold(y, i);
        const bool fire = st == 1;  // the last tail sample
        st -= (st > 0) ? 1 : 0;
        if (fire) emit();
        y2 = y1;
        y1 = y;
    };
    auto feat_step = [&](float x, int i) {
        if (i > 0) fs.fdiff = fmaxf(fs.fdiff, fabsf(__fsub_rn(x, fs.prev)));
        fs.prev = x;
        if (i >= p0 && i < p1) { fs.fmin = fminf(fs.fmin, x); fs.fmax = fmaxf(fs.fmax, x); }
        if (i >= c0 && i < c1) {
            if (!known) { fs.xsum += (double)x; ++fs.nraw; }
            else fs.dsum += (double)(positive ? __fsub_rn(x, b32) : __fsub_rn(b32, x));
        }
    };

    auto issue = [&](int s) {
        const int bsel = s % kNBuf;
        const int cb = s * sc;
        const int clo = cb, chi = min(nch, (s + 1) * sc);
        const unsigned bytes = chi > clo ? (unsigned)(chi - clo) * 16u : 0u;
        fence_proxy_async();
        if (ring.use2d) {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(&ring.bars[bsel], (unsigned)ring.buf_stride);
                tma_tensor2d_g2s(ring.slot + bsel * ring.buf_stride, ring.tmap, cb * 4, ring.row0, &ring.bars[bsel]);
            }
            return;
        }
        const unsigned total = __reduce_add_sync(kFull, bytes);
        if (lane == 0) {
            if (total) mbar_arrive_expect_tx(&ring.bars[bsel], total);
            else mbar_arrive(&ring.bars[bsel]);
        }
        __syncwarp();
        if (bytes) tma_bulk_g2s(ring.slot + bsel * ring.buf_stride, pool + (r.off - mis) + (long long)clo * 4, bytes, &ring.bars[bsel]);
    };

    if (nseg > 0) issue(0);
    for (int s = 0; s < nseg; ++s) {
        const int bsel = s % kNBuf;
        if (s + 1 < nseg) issue(s + 1);
        mbar_wait(&ring.bars[bsel], (*ring.phase_bits >> bsel) & 1u);
        *ring.phase_bits ^= 1u << bsel;
        const uint8_t* buf = ring.slot + bsel * ring.buf_stride;
        const int tend = min(sc, nch - s * sc);  // this lane's chunks in the segment
        for (int t = 0; t < tend; ++t) {
            const int vc = s * sc + t;
            const uint4 q = *reinterpret_cast<const uint4*>(buf + t * 16);
            const float x[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
            const int i0 = vc * 4 - mis;
            if (i0 >= 0 && i0 + 4 <= r.len) {  // four samples of the record
                if (FEAT) {
                    const bool hall = p0 <= i0 && p1 >= i0 + 4, hnone = p1 <= p0 || p1 <= i0 || p0 >= i0 + 4;
                    const bool call = c0 <= i0 && c1 >= i0 + 4, cnone = c1 <= c0 || c1 <= i0 || c0 >= i0 + 4;
                    if (i0 > 0 && (hall || hnone) && (call || cnone)) {
                        const float d = fmaxf(fmaxf(fabsf(__fsub_rn(x[0], fs.prev)), fabsf(__fsub_rn(x[1], x[0]))),
                                              fmaxf(fabsf(__fsub_rn(x[2], x[1])), fabsf(__fsub_rn(x[3], x[2]))));
                        fs.fdiff = fmaxf(fs.fdiff, d);
                        fs.prev = x[3];
                        if (hall) {
                            fs.fmin = fminf(fs.fmin, fminf(fminf(x[0], x[1]), fminf(x[2], x[3])));
                            fs.fmax = fmaxf(fs.fmax, fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
                        }
                        if (call) {
                            if (!known) {
                                fs.xsum += (double)x[0]; fs.xsum += (double)x[1]; fs.xsum += (double)x[2]; fs.xsum += (double)x[3];
                                fs.nraw += 4;
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) fs.dsum += (double)(positive ? __fsub_rn(x[j], b32) : __fsub_rn(b32, x[j]));
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) feat_step(x[j], i0 + j);
                    }
                }
                if (HITS) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) hit_step(x[j], i0 + j);
                }
            } else {  // record start / end
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + j;
                    if (i >= 0 && i < r.len) {
                        if (FEAT) feat_step(x[j], i);
                        if (HITS) hit_step(x[j], i);
                    }
                }
            }
        }
        __syncwarp();  // every lane is done with the buffer before it is refilled
    }
    if (HITS) {
        // behind the record: padding samples (true zeros) up to the padded width belong to the window of a run that
        // ends within `right` samples of the record end
        if (r.len > 0 && st != 0) {
            if (st < 0) { re = r.len; st = right; }
            const float yp = __uint_as_float(sgn);  // +-0
            for (int idx = r.len; st > 0 && idx < a.lmax; ++idx, --st) fold(yp, idx);
            emit();
        }
        ws.carry_n[lane] = nh;
    }
}

}  // namespace wfb
