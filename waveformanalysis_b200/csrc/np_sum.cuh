// numpy's float64 pairwise summation restated for the device, so that sums the reference takes with
// np.sum come out bit-identical (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum).
#pragma once
#include "common.cuh"

namespace wfb {

// numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum), iterative
template <typename Src>
__device__ double numpy_pairwise_sum(const Src& x, int n) {
    // explicit stack of (offset, length) blocks; the recursion halves until length <= 128
    int st_off[40], st_len[40];
    double st_val[40];
    int st_state[40];  // 0: to expand, 1: left done (value holds left sum)
    int sp = 0;
    st_off[0] = 0; st_len[0] = n; st_state[0] = 0;
    double ret = 0.0;
    bool have_ret = false;
    while (sp >= 0) {
        const int off = st_off[sp], len = st_len[sp];
        if (st_state[sp] == 0) {
            if (len < 8) {
                double res = 0.0;
                for (int i = 0; i < len; ++i) res = __dadd_rn(res, x(off + i));
                ret = res; have_ret = true; --sp;
            } else if (len <= 128) {
                double r[8];
                for (int k = 0; k < 8; ++k) r[k] = x(off + k);
                int i;
                for (i = 8; i < len - (len % 8); i += 8)
                    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], x(off + i + k));
                double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                       __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (; i < len; ++i) res = __dadd_rn(res, x(off + i));
                ret = res; have_ret = true; --sp;
            } else {
                int n2 = len / 2;
                n2 -= n2 % 8;
                st_state[sp] = 1;
                st_val[sp] = 0.0;
                // push left
                ++sp;
                st_off[sp] = off; st_len[sp] = n2; st_state[sp] = 0;
            }
        } else if (st_state[sp] == 1) {
            // left returned in ret; remember it and descend right
            int n2 = len / 2;
            n2 -= n2 % 8;
            st_val[sp] = ret;
            st_state[sp] = 2;
            ++sp;
            st_off[sp] = off + n2; st_len[sp] = len - n2; st_state[sp] = 0;
        } else {
            ret = __dadd_rn(st_val[sp], ret);
            --sp;
        }
    }
    (void)have_ret;
    return ret;
}

}  // namespace wfb
