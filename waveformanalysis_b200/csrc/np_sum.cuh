// numpy's float64 pairwise summation restated for the device, so that sums the reference takes with
// np.sum come out bit-identical (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum).
#pragma once
#include "common.cuh"

namespace wfb {

// numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum), iterative
__device__ __forceinline__ double np_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float np_add(float a, float b) { return __fadd_rn(a, b); }

// A = accumulator / element type (np.sum keeps float32 arrays in float32)
template <typename A, typename Src>
__device__ A numpy_pairwise_sum_t(const Src& x, int n) {
    // explicit stack of (offset, length) blocks; the recursion halves until length <= 128
    int st_off[40], st_len[40];
    A st_val[40];
    int st_state[40];  // 0: to expand, 1: left done (value holds left sum)
    int sp = 0;
    st_off[0] = 0; st_len[0] = n; st_state[0] = 0;
    A ret = 0;
    bool have_ret = false;
    while (sp >= 0) {
        const int off = st_off[sp], len = st_len[sp];
        if (st_state[sp] == 0) {
            if (len < 8) {
                A res = 0;
                for (int i = 0; i < len; ++i) res = np_add(res, (A)x(off + i));
                ret = res; have_ret = true; --sp;
            } else if (len <= 128) {
                A r[8];
                for (int k = 0; k < 8; ++k) r[k] = (A)x(off + k);
                int i;
                for (i = 8; i < len - (len % 8); i += 8)
                    for (int k = 0; k < 8; ++k) r[k] = np_add(r[k], (A)x(off + i + k));
                A res = np_add(np_add(np_add(r[0], r[1]), np_add(r[2], r[3])), np_add(np_add(r[4], r[5]), np_add(r[6], r[7])));
                for (; i < len; ++i) res = np_add(res, (A)x(off + i));
                ret = res; have_ret = true; --sp;
            } else {
                int n2 = len / 2;
                n2 -= n2 % 8;
                st_state[sp] = 1;
                st_val[sp] = 0;
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
            ret = np_add(st_val[sp], ret);
            --sp;
        }
    }
    (void)have_ret;
    return ret;
}

template <typename Src>
__device__ double numpy_pairwise_sum(const Src& x, int n) {
    return numpy_pairwise_sum_t<double>(x, n);
}

}  // namespace wfb
