// nb2_selection.cuh - index arithmetic of the ArticulationView copy kernels (include/newton_b200.h, nb2_view_layout).
// Host + device so that the CPU test-suite can run exactly this arithmetic against the oracle's NumPy restatement
// (oracle/selection.py) without a GPU; the product only ever calls it from the kernels of nb2_selection.cu.
#pragma once
#include <stdint.h>

#include "../../include/newton_b200.h"
#include "nb2_math.cuh"

namespace nb2 {

struct ViewElem {
    int w, a;          // (world, articulation) row of the view
    long long word;    // word offset inside the attribute array
};

// Flat index i over the contiguous [W, A, K, T] values array -> its (w, a) and the attribute word it mirrors.
// `idx` is the device copy of layout.indices (may be NULL).
// I = unsigned (arrays below 2^32 words: 32-bit divisions) or long long.
template <typename I>
NB2_DEV ViewElem view_elem(const nb2_view_layout& L, const int32_t* idx, I i) {
    const I T = (I)L.row_words, K = (I)L.value_count, A = (I)L.count_per_world;
    const I v = i / T;  // value index in [0, W*A*K)
    const int t = (int)(i - v * T);
    const I r = v / K;  // row = w*A + a
    const int k = (int)(v - r * K);
    ViewElem e;
    e.w = (int)(r / A);
    e.a = (int)(r - (I)e.w * A);
    const int sel = idx ? idx[k] : L.slice_start + k;
    e.word = ((long long)L.offset + (long long)e.w * L.stride_between_worlds + (long long)e.a * L.stride_within_worlds + sel) * T + t;
    return e;
}

NB2_DEV bool view_selected(const nb2_view_layout& L, const uint8_t* mask, int mask_ndim, const ViewElem& e) {
    if (mask_ndim == 1) return mask[e.w] != 0;
    if (mask_ndim == 2) return mask[(long long)e.w * L.count_per_world + e.a] != 0;
    return true;
}

}  // namespace nb2
