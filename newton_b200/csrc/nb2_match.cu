// nb2_match.cu - frame-to-frame contact matching, CollisionPipeline(contact_matching="latest")
// (reference geometry/contact_match.py: _match_contacts_kernel :266-354, _resolve_claims_kernel :357-390,
// _save_sorted_state_kernel :442-477; call sites sim/collide.py:2033-2137).
//
// Runs on the exported, key-sorted `Contacts` arrays (matching implies deterministic=True upstream, collide.py:1269-1271): for
// each new contact the previous frame's contacts of the same shape pair are found by binary search on the sorted keys, the closest
// one (world-space midpoint) whose normal passes the dot threshold is claimed with one 64-bit atomicMin (distance, then key),
// and losers of a claim become MATCH_BROKEN.  The history (keys, midpoints, normals) lives with the nb2_model.
#include <cstdint>

#include "nb2_internal.cuh"
#include "nb2_math.cuh"

namespace nb2 {

enum { MATCH_NOT_FOUND = -1, MATCH_BROKEN = -2 };
static const long long CLAIM_SENTINEL = 0x7FFFFFFFFFFFFFFFll;

__device__ __forceinline__ long long pair_prefix(int s0, int s1) {  // make_contact_sort_key without the sub key (contact_data.py:59-87)
    return ((long long)(s0 & 0xFFFFF) << 43) | ((long long)(s1 & 0xFFFFF) << 23);
}
__device__ __forceinline__ int lower_bound64(const long long* keys, int lo, int hi, long long v) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ unsigned float_flip(float f) {
    const unsigned i = __float_as_uint(f);
    const unsigned mask = (unsigned)(-(int)(i >> 31)) | 0x80000000u;
    return i ^ mask;
}
__device__ __forceinline__ V3 contact_midpoint(const nb2_model_desc& d, const nb2_contacts_view& c, const float* body_q, int i) {
    V3 p0 = ld3(c.point0 + 3 * i), p1 = ld3(c.point1 + 3 * i);
    const int b0 = d.shape_body[c.shape0[i]], b1 = d.shape_body[c.shape1[i]];
    if (b0 != -1) p0 = xpoint(ldx(body_q + 7 * b0), p0);
    if (b1 != -1) p1 = xpoint(ldx(body_q + 7 * b1), p1);
    return 0.5f * (p0 + p1);
}

// sort keys of the new (sorted) contacts: the sub key is the contact's position inside its pair's run
__global__ void __launch_bounds__(256) match_keys_kernel(nb2_contacts_view c, long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(c.rigid_contact_count[0], c.rigid_contact_max);
    if (i >= n) return;
    const long long prefix = pair_prefix(c.shape0[i], c.shape1[i]);
    int lo = 0, hi = i;  // first contact of the run: shapes are sorted, so search on (shape0, shape1) directly
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pair_prefix(c.shape0[mid], c.shape1[mid]) < prefix) lo = mid + 1;
        else hi = mid;
    }
    keys[i] = prefix | (long long)((i - lo) & 0x7FFFFF);
}

__global__ void __launch_bounds__(256) match_contacts_kernel(nb2_model_desc d, nb2_contacts_view c, const float* __restrict__ body_q,
                                                             const long long* __restrict__ new_keys, const long long* __restrict__ prev_keys,
                                                             const float* __restrict__ prev_pos, const float* __restrict__ prev_normal,
                                                             const int* __restrict__ prev_count, long long* prev_claim,
                                                             const uint8_t* __restrict__ reset_mask, float pos_threshold_sq, float normal_dot_threshold,
                                                             int* __restrict__ match_index) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= c.rigid_contact_max) return;
    const int n_new = min(c.rigid_contact_count[0], c.rigid_contact_max);
    if (tid >= n_new) {
        match_index[tid] = MATCH_NOT_FOUND;
        return;
    }
    const int n_old = prev_count[0];
    if (n_old == 0) {
        match_index[tid] = MATCH_NOT_FOUND;
        return;
    }
    if (reset_mask) {  // reset_world_selected (core/reset.py:14-18) on both shapes
        const int w0 = d.shape_world[c.shape0[tid]], w1 = d.shape_world[c.shape1[tid]];
        const bool r0 = (w0 >= 0 && w0 < d.world_count) ? reset_mask[w0] != 0 : (w0 == -1 && reset_mask[d.world_count] != 0);
        const bool r1 = (w1 >= 0 && w1 < d.world_count) ? reset_mask[w1] != 0 : (w1 == -1 && reset_mask[d.world_count] != 0);
        if (r0 || r1) {
            match_index[tid] = MATCH_NOT_FOUND;
            return;
        }
    }
    const long long target = new_keys[tid];
    const V3 pos = contact_midpoint(d, c, body_q, tid);
    const V3 nrm = ld3(c.normal + 3 * tid);
    const long long prefix = target & ~0x7FFFFFll, pair_end = prefix + 0x800000ll;
    const int lo = lower_bound64(prev_keys, 0, n_old, prefix), hi = lower_bound64(prev_keys, lo, n_old, pair_end);
    if (lo >= hi) {
        match_index[tid] = MATCH_NOT_FOUND;
        return;
    }
    int best = -1;
    float best_d = pos_threshold_sq;
    for (int k = lo; k < hi; ++k) {
        const V3 diff = pos - ld3(prev_pos + 3 * k);
        const float dsq = dot(diff, diff);
        if (dsq <= best_d && dot(nrm, ld3(prev_normal + 3 * k)) >= normal_dot_threshold) {
            best_d = dsq;
            best = k;
        }
    }
    if (best >= 0) {
        match_index[tid] = best;
        const long long claim = ((long long)float_flip(best_d) << 32) | (target & 0xFFFFFFFFll);
        atomicMin(prev_claim + best, claim);
    } else {
        match_index[tid] = MATCH_BROKEN;
    }
}

__global__ void __launch_bounds__(256) match_resolve_kernel(nb2_contacts_view c, const long long* __restrict__ new_keys,
                                                            const long long* __restrict__ prev_claim, int* __restrict__ match_index) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= min(c.rigid_contact_count[0], c.rigid_contact_max)) return;
    const int cand = match_index[tid];
    if (cand < 0) return;
    if ((prev_claim[cand] & 0xFFFFFFFFll) != (new_keys[tid] & 0xFFFFFFFFll)) match_index[tid] = MATCH_BROKEN;
}

__global__ void __launch_bounds__(256) match_save_kernel(nb2_model_desc d, nb2_contacts_view c, const float* __restrict__ body_q,
                                                         const long long* __restrict__ new_keys, long long* __restrict__ prev_keys,
                                                         float* __restrict__ prev_pos, float* __restrict__ prev_normal, long long* __restrict__ prev_claim,
                                                         int* __restrict__ prev_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(c.rigid_contact_count[0], c.rigid_contact_max);
    if (i == 0) prev_count[0] = n;
    if (i >= n) return;
    prev_keys[i] = new_keys[i];
    prev_claim[i] = CLAIM_SENTINEL;
    st3(prev_pos + 3 * i, contact_midpoint(d, c, body_q, i));
    st3(prev_normal + 3 * i, ld3(c.normal + 3 * i));
}

}  // namespace nb2

using namespace nb2;

extern "C" nb2_status nb2_contacts_match(nb2_model* m, const float* body_q, const nb2_contacts_view* c, int32_t* match_index, float pos_threshold,
                                         float normal_dot_threshold, const uint8_t* reset_world_mask, int32_t reset_all, void* cuda_stream) {
    if (!m || !c || !match_index || !c->rigid_contact_count || !c->shape0 || !c->shape1 || !c->point0 || !c->point1 || !c->normal ||
        (!body_q && m->dev.d.body_count > 0)) {
        set_error("nb2_contacts_match: NULL argument");
        return NB2_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    const int cap = c->rigid_contact_max;
    nb2_status st = NB2_OK;
    do {
        if (cap > m->match_capacity) {  // history buffers (not capturable: the first call sizes them, like the sort scratch)
            for (void* p : {(void*)m->match_new_keys, (void*)m->match_prev_keys, (void*)m->match_prev_claim, (void*)m->match_prev_pos,
                            (void*)m->match_prev_normal})
                if (p) cudaFree(p);
            void* p = nullptr;
#define NB2_MATCH_ALLOC(field, bytes)                       \
    if (cudaMalloc(&p, (bytes)) != cudaSuccess) {           \
        set_error("nb2_contacts_match: out of device memory"); \
        st = NB2_ERR_CUDA;                                  \
        break;                                              \
    }                                                       \
    field = static_cast<decltype(field)>(p);
            NB2_MATCH_ALLOC(m->match_new_keys, size_t(cap) * 8)
            NB2_MATCH_ALLOC(m->match_prev_keys, size_t(cap) * 8)
            NB2_MATCH_ALLOC(m->match_prev_claim, size_t(cap) * 8)
            NB2_MATCH_ALLOC(m->match_prev_pos, size_t(cap) * 12)
            NB2_MATCH_ALLOC(m->match_prev_normal, size_t(cap) * 12)
#undef NB2_MATCH_ALLOC
            if (!m->match_prev_count) {
                if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) {
                    st = NB2_ERR_CUDA;
                    break;
                }
                m->match_prev_count = static_cast<int*>(p);
            }
            cudaMemsetAsync(m->match_prev_count, 0, sizeof(int), s);
            m->match_capacity = cap;
        }
        if (reset_all) cudaMemsetAsync(m->match_prev_count, 0, sizeof(int), s);
        if (cap > 0) {
            const int blocks = (cap + 255) / 256;
            match_keys_kernel<<<blocks, 256, 0, s>>>(*c, m->match_new_keys);
            match_contacts_kernel<<<blocks, 256, 0, s>>>(m->dev.d, *c, body_q, m->match_new_keys, m->match_prev_keys, m->match_prev_pos,
                                                         m->match_prev_normal, m->match_prev_count, m->match_prev_claim, reset_world_mask,
                                                         pos_threshold * pos_threshold, normal_dot_threshold, match_index);
            match_resolve_kernel<<<blocks, 256, 0, s>>>(*c, m->match_new_keys, m->match_prev_claim, match_index);
            match_save_kernel<<<blocks, 256, 0, s>>>(m->dev.d, *c, body_q, m->match_new_keys, m->match_prev_keys, m->match_prev_pos,
                                                     m->match_prev_normal, m->match_prev_claim, m->match_prev_count);
            count_launch(4);
            if (cudaGetLastError() != cudaSuccess) {
                set_error("nb2_contacts_match: kernel launch failed");
                st = NB2_ERR_CUDA;
            }
        }
    } while (0);
    if (prev >= 0 && prev != m->device) cudaSetDevice(prev);
    return st;
}
