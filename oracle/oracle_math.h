// oracle_math.h - TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Scalar fp32 restatement of the NVIDIA Warp built-ins that the reference kernels call
// (wp.quat_rotate, wp.transform_point, wp.normalize, ...).  The reference's arithmetic for this
// path lives in the un-vendored third-party package `warp-lang` (pinned 1.17.0.dev20260807 at
// /root/reference/uv.lock:7379-7391, floor >=1.16.0 at pyproject.toml:31); it is absent from
// /root/reference, so the formulas below restate Warp's published native headers
// (warp/native/{vec,quat,mat,spatial}.h) and are anchored on the reference's call sites.
// PARITY UNPINNED at the ulp level: operation order inside these built-ins could not be checked
// against the pinned wheel in this container (no network, Warp not installed).
//
// Compile with -ffp-contract=off so every operation rounds exactly as written.
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

struct vec3 {
    float x, y, z;
    vec3() : x(0.f), y(0.f), z(0.f) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, vec3 b) { a = a - b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
    return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length_sq(vec3 a) { return dot(a, a); }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
// wp.normalize(vec): a / |a| when |a| > 0, zero vector otherwise (warp/native/vec.h)
inline vec3 normalize(vec3 a) {
    float l = length(a);
    if (l > 0.0f) return vec3(a.x / l, a.y / l, a.z / l);
    return vec3();
}
inline vec3 cw_mul(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
// wp.min / wp.max on floats: (a < b) ? a : b and (a > b) ? a : b (warp/native/builtin.h)
inline float minf(float a, float b) { return a < b ? a : b; }
inline float maxf(float a, float b) { return a > b ? a : b; }
inline vec3 vmin(vec3 a, vec3 b) { return vec3(minf(a.x, b.x), minf(a.y, b.y), minf(a.z, b.z)); }
inline vec3 vmax(vec3 a, vec3 b) { return vec3(maxf(a.x, b.x), maxf(a.y, b.y), maxf(a.z, b.z)); }
inline vec3 vabs(vec3 a) { return vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }

// Correctly-rounded fp32 inverse trig / trig (double evaluation rounded to float): what glibc's asinf/acosf
// deliver in practice, written explicitly so the strict CUDA build can reproduce it bit for bit.
inline float asin_w(float x) { return (float)std::asin((double)x); }
inline float acos_w(float x) { return (float)std::acos((double)x); }
inline float sin_w(float x) { return (float)std::sin((double)x); }
inline float atan2_w(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float cos_w(float x) { return (float)std::cos((double)x); }
inline float clampf(float x, float a, float b) { return minf(maxf(a, x), b); }
inline float nonzero(float x) { return x != 0.0f ? 1.0f : 0.0f; }
inline float signf(float x) { return x < 0.0f ? -1.0f : 1.0f; }

struct quat {
    float x, y, z, w;
    quat() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    quat(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    quat(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float operator[](int i) const { return (&x)[i]; }
};
inline quat quat_identity() { return quat(0.f, 0.f, 0.f, 1.f); }
inline quat operator+(quat a, quat b) { return quat(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline quat operator*(quat a, float s) { return quat(a.x * s, a.y * s, a.z * s, a.w * s); }
inline quat operator*(float s, quat a) { return quat(a.x * s, a.y * s, a.z * s, a.w * s); }
// quaternion product (warp/native/quat.h mul)
inline quat operator*(quat a, quat b) {
    return quat(a.w * b.x + b.w * a.x + a.y * b.z - b.y * a.z,
                a.w * b.y + b.w * a.y + a.z * b.x - b.z * a.x,
                a.w * b.z + b.w * a.z + a.x * b.y - b.x * a.y,
                a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
inline float dot(quat a, quat b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(quat q) { return std::sqrt(dot(q, q)); }
inline quat normalize(quat q) {
    float l = length(q);
    if (l > 0.0f) {
        float inv = 1.0f / l;
        return quat(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
    }
    return quat(0.f, 0.f, 0.f, 1.f);
}
inline quat quat_inverse(quat q) { return quat(-q.x, -q.y, -q.z, q.w); }
// wp.quat_rotate: v(2w^2-1) + 2(q.v)q + 2w(q x v)
inline vec3 quat_rotate(quat q, vec3 v) {
    float c = 2.0f * q.w * q.w - 1.0f;
    float d = 2.0f * (q.x * v.x + q.y * v.y + q.z * v.z);
    return vec3(v.x * c + q.x * d + (q.y * v.z - q.z * v.y) * q.w * 2.0f,
                v.y * c + q.y * d + (q.z * v.x - q.x * v.z) * q.w * 2.0f,
                v.z * c + q.z * d + (q.x * v.y - q.y * v.x) * q.w * 2.0f);
}
inline vec3 quat_rotate_inv(quat q, vec3 v) {
    float c = 2.0f * q.w * q.w - 1.0f;
    float d = 2.0f * (q.x * v.x + q.y * v.y + q.z * v.z);
    return vec3(v.x * c + q.x * d - (q.y * v.z - q.z * v.y) * q.w * 2.0f,
                v.y * c + q.y * d - (q.z * v.x - q.x * v.z) * q.w * 2.0f,
                v.z * c + q.z * d - (q.x * v.y - q.y * v.x) * q.w * 2.0f);
}
inline quat quat_from_axis_angle(vec3 axis, float angle) {
    float half = angle * 0.5f;
    float w = cos_w(half);
    float s = sin_w(half);
    vec3 v = axis * s;
    return quat(v.x, v.y, v.z, w);
}

struct mat33 {
    float m[3][3];
    mat33() { for (auto& r : m) for (float& v : r) v = 0.f; }
    mat33(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
        m[0][0] = a; m[0][1] = b; m[0][2] = c; m[1][0] = d; m[1][1] = e; m[1][2] = f; m[2][0] = g; m[2][1] = h; m[2][2] = i;
    }
    static mat33 load(const float* p) { return mat33(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]); }
    vec3 col(int j) const { return vec3(m[0][j], m[1][j], m[2][j]); }
};
// mat * vec (warp/native/mat.h mul: sum of columns scaled by components, left to right)
inline vec3 operator*(const mat33& a, vec3 b) {
    vec3 r = a.col(0) * b.x;
    r += a.col(1) * b.y;
    r += a.col(2) * b.z;
    return r;
}
inline mat33 operator*(const mat33& a, const mat33& b) {
    mat33 t;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) t.m[i][j] += a.m[i][k] * b.m[k][j];
    return t;
}
inline mat33 operator*(float s, const mat33& a) {
    mat33 t;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t.m[i][j] = a.m[i][j] * s;
    return t;
}
inline mat33 transpose(const mat33& a) {
    mat33 t;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t.m[i][j] = a.m[j][i];
    return t;
}
// wp.quat_to_matrix: columns are the rotated basis vectors
inline mat33 quat_to_matrix(quat q) {
    vec3 c1 = quat_rotate(q, vec3(1.f, 0.f, 0.f));
    vec3 c2 = quat_rotate(q, vec3(0.f, 1.f, 0.f));
    vec3 c3 = quat_rotate(q, vec3(0.f, 0.f, 1.f));
    return mat33(c1.x, c2.x, c3.x, c1.y, c2.y, c3.y, c1.z, c2.z, c3.z);
}

// wp.matrix_from_cols / wp.quat_from_matrix (warp/native/mat.h, quat.h): trace branch, else the largest diagonal element;
// the result is normalized.  Used by the two-angular-axis D6 joint (sim/articulation.py:37-58).
inline mat33 matrix_from_cols(vec3 c0, vec3 c1, vec3 c2) { return mat33(c0.x, c1.x, c2.x, c0.y, c1.y, c2.y, c0.z, c1.z, c2.z); }
inline quat quat_from_matrix(const mat33& a) {
    const float tr = a.m[0][0] + a.m[1][1] + a.m[2][2];
    float x, y, z, w, h;
    if (tr >= 0.0f) {
        h = std::sqrt(tr + 1.0f);
        w = 0.5f * h;
        h = 0.5f / h;
        x = (a.m[2][1] - a.m[1][2]) * h;
        y = (a.m[0][2] - a.m[2][0]) * h;
        z = (a.m[1][0] - a.m[0][1]) * h;
    } else {
        int max_diag = 0;
        if (a.m[1][1] > a.m[0][0]) max_diag = 1;
        if (a.m[2][2] > a.m[max_diag][max_diag]) max_diag = 2;
        if (max_diag == 0) {
            h = std::sqrt((a.m[0][0] - (a.m[1][1] + a.m[2][2])) + 1.0f);
            x = 0.5f * h;
            h = 0.5f / h;
            y = (a.m[0][1] + a.m[1][0]) * h;
            z = (a.m[2][0] + a.m[0][2]) * h;
            w = (a.m[2][1] - a.m[1][2]) * h;
        } else if (max_diag == 1) {
            h = std::sqrt((a.m[1][1] - (a.m[2][2] + a.m[0][0])) + 1.0f);
            y = 0.5f * h;
            h = 0.5f / h;
            z = (a.m[1][2] + a.m[2][1]) * h;
            x = (a.m[0][1] + a.m[1][0]) * h;
            w = (a.m[0][2] - a.m[2][0]) * h;
        } else {
            h = std::sqrt((a.m[2][2] - (a.m[0][0] + a.m[1][1])) + 1.0f);
            z = 0.5f * h;
            h = 0.5f / h;
            x = (a.m[2][0] + a.m[0][2]) * h;
            y = (a.m[1][2] + a.m[2][1]) * h;
            w = (a.m[1][0] - a.m[0][1]) * h;
        }
    }
    return normalize(quat(x, y, z, w));
}

// newton.math.quat_decompose (reference math/spatial.py:150-175): wp.quat_to_euler(q, 2, 1, 0) with every angle wrapped to
// [-pi, pi).  wp.quat_to_euler is a Warp built-in (not vendored); restated as the direct quaternion -> Euler conversion of
// Bernardes & Viollet (2022) for the Tait-Bryan sequence (i, j, k) = (Z, Y, X).  What the reference relies on - and what pins this
// restatement - is that for q = qx(a) * qy(b) * qz(c) the result is (a, b, c) (invert_3d_rotational_dofs inverts
// compute_3d_rotational_dofs, sim/articulation.py:150-236; test_kinematics.py:1057-1104 to 1e-6).  Bit-level order unpinned.
inline float wrap_angle_pm_pi(float theta) {  // math/spatial.py:133-147; wp.mod = fmod
    const float pi = 3.14159265358979323846f, two_pi = 2.0f * pi;
    float wrapped = std::fmod(theta + pi, two_pi);
    if (wrapped < 0.0f) wrapped += two_pi;
    return wrapped - pi;
}
inline vec3 quat_decompose(quat q) {
    // (i, j, k) = (3, 2, 1) in the paper's 1-based numbering, not proper: sign = (i - j)(j - k)(k - i) / 2 = -1; t = (w, x, y, z)
    const float a = q.w - q.y, b = q.z - q.x, c = q.y + q.w, d = -q.x - q.z;
    const float n_ab = a * a + b * b;
    float theta2 = acos_w(2.0f * n_ab / (n_ab + c * c + d * d) - 1.0f);
    const float theta_plus = atan2_w(b, a), theta_minus = atan2_w(d, c);
    const float theta1 = theta_plus - theta_minus;
    float theta3 = theta_plus + theta_minus;
    theta3 = -theta3;
    theta2 -= 1.57079632679489661923f;
    // theta1 turns about Z, theta2 about Y, theta3 about X: returned in (x, y, z) order
    return vec3(wrap_angle_pm_pi(theta3), wrap_angle_pm_pi(theta2), wrap_angle_pm_pi(theta1));
}

struct transform {
    vec3 p;
    quat q;
    transform() : p(), q(0.f, 0.f, 0.f, 1.f) {}
    transform(vec3 p_, quat q_) : p(p_), q(q_) {}
    static transform load(const float* f) { return transform(vec3(f[0], f[1], f[2]), quat(f[3], f[4], f[5], f[6])); }
    void store(float* f) const { f[0] = p.x; f[1] = p.y; f[2] = p.z; f[3] = q.x; f[4] = q.y; f[5] = q.z; f[6] = q.w; }
};
inline transform transform_identity() { return transform(); }
inline vec3 transform_point(const transform& t, vec3 x) { return t.p + quat_rotate(t.q, x); }
inline vec3 transform_vector(const transform& t, vec3 x) { return quat_rotate(t.q, x); }
inline transform operator*(const transform& a, const transform& b) {
    return transform(quat_rotate(a.q, b.p) + a.p, a.q * b.q);
}
inline transform transform_inverse(const transform& t) {
    quat qi = quat_inverse(t.q);
    return transform(-quat_rotate(qi, t.p), qi);
}

// Newton spatial vectors are (linear, angular): top = first three floats.
struct spatial {
    vec3 top, bot;
    spatial() {}
    spatial(vec3 a, vec3 b) : top(a), bot(b) {}
    static spatial load(const float* f) { return spatial(vec3(f[0], f[1], f[2]), vec3(f[3], f[4], f[5])); }
    void store(float* f) const { f[0] = top.x; f[1] = top.y; f[2] = top.z; f[3] = bot.x; f[4] = bot.y; f[5] = bot.z; }
};
inline spatial operator+(spatial a, spatial b) { return spatial(a.top + b.top, a.bot + b.bot); }
inline spatial operator-(spatial a, spatial b) { return spatial(a.top - b.top, a.bot - b.bot); }
inline spatial operator*(spatial a, float s) { return spatial(a.top * s, a.bot * s); }
inline void atomic_add(float* arr, int idx, spatial v) {
    float* f = arr + 6 * idx;
    f[0] += v.top.x; f[1] += v.top.y; f[2] += v.top.z; f[3] += v.bot.x; f[4] += v.bot.y; f[5] += v.bot.z;
}
inline void atomic_sub(float* arr, int idx, spatial v) {
    float* f = arr + 6 * idx;
    f[0] -= v.top.x; f[1] -= v.top.y; f[2] -= v.top.z; f[3] -= v.bot.x; f[4] -= v.bot.y; f[5] -= v.bot.z;
}
// newton.math.velocity_at_point (reference math/spatial.py:54-79): v + w x r
inline vec3 velocity_at_point(spatial qd, vec3 r) { return cross(qd.bot, r) + qd.top; }

inline vec3 load3(const float* f) { return vec3(f[0], f[1], f[2]); }
inline void store3(float* f, vec3 v) { f[0] = v.x; f[1] = v.y; f[2] = v.z; }

}  // namespace orc
