// oracle_xpbd.h - TEST INFRASTRUCTURE ONLY.
// CPU restatement of the reference SolverXPBD rigid-body path, one C++ function per Warp kernel,
// executed serially in thread-id order (Warp's CPU backend runs `for tid in range(dim)`), which
// fixes the order of the float atomics exactly as on the reference's CPU device.
#pragma once
#include "../include/newton_b200.h"
#include "oracle_math.h"

namespace orc {

enum { JT_PRISMATIC = 0, JT_REVOLUTE = 1, JT_BALL = 2, JT_FIXED = 3, JT_FREE = 4, JT_DISTANCE = 5, JT_D6 = 6, JT_ROD = 7 };
enum { BODY_KINEMATIC = 2 };

// reference solvers/solver.py:64-107
inline void integrate_rigid_body(const transform& q, const spatial& qd, const spatial& f, vec3 com, const mat33& inertia,
                                 float inv_mass, const mat33& inv_inertia, vec3 gravity, float angular_damping, float dt,
                                 transform& q_new, spatial& qd_new) {
    vec3 x0 = q.p;
    quat r0 = q.q;
    vec3 w0 = qd.bot;
    vec3 v0 = qd.top;
    vec3 t0 = f.bot;
    vec3 f0 = f.top;
    vec3 x_com = x0 + quat_rotate(r0, com);
    vec3 v1 = v0 + (f0 * inv_mass + gravity * nonzero(inv_mass)) * dt;
    vec3 x1 = x_com + v1 * dt;
    vec3 wb = quat_rotate_inv(r0, w0);
    vec3 tb = quat_rotate_inv(r0, t0) - cross(wb, inertia * wb);
    vec3 w1 = quat_rotate(r0, wb + inv_inertia * tb * dt);
    quat r1 = normalize(r0 + quat(w1, 0.0f) * r0 * 0.5f * dt);
    w1 *= 1.0f - angular_damping * dt;
    q_new = transform(x1 - quat_rotate(r1, com), r1);
    qd_new = spatial(v1, w1);
}

// reference solvers/solver.py:112-170
inline void integrate_bodies(const nb2_model_desc& m, const float* body_q, const float* body_qd, const float* body_f,
                             float angular_damping, float dt, float* body_q_new, float* body_qd_new) {
    for (int tid = 0; tid < m.body_count; ++tid) {
        if ((m.body_flags[tid] & BODY_KINEMATIC) != 0) {
            for (int k = 0; k < 7; ++k) body_q_new[7 * tid + k] = body_q[7 * tid + k];
            for (int k = 0; k < 6; ++k) body_qd_new[6 * tid + k] = body_qd[6 * tid + k];
            continue;
        }
        transform q = transform::load(body_q + 7 * tid);
        spatial qd = spatial::load(body_qd + 6 * tid);
        spatial f = spatial::load(body_f + 6 * tid);
        float inv_mass = m.body_inv_mass[tid];
        mat33 inertia = mat33::load(m.body_inertia + 9 * tid);
        mat33 inv_inertia = mat33::load(m.body_inv_inertia + 9 * tid);
        vec3 com = load3(m.body_com + 3 * tid);
        int world_idx = m.body_world[tid];
        if (world_idx < 0) world_idx += m.gravity_count;  // negative index wraps (reference model.py:1300-1307)
        vec3 g = load3(m.gravity + 3 * world_idx);
        transform qn;
        spatial qdn;
        integrate_rigid_body(q, qd, f, com, inertia, inv_mass, inv_inertia, g, angular_damping, dt, qn, qdn);
        qn.store(body_q_new + 7 * tid);
        qdn.store(body_qd_new + 6 * tid);
    }
}

// reference solvers/xpbd/kernels.py:945-1075
inline void apply_joint_forces(const nb2_model_desc& m, const float* body_q, const float* joint_f, float dt, float* body_f,
                               float* joint_impulse = nullptr) {
    for (int tid = 0; tid < m.joint_count; ++tid) {
        int type = m.joint_type[tid];
        if (!m.joint_enabled[tid]) continue;
        if (type == JT_FIXED || type == JT_ROD) continue;
        int id_c = m.joint_child[tid];
        int id_p = m.joint_parent[tid];
        transform X_pj = transform::load(m.joint_X_p + 7 * tid);
        transform X_cj = transform::load(m.joint_X_c + 7 * tid);
        transform X_wp = X_pj;
        transform pose_p = X_pj;
        vec3 com_p(0.f);
        if (id_p >= 0) {
            pose_p = transform::load(body_q + 7 * id_p);
            X_wp = pose_p * X_wp;
            com_p = load3(m.body_com + 3 * id_p);
        }
        vec3 r_p = X_wp.p - transform_point(pose_p, com_p);
        transform pose_c = transform::load(body_q + 7 * id_c);
        transform X_wc = pose_c * X_cj;
        vec3 com_c = load3(m.body_com + 3 * id_c);
        vec3 r_c = X_wc.p - transform_point(pose_c, com_c);
        int qd_start = m.joint_qd_start[tid];
        int lin_axis_count = m.joint_dof_dim[2 * tid + 0];
        int ang_axis_count = m.joint_dof_dim[2 * tid + 1];
        vec3 t_total, f_total;
        if (type == JT_FREE || type == JT_DISTANCE) {
            f_total = vec3(joint_f[qd_start + 0], joint_f[qd_start + 1], joint_f[qd_start + 2]);
            t_total = vec3(joint_f[qd_start + 3], joint_f[qd_start + 4], joint_f[qd_start + 5]);
            atomic_add(body_f, id_c, spatial(f_total, t_total));
            if (id_p >= 0) atomic_sub(body_f, id_p, spatial(f_total, t_total));
            if (joint_impulse) atomic_add(joint_impulse, tid, spatial(f_total, t_total) * dt);  // kernels.py:1018-1019
            continue;
        } else if (type == JT_BALL) {
            t_total = vec3(joint_f[qd_start + 0], joint_f[qd_start + 1], joint_f[qd_start + 2]);
        } else if (type == JT_REVOLUTE || type == JT_PRISMATIC || type == JT_D6) {
            for (int k = 0; k < 3; ++k)
                if (lin_axis_count > k) {
                    vec3 axis = load3(m.joint_axis + 3 * (qd_start + k));
                    float f = joint_f[qd_start + k];
                    vec3 a_p = transform_vector(X_wp, axis);
                    f_total += f * a_p;
                }
            for (int k = 0; k < 3; ++k)
                if (ang_axis_count > k) {
                    vec3 axis = load3(m.joint_axis + 3 * (qd_start + lin_axis_count + k));
                    float f = joint_f[qd_start + lin_axis_count + k];
                    vec3 a_p = transform_vector(X_wp, axis);
                    t_total += f * a_p;
                }
        }
        spatial child_wrench(f_total, t_total + cross(r_c, f_total));
        if (id_p >= 0) atomic_sub(body_f, id_p, spatial(f_total, t_total + cross(r_p, f_total)));
        atomic_add(body_f, id_c, child_wrench);
        if (joint_impulse) atomic_add(joint_impulse, tid, child_wrench * dt);  // kernels.py:1074-1075
    }
}

// reference solvers/xpbd/kernels.py:2047-2081
inline float compute_contact_constraint_delta(float err, const transform& tf_a, const transform& tf_b, float m_inv_a,
                                              float m_inv_b, const mat33& I_inv_a, const mat33& I_inv_b, vec3 linear_a,
                                              vec3 linear_b, vec3 angular_a, vec3 angular_b, float relaxation, float dt) {
    float denom = 0.0f;
    denom += length_sq(linear_a) * m_inv_a;
    denom += length_sq(linear_b) * m_inv_b;
    vec3 rot_angular_a = quat_rotate_inv(tf_a.q, angular_a);
    vec3 rot_angular_b = quat_rotate_inv(tf_b.q, angular_b);
    denom += dot(rot_angular_a, I_inv_a * rot_angular_a);
    denom += dot(rot_angular_b, I_inv_b * rot_angular_b);
    float delta_lambda = -err;
    if (denom > 0.0f) delta_lambda /= dt * denom;
    return delta_lambda * relaxation;
}

// reference solvers/xpbd/kernels.py:2084-2124
inline float compute_positional_correction(float err, float derr, const transform& tf_a, const transform& tf_b, float m_inv_a,
                                           float m_inv_b, const mat33& I_inv_a, const mat33& I_inv_b, vec3 linear_a,
                                           vec3 linear_b, vec3 angular_a, vec3 angular_b, float lambda_in, float compliance,
                                           float damping, float dt) {
    float denom = 0.0f;
    denom += length_sq(linear_a) * m_inv_a;
    denom += length_sq(linear_b) * m_inv_b;
    vec3 rot_angular_a = quat_rotate_inv(tf_a.q, angular_a);
    vec3 rot_angular_b = quat_rotate_inv(tf_b.q, angular_b);
    denom += dot(rot_angular_a, I_inv_a * rot_angular_a);
    denom += dot(rot_angular_b, I_inv_b * rot_angular_b);
    float alpha = compliance;
    float gamma = compliance * damping;
    float delta_lambda = -(err + alpha * lambda_in + gamma * derr);
    if (denom + alpha > 0.0f) delta_lambda /= (dt + gamma) * denom + alpha / dt;
    return delta_lambda;
}

// reference solvers/xpbd/kernels.py:2127-2161
inline float compute_angular_correction(float err, float derr, const transform& tf_a, const transform& tf_b,
                                        const mat33& I_inv_a, const mat33& I_inv_b, vec3 angular_a, vec3 angular_b,
                                        float lambda_in, float compliance, float damping, float dt) {
    float denom = 0.0f;
    vec3 rot_angular_a = quat_rotate_inv(tf_a.q, angular_a);
    vec3 rot_angular_b = quat_rotate_inv(tf_b.q, angular_b);
    denom += dot(rot_angular_a, I_inv_a * rot_angular_a);
    denom += dot(rot_angular_b, I_inv_b * rot_angular_b);
    float alpha = compliance;
    float gamma = compliance * damping;
    float delta_lambda = -(err + alpha * lambda_in + gamma * derr);
    if (denom + alpha > 0.0f) delta_lambda /= (dt + gamma) * denom + alpha / dt;
    return delta_lambda;
}

// reference solvers/xpbd/kernels.py:2164-2399 (+ sim/contacts.py:70-115)
inline void solve_body_contact_positions(const nb2_model_desc& m, const float* body_q, const float* body_qd,
                                         const float* body_m_inv, const float* body_I_inv, const nb2_contacts_view& c,
                                         float relaxation, float dt, float* deltas, float* contact_inv_weight,
                                         float* contact_impulse = nullptr) {
    int count = c.rigid_contact_count[0];
    for (int tid = 0; tid < c.rigid_contact_max; ++tid) {
        if (tid >= count) break;
        int shape_a = c.shape0[tid];
        int shape_b = c.shape1[tid];
        if (shape_a == shape_b) continue;
        int body_a = -1, body_b = -1;
        if (shape_a >= 0) body_a = m.shape_body[shape_a];
        if (shape_b >= 0) body_b = m.shape_body[shape_b];
        if (body_a == body_b) continue;
        transform X_wb_a, X_wb_b;
        if (body_a >= 0) X_wb_a = transform::load(body_q + 7 * body_a);
        if (body_b >= 0) X_wb_b = transform::load(body_q + 7 * body_b);
        vec3 p0 = load3(c.point0 + 3 * tid), p1 = load3(c.point1 + 3 * tid);
        vec3 bx_a = transform_point(X_wb_a, p0);
        vec3 bx_b = transform_point(X_wb_b, p1);
        vec3 n = load3(c.normal + 3 * tid);
        float d = dot(n, bx_b - bx_a) - (c.margin0[tid] + c.margin1[tid]);
        if (d >= 0.0f) continue;
        float m_inv_a = 0.f, m_inv_b = 0.f;
        mat33 I_inv_a, I_inv_b;
        vec3 com_a(0.f), com_b(0.f), omega_a(0.f), omega_b(0.f);
        vec3 offset_a = load3(c.offset0 + 3 * tid), offset_b = load3(c.offset1 + 3 * tid);
        if (body_a >= 0) {
            com_a = load3(m.body_com + 3 * body_a);
            m_inv_a = body_m_inv[body_a];
            I_inv_a = mat33::load(body_I_inv + 9 * body_a);
            omega_a = load3(body_qd + 6 * body_a + 3);
        }
        if (body_b >= 0) {
            com_b = load3(m.body_com + 3 * body_b);
            m_inv_b = body_m_inv[body_b];
            I_inv_b = mat33::load(body_I_inv + 9 * body_b);
            omega_b = load3(body_qd + 6 * body_b + 3);
        }
        int mat_nonzero = 0;
        float mu = 0.f, mu_torsional = 0.f, mu_rolling = 0.f;
        if (shape_a >= 0) {
            mat_nonzero += 1;
            mu += m.shape_material_mu[shape_a];
            mu_torsional += m.shape_material_mu_torsional[shape_a];
            mu_rolling += m.shape_material_mu_rolling[shape_a];
        }
        if (shape_b >= 0) {
            mat_nonzero += 1;
            mu += m.shape_material_mu[shape_b];
            mu_torsional += m.shape_material_mu_torsional[shape_b];
            mu_rolling += m.shape_material_mu_rolling[shape_b];
        }
        if (mat_nonzero > 0) {
            mu /= float(mat_nonzero);
            mu_torsional /= float(mat_nonzero);
            mu_rolling /= float(mat_nonzero);
        }
        vec3 r_a = bx_a - transform_point(X_wb_a, com_a);
        vec3 r_b = bx_b - transform_point(X_wb_b, com_b);
        vec3 angular_a = -cross(r_a, n);
        vec3 angular_b = cross(r_b, n);
        if (contact_inv_weight) {
            if (body_a >= 0) contact_inv_weight[body_a] += 1.0f;
            if (body_b >= 0) contact_inv_weight[body_b] += 1.0f;
        }
        float lambda_n = compute_contact_constraint_delta(d, X_wb_a, X_wb_b, m_inv_a, m_inv_b, I_inv_a, I_inv_b, -n, n,
                                                          angular_a, angular_b, relaxation, dt);
        vec3 lin_delta_a = -n * lambda_n;
        vec3 lin_delta_b = n * lambda_n;
        vec3 ang_delta_a = angular_a * lambda_n;
        vec3 ang_delta_b = angular_b * lambda_n;
        if (mu > 0.0f) {
            bx_a = transform_point(X_wb_a, p0 + offset_a);
            bx_b = transform_point(X_wb_b, p1 + offset_b);
            vec3 delta = bx_b - bx_a;
            vec3 friction_delta = delta - dot(n, delta) * n;
            r_a = bx_a - transform_point(X_wb_a, com_a);
            r_b = bx_b - transform_point(X_wb_b, com_b);
            vec3 rel_v_kin_t(0.f);
            if (body_a >= 0 && (m.body_flags[body_a] & BODY_KINEMATIC) != 0) {
                vec3 v_a = velocity_at_point(spatial::load(body_qd + 6 * body_a), r_a);
                rel_v_kin_t = rel_v_kin_t - (v_a - dot(n, v_a) * n);
            }
            if (body_b >= 0 && (m.body_flags[body_b] & BODY_KINEMATIC) != 0) {
                vec3 v_b = velocity_at_point(spatial::load(body_qd + 6 * body_b), r_b);
                rel_v_kin_t = rel_v_kin_t + (v_b - dot(n, v_b) * n);
            }
            friction_delta += rel_v_kin_t * dt;
            vec3 perp = normalize(friction_delta);
            angular_a = -cross(r_a, perp);
            angular_b = cross(r_b, perp);
            float err = length(friction_delta);
            if (err > 0.0f) {
                float lambda_fr = compute_contact_constraint_delta(err, X_wb_a, X_wb_b, m_inv_a, m_inv_b, I_inv_a, I_inv_b,
                                                                   -perp, perp, angular_a, angular_b, relaxation, dt);
                lambda_fr = maxf(lambda_fr, -lambda_n * mu);
                lin_delta_a -= perp * lambda_fr;
                lin_delta_b += perp * lambda_fr;
                ang_delta_a += angular_a * lambda_fr;
                ang_delta_b += angular_b * lambda_fr;
            }
        }
        vec3 delta_omega = omega_b - omega_a;
        if (mu_torsional > 0.0f) {
            float err = dot(delta_omega, n) * dt;
            if (std::fabs(err) > 0.0f) {
                vec3 lin(0.f);
                float lambda_torsion = compute_contact_constraint_delta(err, X_wb_a, X_wb_b, m_inv_a, m_inv_b, I_inv_a,
                                                                        I_inv_b, lin, lin, -n, n, relaxation, dt);
                lambda_torsion = clampf(lambda_torsion, -lambda_n * mu_torsional, lambda_n * mu_torsional);
                ang_delta_a -= n * lambda_torsion;
                ang_delta_b += n * lambda_torsion;
            }
        }
        if (mu_rolling > 0.0f) {
            delta_omega -= dot(n, delta_omega) * n;
            float err = length(delta_omega) * dt;
            if (err > 0.0f) {
                vec3 lin(0.f);
                vec3 roll_n = normalize(delta_omega);
                float lambda_roll = compute_contact_constraint_delta(err, X_wb_a, X_wb_b, m_inv_a, m_inv_b, I_inv_a, I_inv_b,
                                                                     lin, lin, -roll_n, roll_n, relaxation, dt);
                lambda_roll = maxf(lambda_roll, -lambda_n * mu_rolling);
                ang_delta_a -= roll_n * lambda_roll;
                ang_delta_b += roll_n * lambda_roll;
            }
        }
        if (body_a >= 0) atomic_add(deltas, body_a, spatial(lin_delta_a, ang_delta_a));
        if (body_b >= 0) atomic_add(deltas, body_b, spatial(lin_delta_b, ang_delta_b));
        if (contact_impulse) atomic_add(contact_impulse, tid, spatial(lin_delta_a, ang_delta_a));  // kernels.py:2398-2399
    }
}

// reference solvers/xpbd/kernels.py:2402-2461
inline void accumulate_weighted_contact_impulse(const nb2_model_desc& m, const nb2_contacts_view& c, const float* contact_impulse_iter,
                                                const float* constraint_inv_weight, float* contact_impulse) {
    int count = c.rigid_contact_count[0];
    for (int tid = 0; tid < c.rigid_contact_max; ++tid) {
        if (tid >= count) break;
        spatial impulse = spatial::load(contact_impulse_iter + 6 * tid);
        float weight = 1.0f;
        if (constraint_inv_weight) {
            float n_a = 0.0f, n_b = 0.0f;
            int shape_a = c.shape0[tid];
            if (shape_a >= 0) {
                int body_a = m.shape_body[shape_a];
                if (body_a >= 0) n_a = constraint_inv_weight[body_a];
            }
            int shape_b = c.shape1[tid];
            if (shape_b >= 0) {
                int body_b = m.shape_body[shape_b];
                if (body_b >= 0) n_b = constraint_inv_weight[body_b];
            }
            float n_sum = n_a + n_b;
            if (n_sum > 0.0f) {
                if (n_a == 0.0f) weight = 1.0f / n_b;
                else if (n_b == 0.0f) weight = 1.0f / n_a;
                else weight = 2.0f / n_sum;
            }
        }
        atomic_add(contact_impulse, tid, spatial(impulse.top * weight, impulse.bot * weight));
    }
}

// reference solvers/xpbd/kernels.py:2547-2579
inline void update_body_velocities(const nb2_model_desc& m, const float* poses, const float* poses_prev, float dt, float* qd_out) {
    for (int tid = 0; tid < m.body_count; ++tid) {
        transform pose = transform::load(poses + 7 * tid), pose_prev = transform::load(poses_prev + 7 * tid);
        vec3 com = load3(m.body_com + 3 * tid);
        vec3 x_com = pose.p + quat_rotate(pose.q, com);
        vec3 x_com_prev = pose_prev.p + quat_rotate(pose_prev.q, com);
        vec3 v = (x_com - x_com_prev) / dt;
        quat dq = pose.q * quat_inverse(pose_prev.q);
        vec3 omega = (2.0f / dt) * vec3(dq.x, dq.y, dq.z);
        if (dq.w < 0.0f) omega = -omega;
        spatial(v, omega).store(qd_out + 6 * tid);
    }
}

// reference solvers/xpbd/kernels.py:2582-2728
inline void apply_rigid_restitution(const nb2_model_desc& m, const float* body_q, const float* body_qd, const float* body_q_prev,
                                    const float* body_qd_prev, const float* body_m_inv, const float* body_I_inv,
                                    const nb2_contacts_view& c, float dt, float* deltas) {
    (void)body_q;
    int count = c.rigid_contact_count[0];
    for (int tid = 0; tid < c.rigid_contact_max; ++tid) {
        if (tid >= count) break;
        int shape_a = c.shape0[tid], shape_b = c.shape1[tid];
        if (shape_a == shape_b) continue;
        int body_a = -1, body_b = -1, mat_nonzero = 0;
        float restitution = 0.0f;
        if (shape_a >= 0) {
            mat_nonzero += 1;
            restitution += m.shape_material_restitution[shape_a];
            body_a = m.shape_body[shape_a];
        }
        if (shape_b >= 0) {
            mat_nonzero += 1;
            restitution += m.shape_material_restitution[shape_b];
            body_b = m.shape_body[shape_b];
        }
        if (mat_nonzero > 0) restitution /= float(mat_nonzero);
        if (body_a == body_b) continue;
        float m_inv_a = 0.f, m_inv_b = 0.f, inv_mass = 0.f;
        mat33 I_inv_a, I_inv_b;
        transform X_wb_a_prev, X_wb_b_prev;
        vec3 com_a(0.f), com_b(0.f), v_a(0.f), v_b(0.f), v_a_new(0.f), v_b_new(0.f);
        if (body_a >= 0) {
            X_wb_a_prev = transform::load(body_q_prev + 7 * body_a);
            m_inv_a = body_m_inv[body_a];
            I_inv_a = mat33::load(body_I_inv + 9 * body_a);
            com_a = load3(m.body_com + 3 * body_a);
        }
        if (body_b >= 0) {
            X_wb_b_prev = transform::load(body_q_prev + 7 * body_b);
            m_inv_b = body_m_inv[body_b];
            I_inv_b = mat33::load(body_I_inv + 9 * body_b);
            com_b = load3(m.body_com + 3 * body_b);
        }
        // contact_surface_point (sim/contacts.py:96-115): X_wb * (point + offset)
        vec3 bx_a = transform_point(X_wb_a_prev, load3(c.point0 + 3 * tid) + load3(c.offset0 + 3 * tid));
        vec3 bx_b = transform_point(X_wb_b_prev, load3(c.point1 + 3 * tid) + load3(c.offset1 + 3 * tid));
        vec3 n = load3(c.normal + 3 * tid);
        float d = dot(n, bx_b - bx_a);
        if (d >= 0.0f) continue;
        vec3 r_a = bx_a - transform_point(X_wb_a_prev, com_a);
        vec3 r_b = bx_b - transform_point(X_wb_b_prev, com_b);
        vec3 rxn_a(0.f), rxn_b(0.f);
        auto gravity_of = [&](int body) {
            int w = m.body_world[body];
            if (w < 0) w += m.gravity_count;  // Warp negative indexing: world -1 -> last slot
            return load3(m.gravity + 3 * w);
        };
        if (body_a >= 0) {
            v_a = velocity_at_point(spatial::load(body_qd_prev + 6 * body_a), r_a) + gravity_of(body_a) * dt;
            v_a_new = velocity_at_point(spatial::load(body_qd + 6 * body_a), r_a);
            rxn_a = quat_rotate_inv(X_wb_a_prev.q, cross(r_a, n));
            float inv_mass_a = m_inv_a + dot(rxn_a, I_inv_a * rxn_a);
            inv_mass += inv_mass_a;
        }
        if (body_b >= 0) {
            v_b = velocity_at_point(spatial::load(body_qd_prev + 6 * body_b), r_b) + gravity_of(body_b) * dt;
            v_b_new = velocity_at_point(spatial::load(body_qd + 6 * body_b), r_b);
            rxn_b = quat_rotate_inv(X_wb_b_prev.q, cross(r_b, n));
            float inv_mass_b = m_inv_b + dot(rxn_b, I_inv_b * rxn_b);
            inv_mass += inv_mass_b;
        }
        if (inv_mass == 0.0f) continue;
        float rel_vel_old = dot(n, v_b - v_a);
        float rel_vel_new = dot(n, v_b_new - v_a_new);
        if (rel_vel_old >= 0.0f) continue;
        float dv = (-rel_vel_new - restitution * rel_vel_old) / inv_mass;
        if (body_a >= 0) {
            float dv_a = -dv;
            vec3 dq = quat_rotate(X_wb_a_prev.q, I_inv_a * rxn_a * dv_a);
            atomic_add(deltas, body_a, spatial(n * m_inv_a * dv_a, dq));
        }
        if (body_b >= 0) {
            float dv_b = dv;
            vec3 dq = quat_rotate(X_wb_b_prev.q, I_inv_b * rxn_b * dv_b);
            atomic_add(deltas, body_b, spatial(n * m_inv_b * dv_b, dq));
        }
    }
}

// reference solvers/xpbd/kernels.py:2497-2544 (body_parent_f is zeroed by the caller)
inline void convert_joint_impulse_to_parent_f(const nb2_model_desc& m, const float* joint_impulse, float dt, float* body_parent_f) {
    for (int tid = 0; tid < m.joint_count; ++tid) {
        if (!m.joint_enabled[tid]) continue;
        if (m.joint_type[tid] == JT_FREE) continue;
        int id_c = m.joint_child[tid];
        if (id_c < 0) continue;
        float inv_dt = 1.0f / dt;
        spatial impulse = spatial::load(joint_impulse + 6 * tid);
        atomic_add(body_parent_f, id_c, spatial(impulse.top * inv_dt, impulse.bot * inv_dt));
    }
}

// reference solvers/xpbd/kernels.py:864-933
inline void apply_body_deltas(const nb2_model_desc& m, const float* q_in, const float* qd_in, const float* body_inv_m,
                              const float* body_inv_I, const float* deltas, const float* constraint_inv_weights, float dt,
                              float* q_out, float* qd_out) {
    for (int tid = 0; tid < m.body_count; ++tid) {
        float inv_m = body_inv_m[tid];
        if (inv_m == 0.0f) {
            for (int k = 0; k < 7; ++k) q_out[7 * tid + k] = q_in[7 * tid + k];
            for (int k = 0; k < 6; ++k) qd_out[6 * tid + k] = qd_in[6 * tid + k];
            continue;
        }
        mat33 inv_I = mat33::load(body_inv_I + 9 * tid);
        mat33 body_I = mat33::load(m.body_inertia + 9 * tid);
        transform tf = transform::load(q_in + 7 * tid);
        spatial delta = spatial::load(deltas + 6 * tid);
        vec3 v0 = load3(qd_in + 6 * tid);
        vec3 w0 = load3(qd_in + 6 * tid + 3);
        vec3 p0 = tf.p;
        quat q0 = tf.q;
        float weight = 1.0f;
        if (constraint_inv_weights) {
            float inv_weight = constraint_inv_weights[tid];
            if (inv_weight > 0.0f) weight = 1.0f / inv_weight;
        }
        vec3 dp = delta.top * (inv_m * weight);
        vec3 dq = delta.bot * weight;
        vec3 wb = quat_rotate_inv(q0, w0);
        vec3 dwb = inv_I * quat_rotate_inv(q0, dq);
        vec3 tb = cross(dwb, body_I * (wb + dwb)) + cross(wb, body_I * dwb);
        vec3 dw1 = quat_rotate(q0, dwb - dt * inv_I * tb);
        quat q1 = q0 + 0.5f * quat(dw1 * dt, 0.0f) * q0;
        q1 = normalize(q1);
        vec3 com = load3(m.body_com + 3 * tid);
        vec3 x_com = p0 + quat_rotate(q0, com);
        vec3 p1 = x_com + dp * dt;
        p1 -= quat_rotate(q1, com);
        transform(p1, q1).store(q_out + 7 * tid);
        vec3 v1 = v0 + dp;
        vec3 w1 = w0 + dw1;
        if (length(v1) < 1e-4f) v1 = vec3(0.0f);
        if (length(w1) < 1e-4f) w1 = vec3(0.0f);
        spatial(v1, w1).store(qd_out + 6 * tid);
    }
}

// reference solvers/xpbd/kernels.py:1078-1103
inline void update_joint_axis_limits(vec3 axis, float lower, float upper, vec3& lim_lo, vec3& lim_up) {
    vec3 lo_temp = axis * lower;
    vec3 up_temp = axis * upper;
    vec3 lo = vmin(lo_temp, up_temp);
    vec3 up = vmax(lo_temp, up_temp);
    lim_lo = vmin(lim_lo, lo);
    lim_up = vmax(lim_up, up);
}
inline void update_joint_axis_weighted_target(vec3 axis, float target, float weight, vec3& targets, vec3& weights) {
    vec3 weighted_axis = axis * weight;
    targets += weighted_axis * target;
    weights += vabs(weighted_axis);
}

struct AxisSetup {
    vec3 limits_lower, limits_upper, target_pos, stiffness, target_vel, damping;
};

// The "compute joint target, stiffness, damping" blocks at reference kernels.py:1691-1751 and :1911-1973.
inline AxisSetup gather_axes(const nb2_model_desc& m, const nb2_control_view& ctl, int axis_start, int target_axis_start,
                             int offset, int count) {
    AxisSetup s;
    vec3 pos_ke_t, pos_ke_w, vel_kd_t, vel_kd_w;
    for (int k = 0; k < 3; ++k) {
        if (count > k) {
            int axis_idx = axis_start + offset + k;
            int target_axis_idx = target_axis_start + offset + k;
            vec3 axis = load3(m.joint_axis + 3 * axis_idx);
            float lower = m.joint_limit_lower[axis_idx];
            float upper = m.joint_limit_upper[axis_idx];
            if (k == 0) {
                vec3 lo_temp = axis * lower;
                vec3 up_temp = axis * upper;
                s.limits_lower = vmin(lo_temp, up_temp);
                s.limits_upper = vmax(lo_temp, up_temp);
            } else {
                update_joint_axis_limits(axis, lower, upper, s.limits_lower, s.limits_upper);
            }
            float ke = m.joint_target_ke[axis_idx];
            float kd = m.joint_target_kd[axis_idx];
            float target_pos = ctl.joint_target_q[target_axis_idx];
            float target_vel = ctl.joint_target_qd[axis_idx];
            if (ke > 0.0f) update_joint_axis_weighted_target(axis, target_pos, ke, pos_ke_t, pos_ke_w);
            if (kd > 0.0f) update_joint_axis_weighted_target(axis, target_vel, kd, vel_kd_t, vel_kd_w);
        }
    }
    s.target_pos = pos_ke_t;
    s.stiffness = pos_ke_w;
    s.target_vel = vel_kd_t;
    s.damping = vel_kd_w;
    for (int i = 0; i < 3; ++i)
        if (s.stiffness[i] > 0.0f) s.target_pos[i] /= s.stiffness[i];
    for (int i = 0; i < 3; ++i)
        if (s.damping[i] > 0.0f) s.target_vel[i] /= s.damping[i];
    return s;
}

// reference solvers/xpbd/kernels.py:1513-2044
inline void solve_body_joints(const nb2_model_desc& m, const float* body_q, const float* body_qd, const float* body_inv_m,
                              const float* body_inv_I, const nb2_control_view& ctl, float joint_linear_compliance,
                              float joint_angular_compliance, float angular_relaxation, float linear_relaxation, float dt,
                              float* deltas, float* joint_impulse = nullptr) {
    for (int tid = 0; tid < m.joint_count; ++tid) {
        int type = m.joint_type[tid];
        if (!m.joint_enabled[tid]) continue;
        if (type == JT_FREE) continue;
        int id_c = m.joint_child[tid];
        int id_p = m.joint_parent[tid];
        transform X_pj = transform::load(m.joint_X_p + 7 * tid);
        transform X_cj = transform::load(m.joint_X_c + 7 * tid);
        transform X_wp = X_pj;
        float m_inv_p = 0.0f;
        mat33 I_inv_p;
        transform pose_p = X_pj;
        vec3 com_p(0.f), vel_p(0.f), omega_p(0.f);
        if (id_p >= 0) {
            pose_p = transform::load(body_q + 7 * id_p);
            X_wp = pose_p * X_wp;
            com_p = load3(m.body_com + 3 * id_p);
            m_inv_p = body_inv_m[id_p];
            I_inv_p = mat33::load(body_inv_I + 9 * id_p);
            vel_p = load3(body_qd + 6 * id_p);
            omega_p = load3(body_qd + 6 * id_p + 3);
        }
        transform pose_c = transform::load(body_q + 7 * id_c);
        transform X_wc = pose_c * X_cj;
        vec3 com_c = load3(m.body_com + 3 * id_c);
        float m_inv_c = body_inv_m[id_c];
        mat33 I_inv_c = mat33::load(body_inv_I + 9 * id_c);
        vec3 vel_c = load3(body_qd + 6 * id_c);
        vec3 omega_c = load3(body_qd + 6 * id_c + 3);
        if (m_inv_p == 0.0f && m_inv_c == 0.0f) continue;
        vec3 lin_delta_p(0.f), ang_delta_p(0.f), lin_delta_c(0.f), ang_delta_c(0.f);
        transform rel_pose = transform_inverse(X_wp) * X_wc;
        vec3 rel_p = rel_pose.p;
        vec3 x_p = X_wp.p;
        vec3 x_c = X_wc.p;
        float linear_compliance = joint_linear_compliance;
        float angular_compliance = joint_angular_compliance;
        int axis_start = m.joint_qd_start[tid];
        int target_axis_start = m.joint_target_q_start[tid];
        int lin_axis_count = m.joint_dof_dim[2 * tid + 0];
        int ang_axis_count = m.joint_dof_dim[2 * tid + 1];
        vec3 world_com_p = transform_point(pose_p, com_p);
        vec3 world_com_c = transform_point(pose_c, com_c);

        if (type == JT_DISTANCE) {
            vec3 r_p = x_p - world_com_p;
            vec3 r_c = x_c - world_com_c;
            float lower = m.joint_limit_lower[axis_start];
            float upper = m.joint_limit_upper[axis_start];
            if (lower < 0.0f && upper < 0.0f) continue;
            vec3 anchor_delta = x_c - x_p;
            float d = length(anchor_delta);
            float err = 0.0f;
            if (lower >= 0.0f && d < lower) err = d - lower;
            else if (upper >= 0.0f && d > upper) err = d - upper;
            if (std::fabs(err) > 1e-9f) {
                vec3 linear_c;
                if (d > 1e-9f) {
                    linear_c = anchor_delta / d;
                } else {
                    vec3 com_delta = world_com_c - world_com_p;
                    if (length_sq(com_delta) > 1e-18f) linear_c = normalize(com_delta);
                    else linear_c = transform_vector(X_wp, vec3(1.f, 0.f, 0.f));
                }
                vec3 linear_p = -linear_c;
                vec3 angular_p = -cross(r_p, linear_c);
                vec3 angular_c = cross(r_c, linear_c);
                float derr = dot(linear_p, vel_p) + dot(linear_c, vel_c) + dot(angular_p, omega_p) + dot(angular_c, omega_c);
                float compliance = linear_compliance;
                float ke = m.joint_target_ke[axis_start];
                if (ke > 0.0f) compliance = 1.0f / ke;
                float damping = m.joint_target_kd[axis_start];
                float d_lambda = compute_positional_correction(err, derr, pose_p, pose_c, m_inv_p, m_inv_c, I_inv_p, I_inv_c,
                                                               linear_p, linear_c, angular_p, angular_c, 0.0f, compliance,
                                                               damping, dt);
                lin_delta_p += linear_p * (d_lambda * linear_relaxation);
                ang_delta_p += angular_p * (d_lambda * angular_relaxation);
                lin_delta_c += linear_c * (d_lambda * linear_relaxation);
                ang_delta_c += angular_c * (d_lambda * angular_relaxation);
            }
        } else {
            AxisSetup s = gather_axes(m, ctl, axis_start, target_axis_start, 0, lin_axis_count);
            vec3 projected_rel_p = rel_p;
            for (int dim = 0; dim < 3; ++dim) {
                float lower = s.limits_lower[dim];
                float upper = s.limits_upper[dim];
                if (rel_p[dim] < lower) projected_rel_p[dim] = lower;
                else if (rel_p[dim] > upper) projected_rel_p[dim] = upper;
                else if (s.stiffness[dim] > 0.0f) projected_rel_p[dim] = clampf(s.target_pos[dim], lower, upper);
            }
            mat33 frame_p = quat_to_matrix(X_wp.q);
            vec3 r_p = transform_point(X_wp, projected_rel_p) - world_com_p;
            vec3 r_c = x_c - world_com_c;
            for (int dim = 0; dim < 3; ++dim) {
                float e = rel_p[dim];
                vec3 linear_c(frame_p.m[0][dim], frame_p.m[1][dim], frame_p.m[2][dim]);
                vec3 linear_p = -linear_c;
                vec3 angular_p = -cross(r_p, linear_c);
                vec3 angular_c = cross(r_c, linear_c);
                float derr = dot(linear_p, vel_p) + dot(linear_c, vel_c) + dot(angular_p, omega_p) + dot(angular_c, omega_c);
                float err = 0.0f;
                float compliance = linear_compliance;
                float damping = 0.0f;
                float target_vel = s.target_vel[dim];
                float derr_rel = derr - target_vel;
                float lower = s.limits_lower[dim];
                float upper = s.limits_upper[dim];
                if (e < lower) err = e - lower;
                else if (e > upper) err = e - upper;
                else {
                    float target_pos = clampf(s.target_pos[dim], lower, upper);
                    if (s.stiffness[dim] > 0.0f) {
                        err = e - target_pos;
                        compliance = 1.0f / s.stiffness[dim];
                        damping = s.damping[dim];
                    } else if (s.damping[dim] > 0.0f) {
                        compliance = 1.0f / s.damping[dim];
                        damping = s.damping[dim];
                    }
                }
                if (std::fabs(err) > 1e-9f || std::fabs(derr_rel) > 1e-9f) {
                    float d_lambda = compute_positional_correction(err, derr_rel, pose_p, pose_c, m_inv_p, m_inv_c, I_inv_p,
                                                                   I_inv_c, linear_p, linear_c, angular_p, angular_c, 0.0f,
                                                                   compliance, damping, dt);
                    lin_delta_p += linear_p * (d_lambda * linear_relaxation);
                    ang_delta_p += angular_p * (d_lambda * angular_relaxation);
                    lin_delta_c += linear_c * (d_lambda * linear_relaxation);
                    ang_delta_c += angular_c * (d_lambda * angular_relaxation);
                }
            }
        }

        if (type == JT_FIXED || type == JT_PRISMATIC || type == JT_REVOLUTE || type == JT_D6) {
            quat q_p = X_wp.q;
            quat q_c = X_wc.q;
            if (dot(q_p, q_c) < 0.0f) q_c = q_c * -1.0f;
            quat rel_q = quat_inverse(q_p) * q_c;
            quat qtwist = normalize(quat(rel_q.x, 0.0f, 0.0f, rel_q.w));
            quat qswing = rel_q * quat_inverse(qtwist);
            float s_ = std::sqrt(rel_q.x * rel_q.x + rel_q.w * rel_q.w);
            float invs = 1.0f / s_;
            float invscube = invs * invs * invs;
            float err_0 = 2.0f * asin_w(clampf(qtwist.x, -1.0f, 1.0f));
            float err_1 = qswing.y;
            float err_2 = qswing.z;
            quat grad_0(invs - rel_q.x * rel_q.x * invscube, 0.0f, 0.0f, -(rel_q.w * rel_q.x) * invscube);
            quat grad_1(-rel_q.w * (rel_q.w * rel_q.z + rel_q.x * rel_q.y) * invscube, rel_q.w * invs, -rel_q.x * invs,
                        rel_q.x * (rel_q.w * rel_q.z + rel_q.x * rel_q.y) * invscube);
            quat grad_2(rel_q.w * (rel_q.w * rel_q.y - rel_q.x * rel_q.z) * invscube, rel_q.x * invs, rel_q.w * invs,
                        rel_q.x * (rel_q.z * rel_q.x - rel_q.w * rel_q.y) * invscube);
            grad_0 = grad_0 * (2.0f / std::fabs(qtwist.w));
            float swing_sq = qswing.w * qswing.w;
            const float angularEps = 1.0e-4f;
            if (swing_sq + angularEps < 1.0f) {
                float d = std::sqrt(1.0f - qswing.w * qswing.w);
                float theta = 2.0f * acos_w(clampf(qswing.w, -1.0f, 1.0f));
                float scale = theta / d;
                err_1 *= scale;
                err_2 *= scale;
                grad_1 = grad_1 * scale;
                grad_2 = grad_2 * scale;
            }
            vec3 errs(err_0, err_1, err_2);
            vec3 grad_x(grad_0.x, grad_1.x, grad_2.x);
            vec3 grad_y(grad_0.y, grad_1.y, grad_2.y);
            vec3 grad_z(grad_0.z, grad_1.z, grad_2.z);
            vec3 grad_w(grad_0.w, grad_1.w, grad_2.w);
            AxisSetup s = gather_axes(m, ctl, axis_start, target_axis_start, lin_axis_count, ang_axis_count);
            for (int dim = 0; dim < 3; ++dim) {
                float e = errs[dim];
                quat grad(grad_x[dim], grad_y[dim], grad_z[dim], grad_w[dim]);
                quat quat_c = 0.5f * q_p * grad * quat_inverse(q_c);
                vec3 angular_c(quat_c.x, quat_c.y, quat_c.z);
                vec3 angular_p = -angular_c;
                float derr = dot(angular_p, omega_p) + dot(angular_c, omega_c);
                float err = 0.0f;
                float compliance = angular_compliance;
                float damping = 0.0f;
                float target_vel = s.target_vel[dim];
                float angular_c_len = length(angular_c);
                float derr_rel = derr - target_vel * angular_c_len;
                float lower = s.limits_lower[dim];
                float upper = s.limits_upper[dim];
                if (e < lower) err = e - lower;
                else if (e > upper) err = e - upper;
                else {
                    float target_pos = clampf(s.target_pos[dim], lower, upper);
                    if (s.stiffness[dim] > 0.0f) {
                        err = e - target_pos;
                        compliance = 1.0f / s.stiffness[dim];
                        damping = s.damping[dim];
                    } else if (s.damping[dim] > 0.0f) {
                        damping = s.damping[dim];
                        compliance = 1.0f / s.damping[dim];
                    }
                }
                float d_lambda = compute_angular_correction(err, derr_rel, pose_p, pose_c, I_inv_p, I_inv_c, angular_p,
                                                            angular_c, 0.0f, compliance, damping, dt) *
                                 angular_relaxation;
                ang_delta_p += angular_p * d_lambda;
                ang_delta_c += angular_c * d_lambda;
            }
        }
        if (id_p >= 0) atomic_add(deltas, id_p, spatial(lin_delta_p, ang_delta_p));
        if (id_c >= 0) atomic_add(deltas, id_c, spatial(lin_delta_c, ang_delta_c));
        if (joint_impulse) atomic_add(joint_impulse, tid, spatial(lin_delta_c, ang_delta_c));  // kernels.py:2043-2044
    }
}

}  // namespace orc
