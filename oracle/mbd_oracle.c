/* mbd_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference hot path `reverse_once`
 * (/root/reference/mbd/planners/mbd_planner.py:97-135) below the planner statistics:
 *   - jax.random.normal / uniform / split              (App. C of SURVEY.md; jax/_src/prng.py)
 *   - Car2d.step / get_reward / eval_xref_logpd        (/root/reference/mbd/envs/car2d.py:10-102)
 *   - mbd.utils.rollout_us                             (/root/reference/mbd/utils.py:14-20)
 *   - HumanoidRun.step/_get_reward                     (/root/reference/mbd/envs/humanoidrun.py:34-51)
 *   - HumanoidTrack.step/_get_reward/eval_xref_logpd   (/root/reference/mbd/envs/humanoidtrack.py:63-106)
 *   - HumanoidStandup.step/_get_reward                 (/root/reference/mbd/envs/humanoidstandup.py:40-56)
 *   - brax PipelineEnv.pipeline_step + brax.positional.pipeline.step (call site humanoidrun.py:36)
 *
 * PARITY UNPINNED for the Brax part: Brax is an un-vendored, un-pinned third-party
 * dependency of the reference (absent from /root/reference and from this image, no JAX
 * either), and the reference ships no tests or golden vectors.  The positional step below
 * restates Brax 0.10.x's published algorithm (brax/positional/{pipeline,joints,collisions,
 * integrator}.py, brax/{kinematics,com,actuator,math}.py; XPBD after Mueller et al. 2020)
 * function by function; where Brax's exact expression could not be confirmed the choice is
 * marked [restated].  Association order / FMA placement is OURS and is fixed (see
 * include/mbd_fp32.h) so that this oracle and the CUDA kernel agree bit for bit.
 * car2d and the planner arithmetic are restated from in-repo reference source and ARE pinned
 * up to libm-level differences; the PRNG is pinned by the JAX known-answer vectors in
 * tests/test_prng.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off -mfma -fopenmp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "mbd_fp32.h"
#include "mbd_model.h"

#define ORC_API __attribute__((visibility("default")))

/* ---- [restated] choices behind named compile-time switches ------------------------------------------------------
 * Every place where Brax's exact expression could not be confirmed has a switch; the DEFAULT values are what the CUDA
 * kernels implement (tests compare against the default build only).  scripts/pin_against_brax.py rebuilds this file
 * with other values (make -C oracle variant DEFS="-DORC_...=0" OUT=...) and reports which combination reproduces a real
 * Brax install, substep by substep.  Alternative readings are listed in DESIGN.md section 2. */
#ifndef ORC_JOINT_PASSIVE_IN_ACCEL /* 1: dof stiffness / damping torques enter joints.acceleration_update next to the motor torque;
                                      0: motor torque and constraint_ang_damping only (passive dof terms ignored) */
#define ORC_JOINT_PASSIVE_IN_ACCEL 1
#endif
#ifndef ORC_ANG_DAMP_IN_ACCEL      /* 1: - constraint_ang_damping * jd.ang is part of the joint torque; 0: it is not */
#define ORC_ANG_DAMP_IN_ACCEL 1
#endif
#ifndef ORC_EPS                    /* regulariser added to every XPBD denominator (w1 + w2 + eps) */
#define ORC_EPS 1e-6f
#endif
#ifndef ORC_STATIC_FRICTION_MU     /* static-friction test of resolve_position: 1: |dlambda_t| < mu |dlambda|; 0: |dlambda_t| < |dlambda|
                                      (the form of Brax v1's colliders; identical for the humanoids, whose mu is 1) */
#define ORC_STATIC_FRICTION_MU 1
#endif
#ifndef ORC_SINKING_GATE           /* 1: the normal velocity impulse acts only when the contact was approaching (v_n_old <= 0); 0: always */
#define ORC_SINKING_GATE 1
#endif
#ifndef ORC_CONTACT_MIDPOINT       /* 1: contact position c - n (r + dist/2) (MJX plane_sphere); 0: the sphere's surface point c - n r */
#define ORC_CONTACT_MIDPOINT 1
#endif
#ifndef ORC_TANGENT_EPS_FORM       /* tangent direction of the friction terms: 0: exact normalise with a zero guard (math.normalize);
                                      1: dx / (1e-6 + |dx|), the form Brax v1's colliders use [brax-recalled] */
#define ORC_TANGENT_EPS_FORM 0
#endif
#ifndef ORC_EPS_TANGENT            /* 1: the tangential XPBD denominators carry the same regulariser as the normal ones; 0: they do not */
#define ORC_EPS_TANGENT 1
#endif
#ifndef ORC_EULER_ACOS             /* 0: joint angles from matrix entries with atan2 only; 1: the middle angle as acos(clip(cos)) * sign(sin),
                                      the line-of-nodes form (libm acosf: oracle-only, never bit-compared) */
#define ORC_EULER_ACOS 0
#endif

/* ================================================================================== */
/* brax/math.py                                                                        */
/* ================================================================================== */
typedef struct { float x, y, z; } v3;
typedef struct { float w, x, y, z; } q4;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vscale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V3(-a.x, -a.y, -a.z); }
/* a + b*s */
static inline v3 vfma(v3 b, float s, v3 a) { return V3(fmaf(b.x, s, a.x), fmaf(b.y, s, a.y), fmaf(b.z, s, a.z)); }
static inline float vdot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 vcross(v3 a, v3 b) {
  return V3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
/* math.normalize: x / (norm + 1e-6*(norm==0)) — a zero vector stays zero */
static inline v3 vnormalize(v3 a, float* norm) {
  float n = sqrtf(vdot(a, a));
  float inv = (n == 0.0f) ? 0.0f : 1.0f / n;
  *norm = n;
  return vscale(a, inv);
}
static inline q4 Q4(float w, float x, float y, float z) { q4 r = {w, x, y, z}; return r; }
static inline q4 qconj(q4 q) { return Q4(q.w, -q.x, -q.y, -q.z); }
/* math.quat_mul (Hamilton) */
static inline q4 qmul(q4 u, q4 v) {
  return Q4(fmaf(-u.z, v.z, fmaf(-u.y, v.y, fmaf(-u.x, v.x, u.w * v.w))),
            fmaf(-u.z, v.y, fmaf(u.y, v.z, fmaf(u.x, v.w, u.w * v.x))),
            fmaf(u.z, v.x, fmaf(u.y, v.w, fmaf(-u.x, v.z, u.w * v.y))),
            fmaf(u.z, v.w, fmaf(-u.y, v.x, fmaf(u.x, v.y, u.w * v.z))));
}
/* quat_mul(ang_to_quat(a), q) */
static inline q4 vqmul(v3 a, q4 q) {
  return Q4(fmaf(-a.z, q.z, fmaf(-a.y, q.y, -(a.x * q.x))),
            fmaf(-a.z, q.y, fmaf(a.y, q.z, a.x * q.w)),
            fmaf(a.z, q.x, fmaf(a.y, q.w, -(a.x * q.z))),
            fmaf(a.z, q.w, fmaf(-a.y, q.x, a.x * q.y)));
}
/* math.rotate for unit quaternions: v + 2 s (u x v) + 2 u x (u x v)  [restated: same map as
 * Brax's 2(u.v)u + (s^2-u.u)v + 2s(u x v) when |q| = 1] */
static inline v3 vrotate(v3 v, q4 q) {
  v3 u = V3(q.x, q.y, q.z);
  v3 t = vcross(u, v);
  t = vadd(t, t);
  v3 c = vcross(u, t);
  return V3(fmaf(q.w, t.x, v.x) + c.x, fmaf(q.w, t.y, v.y) + c.y, fmaf(q.w, t.z, v.z) + c.z);
}
static inline v3 vinv_rotate(v3 v, q4 q) { return vrotate(v, qconj(q)); }
static inline q4 qnormalize(q4 q) {
  float n = sqrtf(fmaf(q.z, q.z, fmaf(q.y, q.y, fmaf(q.x, q.x, q.w * q.w))));
  float inv = 1.0f / n;
  return Q4(q.w * inv, q.x * inv, q.y * inv, q.z * inv);
}
static inline q4 qadd(q4 a, q4 b) { return Q4(a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z); }
static inline q4 qscale(q4 a, float s) { return Q4(a.w * s, a.x * s, a.y * s, a.z * s); }
static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ================================================================================== */
/* model blob access                                                                   */
/* ================================================================================== */
typedef struct { v3 p; q4 q; v3 w; v3 v; } Link; /* x_i.pos, x_i.rot, xd_i.ang, xd_i.vel */

typedef struct {
  const float* f;
  const int32_t* i;
  int L, nu, n_frames, reward, ntrack;
  float dt, inv_dt, half_dt, two_inv_dt, vel_damp, ang_damp, scale_pos, scale_ang, collide_scale, elasticity;
  v3 g;
} Model;

static inline float LFf(const Model* m, int field, int l) { return m->f[MBD_HDR_WORDS + field * MBD_MAXL + l]; }
static inline int LFi(const Model* m, int field, int l) { return m->i[MBD_HDR_WORDS + field * MBD_MAXL + l]; }
static inline v3 LF3(const Model* m, int field, int l) { return V3(LFf(m, field, l), LFf(m, field + 1, l), LFf(m, field + 2, l)); }
static inline q4 LF4(const Model* m, int field, int l) {
  return Q4(LFf(m, field, l), LFf(m, field + 1, l), LFf(m, field + 2, l), LFf(m, field + 3, l));
}

static int model_open(Model* m, const uint32_t* blob) {
  if (blob[MBD_H_MAGIC] != MBD_MODEL_MAGIC) return -1;
  m->f = (const float*)blob;
  m->i = (const int32_t*)blob;
  m->L = m->i[MBD_H_NLINK]; m->nu = m->i[MBD_H_NU]; m->n_frames = m->i[MBD_H_NFRAMES];
  m->reward = m->i[MBD_H_REWARD]; m->ntrack = m->i[MBD_H_NTRACK];
  m->dt = m->f[MBD_H_DT]; m->inv_dt = m->f[MBD_H_INV_DT]; m->half_dt = m->f[MBD_H_HALF_DT];
  m->two_inv_dt = m->f[MBD_H_TWO_INV_DT];
  m->vel_damp = m->f[MBD_H_VEL_DAMP]; m->ang_damp = m->f[MBD_H_ANG_DAMP];
  m->scale_pos = m->f[MBD_H_SCALE_POS]; m->scale_ang = m->f[MBD_H_SCALE_ANG];
  m->collide_scale = m->f[MBD_H_COLLIDE_SCALE]; m->elasticity = m->f[MBD_H_ELASTICITY];
  m->g = V3(m->f[MBD_H_GX], m->f[MBD_H_GY], m->f[MBD_H_GZ]);
  return 0;
}

/* ================================================================================== */
/* brax/kinematics.py: world_to_joint (rotational part) + axis_angle_ang               */
/* ================================================================================== */
typedef struct {
  float ang[3];   /* joint angles (psi, theta, parity*phi): intrinsic x-y'-z'' about the joint-frame axes */
  v3 ax[3];       /* instantaneous rotation axes in the PARENT joint frame (a_p) */
  float r10, r20; /* extra entries of R(j.rot), used by the 1-dof axis alignment */
} JointAngles;

/* kinematics.axis_angle_ang [restated]: Brax builds (psi,theta,phi) from a line of nodes
 * cross(axis_3_c, axis_1_p) with signed_angle / arccos; that is the intrinsic Euler
 * decomposition R(j) = Rx(psi) Ry(theta) Rz(phi) in the joint frame, extracted here from the
 * matrix entries with atan2 only (theta = atan2(sin, cos) instead of arccos(cos)*sign(sin):
 * same angle, better conditioned near 0).  A left-handed axis triple carries parity = -1 on
 * the third axis/angle (kinematics.link_to_joint_frame). */
static inline void axis_angle_ang(q4 j, float parity, JointAngles* o) {
  float w = j.w, x = j.x, y = j.y, z = j.z;
  float r00 = 1.0f - 2.0f * fmaf(z, z, y * y);
  float r01 = 2.0f * fmaf(x, y, -(w * z));
  float r02 = 2.0f * fmaf(x, z, w * y);
  float r12 = 2.0f * fmaf(y, z, -(w * x));
  float r22 = 1.0f - 2.0f * fmaf(y, y, x * x);
  o->r10 = 2.0f * fmaf(x, y, w * z);
  o->r20 = 2.0f * fmaf(x, z, -(w * y));
  float psi = mbd_atan2f(-r12, r22);
  float cth = sqrtf(fmaf(r01, r01, r00 * r00));
  float theta = ORC_EULER_ACOS ? copysignf(acosf(clampf(cth, -1.0f, 1.0f)), r02) : mbd_atan2f(r02, cth);
  float phi = mbd_atan2f(-r01, r00);
  float ln;
  v3 lon = vnormalize(V3(0.0f, r22, -r12), &ln);
  o->ang[0] = psi; o->ang[1] = theta; o->ang[2] = parity * phi;
  o->ax[0] = V3(1.0f, 0.0f, 0.0f);
  o->ax[1] = lon;
  o->ax[2] = V3(parity * r02, parity * r12, parity * r22);
}


/* ---- contact.get (MJX plane_sphere) + collisions.resolve_position / resolve_velocity -----------------
 * Written for the ground-plane normal n = +z, the only contact class of the positional envs in
 * scope (every vendored model collides spheres/capsule caps with the z = 0 floor): with n
 * substituted, r x n = (r.y, -r.x, 0), P = dl n = (0,0,dl), r x P = (r.y dl, -r.x dl, 0) and
 * tangential vectors have z = 0.  Same XPBD contact as the general formulation
 * (brax/positional/collisions.py), term for term. */
static inline q4 vqmul_xy(float ax, float ay, q4 q) {
  return Q4(fmaf(-ay, q.y, -(ax * q.x)), fmaf(ay, q.z, ax * q.w), fmaf(ay, q.w, -(ax * q.z)), fmaf(-ay, q.x, ax * q.y));
}
static inline void contact_position_plane(const Model* m, int l, int ci, float im, v3 p, q4 q, v3 p_prev, q4 q_prev, v3* dp, q4* dq,
                                          float* dl_out, v3* cp_out) {
  const int base = MBD_F_CON0 + ci * MBD_CON_STRIDE;
  const float radius = LFf(m, base + 3, l), mu = LFf(m, base + 4, l);
  v3 centre = vadd(p, vrotate(LF3(m, base, l), q));
  float dist = centre.z - radius;                                       /* dist = (c - plane).n - r */
  v3 cp = V3(centre.x, centre.y, centre.z - (radius + (ORC_CONTACT_MIDPOINT ? 0.5f * dist : 0.0f)));    /* pos = c - n (r + dist/2) */
  int coll = dist < 0.0f;
  v3 r = vsub(cp, p);
  float w = im + fmaf(r.x, r.x, r.y * r.y);                             /* 1/m + |r x n|^2 (identity inertia) */
  float dl = coll ? (-dist / (w + ORC_EPS)) : 0.0f;
  dp->z = dp->z + dl * im;
  *dq = qadd(*dq, qscale(vqmul_xy(r.y * dl, -(r.x * dl), q), 0.5f));
  /* static friction: cancel the tangential travel of the contact point since x_i_prev */
  v3 rl = vinv_rotate(r, q);
  v3 pbar = vadd(p_prev, vrotate(rl, q_prev));
  float dx = cp.x - pbar.x, dy = cp.y - pbar.y;
  float ct = sqrtf(fmaf(dy, dy, dx * dx));
  float inv = ORC_TANGENT_EPS_FORM ? 1.0f / (1e-6f + ct) : ((ct == 0.0f) ? 0.0f : 1.0f / ct);
  float ntx = dx * inv, nty = dy * inv;
  float c1 = -(r.z * nty), c2 = r.z * ntx, c3 = fmaf(r.x, nty, -(r.y * ntx));
  float wt = im + fmaf(c3, c3, fmaf(c2, c2, c1 * c1));
  float dlt = -ct / (wt + (ORC_EPS_TANGENT ? ORC_EPS : 0.0f));
  int stat = coll && (fabsf(dlt) < (ORC_STATIC_FRICTION_MU ? mu : 1.0f) * fabsf(dl));
  float mm = stat ? dlt : 0.0f;
  float ptx = ntx * mm, pty = nty * mm;
  dp->x = dp->x + ptx * im;
  dp->y = dp->y + pty * im;
  *dq = qadd(*dq, qscale(vqmul(V3(-(r.z * pty), r.z * ptx, fmaf(r.x, pty, -(r.y * ptx))), q), 0.5f));
  *dl_out = dl;
  *cp_out = cp;
}
static inline void contact_velocity_plane(const Model* m, int l, int ci, float im, v3 p, v3 v, v3 w, v3 v_before, v3 w_before, v3 cp,
                                          float dl, v3* dv, v3* dw) {
  const float mu = LFf(m, MBD_F_CON0 + ci * MBD_CON_STRIDE + 4, l);
  v3 r = vsub(cp, p);
  v3 rel = vadd(v, vcross(w, r));
  float vn = rel.z;
  float vtn = sqrtf(fmaf(rel.y, rel.y, rel.x * rel.x));
  float inv = ORC_TANGENT_EPS_FORM ? 1.0f / (1e-6f + vtn) : ((vtn == 0.0f) ? 0.0f : 1.0f / vtn);
  float tdx = rel.x * inv, tdy = rel.y * inv;
  float fr = mu * fabsf(dl) * m->inv_dt;                                /* dynamic friction bound mu |dlambda| / dt */
  float mag = fr < vtn ? fr : vtn;
  float c1 = -(r.z * tdy), c2 = r.z * tdx, c3 = fmaf(r.x, tdy, -(r.y * tdx));
  float wd = im + fmaf(c3, c3, fmaf(c2, c2, c1 * c1));
  float kd = 1.0f / (wd + (ORC_EPS_TANGENT ? ORC_EPS : 0.0f));
  float pdx = (tdx * -mag) * kd, pdy = (tdy * -mag) * kd;
  v3 rel_old = vadd(v_before, vcross(w_before, r));
  float vn_old = rel_old.z;
  float rest = -m->elasticity * vn_old;                                 /* restitution: min(-e v_n_old, 0) */
  rest = rest < 0.0f ? rest : 0.0f;
  float wn = im + fmaf(r.x, r.x, r.y * r.y);
  float prz = (-vn + rest) * (1.0f / (wn + ORC_EPS));
  v3 P = V3(pdx, pdy, (!ORC_SINKING_GATE || vn_old <= 0.0f) ? prz : 0.0f);                   /* "sinking" gate on the normal impulse */
  if (dl == 0.0f) P = V3(0, 0, 0);
  *dv = vadd(*dv, vscale(P, im));
  *dw = vadd(*dw, vcross(r, P));
}

/* ---- slide (prismatic) dofs -------------------------------------------------------------------------
 * [restated] Brax's positional joints treat a translational dof like XPBD's prismatic joint (Mueller et al.
 * 2020, sec. 3.4.2): the component of the anchor offset along the free axis is removed from the positional error
 * (what exceeds the limits stays in), and the dof's spring / damper / motor act as a force along the axis at the
 * child anchor.  Brax's exact expressions are unpinned like the rest of the physics; the form below is this
 * repo's.  Supported arrangement (blob.pack enforces it): the link's parent is the WORLD and its slide dofs
 * precede its hinges, so the slide axes are the fixed columns of the parent-side joint frame a_p = PQ —
 * planar roots (hopper, walker2d, halfcheetah: slide x, slide z, hinge y) and the cartpole cart.
 * Axis of dof k in the joint frame: e_k (the third column carries the parity, like the third hinge axis). */
static inline v3 slide_axis(int k, float parity, q4 a_p) {
  v3 e = k == 0 ? V3(1.0f, 0.0f, 0.0f) : (k == 1 ? V3(0.0f, 1.0f, 0.0f) : V3(0.0f, 0.0f, parity));
  return vrotate(e, a_p);
}

/* ================================================================================== */
/* brax/positional/pipeline.py: step                                                   */
/* ================================================================================== */
static void positional_step(const Model* m, Link* s, const float* act) {
  const int L = m->L;
  Link prev[MBD_MAXL];
  v3 T[MBD_MAXL];
  v3 Fa[MBD_MAXL]; /* linear acceleration from slide-dof forces (zero for every other link) */
  memcpy(prev, s, sizeof(Link) * L); /* x_i_prev = state.x_i */

  /* ---- actuator.to_tau + joints.acceleration_update -------------------------------- *
   * per jointed link, in the parent joint frame a_p:
   *   torque = sum_k axis_k * (tau_k - stiffness_k*angle_k - damping_k*vel_k) - constraint_ang_damping * jd.ang
   * rotated to world, +T on the child and -T on the parent; xdd.ang = inv_inertia @ T = T
   * (spring_inertia_scale = 1 -> identity inertia), xdd.vel = gravity.                  */
  for (int l = 0; l < L; ++l) {
    T[l] = V3(0, 0, 0);
    Fa[l] = V3(0, 0, 0);
    int ndof = LFi(m, MBD_F_NDOF, l);
    if (ndof <= 0) continue;
    const int smask = LFi(m, MBD_F_SLIDE, l);
    int par = LFi(m, MBD_F_PARENT, l);
    q4 qp = par >= 0 ? s[par].q : Q4(1, 0, 0, 0);
    v3 wp = par >= 0 ? s[par].w : V3(0, 0, 0);
    q4 a_p = qmul(qp, LF4(m, MBD_F_PQ, l));
    q4 a_c = qmul(s[l].q, LF4(m, MBD_F_JQ, l));
    q4 j = qmul(qconj(a_p), a_c);
    v3 jd = vinv_rotate(vsub(s[l].w, wp), a_p);
    JointAngles ja;
    axis_angle_ang(j, LFf(m, MBD_F_PARITY, l), &ja);
    v3 tq = vscale(jd, ORC_ANG_DAMP_IN_ACCEL ? -LFf(m, MBD_F_ANG_DAMP, l) : -0.0f);
    v3 rcw_s = V3(0, 0, 0), d_s = V3(0, 0, 0), va_s = V3(0, 0, 0), Fw = V3(0, 0, 0);
    if (smask) { /* anchor offset and anchor velocity of the child against the (fixed) world anchor RP */
      rcw_s = vrotate(LF3(m, MBD_F_RC, l), s[l].q);
      d_s = vsub(vadd(s[l].p, rcw_s), LF3(m, MBD_F_RP, l));
      va_s = vadd(s[l].v, vcross(s[l].w, rcw_s));
    }
    for (int k = 0; k < ndof; ++k) {
      int base = MBD_F_DOF0 + k * MBD_DOF_STRIDE;
      float tau = 0.0f;
      int a_id = LFi(m, base + MBD_D_ACT, l);
      if (a_id >= 0) /* actuator.to_tau: clip(act, ctrl_range) * gear (motor: gain 1, bias 0) */
        tau = LFf(m, base + MBD_D_GEAR, l) * clampf(act[a_id], LFf(m, base + MBD_D_CLO, l), LFf(m, base + MBD_D_CHI, l));
      if ((smask >> k) & 1) { /* slide dof: force along the axis at the child anchor */
        v3 ak = slide_axis(k, LFf(m, MBD_F_PARITY, l), a_p);
        float x = vdot(d_s, ak), xd = vdot(va_s, ak);
        float f = fmaf(-LFf(m, base + MBD_D_DAMP, l), xd, fmaf(-LFf(m, base + MBD_D_STIFF, l), x, tau));
        Fw = vfma(ak, f, Fw);
        continue;
      }
      float vel = vdot(ja.ax[k], jd);
      float t = ORC_JOINT_PASSIVE_IN_ACCEL ? fmaf(-LFf(m, base + MBD_D_DAMP, l), vel, fmaf(-LFf(m, base + MBD_D_STIFF, l), ja.ang[k], tau)) : tau;
      tq = vfma(ja.ax[k], t, tq);
    }
    T[l] = vrotate(tq, a_p);
    if (smask) { /* F at the anchor: linear acceleration F/m, torque rcw x F about the COM (identity inertia) */
      Fa[l] = vscale(Fw, LFf(m, MBD_F_INV_MASS, l));
      T[l] = vadd(T[l], vcross(rcw_s, Fw));
    }
  }
  /* ---- integrator.integrate_xdd (semi-implicit Euler) ------------------------------- */
  Link before[MBD_MAXL]; /* xd_i right after integration = "xd_i_before" for resolve_velocity */
  for (int l = 0; l < L; ++l) {
    v3 acc = T[l];
    for (int c = 0; c < MBD_MAXCHILD; ++c) {
      int ch = LFi(m, MBD_F_CHILD0 + c, l);
      if (ch >= 0) acc = vsub(acc, T[ch]);
    }
    v3 w = s[l].w, v = s[l].v;
    w = V3(fmaf(acc.x, m->dt, w.x * m->ang_damp), fmaf(acc.y, m->dt, w.y * m->ang_damp), fmaf(acc.z, m->dt, w.z * m->ang_damp));
    const v3 al = vadd(m->g, Fa[l]); /* g + 0 = g bit for bit on links without slide dofs */
    v = V3(fmaf(al.x, m->dt, v.x * m->vel_damp), fmaf(al.y, m->dt, v.y * m->vel_damp), fmaf(al.z, m->dt, v.z * m->vel_damp));
    q4 q = s[l].q;
    q = qnormalize(qadd(q, vqmul(vscale(w, m->half_dt), q)));
    s[l].p = vfma(v, m->dt, s[l].p);
    s[l].q = q; s[l].w = w; s[l].v = v;
    before[l] = s[l];
  }
  /* ---- joints.position_update (XPBD, Jacobi over joints) ---------------------------- */
  v3 dpc[MBD_MAXL], dpp[MBD_MAXL];
  q4 dqc[MBD_MAXL], dqp[MBD_MAXL];
  for (int l = 0; l < L; ++l) {
    dpc[l] = dpp[l] = V3(0, 0, 0);
    dqc[l] = dqp[l] = Q4(0, 0, 0, 0);
    int ndof = LFi(m, MBD_F_NDOF, l);
    if (ndof <= 0) continue;
    int par = LFi(m, MBD_F_PARENT, l);
    v3 pp = par >= 0 ? s[par].p : V3(0, 0, 0);
    q4 qp = par >= 0 ? s[par].q : Q4(1, 0, 0, 0);
    float im_c = LFf(m, MBD_F_INV_MASS, l), im_p = LFf(m, MBD_F_PINV_MASS, l), ii_p = LFf(m, MBD_F_PINV_INERTIA, l);
    /* translation: pull the child anchor a_c.pos onto the parent anchor a_p.pos */
    v3 rpw = vrotate(LF3(m, MBD_F_RP, l), qp);
    v3 rcw = vrotate(LF3(m, MBD_F_RC, l), s[l].q);
    v3 e = vsub(vadd(s[l].p, rcw), vadd(pp, rpw));
    const int smask = LFi(m, MBD_F_SLIDE, l);
    if (smask) { /* prismatic dofs: what the limits allow along each slide axis is not an error */
      q4 a_ps = qmul(qp, LF4(m, MBD_F_PQ, l));
      for (int k = 0; k < ndof; ++k)
        if ((smask >> k) & 1) {
          int base = MBD_F_DOF0 + k * MBD_DOF_STRIDE;
          v3 ak = slide_axis(k, LFf(m, MBD_F_PARITY, l), a_ps);
          float x = vdot(e, ak);
          e = vfma(ak, -clampf(x, LFf(m, base + MBD_D_LO, l), LFf(m, base + MBD_D_HI, l)), e);
        }
    }
    float c;
    v3 n = vnormalize(e, &c);
    v3 crc = vcross(rcw, n), crp = vcross(rpw, n);
    float w_c = im_c + vdot(crc, crc);
    float w_p = fmaf(ii_p, vdot(crp, crp), im_p);
    float dl = -c / (w_p + w_c + ORC_EPS);
    v3 P = vscale(n, dl);
    v3 dp_c = vscale(P, im_c);
    q4 dq_c = qscale(vqmul(vcross(rcw, P), s[l].q), 0.5f);
    v3 dp_p = vscale(P, -im_p);
    q4 dq_p = qscale(vqmul(vcross(rpw, P), qp), -0.5f * ii_p);
    /* rotation: hinge axis alignment (1 dof) / zero the missing Euler angle (2 dof) + limits */
    q4 a_p = qmul(qp, LF4(m, MBD_F_PQ, l));
    q4 a_c = qmul(s[l].q, LF4(m, MBD_F_JQ, l));
    q4 j = qmul(qconj(a_p), a_c);
    JointAngles ja;
    axis_angle_ang(j, LFf(m, MBD_F_PARITY, l), &ja);
    v3 dqj;
    {
      int b0 = MBD_F_DOF0;
      /* a slide dof leaves no rotational freedom in its slot: the whole angle is the error */
      float e0 = (smask & 1) ? ja.ang[0] : ja.ang[0] - clampf(ja.ang[0], LFf(m, b0 + MBD_D_LO, l), LFf(m, b0 + MBD_D_HI, l));
      if (ndof == 1) {
        /* dq = cross(axis_p, axis_c) + axis * (angle - clip(angle)); in the a_p frame axis_p = e_x,
         * axis_c = first column of R(j): cross = (0, -r20, r10) */
        dqj = V3(e0, -ja.r20, ja.r10);
      } else {
        int b1 = MBD_F_DOF0 + MBD_DOF_STRIDE, b2 = MBD_F_DOF0 + 2 * MBD_DOF_STRIDE;
        float e1 = (smask & 2) ? ja.ang[1] : ja.ang[1] - clampf(ja.ang[1], LFf(m, b1 + MBD_D_LO, l), LFf(m, b1 + MBD_D_HI, l));
        float e2 = (smask & 4) ? ja.ang[2] : ja.ang[2] - clampf(ja.ang[2], LFf(m, b2 + MBD_D_LO, l), LFf(m, b2 + MBD_D_HI, l));
        dqj = vscale(ja.ax[0], e0);
        dqj = vfma(ja.ax[1], e1, dqj);
        dqj = vfma(ja.ax[2], e2, dqj);
      }
    }
    v3 dq = vrotate(dqj, a_p);
    float th;
    v3 na = vnormalize(dq, &th);
    float nn = vdot(na, na);
    float dla = -th / (fmaf(ii_p, nn, nn) + ORC_EPS);
    v3 Pa = vscale(na, dla);
    q4 dqa_c = qscale(vqmul(Pa, s[l].q), 0.5f);
    q4 dqa_p = qscale(vqmul(Pa, qp), -0.5f * ii_p);
    dpc[l] = vscale(dp_c, m->scale_pos);
    dpp[l] = vscale(dp_p, m->scale_pos);
    dqc[l] = qadd(qscale(dq_c, m->scale_pos), qscale(dqa_c, m->scale_ang));
    dqp[l] = qadd(qscale(dq_p, m->scale_pos), qscale(dqa_p, m->scale_ang));
  }
  for (int l = 0; l < L; ++l) {
    v3 dp = dpc[l];
    q4 dq = dqc[l];
    for (int c = 0; c < MBD_MAXCHILD; ++c) {
      int ch = LFi(m, MBD_F_CHILD0 + c, l);
      if (ch >= 0) { dp = vadd(dp, dpp[ch]); dq = qadd(dq, dqp[ch]); }
    }
    s[l].p = vadd(s[l].p, dp);
    s[l].q = qnormalize(qadd(s[l].q, dq));
  }
  /* ---- contact.get (sphere-plane, MJX plane_sphere) + collisions.resolve_position ---- */
  float dlam[MBD_MAXL][MBD_MAXCON];
  v3 cpos[MBD_MAXL][MBD_MAXCON];
  for (int l = 0; l < L; ++l) {
    int ncon = LFi(m, MBD_F_NCON, l);
    if (ncon <= 0) continue;
    float im = LFf(m, MBD_F_INV_MASS, l);
    v3 dp = V3(0, 0, 0);
    q4 dq = Q4(0, 0, 0, 0);
    const v3 p0 = s[l].p;
    const q4 q0 = s[l].q;
    for (int ci = 0; ci < ncon; ++ci)
      contact_position_plane(m, l, ci, im, p0, q0, prev[l].p, prev[l].q, &dp, &dq, &dlam[l][ci], &cpos[l][ci]);
    s[l].p = vfma(dp, m->collide_scale, s[l].p);
    s[l].q = qnormalize(qadd(s[l].q, qscale(dq, m->collide_scale)));
  }
  /* ---- integrator.project_xd ---------------------------------------------------------- */
  for (int l = 0; l < L; ++l) {
    s[l].v = vscale(vsub(s[l].p, prev[l].p), m->inv_dt);
    q4 dq = qmul(s[l].q, qconj(prev[l].q)); /* math.relative_quat(prev, cur) */
    float sc = dq.w >= 0.0f ? m->two_inv_dt : -m->two_inv_dt;
    s[l].w = V3(dq.x * sc, dq.y * sc, dq.z * sc);
  }
  /* ---- collisions.resolve_velocity ------------------------------------------------------ */
  for (int l = 0; l < L; ++l) {
    int ncon = LFi(m, MBD_F_NCON, l);
    if (ncon <= 0) continue;
    float im = LFf(m, MBD_F_INV_MASS, l);
    v3 dv = V3(0, 0, 0), dw = V3(0, 0, 0);
    const v3 v0 = s[l].v, w0 = s[l].w;
    for (int ci = 0; ci < ncon; ++ci)
      contact_velocity_plane(m, l, ci, im, s[l].p, v0, w0, before[l].v, before[l].w, cpos[l][ci], dlam[l][ci], &dv, &dw);
    s[l].v = vadd(s[l].v, dv);
    s[l].w = vadd(s[l].w, dw);
  }
}

/* com.to_world: x.pos = x_i.pos - rotate(com, rot); xd.vel = xd_i.vel + cross(rotate(com,rot), ang) */
static inline v3 link_origin(const Model* m, const Link* s, int l) {
  return vsub(s[l].p, vrotate(LF3(m, MBD_F_COM, l), s[l].q));
}
static inline v3 link_origin_vel(const Model* m, const Link* s, int l) {
  v3 rc = vrotate(LF3(m, MBD_F_COM, l), s[l].q);
  return vadd(s[l].v, vcross(rc, s[l].w));
}

static float reward_post(const Model* m, const Link* s) {
  v3 x0 = link_origin(m, s, 0);
  if (m->reward == MBD_REWARD_HUMANOIDRUN) {
    /* humanoidrun.py:46-51 */
    float dz = clampf(fabsf(x0.z - 1.3f), -1.0f, 1.0f);
    return (x0.x - dz) - fabsf(x0.y) * 0.1f;
  }
  if (m->reward == MBD_REWARD_HUMANOIDSTANDUP) {
    /* humanoidstandup.py:50-56 */
    return ((1.5f - clampf(fabsf(x0.z - 1.3f), -2.0f, 1.0f)) - fabsf(x0.x) * 0.1f) - fabsf(x0.y) * 0.1f;
  }
  if (m->reward == MBD_REWARD_HOPPER) {
    /* hopper.py:57-65 (RW0 = 1.0), walker2d.py:56-61 (RW0 = 1.1) */
    return x0.x - clampf(fabsf(x0.z - m->f[MBD_H_RW0]), -1.0f, 1.0f) * 0.5f;
  }
  if (m->reward == MBD_REWARD_CARTPOLE) {
    /* cartpole.py:44: cos(q[1]) - |qd[0]|.  q[1] = hinge angle of link 1 against link 0 (kinematics.inverse: the
     * angle psi of the joint-frame relative rotation), qd[0] = velocity of link 0's slide dof (anchor velocity
     * along the slide axis; the world anchor is at rest) */
    q4 a_p = qmul(s[0].q, LF4(m, MBD_F_PQ, 1));
    q4 a_c = qmul(s[1].q, LF4(m, MBD_F_JQ, 1));
    JointAngles ja;
    axis_angle_ang(qmul(qconj(a_p), a_c), LFf(m, MBD_F_PARITY, 1), &ja);
    v3 rcw = vrotate(LF3(m, MBD_F_RC, 0), s[0].q);
    v3 va = vadd(s[0].v, vcross(s[0].w, rcw));
    float xd = vdot(va, slide_axis(0, LFf(m, MBD_F_PARITY, 0), LF4(m, MBD_F_PQ, 0)));
    return mbd_cosf(ja.ang[0]) - fabsf(xd);
  }
  return 0.0f;
}
/* brax/envs/ant.py [brax-recalled; the env is Brax's stock one, /root/reference/mbd/envs/__init__.py:30-31]:
 *   velocity = (x.pos[0] - x0.pos[0]) / dt;  reward = forward_reward + healthy_reward - ctrl_cost - contact_cost
 * with healthy_reward paid unconditionally (terminate_when_unhealthy=True), contact cost off, ctrl_cost =
 * weight * sum(action^2) on the action handed to env.step.  Sum of squares taken in action order. */
static float reward_ant(const Model* m, float x_before, float x_after, const float* u) {
  const float env_dt = m->f[MBD_H_RW0], healthy = m->f[MBD_H_RW0 + 1], wc = m->f[MBD_H_RW0 + 2];
  float fwd = (x_after - x_before) / env_dt;
  float ss = 0.0f;
  for (int k = 0; k < m->nu; ++k) ss = ss + u[k] * u[k];
  return (fwd + healthy) - wc * ss;
}
static float reward_pre(const Model* m, const Link* s) {
  /* humanoidtrack.py:87-96 — evaluated on the state BEFORE the step */
  v3 x0 = link_origin(m, s, 0);
  v3 v0 = link_origin_vel(m, s, 0);
  return 1.0f + ((-fabsf(v0.x - 1.6f) - fabsf(x0.z - 1.3f)) - fabsf(x0.y) * 0.1f);
}

static void load_state(Link* s, const float* st, int L) {
  for (int l = 0; l < L; ++l) {
    const float* a = st + l * MBD_STATE_STRIDE;
    s[l].p = V3(a[0], a[1], a[2]); s[l].q = Q4(a[3], a[4], a[5], a[6]);
    s[l].w = V3(a[7], a[8], a[9]); s[l].v = V3(a[10], a[11], a[12]);
  }
}
static void store_state(const Link* s, float* st, int L) {
  for (int l = 0; l < L; ++l) {
    float* a = st + l * MBD_STATE_STRIDE;
    a[0] = s[l].p.x; a[1] = s[l].p.y; a[2] = s[l].p.z;
    a[3] = s[l].q.w; a[4] = s[l].q.x; a[5] = s[l].q.y; a[6] = s[l].q.z;
    a[7] = s[l].w.x; a[8] = s[l].w.y; a[9] = s[l].w.z;
    a[10] = s[l].v.x; a[11] = s[l].v.y; a[12] = s[l].v.z;
  }
}

/* vmap(rollout_us)(state_init, Y0s) — mbd_planner.py:109, utils.py:14-20.
 * state_init [L,13] shared by all samples; Y0s [n,H,nu]; outputs:
 *   rewss [n,H] (may be NULL), rews [n] = rewss.mean(-1), logpd [n] (NULL unless xref given),
 *   final_state [n,L,13] (may be NULL), track_pos [n,H,ntrack,3] (may be NULL).
 * xref [ntrack, href, 3].  nsub_override > 0 replaces n_frames (per-substep parity tests). */
ORC_API int orc_xpbd_rollout(const uint32_t* blob, const float* state_init, const float* Y0s, int n, int H,
                             float* rewss, float* rews, const float* xref, int href, float* logpd,
                             float* final_state, float* track_pos, int nsub_override, int nthreads) {
  Model m;
  if (model_open(&m, blob)) return -1;
  const int L = m.L, nu = m.nu;
  const int nsub = nsub_override > 0 ? nsub_override : m.n_frames;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    Link s[MBD_MAXL];
    load_state(s, state_init, L);
    float sum = 0.0f;
    float acc_k[MBD_MAXTRACK];
    for (int k = 0; k < MBD_MAXTRACK; ++k) acc_k[k] = 0.0f;
    for (int t = 0; t < H; ++t) {
      const float* u = Y0s + ((size_t)i * H + t) * nu;
      float r_pre = (m.reward == MBD_REWARD_HUMANOIDTRACK) ? reward_pre(&m, s) : 0.0f;
      const float x_before = (m.reward == MBD_REWARD_ANT) ? link_origin(&m, s, 0).x : 0.0f;
      for (int f = 0; f < nsub; ++f) positional_step(&m, s, u); /* PipelineEnv.pipeline_step */
      float r = (m.reward == MBD_REWARD_HUMANOIDTRACK) ? r_pre
                : (m.reward == MBD_REWARD_ANT) ? reward_ant(&m, x_before, link_origin(&m, s, 0).x, u) : reward_post(&m, s);
      if (rewss) rewss[(size_t)i * H + t] = r;
      sum += r;
      for (int k = 0; k < m.ntrack; ++k) {
        v3 x = link_origin(&m, s, m.i[MBD_H_TRACK0 + k]);
        if (track_pos) {
          float* o = track_pos + (((size_t)i * H + t) * m.ntrack + k) * 3;
          o[0] = x.x; o[1] = x.y; o[2] = x.z;
        }
        if (xref) {
          /* humanoidtrack.py:98-106: ((clip(|xs - xref|, 0, .5)/.5)^2).mean(); t >= href clamps
           * to the last reference row (extension, not reference behaviour: SURVEY F9) */
          int tt = t < href ? t : href - 1;
          const float* xr = xref + ((size_t)k * href + tt) * 3;
          v3 d = V3(x.x - xr[0], x.y - xr[1], x.z - xr[2]);
          float nr = sqrtf(vdot(d, d));
          float cl = nr < 0.5f ? nr : 0.5f;
          float q = cl / 0.5f;
          acc_k[k] = fmaf(q, q, acc_k[k]);
        }
      }
    }
    rews[i] = sum / (float)H;
    if (logpd && xref) {
      float tot = 0.0f;
      for (int k = 0; k < m.ntrack; ++k) tot += acc_k[k];
      logpd[i] = 0.0f - tot / (float)(m.ntrack * H);
    }
    if (final_state) store_state(s, final_state + (size_t)i * L * MBD_STATE_STRIDE, L);
  }
  return 0;
}

/* ================================================================================== */
/* car2d (/root/reference/mbd/envs/car2d.py)                                           */
/* ================================================================================== */
#define CAR_NOBS 11
static inline void car_dynamics(const float* x, const float* u, float* o) { /* car2d.py:10-19 */
  float s, c;
  mbd_sincosf(x[2], &s, &c);
  o[0] = u[1] * s * 3.0f;
  o[1] = u[1] * c * 3.0f;
  o[2] = u[0] * 3.14159274101257324f / 3.0f * 2.0f;
}
static inline void car_rk4(const float* x, const float* u, float dt, float hdt, float sdt, float* xn) { /* car2d.py:22-27 */
  float k1[3], k2[3], k3[3], k4[3], y[3]; /* hdt = f32(dt/2), sdt = f32(dt/6): Python-double constants */
  car_dynamics(x, u, k1);
  for (int i = 0; i < 3; ++i) y[i] = x[i] + hdt * k1[i];
  car_dynamics(y, u, k2);
  for (int i = 0; i < 3; ++i) y[i] = x[i] + hdt * k2[i];
  car_dynamics(y, u, k3);
  for (int i = 0; i < 3; ++i) y[i] = x[i] + dt * k3[i];
  car_dynamics(y, u, k4);
  for (int i = 0; i < 3; ++i) xn[i] = x[i] + sdt * (((k1[i] + 2.0f * k2[i]) + 2.0f * k3[i]) + k4[i]);
}
static inline float car_reward(const float* q) { /* car2d.py:88-93 */
  float dx = q[0] - 0.5f, dy = q[1] - 0.0f;
  float d = sqrtf(dx * dx + dy * dy);
  float c = clampf(d, 0.0f, 0.2f) / 0.2f;
  return 1.0f - c * c;
}
/* params: [obs_center (11x2), obs_radius, dt, dt/2, dt/6]; x0[3]; Y0s [n,H,2]; xref [href,2] or NULL */
ORC_API int orc_car2d_rollout(const float* params, const float* x0, const float* Y0s, int n, int H,
                              float* rewss, float* rews, const float* xref, int href, float* logpd,
                              float* traj, int nthreads) {
  const float* oc = params;
  const float orad = params[2 * CAR_NOBS], dt = params[2 * CAR_NOBS + 1];
  const float hdt = params[2 * CAR_NOBS + 2], sdt = params[2 * CAR_NOBS + 3];
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    float q[3] = {x0[0], x0[1], x0[2]};
    float sum = 0.0f, acc = 0.0f;
    for (int t = 0; t < H; ++t) {
      const float* ur = Y0s + ((size_t)i * H + t) * 2;
      float u[2] = {clampf(ur[0], -1.0f, 1.0f), clampf(ur[1], -1.0f, 1.0f)}; /* car2d.py:80 */
      float qn[3];
      car_rk4(q, u, dt, hdt, sdt, qn);
      int collide = 0; /* car2d.py:30-32 */
      for (int k = 0; k < CAR_NOBS; ++k) {
        float dx = qn[0] - oc[2 * k], dy = qn[1] - oc[2 * k + 1];
        if (sqrtf(dx * dx + dy * dy) < orad) collide = 1;
      }
      if (!collide) { q[0] = qn[0]; q[1] = qn[1]; q[2] = qn[2]; }
      float r = car_reward(q);
      if (rewss) rewss[(size_t)i * H + t] = r;
      sum += r;
      if (traj) { float* o = traj + ((size_t)i * H + t) * 3; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; }
      if (xref) { /* car2d.py:95-102 */
        int tt = t < href ? t : href - 1;
        float ex = q[0] - xref[2 * tt], ey = q[1] - xref[2 * tt + 1];
        float d = sqrtf(ex * ex + ey * ey);
        float c = clampf(d, 0.0f, 0.5f) / 0.5f;
        acc += c * c;
      }
    }
    rews[i] = sum / (float)H;
    if (logpd && xref) logpd[i] = 0.0f - acc / (float)H;
  }
  return 0;
}

/* ================================================================================== */
/* JAX PRNG                                                                            */
/* ================================================================================== */
/* threefry layout of the samplers below: 0 legacy (pinned by JAX's known-answer vectors), 1 partitionable [jax-recalled] */
static int g_orc_prng_part = 0;
ORC_API void orc_set_prng_layout(int partitionable) { g_orc_prng_part = partitionable ? 1 : 0; }
#define ORC_TOTAL(t) (g_orc_prng_part ? 0u : (uint32_t)(t))

ORC_API void orc_threefry2x32(const uint32_t* key, const uint32_t* ctr, uint32_t* out) {
  mbd_threefry2x32(key[0], key[1], ctr[0], ctr[1], &out[0], &out[1]);
}
/* jax.random.bits(key, (total,)) */
ORC_API void orc_random_bits(const uint32_t* key, uint32_t total, uint32_t* out) {
  for (uint32_t i = 0; i < total; ++i) out[i] = mbd_random_bits_at(key[0], key[1], i, ORC_TOTAL(total));
}
/* jax.random.normal(key, shape) flattened; [begin, end) of `total` elements */
ORC_API void orc_normal(const uint32_t* key, uint32_t total, uint32_t begin, uint32_t end, float* out, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t i = begin; i < (int64_t)end; ++i)
    out[i - begin] = mbd_bits_to_normal(mbd_random_bits_at(key[0], key[1], (uint32_t)i, ORC_TOTAL(total)));
}
/* reverse_once sampling, mbd_planner.py:103-106: clip(eps*sigma + Ybar, -1, 1); Ybar [H*nu] */
ORC_API void orc_sample_Y0s(const uint32_t* key, int n_total, int n_begin, int n_end, int HNu, float sigma,
                            const float* Ybar, float* out, int nthreads) {
  const uint32_t total = (uint32_t)n_total * (uint32_t)HNu;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int i = n_begin; i < n_end; ++i)
    for (int e = 0; e < HNu; ++e) {
      uint32_t idx = (uint32_t)i * (uint32_t)HNu + (uint32_t)e;
      float eps = mbd_bits_to_normal(mbd_random_bits_at(key[0], key[1], idx, ORC_TOTAL(total)));
      float y = eps * sigma + Ybar[e];
      out[(size_t)(i - n_begin) * HNu + e] = clampf(y, -1.0f, 1.0f);
    }
}

/* scalar math spec, exported so tests can check accuracy against float64 */
ORC_API float orc_atan2f(float y, float x) { return mbd_atan2f(y, x); }
ORC_API float orc_sinf(float x) { return mbd_sinf(x); }
ORC_API float orc_cosf(float x) { return mbd_cosf(x); }
ORC_API float orc_logf(float x) { return mbd_logf(x); }
ORC_API float orc_expf(float x) { return mbd_expf(x); }
ORC_API float orc_erfinvf(float x) { return mbd_erfinvf(x); }
ORC_API void orc_map(int fn, const float* a, const float* b, float* out, int n) {
  for (int i = 0; i < n; ++i) {
    switch (fn) {
      case 0: out[i] = mbd_atan2f(a[i], b[i]); break;
      case 1: out[i] = mbd_sinf(a[i]); break;
      case 2: out[i] = mbd_cosf(a[i]); break;
      case 3: out[i] = mbd_logf(a[i]); break;
      case 4: out[i] = mbd_expf(a[i]); break;
      case 5: out[i] = mbd_erfinvf(a[i]); break;
    }
  }
}
ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
