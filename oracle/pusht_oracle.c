/* pusht_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, never on the product path) for the pushT env.
 *
 * Restates what /root/reference/mbd/envs/pushT.py:16-66 computes: `pipeline_step` of Brax's GENERALIZED backend
 * (brax/generalized/pipeline.py::step — an un-vendored dependency of the reference, restated from the published algorithm
 * [brax-recalled]) on the planar three-body model of /root/reference/mbd/assets/pushT.xml, plus the env's reward.
 *
 * PARITY UNPINNED: neither Brax nor JAX can be installed here, the reference holds no golden vectors for this env.  What is
 * restated and how sure it is:
 *   structure of a step            tau -> qf_smooth -> qf_constraint -> integrate                      [brax-recalled]
 *   qf_smooth                      actuation + joint damping - centrifugal bias (RNE of a planar body)  [rigid-body mechanics]
 *   soft constraints               MuJoCo solref (0.02, 1) / solimp (0.9, 0.95, 0.001, 0.5, 2):
 *                                  imp(pos), aref = -b vel - k imp pos                                  [brax-recalled == MuJoCo docs]
 *   constraint rows                joint limits; one sphere-box contact per box, 4-sided friction pyramid
 *                                  n -+ mu t for t in the two tangents [brax-recalled]; the out-of-plane pair has
 *                                  the in-plane Jacobian n twice and is merged into ONE row with half the
 *                                  regulariser (same QP minimiser, well-conditioned for Gauss-Seidel)   [own choice]
 *   regulariser                    R = (1 - imp) / imp * diag(J M^-1 J^T): the EXACT diagonal where Brax
 *                                  uses an inverse-weight approximation                                 [own choice]
 *   constraint solve               projected Gauss-Seidel, at most MBD_PT_ITERS sweeps, stopped when a sweep moves the
 *                                  constraint force J^T x by no more than MBD_PT_TOL (1e-6) of its largest component, on the QP
 *                                  min 1/2 x^T (J M^-1 J^T + R) x + x^T (J M^-1 qf_smooth - aref), x >= 0;
 *                                  A is SPD, the minimiser is unique, so any convergent solver agrees with
 *                                  Brax's up to its truncation error                                   [own solver]
 *   integration                    qd += dt (M + dt D)^-1 (qf_smooth + J^T x);  q += dt qd             [brax-recalled]
 *   contact point                  midway between the two surfaces                                     [brax-recalled]
 * All of it is fp32 with the contraction rules of include/mbd_fp32.h; the CUDA kernel (mbd_b200/csrc/pusht.cuh) is
 * compared with this file bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "mbd_fp32.h"
#include "mbd_pusht.h"

#define ORC_API __attribute__((visibility("default")))

/* Variant switches for the day a Brax install can be compared (scripts/pin_against_brax.py builds the grid with `make variant`).
 * The defaults are what the CUDA kernel implements; every other value is a candidate reading of Brax [brax-recalled]. */
#ifndef ORC_PT_REG_INVWEIGHT   /* 1: regulariser from MuJoCo-style inverse weights at qpos0 instead of the exact diagonal:        */
#define ORC_PT_REG_INVWEIGHT 0 /*    limits R = w_dof (1-imp)/imp; pyramid edges R = 2 mu^2 (1 + mu^2) (w_a + w_b) (1-imp)/imp    */
#endif
#ifndef ORC_PT_CONTACT_MIDPOINT /* 0: contact point on the box surface instead of midway between the two surfaces */
#define ORC_PT_CONTACT_MIDPOINT 1
#endif

static inline float pt_clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* impedance and reference acceleration of one constraint row (MuJoCo solver parameters, power = 2) */
static void pt_imp_aref(const float* P, float pos, float vel, float* imp, float* aref) {
  const float dmin = P[MBD_PT_DMIN], dmax = P[MBD_PT_DMAX], mid = P[MBD_PT_MID];
  float x = fabsf(pos) / P[MBD_PT_WIDTH];
  float a = (1.0f / mid) * (x * x);
  float omx = 1.0f - x;
  float b = 1.0f - (1.0f / (1.0f - mid)) * (omx * omx);
  float y = x < mid ? a : b;
  float d = pt_clampf(dmin + y * (dmax - dmin), dmin, dmax);
  if (x > 1.0f) d = dmax;
  *imp = d;
  *aref = (0.0f - P[MBD_PT_KB] * vel) - (P[MBD_PT_KK] * d) * pos;
}

typedef struct {
  float J[MBD_PT_NROW][5]; /* pusher x, y | slider x, y, theta */
  float pos[MBD_PT_NROW];
  float rscale[MBD_PT_NROW]; /* regulariser weight: 1, or 1/2 for a row that stands for two identical pyramid rows */
  int active[MBD_PT_NROW];
} pt_rows;

/* one physics step; u = the two motor controls (already clipped) */
static void pt_substep(const float* P, float* q, float* qd, const float* u) {
  const float dt = P[MBD_PT_DT];
  const float ms = P[MBD_PT_MS], ims = P[MBD_PT_IMS], Is = P[MBD_PT_IS], iIs = P[MBD_PT_IIS];
  float s, c;
  mbd_sincosf(q[4], &s, &c);
  /* slider COM offset in the world frame, r = R(theta) c_body */
  const float rx = c * P[MBD_PT_CX] - s * P[MBD_PT_CY];
  const float ry = s * P[MBD_PT_CX] + c * P[MBD_PT_CY];
  /* ---- qf_smooth = actuation + passive (joint damping) - bias (centrifugal term of the offset COM) */
  float f[5];
  const float w = qd[4], mw2 = ms * (w * w);
  f[0] = P[MBD_PT_GEAR0] * u[0] - P[MBD_PT_DPX] * qd[0];
  f[1] = P[MBD_PT_GEAR1] * u[1] - P[MBD_PT_DPY] * qd[1];
  f[2] = mw2 * rx - P[MBD_PT_DSX] * qd[2];
  f[3] = mw2 * ry - P[MBD_PT_DSY] * qd[3];
  f[4] = 0.0f - P[MBD_PT_DSTH] * w;
  /* ---- M^-1 of the slider block: M = [[m, 0, -m ry], [0, m, m rx], [-m ry, m rx, I + m |r|^2]]; its Schur complement is I */
  const float A00 = ims + (ry * ry) * iIs, A01 = 0.0f - (rx * ry) * iIs, A02 = ry * iIs;
  const float A11 = ims + (rx * rx) * iIs, A12 = 0.0f - rx * iIs, A22 = iIs;
  /* ---- constraint rows */
  pt_rows R;
  memset(&R, 0, sizeof(R));
  for (int k = 0; k < MBD_PT_NLIM; ++k) {
    const float pmin = q[k] - P[MBD_PT_LIM0 + 2 * k], pmax = P[MBD_PT_LIM0 + 2 * k + 1] - q[k];
    const float pm = pmin < pmax ? pmin : pmax;
    R.pos[k] = pm < 0.0f ? pm : 0.0f;
    R.active[k] = pm < 0.0f;
    R.J[k][k] = pmin < pmax ? 1.0f : -1.0f;
    R.rscale[k] = 1.0f;
  }
  const float mu = P[MBD_PT_MU], rp = P[MBD_PT_RP];
  for (int b = 0; b < MBD_PT_NBOX; ++b) {
    const float* B = P + MBD_PT_BOX0 + 4 * b;
    const float bx = q[2] + (c * B[0] - s * B[1]), by = q[3] + (s * B[0] + c * B[1]);
    const float dx = q[0] - bx, dy = q[1] - by;
    const float lx = c * dx + s * dy, ly = c * dy - s * dx; /* sphere centre in the box frame */
    const float clx = pt_clampf(lx, -B[2], B[2]), cly = pt_clampf(ly, -B[3], B[3]);
    const float ex = lx - clx, ey = ly - cly;
    const float d2 = ex * ex + ey * ey;
    float nlx, nly, dist, sx = clx, sy = cly; /* normal (box -> sphere) and the box surface point, box frame */
    if (d2 > 0.0f) {
      const float d = sqrtf(d2);
      nlx = ex / d; nly = ey / d;
      dist = d - rp;
    } else { /* centre inside the box: leave through the nearest face */
      const float px = B[2] - fabsf(lx), py = B[3] - fabsf(ly);
      if (px < py) { nlx = lx < 0.0f ? -1.0f : 1.0f; nly = 0.0f; dist = (0.0f - px) - rp; sx = nlx * B[2]; }
      else { nlx = 0.0f; nly = ly < 0.0f ? -1.0f : 1.0f; dist = (0.0f - py) - rp; sy = nly * B[3]; }
    }
    const float nx = c * nlx - s * nly, ny = s * nlx + c * nly;
    const float half = ORC_PT_CONTACT_MIDPOINT ? 0.5f * dist : 0.0f;
    const float ax = B[0] + (sx + nlx * half), ay = B[1] + (sy + nly * half); /* contact point, body frame */
    const float rhox = c * ax - s * ay, rhoy = s * ax + c * ay;                /* arm from the slider origin */
    const float tx = 0.0f - ny, ty = nx;
    /* the 4-sided pyramid: n -+ mu t (in plane) and n -+ mu z.  The out-of-plane pair has the SAME in-plane Jacobian n: two
     * identical rows with regulariser R each are, for the QP, one row with regulariser R/2 carrying the sum of the two
     * multipliers — merged here, because Gauss-Seidel converges at rate ~imp (0.9-0.95 per sweep) on the duplicated pair. */
    const float dirs[MBD_PT_NCROW][2] = {{nx - mu * tx, ny - mu * ty}, {nx + mu * tx, ny + mu * ty}, {nx, ny}};
    for (int j = 0; j < MBD_PT_NCROW; ++j) {
      const int r = MBD_PT_NLIM + MBD_PT_NCROW * b + j;
      const float ddx = dirs[j][0], ddy = dirs[j][1];
      R.J[r][0] = ddx; R.J[r][1] = ddy;
      R.J[r][2] = 0.0f - ddx; R.J[r][3] = 0.0f - ddy;
      R.J[r][4] = 0.0f - (rhox * ddy - rhoy * ddx);
      R.pos[r] = dist;
      R.rscale[r] = j == 2 ? 0.5f : 1.0f;
      R.active[r] = dist < 0.0f;
    }
  }
  /* ---- constraint QP and projected Gauss-Seidel.  The active rows are compacted to the front and the system is padded with
   * identity rows (x stays 0 there) to 4, 8 or 12 rows: the kernel keeps the 4- and 8-row systems in registers, and both sides
   * sum the same padded terms in the same order. */
  float MiJ[12][5];
  float Mif[5];
  const float imp_ = P[MBD_PT_IMP];
  Mif[0] = imp_ * f[0]; Mif[1] = imp_ * f[1];
  Mif[2] = (A00 * f[2] + A01 * f[3]) + A02 * f[4];
  Mif[3] = (A01 * f[2] + A11 * f[3]) + A12 * f[4];
  Mif[4] = (A02 * f[2] + A12 * f[3]) + A22 * f[4];
  int idx[MBD_PT_NROW], nr = 0;
  for (int r = 0; r < MBD_PT_NROW; ++r)
    if (R.active[r]) idx[nr++] = r;
  float ftot[5] = {f[0], f[1], f[2], f[3], f[4]};
  if (nr > 0) {
    const int nrp = nr <= 4 ? 4 : (nr <= 8 ? 8 : 12);
    float A[12][12], bq[12], invD[12], x[12]; /* padded sizes: 4, 8 or 12 >= MBD_PT_NROW */
    for (int i = 0; i < nr; ++i) {
      const float* J = R.J[idx[i]];
      MiJ[i][0] = imp_ * J[0]; MiJ[i][1] = imp_ * J[1];
      MiJ[i][2] = (A00 * J[2] + A01 * J[3]) + A02 * J[4];
      MiJ[i][3] = (A01 * J[2] + A11 * J[3]) + A12 * J[4];
      MiJ[i][4] = (A02 * J[2] + A12 * J[3]) + A22 * J[4];
    }
    for (int i = 0; i < nrp; ++i)
      for (int j = 0; j < nrp; ++j) {
        float a = 0.0f;
        if (i < nr && j < nr) {
          const float* J = R.J[idx[i]];
          a = (((J[0] * MiJ[j][0] + J[1] * MiJ[j][1]) + J[2] * MiJ[j][2]) + J[3] * MiJ[j][3]) + J[4] * MiJ[j][4];
        }
        A[i][j] = a;
      }
    for (int i = 0; i < nrp; ++i) {
      x[i] = 0.0f;
      if (i < nr) {
        const float* J = R.J[idx[i]];
        const float vel = (((J[0] * qd[0] + J[1] * qd[1]) + J[2] * qd[2]) + J[3] * qd[3]) + J[4] * qd[4];
        float imp, aref;
        pt_imp_aref(P, R.pos[idx[i]], vel, &imp, &aref);
        const float arr = A[i][i];
#if ORC_PT_REG_INVWEIGHT
        {
          /* translational inverse weights at qpos0 (MuJoCo body_invweight0: trace(J M^-1 J^T) / 3 at the body COM, the z direction
           * has no dof) and dof inverse weights (diagonal of M^-1 at qpos0) */
          const float wp = (2.0f * P[MBD_PT_IMP]) / 3.0f, ws = (2.0f * ims) / 3.0f;
          const int r = idx[i];
          float wr;
          if (r < MBD_PT_NLIM) wr = r < 2 ? P[MBD_PT_IMP] : (r == 2 ? ims + (P[MBD_PT_CY] * P[MBD_PT_CY]) * iIs : ims + (P[MBD_PT_CX] * P[MBD_PT_CX]) * iIs);
          else wr = 2.0f * (mu * mu) * ((1.0f + mu * mu) * (wp + ws));
          A[i][i] = arr + (R.rscale[r] * ((1.0f - imp) / imp)) * wr;
        }
#else
        A[i][i] = arr + (R.rscale[idx[i]] * ((1.0f - imp) / imp)) * arr;
#endif
        invD[i] = 1.0f / A[i][i];
        bq[i] = ((((J[0] * Mif[0] + J[1] * Mif[1]) + J[2] * Mif[2]) + J[3] * Mif[3]) + J[4] * Mif[4]) - aref;
      } else {
        A[i][i] = 1.0f; invD[i] = 1.0f; bq[i] = 0.0f;
      }
    }
    const int iters = (int)P[MBD_PT_ITERS];
    /* Convergence is judged on the generalized constraint FORCE J^T x, not on x: the rows of one contact are linearly dependent
     * (n is the mean of n - mu t and n + mu t), so x keeps redistributing along the null space of J^T at rate ~imp per sweep long
     * after the force — the only thing the dynamics sees — has converged. */
    const float tol = P[MBD_PT_TOL];
    float F[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int it = 0; it < iters; ++it) {
      float dF[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      for (int i = 0; i < nrp; ++i) {
        float r0 = bq[i], r1 = 0.0f; /* two interleaved fused accumulators (even / odd columns): half the dependent chain */
        for (int j = 0; j < nrp; j += 2) { r0 = fmaf(A[i][j], x[j], r0); r1 = fmaf(A[i][j + 1], x[j + 1], r1); }
        const float res = r0 + r1;
        const float xn = x[i] - res * invD[i];
        const float xc = xn > 0.0f ? xn : 0.0f;
        const float dxi = xc - x[i];
        if (i < nr)
          for (int k = 0; k < 5; ++k) dF[k] = fmaf(R.J[idx[i]][k], dxi, dF[k]);
        x[i] = xc;
      }
      float dmax = 0.0f, fmx = 0.0f;
      for (int k = 0; k < 5; ++k) {
        dmax = fmaxf(dmax, fabsf(dF[k]));
        F[k] = F[k] + dF[k];
        fmx = fmaxf(fmx, fabsf(F[k]));
      }
      /* the force moved by at most TOL of its largest component (TOL = 0: no change at all) */
      if (dmax <= tol * fmx) break;
    }
    for (int i = 0; i < nr; ++i)
      for (int k = 0; k < 5; ++k) ftot[k] = ftot[k] + R.J[idx[i]][k] * x[i];
  }
  /* ---- semi-implicit Euler, joint damping folded into the mass matrix: (M + dt D) qdd = ftot */
  float qdd[5];
  qdd[0] = ftot[0] / (P[MBD_PT_MP] + dt * P[MBD_PT_DPX]);
  qdd[1] = ftot[1] / (P[MBD_PT_MP] + dt * P[MBD_PT_DPY]);
  {
    const float m1 = ms + dt * P[MBD_PT_DSX], m2 = ms + dt * P[MBD_PT_DSY];
    const float a = 0.0f - ms * ry, b = ms * rx;
    const float J3 = (Is + ms * (rx * rx + ry * ry)) + dt * P[MBD_PT_DSTH];
    const float g1 = ftot[2] / m1, g2 = ftot[3] / m2;
    const float den = (J3 - (a * a) / m1) - (b * b) / m2;
    const float x3 = ((ftot[4] - a * g1) - b * g2) / den;
    qdd[2] = g1 - (a * x3) / m1;
    qdd[3] = g2 - (b * x3) / m2;
    qdd[4] = x3;
  }
  for (int k = 0; k < 5; ++k) {
    qd[k] = qd[k] + qdd[k] * dt;
    q[k] = q[k] + qd[k] * dt;
  }
}

/* pushT.py:50-62 */
static float pt_reward(const float* q) {
  const float gx = q[5] - q[2], gy = q[6] - q[3];
  const float px = q[0] - q[2], py = q[1] - q[3];
  const float dps = sqrtf(px * px + py * py) - 0.2f;
  const float d_pusher2slider = dps > 0.0f ? dps : 0.0f;
  return 1.0f - ((sqrtf(gx * gx + gy * gy) + fabsf(q[7] - q[4]) / MBD_PI_F) + d_pusher2slider);
}

/* params [MBD_PT_NPARAM]; x0 [16] = q | qd; Y0s [n, H, 2]; rewss [n, H] or NULL; rews [n]; final [n, 16] or NULL;
 * traj [n, H, 16] or NULL (state after every env step) */
ORC_API int orc_pusht_rollout(const float* params, const float* x0, const float* Y0s, int n, int H, float* rewss, float* rews,
                              float* final_state, float* traj, int nthreads) {
  const int nsub = (int)params[MBD_PT_NSUB];
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    float q[MBD_PT_NQ], qd[MBD_PT_NQ];
    for (int k = 0; k < MBD_PT_NQ; ++k) { q[k] = x0[k]; qd[k] = x0[MBD_PT_NQ + k]; }
    float sum = 0.0f;
    for (int t = 0; t < H; ++t) {
      const float* ur = Y0s + ((size_t)i * H + t) * 2;
      const float u[2] = {pt_clampf(ur[0], -1.0f, 1.0f), pt_clampf(ur[1], -1.0f, 1.0f)}; /* motor ctrlrange */
      for (int k = 0; k < nsub; ++k) pt_substep(params, q, qd, u);
      const float r = pt_reward(q);
      if (rewss) rewss[(size_t)i * H + t] = r;
      sum += r;
      if (traj) {
        float* o = traj + ((size_t)i * H + t) * MBD_PT_STATE;
        for (int k = 0; k < MBD_PT_NQ; ++k) { o[k] = q[k]; o[MBD_PT_NQ + k] = qd[k]; }
      }
    }
    rews[i] = sum / (float)H;
    if (final_state)
      for (int k = 0; k < MBD_PT_NQ; ++k) { final_state[(size_t)i * MBD_PT_STATE + k] = q[k]; final_state[(size_t)i * MBD_PT_STATE + MBD_PT_NQ + k] = qd[k]; }
  }
  return 0;
}
