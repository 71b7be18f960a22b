// pusht.cuh — the pushT env (/root/reference/mbd/envs/pushT.py:16-66) on the GPU: one sample per thread.
//
// pushT is the reference's only env on Brax's `generalized` backend (reduced coordinates, mass matrix, constraint solve).  Its
// model (mbd/assets/pushT.xml) is three world-parented planar bodies with 8 dofs, so a whole physics step — smooth forces,
// closed-form 3x3 inverse mass matrix of the slider, up to 12 soft constraint rows (4 joint limits, 2 sphere-box contacts x a
// 4-sided friction pyramid), 100 projected Gauss-Seidel sweeps, semi-implicit Euler with implicit joint damping — fits in one
// thread's registers / local memory; there is nothing to exchange between threads and no tensor-core shaped work (the "mass
// matrix" is 3x3).  Layout and the restated algorithm: include/mbd_pusht.h; the arithmetic (association order included) is
// the one of oracle/pusht_oracle.c, which the tests compare bit for bit.  Division and square root are the IEEE ones (nvcc
// defaults -prec-div / -prec-sqrt; this kernel is latency-trivial next to the XPBD rollouts).
#pragma once

#include "mbd_pusht.h"

namespace mbd {

struct PushTArgs {
  const float* params; const float* x0; float* Y0s; int n, H;
  float* rewss; float* rews; float* final_state; float* traj;
  int fused; uint32_t k0, k1; int n_total, n_begin; float sigma; const float* Ybar;
  const mbd_step_params* sp; const mbd_step_ctl* ctl; const float* Ybars;   // device-resident step parameters (see RolloutArgs)
  int prng_part;
};

__device__ __forceinline__ void pt_imp_aref(const float* P, float pos, float vel, float& imp, float& aref) {
  const float dmin = P[MBD_PT_DMIN], dmax = P[MBD_PT_DMAX], mid = P[MBD_PT_MID];
  const float x = fabsf(pos) / P[MBD_PT_WIDTH];
  const float a = (1.0f / mid) * (x * x);
  const float omx = 1.0f - x;
  const float b = 1.0f - (1.0f / (1.0f - mid)) * (omx * omx);
  const float y = x < mid ? a : b;
  float d = clampf(dmin + y * (dmax - dmin), dmin, dmax);
  if (x > 1.0f) d = dmax;
  imp = d;
  aref = (0.0f - P[MBD_PT_KB] * vel) - (P[MBD_PT_KK] * d) * pos;
}

// the padded NRP x NRP constraint system and its projected Gauss-Seidel sweeps; for NRP = 4 and 8 every array below is indexed
// by compile-time constants after unrolling, i.e. lives in registers
template <int NRP>
__device__ __forceinline__ void pt_solve(const float* P, const float (*J)[5], const float (*MiJ)[5], const float* Mif, const float* pos,
                                         const int* idx, int nr, const float* qd, int iters, float* xout) {
  float A[NRP][NRP], bq[NRP], invD[NRP], x[NRP];
#pragma unroll
  for (int i = 0; i < NRP; ++i) {
    float Ji[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (i < nr) { const float* Jr = J[idx[i]]; Ji[0] = Jr[0]; Ji[1] = Jr[1]; Ji[2] = Jr[2]; Ji[3] = Jr[3]; Ji[4] = Jr[4]; }
#pragma unroll
    for (int j = 0; j < NRP; ++j) {
      float a = 0.0f;
      if (i < nr && j < nr) a = (((Ji[0] * MiJ[j][0] + Ji[1] * MiJ[j][1]) + Ji[2] * MiJ[j][2]) + Ji[3] * MiJ[j][3]) + Ji[4] * MiJ[j][4];
      A[i][j] = a;
    }
    x[i] = 0.0f;
    if (i < nr) {
      const float vel = (((Ji[0] * qd[0] + Ji[1] * qd[1]) + Ji[2] * qd[2]) + Ji[3] * qd[3]) + Ji[4] * qd[4];
      float imp, aref;
      pt_imp_aref(P, pos[idx[i]], vel, imp, aref);
      const float arr = A[i][i];
      A[i][i] = arr + ((1.0f - imp) / imp) * arr;
      invD[i] = 1.0f / A[i][i];
      bq[i] = ((((Ji[0] * Mif[0] + Ji[1] * Mif[1]) + Ji[2] * Mif[2]) + Ji[3] * Mif[3]) + Ji[4] * Mif[4]) - aref;
    } else {
      A[i][i] = 1.0f; invD[i] = 1.0f; bq[i] = 0.0f;
    }
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NRP; ++i) {
      float r0 = bq[i], r1 = 0.0f;   // two interleaved fused accumulators (even / odd columns): half the dependent chain
#pragma unroll
      for (int j = 0; j < NRP; j += 2) { r0 = fmaf(A[i][j], x[j], r0); r1 = fmaf(A[i][j + 1], x[j + 1], r1); }
      const float res = r0 + r1;
      const float xn = x[i] - res * invD[i];
      x[i] = xn > 0.0f ? xn : 0.0f;
    }
  }
#pragma unroll
  for (int i = 0; i < NRP; ++i) xout[i] = x[i];
}

// one brax.generalized.pipeline.step of the planar model; P = parameter table in shared memory
__device__ void pusht_substep(const float* P, float* q, float* qd, float u0, float u1) {
  const float dt = P[MBD_PT_DT];
  const float ms = P[MBD_PT_MS], ims = P[MBD_PT_IMS], Is = P[MBD_PT_IS], iIs = P[MBD_PT_IIS];
  float s, c;
  mbd_sincosf(q[4], &s, &c);
  const float rx = c * P[MBD_PT_CX] - s * P[MBD_PT_CY];
  const float ry = s * P[MBD_PT_CX] + c * P[MBD_PT_CY];
  // qf_smooth
  float f[5];
  const float w = qd[4], mw2 = ms * (w * w);
  f[0] = P[MBD_PT_GEAR0] * u0 - P[MBD_PT_DPX] * qd[0];
  f[1] = P[MBD_PT_GEAR1] * u1 - P[MBD_PT_DPY] * qd[1];
  f[2] = mw2 * rx - P[MBD_PT_DSX] * qd[2];
  f[3] = mw2 * ry - P[MBD_PT_DSY] * qd[3];
  f[4] = 0.0f - P[MBD_PT_DSTH] * w;
  // inverse mass matrix of the slider block
  const float A00 = ims + (ry * ry) * iIs, A01 = 0.0f - (rx * ry) * iIs, A02 = ry * iIs;
  const float A11 = ims + (rx * rx) * iIs, A12 = 0.0f - rx * iIs, A22 = iIs;
  // constraint rows
  float J[MBD_PT_NROW][5], pos[MBD_PT_NROW];
  bool active[MBD_PT_NROW];
  for (int r = 0; r < MBD_PT_NROW; ++r)
    for (int k = 0; k < 5; ++k) J[r][k] = 0.0f;
  for (int k = 0; k < MBD_PT_NLIM; ++k) {
    const float pmin = q[k] - P[MBD_PT_LIM0 + 2 * k], pmax = P[MBD_PT_LIM0 + 2 * k + 1] - q[k];
    const float pm = pmin < pmax ? pmin : pmax;
    pos[k] = pm < 0.0f ? pm : 0.0f;
    active[k] = pm < 0.0f;
    J[k][k] = pmin < pmax ? 1.0f : -1.0f;
  }
  const float mu = P[MBD_PT_MU], rp = P[MBD_PT_RP];
  for (int b = 0; b < MBD_PT_NBOX; ++b) {
    const float* B = P + MBD_PT_BOX0 + 4 * b;
    const float bx = q[2] + (c * B[0] - s * B[1]), by = q[3] + (s * B[0] + c * B[1]);
    const float dx = q[0] - bx, dy = q[1] - by;
    const float lx = c * dx + s * dy, ly = c * dy - s * dx;
    const float clx = clampf(lx, -B[2], B[2]), cly = clampf(ly, -B[3], B[3]);
    const float ex = lx - clx, ey = ly - cly;
    const float d2 = ex * ex + ey * ey;
    float nlx, nly, dist, sx = clx, sy = cly;
    if (d2 > 0.0f) {
      const float d = sqrtf(d2);
      nlx = ex / d; nly = ey / d;
      dist = d - rp;
    } else {
      const float px = B[2] - fabsf(lx), py = B[3] - fabsf(ly);
      if (px < py) { nlx = lx < 0.0f ? -1.0f : 1.0f; nly = 0.0f; dist = (0.0f - px) - rp; sx = nlx * B[2]; }
      else { nlx = 0.0f; nly = ly < 0.0f ? -1.0f : 1.0f; dist = (0.0f - py) - rp; sy = nly * B[3]; }
    }
    const float nx = c * nlx - s * nly, ny = s * nlx + c * nly;
    const float half = 0.5f * dist;
    const float ax = B[0] + (sx + nlx * half), ay = B[1] + (sy + nly * half);
    const float rhox = c * ax - s * ay, rhoy = s * ax + c * ay;
    const float tx = 0.0f - ny, ty = nx;
    for (int j = 0; j < 4; ++j) {
      const int r = MBD_PT_NLIM + 4 * b + j;
      float ddx = nx, ddy = ny;
      if (j == 0) { ddx = nx - mu * tx; ddy = ny - mu * ty; }
      if (j == 1) { ddx = nx + mu * tx; ddy = ny + mu * ty; }
      J[r][0] = ddx; J[r][1] = ddy;
      J[r][2] = 0.0f - ddx; J[r][3] = 0.0f - ddy;
      J[r][4] = 0.0f - (rhox * ddy - rhoy * ddx);
      pos[r] = dist;
      active[r] = dist < 0.0f;
    }
  }
  // constraint QP + projected Gauss-Seidel: active rows compacted to the front, system padded with identity rows to 4 / 8 / 12
  // rows (oracle/pusht_oracle.c does the same sums in the same order); the 4- and 8-row systems live in registers
  float MiJ[MBD_PT_NROW][5];
  float Mif[5];
  const float imp_ = P[MBD_PT_IMP];
  Mif[0] = imp_ * f[0]; Mif[1] = imp_ * f[1];
  Mif[2] = (A00 * f[2] + A01 * f[3]) + A02 * f[4];
  Mif[3] = (A01 * f[2] + A11 * f[3]) + A12 * f[4];
  Mif[4] = (A02 * f[2] + A12 * f[3]) + A22 * f[4];
  int idx[MBD_PT_NROW], nr = 0;
  for (int r = 0; r < MBD_PT_NROW; ++r)
    if (active[r]) idx[nr++] = r;
  float ftot[5] = {f[0], f[1], f[2], f[3], f[4]};
  if (nr > 0) {
    for (int i = 0; i < nr; ++i) {
      const float* Jr = J[idx[i]];
      MiJ[i][0] = imp_ * Jr[0]; MiJ[i][1] = imp_ * Jr[1];
      MiJ[i][2] = (A00 * Jr[2] + A01 * Jr[3]) + A02 * Jr[4];
      MiJ[i][3] = (A01 * Jr[2] + A11 * Jr[3]) + A12 * Jr[4];
      MiJ[i][4] = (A02 * Jr[2] + A12 * Jr[3]) + A22 * Jr[4];
    }
    const int iters = (int)P[MBD_PT_ITERS];
    float x[MBD_PT_NROW];
    if (nr <= 4) pt_solve<4>(P, J, MiJ, Mif, pos, idx, nr, qd, iters, x);
    else if (nr <= 8) pt_solve<8>(P, J, MiJ, Mif, pos, idx, nr, qd, iters, x);
    else pt_solve<12>(P, J, MiJ, Mif, pos, idx, nr, qd, iters, x);
    for (int i = 0; i < nr; ++i)
      for (int k = 0; k < 5; ++k) ftot[k] = ftot[k] + J[idx[i]][k] * x[i];
  }
  // (M + dt D) qdd = ftot
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

// pushT.py:50-62
__device__ __forceinline__ float pusht_reward(const float* q) {
  const float gx = q[5] - q[2], gy = q[6] - q[3];
  const float px = q[0] - q[2], py = q[1] - q[3];
  const float dps = sqrtf(px * px + py * py) - 0.2f;
  const float d_pusher2slider = dps > 0.0f ? dps : 0.0f;
  return 1.0f - ((sqrtf(gx * gx + gy * gy) + fabsf(q[7] - q[4]) / MBD_PI_F) + d_pusher2slider);
}

// sample_elem: the planner's per-element sampler (defined in mbd_b200.cu before this header is included)
__global__ void __launch_bounds__(64) k_pusht(PushTArgs a) {
  __shared__ float P[MBD_PT_NPARAM];
  for (int k = threadIdx.x; k < MBD_PT_NPARAM; k += blockDim.x) P[k] = a.params[k];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int HNu = a.H * MBD_PT_NU;
  const int nsub = (int)P[MBD_PT_NSUB];
  const uint32_t total = a.prng_part ? 0u : (uint32_t)a.n_total * (uint32_t)HNu;
  uint32_t ck0 = a.k0, ck1 = a.k1; float csigma = a.sigma; const float* cYbar = a.Ybar;
  if (a.sp != nullptr) {
    const int si = a.ctl->i;
    ck0 = a.sp[si].key[0]; ck1 = a.sp[si].key[1]; csigma = a.sp[si].sigma; cYbar = a.Ybars + (size_t)si * HNu;
  }
  float q[MBD_PT_NQ], qd[MBD_PT_NQ];
  for (int k = 0; k < MBD_PT_NQ; ++k) { q[k] = a.x0[k]; qd[k] = a.x0[MBD_PT_NQ + k]; }
  float sum = 0.0f;
  for (int t = 0; t < a.H; ++t) {
    float* ur = a.Y0s + ((size_t)i * a.H + t) * 2;
    float u0, u1;
    if (a.fused) {
      const uint32_t idx = (uint32_t)(a.n_begin + i) * (uint32_t)HNu + (uint32_t)(2 * t);
      u0 = sample_elem(ck0, ck1, idx, total, csigma, cYbar[2 * t]);
      u1 = sample_elem(ck0, ck1, idx + 1, total, csigma, cYbar[2 * t + 1]);
      ur[0] = u0; ur[1] = u1;
    } else {
      u0 = ur[0]; u1 = ur[1];
    }
    u0 = clampf(u0, -1.0f, 1.0f); u1 = clampf(u1, -1.0f, 1.0f);   // motor ctrlrange
    for (int k = 0; k < nsub; ++k) pusht_substep(P, q, qd, u0, u1);
    const float r = pusht_reward(q);
    if (a.rewss) a.rewss[(size_t)i * a.H + t] = r;
    sum += r;
    if (a.traj) {
      float* o = a.traj + ((size_t)i * a.H + t) * MBD_PT_STATE;
      for (int k = 0; k < MBD_PT_NQ; ++k) { o[k] = q[k]; o[MBD_PT_NQ + k] = qd[k]; }
    }
  }
  a.rews[i] = sum / (float)a.H;
  if (a.final_state)
    for (int k = 0; k < MBD_PT_NQ; ++k) { a.final_state[(size_t)i * MBD_PT_STATE + k] = q[k]; a.final_state[(size_t)i * MBD_PT_STATE + MBD_PT_NQ + k] = qd[k]; }
}

}  // namespace mbd
