// mbd_oracle_simd.cpp — the TIMED CPU ARM of bench.py (test / benchmark infrastructure, NOT product code).
//
// The scalar C oracle (mbd_oracle.c) runs one rollout per OpenMP iteration; the reference's own CPU path (XLA-CPU under
// jax.vmap) would vectorise ACROSS samples.  This file gives the CPU arm the same advantage: the templated physics the packed
// GPU kernel instantiates with T = f2 (mbd_b200/csrc/xpbd_pk.cuh, written against the scalar layer of pk_scalar.cuh) is
// instantiated here with T = vN — 16 samples per SIMD lane group (one AVX-512 register, or two AVX2 registers), every
// operation the IEEE round-to-nearest operation per lane — and the lane groups are spread over the host threads with OpenMP.
// Same expressions, same association, same fmaf placement: the results are the scalar oracle's bit for bit
// (tests/test_pk_host.py::test_simd_cpu_arm_matches_the_scalar_oracle).  Humanoid models (11 links, hinges only) as in the
// packed kernel; bench.py falls back to the scalar oracle for anything else or when this file does not compile.
//
//   g++ -O3 -march=native -std=c++17 -ffp-contract=off -fno-math-errno -fopenmp -shared -fPIC -Iinclude -Imbd_b200/csrc \
//       oracle/mbd_oracle_simd.cpp -o oracle/libmbd_oracle_simd.so
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "pk_scalar.cuh"

namespace mbd {
namespace pk {
constexpr int kW = 16;
typedef float vf_t __attribute__((vector_size(4 * kW)));
typedef int vi_t __attribute__((vector_size(4 * kW)));
struct vN { vf_t v; };
struct mN { vi_t m; };
static inline vN mkv(vf_t v) { vN r; r.v = v; return r; }
static inline vN mul(vN a, vN b) { return mkv(a.v * b.v); }
static inline vN add(vN a, vN b) { return mkv(a.v + b.v); }
static inline vN sub(vN a, vN b) { return mkv(a.v - b.v); }
static inline vN add_nf(vN a, vN b) { return add(a, b); }
static inline vN sub_nf(vN a, vN b) { return sub(a, b); }
static inline vN fma(vN a, vN b, vN c) {
  vN r;
  for (int i = 0; i < kW; ++i) r.v[i] = __builtin_fmaf(a.v[i], b.v[i], c.v[i]);   // vectorised to vfmadd (explicit fmaf, contraction stays off)
  return r;
}
static inline vN neg(vN a) { return mkv(-a.v); }
static inline vN abs_(vN a) { vi_t b = (vi_t)a.v & 0x7fffffff; return mkv((vf_t)b); }
static inline mN lt(vN x, vN y) { mN m; m.m = x.v < y.v; return m; }
static inline mN le(vN x, vN y) { mN m; m.m = x.v <= y.v; return m; }
static inline mN gt(vN x, vN y) { mN m; m.m = x.v > y.v; return m; }
static inline mN ge(vN x, vN y) { mN m; m.m = x.v >= y.v; return m; }
static inline mN eq(vN x, vN y) { mN m; m.m = x.v == y.v; return m; }
static inline mN mand(mN p, mN q) { mN m; m.m = p.m & q.m; return m; }
static inline vN sel(mN m, vN x, vN y) { return mkv(m.m ? x.v : y.v); }
static inline vN div_(vN a, vN b) { return mkv(a.v / b.v); }
static inline vN rcp_(vN x) { vf_t one = {}; one = one + 1.0f; return mkv(one / x.v); }
static inline vN sqrt_(vN x) {
  vN r;
  for (int i = 0; i < kW; ++i) r.v[i] = __builtin_sqrtf(x.v[i]);
  return r;
}
template <> struct Bc<vN> { static vN of(float c) { vf_t z = {}; return mkv(z + c); } };
}  // namespace pk
}  // namespace mbd

#include "xpbd_pk.cuh"

using namespace mbd::pk;

static inline float clampf_h(float v, float lo_, float hi_) { return v < lo_ ? lo_ : (v > hi_ ? hi_ : v); }
static inline float reward_post_h(int kind, float x, float y, float z) {
  if (kind == MBD_REWARD_HUMANOIDRUN) { float dz = clampf_h(fabsf(z - 1.3f), -1.0f, 1.0f); return (x - dz) - fabsf(y) * 0.1f; }
  return ((1.5f - clampf_h(fabsf(z - 1.3f), -2.0f, 1.0f)) - fabsf(x) * 0.1f) - fabsf(y) * 0.1f;   // humanoidstandup
}

// vmap(rollout_us)(state_init, Y0s) for an 11-link hinge-only model with a post-step reward (humanoidrun / humanoidstandup):
// rews [n], optionally final_state [n, L, 13].  Returns -2 for a model this instantiation does not cover.
extern "C" __attribute__((visibility("default")))
int orc_simd_rollout(const uint32_t* blob, const float* state_init, const float* Y0s, int n, int H, float* rews, float* final_state, int nthreads) {
  const float* bf = reinterpret_cast<const float*>(blob);
  const int32_t* bi = reinterpret_cast<const int32_t*>(blob);
  if (blob[MBD_H_MAGIC] != MBD_MODEL_MAGIC) return -1;
  const int L = bi[MBD_H_NLINK], nu = bi[MBD_H_NU], nsub = bi[MBD_H_NFRAMES], kind = bi[MBD_H_REWARD];
  if (kind != MBD_REWARD_HUMANOIDRUN && kind != MBD_REWARD_HUMANOIDSTANDUP) return -2;
  for (int l = 0; l < L; ++l)
    if (bi[MBD_HDR_WORDS + MBD_F_NDOF * MBD_MAXL + l] > 0 && bi[MBD_HDR_WORDS + MBD_F_SLIDE * MBD_MAXL + l] != 0) return -2;
  std::vector<vN> table(MBD_BLOB_WORDS);
  for (int i = 0; i < MBD_BLOB_WORDS; ++i) table[i] = bc<vN>(bf[i]);
  Model<vN> M; M.t = table.data(); M.f = bf;
  std::vector<Cfg> cfg(L);
  for (int l = 0; l < L; ++l) load_cfg(M, l, cfg[l]);
  const int ngroups = (n + kW - 1) / kW;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    std::vector<vN> X((size_t)L * kXF * kLanes), E((size_t)L * kEF * kLanes);   // one exchange area per thread (lane slot 0 is used)
    Smem<vN> S; S.X = X.data(); S.E = E.data(); S.lane = 0;
    std::vector<State<vN>> st(L);
    std::vector<Carry<vN, MBD_MAXCON>> car(L);
    std::vector<vN> tau((size_t)L * MBD_MAXDOF);
#pragma omp for schedule(dynamic, 1)
    for (int g = 0; g < ngroups; ++g) {
      int idx[kW];
      for (int i = 0; i < kW; ++i) idx[i] = g * kW + i < n ? g * kW + i : n - 1;
      for (int l = 0; l < L; ++l) {
        const float* s0 = state_init + l * MBD_STATE_STRIDE;
        auto b = [&](int i) { return bc<vN>(s0[i]); };
        st[l].p = mkV(b(0), b(1), b(2)); st[l].q = mkQ(b(3), b(4), b(5), b(6)); st[l].w = mkV(b(7), b(8), b(9)); st[l].v = mkV(b(10), b(11), b(12));
        S.put_p(l, st[l].p); S.put_q(l, st[l].q); S.put_w(l, st[l].w);
      }
      float rsum[kW];
      for (int i = 0; i < kW; ++i) rsum[i] = 0.0f;
      for (int t = 0; t < H; ++t) {
        for (int l = 0; l < L; ++l)
          for (int d = 0; d < MBD_MAXDOF; ++d) {
            const int base = MBD_F_DOF0 + d * MBD_DOF_STRIDE;
            const int ak = d < cfg[l].ndof ? M.li(base + MBD_D_ACT, l) : -1;
            vN tv = bc<vN>(0.0f);
            if (ak >= 0) {
              vN u;
              for (int i = 0; i < kW; ++i) u.v[i] = Y0s[((size_t)idx[i] * H + t) * nu + ak];
              tv = mul(M.l(base + MBD_D_GEAR, l), clamp_(u, M.l(base + MBD_D_CLO, l), M.l(base + MBD_D_CHI, l)));
            }
            tau[(size_t)l * MBD_MAXDOF + d] = tv;
          }
        for (int f = 0; f < nsub; ++f) {
          for (int l = 0; l < L; ++l) phase_A<vN, MBD_MAXCON>(M, cfg[l], S, st[l], &tau[(size_t)l * MBD_MAXDOF], car[l]);
          for (int l = 0; l < L; ++l) phase_B<vN, MBD_MAXCON>(M, cfg[l], S, st[l], car[l]);
          for (int l = 0; l < L; ++l) phase_C<vN, MBD_MAXCON>(M, cfg[l], S, st[l], car[l]);
          for (int l = 0; l < L; ++l) phase_D<vN, MBD_MAXCON>(M, cfg[l], S, st[l], car[l]);
        }
        V<vN> x0 = link_origin_w(M, 0, st[0]);
        for (int i = 0; i < kW; ++i) rsum[i] += reward_post_h(kind, x0.x.v[i], x0.y.v[i], x0.z.v[i]);
      }
      for (int i = 0; i < kW && g * kW + i < n; ++i) {
        rews[g * kW + i] = rsum[i] / (float)H;
        if (final_state)
          for (int l = 0; l < L; ++l) {
            float* o = final_state + ((size_t)(g * kW + i) * L + l) * MBD_STATE_STRIDE;
            const State<vN>& s = st[l];
            const vN f[13] = {s.p.x, s.p.y, s.p.z, s.q.w, s.q.x, s.q.y, s.q.z, s.w.x, s.w.y, s.w.z, s.v.x, s.v.y, s.v.z};
            for (int j = 0; j < 13; ++j) o[j] = f[j].v[i];
          }
      }
    }
  }
  return 0;
}
extern "C" __attribute__((visibility("default"))) int orc_simd_width(void) { return kW; }
