// Host check build of the packed-kernel physics (mbd_b200/csrc/xpbd_pk.cuh) — TEST INFRASTRUCTURE.
// Compiles the SAME templated phase functions the sm_100a kernel k_rollout_pk instantiates, with plain g++
// (T = float: one sample; T = f2: the two-sample type, emulated as {float, float}), and runs them link by link,
// phase by phase.  tests/test_pk_host.py compares the result bit for bit with the CPU oracle, so the translation
// of the physics into the scalar layer is verified without a GPU; on the GPU only the packed instructions
// themselves (IEEE per component) and the data movement remain to be checked.
//   g++ -O2 -std=c++17 -ffp-contract=off -mfma -shared -fPIC -Iinclude -Imbd_b200/csrc tests/host_pk/pk_harness.cpp
#include <stdint.h>
#include <string.h>
#include <vector>

#include "pk_scalar.cuh"

// ---- a third scalar type: one sample, every operation counted (algorithmic flops "counted by instrumenting
// the CPU restatement").  Lives in mbd::pk so that the templated physics finds the overloads by ADL. ------------------
namespace mbd {
namespace pk {
struct OpCount { unsigned long long mul, add, fma, div, rcp, sqrt, cmp, sel, neg; };
static OpCount g_ops;
struct c1 { float v; };
static inline c1 mkc(float v) { c1 r; r.v = v; return r; }
static inline c1 mul(c1 a, c1 b) { ++g_ops.mul; return mkc(a.v * b.v); }
static inline c1 add(c1 a, c1 b) { ++g_ops.add; return mkc(a.v + b.v); }
static inline c1 sub(c1 a, c1 b) { ++g_ops.add; return mkc(a.v - b.v); }
static inline c1 add_nf(c1 a, c1 b) { return add(a, b); }
static inline c1 sub_nf(c1 a, c1 b) { return sub(a, b); }
static inline c1 fma(c1 a, c1 b, c1 c) { ++g_ops.fma; return mkc(fmaf(a.v, b.v, c.v)); }
static inline c1 neg(c1 a) { ++g_ops.neg; return mkc(-a.v); }
static inline c1 abs_(c1 a) { ++g_ops.neg; return mkc(fabsf(a.v)); }
static inline bool lt(c1 x, c1 y) { ++g_ops.cmp; return x.v < y.v; }
static inline bool le(c1 x, c1 y) { ++g_ops.cmp; return x.v <= y.v; }
static inline bool gt(c1 x, c1 y) { ++g_ops.cmp; return x.v > y.v; }
static inline bool ge(c1 x, c1 y) { ++g_ops.cmp; return x.v >= y.v; }
static inline bool eq(c1 x, c1 y) { ++g_ops.cmp; return x.v == y.v; }
static inline c1 sel(bool m, c1 x, c1 y) { ++g_ops.sel; return m ? x : y; }
static inline c1 div_(c1 a, c1 b) { ++g_ops.div; return mkc(a.v / b.v); }
static inline c1 rcp_(c1 x) { ++g_ops.rcp; return mkc(1.0f / x.v); }
static inline c1 sqrt_(c1 x) { ++g_ops.sqrt; return mkc(sqrtf(x.v)); }
template <> struct Bc<c1> { static c1 of(float c) { return mkc(c); } };
}  // namespace pk
}  // namespace mbd

#include "xpbd_pk.cuh"

using namespace mbd::pk;

static float reward_post_host(int kind, float x, float y, float z) {
  auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
  if (kind == MBD_REWARD_HUMANOIDRUN) {
    float dz = clampf(fabsf(z - 1.3f), -1.0f, 1.0f);
    return (x - dz) - fabsf(y) * 0.1f;
  }
  if (kind == MBD_REWARD_HUMANOIDSTANDUP) return ((1.5f - clampf(fabsf(z - 1.3f), -2.0f, 1.0f)) - fabsf(x) * 0.1f) - fabsf(y) * 0.1f;
  return x - clampf(fabsf(z - 1.0f), -1.0f, 1.0f) * 0.5f;
}

template <class T> struct Lanes;
template <> struct Lanes<float> {
  static constexpr int N = 1;
  static float make(const float* v) { return v[0]; }
  static float get(float x, int) { return x; }
};
template <> struct Lanes<c1> {
  static constexpr int N = 1;
  static c1 make(const float* v) { return mkc(v[0]); }
  static float get(c1 x, int) { return x.v; }
};
template <> struct Lanes<f2> {
  static constexpr int N = 2;
  static f2 make(const float* v) { return mk2(v[0], v[1]); }
  static float get(f2 x, int i) { return i == 0 ? lo(x) : hi(x); }
};

template <class T, int CMAX>
static void rollout(const uint32_t* blob, const float* state_init, const float* Y0s, int n, int H, int nsub_override, float* rews,
                    float* final_state) {
  constexpr int NS = Lanes<T>::N;
  const float* bf = reinterpret_cast<const float*>(blob);
  std::vector<T> table(MBD_BLOB_WORDS);
  for (int i = 0; i < MBD_BLOB_WORDS; ++i) { float v[2] = {bf[i], bf[i]}; table[i] = Lanes<T>::make(v); }
  Model<T> M; M.t = table.data(); M.f = bf;
  const int L = M.hi(MBD_H_NLINK), nu = M.hi(MBD_H_NU);
  const int nsub = nsub_override > 0 ? nsub_override : M.hi(MBD_H_NFRAMES);
  const int kind = M.hi(MBD_H_REWARD);
  std::vector<T> X((size_t)L * kXF * kLanes), E((size_t)L * kEF * kLanes);
  Smem<T> S; S.X = X.data(); S.E = E.data(); S.lane = 0;
  std::vector<Cfg> cfg(L);
  for (int l = 0; l < L; ++l) load_cfg(M, l, cfg[l]);
  for (int n0 = 0; n0 < n; n0 += NS) {
    int idx[2] = {n0, n0 + 1 < n ? n0 + 1 : n - 1};
    std::vector<State<T>> st(L);
    std::vector<Carry<T, CMAX>> car(L);
    for (int l = 0; l < L; ++l) {
      const float* s0 = state_init + l * MBD_STATE_STRIDE;
      auto b = [&](int i) { return bc<T>(s0[i]); };
      st[l].p = mkV(b(0), b(1), b(2)); st[l].q = mkQ(b(3), b(4), b(5), b(6)); st[l].w = mkV(b(7), b(8), b(9)); st[l].v = mkV(b(10), b(11), b(12));
      S.put_p(l, st[l].p); S.put_q(l, st[l].q); S.put_w(l, st[l].w);
    }
    float rsum[2] = {0.0f, 0.0f};
    for (int t = 0; t < H; ++t) {
      std::vector<T> tau((size_t)L * MBD_MAXDOF);
      for (int l = 0; l < L; ++l)
        for (int d = 0; d < MBD_MAXDOF; ++d) {
          const int base = MBD_F_DOF0 + d * MBD_DOF_STRIDE;
          int ak = d < cfg[l].ndof ? M.li(base + MBD_D_ACT, l) : -1;
          T tv = bc<T>(0.0f);
          if (ak >= 0) {
            float u[2];
            for (int i = 0; i < 2; ++i) u[i] = Y0s[((size_t)idx[i] * H + t) * nu + ak];
            tv = mul(M.l(base + MBD_D_GEAR, l), clamp_(Lanes<T>::make(u), M.l(base + MBD_D_CLO, l), M.l(base + MBD_D_CHI, l)));
          }
          tau[(size_t)l * MBD_MAXDOF + d] = tv;
        }
      for (int f = 0; f < nsub; ++f) {
        for (int l = 0; l < L; ++l) phase_A<T, CMAX>(M, cfg[l], S, st[l], &tau[(size_t)l * MBD_MAXDOF], car[l]);
        for (int l = 0; l < L; ++l) phase_B<T, CMAX>(M, cfg[l], S, st[l], car[l]);
        for (int l = 0; l < L; ++l) phase_C<T, CMAX>(M, cfg[l], S, st[l], car[l]);
        for (int l = 0; l < L; ++l) phase_D<T, CMAX>(M, cfg[l], S, st[l], car[l]);
      }
      V<T> x0 = link_origin_w(M, 0, st[0]);
      for (int i = 0; i < NS; ++i) rsum[i] += reward_post_host(kind, Lanes<T>::get(x0.x, i), Lanes<T>::get(x0.y, i), Lanes<T>::get(x0.z, i));
    }
    for (int i = 0; i < NS && n0 + i < n; ++i) {
      rews[n0 + i] = rsum[i] / (float)H;
      if (final_state)
        for (int l = 0; l < L; ++l) {
          float* o = final_state + ((size_t)(n0 + i) * L + l) * MBD_STATE_STRIDE;
          const State<T>& s = st[l];
          const T f[13] = {s.p.x, s.p.y, s.p.z, s.q.w, s.q.x, s.q.y, s.q.z, s.w.x, s.w.y, s.w.z, s.v.x, s.v.y, s.v.z};
          for (int j = 0; j < 13; ++j) o[j] = Lanes<T>::get(f[j], i);
        }
    }
  }
}

// Operation counts of ONE rollout (n = 1): ops[9] = mul, add/sub, fma, div, rcp, sqrt, compare, select, neg/abs —
// physics only (the phases; reward and action clipping excluded by resetting around them would be overkill: they are < 0.1 %).
extern "C" int pk_host_count_ops(const uint32_t* blob, const float* state_init, const float* Y0s, int H, int nsub_override,
                                 unsigned long long* ops, float* rews) {
  memset(&g_ops, 0, sizeof(g_ops));
  rollout<c1, MBD_MAXCON>(blob, state_init, Y0s, 1, H, nsub_override, rews, nullptr);
  const unsigned long long v[9] = {g_ops.mul, g_ops.add, g_ops.fma, g_ops.div, g_ops.rcp, g_ops.sqrt, g_ops.cmp, g_ops.sel, g_ops.neg};
  memcpy(ops, v, sizeof(v));
  return 0;
}

extern "C" int pk_host_rollout(const uint32_t* blob, const float* state_init, const float* Y0s, int n, int H, int packed, int nsub_override,
                               float* rews, float* final_state) {
  if (packed) rollout<f2, MBD_MAXCON>(blob, state_init, Y0s, n, H, nsub_override, rews, final_state);
  else rollout<float, MBD_MAXCON>(blob, state_init, Y0s, n, H, nsub_override, rews, final_state);
  return 0;
}
