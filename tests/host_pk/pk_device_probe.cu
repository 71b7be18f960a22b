// Compile-only probe (tests/test_pk_host.py::test_no_packed_contraction): one instantiation of the four packed phases,
// so that the packed instruction counts of its PTX (what the source asks for) and of its SASS (what ptxas made of it)
// can be compared.  ptxas contracts a packed multiply feeding a packed add into FFMA2 regardless of .rn / -fmad=false;
// a drop of FMUL2 / FADD2 against mul / add+sub.f32x2 means a product slipped into a fusable add (see pk_scalar.cuh).
#include "xpbd_pk.cuh"
using namespace mbd::pk;
template <int CMAX>
__device__ __forceinline__ void one_step(const f2* tab, const float* blob, f2* X, f2* E, f2* state, const f2* tau, int l) {
  Model<f2> M; M.t = tab; M.f = blob;
  Smem<f2> S; S.X = X; S.E = E; S.lane = threadIdx.x & 31;
  Cfg c; load_cfg(M, l, c);
  State<f2> s;
  f2* st = state + threadIdx.x * 13;
  s.p = mkV(st[0], st[1], st[2]); s.q = mkQ(st[3], st[4], st[5], st[6]); s.w = mkV(st[7], st[8], st[9]); s.v = mkV(st[10], st[11], st[12]);
  Carry<f2, CMAX> k;
  f2 t3[MBD_MAXDOF] = {tau[0], tau[1], tau[2]};
  phase_A<f2, CMAX>(M, c, S, s, t3, k); __syncthreads();
  phase_B<f2, CMAX>(M, c, S, s, k); __syncthreads();
  phase_C<f2, CMAX>(M, c, S, s, k); __syncthreads();
  phase_D<f2, CMAX>(M, c, S, s, k);
  st[0] = s.p.x; st[1] = s.p.y; st[2] = s.p.z; st[3] = s.q.w; st[4] = s.q.x; st[5] = s.q.y; st[6] = s.q.z;
  st[7] = s.w.x; st[8] = s.w.y; st[9] = s.w.z; st[10] = s.v.x; st[11] = s.v.y; st[12] = s.v.z;
}
extern "C" __global__ void pk_probe(const f2* tab, const float* blob, f2* X, f2* E, f2* state, const f2* tau, int l) {
  one_step<MBD_MAXCON>(tab, blob, X, E, state, tau, l);
}
