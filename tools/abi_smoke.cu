// abi_smoke.cu - plain C++/CUDA-runtime caller of the C ABI (no torch): exercises every entry point on
// small random problems, incl. shapes whose spans are not 16-byte aligned and tail CTAs.  Meant to
// run under `compute-sanitizer --tool memcheck` (fast: no Python start-up).
//   nvcc -O2 -o abi_smoke tools/abi_smoke.cu -Iinclude -Lmpc/pytorch_b200 -lmpcb200 -Xlinker -rpath=$PWD/mpc/pytorch_b200
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mpcb200.h"

template <typename R>
struct Dev {
  R* p = nullptr;
  size_t n = 0;
  explicit Dev(size_t n_) : n(n_) { cudaMalloc(&p, (n ? n : 1) * sizeof(R)); }
  ~Dev() { cudaFree(p); }
  void up(const std::vector<R>& h) { cudaMemcpy(p, h.data(), n * sizeof(R), cudaMemcpyHostToDevice); }
  std::vector<R> down() const {
    std::vector<R> h(n);
    cudaMemcpy(h.data(), p, n * sizeof(R), cudaMemcpyDeviceToHost);
    return h;
  }
};
static float rnd() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int run_case(int B, int T, int n, int m, int bounds_kind, int with_mask) {
  const int p = n + m;
  std::vector<float> C((size_t)T * B * p * p), c((size_t)T * B * p), F((size_t)(T - 1) * B * n * p),
      f((size_t)(T - 1) * B * n), x0((size_t)B * n), cx((size_t)T * B * n, 0.f), cu((size_t)T * B * m, 0.f),
      lo((size_t)T * B * m, -0.25f), hi((size_t)T * B * m, 0.25f);
  std::vector<unsigned char> mask((size_t)T * B * m, 0);
  for (size_t tb = 0; tb < (size_t)T * B; ++tb) {          // C = L L' / p + I
    std::vector<float> L((size_t)p * p);
    for (auto& v : L) v = rnd();
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) {
        float s = i == j ? 1.f : 0.f;
        for (int k = 0; k < p; ++k) s += L[i * p + k] * L[j * p + k] / p;
        C[(tb * p + i) * p + j] = s;
      }
  }
  for (auto& v : c) v = rnd();
  for (size_t i = 0; i < F.size(); ++i) F[i] = 0.3f * rnd();
  for (size_t tb = 0; tb < (size_t)(T - 1) * B; ++tb)
    for (int i = 0; i < n; ++i) F[(tb * n + i) * p + i] += 0.9f;
  for (auto& v : f) v = 0.1f * rnd();
  for (auto& v : x0) v = rnd();
  for (auto& v : mask) v = (rand() % 4) == 0;
  Dev<float> dC(C.size()), dc(c.size()), dF(F.size()), df(f.size()), dx0(x0.size()), dcx(cx.size()), dcu(cu.size()),
      dlo(lo.size()), dhi(hi.size()), nx(cx.size()), nu(cu.size()), costs(B), fdn(B), al(B), du1(cu.size()),
      Ks((size_t)T * B * m * n), ks((size_t)T * B * m);
  Dev<unsigned char> dmask(mask.size()), fmask(mask.size());
  Dev<int> qp((size_t)T * B), st(B);
  dC.up(C); dc.up(c); dF.up(F); df.up(f); dx0.up(x0); dcu.up(cu); dlo.up(lo); dhi.up(hi);
  cudaMemcpy(dmask.p, mask.data(), mask.size(), cudaMemcpyHostToDevice);
  mpcb200_dims d = {B, T, n, m, T - 1, 1, bounds_kind, with_mask, 0, 10, 20, 1};
  mpcb200_params prm = {-0.25, 0.25, 0.0, 0.2};
  int rc = mpcb200_rollout_f32(&d, dF.p, df.p, dx0.p, dcu.p, dcx.p, nullptr);
  if (rc) return printf("rollout rc=%d\n", rc), 1;
  rc = mpcb200_lqr_step_f32(&d, &prm, dC.p, dc.p, dF.p, df.p, dx0.p, dcx.p, dcu.p,
                            bounds_kind == 2 ? dlo.p : nullptr, bounds_kind == 2 ? dhi.p : nullptr,
                            with_mask ? dmask.p : nullptr, nx.p, nu.p, costs.p, fdn.p, al.p, du1.p, qp.p,
                            fmask.p, st.p, Ks.p, ks.p, nullptr);
  if (rc) return printf("step rc=%d (%s)\n", rc, mpcb200_strerror(rc)), 1;
  // adjoint: masked step on (C, -r) from zeros, then the gradient assembly (both paths)
  Dev<float> r(c.size()), zx(cx.size()), zu(cu.size()), z0(x0.size()), ax(cx.size()), au(cu.size()), rx(cx.size()),
      gx0(x0.size()), gC(C.size()), gc(c.size()), gF(F.size()), gf(f.size()), ws((size_t)2 * T * B * n);
  std::vector<float> rr(c.size());
  for (auto& v : rr) v = rnd();
  r.up(rr);
  cudaMemset(zx.p, 0, cx.size() * 4); cudaMemset(zu.p, 0, cu.size() * 4); cudaMemset(z0.p, 0, x0.size() * 4);
  cudaMemset(rx.p, 0, cx.size() * 4);
  mpcb200_dims da = d;
  da.has_f = 0; da.bounds_kind = 0; da.has_zero_mask = 1;
  rc = mpcb200_lqr_step_f32(&da, &prm, dC.p, r.p, dF.p, nullptr, z0.p, zx.p, zu.p, nullptr, nullptr, dmask.p, ax.p,
                            au.p, costs.p, fdn.p, al.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc) return printf("adjoint step rc=%d\n", rc), 1;
  for (int pass = 0; pass < 2; ++pass) {
    rc = mpcb200_lqr_grad_f32(&d, dC.p, dc.p, dF.p, nx.p, nu.p, ax.p, au.p, rx.p, gx0.p, gC.p, gc.p, gF.p, gf.p,
                              pass ? ws.p : nullptr, nullptr);
    if (rc) return printf("grad rc=%d\n", rc), 1;
  }
  {  // the whole KKT adjoint in one call (prep + nested masked step + costates + outer products)
    const size_t wsb = mpcb200_adjoint_workspace_bytes(&d, 4);
    Dev<unsigned char> aws(wsb);
    Dev<float> wu(cu.size());
    std::vector<float> hw(cu.size());
    for (auto& v : hw) v = rnd();
    wu.up(hw);
    rc = mpcb200_lqr_adjoint_f32(&d, &prm, dC.p, dc.p, dF.p, nx.p, nu.p, r.p /* [T,B,p] >= [T,B,n] */, wu.p,
                                 bounds_kind == 2 ? dlo.p : nullptr, bounds_kind == 2 ? dhi.p : nullptr, gx0.p, gC.p,
                                 gc.p, gF.p, gf.p, aws.p, wsb, nullptr);
    if (rc && rc != MPCB200_ERR_SMEM) return printf("one-call adjoint rc=%d (%s)\n", rc, mpcb200_strerror(rc)), 1;
    // time-invariant dynamics / cost through the stride fields (one slice read for every t)
    mpcb200_dims ds = d;
    ds.F_tstride = MPCB200_TIME_INVARIANT;
    ds.C_tstride = MPCB200_TIME_INVARIANT;
    rc = mpcb200_lqr_step_f32(&ds, &prm, dC.p, dc.p, dF.p, df.p, dx0.p, dcx.p, dcu.p,
                              bounds_kind == 2 ? dlo.p : nullptr, bounds_kind == 2 ? dhi.p : nullptr,
                              with_mask ? dmask.p : nullptr, nx.p, nu.p, costs.p, fdn.p, al.p, nullptr, nullptr,
                              nullptr, st.p, Ks.p, ks.p, nullptr);
    if (rc) return printf("strided step rc=%d\n", rc), 1;
  }
  if (m <= 8) {   // standalone pnqp on the last step's C_uu blocks
    std::vector<float> H((size_t)B * m * m), q((size_t)B * m), l2((size_t)B * m, -0.25f), h2((size_t)B * m, 0.25f);
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < m; ++i) {
        q[b * m + i] = rnd();
        for (int j = 0; j < m; ++j) H[(b * m + i) * m + j] = C[(((size_t)(T - 1) * B + b) * p + n + i) * p + n + j];
      }
    Dev<float> dH(H.size()), dq(q.size()), dl(l2.size()), dh(h2.size()), ox(q.size()), oH(H.size());
    Dev<unsigned char> oI(q.size());
    Dev<int> oit(B), ost(B);
    dH.up(H); dq.up(q); dl.up(l2); dh.up(h2);
    rc = mpcb200_pnqp_f32(B, m, dH.p, dq.p, dl.p, dh.p, nullptr, 20, ox.p, oH.p, oI.p, oit.p, ost.p, nullptr);
    if (rc) return printf("pnqp rc=%d\n", rc), 1;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())), 1;
  double s = 0;
  int bad = 0;
  for (float v : nu.down()) { s += v; bad += !std::isfinite(v); }
  for (float v : gC.down()) { s += v; bad += !std::isfinite(v); }
  printf("B=%d T=%d n=%d m=%d bounds=%d mask=%d: checksum %.6f nonfinite %d\n", B, T, n, m, bounds_kind, with_mask, s, bad);
  return bad != 0;
}

// known systems: rollout, exact Jacobians, and the step kernel with the system inside its line search
static int run_dyn(int kind, int B, int T) {
  const int n = kind == MPCB200_DYN_CARTPOLE ? 5 : 3, m = 1, p = n + m;
  const double prm_c[8] = {9.8, 1.0, 0.1, 0.5, 100.0, 0.05, 0, 0}, prm_p[8] = {10.0, 1.0, 1.0, 0.0, 2.0, 0.05, 0, 0};
  const double* dyn = kind == MPCB200_DYN_CARTPOLE ? prm_c : prm_p;
  std::vector<float> x0((size_t)B * n), u((size_t)T * B * m), C((size_t)T * B * p * p, 0.f), c((size_t)T * B * p);
  for (int b = 0; b < B; ++b) {
    const float th = 3.f * rnd();
    float* s = &x0[(size_t)b * n];
    if (n == 5) { s[0] = rnd(); s[1] = rnd(); s[2] = cosf(th); s[3] = sinf(th); s[4] = rnd(); }
    else { s[0] = cosf(th); s[1] = sinf(th); s[2] = rnd(); }
  }
  for (auto& v : u) v = 0.5f * rnd();
  for (auto& v : c) v = 0.1f * rnd();
  for (size_t tb = 0; tb < (size_t)T * B; ++tb)
    for (int i = 0; i < p; ++i) C[(tb * p + i) * p + i] = 1.f;
  Dev<float> dx0(x0.size()), du(u.size()), dx((size_t)T * B * n), dF((size_t)(T - 1) * B * n * p), df((size_t)(T - 1) * B * n),
      dC(C.size()), dc(c.size()), nx((size_t)T * B * n), nu(u.size()), costs(B), fdn(B), al(B);
  dx0.up(x0); du.up(u); dC.up(C); dc.up(c);
  int rc = mpcb200_dyn_rollout_f32(kind, dyn, B, T, dx0.p, du.p, dx.p, nullptr);
  if (rc) return printf("dyn rollout rc=%d\n", rc), 1;
  rc = mpcb200_dyn_linearize_f32(kind, dyn, B, T, dx.p, du.p, dF.p, df.p, nullptr);
  if (rc) return printf("dyn linearize rc=%d\n", rc), 1;
  mpcb200_dims d = {B, T, n, m, T - 1, 1, 1, 0, 0, 3, 20, 1, kind};
  mpcb200_params prm = {-2.0, 2.0, 0.0, 0.5, {0}};
  for (int i = 0; i < 8; ++i) prm.dyn[i] = dyn[i];
  rc = mpcb200_lqr_step_f32(&d, &prm, dC.p, dc.p, dF.p, df.p, dx0.p, dx.p, du.p, nullptr, nullptr, nullptr, nx.p, nu.p,
                            costs.p, fdn.p, al.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc) return printf("dyn step rc=%d (%s)\n", rc, mpcb200_strerror(rc)), 1;
  if (cudaDeviceSynchronize() != cudaSuccess) return printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())), 1;
  double s = 0;
  int bad = 0;
  for (float v : nx.down()) { s += v; bad += !std::isfinite(v); }
  for (float v : dF.down()) { s += v; bad += !std::isfinite(v); }
  printf("known system %d B=%d T=%d: checksum %.6f nonfinite %d\n", kind, B, T, s, bad);
  return bad != 0;
}

int main() {
  int fails = 0;
  fails += run_dyn(MPCB200_DYN_CARTPOLE, 37, 9);
  fails += run_dyn(MPCB200_DYN_PENDULUM, 20, 7);
  const int cases[][6] = {{13, 6, 8, 2, 0, 0}, {13, 6, 8, 2, 1, 0}, {12, 5, 8, 2, 2, 1}, {7, 4, 3, 1, 1, 0},
                          {5, 4, 16, 4, 2, 0}, {1, 3, 2, 2, 0, 0}, {33, 7, 5, 1, 1, 1}, {9, 3, 3, 4, 2, 0},
                          {64, 40, 8, 2, 1, 0}, {16, 6, 4, 2, 1, 0}, {12, 5, 16, 4, 0, 0}, {24, 6, 8, 2, 2, 1},
                          {20, 9, 8, 4, 1, 0}};
  for (auto& cs : cases) fails += run_case(cs[0], cs[1], cs[2], cs[3], cs[4], cs[5]);
  printf("launches: %llu, failures: %d\n", (unsigned long long)mpcb200_launch_count(), fails);
  return fails;
}
