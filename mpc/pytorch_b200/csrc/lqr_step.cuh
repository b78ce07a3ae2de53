// lqr_step.cuh - one box-constrained LQR step as ONE persistent-per-CTA kernel (sm_100a).
//
// Replaces the body of LQRStepFn.forward (reference mpc/lqr_step.py:277-309):
//   c_back (:289-295), lqr_backward (:52-160) with pnqp (mpc/pnqp.py:5-82) or the
//   u_zero_I masked solve (:100-127), and lqr_forward's rollout + line search (:164-261).
//
// Mapping (designed for B200, not translated from the reference's per-op loop):
//  * P = n+m lanes own one problem; lane j owns COLUMN j of every p-wide matrix of that
//    problem (Q_t, F_t, C_t) and, for j < n, column j of the value matrix V and of K_t.
//    32/P problems share a warp; NW consumer warps (the smallest count whose spans stay 16-byte
//    aligned: small CTAs = fine load-balance granularity, the batch is < 1 wave) + 1 producer warp
//    form a CTA.  The dense products use packed FFMA2 (two fp32 FMAs per lane per instruction).
//  * the producer warp streams the per-time-step tiles C[t],F[t],c[t],f[t],x_bar[t],u_bar[t]
//    (+ tensor bounds) of the CTA's W consecutive problems - contiguous in the reference's
//    time-major layout - into a 3-stage shared-memory ring with 1-D bulk TMA
//    (cp.async.bulk + mbarrier complete_tx); full/empty mbarriers pace it.  Shapes whose
//    spans are not 16-byte aligned take a plain-load path in the same warp.
//  * V (n x n, stored transposed), v, K_t and Q_xu live in a per-problem shared scratch and are re-read as
//    broadcast vector loads; the m x m solve / pnqp runs redundantly on every lane of the
//    problem from shuffled copies of Q_uu, q_u (registers only, no divergence inside a problem).
//  * K_t,k_t for all t stay in shared memory between the backward sweep and the rollout;
//    the rollout re-streams the tiles (L2 hits) and never round-trips gains through HBM
//    (unless T is too long for shared memory, then a caller-provided Ks/ks buffer is used).
//  * line search: per-problem alpha; a CTA repeats the rollout while any of its problems is
//    worse and iterations remain - per problem this is exactly the reference's batch loop.
//  * MODE (template): PLAIN / BOX (pnqp) / MASK (u_zero_I adjoint solve) - no mode branches at run time.
// This is the GENERIC kernel (any n, m <= 32 lanes, unaligned spans): shapes with even n, m and 16-byte
// aligned tensors run the column-pair kernel in lqr_step2.cuh.  The round-1 experiments that lived here as
// compile-time knobs (two columns per lane, V in registers, padded tiles, mma.sync products, phase timers)
// are described with their measurements in DESIGN.md section 7; their code was removed.
#pragma once
#include <atomic>
#include <cstdio>
#include "common.cuh"
#include "dynamics.cuh"

#ifndef MPCB_STAGES
#define MPCB_STAGES 3
#endif

namespace mpcb200 {

struct StepArgs {
  int B, T, F_T;
  int has_f, bounds_kind, has_mask, has_delta, max_ls, pnqp_iters, do_rollout;
  double u_lo, u_hi, delta_u, ls_decay;
  const void *C, *c, *F, *f, *x_init, *cur_x, *cur_u, *u_lower, *u_upper;
  const unsigned char* zero_mask;
  void *new_x, *new_u, *costs, *full_du_norm, *alphas, *du_first;
  int* qp_iters;
  unsigned char* free_mask;
  int* status;
  void *Ks, *ks;
  int bulk_ok;    // host-verified: all tensor bases and per-time-step strides are 16-byte aligned
  int k_in_smem;  // gains of all T steps fit in shared memory
  int impl;       // 0 pick, 1 generic (column per lane), 2 column-pair kernel (lqr_step2.cuh)
  long long C_ts, c_ts, F_ts, f_ts;   // elements between consecutive time slices of C, c, F, f (0 = time invariant)
  // fused KKT adjoint (column-pair kernel only): after the masked solve, a third sweep computes the costates and
  // writes dC, dc, dF, df, dx_init; `c` carries -r, the x_bar/u_bar tile slots carry the forward solution tau*
  int adj, adj_has_df;
  const void *adj_c, *adj_x, *adj_u;
  void *adj_dC, *adj_dc, *adj_dF, *adj_df, *adj_dx_init;
  long long adj_c_ts;
  int dyn_kind;   // true dynamics of the rollout: DYN_LINEAR (F,f) or a known system evaluated in the kernel
  DynParams dp;
};

template <typename R, int N, int M>
struct StepCfg {
  static constexpr int P = N + M;
  static constexpr int EA = 16 / (int)sizeof(R);
  static_assert(P <= 32, "one problem must fit a warp");
  static constexpr int CPL = 1;                     // columns per lane
  static constexpr int LP = (P + CPL - 1) / CPL;   // lanes per problem
  static constexpr int PPW = 32 / LP;               // problems per warp
  // consumer warps per CTA: the smallest count whose per-time-step spans stay 16-byte aligned for every
  // tensor (so the bulk-TMA path applies).  Small CTAs matter: 4096 problems are < 1 wave, and the
  // kernel time is set by the most loaded SM, so the CTA granularity is the load-balance granularity.
#ifndef MPCB_MAXNW
#define MPCB_MAXNW 4
#endif
  static constexpr bool span_ok(int nw) {
    return (nw * PPW * M * (int)sizeof(R)) % 16 == 0 && (nw * PPW * N * (int)sizeof(R)) % 16 == 0;
  }
  static constexpr int NW = (span_ok(1) && MPCB_MAXNW >= 1) ? 1 : (span_ok(2) && MPCB_MAXNW >= 2) ? 2 : 4;
  static constexpr int W = NW * PPW;      // problems per CTA
  static_assert(span_ok(NW), "CTA problem count must keep spans 16-byte aligned");
  static constexpr int THREADS = (NW + 1) * 32;
  static constexpr int S = MPCB_STAGES;   // ring stages
  static constexpr int VS = round_up(N, 4);
  static constexpr int CS = P * P, FS = N * P;     // per-problem strides of the C and F tiles inside a stage
  // stage tile offsets (elements); every sub-tile starts 16-byte aligned (span_ok / padded strides)
  static constexpr int OFF_C = 0;
  static constexpr int OFF_F = OFF_C + W * CS;
  static constexpr int OFF_c = OFF_F + W * FS;
  static constexpr int OFF_f = OFF_c + W * P;
  static constexpr int OFF_x = OFF_f + W * N;
  static constexpr int OFF_u = OFF_x + W * N;
  static constexpr int OFF_lo = OFF_u + W * M;
  static constexpr int OFF_hi = OFF_lo + W * M;
  static constexpr int OFF_END = OFF_hi + W * M;
  static constexpr int STAGE_BYTES = round_up(OFF_END * (int)sizeof(R) + round_up(W * M, 16), 128);
  // per-problem scratch (elements)
  static constexpr int SC_V = 0;                 // N x VS value matrix
  static constexpr int SC_v = SC_V + N * VS;     // VS      value vector
  static constexpr int SC_K = SC_v + VS;         // M x VS (+ M) K_t,k_t exchange when gains are not smem resident
  static constexpr int KT = M * VS + round_up(M, 4);  // elements per (problem, t) of the gain store
  static constexpr int SC_Q = SC_K + KT;         // M x VS  Q_xu exchange (row a = Q[:n, n+a])
  static constexpr int SC_X = SC_Q + M * VS;     // 2 x VS  rollout state exchange
  static constexpr int SC_R = SC_X + 2 * VS;     // cost reduction
  static constexpr int SC_RAW = SC_R + round_up(P, 4);
  static constexpr int SCR = (SC_RAW % 32 == 0 || SC_RAW % 32 == 16) ? SC_RAW + 4 : SC_RAW;
  // KREDUCE: when the gains live in the caller's Ks/ks buffer (long horizons / large n), lane i reads only
  // column i of K_t and the products are butterfly-reduced over the n state lanes (needs one problem per
  // warp and n a power of two).  For such shapes the gain store is moved out of shared memory on purpose
  // (GAIN_SMEM_LIMIT bytes per problem): it is what limits the resident warps per SM.
  static constexpr bool KREDUCE = CPL == 1 && PPW == 1 && (N & (N - 1)) == 0;
  static constexpr int GAIN_SMEM_LIMIT = 6144;
  static bool prefers_workspace(int T, int max_smem_optin) {
    return smem_bytes(T, true) > (size_t)max_smem_optin ||
           (KREDUCE && (size_t)T * KT * sizeof(R) > (size_t)GAIN_SMEM_LIMIT);
  }
  static constexpr int HDR_BYTES = 256;          // 2*S mbarriers + 32 vote words
  static size_t smem_bytes(int T, bool k_in_smem) {
    size_t b = HDR_BYTES + (size_t)S * STAGE_BYTES + (size_t)W * SCR * sizeof(R);
    if (k_in_smem) b += (size_t)W * T * KT * sizeof(R);
    return b;
  }
};

// ---------------------------------------------------------------------------------------------
// pnqp for one problem, executed redundantly by every lane of the problem (registers only).
// Control flow is what the reference takes for n_batch == 1 (mpc/pnqp.py:5-82).
// ---------------------------------------------------------------------------------------------
template <typename R, int M>
MPCB_DEV void pnqp_lane(const R (&H)[M][M], const R (&q)[M], const R (&lo)[M], const R (&hi)[M],
                        bool warm, R (&x)[M], Ldl<R, M>& fac, unsigned& fmask, int& iters,
                        bool& conv, bool& badpiv, int max_iter) {
  const R GAMMA = R(0.1);
  auto obj = [&](const R(&z)[M]) {            // pnqp.py:11-12
    R s = R(0);
#pragma unroll
    for (int a = 0; a < M; ++a) {
      R hz = R(0);
#pragma unroll
      for (int b = 0; b < M; ++b) hz += H[a][b] * z[b];
      s += z[a] * (R(0.5) * hz + q[a]);
    }
    return s;
  };
  badpiv = false;
  if (!warm) {                                 // pnqp.py:14-19
    fac.factor(H);
    badpiv = fac.bad;
    R t[M];
    fac.solve(q, t);
#pragma unroll
    for (int a = 0; a < M; ++a) x[a] = -t[a];
  }
#pragma unroll
  for (int a = 0; a < M; ++a) {                // :23  util.eclamp: lower bound first, then upper
    x[a] = x[a] < lo[a] ? lo[a] : x[a];
    x[a] = x[a] > hi[a] ? hi[a] : x[a];
  }

  for (int i = 0; i < max_iter; ++i) {
    R g[M];
#pragma unroll
    for (int a = 0; a < M; ++a) {              // :29
      R s = q[a];
#pragma unroll
      for (int b = 0; b < M; ++b) s += H[a][b] * x[b];
      g[a] = s;
    }
    fmask = 0u;
#pragma unroll
    for (int a = 0; a < M; ++a) {              // :32 exact equality with the assigned bound
      const bool cl = ((x[a] == lo[a]) && (g[a] > R(0))) || ((x[a] == hi[a]) && (g[a] < R(0)));
      if (!cl) fmask |= (1u << a);
    }
    R A[M][M], gm[M], dx[M];
#pragma unroll
    for (int a = 0; a < M; ++a) {              // :44-48
      const bool fa = (fmask >> a) & 1u;
      gm[a] = fa ? g[a] : R(0);
#pragma unroll
      for (int b = 0; b < M; ++b) {
        const bool fb = (fmask >> b) & 1u;
        A[a][b] = (fa && fb) ? H[a][b] : R(0);
      }
      A[a][a] += R(1e-11);
    }
    fac.factor(A);
    badpiv = badpiv || fac.bad;
    fac.solve(gm, dx);                         // :53-54
    R nrm2 = R(0);
#pragma unroll
    for (int a = 0; a < M; ++a) {
      dx[a] = -dx[a];
      nrm2 += dx[a] * dx[a];
    }
    if (!(sqrt(nrm2) >= R(1e-4))) {            // :56-59
      iters = i;
      conv = true;
      return;
    }
    R alpha = R(1), mx[M];
    const R fx = obj(x);
    int count = 0;
    bool again;
    do {                                       // :65-76 with n_batch == 1
#pragma unroll
      for (int a = 0; a < M; ++a) {
        const R v = x[a] + alpha * dx[a];
        const R vl = v < lo[a] ? lo[a] : v;
        mx[a] = vl > hi[a] ? hi[a] : vl;
      }
      R den = R(0);
#pragma unroll
      for (int a = 0; a < M; ++a) den += g[a] * (x[a] - mx[a]);
      const R arm = (fx - obj(mx)) / den;
      again = arm <= GAMMA;                    // NaN compares false, like torch
      if (again) alpha *= R(0.1);
      ++count;
    } while (again && count < 10);
    // A step that does not move x (bitwise) is a fixed point of the whole iteration: g, the active set,
    // H_, dx and the Armijo trials of every later iteration are identical.  In fp32 this is how the
    // reference fails to converge (|dx| stays just above 1e-4 while x + alpha dx rounds back to x);
    // returning now with the outcome of iteration max_iter-1 is exactly what the remaining iterations
    // would produce, without ~18 x 10 wasted Armijo trials that made this warp the kernel's tail.
    bool moved = false;
#pragma unroll
    for (int a = 0; a < M; ++a) {
      moved = moved || !(mx[a] == x[a]);
      x[a] = mx[a];                            // :78
    }
    if (!moved) break;
  }
  iters = max_iter - 1;                        // :80-82
  conv = false;
}

// ---------------------------------------------------------------------------------------------
// producer warp: stream one (t) tile set of the CTA's problems into ring stage `tile % S`
// ---------------------------------------------------------------------------------------------
template <typename R, int N, int M>
MPCB_DEV void step_producer(const StepArgs& a, unsigned char* stage_base, uint64_t* full,
                            uint64_t* empty, volatile int* votes, int b0, int cnt, int lane) {
  using K = StepCfg<R, N, M>;
  constexpr int P = K::P;
  constexpr uint32_t SZ = sizeof(R);
  const R* gC = (const R*)a.C;
  const R* gc = (const R*)a.c;
  const R* gF = (const R*)a.F;
  const R* gf = (const R*)a.f;
  const R* gx = (const R*)a.cur_x;
  const R* gu = (const R*)a.cur_u;
  const R* glo = (const R*)a.u_lower;
  const R* ghi = (const R*)a.u_upper;
  const bool tail_ok = (cnt == K::W) || (((cnt * M * SZ) % 16 == 0) && ((cnt * N * SZ) % 16 == 0));
  const bool bulk = a.bulk_ok && tail_ok;
  const int T = a.T;
  int s = 0;
  uint32_t ph = 0;

  auto issue = [&](int t, bool fwd) {
    mbar_wait(&empty[s], ph ^ 1u);
    R* st = (R*)(stage_base + (size_t)s * K::STAGE_BYTES);
    const size_t tb = (size_t)t * a.B + b0;
    const size_t tC = (size_t)t * a.C_ts + (size_t)b0 * P * P, tF = (size_t)t * a.F_ts + (size_t)b0 * N * P;
    const size_t tc = (size_t)t * a.c_ts + (size_t)b0 * P, tf_ = (size_t)t * a.f_ts + (size_t)b0 * N;
    const bool needF = t < T - 1;
    const bool needf = fwd && needF && a.has_f;
    if (a.has_mask) {
      unsigned char* mk = (unsigned char*)(st + K::OFF_END);
      for (int i = lane; i < cnt * M; i += 32) mk[i] = a.zero_mask[tb * M + i];
    }
    if (bulk) {
      __syncwarp();
      if (lane == 0) {
        uint32_t bytes = (uint32_t)cnt * (P * P + P + N + M) * SZ;
        if (needF) bytes += (uint32_t)cnt * N * P * SZ;
        if (needf) bytes += (uint32_t)cnt * N * SZ;
        if (a.bounds_kind == 2) bytes += 2u * cnt * M * SZ;
        mbar_arrive_expect_tx(&full[s], bytes);
        if constexpr (K::CS == P * P) {
          bulk_g2s(st + K::OFF_C, gC + tC, (uint32_t)cnt * P * P * SZ, &full[s]);
        } else {
          for (int q = 0; q < cnt; ++q)
            bulk_g2s(st + K::OFF_C + q * K::CS, gC + tC + (size_t)q * P * P, (uint32_t)P * P * SZ, &full[s]);
        }
        if (needF) {
          if constexpr (K::FS == N * P) {
            bulk_g2s(st + K::OFF_F, gF + tF, (uint32_t)cnt * N * P * SZ, &full[s]);
          } else {
            for (int q = 0; q < cnt; ++q)
              bulk_g2s(st + K::OFF_F + q * K::FS, gF + tF + (size_t)q * N * P, (uint32_t)N * P * SZ, &full[s]);
          }
        }
        bulk_g2s(st + K::OFF_c, gc + tc, (uint32_t)cnt * P * SZ, &full[s]);
        if (needf) bulk_g2s(st + K::OFF_f, gf + tf_, (uint32_t)cnt * N * SZ, &full[s]);
        bulk_g2s(st + K::OFF_x, gx + tb * N, (uint32_t)cnt * N * SZ, &full[s]);
        bulk_g2s(st + K::OFF_u, gu + tb * M, (uint32_t)cnt * M * SZ, &full[s]);
        if (a.bounds_kind == 2) {
          bulk_g2s(st + K::OFF_lo, glo + tb * M, (uint32_t)cnt * M * SZ, &full[s]);
          bulk_g2s(st + K::OFF_hi, ghi + tb * M, (uint32_t)cnt * M * SZ, &full[s]);
        }
      }
    } else {
      auto cp = [&](R* dst, const R* src, int nelem) {
        for (int i = lane; i < nelem; i += 32) dst[i] = __ldg(src + i);
      };
      for (int q = 0; q < cnt; ++q) cp(st + K::OFF_C + q * K::CS, gC + tC + (size_t)q * P * P, P * P);
      if (needF)
        for (int q = 0; q < cnt; ++q) cp(st + K::OFF_F + q * K::FS, gF + tF + (size_t)q * N * P, N * P);
      cp(st + K::OFF_c, gc + tc, cnt * P);
      if (needf) cp(st + K::OFF_f, gf + tf_, cnt * N);
      cp(st + K::OFF_x, gx + tb * N, cnt * N);
      cp(st + K::OFF_u, gu + tb * M, cnt * M);
      if (a.bounds_kind == 2) {
        cp(st + K::OFF_lo, glo + tb * M, cnt * M);
        cp(st + K::OFF_hi, ghi + tb * M, cnt * M);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    if (++s == K::S) { s = 0; ph ^= 1u; }
  };

  for (int t = T - 1; t >= 0; --t) issue(t, false);
  if (a.do_rollout) {
    for (int pass = 0;; ++pass) {
      for (int t = 0; t < T; ++t) issue(t, true);
      named_bar_sync(1, K::THREADS);
      const int cont = votes[pass & 31];
      if (!cont || pass + 1 >= a.max_ls) break;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// consumer warps.  MODE (compile time): 0 plain (no bounds, no mask), 1 box (pnqp; optional
// u_zero_I), 2 mask (u_zero_I only - the adjoint solve).
// ---------------------------------------------------------------------------------------------
enum { MODE_PLAIN = 0, MODE_BOX = 1, MODE_MASK = 2 };

template <typename R, int N, int M, int MODE>
MPCB_DEV void step_consumer(const StepArgs& a, unsigned char* stage_base, uint64_t* full,
                            uint64_t* empty, volatile int* votes, R* scratch_all, R* kstore_all,
                            int b0, int warp, int lane) {
  using K = StepCfg<R, N, M>;
  constexpr int P = K::P, LP = K::LP, PPW = K::PPW, CPL = K::CPL, VS = K::VS, EA = K::EA, KT = K::KT;
  constexpr unsigned FULLM = (1u << M) - 1u;
  constexpr bool BOX = MODE == MODE_BOX;
  constexpr int A_N = align_elems<R>(N), A_M = align_elems<R>(M);
  constexpr int A_ROW = (P % 2 == 0) ? 2 : 1;   // row r*P of a per-problem tile (pair aligned for even P)
  constexpr int A_2P = align_elems<R>(2 * P), A_NP = align_elems<R>(K::FS);
  const int T = a.T, B = a.B;
  const bool writer_lane = lane < PPW * LP;
  const int pi = writer_lane ? lane / LP : PPW - 1;
  const int base = pi * LP;
  const int j = writer_lane ? lane - base : LP - 1;
  const int pw = warp * PPW + pi;
  const int b = b0 + pw;
  const bool valid = b < B;
  const bool wr = writer_lane && valid;

  // the CPL columns this lane owns: col = j + s*LP (slot 0 is always a state column)
  int cc[CPL], ua[CPL], fr[CPL];
  bool isx[CPL], wsl[CPL];
#pragma unroll
  for (int sl = 0; sl < CPL; ++sl) {
    const int col = j + sl * LP;
    wsl[sl] = writer_lane && col < P;
    cc[sl] = col < P ? col : P - 1;
    isx[sl] = cc[sl] < N;
    ua[sl] = isx[sl] ? 0 : cc[sl] - N;       // control index of a u column
    fr[sl] = isx[sl] ? cc[sl] : N - 1;       // a valid row of F / V for every slot
  }

  // per-problem element offsets inside a stage (loop invariant)
  const int oC = K::OFF_C + pw * K::CS, oF = K::OFF_F + pw * K::FS;
  const int oc = K::OFF_c + pw * P, of_ = K::OFF_f + pw * N, ox = K::OFF_x + pw * N, ou = K::OFF_u + pw * M;
  const int olo = K::OFF_lo + pw * M, ohi = K::OFF_hi + pw * M;

  R* scr = scratch_all + (size_t)pw * K::SCR;
  R* Vs = scr + K::SC_V;
  R* vs = scr + K::SC_v;
  R* Qx = scr + K::SC_Q;
  R* xs = scr + K::SC_X;
  R* red = scr + K::SC_R;
  R* kst = a.k_in_smem ? kstore_all + (size_t)pw * T * KT : scr + K::SC_K;
  R* gKs = (R*)a.Ks;
  R* gks = (R*)a.ks;

  const R s_lo = (R)a.u_lo, s_hi = (R)a.u_hi, s_du = (R)a.delta_u;
  const R decay = (R)a.ls_decay;
  const bool has_mask = MODE == MODE_MASK || (BOX && a.has_mask);

  int stg = 0;
  uint32_t ph = 0;
  unsigned status = 0u;
  R oldcost_part = R(0);
  R kprev[M];
#pragma unroll
  for (int q = 0; q < M; ++q) kprev[q] = R(0);

  // ======================= backward Riccati sweep (lqr_step.py:61-158) =======================
  for (int t = T - 1; t >= 0; --t) {
    mbar_wait(&full[stg], ph);
    const R* st = (const R*)(stage_base + (size_t)stg * K::STAGE_BYTES);
    const unsigned char* mk = (const unsigned char*)(st + K::OFF_END) + pw * M;

    // nominal point tau_bar = [x_bar; u_bar] replicated on every lane
    Vec<R, P> tb;
    {
      Vec<R, N> tx;
      Vec<R, M> tu;
      tx.template load<A_N>(st + ox);
      tu.template load<A_M>(st + ou);
#pragma unroll
      for (int i = 0; i < N; ++i) tb.set(i, tx.get(i));
#pragma unroll
      for (int q = 0; q < M; ++q) tb.set(N + q, tu.get(q));
      if constexpr (P & 1) tb.p[Vec<R, P>::NP - 1].y = R(0);
    }
    // owned columns of C_t; c_back = (C_t tau_bar) + c for the owned rows  (lqr_step.py:289-295)
    Vec<R, P> Qc[CPL];
    R qj[CPL];
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) {
      Qc[sl].gather(st + oC + cc[sl], P);
      Vec<R, P> Crow;
      Crow.template load<A_ROW>(st + oC + cc[sl] * P);
      const R Ct = Crow.dot(tb);
      const R cj = st[oc + cc[sl]];
      const R tbj = st[(isx[sl] ? ox : ou - N) + cc[sl]];
      if (wsl[sl]) oldcost_part += tbj * (R(0.5) * Ct + cj);   // util.get_cost of the nominal trajectory (:169)
      qj[sl] = Ct + cj;
    }
    if (t < T - 1) {                            // Q = C + F'VF, q = c_back + F'v  (:66-70)
      Vec<R, N> Fcol[CPL], Wc[CPL];
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl) Fcol[sl].gather(st + oF + cc[sl], P);
      {
#pragma unroll
        for (int sl = 0; sl < CPL; ++sl) Wc[sl].zero();
#pragma unroll
        for (int k = 0; k < N; ++k) {             // W[:, c] = V F[:, c]; each V column load feeds CPL columns
          Vec<R, N> Vcol;                         // Vs holds V transposed: row k of Vs == column k of V
          Vcol.template load<EA>(Vs + k * VS);
#pragma unroll
          for (int sl = 0; sl < CPL; ++sl) Wc[sl].axpy(Vcol, Fcol[sl].get(k));
        }
        static_for<0, N>([&](auto kc) {           // Q[:, c] += F' W[:, c]; rows of F are contiguous
          constexpr int k = decltype(kc)::value;
          constexpr int AK = k % 2 == 0 ? A_2P : align_elems<R>(P);
          Vec<R, P> Frow;
          Frow.template load<(AK < A_NP ? AK : A_NP)>(st + oF + k * P);
#pragma unroll
          for (int sl = 0; sl < CPL; ++sl) Qc[sl].axpy(Frow, Wc[sl].get(k));
        });
      }
      Vec<R, N> vv;
      vv.template load<EA>(vs);
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl) qj[sl] += Fcol[sl].dot(vv);
    }
    // replicate Q_uu, q_u on every lane of the problem (column n+b2 lives in lane (n+b2)%LP, slot (n+b2)/LP)
    R Quu[M][M], qu[M];
#pragma unroll
    for (int p2 = 0; p2 < M; ++p2) {
      constexpr int dummy = 0;
      (void)dummy;
      const int src = base + (N + p2) % LP;
#pragma unroll
      for (int p1 = 0; p1 < M; ++p1) Quu[p1][p2] = shfl(Qc[(N + p2) / LP].get(N + p1), src);
      qu[p2] = shfl(qj[(N + p2) / LP], src);
    }
    R kk[M];
    unsigned fm = FULLM;
    int it = 0;
    Ldl<R, M> fac;
    if constexpr (BOX) {                         // (:129-148)
      R lb[M], ub[M];
#pragma unroll
      for (int q = 0; q < M; ++q) {
        const R lo_abs = a.bounds_kind == 2 ? st[olo + q] : s_lo;
        const R hi_abs = a.bounds_kind == 2 ? st[ohi + q] : s_hi;
        const R ubq = tb.get(N + q);
        lb[q] = lo_abs - ubq;
        ub[q] = hi_abs - ubq;
        if (a.has_delta) {
          if (lb[q] < -s_du) lb[q] = -s_du;
          if (ub[q] > s_du) ub[q] = s_du;
        }
        kk[q] = kprev[q];
      }
      if (!valid) {   // padding problems of a tail CTA compute on stale shared memory: give their
                      // (data dependent) pnqp loop a trivial QP so they never become the slowest warp
#pragma unroll
        for (int p1 = 0; p1 < M; ++p1) {
#pragma unroll
          for (int p2 = 0; p2 < M; ++p2) Quu[p1][p2] = p1 == p2 ? R(1) : R(0);
          qu[p1] = R(0);
          lb[p1] = R(-1);
          ub[p1] = R(1);
          kk[p1] = R(0);
        }
      }
      bool conv, badpiv;
      pnqp_lane<R, M>(Quu, qu, lb, ub, t < T - 1, kk, fac, fm, it, conv, badpiv, a.pnqp_iters);
      if (!conv) status |= 1u;
      if (badpiv) status |= 4u;
#pragma unroll
      for (int q = 0; q < M; ++q) kprev[q] = kk[q];
    } else {                                     // unconstrained (:84-94) or u_zero_I masked (:100-127)
      if constexpr (MODE == MODE_MASK) {
        unsigned zm = 0u;
#pragma unroll
        for (int q = 0; q < M; ++q) zm |= (mk[q] ? 1u : 0u) << q;
        fm = FULLM & ~zm;
      }
      R A[M][M], rhs[M], sol[M];
#pragma unroll
      for (int p1 = 0; p1 < M; ++p1) {
        const bool f1 = (fm >> p1) & 1u;
        rhs[p1] = f1 ? qu[p1] : R(0);
#pragma unroll
        for (int p2 = 0; p2 < M; ++p2) A[p1][p2] = (f1 && ((fm >> p2) & 1u)) ? Quu[p1][p2] : R(0);
        if (!f1) A[p1][p1] += R(1e-8);
      }
      fac.factor(A);
      if (fac.bad) status |= 4u;
      fac.solve(rhs, sol);
#pragma unroll
      for (int q = 0; q < M; ++q) kk[q] = -sol[q];
    }
    // K[:, c] = -Hff^{-1} Qux_f[:, c] for the owned columns (rows of clamped / masked controls are zero)
    R Kc[CPL][M];
    R* Kt = a.k_in_smem ? kst + (size_t)t * KT : kst;
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) {
      R rhs[M], sol[M];
#pragma unroll
      for (int q = 0; q < M; ++q) rhs[q] = ((fm >> q) & 1u) ? Qc[sl].get(N + q) : R(0);
      fac.solve(rhs, sol);
#pragma unroll
      for (int q = 0; q < M; ++q) Kc[sl][q] = -sol[q];
      if (wsl[sl]) {
        if (isx[sl]) {
#pragma unroll
          for (int q = 0; q < M; ++q) Kt[q * VS + cc[sl]] = Kc[sl][q];
        } else {                          // u column: publish Q[:n, n+ua] = Q_xu[:, ua]
          Vec<R, N> qxu;                    // rows < n of the control column
#pragma unroll
          for (int i = 0; i < N; ++i) qxu.set(i, Qc[sl].get(i));
          qxu.template store<EA>(Qx + ua[sl] * VS);
        }
      }
    }
    if (j == 0) {
#pragma unroll
      for (int q = 0; q < M; ++q) Kt[M * VS + q] = kk[q];
    }
    if (wr) {
      const size_t tbo = (size_t)t * B + b;
      if (gKs != nullptr) {
#pragma unroll
        for (int sl = 0; sl < CPL; ++sl) {
          if (wsl[sl] && isx[sl]) {
#pragma unroll
            for (int q = 0; q < M; ++q) gKs[(tbo * M + q) * N + cc[sl]] = Kc[sl][q];
          }
        }
        if (j == 0) {
#pragma unroll
          for (int q = 0; q < M; ++q) gks[tbo * M + q] = kk[q];
        }
      }
      if (BOX && a.qp_iters != nullptr && j == 0) a.qp_iters[tbo] = it;
      if (a.free_mask != nullptr) {
#pragma unroll
        for (int sl = 0; sl < CPL; ++sl)
          if (wsl[sl] && !isx[sl]) a.free_mask[tbo * M + ua[sl]] = (fm >> ua[sl]) & 1u;
      }
    }
    __syncwarp();

    // V = Qxx + Qxu K + K'Qux + K'Quu K ; v = qx + Qxu k + K'qu + K'Quu k   (:155-158)
    Vec<R, N> Vn[CPL];
    R vn[CPL], G[CPL][M];
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) {
#pragma unroll
      for (int p1 = 0; p1 < M; ++p1) {
        R sacc = Qc[sl].get(N + p1);
#pragma unroll
        for (int p2 = 0; p2 < M; ++p2) sacc += Quu[p1][p2] * Kc[sl][p2];
        G[sl][p1] = sacc;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) Vn[sl].set(i, Qc[sl].get(i));
      if constexpr (N & 1) Vn[sl].p[Vec<R, N>::NP - 1].y = R(0);
      vn[sl] = qj[sl];
    }
#pragma unroll
    for (int q = 0; q < M; ++q) {
      Vec<R, N> Qrow, Krow;
      Qrow.template load<EA>(Qx + q * VS);
      Krow.template load<EA>(Kt + q * VS);
      R sacc = qu[q];
#pragma unroll
      for (int p2 = 0; p2 < M; ++p2) sacc += Quu[q][p2] * kk[p2];
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl) {
        Vn[sl].axpy(Qrow, Kc[sl][q]);
        Vn[sl].axpy(Krow, G[sl][q]);
        vn[sl] += Qx[q * VS + fr[sl]] * kk[q] + Kc[sl][q] * sacc;
      }
    }
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) {
      if (wsl[sl] && isx[sl]) {
        Vn[sl].template store<EA>(Vs + cc[sl] * VS);        // column c of V, stored as row c (vector stores)
        vs[cc[sl]] = vn[sl];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stg]);
    if (++stg == K::S) { stg = 0; ph ^= 1u; }
  }

  // nominal cost  (sum of the lanes' partial sums, fixed order)
  if (writer_lane) red[j] = oldcost_part;
  __syncwarp();
  R oldcost = R(0);
#pragma unroll
  for (int i = 0; i < LP; ++i) oldcost += red[i];
  __syncwarp();

  if (!a.do_rollout) {
    if (wr && j == 0 && a.status != nullptr) a.status[b] = (int)status;
    return;
  }

  // ======================= rollout + line search (lqr_step.py:164-261) =======================
  const R* gx0 = (const R*)a.x_init;
  R* gnx = (R*)a.new_x;
  R* gnu = (R*)a.new_u;
  R* gdu1 = (R*)a.du_first;
  R alpha = R(1), fdn = R(0), cost = R(0);
  bool worse = false;
  for (int pass = 0;; ++pass) {
    Vec<R, N> xr;
#pragma unroll
    for (int i = 0; i < N; ++i) xr.set(i, valid ? gx0[(size_t)b * N + i] : R(0));
    if constexpr (N & 1) xr.p[Vec<R, N>::NP - 1].y = R(0);
    R xown[CPL];
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) xown[sl] = valid ? gx0[(size_t)b * N + fr[sl]] : R(0);
    R cpart = R(0), dun2 = R(0);
    R kcol[M], kff[M];                           // KREDUCE with gains in global memory: column of K_t, k_t
#pragma unroll
    for (int q = 0; q < M; ++q) {
      kcol[q] = R(0);
      kff[q] = R(0);
    }
    if constexpr (K::KREDUCE) {
      if (!a.k_in_smem) {
        const size_t kb = (size_t)(valid ? b : 0) * M;
#pragma unroll
        for (int q = 0; q < M; ++q) {
          kcol[q] = __ldcg(gKs + (kb + q) * N + fr[0]);
          kff[q] = __ldcg(gks + kb + q);
        }
      }
    }
    size_t orow = (size_t)b;                     // t*B + b
    for (int t = 0; t < T; ++t, orow += (size_t)B) {
      mbar_wait(&full[stg], ph);
      const R* st = (const R*)(stage_base + (size_t)stg * K::STAGE_BYTES);
      const unsigned char* mk = (const unsigned char*)(st + K::OFF_END) + pw * M;

      Vec<R, N> xb, dxv;
      Vec<R, M> ubar;
      xb.template load<A_N>(st + ox);
      ubar.template load<A_M>(st + ou);
#pragma unroll
      for (int k2 = 0; k2 < Vec<R, N>::NP; ++k2) dxv.p[k2] = fma2(xb.p[k2], P2<R>{R(-1), R(-1)}, xr.p[k2]);
      R u[M];
      if (a.k_in_smem) {
        const R* Kt = kst + (size_t)t * KT;
#pragma unroll
        for (int q = 0; q < M; ++q) {
          Vec<R, N> Krow;
          Krow.template load<EA>(Kt + q * VS);
          u[q] = (Krow.dot(dxv) + ubar.get(q)) + alpha * Kt[M * VS + q];      // (:192)
        }
      } else if constexpr (K::KREDUCE) {
        // gains in the caller's buffer: this lane holds column cc of K_t and k_t (prefetched one step
        // ahead); K dx is reduced over the n state lanes with a butterfly, then broadcast
        const R dxi = xown[0] - st[ox + fr[0]];
        R part[M];
#pragma unroll
        for (int q = 0; q < M; ++q) part[q] = isx[0] ? kcol[q] * dxi : R(0);
#pragma unroll
        for (int off = N / 2; off >= 1; off >>= 1) {
#pragma unroll
          for (int q = 0; q < M; ++q) part[q] += __shfl_xor_sync(0xffffffffu, part[q], off);
        }
#pragma unroll
        for (int q = 0; q < M; ++q) u[q] = (shfl(part[q], base) + ubar.get(q)) + alpha * kff[q];
        if (t + 1 < T) {                         // prefetch the next step's column (L2 hit) behind this step
          const size_t kb = ((size_t)(t + 1) * B + (valid ? b : 0)) * M;
#pragma unroll
          for (int q = 0; q < M; ++q) {
            kcol[q] = __ldcg(gKs + (kb + q) * N + fr[0]);
            kff[q] = __ldcg(gks + kb + q);
          }
        }
      } else {
        const R* Kg = gKs + ((size_t)t * B + (valid ? b : 0)) * M * N;
        const R* kg = gks + ((size_t)t * B + (valid ? b : 0)) * M;
#pragma unroll
        for (int q = 0; q < M; ++q) {
          R sacc = R(0);
#pragma unroll
          for (int i = 0; i < N; ++i) sacc += __ldcg(Kg + q * N + i) * dxv.get(i);
          u[q] = (sacc + ubar.get(q)) + alpha * __ldcg(kg + q);
        }
      }
#pragma unroll
      for (int q = 0; q < M; ++q) {
        if constexpr (MODE != MODE_PLAIN) {
          if (has_mask && mk[q]) u[q] = R(0);                     // (:197-198)
        }
        if constexpr (BOX) {                                      // (:200-213)
          R lo = a.bounds_kind == 2 ? st[olo + q] : s_lo;
          R hi = a.bounds_kind == 2 ? st[ohi + q] : s_hi;
          if (a.has_delta) {
            const R l2 = ubar.get(q) - s_du, h2 = ubar.get(q) + s_du;
            lo = l2 < lo ? lo : l2;
            hi = h2 > hi ? hi : h2;
          }
          u[q] = u[q] < lo ? lo : u[q];                                   // util.eclamp order: lower, then upper
          u[q] = u[q] > hi ? hi : u[q];
        }
        const R d = ubar.get(q) - u[q];
        dun2 += d * d;
      }
      Vec<R, P> tau;
#pragma unroll
      for (int i = 0; i < N; ++i) tau.set(i, xr.get(i));
#pragma unroll
      for (int q = 0; q < M; ++q) tau.set(N + q, u[q]);
      if constexpr (P & 1) tau.p[Vec<R, P>::NP - 1].y = R(0);

      R xn[CPL];
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl) {
        R tj = xown[sl];
        if (!isx[sl]) {
#pragma unroll
          for (int q = 0; q < M; ++q)
            if (q == ua[sl]) tj = u[q];
        }
        Vec<R, P> Crow;
        Crow.template load<A_ROW>(st + oC + cc[sl] * P);
        const R Ct = Crow.dot(tau);
        if (wsl[sl]) cpart += tj * (R(0.5) * Ct + st[oc + cc[sl]]);           // (:232)
        if (wr && wsl[sl]) {
          if (isx[sl]) {
            gnx[orow * N + cc[sl]] = tj;
          } else {
            gnu[orow * M + ua[sl]] = tj;
            if (pass == 0 && gdu1 != nullptr) gdu1[orow * M + ua[sl]] = st[ou + ua[sl]] - tj;
          }
        }
        xn[sl] = R(0);
        if (t < T - 1) {                                          // (:217-222), or true_dynamics(x, u) (:224-225)
          bool known = false;
          if constexpr (M == 1 && (N == DynDims<DYN_CARTPOLE>::N || N == DynDims<DYN_PENDULUM>::N)) {
            if (a.dyn_kind != DYN_LINEAR) {       // every lane evaluates the step function and keeps its row
              known = true;
              R sv[N], ov[N];
#pragma unroll
              for (int i = 0; i < N; ++i) sv[i] = tau.get(i);
              dyn_step<R, N == DynDims<DYN_CARTPOLE>::N ? DYN_CARTPOLE : DYN_PENDULUM, R>(a.dp, sv, tau.get(N), ov);
              xn[sl] = ov[0];
#pragma unroll
              for (int i = 1; i < N; ++i)
                if (fr[sl] == i) xn[sl] = ov[i];
            }
          }
          if (!known) {
            Vec<R, P> Frow;
            Frow.template load<A_ROW>(st + oF + fr[sl] * P);
            xn[sl] = Frow.dot(tau);
            if (a.has_f) xn[sl] += st[of_ + fr[sl]];
          }
        }
      }
      if (t < T - 1) {
        R* xsb = xs + (t & 1) * VS;
#pragma unroll
        for (int sl = 0; sl < CPL; ++sl) {
          if (wsl[sl] && isx[sl]) xsb[cc[sl]] = xn[sl];
          xown[sl] = xn[sl];
        }
        __syncwarp();
        xr.template load<EA>(xsb);
        if constexpr (N & 1) xr.p[Vec<R, N>::NP - 1].y = R(0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stg]);
      if (++stg == K::S) { stg = 0; ph ^= 1u; }
    }
    if (writer_lane) red[j] = cpart;
    __syncwarp();
    cost = R(0);
#pragma unroll
    for (int i = 0; i < LP; ++i) cost += red[i];
    __syncwarp();
    if (pass == 0) fdn = sqrt(dun2);                              // (:243-245)
    worse = cost > oldcost;
    const bool more = pass + 1 < a.max_ls;
    if (worse) alpha *= decay;                                    // (:247)
    if (wr && j == 0 && worse && more) votes[pass & 31] = 1;
    if (warp == 0 && lane == 0) votes[(pass + 16) & 31] = 0;
    named_bar_sync(1, K::THREADS);
    const int cont = votes[pass & 31];
    if (!cont || !more) break;
  }
  if (worse) alpha /= decay;                                      // (:252)
  if (wr && j == 0) {
    ((R*)a.costs)[b] = cost;
    ((R*)a.full_du_norm)[b] = fdn;
    ((R*)a.alphas)[b] = alpha;
    if (!(cost - cost == R(0))) status |= 2u;
    if (a.status != nullptr) a.status[b] = (int)status;
  }
}

template <typename R, int N, int M, int MODE>
__global__ void __launch_bounds__(StepCfg<R, N, M>::THREADS)
lqr_step_kernel(const StepArgs a) {
  using K = StepCfg<R, N, M>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + K::S;
  volatile int* votes = reinterpret_cast<volatile int*>(smem_raw + 128);
  unsigned char* stage_base = smem_raw + K::HDR_BYTES;
  R* scratch = reinterpret_cast<R*>(stage_base + (size_t)K::S * K::STAGE_BYTES);
  R* kstore = scratch + (size_t)K::W * K::SCR;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * K::W;
  const int cnt = min(K::W, a.B - b0);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < K::S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], K::NW);
    }
    mbar_fence_init();
  }
  if (tid < 32) votes[tid] = 0;
  __syncthreads();
  if (warp == K::NW) {
    step_producer<R, N, M>(a, stage_base, full, empty, votes, b0, cnt, lane);
  } else {

    step_consumer<R, N, M, MODE>(a, stage_base, full, empty, votes, scratch, kstore, b0, warp, lane);
  }
}

template <typename R, int N, int M, int MODE>
int launch_step_mode(const StepArgs& args, int max_smem_optin, cudaStream_t stream) {
  using K = StepCfg<R, N, M>;
  StepArgs a = args;
  size_t smem = K::smem_bytes(a.T, true);
  a.k_in_smem = 1;
  const bool have_ws = a.Ks != nullptr && a.ks != nullptr;
  if (smem > (size_t)max_smem_optin || (K::prefers_workspace(a.T, max_smem_optin) && have_ws && a.do_rollout)) {
    a.k_in_smem = 0;
    smem = K::smem_bytes(a.T, false);
    if (smem > (size_t)max_smem_optin) return 4;
    if (a.do_rollout && !have_ws) return 4;
  }
  auto kern = lqr_step_kernel<R, N, M, MODE>;
  // per device: the opt-in shared-memory attribute is per context.  Atomic flags: concurrent callers (one host
  // thread per GPU is the expected pattern) may both set the attribute - idempotent - but never read a torn value.
  static std::atomic<int> configured[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 6;
  if (configured[dev].load(std::memory_order_acquire) < (int)smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin) !=
        cudaSuccess)
      return 5;
    configured[dev].store(max_smem_optin, std::memory_order_release);
  }
  const int grid = (a.B + K::W - 1) / K::W;
  kern<<<grid, K::THREADS, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 5;
}

template <typename R, int N, int M>
int launch_step(const StepArgs& a, int max_smem_optin, cudaStream_t stream) {
  if (a.bounds_kind != 0) return launch_step_mode<R, N, M, MODE_BOX>(a, max_smem_optin, stream);
  if (a.has_mask) return launch_step_mode<R, N, M, MODE_MASK>(a, max_smem_optin, stream);
  return launch_step_mode<R, N, M, MODE_PLAIN>(a, max_smem_optin, stream);
}

template <typename R, int N, int M>
int step_prefers_workspace(int T, int max_smem_optin) {
  return StepCfg<R, N, M>::prefers_workspace(T, max_smem_optin) ? 1 : 0;
}

template <typename R, int N, int M>
size_t step_smem_query(int T) {
  return StepCfg<R, N, M>::smem_bytes(T, true);
}

}  // namespace mpcb200
