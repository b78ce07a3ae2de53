/*
 * mpcb200.h - C ABI of the B200-native batched box-constrained LQR step.
 *
 * This is the drop-in boundary for ONE path of locuslab/mpc.pytorch: the body of
 * LQRStepFn.forward / LQRStepFn.backward (reference mpc/lqr_step.py:277-309 and
 * :312-407).  The reference has no FFI of its own (it is pure Python on top of
 * ATen); these entry points are what a binding for that path would call.  The
 * Python host side (mpc/pytorch_b200/) mirrors the reference's LQRStep / MPC
 * signatures on top of this ABI through ctypes; see INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types, no exceptions, no prints.
 *  - every pointer is a DEVICE pointer on the current device; all tensors are
 *    dense, row-major, time-major / batch-second exactly like the reference:
 *      C[T,B,p,p]  c[T,B,p]  F[F_T,B,n,p] (F_T = T-1 or T)  f[T-1,B,n] or NULL
 *      x_init[B,n] cur_x[T,B,n] cur_u[T,B,m]  (p = n+m)
 *  - the caller owns every buffer; the library allocates nothing, frees nothing
 *    and never writes an input.  Calls are asynchronous on `stream`, re-entrant,
 *    and keep no global state besides a launch counter.
 *  - return value: 0 on success, MPCB200_ERR_* otherwise (mpcb200_strerror()).
 *  - optional outputs may be NULL.
 */
#ifndef MPCB200_H_
#define MPCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB200_VERSION 2

enum {
  MPCB200_OK = 0,
  MPCB200_ERR_NULL_POINTER = 1,      /* a required pointer is NULL                          */
  MPCB200_ERR_BAD_DIMS = 2,          /* B,T,n,m <= 0, F_T not in {T-1,T}, bad option combo  */
  MPCB200_ERR_UNSUPPORTED_DIMS = 3,  /* (n,m) has no compiled kernel instance               */
  MPCB200_ERR_SMEM = 4,              /* problem does not fit shared memory and no workspace */
  MPCB200_ERR_LAUNCH = 5,            /* cudaGetLastError() after launch != cudaSuccess      */
  MPCB200_ERR_NO_DEVICE = 6          /* no usable sm_100 device / wrong architecture        */
};

/* Problem sizes and options.  Mirrors the closure arguments of
 * LQRStep(...) (reference mpc/lqr_step.py:22-38). */
typedef struct mpcb200_dims {
  int32_t B;              /* n_batch                                                   */
  int32_t T;              /* horizon                                                   */
  int32_t n;              /* n_state                                                   */
  int32_t m;              /* n_ctrl                                                    */
  int32_t F_T;            /* time slices present in F: T-1 or T (only F[:T-1] is read) */
  int32_t has_f;          /* f given ([T-1,B,n]); 0 = reference's "empty tensor"       */
  int32_t bounds_kind;    /* 0 none, 1 scalar (params.u_lo/u_hi), 2 tensors [T,B,m]    */
  int32_t has_zero_mask;  /* u_zero_I given: uint8 [T,B,m], nonzero = forced-zero ctrl */
  int32_t has_delta_u;    /* trust region |du| <= params.delta_u (needs bounds)        */
  int32_t max_ls_iter;    /* max_linesearch_iter (>=1)                                 */
  int32_t pnqp_max_iter;  /* projected-Newton iteration cap (reference: 20)            */
  int32_t do_rollout;     /* 1: Riccati sweep + line-search rollout (LinDx/QuadCost)
                             0: Riccati sweep only; Ks/ks must be given                */
  int32_t dynamics_kind;  /* true dynamics of the rollout (reference lqr_step.py:217-225):
                             MPCB200_DYN_LINEAR = LinDx(F,f); MPCB200_DYN_CARTPOLE / _PENDULUM = the step
                             function of that system evaluated inside the kernel (params.dyn); F,f are
                             then its linearisation and are used by the Riccati sweep only (ABI v2)  */
  int32_t reserved0;      /* 0 (keeps the 64-bit fields below naturally aligned)              */
  /* Elements between consecutive TIME slices of C, c, F, f (ABI v2).  0: dense ([T,B,...] contiguous; what a
   * zero-initialised struct means).  > 0: that many elements.  MPCB200_TIME_INVARIANT (-1): one [B,...] slice
   * reused for every t (a torch stride of 0), which is what the reference's
   * `C.unsqueeze(0).expand(T, ...)` hands over (mpc/mpc.py:205-226) and what an LTI system is; the slice is
   * read from HBM / copied from the host once instead of T times.  The batch dimension stays dense. */
  int64_t C_tstride, c_tstride, F_tstride, f_tstride;
} mpcb200_dims;

typedef struct mpcb200_params {
  double u_lo, u_hi;      /* scalar bounds (bounds_kind == 1)   */
  double delta_u;         /* has_delta_u                        */
  double ls_decay;        /* linesearch_decay                   */
  double dyn[8];          /* parameters of a known system (dims.dynamics_kind != 0), see mpcb200_dyn_* (ABI v2) */
} mpcb200_params;

/* Known nonlinear systems (reference mpc/env_dx/cartpole.py:63-96, mpc/env_dx/pendulum.py:49-84).
 * dyn[] = cartpole: gravity, masscart, masspole, length, force_mag, dt   (state x,dx,cos th,sin th,dth; n=5, m=1)
 *         pendulum: g, m, l, (unused), max_torque, dt                    (state cos th,sin th,dth; n=3, m=1) */
enum { MPCB200_DYN_LINEAR = 0, MPCB200_DYN_CARTPOLE = 1, MPCB200_DYN_PENDULUM = 2 };
#define MPCB200_TIME_INVARIANT (-1)

/* Per-problem status bits written to `status[B]`. */
#define MPCB200_ST_PNQP_UNCONVERGED 1u /* some time step hit pnqp_max_iter (reference prints a warning, pnqp.py:81) */
#define MPCB200_ST_NONFINITE 2u        /* final cost is not finite                                              */
#define MPCB200_ST_BAD_PIVOT 4u        /* an LDL^T pivot of the free block was <= 0 (Quu not positive definite) */

/*
 * One LQR step: replaces LQRStepFn.forward (reference mpc/lqr_step.py:277-309):
 *   c_back = C*tau_bar + c; lqr_backward (:52-160) incl. pnqp (mpc/pnqp.py:5-82);
 *   lqr_forward rollout + line search (:164-261) with true model QuadCost(C,c)/LinDx(F,f).
 *
 * Outputs: new_x[T,B,n] new_u[T,B,m] costs[B] full_du_norm[B] alphas[B]
 *          (mean_alphas of the reference is mean(alphas); reduced by the caller)
 * Optional outputs (NULL to skip):
 *   du_first[T,B,m]  cur_u - new_u of the FIRST (alpha = 1) rollout pass, i.e. the vector whose
 *                  per-problem 2-norm is full_du_norm.  (The reference computes its
 *                  full_du_norm from a [T,m,B]-ordered buffer viewed as [B,T*m]
 *                  (lqr_step.py:244-245), which mixes batch elements when B > 1; the
 *                  Python host side reproduces that from du_first, this ABI returns the
 *                  per-problem norm.)
 *   qp_iters[T,B]  int32  pnqp iterations "i" per (t,b); the reference's
 *                  n_total_qp_iter is sum_t (1 + max_b qp_iters[t,b])   (lqr_step.py:140)
 *   free_mask[T,B,m] uint8  pnqp free set If (1 = free) / complement of u_zero_I
 *   status[B]      int32  MPCB200_ST_* bits
 *   Ks[T,B,m,n] ks[T,B,m]  feedback gains in FORWARD time order
 * pnqp semantics are per problem (what the reference computes for n_batch=1).
 */
int mpcb200_lqr_step_f32(const mpcb200_dims* dims, const mpcb200_params* params,
                         const float* C, const float* c, const float* F, const float* f,
                         const float* x_init, const float* cur_x, const float* cur_u,
                         const float* u_lower, const float* u_upper, const uint8_t* u_zero_I,
                         float* new_x, float* new_u, float* costs, float* full_du_norm,
                         float* alphas, float* du_first, int32_t* qp_iters, uint8_t* free_mask, int32_t* status,
                         float* Ks, float* ks, void* stream);

int mpcb200_lqr_step_f64(const mpcb200_dims* dims, const mpcb200_params* params,
                         const double* C, const double* c, const double* F, const double* f,
                         const double* x_init, const double* cur_x, const double* cur_u,
                         const double* u_lower, const double* u_upper, const uint8_t* u_zero_I,
                         double* new_x, double* new_u, double* costs, double* full_du_norm,
                         double* alphas, double* du_first, int32_t* qp_iters, uint8_t* free_mask, int32_t* status,
                         double* Ks, double* ks, void* stream);

/*
 * Gradient assembly of the KKT adjoint: replaces the second half of
 * LQRStepFn.backward (reference mpc/lqr_step.py:342-404).  The caller first runs
 * mpcb200_lqr_step_* with (c = -[dl_dx;dl_du], f = NULL, x_init = 0, cur_x = cur_u = 0,
 * u_zero_I = active set, no bounds) to obtain (dx,du) (reference :328-340), then:
 *   dC[t] = -0.5 (dtau tau' + tau dtau'), dc = -dtau,
 *   lambda / dlambda costate recursions, dF[t] = -(dlam_{t+1} tau_t' + lam_{t+1} dtau_t'),
 *   df = -dlam[1:], dx_init = -dlam[0].
 * r = [dl_dx; dl_du] enters only through r_x.  dF has F_T slices (slice T-1, if present, is zeroed).
 * df may be NULL (reference returns an empty tensor when f is empty).
 * workspace: optional device buffer of 2*T*B*n elements (lambda_t, dlambda_t).  With it the
 * call runs as two kernels (sequential costates, then fully parallel outer products - the fast
 * path); with NULL it runs as one fused kernel.  Results are identical.
 */
int mpcb200_lqr_grad_f32(const mpcb200_dims* dims,
                         const float* C, const float* c, const float* F,
                         const float* new_x, const float* new_u,
                         const float* dx, const float* du, const float* dl_dx,
                         float* dx_init, float* dC, float* dc, float* dF, float* df,
                         void* workspace, void* stream);

int mpcb200_lqr_grad_f64(const mpcb200_dims* dims,
                         const double* C, const double* c, const double* F,
                         const double* new_x, const double* new_u,
                         const double* dx, const double* du, const double* dl_dx,
                         double* dx_init, double* dC, double* dc, double* dF, double* df,
                         void* workspace, void* stream);

/*
 * The whole KKT-adjoint backward in ONE call: replaces LQRStepFn.backward (reference mpc/lqr_step.py:312-407).
 * Given the solution (new_x, new_u) of the forward solve and the upstream gradients dl_dx[T,B,n], dl_du[T,B,m]:
 *   1. prep kernel: r = -[dl_dx; dl_du]; active set I = (|u* - u_lower| <= 1e-8) | (|u* - u_upper| <= 1e-8)
 *      from the box (dims->bounds_kind / params->u_lo,u_hi / u_lower,u_upper), reference :316-326;
 *   2. the masked LQR step from the zero trajectory (c = r, x_init = 0, u_zero_I = I, default line search),
 *      i.e. the nested MPC(lqr_iter=1) of :328-340, with the step kernel;
 *   3. costates and outer products (mpcb200_lqr_grad_*, :342-404).
 * Outputs dx_init[B,n] dC[T,B,p,p] dc[T,B,p] dF[F_T,B,n,p] df[T-1,B,n] (df only if dims->has_f).
 * workspace: device buffer of mpcb200_adjoint_workspace_bytes(dims, elem_size) bytes (scratch; contents
 * undefined on return).  Only dims->{B,T,n,m,F_T,has_f,bounds_kind} and params->{u_lo,u_hi} are read.
 */
size_t mpcb200_adjoint_workspace_bytes(const mpcb200_dims* dims, int32_t elem_size);
int mpcb200_lqr_adjoint_f32(const mpcb200_dims* dims, const mpcb200_params* params,
                            const float* C, const float* c, const float* F,
                            const float* new_x, const float* new_u, const float* dl_dx, const float* dl_du,
                            const float* u_lower, const float* u_upper,
                            float* dx_init, float* dC, float* dc, float* dF, float* df,
                            void* workspace, size_t workspace_bytes, void* stream);
int mpcb200_lqr_adjoint_f64(const mpcb200_dims* dims, const mpcb200_params* params,
                            const double* C, const double* c, const double* F,
                            const double* new_x, const double* new_u, const double* dl_dx, const double* dl_du,
                            const double* u_lower, const double* u_upper,
                            double* dx_init, double* dC, double* dc, double* dF, double* df,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * Nominal trajectory under LinDx dynamics: replaces util.get_traj for LinDx (reference
 * mpc/util.py:102-126, called once per iLQR iteration at mpc/mpc.py:251):
 *   x[0] = x_init;  x[t+1] = F[t] [x[t]; u[t]] + f[t]   (f may be NULL; dims->has_f)
 * x[T,B,n] is written; only dims->{B,T,n,m,F_T,has_f} are read.
 */
int mpcb200_rollout_f32(const mpcb200_dims* dims, const float* F, const float* f, const float* x_init,
                        const float* u, float* x, void* stream);
int mpcb200_rollout_f64(const mpcb200_dims* dims, const double* F, const double* f, const double* x_init,
                        const double* u, double* x, void* stream);

/*
 * Known nonlinear dynamics on the device.
 *   mpcb200_dyn_rollout_*:   x[0] = x_init, x[t+1] = step(x[t], u[t])  - util.get_traj for a Module
 *                            (reference mpc/util.py:102-126 with dynamics(x,u), one Python call per step).
 *   mpcb200_dyn_linearize_*: F[t,b] = [d step/dx, d step/du], f[t,b] = step(x,u) - F [x;u] at (x[t,b], u[t,b]),
 *                            t < T-1 - linearize_dynamics (reference mpc/mpc.py:490-601; its AUTO_DIFF mode does
 *                            (T-1)*n_state autograd passes).  Exact Jacobians by forward-mode dual numbers.
 * kind = MPCB200_DYN_CARTPOLE | MPCB200_DYN_PENDULUM; dyn = HOST pointer to 8 doubles (see mpcb200_params.dyn);
 * x_init[B,n] u[T,B,m] x[T,B,n] F[T-1,B,n,n+m] f[T-1,B,n]; n, m are those of the system.
 */
int mpcb200_dyn_rollout_f32(int32_t kind, const double* dyn, int32_t B, int32_t T, const float* x_init,
                            const float* u, float* x, void* stream);
int mpcb200_dyn_rollout_f64(int32_t kind, const double* dyn, int32_t B, int32_t T, const double* x_init,
                            const double* u, double* x, void* stream);
int mpcb200_dyn_linearize_f32(int32_t kind, const double* dyn, int32_t B, int32_t T, const float* x,
                              const float* u, float* F, float* f, void* stream);
int mpcb200_dyn_linearize_f64(int32_t kind, const double* dyn, int32_t B, int32_t T, const double* x,
                              const double* u, double* F, double* f, void* stream);

/*
 * Standalone projected-Newton box QP, n <= 8: replaces pnqp(H,q,lower,upper,x_init,n_iter) of the
 * reference (mpc/pnqp.py:5-82) for batches of small QPs  min 0.5 x'Hx + q'x, lower <= x <= upper.
 * H[B,n,n] q,lower,upper[B,n]; x_init[B,n] or NULL (cold start -H^{-1}q).  Outputs x[B,n],
 * H_free[B,n,n] (the masked matrix H_ of the returning iteration, whose LU the reference returns),
 * If[B,n] uint8 free set, iters[B] (the reference's `i`), status[B] optional (MPCB200_ST_* bits).
 * Per-problem control flow (what the reference computes for n_batch = 1).
 */
int mpcb200_pnqp_f32(int32_t B, int32_t n, const float* H, const float* q, const float* lower,
                     const float* upper, const float* x_init, int32_t n_iter, float* x, float* H_free,
                     uint8_t* If, int32_t* iters, int32_t* status, void* stream);
int mpcb200_pnqp_f64(int32_t B, int32_t n, const double* H, const double* q, const double* lower,
                     const double* upper, const double* x_init, int32_t n_iter, double* x, double* H_free,
                     uint8_t* If, int32_t* iters, int32_t* status, void* stream);

/* 1 if a kernel instance for (n_state, n_ctrl) is compiled in, else 0. */
int mpcb200_supported(int32_t n_state, int32_t n_ctrl);

/* Fills `out` with up to `cap` supported (n,m) pairs (n0,m0,n1,m1,...); returns the count of pairs. */
int mpcb200_supported_list(int32_t* out, int32_t cap);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t mpcb200_launch_count(void);

/* Dynamic shared memory (bytes) the step kernel needs for these dims / element size (4 or 8); 0 if unsupported. */
size_t mpcb200_step_smem_bytes(const mpcb200_dims* dims, int32_t elem_size);

/* 1 if the step kernel wants caller-provided Ks/ks buffers for these dims: either the gain store of all
 * T steps does not fit shared memory, or (one-problem-per-warp shapes such as n=16) moving it out of shared
 * memory is what lets enough warps be resident.  Pass Ks[T,B,m,n], ks[T,B,m] then. */
int mpcb200_step_prefers_workspace(const mpcb200_dims* dims, int32_t elem_size);

int mpcb200_version(void);
const char* mpcb200_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif /* MPCB200_H_ */
