// kernels_cg.cuh -- amplitude draw: compute_rhs + compute_sample_vector +
// the block form of compute_Ax + cg_search, src/dang_cg_mod.f90:179-1100.
//
// The reference applies A = sum_nu T^t N^-1 T matrix-free, re-evaluating every SED twice per
// band per CG iteration.  For diffuse components A is block diagonal per (pixel, Stokes) with a
// CxC symmetric block M[c,c'] = sum_nu sed_c sed_c' / sigma^2, so K1 evaluates the SEDs once,
// stores the C(C+1)/2 block entries, and every CG iteration is a pure HBM stream over
// {M, x, r, d}: 15 doubles per (pixel, Stokes) for C = 2.
#pragma once
#include "common.cuh"

// Every CG pass form runs on the same grid (DG_CG_BLOCKS_PER_SM resident blocks per SM x SM count, or fewer when the
// sky slice is small): the deterministic grid reduction depends on the grid size, and the forms must agree to
// the last bit; the persistent solve kernel additionally needs every block resident.
#define DG_CG_BLOCKS_PER_SM 2

template <int C>
struct CgView {
  int comp[C];               // indices into ModelView::comp, in component_list order
  int S;                     // Stokes planes in this solve (1 or 2)
  int plane[2];              // 0-based plane numbers
  int nog;                   // components whose signal is subtracted from the data (:427-443)
  int og[DG_MAX_COMPS];
  int fluct;                 // 0: none (optimize), 1: reference indexing (Q1), 2: per-component
  double *M;                 // [T][S][Ppad], T = C(C+1)/2, row-major upper triangle
  double *x, *r, *d;         // [C][S][Ppad]
  const double *eta;         // [S][Ppad] injected normals, or nullptr -> Philox
  uint64_t seed;
  int store_d;               // 0: the recompute CG form never reads the initial d (= r)
};

// destination of the amplitude planes when unpack_amplitudes is folded into the final CG pass
// (per component: c%amplitude + plane[0] * Ppad, S contiguous planes), or nulls
template <int C>
struct CgAmpOut {
  double *p[C];
};

template <int C>
__device__ __forceinline__ int tri(int a, int b) {  // a <= b
  return a * C - a * (a - 1) / 2 + (b - a);
}

// One block-local CG step on a pair of elements, shared by every CG form (streaming, checkpointed
// recompute, persistent solve) with the roundings written out -- explicit FMAs, fixed order -- so that the
// forms agree to the last bit whatever the compiler would have contracted in each kernel:
//   d <- r + beta d (:305);  q = M d (:296);  r <- r - alpha q (:300)
template <int C>
__device__ __forceinline__ void cg_block_step(const double2 (&m)[C * (C + 1) / 2], double2 (&dv)[C], double2 (&rv)[C],
                                              double2 (&q)[C], double alpha, double beta) {
#pragma unroll
  for (int c = 0; c < C; c++) {
    dv[c].x = fma(beta, dv[c].x, rv[c].x);
    dv[c].y = fma(beta, dv[c].y, rv[c].y);
  }
#pragma unroll
  for (int c = 0; c < C; c++) {
    double qx = 0.0, qy = 0.0;
#pragma unroll
    for (int c2 = 0; c2 < C; c2++) {
      const double2 mm = m[c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
      qx = fma(mm.x, dv[c2].x, qx);
      qy = fma(mm.y, dv[c2].y, qy);
    }
    q[c].x = qx;
    q[c].y = qy;
  }
#pragma unroll
  for (int c = 0; c < C; c++) {
    rv[c].x = fma(-alpha, q[c].x, rv[c].x);
    rv[c].y = fma(-alpha, q[c].y, rv[c].y);
  }
}
// x <- x + alpha d (:298)
template <int C>
__device__ __forceinline__ void cg_block_x(double2 (&xv)[C], const double2 (&dv)[C], double alpha) {
#pragma unroll
  for (int c = 0; c < C; c++) {
    xv[c].x = fma(alpha, dv[c].x, xv[c].x);
    xv[c].y = fma(alpha, dv[c].y, xv[c].y);
  }
}
// the four sums of a pass: acc += {r.r, r.Mr, r.q, d.q}
template <int C>
__device__ __forceinline__ void cg_block_sums(const double2 (&m)[C * (C + 1) / 2], const double2 (&dv)[C],
                                              const double2 (&rv)[C], const double2 (&q)[C], double (&acc)[4]) {
#pragma unroll
  for (int c = 0; c < C; c++) {
    double mx = 0.0, my = 0.0;
#pragma unroll
    for (int c2 = 0; c2 < C; c2++) {
      const double2 mm = m[c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
      mx = fma(mm.x, rv[c2].x, mx);
      my = fma(mm.y, rv[c2].y, my);
    }
    acc[0] = fma(rv[c].y, rv[c].y, fma(rv[c].x, rv[c].x, acc[0]));
    acc[1] = fma(rv[c].y, my, fma(rv[c].x, mx, acc[1]));
    acc[2] = fma(rv[c].y, q[c].y, fma(rv[c].x, q[c].x, acc[2]));
    acc[3] = fma(dv[c].y, q[c].y, fma(dv[c].x, q[c].x, acc[3]));
  }
}

// K1 epilogue: when the last block of K1 can see every rank's sums (one rank, or NVLink mailboxes) it starts
// the solve's scalar state itself -- no gather kernel, no 1-thread init launch.
struct CgInit {
  CgScalars *st;      // nullptr: the host launches cg_init_scalars_kernel after its own gather
  int i_max, m;
  double converge;
  double *gathered;
};
__device__ __forceinline__ void cg_init_fold(const CgInit &ci, const PeerComm &pc, double *out, bool last);

// K1: one pass over sig/rms builds b2 = b + fluctuation, the blocks M, and the initial
// residual r = b2 - M x (warm start x, Q10), d = r; reduces {r.r, r.M r}.
// out[0] = sum r^2, out[1] = sum r.Mr
template <int C>
__global__ void __launch_bounds__(DG_THREADS)
rhs_blocks_kernel(const ModelView mv, const CgView<C> cg, double *partials, unsigned int *ticket,
                  double *out, const CgInit ci, const PeerComm pc) {
  constexpr int T = C * (C + 1) / 2;
  __shared__ double smem[2 * 32];
  double acc[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride) {
    const bool use = mv.mask[p] != 0;
    if (!use) {
      // masked pixels: zero rows/cols (:474-485, :695); x keeps its initial value
      for (int s = 0; s < cg.S; s++) {
        const size_t e = (size_t)s * mv.Ppad + p;
        for (int t = 0; t < T; t++) cg.M[(size_t)t * cg.S * mv.Ppad + e] = 0.0;
        for (int c = 0; c < C; c++) {
          cg.r[(size_t)c * cg.S * mv.Ppad + e] = 0.0;
          cg.d[(size_t)c * cg.S * mv.Ppad + e] = 0.0;
        }
      }
      continue;
    }
    // per-pixel parameters of the group components and of the subtracted components
    double th[2][C][DG_MAXIND];
    double oa[2][DG_MAX_COMPS], oth[2][DG_MAX_COMPS][DG_MAXIND];
    #pragma unroll
    for (int s = 0; s < 2; s++) {
      if (s >= cg.S) continue;
      const size_t kp = (size_t)cg.plane[s] * mv.Ppad + p;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const CompView &cv = mv.comp[cg.comp[c]];
        th[s][c][0] = cv.nind > 0 ? cv.idx[0][kp] : 0.0;
        th[s][c][1] = cv.nind > 1 ? cv.idx[1][kp] : 0.0;
      }
      for (int o = 0; o < cg.nog; o++) {
        const CompView &cv = mv.comp[cg.og[o]];
        oa[s][o] = cv.amp[kp];
        oth[s][o][0] = cv.nind > 0 ? cv.idx[0][kp] : 0.0;
        oth[s][o][1] = cv.nind > 1 ? cv.idx[1][kp] : 0.0;
      }
    }
    double eta[2] = {0.0, 0.0};
    if (cg.fluct)
#pragma unroll
      for (int s = 0; s < 2; s++)
        if (s < cg.S) eta[s] = cg.eta ? cg.eta[(size_t)s * mv.Ppad + p]
                        : philox_normal(cg.seed, DG_STREAM_ETA,
                                        (uint64_t)s * (uint64_t)mv.npix + (uint64_t)(mv.pix_lo + p));
    // Q and U share the SED whenever their index maps agree (always, after a Q+U draw)
    bool same = cg.S == 2;
    if (same) {
#pragma unroll
      for (int c = 0; c < C; c++)
        same = same && sed_uniform(mv, cg.comp[c], cg.plane[0]) == sed_uniform(mv, cg.comp[c], cg.plane[1]);
      for (int o = 0; o < cg.nog; o++)  // (a template's "SED" is template_amplitudes(band, plane): per plane)
        same = same && mv.comp[cg.og[o]].tamp == nullptr &&
               sed_uniform(mv, cg.og[o], cg.plane[0]) == sed_uniform(mv, cg.og[o], cg.plane[1]);
#pragma unroll
      for (int c = 0; c < C; c++)
        same = same && th[0][c][0] == th[1][c][0] && th[0][c][1] == th[1][c][1];
      for (int o = 0; o < cg.nog; o++)
        same = same && oth[0][o][0] == oth[1][o][0] && oth[0][o][1] == oth[1][o][1];
    }

    double b[2][C], f[2][C], M[2][T];
#pragma unroll
    for (int s = 0; s < 2; s++) {
#pragma unroll
      for (int c = 0; c < C; c++) b[s][c] = f[s][c] = 0.0;
#pragma unroll
      for (int t = 0; t < T; t++) M[s][t] = 0.0;
    }
    for (int j = 0; j < mv.nbands; j++) {
      double sed[2][C], osed[2][DG_MAX_COMPS];
#pragma unroll
      for (int c = 0; c < C; c++) sed[0][c] = sed_eval(mv, cg.comp[c], cg.plane[0], j, th[0][c][0], th[0][c][1]);
      for (int o = 0; o < cg.nog; o++) osed[0][o] = sed_eval(mv, cg.og[o], cg.plane[0], j, oth[0][o][0], oth[0][o][1]);
      if (cg.S == 2) {
        if (same) {
#pragma unroll
          for (int c = 0; c < C; c++) sed[1][c] = sed[0][c];
          for (int o = 0; o < cg.nog; o++) osed[1][o] = osed[0][o];
        } else {
#pragma unroll
          for (int c = 0; c < C; c++) sed[1][c] = sed_eval(mv, cg.comp[c], cg.plane[1], j, th[1][c][0], th[1][c][1]);
          for (int o = 0; o < cg.nog; o++) osed[1][o] = sed_eval(mv, cg.og[o], cg.plane[1], j, oth[1][o][0], oth[1][o][1]);
        }
      }
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (s >= cg.S) continue;
        const int k = cg.plane[s];
        const size_t off = plane_off(mv, j, k) + p;
        double data = ldg_stream(mv.sig + off);
        const double rms = ldg_stream(mv.rms + off);
        if (k == 0) data = data / mv.gain[j];  // :369-373 (no offset here, unlike update_sky_model)
        for (int o = 0; o < cg.nog; o++) data = data - oa[s][o] * osed[s][o];
        const double w = 1.0 / (rms * rms);
        const double tn = eta[s] / rms;  // N^{-1/2} eta, :1005-1017
#pragma unroll
        for (int c = 0; c < C; c++) {
          b[s][c] += data * sed[s][c] * w;       // :489-494
          f[s][c] += tn * sed[s][c];             // :1030-1042
#pragma unroll
          for (int c2 = c; c2 < C; c2++) M[s][tri<C>(c, c2)] += sed[s][c] * sed[s][c2] * w;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
      if (s >= cg.S) continue;
      const size_t e = (size_t)s * mv.Ppad + p;
      if (cg.fluct == 1) {
        b[s][0] += f[s][C - 1];  // Q1: last diffuse component's term lands in slot 1 only
      } else if (cg.fluct == 2) {
#pragma unroll
        for (int c = 0; c < C; c++) b[s][c] += f[s][c];
      }
      double xv[C], rv[C], Mr[C];
#pragma unroll
      for (int c = 0; c < C; c++) xv[c] = cg.x[(size_t)c * cg.S * mv.Ppad + e];
#pragma unroll
      for (int c = 0; c < C; c++) {
        double ax = 0.0;
#pragma unroll
        for (int c2 = 0; c2 < C; c2++) ax += M[s][c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)] * xv[c2];
        rv[c] = b[s][c] - ax;
      }
#pragma unroll
      for (int c = 0; c < C; c++) {
        double a = 0.0;
#pragma unroll
        for (int c2 = 0; c2 < C; c2++) a += M[s][c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)] * rv[c2];
        Mr[c] = a;
      }
#pragma unroll
      for (int t = 0; t < T; t++) cg.M[(size_t)t * cg.S * mv.Ppad + e] = M[s][t];
#pragma unroll
      for (int c = 0; c < C; c++) {
        cg.r[(size_t)c * cg.S * mv.Ppad + e] = rv[c];
        if (cg.store_d) cg.d[(size_t)c * cg.S * mv.Ppad + e] = rv[c];
        acc[0] += rv[c] * rv[c];
        acc[1] += rv[c] * Mr[c];
      }
    }
  }
  cg_init_fold(ci, pc, out, grid_reduce<2>(acc, smem, partials, ticket, out));
}

// ---------------------------------------------------------------- scalar control
// `sums` holds one row of 4 doubles per rank (rank order => identical bits on every rank).
// After K1: {r.r, r.Mr}.  d_1 = r_1 so d.q = r.Mr.
__device__ __forceinline__ void cg_init_update(CgScalars *st, const double *sums, int nranks,
                                               int i_max, double converge) {
  double rr = 0.0, rMr = 0.0;
  for (int g = 0; g < nranks; g++) {
    rr += sums[g * 4 + 0];
    rMr += sums[g * 4 + 1];
  }
  st->delta_new = rr;
  st->delta_old = rr;
  st->dq = rMr;
  st->alpha = rr / rMr;
  st->alpha_prev = 0.0;
  st->beta = 0.0;
  st->iter = 1;  // the reference's loop counter starts at 1 (:287, Q3)
  st->i_max = i_max;
  st->converge = converge;
  st->done = !(1 < i_max && rr > converge);
  st->trace[0] = rr;
  st->ckpt = 0;
  st->x_at = 0;
  st->gen = 0u;
  st->ah[1] = st->alpha;
  st->bh[1] = 0.0;
}

__device__ __forceinline__ void cg_init_fold(const CgInit &ci, const PeerComm &pc, double *out, bool last) {
  if (!ci.st || !last) return;  // `last`: warp 0 of the block that finished the grid reduction
  if (pc.nranks > 1) peer_exchange(pc, out, 4, ci.gathered);
  if (threadIdx.x == 0) {
    cg_init_update(ci.st, pc.nranks > 1 ? ci.gathered : out, pc.nranks, ci.i_max, ci.converge);
    ci.st->m = ci.m;
  }
}

// After a fused pass: {r'.r', r'.Mr', r'.q, d.q} with q = M d.
// delta' = r'.r'; beta' = delta'/delta; d' = r' + beta' d  =>  d'.Md' = r'.Mr' + 2 beta' r'.q + beta'^2 d.q
__device__ __forceinline__ void cg_fused_update(CgScalars *st, const double *sums, int nranks) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int g = 0; g < nranks; g++)
    for (int i = 0; i < 4; i++) s[i] += sums[g * 4 + i];
  const double delta_old = st->delta_new;
  const double delta_new = s[0];
  const double beta = delta_new / delta_old;
  // explicit roundings: every CG form (and every kernel this is inlined into) must produce the same bits
  const double dq = __dadd_rn(__dadd_rn(s[1], __dmul_rn(__dmul_rn(2.0, beta), s[2])), __dmul_rn(__dmul_rn(beta, beta), s[3]));
  st->delta_old = delta_old;
  st->delta_new = delta_new;
  st->beta = beta;
  st->dq = dq;
  st->alpha_prev = st->alpha;
  st->alpha = delta_new / dq;
  const int it = st->iter + 1;
  st->iter = it;
  if (it - 1 < 256) st->trace[it - 1] = delta_new;
  st->done = !(it < st->i_max && delta_new > st->converge);
  if (it < DG_CG_HIST) {
    st->ah[it] = st->alpha;
    st->bh[it] = beta;
  }
  if (st->m > 0 && (it - 1) - st->ckpt == st->m) st->ckpt = it - 1;  // pass it-1 stored its state
}

static __global__ void cg_init_scalars_kernel(CgScalars *st, const double *gathered, int nranks,
                                       int i_max, double converge, int m) {
  cg_init_update(st, gathered, nranks, i_max, converge);
  st->m = m;
}
static __global__ void cg_fused_scalars_kernel(CgScalars *st, const double *gathered, int nranks) {
  if (st->done) return;
  cg_fused_update(st, gathered, nranks);
}

// two-pass form: after the d.q pass
static __global__ void cg_dq_scalars_kernel(CgScalars *st, const double *gathered, int nranks) {
  if (st->done) return;
  double dq = 0.0;
  for (int g = 0; g < nranks; g++) dq += gathered[g * 4 + 0];
  st->dq = dq;
  st->alpha = st->delta_new / dq;  // :297
}
// two-pass form: after the update pass
static __global__ void cg_rr_scalars_kernel(CgScalars *st, const double *gathered, int nranks) {
  if (st->done) return;
  double rr = 0.0;
  for (int g = 0; g < nranks; g++) rr += gathered[g * 4 + 0];
  st->delta_old = st->delta_new;  // :302-304
  st->delta_new = rr;
  st->beta = rr / st->delta_old;
  const int it = st->iter + 1;
  st->iter = it;
  if (it - 1 < 256) st->trace[it - 1] = rr;
  st->done = !(it < st->i_max && rr > st->converge);
}

// ---------------------------------------------------------------- K2: fused CG pass
// One HBM pass per CG iteration over vec2 elements e of the flattened [S][Ppad] index:
//   d <- r + beta d (beta = 0 on the first pass: d = r was stored by K1)
//   q  = M d;  r' = r - alpha q;  sums {r'.r', r'.M r', r'.q, d.q}
//   x: on even passes  x <- (x + alpha_prev d_prev) + alpha d  -- the two updates the reference
//      performs in consecutive iterations (:298), in the same order and therefore bit-identical,
//      but with x read and written every other pass only; cg_x_fixup_kernel applies the pending
//      term if the solve ends on an odd pass.
// Compulsory traffic per element: T + 2C reads + 2C writes, plus C + C for x on even passes
// (11 / 15 doubles for C = 2) against the 15 of SURVEY 8d's bytes_cg_it.
// fold = 1 (single rank): the last block also advances the scalars, saving a launch.
template <int C>
__global__ void __launch_bounds__(DG_THREADS, DG_CG_BLOCKS_PER_SM)
cg_fused_pass_kernel(CgScalars *st, const double *__restrict__ M, double *__restrict__ x,
                     double *__restrict__ r, double *__restrict__ d, int64_t n2 /* vec2 elements */,
                     double *partials, unsigned int *ticket, double *out, int fold, PeerComm pc,
                     double *gathered) {
  constexpr int T = C * (C + 1) / 2;
  if (st->done) return;
  const double alpha = st->alpha, beta = st->beta, alpha_prev = st->alpha_prev;
  const bool with_x = (st->iter & 1) == 0;  // passes are numbered by the loop counter, from 1
  __shared__ double smem[4 * 32];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)n2 * 2;  // doubles per component / block-entry plane-set
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
    double2 m[T], xv[C], rv[C], dv[C], dp[C];
#pragma unroll
    for (int t = 0; t < T; t++) m[t] = ldg_stream2(M + t * vs + 2 * e);
#pragma unroll
    for (int c = 0; c < C; c++) {
      rv[c] = *reinterpret_cast<const double2 *>(r + c * vs + 2 * e);
      dp[c] = *reinterpret_cast<const double2 *>(d + c * vs + 2 * e);
      if (with_x) xv[c] = *reinterpret_cast<const double2 *>(x + c * vs + 2 * e);
    }
#pragma unroll
    for (int c = 0; c < C; c++) dv[c] = dp[c];
    double2 q[C];
    cg_block_step<C>(m, dv, rv, q, alpha, beta);  // :305 (of the previous iteration), :296, :300
    if (with_x) {                                 // :298, twice
      cg_block_x<C>(xv, dp, alpha_prev);
      cg_block_x<C>(xv, dv, alpha);
    }
    cg_block_sums<C>(m, dv, rv, q, acc);
#pragma unroll
    for (int c = 0; c < C; c++) {
      *reinterpret_cast<double2 *>(d + c * vs + 2 * e) = dv[c];
      if (with_x) *reinterpret_cast<double2 *>(x + c * vs + 2 * e) = xv[c];
      *reinterpret_cast<double2 *>(r + c * vs + 2 * e) = rv[c];
    }
  }
  const bool last = grid_reduce<4>(acc, smem, partials, ticket, out);
  if (fold && last) {  // warp 0 of the last block: exchange over NVLink (if any), advance the scalars
    if (pc.nranks > 1) peer_exchange(pc, out, 4, gathered);
    if (threadIdx.x == 0) cg_fused_update(st, pc.nranks > 1 ? gathered : out, pc.nranks);
  }
}

// ---------------------------------------------------------------- K2r: recompute form
// The CG state of a block depends on the rest of the sky only through the scalars (alpha_i,
// beta_i), so a pass does not have to STORE r and d: given the state after some earlier pass c
// ("checkpoint") it re-runs the block-local recurrences for passes c+1..k in registers -- the
// same operations in the same order, hence the same bits -- and only every m-th pass writes
// r, d and x back.  HBM traffic per element and pass drops from T+4C / T+6C doubles to
// T+2C reads, plus 2C (x) reads and 4C... writes on checkpoint passes: ~8 instead of 13 doubles
// on average for C = 2, m = 8, paid for with ~m/2 extra 2x2 mat-vecs per element on an FP64 pipe
// that the streaming form leaves idle.
//   stored (r, d) = (r_{c+1}, d_c);   pass i:  d_i = r_i + beta_i d_{i-1};  q = M d_i;
//   x += alpha_i d_i;  r_{i+1} = r_i - alpha_i q          (cg_search :296-305)
// final = 1: the solve is over; bring x up to date from the last checkpoint (no reduction).
template <int C, bool UNPACK>
__global__ void __launch_bounds__(DG_THREADS, DG_CG_BLOCKS_PER_SM)
cg_recompute_pass_kernel(CgScalars *st, const double *__restrict__ M, double *__restrict__ x,
                         double *__restrict__ r, double *__restrict__ d, int64_t n2,
                         double *partials, unsigned int *ticket, double *out, int fold, int final,
                         PeerComm pc, double *gathered, CgAmpOut<C> ao) {
  constexpr int T = C * (C + 1) / 2;
  if (st->done && !final) return;
  const int c0 = st->ckpt;
  const int k = final ? st->iter - 1 : st->iter;  // last pass to (re)run
  const int nstep = k - c0;
  if (final && nstep <= 0) return;
  const bool store = !final && nstep == st->m;     // checkpoint pass
  const bool with_x = store || final;
  __shared__ double smem[4 * 32];
  __shared__ double sa[DG_CG_MAXM + 1], sb[DG_CG_MAXM + 1];
  if (threadIdx.x < nstep) {
    sa[threadIdx.x] = st->ah[c0 + 1 + threadIdx.x];
    sb[threadIdx.x] = st->bh[c0 + 1 + threadIdx.x];
  }
  __syncthreads();
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)n2 * 2;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
    double2 m[T], xv[C], rv[C], dv[C], q[C];
#pragma unroll
    for (int t = 0; t < T; t++) m[t] = ldg_stream2(M + t * vs + 2 * e);
#pragma unroll
    for (int c = 0; c < C; c++) {
      rv[c] = *reinterpret_cast<const double2 *>(r + c * vs + 2 * e);
      // before the first checkpoint the stored direction is d_0 = r_1 itself (beta_1 = 0)
      dv[c] = c0 == 0 ? rv[c] : *reinterpret_cast<const double2 *>(d + c * vs + 2 * e);
      if (with_x) xv[c] = *reinterpret_cast<const double2 *>(x + c * vs + 2 * e);
    }
    for (int i = 0; i < nstep; i++) {
      const double alpha = sa[i], beta = sb[i];
      cg_block_step<C>(m, dv, rv, q, alpha, beta);  // :305, :296, :300
      if (with_x) cg_block_x<C>(xv, dv, alpha);     // :298
    }
    if (!final) cg_block_sums<C>(m, dv, rv, q, acc);
#pragma unroll
    for (int c = 0; c < C; c++) {
      if (store) {
        *reinterpret_cast<double2 *>(r + c * vs + 2 * e) = rv[c];
        *reinterpret_cast<double2 *>(d + c * vs + 2 * e) = dv[c];
      }
      if (with_x) *reinterpret_cast<double2 *>(x + c * vs + 2 * e) = xv[c];
      // unpack_amplitudes (:1327-1335) folded into the last pass: x -> c%amplitude planes
      if (UNPACK && final && ao.p[c]) *reinterpret_cast<double2 *>(ao.p[c] + 2 * e) = xv[c];
    }
  }
  if (final) return;
  const bool last = grid_reduce<4>(acc, smem, partials, ticket, out);
  if (fold && last) {  // warp 0 of the last block: exchange over NVLink (if any), advance the scalars
    if (pc.nranks > 1) peer_exchange(pc, out, 4, gathered);
    if (threadIdx.x == 0) cg_fused_update(st, pc.nranks > 1 ? gathered : out, pc.nranks);
  }
}

// pending x update when the solve stopped after an odd number of passes
static __global__ void __launch_bounds__(DG_THREADS)
cg_x_fixup_kernel(const CgScalars *st, double *__restrict__ x, const double *__restrict__ d, int64_t n) {
  if (((st->iter - 1) & 1) == 0) return;
  const double a = st->alpha_prev;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n / 2; e += stride) {
    double2 xv = *reinterpret_cast<const double2 *>(x + 2 * e);
    const double2 dv = *reinterpret_cast<const double2 *>(d + 2 * e);
    xv.x = fma(a, dv.x, xv.x);
    xv.y = fma(a, dv.y, xv.y);
    *reinterpret_cast<double2 *>(x + 2 * e) = xv;
  }
}

// ---------------------------------------------------------------- classic two-pass form
// pass A: d <- r + beta d (skipped on the first iteration), sum d.(M d)      (:296-297)
template <int C>
__global__ void __launch_bounds__(DG_THREADS, DG_CG_BLOCKS_PER_SM)
cg_dq_pass_kernel(const CgScalars *st, const double *__restrict__ M, const double *__restrict__ r,
                  double *__restrict__ d, int64_t n2, double *partials, unsigned int *ticket,
                  double *out) {
  constexpr int T = C * (C + 1) / 2;
  if (st->done) return;
  const double beta = st->beta;
  const bool first = st->iter == 1;
  __shared__ double smem[4 * 32];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)n2 * 2;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
    double2 m[T], rv[C], dv[C];
#pragma unroll
    for (int t = 0; t < T; t++) m[t] = ldg_stream2(M + t * vs + 2 * e);
#pragma unroll
    for (int c = 0; c < C; c++) {
      dv[c] = *reinterpret_cast<const double2 *>(d + c * vs + 2 * e);
      if (!first) {
        rv[c] = *reinterpret_cast<const double2 *>(r + c * vs + 2 * e);
        dv[c].x = rv[c].x + beta * dv[c].x;
        dv[c].y = rv[c].y + beta * dv[c].y;
        *reinterpret_cast<double2 *>(d + c * vs + 2 * e) = dv[c];
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      double qx = 0.0, qy = 0.0;
#pragma unroll
      for (int c2 = 0; c2 < C; c2++) {
        const double2 mm = m[c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
        qx += mm.x * dv[c2].x;
        qy += mm.y * dv[c2].y;
      }
      acc[0] += dv[c].x * qx + dv[c].y * qy;
    }
  }
  grid_reduce<4>(acc, smem, partials, ticket, out);
}

// pass B: x += alpha d; r -= alpha (M d); sum r.r                            (:298-303)
template <int C>
__global__ void __launch_bounds__(DG_THREADS, DG_CG_BLOCKS_PER_SM)
cg_update_pass_kernel(const CgScalars *st, const double *__restrict__ M, double *__restrict__ x,
                      double *__restrict__ r, const double *__restrict__ d, int64_t n2,
                      double *partials, unsigned int *ticket, double *out) {
  constexpr int T = C * (C + 1) / 2;
  if (st->done) return;
  const double alpha = st->alpha;
  __shared__ double smem[4 * 32];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)n2 * 2;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
    double2 m[T], xv[C], rv[C], dv[C];
#pragma unroll
    for (int t = 0; t < T; t++) m[t] = ldg_stream2(M + t * vs + 2 * e);
#pragma unroll
    for (int c = 0; c < C; c++) {
      rv[c] = *reinterpret_cast<const double2 *>(r + c * vs + 2 * e);
      dv[c] = *reinterpret_cast<const double2 *>(d + c * vs + 2 * e);
      xv[c] = *reinterpret_cast<const double2 *>(x + c * vs + 2 * e);
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      double qx = 0.0, qy = 0.0;
#pragma unroll
      for (int c2 = 0; c2 < C; c2++) {
        const double2 mm = m[c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
        qx += mm.x * dv[c2].x;
        qy += mm.y * dv[c2].y;
      }
      xv[c].x = xv[c].x + alpha * dv[c].x;
      xv[c].y = xv[c].y + alpha * dv[c].y;
      rv[c].x = rv[c].x - alpha * qx;
      rv[c].y = rv[c].y - alpha * qy;
      acc[0] += rv[c].x * rv[c].x + rv[c].y * rv[c].y;
      *reinterpret_cast<double2 *>(x + c * vs + 2 * e) = xv[c];
      *reinterpret_cast<double2 *>(r + c * vs + 2 * e) = rv[c];
    }
  }
  grid_reduce<4>(acc, smem, partials, ticket, out);
}

// initialize_x / unpack_amplitudes, src/dang_cg_mod.f90:1173-1282, 1284-1396: plane copies
static __global__ void copy_planes_kernel(double *dst, const double *src, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
