// kernels_tmpl.cuh -- amplitude draw for a CG group that holds "border" components: `template`, `monopole` and
// `hi_fit`.  compute_rhs :326-596 (border rows :522-587, the extra subtraction :444-460),
// compute_Ax :598-911 (border columns / rows :717-768, :833-893), compute_sample_vector :913-1100
// (:1044-1096), cg_search :179-324, src/dang_cg_mod.f90.
//
// A border component contributes  x_b(band) * t_b(pix, band)  to the bands it is fitted to (corr(band)):
//     template   t = template(pix, plane)                       on the planes of the solve
//     monopole   t = template(pix, 1) = 1                       Stokes I only
//     hi_fit     t = template(pix, 1) * B_nu(T_d(pix)) in RJ    Stokes I only
// so the solution vector is [diffuse amplitudes per (pixel, Stokes) ..., one scalar per (border component,
// fitted band)] and A = sum_nu T^t N^-1 T is block diagonal plus `nt` dense border rows / columns.  The apply is
// matrix-free, band by band as the reference writes it (T x, / sigma^2, T^t), one thread per pixel; the border rows
// are masked, noise-weighted sums over all pixels (deterministic grid reduction + rank-ordered gather), the `nt`
// tail entries of every CG vector live in TmplScalars and are updated by one-thread kernels.  Each CG iteration is
// two map passes (apply with the direction update folded in; x / r update) -- this path is about coverage of the
// reference's template / monopole / HI fits, the block-diagonal kernels of kernels_cg.cuh remain the fast path.
//
// Tail layout: compute_Ax / compute_rhs / initialize_x lay the border components out one after the other in
// component_list order (`slot_ax`); compute_sample_vector uses ONE running counter over (band, component) pairs
// (:970, never reset: SURVEY Q8), so with several border components its fluctuation terms land in interleaved slots
// (`slot_sv`).  Both maps are built on the host; with one border component they coincide.
#pragma once
#include "common.cuh"

#define DG_TMPL_MAX 32   // tail length: fitted bands summed over the group's border components
#define DG_TMPL_CMAX 2   // diffuse components next to them in the group
#define DG_TMPL_BMAX 3   // border components in one group

enum { DG_BORDER_TEMPLATE = 0, DG_BORDER_MONOPOLE = 1, DG_BORDER_HI_FIT = 2 };

struct TmplView {
  int C;                       // diffuse components in the group with sample_amplitude
  int comp[DG_TMPL_CMAX];
  int nb;                      // border components
  int bcomp[DG_TMPL_BMAX];     // ModelView::comp index; amp = template map
  int bkind[DG_TMPL_BMAX];
  int S, plane[2];
  int nog, og[DG_MAX_COMPS];   // components subtracted from the data (:427-443)
  int nt;                      // tail length
  int slot_ax[DG_TMPL_BMAX][DG_MAX_BANDS];  // tail slot of (border, band) in compute_Ax's layout, -1 if not fitted
  int slot_sv[DG_TMPL_BMAX][DG_MAX_BANDS];  // ... in compute_sample_vector's layout (Q8)
  int fluct;                   // 0 none, 1 reference indexing (Q1), 2 per component
  const double *eta;           // [S][Ppad] or nullptr -> Philox
  uint64_t seed;
  double *b, *x, *r, *d, *q;   // diffuse planes [C][S][Ppad]
};

struct TmplScalars {
  double xt[DG_TMPL_MAX], bt[DG_TMPL_MAX], rt[DG_TMPL_MAX], dt[DG_TMPL_MAX], qt[DG_TMPL_MAX];
  double delta_new, delta_old, alpha, beta, dq, converge;
  int iter, i_max, done, pad;
  double trace[256];
};

#define DG_TMPL_NV (DG_TMPL_MAX + 2)

// t_b(pix, band) on plane k of the solve: eval_sed of the border component (src/dang_component_mod.f90:803-808)
__device__ __forceinline__ double border_factor(const ModelView &mv, const TmplView &tv, int b, int k, int64_t p, int j) {
  const CompView &bc = mv.comp[tv.bcomp[b]];
  if (tv.bkind[b] == DG_BORDER_TEMPLATE) return bc.amp[(size_t)k * mv.Ppad + p];
  const double t = bc.amp[p];  // template(pix, 1): hi_fit and monopole rows always read plane 1 (:722, :736)
  if (tv.bkind[b] == DG_BORDER_MONOPOLE) return t;
  return t * sed_planck_rj(mv, j, bc.idx[0][p]);
}

// compute_rhs (+ compute_sample_vector): b planes and the tail sums.
// out[0..nt) = b_t (+ fluctuation), out[DG_TMPL_MAX] unused
static __global__ void __launch_bounds__(DG_THREADS)
tmpl_rhs_kernel(const ModelView mv, const TmplView tv, double *partials, unsigned int *ticket, double *out) {
  __shared__ double smem[DG_TMPL_NV * 32];
  double acc[DG_TMPL_NV];
#pragma unroll
  for (int i = 0; i < DG_TMPL_NV; i++) acc[i] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride) {
    const bool use = mv.mask[p] != 0;
    for (int s = 0; s < tv.S; s++) {
      const int k = tv.plane[s];
      const size_t kp = (size_t)k * mv.Ppad + p, e = (size_t)s * mv.Ppad + p;
      double b[DG_TMPL_CMAX] = {0.0, 0.0}, f[DG_TMPL_CMAX] = {0.0, 0.0};
      if (use) {
        double eta = 0.0;
        if (tv.fluct)
          eta = tv.eta ? tv.eta[e]
                       : philox_normal(tv.seed, DG_STREAM_ETA, (uint64_t)s * (uint64_t)mv.npix + (uint64_t)(mv.pix_lo + p));
        for (int j = 0; j < mv.nbands; j++) {
          const size_t off = plane_off(mv, j, k) + p;
          double data = ldg_stream(mv.sig + off);
          const double rms = ldg_stream(mv.rms + off);
          if (k == 0) data = data / mv.gain[j];  // :369-373
          for (int o = 0; o < tv.nog; o++) {     // :427-443
            const CompView &cc = mv.comp[tv.og[o]];
            const double t0 = cc.nind > 0 ? cc.idx[0][kp] : 0.0, t1 = cc.nind > 1 ? cc.idx[1][kp] : 0.0;
            data = data - cc.amp[kp] * sed_eval(mv, tv.og[o], k, j, t0, t1);
          }
          // :444-460: a template / monopole is also removed from the bands it is NOT fitted to (eval_signal =
          // template_amplitudes(band, plane) * template(pix, plane); hi_fit is not part of that block)
          for (int bb = 0; bb < tv.nb; bb++)
            if (tv.slot_ax[bb][j] < 0 && tv.bkind[bb] != DG_BORDER_HI_FIT)
              data = data - mv.tab->sed[tv.bcomp[bb] * 3 + k][j] * mv.comp[tv.bcomp[bb]].amp[kp];
          const double tn = eta / rms;  // :1005-1017
          for (int c = 0; c < tv.C; c++) {
            const CompView &cc = mv.comp[tv.comp[c]];
            const double t0 = cc.nind > 0 ? cc.idx[0][kp] : 0.0, t1 = cc.nind > 1 ? cc.idx[1][kp] : 0.0;
            const double sed = sed_eval(mv, tv.comp[c], k, j, t0, t1);
            b[c] = b[c] + (data * sed) / (rms * rms);  // :489-494
            f[c] += tn * sed;                          // :1030-1042
          }
          for (int bb = 0; bb < tv.nb; bb++) {
            const int slot = tv.slot_ax[bb][j];
            if (slot < 0) continue;
            const double t = border_factor(mv, tv, bb, k, p, j);
            acc[slot] += data / (rms * rms) * t;                 // :530, :548, :568-574
            if (tv.fluct) acc[tv.slot_sv[bb][j]] += tn * t;      // :1049, :1061 (t = 1), :1083-1094
          }
        }
        if (tv.C > 0) {
          if (tv.fluct == 1) b[0] += f[tv.C - 1];      // Q1
          else if (tv.fluct == 2)
            for (int c = 0; c < tv.C; c++) b[c] += f[c];
        }
      }
      for (int c = 0; c < tv.C; c++) tv.b[(size_t)c * tv.S * mv.Ppad + e] = b[c];  // masked: 0 (:474-485)
    }
  }
  grid_reduce<DG_TMPL_NV>(acc, smem, partials, ticket, out);
}

// compute_Ax on v = (diffuse planes, tail vt): q planes, out[0..nt) = border-row sums, out[DG_TMPL_MAX] = the
// diffuse part of v.q.  mode 1: v is the new search direction d = r + beta d (stored), :305.
static __global__ void __launch_bounds__(DG_THREADS)
tmpl_apply_kernel(const ModelView mv, const TmplView tv, const TmplScalars *sc, int mode, double *partials,
                  unsigned int *ticket, double *out) {
  __shared__ double smem[DG_TMPL_NV * 32];
  __shared__ double vt[DG_TMPL_MAX];
  if (threadIdx.x < DG_TMPL_MAX) vt[threadIdx.x] = mode ? sc->dt[threadIdx.x] : sc->xt[threadIdx.x];
  __syncthreads();
  const double beta = sc->beta;
  double acc[DG_TMPL_NV];
#pragma unroll
  for (int i = 0; i < DG_TMPL_NV; i++) acc[i] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (sc->done && mode) return;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride) {
    const bool use = mv.mask[p] != 0;
    for (int s = 0; s < tv.S; s++) {
      const int k = tv.plane[s];
      const size_t kp = (size_t)k * mv.Ppad + p, e = (size_t)s * mv.Ppad + p;
      double v[DG_TMPL_CMAX] = {0.0, 0.0}, q[DG_TMPL_CMAX] = {0.0, 0.0};
      for (int c = 0; c < tv.C; c++) {
        const size_t ce = (size_t)c * tv.S * mv.Ppad + e;
        if (mode) {
          v[c] = tv.r[ce] + beta * tv.d[ce];
          tv.d[ce] = v[c];
        } else {
          v[c] = tv.x[ce];
        }
      }
      if (use) {  // masked pixels contribute nothing (:695)
        for (int j = 0; j < mv.nbands; j++) {
          const double rms = ldg_stream(mv.rms + plane_off(mv, j, k) + p);
          double sed[DG_TMPL_CMAX] = {0.0, 0.0}, tb[DG_TMPL_BMAX] = {0.0, 0.0, 0.0};
          double temp1 = 0.0;
          for (int c = 0; c < tv.C; c++) {
            const CompView &cc = mv.comp[tv.comp[c]];
            const double t0 = cc.nind > 0 ? cc.idx[0][kp] : 0.0, t1 = cc.nind > 1 ? cc.idx[1][kp] : 0.0;
            sed[c] = sed_eval(mv, tv.comp[c], k, j, t0, t1);
            temp1 = temp1 + v[c] * sed[c];             // :697-704
          }
          for (int bb = 0; bb < tv.nb; bb++) {
            const int slot = tv.slot_ax[bb][j];
            if (slot < 0) continue;
            tb[bb] = border_factor(mv, tv, bb, k, p, j);
            temp1 = temp1 + vt[slot] * tb[bb];          // :722, :736, :750-751
          }
          temp1 = temp1 / (rms * rms);                  // :775-791
          for (int c = 0; c < tv.C; c++) q[c] += temp1 * sed[c];  // :813-820, :904
          for (int bb = 0; bb < tv.nb; bb++) {
            const int slot = tv.slot_ax[bb][j];
            if (slot >= 0) acc[slot] += temp1 * tb[bb];  // :840, :856 (t = 1), :872-887
          }
        }
      }
      for (int c = 0; c < tv.C; c++) {
        tv.q[(size_t)c * tv.S * mv.Ppad + e] = q[c];
        acc[DG_TMPL_MAX] += v[c] * q[c];
      }
    }
  }
  grid_reduce<DG_TMPL_NV>(acc, smem, partials, ticket, out);
}

// r = b - q (first residual, :283); out[0] = sum r^2 over the diffuse planes
static __global__ void __launch_bounds__(DG_THREADS)
tmpl_resid_kernel(const TmplView tv, int64_t n, double *partials, unsigned int *ticket, double *out) {
  __shared__ double smem[2 * 32];
  double acc[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double r = tv.b[i] - tv.q[i];
    tv.r[i] = r;
    tv.d[i] = 0.0;
    acc[0] += r * r;
  }
  grid_reduce<2>(acc, smem, partials, ticket, out);
}

// x += alpha d, r -= alpha q (:298-299); out[0] = sum r^2 over the diffuse planes
static __global__ void __launch_bounds__(DG_THREADS)
tmpl_update_kernel(const TmplView tv, const TmplScalars *sc, int64_t n, double *partials, unsigned int *ticket,
                   double *out) {
  __shared__ double smem[2 * 32];
  double acc[2] = {0.0, 0.0};
  if (sc->done) return;
  const double alpha = sc->alpha;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    tv.x[i] = tv.x[i] + alpha * tv.d[i];
    const double r = tv.r[i] - alpha * tv.q[i];
    tv.r[i] = r;
    acc[0] += r * r;
  }
  grid_reduce<2>(acc, smem, partials, ticket, out);
}

// ---- one-thread control kernels; `g` holds one row per rank (rank order => identical bits everywhere)
static __global__ void tmpl_s_rhs_kernel(TmplScalars *sc, const double *g, int nranks, int cnt, int nt, int i_max,
                                         double converge) {
  for (int l = 0; l < DG_TMPL_MAX; l++) {
    double s = 0.0;
    for (int r = 0; r < nranks && l < nt; r++) s += g[r * cnt + l];
    sc->bt[l] = s;
    sc->rt[l] = sc->dt[l] = sc->qt[l] = 0.0;
    if (l >= nt) sc->xt[l] = 0.0;
  }
  sc->i_max = i_max;
  sc->converge = converge;
  sc->alpha = sc->beta = 0.0;
  sc->done = 0;
  sc->iter = 1;
}
// after apply(x0): q_t, r_t = b_t - q_t
static __global__ void tmpl_s_resid_a_kernel(TmplScalars *sc, const double *g, int nranks, int cnt, int nt) {
  for (int l = 0; l < nt; l++) {
    double s = 0.0;
    for (int r = 0; r < nranks; r++) s += g[r * cnt + l];
    sc->qt[l] = s;
    sc->rt[l] = sc->bt[l] - s;
    sc->dt[l] = 0.0;
  }
}
// after the residual kernel: delta = sum r^2 (diffuse + tail); loop counter starts at 1 (Q3)
static __global__ void tmpl_s_resid_b_kernel(TmplScalars *sc, const double *g, int nranks, int cnt, int nt) {
  double rr = 0.0;
  for (int r = 0; r < nranks; r++) rr += g[r * cnt];
  for (int l = 0; l < nt; l++) rr += sc->rt[l] * sc->rt[l];
  sc->delta_new = rr;
  sc->delta_old = rr;
  sc->beta = 0.0;  // the first direction is d = r
  sc->trace[0] = rr;
  for (int l = 0; l < nt; l++) sc->dt[l] = sc->rt[l];
  sc->done = !(sc->iter < sc->i_max && rr > sc->converge);
}
// after apply(d): q_t, alpha = delta / d.q, tail of x and r (:297-299)
static __global__ void tmpl_s_apply_kernel(TmplScalars *sc, const double *g, int nranks, int cnt, int nt) {
  if (sc->done) return;
  double dq = 0.0;
  for (int r = 0; r < nranks; r++) dq += g[r * cnt + DG_TMPL_MAX];
  for (int l = 0; l < nt; l++) {
    double s = 0.0;
    for (int r = 0; r < nranks; r++) s += g[r * cnt + l];
    sc->qt[l] = s;
    dq += sc->dt[l] * s;
  }
  sc->dq = dq;
  sc->alpha = sc->delta_new / dq;
  for (int l = 0; l < nt; l++) {
    sc->xt[l] = sc->xt[l] + sc->alpha * sc->dt[l];
    sc->rt[l] = sc->rt[l] - sc->alpha * sc->qt[l];
  }
}
// after the update kernel: delta, beta, tail of d (:300-305), loop test (:293)
static __global__ void tmpl_s_update_kernel(TmplScalars *sc, const double *g, int nranks, int cnt, int nt) {
  if (sc->done) return;
  double rr = 0.0;
  for (int r = 0; r < nranks; r++) rr += g[r * cnt];
  for (int l = 0; l < nt; l++) rr += sc->rt[l] * sc->rt[l];
  sc->delta_old = sc->delta_new;
  sc->delta_new = rr;
  sc->beta = sc->delta_new / sc->delta_old;
  for (int l = 0; l < nt; l++) sc->dt[l] = sc->rt[l] + sc->beta * sc->dt[l];
  if (sc->iter < 256) sc->trace[sc->iter] = rr;
  sc->iter = sc->iter + 1;
  sc->done = !(sc->iter < sc->i_max && rr > sc->converge);
}
