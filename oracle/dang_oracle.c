/*
 * dang_oracle.c -- CPU restatement of hermda02/dang's Gibbs hot path (see dang_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no reference fixtures exist, reference
 * cannot be built here).  Every function cites the reference file:line it follows.
 * Loop nests, operation order and quirks are kept; the
 * dead "T+Q+U" branches (iand(flag,0), Q2) and the monopole / hi_fit / T_cmb components
 * are not restated; the `template` type is (Q+U fits, one template per CG group).
 *
 * OpenMP pragmas sit on the same pixel loops as the reference's !$OMP PARALLEL DO and
 * are only active when built with -fopenmp (the bench CPU baseline); tests build
 * without it, i.e. with the serial semantics that define parity (Q7).  Where the
 * reference has a data race (lnL accumulation, src/dang_lnl_mod.f90:168-180) the
 * OpenMP build uses a reduction instead.
 */
#include "dang_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* src/dang_util_mod.f90:12-15; pi from healpix_types */
static const double ORA_PI = 3.141592653589793238462643383279502884197;
static const double ORA_KB = 1.3806503e-23;
#define ORA_H (1.0545726691251021e-34 * 2.0 * ORA_PI)
static const double ORA_C = 2.99792458e8;
static double ORA_TCMB = 2.7255; /* module-global T_CMB: a 'T_cmb' component overwrites it (dang_sample_mod.f90:76-78) */

typedef struct {
  int n;        /* 0 <=> bp%id == 'delta' */
  double nu_c;  /* Hz */
  double *nu0;  /* Hz */
  double *tau0; /* unit sum */
} ora_band;

typedef struct {
  int type;
  char label[17];
  double nu_ref;
  int cg_group;
  int sample_amplitude;
  int nindices;
  double *amplitude; /* [nmaps][npix] */
  double *indices;   /* [nindices][nmaps][npix] */
  int sample_index[ORA_MAXIND];
  int index_mode[ORA_MAXIND];
  int lnl_type[ORA_MAXIND];
  int prior_type[ORA_MAXIND];
  double gauss_prior[ORA_MAXIND][2];
  double uni_prior[ORA_MAXIND][2];
  double step_size[ORA_MAXIND];
  int tuned[ORA_MAXIND];
  int sample_nside[ORA_MAXIND];
  int nflag[ORA_MAXIND];
  int pol_flag[ORA_MAXIND][3];
  /* type 'template' (src/dang_component_mod.f90:536-577): template map (already divided by its
   * maximum, :574-577), one amplitude per (band, plane), which bands are fitted, how many */
  double *template_map;        /* [nmaps][npix] */
  double *template_amplitudes; /* [nmaps][nbands] */
  int *corr;                   /* [nbands] */
  int nfit;
} ora_comp;

struct ora_state {
  int nside, npix, nmaps, nbands, ncomp;
  ora_band *bp;
  ora_comp *comp;
  double *sig_map, *rms_map, *res_map, *sky_model; /* [nbands][nmaps][npix] */
  double *chi_map;                                 /* [nmaps][npix] */
  double *masks;                                   /* [npix] plane 1 */
  double *gain, *offset;
  int pol_lo, pol_hi;
  double chisq;
  long nump;
};

struct ora_cg {
  ora_state *st;
  int cg_group, i_max, nflag;
  double converge;
  int pol_flag[3];
  double *x[3]; /* Q10: allocated and seeded on first use, then warm-started */
};

#define IDX3(st, pix, k, j) (((size_t)(j) * (st)->nmaps + (size_t)(k)) * (size_t)(st)->npix + (size_t)(pix))
#define IDX2(st, pix, k) ((size_t)(k) * (size_t)(st)->npix + (size_t)(pix))

static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) {
    fprintf(stderr, "dang_oracle: out of memory (%zu x %zu)\n", n, sz);
    abort();
  }
  return p;
}

int ora_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Thread count of the OpenMP build (bench.py sets it explicitly: torch.distributed.run exports
 * OMP_NUM_THREADS=1 to its workers, which would silently turn the CPU baseline into a 1-core run). */
void ora_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ construction */

ora_state *ora_create(int nside, int npix, int nmaps, int nbands, int ncomp) {
  ora_state *st = xcalloc(1, sizeof *st);
  st->nside = nside;
  st->npix = npix;
  st->nmaps = nmaps;
  st->nbands = nbands;
  st->ncomp = ncomp;
  st->bp = xcalloc(nbands, sizeof(ora_band));
  st->comp = xcalloc(ncomp, sizeof(ora_comp));
  size_t n3 = (size_t)npix * nmaps * nbands;
  st->sig_map = xcalloc(n3, sizeof(double));
  st->rms_map = xcalloc(n3, sizeof(double));
  st->res_map = xcalloc(n3, sizeof(double));
  st->sky_model = xcalloc(n3, sizeof(double));
  st->chi_map = xcalloc((size_t)npix * nmaps, sizeof(double));
  st->masks = xcalloc(npix, sizeof(double));
  st->gain = xcalloc(nbands, sizeof(double));
  st->offset = xcalloc(nbands, sizeof(double));
  for (int j = 0; j < nbands; j++) st->gain[j] = 1.0; /* dang_data_mod.f90:127-128 */
  st->pol_lo = 2;
  st->pol_hi = 3;
  return st;
}

void ora_destroy(ora_state *st) {
  if (!st) return;
  for (int j = 0; j < st->nbands; j++) {
    free(st->bp[j].nu0);
    free(st->bp[j].tau0);
  }
  for (int i = 0; i < st->ncomp; i++) {
    free(st->comp[i].amplitude);
    free(st->comp[i].indices);
    free(st->comp[i].template_map);
    free(st->comp[i].template_amplitudes);
    free(st->comp[i].corr);
  }
  free(st->bp);
  free(st->comp);
  free(st->sig_map);
  free(st->rms_map);
  free(st->res_map);
  free(st->sky_model);
  free(st->chi_map);
  free(st->masks);
  free(st->gain);
  free(st->offset);
  free(st);
}

int ora_set_band(ora_state *st, int band, double nu_c, int n, const double *nu0_ghz,
                 const double *tau) {
  if (band < 0 || band >= st->nbands) return 1;
  ora_band *b = &st->bp[band];
  b->nu_c = nu_c;
  if (b->nu_c < 1e9) b->nu_c = b->nu_c * 1e9; /* dang_bp_mod.f90:35-37 */
  free(b->nu0);
  free(b->tau0);
  b->nu0 = b->tau0 = NULL;
  b->n = n;
  if (n > 0) {
    b->nu0 = xcalloc(n, sizeof(double));
    b->tau0 = xcalloc(n, sizeof(double));
    double total = 0.0; /* normalize_bandpass, dang_bp_mod.f90:76-79 */
    for (int i = 0; i < n; i++) total = total + tau[i];
    for (int i = 0; i < n; i++) {
      b->nu0[i] = nu0_ghz[i] * 1.e9; /* read_bandpass :138 */
      b->tau0[i] = tau[i] / total;
    }
  }
  return 0;
}

int ora_set_maps(ora_state *st, const double *sig, const double *rms, const double *mask,
                 const double *gain, const double *offset) {
  size_t n3 = (size_t)st->npix * st->nmaps * st->nbands;
  memcpy(st->sig_map, sig, n3 * sizeof(double));
  memcpy(st->rms_map, rms, n3 * sizeof(double));
  st->nump = 0;
  for (int i = 0; i < st->npix; i++) { /* dang_data_mod.f90:153-161 with Q9 convention */
    if (mask[i] == 0.0 || mask[i] == ORA_MISSVAL) {
      st->masks[i] = 0.0;
    } else {
      st->masks[i] = mask[i];
      st->nump += st->nmaps;
    }
  }
  if (gain) memcpy(st->gain, gain, st->nbands * sizeof(double));
  if (offset) memcpy(st->offset, offset, st->nbands * sizeof(double));
  return 0;
}

void ora_set_pol_type(ora_state *st, int lo, int hi) {
  st->pol_lo = lo;
  st->pol_hi = hi;
}

void ora_set_gain(ora_state *st, int band, double g) { st->gain[band] = g; }

int ora_set_component(ora_state *st, int ic, int type, const char *label, double nu_ref,
                      int cg_group, int sample_amplitude, const double *amplitude,
                      const double *indices) {
  if (ic < 0 || ic >= st->ncomp) return 1;
  ora_comp *c = &st->comp[ic];
  c->type = type;
  strncpy(c->label, label ? label : "", 16);
  c->label[16] = 0;
  c->nu_ref = nu_ref;
  if (c->nu_ref < 1e7) c->nu_ref = c->nu_ref * 1e9; /* dang_param_mod.f90:571-573 */
  c->cg_group = cg_group;
  c->sample_amplitude = sample_amplitude;
  switch (type) { /* dang_component_mod.f90:110,197,285 ... */
    case ORA_MBB: c->nindices = 2; break;
    case ORA_POWERLAW: c->nindices = 1; break;
    case ORA_FREEFREE: c->nindices = 1; break;
    case ORA_LOGNORMAL: c->nindices = 2; break;
    case ORA_CMB: c->nindices = 0; break;
    case ORA_TEMPLATE: c->nindices = 0; break; /* :537 */
    case ORA_T_CMB: c->nindices = 1; break;    /* :430 ff. */
    case ORA_MONOPOLE: c->nindices = 0; break; /* :579-597 */
    case ORA_HI_FIT: c->nindices = 1; break;   /* :599-700 */
    default: return 2;
  }
  size_t n2 = (size_t)st->npix * st->nmaps;
  free(c->amplitude);
  free(c->indices);
  c->amplitude = xcalloc(n2, sizeof(double));
  c->indices = xcalloc(n2 * (c->nindices ? c->nindices : 1), sizeof(double));
  if (amplitude) memcpy(c->amplitude, amplitude, n2 * sizeof(double));
  if (indices && c->nindices) memcpy(c->indices, indices, n2 * c->nindices * sizeof(double));
  for (int l = 0; l < ORA_MAXIND; l++) {
    c->tuned[l] = 1;
    c->sample_nside[l] = st->nside;
    c->index_mode[l] = 2;
    c->uni_prior[l][0] = -HUGE_VAL;
    c->uni_prior[l][1] = HUGE_VAL;
  }
  return 0;
}

/* type 'template', src/dang_component_mod.f90:536-577: template map normalised by its maximum per
 * plane (:574-577, done here as the constructor does), template_amplitudes(nbands,nmaps), corr, nfit */
int ora_set_template(ora_state *st, int ic, const double *template_map, const double *template_amplitudes,
                     const int *corr, int nfit) {
  if (ic < 0 || ic >= st->ncomp) return 1;
  ora_comp *c = &st->comp[ic];
  if (c->type != ORA_TEMPLATE && c->type != ORA_MONOPOLE && c->type != ORA_HI_FIT) return 1;
  const size_t n2 = (size_t)st->npix * st->nmaps;
  free(c->template_map);
  free(c->template_amplitudes);
  free(c->corr);
  c->template_map = xcalloc(n2, sizeof(double));
  c->template_amplitudes = xcalloc((size_t)st->nmaps * st->nbands, sizeof(double));
  c->corr = xcalloc(st->nbands, sizeof(int));
  if (c->type == ORA_MONOPOLE) { /* :591-594: offset map = 1 in intensity, 0 in polarisation */
    for (int i = 0; i < st->npix; i++) c->template_map[IDX2(st, i, 0)] = 1.0;
  } else {
    memcpy(c->template_map, template_map, n2 * sizeof(double));
  }
  for (int k = 0; k < st->nmaps && c->type == ORA_TEMPLATE; k++) { /* :574-577 (type 'template' only) */
    double mx = c->template_map[IDX2(st, 0, k)];
    for (int i = 1; i < st->npix; i++)
      if (c->template_map[IDX2(st, i, k)] > mx) mx = c->template_map[IDX2(st, i, k)];
    for (int i = 0; i < st->npix; i++) c->template_map[IDX2(st, i, k)] = c->template_map[IDX2(st, i, k)] / mx;
  }
  if (template_amplitudes)
    memcpy(c->template_amplitudes, template_amplitudes, (size_t)st->nmaps * st->nbands * sizeof(double));
  int count = 0;
  for (int j = 0; j < st->nbands; j++) {
    c->corr[j] = corr[j] != 0;
    count += c->corr[j];
  }
  c->nfit = nfit;
  return count == nfit ? 0 : 2;
}
double *ora_template_map(ora_state *st, int ic) { return st->comp[ic].template_map; }
double ora_get_T_CMB(void) { return ORA_TCMB; }
void ora_set_T_CMB(double t) { ORA_TCMB = t; }
double *ora_offset(ora_state *st) { return st->offset; }
double *ora_template_amplitudes(ora_state *st, int ic) { return st->comp[ic].template_amplitudes; }

int ora_set_index(ora_state *st, int ic, int nind, int sample_index, int index_mode, int lnl_type,
                  int prior_type, const double gauss[2], const double uni[2], double step_size,
                  int tuned, int sample_nside, const int *pol_flags, int nflag) {
  ora_comp *c = &st->comp[ic];
  if (nind < 0 || nind >= c->nindices || nflag > 3) return 1;
  c->sample_index[nind] = sample_index;
  c->index_mode[nind] = index_mode;
  c->lnl_type[nind] = lnl_type;
  c->prior_type[nind] = prior_type;
  c->gauss_prior[nind][0] = gauss[0];
  c->gauss_prior[nind][1] = gauss[1];
  c->uni_prior[nind][0] = uni[0];
  c->uni_prior[nind][1] = uni[1];
  c->step_size[nind] = step_size;
  c->tuned[nind] = tuned;
  c->sample_nside[nind] = sample_nside;
  c->nflag[nind] = nflag;
  for (int k = 0; k < nflag; k++) c->pol_flag[nind][k] = pol_flags[k];
  return 0;
}

double *ora_amplitude(ora_state *st, int ic) { return st->comp[ic].amplitude; }
double *ora_indices(ora_state *st, int ic) { return st->comp[ic].indices; }
double *ora_sky_model(ora_state *st) { return st->sky_model; }
double *ora_res_map(ora_state *st) { return st->res_map; }
double *ora_chi_map(ora_state *st) { return st->chi_map; }
double ora_step_size(ora_state *st, int ic, int nind) { return st->comp[ic].step_size[nind]; }
int ora_nindices(ora_state *st, int ic) { return st->comp[ic].nindices; }

/* ------------------------------------------------------------------ SEDs */

/* a2t, src/dang_bp_mod.f90:211-243 */
static double ora_a2t(const ora_band *bp) {
  double sum = 0.0, y;
  if (bp->n == 0) {
    if (bp->nu_c > 1e7)
      y = (ORA_H * bp->nu_c) / (ORA_KB * ORA_TCMB);
    else
      y = (ORA_H * bp->nu_c * 1e9) / (ORA_KB * ORA_TCMB);
    sum = ((exp(y) - 1.0) * (exp(y) - 1.0)) / ((y * y) * exp(y));
  } else {
    for (int i = 0; i < bp->n; i++) {
      if (bp->nu0[i] == 0.0) continue;
      if (bp->nu0[i] > 1e7)
        y = (ORA_H * bp->nu0[i]) / (ORA_KB * ORA_TCMB);
      else
        y = (ORA_H * bp->nu0[i] * 1e9) / (ORA_KB * ORA_TCMB);
      sum = sum + bp->tau0[i] * ((exp(y) - 1.0) * (exp(y) - 1.0)) / ((y * y) * exp(y));
    }
  }
  return sum;
}

/* evaluate_powerlaw, src/dang_component_mod.f90:886-918 */
static double eval_powerlaw(const ora_state *st, const ora_comp *c, int band, int pix, int k,
                            const double *theta) {
  const ora_band *bp = &st->bp[band];
  double spectrum = 0.0;
  double beta = theta ? theta[0] : c->indices[IDX2(st, pix, k)];
  if (bp->n == 0) {
    spectrum = pow(bp->nu_c / c->nu_ref, beta);
  } else {
    for (int i = 0; i < bp->n; i++) {
      if (bp->nu0[i] == 0.0) continue;
      spectrum = spectrum + bp->tau0[i] * pow(bp->nu0[i] / c->nu_ref, beta);
    }
  }
  return spectrum;
}

/* evaluate_mbb, src/dang_component_mod.f90:920-958 */
static double eval_mbb(const ora_state *st, const ora_comp *c, int band, int pix, int k,
                       const double *theta) {
  const ora_band *bp = &st->bp[band];
  size_t n2 = (size_t)st->npix * st->nmaps;
  double spectrum = 0.0, beta, td;
  if (theta) {
    beta = theta[0];
    td = theta[1];
  } else {
    beta = c->indices[IDX2(st, pix, k)];
    td = c->indices[n2 + IDX2(st, pix, k)];
  }
  double z = ORA_H / (ORA_KB * td);
  if (bp->n == 0) {
    spectrum = (exp(z * c->nu_ref) - 1.0) / (exp(z * bp->nu_c) - 1.0) *
               pow(bp->nu_c / c->nu_ref, beta + 1.0);
  } else {
    for (int i = 0; i < bp->n; i++) {
      if (bp->nu0[i] == 0.0) continue;
      spectrum = spectrum + bp->tau0[i] * (exp(z * c->nu_ref) - 1.0) /
                                (exp(z * bp->nu0[i]) - 1.0) *
                                pow(bp->nu0[i] / c->nu_ref, beta + 1.0);
    }
  }
  return spectrum;
}

/* evaluate_lognormal, src/dang_component_mod.f90:960-999 (single-precision literals kept) */
static double eval_lognormal(const ora_state *st, const ora_comp *c, int band, int pix, int k,
                             const double *theta) {
  const ora_band *bp = &st->bp[band];
  size_t n2 = (size_t)st->npix * st->nmaps;
  double spectrum = 0.0, nu_p, w_ame;
  if (theta) {
    nu_p = theta[0];
    w_ame = theta[1];
  } else {
    nu_p = c->indices[IDX2(st, pix, k)];
    w_ame = c->indices[n2 + IDX2(st, pix, k)];
  }
  if (bp->n == 0) {
    double t = log(bp->nu_c / (nu_p * (double)1e9f)) / w_ame;
    double r = c->nu_ref / bp->nu_c;
    spectrum = exp(-0.5 * (t * t)) * (r * r);
  } else {
    for (int i = 0; i < bp->n; i++) {
      if (bp->nu0[i] == 0.0) continue;
      double t = log(bp->nu0[i] / (nu_p * (double)1e9f)) / w_ame;
      double r = c->nu_ref / bp->nu0[i];
      spectrum = spectrum + bp->tau0[i] * exp(-0.5 * (t * t)) * (r * r);
    }
  }
  return spectrum;
}

/* evaluate_freefree, src/dang_component_mod.f90:1001-1040 */
static double ff_gaunt(double nu, double T_e) {
  return log(exp(5.960 - sqrt(3.0) / ORA_PI * log(1.0 * nu / 1.e9 * pow(T_e / 1.e4, -1.5))) +
             2.71828);
}
static double eval_freefree(const ora_state *st, const ora_comp *c, int band, int pix, int k,
                            const double *theta) {
  const ora_band *bp = &st->bp[band];
  double spectrum = 0.0;
  double T_e = theta ? theta[0] : c->indices[IDX2(st, pix, k)];
  double S_ref = ff_gaunt(c->nu_ref, T_e);
  if (bp->n == 0) {
    double r = bp->nu_c / c->nu_ref;
    spectrum = ff_gaunt(bp->nu_c, T_e) / S_ref * (1.0 / (r * r));
  } else {
    for (int i = 0; i < bp->n; i++) {
      if (bp->nu0[i] == 0.0) continue;
      double r = bp->nu0[i] / c->nu_ref;
      spectrum = spectrum + bp->tau0[i] * ff_gaunt(bp->nu0[i], T_e) / S_ref * (1.0 / (r * r));
    }
  }
  return spectrum;
}

/* B_nu, src/dang_component_mod.f90:745-752; compute_bnu_prime_RJ, src/dang_bp_mod.f90:160-168 */
static double ora_B_nu(double nu, double T) {
  return ((2.0 * ORA_H * pow(nu, 3.0)) / pow(ORA_C, 2.0)) * (1.0 / (exp((ORA_H * nu) / (ORA_KB * T)) - 1));
}
static double ora_bnu_prime_RJ(double nu) { return 2.0 * ORA_KB * pow(nu, 2.0) / pow(ORA_C, 2.0); }

/* evaluate_T_cmb :815-848 and evaluate_hi_fit :850-884 share one body: the Planck function at T in RJ units */
static double eval_planck_rj(const ora_state *st, const ora_comp *c, int band, int pix, int k,
                             const double *theta) {
  const ora_band *bp = &st->bp[band];
  double spectrum = 0.0;
  double T = theta ? theta[0] : c->indices[IDX2(st, pix, k)];
  if (bp->n == 0) {
    spectrum = ora_B_nu(bp->nu_c, T) / ora_bnu_prime_RJ(bp->nu_c);
  } else {
    for (int i = 0; i < bp->n; i++) {
      if (bp->nu0[i] == 0.0) continue;
      spectrum = spectrum + bp->tau0[i] * ora_B_nu(bp->nu0[i], T) / ora_bnu_prime_RJ(bp->nu0[i]);
    }
  }
  return spectrum * (double)1e6f; /* "*1e6": a default-real literal */
}

/* eval_sed, src/dang_component_mod.f90:778-813.  map_n is 1-based. */
static double eval_sed_c(const ora_state *st, const ora_comp *c, int band, int pix, int map_n,
                         const double *theta) {
  int k = map_n - 1;
  switch (c->type) {
    case ORA_POWERLAW: return eval_powerlaw(st, c, band, pix, k, theta);
    case ORA_MBB: return eval_mbb(st, c, band, pix, k, theta);
    case ORA_FREEFREE: return eval_freefree(st, c, band, pix, k, theta);
    case ORA_LOGNORMAL: return eval_lognormal(st, c, band, pix, k, theta);
    case ORA_CMB: return (double)(1.0f) / ora_a2t(&st->bp[band]);
    case ORA_TEMPLATE: return c->template_map[IDX2(st, pix, map_n - 1)]; /* :803-804 */
    case ORA_T_CMB: return eval_planck_rj(st, c, band, pix, k, theta);     /* :801-802 */
    case ORA_MONOPOLE: return c->template_map[IDX2(st, pix, map_n - 1)];   /* :805-806 */
    case ORA_HI_FIT:                                                        /* :807-808 */
      return c->template_map[IDX2(st, pix, map_n - 1)] * eval_planck_rj(st, c, band, pix, k, theta);
    default: return 0.0;
  }
}

/* eval_signal, src/dang_component_mod.f90:754-776 (template :766-767, diffuse branch :773) */
static double eval_signal_c(const ora_state *st, const ora_comp *c, int band, int pix, int map_n,
                            const double *theta) {
  if (c->type == ORA_TEMPLATE || c->type == ORA_MONOPOLE) /* :766-769 */
    return c->template_amplitudes[(size_t)(map_n - 1) * st->nbands + band] *
           c->template_map[IDX2(st, pix, map_n - 1)];
  if (c->type == ORA_HI_FIT) /* :764-765 */
    return c->template_amplitudes[(size_t)(map_n - 1) * st->nbands + band] * eval_sed_c(st, c, band, pix, map_n, theta);
  if (c->type == ORA_T_CMB) return eval_sed_c(st, c, band, pix, map_n, theta); /* :770-771 */
  return c->amplitude[IDX2(st, pix, map_n - 1)] * eval_sed_c(st, c, band, pix, map_n, theta);
}

double ora_eval_sed(const ora_state *st, int ic, int band, int pix, int map_n,
                    const double *theta) {
  return eval_sed_c(st, &st->comp[ic], band, pix, map_n, theta);
}
double ora_eval_signal(const ora_state *st, int ic, int band, int pix, int map_n,
                       const double *theta) {
  return eval_signal_c(st, &st->comp[ic], band, pix, map_n, theta);
}

/* ------------------------------------------------------------------ amplitude draw */

static int masked(const ora_state *st, int i) {
  return st->masks[i] == 0.0 || st->masks[i] == ORA_MISSVAL;
}

/* pol flag -> planes (1-based).  Only flag 8 (Q+U, S=2) and single planes are live (Q2). */
static int flag_planes(int flag, int planes[2]) {
  if (flag & 8) {
    planes[0] = 2;
    planes[1] = 3;
    return 2;
  }
  int map_n = 1;
  if (flag & 1)
    map_n = 1;
  else if (flag & 2)
    map_n = 2;
  else if (flag & 4)
    map_n = 3;
  planes[0] = planes[1] = map_n;
  return 1;
}

ora_cg *ora_cg_create(ora_state *st, int cg_group, int i_max, double converge,
                      const int *pol_flags, int nflag) {
  ora_cg *g = xcalloc(1, sizeof *g);
  g->st = st;
  g->cg_group = cg_group;
  g->i_max = i_max;
  g->converge = converge;
  g->nflag = nflag;
  for (int f = 0; f < nflag && f < 3; f++) g->pol_flag[f] = pol_flags[f];
  return g;
}

void ora_cg_destroy(ora_cg *g) {
  if (!g) return;
  for (int f = 0; f < 3; f++) free(g->x[f]);
  free(g);
}

long ora_cg_m(const ora_cg *g, int flag_n) {
  int planes[2];
  return (long)flag_planes(g->pol_flag[flag_n], planes) * g->st->npix;
}

long ora_cg_n(const ora_cg *g, int flag_n) {
  long m = ora_cg_m(g, flag_n), n = 0;
  for (int ic = 0; ic < g->st->ncomp; ic++) {
    const ora_comp *c = &g->st->comp[ic];
    if (c->cg_group != g->cg_group || !c->sample_amplitude) continue;
    if (c->type == ORA_TEMPLATE || c->type == ORA_MONOPOLE || c->type == ORA_HI_FIT)
      n += c->nfit; /* :399-414 (templates: only the Q+U branch sizes b correctly) */
    else n += m;
  }
  return n;
}

double *ora_cg_x(ora_cg *g, int flag_n) { return g->x[flag_n]; }

/* 0-based slot of band j among the fitted bands of a template component (the counter l(l_ind) of
 * compute_Ax, :655,886-887), or -1 when the band is not fitted */
static int is_border(const ora_comp *c) {
  return c->type == ORA_TEMPLATE || c->type == ORA_MONOPOLE || c->type == ORA_HI_FIT;
}
/* planes a border component's column / row runs over: a template follows the solve's planes, hi_fit and
 * monopole always use plane 1 and the FIRST npix entries of temp1 (:717-744, :833-866) */
static int border_planes(const ora_comp *c, int S, const int planes[2], int bp[2]) {
  if (c->type == ORA_TEMPLATE) {
    bp[0] = planes[0];
    bp[1] = planes[1];
    return S;
  }
  bp[0] = bp[1] = 1;
  return 1;
}
static int template_slot(const ora_state *st, const ora_comp *c, int j) {
  if (!c->corr[j]) return -1;
  int l = 0;
  for (int jj = 0; jj < j; jj++)
    if (c->corr[jj]) l++;
  (void)st;
  return l;
}

/* compute_rhs, src/dang_cg_mod.f90:326-596 */
void ora_compute_rhs(ora_cg *g, int flag_n, double *b) {
  ora_state *st = g->st;
  const int npix = st->npix, nmaps = st->nmaps, nbands = st->nbands;
  int planes[2];
  const int S = flag_planes(g->pol_flag[flag_n], planes);
  const long n = ora_cg_n(g, flag_n);
  double *data = xcalloc((size_t)npix * nmaps * nbands, sizeof(double));

  for (int k = 0; k < nmaps; k++) /* :368-378 */
    for (int j = 0; j < nbands; j++)
      for (int i = 0; i < npix; i++)
        data[IDX3(st, i, k, j)] =
            (k == 0) ? st->sig_map[IDX3(st, i, k, j)] / st->gain[j] : st->sig_map[IDX3(st, i, k, j)];

  for (long i = 0; i < n; i++) b[i] = 0.0;

  for (int ic = 0; ic < st->ncomp; ic++) { /* :427-461 */
    const ora_comp *c = &st->comp[ic];
    if (c->cg_group != g->cg_group || !c->sample_amplitude) {
#pragma omp parallel for schedule(static)
      for (int i = 0; i < npix; i++) {
        if (masked(st, i)) continue;
        for (int k = 1; k <= nmaps; k++)
          for (int j = 0; j < nbands; j++)
            data[IDX3(st, i, k - 1, j)] =
                data[IDX3(st, i, k - 1, j)] - eval_signal_c(st, c, j, i, k, NULL);
      }
    }
    if (c->type == ORA_TEMPLATE || c->type == ORA_MONOPOLE) { /* :444-460: also removed from the bands they are not fitted to */
      for (int j = 0; j < nbands; j++) {
        if (c->corr[j]) continue;
        for (int i = 0; i < npix; i++) {
          if (masked(st, i)) continue;
          for (int k = 1; k <= nmaps; k++)
            data[IDX3(st, i, k - 1, j)] =
                data[IDX3(st, i, k - 1, j)] - eval_signal_c(st, c, j, i, k, NULL);
        }
      }
    }
  }

  long offset = 0;
  for (int ic = 0; ic < st->ncomp; ic++) { /* :465-521 */
    const ora_comp *c = &st->comp[ic];
    if (c->cg_group != g->cg_group) continue;
    if (!c->sample_amplitude) continue;
    if (is_border(c)) { /* :522-587: one masked, noise-weighted sum per fitted band (hi_fit / monopole: plane 1) */
      int l = 0;
      int bpl[2];
      const int SB = border_planes(c, S, planes, bpl);
      double *val_array = xcalloc((size_t)S * npix, sizeof(double));
      for (int j = 0; j < nbands; j++) {
        if (!c->corr[j]) continue;
        for (long i = 0; i < (long)S * npix; i++) val_array[i] = 0.0;
        for (int i = 0; i < npix; i++) {
          if (masked(st, i)) continue;
          for (int s = 0; s < SB; s++) { /* :568-569: both planes accumulate into val_array(i) */
            const double rms = st->rms_map[IDX3(st, i, bpl[s] - 1, j)];
            val_array[i] = val_array[i] + data[IDX3(st, i, bpl[s] - 1, j)] / (rms * rms) *
                                              eval_sed_c(st, c, j, i, bpl[s], NULL);
          }
        }
        double sum = 0.0;
        for (long i = 0; i < (long)S * npix; i++) sum += val_array[i];
        b[offset + l] = b[offset + l] + sum;
        l++;
      }
      free(val_array);
      offset += c->nfit;
      continue;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < npix; i++) {
      for (int j = 0; j < nbands; j++) {
        if (st->masks[i] == 0.0) { /* :474: only the == 0 test here */
          for (int s = 0; s < S; s++) b[(long)s * npix + i] = 0.0; /* no offset, as in source */
          continue;
        }
        for (int s = 0; s < S; s++) {
          const int map_n = planes[s];
          const double rms = st->rms_map[IDX3(st, i, map_n - 1, j)];
          b[offset + (long)s * npix + i] =
              b[offset + (long)s * npix + i] +
              (data[IDX3(st, i, map_n - 1, j)] * eval_sed_c(st, c, j, i, map_n, NULL)) / (rms * rms);
        }
      }
    }
    offset += (long)S * npix;
  }
  free(data);
}

/* compute_Ax, src/dang_cg_mod.f90:598-911 (diffuse and template components) */
void ora_compute_Ax(ora_cg *g, const double *x, int flag_n, double *res) {
  ora_state *st = g->st;
  const int npix = st->npix, nbands = st->nbands;
  int planes[2];
  const int S = flag_planes(g->pol_flag[flag_n], planes);
  const long m = (long)S * npix, n = ora_cg_n(g, flag_n);
  double *temp1 = xcalloc(m, sizeof(double));
  double *temp3 = xcalloc(n, sizeof(double));
  for (long i = 0; i < n; i++) res[i] = 0.0;

  for (int j = 0; j < nbands; j++) {
    for (long i = 0; i < m; i++) temp1[i] = 0.0;
    for (long i = 0; i < n; i++) temp3[i] = 0.0;
    long offset = 0;
    for (int ic = 0; ic < st->ncomp; ic++) { /* temp1 = T_nu x, :685-769 */
      const ora_comp *c = &st->comp[ic];
      if (c->cg_group != g->cg_group || !c->sample_amplitude) continue;
      if (is_border(c)) { /* :717-768 */
        const int l = template_slot(st, c, j);
        int bpl[2];
        const int SB = border_planes(c, S, planes, bpl);
        if (l >= 0)
          for (int i = 0; i < npix; i++) {
            if (masked(st, i)) continue;
            for (int s = 0; s < SB; s++)
              temp1[(long)s * npix + i] =
                  temp1[(long)s * npix + i] + x[offset + l] * eval_sed_c(st, c, j, i, bpl[s], NULL);
          }
        offset += c->nfit;
        continue;
      }
#pragma omp parallel for schedule(static)
      for (int i = 0; i < npix; i++) {
        if (masked(st, i)) continue;
        for (int s = 0; s < S; s++)
          temp1[(long)s * npix + i] =
              temp1[(long)s * npix + i] +
              x[offset + (long)s * npix + i] * eval_sed_c(st, c, j, i, planes[s], NULL);
      }
      offset += m;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < npix; i++) { /* temp1 = N^-1 temp1, :775-791 */
      if (masked(st, i)) continue;
      for (int s = 0; s < S; s++) {
        const double rms = st->rms_map[IDX3(st, i, planes[s] - 1, j)];
        temp1[(long)s * npix + i] = temp1[(long)s * npix + i] / (rms * rms);
      }
    }
    offset = 0;
    for (int ic = 0; ic < st->ncomp; ic++) { /* temp3 = T_nu^t temp1, :801-894 */
      const ora_comp *c = &st->comp[ic];
      if (c->cg_group != g->cg_group || !c->sample_amplitude) continue;
      if (is_border(c)) { /* :833-893: val_array over the planes, then sum(); the monopole row adds temp1(i) itself (:856) */
        const int l = template_slot(st, c, j);
        int bpl[2];
        const int SB = border_planes(c, S, planes, bpl);
        if (l >= 0) {
          double sum = 0.0;
          for (int s = 0; s < SB; s++)
            for (int i = 0; i < npix; i++) {
              if (masked(st, i)) continue;
              sum += (c->type == ORA_MONOPOLE) ? temp1[(long)s * npix + i]
                                               : temp1[(long)s * npix + i] * eval_sed_c(st, c, j, i, bpl[s], NULL);
            }
          temp3[offset + l] = temp3[offset + l] + sum;
        }
        offset += c->nfit;
        continue;
      }
#pragma omp parallel for schedule(static)
      for (int i = 0; i < npix; i++) {
        if (masked(st, i)) continue;
        for (int s = 0; s < S; s++)
          temp3[offset + (long)s * npix + i] =
              temp1[(long)s * npix + i] * eval_sed_c(st, c, j, i, planes[s], NULL);
      }
      offset += m;
    }
    for (long i = 0; i < n; i++) res[i] = res[i] + temp3[i]; /* :904 */
  }
  free(temp1);
  free(temp3);
}

/* Slot (0-based, relative to the end of the diffuse block) that compute_sample_vector writes the fluctuation of
 * border component `ic` in band `j` to.  The source's counter l (:970) starts at 1 before the band loop and is
 * incremented after every fitted (component, band) pair in loop order bands-outer / components-inner, never reset:
 * with ONE border component this is the band's rank among its fitted bands (the same slot compute_Ax uses); with
 * several the slots interleave across components and run past the component's own block (Q8). */
static long ora_sv_slot(const ora_cg *g, int ic, int j) {
  const ora_state *st = g->st;
  long l = 0;
  for (int jj = 0; jj <= j; jj++)
    for (int c2 = 0; c2 < st->ncomp; c2++) {
      const ora_comp *c = &st->comp[c2];
      if (c->cg_group != g->cg_group || !c->sample_amplitude || !is_border(c)) continue;
      if (jj == j && c2 == ic) return l;
      if (c->corr[jj]) l++;
    }
  return l;
}

/* compute_sample_vector, src/dang_cg_mod.f90:913-1100.
 * fix_q1 == 0 reproduces Q1 (diffuse components all write temp2(i), no offset, :1033-1034);
 * fix_q1 == 1 applies the offset compute_Ax uses (:813-814). */
void ora_compute_sample_vector(ora_cg *g, const double *eta, int flag_n, double *res,
                               int fix_q1) {
  ora_state *st = g->st;
  const int npix = st->npix, nbands = st->nbands;
  int planes[2];
  const int S = flag_planes(g->pol_flag[flag_n], planes);
  const long n = (long)S * npix, m = ora_cg_n(g, flag_n);
  double *temp1 = xcalloc(n, sizeof(double));
  double *temp2 = xcalloc(m, sizeof(double));
  for (long i = 0; i < m; i++) res[i] = 0.0;

  for (int j = 0; j < nbands; j++) {
    for (long i = 0; i < n; i++) temp1[i] = 0.0;
    for (long i = 0; i < m; i++) temp2[i] = 0.0;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < npix; i++) { /* temp1 = N^{-1/2} eta, :1005-1017 */
      if (masked(st, i)) continue;
      for (int s = 0; s < S; s++)
        temp1[(long)s * npix + i] =
            eta[(long)s * npix + i] / st->rms_map[IDX3(st, i, planes[s] - 1, j)];
    }
    long offset = 0;
    long diffuse_len = 0; /* :950-964: templates sit after ALL diffuse entries, counter l never reset (Q8) */
    for (int ic = 0; ic < st->ncomp; ic++) {
      const ora_comp *c = &st->comp[ic];
      if (c->cg_group != g->cg_group || !c->sample_amplitude) continue;
      if (!is_border(c)) diffuse_len += n;
    }
    int l_run = 0; /* the counter `l` of :970: set once, never reset per component (Q8), restarted here per band
                    * because every band revisits the same slots in the same order */
    for (int ic = 0; ic < st->ncomp; ic++) { /* temp2 = T^t temp1, :1021-1097 */
      const ora_comp *c = &st->comp[ic];
      if (c->cg_group != g->cg_group || !c->sample_amplitude) continue;
      if (is_border(c)) { /* :1044-1096 */
        const int l = template_slot(st, c, j);
        int bpl[2];
        const int SB = border_planes(c, S, planes, bpl);
        (void)l_run;
        if (l >= 0) {
          double sum = 0.0;
          for (int s = 0; s < SB; s++)
            for (int i = 0; i < npix; i++) {
              if (masked(st, i)) continue;
              sum += (c->type == ORA_MONOPOLE) ? temp1[(long)s * npix + i]
                                               : temp1[(long)s * npix + i] * eval_sed_c(st, c, j, i, bpl[s], NULL);
            }
          /* Q8: the slot is diffuse_len + (running count of fitted (component, band) pairs met so far in the
           * whole band loop), NOT this component's own block; ora_sv_slot reproduces the running counter */
          const long slot = diffuse_len + ora_sv_slot(g, ic, j);
          temp2[slot] = temp2[slot] + sum;
        }
        continue;
      }
      const long off = fix_q1 ? offset : 0;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < npix; i++) {
        if (masked(st, i)) continue;
        for (int s = 0; s < S; s++)
          temp2[off + (long)s * npix + i] =
              temp1[(long)s * npix + i] * eval_sed_c(st, c, j, i, planes[s], NULL);
      }
      offset += n;
    }
    for (long i = 0; i < m; i++) res[i] = res[i] + temp2[i]; /* :1098 */
  }
  free(temp1);
  free(temp2);
}

/* initialize_x, src/dang_cg_mod.f90:1173-1282 (diffuse) */
static void initialize_x(ora_cg *g, int flag_n) {
  ora_state *st = g->st;
  int planes[2];
  const int S = flag_planes(g->pol_flag[flag_n], planes);
  long offset = 0;
  for (int ic = 0; ic < st->ncomp; ic++) {
    const ora_comp *c = &st->comp[ic];
    if (c->cg_group != g->cg_group || !c->sample_amplitude) continue;
    if (is_border(c)) { /* :1226-1279: Q+U takes plane 2's amplitudes; hi_fit / monopole plane 1 */
      int l = 0;
      const int kp = (c->type == ORA_TEMPLATE) ? planes[0] : 1;
      for (int j = 0; j < st->nbands; j++) {
        if (c->corr[j]) {
          g->x[flag_n][offset + l] = c->template_amplitudes[(size_t)(kp - 1) * st->nbands + j];
          l++;
        }
        if (l >= c->nfit) break;
      }
      offset += l;
      continue;
    }
    for (int s = 0; s < S; s++) {
      for (int i = 0; i < st->npix; i++)
        g->x[flag_n][offset + i] = c->amplitude[IDX2(st, i, planes[s] - 1)];
      offset += st->npix;
    }
  }
}

/* unpack_amplitudes, src/dang_cg_mod.f90:1284-1396 (diffuse) */
void ora_unpack_amplitudes(ora_cg *g, int flag_n) {
  ora_state *st = g->st;
  int planes[2];
  const int S = flag_planes(g->pol_flag[flag_n], planes);
  long offset = 0;
  for (int ic = 0; ic < st->ncomp; ic++) {
    ora_comp *c = &st->comp[ic];
    if (c->cg_group != g->cg_group || !c->sample_amplitude) continue;
    if (is_border(c)) { /* :1337-1392: Q+U writes the fitted value to planes 2 and 3; hi_fit / monopole to plane 1 */
      int l = 0;
      int bpl[2];
      const int SB = border_planes(c, S, planes, bpl);
      for (int j = 0; j < st->nbands; j++) {
        if (c->corr[j]) {
          for (int s = 0; s < SB; s++)
            c->template_amplitudes[(size_t)(bpl[s] - 1) * st->nbands + j] = g->x[flag_n][offset + l];
          l++;
        }
        if (l >= c->nfit) break;
      }
      offset += l;
      continue;
    }
    for (int s = 0; s < S; s++) {
      for (int i = 0; i < st->npix; i++)
        c->amplitude[IDX2(st, i, planes[s] - 1)] = g->x[flag_n][offset + i];
      offset += st->npix;
    }
  }
}

static double vsum_prod(const double *a, const double *b, long n) { /* sum(a*b), sequential */
  double s = 0.0;
  for (long i = 0; i < n; i++) s = s + a[i] * b[i];
  return s;
}

/* cg_search, src/dang_cg_mod.f90:179-324 */
int ora_cg_search(ora_cg *g, int flag_n, const double *b, int ml_mode, const double *eta,
                  int fix_q1, double *delta_final, double *delta_trace, int trace_len) {
  const long n = ora_cg_n(g, flag_n);
  if (!g->x[flag_n]) { /* :227-239 (iter == 1), Q10 */
    g->x[flag_n] = xcalloc(n, sizeof(double));
    initialize_x(g, flag_n);
  }
  double *b2 = xcalloc(n, sizeof(double));
  double *xi = xcalloc(n, sizeof(double));
  double *r = xcalloc(n, sizeof(double));
  double *d = xcalloc(n, sizeof(double));
  double *q = xcalloc(n, sizeof(double));

  if (ml_mode == ORA_SAMPLE && eta) { /* :254-264 */
    ora_compute_sample_vector(g, eta, flag_n, q, fix_q1);
    for (long i = 0; i < n; i++) b2[i] = b[i] + q[i];
  } else { /* :265-266 */
    for (long i = 0; i < n; i++) b2[i] = b[i];
  }
  for (long i = 0; i < n; i++) xi[i] = g->x[flag_n][i]; /* :279 */

  ora_compute_Ax(g, xi, flag_n, q); /* :283 */
  for (long i = 0; i < n; i++) r[i] = b2[i] - q[i];
  for (long i = 0; i < n; i++) d[i] = r[i];
  double delta_new = vsum_prod(r, r, n), delta_old;
  int i = 1;
  if (delta_trace && trace_len > 0) delta_trace[0] = delta_new;

  while (i < g->i_max && delta_new > g->converge) { /* :293-314 */
    ora_compute_Ax(g, d, flag_n, q);
    const double alpha = delta_new / vsum_prod(d, q, n);
    for (long k = 0; k < n; k++) xi[k] = xi[k] + alpha * d[k];
    for (long k = 0; k < n; k++) r[k] = r[k] - alpha * q[k];
    delta_old = delta_new;
    delta_new = vsum_prod(r, r, n);
    const double beta = delta_new / delta_old;
    for (long k = 0; k < n; k++) d[k] = r[k] + beta * d[k];
    i = i + 1;
    if (delta_trace && i - 1 < trace_len) delta_trace[i - 1] = delta_new;
  }
  for (long k = 0; k < n; k++) g->x[flag_n][k] = xi[k]; /* :319 */
  if (delta_final) *delta_final = delta_new;
  free(b2);
  free(xi);
  free(r);
  free(d);
  free(q);
  return i;
}

/* sample_cg_groups loop body, src/dang_cg_mod.f90:166-172 */
int ora_sample_cg_group(ora_cg *g, int ml_mode, const double *eta, int fix_q1, int *niter,
                        double *delta_final) {
  for (int f = 0; f < g->nflag; f++) {
    const long n = ora_cg_n(g, f);
    double *b = xcalloc(n, sizeof(double));
    ora_compute_rhs(g, f, b);
    int it = ora_cg_search(g, f, b, ml_mode, eta, fix_q1, delta_final ? &delta_final[f] : NULL,
                           NULL, 0);
    if (niter) niter[f] = it;
    ora_unpack_amplitudes(g, f);
    free(b);
    if (eta) eta += ora_cg_m(g, f);
  }
  ora_update_sky_model(g->st);
  return 0;
}

/* ------------------------------------------------------------------ data object */

/* update_sky_model, src/dang_data_mod.f90:339-396 */
void ora_update_sky_model(ora_state *st) {
  const int npix = st->npix, nmaps = st->nmaps, nbands = st->nbands;
  size_t n3 = (size_t)npix * nmaps * nbands;
  for (size_t i = 0; i < n3; i++) st->sky_model[i] = 0.0;
  for (int l = 0; l < st->ncomp; l++) {
    const ora_comp *c = &st->comp[l];
    if (c->type == ORA_MONOPOLE) { /* :357-361: the band monopoles become the offsets, not part of the sky model */
      for (int j = 0; j < nbands; j++) st->offset[j] = c->template_amplitudes[j];
      continue;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < npix; i++)
      for (int k = 1; k <= nmaps; k++)
        for (int j = 0; j < nbands; j++)
          st->sky_model[IDX3(st, i, k - 1, j)] =
              st->sky_model[IDX3(st, i, k - 1, j)] + eval_signal_c(st, c, j, i, k, NULL);
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < npix; i++)
    for (int k = 1; k <= nmaps; k++)
      for (int j = 0; j < nbands; j++) {
        if (k == 1)
          st->res_map[IDX3(st, i, 0, j)] = (st->sig_map[IDX3(st, i, 0, j)] - st->offset[j]) /
                                               st->gain[j] -
                                           st->sky_model[IDX3(st, i, 0, j)];
        else
          st->res_map[IDX3(st, i, k - 1, j)] =
              st->sig_map[IDX3(st, i, k - 1, j)] - st->sky_model[IDX3(st, i, k - 1, j)];
      }
}

/* compute_chisq, src/dang_data_mod.f90:494-526.  Returns chisq = sum(chi_map)/nump with the
 * Q9 convention nump = nmaps * #unmasked; chi_sum_planes gets the un-normalised per-plane sums. */
double ora_compute_chisq(ora_state *st, double *chi_sum_planes) {
  const int npix = st->npix, nmaps = st->nmaps, nbands = st->nbands;
  for (size_t i = 0; i < (size_t)npix * nmaps; i++) st->chi_map[i] = 0.0;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < npix; i++) {
    if (masked(st, i)) continue;
    for (int k = st->pol_lo; k <= st->pol_hi; k++) {
      for (int j = 0; j < nbands; j++) {
        const double rms = st->rms_map[IDX3(st, i, k - 1, j)];
        double t;
        if (k == 1)
          t = (st->sig_map[IDX3(st, i, 0, j)] - st->offset[j]) / st->gain[j] -
              st->sky_model[IDX3(st, i, 0, j)];
        else
          t = st->sig_map[IDX3(st, i, k - 1, j)] - st->sky_model[IDX3(st, i, k - 1, j)];
        st->chi_map[IDX2(st, i, k - 1)] = st->chi_map[IDX2(st, i, k - 1)] + (t * t) / (rms * rms);
      }
    }
  }
  for (size_t i = 0; i < (size_t)npix * nmaps; i++) st->chi_map[i] = st->chi_map[i] / nbands;
  double total = 0.0; /* sum(self%chi_map): column-major order = plane by plane */
  for (int k = 0; k < nmaps; k++) {
    double sk = 0.0;
    for (int i = 0; i < npix; i++) {
      total = total + st->chi_map[IDX2(st, i, k)];
      sk = sk + st->chi_map[IDX2(st, i, k)];
    }
    if (chi_sum_planes) chi_sum_planes[k] = sk;
  }
  st->chisq = total / (double)st->nump;
  return st->chisq;
}

/* mask_avg, src/dang_util_mod.f90:186-206 */
double ora_mask_avg(const ora_state *st, int ic, int nind, int map_n) {
  const ora_comp *c = &st->comp[ic];
  size_t n2 = (size_t)st->npix * st->nmaps;
  double sum = 0.0;
  long cnt = 0;
  for (int i = 0; i < st->npix; i++) {
    if (masked(st, i)) continue;
    sum = sum + c->indices[n2 * nind + IDX2(st, i, map_n - 1)];
    cnt++;
  }
  return sum / (double)cnt;
}

/* ------------------------------------------------------------------ likelihood */

/* evaluate_lnL, src/dang_lnl_mod.f90:126-182 */
double ora_evaluate_lnL(const ora_state *st, const double *data, const double *rms,
                        const double *model, const int map_inds[2], int pixel,
                        const double *mask) {
  const int nbands = st->nbands;
  int lo = 0, hi = st->npix - 1;
  if (pixel > -1) lo = hi = pixel;
  double lnL_local = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : lnL_local) if (pixel < 0)
  for (int i = lo; i <= hi; i++) {
    if (mask[i] == 0.0 || mask[i] == ORA_MISSVAL) continue;
    for (int k = map_inds[0]; k <= map_inds[1]; k++)
      for (int j = 0; j < nbands; j++) {
        const double t = (data[IDX3(st, i, k - 1, j)] - model[IDX3(st, i, k - 1, j)]) /
                         rms[IDX3(st, i, k - 1, j)];
        lnL_local = lnL_local - 0.5 * (t * t);
      }
  }
  return 0.0 + lnL_local;
}

/* evaluate_marginal_lnL, src/dang_lnl_mod.f90:47-124 (mask is ignored by the source) */
double ora_evaluate_marginal_lnL(const ora_state *st, const double *data, const double *rms,
                                 const double *model, const int map_inds[2], int pixel) {
  int lo = 0, hi = st->npix - 1;
  if (pixel > -1) lo = hi = pixel;
  double lnL = 0.0;
  for (int j = 0; j < st->nbands; j++)
    for (int k = map_inds[0]; k <= map_inds[1]; k++) {
      double TNd = 0.0, TNT = 0.0;
      for (int i = lo; i <= hi; i++) {
        const double r = rms[IDX3(st, i, k - 1, j)];
        const double TN = model[IDX3(st, i, k - 1, j)] / (r * r);
        TNd = TNd + TN * data[IDX3(st, i, k - 1, j)];
      }
      for (int i = lo; i <= hi; i++) {
        const double r = rms[IDX3(st, i, k - 1, j)];
        const double TN = model[IDX3(st, i, k - 1, j)] / (r * r);
        TNT = TNT + TN * model[IDX3(st, i, k - 1, j)];
      }
      const double invTNT = 1.0 / TNT;
      lnL = lnL - 0.5 * TNd * invTNT * TNd;
    }
  return lnL;
}

/* eval_normal_prior, src/dang_util_mod.f90:112-121 */
double ora_eval_normal_prior(double prop, double mean, double std) {
  const double var = std * std;
  const double num = exp(-((prop - mean) * (prop - mean)) / (2 * var));
  const double denom = std * sqrt(2.0 * ORA_PI);
  return num / denom;
}

/* rand_normal, src/dang_util_mod.f90:100-110 (sine branch of Box-Muller) */
double ora_rand_normal_from_uniform(double mean, double stdev, double u1, double u2) {
  const double r = pow(-2.0 * log(u1), 0.5);
  const double theta = 2.0 * ORA_PI * u2;
  return mean + stdev * r * sin(theta);
}

/* eval_jeffreys_prior, src/dang_lnl_mod.f90:242-304 */
static double eval_jeffreys_prior(const ora_state *st, const ora_comp *c, const double *rms,
                                  const int map_inds[2], int pixel, const double *mask,
                                  double val) {
  double theta[2] = {val, 0.0};
  double sum = 0.0;
  int lo = 0, hi = st->npix - 1;
  if (pixel > -1) lo = hi = pixel;
  if (strcmp(c->label, "synch") == 0) {
    for (int i = lo; i <= hi; i++) {
      if (mask[i] == 0.0 || mask[i] == ORA_MISSVAL) continue;
      for (int k = map_inds[0]; k <= map_inds[1]; k++)
        for (int j = 0; j < st->nbands; j++) {
          const double ss = eval_signal_c(st, c, j, i, k, theta);
          const double ir = 1.0 / rms[IDX3(st, i, k - 1, j)];
          const double t =
              ((ir * ir) * (ss / c->amplitude[IDX2(st, i, k - 1)]) * log(st->bp[j].nu_c / c->nu_ref));
          sum = sum + t * t;
        }
    }
  }
  return sqrt(sum);
}

/* ------------------------------------------------------------------ spectral-parameter draw */

/* update_sample_model, src/dang_sample_mod.f90:520-568 */
static void update_sample_model(const ora_state *st, double *model, const ora_comp *c,
                                const int map_inds[2], const double *sample, int pixel) {
  int lo = 0, hi = st->npix - 1;
  if (pixel > -1) lo = hi = pixel;
#pragma omp parallel for schedule(static) if (pixel < 0)
  for (int i = lo; i <= hi; i++)
    for (int k = map_inds[0]; k <= map_inds[1]; k++)
      for (int j = 0; j < st->nbands; j++)
        model[IDX3(st, i, k - 1, j)] = eval_signal_c(st, c, j, i, k, sample);
}

static void set_map_inds(int map_n, int map_inds[2]) { /* :157-163 */
  if (map_n == -1) {
    map_inds[0] = 2;
    map_inds[1] = 3;
  } else if (map_n == -2) {
    map_inds[0] = 1;
    map_inds[1] = 3;
  } else {
    map_inds[0] = map_inds[1] = map_n;
  }
}

/* data_raw construction, src/dang_sample_mod.f90:168-196 */
static double *build_mh_data(const ora_state *st, const ora_comp *c) {
  const int npix = st->npix, nmaps = st->nmaps, nbands = st->nbands;
  double *data = xcalloc((size_t)npix * nmaps * nbands, sizeof(double));
  for (int j = 0; j < nbands; j++)
    for (int k = 0; k < nmaps; k++)
      for (int i = 0; i < npix; i++)
        data[IDX3(st, i, k, j)] = (k == 0)
                                      ? (st->sig_map[IDX3(st, i, 0, j)] - st->offset[j]) / st->gain[j]
                                      : st->sig_map[IDX3(st, i, k, j)];
  for (int l = 0; l < st->ncomp; l++) {
    const ora_comp *c2 = &st->comp[l];
    if (strcmp(c2->label, c->label) == 0) continue;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < npix; i++)
      for (int k = 1; k <= nmaps; k++)
        for (int j = 0; j < nbands; j++)
          data[IDX3(st, i, k - 1, j)] =
              data[IDX3(st, i, k - 1, j)] - eval_signal_c(st, c2, j, i, k, NULL);
  }
  return data;
}

static double lnl_dispatch(const ora_state *st, int lnl_type, const double *data,
                           const double *model, const int map_inds[2], int pixel, double prev) {
  if (lnl_type == ORA_LNL_CHISQ)
    return ora_evaluate_lnL(st, data, st->rms_map, model, map_inds, pixel, st->masks);
  if (lnl_type == ORA_LNL_MARGINAL)
    return ora_evaluate_marginal_lnL(st, data, st->rms_map, model, map_inds, pixel);
  return prev;
}

static double prior_dispatch(const ora_state *st, const ora_comp *c, int nind,
                             const int map_inds[2], int pixel, double val, double prev) {
  if (c->prior_type[nind] == ORA_PRIOR_GAUSSIAN)
    return log(ora_eval_normal_prior(val, c->gauss_prior[nind][0], c->gauss_prior[nind][1]));
  if (c->prior_type[nind] == ORA_PRIOR_JEFFREYS)
    return log(eval_jeffreys_prior(st, c, st->rms_map, map_inds, pixel, st->masks, val));
  if (c->prior_type[nind] == ORA_PRIOR_UNIFORM) return 0.0;
  return prev;
}

/* sample_index_mh, src/dang_sample_mod.f90:88-485 (sample_nside == nside only) */
int ora_sample_index_mh(ora_state *st, int ic, int nind, int map_n, int nsample, int ml_mode,
                        const double *z, const double *u, double *accept_out,
                        unsigned char *decisions, double *lnl_trace) {
  ora_comp *c = &st->comp[ic];
  const int npix = st->npix, nmaps = st->nmaps;
  const size_t n2 = (size_t)npix * nmaps;
  if (c->sample_nside[nind] != st->nside) return 2; /* udgrade path not restated (HEALPix) */
  if (!c->tuned[nind]) return 3;                    /* call ora_tune_step first */
  int map_inds[2];
  set_map_inds(map_n, map_inds);

  double *data = build_mh_data(st, c);
  double *index_full_res = xcalloc(n2, sizeof(double));
  double accept_total = 0.0;

  if (c->index_mode[nind] == 1) { /* full sky, :229-329 */
    double *model = xcalloc(n2 * st->nbands, sizeof(double));
    double sample[ORA_MAXIND] = {0, 0}, theta[ORA_MAXIND] = {0, 0};
    double lnl = 0.0, lnl_prior = 0.0, lnl_old, lnl_new;
    int sample_it = 1;
    for (int l = 0; l < c->nindices; l++)
      sample[l] = c->indices[n2 * l + IDX2(st, 0, map_inds[0] - 1)];
    for (int l = 0; l < c->nindices; l++) theta[l] = sample[l];
    update_sample_model(st, model, c, map_inds, sample, -1);
    if (c->lnl_type[nind] == ORA_LNL_PRIOR) { /* :255-257 */
      sample_it = 0;
      sample[nind] = c->gauss_prior[nind][0] + c->gauss_prior[nind][1] * z[0];
    } else {
      lnl = lnl_dispatch(st, c->lnl_type[nind], data, model, map_inds, -1, lnl);
    }
    lnl_prior = prior_dispatch(st, c, nind, map_inds, -1, sample[nind], lnl_prior);
    lnl_old = lnl + lnl_prior;
    if (sample_it) {
      for (int l = 0; l < c->nindices; l++)
        sample[l] = c->indices[n2 * l + IDX2(st, 0, map_inds[0] - 1)];
      for (int l = 0; l < c->nindices; l++) theta[l] = sample[l];
      for (int l = 0; l < nsample; l++) { /* :282-324 */
        if (lnl_trace) lnl_trace[l] = NAN;
        theta[nind] = sample[nind] + (0.0 + c->step_size[nind] * z[l]);
        if (theta[nind] < c->uni_prior[nind][0] || theta[nind] > c->uni_prior[nind][1]) {
          if (decisions) decisions[l] = 2;
          continue;
        }
        update_sample_model(st, model, c, map_inds, theta, -1);
        lnl = lnl_dispatch(st, c->lnl_type[nind], data, model, map_inds, -1, lnl);
        lnl_prior = prior_dispatch(st, c, nind, map_inds, -1, theta[nind], lnl_prior);
        lnl_new = lnl + lnl_prior;
        if (lnl_trace) lnl_trace[l] = lnl_new;
        const double diff = lnl_new - lnl_old;
        const double ratio = exp(diff);
        int acc = 0;
        if (ml_mode == ORA_OPTIMIZE)
          acc = ratio > 1.0;
        else
          acc = ratio > u[l]; /* :318-319 */
        if (acc) {
          sample[nind] = theta[nind];
          lnl_old = lnl_new;
          accept_total += 1.0;
        }
        if (decisions) decisions[l] = (unsigned char)acc;
      }
    }
    for (int k = map_inds[0]; k <= map_inds[1]; k++) /* :329 */
      for (int i = 0; i < npix; i++) index_full_res[IDX2(st, i, k - 1)] = sample[nind];
    free(model);
  } else { /* per pixel, :332-481 */
    double *model = xcalloc(n2 * st->nbands, sizeof(double));
#pragma omp parallel for schedule(static) reduction(+ : accept_total)
    for (int i = 0; i < npix; i++) {
      if (decisions)
        for (int l = 0; l < nsample; l++) decisions[(size_t)l * npix + i] = 3;
      if (lnl_trace)
        for (int l = 0; l < nsample; l++) lnl_trace[(size_t)l * npix + i] = NAN;
      if (st->masks[i] == ORA_MISSVAL || st->masks[i] == 0.0) continue; /* :362 */
      int sample_it = 1;
      double lnl = 0.0, lnl_old = 0.0, lnl_new = 0.0, lnl_prior = 0.0;
      double sample[ORA_MAXIND] = {0, 0}, theta[ORA_MAXIND] = {0, 0};
      for (int l = 0; l < c->nindices; l++)
        sample[l] = c->indices[n2 * l + IDX2(st, i, map_inds[0] - 1)];
      for (int l = 0; l < c->nindices; l++) theta[l] = sample[l];
      update_sample_model(st, model, c, map_inds, sample, i);
      if (c->lnl_type[nind] == ORA_LNL_PRIOR) {
        sample_it = 0;
        sample[nind] = c->gauss_prior[nind][0] + c->gauss_prior[nind][1] * z[i];
      } else {
        lnl = lnl_dispatch(st, c->lnl_type[nind], data, model, map_inds, i, lnl);
      }
      lnl_prior = prior_dispatch(st, c, nind, map_inds, i, sample[nind], lnl_prior);
      lnl_old = lnl + lnl_prior;
      if (sample_it) {
        for (int l = 0; l < nsample; l++) { /* :410-455 */
          const size_t slot = (size_t)l * npix + i;
          theta[nind] = sample[nind] + (0.0 + c->step_size[nind] * z[slot]);
          if (theta[nind] < c->uni_prior[nind][0] || theta[nind] > c->uni_prior[nind][1]) {
            if (decisions) decisions[slot] = 2;
            continue;
          }
          update_sample_model(st, model, c, map_inds, theta, i);
          lnl = lnl_dispatch(st, c->lnl_type[nind], data, model, map_inds, i, lnl);
          lnl_prior = prior_dispatch(st, c, nind, map_inds, i, theta[nind], lnl_prior);
          lnl_new = lnl + lnl_prior;
          if (lnl_trace) lnl_trace[slot] = lnl_new;
          const double diff = lnl_new - lnl_old;
          int acc = 0;
          if (ml_mode == ORA_OPTIMIZE)
            acc = diff > 0.0;
          else
            acc = diff > log(u[slot]); /* :449-450, Q4 */
          if (acc) {
            sample[nind] = theta[nind];
            lnl_old = lnl_new;
            accept_total += 1.0;
          }
          if (decisions) decisions[slot] = (unsigned char)acc;
        }
      }
      for (int k = map_inds[0]; k <= map_inds[1]; k++) /* :465 */
        index_full_res[IDX2(st, i, k - 1)] = sample[nind];
    }
    free(model);
  }
  for (int k = map_inds[0]; k <= map_inds[1]; k++) /* :483 */
    for (int i = 0; i < npix; i++)
      c->indices[n2 * nind + IDX2(st, i, k - 1)] = index_full_res[IDX2(st, i, k - 1)];
  if (accept_out) *accept_out = accept_total;
  free(data);
  free(index_full_res);
  return 0;
}

/* sample_spectral_parameters, src/dang_sample_mod.f90:21-86 */
int ora_sample_spectral_parameters(ora_state *st, int nsample, int ml_mode, const double *z,
                                   const double *u) {
  int ncall = 0, sampled = 0;
  const size_t stride = (size_t)nsample * st->npix;
  for (int i = 0; i < st->ncomp; i++) {
    ora_comp *c = &st->comp[i];
    if (c->nindices == 0) continue;
    int any = 0;
    for (int j = 0; j < c->nindices; j++) any |= c->sample_index[j];
    if (!any) continue;
    sampled = 1;
    for (int j = 0; j < c->nindices; j++) {
      if (!c->sample_index[j]) continue;
      for (int k = 0; k < c->nflag[j]; k++) {
        int map_n;
        const int f = c->pol_flag[j][k];
        if (f & 1)
          map_n = 1;
        else if (f & 2)
          map_n = 2;
        else if (f & 4)
          map_n = 3;
        else if (f & 8)
          map_n = -1;
        else
          continue;
        ora_sample_index_mh(st, i, j, map_n, nsample, ml_mode, z + stride * ncall,
                            u + stride * ncall, NULL, NULL, NULL);
        ncall++;
      }
    }
    if (c->type == ORA_T_CMB) ORA_TCMB = c->indices[IDX2(st, 0, 0)]; /* :76-78: T_CMB = c%indices(0,1,1) */
  }
  if (sampled) ora_update_sky_model(st);
  return ncall;
}

/* tune_spectral_parameter_length, src/dang_sample_mod.f90:623-717 (full-sky chain started at
 * indices(0, map_inds(1), :), the full-sky call site :272-275) */
int ora_tune_step_from(ora_state *st, int ic, int nind, int map_n, int nsample, int ml_mode, const double *z,
                       const double *u, int max_blocks, const double *theta_init);
int ora_tune_step(ora_state *st, int ic, int nind, int map_n, int nsample, int ml_mode,
                  const double *z, const double *u, int max_blocks) {
  return ora_tune_step_from(st, ic, nind, map_n, nsample, ml_mode, z, u, max_blocks, NULL);
}
/* the per-pixel call site's start, src/dang_sample_mod.f90:341-347: sample(l) = sum(c%indices(:,map_inds(1),l)) /
 * sum(mask(:,1)) -- the numerator runs over EVERY pixel; the tuner is called inside the loop over l, so for
 * l = 1 the later entries of sample are still 0 (returned here for l = nind only, the others 0) */
void ora_perpixel_tune_start(const ora_state *st, int ic, int nind, int map_n, double *theta_init) {
  const ora_comp *c = &st->comp[ic];
  const size_t n2 = (size_t)st->npix * st->nmaps;
  int map_inds[2];
  set_map_inds(map_n, map_inds);
  double num = 0.0, den = 0.0;
  for (int i = 0; i < st->npix; i++) {
    num += c->indices[n2 * nind + IDX2(st, i, map_inds[0] - 1)];
    den += st->masks[i];
  }
  for (int l = 0; l < ORA_MAXIND; l++) theta_init[l] = 0.0;
  theta_init[nind] = num / den;
}
/* theta_init: the tuner's start (theta_init argument of :623); NULL = indices(0, map_inds(1), :) (:240-243) */
int ora_tune_step_from(ora_state *st, int ic, int nind, int map_n, int nsample, int ml_mode, const double *z,
                       const double *u, int max_blocks, const double *theta_init) {
  ora_comp *c = &st->comp[ic];
  const size_t n2 = (size_t)st->npix * st->nmaps;
  int map_inds[2];
  set_map_inds(map_n, map_inds);
  double *data = build_mh_data(st, c);
  double *model = xcalloc(n2 * st->nbands, sizeof(double));
  double sample[ORA_MAXIND] = {0, 0}, theta[ORA_MAXIND] = {0, 0};
  double lnl = 0.0, lnl_new = 0.0, lnl_old = 0.0;
  for (int l = 0; l < c->nindices; l++)
    sample[l] = theta[l] = theta_init ? theta_init[l] : c->indices[n2 * l + IDX2(st, 0, map_inds[0] - 1)];
  update_sample_model(st, model, c, map_inds, sample, -1);
  lnl = lnl_dispatch(st, c->lnl_type[nind], data, model, map_inds, -1, lnl);
  if (c->prior_type[nind] == ORA_PRIOR_GAUSSIAN)
    lnl_old = lnl + log(ora_eval_normal_prior(sample[nind], c->gauss_prior[nind][0],
                                              c->gauss_prior[nind][1]));
  else if (c->prior_type[nind] == ORA_PRIOR_UNIFORM)
    lnl_old = lnl;
  int blk = 0;
  while (!c->tuned[nind] && blk < max_blocks) {
    double accept = 0.0;
    for (int l = 0; l < nsample; l++) {
      const size_t slot = (size_t)blk * nsample + l;
      theta[nind] = sample[nind] + (0.0 + c->step_size[nind] * z[slot]);
      if (theta[nind] < c->uni_prior[nind][0] || theta[nind] > c->uni_prior[nind][1]) continue;
      update_sample_model(st, model, c, map_inds, theta, -1);
      lnl = lnl_dispatch(st, c->lnl_type[nind], data, model, map_inds, -1, lnl);
      if (c->prior_type[nind] == ORA_PRIOR_GAUSSIAN)
        lnl_new = lnl + log(ora_eval_normal_prior(theta[nind], c->gauss_prior[nind][0],
                                                  c->gauss_prior[nind][1]));
      else if (c->prior_type[nind] == ORA_PRIOR_UNIFORM)
        lnl_new = lnl;
      const double diff = lnl_new - lnl_old;
      const double ratio = exp(diff);
      int acc = (ml_mode == ORA_OPTIMIZE) ? (ratio > 1.0) : (ratio > u[slot]);
      if (acc) {
        sample[nind] = theta[nind];
        lnl_old = lnl_new;
        accept = accept + 1;
      }
      lnl = 0.0;
    }
    /* after a Fortran "do l = 1, nsample" the counter is nsample+1 (:707) */
    const double rate = accept / (double)(nsample + 1);
    if (rate < (double)0.4f)
      c->step_size[nind] = c->step_size[nind] - (double)0.5f * c->step_size[nind];
    else if (rate > (double)0.6f)
      c->step_size[nind] = c->step_size[nind] + (double)0.5f * c->step_size[nind];
    else
      for (int l = 0; l < c->nindices; l++) c->tuned[l] = 1; /* c%tuned = .true. (all) */
    blk++;
  }
  free(data);
  free(model);
  return blk;
}

/* fit_band_gain, src/dang_sample_mod.f90:570-621 */
double ora_fit_band_gain(ora_state *st, int map_n, int band, int ml_mode, double z) {
  double mu = 0.0, sigma = 0.0;
  for (int i = 0; i < st->npix; i++) {
    if (masked(st, i)) continue;
    const double noise = st->rms_map[IDX3(st, i, map_n - 1, band)];
    const double N_inv = 1.0 / (noise * noise);
    const double map1 = st->sky_model[IDX3(st, i, map_n - 1, band)];
    const double map2 = st->res_map[IDX3(st, i, map_n - 1, band)] + map1;
    mu = mu + map2 * N_inv * map1;
    sigma = sigma + map1 * N_inv * map1;
  }
  mu = mu / sigma;
  sigma = sqrt(1.0 / sigma);
  const double gain = (ml_mode == ORA_OPTIMIZE) ? mu : mu + sigma * (0.0 + 1.0 * z);
  st->gain[band] = gain;
  return gain;
}

/* ------------------------------------------------------------------ Philox4x32-10 */

static void philox4x32_10(unsigned int c[4], unsigned int k0, unsigned int k1) {
  for (int r = 0; r < 10; r++) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0];
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[2];
    const unsigned int n0 = (unsigned int)(p1 >> 32) ^ c[1] ^ k0;
    const unsigned int n1 = (unsigned int)p1;
    const unsigned int n2 = (unsigned int)(p0 >> 32) ^ c[3] ^ k1;
    const unsigned int n3 = (unsigned int)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

void ora_philox_raw(unsigned int ctr[4], unsigned int k0, unsigned int k1) { philox4x32_10(ctr, k0, k1); }

void ora_philox_uniform2(unsigned long long seed, unsigned int stream, unsigned long long slot,
                         double *u1, double *u2) {
  unsigned int c[4] = {(unsigned int)slot, (unsigned int)(slot >> 32), stream, 0x44414e47u};
  philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
  const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1];
  const unsigned long long b = ((unsigned long long)c[2] << 32) | c[3];
  /* 53-bit mantissa, offset by half an ulp: strictly inside (0,1) */
  *u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  *u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

void ora_philox_normals(unsigned long long seed, unsigned int stream, unsigned long long slot0,
                        long n, double *z) {
  for (long i = 0; i < n; i++) {
    double u1, u2;
    ora_philox_uniform2(seed, stream, slot0 + (unsigned long long)i, &u1, &u2);
    z[i] = ora_rand_normal_from_uniform(0.0, 1.0, u1, u2);
  }
}

void ora_philox_uniforms(unsigned long long seed, unsigned int stream, unsigned long long slot0,
                         long n, double *u) {
  for (long i = 0; i < n; i++) {
    double u1, u2;
    ora_philox_uniform2(seed, stream, slot0 + (unsigned long long)i, &u1, &u2);
    u[i] = u1;
  }
}

/* ------------------------------------------------------------------ udgrade
 * HEALPix-F90 udgrade_ring (module udgrade_nr, Healpix 3.8x: un-vendored, src/Makefile_gnu:28) as published:
 * RING -> NESTED, sub_udgrade_nest, NESTED -> RING.  Degrading averages the (nside_in/nside_out)^2 NESTED children of
 * every output pixel, skipping children equal to the bad value -1.6375e30 (all bad -> bad value; pessimistic = off);
 * upgrading copies the parent into its children.  Index conversions follow healpix_base (ring2xyf / xyf2nest ...).
 * Call sites: src/dang_sample_mod.f90:204-217, 480; wrappers udgrade_rms / udgrade_mask src/dang_util_mod.f90:341-376. */
static const int hp_jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
static const int hp_jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

static long hp_isqrt(long v) {
  long r = (long)sqrt((double)v + 0.5);
  while (r * r > v) r--;
  while ((r + 1) * (r + 1) <= v) r++;
  return r;
}
static long hp_spread_bits(long v) { /* bit i of v -> bit 2i */
  long r = 0;
  for (int i = 0; i < 31; i++) r |= ((v >> i) & 1L) << (2 * i);
  return r;
}
static long hp_compress_bits(long v) { /* bit 2i of v -> bit i */
  long r = 0;
  for (int i = 0; i < 31; i++) r |= ((v >> (2 * i)) & 1L) << i;
  return r;
}
long ora_nest2ring(long nside, long pix) {
  const long npface = nside * nside, npix = 12 * npface, ncap = 2 * nside * (nside - 1), nl4 = 4 * nside;
  const int face = (int)(pix / npface);
  const long p = pix & (npface - 1);
  const long ix = hp_compress_bits(p), iy = hp_compress_bits(p >> 1);
  const long jr = hp_jrll[face] * nside - ix - iy - 1;
  long nr, n_before, kshift;
  if (jr < nside) {
    nr = jr;
    n_before = 2 * nr * (nr - 1);
    kshift = 0;
  } else if (jr > 3 * nside) {
    nr = nl4 - jr;
    n_before = npix - 2 * (nr + 1) * nr;
    kshift = 0;
  } else {
    nr = nside;
    n_before = ncap + (jr - nside) * nl4;
    kshift = (jr - nside) & 1;
  }
  long jp = (hp_jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  else if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}
long ora_ring2nest(long nside, long pix) {
  const long npface = nside * nside, npix = 12 * npface, ncap = 2 * nside * (nside - 1), nl2 = 2 * nside;
  long iring, iphi, kshift, nr;
  int face;
  if (pix < ncap) {
    iring = (1 + hp_isqrt(1 + 2 * pix)) >> 1;
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0;
    nr = iring;
    face = (int)((iphi - 1) / nr);
  } else if (pix < npix - ncap) {
    const long ip = pix - ncap;
    iring = ip / (4 * nside) + nside;
    iphi = ip % (4 * nside) + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    const long ire = iring - nside + 1, irm = nl2 + 2 - ire;
    const long ifm = (iphi - ire / 2 + nside - 1) / nside, ifp = (iphi - irm / 2 + nside - 1) / nside;
    if (ifp == ifm) face = (ifp == 4) ? 4 : (int)ifp + 4;
    else if (ifp < ifm) face = (int)ifp;
    else face = (int)ifm + 8;
  } else {
    const long ip = npix - pix;
    iring = (1 + hp_isqrt(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0;
    nr = iring;
    iring = 2 * nl2 - iring;
    face = 8 + (int)((iphi - 1) / nr);
  }
  const long irt = iring - hp_jrll[face] * nside + 1;
  long ipt = 2 * iphi - hp_jpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  const long ix = (ipt - irt) >> 1, iy = (-(ipt + irt)) >> 1;
  return (long)face * npface + hp_spread_bits(ix) + (hp_spread_bits(iy) << 1);
}

/* udgrade_ring on `nmaps` planes, Fortran (npix, nmaps) == C [map][pix] */
void ora_udgrade_ring(const double *in, long nside_in, double *out, long nside_out, int nmaps) {
  const long npix_in = 12 * nside_in * nside_in, npix_out = 12 * nside_out * nside_out;
  const double bad = ORA_MISSVAL; /* HPX_DBADVAL == dang's missval, -1.6375e30 */
  for (int k = 0; k < nmaps; k++) {
    const double *mi = in + (size_t)k * npix_in;
    double *mo = out + (size_t)k * npix_out;
    if (nside_out < nside_in) {
      const long npratio = npix_in / npix_out;
      for (long id = 0; id < npix_out; id++) { /* id: NESTED output pixel */
        double total = 0.0;
        long nobs = 0;
        for (long ip = 0; ip < npratio; ip++) {
          const double v = mi[ora_nest2ring(nside_in, id * npratio + ip)];
          if (v != bad) {
            total = total + v;
            nobs++;
          }
        }
        mo[ora_nest2ring(nside_out, id)] = nobs ? total / (double)nobs : bad;
      }
    } else {
      const long npratio = npix_out / npix_in;
      for (long iu = 0; iu < npix_out; iu++) mo[ora_nest2ring(nside_out, iu)] = mi[ora_nest2ring(nside_in, iu / npratio)];
    }
  }
}
/* udgrade_rms, src/dang_util_mod.f90:341-356 */
void ora_udgrade_rms(const double *in, long nside_in, double *out, long nside_out, int nmaps) {
  const long npix_in = 12 * nside_in * nside_in, npix_out = 12 * nside_out * nside_out;
  double *buf = xcalloc((size_t)npix_in * nmaps, sizeof(double));
  for (size_t i = 0; i < (size_t)npix_in * nmaps; i++) buf[i] = in[i] * in[i];
  ora_udgrade_ring(buf, nside_in, out, nside_out, nmaps);
  for (size_t i = 0; i < (size_t)npix_out * nmaps; i++) out[i] = sqrt(out[i]) * ((double)nside_out * 1.0 / (double)nside_in);
  free(buf);
}
/* udgrade_mask, src/dang_util_mod.f90:358-376 */
void ora_udgrade_mask(const double *in, long nside_in, double *out, long nside_out, int nmaps, double threshold) {
  const long npix_out = 12 * nside_out * nside_out;
  ora_udgrade_ring(in, nside_in, out, nside_out, nmaps);
  if (nside_in > nside_out)
    for (size_t i = 0; i < (size_t)npix_out * nmaps; i++) out[i] = (out[i] < threshold) ? 0.0 : 1.0;
}
