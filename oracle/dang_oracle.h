/*
 * dang_oracle.h -- CPU restatement of hermda02/dang's per-Gibbs-iteration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dang_b200/ (the product) may include,
 * link, import or execute this file.  Allowed users: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 *  and cannot be compiled in this image (no Fortran compiler,
 * HEALPix, CFITSIO, MPI).  This oracle is pinned only by (i) line-by-line
 * correspondence with the cited reference source and (ii) analytic known-answer
 * tests in tests/test_oracle_kat.py.
 *
 * Conventions
 *   - All arithmetic is fp64, serial semantics of the source (OMP_NUM_THREADS=1,
 *     the reference's -O0 build flags: no FMA contraction, sequential sums).
 *   - Arrays use the reference's Fortran layout A(0:npix-1, nmaps, nbands), i.e.
 *     in C  A[(band*nmaps + k)*npix + pix]  with k = 0,1,2 for I,Q,U.
 *   - "map_n" / plane numbers in the API are the reference's 1-based values
 *     (1=I, 2=Q, 3=U, -1=Q+U, -2=I+Q+U); pixel indices are 0-based as in the source.
 *   - Random deviates are INPUTS (the reference calls RANDOM_SEED() with no
 *     argument, src/dang.f90:67, so its chain is irreproducible by construction).
 *     A "normal" deviate z stands for the value r*sin(theta) inside rand_normal
 *     (src/dang_util_mod.f90:100-110); rand_normal(mean,stdev) == mean + stdev*z.
 */
#ifndef DANG_ORACLE_H
#define DANG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MISSVAL (-1.6375e30) /* src/dang_util_mod.f90:19 */
#define ORA_MAXIND 2

/* component types, src/dang_component_mod.f90:791-809 */
enum { ORA_POWERLAW = 1, ORA_MBB = 2, ORA_FREEFREE = 3, ORA_LOGNORMAL = 4, ORA_CMB = 5, ORA_TEMPLATE = 6,
       ORA_T_CMB = 7, ORA_MONOPOLE = 8, ORA_HI_FIT = 9 };
/* *_LNL_TYPE, src/dang_sample_mod.f90:249-258 */
enum { ORA_LNL_CHISQ = 0, ORA_LNL_MARGINAL = 1, ORA_LNL_PRIOR = 2 };
/* *_PRIOR, src/dang_sample_mod.f90:260-266 */
enum { ORA_PRIOR_UNIFORM = 0, ORA_PRIOR_GAUSSIAN = 1, ORA_PRIOR_JEFFREYS = 2 };
/* ML_MODE, src/dang_cg_mod.f90:254,265 */
enum { ORA_OPTIMIZE = 0, ORA_SAMPLE = 1 };

typedef struct ora_state ora_state;

/* ---- construction (mirrors dang.f90:43-75 initialisation order) ---- */
ora_state *ora_create(int nside, int npix, int nmaps, int nbands, int ncomp);
void ora_destroy(ora_state *st);

/* bp(j): src/dang_bp_mod.f90:7-12,19-60.  n == 0  <=> id == 'delta'.
 * nu_c given in Hz or GHz (GHz -> Hz if < 1e9, :35-37); nu0 in GHz (x1e9, :138);
 * tau0 is normalised to unit sum here (normalize_bandpass, :62-81). */
int ora_set_band(ora_state *st, int band, double nu_c, int n, const double *nu0_ghz,
                 const double *tau);

/* ddata: sig_map, rms_map (npix*nmaps*nbands), masks plane 1 (npix), gain, offset (nbands).
 * The mask is normalised as read_data_maps does (missval -> 0, :153-161). */
int ora_set_maps(ora_state *st, const double *sig, const double *rms, const double *mask,
                 const double *gain, const double *offset);
void ora_set_pol_type(ora_state *st, int lo, int hi); /* ddata%pol_type(1), (size) */

/* component_list(ic): type dang_comps, src/dang_component_mod.f90:12-55 */
int ora_set_component(ora_state *st, int ic, int type, const char *label, double nu_ref,
                      int cg_group, int sample_amplitude, const double *amplitude /*npix*nmaps*/,
                      const double *indices /*npix*nmaps*nindices*/);
/* type 'template' (src/dang_component_mod.f90:536-577); template_map [nmaps][npix] is normalised by its
 * maximum per plane inside, as the constructor does; template_amplitudes [nmaps][nbands] */
int ora_set_template(ora_state *st, int ic, const double *template_map, const double *template_amplitudes,
                     const int *corr, int nfit);
double *ora_template_map(ora_state *st, int ic);        /* [nmaps][npix], normalised */
double *ora_template_amplitudes(ora_state *st, int ic); /* [nmaps][nbands] */
int ora_set_index(ora_state *st, int ic, int nind /*0-based*/, int sample_index, int index_mode,
                  int lnl_type, int prior_type, const double gauss[2], const double uni[2],
                  double step_size, int tuned, int sample_nside, const int *pol_flags, int nflag);

double *ora_amplitude(ora_state *st, int ic);          /* [nmaps][npix] */
double *ora_indices(ora_state *st, int ic);            /* [nindices][nmaps][npix] */
double *ora_sky_model(ora_state *st);                  /* [nbands][nmaps][npix] */
double *ora_res_map(ora_state *st);
double *ora_chi_map(ora_state *st);                    /* [nmaps][npix] */
double ora_step_size(ora_state *st, int ic, int nind);
int ora_nindices(ora_state *st, int ic);
void ora_set_gain(ora_state *st, int band, double g);

/* ---- SED / signal: src/dang_component_mod.f90:754-813, 886-1040 ---- */
double ora_eval_sed(const ora_state *st, int ic, int band, int pix, int map_n, const double *theta);
double ora_eval_signal(const ora_state *st, int ic, int band, int pix, int map_n,
                       const double *theta);

/* ---- amplitude draw: src/dang_cg_mod.f90 ---- */
typedef struct ora_cg ora_cg;
ora_cg *ora_cg_create(ora_state *st, int cg_group, int i_max, double converge,
                      const int *pol_flags, int nflag);
void ora_cg_destroy(ora_cg *g);
long ora_cg_n(const ora_cg *g, int flag_n);            /* length of b / x            */
long ora_cg_m(const ora_cg *g, int flag_n);            /* length of eta (S*npix)     */
double *ora_cg_x(ora_cg *g, int flag_n);               /* saved x (Q10)              */
void ora_compute_rhs(ora_cg *g, int flag_n, double *b);                     /* :326-596  */
void ora_compute_Ax(ora_cg *g, const double *x, int flag_n, double *res);   /* :598-911  */
void ora_compute_sample_vector(ora_cg *g, const double *eta, int flag_n,
                               double *res, int fix_q1);                    /* :913-1100 */
/* cg_search :179-324.  eta == NULL or ml_mode == ORA_OPTIMIZE skips the fluctuation.
 * Returns the final value of the loop counter i (the reference prints it). */
int ora_cg_search(ora_cg *g, int flag_n, const double *b, int ml_mode, const double *eta,
                  int fix_q1, double *delta_final, double *delta_trace, int trace_len);
void ora_unpack_amplitudes(ora_cg *g, int flag_n);                          /* :1284-1396 */
/* sample_cg_groups body for one group (:166-172): rhs -> cg -> unpack per flag, then sky model */
int ora_sample_cg_group(ora_cg *g, int ml_mode, const double *eta, int fix_q1, int *niter,
                        double *delta_final);

/* ---- data object: src/dang_data_mod.f90:339-396, 494-526 ---- */
void ora_update_sky_model(ora_state *st);
double ora_compute_chisq(ora_state *st, double *chi_sum_planes /*nmaps, un-normalised sums*/);
double ora_mask_avg(const ora_state *st, int ic, int nind, int map_n);      /* util :186-206 */

/* ---- likelihood: src/dang_lnl_mod.f90 ---- */
double ora_evaluate_lnL(const ora_state *st, const double *data, const double *rms,
                        const double *model, const int map_inds[2], int pixel,
                        const double *mask);                                /* :126-182 */
double ora_evaluate_marginal_lnL(const ora_state *st, const double *data, const double *rms,
                                 const double *model, const int map_inds[2], int pixel);
double ora_eval_normal_prior(double prop, double mean, double std);         /* util :112-121 */
double ora_rand_normal_from_uniform(double mean, double stdev, double u1, double u2); /* :100-110 */

/* ---- spectral-parameter draw: src/dang_sample_mod.f90:88-485 ----
 * z, u: injected deviates, slot-indexed.  Full-sky: z[l], u[l].  Per-pixel:
 * z[l*npix + pix], u[l*npix + pix].  A slot is consumed only if the reference
 * would have drawn it (Q5); unused slots are ignored.
 * accept_out (optional): number of accepted proposals (full-sky) or total over pixels.
 * decisions (optional): full-sky: nsample bytes; per-pixel: nsample*npix bytes,
 *   0 = rejected, 1 = accepted, 2 = out of bounds (no uniform drawn), 3 = pixel masked.
 * lnl_trace (optional): the lnl_new of every evaluated proposal, same indexing (NaN if none). */
int ora_sample_index_mh(ora_state *st, int ic, int nind, int map_n, int nsample, int ml_mode,
                        const double *z, const double *u, double *accept_out,
                        unsigned char *decisions, double *lnl_trace);
/* sample_spectral_parameters (:21-86): loops comps / indices / flags in reference order.
 * Deviate arrays are consumed call by call: call number q uses z + q*stride, u + q*stride
 * with stride = nsample*npix. Returns the number of sample_index_mh calls made. */
int ora_sample_spectral_parameters(ora_state *st, int nsample, int ml_mode, const double *z,
                                   const double *u);
/* tune_spectral_parameter_length (:623-717), full-sky, driven by slot-indexed deviates
 * z[blk*nsample + l]; stops after max_blocks.  Returns number of blocks run. */
int ora_tune_step(ora_state *st, int ic, int nind, int map_n, int nsample, int ml_mode,
                  const double *z, const double *u, int max_blocks);

/* ---- band calibration: src/dang_sample_mod.f90:570-621 (Stokes I) ---- */
double ora_fit_band_gain(ora_state *st, int map_n, int band, int ml_mode, double z);

/* ---- counter-based RNG shared by definition with the device path (Philox4x32-10) ----
 * Stream definition (DESIGN.md "RNG"): uniform pair for (seed, stream, slot) is taken from
 * philox4x32_10(counter = {slot_lo, slot_hi, stream, 0x44414e47}, key = {seed_lo, seed_hi});
 * u1 = (x0*2^32 + x1 + 0.5) * 2^-64 ... see ora_philox_uniform2. */
void ora_philox_raw(unsigned int ctr[4], unsigned int k0, unsigned int k1); /* Random123 KATs */
void ora_philox_uniform2(unsigned long long seed, unsigned int stream, unsigned long long slot,
                         double *u1, double *u2);
void ora_philox_normals(unsigned long long seed, unsigned int stream, unsigned long long slot0,
                        long n, double *z);
void ora_philox_uniforms(unsigned long long seed, unsigned int stream, unsigned long long slot0,
                         long n, double *u);

int ora_num_threads(void);
long ora_nest2ring(long nside, long pix);
long ora_ring2nest(long nside, long pix);
void ora_udgrade_ring(const double *in, long nside_in, double *out, long nside_out, int nmaps);
void ora_udgrade_rms(const double *in, long nside_in, double *out, long nside_out, int nmaps);
void ora_udgrade_mask(const double *in, long nside_in, double *out, long nside_out, int nmaps, double threshold);
int ora_tune_step_from(ora_state *st, int ic, int nind, int map_n, int nsample, int ml_mode, const double *z,
                       const double *u, int max_blocks, const double *theta_init);
void ora_perpixel_tune_start(const ora_state *st, int ic, int nind, int map_n, double *theta_init);
double ora_get_T_CMB(void);
void ora_set_T_CMB(double t);
double *ora_offset(ora_state *st);
void ora_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
