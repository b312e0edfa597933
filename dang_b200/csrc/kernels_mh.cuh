// kernels_mh.cuh -- spectral-parameter draw: sample_index_mh,
// src/dang_sample_mod.f90:88-485, with update_sample_model (:520-568), evaluate_lnL /
// evaluate_marginal_lnL / eval_jeffreys_prior (src/dang_lnl_mod.f90) and eval_normal_prior
// (src/dang_util_mod.f90:112-121) fused in.
//
// Per-pixel mode (K5): one thread owns one pixel's chain; the pixel's residual data, rms and
// the proposal's SED live in shared memory; no communication.
// Full-sky mode (K4): the chain itself runs in a single-thread scalar kernel on the device;
// the likelihood comes either from per-band sufficient statistics gathered in ONE pass over the
// maps (default) or from one streaming pass per proposal (the reference's structure).
#pragma once
#include "common.cuh"

// log(eval_normal_prior(x, mean, std)), src/dang_util_mod.f90:112-121 + dang_sample_mod.f90:261
__device__ __forceinline__ double log_normal_prior(double x, double mean, double sd) {
  const double var = sd * sd;
  const double num = exp(-((x - mean) * (x - mean)) / (2 * var));
  const double denom = sd * sqrt(2.0 * DG_PI);
  return log(num / denom);
}

// data_raw(i,k,j) = sig - sum_{c2 /= c} signal_c2, src/dang_sample_mod.f90:173-196
__device__ __forceinline__ double mh_data_value(const ModelView &mv, int ic, int j, int k,
                                                int64_t p) {
  const size_t off = plane_off(mv, j, k) + p;
  double v = ldg_stream(mv.sig + off);
  if (k == 0) v = (v - mv.offset[j]) / mv.gain[j];
  const size_t kp = (size_t)k * mv.Ppad + p;
  for (int c2 = 0; c2 < mv.ncomp; c2++) {
    if (c2 == ic) continue;
    const CompView &cc = mv.comp[c2];
    const double t0 = cc.nind > 0 ? cc.idx[0][kp] : 0.0;
    const double t1 = cc.nind > 1 ? cc.idx[1][kp] : 0.0;
    v = v - cc.amp[kp] * sed_eval(mv, c2, k, j, t0, t1);
  }
  return v;
}

// ---------------------------------------------------------------- K5 (strict order): per-pixel chains
// Selected by DANG_OPT_PERPIXEL_SERIAL: lnL accumulated in exactly the reference's order.
// One thread owns one pixel's chain (:353-470).  Shared memory per thread: the residual data
// sD[B][S], the inverse noise sW[B][S], the proposal's SED sS[B] and sF[B], the factor of the SED
// that does not depend on the index being sampled:
//     power-law            sed_j = exp((beta) L_j)                      F_j = 1
//     mbb, beta sampled    sed_j = [eref/(exp(z nu_j)-1)] exp((beta+1) L_j)   F_j = Planck ratio (T fixed)
//     mbb, T sampled       sed_j = eref(T)/(exp(z nu_j)-1) [exp((beta+1) L_j)] F_j = power law (beta fixed)
// so a proposal costs one exp per band (+1 for T) instead of three transcendentals; the products
// are formed in the same order as sed_mbb, i.e. the same bits.  Bandpass-integrated bands keep
// the full sum over bandpass samples.  This kernel is FP64-pipe bound (nsample*B exp per pixel
// against 2*B*S*8 bytes of traffic), not HBM bound.
__device__ __forceinline__ double mh_fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

static __global__ void __launch_bounds__(DG_MH_THREADS)
mh_perpixel_serial_kernel(const ModelView mv, const MhView mh, double *partials, unsigned int *ticket,
                          double *out) {
  extern __shared__ double dyn[];
  const int T = DG_MH_THREADS, tid = threadIdx.x, B = mv.nbands, S = mh.S;
  double *sD = dyn, *sW = dyn + (size_t)B * S * T, *sS = dyn + (size_t)2 * B * S * T,
         *sF = dyn + (size_t)(2 * S + 1) * B * T;
  __shared__ double smem[32];
  double acc[1] = {0.0};
  const CompView &cv = mv.comp[mh.ic];
  const SedTable &tab = *mv.tab;
  const bool is_mbb = cv.type == 2;
  const bool is_pl = cv.type == 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + tid; p < mv.P; p += stride) {
    const uint64_t gpix = (uint64_t)(mv.pix_lo + p);
    if (!mv.mask[p]) {  // :362; index_map stays 0 for masked pixels (:223, :465, :483)
      for (int s = 0; s < S; s++) cv.idx[mh.nind][(size_t)mh.plane[s] * mv.Ppad + p] = 0.0;
      if (mh.decisions)
        for (int l = 0; l < mh.nsample; l++) mh.decisions[(size_t)l * mv.P + p] = 3;
      continue;
    }
    const size_t kp0 = (size_t)mh.plane[0] * mv.Ppad + p;
    double sample[DG_MAXIND], theta[DG_MAXIND], amp[2];
    sample[0] = cv.nind > 0 ? cv.idx[0][kp0] : 0.0;  // :372-374
    sample[1] = cv.nind > 1 ? cv.idx[1][kp0] : 0.0;
    theta[0] = sample[0];
    theta[1] = sample[1];
    for (int s = 0; s < S; s++) amp[s] = cv.amp[(size_t)mh.plane[s] * mv.Ppad + p];
    for (int j = 0; j < B; j++)
      for (int s = 0; s < S; s++) {
        sD[((size_t)j * S + s) * T + tid] = mh_data_value(mv, mh.ic, j, mh.plane[s], p);
        sW[((size_t)j * S + s) * T + tid] = ldg_stream(mv.rms + plane_off(mv, j, mh.plane[s]) + p);
      }
    // the factor of each delta band's SED that stays fixed during this chain
    if (is_mbb) {
      if (mh.nind == 0) {
        const double z = DG_H / (DG_KB * sample[1]);
        const double eref = exp(z * cv.nu_ref) - 1.0;
        for (int j = 0; j < B; j++)
          if (mv.band[j].n == 0) sF[(size_t)j * T + tid] = eref / (exp(z * mv.band[j].nu_c) - 1.0);
      } else {
        for (int j = 0; j < B; j++)
          if (mv.band[j].n == 0)
            sF[(size_t)j * T + tid] = exp_scaled(sample[0] + 1.0, tab.lnr_hi[mh.ic][j], tab.lnr_lo[mh.ic][j]);
      }
    }

    // SED of the proposal, band by band, into sS
    auto eval_sed = [&](const double *th) {
      if (is_pl) {
#pragma unroll 4
        for (int j = 0; j < B; j++)
          sS[(size_t)j * T + tid] = sed_powerlaw(mv, mh.ic, j, th[0]);
      } else if (!is_mbb) {
        for (int j = 0; j < B; j++) sS[(size_t)j * T + tid] = sed_theta(mv, mh.ic, j, th[0], th[1], mh.plane[0]);
      } else if (mh.nind == 0) {
#pragma unroll 4
        for (int j = 0; j < B; j++)
          sS[(size_t)j * T + tid] =
              mv.band[j].n == 0 ? sF[(size_t)j * T + tid] * exp_scaled(th[0] + 1.0, tab.lnr_hi[mh.ic][j], tab.lnr_lo[mh.ic][j])
                                : sed_mbb(mv, mh.ic, j, th[0], th[1]);
      } else {
        const double z = DG_H / (DG_KB * th[1]);
        const double eref = exp(z * cv.nu_ref) - 1.0;
#pragma unroll 4
        for (int j = 0; j < B; j++)
          sS[(size_t)j * T + tid] = mv.band[j].n == 0
                                        ? eref / (exp(z * mv.band[j].nu_c) - 1.0) * sF[(size_t)j * T + tid]
                                        : sed_mbb(mv, mh.ic, j, th[0], th[1]);
      }
    };
    // lnL of theta; order of accumulation as evaluate_lnL: Stokes outer, band inner (:172-176)
    auto eval_lnl = [&](const double *th) -> double {
      eval_sed(th);
      double lnl = 0.0;
      if (mh.lnl_type == 0) {
        for (int s = 0; s < S; s++)
#pragma unroll 4
          for (int j = 0; j < B; j++) {
            const double model = amp[s] * sS[(size_t)j * T + tid];  // eval_signal :773
            const double t = (sD[((size_t)j * S + s) * T + tid] - model) / sW[((size_t)j * S + s) * T + tid];
            lnl = lnl - 0.5 * (t * t);
          }
      } else {  // marginal, src/dang_lnl_mod.f90:113-122 (band outer, Stokes inner)
        for (int j = 0; j < B; j++)
          for (int s = 0; s < S; s++) {
            const double model = amp[s] * sS[(size_t)j * T + tid];
            const double rms = sW[((size_t)j * S + s) * T + tid];
            const double TN = model / (rms * rms);
            const double TNd = TN * sD[((size_t)j * S + s) * T + tid];
            const double TNT = TN * model;
            const double invTNT = 1.0 / TNT;
            lnl = lnl - 0.5 * TNd * invTNT * TNd;
          }
      }
      return lnl;
    };
    auto eval_prior = [&](double val, double prev) -> double {
      if (mh.prior_type == 1) return log_normal_prior(val, mh.gauss[0], mh.gauss[1]);
      if (mh.prior_type == 2) {  // eval_jeffreys_prior, src/dang_lnl_mod.f90:242-304
        double sum = 0.0;
        if (mh.is_synch) {
          for (int s = 0; s < S; s++)
            for (int j = 0; j < B; j++) {
              const double ss = amp[s] * sed_theta(mv, mh.ic, j, val, 0.0, mh.plane[0]);
              const double ir = 1.0 / sW[((size_t)j * S + s) * T + tid];
              const double t = (ir * ir) * (ss / amp[s]) * log(mv.band[j].nu_c / cv.nu_ref);
              sum = sum + t * t;
            }
        }
        return log(sqrt(sum));
      }
      if (mh.prior_type == 0) return 0.0;
      return prev;
    };

    double lnl = 0.0, lnl_prior = 0.0, lnl_old, lnl_new;
    bool sample_it = true;
    if (mh.lnl_type == 2) {  // 'prior': draw from the Gaussian prior (:389-391)
      sample_it = false;
      const double zz = mh.z ? mh.z[p] : philox_normal(mh.seed, DG_STREAM_MH_Z, gpix);
      sample[mh.nind] = mh.gauss[0] + mh.gauss[1] * zz;
    } else {
      lnl = eval_lnl(sample);
    }
    lnl_prior = eval_prior(sample[mh.nind], lnl_prior);
    lnl_old = lnl + lnl_prior;
    if (sample_it) {
      for (int l = 0; l < mh.nsample; l++) {  // :410-455
        const size_t slot = (size_t)l * mv.P + p;
        const uint64_t gslot = (uint64_t)l * (uint64_t)mv.npix + gpix;
        const double zz = mh.z ? mh.z[slot] : philox_normal(mh.seed, DG_STREAM_MH_Z, gslot);
        theta[mh.nind] = sample[mh.nind] + (0.0 + mh.step * zz);
        if (theta[mh.nind] < mh.uni[0] || theta[mh.nind] > mh.uni[1]) {  // :415, Q5
          if (mh.decisions) mh.decisions[slot] = 2;
          continue;
        }
        lnl = eval_lnl(theta);
        lnl_prior = eval_prior(theta[mh.nind], lnl_prior);
        lnl_new = lnl + lnl_prior;
        if (mh.lnl_trace) mh.lnl_trace[slot] = lnl_new;
        const double diff = lnl_new - lnl_old;
        bool accept;
        if (mh.ml_mode == 0) {
          accept = diff > 0.0;
        } else {
          double uu;
          if (mh.u) {
            uu = mh.u[slot];
          } else {
            double u2;
            philox_uniform2(mh.seed, DG_STREAM_MH_U, gslot, uu, u2);
          }
          accept = diff > log(uu);  // :450, Q4
        }
        if (accept) {
          sample[mh.nind] = theta[mh.nind];
          lnl_old = lnl_new;
          acc[0] += 1.0;
        }
        if (mh.decisions) mh.decisions[slot] = accept ? 1 : 0;
      }
    }
    for (int s = 0; s < S; s++)  // :465, :483
      cv.idx[mh.nind][(size_t)mh.plane[s] * mv.Ppad + p] = sample[mh.nind];
  }
  grid_reduce<1>(acc, smem, partials, ticket, out);
}

// ---------------------------------------------------------------- K5 (default): per-pixel chains
// chisq likelihood with uniform / Gaussian prior (every BASELINE config); the rare variants
// (marginal lnL, 'prior' draws, Jeffreys prior) run on the strict kernel above.
//
// DG_MH_LANES lanes cooperate on one pixel's chain: lane r owns bands r, r+L, r+2L, ... with their
// residual data, inverse noise, ln(nu/nu_ref) and fixed SED factor in REGISTERS (BPL = bands per
// lane and the SED form are template parameters), evaluates its share of the proposal's SED and
// chi-square, and the group forms lnL with a fixed-order butterfly.  The chain's deviates
// (z_l, ln u_l) for the warp's 8 pixels are generated by all 32 lanes and parked in shared memory.
// Compared with the strict kernel the numerical differences are the (fixed, tree) summation order
// over bands, 1/sigma by multiplication and the Gaussian log-prior evaluated as
// -(x-mu)^2/(2 sigma^2) - ln(sigma sqrt(2 pi)): lnL agrees to ~1e-16 relative, so a decision can
// differ from the reference's only when |diff - ln u| < ~1e-14.
#define DG_MH_LANES 4
enum { MH_SED_POWERLAW = 0, MH_SED_MBB_BETA = 1, MH_SED_MBB_T = 2, MH_SED_GENERIC = 3,
       MH_SED_BP_POWERLAW = 4, MH_SED_BP_MBB_BETA = 5 };

// Tabulated bandpasses (MH_SED_BP_*): the bandpass-integrated SED of a proposal beta = beta_0 + delta is
//     sum_i w_i exp(delta L_i),  w_i = tau_i [Planck ratio_i] exp((beta_0 [+1]) L_i),  L_i = ln(nu_i / nu_ref),
// and across one band L_i stays within ~0.15 of the band centre's L_c, so
//     = exp(delta L_c) sum_k (delta^k / k!) m_k,   m_k = sum_i w_i (L_i - L_c)^k.
// The moments m_0..m_8 are formed ONCE per pixel and band at the chain's first point (n_bp exp, the summand
// of evaluate_powerlaw / evaluate_mbb, src/dang_component_mod.f90:909-914,949-955, in its order, so m_0 is
// the reference's SED there bit for bit); a proposal then costs one exp and 8 FMAs per band instead of
// n_bp exp -- 128x fewer transcendentals at config c3's n_bp = 128.  The series is cut at
// |delta| max|L_i - L_c| <= 0.1 (remainder < 3e-15 relative, below the rounding of the n_bp-term sum it
// replaces); beyond that the proposal is summed directly.
#define DG_MH_KM 8

template <int BPL, int MODE>
__global__ void __launch_bounds__(DG_MH_THREADS)
mh_perpixel_kernel(const ModelView mv, const MhView mh, double *partials, unsigned int *ticket,
                   double *out) {
  constexpr int L = DG_MH_LANES;
  constexpr int PB = DG_MH_THREADS / L;  // pixels per block iteration
  constexpr int PW = 32 / L;             // pixels per warp
  extern __shared__ double dyn[];        // zs[PB][nsample], lus[PB][nsample]
  __shared__ double smem[32];
  const int tid = threadIdx.x, lane = tid & 31, r = tid % L, g = tid / L, B = mv.nbands, S = mh.S;
  const int nsample = mh.nsample;
  const unsigned full = 0xffffffffu;
  double *zs = dyn + (size_t)g * nsample, *lus = dyn + (size_t)(PB + g) * nsample;
  double *wzs = dyn + (size_t)(g - lane / L) * nsample;  // first pixel of this warp
  double *wlus = wzs + (size_t)PB * nsample;
  constexpr bool BPM = MODE == MH_SED_BP_POWERLAW || MODE == MH_SED_BP_MBB_BETA;
  constexpr bool PLAW = MODE == MH_SED_POWERLAW || MODE == MH_SED_BP_POWERLAW;
  constexpr bool MBBB = MODE == MH_SED_MBB_BETA || MODE == MH_SED_BP_MBB_BETA;
  double *mks = dyn + (size_t)2 * PB * nsample + tid;  // BPM: m_k / k! and max|L_i - L_c| per band, [.][threads]
  double acc[1] = {0.0};
  const CompView &cv = mv.comp[mh.ic];
  const SedTable &tab = *mv.tab;
  const double nu_ref = cv.nu_ref;
  const double ln_denom = log(mh.gauss[1] * sqrt(2.0 * DG_PI));
  const double inv2var = 1.0 / (2 * (mh.gauss[1] * mh.gauss[1]));

  // per-lane band constants
  double Lh[BPL], Ll[BPL], nuc[BPL];
  bool bp[BPL];
#pragma unroll
  for (int i = 0; i < BPL; i++) {
    const int j = r + i * L;
    Lh[i] = j < B ? tab.lnr_hi[mh.ic][j] : 0.0;
    Ll[i] = j < B ? tab.lnr_lo[mh.ic][j] : 0.0;
    nuc[i] = j < B ? mv.band[j].nu_c : 1.0;
    bp[i] = j < B && mv.band[j].n != 0;
  }

  const int64_t ngroups = (int64_t)gridDim.x * PB;
  const int64_t niter = (mv.P + ngroups - 1) / ngroups;
  for (int64_t itp = 0; itp < niter; itp++) {
    const int64_t p = itp * ngroups + (int64_t)blockIdx.x * PB + g;
    const bool valid = p < mv.P;
    const bool use = valid && mv.mask[p] != 0;
    if (!use && valid && r == 0) {  // :362; index_map stays 0 for masked pixels (:223, :465, :483)
      for (int s = 0; s < S; s++) cv.idx[mh.nind][(size_t)mh.plane[s] * mv.Ppad + p] = 0.0;
      if (mh.decisions)
        for (int l = 0; l < nsample; l++) mh.decisions[(size_t)l * mv.P + p] = 3;
    }
    // a warp leaves only if all of its 8 pixels are idle; otherwise idle groups run along on
    // pixel 0's data so the shuffles stay convergent (use == false guards every store)
    if (!__any_sync(full, use)) continue;
    const int64_t pp = use ? p : 0;
    const size_t kp0 = (size_t)mh.plane[0] * mv.Ppad + pp;
    const double idx0 = cv.nind > 0 ? cv.idx[0][kp0] : 0.0;  // :372-374
    const double idx1 = cv.nind > 1 ? cv.idx[1][kp0] : 0.0;
    double cur = mh.nind == 0 ? idx0 : idx1;
    const double amp0 = cv.amp[(size_t)mh.plane[0] * mv.Ppad + pp];
    const double amp1 = S > 1 ? cv.amp[(size_t)mh.plane[1] * mv.Ppad + pp] : 0.0;

    double D0[BPL], D1[BPL], W0[BPL], W1[BPL], F[BPL];
#pragma unroll
    for (int i = 0; i < BPL; i++) {
      const int j = r + i * L;
      D0[i] = D1[i] = W0[i] = W1[i] = 0.0;
      F[i] = 1.0;
      if (j < B) {
        D0[i] = mh_data_value(mv, mh.ic, j, mh.plane[0], pp);
        W0[i] = 1.0 / ldg_stream(mv.rms + plane_off(mv, j, mh.plane[0]) + pp);
        if (S > 1) {
          D1[i] = mh_data_value(mv, mh.ic, j, mh.plane[1], pp);
          W1[i] = 1.0 / ldg_stream(mv.rms + plane_off(mv, j, mh.plane[1]) + pp);
        }
        if (MBBB) {
          const double z = DG_H / (DG_KB * idx1);
          F[i] = (exp(z * nu_ref) - 1.0) / (exp(z * nuc[i]) - 1.0);
        } else if (MODE == MH_SED_MBB_T) {
          F[i] = exp_scaled(idx0 + 1.0, Lh[i], Ll[i]);
        }
        if (BPM && bp[i]) {  // moments of the bandpass-integrated SED about the chain's first point
          const BandView &bv = mv.band[j];
          const double *lh = mv.bp_lnr_hi + (size_t)mh.ic * mv.nbp + bv.off, *ll = mv.bp_lnr_lo + (size_t)mh.ic * mv.nbp + bv.off;
          const double *nu0 = mv.bp_nu0 + bv.off, *tau = mv.bp_tau0 + bv.off;
          const double z = MBBB ? DG_H / (DG_KB * idx1) : 0.0;
          const double eref = MBBB ? exp(z * nu_ref) - 1.0 : 0.0;
          double m[DG_MH_KM + 1], dlmax = 0.0;
#pragma unroll
          for (int k = 0; k <= DG_MH_KM; k++) m[k] = 0.0;
          for (int q = 0; q < bv.n; q++) {
            if (nu0[q] == 0.0) continue;  // :911, :951
            double w;
            if (PLAW) w = tau[q] * exp_scaled(cur, lh[q], ll[q]);
            else w = tau[q] * eref / (exp(z * nu0[q]) - 1.0) * exp_scaled(cur + 1.0, lh[q], ll[q]);
            const double dl = (lh[q] - Lh[i]) + (ll[q] - Ll[i]);
            dlmax = fmax(dlmax, fabs(dl));
            double pw = w;
#pragma unroll
            for (int k = 0; k <= DG_MH_KM; k++) {
              m[k] = m[k] + pw;   // m[0] accumulates in the reference's order: the reference's SED at `cur`
              pw *= dl;
            }
          }
          double fact = 1.0;
#pragma unroll
          for (int k = 0; k <= DG_MH_KM; k++) {
            if (k > 1) fact *= (double)k;
            mks[((size_t)i * (DG_MH_KM + 2) + k) * DG_MH_THREADS] = m[k] / fact;
          }
          mks[((size_t)i * (DG_MH_KM + 2) + DG_MH_KM + 1) * DG_MH_THREADS] = dlmax;
        }
      }
    }
    const double theta_ref = cur;
    // deviates of this warp's PW chains, slot-indexed (Q5), generated by all 32 lanes
    __syncwarp();
    {
      const int64_t p_first = itp * ngroups + (int64_t)blockIdx.x * PB + (g - lane / L);
      for (int q = lane; q < PW * nsample; q += 32) {
        const int gq = q / nsample, l = q - gq * nsample;
        int64_t pq = p_first + gq;
        if (pq >= mv.P) pq = 0;
        const size_t slot = (size_t)l * mv.P + pq;
        const uint64_t gslot = (uint64_t)l * (uint64_t)mv.npix + (uint64_t)(mv.pix_lo + pq);
        wzs[(size_t)gq * nsample + l] = mh.z ? mh.z[slot] : philox_normal(mh.seed, DG_STREAM_MH_Z, gslot);
        double uu = 1.0, u2;
        if (mh.ml_mode != 0) {
          if (mh.u) uu = mh.u[slot];
          else philox_uniform2(mh.seed, DG_STREAM_MH_U, gslot, uu, u2);
        }
        wlus[(size_t)gq * nsample + l] = log(uu);  // :450
      }
    }
    __syncwarp();

    double lnl_old = 0.0, naccept = 0.0;
    for (int l = -1; l < nsample; l++) {  // l = -1: lnL of the starting point (:380-402)
      double x = cur;
      bool oob = false;
      if (l >= 0) {
        x = cur + (0.0 + mh.step * zs[l]);             // :414
        oob = x < mh.uni[0] || x > mh.uni[1];          // :415, Q5
      }
      const double xe = oob ? cur : x;  // out-of-bounds groups evaluate a harmless value, discarded
      // ---- this lane's share of lnL(xe)
      double zT = 0.0, eref = 0.0;
      if (MODE == MH_SED_MBB_T) {
        zT = DG_H / (DG_KB * xe);
        eref = exp(zT * nu_ref) - 1.0;
      }
      double part = 0.0;
#pragma unroll
      for (int i = 0; i < BPL; i++) {
        const int j = r + i * L;
        if (j < B) {
          double sed;
          if (BPM && bp[i]) {
            const double *mk = mks + (size_t)i * (DG_MH_KM + 2) * DG_MH_THREADS;
            const double dlt = xe - theta_ref;
            if (fabs(dlt) * mk[(DG_MH_KM + 1) * DG_MH_THREADS] <= 0.1) {
              double poly = mk[DG_MH_KM * DG_MH_THREADS];
#pragma unroll
              for (int k = DG_MH_KM - 1; k >= 0; k--) poly = fma(poly, dlt, mk[k * DG_MH_THREADS]);
              sed = exp_scaled(dlt, Lh[i], Ll[i]) * poly;
            } else {
              sed = sed_theta(mv, mh.ic, j, xe, idx1, mh.plane[0]);  // far from the first point: sum the bandpass directly
            }
          } else if (PLAW) sed = exp_scaled(xe, Lh[i], Ll[i]);
          else if (MBBB) sed = F[i] * exp_scaled(xe + 1.0, Lh[i], Ll[i]);
          else if (MODE == MH_SED_MBB_T) sed = eref * mh_fast_rcp(exp(zT * nuc[i]) - 1.0) * F[i];
          else sed = sed_theta(mv, mh.ic, j, mh.nind == 0 ? xe : idx0, mh.nind == 0 ? idx1 : xe, mh.plane[0]);
          const double t0 = (D0[i] - amp0 * sed) * W0[i];
          part = part - 0.5 * (t0 * t0);
          if (S > 1) {
            const double t1 = (D1[i] - amp1 * sed) * W1[i];
            part = part - 0.5 * (t1 * t1);
          }
        }
      }
      part += __shfl_xor_sync(full, part, 1);
      part += __shfl_xor_sync(full, part, 2);
      double prior = 0.0;
      if (mh.prior_type == 1) {
        const double a = ((xe - mh.gauss[0]) * (xe - mh.gauss[0])) * inv2var;
        prior = a < 700.0 ? -a - ln_denom : log_normal_prior(xe, mh.gauss[0], mh.gauss[1]);
      }
      const double lnl_new = part + prior;
      if (l < 0) {
        lnl_old = lnl_new;
        continue;
      }
      const size_t slot = (size_t)l * mv.P + pp;
      if (oob) {
        if (use && r == 0 && mh.decisions) mh.decisions[slot] = 2;
        continue;
      }
      const double diff = lnl_new - lnl_old;
      const bool accept = (mh.ml_mode == 0) ? (diff > 0.0) : (diff > lus[l]);  // :444, :450 (Q4)
      if (accept) {
        cur = x;
        lnl_old = lnl_new;
        naccept += 1.0;
      }
      if (use && r == 0) {
        if (mh.lnl_trace) mh.lnl_trace[slot] = lnl_new;
        if (mh.decisions) mh.decisions[slot] = accept ? 1 : 0;
      }
    }
    if (use && r == 0) {
      for (int s = 0; s < S; s++) cv.idx[mh.nind][(size_t)mh.plane[s] * mv.Ppad + p] = cur;  // :465, :483
      acc[0] += naccept;
    }
  }
  grid_reduce<1>(acc, smem, partials, ticket, out);
}

// ---------------------------------------------------------------- K4: full-sky chain
// start of the chain: sample(:) = c%indices(0, map_inds(1), :) (:240-243), broadcast from the
// rank that owns global pixel 0 (gathered[0..1] of rank 0)
static __global__ void mh_fullsky_init_kernel(const ModelView mv, const MhView mh, MhScalars *ms,
                                       const double *gathered, int cnt) {
  ms->sample[0] = gathered[0];
  ms->sample[1] = gathered[1];
  ms->theta[0] = ms->sample[0];
  ms->theta[1] = ms->sample[1];
  ms->lnl_old = 0.0;
  ms->accept = 0.0;
  ms->l = 0;
  ms->phase = 0;
  ms->skip = 0;
  for (int j = 0; j < mv.nbands; j++) {
    const double s = sed_theta(mv, mh.ic, j, ms->sample[0], ms->sample[1], mh.plane[0]);
    ms->sed[j] = s;
    ms->s0[j] = s;
  }
  (void)cnt;
}

// local values of the index maps at this handle's first pixel (rank 0 owns global pixel 0)
static __global__ void mh_first_pixel_kernel(const ModelView mv, const MhView mh, double *out) {
  const CompView &cv = mv.comp[mh.ic];
  const size_t kp0 = (size_t)mh.plane[0] * mv.Ppad;
  out[0] = cv.nind > 0 ? cv.idx[0][kp0] : 0.0;
  out[1] = cv.nind > 1 ? cv.idx[1][kp0] : 0.0;
}

__device__ __forceinline__ double mh_draw_z(const MhView &mh, int l) {
  return mh.z ? mh.z[l]
              : philox_normal(mh.seed, mh.rng_z_stream ? (uint32_t)mh.rng_z_stream : (uint32_t)DG_STREAM_MH_Z,
                              (uint64_t)(mh.rng_slot0 + l));
}
__device__ __forceinline__ double mh_draw_u(const MhView &mh, int l) {
  if (mh.u) return mh.u[l];
  double u1, u2;
  philox_uniform2(mh.seed, mh.rng_u_stream ? (uint32_t)mh.rng_u_stream : (uint32_t)DG_STREAM_MH_U,
                  (uint64_t)(mh.rng_slot0 + l), u1, u2);
  return u1;
}

// advance to the next in-bounds proposal (out-of-bounds ones consume z but not u, Q5)
__device__ __forceinline__ void mh_next_proposal(const ModelView &mv, const MhView &mh,
                                                 MhScalars *ms) {
  while (ms->l < mh.nsample) {
    const double th = ms->sample[mh.nind] + (0.0 + mh.step * mh_draw_z(mh, ms->l));  // :286
    ms->theta[mh.nind] = th;
    if (th < mh.uni[0] || th > mh.uni[1]) {  // :287
      if (mh.decisions) mh.decisions[ms->l] = 2;
      ms->l++;
      continue;
    }
    for (int j = 0; j < mv.nbands; j++) ms->sed[j] = sed_theta(mv, mh.ic, j, ms->theta[0], ms->theta[1], mh.plane[0]);
    return;
  }
  ms->skip = 1;
}

__device__ __forceinline__ void mh_accept_step(const MhView &mh, MhScalars *ms, double lnl, double prior) {
  const double lnl_new = lnl + prior;  // :306
  const int l = ms->l;
  if (mh.lnl_trace) mh.lnl_trace[l] = lnl_new;
  const double diff = lnl_new - ms->lnl_old;
  const double ratio = exp(diff);  // :310, Q4
  const bool accept = (mh.ml_mode == 0) ? (ratio > 1.0) : (ratio > mh_draw_u(mh, l));
  if (accept) {
    ms->sample[mh.nind] = ms->theta[mh.nind];
    ms->lnl_old = lnl_new;
    ms->accept += 1.0;
  }
  if (mh.decisions) mh.decisions[l] = accept ? 1 : 0;
  ms->l = l + 1;
}

// streaming mode: consume the sums just reduced (gathered over ranks), decide, propose next.
// Row layout: [0] chi-square lnL, [1] Jeffreys sum, [2 + 4*j + 2*s + {0,1}] = TNd, TNT (marginal).
static __global__ void mh_fullsky_step_kernel(const ModelView mv, const MhView mh, MhScalars *ms,
                                       const double *gathered, int nranks, int cnt) {
  if (ms->skip) return;
  double lnl = 0.0;
  if (mh.lnl_type == 0) {
    for (int g = 0; g < nranks; g++) lnl += gathered[g * cnt];
  } else if (mh.lnl_type == 1) {  // evaluate_marginal_lnL, src/dang_lnl_mod.f90:113-122
    for (int j = 0; j < mv.nbands; j++)
      for (int s = 0; s < mh.S; s++) {
        double TNd = 0.0, TNT = 0.0;
        for (int g = 0; g < nranks; g++) {
          TNd += gathered[g * cnt + 2 + 4 * j + 2 * s];
          TNT += gathered[g * cnt + 2 + 4 * j + 2 * s + 1];
        }
        const double invTNT = 1.0 / TNT;
        lnl = lnl - 0.5 * TNd * invTNT * TNd;
      }
  }
  const double val = ms->phase == 0 ? ms->sample[mh.nind] : ms->theta[mh.nind];
  double prior = 0.0;
  if (mh.prior_type == 1) {
    prior = log_normal_prior(val, mh.gauss[0], mh.gauss[1]);
  } else if (mh.prior_type == 2) {  // eval_jeffreys_prior, src/dang_lnl_mod.f90:242-304 (:263, :302)
    double sum = 0.0;
    for (int g = 0; g < nranks; g++) sum += gathered[g * cnt + 1];
    prior = log(sqrt(sum));
  }
  if (ms->phase == 0) {  // lnL of the starting point (:250, :261, :268)
    ms->lnl_old = lnl + prior;
    ms->phase = 1;
  } else {
    mh_accept_step(mh, ms, lnl, prior);
  }
  mh_next_proposal(mv, mh, ms);
}

// step-size tuner, streaming form: the chain start given explicitly (the per-pixel call site starts the tuner at
// the map's mean, src/dang_sample_mod.f90:341-347) ...
static __global__ void mh_fullsky_override_kernel(const ModelView mv, const MhView mh, MhScalars *ms, double t0, double t1) {
  ms->sample[0] = ms->theta[0] = t0;
  ms->sample[1] = ms->theta[1] = t1;
  for (int j = 0; j < mv.nbands; j++) {
    const double s = sed_theta(mv, mh.ic, j, t0, t1, mh.plane[0]);
    ms->sed[j] = s;
    ms->s0[j] = s;
  }
}
// ... and the start of the next block of nsample proposals: the chain goes on (sample, lnl_old stay), the
// acceptance count restarts (:664-665)
static __global__ void mh_tune_block_start_kernel(const ModelView mv, const MhView mh, MhScalars *ms) {
  ms->accept = 0.0;
  ms->l = 0;
  ms->skip = 0;
  mh_next_proposal(mv, mh, ms);
}

// lnl_type 'prior' (:255-257): no chain, the index is drawn from its Gaussian prior
static __global__ void mh_fullsky_prior_draw_kernel(const MhView mh, MhScalars *ms) {
  ms->sample[mh.nind] = mh.gauss[0] + mh.gauss[1] * mh_draw_z(mh, 0);
  ms->skip = 1;
}

// D[j][s][Ppad] = data_raw for the sampled planes (streaming mode only)
static __global__ void __launch_bounds__(DG_THREADS)
mh_data_kernel(const ModelView mv, const MhView mh, double *D) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride)
    for (int j = 0; j < mv.nbands; j++)
      for (int s = 0; s < mh.S; s++)
        D[((size_t)j * mh.S + s) * mv.Ppad + p] = mh_data_value(mv, mh.ic, j, mh.plane[s], p);
}

// evaluate_lnL full sky, src/dang_lnl_mod.f90:126-182, for model = amplitude * ms->sed[j]
static __global__ void __launch_bounds__(DG_THREADS)
mh_fullsky_lnl_kernel(const ModelView mv, const MhView mh, const MhScalars *ms, const double *D,
                      double *partials, unsigned int *ticket, double *out) {
  if (ms->skip) return;
  __shared__ double smem[2 * 32];
  __shared__ double ssed[DG_MAX_BANDS];
  if (threadIdx.x < mv.nbands) ssed[threadIdx.x] = ms->sed[threadIdx.x];
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  const CompView &cv = mv.comp[mh.ic];
  const bool jeff = mh.prior_type == 2 && mh.is_synch;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride) {
    if (!mv.mask[p]) continue;
    double lnl = 0.0, js = 0.0;
    for (int s = 0; s < mh.S; s++) {
      const double a = cv.amp[(size_t)mh.plane[s] * mv.Ppad + p];
      for (int j = 0; j < mv.nbands; j++) {
        const double rms = ldg_stream(mv.rms + plane_off(mv, j, mh.plane[s]) + p);
        if (mh.lnl_type == 0) {
          const double d = ldg_stream(D + ((size_t)j * mh.S + s) * mv.Ppad + p);
          const double t = (d - a * ssed[j]) / rms;
          lnl = lnl - 0.5 * (t * t);
        }
        if (jeff) {  // (((1/rms)**2) * (ss/amplitude) * log(nu_c/nu_ref))**2, :296-297
          const double ir = 1.0 / rms;
          const double t = (ir * ir) * ((a * ssed[j]) / a) * log(mv.band[j].nu_c / cv.nu_ref);
          js = js + t * t;
        }
      }
    }
    acc[0] += lnl;
    acc[1] += js;
  }
  grid_reduce<2>(acc, smem, partials, ticket, out);
}

// evaluate_marginal_lnL full sky (:113-122): per (band, Stokes) the sums over ALL pixels (the
// source ignores the mask) of TN*data and TN*model, TN = model / rms^2, model = amplitude * sed.
// out[2 + 4*j + 2*s + {0,1}]; bands in chunks of DG_SUFF_CHUNK.
static __global__ void __launch_bounds__(DG_THREADS)
mh_fullsky_marginal_kernel(const ModelView mv, const MhView mh, const MhScalars *ms, const double *D,
                           double *partials, unsigned int *tickets, double *out) {
  if (ms->skip) return;
  constexpr int NV = 4 * DG_SUFF_CHUNK;  // (TNd, TNT) x 2 planes per band
  __shared__ double smem[NV * 32];
  __shared__ double ssed[DG_MAX_BANDS];
  if (threadIdx.x < mv.nbands) ssed[threadIdx.x] = ms->sed[threadIdx.x];
  __syncthreads();
  const CompView &cv = mv.comp[mh.ic];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int nchunk = (mv.nbands + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
  for (int ch = 0; ch < nchunk; ch++) {
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride) {
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (s >= mh.S) continue;
        const double a = cv.amp[(size_t)mh.plane[s] * mv.Ppad + p];
#pragma unroll
        for (int jj = 0; jj < DG_SUFF_CHUNK; jj++) {
          const int j = ch * DG_SUFF_CHUNK + jj;
          if (j < mv.nbands) {
            const double d = ldg_stream(D + ((size_t)j * mh.S + s) * mv.Ppad + p);
            const double rms = ldg_stream(mv.rms + plane_off(mv, j, mh.plane[s]) + p);
            const double model = a * ssed[j];
            const double TN = model / (rms * rms);
            acc[4 * jj + 2 * s + 0] += TN * d;
            acc[4 * jj + 2 * s + 1] += TN * model;
          }
        }
      }
    }
    // chunk ch holds bands ch*CHUNK .. : slot 2 + 4*j + 2*s + {0,1}
    grid_reduce<NV>(acc, smem, partials + (size_t)ch * NV * gridDim.x, tickets + ch, out + 2 + ch * NV);
    __syncthreads();
  }
}

// Sufficient statistics of the full-sky chi-square about the chain's starting SED s0:
//   R = data - a s0,  t = R / sigma,  u = a / sigma
//   X_j = sum t^2,  Y_j = sum t u,  Z_j = sum u^2      (unmasked pixels, sampled planes)
// so that for a proposal with SED s0 + delta:  sum ((data - a s)/sigma)^2 = X - 2 delta Y + delta^2 Z.
// Expanding about s0 (not about 0) keeps X at the chi-square itself: no cancellation.
// out[(s*nchunk + chunk)*12 + 3*jj + {0,1,2}]: per plane s, bands in chunks of DG_SUFF_CHUNK
static __global__ void __launch_bounds__(DG_THREADS)
mh_suffstat_kernel(const ModelView mv, const MhView mh, const MhScalars *ms, double *partials,
                   unsigned int *tickets, double *out) {
  constexpr int NV = 3 * DG_SUFF_CHUNK;
  __shared__ double smem[NV * 32];
  const CompView &cv = mv.comp[mh.ic];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int nchunk = (mv.nbands + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
  for (int s = 0; s < mh.S; s++)
  for (int ch = 0; ch < nchunk; ch++) {
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.0;
    double s0[DG_SUFF_CHUNK];
#pragma unroll
    for (int jj = 0; jj < DG_SUFF_CHUNK; jj++) {
      const int j = ch * DG_SUFF_CHUNK + jj;
      s0[jj] = j < mv.nbands ? ms->s0[j] : 0.0;
    }
    const int k = mh.plane[s];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride) {
      if (!mv.mask[p]) continue;
      const double a = cv.amp[(size_t)k * mv.Ppad + p];
#pragma unroll
      for (int jj = 0; jj < DG_SUFF_CHUNK; jj++) {
        const int j = ch * DG_SUFF_CHUNK + jj;
        if (j < mv.nbands) {
          const double d = mh_data_value(mv, mh.ic, j, k, p);
          const double rms = ldg_stream(mv.rms + plane_off(mv, j, k) + p);
          const double t = (d - a * s0[jj]) / rms;
          const double u = a / rms;
          acc[3 * jj + 0] += t * t;
          acc[3 * jj + 1] += t * u;
          acc[3 * jj + 2] += u * u;
        }
      }
    }
    const int sc = s * nchunk + ch;
    grid_reduce<NV>(acc, smem, partials + (size_t)sc * NV * gridDim.x, tickets + sc, out + sc * NV);
    __syncthreads();
  }
}

// The whole full-sky chain from the statistics (:282-324), no map traffic.  One warp: lane j owns
// band j (its statistics, s0 and the proposal's SED), lnL is a fixed-order butterfly over lanes;
// every lane carries the same chain state, lane 0 publishes it.
// tab_out != nullptr: the component's planes were tabulated already (constant index maps) and stay so, and the
// next full-sky draw of this index starts where this one ends: the kernel leaves the SED of the final sample in the
// table (the very value sed_table_kernel would compute) and in ms->s0, so that neither the table kernel nor the
// chain-start kernels (first pixel, gather, init) have to run before the next statistics pass.
// Injected deviates of a short chain travel as a kernel ARGUMENT (FsDeviates), not through a host-to-device copy:
// a cudaMemcpyAsync of 160 bytes on the compute stream queues on a copy engine behind whatever bulk download is
// running and stalls the whole spectral-parameter block for hundreds of microseconds (profiles/r02_timeline_*).
#define DG_FS_DEV_MAX 128
struct FsDeviates {
  int n;  // 0: not used (device RNG, or the deviates sit in device buffers mh.z / mh.u)
  double z[DG_FS_DEV_MAX], u[DG_FS_DEV_MAX];
};
static __global__ void __launch_bounds__(32)
mh_suff_chain_kernel(const ModelView mv, const MhView mh_in, MhScalars *ms, const double *gathered,
                     int nranks, int cnt, SedTable *tab_out, const FsDeviates dev) {
  __shared__ double sz[DG_FS_DEV_MAX], su[DG_FS_DEV_MAX];
  MhView mh = mh_in;
  if (dev.n > 0) {
    for (int i = threadIdx.x; i < dev.n; i += 32) {
      sz[i] = dev.z[i];
      su[i] = dev.u[i];
    }
    __syncwarp();
    mh.z = sz;
    mh.u = su;
  }
  const int B = mv.nbands, j = threadIdx.x;
  double X = 0.0, Y = 0.0, Z = 0.0, s0 = 0.0;
  double Xk[2] = {0.0, 0.0}, Yk[2] = {0.0, 0.0}, Zk[2] = {0.0, 0.0};  // per plane, summed over ranks in rank order
  if (j < B) {
    const int nchunk = (B + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
    const int o = (j / DG_SUFF_CHUNK) * 3 * DG_SUFF_CHUNK + 3 * (j % DG_SUFF_CHUNK);
    for (int s = 0; s < mh.S; s++) {
      for (int g = 0; g < nranks; g++) {
        const double *row = gathered + (size_t)g * cnt + (size_t)s * nchunk * 3 * DG_SUFF_CHUNK + o;
        Xk[s] += row[0];
        Yk[s] += row[1];
        Zk[s] += row[2];
      }
      X += Xk[s];
      Y += Yk[s];
      Z += Zk[s];
    }
    s0 = ms->s0[j];
  }
  double sample[DG_MAXIND] = {ms->sample[0], ms->sample[1]};
  // mbb beta chain on a delta band: the Planck factor eref / (exp(z nu_c) - 1) does not move (T is fixed),
  // so it is formed once; sed_mbb multiplies exactly this quotient by the power law: same bits, 1 exp per
  // proposal on the chain's serial path instead of 3
  const bool planck_fixed = j < B && mv.comp[mh.ic].type == 2 && mh.nind == 0 && mv.band[j].n == 0;
  double Fj = 0.0, lh = 0.0, ll = 0.0;
  if (planck_fixed) {
    const double z = DG_H / (DG_KB * sample[1]);
    Fj = (exp(z * mv.comp[mh.ic].nu_ref) - 1.0) / (exp(z * mv.band[j].nu_c) - 1.0);
    lh = mv.tab->lnr_hi[mh.ic][j];
    ll = mv.tab->lnr_lo[mh.ic][j];
  }
  auto sed_of = [&](double t0, double t1) -> double {
    if (j >= B) return 0.0;
    return planck_fixed ? Fj * exp_scaled(t0 + 1.0, lh, ll) : sed_theta(mv, mh.ic, j, t0, t1, mh.plane[0]);
  };
  // lnL(theta) = -1/2 sum_j (X_j - 2 delta_j Y_j + delta_j^2 Z_j), delta_j = sed_j(theta) - s0_j
  auto lnl_of = [&](double sed) -> double {
    const double dl = sed - s0;
    double v = (j < B) ? -0.5 * (X - 2.0 * dl * Y + dl * dl * Z) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  auto prior_of = [&](double val) -> double {
    return mh.prior_type == 1 ? log_normal_prior(val, mh.gauss[0], mh.gauss[1]) : 0.0;
  };
  double lnl_old = lnl_of(s0) + prior_of(sample[mh.nind]);  // :250, :261, :268
  double accept = 0.0;
  // ---- the chain, speculatively parallel over proposals.  Until a proposal is accepted the chain's state does not
  // move, so lane i evaluates proposal l0 + i against the CURRENT state -- all bands itself, the band sum in the
  // very order of the lane butterfly above (same bits as the one-proposal-at-a-time loop this replaces) -- and the
  // first accepted lane ends the round: proposals before it were rejected against the right state, it becomes the
  // state, everything after it is re-evaluated.  Rounds = accepted moves + 1 instead of nsample serial steps.
  __shared__ double cX[32], cY[32], cZ[32], cS0[32], cF[32], cLh[32], cLl[32];
  __shared__ double gz[DG_FS_DEV_MAX], gu[DG_FS_DEV_MAX];
  __shared__ double sv[32][33];
  cX[j] = X;
  cY[j] = Y;
  cZ[j] = Z;
  cS0[j] = s0;
  cF[j] = Fj;
  cLh[j] = lh;
  cLl[j] = ll;
  const bool pf_all = __all_sync(0xffffffffu, planck_fixed || j >= B);  // (the same SED form in every band)
  const bool pregen = mh.nsample <= DG_FS_DEV_MAX;
  if (pregen)
    for (int i = j; i < mh.nsample; i += 32) {  // deviates of the whole chain, off the serial path
      gz[i] = mh_draw_z(mh, i);
      gu[i] = mh.ml_mode == 0 ? 1.0 : mh_draw_u(mh, i);
    }
  __syncwarp();
  for (int l0 = 0; l0 < mh.nsample;) {
    const int l = l0 + j;
    const bool live = l < mh.nsample;
    double th[DG_MAXIND] = {sample[0], sample[1]};
    double lnl_new = 0.0, u_l = 1.0;
    bool oob = false, acc = false;
    if (live) {
      const double z_l = pregen ? gz[l] : mh_draw_z(mh, l);
      th[mh.nind] = sample[mh.nind] + (0.0 + mh.step * z_l);  // :286
      oob = th[mh.nind] < mh.uni[0] || th[mh.nind] > mh.uni[1];  // :287, Q5
      if (!oob) {
        u_l = pregen ? gu[l] : (mh.ml_mode == 0 ? 1.0 : mh_draw_u(mh, l));
        if (pf_all) {  // independent exps: unrolled so that they overlap
#pragma unroll 4
          for (int jb = 0; jb < 32; jb++) {
            double v = 0.0;
            if (jb < B) {
              const double sed = cF[jb] * exp_scaled(th[0] + 1.0, cLh[jb], cLl[jb]);
              const double dl = sed - cS0[jb];
              v = -0.5 * (cX[jb] - 2.0 * dl * cY[jb] + dl * dl * cZ[jb]);
            }
            sv[j][jb] = v;
          }
        } else {
#pragma unroll 1
          for (int jb = 0; jb < 32; jb++) {
            double v = 0.0;
            if (jb < B) {
              const double sed = sed_theta(mv, mh.ic, jb, th[0], th[1], mh.plane[0]);
              const double dl = sed - cS0[jb];
              v = -0.5 * (cX[jb] - 2.0 * dl * cY[jb] + dl * dl * cZ[jb]);
            }
            sv[j][jb] = v;
          }
        }
        for (int o = 16; o > 0; o >>= 1)  // the butterfly's pairing: (i, i ^ o) level by level
          for (int i = 0; i < o; i++) sv[j][i] += sv[j][i + o];
        lnl_new = sv[j][0] + prior_of(th[mh.nind]);  // :306
        const double diff = lnl_new - lnl_old;
        const double ratio = exp(diff);  // :310, Q4
        acc = (mh.ml_mode == 0) ? (ratio > 1.0) : (ratio > u_l);
      }
    }
    const unsigned hit = __ballot_sync(0xffffffffu, live && !oob && acc);
    const int first = hit ? __ffs(hit) - 1 : 32;  // lane of the first accepted proposal of this round
    if (live && j <= first) {                     // these proposals are final
      if (mh.decisions) mh.decisions[l] = oob ? 2 : (acc ? 1 : 0);
      if (mh.lnl_trace && !oob) mh.lnl_trace[l] = lnl_new;
    }
    if (hit) {
      sample[mh.nind] = __shfl_sync(0xffffffffu, th[mh.nind], first);
      lnl_old = __shfl_sync(0xffffffffu, lnl_new, first);
      accept += 1.0;
      l0 += first + 1;
    } else {
      l0 += 32;
    }
  }
  // chi-square of the final state per plane, from the same statistics (what compute_chisq would
  // sum after the draw: sum_j sum_pix ((d - sky)/sigma)^2 / nbands, src/dang_data_mod.f90:514-523)
  double chi[2] = {0.0, 0.0};
  {
    const double sed = sed_of(sample[0], sample[1]);
    if (tab_out && j < B) {
      for (int s = 0; s < mh.S; s++) tab_out->sed[mh.ic * 3 + mh.plane[s]][j] = sed;
      ms->s0[j] = sed;
      ms->sed[j] = sed;
    }
    const double dl = sed - s0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      double v = (j < B && s < mh.S) ? (Xk[s] - 2.0 * dl * Yk[s] + dl * dl * Zk[s]) / (double)B : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      chi[s] = v;
    }
  }
  if (j == 0) {
    ms->sample[0] = sample[0];
    ms->sample[1] = sample[1];
    ms->theta[0] = sample[0];
    ms->theta[1] = sample[1];
    ms->lnl_old = lnl_old;
    ms->accept = accept;
    ms->l = mh.nsample;
    ms->phase = 1;
    ms->skip = 1;
    ms->chisq[0] = chi[0];
    ms->chisq[1] = chi[1];
  }
}

// tune_spectral_parameter_length, src/dang_sample_mod.f90:623-717, on the same statistics:
// blocks of nsample proposals; halve the step if accept/(nsample+1) < 0.4, x1.5 if > 0.6
// (single-precision literals as in the source), stop when it lands in between.
// out[0] = tuned step, out[1] = blocks run, out[2] = tuned flag.  Deviate slots: blk*nsample + l.
static __global__ void __launch_bounds__(32)
mh_suff_tune_kernel(const ModelView mv, const MhView mh, const MhScalars *ms, const double *gathered,
                    int nranks, int cnt, int max_blocks, double *out) {
  const int B = mv.nbands, j = threadIdx.x;
  double X = 0.0, Y = 0.0, Z = 0.0, s0 = 0.0;
  double Xk[2] = {0.0, 0.0}, Yk[2] = {0.0, 0.0}, Zk[2] = {0.0, 0.0};  // per plane, summed over ranks in rank order
  if (j < B) {
    const int nchunk = (B + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
    const int o = (j / DG_SUFF_CHUNK) * 3 * DG_SUFF_CHUNK + 3 * (j % DG_SUFF_CHUNK);
    for (int s = 0; s < mh.S; s++) {
      for (int g = 0; g < nranks; g++) {
        const double *row = gathered + (size_t)g * cnt + (size_t)s * nchunk * 3 * DG_SUFF_CHUNK + o;
        Xk[s] += row[0];
        Yk[s] += row[1];
        Zk[s] += row[2];
      }
      X += Xk[s];
      Y += Yk[s];
      Z += Zk[s];
    }
    s0 = ms->s0[j];
  }
  double sample[DG_MAXIND] = {ms->sample[0], ms->sample[1]};
  double theta[DG_MAXIND] = {sample[0], sample[1]};
  auto lnl_of = [&](double sed) -> double {
    const double dl = sed - s0;
    double v = (j < B) ? -0.5 * (X - 2.0 * dl * Y + dl * dl * Z) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  double lnl_old = 0.0, lnl_new = 0.0;
  {
    const double lnl = lnl_of(s0);
    if (mh.prior_type == 1) lnl_old = lnl + log_normal_prior(sample[mh.nind], mh.gauss[0], mh.gauss[1]);  // :658-662
    else if (mh.prior_type == 0) lnl_old = lnl;
  }
  double step = mh.step;
  int blk = 0, tuned = 0;
  while (!tuned && blk < max_blocks) {
    double accept = 0.0;
    for (int l = 0; l < mh.nsample; l++) {
      const int slot = blk * mh.nsample + l;
      const double zz = mh.z ? mh.z[slot] : philox_normal(mh.seed, DG_STREAM_TUNE_Z, (uint64_t)slot);
      theta[mh.nind] = sample[mh.nind] + (0.0 + step * zz);  // :668
      if (theta[mh.nind] < mh.uni[0] || theta[mh.nind] > mh.uni[1]) continue;
      const double sed = (j < B) ? sed_theta(mv, mh.ic, j, theta[0], theta[1], mh.plane[0]) : 0.0;
      const double lnl = lnl_of(sed);
      if (mh.prior_type == 1) lnl_new = lnl + log_normal_prior(theta[mh.nind], mh.gauss[0], mh.gauss[1]);
      else if (mh.prior_type == 0) lnl_new = lnl;
      const double ratio = exp(lnl_new - lnl_old);
      bool acc;
      if (mh.ml_mode == 0) {
        acc = ratio > 1.0;
      } else {
        double uu, u2;
        if (mh.u) uu = mh.u[slot];
        else philox_uniform2(mh.seed, DG_STREAM_TUNE_U, (uint64_t)slot, uu, u2);
        acc = ratio > uu;
      }
      if (acc) {
        sample[mh.nind] = theta[mh.nind];
        lnl_old = lnl_new;
        accept = accept + 1;
      }
    }
    const double rate = accept / (double)(mh.nsample + 1);  // Fortran loop counter after the loop (:707)
    if (rate < (double)0.4f) step = step - (double)0.5f * step;
    else if (rate > (double)0.6f) step = step + (double)0.5f * step;
    else tuned = 1;
    blk++;
  }
  if (j == 0) {
    out[0] = step;
    out[1] = (double)blk;
    out[2] = (double)tuned;
  }
}

// index_full_res(:, map_inds) = sample(nind) -> c%indices (:329, :483)
static __global__ void mh_fullsky_store_kernel(const ModelView mv, const MhView mh, const MhScalars *ms) {
  const CompView &cv = mv.comp[mh.ic];
  const double v = ms->sample[mh.nind];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < mv.P; p += stride)
    for (int s = 0; s < mh.S; s++) cv.idx[mh.nind][(size_t)mh.plane[s] * mv.Ppad + p] = v;
}
