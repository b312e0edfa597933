// kernels_mh_pix.cuh -- K5, screened form with ONE THREAD PER PIXEL and a two-number band state.
//
// Same contract as kernels_mh_fast.cuh (per-pixel branch of sample_index_mh, src/dang_sample_mod.f90:332-481, chisq
// likelihood, uniform / Gaussian prior, delta bands, power-law beta / mbb beta / mbb T): every proposal is decided
// from a single-precision evaluation when its error bound allows, from the fp64 arithmetic of mh_perpixel_kernel
// otherwise, so the decisions -- and the stored index maps -- are those of the fp64 kernel.
//
// What changed against the lane-cooperative form (4 lanes per pixel, t and g per band and Stokes in registers):
//   * the index is common to the sampled planes, so rho_j (the relative SED change of band j) is too, and
//         lnL(theta') - lnL(cur) = 1/2 sum_j rho_j (2 P_j - rho_j Q_j),
//         P_j = sum_s g_js t_js,   Q_j = sum_s g_js^2        (t = (d - a s_cur)/sigma, g = a s_cur/sigma)
//     needs TWO numbers per band instead of four, and two FMAs per band instead of six.  An accepted move maps the
//     state onto itself:  P' = (1 + rho)(P - rho Q),  Q' = (1 + rho)^2 Q.
//   * one thread owns one pixel: no shuffles, and the scalar part of a proposal (the fp64 step, bounds check, prior,
//     ln u, the error budget) is done once per pixel instead of once per lane -- it was ~40 % of the executed
//     instructions.  The band state {2P, Q, e_P} lives in shared memory as one float4 per (band, thread)
//     (+ K_j in T mode); band constants come from the constant bank (kernel parameter).
//   * the fp64 fallback re-reads the maps (5e-5 of the proposals), so nothing fp64 is parked on chip.
//   * exp(x) - 1 of the hot loop is ONE polynomial (degree 9, |x| < 0.7, ~3 ulp): with the step sizes the tuner
//     leaves, |x| = |step z ln(nu/nu_ref)| stays far below that.  A proposal that leaves the range in some band
//     takes the general two-branch k5_em1f in a rolled loop instead.
//   * rho_j of the proposal is parked in the spare lane of the band's float4, so an accepted move does not
//     evaluate it again; the state-error growth uses b_max = max_j b_j for every band.
//
// Error budget of the screened difference lam = sum_j rho_j (2P_j - rho_j Q_j), all in units of lnL:
//   |d rho_j| <= kappa b_j  (b_j as in kernels_mh_fast.cuh)    ->  2 kappa sum_j b_j m_j,   m_j = |2P_j| + |rho_j| Q_j
//   fp32 rounding of the NB-term sum                            ->  (NB + 4) 2^-24 sum_j |rho_j| m_j =: c_r W
//   state: |d(2P_j)| <= e_P,j (carried per band), Q relative eQ ->  sum_j |rho_j| e_P,j + eQ W
// diff~ = lam / 2 + (prior' - prior_cur) is certain when |diff~ - ln u| > eps = 0.55 (...) + eps_u.
// e_P and eQ start at the fp64 -> fp32 rounding of the state and grow with every accepted move by the first-order
// propagation of the same three error sources through the update (see `accept` below); a move with 1 + rho < 1/2
// in some band (never for tuned step sizes) voids the budget and sends the rest of that chain to the fp64 path.
// In record mode every proposal is also evaluated in fp64 and out[2] counts broken bounds / wrong certain decisions.
#pragma once
#include "kernels_mh_fast.cuh"

#define DG_K5P_MAXB 32
#define DG_K5P_XMAX 0.7f
struct K5Bands {
  float cf[DG_K5P_MAXB];  // ln(nu_j / nu_ref) (beta modes) or h nu_j / k (T mode), 0 beyond nbands
  float cref;             // h nu_ref / k
  float cfmax;            // max |cf_j| (T mode: and cref)
};

// exp(x) - 1 for |x| < DG_K5P_XMAX: Taylor to x^9 in Horner form (truncation < 1.2e-8 relative, rounding ~3 ulp)
__device__ __forceinline__ float k5p_rcp(float x) {  // MUFU.RCP: 1 ulp
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float k5p_em1(float x) {
  float p = fmaf(x, 1.0f / 362880.0f, 1.0f / 40320.0f);
  p = fmaf(p, x, 1.0f / 5040.0f);
  p = fmaf(p, x, 1.0f / 720.0f);
  p = fmaf(p, x, 1.0f / 120.0f);
  p = fmaf(p, x, 1.0f / 24.0f);
  p = fmaf(p, x, 1.0f / 6.0f);
  p = fmaf(p, x, 0.5f);
  p = fmaf(p, x, 1.0f);
  return p * x;
}

// lnL(xe) in fp64 straight from the maps, in the arithmetic AND summation order of mh_perpixel_kernel: four partial
// sums over the bands j = r, r + 4, ... (its four lanes), combined as (p0 + p1) + (p2 + p3)
template <int MODE>
__device__ __noinline__ double k5p_exact(const ModelView &mv, const MhView &mh, int64_t pp, double xe, double idx0,
                                         double idx1, double amp0, double amp1) {
  const int B = mv.nbands, S = mh.S;
  const CompView &cv = mv.comp[mh.ic];
  const SedTable &tab = *mv.tab;
  const double nu_ref = cv.nu_ref;
  double zT = 0.0, eref = 0.0, zF = 0.0, erefF = 0.0;
  if (MODE == MH_SED_MBB_T) {
    zT = DG_H / (DG_KB * xe);
    eref = exp(zT * nu_ref) - 1.0;
  }
  if (MODE == MH_SED_MBB_BETA) {
    zF = DG_H / (DG_KB * idx1);
    erefF = exp(zF * nu_ref) - 1.0;
  }
  double part[DG_MH_LANES];
#pragma unroll
  for (int r = 0; r < DG_MH_LANES; r++) {
    double acc = 0.0;
#pragma unroll 1
    for (int j = r; j < B; j += DG_MH_LANES) {
      const double Lh = tab.lnr_hi[mh.ic][j], Ll = tab.lnr_lo[mh.ic][j], nuc = mv.band[j].nu_c;
      double D0, s0;
      k5_fetch(mv, mh.ic, j, mh.plane[0], pp, D0, s0);
      const double W0 = 1.0 / s0;
      double sed;
      if (MODE == MH_SED_POWERLAW) {
        sed = exp_scaled(xe, Lh, Ll);
      } else if (MODE == MH_SED_MBB_BETA) {
        const double F = erefF / (exp(zF * nuc) - 1.0);
        sed = F * exp_scaled(xe + 1.0, Lh, Ll);
      } else {
        const double F = exp_scaled(idx0 + 1.0, Lh, Ll);
        sed = eref * mh_fast_rcp(exp(zT * nuc) - 1.0) * F;
      }
      const double t0 = (D0 - amp0 * sed) * W0;
      acc = acc - 0.5 * (t0 * t0);
      if (S > 1) {
        double D1, s1;
        k5_fetch(mv, mh.ic, j, mh.plane[1], pp, D1, s1);
        const double W1 = 1.0 / s1;
        const double t1 = (D1 - amp1 * sed) * W1;
        acc = acc - 0.5 * (t1 * t1);
      }
    }
    part[r] = acc;
  }
  static_assert(DG_MH_LANES == 4, "summation order of the four-lane kernel");
  return (part[0] + part[1]) + (part[2] + part[3]);
}

// sum over the OTHER components of amplitude x SED at (band j, plane k, pixel p): what mh_data_value subtracts from
// the map (src/dang_sample_mod.f90:173-196).  Out of line: the SED dispatch exists once.
static __device__ __noinline__ double k5p_others(const ModelView &mv, int ic, int j, int k, int64_t p) {
  const size_t kp = (size_t)k * mv.Ppad + p;
  double v = 0.0;
  for (int c2 = 0; c2 < mv.ncomp; c2++) {
    if (c2 == ic) continue;
    const CompView &cc = mv.comp[c2];
    const double t0 = cc.nind > 0 ? cc.idx[0][kp] : 0.0;
    const double t1 = cc.nind > 1 ? cc.idx[1][kp] : 0.0;
    v = v + cc.amp[kp] * sed_eval(mv, c2, k, j, t0, t1);
  }
  return v;
}

// Three blocks per SM (168 registers: at 128 the loop spills around every proposal) and the band loops unrolled by
// four only: measured on c4 / nside 512, B200 -- (4 blocks, full unroll) 7.6 ms per launch, (3, full) 7.5, (3, 10) 7.1,
// (3, 5) 6.5, (3, 4) 6.3, (3, 2) 6.3, (3, 1) 6.8, (2, 5) 7.9; the lane-cooperative form 7.9 (profiles/r02_k5.md).
template <int NB, int MODE, int MINB = 3, int UNR = 4>
__global__ void __launch_bounds__(DG_MH_THREADS, MINB)
mh_perpixel_pix_kernel(const __grid_constant__ ModelView mv, const __grid_constant__ MhView mh,
                       const __grid_constant__ K5Bands kb, double *partials, unsigned int *ticket, double *out) {
  extern __shared__ __align__(16) unsigned char k5p_raw[];
  float4 *st = reinterpret_cast<float4 *>(k5p_raw) + threadIdx.x;                                   // [NB][threads]: {2P, Q, e_P, -}
  float *kjs = reinterpret_cast<float *>(k5p_raw + (size_t)NB * DG_MH_THREADS * sizeof(float4)) + threadIdx.x;  // T mode: K_j
  __shared__ double smem[4 * 32];
  constexpr int TH = DG_MH_THREADS;
  const int B = mv.nbands, S = mh.S, nsample = mh.nsample;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const CompView &cv = mv.comp[mh.ic];
  const SedTable &tab = *mv.tab;
  const double nu_ref = cv.nu_ref;
  const double ln_denom = log(mh.gauss[1] * sqrt(2.0 * DG_PI));
  const double inv2var = 1.0 / (2 * (mh.gauss[1] * mh.gauss[1]));
  const bool record = mh.decisions != nullptr;
  constexpr float kappa = MODE == MH_SED_MBB_T ? DG_K5_KAPPA_T : DG_K5_KAPPA_BETA;
  constexpr float c_round = (NB + 4) * 5.97e-8f;

  const int64_t stride = (int64_t)gridDim.x * TH;
  for (int64_t p = (int64_t)blockIdx.x * TH + threadIdx.x; p < mv.P; p += stride) {
    if (mv.mask[p] == 0) {  // :362; index_map stays 0 for masked pixels (:223, :465, :483)
      for (int s = 0; s < S; s++) cv.idx[mh.nind][(size_t)mh.plane[s] * mv.Ppad + p] = 0.0;
      if (mh.decisions)
        for (int l = 0; l < nsample; l++) mh.decisions[(size_t)l * mv.P + p] = 3;
      continue;
    }
    const size_t kp0 = (size_t)mh.plane[0] * mv.Ppad + p;
    const double idx0 = cv.nind > 0 ? cv.idx[0][kp0] : 0.0;  // :372-374
    const double idx1 = cv.nind > 1 ? cv.idx[1][kp0] : 0.0;
    double cur = mh.nind == 0 ? idx0 : idx1;
    const double amp0 = cv.amp[(size_t)mh.plane[0] * mv.Ppad + p];
    const double amp1 = S > 1 ? cv.amp[(size_t)mh.plane[1] * mv.Ppad + p] : 0.0;

    // ---- state at the chain's first point: fp64 evaluation, fp32 copy
    // T mode keeps K - 1 = 1 / em1(h nu / k T_cur) for the reference frequency (kref1) and per band (kjs): an accepted
    // move rescales them, K' - 1 = (K - 1) / (1 + n), with the factors the proposal already formed
    float kref1 = 0.0f;
    {
      double zF = 0.0, erefF = 0.0;
      if (MODE != MH_SED_POWERLAW) {
        zF = DG_H / (DG_KB * idx1);  // T of the first point (idx1 == cur in T mode)
        erefF = exp(zF * nu_ref) - 1.0;
        kref1 = (float)(1.0 / erefF);
      }
      // the maps of four bands at a time: 16 independent loads in flight per thread before anything waits on them
      static_assert(NB % 4 == 0, "band batches of four");
#pragma unroll 1
      for (int j0 = 0; j0 < NB; j0 += 4) {
        double sg[4][2], rm[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int j = j0 + i;
          sg[i][0] = sg[i][1] = 0.0;
          rm[i][0] = rm[i][1] = 1.0;
          if (j < B) {
            sg[i][0] = ldg_stream(mv.sig + plane_off(mv, j, mh.plane[0]) + p);
            rm[i][0] = ldg_stream(mv.rms + plane_off(mv, j, mh.plane[0]) + p);
            if (S > 1) {
              sg[i][1] = ldg_stream(mv.sig + plane_off(mv, j, mh.plane[1]) + p);
              rm[i][1] = ldg_stream(mv.rms + plane_off(mv, j, mh.plane[1]) + p);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int j = j0 + i;
          float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          float kj = kref1;  // (padding bands carry the reference frequency in T mode: rho = 0)
          if (j < B) {
            const double Lh = tab.lnr_hi[mh.ic][j], Ll = tab.lnr_lo[mh.ic][j];
            double sed;
            if (MODE == MH_SED_POWERLAW) {
              sed = exp_scaled(idx0, Lh, Ll);
            } else {
              const double em1 = exp(zF * mv.band[j].nu_c) - 1.0;
              const double iem1 = mh_fast_rcp(em1);
              sed = erefF * iem1 * exp_scaled(idx0 + 1.0, Lh, Ll);
              kj = (float)iem1;
            }
            double P2 = 0.0, Q = 0.0, Pabs = 0.0;
#pragma unroll
            for (int sI = 0; sI < 2; sI++) {
              if (sI >= S) break;
              const int k = mh.plane[sI];
              double D = sg[i][sI];
              if (k == 0) D = (D - mv.offset[j]) / mv.gain[j];
              D -= k5p_others(mv, mh.ic, j, k, p);
              const double Wn = mh_fast_rcp(rm[i][sI]), m = (sI == 0 ? amp0 : amp1) * sed;
              const double t = (D - m) * Wn, g = m * Wn;
              P2 += 2.0 * (g * t);
              Q += g * g;
              Pabs += 2.0 * fabs(g * t);
            }
            o.x = (float)P2;
            o.y = (float)Q;
            o.z = (float)(1.0e-7 * Pabs) + 1.0e-37f;  // fp32 rounding of 2P (and ~1e-15 of fp64 cancellation in it)
          }
          st[j * TH] = o;
          if (MODE == MH_SED_MBB_T) kjs[j * TH] = kj;
        }
      }
    }

    auto prior_of = [&](double xe) -> double {
      if (mh.prior_type != 1) return 0.0;
      const double a = ((xe - mh.gauss[0]) * (xe - mh.gauss[0])) * inv2var;
      return a < 700.0 ? -a - ln_denom : log_normal_prior(xe, mh.gauss[0], mh.gauss[1]);
    };
    auto exact_lnl = [&](double xe) -> double {
      return k5p_exact<MODE>(mv, mh, p, xe, idx0, idx1, amp0, amp1) + prior_of(xe);
    };
    // rho_j and b_j (the magnitude the rounding error of rho_j scales with) of band j for the step d: the hot form
    // (every |x| < DG_K5P_XMAX) and the general one
    auto rho_of = [&](int j, float d, float n1, float &rho, float &b) {
      if (MODE == MH_SED_MBB_T) {
        const float e = k5p_em1(kb.cf[j] * d);
        const float n2 = fmaf(kjs[j * TH], e, e);  // K_j e, with K_j - 1 stored
        const float inv = k5p_rcp(1.0f + n2);
        rho = (n1 - n2) * inv;
        b = (fabsf(n1) + fabsf(n2)) * inv;
      } else {
        const float xx = d * kb.cf[j];
        rho = k5p_em1(xx);
        b = fmaf(fabsf(rho), fabsf(xx), fabsf(rho));
      }
    };
    auto rho_of_general = [&](int j, float d, float n1, float w1, float &rho, float &b) {
      float w2;
      if (MODE == MH_SED_MBB_T) {
        const float e2 = k5_em1f(kb.cf[j] * d, w2);
        const float n2 = fmaf(kjs[j * TH], e2, e2);
        const float inv = __frcp_rn(1.0f + n2);
        rho = (n1 - n2) * inv;
        b = fmaf(fabsf(n1), w1, fabsf(n2) * w2) * inv;
      } else {
        const float xx = d * kb.cf[j];
        rho = k5_em1f(xx, w2);
        b = fabsf(rho) * w2 * (1.0f + fabsf(xx));
      }
    };

    float epsQ = 1.2e-7f;  // relative error bound of the Q_j
    float kap = kappa;     // T mode: + the relative error the rescaled K - 1 have picked up
    double prior_cur = prior_of(cur), naccept = 0.0;
    double lnl_cur_x = 0.0;  // fp64 lnL(cur) when a fallback has already evaluated it
    bool have_cur_x = false;
#pragma unroll 1
    for (int l = 0; l < nsample; l++) {
      const size_t slot = (size_t)l * mv.P + p;
      const uint64_t gslot = (uint64_t)l * (uint64_t)mv.npix + (uint64_t)(mv.pix_lo + p);  // slot-indexed deviates (Q5)
      const double z = mh.z ? mh.z[slot] : philox_normal(mh.seed, DG_STREAM_MH_Z, gslot);
      double uu = 1.0, u2;
      if (mh.ml_mode != 0) {
        if (mh.u) uu = mh.u[slot];
        else philox_uniform2(mh.seed, DG_STREAM_MH_U, gslot, uu, u2);
      }
      const double x = cur + (0.0 + mh.step * z);       // :414
      const bool oob = x < mh.uni[0] || x > mh.uni[1];  // :415, Q5
      float lam = 0.0f, E1 = 0.0f, W = 0.0f, SP = 0.0f, d = 0.0f, bmax = 0.0f, n1s = 0.0f;
      bool big = false;  // some |x| outside the polynomial's range: the general exp(x) - 1, in a rolled loop
      if (!oob) {
        if (MODE == MH_SED_MBB_T) d = (float)((cur - x) * mh_fast_rcp(x * cur));  // 1/T' - 1/T_cur
        else d = (float)(x - cur);
        big = !(fabsf(d) * kb.cfmax < DG_K5P_XMAX);
        if (!big) {
          float n1 = 0.0f;
          if (MODE == MH_SED_MBB_T) {
            const float e1 = k5p_em1(kb.cref * d);
            n1 = fmaf(kref1, e1, e1);
          }
          n1s = n1;
#pragma unroll UNR
          for (int j = 0; j < NB; j++) {
            float rho, b;
            rho_of(j, d, n1, rho, b);
            const float4 s = st[j * TH];
            st[j * TH].w = rho;  // parked for the state update of an accepted move
            const float ar = fabsf(rho);
            const float m = fmaf(ar, s.y, fabsf(s.x));
            lam = fmaf(rho, fmaf(-rho, s.y, s.x), lam);
            E1 = fmaf(b, m, E1);
            W = fmaf(ar, m, W);
            SP = fmaf(ar, s.z, SP);
            bmax = fmaxf(bmax, b);
          }
        } else {
          float n1 = 0.0f, w1 = 1.0f;
          if (MODE == MH_SED_MBB_T) {
            const float e1 = k5_em1f(kb.cref * d, w1);
            n1 = fmaf(kref1, e1, e1);
          }
          n1s = n1;
#pragma unroll 1
          for (int j = 0; j < NB; j++) {
            float rho, b;
            rho_of_general(j, d, n1, w1, rho, b);
            const float4 s = st[j * TH];
            st[j * TH].w = rho;
            const float ar = fabsf(rho);
            const float m = fmaf(ar, s.y, fabsf(s.x));
            lam = fmaf(rho, fmaf(-rho, s.y, s.x), lam);
            E1 = fmaf(b, m, E1);
            W = fmaf(ar, m, W);
            SP = fmaf(ar, s.z, SP);
            bmax = fmaxf(bmax, b);
          }
        }
      }
      const double prior_new = oob ? prior_cur : prior_of(x);
      const double diff_s = 0.5 * (double)lam + (prior_new - prior_cur);
      float lu = 0.0f, eps_u = 0.0f;
      if (mh.ml_mode != 0) {
        lu = __logf((float)uu);  // :450 (Q4), screened
        eps_u = 1.5e-6f * (1.0f + fabsf(lu));
      }
      const float eps = 0.55f * (2.0f * kap * E1 + (c_round + epsQ) * W + SP) + eps_u + 1.0e-30f;
      const bool certain = fabs(diff_s - (double)lu) > (double)eps;  // false for NaN / inf
      bool accept = diff_s > (double)lu;
      const bool need = !oob && !certain;
      if (need || record) {  // fp64 re-evaluation, this thread only
        const double xe = oob ? cur : x;
        const double lnl_new = exact_lnl(xe);
        if (!have_cur_x) {
          lnl_cur_x = exact_lnl(cur);
          have_cur_x = true;
        }
        const double lnl_old = lnl_cur_x;
        const double diff = lnl_new - lnl_old;
        const bool acc_x = (mh.ml_mode == 0) ? (diff > 0.0) : (diff > log(uu));
        if (record && !oob) {
          if (mh.lnl_trace) mh.lnl_trace[slot] = lnl_new;
          const double scale = fmax(1.0, fmax(fabs(lnl_new), fabs(lnl_old)));
          const bool bound_ok = fabs(diff_s - diff) <= (double)eps + 1e-13 * scale;
          if (!bound_ok || (certain && acc_x != accept)) acc[2] += 1.0;
        }
        if (need) {
          accept = acc_x;
          acc[1] += 1.0;
        }
        if (accept && !oob) lnl_cur_x = lnl_new;  // stays valid for the new point
      } else if (accept && !oob) {
        have_cur_x = false;
      }
      if (oob) accept = false;
      if (accept) {  // the state follows the chain
        cur = x;
        prior_cur = prior_new;
        naccept += 1.0;
        float opmin = 1.0f;
        const float kbm = 1.1f * kap * bmax, c2 = 1.1f * epsQ;
        float inv1 = 1.0f;
        if (MODE == MH_SED_MBB_T) inv1 = __frcp_rn(1.0f + n1s);
#pragma unroll UNR
        for (int j = 0; j < NB; j++) {
          float4 s = st[j * TH];
          const float rho = s.w, op = 1.0f + rho;
          const float q2 = (rho + rho) * s.y, aq2 = fabsf(q2), mm = fabsf(s.x) + aq2;
          // first-order propagation of d rho (kappa b_max), the roundings (3 ulp of the magnitudes) and eQ through
          // 2P' = (1 + rho)(2P - 2 rho Q), with 10 % on the new terms for the fp32 evaluation of the bound itself
          const float t1 = fmaf(aq2, c2, mm * 2.2e-7f), t2 = fmaf(op, s.y + s.y, mm);
          s.z = fmaf(op * 1.000001f, s.z, fmaf(kbm, t2, op * t1)) + 1.0e-37f;
          s.x = op * (s.x - q2);
          s.y = op * op * s.y;
          st[j * TH] = s;
          opmin = fminf(opmin, op);
          if (MODE == MH_SED_MBB_T) kjs[j * TH] *= op * inv1;  // 1 / (1 + n2_j) = (1 + rho_j) / (1 + n1)
        }
        // Q' = (1 + rho)^2 Q: relative error 2 d rho / (1 + rho) + 3 ulp
        epsQ += 4.0f * kap * bmax + 2.5e-7f;
        if (!(opmin > 0.5f)) epsQ = __int_as_float(0x7f800000);  // budget void: the rest of the chain runs in fp64
        if (MODE == MH_SED_MBB_T) {  // relative error of the rescaled K - 1: d rho / (1 + rho), d n1 / (1 + n1), 3 ulp
          kref1 *= inv1;
          kap += 4.0f * kap * bmax + 5.0e-7f;
        }
      }
      if (mh.decisions) mh.decisions[slot] = oob ? 2 : (accept ? 1 : 0);
    }
    for (int s = 0; s < S; s++) cv.idx[mh.nind][(size_t)mh.plane[s] * mv.Ppad + p] = cur;  // :465, :483
    acc[0] += naccept;
  }
  grid_reduce<4>(acc, smem, partials, ticket, out);
}
