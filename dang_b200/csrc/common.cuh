// common.cuh -- shared device helpers for the dang Gibbs hot path (sm_100a).
//
// Everything here is fp64: the reference's arithmetic is real(dp) throughout and the parity
// contract is 1e-10 relative.  The per-pixel normal-equation blocks are 2x2..4x4, so there is
// no tensor-core work anywhere; the kernels are HBM-streaming with warp-shuffle + block
// reductions (BASELINE.json north_star).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define DG_MAX_BANDS 32
#define DG_MAX_COMPS 8
#define DG_MAX_CG 4      // diffuse components per CG group
#define DG_MAXIND 2
#define DG_THREADS 256

// src/dang_util_mod.f90:12-13 (pi from healpix_types)
#define DG_PI 3.141592653589793238462643383279502884197
#define DG_KB 1.3806503e-23
#define DG_H (1.0545726691251021e-34 * 2.0 * DG_PI)

struct BandView {
  double nu_c;  // Hz
  int n;        // 0 <=> delta band
  int off;      // offset into bp_nu0 / bp_tau0
};

// Device-resident per-handle tables (refreshed whenever an index map changes)
struct SedTable {
  int nonuni[DG_MAX_COMPS * 3][DG_MAXIND];     // scratch of uniform_check_kernel
  int uni[DG_MAX_COMPS * 3];                   // all index maps of comp c constant on plane k
  double sed[DG_MAX_COMPS * 3][DG_MAX_BANDS];  // tabulated SED for uniform (c, k)
  double lnr_hi[DG_MAX_COMPS][DG_MAX_BANDS];   // ln(nu_c / nu_ref) as a double-double
  double lnr_lo[DG_MAX_COMPS][DG_MAX_BANDS];
};

struct CompView {
  int type;     // DANG_COMP_*
  int nind;
  double nu_ref;
  double *amp;             // [nmaps][Ppad]; type 'template': the (normalised) template map
  double *idx[DG_MAXIND];  // each [nmaps][Ppad]
  const double *tamp;      // 'template' / 'monopole' / 'hi_fit': template_amplitudes [3][DG_MAX_BANDS], else null
  int in_sky;              // 0: left out of the sky model ('monopole', src/dang_data_mod.f90:357-361)
};

// Model description handed to kernels by value (lives in the constant bank: every thread reads
// the same entries, which is exactly what the LDC path is for).
struct ModelView {
  int nbands, ncomp, nmaps;
  int64_t P;     // pixels owned by this handle
  int64_t Ppad;  // plane stride (multiple of 64 doubles so every plane is 512 B aligned)
  int64_t pix_lo, npix;  // global offset / full-sky size (RNG slots are global)
  BandView band[DG_MAX_BANDS];
  CompView comp[DG_MAX_COMPS];
  const double *bp_nu0, *bp_tau0;          // flattened bandpass tables, nbp samples in total
  const double *bp_lnr_hi, *bp_lnr_lo;     // [ncomp][nbp] ln(nu0 / nu_ref)
  int nbp;
  const SedTable *tab;
  const double *sig, *rms;   // [nbands][nmaps][Ppad]
  const unsigned char *mask; // [Ppad], 1 = use pixel (mask /= 0 and /= missval)
  double gain[DG_MAX_BANDS], offset[DG_MAX_BANDS];
  double T_cmb;              // the module-global T_CMB (src/dang_util_mod.f90:15; a 'T_cmb' component overwrites it)
};

__device__ __forceinline__ size_t plane_off(const ModelView &mv, int band, int k) {
  return ((size_t)band * mv.nmaps + (size_t)k) * (size_t)mv.Ppad;
}

// streaming loads: read-only path, do not pollute L1
__device__ __forceinline__ double ldg_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ldg_stream2(const double *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p));
  return v;
}

// ---------------------------------------------------------------- device-resident solver / sampler state
// scalars of one solve, device resident (DESIGN.md "CG control")
#define DG_CG_HIST 1024   // passes whose (alpha, beta) are kept for the recompute form
#define DG_CG_MAXM 32     // largest checkpoint interval

struct CgScalars {
  double delta_new, delta_old, alpha, beta, dq;
  double alpha_prev;  // alpha of the pass that ran last (deferred x update, see cg_fused_pass_kernel)
  double converge;
  int iter, i_max, done, pad;
  int ckpt, m;        // recompute form: pass whose state is stored in (r, d); checkpoint interval
  int x_at;           // persistent solve kernel: x (and the amplitude planes) hold the state after pass x_at
  unsigned int gen;   // ... and its grid barrier: passes completed in the running launch
  int k_pred;         // passes the previous solve ran (set when a solve ends, NOT reset by cg_init_update): lets a
                      // solve whose host has not read the previous result yet still predict its last pass
  int pad2;
  double trace[256];
  double ah[DG_CG_HIST], bh[DG_CG_HIST];  // alpha_i, beta_i used IN pass i (1-based)
};

#define DG_MH_THREADS 128
#define DG_SUFF_CHUNK 4   // bands per register-resident accumulator chunk

struct MhView {
  int ic;          // component being sampled
  int nind;        // which of its indices
  int S;           // planes map_inds(1)..map_inds(2)
  int plane[2];    // 0-based
  int nsample, ml_mode, lnl_type, prior_type;
  int is_synch;    // label == 'synch' (eval_jeffreys_prior, src/dang_lnl_mod.f90:289)
  double gauss[2], uni[2], step;
  const double *z, *u;        // injected deviates (device copies) or nullptr -> Philox(seed)
  uint64_t seed;
  unsigned char *decisions;   // optional instrumentation
  double *lnl_trace;
  // device RNG of the full-sky chain kernels: Philox stream ids (0: the draw's own, DG_STREAM_MH_Z / _U) and the
  // slot of proposal 0 -- the step-size tuner numbers its proposals blk * nsample + l on its own streams
  int rng_z_stream, rng_u_stream;
  long long rng_slot0;
};

struct MhScalars {
  double sample[DG_MAXIND], theta[DG_MAXIND];
  double lnl_old, accept;
  int l, phase, skip, pad;
  double sed[DG_MAX_BANDS];  // SED of the proposal per band (streaming lnL kernel)
  double s0[DG_MAX_BANDS];   // SED at the chain's starting point (sufficient statistics)
  double chisq[2];           // chi-square per sampled plane at the chain's final state (statistics form)
};

// ---------------------------------------------------------------- SEDs
// eval_sed, src/dang_component_mod.f90:778-813; evaluate_powerlaw :886-918; evaluate_mbb :920-958.
//
// (nu/nu_ref)**beta is evaluated as exp(beta * ln(nu/nu_ref)) with the logarithm precomputed on
// the host in extended precision and carried as a double-double (hi, lo), so the result is
// within ~1 ulp of a correctly rounded pow at a third of its cost; the Planck factor keeps the
// reference's exp()-1 form and operation order.
__device__ __forceinline__ double exp_scaled(double beta, double l_hi, double l_lo) {
  const double p = beta * l_hi;
  const double e = fma(beta, l_hi, -p) + beta * l_lo;
  const double r = exp(p);
  return fma(r, e, r);
}

__device__ __forceinline__ double sed_powerlaw(const ModelView &mv, int ic, int band, double beta) {
  const BandView &b = mv.band[band];
  const SedTable &t = *mv.tab;
  if (b.n == 0) return exp_scaled(beta, t.lnr_hi[ic][band], t.lnr_lo[ic][band]);
  const double *lh = mv.bp_lnr_hi + (size_t)ic * mv.nbp, *ll = mv.bp_lnr_lo + (size_t)ic * mv.nbp;
  double spectrum = 0.0;
  for (int i = 0; i < b.n; i++) {
    if (mv.bp_nu0[b.off + i] == 0.0) continue;
    spectrum = spectrum + mv.bp_tau0[b.off + i] * exp_scaled(beta, lh[b.off + i], ll[b.off + i]);
  }
  return spectrum;
}

__device__ __forceinline__ double sed_mbb(const ModelView &mv, int ic, int band, double beta,
                                          double td) {
  const BandView &b = mv.band[band];
  const SedTable &t = *mv.tab;
  const double nu_ref = mv.comp[ic].nu_ref;
  const double z = DG_H / (DG_KB * td);
  const double eref = exp(z * nu_ref) - 1.0;
  if (b.n == 0)
    return eref / (exp(z * b.nu_c) - 1.0) * exp_scaled(beta + 1.0, t.lnr_hi[ic][band], t.lnr_lo[ic][band]);
  const double *lh = mv.bp_lnr_hi + (size_t)ic * mv.nbp, *ll = mv.bp_lnr_lo + (size_t)ic * mv.nbp;
  double spectrum = 0.0;
  for (int i = 0; i < b.n; i++) {
    const double nu0 = mv.bp_nu0[b.off + i];
    if (nu0 == 0.0) continue;
    spectrum = spectrum + mv.bp_tau0[b.off + i] * eref / (exp(z * nu0) - 1.0) *
                              exp_scaled(beta + 1.0, lh[b.off + i], ll[b.off + i]);
  }
  return spectrum;
}

// evaluate_freefree, src/dang_component_mod.f90:1001-1040 (T_e is the only index)
__device__ __forceinline__ double ff_gaunt(double nu, double T_e) {
  return log(exp(5.960 - sqrt(3.0) / DG_PI * log(1.0 * nu / 1.e9 * pow(T_e / 1.e4, -1.5))) + 2.71828);
}
__device__ __forceinline__ double sed_freefree(const ModelView &mv, int ic, int band, double T_e) {
  const BandView &b = mv.band[band];
  const double nu_ref = mv.comp[ic].nu_ref;
  const double S_ref = ff_gaunt(nu_ref, T_e);
  if (b.n == 0) {
    const double r = b.nu_c / nu_ref;
    return ff_gaunt(b.nu_c, T_e) / S_ref * (1.0 / (r * r));
  }
  double spectrum = 0.0;
  for (int i = 0; i < b.n; i++) {
    const double nu0 = mv.bp_nu0[b.off + i];
    if (nu0 == 0.0) continue;
    const double r = nu0 / nu_ref;
    spectrum = spectrum + mv.bp_tau0[b.off + i] * ff_gaunt(nu0, T_e) / S_ref * (1.0 / (r * r));
  }
  return spectrum;
}

// evaluate_lognormal, src/dang_component_mod.f90:960-999 (nu_p [GHz], w_ame)
__device__ __forceinline__ double sed_lognormal(const ModelView &mv, int ic, int band, double nu_p,
                                                double w_ame) {
  const BandView &b = mv.band[band];
  const double nu_ref = mv.comp[ic].nu_ref;
  if (b.n == 0) {
    const double t = log(b.nu_c / (nu_p * 1e9)) / w_ame, r = nu_ref / b.nu_c;
    return exp(-0.5 * (t * t)) * (r * r);
  }
  double spectrum = 0.0;
  for (int i = 0; i < b.n; i++) {
    const double nu0 = mv.bp_nu0[b.off + i];
    if (nu0 == 0.0) continue;
    const double t = log(nu0 / (nu_p * 1e9)) / w_ame, r = nu_ref / nu0;
    spectrum = spectrum + mv.bp_tau0[b.off + i] * exp(-0.5 * (t * t)) * (r * r);
  }
  return spectrum;
}

// type 'cmb': 1/a2t(bp(band)), src/dang_component_mod.f90:799-800 with a2t from dang_bp_mod.f90:211-243
__device__ __forceinline__ double sed_cmb(const ModelView &mv, int band) {
  const BandView &b = mv.band[band];
  const double T_CMB = mv.T_cmb;  // src/dang_util_mod.f90:15 (2.7255 until a 'T_cmb' draw changes it)
  double sum = 0.0;
  if (b.n == 0) {
    const double y = (b.nu_c > 1e7) ? (DG_H * b.nu_c) / (DG_KB * T_CMB) : (DG_H * b.nu_c * 1e9) / (DG_KB * T_CMB);
    sum = ((exp(y) - 1.0) * (exp(y) - 1.0)) / ((y * y) * exp(y));
  } else {
    for (int i = 0; i < b.n; i++) {
      const double nu0 = mv.bp_nu0[b.off + i];
      if (nu0 == 0.0) continue;
      const double y = (nu0 > 1e7) ? (DG_H * nu0) / (DG_KB * T_CMB) : (DG_H * nu0 * 1e9) / (DG_KB * T_CMB);
      sum = sum + mv.bp_tau0[b.off + i] * ((exp(y) - 1.0) * (exp(y) - 1.0)) / ((y * y) * exp(y));
    }
  }
  return 1.0 / sum;
}

// evaluate_T_cmb :815-848 and evaluate_hi_fit :850-884 (one body): B_nu(nu, T) / compute_bnu_prime_RJ(nu) * 1e6,
// B_nu src/dang_component_mod.f90:745-752, compute_bnu_prime_RJ src/dang_bp_mod.f90:160-168; operation order kept.
#define DG_C 2.99792458e8
__device__ __forceinline__ double planck_rj_at(double nu, double T) {
  const double B = ((2.0 * DG_H * (nu * nu * nu)) / (DG_C * DG_C)) * (1.0 / (exp((DG_H * nu) / (DG_KB * T)) - 1.0));
  return B / (2.0 * DG_KB * (nu * nu) / (DG_C * DG_C));
}
__device__ __forceinline__ double sed_planck_rj(const ModelView &mv, int band, double T) {
  const BandView &b = mv.band[band];
  double spectrum = 0.0;
  if (b.n == 0) {
    spectrum = planck_rj_at(b.nu_c, T);
  } else {
    for (int i = 0; i < b.n; i++) {
      const double nu0 = mv.bp_nu0[b.off + i];
      if (nu0 == 0.0) continue;
      spectrum = spectrum + mv.bp_tau0[b.off + i] * planck_rj_at(nu0, T);
    }
  }
  return spectrum * 1e6;
}

// SED of component ic in `band` for explicit parameters (a Metropolis proposal, or a pixel).  `k`: plane, needed
// by 'hi_fit' only (its SED carries template_amplitudes(band, plane), see sed_table_kernel).
__device__ __forceinline__ double sed_theta(const ModelView &mv, int ic, int band, double t0,
                                            double t1, int k = 0) {
  switch (mv.comp[ic].type) {
    case 1: return sed_powerlaw(mv, ic, band, t0);
    case 2: return sed_mbb(mv, ic, band, t0, t1);
    case 3: return sed_freefree(mv, ic, band, t0);
    case 4: return sed_lognormal(mv, ic, band, t0, t1);
    case 6: return 0.0;  // 'template': always tabulated (sed_table_kernel copies template_amplitudes)
    case 7: return sed_planck_rj(mv, band, t0);                                           // 'T_cmb'
    case 8: return 0.0;  // 'monopole': always tabulated
    case 9: return mv.comp[ic].tamp[k * DG_MAX_BANDS + band] * sed_planck_rj(mv, band, t0);  // 'hi_fit'
    default: return sed_cmb(mv, band);
  }
}

// SED of component ic at a pixel of plane k.  When every index map of the component is constant
// over this handle's pixels of that plane (always true for full-sky-sampled or never-sampled
// indices) the per-band value was tabulated once by sed_table_kernel with the very same
// arithmetic, and the pixel kernels become pure HBM streams.
__device__ __forceinline__ bool sed_uniform(const ModelView &mv, int ic, int k) {
  return mv.tab->uni[ic * 3 + k] != 0;
}
__device__ __forceinline__ double sed_eval(const ModelView &mv, int ic, int k, int band, double t0,
                                           double t1) {
  if (sed_uniform(mv, ic, k)) return mv.tab->sed[ic * 3 + k][band];
  return sed_theta(mv, ic, band, t0, t1, k);
}

// ---------------------------------------------------------------- Philox4x32-10
// Stream definition shared with the test oracle (DESIGN.md "RNG"):
// counter = {slot_lo, slot_hi, stream, 'DANG'}, key = {seed_lo, seed_hi}.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__host__ __device__ __forceinline__ void philox_uniform2(uint64_t seed, uint32_t stream,
                                                         uint64_t slot, double &u1, double &u2) {
  uint32_t c[4] = {(uint32_t)slot, (uint32_t)(slot >> 32), stream, 0x44414e47u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t a = ((uint64_t)c[0] << 32) | c[1];
  const uint64_t b = ((uint64_t)c[2] << 32) | c[3];
  u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// rand_normal(0,1), src/dang_util_mod.f90:100-110: Box-Muller, sine branch
__host__ __device__ __forceinline__ double philox_normal(uint64_t seed, uint32_t stream,
                                                         uint64_t slot) {
  double u1, u2;
  philox_uniform2(seed, stream, slot, u1, u2);
  const double r = sqrt(-2.0 * log(u1));
#ifdef __CUDA_ARCH__
  return r * sinpi(2.0 * u2);  // == sin(2 pi u2) up to rounding, without reducing by an inexact pi
#else
  return r * sin(2.0 * DG_PI * u2);
#endif
}

enum { DG_STREAM_ETA = 1, DG_STREAM_MH_Z = 2, DG_STREAM_MH_U = 3, DG_STREAM_TUNE_Z = 4, DG_STREAM_TUNE_U = 5, DG_STREAM_GAIN = 6 };

// ---------------------------------------------------------------- cross-rank scalar exchange
// Every rank owns a mailbox in its own HBM, mapped into all peers with CUDA IPC.  An exchange is: write
// my row of <= DG_MAIL_VALS doubles into slot (seq mod DG_MAIL_SLOTS) of EVERY rank's mailbox over
// NVLink, then wait for every rank's row in my own mailbox and copy the rows out in rank order (=> the
// sums formed from them are bit-identical everywhere).  The rows travel in the low-latency form NCCL's
// LL protocol uses: every 8-byte store carries 4 bytes of payload and the 4-byte sequence number of the
// exchange, so a word is valid exactly when its flag matches -- no fence, no separate "ready" store,
// one NVLink write latency.  One warp does it inside the kernel that produced the partial sums, so a CG
// iteration on N GPUs is still ONE launch.
// All ranks execute the same sequence of exchanges (they take identical decisions from identical
// sums), so sequence numbers stay aligned; a rank can lead by at most one exchange, and a slot is
// reused only DG_MAIL_SLOTS exchanges later.
#define DG_MAIL_VALS 256
#define DG_MAIL_SLOTS 4
#define DG_MAX_RANKS 32

struct Mail {
  unsigned long long w[2 * DG_MAIL_VALS];  // {payload half (low 32 bits), flag (high 32 bits)} x 2 per double
};

struct PeerComm {
  int nranks, rank;
  unsigned long long *seq;       // device counter of exchanges done by this rank
  int *error;                    // set on timeout; lives in mapped pinned host memory, the host checks it at
                                 // the end of every API call (a timed-out exchange is fatal: DANG_GPU_ENCCL)
  long long timeout_cycles;      // DANG_GPU_PEER_TIMEOUT_S (default 30 s) in SM clocks
  Mail *box[DG_MAX_RANKS];       // box[g]: rank g's mailbox [DG_MAIL_SLOTS][nranks] (peer-mapped)
};

__device__ __forceinline__ void st_ll(unsigned long long *p, unsigned int data, unsigned int flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ void ld_ll(const unsigned long long *p, unsigned int &data, unsigned int &flag) {
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(data), "=r"(flag) : "l"(p) : "memory");
}

// Called by ONE full warp.  local[cnt] -> gathered[nranks][cnt] (rank order).
__device__ __forceinline__ void peer_exchange(const PeerComm &pc, const double *local, int cnt,
                                              double *gathered) {
  const int lane = threadIdx.x & 31;
  __syncwarp();  // `local` was written by one lane of this warp
  const unsigned long long seq = *pc.seq + 1;
  const unsigned int flag = (unsigned int)seq;  // never 0 in practice; a stale slot holds seq - DG_MAIL_SLOTS
  const int slot = (int)(seq % DG_MAIL_SLOTS);
  const int nw = 2 * cnt, total = pc.nranks * nw;
  // send: word k of my row to every rank (item = peer * nw + k, strided over the warp)
  for (int it = lane; it < total; it += 32) {
    const int g = it / nw, k = it - g * nw;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(local[k >> 1]);
    const unsigned int half = (k & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits;
    Mail *dst = pc.box[g] + (size_t)slot * pc.nranks + pc.rank;  // my row in rank g's box
    st_ll(&dst->w[k], half, flag);
  }
  // receive: every rank's row from my own box (all 32 lanes stay in step for the shuffle)
  bool ok = true;
  const long long t0 = clock64();
  for (int base = 0; base < total; base += 32) {
    const int it = base + lane;
    const bool active = it < total;
    const int g = active ? it / nw : 0, k = active ? it - g * nw : 0;
    unsigned int half = 0, f = flag;
    if (active) {
      const Mail *src = pc.box[pc.rank] + (size_t)slot * pc.nranks + g;
      ld_ll(&src->w[k], half, f);
      while (f != flag) {
        if (clock64() - t0 > pc.timeout_cycles) {  // a peer died: fail loudly instead of hanging.  The missing
          ok = false;                              // words become a NaN, so every sum formed from this exchange
          half = 0xfff80000u;                      // is poisoned (a CG solve stops at once: NaN > converge is
          break;                                   // false) and the host turns the flag into DANG_GPU_ENCCL
        }
        ld_ll(&src->w[k], half, f);
      }
    }
    __syncwarp();
    // the two halves of a double sit in neighbouring lanes (nw is even, so lane parity == k parity)
    const unsigned int other = __shfl_xor_sync(0xffffffffu, half, 1);
    if (active && !(k & 1)) {
      const unsigned long long bits = ((unsigned long long)other << 32) | half;
      gathered[g * cnt + (k >> 1)] = __longlong_as_double((long long)bits);
    }
  }
  if (!ok) {
    *(volatile int *)pc.error = 1;
    __threadfence_system();
  }
  __syncwarp();
  if (lane == 0) *pc.seq = seq;
  __syncwarp();
}

static __global__ void peer_exchange_kernel(PeerComm pc, const double *local, int cnt, double *gathered) {
  peer_exchange(pc, local, cnt, gathered);
}
// `reps` back-to-back exchanges of `cnt` doubles by one warp: the latency floor of a scalar exchange (dang_gpu_comm_probe)
static __global__ void peer_exchange_probe_kernel(PeerComm pc, double *local, int cnt, double *gathered, int reps) {
  for (int i = 0; i < reps; i++) {
    if (threadIdx.x < cnt) local[threadIdx.x] = (double)(i + pc.rank);
    peer_exchange(pc, local, cnt, gathered);
  }
}

// Small results go back to the host through stores into mapped pinned memory, not through the
// copy engine: a cudaMemcpyAsync of a few bytes queues behind whatever bulk download (amplitude
// maps on the d2h stream) the device-to-host engine is busy with and would stall the compute stream.
static __global__ void readback_kernel(unsigned int *dst_host, const unsigned int *src, int nwords) {
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst_host[i] = src[i];
  __threadfence_system();
}

// three small device regions -> one pinned snapshot slot (dang_gpu_iteration_mark)
static __global__ void snapshot_kernel(unsigned int *d0, const unsigned int *s0, int n0, unsigned int *d1, const unsigned int *s1,
                                       int n1, unsigned int *d2, const unsigned int *s2, int n2) {
  for (int i = threadIdx.x; i < n0; i += blockDim.x) d0[i] = s0[i];
  for (int i = threadIdx.x; i < n1; i += blockDim.x) d1[i] = s1[i];
  for (int i = threadIdx.x; i < n2; i += blockDim.x) d2[i] = s2[i];
  __threadfence_system();
}

// ---------------------------------------------------------------- deterministic reductions
// Two-stage tree: per-thread serial partial -> warp shuffle -> shared-memory tree over warps
// -> one partial per block in global memory -> fixed-order final sum by the block that
// finishes last.  The result depends only on (grid, block) sizes, never on scheduling.
template <int NV>
__device__ __forceinline__ void warp_reduce(double (&v)[NV]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}

// Reduces v over the block; valid in warp 0 afterwards (all lanes).
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *smem /* NV*32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  warp_reduce<NV>(v);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; i++) smem[i * 32 + warp] = v[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = (lane < nwarp) ? smem[i * 32 + lane] : 0.0;
    warp_reduce<NV>(v);
  }
}

// Block partial -> global; the last block to arrive sums all partials in block order and
// writes `out[NV]`.  `ticket` must be zero on entry and is reset to zero on exit.
// Returns true in the threads of warp 0 of the last block (so a caller can append work).
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *smem, double *partials,
                                            unsigned int *ticket, double *out) {
  __shared__ bool is_last;
  block_reduce<NV>(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) partials[(size_t)i * gridDim.x + blockIdx.x] = v[i];
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double a = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      a += __ldcg(&partials[(size_t)i * gridDim.x + b]);
    acc[i] = a;
  }
  __syncthreads();  // smem reuse
  block_reduce<NV>(acc, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) out[i] = acc[i];
    *ticket = 0;
    __threadfence();
  }
  return threadIdx.x < 32;
}

// Grouped variant: block b belongs to group (b % ngroup) and the partials of each group are summed
// separately (in block order) by the last block: out[g * NV + i].  gridDim.x must be a multiple of
// ngroup.  Lets one launch run several independent reductions side by side instead of one after the
// other (each with its own grid-wide tail).
template <int NV>
__device__ __forceinline__ bool grid_reduce_grouped(double (&v)[NV], double *smem, double *partials,
                                                    unsigned int *ticket, double *out, int ngroup) {
  __shared__ bool is_last_g;
  block_reduce<NV>(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) partials[(size_t)i * gridDim.x + blockIdx.x] = v[i];
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last_g = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last_g) return false;
  __threadfence();
  const unsigned int nsub = gridDim.x / ngroup;
  for (int g = 0; g < ngroup; g++) {
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double a = 0.0;
      for (unsigned int k = threadIdx.x; k < nsub; k += blockDim.x)
        a += __ldcg(&partials[(size_t)i * gridDim.x + (size_t)k * ngroup + g]);
      acc[i] = a;
    }
    __syncthreads();  // smem reuse
    block_reduce<NV>(acc, smem);
    if (threadIdx.x == 0)
#pragma unroll
      for (int i = 0; i < NV; i++) out[g * NV + i] = acc[i];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *ticket = 0;
    __threadfence();
  }
  return threadIdx.x < 32;
}
