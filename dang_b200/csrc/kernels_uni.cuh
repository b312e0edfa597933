// kernels_uni.cuh -- streaming variants of K1 (rhs + blocks), K6 (chi-square) and the full-sky
// sufficient statistics for the case where every component involved has spatially constant
// spectral indices on the planes in play (full-sky-sampled or never-sampled indices: the
// BASELINE headline config c2).  The per-band SEDs then come from the SedTable that
// sed_table_kernel built with the same arithmetic as the per-pixel path, so these kernels do
// no transcendental work at all: they are pure HBM streams over sig/rms with 16-byte loads,
// two pixels per thread, the band loop unrolled for memory-level parallelism.
// Element-wise arithmetic (operation order) is identical to the general kernels.
#pragma once
#include "common.cuh"
#include "kernels_cg.cuh"
#include "kernels_data.cuh"
#include "kernels_mh.cuh"

// Band loops run in batches of DG_UB bands: all 2*DG_UB 16-byte loads of a batch are issued
// before any arithmetic (indices clamped so the loads are unconditional), which is what keeps
// enough bytes in flight per SM to reach the HBM roofline at 1-2 resident blocks per SM.
#define DG_UB 4

// 1/x without the IEEE slow-path branches of the compiler's fp64 division (those BSSY/BSYNC
// regions stop the scheduler from hoisting the next loads): hardware seed + two Newton steps,
// <= 1 ulp for the O(1) noise levels it is applied to.  Inf/NaN lanes (masked pixels, rms = 0)
// are discarded by the callers.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

__device__ __forceinline__ double2 ld2(const double *p) { return ldg_stream2(p); }

// Per-thread SED staging for components whose indices vary from pixel to pixel: the SEDs of the
// thread's two pixels, all bands, go to dynamic shared memory dst[(j*2 + lane)*nthr + tid] (one
// exp per band for a power law, two for a modified blackbody with the reference-frequency Planck
// term hoisted), after which the band loop streams exactly as in the all-uniform case.
__device__ __forceinline__ void sed_pair_to_smem(const ModelView &mv, int ic, int k, int64_t p,
                                                 double *dst, int nthr, int tid) {
  const CompView &cv = mv.comp[ic];
  const size_t kp = (size_t)k * mv.Ppad + p;
  const double2 t0 = cv.nind > 0 ? *reinterpret_cast<const double2 *>(cv.idx[0] + kp) : make_double2(0.0, 0.0);
  const double2 t1 = cv.nind > 1 ? *reinterpret_cast<const double2 *>(cv.idx[1] + kp) : make_double2(0.0, 0.0);
  const bool same = t0.x == t0.y && t1.x == t1.y;
  const SedTable &tab = *mv.tab;
  if (cv.type > 2) {  // free-free, lognormal, cmb: the generic per-band evaluation
    for (int j = 0; j < mv.nbands; j++) {
      const double a = sed_theta(mv, ic, j, t0.x, t1.x, k);
      dst[(size_t)(j * 2 + 0) * nthr + tid] = a;
      dst[(size_t)(j * 2 + 1) * nthr + tid] = same ? a : sed_theta(mv, ic, j, t0.y, t1.y, k);
    }
  } else if (cv.type == 1) {
    for (int j = 0; j < mv.nbands; j++) {
      const double a = sed_powerlaw(mv, ic, j, t0.x);
      dst[(size_t)(j * 2 + 0) * nthr + tid] = a;
      dst[(size_t)(j * 2 + 1) * nthr + tid] = same ? a : sed_powerlaw(mv, ic, j, t0.y);
    }
  } else {
    const double zx = DG_H / (DG_KB * t1.x), zy = DG_H / (DG_KB * t1.y);
    const double ex = exp(zx * cv.nu_ref) - 1.0, ey = same ? ex : exp(zy * cv.nu_ref) - 1.0;
    for (int j = 0; j < mv.nbands; j++) {
      double a, b;
      if (mv.band[j].n == 0) {  // same operation order as sed_mbb
        a = ex / (exp(zx * mv.band[j].nu_c) - 1.0) * exp_scaled(t0.x + 1.0, tab.lnr_hi[ic][j], tab.lnr_lo[ic][j]);
        b = same ? a : ey / (exp(zy * mv.band[j].nu_c) - 1.0) * exp_scaled(t0.y + 1.0, tab.lnr_hi[ic][j], tab.lnr_lo[ic][j]);
      } else {
        a = sed_mbb(mv, ic, j, t0.x, t1.x);
        b = same ? a : sed_mbb(mv, ic, j, t0.y, t1.y);
      }
      dst[(size_t)(j * 2 + 0) * nthr + tid] = a;
      dst[(size_t)(j * 2 + 1) * nthr + tid] = b;
    }
  }
}
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

// do the thread's two pixels carry the same indices of component ic on planes k and k2?
// (always true after a Q+U draw: then the staged SEDs are reused instead of recomputed)
__device__ __forceinline__ bool same_indices(const ModelView &mv, int ic, int k, int k2, int64_t p) {
  const CompView &cv = mv.comp[ic];
  if (sed_uniform(mv, ic, k) != sed_uniform(mv, ic, k2)) return false;
  bool same = true;
  for (int l = 0; l < cv.nind; l++) {
    const double2 a = *reinterpret_cast<const double2 *>(cv.idx[l] + (size_t)k * mv.Ppad + p);
    const double2 b = *reinterpret_cast<const double2 *>(cv.idx[l] + (size_t)k2 * mv.Ppad + p);
    same = same && a.x == b.x && a.y == b.y;
  }
  return same;
}

// K1, uniform-SED form.  Same outputs as rhs_blocks_kernel.
template <int C>
__global__ void __launch_bounds__(DG_THREADS, (C <= 2 ? 2 : 1))
rhs_blocks_uni_kernel(const ModelView mv, const CgView<C> cg, double *partials,
                      unsigned int *ticket, double *out, unsigned nu_mask, const CgInit ci, const PeerComm pc) {
  constexpr int T = C * (C + 1) / 2;
  extern __shared__ double dsed[];  // [slot][band][2][blockDim] for the components in nu_mask
  const int nthr = blockDim.x, tid = threadIdx.x;
  __shared__ double smem[2 * 32];
  __shared__ double ssed[2][C][DG_MAX_BANDS];
  __shared__ double sog[2][DG_MAX_COMPS][DG_MAX_BANDS];
  const int B = mv.nbands;
  for (int i = threadIdx.x; i < 2 * C * B; i += blockDim.x) {
    const int s = i / (C * B), c = (i / B) % C, j = i % B;
    if (s < cg.S && !((nu_mask >> c) & 1u)) ssed[s][c][j] = mv.tab->sed[cg.comp[c] * 3 + cg.plane[s]][j];
  }
  for (int i = threadIdx.x; i < 2 * cg.nog * B; i += blockDim.x) {
    const int s = i / (cg.nog * B), o = (i / B) % cg.nog, j = i % B;
    if (s < cg.S) sog[s][o][j] = mv.tab->sed[cg.og[o] * 3 + cg.plane[s]][j];
  }
  __syncthreads();

  double acc[2] = {0.0, 0.0};
  const int64_t n2 = mv.Ppad / 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)cg.S * mv.Ppad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
    const int64_t p = 2 * e;
    const uchar2 mk = *reinterpret_cast<const uchar2 *>(mv.mask + p);
    const bool use0 = mk.x != 0, use1 = mk.y != 0;
    for (int s = 0; s < cg.S; s++) {
      const int k = cg.plane[s];
      const size_t es = (size_t)s * mv.Ppad + p;
      if (nu_mask) {  // stage this pixel pair's SEDs of the varying-index components
        int slot = 0;
#pragma unroll
        for (int c = 0; c < C; c++)
          if ((nu_mask >> c) & 1u) {
            if (s == 0 || !same_indices(mv, cg.comp[c], k, cg.plane[0], p))
              sed_pair_to_smem(mv, cg.comp[c], k, p, dsed + (size_t)slot * B * 2 * nthr, nthr, tid);
            slot++;
          }
      }
      double2 b[C], f[C], M[T];
#pragma unroll
      for (int c = 0; c < C; c++) b[c] = f[c] = make_double2(0.0, 0.0);
#pragma unroll
      for (int t = 0; t < T; t++) M[t] = make_double2(0.0, 0.0);
      double2 eta = make_double2(0.0, 0.0);
      if (cg.fluct) {
        if (cg.eta) {
          eta = ld2(cg.eta + es);
        } else {
          const uint64_t g0 = (uint64_t)s * (uint64_t)mv.npix + (uint64_t)(mv.pix_lo + p);
          eta.x = use0 ? philox_normal(cg.seed, DG_STREAM_ETA, g0) : 0.0;
          eta.y = use1 ? philox_normal(cg.seed, DG_STREAM_ETA, g0 + 1) : 0.0;
        }
      }
      double2 oa[DG_MAX_COMPS];
      for (int o = 0; o < cg.nog; o++) oa[o] = ld2(mv.comp[cg.og[o]].amp + (size_t)k * mv.Ppad + p);
      for (int j0 = 0; j0 < B; j0 += DG_UB) {
        double2 sg[DG_UB], rm[DG_UB];
#pragma unroll
        for (int u = 0; u < DG_UB; u++) {
          const int j = min(j0 + u, B - 1);
          const size_t off = plane_off(mv, j, k) + p;
          sg[u] = ld2(mv.sig + off);
          rm[u] = ld2(mv.rms + off);
        }
#pragma unroll
        for (int u = 0; u < DG_UB; u++) {
          const int j = j0 + u;
          if (j < B) {
            double2 data = sg[u];
            if (k == 0) {
              data.x = data.x / mv.gain[j];
              data.y = data.y / mv.gain[j];
            }
            for (int o = 0; o < cg.nog; o++) {
              data.x = data.x - oa[o].x * sog[s][o][j];
              data.y = data.y - oa[o].y * sog[s][o][j];
            }
            // one reciprocal per lane: 1/sigma, then 1/sigma^2 and eta/sigma by multiplication
            const double ix = fast_rcp(rm[u].x), iy = fast_rcp(rm[u].y);
            const double wx = ix * ix, wy = iy * iy;
            const double tx = eta.x * ix, ty = eta.y * iy;
            double2 sc[C];
            {
              int slot = 0;
#pragma unroll
              for (int c = 0; c < C; c++) {
                if ((nu_mask >> c) & 1u) {
                  const double *q = dsed + ((size_t)slot * B + j) * 2 * nthr + tid;
                  sc[c] = make_double2(q[0], q[nthr]);
                  slot++;
                } else {
                  sc[c] = make_double2(ssed[s][c][j], ssed[s][c][j]);
                }
              }
            }
#pragma unroll
            for (int c = 0; c < C; c++) {
              b[c].x += data.x * sc[c].x * wx;
              b[c].y += data.y * sc[c].y * wy;
              f[c].x += tx * sc[c].x;
              f[c].y += ty * sc[c].y;
#pragma unroll
              for (int c2 = c; c2 < C; c2++) {
                M[tri<C>(c, c2)].x += sc[c].x * sc[c2].x * wx;
                M[tri<C>(c, c2)].y += sc[c].y * sc[c2].y * wy;
              }
            }
          }
        }
      }
      if (cg.fluct == 1) {
        b[0].x += f[C - 1].x;
        b[0].y += f[C - 1].y;
      } else if (cg.fluct == 2) {
#pragma unroll
        for (int c = 0; c < C; c++) {
          b[c].x += f[c].x;
          b[c].y += f[c].y;
        }
      }
      // masked lanes: zero rows / columns (rms may be anything there, so select, do not scale)
#pragma unroll
      for (int t = 0; t < T; t++) {
        if (!use0) M[t].x = 0.0;
        if (!use1) M[t].y = 0.0;
      }
      double2 xv[C], rv[C];
#pragma unroll
      for (int c = 0; c < C; c++) xv[c] = *reinterpret_cast<const double2 *>(cg.x + c * vs + es);
#pragma unroll
      for (int c = 0; c < C; c++) {
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int c2 = 0; c2 < C; c2++) {
          const double2 mm = M[c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
          ax += mm.x * xv[c2].x;
          ay += mm.y * xv[c2].y;
        }
        rv[c].x = use0 ? b[c].x - ax : 0.0;
        rv[c].y = use1 ? b[c].y - ay : 0.0;
      }
#pragma unroll
      for (int c = 0; c < C; c++) {
        double mx = 0.0, my = 0.0;
#pragma unroll
        for (int c2 = 0; c2 < C; c2++) {
          const double2 mm = M[c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
          mx += mm.x * rv[c2].x;
          my += mm.y * rv[c2].y;
        }
        acc[0] += rv[c].x * rv[c].x + rv[c].y * rv[c].y;
        acc[1] += rv[c].x * mx + rv[c].y * my;
      }
#pragma unroll
      for (int t = 0; t < T; t++) st2(cg.M + t * vs + es, M[t]);
#pragma unroll
      for (int c = 0; c < C; c++) {
        st2(cg.r + c * vs + es, rv[c]);
        if (cg.store_d) st2(cg.d + c * vs + es, rv[c]);
      }
    }
  }
  cg_init_fold(ci, pc, out, grid_reduce<2>(acc, smem, partials, ticket, out));
}

// K6, uniform-SED form (chi-square reduction only; map output stays in chisq_kernel).
template <int NC>
__global__ void __launch_bounds__(DG_THREADS, 3)
chisq_uni_kernel(const ModelView mv, const ChisqView cv, double *partials, unsigned int *ticket,
                 double *out, unsigned nu_mask) {
  extern __shared__ double dsed[];  // [slot][band][2][blockDim] for the components in nu_mask
  const int nthr = blockDim.x, tid = threadIdx.x;
  __shared__ double smem[4 * 32];
  __shared__ double ssed[3][NC][DG_MAX_BANDS];
  const int B = mv.nbands;
  for (int i = threadIdx.x; i < 3 * NC * B; i += blockDim.x) {
    const int k = i / (NC * B), c = (i / B) % NC, j = i % B;
    ssed[k][c][j] = (c < mv.ncomp && k < mv.nmaps && mv.comp[c].in_sky) ? mv.tab->sed[c * 3 + k][j] : 0.0;  // (monopoles: offsets, not sky)
  }
  __syncthreads();
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t n2 = mv.Ppad / 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
    const int64_t p = 2 * e;
    const uchar2 mk = *reinterpret_cast<const uchar2 *>(mv.mask + p);
    const bool use0 = mk.x != 0, use1 = mk.y != 0;
    acc[3] += (use0 ? 1.0 : 0.0) + (use1 ? 1.0 : 0.0);
    if (!use0 && !use1) continue;
    for (int k = cv.k_lo; k <= cv.k_hi; k++) {
      if (nu_mask) {
        int slot = 0;
#pragma unroll
        for (int c = 0; c < NC; c++)
          if (c < mv.ncomp && ((nu_mask >> c) & 1u)) {
            // components whose index maps vary on this plane; planes with a tabulated SED just copy
            if (k > cv.k_lo && same_indices(mv, c, k, k - 1, p)) {
              // staged values of the previous plane are still valid
            } else if (sed_uniform(mv, c, k)) {
              for (int j = 0; j < B; j++) {
                dsed[((size_t)slot * B + j) * 2 * nthr + tid] = ssed[k][c][j];
                dsed[((size_t)slot * B + j) * 2 * nthr + nthr + tid] = ssed[k][c][j];
              }
            } else {
              sed_pair_to_smem(mv, c, k, p, dsed + (size_t)slot * B * 2 * nthr, nthr, tid);
            }
            slot++;
          }
      }
      double2 a[NC];
#pragma unroll
      for (int c = 0; c < NC; c++)
        a[c] = c < mv.ncomp ? ld2(mv.comp[c].amp + (size_t)k * mv.Ppad + p) : make_double2(0.0, 0.0);
      double2 chi = make_double2(0.0, 0.0);
      for (int j0 = 0; j0 < B; j0 += DG_UB) {
        double2 sg[DG_UB], rm[DG_UB];
#pragma unroll
        for (int u = 0; u < DG_UB; u++) {
          const int j = min(j0 + u, B - 1);
          const size_t off = plane_off(mv, j, k) + p;
          sg[u] = ld2(mv.sig + off);
          rm[u] = ld2(mv.rms + off);
        }
#pragma unroll
        for (int u = 0; u < DG_UB; u++) {
          const int j = j0 + u;
          if (j < B) {
            double skx = 0.0, sky = 0.0;
            {
              int slot = 0;
#pragma unroll
              for (int c = 0; c < NC; c++)
                if (c < mv.ncomp) {
                  if ((nu_mask >> c) & 1u) {
                    const double *q = dsed + ((size_t)slot * B + j) * 2 * nthr + tid;
                    skx = skx + a[c].x * q[0];
                    sky = sky + a[c].y * q[nthr];
                    slot++;
                  } else {
                    skx = skx + a[c].x * ssed[k][c][j];
                    sky = sky + a[c].y * ssed[k][c][j];
                  }
                }
            }
            double tx, ty;
            if (k == 0) {
              tx = (sg[u].x - mv.offset[j]) / mv.gain[j] - skx;
              ty = (sg[u].y - mv.offset[j]) / mv.gain[j] - sky;
            } else {
              tx = sg[u].x - skx;
              ty = sg[u].y - sky;
            }
            chi.x = chi.x + (tx * tx) * fast_rcp(rm[u].x * rm[u].x);
            chi.y = chi.y + (ty * ty) * fast_rcp(rm[u].y * rm[u].y);
          }
        }
      }
      acc[k] += (use0 ? chi.x / B : 0.0) + (use1 ? chi.y / B : 0.0);
    }
  }
  grid_reduce<4>(acc, smem, partials, ticket, out);
}

// Full-sky sufficient statistics, uniform-SED form: every component other than the sampled one
// has tabulated SEDs, so data_raw = sig - sum_c2 a_c2 sed_c2 needs no transcendental either.
template <int NC>
__global__ void __launch_bounds__(DG_THREADS, 2)
mh_suffstat_uni_kernel(const ModelView mv, const MhView mh, const MhScalars *ms, double *partials,
                       unsigned int *tickets, double *out) {
  constexpr int NV = 3 * DG_SUFF_CHUNK;
  __shared__ double smem[NV * 32];
  __shared__ double ssed[2][NC][DG_MAX_BANDS];
  __shared__ double s0s[DG_MAX_BANDS];
  const int B = mv.nbands;
  for (int i = threadIdx.x; i < 2 * NC * B; i += blockDim.x) {
    const int s = i / (NC * B), c = (i / B) % NC, j = i % B;
    ssed[s][c][j] = (c < mv.ncomp && c != mh.ic && s < mh.S) ? mv.tab->sed[c * 3 + mh.plane[s]][j] : 0.0;
  }
  if (threadIdx.x < B) s0s[threadIdx.x] = ms->s0[threadIdx.x];
  __syncthreads();
  const int64_t n2 = mv.Ppad / 2;
  const int nchunk = (B + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
  // One (plane, band chunk) combination per block: the statistics are kept per plane so that the chi-square
  // of compute_chisq (per plane) can be formed from them as well (DESIGN.md "statistics cache"), and the
  // S * nchunk sweeps run side by side in ONE grid phase (blocks b, b + ncombo, ... share a combination)
  // instead of one after the other, each with its own reduction tail.  gridDim.x is a multiple of ncombo.
  const int ncombo = mh.S * nchunk;
  const int combo = blockIdx.x % ncombo, sub = blockIdx.x / ncombo, nsub = gridDim.x / ncombo;
  const int64_t stride = (int64_t)nsub * blockDim.x;
  {
    const int s = combo / nchunk, ch = combo % nchunk;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.0;
    const int k = mh.plane[s];
    for (int64_t e = (int64_t)sub * blockDim.x + threadIdx.x; e < n2; e += stride) {
      const int64_t p = 2 * e;
      const uchar2 mk = *reinterpret_cast<const uchar2 *>(mv.mask + p);
      const bool use0 = mk.x != 0, use1 = mk.y != 0;
      if (!use0 && !use1) continue;
      {
        double2 a[NC];
#pragma unroll
        for (int c = 0; c < NC; c++)
          a[c] = c < mv.ncomp ? ld2(mv.comp[c].amp + (size_t)k * mv.Ppad + p) : make_double2(0.0, 0.0);
        double2 am = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < NC; c++)
          if (c == mh.ic) am = a[c];
        double2 sg[DG_SUFF_CHUNK], rm[DG_SUFF_CHUNK];
#pragma unroll
        for (int jj = 0; jj < DG_SUFF_CHUNK; jj++) {
          const int j = min(ch * DG_SUFF_CHUNK + jj, B - 1);
          const size_t off = plane_off(mv, j, k) + p;
          sg[jj] = ld2(mv.sig + off);
          rm[jj] = ld2(mv.rms + off);
        }
#pragma unroll
        for (int jj = 0; jj < DG_SUFF_CHUNK; jj++) {
          const int j = ch * DG_SUFF_CHUNK + jj;
          if (j < B) {
            double2 d = sg[jj];
            if (k == 0) {
              d.x = (d.x - mv.offset[j]) / mv.gain[j];
              d.y = (d.y - mv.offset[j]) / mv.gain[j];
            }
#pragma unroll
            for (int c = 0; c < NC; c++)
              if (c < mv.ncomp && c != mh.ic) {
                d.x = d.x - a[c].x * ssed[s][c][j];
                d.y = d.y - a[c].y * ssed[s][c][j];
              }
            const double ix = fast_rcp(rm[jj].x), iy = fast_rcp(rm[jj].y);
            const double tx = (d.x - am.x * s0s[j]) * ix, ty = (d.y - am.y * s0s[j]) * iy;
            const double ux = am.x * ix, uy = am.y * iy;
            acc[3 * jj + 0] += (use0 ? tx * tx : 0.0) + (use1 ? ty * ty : 0.0);
            acc[3 * jj + 1] += (use0 ? tx * ux : 0.0) + (use1 ? ty * uy : 0.0);
            acc[3 * jj + 2] += (use0 ? ux * ux : 0.0) + (use1 ? uy * uy : 0.0);
          }
        }
      }
    }
    grid_reduce_grouped<NV>(acc, smem, partials, tickets, out, ncombo);  // out[(s * nchunk + ch) * NV + ...]
  }
}

// ======================================================================================
// TMA-staged variants (Blackwell bulk-copy engine).  K1 and the sufficient-statistics pass keep
// ~30 live accumulators per thread, which caps them at two resident blocks per SM -- too few
// 16-byte loads in flight to saturate HBM.  Here one elected thread streams whole [band][plane]
// rows of a 512-pixel tile into a 3-stage shared-memory ring with cp.async.bulk (1-D TMA,
// completion counted on an mbarrier), so ~128 KB per SM are in flight independent of occupancy
// and registers, and the 128 consumer threads read their pixel pair from shared memory.
// The arithmetic is the same, in the same order, as in the LDG kernels above.
// MEASURED (B200, nside 512, 8 bands): 313 us against 285 us for rhs_blocks_uni_kernel -- with the
// loads decoupled from the registers the kernel turns out to be bound by FP64 / issue latency at
// the 16 warps per SM its accumulators allow, not by loads in flight.  It is therefore OFF by
// default (DANG_OPT_TMA) and kept as a measured experiment; parity is covered by
// tests/test_gpu_parity.py::test_tma_staged_k1_matches.
// ======================================================================================
#define DG_TMA_TILE 1024    // pixels per tile (2 per consumer thread)
#define DG_TMA_THREADS 512
#define DG_TMA_STAGES 3
#define DG_TMA_BANDS 2      // bands per stage

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// One work item = (tile, band chunk).  Rows of a stage: [(jj*S + s)*2 + {sig, rms}][DG_TMA_TILE].
struct TmaRing {
  double *buf;       // DG_TMA_STAGES * rows_max * DG_TMA_TILE doubles
  uint64_t *full;    // DG_TMA_STAGES mbarriers
  int rows_max;
};

__device__ __forceinline__ void tma_issue_item(const ModelView &mv, const TmaRing &ring, int stage, int64_t tile,
                                               int chunk, int S, const int *plane) {
  const int64_t p0 = tile * DG_TMA_TILE;
  const int npx = (int)min((int64_t)DG_TMA_TILE, mv.Ppad - p0);
  const int j0 = chunk * DG_TMA_BANDS, nb = min(DG_TMA_BANDS, mv.nbands - j0);
  const uint32_t row_bytes = (uint32_t)npx * 8u;
  double *base = ring.buf + (size_t)stage * ring.rows_max * DG_TMA_TILE;
  mbar_expect_tx(&ring.full[stage], row_bytes * (uint32_t)(nb * S * 2));
  for (int jj = 0; jj < nb; jj++)
    for (int s = 0; s < S; s++) {
      const size_t off = plane_off(mv, j0 + jj, plane[s]) + p0;
      double *dst = base + (size_t)((jj * S + s) * 2) * DG_TMA_TILE;
      bulk_g2s(dst, mv.sig + off, row_bytes, &ring.full[stage]);
      bulk_g2s(dst + DG_TMA_TILE, mv.rms + off, row_bytes, &ring.full[stage]);
    }
}

// K1, all SEDs tabulated, no subtracted components (the headline configuration).
template <int C>
__global__ void __launch_bounds__(DG_TMA_THREADS, 1)
rhs_blocks_tma_kernel(const ModelView mv, const CgView<C> cg, double *partials, unsigned int *ticket, double *out,
                      const CgInit ci, const PeerComm pc) {
  constexpr int T = C * (C + 1) / 2;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ double smem[2 * 32];
  __shared__ double ssed[2][C][DG_MAX_BANDS];
  __shared__ __align__(8) uint64_t full[DG_TMA_STAGES];
  const int B = mv.nbands, S = cg.S, tid = threadIdx.x;
  TmaRing ring;
  ring.buf = reinterpret_cast<double *>(tma_smem);
  ring.full = full;
  ring.rows_max = DG_TMA_BANDS * 2 * 2;
  for (int i = tid; i < 2 * C * B; i += blockDim.x) {
    const int s = i / (C * B), c = (i / B) % C, j = i % B;
    if (s < S) ssed[s][c][j] = mv.tab->sed[cg.comp[c] * 3 + cg.plane[s]][j];
  }
  if (tid == 0) {
    for (int st = 0; st < DG_TMA_STAGES; st++) mbar_init(&full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t ntiles = (mv.Ppad + DG_TMA_TILE - 1) / DG_TMA_TILE;
  const int nchunk = (B + DG_TMA_BANDS - 1) / DG_TMA_BANDS;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t nitems = my_tiles * nchunk;
  auto item_tile = [&](int64_t i) { return (int64_t)blockIdx.x + (i / nchunk) * gridDim.x; };
  if (tid == 0)
    for (int64_t i = 0; i < DG_TMA_STAGES - 1 && i < nitems; i++)
      tma_issue_item(mv, ring, (int)(i % DG_TMA_STAGES), item_tile(i), (int)(i % nchunk), S, cg.plane);

  double acc[2] = {0.0, 0.0};
  const size_t vs = (size_t)S * mv.Ppad;
  double2 b[2][C], f[2][C], M[2][T], eta[2];
  for (int64_t i = 0; i < nitems; i++) {
    const int stage = (int)(i % DG_TMA_STAGES);
    const uint32_t parity = (uint32_t)((i / DG_TMA_STAGES) & 1);
    const int64_t tile = item_tile(i);
    const int chunk = (int)(i % nchunk);
    if (tid == 0 && i + DG_TMA_STAGES - 1 < nitems) {
      const int64_t n = i + DG_TMA_STAGES - 1;  // its stage was consumed in iteration i-1 (barrier below)
      tma_issue_item(mv, ring, (int)(n % DG_TMA_STAGES), item_tile(n), (int)(n % nchunk), S, cg.plane);
    }
    const int64_t p = tile * DG_TMA_TILE + 2 * tid;
    const bool inb = p < mv.Ppad;
    uchar2 mk = make_uchar2(0, 0);
    if (inb) mk = *reinterpret_cast<const uchar2 *>(mv.mask + p);
    const bool use0 = mk.x != 0, use1 = mk.y != 0;
    if (chunk == 0) {
#pragma unroll
      for (int s = 0; s < 2; s++) {
#pragma unroll
        for (int c = 0; c < C; c++) b[s][c] = f[s][c] = make_double2(0.0, 0.0);
#pragma unroll
        for (int t = 0; t < T; t++) M[s][t] = make_double2(0.0, 0.0);
        eta[s] = make_double2(0.0, 0.0);
        if (cg.fluct && s < S && inb) {
          if (cg.eta) {
            eta[s] = ld2(cg.eta + (size_t)s * mv.Ppad + p);
          } else {
            const uint64_t g0 = (uint64_t)s * (uint64_t)mv.npix + (uint64_t)(mv.pix_lo + p);
            eta[s].x = use0 ? philox_normal(cg.seed, DG_STREAM_ETA, g0) : 0.0;
            eta[s].y = use1 ? philox_normal(cg.seed, DG_STREAM_ETA, g0 + 1) : 0.0;
          }
        }
      }
    }
    mbar_wait(&full[stage], parity);
    const double *base = ring.buf + (size_t)stage * ring.rows_max * DG_TMA_TILE;
    const int j0 = chunk * DG_TMA_BANDS, nb = min(DG_TMA_BANDS, B - j0);
    if (inb) {
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (s >= S) continue;
        const int k = cg.plane[s];
#pragma unroll
        for (int jj = 0; jj < DG_TMA_BANDS; jj++) {
          if (jj >= nb) continue;
          const int j = j0 + jj;
          const double *row = base + (size_t)((jj * S + s) * 2) * DG_TMA_TILE + 2 * tid;
          double2 data = *reinterpret_cast<const double2 *>(row);
          const double2 rm = *reinterpret_cast<const double2 *>(row + DG_TMA_TILE);
          if (k == 0) {
            data.x = data.x / mv.gain[j];
            data.y = data.y / mv.gain[j];
          }
          const double ix = fast_rcp(rm.x), iy = fast_rcp(rm.y);
          const double wx = ix * ix, wy = iy * iy;
          const double tx = eta[s].x * ix, ty = eta[s].y * iy;
#pragma unroll
          for (int c = 0; c < C; c++) {
            const double sc = ssed[s][c][j];
            b[s][c].x += data.x * sc * wx;
            b[s][c].y += data.y * sc * wy;
            f[s][c].x += tx * sc;
            f[s][c].y += ty * sc;
#pragma unroll
            for (int c2 = c; c2 < C; c2++) {
              const double sc2 = ssed[s][c2][j];
              M[s][tri<C>(c, c2)].x += sc * sc2 * wx;
              M[s][tri<C>(c, c2)].y += sc * sc2 * wy;
            }
          }
        }
      }
    }
    if (chunk == nchunk - 1 && inb) {
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (s >= S) continue;
        const size_t es = (size_t)s * mv.Ppad + p;
        if (cg.fluct == 1) {
          b[s][0].x += f[s][C - 1].x;
          b[s][0].y += f[s][C - 1].y;
        } else if (cg.fluct == 2) {
#pragma unroll
          for (int c = 0; c < C; c++) {
            b[s][c].x += f[s][c].x;
            b[s][c].y += f[s][c].y;
          }
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
          if (!use0) M[s][t].x = 0.0;
          if (!use1) M[s][t].y = 0.0;
        }
        double2 xv[C], rv[C];
#pragma unroll
        for (int c = 0; c < C; c++) xv[c] = *reinterpret_cast<const double2 *>(cg.x + c * vs + es);
#pragma unroll
        for (int c = 0; c < C; c++) {
          double ax = 0.0, ay = 0.0;
#pragma unroll
          for (int c2 = 0; c2 < C; c2++) {
            const double2 mm = M[s][c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
            ax += mm.x * xv[c2].x;
            ay += mm.y * xv[c2].y;
          }
          rv[c].x = use0 ? b[s][c].x - ax : 0.0;
          rv[c].y = use1 ? b[s][c].y - ay : 0.0;
        }
#pragma unroll
        for (int c = 0; c < C; c++) {
          double mx = 0.0, my = 0.0;
#pragma unroll
          for (int c2 = 0; c2 < C; c2++) {
            const double2 mm = M[s][c <= c2 ? tri<C>(c, c2) : tri<C>(c2, c)];
            mx += mm.x * rv[c2].x;
            my += mm.y * rv[c2].y;
          }
          acc[0] += rv[c].x * rv[c].x + rv[c].y * rv[c].y;
          acc[1] += rv[c].x * mx + rv[c].y * my;
        }
#pragma unroll
        for (int t = 0; t < T; t++) st2(cg.M + t * vs + es, M[s][t]);
#pragma unroll
        for (int c = 0; c < C; c++) {
          st2(cg.r + c * vs + es, rv[c]);
          if (cg.store_d) st2(cg.d + c * vs + es, rv[c]);
        }
      }
    }
    __syncthreads();  // every consumer is done with this stage before it is refilled
  }
  cg_init_fold(ci, pc, out, grid_reduce<2>(acc, smem, partials, ticket, out));
}
