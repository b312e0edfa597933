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
      const double a = sed_theta(mv, ic, j, t0.x, t1.x);
      dst[(size_t)(j * 2 + 0) * nthr + tid] = a;
      dst[(size_t)(j * 2 + 1) * nthr + tid] = same ? a : sed_theta(mv, ic, j, t0.y, t1.y);
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
                      unsigned int *ticket, double *out, unsigned nu_mask) {
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
  grid_reduce<2>(acc, smem, partials, ticket, out);
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
    ssed[k][c][j] = (c < mv.ncomp && k < mv.nmaps) ? mv.tab->sed[c * 3 + k][j] : 0.0;
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
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int nchunk = (B + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
  for (int ch = 0; ch < nchunk; ch++) {
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
      const int64_t p = 2 * e;
      const uchar2 mk = *reinterpret_cast<const uchar2 *>(mv.mask + p);
      const bool use0 = mk.x != 0, use1 = mk.y != 0;
      if (!use0 && !use1) continue;
      for (int s = 0; s < mh.S; s++) {
        const int k = mh.plane[s];
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
    grid_reduce<NV>(acc, smem, partials + (size_t)ch * NV * gridDim.x, tickets + ch, out + ch * NV);
    __syncthreads();
  }
}
