// kernels_stream.cuh -- the two passes over sig / rms of the headline configuration (every component with
// tabulated SEDs, nothing subtracted from the data) as asynchronous streams:
//
//   rhs_blocks_ring_kernel   K1: compute_rhs + compute_sample_vector + block build + first residual
//                            (src/dang_cg_mod.f90:326-596, :913-1100, the SED part of :598-911)
//   mh_suffstat_ring_kernel  the per-band sufficient statistics of a full-sky draw, which also serve the two
//                            chi-squares around it (src/dang_sample_mod.f90:282-324, dang_data_mod.f90:494-526)
//
// Same arithmetic, same order, same grid reduction as rhs_blocks_uni_kernel / mh_suffstat_uni_kernel
// (kernels_uni.cuh) -- bit-identical results -- but the loads are decoupled from the registers: every thread
// copies its own next batches of sig / rms (4 bands x {sig, rms} x 16 bytes) into a private three-stage ring
// in shared memory with cp.async (LDGSTS, 16-byte, L1-bypassing) two batches ahead of the arithmetic.  No
// thread ever reads another thread's slots, so there is no block-level barrier in the loop, only
// cp.async.wait_group.  The LDG forms keep 8 loads in flight per thread only while the thread is stalled on
// them and nothing during its FP64 phase (Philox + Box-Muller + the band math are ~45 % of K1's issue slots);
// at the 16 warps per SM the accumulators allow, that left K1 at 0.63 and the statistics pass at 0.68 of the
// measured HBM peak.  Here 2 x 128 bytes per thread (128 KB per SM) stay in flight whatever the warps are doing.
#pragma once
#include "cp_async.cuh"
#include "kernels_uni.cuh"

#define DG_RING_STAGES 3
#define DG_RING_SLOTS (2 * DG_UB)  // 16-byte slots per stage and thread: {sig, rms} x DG_UB bands

// A thread's work is the sequence of items (pixel pair e, plane s, band batch jb), e = e0, e0 + stride, ...;
// item i -> stage i % DG_RING_STAGES.  Slot layout [stage][slot][thread] keeps the 16-byte LDS conflict-free.
struct RingCursor {
  int64_t e;   // pixel pair
  int s, jb;   // plane number within the solve / draw, band batch
};
__device__ __forceinline__ void ring_advance(RingCursor &c, int S, int nbatch, int64_t stride) {
  if (++c.jb == nbatch) {
    c.jb = 0;
    if (++c.s == S) {
      c.s = 0;
      c.e += stride;
    }
  }
}
__device__ __forceinline__ void ring_issue(const ModelView &mv, const RingCursor &c, const int *plane, int64_t n2,
                                           double2 *ring, int stage) {
  if (c.e < n2) {
    const int k = plane[c.s];
    const int64_t p = 2 * c.e;
    double2 *dst = ring + (size_t)stage * DG_RING_SLOTS * blockDim.x + threadIdx.x;
#pragma unroll
    for (int u = 0; u < DG_UB; u++) {
      const int j = min(c.jb * DG_UB + u, mv.nbands - 1);
      const size_t off = plane_off(mv, j, k) + p;
      cp_async16(dst + (size_t)(2 * u) * blockDim.x, mv.sig + off);
      cp_async16(dst + (size_t)(2 * u + 1) * blockDim.x, mv.rms + off);
    }
  }
  cp_async_commit();  // (an empty group when the thread has run out of items: keeps the group count in step)
}

// ---------------------------------------------------------------- K1
template <int C>
__global__ void __launch_bounds__(DG_THREADS, 2)
rhs_blocks_ring_kernel(const ModelView mv, const CgView<C> cg, double *partials, unsigned int *ticket, double *out,
                       const CgInit ci, const PeerComm pc) {
  constexpr int T = C * (C + 1) / 2;
  extern __shared__ __align__(16) unsigned char ring_raw[];
  double2 *ring = reinterpret_cast<double2 *>(ring_raw);
  __shared__ double smem[2 * 32];
  __shared__ double ssed[2][C][DG_MAX_BANDS];
  __shared__ int splane[2];
  const int B = mv.nbands, S = cg.S;
  for (int i = threadIdx.x; i < 2 * C * B; i += blockDim.x) {
    const int s = i / (C * B), c = (i / B) % C, j = i % B;
    if (s < S) ssed[s][c][j] = mv.tab->sed[cg.comp[c] * 3 + cg.plane[s]][j];
  }
  if (threadIdx.x < 2) splane[threadIdx.x] = cg.plane[threadIdx.x];
  __syncthreads();

  double acc[2] = {0.0, 0.0};
  const int64_t n2 = mv.Ppad / 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)S * mv.Ppad;
  const int nbatch = (B + DG_UB - 1) / DG_UB;
  RingCursor cur{(int64_t)blockIdx.x * blockDim.x + threadIdx.x, 0, 0}, pre = cur;
#pragma unroll
  for (int st = 0; st < DG_RING_STAGES - 1; st++) {
    ring_issue(mv, pre, splane, n2, ring, st);
    ring_advance(pre, S, nbatch, stride);
  }
  int stage = 0, pstage = DG_RING_STAGES - 1;
  double2 b[C], f[C], M[T], xv[C], eta = make_double2(0.0, 0.0);
  bool use0 = false, use1 = false;
  while (cur.e < n2) {
    // keep two batches in flight behind the one about to be consumed
    ring_issue(mv, pre, splane, n2, ring, pstage);
    ring_advance(pre, S, nbatch, stride);
    pstage = pstage + 1 == DG_RING_STAGES ? 0 : pstage + 1;
    const int64_t p = 2 * cur.e;
    const int s = cur.s, k = splane[s];
    const size_t es = (size_t)s * mv.Ppad + p;
    if (cur.jb == 0) {
      if (s == 0) {
        const uchar2 mk = *reinterpret_cast<const uchar2 *>(mv.mask + p);
        use0 = mk.x != 0;
        use1 = mk.y != 0;
      }
#pragma unroll
      for (int c = 0; c < C; c++) b[c] = f[c] = make_double2(0.0, 0.0);
#pragma unroll
      for (int t = 0; t < T; t++) M[t] = make_double2(0.0, 0.0);
      // the warm-start x of this (pair, plane) is needed only after the last band batch: load it now
#pragma unroll
      for (int c = 0; c < C; c++) xv[c] = *reinterpret_cast<const double2 *>(cg.x + c * vs + es);
      eta = make_double2(0.0, 0.0);
      if (cg.fluct) {
        if (cg.eta) {
          eta = ld2(cg.eta + es);
        } else {
          const uint64_t g0 = (uint64_t)s * (uint64_t)mv.npix + (uint64_t)(mv.pix_lo + p);
          eta.x = use0 ? philox_normal(cg.seed, DG_STREAM_ETA, g0) : 0.0;
          eta.y = use1 ? philox_normal(cg.seed, DG_STREAM_ETA, g0 + 1) : 0.0;
        }
      }
    }
    cp_async_wait<DG_RING_STAGES - 1>();  // the oldest group (this item's) has landed
    const double2 *src = ring + (size_t)stage * DG_RING_SLOTS * blockDim.x + threadIdx.x;
    const int j0 = cur.jb * DG_UB;
#pragma unroll
    for (int u = 0; u < DG_UB; u++) {
      const int j = j0 + u;
      if (j < B) {
        double2 data = src[(size_t)(2 * u) * blockDim.x];
        const double2 rm = src[(size_t)(2 * u + 1) * blockDim.x];
        if (k == 0) {  // :369-373
          data.x = data.x / mv.gain[j];
          data.y = data.y / mv.gain[j];
        }
        // one reciprocal per lane: 1/sigma, then 1/sigma^2 and eta/sigma by multiplication
        const double ix = fast_rcp(rm.x), iy = fast_rcp(rm.y);
        const double wx = ix * ix, wy = iy * iy;
        const double tx = eta.x * ix, ty = eta.y * iy;
        double2 sc[C];
#pragma unroll
        for (int c = 0; c < C; c++) sc[c] = make_double2(ssed[s][c][j], ssed[s][c][j]);
#pragma unroll
        for (int c = 0; c < C; c++) {
          b[c].x += data.x * sc[c].x * wx;  // :489-494
          b[c].y += data.y * sc[c].y * wy;
          f[c].x += tx * sc[c].x;           // :1030-1042
          f[c].y += ty * sc[c].y;
#pragma unroll
          for (int c2 = c; c2 < C; c2++) {
            M[tri<C>(c, c2)].x += sc[c].x * sc[c2].x * wx;
            M[tri<C>(c, c2)].y += sc[c].y * sc[c2].y * wy;
          }
        }
      }
    }
    if (cur.jb == nbatch - 1) {  // last batch of this (pair, plane): first residual, stores
      if (cg.fluct == 1) {       // Q1
        b[0].x += f[C - 1].x;
        b[0].y += f[C - 1].y;
      } else if (cg.fluct == 2) {
#pragma unroll
        for (int c = 0; c < C; c++) {
          b[c].x += f[c].x;
          b[c].y += f[c].y;
        }
      }
#pragma unroll
      for (int t = 0; t < T; t++) {  // masked lanes: zero rows / columns
        if (!use0) M[t].x = 0.0;
        if (!use1) M[t].y = 0.0;
      }
      double2 rv[C];
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
    ring_advance(cur, S, nbatch, stride);
    stage = stage + 1 == DG_RING_STAGES ? 0 : stage + 1;
  }
  cp_async_wait<0>();
  cg_init_fold(ci, pc, out, grid_reduce<2>(acc, smem, partials, ticket, out));
}

// ---------------------------------------------------------------- full-sky sufficient statistics
// One (plane, band chunk) combination per block as in mh_suffstat_uni_kernel (DG_SUFF_CHUNK == DG_UB bands): the
// thread's items are its pixel pairs, one batch each.
template <int NC>
__global__ void __launch_bounds__(DG_THREADS, 2)
mh_suffstat_ring_kernel(const ModelView mv, const MhView mh, const MhScalars *ms, double *partials,
                        unsigned int *tickets, double *out) {
  static_assert(DG_SUFF_CHUNK == DG_UB, "one ring stage holds one band chunk");
  constexpr int NV = 3 * DG_SUFF_CHUNK;
  extern __shared__ __align__(16) unsigned char ring_raw[];
  double2 *ring = reinterpret_cast<double2 *>(ring_raw);
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
  const int ncombo = mh.S * nchunk;
  const int combo = blockIdx.x % ncombo, sub = blockIdx.x / ncombo, nsub = gridDim.x / ncombo;
  const int64_t stride = (int64_t)nsub * blockDim.x;
  const int s = combo / nchunk, ch = combo % nchunk;
  const int k = mh.plane[s];
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) acc[i] = 0.0;
  auto issue = [&](int64_t e, int stage) {
    if (e < n2) {
      double2 *dst = ring + (size_t)stage * DG_RING_SLOTS * blockDim.x + threadIdx.x;
#pragma unroll
      for (int jj = 0; jj < DG_SUFF_CHUNK; jj++) {
        const int j = min(ch * DG_SUFF_CHUNK + jj, B - 1);
        const size_t off = plane_off(mv, j, k) + 2 * e;
        cp_async16(dst + (size_t)(2 * jj) * blockDim.x, mv.sig + off);
        cp_async16(dst + (size_t)(2 * jj + 1) * blockDim.x, mv.rms + off);
      }
    }
    cp_async_commit();
  };
  int64_t e = (int64_t)sub * blockDim.x + threadIdx.x, epre = e;
#pragma unroll
  for (int st = 0; st < DG_RING_STAGES - 1; st++) {
    issue(epre, st);
    epre += stride;
  }
  int stage = 0, pstage = DG_RING_STAGES - 1;
  // amplitudes and mask of the next TWO pixel pairs travel in registers (a two-deep pipeline, like the ring: one
  // iteration of this loop is shorter than the memory latency under load)
  double2 a_n1[NC], a_n2[NC];
  uchar2 mk_n1 = make_uchar2(0, 0), mk_n2 = make_uchar2(0, 0);
  auto fetch = [&](int64_t en, double2 (&an)[NC], uchar2 &mkn) {
    const int64_t pn = 2 * (en < n2 ? en : e);  // (a valid address when the thread has run out of items)
    mkn = *reinterpret_cast<const uchar2 *>(mv.mask + pn);
#pragma unroll
    for (int c = 0; c < NC; c++)
      an[c] = c < mv.ncomp ? ld2(mv.comp[c].amp + (size_t)k * mv.Ppad + pn) : make_double2(0.0, 0.0);
  };
  if (e < n2) {
    fetch(e, a_n1, mk_n1);
    fetch(e + stride, a_n2, mk_n2);
  }
  for (; e < n2; e += stride) {
    issue(epre, pstage);
    epre += stride;
    pstage = pstage + 1 == DG_RING_STAGES ? 0 : pstage + 1;
    const uchar2 mk = mk_n1;
    const bool use0 = mk.x != 0, use1 = mk.y != 0;
    double2 a[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      a[c] = a_n1[c];
      a_n1[c] = a_n2[c];
    }
    mk_n1 = mk_n2;
    fetch(e + 2 * stride, a_n2, mk_n2);
    double2 am = make_double2(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < NC; c++)
      if (c == mh.ic) am = a[c];
    cp_async_wait<DG_RING_STAGES - 1>();
    const double2 *src = ring + (size_t)stage * DG_RING_SLOTS * blockDim.x + threadIdx.x;
    stage = stage + 1 == DG_RING_STAGES ? 0 : stage + 1;
    if (!use0 && !use1) continue;
#pragma unroll
    for (int jj = 0; jj < DG_SUFF_CHUNK; jj++) {
      const int j = ch * DG_SUFF_CHUNK + jj;
      if (j < B) {
        double2 d = src[(size_t)(2 * jj) * blockDim.x];
        const double2 rm = src[(size_t)(2 * jj + 1) * blockDim.x];
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
        const double ix = fast_rcp(rm.x), iy = fast_rcp(rm.y);
        const double tx = (d.x - am.x * s0s[j]) * ix, ty = (d.y - am.y * s0s[j]) * iy;
        const double ux = am.x * ix, uy = am.y * iy;
        acc[3 * jj + 0] += (use0 ? tx * tx : 0.0) + (use1 ? ty * ty : 0.0);
        acc[3 * jj + 1] += (use0 ? tx * ux : 0.0) + (use1 ? ty * uy : 0.0);
        acc[3 * jj + 2] += (use0 ? ux * ux : 0.0) + (use1 ? uy * uy : 0.0);
      }
    }
  }
  cp_async_wait<0>();
  grid_reduce_grouped<NV>(acc, smem, partials, tickets, out, ncombo);
}
