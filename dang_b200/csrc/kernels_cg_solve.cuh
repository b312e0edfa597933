// kernels_cg_solve.cuh -- the whole of cg_search's loop (src/dang_cg_mod.f90:293-314) as ONE persistent,
// cooperatively launched kernel: every CG iteration is a sweep of the checkpointed-recompute form
// (kernels_cg.cuh, cg_recompute_pass_kernel) followed by a grid barrier; the block that finishes the
// deterministic grid reduction exchanges the four sums with the other ranks over the NVLink mailboxes
// (peer_exchange), advances the scalar state (alpha, beta, delta, the stop rule of :293) and releases the
// barrier.  No launch, no host round trip and no scalar kernel between iterations; on a rank whose CG state
// fits the 126 MB L2 (nside 512 on >= 4 GPUs) the sweeps never touch HBM.
//
// The sweeps are asynchronous streams: every thread copies the block matrices and the stored (r, d) of its next
// two element pairs into a private three-stage ring in shared memory with cp.async while it replays the
// recurrences of the current pair, so the FP64 work of the replay (up to m block steps per element) overlaps the
// memory stream instead of alternating with it.
//
// unpack_amplitudes (:1284-1396) rides along speculatively: from the pass the previous solve of this
// (group, flag) ended on (`k_pred`) onwards a sweep also brings x up to date and writes the amplitude
// planes, so a solve that converges where the previous one did needs no extra sweep at the end.  x carries
// its own marker (CgScalars::x_at): the additions x += alpha_i d_i happen once each, in iteration order,
// exactly as :298 performs them -- the result is bit-identical to the pass-per-launch forms.
#pragma once
#include "cp_async.cuh"
#include "kernels_cg.cuh"

#define DG_CGR_STAGES 3

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int *p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// dynamic shared memory of cg_solve_kernel<C>: [stage][slot][thread] 16-byte slots, slots = T (M) + C (r) + C (d)
template <int C>
constexpr size_t cg_ring_bytes() {
  return (size_t)DG_CGR_STAGES * (C * (C + 1) / 2 + 2 * C) * DG_THREADS * sizeof(double2);
}

// One sweep over this rank's elements: replay steps c0+1 .. c0+nstep of the block-local recurrences from the
// stored state (r_{c0+1}, d_{c0}); steps with index > x_at also advance x.  REDUCE: accumulate the four sums
// of the last step (a CG pass); otherwise it is the closing sweep of a solve.
//   store_rd: write (r, d) back (checkpoint pass);  with_x: read x, write x and the amplitude planes.
template <int C, bool REDUCE>
__device__ __forceinline__ void cg_replay_sweep(const double *__restrict__ M, double *__restrict__ x,
                                                double *__restrict__ r, double *__restrict__ d, int64_t n2,
                                                int c0, int nstep, int x_at, bool store_rd, bool with_x,
                                                const double *sa, const double *sb, const CgAmpOut<C> &ao,
                                                double2 *ring, double (&acc)[4]) {
  constexpr int T = C * (C + 1) / 2;
  constexpr int SLOTS = T + 2 * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)n2 * 2;
  const int xskip = x_at - c0;  // replay steps i < xskip are already in x
  const bool read_d = c0 != 0;  // before the first checkpoint the stored direction is d_0 = r_1 itself (beta_1 = 0)
  auto issue = [&](int64_t e, int stage) {
    if (e < n2) {
      double2 *dst = ring + (size_t)stage * SLOTS * blockDim.x + threadIdx.x;
#pragma unroll
      for (int t = 0; t < T; t++) cp_async16(dst + (size_t)t * blockDim.x, M + t * vs + 2 * e);
#pragma unroll
      for (int c = 0; c < C; c++) {
        cp_async16(dst + (size_t)(T + c) * blockDim.x, r + c * vs + 2 * e);
        if (read_d) cp_async16(dst + (size_t)(T + C + c) * blockDim.x, d + c * vs + 2 * e);
      }
    }
    cp_async_commit();
  };
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, epre = e;
#pragma unroll
  for (int st = 0; st < DG_CGR_STAGES - 1; st++) {
    issue(epre, st);
    epre += stride;
  }
  int stage = 0, pstage = DG_CGR_STAGES - 1;
  for (; e < n2; e += stride) {
    issue(epre, pstage);
    epre += stride;
    pstage = pstage + 1 == DG_CGR_STAGES ? 0 : pstage + 1;
    double2 m[T], xv[C], rv[C], dv[C], q[C];
    if (with_x) {
#pragma unroll
      for (int c = 0; c < C; c++) xv[c] = __ldcg(reinterpret_cast<const double2 *>(x + c * vs + 2 * e));
    }
    cp_async_wait<DG_CGR_STAGES - 1>();
    const double2 *src = ring + (size_t)stage * SLOTS * blockDim.x + threadIdx.x;
    stage = stage + 1 == DG_CGR_STAGES ? 0 : stage + 1;
#pragma unroll
    for (int t = 0; t < T; t++) m[t] = src[(size_t)t * blockDim.x];
#pragma unroll
    for (int c = 0; c < C; c++) {
      rv[c] = src[(size_t)(T + c) * blockDim.x];
      dv[c] = read_d ? src[(size_t)(T + C + c) * blockDim.x] : rv[c];
    }
    for (int i = 0; i < nstep; i++) {
      const double alpha = sa[i], beta = sb[i];
      cg_block_step<C>(m, dv, rv, q, alpha, beta);                // :305, :296, :300
      if (with_x && i >= xskip) cg_block_x<C>(xv, dv, alpha);     // :298
    }
    if (REDUCE) cg_block_sums<C>(m, dv, rv, q, acc);
#pragma unroll
    for (int c = 0; c < C; c++) {
      if (store_rd) {
        *reinterpret_cast<double2 *>(r + c * vs + 2 * e) = rv[c];
        *reinterpret_cast<double2 *>(d + c * vs + 2 * e) = dv[c];
      }
      if (with_x) {
        *reinterpret_cast<double2 *>(x + c * vs + 2 * e) = xv[c];
        *reinterpret_cast<double2 *>(ao.p[c] + 2 * e) = xv[c];  // unpack_amplitudes :1327-1335
      }
    }
  }
  cp_async_wait<0>();
}

template <int C>
__global__ void __launch_bounds__(DG_THREADS, DG_CG_BLOCKS_PER_SM)
cg_solve_kernel(CgScalars *st, const double *__restrict__ M, double *__restrict__ x, double *__restrict__ r,
                double *__restrict__ d, int64_t n2, double *partials, unsigned int *ticket, double *out,
                PeerComm pc, double *gathered, CgAmpOut<C> ao, int k_pred) {
  extern __shared__ __align__(16) unsigned char cg_ring_raw[];
  double2 *ring = reinterpret_cast<double2 *>(cg_ring_raw);
  __shared__ double smem[4 * 32];
  __shared__ double sa[DG_CG_MAXM + 1], sb[DG_CG_MAXM + 1];
  __shared__ int ctl[6];
  unsigned int gen = 0;  // st->gen was zeroed by cg_init_update
  for (;;) {
    if (threadIdx.x == 0) {  // the scalar state lives in L2: read around L1 (it changes between passes)
      ctl[0] = __ldcg(&st->done);
      ctl[1] = __ldcg(&st->iter);
      ctl[2] = __ldcg(&st->ckpt);
      ctl[3] = __ldcg(&st->x_at);
      ctl[4] = __ldcg(&st->m);
    }
    __syncthreads();
    const int done = ctl[0], k = ctl[1], c0 = ctl[2], x_at = ctl[3], m = ctl[4];
    if (done) {
      // closing sweep, only when x is behind the last pass (the solve ended before the predicted pass)
      const int nstep = (k - 1) - c0;
      if ((k - 1) > x_at) {
        if (threadIdx.x < nstep) {
          sa[threadIdx.x] = __ldcg(&st->ah[c0 + 1 + threadIdx.x]);
          sb[threadIdx.x] = __ldcg(&st->bh[c0 + 1 + threadIdx.x]);
        }
        __syncthreads();
        double acc[4];
        cg_replay_sweep<C, false>(M, x, r, d, n2, c0, nstep, x_at, false, true, sa, sb, ao, ring, acc);
      }
      return;
    }
    const int nstep = k - c0;
    const bool store = nstep == m;  // checkpoint pass
    const bool with_x = store || k >= k_pred;
    if (threadIdx.x < nstep) {
      sa[threadIdx.x] = __ldcg(&st->ah[c0 + 1 + threadIdx.x]);
      sb[threadIdx.x] = __ldcg(&st->bh[c0 + 1 + threadIdx.x]);
    }
    __syncthreads();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    cg_replay_sweep<C, true>(M, x, r, d, n2, c0, nstep, x_at, store, with_x, sa, sb, ao, ring, acc);
    const bool last = grid_reduce<4>(acc, smem, partials, ticket, out);
    if (last) {  // warp 0 of the block that arrived last: exchange over NVLink (if any), advance the scalars
      if (pc.nranks > 1) peer_exchange(pc, out, 4, gathered);
      if (threadIdx.x == 0) {
        cg_fused_update(st, pc.nranks > 1 ? gathered : out, pc.nranks);
        if (with_x) st->x_at = k;
        __threadfence();
        red_release_gpu_add(&st->gen, 1u);  // releases the barrier: scalars are visible before the count
      }
    }
    gen++;
    if (threadIdx.x == 0)
      while (ld_acquire_gpu(&st->gen) < gen) __nanosleep(40);
    __syncthreads();
  }
}
