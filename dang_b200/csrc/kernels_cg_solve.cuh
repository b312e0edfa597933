// kernels_cg_solve.cuh -- the whole of cg_search's loop (src/dang_cg_mod.f90:293-314) as ONE persistent,
// cooperatively launched kernel.  Every CG iteration is a sweep of the checkpointed-recompute form
// (kernels_cg.cuh, cg_recompute_pass_kernel): an element's state is re-derived from the last checkpoint by
// replaying the block-local recurrences with the (alpha_i, beta_i) history, and only the LAST replay step of pass k
// needs the two scalars that the global sums of pass k-1 produce.  So the passes are pipelined instead of separated
// by a grid barrier:
//   * a block that has finished its sweep of pass k-1 posts its partial sums and goes straight on to pass k: it
//     prefetches (cp.async ring) and replays the known steps c0+1 .. k-1 of its first elements;
//   * the block that arrives last finishes the deterministic grid reduction, exchanges the four sums with the other
//     ranks over the NVLink mailboxes (peer_exchange), advances the scalar state (alpha, beta, delta, the stop rule of
//     :293) and publishes pass k-1 as complete (CgScalars::gen, release);
//   * every warp picks up (alpha_k, beta_k, done) once, at the point where its first element needs them (acquire on
//     gen), and keeps them in registers for the rest of the sweep.
// The reduction tail, the exchange latency and the cold start of the next sweep overlap; nothing is stored before
// a warp knows that the solve is still running, so a finished solve is never touched again.  No launch, no host round
// trip, no scalar kernel between iterations; on a rank whose CG state fits the 126 MB L2 (nside 512 on >= 4 GPUs) the
// sweeps never touch HBM.
//
// unpack_amplitudes (:1284-1396) rides along speculatively: from the pass the previous solve of this
// (group, flag) ended on (`k_pred`) onwards a sweep also brings x up to date and writes the amplitude
// planes, so a solve that converges where the previous one did needs no extra sweep at the end.  x carries
// its own marker (x_at): the additions x += alpha_i d_i happen once each, in iteration order,
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
__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// dynamic shared memory of cg_solve_kernel<C>: [stage][slot][thread] 16-byte slots, slots = T (M) + C (r) + C (d)
template <int C>
constexpr size_t cg_ring_bytes() {
  return (size_t)DG_CGR_STAGES * (C * (C + 1) / 2 + 2 * C) * DG_THREADS * sizeof(double2);
}

// what a warp learns when pass k-1 is published
struct CgLate {
  double alpha, beta;  // of pass k
  int done;            // the solve ended with pass k-1
};

// One sweep over this rank's elements for pass k: replay steps c0+1 .. k of the block-local recurrences from the stored
// state (r_{c0+1}, d_{c0}); steps with index > x_at also advance x.  The scalars of steps c0+1 .. k-1 come from the
// history in shared memory; those of step k (and whether the solve is still running) are fetched by `late()` the
// first time they are needed.  Returns false when the solve turned out to be over (nothing was stored).
//   store_rd: write (r, d) back (checkpoint pass);  with_x: read x, write x and the amplitude planes.
// closing == true: every step is known (the solve is over, x is brought up to date, no sums).
template <int C, typename Late>
__device__ __forceinline__ bool cg_replay_sweep(const double *__restrict__ M, double *__restrict__ x,
                                                double *__restrict__ r, double *__restrict__ d, int64_t n2, int c0, int k,
                                                int x_at, bool store_rd, bool with_x, bool closing, const double *ha,
                                                const double *hb, const CgAmpOut<C> &ao, double2 *ring, Late late,
                                                CgLate &lt, bool &have_late, double (&acc)[4]) {
  constexpr int T = C * (C + 1) / 2;
  constexpr int SLOTS = T + 2 * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t vs = (size_t)n2 * 2;
  const int nknown = closing ? k - c0 : k - 1 - c0;  // replay steps with scalars from the history
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
  bool alive = true;
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
    for (int i = 0; i < nknown; i++) {
      const double alpha = ha[c0 + 1 + i], beta = hb[c0 + 1 + i];
      cg_block_step<C>(m, dv, rv, q, alpha, beta);                        // :305, :296, :300
      if (with_x && c0 + 1 + i > x_at) cg_block_x<C>(xv, dv, alpha);       // :298
    }
    if (!closing) {
      if (!have_late) {  // (warp-uniform: every lane of a warp reaches its first element together or not at all)
        lt = late();
        have_late = true;
      }
      if (lt.done) {
        alive = false;
        break;
      }
      cg_block_step<C>(m, dv, rv, q, lt.alpha, lt.beta);
      if (with_x) cg_block_x<C>(xv, dv, lt.alpha);  // (k > x_at always)
      cg_block_sums<C>(m, dv, rv, q, acc);
    }
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
  return alive;
}

template <int C>
__global__ void __launch_bounds__(DG_THREADS, DG_CG_BLOCKS_PER_SM)
cg_solve_kernel(CgScalars *st, const double *__restrict__ M, double *__restrict__ x, double *__restrict__ r,
                double *__restrict__ d, int64_t n2, double *partials, unsigned int *ticket, double *out,
                PeerComm pc, double *gathered, CgAmpOut<C> ao, int k_pred) {
  extern __shared__ __align__(16) unsigned char cg_ring_raw[];
  double2 *ring = reinterpret_cast<double2 *>(cg_ring_raw);
  __shared__ double smem[4 * 32];
  __shared__ double ha[DG_CG_HIST], hb[DG_CG_HIST];  // alpha_i, beta_i used IN pass i (1-based), as they become known
  __shared__ int s_done0, s_m, s_done, s_gen, s_kpred;  // s_gen: last pass whose outcome warp 0 has copied into this block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {  // state left by cg_init_update (K1's last block): pass 1 is fully known
    s_done0 = __ldcg(&st->done);
    s_done = 0;
    s_gen = 0;
    s_m = __ldcg(&st->m);
    s_kpred = k_pred >= 0 ? k_pred : __ldcg(&st->k_pred);  // (< 0: the host has not seen the previous solve's count)
    ha[1] = __ldcg(&st->ah[1]);
    hb[1] = __ldcg(&st->bh[1]);
  }
  __syncthreads();
  const int m = s_m;
  k_pred = s_kpred;
  int k = 1, c0 = 0, x_at = 0;  // pass to run, pass of the last checkpoint, pass x is current for: block-uniform
  bool done = s_done0 != 0;
  while (!done) {
    const int nstep = k - c0;
    const bool store = nstep == m;  // checkpoint pass
    const bool with_x = store || k >= k_pred;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    CgLate lt;
    lt.alpha = ha[k];  // valid for k == 1 only; otherwise fetched by late()
    lt.beta = hb[k];
    lt.done = 0;
    bool have_late = k == 1;
    // Pass k-1 is published once gen >= k-1.  ONE warp per block watches the global counter (a few hundred pollers on
    // that L2 line, not a few thousand) and copies the outcome into shared memory; the other warps watch the copy.
    auto late = [&]() -> CgLate {
      CgLate v;
      v.alpha = 0.0;
      v.beta = 0.0;
      v.done = 0;
      if (lane == 0) {
        if (warp == 0) {
          while (ld_acquire_gpu(&st->gen) < (unsigned)(k - 1)) __nanosleep(40);
          const int kk = k < DG_CG_HIST ? k : DG_CG_HIST - 1;
          *(volatile double *)&ha[kk] = __ldcg(&st->ah[kk]);
          *(volatile double *)&hb[kk] = __ldcg(&st->bh[kk]);
          *(volatile int *)&s_done = __ldcg(&st->done);
          __threadfence_block();
          *(volatile int *)&s_gen = k - 1;
        } else {
          while (*(volatile int *)&s_gen < k - 1) __nanosleep(40);
          __threadfence_block();
        }
        const int kk = k < DG_CG_HIST ? k : DG_CG_HIST - 1;
        v.alpha = *(volatile double *)&ha[kk];
        v.beta = *(volatile double *)&hb[kk];
        v.done = *(volatile int *)&s_done;
      }
      v.alpha = __shfl_sync(0xffffffffu, v.alpha, 0);
      v.beta = __shfl_sync(0xffffffffu, v.beta, 0);
      v.done = __shfl_sync(0xffffffffu, v.done, 0);
      return v;
    };
    cg_replay_sweep<C>(M, x, r, d, n2, c0, k, x_at, store, with_x, false, ha, hb, ao, ring, late, lt, have_late, acc);
    if (!have_late) lt = late();  // warps without an element in this sweep still have to learn how pass k-1 ended
    if (lt.done) {  // the solve ended with pass k-1 (every warp of every block reads the same published flag)
      done = true;
      k = k - 1;
      break;
    }
    const bool last = grid_reduce<4>(acc, smem, partials, ticket, out);  // (its barriers order ha[k] / hb[k] for later passes)
    if (last) {  // warp 0 of the block that arrived last: exchange over NVLink (if any), advance the scalars
      if (pc.nranks > 1) peer_exchange(pc, out, 4, gathered);
      if (threadIdx.x == 0) {
        cg_fused_update(st, pc.nranks > 1 ? gathered : out, pc.nranks);  // iter = k + 1, done, ah / bh[k + 1], ckpt
        if (with_x) st->x_at = k;
        if (st->done) st->k_pred = k;  // the solve ends with this pass: the next solve's prediction
        __threadfence();
        st_release_gpu(&st->gen, (unsigned)k);  // pass k is complete: its scalars are visible before the count
      }
    }
    if (store) c0 = k;
    if (with_x) x_at = k;
    k++;
  }
  if (s_done0) return;  // nothing to solve
  // k passes ran.  Closing sweep, only when x is behind the last pass (the solve ended before the predicted pass).
  __syncthreads();      // (ha / hb of the last pass)
  if (k > x_at) {
    double acc[4];
    CgLate lt;
    lt.alpha = lt.beta = 0.0;
    lt.done = 0;
    bool have_late = true;
    auto none = [&]() -> CgLate { return lt; };
    cg_replay_sweep<C>(M, x, r, d, n2, c0, k, x_at, false, true, true, ha, hb, ao, ring, none, lt, have_late, acc);
  }
}
