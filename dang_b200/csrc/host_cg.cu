// host_cg.cu -- amplitude draw: compute_rhs + cg_search + unpack_amplitudes for one (group, flag),
// src/dang_cg_mod.f90:167-169 (launch logic; kernels in kernels_cg.cuh / kernels_uni.cuh).
#include <utility>

#include "host.cuh"
#include "kernels_cg.cuh"
#include "kernels_cg_solve.cuh"
#include "kernels_uni.cuh"
#include "kernels_stream.cuh"

namespace {
// ---------------------------------------------------------------- amplitude draw
template <int C>
void cg_solve_impl(dang_gpu *h, CgGroupHost &g, int flag_n, int ml_mode, const double *eta,
                   uint64_t seed, const int *comps, const int *og, int nog, int *n_iter,
                   double *delta_final) {
  constexpr int T = C * (C + 1) / 2;
  ModelView mv = model_view(h);
  CgView<C> cv;
  memset(&cv, 0, sizeof cv);
  cv.S = flag_planes(g.pol_flag[flag_n], cv.plane);
  for (int c = 0; c < C; c++) cv.comp[c] = comps[c];
  cv.nog = nog;
  for (int o = 0; o < nog; o++) cv.og[o] = og[o];
  const int S = cv.S;
  for (int s = 0; s < S; s++)
    if (cv.plane[s] >= h->nmaps) fail(DANG_GPU_EINVAL, "pol flag needs plane %d, nmaps = %d", cv.plane[s] + 1, h->nmaps);
  const size_t vs = (size_t)S * h->Ppad;  // doubles per component
  const int64_t n2 = (int64_t)(vs / 2);

  // self%x: allocate + seed from c%amplitude on first use only (cg_search :227-239, Q10)
  if (!g.x[flag_n] || g.x_len[flag_n] != C * vs) {
    if (g.x[flag_n]) CK(cudaFree(g.x[flag_n]));
    CK(cudaMalloc(&g.x[flag_n], C * vs * sizeof(double)));
    g.x_len[flag_n] = C * vs;
    CK(cudaMemsetAsync(g.x[flag_n], 0, C * vs * sizeof(double), h->stream));
    for (int c = 0; c < C; c++)
      for (int s = 0; s < S; s++)  // initialize_x :1216-1224
        CK(cudaMemcpyAsync(g.x[flag_n] + c * vs + (size_t)s * h->Ppad,
                           h->comp[comps[c]].amp + (size_t)cv.plane[s] * h->Ppad,
                           h->P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (h->M_len < T * vs) {
    ensure(h->M, h->M_len, T * vs);
    h->cg_layout = -1;
  }
  if (h->v_len < C * vs) {
    size_t l1 = h->v_len, l2 = h->v_len;
    ensure(h->r, l1, C * vs);
    ensure(h->d, l2, C * vs);
    h->v_len = C * vs;
    h->cg_layout = -1;
  }
  // padding lanes must hold zeros (they are swept by the vectorised passes and nothing ever
  // writes a nonzero there), so clear only when the plane layout of the scratch changes
  if (h->cg_layout != C * 16 + S) {
    CK(cudaMemsetAsync(h->M, 0, T * vs * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->r, 0, C * vs * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->d, 0, C * vs * sizeof(double), h->stream));
    h->cg_layout = C * 16 + S;
  }
  cv.M = h->M;
  cv.r = h->r;
  cv.d = h->d;
  cv.x = g.x[flag_n];
  cv.seed = seed;
  cv.fluct = 0;
  cv.eta = nullptr;
  int staged_slot = -1;
  // recompute form needs the (alpha, beta) history: fall back to streaming for very long solves
  const int ckpt_m = (!h->cg_two_pass && g.i_max < DG_CG_HIST) ? h->cg_ckpt : 0;
  cv.store_d = ckpt_m ? 0 : 1;
  if (ml_mode == DANG_ML_SAMPLE) {  // :254-264
    cv.fluct = h->fix_q1 ? 2 : 1;
    if (eta) {
      ensure(h->eta, h->eta_len, vs);
      h2d_planes(h, h->eta, eta, S);  // host eta is [stokes][npix]
      cv.eta = h->eta;
    } else if (h->eta_count > 0 && h->eta_stage_planes[h->eta_head] == S) {
      CK(cudaStreamWaitEvent(h->stream, h->ev_eta[h->eta_head], 0));  // uploaded by dang_gpu_stage_eta
      cv.eta = h->eta_stage[h->eta_head];
      staged_slot = h->eta_head;
      h->eta_head = (h->eta_head + 1) % 2;
      h->eta_count--;
    }
  }

  CK(cudaMemsetAsync(h->sums_local, 0, GATHER_MAX * sizeof(double), h->stream));
  // one rank, or NVLink mailboxes: the last block of K1 sees every rank's sums and starts the scalar state itself
  const bool fold_all = h->nranks == 1 || h->use_mail;
  const CgInit ci{fold_all ? h->cg_scalars : nullptr, g.i_max, ckpt_m, g.converge, h->gathered};
  {
    const double n_el = (double)S * h->P;
    KTimer kt(h, DANG_K_RHS_BLOCKS,
              bytes_w(n_el * (2.0 * h->nbands + 1 + C + T + (ckpt_m ? 1.0 : 2.0) * C)) + bytes_w((double)h->P * 3));
    // streaming kernel: out-of-group components must have tabulated SEDs; group components with
    // varying indices get their SEDs staged per thread in dynamic shared memory
    bool og_uni = true;
    unsigned nu_mask = 0;
    for (int s = 0; s < S; s++) {
      for (int c = 0; c < C; c++)
        if (!comp_uniform(h, comps[c], cv.plane[s])) nu_mask |= 1u << c;
      for (int o = 0; o < nog; o++) og_uni = og_uni && comp_uniform(h, og[o], cv.plane[s]);
    }
    const size_t dsm = (size_t)__builtin_popcount(nu_mask) * h->nbands * 2 * DG_THREADS * sizeof(double);
    const size_t tma_smem = (size_t)DG_TMA_STAGES * DG_TMA_BANDS * 2 * 2 * DG_TMA_TILE * sizeof(double);
    const size_t ring_smem = (size_t)DG_RING_STAGES * DG_RING_SLOTS * DG_THREADS * sizeof(double2);
    if (h->stream_ring && !h->use_tma && nu_mask == 0 && nog == 0) {
      // asynchronous stream: per-thread cp.async ring in shared memory (kernels_stream.cuh)
      CK(cudaFuncSetAttribute(rhs_blocks_ring_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
      const int g4 = occ_grid(h, rhs_blocks_ring_kernel<C>, h->Ppad / 2, DG_THREADS, ring_smem);
      rhs_blocks_ring_kernel<C><<<g4, DG_THREADS, ring_smem, h->stream>>>(mv, cv, h->partials, h->tickets, h->sums_local, ci, h->peer);
    } else if (h->use_tma && nu_mask == 0 && nog == 0) {
      // TMA-staged stream: one block per SM, 3-stage shared-memory ring filled by cp.async.bulk
      CK(cudaFuncSetAttribute(rhs_blocks_tma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem));
      const int64_t ntiles = (h->Ppad + DG_TMA_TILE - 1) / DG_TMA_TILE;
      const int g3 = (int)(ntiles < h->num_sms ? ntiles : h->num_sms);
      rhs_blocks_tma_kernel<C><<<g3, DG_TMA_THREADS, tma_smem, h->stream>>>(mv, cv, h->partials, h->tickets, h->sums_local, ci, h->peer);
    } else if (og_uni && dsm <= 160 * 1024) {
      if (dsm > 48 * 1024) CK(cudaFuncSetAttribute(rhs_blocks_uni_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
      const int g2 = occ_grid(h, rhs_blocks_uni_kernel<C>, h->Ppad / 2, DG_THREADS, dsm);
      rhs_blocks_uni_kernel<C><<<g2, DG_THREADS, dsm, h->stream>>>(mv, cv, h->partials, h->tickets, h->sums_local, nu_mask, ci, h->peer);
    } else {
      const int g1 = occ_grid(h, rhs_blocks_kernel<C>, h->P, DG_THREADS);
      rhs_blocks_kernel<C><<<g1, DG_THREADS, 0, h->stream>>>(mv, cv, h->partials, h->tickets, h->sums_local, ci, h->peer);
    }
    kt.done();
    if (staged_slot >= 0) {  // the slot may be refilled once K1 has read it
      CK(cudaEventRecord(h->ev_eta_used[staged_slot], h->stream));
      h->eta_used_recorded[staged_slot] = true;
    }
  }
  if (!fold_all) {  // NCCL transport: gather on the stream, then a 1-thread kernel
    gather(h, 4);
    KTimer kt(h, DANG_K_SCALAR, 0);
    cg_init_scalars_kernel<<<1, 1, 0, h->stream>>>(h->cg_scalars, h->gathered, h->nranks, g.i_max, g.converge, ckpt_m);
    kt.done();
  }

  struct Snap { double delta_new; int iter, done; };
  bool have_state = false;
  auto read_state = [&]() -> Snap {  // scalars + residual trace in one small copy
    CgScalars *hs = (CgScalars *)h->pinned;
    readback(h, hs, h->cg_scalars, offsetof(CgScalars, ah));
    CK(cudaStreamSynchronize(h->stream));
    have_state = true;
    return Snap{hs->delta_new, hs->iter, hs->done};
  };

  // unpack_amplitudes :1327-1335: x -> c%amplitude planes.  If an asynchronous download is still reading a
  // component's planes the new state goes into its second buffer (solved planes from x, the others carried
  // over) and the buffers swap roles; the download is never waited for.
  for (int c = 0; c < C; c++) {
    CompHost &cc = h->comp[comps[c]];
    if (cc.read_pending) {
      const size_t n2a = (size_t)h->nmaps * h->Ppad;
      if (!cc.amp_alt) {
        CK(cudaMalloc(&cc.amp_alt, n2a * sizeof(double)));
        CK(cudaMemsetAsync(cc.amp_alt, 0, n2a * sizeof(double), h->stream));
      }
      if (cc.read_pending_alt) {  // the download before the pending one read the buffer we are about to write
        CK(cudaStreamWaitEvent(h->stream, cc.ev_read_alt, 0));
        cc.read_pending_alt = false;
      }
      for (int k = 0; k < h->nmaps; k++) {
        bool solved = false;
        for (int s = 0; s < S; s++) solved = solved || cv.plane[s] == k;
        if (!solved)
          CK(cudaMemcpyAsync(cc.amp_alt + (size_t)k * h->Ppad, cc.amp + (size_t)k * h->Ppad, h->P * sizeof(double),
                             cudaMemcpyDeviceToDevice, h->stream));
      }
      std::swap(cc.amp, cc.amp_alt);
      std::swap(cc.ev_read, cc.ev_read_alt);
      std::swap(cc.read_pending, cc.read_pending_alt);
    }
  }
  // ---- default form: the whole loop of cg_search in one persistent cooperative kernel (kernels_cg_solve.cuh)
  const bool persistent = h->cg_persistent && ckpt_m > 0 && fold_all && C <= 2;  // (C > 2: the prefetch ring of two resident blocks does not fit in shared memory)
  if (persistent) {
    CgAmpOut<C> ao{};
    for (int c = 0; c < C; c++) ao.p[c] = h->comp[comps[c]].amp + (size_t)cv.plane[0] * h->Ppad;  // S contiguous planes
    int nsolve = 0;
    for (auto &gg : h->cg)
      if (gg.set) nsolve += gg.nflag;
    // deferred scalars (host.cuh): with one solve per Gibbs iteration the device carries the previous count itself
    const bool defer = h->defer_scalars && nsolve == 1;
    int k_pred = defer ? -1 : g.last_iter[flag_n] > 1 ? g.last_iter[flag_n] - 1 : 0x7fffffff;  // pass the previous solve ended on
    int per_sm = 0;
    const size_t cg_smem = cg_ring_bytes<C>();
    CK(cudaFuncSetAttribute(cg_solve_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cg_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_solve_kernel<C>, DG_THREADS, cg_smem));
    if (per_sm < DG_CG_BLOCKS_PER_SM) fail(DANG_GPU_ECUDA, "cg_solve_kernel: %d resident blocks per SM, %d needed", per_sm, DG_CG_BLOCKS_PER_SM);
    const int64_t cgrid = grid_for(h, n2, DG_THREADS, DG_CG_BLOCKS_PER_SM);  // every block resident: the kernel has a grid barrier
    CgScalars *st = h->cg_scalars;
    const double *Mp = h->M;
    double *xp = g.x[flag_n], *rp = h->r, *dp = h->d, *part = h->partials, *outp = h->sums_local, *gath = h->gathered;
    unsigned int *tick = h->tickets;
    int64_t n2v = n2;
    PeerComm pcv = h->peer;
    void *args[] = {&st, &Mp, &xp, &rp, &dp, &n2v, &part, &tick, &outp, &pcv, &gath, &ao, &k_pred};
    // deferred amplitude downloads (host.cuh): they start when this solve starts -- the event fires once K1 is done --
    // and read the buffers swapped out above while the solve writes the fresh ones
    if (!h->deferred.empty()) CK(cudaEventRecord(h->ev_compute, h->stream));
    KTimer kt(h, DANG_K_CG_PASS, 0, true);
    CK(cudaLaunchCooperativeKernel((void *)cg_solve_kernel<C>, dim3((unsigned)cgrid), dim3(DG_THREADS), args, cg_smem, h->stream));
    kt.done();
    if (!h->deferred.empty()) issue_deferred_d2h(h, h->ev_compute);
    // The scalars come back right behind the solve and the host waits for that point of the stream only.  When this
    // is the run's only solve per Gibbs iteration and a full-sky draw follows, the draw's statistics pass (which
    // also serves the chi-square printed after the amplitude draw) is enqueued first, so the device keeps
    // streaming while the host turns the solve around.
    if (defer) {  // no host wait here: dang_gpu_iteration_mark / _scalars pick the outcome up later
      prefetch_statistics(h);
      kt.bytes = 0.0;  // booked by dang_gpu_iteration_scalars once the pass count is known
      kt.commit(true);
      h->pend_cg = true;
      h->pend_g = (int)(&g - h->cg.data());
      h->pend_flag = flag_n;
      h->pend_cg_T = T;
      h->pend_cg_C = C;
      h->pend_cg_m = ckpt_m;
      h->pend_cg_vs = (double)vs;
      h->last_trace.clear();
      if (n_iter) *n_iter = -1;
      if (delta_final) *delta_final = nan("");
      return;
    }
    CgScalars *hs = (CgScalars *)h->pinned;
    readback(h, hs, h->cg_scalars, offsetof(CgScalars, ah));
    CK(cudaEventRecord(h->ev_sync, h->stream));
    if (nsolve == 1) prefetch_statistics(h);
    CK(cudaEventSynchronize(h->ev_sync));
    kt.bytes = bytes_w(cg_solve_bytes(hs->iter - 1, k_pred, ckpt_m, T, C, (double)vs));
    kt.commit(true);
    int n = hs->iter < 256 ? hs->iter : 256;
    h->last_trace.assign(hs->trace, hs->trace + n);
    g.last_iter[flag_n] = hs->iter;
    if (n_iter) *n_iter = hs->iter;
    if (delta_final) *delta_final = hs->delta_new;
    return;
  }

  if (!h->deferred.empty()) {  // pass-per-launch forms: no long launch to hide a bulk copy behind, start it now
    CK(cudaEventRecord(h->ev_compute, h->stream));
    issue_deferred_d2h(h, h->ev_compute);
  }
  const int fold = (h->nranks == 1 || h->use_mail) ? 1 : 0;
  const int grid = grid_for(h, n2, DG_THREADS, DG_CG_BLOCKS_PER_SM);  // the same grid for every form (bit-equal sums)
  const double el = (double)vs;
  std::vector<std::pair<int, KTimer>> pass_recs;  // (pass number, launch) booked once the iteration count is known
  auto enqueue_pass = [&](int pass_no) {
    if (!h->cg_two_pass) {
      if (ckpt_m) {
        // compulsory traffic of this launch: M, r, d in; on checkpoint passes also x in, r, d, x out
        const double per_el = (pass_no % ckpt_m == 0) ? (T + (pass_no == ckpt_m ? 5.0 : 6.0) * C)
                                                      : (T + (pass_no < ckpt_m ? 1.0 : 2.0) * C);
        KTimer kt(h, DANG_K_CG_PASS, bytes_w(el * per_el), true);
        cg_recompute_pass_kernel<C, false><<<grid, DG_THREADS, 0, h->stream>>>(
            h->cg_scalars, h->M, g.x[flag_n], h->r, h->d, n2, h->partials, h->tickets, h->sums_local, fold, 0, h->peer, h->gathered,
            CgAmpOut<C>{});
        kt.done();
        pass_recs.emplace_back(pass_no, kt);
      } else {
        // compulsory traffic of this launch: x is touched on even passes only
        const double per_el = (pass_no & 1) ? (T + 4.0 * C) : (T + 6.0 * C);
        KTimer kt(h, DANG_K_CG_PASS, bytes_w(el * per_el), true);
        cg_fused_pass_kernel<C><<<grid, DG_THREADS, 0, h->stream>>>(
            h->cg_scalars, h->M, g.x[flag_n], h->r, h->d, n2, h->partials, h->tickets, h->sums_local, fold, h->peer, h->gathered);
        kt.done();
        pass_recs.emplace_back(pass_no, kt);
      }
      if (!fold) {
        gather(h, 4);
        KTimer ks(h, DANG_K_SCALAR, 0);
        cg_fused_scalars_kernel<<<1, 1, 0, h->stream>>>(h->cg_scalars, h->gathered, h->nranks);
        ks.done();
      }
    } else {
      {
        KTimer kt(h, DANG_K_CG_DQ, bytes_w(el * (T + 3.0 * C)));
        cg_dq_pass_kernel<C><<<grid, DG_THREADS, 0, h->stream>>>(h->cg_scalars, h->M, h->r, h->d, n2,
                                                                  h->partials, h->tickets, h->sums_local);
        kt.done();
      }
      gather(h, 4);
      {
        KTimer ks(h, DANG_K_SCALAR, 0);
        cg_dq_scalars_kernel<<<1, 1, 0, h->stream>>>(h->cg_scalars, h->gathered, h->nranks);
        ks.done();
      }
      {
        KTimer kt(h, DANG_K_CG_UPDATE, bytes_w(el * (T + 5.0 * C)));
        cg_update_pass_kernel<C><<<grid, DG_THREADS, 0, h->stream>>>(
            h->cg_scalars, h->M, g.x[flag_n], h->r, h->d, n2, h->partials, h->tickets, h->sums_local);
        kt.done();
      }
      gather(h, 4);
      KTimer ks(h, DANG_K_SCALAR, 0);
      cg_rr_scalars_kernel<<<1, 1, 0, h->stream>>>(h->cg_scalars, h->gathered, h->nranks);
      ks.done();
    }
  };
  // Passes are enqueued without waiting for the convergence flag: a pass launched after the
  // solve is done returns at once (device-side early exit).  The first batch is sized from the
  // previous solve of this (group, flag) -- successive Gibbs iterations converge in almost the
  // same number of steps -- and further batches of cg_chunk follow until the flag is seen.
  // L2 residency for the block matrices: every pass re-reads M (and r); the working set (250 MB at
  // nside 512) exceeds the 126 MB L2 and a cyclic stream gets no hits from LRU, so a slice of M is
  // marked persisting for the passes of this solve (DANG_OPT_L2_PERSIST_MB) and the rest streams.
  bool l2_window = false;
  if (h->l2_persist_mb > 0 && h->l2_persist_max > 0) {
    size_t want = (size_t)h->l2_persist_mb << 20;
    if (want > h->l2_persist_max) want = h->l2_persist_max;
    size_t win = (size_t)T * vs * sizeof(double);
    if (win > h->l2_window_max) win = h->l2_window_max;
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    av.accessPolicyWindow.base_ptr = h->M;
    av.accessPolicyWindow.num_bytes = win;
    av.accessPolicyWindow.hitRatio = want >= win ? 1.0f : (float)((double)want / (double)win);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    l2_window = true;
  }
  const int max_pass = g.i_max - 1;
  int enq = 0;
  int batch = g.last_iter[flag_n] > 1 ? g.last_iter[flag_n] - 1 : h->cg_chunk;
  Snap sn{0.0, 1, max_pass < 1};
  while (!sn.done && enq < max_pass) {
    if (batch > max_pass - enq) batch = max_pass - enq;
    for (int it = 0; it < batch; it++) enqueue_pass(enq + it + 1);
    enq += batch;
    sn = read_state();
    batch = h->cg_chunk;
  }
  // The last pass of the recompute form brings x up to date and writes the amplitude planes in the same
  // sweep -- when it has anything to do (a solve that stopped exactly on a checkpoint needs no pass).
  bool unpacked = false;
  if (!h->cg_two_pass) {  // bring x up to date (pending term of the deferred / checkpointed update)
    const CgScalars *hs0 = (const CgScalars *)h->pinned;
    // the final pass works only if passes ran since the last checkpoint (recompute form) / an odd number of
    // passes ran (streaming form); when the host does not know (no state read yet) it is booked as working
    const bool final_works = !(sn.done && have_state) || (ckpt_m ? (hs0->iter - 1) - hs0->ckpt > 0 : ((hs0->iter - 1) & 1) != 0);
    KTimer kt(h, DANG_K_CG_FIXUP, bytes_w(el * (ckpt_m ? T + 5.0 * C : 3.0 * C)), true);
    if (ckpt_m) {
      CgAmpOut<C> ao{};
      const bool planes_contiguous = S == 1 || cv.plane[1] == cv.plane[0] + 1;
      if (sn.done && have_state && planes_contiguous && (hs0->iter - 1) - hs0->ckpt > 0) {
        for (int c = 0; c < C; c++) ao.p[c] = h->comp[comps[c]].amp + (size_t)cv.plane[0] * h->Ppad;
        unpacked = true;
      }
      if (unpacked)
        cg_recompute_pass_kernel<C, true><<<grid, DG_THREADS, 0, h->stream>>>(
            h->cg_scalars, h->M, g.x[flag_n], h->r, h->d, n2, h->partials, h->tickets, h->sums_local, 0, 1, h->peer, h->gathered, ao);
      else
        cg_recompute_pass_kernel<C, false><<<grid, DG_THREADS, 0, h->stream>>>(
            h->cg_scalars, h->M, g.x[flag_n], h->r, h->d, n2, h->partials, h->tickets, h->sums_local, 0, 1, h->peer, h->gathered, ao);
    } else {
      cg_x_fixup_kernel<<<grid, DG_THREADS, 0, h->stream>>>(h->cg_scalars, g.x[flag_n], h->d, (int64_t)(C * vs));
    }
    kt.done();
    kt.commit(final_works);
  }

  if (l2_window) {  // later kernels stream: drop the window and release the persisting lines
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    av.accessPolicyWindow.num_bytes = 0;
    CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    CK(cudaCtxResetPersistingL2Cache());
  }
  if (!unpacked)
    for (int c = 0; c < C; c++) {
      CompHost &cc = h->comp[comps[c]];
      for (int s = 0; s < S; s++)
        CK(cudaMemcpyAsync(cc.amp + (size_t)cv.plane[s] * h->Ppad, g.x[flag_n] + c * vs + (size_t)s * h->Ppad,
                           h->P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
  {
    CgScalars *hs = (CgScalars *)h->pinned;  // filled by the last read_state (the solve was done then)
    if (!sn.done || !have_state) {  // ran out of passes (i_max) without seeing the flag: read the final state
      readback(h, hs, h->cg_scalars, offsetof(CgScalars, ah));
      CK(cudaStreamSynchronize(h->stream));
    }
    int n = hs->iter < 256 ? hs->iter : 256;
    h->last_trace.assign(hs->trace, hs->trace + n);
    g.last_iter[flag_n] = hs->iter;
    if (n_iter) *n_iter = hs->iter;
    if (delta_final) *delta_final = hs->delta_new;
    // passes enqueued after convergence returned at once: they are booked as scalar launches, without bytes
    for (auto &pr : pass_recs) pr.second.commit(pr.first <= hs->iter - 1);
  }
}
}  // namespace

void cg_solve(dang_gpu *h, int cg_group, int flag_n, int ml_mode, const double *eta, uint64_t seed,
              int *n_iter, double *delta_final) {
  CgGroupHost *g = nullptr;
  for (auto &gg : h->cg)
    if (gg.set && gg.cg_group == cg_group) g = &gg;
  if (!g) fail(DANG_GPU_ESTATE, "CG group %d has not been set", cg_group);
  if (flag_n < 0 || flag_n >= g->nflag) fail(DANG_GPU_EINVAL, "flag_n %d out of range", flag_n);
  int comps[DG_MAX_COMPS], og[DG_MAX_COMPS], borders[DG_MAX_COMPS], C = 0, nog = 0, nb = 0;
  for (int c = 0; c < h->ncomp; c++) {
    const CompHost &cc = h->comp[c];
    if (!cc.set) fail(DANG_GPU_ESTATE, "component %d has not been set", c);
    if (cc.cg_group == cg_group && cc.sample_amplitude) {
      if (cc.type == DANG_COMP_T_CMB)
        fail(DANG_GPU_EUNSUPPORTED, "a 'T_cmb' component has no amplitude to fit (eval_signal = eval_sed, src/dang_component_mod.f90:770-771)");
      if (cc.is_template) borders[nb++] = c;  // template / monopole / hi_fit: border rows
      else comps[C++] = c;
    } else {
      og[nog++] = c;  // :430
      // :444-460 removes a template / monopole a second time from the bands it is not fitted to; with zero amplitudes
      // there (the constructor's default) that is a no-op, anything else is not reproduced on this path
      if (cc.type == DANG_COMP_TEMPLATE || cc.type == DANG_COMP_MONOPOLE)
        for (int j = 0; j < h->nbands; j++)
          for (int k = 0; k < h->nmaps; k++)
            if (!cc.corr[j] && cc.tamp_host[k][j] != 0.0)
              fail(DANG_GPU_EUNSUPPORTED, "an out-of-group template with non-zero amplitude in an unfitted band (component %d, band %d)", c, j);
    }
  }
  if (nb > 0) {
    cg_solve_template(h, *g, flag_n, ml_mode, eta, seed, comps, C, borders, nb, og, nog, n_iter, delta_final);
    return;
  }
  if (C == 0) fail(DANG_GPU_EINVAL, "Woah there, number of CG components = 0 for CG group %d", cg_group);
  if (C > DG_MAX_CG) fail(DANG_GPU_EUNSUPPORTED, "%d diffuse components in one CG group (max %d)", C, DG_MAX_CG);
  switch (C) {
    case 1: cg_solve_impl<1>(h, *g, flag_n, ml_mode, eta, seed, comps, og, nog, n_iter, delta_final); break;
    case 2: cg_solve_impl<2>(h, *g, flag_n, ml_mode, eta, seed, comps, og, nog, n_iter, delta_final); break;
    case 3: cg_solve_impl<3>(h, *g, flag_n, ml_mode, eta, seed, comps, og, nog, n_iter, delta_final); break;
    default: cg_solve_impl<4>(h, *g, flag_n, ml_mode, eta, seed, comps, og, nog, n_iter, delta_final); break;
  }
}
