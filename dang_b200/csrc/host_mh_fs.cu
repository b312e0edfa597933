// host_mh_fs.cu -- full-sky branch of sample_index_mh (src/dang_sample_mod.f90:229-329) and the
// step-size tuner (:623-717).
#include "host.cuh"
#include "kernels_mh.cuh"
#include "kernels_uni.cuh"
#include "kernels_stream.cuh"

void mh_view(dang_gpu *h, int ic, int nind, int map_n, int nsample, int ml_mode, MhView &mh) {
  if (ic < 0 || ic >= h->ncomp || !h->comp[ic].set) fail(DANG_GPU_EINVAL, "bad component %d", ic);
  const CompHost &c = h->comp[ic];
  if (nind < 0 || nind >= c.nind) fail(DANG_GPU_EINVAL, "component %d has no index %d", ic, nind);
  const IndexHost &ix = c.index[nind];
  if (ix.sample_nside != h->nside)
    fail(DANG_GPU_EUNSUPPORTED, "sample_nside %d /= nside %d needs HEALPix udgrade_ring (DESIGN.md, out of scope)",
         ix.sample_nside, h->nside);
  memset(&mh, 0, sizeof mh);
  mh.ic = ic;
  mh.nind = nind;
  if (map_n == -1) {  // :157-163
    mh.S = 2;
    mh.plane[0] = 1;
    mh.plane[1] = 2;
  } else if (map_n >= 1 && map_n <= 3) {
    mh.S = 1;
    mh.plane[0] = mh.plane[1] = map_n - 1;
  } else {
    fail(DANG_GPU_EUNSUPPORTED, "map_n = %d (T+Q+U) is unreachable in the reference (SURVEY Q2)", map_n);
  }
  for (int s = 0; s < mh.S; s++)
    if (mh.plane[s] >= h->nmaps) fail(DANG_GPU_EINVAL, "map_n %d needs plane %d, nmaps = %d", map_n, mh.plane[s] + 1, h->nmaps);
  mh.nsample = nsample;
  mh.ml_mode = ml_mode;
  mh.lnl_type = ix.lnl_type;
  mh.prior_type = ix.prior_type;
  mh.is_synch = c.label == "synch";
  mh.gauss[0] = ix.gauss[0];
  mh.gauss[1] = ix.gauss[1];
  mh.uni[0] = ix.uni[0];
  mh.uni[1] = ix.uni[1];
  mh.step = ix.step;
}

void ensure_zu(dang_gpu *h, size_t n) {
  if (h->zu_len >= n) return;
  dfree(h->zbuf);
  dfree(h->ubuf);
  CK(cudaMalloc(&h->zbuf, n * sizeof(double)));
  CK(cudaMalloc(&h->ubuf, n * sizeof(double)));
  h->zu_len = n;
}

void ensure_decisions(dang_gpu *h, size_t n) {
  if (h->dec_len >= n) return;
  dfree(h->decisions);
  dfree(h->lnl_trace);
  CK(cudaMalloc(&h->decisions, n));
  CK(cudaMalloc(&h->lnl_trace, n * sizeof(double)));
  h->dec_len = n;
}

namespace {
// chain start (sample <- indices at global pixel 0) + sufficient statistics, gathered over ranks
int fullsky_statistics(dang_gpu *h, const ModelView &mv, MhView &mh) {
  CK(cudaMemsetAsync(h->sums_local, 0, GATHER_MAX * sizeof(double), h->stream));
  // chain start: sample <- indices at global pixel 0, broadcast, SED at the start -- unless the previous draw of
  // this very index left all of that on the device (chain continuity, host.cuh)
  const bool cont = !h->fullsky_stream && h->fs_cont_valid && h->fs_cont_ic == mh.ic && h->fs_cont_nind == mh.nind &&
                    h->fs_cont_S == mh.S && h->fs_cont_plane0 == mh.plane[0] && h->fs_cont_epoch == h->idx_epoch;
  if (!cont) {
    {
      KTimer kt(h, DANG_K_SCALAR, 0);
      mh_first_pixel_kernel<<<1, 1, 0, h->stream>>>(mv, mh, h->sums_local);
      kt.done();
    }
    gather(h, 2);
    {
      KTimer kt(h, DANG_K_SCALAR, 0);
      mh_fullsky_init_kernel<<<1, 1, 0, h->stream>>>(mv, mh, h->mh_scalars, h->gathered, 2);
      kt.done();
    }
  }
  h->fs_cont_valid = false;  // (re-armed by the draw that consumes these statistics)
  if (h->fullsky_stream) return 0;
  const double n_el = (double)mh.S * h->P;
  const int nchunk = (h->nbands + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
  const int cnt = mh.S * nchunk * 3 * DG_SUFF_CHUNK;  // per plane, per band chunk: X, Y, Z
  if (cnt > GATHER_MAX) fail(DANG_GPU_EUNSUPPORTED, "sufficient statistics of %d bands x %d planes", h->nbands, mh.S);
  KTimer kt(h, DANG_K_MH_SUFFSTAT, bytes_w(n_el * (2.0 * h->nbands + h->ncomp)));
  bool uni = h->ncomp <= 4;
  for (int s = 0; s < mh.S && uni; s++)
    for (int c = 0; c < h->ncomp; c++)
      if (c != mh.ic) uni = uni && comp_uniform(h, c, mh.plane[s]);
  if (uni) {
    // one (plane, band chunk) combination per block: the grid is a multiple of the number of combinations
    const int ncombo = mh.S * nchunk;
    int g2 = occ_grid(h, mh_suffstat_uni_kernel<4>, h->Ppad / 2 * ncombo, DG_THREADS);
    g2 = g2 / ncombo * ncombo;
    if (g2 < ncombo) g2 = ncombo;
    if (h->stream_ring) {  // asynchronous stream: per-thread cp.async ring in shared memory (kernels_stream.cuh)
      const size_t ring_smem = (size_t)DG_RING_STAGES * DG_RING_SLOTS * DG_THREADS * sizeof(double2);
      auto launch = [&](auto kernel) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
        int g3 = occ_grid(h, kernel, h->Ppad / 2 * ncombo, DG_THREADS, ring_smem);
        g3 = g3 / ncombo * ncombo;
        if (g3 < ncombo) g3 = ncombo;
        kernel<<<g3, DG_THREADS, ring_smem, h->stream>>>(mv, mh, h->mh_scalars, h->partials, h->tickets, h->sums_local);
      };
      if (h->ncomp <= 2) launch(mh_suffstat_ring_kernel<2>);
      else launch(mh_suffstat_ring_kernel<4>);
    } else if (h->ncomp <= 2) mh_suffstat_uni_kernel<2><<<g2, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars, h->partials, h->tickets, h->sums_local);
    else mh_suffstat_uni_kernel<4><<<g2, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars, h->partials, h->tickets, h->sums_local);
  } else {
    const int grid = occ_grid(h, mh_suffstat_kernel, h->P, DG_THREADS);
    mh_suffstat_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars, h->partials, h->tickets, h->sums_local);
  }
  kt.done();
  gather(h, cnt);
  // keep the gathered rows: they serve this draw, and the chi-square before and after it
  if (!h->stat_buf) CK(cudaMalloc(&h->stat_buf, (size_t)GATHER_MAX * DG_MAX_RANKS * sizeof(double)));
  CK(cudaMemcpyAsync(h->stat_buf, h->gathered, (size_t)h->nranks * cnt * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->stat_valid = true;
  h->stat_ic = mh.ic;
  h->stat_nind = mh.nind;
  h->stat_S = mh.S;
  h->stat_plane[0] = mh.plane[0];
  h->stat_plane[1] = mh.plane[1];
  h->stat_cnt = cnt;
  h->stat_version = h->version;
  return cnt;
}

bool stat_cache_hit(const dang_gpu *h, const MhView &mh) {
  return h->stat_cache && h->stat_valid && h->stat_version == h->version && h->stat_ic == mh.ic &&
         h->stat_nind == mh.nind && h->stat_S == mh.S && h->stat_plane[0] == mh.plane[0] &&
         h->stat_plane[1] == mh.plane[1];
}

// the sufficient-statistics form covers the chisq likelihood with uniform / Gaussian prior; the
// marginal likelihood, the Jeffreys prior and 'prior' draws stream the maps per proposal
bool fullsky_needs_stream(const MhView &mh) {
  return mh.lnl_type != DANG_LNL_CHISQ || mh.prior_type == DANG_PRIOR_JEFFREYS;
}

void upload_fullsky_deviates(dang_gpu *h, MhView &mh, const double *z, const double *u, size_t n) {
  if (!z) return;
  ensure_zu(h, n > 0 ? n : 1);
  CK(cudaMemcpyAsync(h->zbuf, z, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  mh.z = h->zbuf;
  if (u) {
    CK(cudaMemcpyAsync(h->ubuf, u, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    mh.u = h->ubuf;
  } else if (mh.ml_mode == DANG_ML_SAMPLE) {
    fail(DANG_GPU_EINVAL, "z injected without u");
  }
}
}  // namespace

void sample_fullsky(dang_gpu *h, MhView &mh, const double *z, const double *u, uint64_t seed,
                    double *accept) {
  const int saved_stream = h->fullsky_stream;
  struct Restore {
    dang_gpu *h; int v;
    ~Restore() { h->fullsky_stream = v; }
  } restore{h, saved_stream};
  if (fullsky_needs_stream(mh)) h->fullsky_stream = 1;
  ModelView mv = model_view(h);
  mh.seed = seed;
  const size_t n = (size_t)mh.nsample;
  // short chains on the statistics form take their injected deviates as a kernel argument (no copy-engine work)
  static thread_local FsDeviates dev;
  dev.n = 0;
  if (z && !h->fullsky_stream && n > 0 && n <= DG_FS_DEV_MAX && (u || mh.ml_mode != DANG_ML_SAMPLE)) {
    dev.n = (int)n;
    for (size_t i = 0; i < n; i++) {
      dev.z[i] = z[i];
      dev.u[i] = u ? u[i] : 0.0;
    }
  } else {
    upload_fullsky_deviates(h, mh, z, u, n);
  }
  ensure_decisions(h, n > 0 ? n : 1);
  CK(cudaMemsetAsync(h->decisions, 3, n, h->stream));
  CK(cudaMemsetAsync(h->lnl_trace, 0xff, n * sizeof(double), h->stream));
  mh.decisions = h->decisions;
  mh.lnl_trace = h->lnl_trace;
  h->dec_mode = 1;
  h->dec_nsample = mh.nsample;

  // statistics gathered by the chi-square call that preceded this draw are still valid when the
  // model state has not changed since (same chain start, same data)
  h->fs_tab_written = false;
  const bool reuse = !h->fullsky_stream && stat_cache_hit(h, mh);
  const int cnt = reuse ? h->stat_cnt : fullsky_statistics(h, mv, mh);
  const double n_el = (double)mh.S * h->P;
  if (h->fullsky_stream) {
    const size_t dl = (size_t)h->nbands * mh.S * h->Ppad;
    ensure(h->D, h->D_len, dl);
    const int grid = grid_for(h, h->P, DG_THREADS, 4);
    {
      KTimer kt(h, DANG_K_MH_DATA, bytes_w(n_el * (2.0 * h->nbands + h->ncomp * 2.0)));
      mh_data_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->D);
      kt.done();
    }
    const int nchunk = (h->nbands + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
    const int cnt = mh.lnl_type == DANG_LNL_MARGINAL ? 2 + 4 * DG_SUFF_CHUNK * nchunk : 2;
    if (cnt > GATHER_MAX) fail(DANG_GPU_EUNSUPPORTED, "full-sky marginal lnL with %d bands", h->nbands);
    if (mh.lnl_type == DANG_LNL_PRIOR) {  // :255-257: no chain, draw the index from its Gaussian prior
      KTimer ks(h, DANG_K_SCALAR, 0);
      mh_fullsky_prior_draw_kernel<<<1, 1, 0, h->stream>>>(mh, h->mh_scalars);
      ks.done();
    }
    for (int l = 0; l <= mh.nsample && mh.lnl_type != DANG_LNL_PRIOR; l++) {  // starting point + proposals
      CK(cudaMemsetAsync(h->sums_local, 0, GATHER_MAX * sizeof(double), h->stream));
      if (mh.lnl_type == DANG_LNL_CHISQ || mh.prior_type == DANG_PRIOR_JEFFREYS) {
        KTimer kt(h, DANG_K_MH_FULLSKY_LNL, bytes_w(n_el * (2.0 * h->nbands + 1)));
        mh_fullsky_lnl_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars, h->D, h->partials,
                                                                 h->tickets, h->sums_local);
        kt.done();
      }
      if (mh.lnl_type == DANG_LNL_MARGINAL) {
        KTimer kt(h, DANG_K_MH_FULLSKY_LNL, bytes_w(n_el * (2.0 * h->nbands + 1)));
        mh_fullsky_marginal_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars, h->D, h->partials,
                                                                      h->tickets, h->sums_local);
        kt.done();
      }
      gather(h, cnt);
      KTimer ks(h, DANG_K_SCALAR, 0);
      mh_fullsky_step_kernel<<<1, 1, 0, h->stream>>>(mv, mh, h->mh_scalars, h->gathered, h->nranks, cnt);
      ks.done();
    }
  } else {
    // planes already tabulated (constant index maps) and no per-plane factor in the SED: the chain kernel refreshes
    // the table entry itself and leaves the next chain's start behind
    bool keep = h->comp[mh.ic].type != DANG_COMP_HI_FIT && !h->record;
    for (int s = 0; s < mh.S; s++) keep = keep && comp_uniform(h, mh.ic, mh.plane[s]) && !h->tab_dirty;
    KTimer ks(h, DANG_K_SCALAR, 0);
    mh_suff_chain_kernel<<<1, 32, 0, h->stream>>>(mv, mh, h->mh_scalars, h->stat_buf, h->nranks, cnt, keep ? h->tab : nullptr, dev);
    ks.done();
    h->fs_tab_written = keep;
    if (keep) {
      h->fs_cont_valid = true;
      h->fs_cont_ic = mh.ic;
      h->fs_cont_nind = mh.nind;
      h->fs_cont_S = mh.S;
      h->fs_cont_plane0 = mh.plane[0];
      h->fs_cont_epoch = h->idx_epoch;
    }
  }
  // The chain's results go back first and the host waits for THAT point of the stream only; the kernel that
  // writes the final sample into the index planes (:329, :483) runs while the host is already enqueuing the
  // next call.
  bool has_monopole = false;
  for (int c = 0; c < h->ncomp; c++) has_monopole = has_monopole || (h->comp[c].set && h->comp[c].type == DANG_COMP_MONOPOLE);
  if (h->defer_scalars && !h->fullsky_stream) {  // deferred scalars (host.cuh): nothing comes back here
    const int grid = grid_for(h, h->P, DG_THREADS, 4);
    KTimer kt(h, DANG_K_SCALAR, 0);
    mh_fullsky_store_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars);
    kt.done();
    if (accept) *accept = nan("");
    touch(h, 2);
    h->stat_valid = false;
    h->chisq_valid = false;
    h->pend_draw = true;
    h->pend_ic = mh.ic;
    h->pend_nind = mh.nind;
    h->pend_S = mh.S;
    h->pend_plane[0] = mh.plane[0];
    h->pend_plane[1] = mh.plane[1];
    h->pend_draw_chisq = h->stat_cache && !has_monopole;
    h->pend_draw_version = h->version;
    h->comp[mh.ic].index[mh.nind].last_value_planes = -1;  // the host does not know the new value (sample_index checks)
    return;
  }
  MhScalars *hs = (MhScalars *)h->pinned;
  readback(h, hs, h->mh_scalars, sizeof(MhScalars));
  CK(cudaEventRecord(h->ev_sync, h->stream));
  {
    const int grid = grid_for(h, h->P, DG_THREADS, 4);
    KTimer kt(h, DANG_K_SCALAR, 0);
    mh_fullsky_store_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars);
    kt.done();
  }
  CK(cudaEventSynchronize(h->ev_sync));
  if (accept) *accept = hs->accept;
  h->comp[mh.ic].index[mh.nind].last_value = hs->sample[mh.nind];
  touch(h, 2);
  h->stat_valid = false;
  if (!h->fullsky_stream && h->stat_cache && !has_monopole) {  // chi-square of the new state, from the statistics
    h->chisq_valid = true;
    h->chisq_version = h->version;
    h->chisq_lo = mh.plane[0] + 1;
    h->chisq_hi = mh.plane[mh.S - 1] + 1;
    for (int k = 0; k < 4; k++) h->chisq_vals[k] = 0.0;
    for (int sI = 0; sI < mh.S; sI++) h->chisq_vals[mh.plane[sI]] = hs->chisq[sI];
  }
}

// compute_chisq right after an amplitude draw, when the next call in sample_spectral_parameters'
// order (components -> indices -> pol flags, src/dang_sample_mod.f90:39-74) is a full-sky draw with
// the chisq likelihood over the same planes: gather that draw's per-plane statistics now.  X_j is
// the chi-square term of band j about the current state, so chisq_planes = sum_j X_j / nbands, and
// the draw that follows (and the chi-square after it) need no further pass over the maps.
// The full-sky chisq draw that follows an amplitude draw in sample_spectral_parameters' order, if there is one
// whose statistics can be gathered ahead of time.
static bool upcoming_fullsky_draw(dang_gpu *h, MhView &mh) {
  if (!h->stat_cache || h->fullsky_stream || h->last_mutation != 1) return false;
  // with a monopole the draw's data ((sig - offset) / gain minus EVERY other component, monopole included,
  // src/dang_sample_mod.f90:173-196) is not the chi-square's residual (monopole left out, dang_data_mod.f90:357-361)
  for (int c = 0; c < h->ncomp; c++)
    if (h->comp[c].set && h->comp[c].type == DANG_COMP_MONOPOLE) return false;
  int ic = -1, nind = -1;
  for (int c = 0; c < h->ncomp && ic < 0; c++)
    for (int l = 0; l < h->comp[c].nind && ic < 0; l++)
      if (h->comp[c].set && h->comp[c].index[l].sample_index && h->comp[c].index[l].nflag > 0) {
        ic = c;
        nind = l;
      }
  if (ic < 0) return false;
  const IndexHost &ix = h->comp[ic].index[nind];
  if (ix.index_mode != DANG_INDEX_FULLSKY || ix.sample_nside != h->nside) return false;
  const int flag = ix.pol_flag[0];
  const int map_n = (flag & 8) ? -1 : (flag & 1) ? 1 : (flag & 2) ? 2 : (flag & 4) ? 3 : 0;
  if (map_n == 0) return false;
  try {
    mh_view(h, ic, nind, map_n, 0, DANG_ML_SAMPLE, mh);
  } catch (const DgError &) {
    return false;  // the draw itself will report what is wrong with it
  }
  return !fullsky_needs_stream(mh);
}

// Called by the amplitude draw right after its last kernel is enqueued and BEFORE the host waits for the solve's
// scalars: when the configuration has a full-sky chisq draw coming, its statistics pass (which also serves the
// chi-square the reference prints after the draw, write_stats_to_term) goes onto the stream now, so the device
// never idles while the host turns the solve around.
void prefetch_statistics(dang_gpu *h) {
  MhView mh;
  if (!upcoming_fullsky_draw(h, mh)) return;
  if (h->maps_set && h->n_unmasked < 0) return;  // first call: let compute_chisq count the mask first
  ModelView mv = model_view(h);
  if (!stat_cache_hit(h, mh)) fullsky_statistics(h, mv, mh);
}

bool chisq_from_statistics(dang_gpu *h, int pol_lo, int pol_hi, double out4[4]) {
  MhView mh;
  if (!upcoming_fullsky_draw(h, mh)) return false;
  if (mh.plane[0] != pol_lo - 1 || mh.plane[mh.S - 1] != pol_hi - 1 || pol_hi - pol_lo + 1 != mh.S) return false;
  ModelView mv = model_view(h);
  const int64_t n_unmasked = unmasked_count(h);
  int cnt = h->stat_cnt;
  if (!stat_cache_hit(h, mh)) cnt = fullsky_statistics(h, mv, mh);
  if (h->defer_scalars) {  // deferred scalars: the rows stay in stat_buf until dang_gpu_iteration_mark snapshots them
    for (int k = 0; k < 3; k++) out4[k] = nan("");
    out4[3] = (double)n_unmasked;
    h->pend_chisq_cg = true;
    h->pend_S = mh.S;
    h->pend_plane[0] = mh.plane[0];
    h->pend_plane[1] = mh.plane[1];
    h->pend_stat_cnt = cnt;
    return true;
  }
  double *hp = (double *)h->pinned;
  readback(h, hp, h->stat_buf, (size_t)h->nranks * cnt * sizeof(double));
  CK(cudaStreamSynchronize(h->stream));
  chisq_of_statistics(h, hp, cnt, mh.S, mh.plane, out4);
  out4[3] = (double)n_unmasked;
  return true;
}

// chisq_planes = sum_j X_j / nbands from gathered statistics rows (host copy)
void chisq_of_statistics(const dang_gpu *h, const double *hp, int cnt, int S, const int *plane, double out4[4]) {
  const int B = h->nbands, nchunk = (B + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
  for (int k = 0; k < 4; k++) out4[k] = 0.0;
  for (int s = 0; s < S; s++) {
    double tot = 0.0;
    for (int j = 0; j < B; j++) {
      const int o = (s * nchunk + j / DG_SUFF_CHUNK) * 3 * DG_SUFF_CHUNK + 3 * (j % DG_SUFF_CHUNK);
      double X = 0.0;
      for (int g = 0; g < h->nranks; g++) X += hp[(size_t)g * cnt + o];  // rank order: same bits on every rank
      tot += X / (double)B;
    }
    out4[plane[s]] = tot;
  }
}

// tune_spectral_parameter_length, src/dang_sample_mod.f90:623-717.  `start`: the chain start when the call site gives
// one (per-pixel indices: the map's mean, :341-347), else indices(0, map_inds(1), :) (:240-243).  The chisq likelihood
// runs on the sufficient statistics (one pass over the maps, all blocks in one kernel); the marginal likelihood
// (:650-651, :676-677) streams the maps once per proposal like the draw itself, one host round trip per block.
void tune_fullsky(dang_gpu *h, int ic, int nind, MhView &mh, const double *z, const double *u, uint64_t seed,
                  int max_blocks, int *blocks_run, double *step_size, const double *start) {
  if (mh.prior_type == DANG_PRIOR_JEFFREYS || mh.lnl_type == DANG_LNL_PRIOR)
    fail(DANG_GPU_EUNSUPPORTED, "the step-size tuner covers the chisq / marginal likelihoods with uniform / Gaussian prior "
                                "(the reference's tuner has no other branch, src/dang_sample_mod.f90:650-662)");
  if (max_blocks < 0) fail(DANG_GPU_EINVAL, "max_blocks = %d", max_blocks);
  ModelView mv = model_view(h);
  mh.seed = seed;
  mh.rng_z_stream = DG_STREAM_TUNE_Z;
  mh.rng_u_stream = DG_STREAM_TUNE_U;
  upload_fullsky_deviates(h, mh, z, u, (size_t)mh.nsample * max_blocks);
  const bool stream = mh.lnl_type == DANG_LNL_MARGINAL;
  const int saved_stream = h->fullsky_stream;
  struct Restore {
    dang_gpu *h; int v;
    ~Restore() { h->fullsky_stream = v; }
  } restore{h, saved_stream};
  h->fs_cont_valid = false;
  if (start) {  // the start is not what the maps hold at pixel 0: set it, then gather the statistics about it
    h->fullsky_stream = 1;  // (chain-start kernels only)
    fullsky_statistics(h, mv, mh);
    mh_fullsky_override_kernel<<<1, 1, 0, h->stream>>>(mv, mh, h->mh_scalars, start[0], start[1]);
    CK(cudaGetLastError());
    h->launches++;
  }
  if (!stream) {
    h->fullsky_stream = 0;  // the tuner runs on the sufficient statistics
    if (start) h->fs_cont_valid = true, h->fs_cont_ic = mh.ic, h->fs_cont_nind = mh.nind, h->fs_cont_S = mh.S,
               h->fs_cont_plane0 = mh.plane[0], h->fs_cont_epoch = h->idx_epoch;  // keep the overridden start
    const int cnt = fullsky_statistics(h, mv, mh);
    h->fs_cont_valid = false;
    h->stat_valid = false;  // (statistics about a start that is not the maps' state must not serve a chi-square)
    double *d_out = h->sums_local + GATHER_MAX - 4;  // scratch beyond the statistics rows
    {
      KTimer ks(h, DANG_K_SCALAR, 0);
      mh_suff_tune_kernel<<<1, 32, 0, h->stream>>>(mv, mh, h->mh_scalars, h->stat_buf, h->nranks, cnt, max_blocks, d_out);
      ks.done();
    }
    double *hp = (double *)h->pinned;
    readback(h, hp, d_out, 3 * sizeof(double));
    CK(cudaStreamSynchronize(h->stream));
    h->comp[ic].index[nind].step = hp[0];  // c%step_size(nind), :708-710
    if (blocks_run) *blocks_run = (int)hp[1];
    if (step_size) *step_size = hp[0];
    return;
  }
  // ---- streaming form (marginal likelihood)
  h->fullsky_stream = 1;
  if (!start) fullsky_statistics(h, mv, mh);  // chain start from the maps
  const double n_el = (double)mh.S * h->P;
  const size_t dl = (size_t)h->nbands * mh.S * h->Ppad;
  ensure(h->D, h->D_len, dl);
  const int grid = grid_for(h, h->P, DG_THREADS, 4);
  {
    KTimer kt(h, DANG_K_MH_DATA, bytes_w(n_el * (2.0 * h->nbands + h->ncomp * 2.0)));
    mh_data_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->D);
    kt.done();
  }
  const int nchunk = (h->nbands + DG_SUFF_CHUNK - 1) / DG_SUFF_CHUNK;
  const int cnt = 2 + 4 * DG_SUFF_CHUNK * nchunk;
  if (cnt > GATHER_MAX) fail(DANG_GPU_EUNSUPPORTED, "full-sky marginal lnL with %d bands", h->nbands);
  mh.decisions = nullptr;
  mh.lnl_trace = nullptr;
  const double *z0 = mh.z, *u0 = mh.u;
  auto lnl_pass = [&]() {  // lnL of ms->sed (the start, or the pending proposal), then accept / reject + next proposal
    CK(cudaMemsetAsync(h->sums_local, 0, GATHER_MAX * sizeof(double), h->stream));
    KTimer kt(h, DANG_K_MH_FULLSKY_LNL, bytes_w(n_el * (2.0 * h->nbands + 1)));
    mh_fullsky_marginal_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, mh, h->mh_scalars, h->D, h->partials, h->tickets, h->sums_local);
    kt.done();
    gather(h, cnt);
    KTimer ks(h, DANG_K_SCALAR, 0);
    mh_fullsky_step_kernel<<<1, 1, 0, h->stream>>>(mv, mh, h->mh_scalars, h->gathered, h->nranks, cnt);
    ks.done();
  };
  double step = mh.step;
  int blk = 0, tuned = 0;
  MhScalars *hs = (MhScalars *)h->pinned;
  while (!tuned && blk < max_blocks) {
    mh.step = step;
    mh.rng_slot0 = (long long)blk * mh.nsample;
    mh.z = z0 ? z0 + (size_t)blk * mh.nsample : nullptr;
    mh.u = u0 ? u0 + (size_t)blk * mh.nsample : nullptr;
    if (blk == 0) {
      lnl_pass();  // phase 0: lnL of the start (:647-662); proposes the first candidate of block 0
    } else {
      KTimer ks(h, DANG_K_SCALAR, 0);
      mh_tune_block_start_kernel<<<1, 1, 0, h->stream>>>(mv, mh, h->mh_scalars);
      ks.done();
    }
    for (int l = 0; l < mh.nsample; l++) lnl_pass();  // passes behind the block's last proposal return at once (ms->skip)
    readback(h, hs, h->mh_scalars, sizeof(MhScalars));
    CK(cudaStreamSynchronize(h->stream));
    const double rate = hs->accept / (double)(mh.nsample + 1);  // Fortran loop counter after the loop (:707)
    if (rate < (double)0.4f) step = step - (double)0.5f * step;
    else if (rate > (double)0.6f) step = step + (double)0.5f * step;
    else tuned = 1;
    blk++;
  }
  h->comp[ic].index[nind].step = step;
  if (blocks_run) *blocks_run = blk;
  if (step_size) *step_size = step;
}
