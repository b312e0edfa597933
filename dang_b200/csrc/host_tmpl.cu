// host_tmpl.cu -- amplitude draw for a CG group holding border components -- `template`, `monopole`, `hi_fit`
// (SURVEY 8f-1): compute_rhs + compute_sample_vector + cg_search + unpack_amplitudes with their border rows,
// src/dang_cg_mod.f90:167-169 (kernels in kernels_tmpl.cuh).
#include "host.cuh"
#include "kernels_tmpl.cuh"

void cg_solve_template(dang_gpu *h, CgGroupHost &g, int flag_n, int ml_mode, const double *eta, uint64_t seed,
                       const int *comps, int C, const int *borders, int nb, const int *og, int nog, int *n_iter,
                       double *delta_final) {
  if (C > DG_TMPL_CMAX) fail(DANG_GPU_EUNSUPPORTED, "%d diffuse components next to a template / monopole / hi_fit in one CG group (max %d)", C, DG_TMPL_CMAX);
  if (nb > DG_TMPL_BMAX) fail(DANG_GPU_EUNSUPPORTED, "%d template / monopole / hi_fit components in one CG group (max %d)", nb, DG_TMPL_BMAX);
  ModelView mv = model_view(h);
  TmplView tv;
  memset(&tv, 0, sizeof tv);
  tv.C = C;
  for (int c = 0; c < C; c++) tv.comp[c] = comps[c];
  tv.S = flag_planes(g.pol_flag[flag_n], tv.plane);
  tv.nb = nb;
  tv.nt = 0;
  for (int b = 0; b < nb; b++) {
    const CompHost &bc = h->comp[borders[b]];
    tv.bcomp[b] = borders[b];
    tv.bkind[b] = bc.type == DANG_COMP_TEMPLATE ? DG_BORDER_TEMPLATE : bc.type == DANG_COMP_MONOPOLE ? DG_BORDER_MONOPOLE : DG_BORDER_HI_FIT;
    if (bc.nfit < 1) fail(DANG_GPU_EINVAL, "component %d is fitted to no band", borders[b]);
    // compute_sample_vector puts every border entry after ALL diffuse entries (:950-964) while compute_Ax follows
    // component_list order (:685-768): the two agree only when the diffuse components come first
    for (int c = 0; c < C; c++)
      if (comps[c] > borders[b])
        fail(DANG_GPU_EUNSUPPORTED, "a diffuse component after a template / monopole / hi_fit component in component_list: "
                                    "compute_sample_vector and compute_Ax lay x out differently in the reference "
                                    "(src/dang_cg_mod.f90:950-964 vs :745-768)");
    if (tv.bkind[b] == DG_BORDER_TEMPLATE) {
      if (tv.S != 2)  // compute_rhs sizes b for the template rows only in the Q+U branch (:409-414): single planes overrun it
        fail(DANG_GPU_EUNSUPPORTED, "template fits are defined for CG_POLTYPE = Q+U only (src/dang_cg_mod.f90:409-414)");
    } else if (tv.S != 1 || tv.plane[0] != 0) {
      // hi_fit / monopole columns and rows always address plane 1 and the FIRST npix entries of the work vectors
      // (:722, :736, :840, :856), which is the Stokes-I slot only when CG_POLTYPE = T
      fail(DANG_GPU_EUNSUPPORTED, "monopole / hi_fit fits are Stokes-I only: CG_POLTYPE must be T (src/dang_cg_mod.f90:717-744)");
    }
    for (int j = 0; j < DG_MAX_BANDS; j++) tv.slot_ax[b][j] = (j < h->nbands && bc.corr[j]) ? tv.nt++ : -1;
  }
  if (tv.nt > DG_TMPL_MAX) fail(DANG_GPU_EUNSUPPORTED, "%d fitted (component, band) pairs in one CG group (max %d)", tv.nt, DG_TMPL_MAX);
  {  // compute_sample_vector's running counter (:970): bands outer, border components inner, never reset
    int l = 0;
    for (int j = 0; j < DG_MAX_BANDS; j++)
      for (int b = 0; b < nb; b++) tv.slot_sv[b][j] = (j < h->nbands && h->comp[borders[b]].corr[j]) ? l++ : -1;
  }
  tv.nog = nog;
  for (int o = 0; o < nog; o++) tv.og[o] = og[o];
  const int S = tv.S;
  for (int s = 0; s < S; s++)
    if (tv.plane[s] >= h->nmaps) fail(DANG_GPU_EINVAL, "pol flag needs plane %d, nmaps = %d", tv.plane[s] + 1, h->nmaps);
  const size_t vs = (size_t)S * h->Ppad;   // doubles per diffuse component
  const size_t nd = (size_t)(C > 0 ? C : 1) * vs;
  const int64_t n_el = (int64_t)C * (int64_t)vs;
  // plane whose template_amplitudes seed / receive the tail (initialize_x :1226-1279, unpack :1337-1392)
  auto tail_plane = [&](int b, int s) { return tv.bkind[b] == DG_BORDER_TEMPLATE ? tv.plane[s] : 0; };

  // self%x: allocate + seed on first use only (:227-239, Q10): diffuse planes from c%amplitude, the tail from
  // template_amplitudes (initialize_x :1264-1279)
  if (!g.x[flag_n] || g.x_len[flag_n] != nd) {
    if (g.x[flag_n]) CK(cudaFree(g.x[flag_n]));
    CK(cudaMalloc(&g.x[flag_n], nd * sizeof(double)));
    g.x_len[flag_n] = nd;
    CK(cudaMemsetAsync(g.x[flag_n], 0, nd * sizeof(double), h->stream));
    for (int c = 0; c < C; c++)
      for (int s = 0; s < S; s++)
        CK(cudaMemcpyAsync(g.x[flag_n] + c * vs + (size_t)s * h->Ppad, h->comp[comps[c]].amp + (size_t)tv.plane[s] * h->Ppad,
                           h->P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    g.xt_set[flag_n] = false;
  }
  if (!g.xt_set[flag_n]) {
    for (int b = 0; b < nb; b++) {
      const CompHost &bc = h->comp[borders[b]];
      for (int j = 0; j < h->nbands; j++)
        if (tv.slot_ax[b][j] >= 0) g.xt[flag_n][tv.slot_ax[b][j]] = bc.tamp_host[tail_plane(b, 0)][j];
    }
    g.xt_set[flag_n] = true;
  }
  if (h->v_len < nd) {
    size_t l1 = h->v_len, l2 = h->v_len;
    ensure(h->r, l1, nd);
    ensure(h->d, l2, nd);
    h->v_len = nd;
    h->cg_layout = -1;
  }
  if (h->t_len < nd) {
    size_t l1 = h->t_len, l2 = h->t_len;
    ensure(h->tb, l1, nd);
    ensure(h->tq, l2, nd);
    h->t_len = nd;
  }
  h->cg_layout = -1;  // the block-diagonal kernels must re-zero their padding after this solve used r / d
  CK(cudaMemsetAsync(h->r, 0, nd * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->d, 0, nd * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->tb, 0, nd * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->tq, 0, nd * sizeof(double), h->stream));
  if (!h->tmpl_scalars) CK(cudaMalloc(&h->tmpl_scalars, sizeof(TmplScalars)));
  CK(cudaMemsetAsync(h->tmpl_scalars, 0, sizeof(TmplScalars), h->stream));
  CK(cudaMemcpyAsync((char *)h->tmpl_scalars + offsetof(TmplScalars, xt), g.xt[flag_n], DG_TMPL_MAX * sizeof(double),
                     cudaMemcpyHostToDevice, h->stream));
  tv.b = h->tb;
  tv.q = h->tq;
  tv.r = h->r;
  tv.d = h->d;
  tv.x = g.x[flag_n];
  tv.seed = seed;
  tv.fluct = 0;
  tv.eta = nullptr;
  if (ml_mode == DANG_ML_SAMPLE) {
    tv.fluct = h->fix_q1 ? 2 : 1;
    if (eta) {
      ensure(h->eta, h->eta_len, vs);
      h2d_planes(h, h->eta, eta, S);
      tv.eta = h->eta;
    }
  }
  const int cnt = DG_TMPL_NV;
  const int grid = occ_grid(h, tmpl_apply_kernel, h->P, DG_THREADS);
  const int gridv = grid_for(h, n_el > 0 ? n_el : 1, DG_THREADS, 4);
  const double el = (double)S * h->P;
  TmplScalars *sc = h->tmpl_scalars;

  CK(cudaMemsetAsync(h->sums_local, 0, GATHER_MAX * sizeof(double), h->stream));
  {
    KTimer kt(h, DANG_K_RHS_BLOCKS, bytes_w(el * (2.0 * h->nbands + 2 + C)));
    tmpl_rhs_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, tv, h->partials, h->tickets, h->sums_local);
    kt.done();
  }
  gather(h, cnt);
  tmpl_s_rhs_kernel<<<1, 1, 0, h->stream>>>(sc, h->gathered, h->nranks, cnt, tv.nt, g.i_max, g.converge);
  auto apply = [&](int mode) {
    KTimer kt(h, DANG_K_CG_DQ, bytes_w(el * (h->nbands + 1.0 + 3.0 * C)));
    tmpl_apply_kernel<<<grid, DG_THREADS, 0, h->stream>>>(mv, tv, sc, mode, h->partials, h->tickets, h->sums_local);
    kt.done();
    gather(h, cnt);
  };
  apply(0);  // q = A x0
  tmpl_s_resid_a_kernel<<<1, 1, 0, h->stream>>>(sc, h->gathered, h->nranks, cnt, tv.nt);
  {
    KTimer kt(h, DANG_K_CG_UPDATE, bytes_w(el * 4.0 * C));
    tmpl_resid_kernel<<<gridv, DG_THREADS, 0, h->stream>>>(tv, n_el, h->partials, h->tickets, h->sums_local);
    kt.done();
  }
  gather(h, cnt);
  tmpl_s_resid_b_kernel<<<1, 1, 0, h->stream>>>(sc, h->gathered, h->nranks, cnt, tv.nt);
  CK(cudaGetLastError());

  TmplScalars *hs = (TmplScalars *)h->pinned;
  static_assert(sizeof(TmplScalars) <= 32 * 1024, "TmplScalars must fit the pinned read-back buffer");
  auto read_state = [&]() {
    readback(h, hs, sc, sizeof(TmplScalars));
    CK(cudaStreamSynchronize(h->stream));
  };
  const int max_pass = g.i_max - 1;
  int enq = 0;
  int batch = g.last_iter[flag_n] > 1 ? g.last_iter[flag_n] - 1 : h->cg_chunk;
  bool done = max_pass < 1;
  while (!done && enq < max_pass) {
    if (batch > max_pass - enq) batch = max_pass - enq;
    for (int it = 0; it < batch; it++) {
      apply(1);  // d = r + beta d (stored), q = A d
      tmpl_s_apply_kernel<<<1, 1, 0, h->stream>>>(sc, h->gathered, h->nranks, cnt, tv.nt);
      {
        KTimer kt(h, DANG_K_CG_UPDATE, bytes_w(el * 6.0 * C));
        tmpl_update_kernel<<<gridv, DG_THREADS, 0, h->stream>>>(tv, sc, n_el, h->partials, h->tickets, h->sums_local);
        kt.done();
      }
      gather(h, cnt);
      tmpl_s_update_kernel<<<1, 1, 0, h->stream>>>(sc, h->gathered, h->nranks, cnt, tv.nt);
      CK(cudaGetLastError());
    }
    enq += batch;
    read_state();
    done = hs->done != 0;
    batch = h->cg_chunk;
  }
  if (enq == 0 || (!done && enq >= max_pass)) read_state();

  // unpack_amplitudes: diffuse planes (:1327-1335); template amplitudes to planes 2 and 3 (:1374-1392), hi_fit and
  // monopole amplitudes to plane 1 (:1337-1372) -- and, for a monopole, on into the band offsets (update_sky_model)
  for (int c = 0; c < C; c++) {
    CompHost &cc = h->comp[comps[c]];
    amp_write_barrier(h, cc);
    for (int s = 0; s < S; s++)
      CK(cudaMemcpyAsync(cc.amp + (size_t)tv.plane[s] * h->Ppad, g.x[flag_n] + c * vs + (size_t)s * h->Ppad,
                         h->P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  for (int b = 0; b < nb; b++) {
    CompHost &bc = h->comp[borders[b]];
    const int SB = tv.bkind[b] == DG_BORDER_TEMPLATE ? S : 1;
    for (int j = 0; j < h->nbands; j++) {
      const int slot = tv.slot_ax[b][j];
      if (slot < 0) continue;
      g.xt[flag_n][slot] = hs->xt[slot];
      for (int s = 0; s < SB; s++) bc.tamp_host[tail_plane(b, s)][j] = hs->xt[slot];
    }
    upload_tamp(h, bc);
    monopole_to_offset(h, bc);
  }
  const int n = hs->iter < 256 ? hs->iter : 256;
  h->last_trace.assign(hs->trace, hs->trace + n);
  g.last_iter[flag_n] = hs->iter;
  if (n_iter) *n_iter = hs->iter;
  if (delta_final) *delta_final = hs->delta_new;
}
