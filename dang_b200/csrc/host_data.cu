// host_data.cu -- update_sky_model + compute_chisq, src/dang_data_mod.f90:339-396,494-526.
#include "host.cuh"
#include "kernels_data.cuh"
#include "kernels_uni.cuh"

namespace {
template <int NC>
void launch_chisq(dang_gpu *h, const ModelView &mv, const ChisqView &cv, int) {
  const int grid = occ_grid(h, chisq_kernel<NC>, h->P, DG_THREADS);
  chisq_kernel<NC><<<grid, DG_THREADS, 0, h->stream>>>(mv, cv, h->partials, h->tickets, h->sums_local);
}
}  // namespace

void run_chisq(dang_gpu *h, int pol_lo, int pol_hi, double *sky, double *res, double *chi_map,
               double out4[4]) {
  if (pol_lo < 1 || pol_hi > h->nmaps || pol_lo > pol_hi) fail(DANG_GPU_EINVAL, "bad pol_type range %d..%d", pol_lo, pol_hi);
  if (!sky && !res && !chi_map && h->stat_cache) {
    if (h->chisq_valid && h->chisq_version == h->version && h->chisq_lo == pol_lo && h->chisq_hi == pol_hi) {
      // the full-sky draw that produced the current state left its chi-square behind
      for (int k = 0; k < 3; k++) out4[k] = h->chisq_vals[k];
      out4[3] = (double)unmasked_count(h);
      return;
    }
    if (h->defer_scalars && h->pend_draw && h->pend_draw_chisq && h->pend_draw_version == h->version &&
        h->pend_plane[0] + 1 == pol_lo && h->pend_plane[h->pend_S - 1] + 1 == pol_hi && pol_hi - pol_lo + 1 == h->pend_S) {
      // deferred scalars: the pending full-sky draw left this chi-square in its MhScalars (dang_gpu_iteration_scalars)
      for (int k = 0; k < 3; k++) out4[k] = nan("");
      out4[3] = (double)unmasked_count(h);
      h->pend_chisq_draw = true;
      return;
    }
    if (chisq_from_statistics(h, pol_lo, pol_hi, out4)) return;
  }
  ModelView mv = model_view(h);
  ChisqView cv;
  cv.k_lo = pol_lo - 1;
  cv.k_hi = pol_hi - 1;
  cv.sky = sky;
  cv.res = res;
  cv.chi_map = chi_map;
  const int grid = grid_for(h, h->P, DG_THREADS, 4);
  const bool maps = sky || res;
  const double nk = maps ? h->nmaps : (pol_hi - pol_lo + 1);
  double bytes = bytes_w((double)h->P * nk * (2.0 * h->nbands + h->ncomp * 2.0));
  if (maps) bytes += bytes_w((double)h->P * h->nmaps * h->nbands * ((sky ? 1 : 0) + (res ? 1 : 0)));
  KTimer kt(h, maps ? DANG_K_SKYMODEL : DANG_K_CHISQ, bytes);
  bool uni = !maps && !chi_map && h->ncomp <= 4;
  unsigned nu_mask = 0;
  for (int k = cv.k_lo; k <= cv.k_hi && uni; k++)
    for (int c = 0; c < h->ncomp; c++)
      if (!comp_uniform(h, c, k)) nu_mask |= 1u << c;
  const size_t dsm = (size_t)__builtin_popcount(nu_mask) * h->nbands * 2 * DG_THREADS * sizeof(double);
  if (dsm > 160 * 1024) uni = false;
  if (uni) {
#define LAUNCH_CHISQ_UNI(NC)                                                                                  \
    {                                                                                                         \
      if (dsm > 48 * 1024) CK(cudaFuncSetAttribute(chisq_uni_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm)); \
      const int g2 = occ_grid(h, chisq_uni_kernel<NC>, h->Ppad / 2, DG_THREADS, dsm);                         \
      chisq_uni_kernel<NC><<<g2, DG_THREADS, dsm, h->stream>>>(mv, cv, h->partials, h->tickets, h->sums_local, nu_mask); \
    }
    if (h->ncomp <= 1) LAUNCH_CHISQ_UNI(1)
    else if (h->ncomp == 2) LAUNCH_CHISQ_UNI(2)
    else if (h->ncomp == 3) LAUNCH_CHISQ_UNI(3)
    else LAUNCH_CHISQ_UNI(4)
#undef LAUNCH_CHISQ_UNI
  }
  else if (h->ncomp <= 1) launch_chisq<1>(h, mv, cv, grid);
  else if (h->ncomp == 2) launch_chisq<2>(h, mv, cv, grid);
  else if (h->ncomp == 3) launch_chisq<3>(h, mv, cv, grid);
  else if (h->ncomp == 4) launch_chisq<4>(h, mv, cv, grid);
  else launch_chisq<DG_MAX_COMPS>(h, mv, cv, grid);
  kt.done();
  gather(h, 4);
  double *hp = (double *)h->pinned;
  readback(h, hp, h->gathered, (size_t)h->nranks * 4 * sizeof(double));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 4; i++) {
    out4[i] = 0.0;
    for (int g = 0; g < h->nranks; g++) out4[i] += hp[g * 4 + i];
  }
  h->n_unmasked = (int64_t)(out4[3] + 0.5);
}
