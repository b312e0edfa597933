// host_mh_pp.cu -- per-pixel branch of sample_index_mh, src/dang_sample_mod.f90:332-481.
#include "host.cuh"
#include "kernels_mh.cuh"

void sample_perpixel(dang_gpu *h, MhView &mh, const double *z, const double *u, uint64_t seed,
                     double *accept) {
  ModelView mv = model_view(h);
  const size_t n = (size_t)mh.nsample * h->P;
  mh.seed = seed;
  if (z) {
    ensure_zu(h, n > 0 ? n : 1);
    // host [l][npix] -> device [l][P]
    CK(cudaMemcpy2DAsync(h->zbuf, h->P * sizeof(double), z + h->lo, h->npix * sizeof(double),
                         h->P * sizeof(double), mh.nsample, cudaMemcpyHostToDevice, h->stream));
    mh.z = h->zbuf;
    if (u) {
      CK(cudaMemcpy2DAsync(h->ubuf, h->P * sizeof(double), u + h->lo, h->npix * sizeof(double),
                           h->P * sizeof(double), mh.nsample, cudaMemcpyHostToDevice, h->stream));
      mh.u = h->ubuf;
    } else if (mh.ml_mode == DANG_ML_SAMPLE) {
      fail(DANG_GPU_EINVAL, "z injected without u");
    }
  }
  bool pp_fast_ran = false;
  h->dec_mode = 0;
  if (h->record) {
    ensure_decisions(h, n > 0 ? n : 1);
    CK(cudaMemsetAsync(h->decisions, 3, n, h->stream));
    CK(cudaMemsetAsync(h->lnl_trace, 0xff, n * sizeof(double), h->stream));  // NaN pattern
    mh.decisions = h->decisions;
    mh.lnl_trace = h->lnl_trace;
    h->dec_mode = 2;
    h->dec_nsample = mh.nsample;
  }
  const double n_el = (double)mh.S * h->P;
  const double kbytes = bytes_w(n_el * (2.0 * h->nbands + h->ncomp + 1) + (double)h->P * 4);
  // the lane-cooperative kernel covers the chisq likelihood with uniform / Gaussian prior;
  // marginal lnL, 'prior' draws and the Jeffreys prior run on the strict kernel
  const bool strict = h->perpixel_serial || mh.lnl_type != DANG_LNL_CHISQ || mh.prior_type == DANG_PRIOR_JEFFREYS;
  if (strict) {
    const size_t smem = (size_t)(2 * h->nbands * mh.S + 2 * h->nbands) * DG_MH_THREADS * sizeof(double);
    if (smem > 200 * 1024) fail(DANG_GPU_EUNSUPPORTED, "per-pixel chain needs %zu B of shared memory", smem);
    CK(cudaFuncSetAttribute(mh_perpixel_serial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = occ_grid(h, mh_perpixel_serial_kernel, h->P, DG_MH_THREADS, smem);
    KTimer kt(h, DANG_K_MH_PERPIXEL, kbytes);
    mh_perpixel_serial_kernel<<<grid, DG_MH_THREADS, smem, h->stream>>>(mv, mh, h->partials, h->tickets, h->sums_local);
    kt.done();
  } else {
    const size_t smem = (size_t)2 * (DG_MH_THREADS / DG_MH_LANES) * (mh.nsample > 0 ? mh.nsample : 1) * sizeof(double);
    if (smem > 160 * 1024) fail(DANG_GPU_EUNSUPPORTED, "nsample = %d needs %zu B of shared memory", mh.nsample, smem);
    const int bpl = (h->nbands + DG_MH_LANES - 1) / DG_MH_LANES;
    const int64_t work = h->P * DG_MH_LANES;
    bool any_bp = false;
    for (int j = 0; j < h->nbands; j++) any_bp = any_bp || h->band[j].n != 0;
    int mode = MH_SED_GENERIC;
    size_t bp_smem = 0;
    if (!any_bp) {
      if (h->comp[mh.ic].type == DANG_COMP_POWERLAW) mode = MH_SED_POWERLAW;
      else if (h->comp[mh.ic].type == DANG_COMP_MBB) mode = mh.nind == 0 ? MH_SED_MBB_BETA : MH_SED_MBB_T;
    } else if (h->pp_bp_series) {  // tabulated bandpasses: moment series about the chain's first point (kernels_mh.cuh)
      if (h->comp[mh.ic].type == DANG_COMP_POWERLAW) mode = MH_SED_BP_POWERLAW;
      else if (h->comp[mh.ic].type == DANG_COMP_MBB && mh.nind == 0) mode = MH_SED_BP_MBB_BETA;
      if (mode != MH_SED_GENERIC) {
        const int bplr = bpl <= 2 ? 2 : bpl <= 3 ? 3 : bpl <= 5 ? 5 : 8;
        bp_smem = (size_t)bplr * (DG_MH_KM + 2) * DG_MH_THREADS * sizeof(double);
      }
    }
    // certified fp32 screening (kernels_mh_fast.cuh) for delta-band power-law / mbb draws; everything
    // else (tabulated bandpasses, other SED types) evaluates every proposal in fp64
    const bool fast = h->pp_fast && mode != MH_SED_GENERIC && mode < MH_SED_BP_POWERLAW;
    KTimer kt(h, DANG_K_MH_PERPIXEL, kbytes);
    if (fast && h->pp_pix && h->nbands <= 32) {
      launch_perpixel_pix(h, mv, mh, mode);
    } else if (fast && h->pp_split && !h->record) {
      launch_perpixel_split(h, mv, mh, bpl, mode, work);
    } else if (fast) {
      const int bplr = bpl <= 2 ? 2 : bpl <= 3 ? 3 : bpl <= 5 ? 5 : 8;
      launch_perpixel_fast(h, mv, mh, bpl, mode, work, smem + (size_t)4 * bplr * DG_MH_THREADS * sizeof(double));
    }
    else if (bpl <= 2) launch_perpixel_fp64_2(h, mv, mh, mode, work, smem + bp_smem);
    else if (bpl <= 3) launch_perpixel_fp64_3(h, mv, mh, mode, work, smem + bp_smem);
    else if (bpl <= 5) launch_perpixel_fp64_5(h, mv, mh, mode, work, smem + bp_smem);
    else launch_perpixel_fp64_8(h, mv, mh, mode, work, smem + bp_smem);
    kt.done();
    pp_fast_ran = fast;
  }
  const int cnt = pp_fast_ran ? 4 : 1;
  gather(h, cnt);
  double *hp = (double *)h->pinned;
  readback(h, hp, h->gathered, (size_t)h->nranks * cnt * sizeof(double));
  CK(cudaStreamSynchronize(h->stream));
  double a = 0;
  h->pp_fallbacks = h->pp_violations = 0.0;
  for (int g = 0; g < h->nranks; g++) {
    a += hp[g * cnt];
    if (pp_fast_ran) {
      h->pp_fallbacks += hp[g * cnt + 1];
      h->pp_violations += hp[g * cnt + 2];
    }
  }
  if (accept) *accept = a;
}
