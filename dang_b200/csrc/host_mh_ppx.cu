// host_mh_ppx.cu -- launches of the one-thread-per-pixel screened Metropolis kernel (kernels_mh_pix.cuh).
#include "host.cuh"
#include "kernels_mh_pix.cuh"

void launch_perpixel_pix(dang_gpu *h, const ModelView &mv, const MhView &mh, int mode) {
  if (h->nbands > DG_K5P_MAXB) fail(DANG_GPU_EUNSUPPORTED, "screened per-pixel kernel: %d bands (max %d)", h->nbands, DG_K5P_MAXB);
  K5Bands kb;
  const CompHost &cc = h->comp[mh.ic];
  const double hk = DG_H / DG_KB;
  kb.cref = (float)(hk * cc.nu_ref);
  for (int j = 0; j < DG_K5P_MAXB; j++) {
    if (mode == MH_SED_MBB_T) kb.cf[j] = j < h->nbands ? (float)(hk * h->band[j].nu_c) : kb.cref;  // (padding: rho = 0)
    else if (j < h->nbands) {
      double hi, lo;
      dd_log_ratio(h->band[j].nu_c, cc.nu_ref, hi, lo);  // (the table's lnr_hi)
      kb.cf[j] = (float)hi;
    } else kb.cf[j] = 0.0f;
  }
  kb.cfmax = mode == MH_SED_MBB_T ? fabsf(kb.cref) : 0.0f;
  for (int j = 0; j < h->nbands; j++) kb.cfmax = fmaxf(kb.cfmax, fabsf(kb.cf[j]));
  kb.cfmax *= 1.000001f;
#define LAUNCH_PIX(NB, MODE)                                                                                  \
  {                                                                                                           \
    const size_t smem = (size_t)NB * DG_MH_THREADS * (sizeof(float4) + (MODE == MH_SED_MBB_T ? sizeof(float) : 0)); \
    CK(cudaFuncSetAttribute(mh_perpixel_pix_kernel<NB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    const int grid = occ_grid(h, mh_perpixel_pix_kernel<NB, MODE>, h->P, DG_MH_THREADS, smem);                \
    mh_perpixel_pix_kernel<NB, MODE><<<grid, DG_MH_THREADS, smem, h->stream>>>(mv, mh, kb, h->partials, h->tickets, h->sums_local); \
  }
#define LAUNCH_PIX_MODE(NB)                                                  \
  {                                                                          \
    if (mode == MH_SED_POWERLAW) LAUNCH_PIX(NB, MH_SED_POWERLAW)             \
    else if (mode == MH_SED_MBB_BETA) LAUNCH_PIX(NB, MH_SED_MBB_BETA)        \
    else if (mode == MH_SED_MBB_T) LAUNCH_PIX(NB, MH_SED_MBB_T)              \
    else fail(DANG_GPU_EINVAL, "no screened kernel for SED mode %d", mode);  \
  }
  if (h->nbands <= 8) LAUNCH_PIX_MODE(8)
  else if (h->nbands <= 12) LAUNCH_PIX_MODE(12)
  else if (h->nbands <= 20) LAUNCH_PIX_MODE(20)
  else LAUNCH_PIX_MODE(32)
#undef LAUNCH_PIX_MODE
#undef LAUNCH_PIX
}
