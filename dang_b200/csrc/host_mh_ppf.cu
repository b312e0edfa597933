// host_mh_ppf.cu -- launches of the fp32-screened per-pixel Metropolis kernel (kernels_mh_fast.cuh).
#include "host.cuh"
#include "kernels_mh_fast.cuh"

void launch_perpixel_fast(dang_gpu *h, const ModelView &mv, const MhView &mh, int bpl, int mode, int64_t work,
                          size_t smem) {
#define LAUNCH_PPF(BPL, MODE)                                                                              \
  {                                                                                                        \
    CK(cudaFuncSetAttribute(mh_perpixel_fast_kernel<BPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    const int grid = occ_grid(h, mh_perpixel_fast_kernel<BPL, MODE>, work, DG_MH_THREADS, smem);           \
    mh_perpixel_fast_kernel<BPL, MODE><<<grid, DG_MH_THREADS, smem, h->stream>>>(mv, mh, h->partials, h->tickets, h->sums_local); \
  }
#define LAUNCH_PPF_MODE(BPL)                                                 \
  {                                                                          \
    if (mode == MH_SED_POWERLAW) LAUNCH_PPF(BPL, MH_SED_POWERLAW)            \
    else if (mode == MH_SED_MBB_BETA) LAUNCH_PPF(BPL, MH_SED_MBB_BETA)       \
    else if (mode == MH_SED_MBB_T) LAUNCH_PPF(BPL, MH_SED_MBB_T)             \
    else fail(DANG_GPU_EINVAL, "no screened kernel for SED mode %d", mode);  \
  }
  if (bpl <= 2) LAUNCH_PPF_MODE(2)
  else if (bpl <= 3) LAUNCH_PPF_MODE(3)
  else if (bpl <= 5) LAUNCH_PPF_MODE(5)
  else LAUNCH_PPF_MODE(8)
#undef LAUNCH_PPF_MODE
#undef LAUNCH_PPF
}
