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

// Split form: deviates -> global, fp64 state -> fp32 scratch, then the register-light chain kernel.
void launch_perpixel_split(dang_gpu *h, const ModelView &mv, MhView &mh, int bpl, int mode, int64_t work) {
  const int bplr = bpl <= 2 ? 2 : bpl <= 3 ? 3 : bpl <= 5 ? 5 : 8;
  const size_t need = (size_t)bplr * h->P * DG_MH_LANES;
  if (h->k5_len < need) {
    if (h->k5_st4) CK(cudaFree(h->k5_st4));
    if (h->k5_kj) CK(cudaFree(h->k5_kj));
    h->k5_st4 = nullptr;
    h->k5_kj = nullptr;
    CK(cudaMalloc(&h->k5_st4, need * sizeof(float4)));
    CK(cudaMalloc(&h->k5_kj, need * sizeof(float)));
    h->k5_len = need;
  }
  if (!mh.z) {  // device RNG: the whole chain's deviates in one parallel pass
    const size_t n = (size_t)mh.nsample * h->P;
    ensure_zu(h, n > 0 ? n : 1);
    KTimer kt(h, DANG_K_SCALAR, 0);
    k5_rng_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(mv, mh, h->zbuf, h->ubuf);
    kt.done();
    mh.z = h->zbuf;
    mh.u = h->ubuf;
  }
  float4 *st4 = (float4 *)h->k5_st4;
  float *kj = h->k5_kj;
#define LAUNCH_SPLIT(BPL, MODE)                                                                            \
  {                                                                                                        \
    const int g1 = occ_grid(h, k5_state_kernel<BPL, MODE>, work, DG_MH_THREADS);                           \
    k5_state_kernel<BPL, MODE><<<g1, DG_MH_THREADS, 0, h->stream>>>(mv, mh, st4, kj);                      \
    CK(cudaGetLastError());                                                                                \
    const int g2 = occ_grid(h, k5_chain_kernel<BPL, MODE>, work, DG_MH_THREADS);                           \
    k5_chain_kernel<BPL, MODE><<<g2, DG_MH_THREADS, 0, h->stream>>>(mv, mh, st4, kj, h->partials, h->tickets, h->sums_local); \
  }
#define LAUNCH_SPLIT_MODE(BPL)                                                 \
  {                                                                            \
    if (mode == MH_SED_POWERLAW) LAUNCH_SPLIT(BPL, MH_SED_POWERLAW)            \
    else if (mode == MH_SED_MBB_BETA) LAUNCH_SPLIT(BPL, MH_SED_MBB_BETA)       \
    else if (mode == MH_SED_MBB_T) LAUNCH_SPLIT(BPL, MH_SED_MBB_T)             \
    else fail(DANG_GPU_EINVAL, "no screened kernel for SED mode %d", mode);    \
  }
  if (bpl <= 2) LAUNCH_SPLIT_MODE(2)
  else if (bpl <= 3) LAUNCH_SPLIT_MODE(3)
  else if (bpl <= 5) LAUNCH_SPLIT_MODE(5)
  else LAUNCH_SPLIT_MODE(8)
#undef LAUNCH_SPLIT_MODE
#undef LAUNCH_SPLIT
}
