// host_udgrade.cu -- HEALPix resolution changes of the low-resolution sampling branch as standalone device ops
// (SURVEY 8f-2): udgrade_ring as HEALPix-F90 (module udgrade_nr, un-vendored) publishes it, and dang's wrappers
// udgrade_rms / udgrade_mask (src/dang_util_mod.f90:341-376), at the call sites src/dang_sample_mod.f90:204-217, 480.
// The sampler itself stays closed for sample_nside /= nside (the reference's own branch mixes resolutions).
#include "host.cuh"

namespace {
__constant__ int c_jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
__constant__ int c_jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

__device__ __forceinline__ long long compress_bits(long long v) {  // bit 2i -> bit i
  v &= 0x5555555555555555LL;
  v = (v | (v >> 1)) & 0x3333333333333333LL;
  v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0fLL;
  v = (v | (v >> 4)) & 0x00ff00ff00ff00ffLL;
  v = (v | (v >> 8)) & 0x0000ffff0000ffffLL;
  v = (v | (v >> 16)) & 0x00000000ffffffffLL;
  return v;
}
__device__ __forceinline__ long long spread_bits(long long v) {  // bit i -> bit 2i
  v &= 0x00000000ffffffffLL;
  v = (v | (v << 16)) & 0x0000ffff0000ffffLL;
  v = (v | (v << 8)) & 0x00ff00ff00ff00ffLL;
  v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0fLL;
  v = (v | (v << 2)) & 0x3333333333333333LL;
  v = (v | (v << 1)) & 0x5555555555555555LL;
  return v;
}
__device__ __forceinline__ long long isqrt_ll(long long v) {
  long long r = (long long)sqrt((double)v + 0.5);
  while (r * r > v) r--;
  while ((r + 1) * (r + 1) <= v) r++;
  return r;
}
// healpix_base nest2xyf + xyf2ring
__device__ long long nest2ring(long long nside, long long pix) {
  const long long npface = nside * nside, npix = 12 * npface, ncap = 2 * nside * (nside - 1), nl4 = 4 * nside;
  const int face = (int)(pix / npface);
  const long long p = pix & (npface - 1);
  const long long ix = compress_bits(p), iy = compress_bits(p >> 1);
  const long long jr = c_jrll[face] * nside - ix - iy - 1;
  long long nr, n_before, kshift;
  if (jr < nside) {
    nr = jr;
    n_before = 2 * nr * (nr - 1);
    kshift = 0;
  } else if (jr > 3 * nside) {
    nr = nl4 - jr;
    n_before = npix - 2 * (nr + 1) * nr;
    kshift = 0;
  } else {
    nr = nside;
    n_before = ncap + (jr - nside) * nl4;
    kshift = (jr - nside) & 1;
  }
  long long jp = (c_jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  else if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}
// healpix_base ring2xyf + xyf2nest
__device__ long long ring2nest(long long nside, long long pix) {
  const long long npface = nside * nside, npix = 12 * npface, ncap = 2 * nside * (nside - 1), nl2 = 2 * nside;
  long long iring, iphi, kshift, nr;
  int face;
  if (pix < ncap) {
    iring = (1 + isqrt_ll(1 + 2 * pix)) >> 1;
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0;
    nr = iring;
    face = (int)((iphi - 1) / nr);
  } else if (pix < npix - ncap) {
    const long long ip = pix - ncap;
    iring = ip / (4 * nside) + nside;
    iphi = ip % (4 * nside) + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    const long long ire = iring - nside + 1, irm = nl2 + 2 - ire;
    const long long ifm = (iphi - ire / 2 + nside - 1) / nside, ifp = (iphi - irm / 2 + nside - 1) / nside;
    if (ifp == ifm) face = (ifp == 4) ? 4 : (int)ifp + 4;
    else if (ifp < ifm) face = (int)ifp;
    else face = (int)ifm + 8;
  } else {
    const long long ip = npix - pix;
    iring = (1 + isqrt_ll(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0;
    nr = iring;
    iring = 2 * nl2 - iring;
    face = 8 + (int)((iphi - 1) / nr);
  }
  const long long irt = iring - c_jrll[face] * nside + 1;
  long long ipt = 2 * iphi - c_jpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  const long long ix = (ipt - irt) >> 1, iy = (-(ipt + irt)) >> 1;
  return (long long)face * npface + spread_bits(ix) + (spread_bits(iy) << 1);
}

// One thread per OUTPUT pixel (RING order) and plane.  kind 0: udgrade_ring; 1: udgrade_rms (squares in, sqrt(mean) *
// nside_out / nside_in out); 2: udgrade_mask (threshold when degrading).  Children are summed in NESTED order, as
// sub_udgrade_nest does; bad pixels (-1.6375e30) are skipped, all-bad parents get the bad value.
__global__ void __launch_bounds__(DG_THREADS)
udgrade_kernel(const double *__restrict__ in, long long nside_in, double *__restrict__ out, long long nside_out, int nmaps,
               int kind, double threshold) {
  const long long npix_in = 12 * nside_in * nside_in, npix_out = 12 * nside_out * nside_out;
  const double bad = -1.6375e30;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < npix_out * nmaps; t += stride) {
    const int k = (int)(t / npix_out);
    const long long po = t - (long long)k * npix_out;
    const double *mi = in + (size_t)k * npix_in;
    const long long id = ring2nest(nside_out, po);
    double v;
    if (nside_out < nside_in) {
      const long long npratio = npix_in / npix_out;
      double total = 0.0;
      long long nobs = 0;
      for (long long ip = 0; ip < npratio; ip++) {
        double x = mi[nest2ring(nside_in, id * npratio + ip)];
        if (kind == 1) x = x * x;  // data_buffer = data_in*data_in (:350); a bad rms squares away from the bad value
        if (x != bad) {
          total = total + x;
          nobs++;
        }
      }
      v = nobs ? total / (double)nobs : bad;
    } else {
      const long long npratio = npix_out / npix_in;
      v = mi[nest2ring(nside_in, id / npratio)];
      if (kind == 1) v = v * v;
    }
    if (kind == 1) v = sqrt(v) * ((double)nside_out * 1.0 / (double)nside_in);  // :354
    if (kind == 2 && nside_in > nside_out) v = (v < threshold) ? 0.0 : 1.0;      // :367-373
    out[(size_t)k * npix_out + po] = v;
  }
}
}  // namespace

extern "C" int dang_gpu_udgrade(dang_gpu_t *h, int kind, const double *in, int nside_in, double *out, int nside_out, int nmaps,
                                double threshold) {
  if (!h) return DANG_GPU_EINVAL;
  double *d_in = nullptr, *d_out = nullptr;
  try {
    set_device(h);
    auto pow2 = [](int n) { return n > 0 && (n & (n - 1)) == 0; };
    if (kind < 0 || kind > 2 || !in || !out || nmaps < 1 || !pow2(nside_in) || !pow2(nside_out) || nside_in > 8192 || nside_out > 8192)
      fail(DANG_GPU_EINVAL, "udgrade: kind %d, nside %d -> %d, nmaps %d", kind, nside_in, nside_out, nmaps);
    const size_t n_in = (size_t)12 * nside_in * nside_in * nmaps, n_out = (size_t)12 * nside_out * nside_out * nmaps;
    CK(cudaMalloc(&d_in, n_in * sizeof(double)));
    CK(cudaMalloc(&d_out, n_out * sizeof(double)));
    CK(cudaMemcpyAsync(d_in, in, n_in * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int grid = grid_for(h, (int64_t)n_out, DG_THREADS, 8);
    udgrade_kernel<<<grid, DG_THREADS, 0, h->stream>>>(d_in, nside_in, d_out, nside_out, nmaps, kind, threshold);
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaMemcpyAsync(out, d_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_in);
    cudaFree(d_out);
    return DANG_GPU_OK;
  } catch (const DgError &e) {
    h->err = e.what();
    cudaFree(d_in);
    cudaFree(d_out);
    return e.code;
  }
}
