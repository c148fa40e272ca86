// PSF-convolution Hessian on the device: the operator pfb-imaging iterates inside
// pcg / power method / primal-dual,
//     out = beam * crop( IFFT( FFT( pad(beam * x) ) * khat ) ) + eta * x
// (/root/reference/src/pfb_imaging/operators/hessian.py:103-143 hessian_psf_slice,
//  operators/psf.py:8-31 psf_convolve_slice), built from the same shared-memory FFT engine as the
// plane transforms: a row pass, ONE column kernel that transforms, multiplies and transforms back
// without leaving shared memory (DIF forward -> multiply in digit-reversed order -> DIT inverse), and
// a row pass back.  The zero-padded rows / columns are never read or written.
#pragma once
#include "fft.cuh"

struct ConvTabs {
  FftDesc du, dv;      // along u (length nxp) and v (length nyp)
  const void* tw_u;
  const void* tw_v;
  const int* rev_u;
  const int* rev_v;
  const int* pos_v;
  int nx, ny, nxp, nyp;
};

// image rows -> tmp (nx, nyp): FFT along v of the zero-padded row, image placed at columns 0..ny-1
template <typename T>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 ? 3 : 2))
k_conv_rows_fwd(ConvTabs ct, const T* __restrict__ x, const T* __restrict__ beam, cx2<T>* __restrict__ tmp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x, i = blockIdx.x;
  for (int n = tid; n < ct.nyp; n += nthr) s[fft_pad<T>(n)] = {(T)0, (T)0};
  __syncthreads();
  for (int j = tid; j < ct.ny; j += nthr) {
    const int64_t pix = (int64_t)i * ct.ny + j;
    T v = x[pix];
    if (beam) v *= beam[pix];
    s[fft_pad<T>(ct.pos_v[j])] = {v, (T)0};
  }
  __syncthreads();
  fft_dit<T, 1>(s, (const cx2<T>*)ct.tw_v, ct.dv, tid, nthr);
  cx2<T>* dst = tmp + (int64_t)i * ct.nyp;
  for (int n = tid; n < ct.nyp; n += nthr) dst[n] = s[fft_pad<T>(n)];
}

// column block: forward along u, multiply by khat[k][b], inverse along u, keep rows < nx
template <typename T, int C>
__global__ void __launch_bounds__(512)
k_conv_cols(ConvTabs ct, const cx2<T>* __restrict__ khat, cx2<T>* __restrict__ tmp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b0 = blockIdx.x * C;
  cx2<T>* g = tmp + b0;
  for (int w = tid; w < ct.nxp * C; w += nthr) {
    const int a = w / C, c = w - a * C;
    s[fft_pad<T>(w)] = a < ct.nx ? g[(int64_t)a * ct.nyp + c] : cx2<T>{(T)0, (T)0};
  }
  __syncthreads();
  fft_dif<T, C>(s, (const cx2<T>*)ct.tw_u, ct.du, tid, nthr);
  for (int w = tid; w < ct.nxp * C; w += nthr) {
    const int pos = w / C, c = w - pos * C;
    const cx2<T> k = khat[(int64_t)ct.rev_u[pos] * ct.nyp + b0 + c];
    cx2<T> v = cmul(s[fft_pad<T>(w)], k);
    v.y = -v.y;  // conj: the inverse transform is conj o forward o conj
    s[fft_pad<T>(w)] = v;
  }
  __syncthreads();
  fft_dit<T, C>(s, (const cx2<T>*)ct.tw_u, ct.du, tid, nthr);
  for (int w = tid; w < ct.nx * C; w += nthr) {
    const int a = w / C, c = w - a * C;
    cx2<T> v = s[fft_pad<T>(w)];
    v.y = -v.y;
    g[(int64_t)a * ct.nyp + c] = v;
  }
}

// tmp rows -> image: inverse FFT along v, crop, normalise, beam, ridge
template <typename T>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 ? 3 : 2))
k_conv_rows_inv(ConvTabs ct, const cx2<T>* __restrict__ tmp, const T* __restrict__ beam, const T* __restrict__ xin,
                double scale, double eta, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x, i = blockIdx.x;
  const cx2<T>* src = tmp + (int64_t)i * ct.nyp;
  for (int n = tid; n < ct.nyp; n += nthr) {
    cx2<T> v = src[n];
    v.y = -v.y;
    s[fft_pad<T>(n)] = v;
  }
  __syncthreads();
  fft_dif<T, 1>(s, (const cx2<T>*)ct.tw_v, ct.dv, tid, nthr);
  for (int j = tid; j < ct.ny; j += nthr) {
    const int64_t pix = (int64_t)i * ct.ny + j;
    double r = (double)s[fft_pad<T>(ct.pos_v[j])].x * scale;  // real part of conj(.) is the same
    if (beam) r *= (double)beam[pix];
    if (xin) r += eta * (double)xin[pix];
    out[pix] = (T)r;
  }
}

// khat_full[k][l] from the half spectrum (nxp, nyp/2+1) of a real kernel: Hermitian mirror
template <typename T>
__global__ void k_expand_half(int nxp, int nyp, const cx2<T>* __restrict__ half, cx2<T>* __restrict__ full) {
  int l = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (l >= nyp) return;
  const int nh = nyp / 2 + 1;
  cx2<T> v;
  if (l < nh) v = half[(int64_t)k * nh + l];
  else {
    const int km = k == 0 ? 0 : nxp - k;
    v = half[(int64_t)km * nh + (nyp - l)];
    v.y = -v.y;
  }
  full[(int64_t)k * nyp + l] = v;
}
