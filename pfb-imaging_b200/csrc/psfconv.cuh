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
  // fill / drain loops keep several independent global accesses in flight (they were latency-bound, like the
  // plane-transform kernels before the same change)
  constexpr int U = 4;
  const int64_t row = (int64_t)i * ct.ny;
  for (int j0 = tid; j0 < ct.ny; j0 += U * nthr) {
    T xv[U], bv[U];
    int pv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * nthr;
      if (j < ct.ny) {
        xv[u] = x[row + j];
        bv[u] = beam ? beam[row + j] : (T)1;
        pv[u] = ct.pos_v[j];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (j0 + u * nthr < ct.ny) s[fft_pad<T>(pv[u])] = {xv[u] * bv[u], (T)0};
  }
  __syncthreads();
  fft_dit<T, 1>(s, (const cx2<T>*)ct.tw_v, ct.dv, tid, nthr);
  cx2<T>* dst = tmp + (int64_t)i * ct.nyp;
  // x (and beam) are real and khat is the transform of a real kernel, so column nyp - l of every later stage is the
  // complex conjugate of column l: only l <= nyp/2 (rounded up to a column block) is written / transformed
  const int nkeep = min(ct.nyp, ((ct.nyp / 2 + 1 + 3) / 4) * 4);
#pragma unroll 4
  for (int n = tid; n < nkeep; n += nthr) dst[n] = s[fft_pad<T>(n)];
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
  constexpr int U = 8;
  for (int w = ct.nx * C + tid; w < ct.nxp * C; w += nthr) s[fft_pad<T>(w)] = {(T)0, (T)0};
  for (int w0 = tid; w0 < ct.nx * C; w0 += U * nthr) {
    cx2<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int w = w0 + u * nthr;
      if (w < ct.nx * C) v[u] = g[(int64_t)(w / C) * ct.nyp + (w % C)];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int w = w0 + u * nthr;
      if (w < ct.nx * C) s[fft_pad<T>(w)] = v[u];
    }
  }
  __syncthreads();
  fft_dif<T, C>(s, (const cx2<T>*)ct.tw_u, ct.du, tid, nthr);
  for (int w0 = tid; w0 < ct.nxp * C; w0 += U * nthr) {
    cx2<T> kv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int w = w0 + u * nthr;
      if (w < ct.nxp * C) kv[u] = khat[(int64_t)ct.rev_u[w / C] * ct.nyp + b0 + (w % C)];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int w = w0 + u * nthr;
      if (w < ct.nxp * C) {
        cx2<T> v = cmul(s[fft_pad<T>(w)], kv[u]);
        v.y = -v.y;  // conj: the inverse transform is conj o forward o conj
        s[fft_pad<T>(w)] = v;
      }
    }
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
  const int nh = ct.nyp / 2;
#pragma unroll 4
  for (int n = tid; n < ct.nyp; n += nthr) {
    // columns above nyp/2 are the conjugates of their mirrors (never computed); the inverse transform itself is
    // conj o forward o conj, so the mirrored half is loaded as is and the computed half conjugated
    cx2<T> v = src[n <= nh ? n : ct.nyp - n];
    if (n <= nh) v.y = -v.y;
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

// Hermitian part of a full spectrum: Re(IFFT(FFT(x) k)) for real x only sees (k[a,b] + conj k[-a,-b]) / 2, which is
// what the half-column evaluation above assumes.  out may not alias in.
template <typename T>
__global__ void k_hermitize(int nxp, int nyp, const cx2<T>* __restrict__ in, cx2<T>* __restrict__ out) {
  int l = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (l >= nyp) return;
  const int km = k == 0 ? 0 : nxp - k, lm = l == 0 ? 0 : nyp - l;
  const cx2<T> a = in[(int64_t)k * nyp + l], b = in[(int64_t)km * nyp + lm];
  out[(int64_t)k * nyp + l] = {(T)0.5 * (a.x + b.x), (T)0.5 * (a.y - b.y)};
}
