// Fused, pruned plane transforms (kernel 4 + the 2-D FFT in one): replaces
//   k_img2grid + cuFFT (degrid direction)   by  k_rows_fwd  + k_cols_fwd
//   cuFFT + k_grid2img (grid direction)     by  k_cols_inv  + k_rows_inv
//
// Per direction the plane stack is touched 2.3 times instead of 7 (SURVEY §8d counts 6):
//   rows_fwd : reads the image row (L2), builds pad + beam + correction + w-screen in shared memory,
//              FFT along v (DIT, natural output), writes ONLY the nx image rows of the plane
//   cols_fwd : reads ONLY those nx rows of a 32-byte wide column block, FFT along u, writes nu rows
//   cols_inv : reads nu rows of a column block, inverse FFT along u, writes ONLY the nx image rows
//   rows_inv : per image row, loops over the planes: reads the row, inverse FFT along v, applies the
//              conjugate w-screen to the ny kept outputs and accumulates them in fp64 registers;
//              one write of the image row with correction / beam / wsum / ridge fused
// Column blocks whose cells no bound visibility touches are skipped (cb_lo..cb_hi).
#pragma once
#include "common.cuh"
#include "fft.cuh"

struct FusedTabs {
  FftDesc du, dv;            // transforms along u (length nu) and v (length nv)
  const void* tw_u;          // cx2<T>[nu]
  const void* tw_v;
  const int* rev_u;          // pos -> k
  const int* rev_v;
  const int* pos_v;          // k -> pos (DIT input scatter)
  // cells any bound sample can touch: rows [a_lo, a_lo+a_len) and columns [b_lo, b_lo+b_len), circular
  int a_lo, a_len, b_lo, b_len;
};

__device__ __forceinline__ bool in_window(int n, int lo, int len, int size) {
  int rel = n - lo;
  if (rel < 0) rel += size;
  return rel < len;
}

#define ROWS_THREADS 256
#define RINV_THREADS 512
#define RINV_MAXPER 16     /* ceil(nv / RINV_THREADS) <= 16  (nv <= 8192) */

// --------------------------------------------------------------------------- degrid direction
template <typename T>
__global__ void __launch_bounds__(ROWS_THREADS)
k_rows_fwd(GParams p, FusedTabs ft, const T* __restrict__ x, const T* __restrict__ beam, const T* __restrict__ corr,
           typename cplx_of<T>::type* __restrict__ grid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, i = blockIdx.x, q = blockIdx.y;
  const int nv = p.nv, hy = p.ny / 2;
  const int ip = i - p.nx / 2;
  const int a = ip < 0 ? ip + p.nu : ip;
  for (int n = tid; n < nv; n += ROWS_THREADS) s[n] = {(T)0, (T)0};
  __syncthreads();
  const double wq = p.w0 + q * p.dw;
  for (int j = tid; j < p.ny; j += ROWS_THREADS) {
    const int64_t pix = (int64_t)i * p.ny + j;
    T val = x[pix] * corr[pix];
    if (beam) val *= beam[pix];
    cx2<T> v = {val, (T)0};
    if (p.do_wgridding && val != (T)0) {
      T c, sn;
      cis_turns(wq * (pixel_nm1(p, i, j) + p.nshift), c, sn);
      v = {val * c, val * sn};
    }
    const int jp = j - hy;
    s[ft.pos_v[jp < 0 ? jp + nv : jp]] = v;
  }
  __syncthreads();
  fft_dit<T, 1>(s, (const cx2<T>*)ft.tw_v, ft.dv, tid, ROWS_THREADS);
  cx2<T>* dst = reinterpret_cast<cx2<T>*>(grid) + ((int64_t)q * p.nu + a) * nv;
  for (int n = tid; n < nv; n += ROWS_THREADS)
    if (in_window(n, ft.b_lo, ft.b_len, nv)) dst[n] = s[n];
}

// column block of C columns starting at b0; rows of the image band only are read
template <typename T, int C>
__global__ void __launch_bounds__(512)
k_cols_fwd(GParams p, FusedTabs ft, typename cplx_of<T>::type* __restrict__ grid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b0 = (ft.b_lo + blockIdx.x * C) % p.nv, q = blockIdx.y;
  const int nu = p.nu, hx = p.nx / 2;
  cx2<T>* g = reinterpret_cast<cx2<T>*>(grid) + (int64_t)q * nu * p.nv + b0;
  for (int w = tid; w < nu * C; w += nthr) {
    const int a = w / C, c = w - a * C;
    const bool band = a < hx || a >= nu - hx;
    s[w] = band ? g[(int64_t)a * p.nv + c] : cx2<T>{(T)0, (T)0};
  }
  __syncthreads();
  fft_dif<T, C>(s, (const cx2<T>*)ft.tw_u, ft.du, tid, nthr);
  for (int w = tid; w < nu * C; w += nthr) {
    const int pos = w / C, c = w - pos * C;
    const int k = ft.rev_u[pos];
    if (in_window(k, ft.a_lo, ft.a_len, nu)) g[(int64_t)k * p.nv + c] = s[w];
  }
}

// --------------------------------------------------------------------------- grid direction
template <typename T, int C>
__global__ void __launch_bounds__(512)
k_cols_inv(GParams p, FusedTabs ft, typename cplx_of<T>::type* __restrict__ grid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b0 = (ft.b_lo + blockIdx.x * C) % p.nv, q = blockIdx.y;
  const int nu = p.nu, hx = p.nx / 2;
  cx2<T>* g = reinterpret_cast<cx2<T>*>(grid) + (int64_t)q * nu * p.nv + b0;
  for (int w = tid; w < nu * C; w += nthr) {
    const int a = w / C, c = w - a * C;
    cx2<T> v = {(T)0, (T)0};
    if (in_window(a, ft.a_lo, ft.a_len, nu)) {  // rows outside the window are known to be zero
      v = g[(int64_t)a * p.nv + c];
      v.y = -v.y;  // inverse = conj o forward o conj
    }
    s[w] = v;
  }
  __syncthreads();
  fft_dif<T, C>(s, (const cx2<T>*)ft.tw_u, ft.du, tid, nthr);
  for (int w = tid; w < nu * C; w += nthr) {
    const int pos = w / C, c = w - pos * C;
    const int k = ft.rev_u[pos];
    if (k < hx || k >= nu - hx) {
      cx2<T> v = s[w];
      v.y = -v.y;
      g[(int64_t)k * p.nv + c] = v;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(RINV_THREADS)
k_rows_inv(GParams p, FusedTabs ft, const typename cplx_of<T>::type* __restrict__ grid,
           const T* __restrict__ corr, const T* __restrict__ beam, const T* __restrict__ xin, double inv_wsum,
           double eta, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, i = blockIdx.x;
  const int nv = p.nv, hy = p.ny / 2;
  const int ip = i - p.nx / 2;
  const int a = ip < 0 ? ip + p.nu : ip;
  // positions owned by this thread: pos = tid + m*RINV_THREADS; k = rev_v[pos]; kept if k is an image column
  int jcol[RINV_MAXPER];
  double nuv[RINV_MAXPER], acc[RINV_MAXPER];
#pragma unroll
  for (int m = 0; m < RINV_MAXPER; ++m) {
    const int pos = tid + m * RINV_THREADS;
    jcol[m] = -1;
    acc[m] = 0.0;
    nuv[m] = 0.0;
    if (pos < nv) {
      const int k = ft.rev_v[pos];
      if (k < hy) jcol[m] = k + hy;
      else if (k >= nv - hy) jcol[m] = k - nv + hy;
      if (jcol[m] >= 0 && p.do_wgridding) nuv[m] = pixel_nm1(p, i, jcol[m]) + p.nshift;
    }
  }
  for (int q = 0; q < p.nplanes; ++q) {
    const cx2<T>* src = reinterpret_cast<const cx2<T>*>(grid) + ((int64_t)q * p.nu + a) * nv;
    for (int n = tid; n < nv; n += RINV_THREADS) {
      cx2<T> v = {(T)0, (T)0};  // columns outside the window are known to be zero
      if (in_window(n, ft.b_lo, ft.b_len, nv)) { v = src[n]; v.y = -v.y; }
      s[n] = v;
    }
    __syncthreads();
    fft_dif<T, 1>(s, (const cx2<T>*)ft.tw_v, ft.dv, tid, RINV_THREADS);
    const double wq = p.w0 + q * p.dw;
#pragma unroll
    for (int m = 0; m < RINV_MAXPER; ++m) {
      if (jcol[m] >= 0) {
        const cx2<T> v = s[tid + m * RINV_THREADS];  // conj(v) is the inverse transform
        if (p.do_wgridding) {
          T c, sn;
          cis_turns(wq * nuv[m], c, sn);
          acc[m] += (double)(v.x * c - v.y * sn);  // Re( conj(v) * e^{-i theta} ) = v.x c - v.y s  (v.y is -Im)
        } else {
          acc[m] += (double)v.x;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < RINV_MAXPER; ++m) {
    if (jcol[m] >= 0) {
      const int64_t pix = (int64_t)i * p.ny + jcol[m];
      double r = acc[m] * (double)corr[pix];
      if (beam) r *= (double)beam[pix];
      r *= inv_wsum;
      if (xin) r += eta * (double)xin[pix];
      out[pix] = (T)r;
    }
  }
}

// mark the 32-cell groups of rows / columns touched by the bound samples (host derives the windows)
template <typename Rec>
__global__ void k_mark_cells(const Rec* __restrict__ recs, int64_t nact, int W, int nu, int nv,
                             int* __restrict__ uflag, int* __restrict__ vflag) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nact) return;
  int iu = recs[k].iu, iv = recs[k].iv;
  int iu2 = iu + W - 1, iv2 = iv + W - 1;
  if (iu2 >= nu) iu2 -= nu;
  if (iv2 >= nv) iv2 -= nv;
  uflag[iu >> 5] = 1; uflag[iu2 >> 5] = 1;
  vflag[iv >> 5] = 1; vflag[iv2 >> 5] = 1;
}

// zero the active window of every plane (instead of a memset of the whole stack)
template <typename C>
__global__ void k_zero_window(C* __restrict__ grid, int nu, int nv, int a_lo, int a_len, int b_lo, int b_len) {
  const int q = blockIdx.z;
  const int ra = blockIdx.y;
  int a = a_lo + ra;
  if (a >= nu) a -= nu;
  C* row = grid + ((int64_t)q * nu + a) * nv;
  C z; z.x = 0; z.y = 0;
  for (int rb = blockIdx.x * blockDim.x + threadIdx.x; rb < b_len; rb += gridDim.x * blockDim.x) {
    int b = b_lo + rb;
    if (b >= nv) b -= nv;
    row[b] = z;
  }
}
