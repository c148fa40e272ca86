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
  const double* nutab;       // (nx,ny) fp64: n - 1 + nshift per pixel (w-screen phase = w_p * nutab, up to 1e3 turns)
  // cells any bound sample can touch: rows [a_lo, a_lo+a_len) and columns [b_lo, b_lo+b_len), circular
  int a_lo, a_len, b_lo, b_len;
};

__device__ __forceinline__ bool in_window(int n, int lo, int len, int size) {
  int rel = n - lo;
  if (rel < 0) rel += size;
  return rel < len;
}

#define ROWS_MAX_THREADS 256


// --------------------------------------------------------------------------- degrid direction
template <typename T>
__global__ void __launch_bounds__(ROWS_MAX_THREADS, (sizeof(T) == 4 ? 3 : 1))
k_rows_fwd(GParams p, FusedTabs ft, const T* __restrict__ x, const T* __restrict__ beam, const T* __restrict__ corr,
           typename cplx_of<T>::type* __restrict__ grid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x, i = blockIdx.x, q = blockIdx.y;
  const int nv = p.nv, hy = p.ny / 2;
  const int ip = i - p.nx / 2;
  const int a = ip < 0 ? ip + p.nu : ip;
  for (int n = tid; n < nv; n += nthr) s[fft_pad<T>(n)] = {(T)0, (T)0};
  __syncthreads();
  const double wq = p.w0 + q * p.dw;
  for (int j = tid; j < p.ny; j += nthr) {
    const int64_t pix = (int64_t)i * p.ny + j;
    T val = x[pix] * corr[pix];
    if (beam) val *= beam[pix];
    cx2<T> v = {val, (T)0};
    if (p.do_wgridding && val != (T)0) {
      T c, sn;
      cis_turns(wq * ft.nutab[pix], c, sn);
      v = {val * c, val * sn};
    }
    const int jp = j - hy;
    s[fft_pad<T>(ft.pos_v[jp < 0 ? jp + nv : jp])] = v;
  }
  __syncthreads();
  fft_dit<T, 1>(s, (const cx2<T>*)ft.tw_v, ft.dv, tid, nthr);
  cx2<T>* dst = reinterpret_cast<cx2<T>*>(grid) + ((int64_t)q * p.nu + a) * nv;
  for (int n = tid; n < nv; n += nthr)
    if (in_window(n, ft.b_lo, ft.b_len, nv)) dst[n] = s[fft_pad<T>(n)];
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
    s[fft_pad<T>(w)] = band ? g[(int64_t)a * p.nv + c] : cx2<T>{(T)0, (T)0};
  }
  __syncthreads();
  fft_dif<T, C>(s, (const cx2<T>*)ft.tw_u, ft.du, tid, nthr);
  for (int w = tid; w < nu * C; w += nthr) {
    const int pos = w / C, c = w - pos * C;
    const int k = ft.rev_u[pos];
    if (in_window(k, ft.a_lo, ft.a_len, nu)) g[(int64_t)k * p.nv + c] = s[fft_pad<T>(w)];
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
    s[fft_pad<T>(w)] = v;
  }
  __syncthreads();
  fft_dif<T, C>(s, (const cx2<T>*)ft.tw_u, ft.du, tid, nthr);
  for (int w = tid; w < nu * C; w += nthr) {
    const int pos = w / C, c = w - pos * C;
    const int k = ft.rev_u[pos];
    if (k < hx || k >= nu - hx) {
      cx2<T> v = s[fft_pad<T>(w)];
      v.y = -v.y;
      g[(int64_t)k * p.nv + c] = v;
    }
  }
}

// One CTA per (plane, image row): read the row (active window only), inverse FFT along v, apply the
// conjugate w-screen to the ny kept outputs and add their real parts to the fp64 accumulation image
// (RED.F64; CTAs of one row are adjacent in the grid, so the 32 KB image row stays in L2).
template <typename T>
__global__ void __launch_bounds__(ROWS_MAX_THREADS, (sizeof(T) == 4 ? 3 : 1))
k_rows_inv(GParams p, FusedTabs ft, const typename cplx_of<T>::type* __restrict__ grid, double* __restrict__ accimg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x, q = blockIdx.x, i = blockIdx.y;
  const int nv = p.nv, hy = p.ny / 2;
  const int ip = i - p.nx / 2;
  const int a = ip < 0 ? ip + p.nu : ip;
  const cx2<T>* src = reinterpret_cast<const cx2<T>*>(grid) + ((int64_t)q * p.nu + a) * nv;
  for (int n = tid; n < nv; n += nthr) {
    cx2<T> v = {(T)0, (T)0};  // columns outside the window are known to be zero
    if (in_window(n, ft.b_lo, ft.b_len, nv)) { v = src[n]; v.y = -v.y; }
    s[fft_pad<T>(n)] = v;
  }
  __syncthreads();
  fft_dif<T, 1>(s, (const cx2<T>*)ft.tw_v, ft.dv, tid, nthr);
  const double wq = p.w0 + q * p.dw;
  double* dst = accimg + (int64_t)i * p.ny;
  for (int j = tid; j < p.ny; j += nthr) {
    const int jp = j - hy;
    const cx2<T> v = s[fft_pad<T>(ft.pos_v[jp < 0 ? jp + nv : jp])];  // conj(v) is the inverse transform
    double r;
    if (p.do_wgridding) {
      T c, sn;
      cis_turns(wq * ft.nutab[(int64_t)i * p.ny + j], c, sn);
      r = (double)(v.x * c - v.y * sn);  // Re( conj(v) e^{-i theta} )
    } else {
      r = (double)v.x;
    }
    atomicAdd(dst + j, r);
  }
}

// out = acc * corr [* beam] * inv_wsum [+ eta * xin]
template <typename T>
__global__ void k_finish_image(int64_t npix, const double* __restrict__ acc, const T* __restrict__ corr,
                               const T* __restrict__ beam, const T* __restrict__ xin, double inv_wsum, double eta,
                               T* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= npix) return;
  double r = acc[k] * (double)corr[k];
  if (beam) r *= (double)beam[k];
  r *= inv_wsum;
  if (xin) r += eta * (double)xin[k];
  out[k] = (T)r;
}

__global__ void k_nu_table(GParams p, double* __restrict__ nutab) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= p.ny) return;
  nutab[(int64_t)i * p.ny + j] = pixel_nm1(p, i, j) + p.nshift;
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
