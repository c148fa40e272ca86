// Kernels of the device SARA backward step (include/pfbsara.h): tiled single-level 2-D wavelet analysis /
// synthesis (pywt 'zero' mode, filters up to 10 taps), the fused l21 dual update and the element-wise
// primal-dual pieces.  All of it is streaming work: each level reads its input once and writes its output
// once (both 1-D passes of a level happen in shared memory), so the bound is HBM bandwidth.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SW_KMAX 10
#define SW_TX 16   // coefficient rows per tile (per sub-band)
#define SW_TY 32   // coefficient columns per tile (per sub-band)

struct DwtFilt {
  double lo[SW_KMAX], hi[SW_KMAX];
  int K;
};

// strided batch: band b of the grid's z dimension works on  base + b * stride
struct BandPtr {
  int64_t in_stride, out_stride, aux_stride;
};

// ---------------------------------------------------------------------------
// analysis, one level:  in (nxin, nyin) -> block (2sx, 2sy) = [LL LH; HL HH] (x-first), LL also to `approx`
//   rows:  r[i, o] = sum_k h[k] in[i, 2o+1-k]        cols:  out[o, :] = sum_k h[k] r[2o+1-k, :]
// ---------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(256)
k_dwt_level(const T* __restrict__ in, int ld_in, int nxin, int nyin, T* __restrict__ out, int ld_out, int sx, int sy,
            T* __restrict__ approx, DwtFilt f, BandPtr bp) {
  constexpr int R = 2 * SW_TX + SW_KMAX - 2, CC = 2 * SW_TY + SW_KMAX - 2;
  __shared__ T sin_[R][CC + 1];
  __shared__ T stmp[R][2 * SW_TY];
  const int tid = threadIdx.x;
  const int ox0 = blockIdx.y * SW_TX, oy0 = blockIdx.x * SW_TY;
  T flo[K], fhi[K];  // filter taps in registers, loops fully unrolled (K is a template parameter)
#pragma unroll
  for (int k = 0; k < K; ++k) { flo[k] = (T)f.lo[k]; fhi[k] = (T)f.hi[k]; }
  in += (int64_t)blockIdx.z * bp.in_stride;
  out += (int64_t)blockIdx.z * bp.out_stride;
  if (approx) approx += (int64_t)blockIdx.z * bp.aux_stride;
  const int nr = 2 * SW_TX + K - 2, nc = 2 * SW_TY + K - 2;
  const int r0 = 2 * ox0 + 2 - K, c0 = 2 * oy0 + 2 - K;
  for (int idx = tid; idx < nr * nc; idx += 256) {
    const int r = idx / nc, c = idx - r * nc;
    const int gr = r0 + r, gc = c0 + c;
    T v = 0;
    if (gr >= 0 && gr < nxin && gc >= 0 && gc < nyin) v = in[(int64_t)gr * ld_in + gc];
    sin_[r][c] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < nr * SW_TY; idx += 256) {
    const int r = idx / SW_TY, o = idx - r * SW_TY;
    T lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const T v = sin_[r][2 * o + K - 1 - k];
      lo += flo[k] * v;
      hi += fhi[k] * v;
    }
    stmp[r][o] = lo;
    stmp[r][SW_TY + o] = hi;
  }
  __syncthreads();
  for (int idx = tid; idx < SW_TX * 2 * SW_TY; idx += 256) {
    const int o = idx / (2 * SW_TY), c = idx - o * (2 * SW_TY);
    const int gx = ox0 + o, cc = c < SW_TY ? c : c - SW_TY, gy = oy0 + cc;
    if (gx >= sx || gy >= sy) continue;
    T lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const T v = stmp[2 * o + K - 1 - k][c];
      lo += flo[k] * v;
      hi += fhi[k] * v;
    }
    const int ycol = c < SW_TY ? gy : sy + gy;
    out[(int64_t)gx * ld_out + ycol] = lo;
    out[(int64_t)(sx + gx) * ld_out + ycol] = hi;
    if (approx && c < SW_TY) approx[(int64_t)gx * sy + gy] = lo;
  }
}

// ---------------------------------------------------------------------------
// synthesis, one level: quadrants LL (from `ll`, its own stride), LH/HL/HH (from the block) -> out (nxo, nyo)
//   cols: cb[2t+p, c] = sum_m lo[2m+p] X0[t+H-1-m, c] + hi[2m+p] X1[t+H-1-m, c]     (H = K/2)
//   rows: out[r, 2u+p] = sum_m lo[2m+p] cb[r, u+H-1-m] + hi[2m+p] cb[r, sy+u+H-1-m]
// ---------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(256)
k_idwt_level(const T* __restrict__ ll, int ld_ll, const T* __restrict__ blk, int ld_blk, int sx, int sy,
             T* __restrict__ out, int ld_out, int nxo, int nyo, DwtFilt f, int accumulate, BandPtr bp) {
  constexpr int HM = SW_KMAX / 2, RX = SW_TX + HM - 1, RY = SW_TY + HM - 1;
  __shared__ T sc[2][2][RX][RY + 1];       // [x half][y half]
  __shared__ T cb[2 * SW_TX][2][RY + 1];
  constexpr int H = K / 2;
  const int tid = threadIdx.x;
  const int tx0 = blockIdx.y * SW_TX, ty0 = blockIdx.x * SW_TY;
  T flo[K], fhi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { flo[k] = (T)f.lo[k]; fhi[k] = (T)f.hi[k]; }
  ll += (int64_t)blockIdx.z * bp.aux_stride;
  blk += (int64_t)blockIdx.z * bp.in_stride;
  out += (int64_t)blockIdx.z * bp.out_stride;
  const int nrx = SW_TX + H - 1, nry = SW_TY + H - 1;
  for (int idx = tid; idx < 4 * nrx * nry; idx += 256) {
    const int qd = idx / (nrx * nry), rem = idx - qd * (nrx * nry);
    const int r = rem / nry, c = rem - r * nry;
    const int xh = qd >> 1, yh = qd & 1;
    const int gi = tx0 + r, gj = ty0 + c;
    T v = 0;
    if (gi < sx && gj < sy) {
      if (qd == 0) v = ll[(int64_t)gi * ld_ll + gj];
      else v = blk[(int64_t)(xh * sx + gi) * ld_blk + yh * sy + gj];
    }
    sc[xh][yh][r][c] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < 2 * SW_TX * 2 * nry; idx += 256) {
    const int r = idx / (2 * nry), rem = idx - r * (2 * nry);
    const int yh = rem / nry, c = rem - yh * nry;
    const int t = r >> 1, p = r & 1;
    T acc = 0;
#pragma unroll
    for (int m = 0; m < H; ++m) {
      const T a = sc[0][yh][t + H - 1 - m][c], b = sc[1][yh][t + H - 1 - m][c];
      acc += (p ? flo[2 * m + 1] : flo[2 * m]) * a + (p ? fhi[2 * m + 1] : fhi[2 * m]) * b;
    }
    cb[r][yh][c] = acc;
  }
  __syncthreads();
  for (int idx = tid; idx < 2 * SW_TX * 2 * SW_TY; idx += 256) {
    const int r = idx / (2 * SW_TY), cc = idx - r * (2 * SW_TY);
    const int u = cc >> 1, p = cc & 1;
    const int gr = 2 * tx0 + r, gc = 2 * ty0 + cc;
    if (gr >= nxo || gc >= nyo) continue;
    T acc = 0;
#pragma unroll
    for (int m = 0; m < H; ++m) {
      const T a = cb[r][0][u + H - 1 - m], b = cb[r][1][u + H - 1 - m];
      acc += (p ? flo[2 * m + 1] : flo[2 * m]) * a + (p ? fhi[2 * m + 1] : fhi[2 * m]) * b;
    }
    T* o = out + (int64_t)gr * ld_out + gc;
    *o = accumulate ? *o + acc : acc;
  }
}

// 'self' basis and layout helpers ------------------------------------------------------------
// dst[b][i*ld_d + j] (=|+=) src[b][i*ld_s + j] over an (n0, n1) window
template <typename T>
__global__ void k_copy2d(const T* __restrict__ src, int ld_s, int64_t sstride, T* __restrict__ dst, int ld_d,
                         int64_t dstride, int n0, int n1, int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n1 || i >= n0) return;
  const T v = src[(int64_t)blockIdx.z * sstride + (int64_t)i * ld_s + j];
  T* d = dst + (int64_t)blockIdx.z * dstride + (int64_t)i * ld_d + j;
  *d = accumulate ? *d + v : v;
}

// batched transpose: dst[z][j][i] = src[z][i][j], src (n0, n1)
template <typename T>
__global__ void k_transpose(const T* __restrict__ src, T* __restrict__ dst, int n0, int n1) {
  __shared__ T tile[32][33];
  const int64_t off = (int64_t)blockIdx.z * n0 * n1;
  int i = blockIdx.y * 32 + threadIdx.y, j = blockIdx.x * 32 + threadIdx.x;
  for (int k = 0; k < 32; k += 8)
    if (i + k < n0 && j < n1) tile[threadIdx.y + k][threadIdx.x] = src[off + (int64_t)(i + k) * n1 + j];
  __syncthreads();
  i = blockIdx.x * 32 + threadIdx.y;
  j = blockIdx.y * 32 + threadIdx.x;
  for (int k = 0; k < 32; k += 8)
    if (i + k < n1 && j < n0) dst[off + (int64_t)(i + k) * n0 + j] = tile[threadIdx.x][threadIdx.y + k];
}

// l21 dual update / prox ----------------------------------------------------------------------
// vbar != nullptr: also writes the extrapolated dual 2 v_new - vp (opt/primal_dual.py:16-23) in the same pass, so the
// loop needs neither the separate extrapolation kernel nor the vp <- v copy (the caller swaps the two buffers)
template <typename T>
__global__ void k_dual_update(const T* __restrict__ vp, T* __restrict__ v, const T* __restrict__ w, T lam, T sigma,
                              int nband, int64_t ncoef, T* __restrict__ bsum, int phase, T* __restrict__ vbar) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncoef) return;
  T sum = 0;
  if (phase == 2) {
    sum = bsum[k];
  } else {
    for (int b = 0; b < nband; ++b) {
      const int64_t o = (int64_t)b * ncoef + k;
      const T vt = vp[o] + sigma * v[o];
      v[o] = vt;
      sum += vt;
    }
    if (phase == 1) { bsum[k] = sum; return; }
  }
  const T s = fabs(sum), thr = lam * w[k];
  const bool shrink = s > thr;
  if (!shrink && !vbar) return;
  const T sc = shrink ? thr / s : (T)1;
  for (int b = 0; b < nband; ++b) {
    const int64_t o = (int64_t)b * ncoef + k;
    T vn = v[o];
    if (shrink) { vn *= sc; v[o] = vn; }
    if (vbar) vbar[o] = (T)2 * vn - vp[o];
  }
}

template <typename T>
__global__ void k_prox_21m(const T* __restrict__ v, T* __restrict__ res, const T* __restrict__ w, T lam, T sigma,
                           int nband, int64_t ncoef) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncoef) return;
  T sum = 0;
  for (int b = 0; b < nband; ++b) sum += v[(int64_t)b * ncoef + k];
  sum /= sigma;
  T ratio = 0;
  if (sum != (T)0) {
    const T a = fabs(sum);
    const T soft = fmax(a - lam * w[k] / sigma, (T)0);
    ratio = soft / a / sigma;
  }
  for (int b = 0; b < nband; ++b) res[(int64_t)b * ncoef + k] = v[(int64_t)b * ncoef + k] * ratio;
}

template <typename T>
__global__ void k_extrapolate(const T* __restrict__ v, T* __restrict__ vp, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) vp[k] = (T)2 * v[k] - vp[k];
}

// out = a x + b y (out may alias x or y)
template <typename T>
__global__ void k_axpby(T* out, T a, const T* x, T b, const T* y, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = a * x[k] + b * y[k];
}

// x, xp and xout may be one and the same array (ForwardBackward's in-place positivity pass calls this with
// tau = 0): no __restrict__ here, and tau == 0 leaves xp untouched even where xout is inf / NaN
template <typename T>
__global__ void k_primal_step(T* x, const T* xp, const T* xout, T tau, int positivity, int nband, int64_t npix) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= npix) return;
  bool kill = false;
  for (int b = 0; b < nband; ++b) {
    const int64_t o = (int64_t)b * npix + k;
    T v = tau != (T)0 ? xp[o] - tau * xout[o] : xp[o];
    if (positivity == 1 && v < (T)0) v = 0;
    if (positivity == 2 && v <= (T)0) kill = true;
    x[o] = v;
  }
  if (kill)
    for (int b = 0; b < nband; ++b) x[(int64_t)b * npix + k] = 0;
}

template <typename T>
__global__ void k_norm_diff(const T* __restrict__ x, const T* __restrict__ xp, int64_t n, double* __restrict__ acc) {
  double num = 0, den = 0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const double a = (double)x[k], d = a - (double)xp[k];
    num += d * d;
    den += a * a;
  }
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc, num);
    atomicAdd(acc + 1, den);
  }
}

// acc[0] += <a, b>, acc[1] += <c, d> (fp64 accumulation): the scalar products of the device-resident CG / power method
template <typename T>
__global__ void k_dot2(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, const T* __restrict__ d,
                       int64_t n, double* __restrict__ acc) {
  double s0 = 0, s1 = 0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    s0 += (double)a[k] * (double)b[k];
    s1 += (double)c[k] * (double)d[k];
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc, s0);
    atomicAdd(acc + 1, s1);
  }
}
