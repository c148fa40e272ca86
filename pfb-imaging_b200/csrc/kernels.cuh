// CUDA kernels of the B200 w-gridder (sm_100a).
//   kernel 1  k_bin            : uv-tile / w-plane bucket key per sample
//   kernel 2  k_grid_direct    : spread samples onto the plane stack (vector RED; last-resort path, the
//                                run kernels of runs.cuh are the default)
//   kernel 3  k_degrid_direct  : gather samples from the plane stack
//   kernel 4  k_img2grid / k_grid2img / k_corr_init : w-screen, grid correction,
//             taper (beam), wsum and ridge fused with the pad / crop.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// kernel 1: bucket keys
// ---------------------------------------------------------------------------
__global__ void k_bin(GParams p, const double* __restrict__ uvw, const double* __restrict__ fscale,
                      const uint8_t* __restrict__ mask, int64_t nvis, uint64_t inactive_key,
                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                      unsigned long long* __restrict__ nactive) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool act = false;
  if (k < nvis) {
    int64_t row = k / p.nchan;
    int chan = (int)(k - row * p.nchan);
    act = mask ? (mask[k] != 0) : true;
    uint64_t key = inactive_key;
    if (act) {
      VisCoord c = vis_coord(p, uvw, fscale, row, chan);
      key = bucket_key(p, c);
    }
    keys[k] = key;
    vals[k] = (uint32_t)k;
  }
  unsigned b = __ballot_sync(0xffffffffu, act);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(nactive, (unsigned long long)__popc(b));
}

// debug / parity dump of the per-sample indices (all nvis samples)
__global__ void k_bin_dump(GParams p, const double* __restrict__ uvw, const double* __restrict__ fscale,
                           int64_t nvis, int32_t* iu0, int32_t* iv0, int32_t* ip0, uint64_t* key) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nvis) return;
  int64_t row = k / p.nchan;
  int chan = (int)(k - row * p.nchan);
  VisCoord c = vis_coord(p, uvw, fscale, row, chan);
  if (iu0) iu0[k] = c.iu0;
  if (iv0) iv0[k] = c.iv0;
  if (ip0) ip0[k] = c.ip0;
  if (key) key[k] = bucket_key(p, c);
}

// ---------------------------------------------------------------------------
// kernel 2 (general path): one warp per sample, W*W cells over the lanes,
// W planes in the inner loop, vector RED into the L2-resident plane stack.
// Works for any W <= 16 and both precisions.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_grid_direct(GParams p, const double* __restrict__ uvw, const double* __restrict__ fscale,
              const uint32_t* __restrict__ sorted_idx, int64_t nact,
              const typename cplx_of<T>::type* __restrict__ vis, int64_t vis_rs, int64_t vis_cs,
              const T* __restrict__ wgt, typename cplx_of<T>::type* __restrict__ grid,
              int vis_sorted, int apply_phase) {
  using C = typename cplx_of<T>::type;
  __shared__ T sk[8][3][PFBG_MAXW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = p.W, npl = p.do_wgridding ? W : 1;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; k < nact; k += nwarps) {
    uint32_t idx = sorted_idx[k];
    int64_t row = idx / p.nchan;
    int chan = (int)(idx - row * p.nchan);
    VisCoord c = vis_coord(p, uvw, fscale, row, chan);
    C a = vis_sorted ? vis[k] : vis[row * vis_rs + chan * vis_cs];
    T w = wgt ? wgt[idx] : (T)1;
    T pc = 1, ps = 0;
    if (apply_phase) cis_turns(vis_phase_turns(p, c), pc, ps);
    if (apply_phase && c.conj) a.y = -a.y;
    T are = (a.x * pc - a.y * ps) * w, aim = (a.x * ps + a.y * pc) * w;
    __syncwarp();
    for (int t = lane; t < 3 * W; t += 32) {
      int d = t / W, j = t - d * W;
      T v;
      if (d == 0) v = tap<T>(c.gu, c.iu0, j, W, (T)p.beta);
      else if (d == 1) v = tap<T>(c.gv, c.iv0, j, W, (T)p.beta);
      else v = p.do_wgridding ? tap<T>(c.gw, c.ip0, j, W, (T)p.beta) : (T)1;
      sk[warp][d][j] = v;
    }
    __syncwarp();
    if (are == (T)0 && aim == (T)0) continue;
    for (int cell = lane; cell < W * W; cell += 32) {
      int i = cell / W, j = cell - i * W;
      T wuv = sk[warp][0][i] * sk[warp][1][j];
      int iu = wrap(c.iu0 + i, p.nu), iv = wrap(c.iv0 + j, p.nv);
      for (int q = 0; q < npl; ++q) {
        T ww = wuv * sk[warp][2][q];
        bool cj;
        C* g = grid + plane_cell(p, c.ip0 + q, iu, iv, cj);
        atomic_add_c(g, are * ww, cj ? -aim * ww : aim * ww);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// kernel 3 (general path): one warp per sample, gather + warp-shuffle reduce.
// out_sorted != nullptr : write the result at position k of the bucket order
//                         (Hessian path: model visibilities stay on device);
// otherwise scatter to vis_out[idx] (row-major nrow x nchan).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_degrid_direct(GParams p, const double* __restrict__ uvw, const double* __restrict__ fscale,
                const uint32_t* __restrict__ sorted_idx, int64_t nact,
                const typename cplx_of<T>::type* __restrict__ grid, const T* __restrict__ wgt,
                typename cplx_of<T>::type* __restrict__ vis_out,
                typename cplx_of<T>::type* __restrict__ out_sorted, int apply_phase) {
  using C = typename cplx_of<T>::type;
  __shared__ T sk[8][3][PFBG_MAXW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = p.W, npl = p.do_wgridding ? W : 1;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; k < nact; k += nwarps) {
    uint32_t idx = sorted_idx[k];
    int64_t row = idx / p.nchan;
    int chan = (int)(idx - row * p.nchan);
    VisCoord c = vis_coord(p, uvw, fscale, row, chan);
    __syncwarp();
    for (int t = lane; t < 3 * W; t += 32) {
      int d = t / W, j = t - d * W;
      T v;
      if (d == 0) v = tap<T>(c.gu, c.iu0, j, W, (T)p.beta);
      else if (d == 1) v = tap<T>(c.gv, c.iv0, j, W, (T)p.beta);
      else v = p.do_wgridding ? tap<T>(c.gw, c.ip0, j, W, (T)p.beta) : (T)1;
      sk[warp][d][j] = v;
    }
    __syncwarp();
    T accr = 0, acci = 0;
    for (int cell = lane; cell < W * W; cell += 32) {
      int i = cell / W, j = cell - i * W;
      T wuv = sk[warp][0][i] * sk[warp][1][j];
      int iu = wrap(c.iu0 + i, p.nu), iv = wrap(c.iv0 + j, p.nv);
      for (int q = 0; q < npl; ++q) {
        bool cj;
        C v = grid[plane_cell(p, c.ip0 + q, iu, iv, cj)];
        T ww = wuv * sk[warp][2][q];
        accr += v.x * ww;
        acci += (cj ? -v.y : v.y) * ww;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accr += __shfl_xor_sync(0xffffffffu, accr, o);
      acci += __shfl_xor_sync(0xffffffffu, acci, o);
    }
    if (lane == 0) {
      T re = accr, im = acci;
      if (apply_phase) {
        T pc, ps;
        cis_turns(vis_phase_turns(p, c), pc, ps);  // e^{+i t}; we need e^{-i t}
        re = accr * pc + acci * ps;
        im = acci * pc - accr * ps;
      }
      if (apply_phase && c.conj) im = -im;
      if (wgt) { T w = wgt[idx]; re *= w; im *= w; }
      C o; o.x = re; o.y = im;
      if (out_sorted) out_sorted[k] = o; else vis_out[idx] = o;
    }
  }
}

// ---------------------------------------------------------------------------
// kernel 4a: per-pixel correction image, once per plan:
//   corr[i,j] = (-1)^(i'+j') * cu[i] * cv[j] / psihat_w((nm1+nshift) dw) [/ n]
// psihat_w by Gauss-Legendre quadrature: psihat(xi) = sum_k glw[k] cos(pi W xi glx[k]),
// glw[k] already holds  W * w_k * phi(x_k).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_corr_init(GParams p, const double* __restrict__ cu, const double* __restrict__ cv,
                            const double* __restrict__ glx, const double* __restrict__ glw, int ngl,
                            T* __restrict__ corr) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= p.ny || i >= p.nx) return;
  double nm1 = pixel_nm1(p, i, j);
  double v = cu[i] * cv[j];
  if (p.do_wgridding) {
    double xi = (nm1 + p.nshift) * p.dw;
    double a = (double)p.W * xi;  // cos(pi * a * x)
    double s = 0.0;
    for (int k = 0; k < ngl; ++k) s += glw[k] * cospi(a * glx[k]);
    v /= s;
  }
  if (p.divide_by_n) v /= (nm1 + 1.0);
  if (((i - p.nx / 2) + (j - p.ny / 2)) & 1) v = -v;
  corr[(int64_t)i * p.ny + j] = (T)v;
}

// ---------------------------------------------------------------------------
// kernel 4b (degrid direction): image -> zero-padded, w-screened plane stack
//   plane_p[a,b] = x[i,j] * beam[i,j] * corr[i,j] * e^{+2 pi i w_p (nm1+nshift)}
// One thread per padded cell, planes in the inner loop (pixel factors reused).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_img2grid(GParams p, const T* __restrict__ x, const T* __restrict__ beam, const T* __restrict__ corr,
           typename cplx_of<T>::type* __restrict__ grid) {
  using C = typename cplx_of<T>::type;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int a = blockIdx.y;
  if (b >= p.nv) return;
  const int hx = p.nx / 2, hy = p.ny / 2;
  int ip = a < hx ? a : (a >= p.nu - hx ? a - p.nu : INT_MIN);
  int jp = b < hy ? b : (b >= p.nv - hy ? b - p.nv : INT_MIN);
  const int64_t plane_sz = (int64_t)p.nu * p.nv;
  C* g = grid + (int64_t)a * p.nv + b;
  C z; z.x = 0; z.y = 0;
  if (ip == INT_MIN || jp == INT_MIN) {
    for (int q = 0; q < p.nplanes; ++q) g[q * plane_sz] = z;
    return;
  }
  int i = ip + hx, j = jp + hy;
  int64_t pix = (int64_t)i * p.ny + j;
  T val = x[pix] * corr[pix];
  if (beam) val *= beam[pix];
  if (!p.do_wgridding) {
    C o; o.x = val; o.y = 0;
    g[0] = o;
    return;
  }
  if (val == (T)0) {
    for (int q = 0; q < p.nplanes; ++q) g[q * plane_sz] = z;
    return;
  }
  double nu_ = pixel_nm1(p, i, j) + p.nshift;
  for (int q = 0; q < p.nplanes; ++q) {
    T c, s;
    cis_turns((p.w0 + q * p.dw) * nu_, c, s);
    C o; o.x = val * c; o.y = val * s;
    g[q * plane_sz] = o;
  }
}

// ---------------------------------------------------------------------------
// kernel 4c (grid direction): plane stack -> image, fused epilogue
//   acc   = sum_p Re( F_p[a,b] * e^{-2 pi i w_p (nm1+nshift)} )     (fp64 accumulate)
//   out   = acc * corr [* beam] [/ wsum] [+ eta * xin]
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_grid2img(GParams p, const typename cplx_of<T>::type* __restrict__ grid, const T* __restrict__ corr,
           const T* __restrict__ beam, const T* __restrict__ xin, double inv_wsum, double eta,
           T* __restrict__ out) {
  using C = typename cplx_of<T>::type;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= p.ny) return;
  int ip = i - p.nx / 2, jp = j - p.ny / 2;
  int a = ip < 0 ? ip + p.nu : ip, b = jp < 0 ? jp + p.nv : jp;
  const int64_t plane_sz = (int64_t)p.nu * p.nv;
  const C* g = grid + (int64_t)a * p.nv + b;
  double acc = 0.0;
  if (!p.do_wgridding) {
    acc = (double)g[0].x;
  } else {
    double nu_ = pixel_nm1(p, i, j) + p.nshift;
    for (int q = 0; q < p.nplanes; ++q) {
      C v = g[q * plane_sz];
      T c, s;
      cis_turns((p.w0 + q * p.dw) * nu_, c, s);
      acc += (double)(v.x * c + v.y * s);
    }
  }
  int64_t pix = (int64_t)i * p.ny + j;
  double r = acc * (double)corr[pix];
  if (beam) r *= (double)beam[pix];
  r *= inv_wsum;
  if (xin) r += eta * (double)xin[pix];
  out[pix] = (T)r;
}

// PSF visibilities of an off-centre field (operators/gridder.py:616-622, 877-884; utils/stokes2im.py:483-486):
//   vis[r,c] = exp(sign 2 pi i f_c/c0 (u x0 + v y0 - w (n0 - 1)))      with the UNFLIPPED u, v, w
// generated on the device from the bound uvw (never materialised or copied from the host).
template <typename T>
__global__ void k_psf_ramp(const double* __restrict__ uvw, const double* __restrict__ fscale, int64_t nvis, int nchan,
                           double x0, double y0, double nm1_0, double sign, typename cplx_of<T>::type* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nvis) return;
  const int64_t row = k / nchan;
  const int chan = (int)(k - row * nchan);
  const double t = sign * fscale[chan] * (uvw[3 * row] * x0 + uvw[3 * row + 1] * y0 - uvw[3 * row + 2] * nm1_0);
  T c, sn;
  cis_turns(t, c, sn);
  typename cplx_of<T>::type o;
  o.x = c; o.y = sn;
  out[k] = o;
}

// small utility kernels -----------------------------------------------------
template <typename C>
__global__ void k_zero_vis(C* __restrict__ v, int64_t n) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) { C z; z.x = 0; z.y = 0; v[k] = z; }
}

// zero the ACTIVE samples only (the wide degridding teams add their row partials atomically; with
// PFBG_NO_MASK_ZERO the masked samples of the caller's array must stay as they are)
template <typename C>
__global__ void k_zero_active(C* __restrict__ v, const uint32_t* __restrict__ sorted_idx, int64_t nact) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nact) { C z; z.x = 0; z.y = 0; v[sorted_idx[k]] = z; }
}

template <typename T>
__global__ void k_any_nonzero(const T* __restrict__ x, int64_t n, int* flag) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool nz = k < n && x[k] != (T)0;
  if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) *flag = 1;
}
