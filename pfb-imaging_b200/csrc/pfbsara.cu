// C-ABI of the device SARA backward step (include/pfbsara.h); kernels in sara.cuh.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/pfbsara.h"
#include "sara.cuh"

// error plumbing shared with pfbgrid.cu (same shared library)
int pfbg_fail(int code, const char* fmt, ...);
void pfbg_count_launch();

#define SCK(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return pfbg_fail(PFBG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define SRC(call)                   \
  do {                              \
    int rc_ = (call);               \
    if (rc_ != PFBG_OK) return rc_; \
  } while (0)

struct pfbs_psi {
  int precision = 0, device = 0;
  int nband = 0, nx = 0, ny = 0, nbasis = 0, nlevel = 0, nxmax = 0, nymax = 0;
  std::vector<int> K;
  std::vector<DwtFilt> dec, rec;
  std::vector<int64_t> ix, iy, sx, sy, spx, spy, ntotx, ntoty;
  void *approx[2] = {nullptr, nullptr};  // (nband, sx0max, sy0max) ping-pong: LL of the previous level
  void *img[2] = {nullptr, nullptr};     // (nband, nx+1, ny+1) ping-pong: inner reconstructions
  void *d_x = nullptr, *d_alpha = nullptr, *d_alpha_t = nullptr;  // staging for host-pointer / transposed calls
  size_t approx_elems = 0, img_elems = 0;
  int64_t L(const std::vector<int64_t>& a, int b, int l) const { return a[(size_t)b * nlevel + l]; }
};

static size_t rbytes(int precision) { return precision == PFBG_F32 ? 4 : 8; }

extern "C" int pfbs_psi_destroy(pfbs_psi* p) {
  if (!p) return PFBG_OK;
  cudaSetDevice(p->device);
  void* all[] = {p->approx[0], p->approx[1], p->img[0], p->img[1], p->d_x, p->d_alpha, p->d_alpha_t};
  for (void* q : all)
    if (q) cudaFree(q);
  delete p;
  return PFBG_OK;
}

extern "C" int pfbs_psi_create(int32_t precision, int32_t device, int32_t nband, int32_t nx, int32_t ny, int32_t nbasis,
                               int32_t nlevel, const int32_t* K, const double* filters, const int64_t* ix,
                               const int64_t* iy, const int64_t* sx, const int64_t* sy, const int64_t* spx,
                               const int64_t* spy, const int64_t* ntotx, const int64_t* ntoty, int32_t nxmax,
                               int32_t nymax, pfbs_psi** out) {
  if (!out || !K || !filters || !ix || !iy || !sx || !sy || !spx || !spy || !ntotx || !ntoty)
    return pfbg_fail(PFBG_ERR_ARG, "null argument");
  *out = nullptr;
  if (precision != PFBG_F32 && precision != PFBG_F64) return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  if (nband < 1 || nx < 1 || ny < 1 || nbasis < 1 || nlevel < 1 || nxmax < nx || nymax < ny)
    return pfbg_fail(PFBG_ERR_ARG, "bad sizes");
  int ndev = 0;
  SCK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return pfbg_fail(PFBG_ERR_ARG, "device %d out of range", device);
  SCK(cudaSetDevice(device));
  pfbs_psi* p = new pfbs_psi();
  p->precision = precision; p->device = device; p->nband = nband; p->nx = nx; p->ny = ny;
  p->nbasis = nbasis; p->nlevel = nlevel; p->nxmax = nxmax; p->nymax = nymax;
  p->K.assign(K, K + nbasis);
  p->dec.resize(nbasis); p->rec.resize(nbasis);
  const size_t nl = (size_t)nbasis * nlevel;
  p->ix.assign(ix, ix + 2 * nl); p->iy.assign(iy, iy + 2 * nl);
  p->sx.assign(sx, sx + nl); p->sy.assign(sy, sy + nl); p->spx.assign(spx, spx + nl); p->spy.assign(spy, spy + nl);
  p->ntotx.assign(ntotx, ntotx + nbasis); p->ntoty.assign(ntoty, ntoty + nbasis);
  size_t amax = 1;
  for (int b = 0; b < nbasis; ++b) {
    const int k = K[b];
    if (k == 0) continue;
    if (k < 2 || k > PFBS_KMAX || (k & 1)) { delete p; return pfbg_fail(PFBG_ERR_ARG, "filter length %d not in 2..%d (even)", k, PFBS_KMAX); }
    if (ntotx[b] > nxmax || ntoty[b] > nymax) { delete p; return pfbg_fail(PFBG_ERR_ARG, "basis %d does not fit (nxmax, nymax)", b); }
    const double* f = filters + (size_t)b * 4 * PFBS_KMAX;
    memset(&p->dec[b], 0, sizeof(DwtFilt)); memset(&p->rec[b], 0, sizeof(DwtFilt));
    p->dec[b].K = p->rec[b].K = k;
    for (int t = 0; t < k; ++t) {
      p->dec[b].lo[t] = f[t]; p->dec[b].hi[t] = f[PFBS_KMAX + t];
      p->rec[b].lo[t] = f[2 * PFBS_KMAX + t]; p->rec[b].hi[t] = f[3 * PFBS_KMAX + t];
    }
    for (int l = 0; l < nlevel; ++l) {
      const size_t a = (size_t)p->L(p->sx, b, l) * p->L(p->sy, b, l);
      if (a > amax) amax = a;
    }
  }
  const size_t rb = rbytes(precision);
  p->approx_elems = amax;
  p->img_elems = (size_t)(nx + 2) * (ny + 2);
  cudaError_t e = cudaSuccess;
  for (int t = 0; t < 2 && e == cudaSuccess; ++t) {
    e = cudaMalloc(&p->approx[t], (size_t)nband * p->approx_elems * rb);
    if (e == cudaSuccess) e = cudaMalloc(&p->img[t], (size_t)nband * p->img_elems * rb);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    pfbs_psi_destroy(p);
    return pfbg_fail(PFBG_ERR_NOMEM, "scratch allocation failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return PFBG_OK;
}

static int stage_alloc(void** buf, size_t bytes) {
  if (*buf) return PFBG_OK;
  cudaError_t e = cudaMalloc(buf, bytes);
  if (e != cudaSuccess) { *buf = nullptr; cudaGetLastError(); return pfbg_fail(PFBG_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); }
  return PFBG_OK;
}

template <typename T>
static int psi_dot_t(pfbs_psi* p, const T* x, T* alpha, cudaStream_t s) {
  const int64_t img = (int64_t)p->nx * p->ny, plane = (int64_t)p->nxmax * p->nymax;
  const int64_t astride = (int64_t)p->nbasis * plane;
  SCK(cudaMemsetAsync(alpha, 0, (size_t)p->nband * astride * sizeof(T), s));
  for (int b = 0; b < p->nbasis; ++b) {
    T* ab = alpha + (int64_t)b * plane;
    if (p->K[b] == 0) {
      k_copy2d<T><<<dim3((p->ny + 127) / 128, p->nx, p->nband), 128, 0, s>>>(x, p->ny, img, ab, p->nymax, astride, p->nx, p->ny, 0);
      pfbg_count_launch();
      continue;
    }
    const T* in = x;
    int ld_in = p->ny, nxin = p->nx, nyin = p->ny;
    int64_t in_stride = img;
    for (int l = 0; l < p->nlevel; ++l) {
      const int sx = (int)p->L(p->sx, b, l), sy = (int)p->L(p->sy, b, l);
      const int hx = (int)p->ix[((size_t)b * p->nlevel + l) * 2 + 1], hy = (int)p->iy[((size_t)b * p->nlevel + l) * 2 + 1];
      const int lx = hx - 2 * sx, ly = hy - 2 * sy;
      T* ap = (l + 1 < p->nlevel) ? (T*)p->approx[l & 1] : nullptr;
      BandPtr bp{in_stride, astride, (int64_t)p->approx_elems};
      dim3 grd((sy + SW_TY - 1) / SW_TY, (sx + SW_TX - 1) / SW_TX, p->nband);
      T* blk = ab + (int64_t)lx * p->nymax + ly;
      switch (p->K[b]) {
        case 2: k_dwt_level<T, 2><<<grd, 256, 0, s>>>(in, ld_in, nxin, nyin, blk, p->nymax, sx, sy, ap, p->dec[b], bp); break;
        case 4: k_dwt_level<T, 4><<<grd, 256, 0, s>>>(in, ld_in, nxin, nyin, blk, p->nymax, sx, sy, ap, p->dec[b], bp); break;
        case 6: k_dwt_level<T, 6><<<grd, 256, 0, s>>>(in, ld_in, nxin, nyin, blk, p->nymax, sx, sy, ap, p->dec[b], bp); break;
        case 8: k_dwt_level<T, 8><<<grd, 256, 0, s>>>(in, ld_in, nxin, nyin, blk, p->nymax, sx, sy, ap, p->dec[b], bp); break;
        default: k_dwt_level<T, 10><<<grd, 256, 0, s>>>(in, ld_in, nxin, nyin, blk, p->nymax, sx, sy, ap, p->dec[b], bp); break;
      }
      pfbg_count_launch();
      in = ap; ld_in = sy; nxin = sx; nyin = sy; in_stride = (int64_t)p->approx_elems;
    }
  }
  SCK(cudaGetLastError());
  return PFBG_OK;
}

template <typename T>
static int psi_hdot_t(pfbs_psi* p, const T* alpha, T* x, cudaStream_t s) {
  const int64_t img = (int64_t)p->nx * p->ny, plane = (int64_t)p->nxmax * p->nymax;
  const int64_t astride = (int64_t)p->nbasis * plane;
  SCK(cudaMemsetAsync(x, 0, (size_t)p->nband * img * sizeof(T), s));
  for (int b = 0; b < p->nbasis; ++b) {
    const T* ab = alpha + (int64_t)b * plane;
    if (p->K[b] == 0) {
      k_copy2d<T><<<dim3((p->ny + 127) / 128, p->nx, p->nband), 128, 0, s>>>(ab, p->nymax, astride, x, p->ny, img, p->nx, p->ny, 1);
      pfbg_count_launch();
      continue;
    }
    const T* ll = nullptr;  // LL source of the current level: the block's own quadrant at the deepest level
    int ld_ll = 0;
    int64_t ll_stride = 0;
    for (int l = p->nlevel - 1; l >= 0; --l) {
      const int sx = (int)p->L(p->sx, b, l), sy = (int)p->L(p->sy, b, l);
      const int hx = (int)p->ix[((size_t)b * p->nlevel + l) * 2 + 1], hy = (int)p->iy[((size_t)b * p->nlevel + l) * 2 + 1];
      const int lx = hx - 2 * sx, ly = hy - 2 * sy;
      const T* blk = ab + (int64_t)lx * p->nymax + ly;
      if (l == p->nlevel - 1) { ll = blk; ld_ll = p->nymax; ll_stride = astride; }
      int nxo = (int)p->L(p->spx, b, l), nyo = (int)p->L(p->spy, b, l);
      if (nxo > p->nx) nxo = p->nx;
      if (nyo > p->ny) nyo = p->ny;
      T* dst; int ld_dst; int64_t dst_stride; int acc;
      if (l == 0) { dst = x; ld_dst = p->ny; dst_stride = img; acc = 1; }
      else { dst = (T*)p->img[l & 1]; ld_dst = p->ny + 2; dst_stride = (int64_t)p->img_elems; acc = 0; }
      BandPtr bp{astride, dst_stride, ll_stride};
      dim3 grd(((nyo + 1) / 2 + SW_TY - 1) / SW_TY, ((nxo + 1) / 2 + SW_TX - 1) / SW_TX, p->nband);
      switch (p->K[b]) {
        case 2: k_idwt_level<T, 2><<<grd, 256, 0, s>>>(ll, ld_ll, blk, p->nymax, sx, sy, dst, ld_dst, nxo, nyo, p->rec[b], acc, bp); break;
        case 4: k_idwt_level<T, 4><<<grd, 256, 0, s>>>(ll, ld_ll, blk, p->nymax, sx, sy, dst, ld_dst, nxo, nyo, p->rec[b], acc, bp); break;
        case 6: k_idwt_level<T, 6><<<grd, 256, 0, s>>>(ll, ld_ll, blk, p->nymax, sx, sy, dst, ld_dst, nxo, nyo, p->rec[b], acc, bp); break;
        case 8: k_idwt_level<T, 8><<<grd, 256, 0, s>>>(ll, ld_ll, blk, p->nymax, sx, sy, dst, ld_dst, nxo, nyo, p->rec[b], acc, bp); break;
        default: k_idwt_level<T, 10><<<grd, 256, 0, s>>>(ll, ld_ll, blk, p->nymax, sx, sy, dst, ld_dst, nxo, nyo, p->rec[b], acc, bp); break;
      }
      pfbg_count_launch();
      ll = dst; ld_ll = ld_dst; ll_stride = dst_stride;
    }
  }
  SCK(cudaGetLastError());
  return PFBG_OK;
}

template <typename T>
static int transpose_planes(const T* src, T* dst, int nplanes, int n0, int n1, cudaStream_t s) {
  for (int z0 = 0; z0 < nplanes; z0 += 65535) {
    const int nz = nplanes - z0 < 65535 ? nplanes - z0 : 65535;
    k_transpose<T><<<dim3((n1 + 31) / 32, (n0 + 31) / 32, nz), dim3(32, 8), 0, s>>>(src + (int64_t)z0 * n0 * n1, dst + (int64_t)z0 * n0 * n1, n0, n1);
    pfbg_count_launch();
  }
  SCK(cudaGetLastError());
  return PFBG_OK;
}

template <typename T>
static int psi_apply(pfbs_psi* p, const void* in, void* out, uint32_t flags, cudaStream_t s, bool fwd) {
  const bool dev = flags & PFBG_DEVICE_PTRS, tr = flags & PFBS_TRANSPOSED;
  const size_t xb = (size_t)p->nband * p->nx * p->ny * sizeof(T);
  const size_t ab = (size_t)p->nband * p->nbasis * p->nxmax * p->nymax * sizeof(T);
  const int nplanes = p->nband * p->nbasis;
  if (fwd) {
    const T* dx = (const T*)in;
    if (!dev) { SRC(stage_alloc(&p->d_x, xb)); SCK(cudaMemcpyAsync(p->d_x, in, xb, cudaMemcpyHostToDevice, s)); dx = (const T*)p->d_x; }
    T* da = (T*)out;
    if (!dev || tr) { SRC(stage_alloc(&p->d_alpha, ab)); da = (T*)p->d_alpha; }
    SRC(psi_dot_t<T>(p, dx, da, s));
    if (tr) {
      T* dt = (T*)out;
      if (!dev) { SRC(stage_alloc(&p->d_alpha_t, ab)); dt = (T*)p->d_alpha_t; }
      SRC(transpose_planes<T>(da, dt, nplanes, p->nxmax, p->nymax, s));
      da = dt;
    }
    if (!dev) { SCK(cudaMemcpyAsync(out, da, ab, cudaMemcpyDeviceToHost, s)); SCK(cudaStreamSynchronize(s)); }
  } else {
    const T* da = (const T*)in;
    if (!dev) { SRC(stage_alloc(tr ? &p->d_alpha_t : &p->d_alpha, ab)); void* st = tr ? p->d_alpha_t : p->d_alpha; SCK(cudaMemcpyAsync(st, in, ab, cudaMemcpyHostToDevice, s)); da = (const T*)st; }
    if (tr) {
      SRC(stage_alloc(&p->d_alpha, ab));
      SRC(transpose_planes<T>(da, (T*)p->d_alpha, nplanes, p->nymax, p->nxmax, s));
      da = (const T*)p->d_alpha;
    }
    T* dx = (T*)out;
    if (!dev) { SRC(stage_alloc(&p->d_x, xb)); dx = (T*)p->d_x; }
    SRC(psi_hdot_t<T>(p, da, dx, s));
    if (!dev) { SCK(cudaMemcpyAsync(out, dx, xb, cudaMemcpyDeviceToHost, s)); SCK(cudaStreamSynchronize(s)); }
  }
  return PFBG_OK;
}

extern "C" int pfbs_psi_dot(pfbs_psi* p, const void* x, void* alpha, uint32_t flags, void* stream) {
  if (!p || !x || !alpha) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(p->device));
  return p->precision == PFBG_F32 ? psi_apply<float>(p, x, alpha, flags, (cudaStream_t)stream, true)
                                  : psi_apply<double>(p, x, alpha, flags, (cudaStream_t)stream, true);
}

extern "C" int pfbs_psi_hdot(pfbs_psi* p, const void* alpha, void* x, uint32_t flags, void* stream) {
  if (!p || !x || !alpha) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(p->device));
  return p->precision == PFBG_F32 ? psi_apply<float>(p, alpha, x, flags, (cudaStream_t)stream, false)
                                  : psi_apply<double>(p, alpha, x, flags, (cudaStream_t)stream, false);
}

// ---------------------------------------------------------------------------
// element-wise pieces (device pointers only: they live inside the device-resident primal-dual loop)
// ---------------------------------------------------------------------------
static unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

extern "C" int pfbs_dual_update(int32_t precision, int32_t device, const void* vp, void* v, const void* weight,
                                double lam, double sigma, int32_t nband, int64_t ncoef, void* bsum, int32_t phase,
                                void* vbar, void* stream) {
  if (!v || !weight || (phase != 2 && !vp) || (phase != 0 && !bsum) || (vbar && !vp))
    return pfbg_fail(PFBG_ERR_ARG, "null argument");
  if (vbar && phase == 1) return pfbg_fail(PFBG_ERR_ARG, "the extrapolated dual is written by phase 0 or 2");
  if (phase < 0 || phase > 2 || nband < 1 || ncoef < 0) return pfbg_fail(PFBG_ERR_ARG, "bad phase / sizes");
  SCK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (ncoef == 0) return PFBG_OK;
  if (precision == PFBG_F32)
    k_dual_update<float><<<nblk(ncoef), 256, 0, s>>>((const float*)vp, (float*)v, (const float*)weight, (float)lam, (float)sigma, nband, ncoef, (float*)bsum, phase, (float*)vbar);
  else if (precision == PFBG_F64)
    k_dual_update<double><<<nblk(ncoef), 256, 0, s>>>((const double*)vp, (double*)v, (const double*)weight, lam, sigma, nband, ncoef, (double*)bsum, phase, (double*)vbar);
  else return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  pfbg_count_launch();
  SCK(cudaGetLastError());
  return PFBG_OK;
}

extern "C" int pfbs_prox_21m(int32_t precision, int32_t device, const void* v, void* result, const void* weight,
                             double lam, double sigma, int32_t nband, int64_t ncoef, void* stream) {
  if (!v || !result || !weight) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  if (nband < 1 || ncoef < 0 || !(sigma > 0)) return pfbg_fail(PFBG_ERR_ARG, "bad sizes / sigma");
  SCK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (ncoef == 0) return PFBG_OK;
  if (precision == PFBG_F32)
    k_prox_21m<float><<<nblk(ncoef), 256, 0, s>>>((const float*)v, (float*)result, (const float*)weight, (float)lam, (float)sigma, nband, ncoef);
  else if (precision == PFBG_F64)
    k_prox_21m<double><<<nblk(ncoef), 256, 0, s>>>((const double*)v, (double*)result, (const double*)weight, lam, sigma, nband, ncoef);
  else return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  pfbg_count_launch();
  SCK(cudaGetLastError());
  return PFBG_OK;
}

extern "C" int pfbs_extrapolate(int32_t precision, int32_t device, const void* v, void* vp, int64_t n, void* stream) {
  if (!v || !vp) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) return PFBG_OK;
  if (precision == PFBG_F32) k_extrapolate<float><<<nblk(n), 256, 0, s>>>((const float*)v, (float*)vp, n);
  else if (precision == PFBG_F64) k_extrapolate<double><<<nblk(n), 256, 0, s>>>((const double*)v, (double*)vp, n);
  else return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  pfbg_count_launch();
  SCK(cudaGetLastError());
  return PFBG_OK;
}

extern "C" int pfbs_axpby(int32_t precision, int32_t device, void* out, double a, const void* x, double b, const void* y,
                          int64_t n, void* stream) {
  if (!out || !x || !y) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) return PFBG_OK;
  if (precision == PFBG_F32) k_axpby<float><<<nblk(n), 256, 0, s>>>((float*)out, (float)a, (const float*)x, (float)b, (const float*)y, n);
  else if (precision == PFBG_F64) k_axpby<double><<<nblk(n), 256, 0, s>>>((double*)out, a, (const double*)x, b, (const double*)y, n);
  else return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  pfbg_count_launch();
  SCK(cudaGetLastError());
  return PFBG_OK;
}

extern "C" int pfbs_primal_step(int32_t precision, int32_t device, void* x, const void* xp, const void* xout, double tau,
                                int32_t positivity, int32_t nband, int64_t npix, void* stream) {
  if (!x || !xp || !xout) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  if (positivity < 0 || positivity > 2 || nband < 1) return pfbg_fail(PFBG_ERR_ARG, "bad positivity mode / nband");
  SCK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (npix <= 0) return PFBG_OK;
  if (precision == PFBG_F32)
    k_primal_step<float><<<nblk(npix), 256, 0, s>>>((float*)x, (const float*)xp, (const float*)xout, (float)tau, positivity, nband, npix);
  else if (precision == PFBG_F64)
    k_primal_step<double><<<nblk(npix), 256, 0, s>>>((double*)x, (const double*)xp, (const double*)xout, tau, positivity, nband, npix);
  else return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  pfbg_count_launch();
  SCK(cudaGetLastError());
  return PFBG_OK;
}

// two doubles of device scratch for the reductions, held for the duration of the call (pfbgrid.cu: recycled blocks,
// one per concurrent caller, so host threads / streams on one device never share an accumulator; no cudaMalloc /
// cudaFree, which synchronises the device, inside the solver loops)
void* pfbg_scratch_get(int device);
void pfbg_scratch_put(int device, void* p);

template <typename K>
static int reduce2(int device, cudaStream_t s, double* out2, K launch) {
  struct Hold {
    int dev; void* p;
    ~Hold() { pfbg_scratch_put(dev, p); }
  } hold{device, pfbg_scratch_get(device)};
  double* acc = (double*)hold.p;
  if (!acc) return pfbg_fail(PFBG_ERR_NOMEM, "no reduction scratch on device %d", device);
  SCK(cudaMemsetAsync(acc, 0, 16, s));
  launch(acc);
  pfbg_count_launch();
  SCK(cudaGetLastError());
  SCK(cudaMemcpyAsync(out2, acc, 16, cudaMemcpyDeviceToHost, s));
  SCK(cudaStreamSynchronize(s));
  return PFBG_OK;
}

extern "C" int pfbs_norm_diff(int32_t precision, int32_t device, const void* x, const void* xp, int64_t n,
                              double* num_den, void* stream) {
  if (!x || !xp || !num_den) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  if (precision != PFBG_F32 && precision != PFBG_F64) return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  SCK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  num_den[0] = num_den[1] = 0.0;
  if (n <= 0) return PFBG_OK;
  const unsigned grd = n < (int64_t)1184 * 256 ? nblk(n) : 1184;  // 148 SMs x 8 CTAs
  return reduce2(device, s, num_den, [&](double* acc) {
    if (precision == PFBG_F32) k_norm_diff<float><<<grd, 256, 0, s>>>((const float*)x, (const float*)xp, n, acc);
    else k_norm_diff<double><<<grd, 256, 0, s>>>((const double*)x, (const double*)xp, n, acc);
  });
}

extern "C" int pfbs_dot2(int32_t precision, int32_t device, const void* a, const void* b, const void* c, const void* d,
                         int64_t n, double* out2, void* stream) {
  if (!a || !b || !c || !d || !out2) return pfbg_fail(PFBG_ERR_ARG, "null argument");
  if (precision != PFBG_F32 && precision != PFBG_F64) return pfbg_fail(PFBG_ERR_ARG, "bad precision");
  SCK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  out2[0] = out2[1] = 0.0;
  if (n <= 0) return PFBG_OK;
  const unsigned grd = n < (int64_t)1184 * 256 ? nblk(n) : 1184;
  return reduce2(device, s, out2, [&](double* acc) {
    if (precision == PFBG_F32) k_dot2<float><<<grd, 256, 0, s>>>((const float*)a, (const float*)b, (const float*)c, (const float*)d, n, acc);
    else k_dot2<double><<<grd, 256, 0, s>>>((const double*)a, (const double*)b, (const double*)c, (const double*)d, n, acc);
  });
}
