/* TEST INFRASTRUCTURE ONLY — C/OpenMP restatement of the per-visibility loops of the
 * w-stacked gridder (the CPU baseline timed by bench.py and a fast checker for tests).
 *
 * The algorithm is the published one of ducc0.wgridder 0.41.0 (Arras et al. 2021, A&A 646
 * A58), a third-party dependency of the reference (/root/reference/pyproject.toml:42,
 * uv.lock:1119-1120) whose source is not in the reference tree; call sites:
 * /root/reference/src/pfb_imaging/operators/gridder.py:78-100,128-143 and
 * operators/hessian.py:50-89.  Like ducc0, each thread accumulates a uv tile (+halo) of the
 * W planes a bucket touches in a private buffer and flushes it to the shared grid.
 * Coordinates / bucket order come from oracle/wgridder_np.py:bin_indices (bit-exact spec);
 * the FFTs and image-space factors stay in numpy/scipy (oracle/cwgridder.py).
 *
 * Parity: pinned against the explicit DFT and the golden vectors through
 * tests/test_oracle.py; parity with ducc0's own binary output is unpinned (ducc0 cannot be
 * installed here).  Nothing under pfb-imaging_b200/ links or calls this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define MAXW 16

static inline double es(double x, double beta) {
  double a = (1.0 - x) * (1.0 + x);
  return a < 0.0 ? 0.0 : exp(beta * (sqrt(a) - 1.0));
}

static inline int wrapi(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

int cw_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU baseline sets its thread count explicitly */
void cw_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Spread n samples (already in bucket order) onto grid[P][nu][nv] (interleaved re,im). */
void cw_grid(int64_t n, const int64_t* order, const double* gu, const double* gv, const double* gw,
             const int32_t* iu0, const int32_t* iv0, const int32_t* ip0, const uint64_t* key,
             const double* a /* 2n: re,im of weight*phase*vis */, int W, double beta, int nu, int nv,
             int P, int do_w, double* grid) {
  const int npl = do_w ? W : 1;
  const int side = TILE + MAXW;
  const int64_t plane = (int64_t)nu * nv;
#pragma omp parallel
  {
    double* buf = (double*)calloc((size_t)npl * side * side * 2, sizeof(double));
    int64_t cur = -1;
    int bu = 0, bv = 0, bp = 0;
    double ku[MAXW], kv[MAXW], kw[MAXW];
#pragma omp for schedule(dynamic, 4096)
    for (int64_t s = 0; s < n; ++s) {
      int64_t k = order[s];
      int64_t bucket = (int64_t)(key[k] / (TILE * TILE));
      if (bucket != cur) {
        if (cur >= 0) { /* flush */
          for (int q = 0; q < npl; ++q)
            for (int i = 0; i < TILE + W - 1; ++i) {
              int iu = wrapi(bu + i, nu);
              for (int j = 0; j < TILE + W - 1; ++j) {
                double re = buf[((q * side + i) * side + j) * 2], im = buf[((q * side + i) * side + j) * 2 + 1];
                if (re != 0.0 || im != 0.0) {
                  int iv = wrapi(bv + j, nv);
                  /* planes below zero (mirror planes): Hermitian mirror of plane -p-1, see oracle/wgridder_np._mirror */
                  int pl = bp + q, mu = iu, mv = iv;
                  double sg = 1.0;
                  if (pl < 0) { pl = -pl - 1; mu = iu ? nu - iu : 0; mv = iv ? nv - iv : 0; sg = -1.0; }
                  double* g = grid + ((int64_t)pl * plane + (int64_t)mu * nv + mv) * 2;
#pragma omp atomic
                  g[0] += re;
#pragma omp atomic
                  g[1] += sg * im;
                }
              }
            }
          memset(buf, 0, (size_t)npl * side * side * 2 * sizeof(double));
        }
        cur = bucket;
        bu = (wrapi(iu0[k], nu) / TILE) * TILE;
        bv = (wrapi(iv0[k], nv) / TILE) * TILE;
        bp = ip0[k];
      }
      for (int j = 0; j < W; ++j) {
        ku[j] = es(((double)(iu0[k] + j) - gu[k]) * (2.0 / W), beta);
        kv[j] = es(((double)(iv0[k] + j) - gv[k]) * (2.0 / W), beta);
        kw[j] = do_w ? es(((double)(ip0[k] + j) - gw[k]) * (2.0 / W), beta) : 1.0;
      }
      int ou = wrapi(iu0[k], nu) - bu, ov = wrapi(iv0[k], nv) - bv;
      double are = a[2 * k], aim = a[2 * k + 1];
      for (int q = 0; q < npl; ++q) {
        double qre = are * kw[q], qim = aim * kw[q];
        for (int i = 0; i < W; ++i) {
          double ire = qre * ku[i], iim = qim * ku[i];
          double* row = buf + ((q * side + ou + i) * side + ov) * 2;
          for (int j = 0; j < W; ++j) {
            row[2 * j] += ire * kv[j];
            row[2 * j + 1] += iim * kv[j];
          }
        }
      }
    }
    if (cur >= 0) {
      for (int q = 0; q < npl; ++q)
        for (int i = 0; i < TILE + W - 1; ++i) {
          int iu = wrapi(bu + i, nu);
          for (int j = 0; j < TILE + W - 1; ++j) {
            double re = buf[((q * side + i) * side + j) * 2], im = buf[((q * side + i) * side + j) * 2 + 1];
            if (re != 0.0 || im != 0.0) {
              int iv = wrapi(bv + j, nv);
              int pl = bp + q, mu = iu, mv = iv;
              double sg = 1.0;
              if (pl < 0) { pl = -pl - 1; mu = iu ? nu - iu : 0; mv = iv ? nv - iv : 0; sg = -1.0; }
              double* g = grid + ((int64_t)pl * plane + (int64_t)mu * nv + mv) * 2;
#pragma omp atomic
              g[0] += re;
#pragma omp atomic
              g[1] += sg * im;
            }
          }
        }
    }
    free(buf);
  }
}

/* Gather: out[2k..] = sum over the W^3 support of grid * weights, for every sample. */
void cw_degrid(int64_t n, const int64_t* order, const double* gu, const double* gv, const double* gw,
               const int32_t* iu0, const int32_t* iv0, const int32_t* ip0, int W, double beta, int nu,
               int nv, int P, int do_w, const double* grid, double* out) {
  const int npl = do_w ? W : 1;
  const int64_t plane = (int64_t)nu * nv;
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t s = 0; s < n; ++s) {
    int64_t k = order[s];
    double ku[MAXW], kv[MAXW], kw[MAXW];
    int ivs[MAXW];
    for (int j = 0; j < W; ++j) {
      ku[j] = es(((double)(iu0[k] + j) - gu[k]) * (2.0 / W), beta);
      kv[j] = es(((double)(iv0[k] + j) - gv[k]) * (2.0 / W), beta);
      kw[j] = do_w ? es(((double)(ip0[k] + j) - gw[k]) * (2.0 / W), beta) : 1.0;
      ivs[j] = wrapi(wrapi(iv0[k], nv) + j, nv);
    }
    double re = 0.0, im = 0.0;
    for (int q = 0; q < npl; ++q) {
      double pre = 0.0, pim = 0.0;
      int pl = ip0[k] + q;
      const int mir = pl < 0; /* mirror plane: conj of plane -p-1 at the mirrored cell */
      if (mir) pl = -pl - 1;
      for (int i = 0; i < W; ++i) {
        int iu = wrapi(wrapi(iu0[k], nu) + i, nu);
        if (mir) iu = iu ? nu - iu : 0;
        const double* row = grid + ((int64_t)pl * plane + (int64_t)iu * nv) * 2;
        double rre = 0.0, rim = 0.0;
        for (int j = 0; j < W; ++j) {
          int iv = mir ? (ivs[j] ? nv - ivs[j] : 0) : ivs[j];
          rre += row[2 * iv] * kv[j];
          rim += (mir ? -row[2 * iv + 1] : row[2 * iv + 1]) * kv[j];
        }
        pre += rre * ku[i];
        pim += rim * ku[i];
      }
      re += pre * kw[q];
      im += pim * kw[q];
    }
    out[2 * k] = re;
    out[2 * k + 1] = im;
  }
}

/* Sampled explicit DFT (the truth of /root/reference/tests/test_hessian_approx.py:23-67) in OpenMP, for
 * full-size parity tests: out[p] = Re sum_k a_k exp(+2 pi i (u_k l_p + v_k m_p - w_k nm1_p)), with u, v, w
 * already in wavelengths and sign-flipped.  oracle/dft.py:dft_vis2dirty is the numpy statement of the same
 * sum; tests/test_oracle.py pins one against the other. */
void cw_dft_pixels(int64_t n, const double* u, const double* v, const double* w, const double* a /* 2n */,
                   int64_t npix, const double* l, const double* m, const double* nm1, double* out) {
  const double twopi = 6.283185307179586476925286766559;
  for (int64_t p = 0; p < npix; ++p) {
    double acc = 0.0;
    const double lp = l[p], mp = m[p], np_ = nm1[p];
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t k = 0; k < n; ++k) {
      double ph = u[k] * lp + v[k] * mp - w[k] * np_;
      ph -= rint(ph);
      const double c = cos(twopi * ph), s = sin(twopi * ph);
      acc += a[2 * k] * c - a[2 * k + 1] * s;
    }
    out[p] = acc;
  }
}

/* vis[k] = sum_p flux_p exp(-2 pi i (u_k l_p + v_k m_p - w_k nm1_p)) for n sampled (u, v, w). */
void cw_dft_rows(int64_t n, const double* u, const double* v, const double* w, int64_t npix, const double* l,
                 const double* m, const double* nm1, const double* flux, double* out /* 2n */) {
  const double twopi = 6.283185307179586476925286766559;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < n; ++k) {
    double re = 0.0, im = 0.0;
    for (int64_t p = 0; p < npix; ++p) {
      double ph = u[k] * l[p] + v[k] * m[p] - w[k] * nm1[p];
      ph -= rint(ph);
      re += flux[p] * cos(twopi * ph);
      im -= flux[p] * sin(twopi * ph);
    }
    out[2 * k] = re;
    out[2 * k + 1] = im;
  }
}
