/*
 * pfbsara.h — C ABI of the device SARA backward step: the wavelet dictionary
 * Psi (analysis `dot`, synthesis `hdot`), the fused l21 dual update, and the element-wise pieces of the
 * primal-dual iteration.  Same conventions as pfbgrid.h: every entry point returns 0 or a PFBG_ERR_* code,
 * pfbg_last_error() holds the message, plain pointers and sizes only.
 *
 * Reference interfaces replaced (paths relative to /root/reference/src/pfb_imaging):
 *   - Psi / PsiNocopyt .dot / .hdot  (operators/psi.py:549-664; PsiBand :231-372, PsiBandNocopyt :466-530;
 *     transforms wavelets/wavelets.py:40-343, convolutions wavelets/convolutions.py)        -> pfbs_psi_dot / _hdot
 *   - dual_update_numba_fast (prox/prox_21m.py:104-135)                                     -> pfbs_dual_update
 *   - prox_21m_numba (prox/prox_21m.py:30-62)                                               -> pfbs_prox_21m
 *   - _nb_extrapolate_dual, _nb_primal_step, _nb_norm_diff (opt/primal_dual.py:16-61),
 *     positivity / positivity_band (prox/positivity.py:13-38)                               -> pfbs_extrapolate /
 *                                                                                              pfbs_primal_step / pfbs_norm_diff
 */
#ifndef PFBSARA_H
#define PFBSARA_H

#include <stdint.h>

#include "pfbgrid.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PFBS_KMAX 10          /* longest filter (db5) */
#define PFBS_TRANSPOSED 8u    /* coefficient arrays are (…, nymax, nxmax) (`Psi`) instead of (…, nxmax, nymax) (`PsiNocopyt`) */

typedef struct pfbs_psi pfbs_psi;

/*
 * One dictionary for `nband` images of (nx, ny).  Per basis b: K[b] = filter length (0 = 'self', the identity),
 * filters[b][4][PFBS_KMAX] = dec_lo, dec_hi, rec_lo, rec_hi (zero padded).  The packing arrays are those of
 * operators/psi.py:24-137: ix, iy (nbasis, nlevel, 2); sx, sy, spx, spy (nbasis, nlevel); ntotx, ntoty (nbasis).
 */
int pfbs_psi_create(int32_t precision, int32_t device, int32_t nband, int32_t nx, int32_t ny, int32_t nbasis,
                    int32_t nlevel, const int32_t* K, const double* filters, const int64_t* ix, const int64_t* iy,
                    const int64_t* sx, const int64_t* sy, const int64_t* spx, const int64_t* spy,
                    const int64_t* ntotx, const int64_t* ntoty, int32_t nxmax, int32_t nymax, pfbs_psi** out);
int pfbs_psi_destroy(pfbs_psi* psi);
/* x (nband, nx, ny) -> alpha (nband, nbasis, nxmax, nymax); alpha is fully overwritten (unused cells = 0). */
int pfbs_psi_dot(pfbs_psi* psi, const void* x, void* alpha, uint32_t flags, void* stream);
/* alpha -> x (nband, nx, ny) = sum over bases of the inverse transforms (overwritten). */
int pfbs_psi_hdot(pfbs_psi* psi, const void* alpha, void* x, uint32_t flags, void* stream);

/*
 * Fused dual update on device arrays: v (nband, ncoef) holds Psi^T xp on entry;
 *   vtilde = vp + sigma v;  s = |sum_band vtilde|;  v = vtilde * min(1, lam w / s)
 * phase 0: all in one kernel (every band on this device).
 * phase 1: v = vtilde and bsum (ncoef) = sum over the LOCAL bands;   [caller all-reduces bsum across ranks]
 * phase 2: v *= min(1, lam w / |bsum|).
 * vbar (optional, phases 0 and 2; needs vp): also receives the extrapolated dual 2 v - vp (_nb_extrapolate_dual,
 * opt/primal_dual.py:16-23) in the same pass.
 */
int pfbs_dual_update(int32_t precision, int32_t device, const void* vp, void* v, const void* weight, double lam,
                     double sigma, int32_t nband, int64_t ncoef, void* bsum, int32_t phase, void* vbar, void* stream);
/* result = v * max(|sum_band v / sigma| - lam w / sigma, 0) / |sum_band v / sigma| / sigma   (0 where the sum is 0) */
int pfbs_prox_21m(int32_t precision, int32_t device, const void* v, void* result, const void* weight, double lam,
                  double sigma, int32_t nband, int64_t ncoef, void* stream);
/* out = a x + b y on device arrays (out may alias x or y): the element-wise steps of the forward-backward loop
 * (opt/forward_backward.py:84-121) */
int pfbs_axpby(int32_t precision, int32_t device, void* out, double a, const void* x, double b, const void* y,
               int64_t n, void* stream);
/* vp = 2 v - vp */
int pfbs_extrapolate(int32_t precision, int32_t device, const void* v, void* vp, int64_t n, void* stream);
/* x = xp - tau * xout, then positivity: 0 none, 1 clamp negatives, 2 zero a pixel in all (local) bands if any band <= 0 */
int pfbs_primal_step(int32_t precision, int32_t device, void* x, const void* xp, const void* xout, double tau,
                     int32_t positivity, int32_t nband, int64_t npix, void* stream);
/* num_den[0] = sum (x - xp)^2, num_den[1] = sum x^2 (host output; synchronises the stream) */
int pfbs_norm_diff(int32_t precision, int32_t device, const void* x, const void* xp, int64_t n, double* num_den,
                   void* stream);

/* out2[0] = <a, b>, out2[1] = <c, d> on device arrays (host output; synchronises the stream): the scalar products of
 * the device-resident conjugate gradients (opt/pcg.py:35-41, 77-85) */
int pfbs_dot2(int32_t precision, int32_t device, const void* a, const void* b, const void* c, const void* d, int64_t n,
              double* out2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PFBSARA_H */
