// sj_cep.cu -- pulse parameters (carrier frequency, arrival time, carrier-envelope phase) of every monitor series, on the
// device, from the series the run left there.  This is the front half of the reference's post-processing class `signal`
// (scripts/phases.py:94-157, 205-244, 289-357), the part its phase sweeps evaluate for 1 600 series x phases in a Python
// loop: guess the arrival sample from the peaks of v^2 (:95-106), transform the series rolled by that guess (2 rfft, :318),
// take the magnitude-weighted mean frequency below twice the peak (:228-231), select the band around it whose magnitude
// exceeds CUT_ALPHA = 0.1 of the centre (:108-124), unwrap the phase over the band (fix_angle_seq, :62-74) and regress it on
// the frequency (:126-133): t0_corr = -slope / 2 pi, phi_corr = fix_angle(intercept).  The low-pass retry for noisy series
// (:277-287, 334-347) is detected (status bit 0) but not applied: the host falls back to the Python pipeline for those.
// One thread block per (monitor, field set) series.
#include <math.h>
#include <stdio.h>

#include "sj_internal.h"

#define CEP_COLS 10

static __device__ double block_sum(double v, double *red) {
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

// numpy.rint: round half to even
static __device__ double rint_even(double x) { return rint(x); }

// returns the angle transformed to (-pi, pi]  (phases.py:49-60)
static __device__ double fix_angle(double th) {
    const double pi = 3.141592653589793;
    if (th > pi) { th -= 2 * pi * floor(th / (2 * pi)); if (th > pi) th -= 2 * pi; }
    else if (th <= -pi) { th += 2 * pi * (floor(-th / (2 * pi)) + 1); if (th > pi) th -= 2 * pi; }
    return th;
}

__global__ void cep_kernel(const double *__restrict__ series, int n_mon, int n_sets, int n, double dt, double cut_alpha,
                           double *__restrict__ out) {
    extern __shared__ double sh[];
    const int nf = n / 2 + 1;
    double *v = sh, *re = v + n, *im = re + nf, *mag = im + nf, *red = mag + nf;      // red: 64 doubles
    __shared__ int s_t0, s_f0i, s_fmin, s_fmax, s_status;
    __shared__ double s_f0;
    const int mon = blockIdx.x % n_mon, set = blockIdx.x / n_mon;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = series[((long long)i * n_mon + mon) * n_sets + set];
    __syncthreads();
    // ---- _guess_t0_ind: |v|-weighted mean index over the peaks of v^2 (scipy.signal.find_peaks: strict rise, plateau
    // midpoints, strict fall) ----
    double num = 0.0, den = 0.0;
    for (int i = 1 + threadIdx.x; i < n - 1; i += blockDim.x) {
        const double x = v[i] * v[i], xl = v[i - 1] * v[i - 1];
        if (!(xl < x)) continue;
        int j = i;
        while (j + 1 < n - 1 && v[j + 1] * v[j + 1] == x) ++j;          // plateau
        if (j + 1 <= n - 1 && v[j + 1] * v[j + 1] < x) {
            const int mid = (i + j) / 2;
            num += (double)mid * fabs(v[mid]); den += fabs(v[mid]);
        }
    }
    num = block_sum(num, red); den = block_sum(den, red);
    if (threadIdx.x == 0) { s_t0 = den > 0.0 ? (int)(num / den) : 0; s_status = 0; }
    __syncthreads();
    const int t0 = s_t0;
    // ---- vf = 2 rfft(roll(v, -t0)): bin k = 2 sum_m v[(m + t0) mod n] exp(-2 pi i k m / n) ----
    for (int k = threadIdx.x; k < nf; k += blockDim.x) {
        double ar = 0.0, ai = 0.0;
        for (int m = 0; m < n; ++m) {
            const long long r = ((long long)k * m) % n;
            double sn, cs;
            sincospi(-2.0 * (double)r / (double)n, &sn, &cs);
            int src = m + t0; if (src >= n) src -= n;
            ar += v[src] * cs; ai += v[src] * sn;
        }
        re[k] = 2.0 * ar; im[k] = 2.0 * ai; mag[k] = hypot(2.0 * ar, 2.0 * ai);
    }
    __syncthreads();
    // ---- f0 ('avg' mode): argmax of |vf|, then trapz(mag f) / trapz(mag) over the bins below twice the peak ----
    if (threadIdx.x == 0) {
        const double df = 1.0 / (n * dt);
        int pk = 0;
        for (int k = 1; k < nf; ++k) if (mag[k] > mag[pk]) pk = k;          // numpy argmax: first maximum
        const int m = min(2 * pk, nf);
        double a = 0.0, b = 0.0;
        for (int k = 0; k + 1 < m; ++k) { a += 0.5 * (mag[k] * (k * df) + mag[k + 1] * ((k + 1) * df)); b += 0.5 * (mag[k] + mag[k + 1]); }
        double f0 = b != 0.0 ? a / b : 0.0;
        int f0i = (int)rint_even(f0 / df);
        if (f0i >= nf) f0i = nf - 1;
        if (f0i < 0) f0i = 0;
        s_f0 = f0; s_f0i = f0i;
        // ---- is_noisy (phases.py:277-287): a second prominent peak above three times the centre bin.  Prominence of a
        // peak = its height above the higher of the two lowest points between it and the next higher sample on each side.
        {
            const double need = 0.1 * mag[f0i];
            int n_pk = 0, high = 0;
            for (int k = 1; k < nf - 1; ++k) {
                if (!(mag[k - 1] < mag[k])) continue;
                int j = k; while (j + 1 < nf - 1 && mag[j + 1] == mag[k]) ++j;
                if (!(mag[j + 1] < mag[k])) { k = j; continue; }
                const int mid = (k + j) / 2;
                double lmin = mag[mid], rmin = mag[mid];
                for (int q = mid - 1; q >= 0 && mag[q] <= mag[mid]; --q) lmin = fmin(lmin, mag[q]);
                for (int q = mid + 1; q < nf && mag[q] <= mag[mid]; ++q) rmin = fmin(rmin, mag[q]);
                if (mag[mid] - fmax(lmin, rmin) >= need) { ++n_pk; if (mid > 3 * f0i) high = 1; }
                k = j;
            }
            if (n_pk >= 2 && high) s_status |= 1;
            if (f0i > nf / 2) s_status |= 2;               // suspiciously high centre frequency (phases.py:337-338)
        }
        // ---- _get_freq_fwhm (phases.py:108-124) ----
        const double cut = mag[f0i] * cut_alpha;
        int fmin_ = f0i, fmax_ = f0i;
        const int jb = min(f0i, nf - f0i - 1);
        for (int j = 0; j < jb; ++j) {
            if (mag[f0i - j] > cut) fmin_ = f0i - j - 1;
            if (mag[f0i + j] > cut) fmax_ = f0i + j + 1;
        }
        s_fmin = fmin_; s_fmax = fmax_;
    }
    __syncthreads();
    // ---- _param_est (phases.py:126-155): unwrap arg(vf) over the band, linear regression on the frequency; if the fitted
    // delay is large against the band the series is rolled by it as well and the fit repeated once (MAX_PARAM_EVALS = 1) ----
    __shared__ double s_slope, s_icpt, s_rval;
    __shared__ int s_again, s_t0b;
    for (int pass = 0; pass < 2; ++pass) {
        if (threadIdx.x == 0) {
            const double pi = 3.141592653589793, df = 1.0 / (n * dt);
            const int a = s_fmin, b = s_fmax, m = b - a;
            double slope = 0.0, icpt = 0.0, rval = 0.0;
            if (m >= 2) {
                double sx = 0, sy = 0, prev = 0, shift = 0;
                for (int q = 0; q < m; ++q) {
                    double ang = atan2(im[a + q], re[a + q]) + shift;
                    if (q > 0) {    // fix_angle_seq: a step of +-2 pi that brings the angle closer is applied to the rest of the sequence
                        const double d0 = fabs(ang - prev);
                        if (fabs(ang + 2 * pi - prev) < d0) { ang += 2 * pi; shift += 2 * pi; }
                        else if (fabs(ang - 2 * pi - prev) < d0) { ang -= 2 * pi; shift -= 2 * pi; }
                    }
                    prev = ang; sx += (a + q) * df; sy += ang;
                }
                const double mx = sx / m, my = sy / m;
                double sxx = 0, sxy = 0, syy = 0; prev = 0; shift = 0;
                for (int q = 0; q < m; ++q) {
                    double ang = atan2(im[a + q], re[a + q]) + shift;
                    if (q > 0) {
                        const double d0 = fabs(ang - prev);
                        if (fabs(ang + 2 * pi - prev) < d0) { ang += 2 * pi; shift += 2 * pi; }
                        else if (fabs(ang - 2 * pi - prev) < d0) { ang -= 2 * pi; shift -= 2 * pi; }
                    }
                    prev = ang;
                    const double dx = (a + q) * df - mx, dy = ang - my;
                    sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
                }
                slope = sxx != 0.0 ? sxy / sxx : 0.0;
                icpt = my - slope * mx;
                rval = (sxx > 0 && syy > 0) ? sxy / sqrt(sxx * syy) : 0.0;
            } else s_status |= 4;
            s_slope = slope; s_icpt = icpt; s_rval = rval;
            const double t0c = -slope / (2 * pi);
            s_again = (pass == 0 && m >= 2 && fabs(t0c) > 2.0 / (double)m) ? 1 : 0;
            s_t0b = t0 + (int)(t0c / dt);              // int(): toward zero
        }
        __syncthreads();
        if (!s_again) break;
        // the band bins of the series rolled by the new arrival sample (np.roll: any shift, modulo n)
        int sh = s_t0b % n; if (sh < 0) sh += n;
        for (int k = s_fmin + (int)threadIdx.x; k < s_fmax; k += blockDim.x) {
            double ar = 0.0, ai = 0.0;
            for (int m = 0; m < n; ++m) {
                const long long r = ((long long)k * m) % n;
                double sn, cs;
                sincospi(-2.0 * (double)r / (double)n, &sn, &cs);
                int src = m + sh; if (src >= n) src -= n;
                ar += v[src] * cs; ai += v[src] * sn;
            }
            re[k] = 2.0 * ar; im[k] = 2.0 * ai;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double pi = 3.141592653589793;
        double *o = out + (long long)blockIdx.x * CEP_COLS;
        o[0] = s_f0; o[1] = (double)s_f0i; o[2] = (double)t0; o[3] = -s_slope / (2 * pi); o[4] = fix_angle(s_icpt);
        o[5] = s_slope; o[6] = s_icpt; o[7] = s_rval; o[8] = (double)s_fmin + 65536.0 * (double)s_fmax; o[9] = (double)s_status;
    }
}

extern "C" int sj_extract_cep(sj_sim *s, double dt_sample, double *out) {
    if (!s || !out || !(dt_sample > 0)) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    const int n = s->n_samples, nf = n / 2 + 1;
    if (n < 8 || s->n_mon == 0) { s->err = "sj_extract_cep needs monitors and at least 8 samples"; return SJ_ERR_STATE; }
    const size_t smem = ((size_t)n + 3 * (size_t)nf + 64) * sizeof(double);
    if (smem > 200 * 1024) { s->err = "series too long for the on-device extraction (about 10 000 samples)"; return SJ_ERR_UNSUPPORTED; }
    CK(cudaFuncSetAttribute(cep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_series = s->n_mon * s->g.n_sets;
    double *dev = NULL;
    CK(cudaMalloc((void **)&dev, (size_t)n_series * CEP_COLS * sizeof(double)));
    cep_kernel<<<n_series, 256, smem, s->stream>>>(s->series, s->n_mon, s->g.n_sets, n, dt_sample, 0.1, dev);
    s->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, dev, (size_t)n_series * CEP_COLS * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) { s->err = cudaGetErrorString(e); return SJ_ERR_CUDA; }
    return SJ_OK;
}
