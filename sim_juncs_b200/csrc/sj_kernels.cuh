// sj_kernels.cuh -- sm_100a kernels of the FDTD hot path.
//
// Two passes per time step, every field array touched once per pass:
//   H-pass : B -= C curl E (M1) fused with H = mu^-1 B and the UPML auxiliaries (M2)
//   E-pass : D += C curl H (M3) + source injection (M4) + Drude-Lorentz ADE (M5) +
//            E = chi1inv (D - sum P) (M6), UPML auxiliaries fused
// Interior cells (all PML sigmas zero) run the *_interior kernels: z-marching 2.5-D tiles,
// 128-bit row loads, the k+-1 plane carried in registers, i+-1 neighbours by warp shuffle.  They
// keep only E and H: E^{n+1} = E^n + chi1inv (dD - dP - dS), algebraically meep's
// E = chi1inv (D - P - S) with D eliminated.  The PML shell (six boxes) runs the *_pml kernels,
// which hold D/B (and U where two sigmas overlap) in compact per-box arrays and follow meep's
// step_curl / step_update_EDHB formulas.
#pragma once
#include "sj_internal.h"

template <typename T, int V> struct VecOf;
template <> struct VecOf<double, 2> { typedef double2 type; };
template <> struct VecOf<float, 4> { typedef float4 type; };
template <> struct VecOf<float, 2> { typedef float2 type; };

template <typename T, int V>
struct Vec {
    T v[V];
    __device__ __forceinline__ void load(const T *p) {
        typedef typename VecOf<T, V>::type VT;
        VT t = *reinterpret_cast<const VT *>(p);
        const T *q = reinterpret_cast<const T *>(&t);
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = q[i];
    }
    __device__ __forceinline__ void store(T *p) const {
        typedef typename VecOf<T, V>::type VT;
        VT t;
        T *q = reinterpret_cast<T *>(&t);
#pragma unroll
        for (int i = 0; i < V; ++i) q[i] = v[i];
        *reinterpret_cast<VT *>(p) = t;
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = T(0);
    }
};

// Source amplitude at a Yee point of component c: amp * wx * wy * wz, for set q the real drive
// pair {dS, dt*J}.  Returns the D-side increment  -(A dS) - (A dtJ)  to add to dD.
template <typename T>
__device__ __forceinline__ T source_term(const KParams<T> &p, int c, int i, int j, int k, int set, long long step) {
    T acc = T(0);
    for (int s = 0; s < p.n_src; ++s) {
        const SrcDev<T> &g = p.src[s];
        if (g.comp != c) continue;
        if (k < g.lo[2] || k > g.hi[2] || j < g.lo[1] || j > g.hi[1] || i < g.lo[0] || i > g.hi[0]) continue;
        const T wgt = g.w[0][i - g.lo[0]] * g.w[1][j - g.lo[1]] * g.w[2][k - g.lo[2]];
        const T *d0 = p.drive + ((step * p.n_src + s) * p.n_sets + set) * 2;
        const T *d1 = d0 + (long long)p.n_src * p.n_sets * 2;
        // integrated sources: S_{n+1} - S_n ; current sources: dt*J_n
        acc -= wgt * ((d1[0] - d0[0]) + d0[1]);
    }
    return acc;
}
// Same, but split into the "S" part at step n, at n+1 and the current kick (PML kernels need W_old).
template <typename T>
__device__ __forceinline__ void source_parts(const KParams<T> &p, int c, int i, int j, int k, int set, long long step,
                                             T &S0, T &S1, T &J) {
    S0 = S1 = J = T(0);
    for (int s = 0; s < p.n_src; ++s) {
        const SrcDev<T> &g = p.src[s];
        if (g.comp != c) continue;
        if (k < g.lo[2] || k > g.hi[2] || j < g.lo[1] || j > g.hi[1] || i < g.lo[0] || i > g.hi[0]) continue;
        const T wgt = g.w[0][i - g.lo[0]] * g.w[1][j - g.lo[1]] * g.w[2][k - g.lo[2]];
        const T *d0 = p.drive + ((step * p.n_src + s) * p.n_sets + set) * 2;
        const T *d1 = d0 + (long long)p.n_src * p.n_sets * 2;
        S0 += wgt * d0[0]; S1 += wgt * d1[0]; J += wgt * d0[1];
    }
}

// ADE update of all poles of material m at linear index x (M5): returns sum(P_new - P_cur) and
// optionally sum P_cur / sum P_new.  `drive` is E^n (W^n inside the PML).
template <typename T>
__device__ __forceinline__ T ade_update(const KParams<T> &p, int c, int m, long long x, int parity, T drive,
                                        T *sum_old, T *sum_new) {
    const int np = p.mt_np[m];
    T dP = T(0), so = T(0), sn = T(0);
    for (int s = 0; s < np; ++s) {
        T *cur = parity ? p.PB[s][c] : p.PA[s][c];
        T *prv = parity ? p.PA[s][c] : p.PB[s][c];
        const T *cf = p.mt_coef + ((long long)m * SJ_MAX_POLES + s) * 3;
        const T pc = cur[x], pp = prv[x];
        const T pn = cf[0] * pc + cf[1] * pp + cf[2] * drive;
        prv[x] = pn;  // becomes "current" after the parity flip
        dP += pn - pc; so += pc; sn += pn;
    }
    if (sum_old) *sum_old = so;
    if (sum_new) *sum_new = sn;
    return dP;
}

// ------------------------------------------------------------------------------------------
// Interior H-pass.  Block = 32 lanes (x, V cells each) x blockDim.y rows; marches k over a chunk.
// ------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) h_interior(KParams<T> p, int i_lo, int i_hi, int j_lo, int j_hi, int k_lo,
                                                  int k_hi, int zchunk, int nzc) {
    const int lane = threadIdx.x;
    const int i0 = i_lo + (blockIdx.x * 32 + lane) * V;
    const int j = j_lo + blockIdx.y * blockDim.y + threadIdx.y;
    const int set = blockIdx.z / nzc;
    const int kb = k_lo + (blockIdx.z % nzc) * zchunk;
    const int ke = min(kb + zchunk, k_hi);
    if (j >= j_hi) return;  // whole warp shares j: warp-uniform exit keeps shuffles legal
    const bool ld = (i0 + V <= p.pitch);
    const bool st = (i0 < i_hi);
    const T C = p.courant;
    const long long so = (long long)set * p.set_stride;
    const T *Ex = p.E[0] + so, *Ey = p.E[1] + so, *Ez = p.E[2] + so;
    T *Hx = p.H[0] + so, *Hy = p.H[1] + so, *Hz = p.H[2] + so;
    long long x = (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;

    Vec<T, V> ex0, ey0, ex1, ey1, ez0, ezj, exj, hx, hy, hz;
    if (ld) { ex0.load(Ex + x); ey0.load(Ey + x); } else { ex0.zero(); ey0.zero(); }
    for (int k = kb; k < ke; ++k, x += p.plane) {
        if (ld) {
            ex1.load(Ex + x + p.plane); ey1.load(Ey + x + p.plane);
            ez0.load(Ez + x); ezj.load(Ez + x + p.pitch); exj.load(Ex + x + p.pitch);
        } else { ex1.zero(); ey1.zero(); ez0.zero(); ezj.zero(); exj.zero(); }
        if (st) { hx.load(Hx + x); hy.load(Hy + x); hz.load(Hz + x); }
        T ez_n = __shfl_down_sync(0xffffffffu, ez0.v[0], 1);
        T ey_n = __shfl_down_sync(0xffffffffu, ey0.v[0], 1);
        if (lane == 31 && st) {
            const bool ok = (i0 + V < p.pitch);
            ez_n = ok ? Ez[x + V] : T(0);
            ey_n = ok ? Ey[x + V] : T(0);
        }
        if (st) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T ezi = (v < V - 1) ? ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                hx.v[v] -= C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1.v[v]);
                hy.v[v] -= C * (((ex1.v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
                hz.v[v] -= C * (((eyi - ey0.v[v]) + ex0.v[v]) - exj.v[v]);
            }
            hx.store(Hx + x); hy.store(Hy + x); hz.store(Hz + x);
        }
        ex0 = ex1; ey0 = ey1;
    }
}

// ------------------------------------------------------------------------------------------
// Interior E-pass.
// ------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) e_interior(KParams<T> p, int i_lo, int i_hi, int j_lo, int j_hi, int k_lo,
                                                  int k_hi, int zchunk, int nzc) {
    const int lane = threadIdx.x;
    const int i0 = i_lo + (blockIdx.x * 32 + lane) * V;
    const int j = j_lo + blockIdx.y * blockDim.y + threadIdx.y;
    const int set = blockIdx.z / nzc;
    const int kb = k_lo + (blockIdx.z % nzc) * zchunk;
    const int ke = min(kb + zchunk, k_hi);
    if (j >= j_hi) return;
    const bool ld = (i0 + V <= p.pitch);
    const bool st = (i0 < i_hi);
    const T C = p.courant;
    const long long step = *p.step;
    const int parity = (int)(step & 1);
    const long long so = (long long)set * p.set_stride;
    T *Ex = p.E[0] + so, *Ey = p.E[1] + so, *Ez = p.E[2] + so;
    const T *Hx = p.H[0] + so, *Hy = p.H[1] + so, *Hz = p.H[2] + so;
    long long x = (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;

    Vec<T, V> hxm, hym, hx0, hy0, hz0, hzj, hxj, ex, ey, ez;
    if (ld) { hxm.load(Hx + x - p.plane); hym.load(Hy + x - p.plane); } else { hxm.zero(); hym.zero(); }
    for (int k = kb; k < ke; ++k, x += p.plane) {
        if (ld) {
            hx0.load(Hx + x); hy0.load(Hy + x); hz0.load(Hz + x);
            hzj.load(Hz + x - p.pitch); hxj.load(Hx + x - p.pitch);
        } else { hx0.zero(); hy0.zero(); hz0.zero(); hzj.zero(); hxj.zero(); }
        unsigned char mx[V], my[V], mz[V];
        if (st) {
            ex.load(Ex + x); ey.load(Ey + x); ez.load(Ez + x);
#pragma unroll
            for (int v = 0; v < V; ++v) { mx[v] = p.mat[0][x + v]; my[v] = p.mat[1][x + v]; mz[v] = p.mat[2][x + v]; }
        }
        T hz_p = __shfl_up_sync(0xffffffffu, hz0.v[V - 1], 1);
        T hy_p = __shfl_up_sync(0xffffffffu, hy0.v[V - 1], 1);
        if (lane == 0 && st) {
            hz_p = (i0 > 0) ? Hz[x - 1] : T(0);
            hy_p = (i0 > 0) ? Hy[x - 1] : T(0);
        }
        if (st) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                T dDx = -(C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym.v[v]));
                T dDy = -(C * (((hxm.v[v] - hx0.v[v]) + hz0.v[v]) - hzi));
                T dDz = -(C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]));
                const int i = i0 + v;
                if (p.n_src) {
                    dDx += source_term(p, 0, i, j, k, set, step);
                    dDy += source_term(p, 1, i, j, k, set, step);
                    dDz += source_term(p, 2, i, j, k, set, step);
                }
                const long long xv = so + x + v;
                {
                    const int m = mx[v];
                    T dP = T(0);
                    if (p.mt_np[m]) dP = ade_update(p, 0, m, xv, parity, ex.v[v], (T *)0, (T *)0);
                    ex.v[v] += p.mt_chi[m] * (dDx - dP);
                }
                {
                    const int m = my[v];
                    T dP = T(0);
                    if (p.mt_np[m]) dP = ade_update(p, 1, m, xv, parity, ey.v[v], (T *)0, (T *)0);
                    ey.v[v] += p.mt_chi[m] * (dDy - dP);
                }
                {
                    const int m = mz[v];
                    T dP = T(0);
                    if (p.mt_np[m]) dP = ade_update(p, 2, m, xv, parity, ez.v[v], (T *)0, (T *)0);
                    ez.v[v] += p.mt_chi[m] * (dDz - dP);
                }
            }
            ex.store(Ex + x); ey.store(Ey + x); ez.store(Ez + x);
        }
        hxm = hx0; hym = hy0;
    }
}

// ------------------------------------------------------------------------------------------
// PML shell kernels: one thread per cell of a box, general meep formulas.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T pml_db(T fold, T curl, T sk, T su, T *U, long long xb, T &fnew_out) {
    // returns new f; meep step_curl general branch with kappa = 1
    T fnew;
    if (su != T(0) && sk != T(0)) {
        const T uold = U[xb];
        const T unew = ((T(1) - sk) * uold - curl) / (T(1) + sk);
        U[xb] = unew;
        fnew = ((T(1) - su) * fold + (unew - uold)) / (T(1) + su);
    } else if (su != T(0)) {
        fnew = ((T(1) - su) * fold - curl) / (T(1) + su);
    } else {
        fnew = ((T(1) - sk) * fold - curl) / (T(1) + sk);
    }
    fnew_out = fnew;
    return fnew;
}

template <typename T>
__global__ void __launch_bounds__(256) h_pml(KParams<T> p, PmlBox<T> b, int k_lo, int k_hi) {
    const long long ncell = (long long)b.bx * b.by * (k_hi - k_lo);
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell * p.n_sets) return;
    const int set = (int)(t / ncell);
    t -= (long long)set * ncell;
    const int li = (int)(t % b.bx);
    const int lj = (int)((t / b.bx) % b.by);
    const int lk = (int)(t / ((long long)b.bx * b.by));
    const int i = b.lo[0] + li, j = b.lo[1] + lj, k = k_lo + lk;
    const int idx[3] = {i, j, k};
    const long long so = (long long)set * p.set_stride;
    const long long x = so + (long long)(k - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i;
    const long long xb = (long long)set * b.bset + (long long)(k - b.lo[2]) * b.bplane + (long long)lj * b.bpitch + li;
    const long long str[3] = {1, p.pitch, p.plane};
    const T C = p.courant;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d1 = (c + 1) % 3, d2 = (c + 2) % 3;
        // owned + non-metal range of H component c
        if (idx[c] < 1 || idx[c] > p.n[c] - 1 || idx[d1] > p.n[d1] - 1 || idx[d2] > p.n[d2] - 1) continue;
        const T *g1 = p.E[d2], *g2 = p.E[d1];
        const T curl = C * (((g1[x + str[d1]] - g1[x]) + g2[x]) - g2[x + str[d2]]);
        const T sk = p.sig[d1][2 * idx[d1] + 1], su = p.sig[d2][2 * idx[d2] + 1], sw = p.sig[c][2 * idx[c]];
        const T bold = b.B[c][xb];
        T bnew;
        pml_db(bold, curl, sk, su, b.UB[c], xb, bnew);
        b.B[c][xb] = bnew;
        T *H = p.H[c];
        if (sw != T(0)) H[x] += (T(1) + sw) * bnew - (T(1) - sw) * bold;
        else H[x] = bnew;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) e_pml(KParams<T> p, PmlBox<T> b, int k_lo, int k_hi) {
    const long long ncell = (long long)b.bx * b.by * (k_hi - k_lo);
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell * p.n_sets) return;
    const int set = (int)(t / ncell);
    t -= (long long)set * ncell;
    const int li = (int)(t % b.bx);
    const int lj = (int)((t / b.bx) % b.by);
    const int lk = (int)(t / ((long long)b.bx * b.by));
    const int i = b.lo[0] + li, j = b.lo[1] + lj, k = k_lo + lk;
    const int idx[3] = {i, j, k};
    const long long so = (long long)set * p.set_stride;
    const long long xl = (long long)(k - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i;
    const long long x = so + xl;
    const long long xb = (long long)set * b.bset + (long long)(k - b.lo[2]) * b.bplane + (long long)lj * b.bpitch + li;
    const long long str[3] = {1, p.pitch, p.plane};
    const T C = p.courant;
    const long long step = *p.step;
    const int parity = (int)(step & 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d1 = (c + 1) % 3, d2 = (c + 2) % 3;
        if (idx[c] > p.n[c] - 1 || idx[d1] < 1 || idx[d1] > p.n[d1] - 1 || idx[d2] < 1 || idx[d2] > p.n[d2] - 1) continue;
        const T *g1 = p.H[d2], *g2 = p.H[d1];
        const T curl = C * (((g1[x - str[d1]] - g1[x]) + g2[x]) - g2[x - str[d2]]);
        const T sk = p.sig[d1][2 * idx[d1]], su = p.sig[d2][2 * idx[d2]], sw = p.sig[c][2 * idx[c] + 1];
        const T dold = b.D[c][xb];
        T dnew;
        pml_db(dold, curl, sk, su, b.UD[c], xb, dnew);
        T S0 = T(0), S1 = T(0), J = T(0);
        if (p.n_src) source_parts(p, c, i, j, k, set, step, S0, S1, J);
        dnew -= J;
        b.D[c][xb] = dnew;
        const int m = p.mat[c][xl];
        const T chi = p.mt_chi[m];
        T pold = T(0), pnew = T(0);
        T wold = chi * (dold - S0);
        if (p.mt_np[m]) {
            // W^n = chi (D^n - sum P^n - S^n) needs sum P^n before the poles advance
            T so_ = T(0);
            const int np = p.mt_np[m];
            for (int s = 0; s < np; ++s) so_ += (parity ? p.PB[s][c] : p.PA[s][c])[x];
            wold = chi * ((dold - so_) - S0);
            ade_update(p, c, m, x, parity, wold, &pold, &pnew);
        }
        const T wnew = chi * ((dnew - pnew) - S1);
        T *E = p.E[c];
        if (sw != T(0)) E[x] += (T(1) + sw) * wnew - (T(1) - sw) * wold;
        else E[x] = wnew;
    }
}


// ------------------------------------------------------------------------------------------
// PML shell, tiled: same z-marching / 128-bit / shuffle structure as the interior kernels, plus
// the box-local auxiliary arrays (D or B always, U where two sigmas overlap) and meep's
// step_curl / step_update_EDHB arithmetic with the sig / siginv tables.  LX = lanes of a warp
// along x (32 for the wide z- and y-boxes, 8 for the narrow x-boxes: 4 rows per warp).
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T pml_step_db(T fold, T curl, T sk, T ik, T su, T iu, T *U, long long xb) {
    if (su != T(0) && sk != T(0)) {
        const T uo = U[xb];
        const T un = ((T(1) - sk) * uo - curl) * ik;
        U[xb] = un;
        return ((T(1) - su) * fold + (un - uo)) * iu;
    }
    if (su != T(0)) return ((T(1) - su) * fold - curl) * iu;
    return ((T(1) - sk) * fold - curl) * ik;
}

template <typename T>
__device__ __forceinline__ void pml_e_elem(const KParams<T> &p, int c, T &e, T &d, T curl, T sk, T ik, T su, T iu, T sw,
                                           T *U, long long xb, long long x, long long xl, int i, int j, int k, int set,
                                           long long step, int parity) {
    const T dold = d;
    T dnew = pml_step_db(dold, curl, sk, ik, su, iu, U, xb);
    T S0 = T(0), S1 = T(0), J = T(0);
    if (p.n_src) source_parts(p, c, i, j, k, set, step, S0, S1, J);
    dnew -= J;
    d = dnew;
    const int m = p.mat[c][xl];
    const T chi = p.mt_chi[m];
    T pold = T(0), pnew = T(0);
    T wold = chi * (dold - S0);
    const int np = p.mt_np[m];
    if (np) {
        T so_ = T(0);
        for (int s = 0; s < np; ++s) so_ += (parity ? p.PB[s][c] : p.PA[s][c])[x];
        wold = chi * ((dold - so_) - S0);
        ade_update(p, c, m, x, parity, wold, &pold, &pnew);
    }
    const T wnew = chi * ((dnew - pnew) - S1);
    e = (sw != T(0)) ? e + (T(1) + sw) * wnew - (T(1) - sw) * wold : wnew;
}

template <typename T, int V, int LX>
__global__ void __launch_bounds__(256) h_pml_tile(KParams<T> p, PmlBoxSet<T> bs, const WorkItem *__restrict__ items,
                                                  int k_lo, int k_hi) {
    const WorkItem it = items[blockIdx.x];
    const PmlBox<T> &b = bs.b[it.box];
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = it.i0 + lx * V;
    const int j = it.j0 + warp * RW + ly;
    const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
    if (kb >= ke) return;
    const bool act = (j < b.hi[1]) && (i0 < b.hi[0]);
    const bool rowp = act && (j + 1 <= p.n[1]);
    const T C = p.courant;
    const int set = it.set;
    const long long so = (long long)set * p.set_stride;
    const T *Ex = p.E[0] + so, *Ey = p.E[1] + so, *Ez = p.E[2] + so;
    T *Hx = p.H[0] + so, *Hy = p.H[1] + so, *Hz = p.H[2] + so;
    long long x = (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;
    long long xb = (long long)set * b.bset + (long long)(kb - b.lo[2]) * b.bplane + (long long)(j - b.lo[1]) * b.bpitch + (i0 - b.lo[0]);
    // per-thread PML coefficients along x (per element) and y
    T sxi[V], ixi[V], sxh[V], ixh[V];
    T syi = T(0), iyi = T(1), syh = T(0), iyh = T(1);
    if (act) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int h = 2 * (i0 + v);
            sxi[v] = p.sig[0][h]; ixi[v] = p.siginv[0][h]; sxh[v] = p.sig[0][h + 1]; ixh[v] = p.siginv[0][h + 1];
        }
        syi = p.sig[1][2 * j]; iyi = p.siginv[1][2 * j]; syh = p.sig[1][2 * j + 1]; iyh = p.siginv[1][2 * j + 1];
    }
    Vec<T, V> ex0, ey0, ex1, ey1, ez0, ezj, exj, hx, hy, hz, bx, by, bz;
    if (act) { ex0.load(Ex + x); ey0.load(Ey + x); } else { ex0.zero(); ey0.zero(); }
    for (int k = kb; k < ke; ++k, x += p.plane, xb += b.bplane) {
        if (act) {
            ex1.load(Ex + x + p.plane); ey1.load(Ey + x + p.plane); ez0.load(Ez + x);
            hx.load(Hx + x); hy.load(Hy + x); hz.load(Hz + x);
            bx.load(b.B[0] + xb); by.load(b.B[1] + xb); bz.load(b.B[2] + xb);
        } else { ex1.zero(); ey1.zero(); ez0.zero(); }
        if (rowp) { ezj.load(Ez + x + p.pitch); exj.load(Ex + x + p.pitch); } else { ezj.zero(); exj.zero(); }
        T ez_n = __shfl_down_sync(0xffffffffu, ez0.v[0], 1, LX);
        T ey_n = __shfl_down_sync(0xffffffffu, ey0.v[0], 1, LX);
        if (act && (lx == LX - 1 || i0 + V >= b.hi[0])) {
            const bool ok = (i0 + V < p.pitch);
            ez_n = ok ? Ez[x + V] : T(0);
            ey_n = ok ? Ey[x + V] : T(0);
        }
        if (act) {
            const T szi = p.sig[2][2 * k], szh = p.sig[2][2 * k + 1], izh = p.siginv[2][2 * k + 1];
            const bool kok = (k <= p.n[2] - 1), jok = (j <= p.n[1] - 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int i = i0 + v;
                const T ezi = (v < V - 1) ? ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                const bool iok = (i <= p.n[0] - 1);
                if (i >= 1 && iok && jok && kok) {       // Hx: k-dir y, u-dir z, w-dir x
                    const T curl = C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1.v[v]);
                    const T bo = bx.v[v];
                    const T bn = pml_step_db(bo, curl, syh, iyh, szh, izh, b.UB[0], xb + v);
                    bx.v[v] = bn;
                    hx.v[v] = (sxi[v] != T(0)) ? hx.v[v] + (T(1) + sxi[v]) * bn - (T(1) - sxi[v]) * bo : bn;
                }
                if (j >= 1 && jok && iok && kok) {       // Hy: k-dir z, u-dir x, w-dir y
                    const T curl = C * (((ex1.v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
                    const T bo = by.v[v];
                    const T bn = pml_step_db(bo, curl, szh, izh, sxh[v], ixh[v], b.UB[1], xb + v);
                    by.v[v] = bn;
                    hy.v[v] = (syi != T(0)) ? hy.v[v] + (T(1) + syi) * bn - (T(1) - syi) * bo : bn;
                }
                if (k >= 1 && kok && iok && jok) {       // Hz: k-dir x, u-dir y, w-dir z
                    const T curl = C * (((eyi - ey0.v[v]) + ex0.v[v]) - exj.v[v]);
                    const T bo = bz.v[v];
                    const T bn = pml_step_db(bo, curl, sxh[v], ixh[v], syh, iyh, b.UB[2], xb + v);
                    bz.v[v] = bn;
                    hz.v[v] = (szi != T(0)) ? hz.v[v] + (T(1) + szi) * bn - (T(1) - szi) * bo : bn;
                }
            }
            hx.store(Hx + x); hy.store(Hy + x); hz.store(Hz + x);
            bx.store(b.B[0] + xb); by.store(b.B[1] + xb); bz.store(b.B[2] + xb);
        }
        ex0 = ex1; ey0 = ey1;
    }
    (void)iyi; (void)ixi;
}

template <typename T, int V, int LX>
__global__ void __launch_bounds__(256) e_pml_tile(KParams<T> p, PmlBoxSet<T> bs, const WorkItem *__restrict__ items,
                                                  int k_lo, int k_hi) {
    const WorkItem it = items[blockIdx.x];
    const PmlBox<T> &b = bs.b[it.box];
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = it.i0 + lx * V;
    const int j = it.j0 + warp * RW + ly;
    const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
    if (kb >= ke) return;
    const bool act = (j < b.hi[1]) && (i0 < b.hi[0]);
    const bool rowm = act && (j >= 1);
    const T C = p.courant;
    const int set = it.set;
    const long long step = *p.step;
    const int parity = (int)(step & 1);
    const long long so = (long long)set * p.set_stride;
    T *Ex = p.E[0] + so, *Ey = p.E[1] + so, *Ez = p.E[2] + so;
    const T *Hx = p.H[0] + so, *Hy = p.H[1] + so, *Hz = p.H[2] + so;
    long long x = (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;
    long long xb = (long long)set * b.bset + (long long)(kb - b.lo[2]) * b.bplane + (long long)(j - b.lo[1]) * b.bpitch + (i0 - b.lo[0]);
    T sxi[V], ixi[V], sxh[V];
    T syi = T(0), iyi = T(1), syh = T(0);
    if (act) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int h = 2 * (i0 + v);
            sxi[v] = p.sig[0][h]; ixi[v] = p.siginv[0][h]; sxh[v] = p.sig[0][h + 1];
        }
        syi = p.sig[1][2 * j]; iyi = p.siginv[1][2 * j]; syh = p.sig[1][2 * j + 1];
    }
    Vec<T, V> hxm, hym, hx0, hy0, hz0, hzj, hxj, ex, ey, ez, dx, dy, dz;
    if (act) { hxm.load(Hx + x - p.plane); hym.load(Hy + x - p.plane); } else { hxm.zero(); hym.zero(); }
    for (int k = kb; k < ke; ++k, x += p.plane, xb += b.bplane) {
        if (act) {
            hx0.load(Hx + x); hy0.load(Hy + x); hz0.load(Hz + x);
            ex.load(Ex + x); ey.load(Ey + x); ez.load(Ez + x);
            dx.load(b.D[0] + xb); dy.load(b.D[1] + xb); dz.load(b.D[2] + xb);
        } else { hx0.zero(); hy0.zero(); hz0.zero(); }
        if (rowm) { hzj.load(Hz + x - p.pitch); hxj.load(Hx + x - p.pitch); } else { hzj.zero(); hxj.zero(); }
        T hz_p = __shfl_up_sync(0xffffffffu, hz0.v[V - 1], 1, LX);
        T hy_p = __shfl_up_sync(0xffffffffu, hy0.v[V - 1], 1, LX);
        if (act && (lx == 0 || i0 == b.lo[0])) {
            hz_p = (i0 > 0) ? Hz[x - 1] : T(0);
            hy_p = (i0 > 0) ? Hy[x - 1] : T(0);
        }
        if (act) {
            const T szi = p.sig[2][2 * k], izi = p.siginv[2][2 * k], szh = p.sig[2][2 * k + 1];
            const bool kin = (k >= 1 && k <= p.n[2] - 1), jin = (j >= 1 && j <= p.n[1] - 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int i = i0 + v;
                const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                const bool iin = (i >= 1 && i <= p.n[0] - 1);
                const long long xg = so + x + v, xl = x + v;
                if (i <= p.n[0] - 1 && jin && kin) {     // Ex: k-dir y, u-dir z, w-dir x
                    const T curl = C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym.v[v]);
                    pml_e_elem(p, 0, ex.v[v], dx.v[v], curl, syi, iyi, szi, izi, sxh[v], b.UD[0], xb + v, xg, xl, i, j, k, set, step, parity);
                }
                if (j <= p.n[1] - 1 && iin && kin) {     // Ey: k-dir z, u-dir x, w-dir y
                    const T curl = C * (((hxm.v[v] - hx0.v[v]) + hz0.v[v]) - hzi);
                    pml_e_elem(p, 1, ey.v[v], dy.v[v], curl, szi, izi, sxi[v], ixi[v], syh, b.UD[1], xb + v, xg, xl, i, j, k, set, step, parity);
                }
                if (k <= p.n[2] - 1 && iin && jin) {     // Ez: k-dir x, u-dir y, w-dir z
                    const T curl = C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]);
                    pml_e_elem(p, 2, ez.v[v], dz.v[v], curl, sxi[v], ixi[v], syi, iyi, szh, b.UD[2], xb + v, xg, xl, i, j, k, set, step, parity);
                }
            }
            ex.store(Ex + x); ey.store(Ey + x); ez.store(Ez + x);
            dx.store(b.D[0] + xb); dy.store(b.D[1] + xb); dz.store(b.D[2] + xb);
        }
        hxm = hx0; hym = hy0;
    }
}

// ------------------------------------------------------------------------------------------
// monitors (fields.get_field -> linear interpolation of <= 8 Yee points), step counter
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void sample_monitors(KParams<T> p, MonDev m, long long base_step, int base_cursor, int save_span) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m.n_mon * p.n_sets) return;
    const int set = t / m.n_mon, mon = t % m.n_mon;
    const long long step = *p.step;
    const int cursor = base_cursor + (int)((step - base_step) / save_span);
    const T *F = (m.comp < 3 ? p.E[m.comp] : p.H[m.comp - 3]) + (long long)set * p.set_stride;
    double res = 0.0;
    bool owned = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const long long ix = m.idx[mon * 8 + q];
        const double w = m.w[mon * 8 + q];
        if (ix >= 0) { owned = true; if (w != 0.0) res += w * (double)F[ix]; }
    }
    if (!owned) res = 0.0;
    m.series[((long long)cursor * m.n_mon + mon) * p.n_sets + set] = res;
    if (set == 0 && (res > 1000.0 || res != res)) m.flags[0] = 1;
}

__global__ void tick_kernel(long long *step) { *step += 1; }
