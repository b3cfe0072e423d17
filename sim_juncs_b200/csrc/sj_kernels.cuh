// sj_kernels.cuh -- sm_100a kernels of the FDTD hot path.
//
// Two passes per time step, every field array touched once per pass:
//   H-pass : B -= C curl E (M1) fused with H = mu^-1 B and the UPML auxiliaries (M2)
//   E-pass : D += C curl H (M3) + source injection (M4) + Drude-Lorentz ADE (M5) +
//            E = chi1inv (D - sum P) (M6), UPML auxiliaries fused
// All kernels share one structure: a thread block owns an xy tile and marches along z over a chunk
// of planes; every row access is a 128-bit load/store; the k-1 / k+1 plane is carried in registers;
// the i-1 / i+1 neighbour comes from the adjacent lane by warp shuffle (one scalar load per tile
// edge).  LX = lanes of a warp along x (32, 16 or 8; the warp covers 32/LX rows) so that narrow
// regions still fill their warps.
//   *_interior : cells where every PML sigma is zero.  Only E and H exist there:
//                E^{n+1} = E^n + chi1inv (dD - dP - dS), algebraically meep's E = chi1inv (D - P - S).
//                A per-(tile,chunk) flag selects a uniform-material fast path (no material bytes
//                read, no ADE code) -- block-uniform, no divergence.
//   *_pml_tile : the PML shell (six boxes), driven by a work list.  Box-local auxiliary arrays hold
//                D / B (and U where two sigmas overlap); arithmetic follows meep's step_curl /
//                step_update_EDHB with the sig / siginv tables.
#pragma once
#include "sj_internal.h"

#ifndef SJ_HFACE_BLOCKS
#define SJ_HFACE_BLOCKS 3
#endif
#ifndef SJ_EINT0_BLOCKS
#define SJ_EINT0_BLOCKS 4
#endif
#ifndef SJ_EFACE_BLOCKS
#define SJ_EFACE_BLOCKS 3
#endif

template <typename T, int V> struct VecOf;
template <> struct VecOf<double, 2> { typedef double2 type; };
template <> struct VecOf<float, 4> { typedef float4 type; };
template <> struct VecOf<double, 1> { typedef double type; };
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
    // streaming store (st.global.cs): the value is not read again before a whole pass has gone through the L2
    __device__ __forceinline__ void store_cs(T *p) const {
        typedef typename VecOf<T, V>::type VT;
        VT t;
        T *q = reinterpret_cast<T *>(&t);
#pragma unroll
        for (int i = 0; i < V; ++i) q[i] = v[i];
        __stcs(reinterpret_cast<VT *>(p), t);
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = T(0);
    }
};

template <int V>
__device__ __forceinline__ void load_bytes(const uint8_t *p, unsigned char (&m)[V]) {
    if (V == 1) m[0] = *p;
    else if (V == 2) { const uchar2 t = *reinterpret_cast<const uchar2 *>(p); m[0] = t.x; m[V > 1 ? 1 : 0] = t.y; }
    else { const uchar4 t = *reinterpret_cast<const uchar4 *>(p); m[0] = t.x; m[1] = t.y; m[V > 2 ? 2 : 0] = t.z; m[V > 3 ? 3 : 0] = t.w; }
}

// ---- sources ---------------------------------------------------------------------------------
// bit s of the result: source s touches plane k (block-uniform test, a few scalar instructions)
template <typename T>
__device__ __forceinline__ unsigned src_plane_mask(const KParams<T> &p, int k) {
    unsigned m = 0;
    for (int s = 0; s < p.n_src; ++s)
        if (k >= p.src_klo[s] && k <= p.src_khi[s]) m |= 1u << s;
    return m;
}
// does any source touch the planes [kb, ke)?  Evaluated once per block: away from the source planes
// the per-plane test and the injection code are skipped by one uniform branch.
template <typename T>
__device__ __forceinline__ bool src_in_chunk(const KParams<T> &p, int kb, int ke) {
    bool any = false;
    for (int s = 0; s < p.n_src; ++s) any |= (ke > p.src_klo[s] && kb <= p.src_khi[s]);
    return any;
}
// S_n, S_{n+1} (integrated dipole) and dt*J_n (current) weights at a Yee point of component c.
// Only the SRC = true instantiation of a kernel body contains this code; blocks whose planes hold no
// source take the SRC = false body (one uniform branch per block), which has no source arithmetic.
template <typename T>
__device__ __forceinline__ void source_parts(const KParams<T> &p, unsigned smask, int c, int i, int j, int k, int set,
                                             long long step, T &S0, T &S1, T &J) {
    S0 = S1 = J = T(0);
    for (int s = 0; s < p.n_src; ++s) {
        if (!((smask >> s) & 1u)) continue;
        const SrcDev<T> &g = p.srcd[s];
        if (g.comp != c || j < g.lo[1] || j > g.hi[1] || i < g.lo[0] || i > g.hi[0]) continue;
        const T wgt = g.w[0][i - g.lo[0]] * g.w[1][j - g.lo[1]] * g.w[2][k - g.lo[2]];
        const T *d0 = p.drive + ((step * p.n_src + s) * p.n_sets + set) * 2;
        const T *d1 = d0 + (long long)p.n_src * p.n_sets * 2;
        S0 += wgt * d0[0]; S1 += wgt * d1[0]; J += wgt * d0[1];
    }
}

// ---- ADE (M5) ----------------------------------------------------------------------------------
// Polarisation arrays live in one allocation [parity][slot][comp][set][plane][row][x]; the material
// table is sorted by pole count, so the pole count of id m is a sum of compares (no table load) and
// the P loads of a plane can be issued at the top of the iteration together with the field loads.
template <typename T>
__device__ __forceinline__ int pole_count(const KParams<T> &p, int m) {
    return (m >= p.np_thr[0]) + (m >= p.np_thr[1]) + (m >= p.np_thr[2]) + (m >= p.np_thr[3]);
}
// offset that turns a field index xg (set, local plane, row, x) into the index of the same point in a polarisation array
template <typename T>
__device__ __forceinline__ long long pol_shift(const KParams<T> &p, int set) {
    return (long long)set * (p.set_stride - p.p_set_stride) + (long long)p.p_k0 * p.plane;
}
// base of the parity-`par` half; (slot s, component c) sits at + (3 s + c) * p_comp_stride
template <typename T>
__device__ __forceinline__ T *pol_base(const KParams<T> &p, int par) {
    return p.Pall + (long long)par * p.n_slots * 3 * p.p_comp_stride;
}

template <typename T, int V, int NS>
struct PolState {
    Vec<T, V> cur[3][NS > 0 ? NS : 1], prv[3][NS > 0 ? NS : 1];
    bool need[3][NS > 0 ? NS : 1];
    // issue every polarisation load of this vector (all components, all slots that any element uses)
    __device__ __forceinline__ void load(const KParams<T> &p, int parity, long long xg, const unsigned char (&mx)[V],
                                         const unsigned char (&my)[V], const unsigned char (&mz)[V]) {
        const T *bc = pol_base(p, parity) + xg, *bp = pol_base(p, parity ^ 1) + xg;
        const long long cs = p.p_comp_stride;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int npmax = 0;
#pragma unroll
            for (int v = 0; v < V; ++v) npmax = max(npmax, pole_count(p, c == 0 ? mx[v] : c == 1 ? my[v] : mz[v]));
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                need[c][s] = (npmax > s);
                if (need[c][s]) { cur[c][s].load(bc + (3 * s + c) * cs); prv[c][s].load(bp + (3 * s + c) * cs); }
                else { cur[c][s].zero(); prv[c][s].zero(); }
            }
        }
    }
    // advance the poles of element v of component c; returns sum(P_new - P_cur), also sum P_new
    __device__ __forceinline__ T advance(const KParams<T> &p, int c, int v, int m, T drive, T &sum_new) {
        T dP = T(0), sn = T(0);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const T *cf = p.mt_coef + ((long long)m * SJ_MAX_POLES + s) * 3;   // zero rows beyond the pole count
            const T pc = cur[c][s].v[v];
            const T pn = cf[0] * pc + cf[1] * prv[c][s].v[v] + cf[2] * drive;
            prv[c][s].v[v] = pn;      // written to the "previous" array, which becomes current after the flip
            dP += pn - pc; sn += pn;
        }
        sum_new = sn;
        return dP;
    }
    // same arithmetic with the coefficients of a block-uniform material held in registers (cfu[s] = {a, b, c})
    __device__ __forceinline__ T advance_u(int c, int v, T drive, T &sum_new, const T (&cfu)[NS > 0 ? NS : 1][3]) {
        T dP = T(0), sn = T(0);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const T pc = cur[c][s].v[v];
            const T pn = cfu[s][0] * pc + cfu[s][1] * prv[c][s].v[v] + cfu[s][2] * drive;
            prv[c][s].v[v] = pn;
            dP += pn - pc; sn += pn;
        }
        sum_new = sn;
        return dP;
    }
    // every slot of every component (uniform material with exactly NS poles)
    __device__ __forceinline__ void load_all(const KParams<T> &p, int parity, long long xg) {
        const T *bc = pol_base(p, parity) + xg, *bp = pol_base(p, parity ^ 1) + xg;
        const long long cs = p.p_comp_stride;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int s = 0; s < NS; ++s) { need[c][s] = true; cur[c][s].load(bc + (3 * s + c) * cs); prv[c][s].load(bp + (3 * s + c) * cs); }
    }
    __device__ __forceinline__ T sum_cur(int c, int v) const {
        T so = T(0);
#pragma unroll
        for (int s = 0; s < NS; ++s) so += cur[c][s].v[v];
        return so;
    }
    __device__ __forceinline__ void store(const KParams<T> &p, int parity, long long xg) const {
        T *bp = pol_base(p, parity ^ 1) + xg;
        const long long cs = p.p_comp_stride;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (need[c][s]) prv[c][s].store(bp + (3 * s + c) * cs);
    }
};

// ------------------------------------------------------------------------------------------
// Interior H-pass
// ------------------------------------------------------------------------------------------
template <typename T, int V, int LX>
__global__ void __launch_bounds__(256, 4) h_interior(KParams<T> p, IntGeom g, int k_begin, int k_end) {
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = g.i_lo + (blockIdx.x * LX + lx) * V;
    const int j = g.j_lo + (blockIdx.y * 8 + warp) * RW + ly;
    const int set = blockIdx.z / g.nzc;
    const int kc = g.c0 + blockIdx.z % g.nzc;
    const int kb = max(g.k_lo + kc * g.zchunk, k_begin);
    const int ke = min(min(g.k_lo + (kc + 1) * g.zchunk, g.k_hi), k_end);
    if (kb >= ke) return;
    const bool ld = (j < g.j_hi) && (i0 + V <= p.pitch);
    const bool st = (j < g.j_hi) && (i0 < g.i_hi);
    const T C = p.courant;
    const long long x0 = (long long)set * p.set_stride + (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    const T *pE = p.F + x0;
    T *pH = p.F + 3 * fcs + x0;
    const long long plane = p.plane;
    const int pitch = p.pitch;
    const bool edge = st && (lx == LX - 1) && (i0 + V < pitch);

    Vec<T, V> ex0, ey0, ex1, ey1, ez0, ezj, exj, hx, hy, hz;
    if (ld) { ex0.load(pE); ey0.load((pE + fcs)); } else { ex0.zero(); ey0.zero(); }
    for (int k = kb; k < ke; ++k) {
        if (ld) {
            ex1.load(pE + plane); ey1.load((pE + fcs) + plane);
            ez0.load((pE + fcs2)); ezj.load((pE + fcs2) + pitch); exj.load(pE + pitch);
        } else { ex1.zero(); ey1.zero(); ez0.zero(); ezj.zero(); exj.zero(); }
        if (st) { hx.load(pH); hy.load((pH + fcs)); hz.load((pH + fcs2)); }
        T ez_n = __shfl_down_sync(0xffffffffu, ez0.v[0], 1, LX);
        T ey_n = __shfl_down_sync(0xffffffffu, ey0.v[0], 1, LX);
        if (edge) { ez_n = (pE + fcs2)[V]; ey_n = (pE + fcs)[V]; }
        if (st) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T ezi = (v < V - 1) ? ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                hx.v[v] -= C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1.v[v]);
                hy.v[v] -= C * (((ex1.v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
                hz.v[v] -= C * (((eyi - ey0.v[v]) + ex0.v[v]) - exj.v[v]);
            }
            hx.store(pH); hy.store((pH + fcs)); hz.store((pH + fcs2));
        }
        ex0 = ex1; ey0 = ey1;
        pE += plane; pH += plane;
    }
}

// Software-pipelined variant: the loads of planes k+1 and k+2 are in flight while plane k is computed
// (three register sets, rotated by unrolling the plane loop by three).  Measured on this GPU
// (scripts/microbench/pipe.cu): one plane in flight per thread tops out near 5.1-6.3 TB/s whatever the
// occupancy, two or more reach 6.8-6.9 TB/s even at one or two blocks per SM.
template <typename T, int V>
struct HPlane { Vec<T, V> ex1, ey1, ez0, ezj, exj, hx, hy, hz; T ez_e, ey_e; };

template <typename T, int V, int LX>
__global__ void __launch_bounds__(256, 2) h_interior_pipe(KParams<T> p, IntGeom g, int k_begin, int k_end) {
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = g.i_lo + (blockIdx.x * LX + lx) * V;
    const int j = g.j_lo + (blockIdx.y * 8 + warp) * RW + ly;
    const int set = blockIdx.z / g.nzc;
    const int kc = g.c0 + blockIdx.z % g.nzc;
    const int kb = max(g.k_lo + kc * g.zchunk, k_begin);
    const int ke = min(min(g.k_lo + (kc + 1) * g.zchunk, g.k_hi), k_end);
    if (kb >= ke) return;
    const bool ld = (j < g.j_hi) && (i0 + V <= p.pitch);
    const bool st = (j < g.j_hi) && (i0 < g.i_hi);
    const T C = p.courant;
    const long long x0 = (long long)set * p.set_stride + (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    const T *__restrict__ pE = p.F + x0;
    T *pH = p.F + 3 * fcs + x0;
    const long long plane = p.plane;
    const int pitch = p.pitch;
    const bool edge = st && (lx == LX - 1) && (i0 + V < pitch);

    auto load_plane = [&](HPlane<T, V> &q, long long off) {
        if (ld) {
            q.ex1.load(pE + off + plane); q.ey1.load(pE + fcs + off + plane);
            q.ez0.load(pE + fcs2 + off); q.ezj.load(pE + fcs2 + off + pitch); q.exj.load(pE + off + pitch);
        } else { q.ex1.zero(); q.ey1.zero(); q.ez0.zero(); q.ezj.zero(); q.exj.zero(); }
        if (st) { q.hx.load(pH + off); q.hy.load(pH + fcs + off); q.hz.load(pH + fcs2 + off); }
        q.ez_e = T(0); q.ey_e = T(0);
        if (edge) { q.ez_e = (pE + fcs2 + off)[V]; q.ey_e = (pE + fcs + off)[V]; }
    };
    Vec<T, V> ex0, ey0;
    if (ld) { ex0.load(pE); ey0.load(pE + fcs); } else { ex0.zero(); ey0.zero(); }
    auto finish_plane = [&](HPlane<T, V> &q, long long off) {
        T ez_n = __shfl_down_sync(0xffffffffu, q.ez0.v[0], 1, LX);
        T ey_n = __shfl_down_sync(0xffffffffu, ey0.v[0], 1, LX);
        if (edge) { ez_n = q.ez_e; ey_n = q.ey_e; }
        if (st) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T ezi = (v < V - 1) ? q.ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                q.hx.v[v] -= C * (((q.ezj.v[v] - q.ez0.v[v]) + ey0.v[v]) - q.ey1.v[v]);
                q.hy.v[v] -= C * (((q.ex1.v[v] - ex0.v[v]) + q.ez0.v[v]) - ezi);
                q.hz.v[v] -= C * (((eyi - ey0.v[v]) + ex0.v[v]) - q.exj.v[v]);
            }
            q.hx.store(pH + off); q.hy.store(pH + fcs + off); q.hz.store(pH + fcs2 + off);
        }
        ex0 = q.ex1; ey0 = q.ey1;
    };
    HPlane<T, V> q0, q1, q2;
    load_plane(q0, 0);
    if (kb + 1 < ke) load_plane(q1, plane);
    long long off = 0;
    for (int k = kb; k < ke; k += 3) {
        if (k + 2 < ke) load_plane(q2, off + 2 * plane);
        finish_plane(q0, off);
        if (k + 1 >= ke) break;
        if (k + 3 < ke) load_plane(q0, off + 3 * plane);
        finish_plane(q1, off + plane);
        if (k + 2 >= ke) break;
        if (k + 4 < ke) load_plane(q1, off + 4 * plane);
        finish_plane(q2, off + 2 * plane);
        off += 3 * plane;
    }
}

// ------------------------------------------------------------------------------------------
// Interior E-pass
// ------------------------------------------------------------------------------------------
// NS = 0: uniform non-dispersive material (chi_u), no material bytes read; NS > 0: general path for
// materials with up to NS poles (every polarisation load is issued with the field loads).
template <typename T, int V, int LX, int NS, bool SRC>
__device__ __forceinline__ void e_interior_body(const KParams<T> &p, const IntGeom &g, int i0, int j, int lx, int set, int kb,
                                                int ke, T chi_u) {
    constexpr bool GEN = NS > 0;
    const bool ld = (j < g.j_hi) && (i0 + V <= p.pitch);
    const bool st = (j < g.j_hi) && (i0 < g.i_hi);
    const T C = p.courant;
    const long long step = *p.step;
    const int parity = (int)(step & 1);
    const long long xl0 = (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;
    long long xg = (long long)set * p.set_stride + xl0;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    T *pE = p.F + xg;
    const T *pH = p.F + 3 * fcs + xg;
    const long long mcs = p.set_stride, mcs2 = 2 * p.set_stride;
    const uint8_t *pm = p.mat[0] + xl0;
    const long long plane = p.plane;
    const int pitch = p.pitch;
    const bool edge = st && (lx == 0) && (i0 > 0);

    Vec<T, V> hxm, hym, hx0, hy0, hz0, hzj, hxj, ex, ey, ez;
    unsigned char mx[V], my[V], mz[V];
#pragma unroll
    for (int v = 0; v < V; ++v) mx[v] = my[v] = mz[v] = 0;
    PolState<T, V, NS> pol;
    if (ld) { hxm.load(pH - plane); hym.load((pH + fcs) - plane); } else { hxm.zero(); hym.zero(); }
    if (GEN && st) { load_bytes<V>(pm, mx); load_bytes<V>((pm + mcs), my); load_bytes<V>((pm + mcs2), mz); }
    for (int k = kb; k < ke; ++k) {
        if (ld) {
            hx0.load(pH); hy0.load((pH + fcs)); hz0.load((pH + fcs2));
            hzj.load((pH + fcs2) - pitch); hxj.load(pH - pitch);
        } else { hx0.zero(); hy0.zero(); hz0.zero(); hzj.zero(); hxj.zero(); }
        // the tile-edge neighbours are fetched with the plane's other loads (a load placed after the shuffles
        // would cost the whole warp a second memory round trip per plane)
        T hz_e = T(0), hy_e = T(0);
        if (edge) { hz_e = (pH + fcs2)[-1]; hy_e = (pH + fcs)[-1]; }
        unsigned char nx_[V], ny_[V], nz_[V];
        if (st) {
            ex.load(pE); ey.load((pE + fcs)); ez.load((pE + fcs2));
            if (GEN) {
                pol.load(p, parity, xg - pol_shift(p, set), mx, my, mz);
                // material bytes of the next plane (a halo plane always exists)
                load_bytes<V>(pm + plane, nx_); load_bytes<V>((pm + mcs) + plane, ny_); load_bytes<V>((pm + mcs2) + plane, nz_);
            }
        }
        T hz_p = __shfl_up_sync(0xffffffffu, hz0.v[V - 1], 1, LX);
        T hy_p = __shfl_up_sync(0xffffffffu, hy0.v[V - 1], 1, LX);
        if (edge) { hz_p = hz_e; hy_p = hy_e; }
        const unsigned smask = SRC ? src_plane_mask(p, k) : 0u;
        if (st) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                T dDx = -(C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym.v[v]));
                T dDy = -(C * (((hxm.v[v] - hx0.v[v]) + hz0.v[v]) - hzi));
                T dDz = -(C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]));
                if (SRC && smask) {
                    T S0, S1, J;
                    source_parts(p, smask, 0, i0 + v, j, k, set, step, S0, S1, J); dDx -= (S1 - S0) + J;
                    source_parts(p, smask, 1, i0 + v, j, k, set, step, S0, S1, J); dDy -= (S1 - S0) + J;
                    source_parts(p, smask, 2, i0 + v, j, k, set, step, S0, S1, J); dDz -= (S1 - S0) + J;
                }
                if (GEN) {
                    T sn;
                    ex.v[v] += p.mt_chi[mx[v]] * (dDx - pol.advance(p, 0, v, mx[v], ex.v[v], sn));
                    ey.v[v] += p.mt_chi[my[v]] * (dDy - pol.advance(p, 1, v, my[v], ey.v[v], sn));
                    ez.v[v] += p.mt_chi[mz[v]] * (dDz - pol.advance(p, 2, v, mz[v], ez.v[v], sn));
                } else {
                    ex.v[v] += chi_u * dDx; ey.v[v] += chi_u * dDy; ez.v[v] += chi_u * dDz;
                }
            }
            ex.store(pE); ey.store((pE + fcs)); ez.store((pE + fcs2));
            if (GEN) {
                pol.store(p, parity, xg - pol_shift(p, set));
#pragma unroll
                for (int v = 0; v < V; ++v) { mx[v] = nx_[v]; my[v] = ny_[v]; mz[v] = nz_[v]; }
            }
        }
        hxm = hx0; hym = hy0;
        pE += plane; pH += plane;
        pm += plane; xg += plane;
    }
}

template <typename T, int V, int LX, int NS>
__global__ void __launch_bounds__(256, NS == 0 ? SJ_EINT0_BLOCKS : 1) e_interior(KParams<T> p, IntGeom g, const WorkItem *__restrict__ items,
                                                                   int k_begin, int k_end) {
    const WorkItem it = items[blockIdx.x];
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = it.i0 + lx * V;
    const int j = it.j0 + warp * RW + ly;
    const int kb = max(it.kb, k_begin), ke = min(it.ke, k_end);
    if (kb >= ke) return;
    const T chi_u = NS > 0 ? T(0) : p.mt_chi[it.mat];
    if (src_in_chunk(p, kb, ke)) e_interior_body<T, V, LX, NS, true>(p, g, i0, j, lx, it.set, kb, ke, chi_u);
    else e_interior_body<T, V, LX, NS, false>(p, g, i0, j, lx, it.set, kb, ke, chi_u);
}

// ------------------------------------------------------------------------------------------
// Staged variant of the general interior E-pass: every vector of plane k+1 is fetched with
// cp.async (LDGSTS, 16 B per thread, L2 -> shared, no registers held) while plane k is computed.
// Each thread only ever reads the shared-memory slots it filled itself, so the pipeline needs
// cp.async.wait_group but no block barrier.  Slots: [stage][slot][thread] of 16 bytes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <typename T, int V, int NSLOT, int NT = 256>
struct Stager {
    uint4 *base;   // &smem[threadIdx.x]
    __device__ __forceinline__ uint4 *addr(int st, int slot) const { return base + (st * NSLOT + slot) * NT; }
    __device__ __forceinline__ void issue(int st, int slot, const T *g, bool pred) const {
        if (pred) cp_async16(addr(st, slot), g);
        else *addr(st, slot) = make_uint4(0u, 0u, 0u, 0u);
    }
    __device__ __forceinline__ void get(int st, int slot, Vec<T, V> &v) const {
        const uint4 t = *addr(st, slot);
        const T *q = reinterpret_cast<const T *>(&t);
#pragma unroll
        for (int i = 0; i < V; ++i) v.v[i] = q[i];
    }
};

// UNI: the tile holds one material with exactly NS poles (it.mat): no material bytes, coefficients in registers,
// every polarisation vector loaded unconditionally.  Same arithmetic as the general path.
template <typename T, int V, int LX, int NS, bool SRC, bool UNI>
__device__ __forceinline__ void e_interior_stg_body(const KParams<T> &p, const IntGeom &g, const WorkItem &it, int kb, int ke,
                                                    uint4 *smem) {
    constexpr int NSLOT = 8 + 6 * NS;
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = it.i0 + lx * V;
    const int j = it.j0 + warp * RW + ly;
    const int set = it.set;
    Stager<T, V, NSLOT> sg; sg.base = smem + threadIdx.x;
    const bool ld = (j < g.j_hi) && (i0 + V <= p.pitch);
    const bool st = (j < g.j_hi) && (i0 < g.i_hi);
    const T C = p.courant;
    const long long step = *p.step;
    const int parity = (int)(step & 1);
    const long long xl0 = (long long)(kb - p.kz0 + 1) * p.plane + (long long)j * p.pitch + i0;
    long long xg = (long long)set * p.set_stride + xl0;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    T *pE = p.F + xg;
    const T *pH = p.F + 3 * fcs + xg;
    const long long mcs = p.set_stride, mcs2 = 2 * p.set_stride;
    const uint8_t *pm = p.mat[0] + xl0;
    const long long plane = p.plane;
    const int pitch = p.pitch;
    const bool edge = st && (lx == 0) && (i0 > 0);
    const long long pcs = p.p_comp_stride, psh = pol_shift(p, set);
    const T *bcur = p.Pall + (long long)parity * p.n_slots * 3 * pcs;          // parity half read as "current"
    T *bprv = p.Pall + (long long)(parity ^ 1) * p.n_slots * 3 * pcs;          // read as "previous", written as new

    // which polarisation vectors the materials of a plane need: bit (3 s + c)
    auto need_of = [&](const unsigned char (&ax)[V], const unsigned char (&ay)[V], const unsigned char (&az)[V]) {
        unsigned nd = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int npmax = 0;
#pragma unroll
            for (int v = 0; v < V; ++v) npmax = max(npmax, pole_count(p, c == 0 ? ax[v] : c == 1 ? ay[v] : az[v]));
#pragma unroll
            for (int s = 0; s < NS; ++s) if (npmax > s) nd |= 1u << (3 * s + c);
        }
        return nd;
    };
    // issue every vector load of the plane at offset `off` (0 = plane k of the current pointers)
    auto issue_plane = [&](int stg, long long off, long long xgp, unsigned nd) {
        sg.issue(stg, 0, pH + off, ld); sg.issue(stg, 1, pH + fcs + off, ld); sg.issue(stg, 2, pH + fcs2 + off, ld);
        sg.issue(stg, 3, pH + fcs2 + off - pitch, ld); sg.issue(stg, 4, pH + off - pitch, ld);
        sg.issue(stg, 5, pE + off, st); sg.issue(stg, 6, pE + fcs + off, st); sg.issue(stg, 7, pE + fcs2 + off, st);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const bool nd_ = st && ((nd >> (3 * s + c)) & 1u);
                sg.issue(stg, 8 + (c * NS + s) * 2, bcur + (3 * s + c) * pcs + (xgp - psh), nd_);
                sg.issue(stg, 9 + (c * NS + s) * 2, bprv + (3 * s + c) * pcs + (xgp - psh), nd_);
            }
    };

    Vec<T, V> hxm, hym, hx0, hy0, hz0, hzj, hxj, ex, ey, ez;
    unsigned char mx[V], my[V], mz[V], nx_[V], ny_[V], nz_[V];
#pragma unroll
    for (int v = 0; v < V; ++v) mx[v] = my[v] = mz[v] = nx_[v] = ny_[v] = nz_[v] = 0;
    if (ld) { hxm.load(pH - plane); hym.load((pH + fcs) - plane); } else { hxm.zero(); hym.zero(); }
    if (!UNI && st) {
        load_bytes<V>(pm, mx); load_bytes<V>(pm + mcs, my); load_bytes<V>(pm + mcs2, mz);
        load_bytes<V>(pm + plane, nx_); load_bytes<V>(pm + mcs + plane, ny_); load_bytes<V>(pm + mcs2 + plane, nz_);
    }
    constexpr unsigned NEED_ALL = (1u << (3 * NS)) - 1u;
    T cfu[NS][3], chi_m = T(0);
    if (UNI) {
        chi_m = p.mt_chi[it.mat];
#pragma unroll
        for (int s_ = 0; s_ < NS; ++s_)
#pragma unroll
            for (int q_ = 0; q_ < 3; ++q_) cfu[s_][q_] = p.mt_coef[((long long)it.mat * SJ_MAX_POLES + s_) * 3 + q_];
    }
    unsigned need_c = UNI ? NEED_ALL : need_of(mx, my, mz);
    T hz_pc = T(0), hy_pc = T(0);
    if (edge) { hz_pc = (pH + fcs2)[-1]; hy_pc = (pH + fcs)[-1]; }
    issue_plane(0, 0, xg, need_c);
    cp_async_commit();
    for (int k = kb; k < ke; ++k) {
        const int cur = (k - kb) & 1;
        const unsigned need_n = UNI ? NEED_ALL : need_of(nx_, ny_, nz_);
        unsigned char fx[V], fy[V], fz[V];          // material bytes two planes ahead
#pragma unroll
        for (int v = 0; v < V; ++v) fx[v] = fy[v] = fz[v] = 0;
        T hz_pn = T(0), hy_pn = T(0);
        if (k + 1 < ke) {
            issue_plane(cur ^ 1, plane, xg + plane, need_n);
            if (!UNI && st) { load_bytes<V>(pm + 2 * plane, fx); load_bytes<V>(pm + mcs + 2 * plane, fy); load_bytes<V>(pm + mcs2 + 2 * plane, fz); }
            if (edge) { hz_pn = (pH + fcs2 + plane)[-1]; hy_pn = (pH + fcs + plane)[-1]; }
        }
        cp_async_commit();
        cp_async_wait<1>();
        sg.get(cur, 0, hx0); sg.get(cur, 1, hy0); sg.get(cur, 2, hz0); sg.get(cur, 3, hzj); sg.get(cur, 4, hxj);
        T hz_p = __shfl_up_sync(0xffffffffu, hz0.v[V - 1], 1, LX);
        T hy_p = __shfl_up_sync(0xffffffffu, hy0.v[V - 1], 1, LX);
        if (edge) { hz_p = hz_pc; hy_p = hy_pc; }
        const unsigned smask = SRC ? src_plane_mask(p, k) : 0u;
        if (st) {
            sg.get(cur, 5, ex); sg.get(cur, 6, ey); sg.get(cur, 7, ez);
            Vec<T, V> pc[3][NS], pp[3][NS];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int s = 0; s < NS; ++s) { sg.get(cur, 8 + (c * NS + s) * 2, pc[c][s]); sg.get(cur, 9 + (c * NS + s) * 2, pp[c][s]); }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                T dD[3];
                dD[0] = -(C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym.v[v]));
                dD[1] = -(C * (((hxm.v[v] - hx0.v[v]) + hz0.v[v]) - hzi));
                dD[2] = -(C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]));
                if (SRC && smask) {
                    T S0, S1, J;
#pragma unroll
                    for (int c = 0; c < 3; ++c) { source_parts(p, smask, c, i0 + v, j, k, set, step, S0, S1, J); dD[c] -= (S1 - S0) + J; }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int m = c == 0 ? mx[v] : c == 1 ? my[v] : mz[v];
                    T &e = c == 0 ? ex.v[v] : c == 1 ? ey.v[v] : ez.v[v];
                    T dP = T(0);
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const T *cf = p.mt_coef + ((long long)m * SJ_MAX_POLES + s) * 3;
                        const T c0 = UNI ? cfu[s][0] : cf[0], c1 = UNI ? cfu[s][1] : cf[1], c2 = UNI ? cfu[s][2] : cf[2];
                        const T pcur = pc[c][s].v[v];
                        const T pn = c0 * pcur + c1 * pp[c][s].v[v] + c2 * e;
                        pp[c][s].v[v] = pn;
                        dP += pn - pcur;
                    }
                    e += (UNI ? chi_m : p.mt_chi[m]) * (dD[c] - dP);
                }
            }
            ex.store(pE); ey.store((pE + fcs)); ez.store((pE + fcs2));
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if ((need_c >> (3 * s + c)) & 1u) pp[c][s].store(bprv + (3 * s + c) * pcs + (xg - psh));
        }
        hxm = hx0; hym = hy0;
        hz_pc = hz_pn; hy_pc = hy_pn;
        need_c = need_n;
#pragma unroll
        for (int v = 0; v < V; ++v) { mx[v] = nx_[v]; my[v] = ny_[v]; mz[v] = nz_[v]; nx_[v] = fx[v]; ny_[v] = fy[v]; nz_[v] = fz[v]; }
        pE += plane; pH += plane; pm += plane; xg += plane;
    }
}

template <typename T, int V, int LX, int NS, bool UNI>
__global__ void __launch_bounds__(256, 1) e_interior_stg(KParams<T> p, IntGeom g, const WorkItem *__restrict__ items,
                                                         int k_begin, int k_end) {
    extern __shared__ uint4 sj_smem[];
    const WorkItem it = items[blockIdx.x];
    const int kb = max(it.kb, k_begin), ke = min(it.ke, k_end);
    if (kb >= ke) return;
    if (src_in_chunk(p, kb, ke)) e_interior_stg_body<T, V, LX, NS, true, UNI>(p, g, it, kb, ke, sj_smem);
    else e_interior_stg_body<T, V, LX, NS, false, UNI>(p, g, it, kb, ke, sj_smem);
}

// ------------------------------------------------------------------------------------------
// PML shell, tiled.  PD = 0: general cell (any combination of sigmas; D/B for all components, U
// where two sigmas overlap).  PD = 1,2,3: "face" tile whose cells have a single non-zero sigma,
// along x,y,z: no U; the tangential components need no auxiliary array at all (H doubles as B;
// D is reconstructed as eps E + sum P + S), only the normal component keeps its D / B.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T pml_step_db(T fold, T curl, T sk, T ik, T su, T iu, T &U) {
    if (su != T(0) && sk != T(0)) {
        const T uo = U;
        const T un = ((T(1) - sk) * uo - curl) * ik;
        U = un;
        return ((T(1) - su) * fold + (un - uo)) * iu;
    }
    if (su != T(0)) return ((T(1) - su) * fold - curl) * iu;
    return ((T(1) - sk) * fold - curl) * ik;
}

// MODE 0: general; 1: tangential component of a face (sigma sf/isf, no D array); 2: normal
// component of a face (sigma sw on W, D stored, no sigma on D).
// UNI: block-uniform material (chi_u, eps_u, cfu hold its table entries), also when it has poles
template <typename T, int V, int NS, int MODE, bool SRC, bool UNI>
__device__ __forceinline__ void pml_e_elem(const KParams<T> &p, PolState<T, V, NS> &pol, int c, int v, T &e, T &d, T curl, T sk,
                                           T ik, T su, T iu, T sw, T &U, int m, T chi_u, T eps_u, T S0, T S1, T J,
                                           const T (&cfu)[NS > 0 ? NS : 1][3]) {
    const T chi = (NS > 0 && !UNI) ? p.mt_chi[m] : chi_u;
    const T pold = NS > 0 ? pol.sum_cur(c, v) : T(0);
    T pnew = T(0);
    if (MODE == 1) {
        const T eps = (NS > 0 && !UNI) ? p.mt_eps[m] : eps_u;
        const T dold = SRC ? (eps * e + pold) + S0 : (eps * e + pold);
        T dnew = ((T(1) - sk) * dold - curl) * ik;
        if (SRC) dnew -= J;
        if (NS > 0) { if (UNI) pol.advance_u(c, v, e, pnew, cfu); else pol.advance(p, c, v, m, e, pnew); }   // W == E here
        e = SRC ? chi * ((dnew - pnew) - S1) : chi * (dnew - pnew);
        return;
    }
    const T dold = d;
    T dnew = (MODE == 2) ? dold - curl : pml_step_db(dold, curl, sk, ik, su, iu, U);
    if (SRC) dnew -= J;
    d = dnew;
    // W^n = chi (D^n - sum P^n - S^n) drives the poles (meep update_pols runs on W)
    const T wold = SRC ? chi * ((dold - pold) - S0) : chi * (dold - pold);
    if (NS > 0) { if (UNI) pol.advance_u(c, v, wold, pnew, cfu); else pol.advance(p, c, v, m, wold, pnew); }
    const T wnew = SRC ? chi * ((dnew - pnew) - S1) : chi * (dnew - pnew);
    e = (MODE == 2 || sw != T(0)) ? e + (T(1) + sw) * wnew - (T(1) - sw) * wold : wnew;
}

template <typename T, int V, int LX, int PD>
__device__ __forceinline__ void h_pml_body(const KParams<T> &p, const PmlBox<T> &b, const WorkItem &it, int k_lo, int k_hi) {
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = it.i0 + lx * V;
    const int j = it.j0 + warp * RW + ly;
    const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
    if (kb >= ke) return;
    const bool act = (j < it.j_hi) && (i0 < it.i_hi);
    const bool rowp = act && (j + 1 <= p.n[1]);
    const T C = p.courant;
    const long long plane = p.plane, bplane = b.bplane;
    const int pitch = p.pitch;
    const long long x0 = (long long)it.set * p.set_stride + (long long)(kb - p.kz0 + 1) * plane + (long long)j * pitch + i0;
    const long long xb0 = (long long)it.set * b.bset + (long long)(kb - b.lo[2]) * bplane + (long long)(j - b.lo[1]) * b.bpitch + (i0 - b.lo[0]);
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    const T *pE = p.F + x0;
    T *pH = p.F + 3 * fcs + x0;
    const long long bcs = b.bcs, bcs2 = 2 * b.bcs;
    T *pB = b.B[0] + xb0;
    T *pU = b.UB[0] + xb0;
    const bool last = (lx == LX - 1 || i0 + V >= it.i_hi);
    const bool edge = act && last && (i0 + V < pitch);
    // per-thread PML coefficients along x (per element) and y
    T sxi[V], sxh[V], ixh[V];
    T syi = T(0), syh = T(0), iyh = T(1);
#pragma unroll
    for (int v = 0; v < V; ++v) { sxi[v] = T(0); sxh[v] = T(0); ixh[v] = T(1); }
    if (act) {
        if (PD == 0 || PD == 1) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int h = 2 * (i0 + v);
                sxi[v] = p.sig[0][h]; sxh[v] = p.sig[0][h + 1]; ixh[v] = p.siginv[0][h + 1];
            }
        }
        if (PD == 0 || PD == 2) { syi = p.sig[1][2 * j]; syh = p.sig[1][2 * j + 1]; iyh = p.siginv[1][2 * j + 1]; }
    }
    const bool jok = (j <= p.n[1] - 1);
    Vec<T, V> ex0, ey0, ex1, ey1, ez0, ezj, exj, hx, hy, hz, bx, by, bz, ux, uy, uz;
    bx.zero(); by.zero(); bz.zero(); ux.zero(); uy.zero(); uz.zero();
    if (act) { ex0.load(pE); ey0.load((pE + fcs)); } else { ex0.zero(); ey0.zero(); }
    for (int k = kb; k < ke; ++k) {
        if (act) {
            ex1.load(pE + plane); ey1.load((pE + fcs) + plane); ez0.load((pE + fcs2));
            hx.load(pH); hy.load((pH + fcs)); hz.load((pH + fcs2));
            if (PD == 0 || PD == 1) bx.load(pB);
            if (PD == 0 || PD == 2) by.load((pB + bcs));
            if (PD == 0 || PD == 3) bz.load((pB + bcs2));
            if (PD == 0) { ux.load(pU); uy.load((pU + bcs)); uz.load((pU + bcs2)); }
        } else { ex1.zero(); ey1.zero(); ez0.zero(); }
        if (rowp) { ezj.load((pE + fcs2) + pitch); exj.load(pE + pitch); } else { ezj.zero(); exj.zero(); }
        T ez_e = T(0), ey_e = T(0);       // tile-edge neighbours, issued with the plane's other loads
        if (edge) { ez_e = (pE + fcs2)[V]; ey_e = (pE + fcs)[V]; }
        T ez_n = __shfl_down_sync(0xffffffffu, ez0.v[0], 1, LX);
        T ey_n = __shfl_down_sync(0xffffffffu, ey0.v[0], 1, LX);
        if (last) { ez_n = ez_e; ey_n = ey_e; }
        if (act) {
            T szi = T(0), szh = T(0), izh = T(1);
            if (PD == 0 || PD == 3) { szi = p.sig[2][2 * k]; szh = p.sig[2][2 * k + 1]; izh = p.siginv[2][2 * k + 1]; }
            const bool kok = (k <= p.n[2] - 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int i = i0 + v;
                const T ezi = (v < V - 1) ? ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                const bool iok = (i <= p.n[0] - 1);
                // the face sigma at the half-pixel position (tangential components)
                const T sf = PD == 1 ? sxh[v] : PD == 2 ? syh : szh, isf = PD == 1 ? ixh[v] : PD == 2 ? iyh : izh;
                if (i >= 1 && iok && jok && kok) {       // Hx: k-dir y, u-dir z, w-dir x
                    const T curl = C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1.v[v]);
                    if (PD == 0) {
                        const T bo = bx.v[v];
                        const T bn = pml_step_db(bo, curl, syh, iyh, szh, izh, ux.v[v]);
                        bx.v[v] = bn;
                        hx.v[v] = (sxi[v] != T(0)) ? hx.v[v] + (T(1) + sxi[v]) * bn - (T(1) - sxi[v]) * bo : bn;
                    } else if (PD == 1) {
                        const T bo = bx.v[v], bn = bo - curl;
                        bx.v[v] = bn;
                        hx.v[v] += (T(1) + sxi[v]) * bn - (T(1) - sxi[v]) * bo;
                    } else hx.v[v] = ((T(1) - sf) * hx.v[v] - curl) * isf;
                }
                if (j >= 1 && jok && iok && kok) {       // Hy: k-dir z, u-dir x, w-dir y
                    const T curl = C * (((ex1.v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
                    if (PD == 0) {
                        const T bo = by.v[v];
                        const T bn = pml_step_db(bo, curl, szh, izh, sxh[v], ixh[v], uy.v[v]);
                        by.v[v] = bn;
                        hy.v[v] = (syi != T(0)) ? hy.v[v] + (T(1) + syi) * bn - (T(1) - syi) * bo : bn;
                    } else if (PD == 2) {
                        const T bo = by.v[v], bn = bo - curl;
                        by.v[v] = bn;
                        hy.v[v] += (T(1) + syi) * bn - (T(1) - syi) * bo;
                    } else hy.v[v] = ((T(1) - sf) * hy.v[v] - curl) * isf;
                }
                if (k >= 1 && kok && iok && jok) {       // Hz: k-dir x, u-dir y, w-dir z
                    const T curl = C * (((eyi - ey0.v[v]) + ex0.v[v]) - exj.v[v]);
                    if (PD == 0) {
                        const T bo = bz.v[v];
                        const T bn = pml_step_db(bo, curl, sxh[v], ixh[v], syh, iyh, uz.v[v]);
                        bz.v[v] = bn;
                        hz.v[v] = (szi != T(0)) ? hz.v[v] + (T(1) + szi) * bn - (T(1) - szi) * bo : bn;
                    } else if (PD == 3) {
                        const T bo = bz.v[v], bn = bo - curl;
                        bz.v[v] = bn;
                        hz.v[v] += (T(1) + szi) * bn - (T(1) - szi) * bo;
                    } else hz.v[v] = ((T(1) - sf) * hz.v[v] - curl) * isf;
                }
            }
            hx.store(pH); hy.store((pH + fcs)); hz.store((pH + fcs2));
            if (PD == 0 || PD == 1) bx.store(pB);
            if (PD == 0 || PD == 2) by.store((pB + bcs));
            if (PD == 0 || PD == 3) bz.store((pB + bcs2));
            if (PD == 0) { ux.store(pU); uy.store((pU + bcs)); uz.store((pU + bcs2)); }
        }
        ex0 = ex1; ey0 = ey1;
        pE += plane; pH += plane;
        pB += bplane; pU += bplane;
    }
}

// FACE = false: general tiles (PD 0); FACE = true: face tiles, normal direction taken from the item
template <typename T, int V, int LX, bool FACE>
__global__ void __launch_bounds__(256, (V * sizeof(T) == 8) ? (FACE ? 4 : 3) : (FACE ? SJ_HFACE_BLOCKS : 2)) h_pml_tile(KParams<T> p, PmlBoxSet<T> bs, const WorkItem *__restrict__ items,
                                                     int k_lo, int k_hi) {
    const WorkItem it = items[blockIdx.x];
    const PmlBox<T> &b = bs.b[it.box];
    if (!FACE) h_pml_body<T, V, LX, 0>(p, b, it, k_lo, k_hi);
    else if (it.kind == 1) h_pml_body<T, V, LX, 1>(p, b, it, k_lo, k_hi);
    else if (it.kind == 2) h_pml_body<T, V, LX, 2>(p, b, it, k_lo, k_hi);
    else h_pml_body<T, V, LX, 3>(p, b, it, k_lo, k_hi);
}

template <typename T, int V, int LX, int NS, int PD, bool SRC, bool UNI>
__device__ __forceinline__ void e_pml_body(const KParams<T> &p, const PmlBox<T> &b, const WorkItem &it, int k_lo, int k_hi,
                                           T chi_u, T eps_u) {
    constexpr bool GEN = NS > 0 && !UNI;          // per-cell material bytes and table look-ups
    constexpr bool POL = NS > 0;
    constexpr int RW = 32 / LX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int i0 = it.i0 + lx * V;
    const int j = it.j0 + warp * RW + ly;
    const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
    if (kb >= ke) return;
    const bool act = (j < it.j_hi) && (i0 < it.i_hi);
    const bool rowm = act && (j >= 1);
    const T C = p.courant;
    const int set = it.set;
    const long long step = *p.step;
    const int parity = (int)(step & 1);
    const long long plane = p.plane, bplane = b.bplane;
    const int pitch = p.pitch;
    const long long xl0 = (long long)(kb - p.kz0 + 1) * plane + (long long)j * pitch + i0;
    long long xg = (long long)set * p.set_stride + xl0;
    const long long xb0 = (long long)set * b.bset + (long long)(kb - b.lo[2]) * bplane + (long long)(j - b.lo[1]) * b.bpitch + (i0 - b.lo[0]);
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    T *pE = p.F + xg;
    const T *pH = p.F + 3 * fcs + xg;
    const long long bcs = b.bcs, bcs2 = 2 * b.bcs;
    T *pD = b.D[0] + xb0;
    T *pU = b.UD[0] + xb0;
    const long long mcs = p.set_stride, mcs2 = 2 * p.set_stride;
    const uint8_t *pm = p.mat[0] + xl0;
    const bool first = (lx == 0 || i0 == it.i0);
    const bool edge = act && first && (i0 > 0);
    T sxi[V], ixi[V], sxh[V];
    T syi = T(0), iyi = T(1), syh = T(0);
#pragma unroll
    for (int v = 0; v < V; ++v) { sxi[v] = T(0); ixi[v] = T(1); sxh[v] = T(0); }
    if (act) {
        if (PD == 0 || PD == 1) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int h = 2 * (i0 + v);
                sxi[v] = p.sig[0][h]; ixi[v] = p.siginv[0][h]; sxh[v] = p.sig[0][h + 1];
            }
        }
        if (PD == 0 || PD == 2) { syi = p.sig[1][2 * j]; iyi = p.siginv[1][2 * j]; syh = p.sig[1][2 * j + 1]; }
    }
    const bool jin = (j >= 1 && j <= p.n[1] - 1);
    Vec<T, V> hxm, hym, hx0, hy0, hz0, hzj, hxj, ex, ey, ez, dx, dy, dz, ux, uy, uz;
    dx.zero(); dy.zero(); dz.zero(); ux.zero(); uy.zero(); uz.zero();
    unsigned char mx[V], my[V], mz[V];
#pragma unroll
    for (int v = 0; v < V; ++v) mx[v] = my[v] = mz[v] = 0;
    PolState<T, V, NS> pol;
    T cfu[NS > 0 ? NS : 1][3];
#pragma unroll
    for (int s_ = 0; s_ < (NS > 0 ? NS : 1); ++s_)
#pragma unroll
        for (int q_ = 0; q_ < 3; ++q_) cfu[s_][q_] = (UNI && POL) ? p.mt_coef[((long long)it.mat * SJ_MAX_POLES + s_) * 3 + q_] : T(0);
    if (act) { hxm.load(pH - plane); hym.load((pH + fcs) - plane); } else { hxm.zero(); hym.zero(); }
    if (GEN && act) { load_bytes<V>(pm, mx); load_bytes<V>((pm + mcs), my); load_bytes<V>((pm + mcs2), mz); }
    for (int k = kb; k < ke; ++k) {
        unsigned char nx_[V], ny_[V], nz_[V];
        if (act) {
            hx0.load(pH); hy0.load((pH + fcs)); hz0.load((pH + fcs2));
            ex.load(pE); ey.load((pE + fcs)); ez.load((pE + fcs2));
            if (PD == 0 || PD == 1) dx.load(pD);
            if (PD == 0 || PD == 2) dy.load((pD + bcs));
            if (PD == 0 || PD == 3) dz.load((pD + bcs2));
            if (PD == 0) { ux.load(pU); uy.load((pU + bcs)); uz.load((pU + bcs2)); }
            if (GEN) {
                pol.load(p, parity, xg - pol_shift(p, set), mx, my, mz);
                load_bytes<V>(pm + plane, nx_); load_bytes<V>((pm + mcs) + plane, ny_); load_bytes<V>((pm + mcs2) + plane, nz_);
            } else if (POL) pol.load_all(p, parity, xg - pol_shift(p, set));
        } else { hx0.zero(); hy0.zero(); hz0.zero(); }
        if (rowm) { hzj.load((pH + fcs2) - pitch); hxj.load(pH - pitch); } else { hzj.zero(); hxj.zero(); }
        T hz_e = T(0), hy_e = T(0);       // tile-edge neighbours, issued with the plane's other loads
        if (edge) { hz_e = (pH + fcs2)[-1]; hy_e = (pH + fcs)[-1]; }
        T hz_p = __shfl_up_sync(0xffffffffu, hz0.v[V - 1], 1, LX);
        T hy_p = __shfl_up_sync(0xffffffffu, hy0.v[V - 1], 1, LX);
        if (first) { hz_p = hz_e; hy_p = hy_e; }
        const unsigned smask = SRC ? src_plane_mask(p, k) : 0u;
        if (act) {
            T szi = T(0), izi = T(1), szh = T(0);
            if (PD == 0 || PD == 3) { szi = p.sig[2][2 * k]; izi = p.siginv[2][2 * k]; szh = p.sig[2][2 * k + 1]; }
            const bool kin = (k >= 1 && k <= p.n[2] - 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int i = i0 + v;
                const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                const bool iin = (i >= 1 && i <= p.n[0] - 1);
                // the face sigma at the integer position (tangential components)
                const T sf = PD == 1 ? sxi[v] : PD == 2 ? syi : szi, isf = PD == 1 ? ixi[v] : PD == 2 ? iyi : izi;
                T S0 = T(0), S1 = T(0), J = T(0);
                if (i <= p.n[0] - 1 && jin && kin) {     // Ex: k-dir y, u-dir z, w-dir x
                    const T curl = C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym.v[v]);
                    if (SRC && smask) source_parts(p, smask, 0, i, j, k, set, step, S0, S1, J);
                    if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, syi, iyi, szi, izi, sxh[v], ux.v[v], mx[v], chi_u, eps_u, S0, S1, J, cfu);
                    else if (PD == 1) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, T(0), T(1), T(0), T(1), sxh[v], ux.v[v], mx[v], chi_u, eps_u, S0, S1, J, cfu);
                    else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, sf, isf, T(0), T(1), T(0), ux.v[v], mx[v], chi_u, eps_u, S0, S1, J, cfu);
                }
                if (j <= p.n[1] - 1 && iin && kin) {     // Ey: k-dir z, u-dir x, w-dir y
                    const T curl = C * (((hxm.v[v] - hx0.v[v]) + hz0.v[v]) - hzi);
                    if (SRC && smask) source_parts(p, smask, 1, i, j, k, set, step, S0, S1, J);
                    if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, szi, izi, sxi[v], ixi[v], syh, uy.v[v], my[v], chi_u, eps_u, S0, S1, J, cfu);
                    else if (PD == 2) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, T(0), T(1), T(0), T(1), syh, uy.v[v], my[v], chi_u, eps_u, S0, S1, J, cfu);
                    else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, sf, isf, T(0), T(1), T(0), uy.v[v], my[v], chi_u, eps_u, S0, S1, J, cfu);
                }
                if (k <= p.n[2] - 1 && iin && jin) {     // Ez: k-dir x, u-dir y, w-dir z
                    const T curl = C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]);
                    if (SRC && smask) source_parts(p, smask, 2, i, j, k, set, step, S0, S1, J);
                    if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, sxi[v], ixi[v], syi, iyi, szh, uz.v[v], mz[v], chi_u, eps_u, S0, S1, J, cfu);
                    else if (PD == 3) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, T(0), T(1), T(0), T(1), szh, uz.v[v], mz[v], chi_u, eps_u, S0, S1, J, cfu);
                    else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, sf, isf, T(0), T(1), T(0), uz.v[v], mz[v], chi_u, eps_u, S0, S1, J, cfu);
                }
            }
            ex.store(pE); ey.store((pE + fcs)); ez.store((pE + fcs2));
            if (PD == 0 || PD == 1) dx.store(pD);
            if (PD == 0 || PD == 2) dy.store((pD + bcs));
            if (PD == 0 || PD == 3) dz.store((pD + bcs2));
            if (PD == 0) { ux.store(pU); uy.store((pU + bcs)); uz.store((pU + bcs2)); }
            if (POL) pol.store(p, parity, xg - pol_shift(p, set));
            if (GEN) {
#pragma unroll
                for (int v = 0; v < V; ++v) { mx[v] = nx_[v]; my[v] = ny_[v]; mz[v] = nz_[v]; }
            }
        }
        hxm = hx0; hym = hy0;
        pE += plane; pH += plane;
        pD += bplane; pU += bplane;
        pm += plane; xg += plane;
    }
}

template <typename T, int V, int LX, int NS, bool FACE, bool UNI>
__global__ void __launch_bounds__(256, (V * sizeof(T) == 8) ? (NS == 0 ? (FACE ? 4 : 3) : 2) : (NS == 0 ? (FACE ? SJ_EFACE_BLOCKS : 2) : 1)) e_pml_tile(KParams<T> p, PmlBoxSet<T> bs, const WorkItem *__restrict__ items,
                                                                   int k_lo, int k_hi) {
    const WorkItem it = items[blockIdx.x];
    const PmlBox<T> &b = bs.b[it.box];
    const T chi_u = UNI ? p.mt_chi[it.mat] : T(0), eps_u = UNI ? p.mt_eps[it.mat] : T(0);
    if (src_in_chunk(p, max(it.kb, k_lo), min(it.ke, k_hi))) {
        if (!FACE) e_pml_body<T, V, LX, NS, 0, true, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
        else if (it.kind == 1) e_pml_body<T, V, LX, NS, 1, true, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
        else if (it.kind == 2) e_pml_body<T, V, LX, NS, 2, true, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
        else e_pml_body<T, V, LX, NS, 3, true, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
        return;
    }
    if (!FACE) e_pml_body<T, V, LX, NS, 0, false, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
    else if (it.kind == 1) e_pml_body<T, V, LX, NS, 1, false, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
    else if (it.kind == 2) e_pml_body<T, V, LX, NS, 2, false, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
    else e_pml_body<T, V, LX, NS, 3, false, UNI>(p, b, it, k_lo, k_hi, chi_u, eps_u);
}

// per-plane material flags of work items (one block per item): flags[item * zmax + (k - kb)] = 1 if the tile's plane k
// holds more than one material id (over the three E components), else (id << 8) | (2 if the material has poles)
static __global__ void plane_flags_kernel(const uint8_t *m0, const uint8_t *m1, const uint8_t *m2, const WorkItem *items, int tile_w,
                                          int tile_h, int n0, int n1, int pitch, long long plane, int kz0, int first_disp,
                                          unsigned *flags, int zmax) {
    const WorkItem it = items[blockIdx.x];
    const int i_hi = min(min(it.i0 + tile_w, it.i_hi), n0 + 1), j_hi = min(min(it.j0 + tile_h, it.j_hi), n1 + 1);
    const int nx = max(i_hi - it.i0, 0), ny = max(j_hi - it.j0, 0), nz = min(max(it.ke - it.kb, 0), zmax);
    for (int kk = 0; kk < nz; ++kk) {
        const long long base = (long long)(it.kb + kk - kz0 + 1) * plane;
        const int ref = (nx && ny) ? m0[base + (long long)it.j0 * pitch + it.i0] : 0;
        int gen = 0;
        for (int t = threadIdx.x; t < nx * ny && !gen; t += blockDim.x) {
            const long long x = base + (long long)(it.j0 + t / nx) * pitch + it.i0 + t % nx;
            if (m0[x] != ref || m1[x] != ref || m2[x] != ref) gen = 1;
        }
        gen = __syncthreads_or(gen);
        if (threadIdx.x == 0) flags[(long long)blockIdx.x * zmax + kk] = gen ? 1u : (((unsigned)ref << 8) | (ref >= first_disp ? 2u : 0u));
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
    if (m.wait_flag) {                  // z-slab run: the upper slab's bottom E plane of the last step must be in my halo
        const long long t0 = clock64();
        while (*(volatile const unsigned long long *)m.wait_flag < (unsigned long long)step && clock64() - t0 < 20000000000LL) {}
        __threadfence_system();
    }
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

static __global__ void tick_kernel(long long *step) { *step += 1; }

// remap material bytes through a 256-entry LUT (non-dispersive materials first)
static __global__ void remap_bytes(uint8_t *a, long long n, const uint8_t *lut) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] = lut[a[t]];
}
