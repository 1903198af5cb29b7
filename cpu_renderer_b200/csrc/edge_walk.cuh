// Active-edge walk of ONE triangle (device side), shared by the set-up kernel's row walk.
//
// Restates DrawModel's list handling (projekt.cpp:198-303, 542-572) for the <= 3 edges of a
// single-triangle object, with array storage instead of the intrusive linked list and with
// defined behaviour where the reference dereferences a null list pointer (SURVEY.md 8c,
// "level 1"): a row with fewer than two active edges draws nothing and steps nothing.
#pragma once

#include "raster_device.cuh"

namespace b200r {

struct ActiveEdge
{
    float x, z, c0, c1, c2, c3;         // running XMin, ZMin, MinColor
    float dx, dz, d0, d1, d2, d3;       // Gradient, ZGradient, ColorGradient
    float n0, n1, n2, g0, g1, g2;       // Phong: running MinNormal, NormalGradient
    int ymax;
    int id;                             // slot in the triangle record
};

// nrm: the triangle's three staged vertex normals (Phong), or nullptr
__device__ __forceinline__ void load_edge(ActiveEdge &a, const uint32_t *rec, int id, const float *nrm)
{
    const uint32_t *E = rec + R_EDGE0 + id*kEdgeWords;
    a.ymax = (int)E[E_YMAX];
    a.x = __uint_as_float(E[E_X]);        a.dx = __uint_as_float(E[E_DX]);
    a.z = __uint_as_float(E[E_Z]);        a.dz = __uint_as_float(E[E_DZ]);
    a.c0 = __uint_as_float(E[E_C + 0]);   a.c1 = __uint_as_float(E[E_C + 1]);
    a.c2 = __uint_as_float(E[E_C + 2]);   a.c3 = __uint_as_float(E[E_C + 3]);
    a.d0 = __uint_as_float(E[E_DC + 0]);  a.d1 = __uint_as_float(E[E_DC + 1]);
    a.d2 = __uint_as_float(E[E_DC + 2]);  a.d3 = __uint_as_float(E[E_DC + 3]);
    a.id = id;
    if(nrm)
    {
        // projekt.cpp:4017-4018, 4104-4109: MinNormal = upper vertex normal (not advanced by the
        // top clip), NormalGradient = (MaxNormal - MinNormal)/YDifference
        const int mn = (int)((E[E_LEFT] >> 8) & 3u), mx = (int)((E[E_LEFT] >> 16) & 3u);
        const float ydiff = fsub(__int2float_rn((int)E[E_YMAX]), __int2float_rn((int)E[E_YMIN]));
        a.n0 = nrm[3*mn + 0]; a.n1 = nrm[3*mn + 1]; a.n2 = nrm[3*mn + 2];
        a.g0 = fdiv_zq(fsub(nrm[3*mx + 0], a.n0), ydiff);
        a.g1 = fdiv_zq(fsub(nrm[3*mx + 1], a.n1), ydiff);
        a.g2 = fdiv_zq(fsub(nrm[3*mx + 2], a.n2), ydiff);
    }
}

// Normalize(a) = a * (1/sqrt(a.a)) on three scalars (SURVEY.md Appendix A pin)
__device__ __forceinline__ void normalize3f(float &x, float &y, float &z)
{
    const float s = fdiv(1.0f, __fsqrt_rn(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z))));
    x = fmul(s, x); y = fmul(s, y); z = fmul(s, z);
}

// projekt.cpp:542-552: one row down an edge.
template<bool PHONG>
__device__ __forceinline__ void step_edge(ActiveEdge &a)
{
    a.x = fadd(a.x, a.dx);   a.z = fadd(a.z, a.dz);
    a.c0 = fadd(a.c0, a.d0); a.c1 = fadd(a.c1, a.d1);
    a.c2 = fadd(a.c2, a.d2); a.c3 = fadd(a.c3, a.d3);
    if(PHONG)                                               // :551-552
    {
        a.n0 = fadd(a.n0, a.g0); a.n1 = fadd(a.n1, a.g1); a.n2 = fadd(a.n2, a.g2);
        normalize3f(a.n0, a.n1, a.n2);
    }
}

// The active list at row y (projekt.cpp:202-296): insert, in record order, every edge whose
// YMin == y before the first entry it sorts strictly before (XMin, then Gradient, then Left;
// :212-216, :229-233), then drop entries with YMax <= y (:262-296).  L/R receive the first two
// survivors, keeping the running values of edges that were already active.  For finite vertices
// at most two edges survive a row (a triangle's upper and lower short edges never share a row);
// a third survivor is ignored.  next_ev = the next row at which the list can change.
__device__ __forceinline__ void active_list_event(int y, const uint32_t *rec, int nedges,
                                                  ActiveEdge &L, ActiveEdge &R, int &nact, int &next_ev,
                                                  const float *nrm)
{
    int ids[3] = {0, 0, 0};
    float xs[3] = {0.0f, 0.0f, 0.0f};
    int n = 0;
    if(nact >= 1) { ids[0] = L.id; xs[0] = L.x; n = 1; }
    if(nact >= 2) { ids[1] = R.id; xs[1] = R.x; n = 2; }
#pragma unroll
    for(int e = 0; e < 3; ++e)
    {
        if(e >= nedges) break;
        const uint32_t *E = rec + R_EDGE0 + e*kEdgeWords;
        if((int)E[E_YMIN] != y || n >= 3) continue;
        const float nx = __uint_as_float(E[E_X]), ng = __uint_as_float(E[E_DX]);
        const int nl = (int)(E[E_LEFT] & 1u);
        int at = n;
#pragma unroll
        for(int k = 2; k >= 0; --k)
        {
            if(k >= n) continue;
            const uint32_t *O = rec + R_EDGE0 + ids[k]*kEdgeWords;
            const float ox = xs[k], og = __uint_as_float(O[E_DX]);
            const int ol = (int)(O[E_LEFT] & 1u);
            if(nx < ox || (nx == ox && (ng < og || (ng == og && nl < ol)))) at = k;   // ends as the first such k
        }
#pragma unroll
        for(int k = 2; k >= 1; --k)
        {
            if(k <= n && k > at) { ids[k] = ids[k - 1]; xs[k] = xs[k - 1]; }
        }
#pragma unroll
        for(int k = 0; k < 3; ++k) if(k == at) { ids[k] = e; xs[k] = nx; }
        ++n;
    }
    int kept = 0, kid0 = -1, kid1 = -1;
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
        if(k >= n) continue;
        const int ym = (int)rec[R_EDGE0 + ids[k]*kEdgeWords + E_YMAX];
        if(ym <= y) continue;
        if(kept == 0) kid0 = ids[k]; else if(kept == 1) kid1 = ids[k];
        ++kept;
    }
    const ActiveEdge oldL = L, oldR = R;
    const int oldn = nact;
    if(kept >= 1)
    {
        if(oldn >= 1 && kid0 == oldL.id) L = oldL;
        else if(oldn >= 2 && kid0 == oldR.id) L = oldR;
        else load_edge(L, rec, kid0, nrm);
    }
    if(kept >= 2)
    {
        if(oldn >= 1 && kid1 == oldL.id) R = oldL;
        else if(oldn >= 2 && kid1 == oldR.id) R = oldR;
        else load_edge(R, rec, kid1, nrm);
    }
    nact = (kept > 2) ? 2 : kept;
    int ev = 0x7fffffff;
#pragma unroll
    for(int e = 0; e < 3; ++e)
    {
        if(e >= nedges) break;
        const int ym = (int)rec[R_EDGE0 + e*kEdgeWords + E_YMIN];
        if(ym > y && ym < ev) ev = ym;
    }
    if(nact >= 1 && L.ymax < ev) ev = L.ymax;
    if(nact >= 2 && R.ymax < ev) ev = R.ymax;
    next_ev = ev;
}

// Upper bound of every depth value of a span: z_k = fl(z_{k-1} + zi), k <= n.  Each add rounds by
// at most half an ulp of a value no larger than |z0| + n|zi| (plus the bound itself), so
//   z_k <= max(z0, z0 + n*zi) + n * 2^-23 * (|z0| + n*|zi|)        (2x the worst case),
// evaluated with every operation rounded towards +inf.  NaN / Inf input gives +inf or NaN: the
// raster kernel culls only on a strict, ordered "bound < row minimum", so such spans are kept.
__device__ __forceinline__ float span_depth_bound(float z0, float zi, int n)
{
    const float fn = (float)max(n, 0);
    const float end = __fadd_ru(z0, __fmul_ru(fn, zi));
    const float mag = __fadd_ru(fabsf(z0), __fmul_ru(fn, fabsf(zi)));
    const float slack = __fmul_ru(__fmul_ru(fn, 1.1920929e-7f), mag);
    return __fadd_ru(__fadd_ru(fmaxf(z0, end), slack), 1.0e-30f);    // absolute term: subnormal rounding
}

} // namespace b200r
