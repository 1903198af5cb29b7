// Geometry / triangle set-up kernel (sm_100a).
//
// Restates, per triangle, the body of FillEdgeTable for the Gouraud branch
// (projekt.cpp:3894-4115): translate by Object->P, ProjectVertex (74-93), back-face test
// (3926-3927, 3943), per-edge y-ordering, reject / top-clip, start values and per-row
// gradients (3947-4111), vertex lighting (4020-4064), and the MergeSort order (2-72) of the
// <= 3 edges of the triangle ("one triangle = one object").
//
// It then walks the triangle's rows ONCE (DrawModel's active-edge maintenance, span set-up and
// edge stepping, projekt.cpp:198-412, 542-572; no pixels) and emits one 64-byte SPAN record per
// row -- column range, start values, per-pixel increments -- grouped into segments (one per edge
// pair and tile-row band) that carry the exact tile columns their spans touch.  The raster
// kernel never re-derives any of this: it only replays per-pixel adds and depth-tests.
//
// Mapping: one thread per triangle, one CTA per 128 triangles.  A warp handles 32 triangles:
// their 32 x 120 B of vertex attributes are fetched with fully coalesced loads into shared
// memory.  Segment slots are handed out with a CTA-wide prefix sum of per-triangle segment
// counts (warp __shfl_up_sync scans) and ONE global atomicAdd per CTA.  (A literal
// warp-per-triangle mapping would spend 32 lanes on ~3 edges x 2 ends of scalar work; see
// DESIGN.md.)
#include "raster_device.cuh"
#include "edge_walk.cuh"

#include <algorithm>

namespace b200r {

constexpr int kSetupThreads = 128;
constexpr int kEdgeRec = 3*kEdgeWords;   // 45 words of edge data per triangle in shared memory
constexpr int kSortBins = 64;        // walkers are sorted by min(rows, 63); 3 bins per lane of one warp >= 65

struct V3 { float x, y, z; };

// projekt.cpp:74-93
__device__ __forceinline__ V3 project_vertex(V3 cam, const ViewParams &v)
{
    V3 r = {0.0f, 0.0f, 0.0f};
    float d = fsub(v.dist, cam.z);                          // :81
    if(d > 0.2f)                                            // :82, :86
    {
        float s = fmul(fdiv(1.0f, d), v.focal);             // :88 (1/d)*FocalLength first
        float px = fmul(s, cam.x);
        float py = fmul(s, cam.y);
        r.x = fadd(v.cx, fmul(v.m2p, px));                  // :89
        r.y = fadd(v.cy, fmul(v.m2p, py));
        r.z = fadd(d, fmul(v.m2p, 0.0f));
    }
    return r;
}

__device__ __forceinline__ float inner3(V3 a, V3 b)         // left to right
{
    return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
// Normalize(a) = a * (1/sqrt(a.a))  (SURVEY.md Appendix A pin; projekt.cpp:3926, 4029)
__device__ __forceinline__ V3 normalize3(V3 a)
{
    float s = fdiv(1.0f, __fsqrt_rn(inner3(a, a)));
    V3 r = { fmul(s, a.x), fmul(s, a.y), fmul(s, a.z) };
    return r;
}

// projekt.cpp:4022-4062, non-bitmap branch; value depends only on the vertex.
__device__ __forceinline__ void light_vertex(V3 cam, V3 nrm, const float col[4], const ViewParams &v,
                                             float out[4])
{
    float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for(int l = 0; l < v.nlights; ++l)
    {
        const DevLight &L = v.lights[l];
        V3 to = { fsub(L.px, cam.x), fsub(L.py, cam.y), fsub(L.pz, cam.z) };
        V3 dir = normalize3(to);                            // :4029
        if(l == 0)                                          // :4032-4044
        {
#pragma unroll
            for(int i = 0; i < 4; ++i) c[i] = fmul(col[i], v.amb[i]);
        }
        float dot = clamp01(inner3(dir, nrm));              // :4047
        const float I[4] = { L.ir, L.ig, L.ib, L.ia };
#pragma unroll
        for(int i = 0; i < 4; ++i)                          // :4058
        {
            c[i] = clamp01(fadd(c[i], fmul(dot, fmul(col[i], I[i]))));
        }
    }
#pragma unroll
    for(int i = 0; i < 4; ++i) out[i] = c[i];
}

// ---- z range of a frame's vertices (for the depth buckets): ordered-integer keys ----
__device__ __forceinline__ unsigned float_key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256)
zrange_kernel(MeshParams m, unsigned *zkeys)
{
    // zkeys[0] = max key, zkeys[1] = max of ~key (i.e. ~min key) over camera z = vertex z + P.z
    // (finite values only); both start at 0 so one memset initialises them.
    // The positions are streamed with 128-bit loads, four in flight per thread, and the z components picked out
    // of them (float 3j + 2 of the stream); one strided 4-byte load per vertex read the 36 MB of C2 at 2.1 TB/s.
    unsigned kmax = 0u, kmin = 0xffffffffu;
    const size_t nf = (size_t)m.ntri*9;                       // floats in the stream
    const size_t tid = (size_t)blockIdx.x*blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x*blockDim.x;
    auto take = [&](float zv)
    {
        const float z = zv + m.pz;
        if(fabsf(z) < 3.0e38f) { const unsigned k = float_key(z); kmax = max(kmax, k); kmin = min(kmin, k); }
    };
    size_t done = 0;                                          // floats covered by the vector part
    if((((uintptr_t)m.pos) & 15) == 0)
    {
        const float4 *g4 = reinterpret_cast<const float4 *>(m.pos);
        const size_t n4 = nf/4;
        for(size_t i0 = tid; i0 < n4; i0 += 4*nthreads)
        {
            float4 q[4];
#pragma unroll
            for(int u = 0; u < 4; ++u) { const size_t i = i0 + u*nthreads; q[u] = (i < n4) ? __ldg(g4 + i) : make_float4(0, 0, 0, 0); }
#pragma unroll
            for(int u = 0; u < 4; ++u)
            {
                const size_t i = i0 + u*nthreads;
                if(i >= n4) break;
                const unsigned r = (unsigned)((4*i) % 3);     // component k of this vector is a z iff (r + k) % 3 == 2
                if(r == 2) { take(q[u].x); take(q[u].w); }
                else if(r == 1) take(q[u].y);
                else take(q[u].z);
            }
        }
        done = n4*4;
    }
    for(size_t f = done + tid; f < nf; f += nthreads)
        if(f % 3 == 2) take(__ldg(m.pos + f));
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    if((threadIdx.x & 31) == 0 && kmax != 0u)
    {
        if(kmax > __ldcg(&zkeys[0])) atomicMax(&zkeys[0], kmax);
        if(~kmin > __ldcg(&zkeys[1])) atomicMax(&zkeys[1], ~kmin);
    }
}

__global__ void zrange_finish_kernel(unsigned *zkeys)
{
    const unsigned kmax = zkeys[0], kmin = ~zkeys[1];
    float zmax = 0.0f, inv = 0.0f;
    if(kmax != 0u && kmax >= kmin)
    {
        zmax = key_float(kmax);
        const float range = zmax - key_float(kmin);
        inv = (range > 0.0f) ? 1.0f/range : 0.0f;
        if(!(fabsf(inv) < 3.0e38f)) inv = 0.0f;
    }
    reinterpret_cast<float *>(zkeys)[0] = zmax;
    reinterpret_cast<float *>(zkeys)[1] = inv;
}

// Can a triangle's rows reach this GPU's band (or the row just above it, which may drop an alias
// pixel into the band)?  Rows lie in [Round(min y), Round(max y)), so two rows of margin are
// conservative; NaN compares false and keeps the triangle.  ONE definition, used by the set-up
// kernel's phase 1 and by select_kernel.
__device__ __forceinline__ bool off_band_rows(float y0, float y1, float y2, const ViewParams &v)
{
    const float ymin_p = fminf(y0, fminf(y1, y2));
    const float ymax_p = fmaxf(y0, fmaxf(y1, y2));
    return ymax_p < (float)(v.band_y0 - 2) || ymin_p > (float)(v.band_y1 + 1);
}

// Row-band pre-selection.  With G GPUs rendering row bands of one frame, every GPU used to run the
// set-up kernel's whole front end (120 B of attributes staged, three projections, the back-face
// test) on every triangle although about 1/G of them can reach its band: C4's set-up scaled 1.65x
// on 8 GPUs.  This pass reads the positions only (36 B), applies the same band test to the same
// projected rows, and compacts the surviving indices; the set-up kernel then gathers just those.
// The positions are read ONCE: the same pass accumulates the frame's z range (what zrange_kernel does
// for meshes that are not pre-selected) -- both kernels read every position, and at 8 bands that
// read is a fifth of a GPU's whole frame.
__global__ void __launch_bounds__(256, 8)
select_kernel(ViewParams v, MeshParams m, unsigned *list, unsigned *count, unsigned *zkeys)
{
    // Persistent CTAs over blocks of 256 triangles.  A block's positions (9 KB) are staged through shared memory
    // with coalesced 128-bit loads (a thread's own nine floats are 36 bytes apart), and the NEXT block's loads are
    // issued before this block's projections start, so that every CTA has a block in flight at all times.  (One
    // block per CTA, load -> barrier -> compute -> atomic, read the positions at 2.3 TB/s: while a CTA computed it
    // had nothing in flight, and that pass is a tenth of a GPU's whole frame on 8 row bands.)
    __shared__ __align__(16) float s_p[256*9];
    __shared__ unsigned s_cnt[8], s_base;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned nblocks = (m.ntri + 255u)/256u;
    const bool aligned = (((uintptr_t)m.pos) & 15) == 0;      // a block starts 9216 bytes after the previous one
    unsigned kmax = 0u, kmin = 0xffffffffu;
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, rc = ra;
    auto full = [&](unsigned blk) { return aligned && blk*256u + 256u <= m.ntri; };
    auto issue = [&](unsigned blk)
    {
        const float4 *g4 = reinterpret_cast<const float4 *>(m.pos + (size_t)blk*256u*9u);
        ra = __ldg(g4 + threadIdx.x); rb = __ldg(g4 + 256 + threadIdx.x);
        if(threadIdx.x < 64) rc = __ldg(g4 + 512 + threadIdx.x);
    };
    unsigned blk = blockIdx.x;
    if(blk < nblocks && full(blk)) issue(blk);
    for(; blk < nblocks; blk += gridDim.x)
    {
        const unsigned base = blk*256u;
        const unsigned tri = base + threadIdx.x;
        __syncthreads();                                    // the previous block's readers are done with s_p
        if(full(blk))
        {
            float4 *s4 = reinterpret_cast<float4 *>(s_p);
            s4[threadIdx.x] = ra; s4[256 + threadIdx.x] = rb;
            if(threadIdx.x < 64) s4[512 + threadIdx.x] = rc;
        }
        else
        {
            const float *gp = m.pos + (size_t)base*9;
            const unsigned nf = min(256u, m.ntri - base)*9u;
            for(unsigned i = threadIdx.x; i < nf; i += 256u) s_p[i] = __ldg(gp + i);
        }
        __syncthreads();
        const unsigned next = blk + gridDim.x;
        if(next < nblocks && full(next)) issue(next);       // in flight while this block is projected
        bool keep = false;
        if(tri < m.ntri)
        {
            const float *sp = s_p + threadIdx.x*9;
            float y[3];
#pragma unroll
            for(int k = 0; k < 3; ++k)
            {
                const float pz = sp[3*k + 2];
                V3 cam = { fadd(sp[3*k + 0], m.px), fadd(sp[3*k + 1], m.py), fadd(pz, m.pz) };
                y[k] = project_vertex(cam, v).y;
                const float z = pz + m.pz;                      // as zrange_kernel
                if(fabsf(z) < 3.0e38f) { const unsigned kk = float_key(z); kmax = max(kmax, kk); kmin = min(kmin, kk); }
            }
            keep = !off_band_rows(y[0], y[1], y[2], v);
        }
        // ONE returning atomic per block hands out the list slots: all blocks add to the same word, and
        // same-address atomics are served one at a time (a per-warp atomic made this pass 0.6 ms on C4)
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if(lane == 0) s_cnt[warp] = (unsigned)__popc(bal);
        __syncthreads();
        if(threadIdx.x == 0)
        {
            unsigned tot = 0;
            for(int w = 0; w < 8; ++w) { const unsigned c_ = s_cnt[w]; s_cnt[w] = tot; tot += c_; }
            s_base = tot ? atomicAdd(count, tot) : 0u;
        }
        __syncthreads();
        B200R_ASSERT(!keep || s_base + s_cnt[warp] + __popc(bal & ((1u << lane) - 1u)) < m.ntri);
        if(keep) list[s_base + s_cnt[warp] + __popc(bal & ((1u << lane) - 1u))] = tri;
    }
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    // only the warps that still widen the range touch the two words (one same-address atomic per warp
    // serialises: 1.2 M of them cost C4 on 8 GPUs 0.3 ms)
    if(lane == 0 && kmax != 0u)
    {
        if(kmax > __ldcg(&zkeys[0])) atomicMax(&zkeys[0], kmax);
        if(~kmin > __ldcg(&zkeys[1])) atomicMax(&zkeys[1], ~kmin);
    }
}

// PHONG: the mesh is drawn with per-pixel Phong shading (render_entry_3d_object::PhongShading,
// projekt.cpp:4012-4019): edge colours stay unlit, edges and spans carry interpolated normals.
// TEX: the mesh is textured (MeshParams::uv): the colour interpolants carry u/z, v/z, 1/z.  A
// template parameter, not a run-time branch: the untextured kernel must not pay for the code (this
// kernel is instruction-cache bound; the branches cost 6 % of its time when they were run-time).
// LISTED: the kernel processes the triangles of MeshParams::tri_list (row-band pre-selection) and
// gathers their attributes; a template parameter for the same reason.
// SPLIT: row-parallel walks for frames of FEW, TALL triangles (the host picks it from the triangle count).
// A triangle's rows are sequential state (an edge's value at row r is r rounded adds from its start), so
// a thread that walks a 150-row triangle alone is a 150-row latency chain, and 50 000 such threads leave
// most of the machine idle (C3: set-up 0.30 ms at 16 % of the resident warps; a 500-triangle 4K frame
// still took 0.2 ms).  With SPLIT a CTA stages only MeshParams::tris_per_cta (<= 64) triangles and a
// triangle taller than kSplitMinRows is cut into WALKERS of MeshParams::part_rows screen rows (slabs
// aligned to tile rows, so no segment straddles two walkers).  Every walker replays the edge stepping
// from the triangle's first row -- 14 adds and the crossing test per row, no span set-up -- and emits
// the spans of its own slab only, at offsets it derives from the same closed-form row counts that sized
// the triangle's allocation.  A template parameter: the small-triangle kernels must not pay for it.
// DIRECT_COL: untextured meshes do not stage their vertex colours -- the thread that lights a triangle reads its
// 48 bytes itself (three 128-bit loads issued at the top of phase 2, used by the lighting ~150 instructions
// later) -- which takes the kernel's shared memory from 43.4 to 37.3 KB; held to 80 registers (16 bytes of
// spill) six CTAs fit an SM instead of five.  Measured: C2 and C3 unchanged (the kernel is not occupancy
// bound: 0.334 ms either way), a 1/8 row band of C4 (LISTED: the gather no longer stages colours) 1.85 -> 1.76 ms.
template<bool PHONG, bool TEX, bool LISTED, bool SPLIT>
__global__ void __launch_bounds__(kSetupThreads, (!PHONG && !TEX) ? 6 : 5)
setup_kernel(ViewParams v, MeshParams m, SetupOutputs out)
{
    constexpr bool DIRECT_COL = !TEX;
    __shared__ __align__(16) float s_pos[kSetupThreads*9];
    __shared__ __align__(16) float s_col[DIRECT_COL ? 4 : kSetupThreads*12];
    __shared__ __align__(16) float s_nrm[kSetupThreads*9];
    // per triangle only the 3 x 15 edge words live in shared memory (odd stride: conflict-free);
    // `rec` pointers below are biased by -R_EDGE0 so that rec[R_EDGE0 + ...] addresses them.
    __shared__ __align__(16) uint32_t s_edge[kSetupThreads*kEdgeRec];
    __shared__ unsigned s_binned, s_pairs, s_seg_base, s_span_base, s_fits;
    __shared__ unsigned s_hist[kSortBins + 2 + 32], s_order[kSetupThreads], s_alive[kSetupThreads];
    __shared__ int s_w_tri[kSetupThreads];
    __shared__ int s_w_first[kSetupThreads], s_w_end[kSetupThreads];
    __shared__ unsigned s_w_info[kSetupThreads], s_w_spans[kSetupThreads], s_w_seg_at[kSetupThreads], s_w_span_at[kSetupThreads];
    __shared__ unsigned s_warp_sum[kSetupThreads/32], s_warp_sum2[kSetupThreads/32];

    __shared__ unsigned s_id[kSetupThreads];             // list mode: original index of each staged triangle
    __shared__ unsigned s_wmap[SPLIT ? kSetupThreads : 1];   // SPLIT: the batch's walkers, triangle slot | part << 8
    __shared__ unsigned s_wtotal;
    constexpr bool listed = LISTED;
    const unsigned total = listed ? *m.tri_count : m.ntri;
    const unsigned tpc = SPLIT ? (unsigned)m.tris_per_cta : (unsigned)kSetupThreads;   // triangles per CTA
    const unsigned base = blockIdx.x*tpc;
    if(base >= total) return;                              // list mode launches for ntri; the tail has nothing to do
    const unsigned n = min(tpc, total - base);
    const int t = threadIdx.x;
    if(t == 0) { s_binned = 0; s_pairs = 0; }
    int my_segs = 0, my_spans = 0, walk_end = 0, first_row = 0, max_y = 0, nedges = 0, nonfinite = 0;
    bool have_walk = false;
    unsigned seg_at = 0, span_at = 0;
    int nalive = 0;

    // Coalesced attribute fetch.  Full CTAs with 16-byte aligned streams issue all nine 128-bit
    // loads of a thread back to back and only then store to shared memory: one DRAM round trip
    // per CTA.  (A load->store loop costs one round trip per iteration -- measured: the kernel's
    // time was CTA waves x ~30 us of serialized loads, whatever the arithmetic did.)
    {
        const float *gp = m.pos + (size_t)base*9;
        const float *gc = m.col + (size_t)base*12;
        const float *gn = m.nrm + (size_t)base*9;
        const bool aligned = ((((uintptr_t)gp) | ((uintptr_t)gc) | ((uintptr_t)gn)) & 15) == 0;
        if(listed)
        {
            // gather: thread t stages the triangle list[base + t] (36 + 48 + 36 contiguous bytes each)
            if((unsigned)t < n)
            {
                const unsigned id = __ldg(m.tri_list + base + t);
                s_id[t] = id;
                const float *p9 = m.pos + (size_t)id*9, *n9 = m.nrm + (size_t)id*9;
                float rp[9], rn[9];
#pragma unroll
                for(int i = 0; i < 9; ++i) { rp[i] = __ldg(p9 + i); rn[i] = __ldg(n9 + i); }
                if(TEX)
                {
                    const float *u6 = m.uv + (size_t)id*6;
                    float ru[6];
#pragma unroll
                    for(int i = 0; i < 6; ++i) ru[i] = __ldg(u6 + i);
#pragma unroll
                    for(int i = 0; i < 6; ++i) s_col[t*12 + i] = ru[i];
                }
                else if(!DIRECT_COL)
                {
                    const float *c12 = m.col + (size_t)id*12;
                    float rc[12];
#pragma unroll
                    for(int i = 0; i < 12; ++i) rc[i] = __ldg(c12 + i);
#pragma unroll
                    for(int i = 0; i < 12; ++i) s_col[t*12 + i] = rc[i];
                }
#pragma unroll
                for(int i = 0; i < 9; ++i) { s_pos[t*9 + i] = rp[i]; s_nrm[t*9 + i] = rn[i]; }
            }
        }
        else if(TEX)
        {
            // textured mesh: the vertex colours never reach the image (MeshParams::uv); the colour
            // slots of the staging area take the UVs, two floats per vertex
            const float *gu = m.uv + (size_t)base*6;
            for(unsigned i = t; i < n*9; i += kSetupThreads) { s_pos[i] = __ldg(gp + i); s_nrm[i] = __ldg(gn + i); }
            for(unsigned i = t; i < n*6; i += kSetupThreads) s_col[(i/6)*12 + (i%6)] = __ldg(gu + i);
        }
        else if(n == (unsigned)kSetupThreads && aligned)
        {
            constexpr int kPos4 = kSetupThreads*9/4, kCol4 = kSetupThreads*12/4;     // 288, 384 float4
            const float4 *gp4 = reinterpret_cast<const float4 *>(gp);
            const float4 *gc4 = reinterpret_cast<const float4 *>(gc);
            const float4 *gn4 = reinterpret_cast<const float4 *>(gn);
            float4 rp[3], rc[3], rn[3];
#pragma unroll
            for(int k = 0; k < 3; ++k)
            {
                const int i = t + k*kSetupThreads;
                rp[k] = (i < kPos4) ? __ldg(gp4 + i) : make_float4(0, 0, 0, 0);
                if(!DIRECT_COL) rc[k] = (i < kCol4) ? __ldg(gc4 + i) : make_float4(0, 0, 0, 0);
                rn[k] = (i < kPos4) ? __ldg(gn4 + i) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for(int k = 0; k < 3; ++k)
            {
                const int i = t + k*kSetupThreads;
                if(i < kPos4) { reinterpret_cast<float4 *>(s_pos)[i] = rp[k]; reinterpret_cast<float4 *>(s_nrm)[i] = rn[k]; }
                if(!DIRECT_COL && i < kCol4) reinterpret_cast<float4 *>(s_col)[i] = rc[k];
            }
        }
        else
        {
            for(unsigned i = t; i < n*9; i += kSetupThreads) { s_pos[i] = __ldg(gp + i); s_nrm[i] = __ldg(gn + i); }
            if(!DIRECT_COL) for(unsigned i = t; i < n*12; i += kSetupThreads) s_col[i] = __ldg(gc + i);
        }
    }
    __syncthreads();

    // ---- phase 1 (thread per triangle): project, back-face test (projekt.cpp:3926-3927, :3943,
    //      Eye = (0,0,-1)) and, for multi-GPU row bands, a conservative "can it reach my band" test.
    //      Survivors are compacted so that the expensive phases below run on dense warps. ----
    auto project_triangle = [&](int tri_, V3 cam[3], V3 prj[3])
    {
#pragma unroll
        for(int k = 0; k < 3; ++k)
        {
            cam[k].x = fadd(s_pos[tri_*9 + 3*k + 0], m.px);    // :3900
            cam[k].y = fadd(s_pos[tri_*9 + 3*k + 1], m.py);
            cam[k].z = fadd(s_pos[tri_*9 + 3*k + 2], m.pz);
            prj[k] = project_vertex(cam[k], v);                // :3907
        }
    };
    // Spans and segments of the rows [first, limit) of a triangle, without walking: between consecutive
    // list-change rows the set of active edges {e : YMin <= row < YMax} is constant; a stretch with >= 2 of
    // them yields one span per row.  A segment is the run of a triangle's spans inside one tile-row band.
    auto count_rows = [&](const uint32_t *rec_, int nedges_, int first, int limit, int &segs, int &spans)
    {
        int y = first, last_band = -1;
        segs = 0; spans = 0;
        while(y < limit)
        {
            int nxt = limit, act = 0;
#pragma unroll
            for(int e = 0; e < 3; ++e)
            {
                if(e >= nedges_) break;
                const int ymn = (int)rec_[R_EDGE0 + e*kEdgeWords + E_YMIN];
                const int ymx = (int)rec_[R_EDGE0 + e*kEdgeWords + E_YMAX];
                if(ymn <= y && y < ymx) ++act;
                if(ymn > y && ymn < nxt) nxt = ymn;
                if(ymx > ymn && ymx > y && ymx < nxt) nxt = ymx;
            }
            if(act >= 2)
            {
                const int a = max(y, v.band_y0);
                if(a < nxt)
                {
                    const int b0 = (a - v.band_y0) >> v.tile_h_shift, b1 = (nxt - 1 - v.band_y0) >> v.tile_h_shift;
                    segs += b1 - b0 + ((b0 == last_band) ? 0 : 1);
                    last_band = b1;
                    spans += nxt - a;
                }
            }
            y = nxt;
        }
    };
    bool alive = false;
    if((unsigned)t < n)
    {
        V3 cam[3], prj[3];
        project_triangle(t, cam, prj);
        V3 d1 = { fsub(prj[1].x, prj[0].x), fsub(prj[1].y, prj[0].y), fsub(prj[1].z, prj[0].z) };
        V3 d2 = { fsub(prj[2].x, prj[0].x), fsub(prj[2].y, prj[0].y), fsub(prj[2].z, prj[0].z) };
        V3 n1 = normalize3(d1), n2 = normalize3(d2);
        float crx = fsub(fmul(n1.y, n2.z), fmul(n1.z, n2.y));
        float cry = fsub(fmul(n1.z, n2.x), fmul(n1.x, n2.z));
        float crz = fsub(fmul(n1.x, n2.y), fmul(n1.y, n2.x));
        float facing = fadd(fadd(fmul(0.0f, crx), fmul(0.0f, cry)), fmul(-1.0f, crz));
        // A triangle whose projected rows cannot reach this GPU's band (nor the row just above it,
        // which may drop an alias pixel into the band) needs nothing more here.  Rows lie in
        // [Round(min y), Round(max y)), so two rows of margin are conservative; NaN compares false
        // and keeps the triangle.
        const bool off_band = out.recs == nullptr && off_band_rows(prj[0].y, prj[1].y, prj[2].y, v);
        alive = facing > 0.0f && !off_band;
        // b200r_fill_edge_table only: default record header (no edges), overwritten in phase 2
        if(out.recs)
        {
            uint32_t *g = out.recs + (size_t)(m.prim_base + base + t)*kRecWords;
            g[R_NEDGES] = 0; g[R_FIRSTROW] = 0; g[R_MAXY] = 0; g[R_PRIM] = m.prim_base + base + t;
            g[R_EDGE0 + 3*kEdgeWords] = 0;
        }
    }
    {
        const unsigned lane = t & 31, warp = t >> 5;
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        if(lane == 0) s_warp_sum[warp] = __popc(bal);
        __syncthreads();
        unsigned before = 0;
        for(unsigned w = 0; w < warp; ++w) before += s_warp_sum[w];
        if(alive) s_alive[before + __popc(bal & ((1u << lane) - 1u))] = (unsigned)t;
        unsigned total = 0;
        for(unsigned w = 0; w < kSetupThreads/32; ++w) total += s_warp_sum[w];
        nalive = (int)total;
        __syncthreads();
    }

    // ---- phase 2 (thread per surviving triangle): edges, lighting, MergeSort order, counts ----
    const int tri = (t < nalive) ? (int)s_alive[t] : -1;
    if(tri >= 0)
    {
        uint32_t *rec = s_edge + tri*kEdgeRec - R_EDGE0;
        // DIRECT_COL: the triangle's three vertex colours, requested now and used by the lighting below
        float4 cq[3];
        if(DIRECT_COL)
        {
            const float *c12 = m.col + (size_t)(listed ? s_id[tri] : base + (unsigned)tri)*12;
            if((((uintptr_t)m.col) & 15) == 0)
            {
#pragma unroll
                for(int q = 0; q < 3; ++q) cq[q] = __ldg(reinterpret_cast<const float4 *>(c12) + q);
            }
            else
            {
#pragma unroll
                for(int q = 0; q < 3; ++q) cq[q] = make_float4(__ldg(c12 + 4*q), __ldg(c12 + 4*q + 1), __ldg(c12 + 4*q + 2), __ldg(c12 + 4*q + 3));
            }
        }
        V3 cam[3], prj[3];
        project_triangle(tri, cam, prj);
        uint32_t emit = 0;                                  // slot of the k-th edge FillEdgeTable emits, 2 bits each
        int slot_of[3] = {-1, -1, -1};
        int mn[3], mx[3];                                   // per edge: index of upper / lower end

        {
            // pass 1: which edges survive and where MergeSort puts them (needs only y)
            int key[3];
            bool counted[3];
#pragma unroll
            for(int e = 0; e < 3; ++e)
            {
                int i0 = e, i1 = (e + 1)%3;                 // :3936-3941
                int a = i0, b = i1;
                if(prj[a].y > prj[b].y) { int s = a; a = b; b = s; }     // :3957
                mn[e] = a; mx[e] = b;
                counted[e] = (prj[b].y > 0.0f) && (fsub(prj[a].y, prj[b].y) != 0.0f);   // :3968, :4066
                float rmin = __int2float_rn(round_s32(prj[a].y));
                key[e] = __float2int_rz((0.0f > rmin) ? 0.0f : rmin);                   // :3999
            }
            int order[3]; int k = 0;
#pragma unroll
            for(int e = 0; e < 3; ++e) if(counted[e]) order[k++] = e;
            nedges = k;
            if(k == 2)                                      // projekt.cpp:9-19
            {
                if(key[order[0]] > key[order[1]]) { int s = order[0]; order[0] = order[1]; order[1] = s; }
            }
            else if(k == 3)                                 // projekt.cpp:20-59 with Half0 = 1
            {
                int a = order[1], b = order[2], c = order[0];
                if(key[a] > key[b]) { int s = a; a = b; b = s; }
                if(key[c] < key[a]) { order[0] = c; order[1] = a; order[2] = b; }
                else if(key[c] < key[b]) { order[0] = a; order[1] = c; order[2] = b; }
                else { order[0] = a; order[1] = b; order[2] = c; }
            }
            for(int s = 0; s < k; ++s) slot_of[order[s]] = s;
            {
                int q = 0;
#pragma unroll
                for(int e = 0; e < 3; ++e) if(counted[e]) { emit |= (uint32_t)slot_of[e] << (2*q); ++q; }
            }

            if(k > 0)
            {
                float lit[3][4];
#pragma unroll 1
                for(int q = 0; q < 3; ++q)
                {
                    float4 c4;
                    if(DIRECT_COL) c4 = (q == 0) ? cq[0] : (q == 1) ? cq[1] : cq[2];
                    else c4 = *reinterpret_cast<const float4 *>(&s_col[tri*12 + 4*q]);
                    float col[4] = { c4.x, c4.y, c4.z, c4.w };
                    V3 nr = { s_nrm[tri*9 + 3*q + 0], s_nrm[tri*9 + 3*q + 1], s_nrm[tri*9 + 3*q + 2] };
                    if(TEX) { lit[q][0] = s_col[tri*12 + 2*q]; lit[q][1] = s_col[tri*12 + 2*q + 1]; lit[q][2] = 0.0f; lit[q][3] = 0.0f; }
                    else if(PHONG) { lit[q][0] = col[0]; lit[q][1] = col[1]; lit[q][2] = col[2]; lit[q][3] = col[3]; }   // :4014-4015
                    else
                    {
                        if(m.white) { col[0] = col[1] = col[2] = col[3] = 1.0f; }   // :4034-4060: Hadamard(V4(1,1,1,1), ...)
                        light_vertex(cam[q], nr, col, v, lit[q]);
                    }
                }
                int max_row = (int)0x80000000;
#pragma unroll 1
                for(int e = 0; e < 3; ++e)
                {
                    if(slot_of[e] < 0) continue;
                    uint32_t *E = rec + R_EDGE0 + slot_of[e]*kEdgeWords;
                    const V3 minv = prj[mn[e]], maxv = prj[mx[e]];
                    int ymax = round_s32(maxv.y);                           // :3988
                    float clipped = 0.0f, tt = 0.0f;
                    if(minv.y < 0.0f)                                       // :3993-3997
                    {
                        clipped = -minv.y;
                        tt = fdiv(-minv.y, fsub(maxv.y, minv.y));
                    }
                    int ymin = key[e];
                    float ydiff = fsub(__int2float_rn(ymax), __int2float_rn(ymin));                  // :4070
                    float zg = fdiv_zq(fsub(cam[mx[e]].z, cam[mn[e]].z), ydiff);                        // :4072
                    float g = fdiv(fsub(maxv.x, minv.x), fsub(maxv.y, minv.y));                      // :4073
                    float x = fadd(minv.x, fmul(clipped, g));                                        // :4075
                    float z = fadd(cam[mn[e]].z, fmul(clipped, zg));                                 // :4076
                    E[E_YMIN] = (uint32_t)ymin; E[E_YMAX] = (uint32_t)ymax;
                    E[E_X] = __float_as_uint(x); E[E_DX] = __float_as_uint(g);
                    E[E_Z] = __float_as_uint(z); E[E_DZ] = __float_as_uint(zg);
                    if(TEX)
                    {
                        // projekt.cpp:4002-4008, 4078-4089.  z of a projected vertex is
                        // DistanceAboveTarget - camera z (:81, :89).  UMin is u/z, a true division; the
                        // gradient is built from u*(1/z) -- both forms are kept.
                        const float zmin_p = minv.z, zmax_p = maxv.z;
                        const float u0 = lit[mn[e]][0], v0 = lit[mn[e]][1], u1 = lit[mx[e]][0], v1 = lit[mx[e]][1];
                        float umin = fdiv(u0, zmin_p), vmin = fdiv(v0, zmin_p), wmin = fdiv(1.0f, zmin_p);   // :4002-4004
                        const float inv_max = fdiv(1.0f, zmax_p), inv_min = fdiv(1.0f, zmin_p);
                        const float su = fmul(inv_max, u1), sv = fmul(inv_max, v1);                    // :4006
                        const float fu = fmul(inv_min, u0), fv = fmul(inv_min, v0);                    // :4008
                        const float ug = fdiv_zq(fsub(su, fu), ydiff), vg = fdiv_zq(fsub(sv, fv), ydiff);   // :4080-4081
                        umin = fadd(umin, fmul(clipped, ug)); vmin = fadd(vmin, fmul(clipped, vg));    // :4083-4084
                        const float wg = fdiv_zq(fsub(fdiv(1.0f, zmax_p), wmin), ydiff);               // :4086
                        wmin = fadd(wmin, fmul(clipped, wg));                                          // :4088
                        E[E_C + 0] = __float_as_uint(umin); E[E_C + 1] = __float_as_uint(vmin);
                        E[E_C + 2] = __float_as_uint(wmin); E[E_C + 3] = 0u;
                        E[E_DC + 0] = __float_as_uint(ug); E[E_DC + 1] = __float_as_uint(vg);
                        E[E_DC + 2] = __float_as_uint(wg); E[E_DC + 3] = 0u;
                    }
                    else
                    {
                        float omt = fsub(1.0f, tt);
#pragma unroll
                        for(int i = 0; i < 4; ++i)
                        {
                            float c0 = fadd(fmul(omt, lit[mn[e]][i]), fmul(tt, lit[mx[e]][i]));          // :4091
                            E[E_C + i] = __float_as_uint(c0);
                            E[E_DC + i] = __float_as_uint(fdiv_zq(fsub(lit[mx[e]][i], c0), ydiff));         // :4096
                        }
                    }
                    E[E_LEFT] = ((ymin == round_s32(prj[e].y)) ? 1u : 0u) |                           // :4093
                                ((unsigned)mn[e] << 8) | ((unsigned)mx[e] << 16);
                    if(ymax > max_row) max_row = ymax;
                    if(slot_of[e] == 0) first_row = ymin;                                            // projekt.cpp:173
                }
                max_y = (max_row > v.height) ? v.height : max_row;                                   // :187-196
            }
        }
        if(out.recs)                                        // b200r_fill_edge_table only
        {
            uint32_t *g = out.recs + (size_t)(m.prim_base + base + (unsigned)tri)*kRecWords;
            g[R_NEDGES] = (uint32_t)nedges; g[R_FIRSTROW] = (uint32_t)first_row; g[R_MAXY] = (uint32_t)max_y;
            g[R_EDGE0 + 3*kEdgeWords] = emit;
            for(int w = 0; w < nedges*kEdgeWords; ++w) g[R_EDGE0 + w] = rec[R_EDGE0 + w];
        }

        have_walk = (nedges >= 2) && out.spans != nullptr;
        nonfinite = 0;
        if(have_walk && !TEX)
        {
            // RoundR32ToU32 (cvtss2si) and cvt.rni.s32.f32 agree only for |c*255| < 2^31.  Colours of
            // finite scenes stay near [0,1]; a triangle whose edge colours could leave that range
            // (NaN/Inf or absurd input) takes the guarded pack in the raster kernel.  Edges that
            // never step (YMax <= YMin: inserted and expired in the same row) are never drawn.
#pragma unroll
            for(int e = 0; e < 3; ++e)
            {
                if(e >= nedges) break;
                const uint32_t *E = rec + R_EDGE0 + e*kEdgeWords;
                if((int)E[E_YMAX] <= (int)E[E_YMIN]) continue;
#pragma unroll
                for(int i = 0; i < 4; ++i)
                    if(!(fabsf(__uint_as_float(E[E_C + i])) <= 4.0f) || !(fabsf(__uint_as_float(E[E_DC + i])) <= 4.0f))
                        nonfinite = 1;
            }
        }
        // Row range this GPU's band needs: rows above the band are walked (state), rows below not.
        // (with contiguous rows, the row just above the band can still drop an alias pixel into it)
        walk_end = min(max_y, v.band_y1);
        if(have_walk && !(first_row < walk_end && max_y > v.band_y0 - (v.alias_rows ? 1 : 0))) have_walk = false;

        // Number of spans and segments, without walking: between consecutive list-change rows the
        // set of active edges {e : YMin <= row < YMax} is constant; a stretch with >= 2 of them
        // yields one span per row.  A segment is the run of a triangle's spans inside one tile-row
        // band (its span records are consecutive), so the segment count is the number of distinct
        // bands those rows fall into.  The walk below opens segments by exactly the same rule.
        if(have_walk) count_rows(rec, nedges, first_row, walk_end, my_segs, my_spans);
    }

    // ---- CTA-wide exclusive scans of segment and span counts: warp shuffles, then ONE pair of
    //      global atomicAdds per CTA hands out the output ranges ----
    {
        const unsigned lane = t & 31, warp = t >> 5;
        unsigned incl_g = (unsigned)my_segs, incl_p = (unsigned)my_spans;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1)
        {
            unsigned ug = __shfl_up_sync(0xffffffffu, incl_g, d);
            unsigned up = __shfl_up_sync(0xffffffffu, incl_p, d);
            if(lane >= (unsigned)d) { incl_g += ug; incl_p += up; }
        }
        if(lane == 31) { s_warp_sum[warp] = incl_g; s_warp_sum2[warp] = incl_p; }
        __syncthreads();
        if(t == 0)
        {
            unsigned run_g = 0, run_p = 0;
            for(int w = 0; w < kSetupThreads/32; ++w)
            {
                unsigned cg = s_warp_sum[w], cp = s_warp_sum2[w];
                s_warp_sum[w] = run_g; s_warp_sum2[w] = run_p; run_g += cg; run_p += cp;
            }
            // region-relative bases; fits == false marks the whole CTA as overflowed (host re-issues)
            const unsigned region = blockIdx.x % kSubAllocators;
            const unsigned seg_region = out.seg_capacity/kSubAllocators, span_region = out.span_capacity/kSubAllocators;
            const unsigned g0 = run_g ? atomicAdd(out.seg_fill + region, run_g) : 0u;
            const unsigned p0 = run_p ? atomicAdd(out.span_fill + region, run_p) : 0u;
            s_fits = ((unsigned long long)g0 + run_g <= seg_region && (unsigned long long)p0 + run_p <= span_region) ? 1u : 0u;
            s_seg_base = region*seg_region + g0;
            s_span_base = region*span_region + p0;
        }
        __syncthreads();
        seg_at = s_seg_base + s_warp_sum[warp] + (incl_g - (unsigned)my_segs);
        span_at = s_span_base + s_warp_sum2[warp] + (incl_p - (unsigned)my_spans);
    }

    // ---- hand the walks out again: compact the triangles that have rows to walk and sort them by
    //      row count (counting sort in shared memory), so that warps are dense and hold triangles of
    //      similar height.  Lanes of a warp walk in lock step, so a warp costs its tallest triangle;
    //      without this, culled / off-band / short triangles leave most lanes idle.
    //      SPLIT: the units are walkers (triangle, slab of part_rows rows), handed out in batches of one
    //      per thread. ----
    const bool mine_walks = tri >= 0 && have_walk && s_fits != 0u;
    s_w_first[t] = first_row; s_w_end[t] = walk_end;
    s_w_info[t] = (unsigned)nedges | ((unsigned)nonfinite << 8) | (mine_walks ? 0x8000u : 0u) | ((unsigned)my_segs << 16);
    s_w_spans[t] = (unsigned)my_spans;
    s_w_seg_at[t] = seg_at; s_w_span_at[t] = span_at; s_w_tri[t] = tri;
    // SPLIT: slab s covers the band-relative rows [s*part_rows, (s+1)*part_rows); rows above the band are
    // replayed by every walker and emitted by none
    constexpr int kSplitMinRows = 40;                     // shorter triangles stay one walker
    if(SPLIT) s_alive[t] = 0u;                             // reused: "this triangle has been counted as binned"
    const int part_rows = SPLIT ? m.part_rows : 1;
    int my_parts = mine_walks ? 1 : 0, my_slab0 = 0;
    unsigned my_woff = 0;
    if(SPLIT)
    {
        if(mine_walks && walk_end - first_row > kSplitMinRows)
        {
            my_slab0 = (max(first_row, v.band_y0) - v.band_y0)/part_rows;
            my_parts = (walk_end - 1 - v.band_y0)/part_rows - my_slab0 + 1;
        }
        // exclusive scan of the walker counts over the CTA
        const unsigned lane = t & 31, warp = t >> 5;
        unsigned incl = (unsigned)my_parts;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1)
        {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
            if(lane >= (unsigned)d) incl += up;
        }
        __syncthreads();                                    // s_warp_sum was read by the allocation scan
        if(lane == 31) s_warp_sum[warp] = incl;
        __syncthreads();
        unsigned before = 0, all = 0;
        for(unsigned w = 0; w < kSetupThreads/32; ++w) { if(w < warp) before += s_warp_sum[w]; all += s_warp_sum[w]; }
        my_woff = before + incl - (unsigned)my_parts;
        if(t == 0) s_wtotal = all;
    }
    for(unsigned wb = 0; ; wb += kSetupThreads)
    {
    unsigned wmine = (unsigned)t;                          // this thread's unit before sorting: triangle slot | part << 8
    bool unit = mine_walks;
    if(SPLIT)
    {
        __syncthreads();                                    // s_wtotal written / the previous batch has been read
        if(wb >= s_wtotal) break;
        s_wmap[t] = 0xffffffffu;
        __syncthreads();
        for(int q = 0; q < my_parts; ++q)
        {
            const unsigned w = my_woff + (unsigned)q;
            if(w >= wb && w < wb + kSetupThreads) s_wmap[w - wb] = (unsigned)t | ((unsigned)q << 8);
        }
        __syncthreads();
        wmine = s_wmap[t];
        unit = wmine != 0xffffffffu;
    }
    else if(wb > 0) break;                                // one unit per thread: a single batch
    // a unit's rows: where it starts to walk, where it starts and stops to emit
    auto unit_rows = [&](unsigned wm, int &u_first, int &u_emit, int &u_end)
    {
        const int slot_ = (int)(wm & 0xffu), part_ = (int)(wm >> 8);
        u_first = s_w_first[slot_]; u_emit = v.band_y0; u_end = s_w_end[slot_];
        if(SPLIT && u_end - u_first > kSplitMinRows)
        {
            const int slab = (max(u_first, v.band_y0) - v.band_y0)/part_rows + part_;
            u_emit = max(v.band_y0 + slab*part_rows, v.band_y0);
            u_end = min(u_end, v.band_y0 + (slab + 1)*part_rows);
        }
    };
    {
        int u_first = 0, u_emit = 0, u_end = 0;
        if(unit) unit_rows(wmine, u_first, u_emit, u_end);
        // SPLIT: lanes walk in lock step, so a row costs the warp a span set-up as soon as ONE lane emits it;
        // walkers are therefore grouped by how many rows they replay first (their emitting rows are at most
        // part_rows each): a warp replays together, then emits together
        // (not SPLIT: by height; MeshParams::sort_shift scales heights into the kSortBins bins for frames of
        // tall triangles, which would otherwise all share the last bin)
        const int my_rows = unit ? (SPLIT ? ((u_emit - u_first) >> 2) : ((u_end - u_first) >> m.sort_shift)) : 0;
        const int key = unit ? (kSortBins - 1 - max(min(my_rows, kSortBins - 1), 0)) : kSortBins;   // tall first, idle last
        if(t <= kSortBins) s_hist[t] = 0;
        __syncthreads();
        const unsigned rank = atomicAdd(&s_hist[key], 1u);
        __syncthreads();
        if(t < 32)
        {
            // exclusive scan of the kSortBins + 1 bin counts by one warp (three bins per lane)
            unsigned a = s_hist[3*t], b = (3*t + 1 <= kSortBins) ? s_hist[3*t + 1] : 0u, c3 = (3*t + 2 <= kSortBins) ? s_hist[3*t + 2] : 0u;
            unsigned sum = a + b + c3, incl = sum;
#pragma unroll
            for(int d = 1; d < 32; d <<= 1)
            {
                unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
                if(t >= d) incl += up;
            }
            unsigned run = incl - sum;
            s_hist[3*t] = run; run += a;
            if(3*t + 1 <= kSortBins) { s_hist[3*t + 1] = run; run += b; }
            if(3*t + 2 <= kSortBins) { s_hist[3*t + 2] = run; }
        }
        __syncthreads();
        s_order[s_hist[key] + rank] = unit ? wmine : 0xffffffffu;          // order inside a bin is irrelevant
        __syncthreads();
    }

    // ---- the row walk.  All 32 lanes of a warp stay in the loops below (warp-wide trip counts,
    //      predicated bodies): list events are handled for all lanes that have one at the same
    //      time, then the rows up to each lane's next event run in lock step. ----
    {
        const unsigned wm = s_order[t];                     // the unit this thread walks
        const bool have_unit = wm != 0xffffffffu;
        const int slot = have_unit ? (int)(wm & 0xffu) : 0; // whose phase-2 results
        const int tri = max(s_w_tri[slot], 0);              // that triangle's index inside the CTA
        const unsigned info = s_w_info[slot];
        const int nedges = (int)(info & 0xffu);
        const int nonfinite = (int)((info >> 8) & 0x7fu);
        int my_segs = (int)(info >> 16), my_spans = (int)s_w_spans[slot];
        int first_row = 0, emit_from = 0, walk_end = 0;
        if(have_unit) unit_rows(wm, first_row, emit_from, walk_end);
        unsigned seg_at = s_w_seg_at[slot], span_at = s_w_span_at[slot];
        const bool walking = have_unit && (info & 0x8000u) != 0;
        const uint32_t *rec = s_edge + tri*kEdgeRec - R_EDGE0;
        if(SPLIT && walking && s_w_end[slot] - s_w_first[slot] > kSplitMinRows)
        {
            // this walker's share of the triangle's allocation: what the rows before its slab produce, and
            // what the rows up to its end produce
            int sb, pb, se, pe;
            count_rows(rec, nedges, first_row, emit_from, sb, pb);
            count_rows(rec, nedges, first_row, walk_end, se, pe);
            seg_at += (unsigned)sb; span_at += (unsigned)pb;
            my_segs = se - sb; my_spans = pe - pb;
        }
        const float *nrm = PHONG ? (s_nrm + tri*9) : nullptr;
        const int sw = out.span_words;
        const uint32_t span_flags = (nonfinite ? kSpanNonFinite : 0u) | (PHONG ? kSpanPhong : 0u) |
                                    (TEX ? (kSpanTex | ((uint32_t)m.tex << 8)) : 0u);
        const float wf = (float)v.width, wf_m1 = fsub(wf, 1.0f);
        ActiveEdge L, R;
        L.x = L.z = L.c0 = L.c1 = L.c2 = L.c3 = L.dx = L.dz = L.d0 = L.d1 = L.d2 = L.d3 = 0.0f;
        L.n0 = L.n1 = L.n2 = L.g0 = L.g1 = L.g2 = 0.0f;
        L.ymax = 0; L.id = -1; R = L;
        int nact = 0, next_ev = first_row;
        int y = first_row;
        int seg_band = -1;                                  // band of the open segment, -1: none
        unsigned seg = seg_at, span = span_at;
        int seg_minx = 0x7fffffff, seg_maxx = (int)0x80000000;
        unsigned seg_span0 = 0;
        unsigned pairs = 0;
        const unsigned prim = m.prim_base + (listed ? s_id[tri] : base + (unsigned)tri);
        // depth bucket of the whole triangle: 0 = nearest (largest camera z wins, projekt.cpp:525)
        unsigned bucket = 0;
        if(walking)
        {
            // the centroid's depth orders the per-pixel depths of overlapping triangles better than the nearest
            // vertex does (measured on C3, whose triangles are steeply slanted: raster 0.68 -> 0.61 ms)
            const float ztri = (s_pos[tri*9 + 2] + s_pos[tri*9 + 5] + s_pos[tri*9 + 8])*(1.0f/3.0f) + m.pz;
            const float f = (out.zrange[0] - ztri)*out.zrange[1]*(float)kDepthBuckets;
            if(f > 0.0f) bucket = (unsigned)min((int)f, kDepthBuckets - 1);
        }

        auto close_segment = [&]()
        {
            if(seg_band < 0) return;
            int tx0 = 1, tx1 = 0;
            if(seg_minx <= seg_maxx) { tx0 = seg_minx >> v.tile_w_shift; tx1 = seg_maxx >> v.tile_w_shift; }
            SegInfo si;
            si.tile_row = (unsigned)seg_band | (bucket << 24);
            si.tx = (unsigned)tx0 | ((unsigned)tx1 << 16);
            si.span_base = seg_span0;
            si.nrows = span - seg_span0;
            B200R_ASSERT(seg < seg_at + (unsigned)my_segs && seg < out.seg_capacity);
            out.segs[seg] = si;
            for(int tx = tx0; tx <= tx1; ++tx)
                atomicAdd(&out.tile_count[(seg_band*v.tiles_x + tx)*kDepthBuckets + bucket], si.nrows);
            if(tx0 <= tx1) pairs += (unsigned)(tx1 - tx0 + 1)*si.nrows;
            ++seg; seg_band = -1;
        };

        while(__any_sync(0xffffffffu, walking && y < walk_end))
        {
            // SPLIT: a row costs the whole warp a span set-up as soon as ONE lane emits it, so the lanes of a
            // warp are kept in phase: while any lane still replays rows in front of its slab (stepping only),
            // the lanes that have reached theirs wait; then all emit together
            bool go = walking && y < walk_end;
            int stop = walk_end;
            if(SPLIT)
            {
                const bool replaying = go && y < emit_from;
                if(__any_sync(0xffffffffu, replaying)) { go = replaying; stop = min(emit_from, walk_end); }
            }
            // (a) list events (projekt.cpp:202-296), together
            if(go && y == next_ev) active_list_event(y, rec, nedges, L, R, nact, next_ev, nrm);
            // (b) rows up to the next event
            // (MeshParams::row_chunk caps the stretch for frames of tall triangles: with stretches of ~50 rows a
            // lane whose stretch ends early would idle until the warp's longest one has finished; with the
            // cap it goes through its list event a few rows later and carries on)
            const int nrow = go ? min(min(next_ev, stop) - y, m.row_chunk) : 0;
            const int nrow_max = __reduce_max_sync(0xffffffffu, nrow);
            for(int k = 0; k < nrow_max; ++k)
            {
                if(k >= nrow) continue;
                if(nact == 2)
                {
                    const bool in_band = y >= emit_from;           // rows before: state only (above the band / another walker's)
                    if(in_band || (v.alias_rows && y == v.band_y0 - 1 && emit_from == v.band_y0))
                    {
                        if(in_band)
                        {
                            const int band = (y - v.band_y0) >> v.tile_h_shift;
                            if(band != seg_band)
                            {
                                close_segment();
                                seg_band = band; seg_span0 = span; seg_minx = 0x7fffffff; seg_maxx = (int)0x80000000;
                            }
                        }
                        // ---- span set-up, projekt.cpp:306-412, once per row ----
                        const float xdiff = roundf(fsub(R.x, L.x));                   // :311-312
                        float zi = 0.0f, i0 = 0.0f, i1 = 0.0f, i2 = 0.0f, i3 = 0.0f;
                        float ni0 = 0.0f, ni1 = 0.0f, ni2 = 0.0f;
                        if(xdiff != 0.0f)                                             // :333-363
                        {
                            i0 = fdiv_zn(fsub(R.c0, L.c0), xdiff); i1 = fdiv_zn(fsub(R.c1, L.c1), xdiff);
                            i2 = fdiv_zn(fsub(R.c2, L.c2), xdiff); i3 = fdiv_zn(fsub(R.c3, L.c3), xdiff);
                            zi = fdiv_zn(fsub(R.z, L.z), xdiff);
                            if(PHONG)                                                 // :344-349
                            {
                                ni0 = fdiv_zn(fsub(R.n0, L.n0), xdiff); ni1 = fdiv_zn(fsub(R.n1, L.n1), xdiff);
                                ni2 = fdiv_zn(fsub(R.n2, L.n2), xdiff);
                            }
                        }
                        float xoff = 0.0f, leftx = L.x;                               // :381-390
                        if(leftx < 0.0f) { xoff = -leftx; leftx = 0.0f; }
                        else if(leftx >= wf) { leftx = wf_m1; }
                        float rightx = R.x;                                           // :392-400
                        if(rightx < 0.0f) { rightx = 0.0f; }
                        else if(rightx >= wf) { rightx = wf_m1; }
                        const int minx = round_s32(leftx);                            // :402-406
                        int maxx = round_s32(rightx);
                        if(v.right_end_exclusive) maxx -= 1;                          // AVX fillers: [MinX, MaxX), projekt.cpp:782-794
                        const float z = fadd(L.z, fmul(xoff, zi));                    // :375, :408
                        const float c0 = fadd(L.c0, fmul(xoff, i0)), c1 = fadd(L.c1, fmul(xoff, i1));   // :379, :412
                        const float c2 = fadd(L.c2, fmul(xoff, i2)), c3 = fadd(L.c3, fmul(xoff, i3));
                        float sn0 = 0.0f, sn1 = 0.0f, sn2 = 0.0f;                     // :378, :411
                        if(PHONG) { sn0 = fadd(L.n0, fmul(xoff, ni0)); sn1 = fadd(L.n1, fmul(xoff, ni1)); sn2 = fadd(L.n2, fmul(xoff, ni2)); }
                        if(maxx >= v.width && minx <= maxx)
                        {
                            // An end in [Width-0.5, Width) is not clamped (:387, :397) and rounds up to
                            // column == Width (:402-403); the reference's pointer arithmetic (:414-419)
                            // puts that pixel into column 0 of the NEXT row when rows are contiguous, into
                            // row padding otherwise, past the buffer on the last row.  The next-row write
                            // is reproduced as a one-pixel span of its own; the others are dropped.
                            const int ay = y + 1;
                            if(v.alias_rows && ay < v.height && ay >= v.band_y0 && ay < v.band_y1)
                            {
                                float az = z, a0 = c0, a1 = c1, a2 = c2, a3 = c3;
                                float an0 = sn0, an1 = sn1, an2 = sn2;
                                for(int sx = minx; sx < v.width; ++sx)                // :534-535 / :504-506 up to that column
                                {
                                    if(PHONG) { an0 = fadd(an0, ni0); an1 = fadd(an1, ni1); an2 = fadd(an2, ni2); normalize3f(an0, an1, an2); }
                                    a0 = fadd(a0, i0); a1 = fadd(a1, i1); a2 = fadd(a2, i2); a3 = fadd(a3, i3);
                                    az = fadd(az, zi);
                                }
                                const unsigned ex = atomicAdd(out.extra_total, 1u);
                                if(ex < out.span_capacity && ex < out.seg_capacity)
                                {
                                    const unsigned asp = out.span_capacity - 1u - ex, asg = out.seg_capacity - 1u - ex;
                                    float4 *Q = reinterpret_cast<float4 *>(out.spans + (size_t)asp*sw);
                                    // the aliased pixel is shaded with its own X = Width, Row = y (:455-457);
                                    // the span record says column 0 of row ay, so a Phong alias span carries the
                                    // unprojection coordinates in its (unused) increment fields
                                    Q[0] = make_float4(__uint_as_float(prim), __int_as_float(ay), __int_as_float(0), __int_as_float(0));
                                    Q[1] = make_float4(az, a0, a1, a2);
                                    Q[2] = make_float4(a3, 0.0f, 0.0f, 0.0f);
                                    Q[3] = make_float4(0.0f, 0.0f, __uint_as_float(span_flags | (PHONG ? kSpanAlias : 0u)), az);
                                    if(PHONG) { Q[4] = make_float4(an0, an1, an2, (float)v.width); Q[5] = make_float4((float)y, 0.0f, 0.0f, 0.0f); }
                                    SegInfo si;
                                    const unsigned trow = (unsigned)((ay - v.band_y0) >> v.tile_h_shift);
                                    si.tile_row = trow | (bucket << 24); si.tx = 0u; si.span_base = asp; si.nrows = 1u;
                                    out.segs[asg] = si;
                                    atomicAdd(&out.tile_count[(trow*v.tiles_x)*kDepthBuckets + bucket], 1u);
                                    pairs += 1u;
                                }
                            }
                            maxx = v.width - 1;
                        }
                        if(in_band)
                        {
                            B200R_ASSERT(span < span_at + (unsigned)my_spans && span < out.span_capacity);
                            float4 *Q = reinterpret_cast<float4 *>(out.spans + (size_t)span*sw);
                            Q[0] = make_float4(__uint_as_float(prim), __int_as_float(y), __int_as_float(minx), __int_as_float(maxx));
                            Q[1] = make_float4(z, c0, c1, c2);
                            Q[2] = make_float4(c3, zi, i0, i1);
                            Q[3] = make_float4(i2, i3, __uint_as_float(span_flags), span_depth_bound(z, zi, maxx - minx));
                            if(PHONG) { Q[4] = make_float4(sn0, sn1, sn2, ni0); Q[5] = make_float4(ni1, ni2, 0.0f, 0.0f); }
                            ++span;
                            if(minx <= maxx) { seg_minx = min(seg_minx, minx); seg_maxx = max(seg_maxx, maxx); }
                        }
                    }
                    step_edge<PHONG>(L); step_edge<PHONG>(R);                     // :542-552
                    if(L.x > R.x)                                                 // :562-572 (rare: keep it a branch)
                    {
                        asm volatile("" ::: "memory");
                        ActiveEdge tmp = L; L = R; R = tmp;
                    }
                }
                ++y;
            }
        }
        if(walking)
        {
            close_segment();
            // slots promised by the counts but not produced (cannot happen for finite input) are blanked
            for(; span < span_at + (unsigned)my_spans; ++span)
            {
                float4 *Q = reinterpret_cast<float4 *>(out.spans + (size_t)span*sw);
                Q[0] = make_float4(__uint_as_float(prim), 0.0f, __int_as_float(1), __int_as_float(0));
                Q[1] = Q[2] = Q[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            for(; seg < seg_at + (unsigned)my_segs; ++seg)
            {
                SegInfo si; si.tile_row = 0; si.tx = 1u; si.span_base = 0; si.nrows = 0;
                out.segs[seg] = si;
            }
            if(pairs)
            {
                // SPLIT: a triangle counts as binned once, whichever of its walkers produced pairs
                if(!SPLIT || atomicExch(&s_alive[slot], 1u) == 0u) atomicAdd(&s_binned, 1u);
                atomicAdd(&s_pairs, pairs);
            }
        }
    }
    }   // batches of walkers
    __syncthreads();

    if(t == 0 && s_binned)
    {
        atomicAdd(&out.counters[0], (unsigned long long)s_binned);
        atomicAdd(&out.counters[1], (unsigned long long)s_pairs);
    }
}

void launch_zrange(const MeshParams &m, unsigned *zkeys, cudaStream_t s)
{
    if(m.ntri == 0) return;
    unsigned blocks = (unsigned)std::min<size_t>(((size_t)m.ntri*9/16 + 255)/256 + 1, 148*8);
    zrange_kernel<<<blocks, 256, 0, s>>>(m, zkeys);
}

void launch_zrange_finish(unsigned *zkeys, cudaStream_t s)
{
    zrange_finish_kernel<<<1, 1, 0, s>>>(zkeys);
}

void launch_select(const ViewParams &v, const MeshParams &m, unsigned *list, unsigned *count, unsigned *zkeys, int sm_count, cudaStream_t s)
{
    if(m.ntri == 0) return;
    // persistent: 8 CTAs of 256 threads per SM, fewer for small meshes
    const unsigned nblocks = (m.ntri + 255u)/256u;
    select_kernel<<<std::min(nblocks, (unsigned)std::max(sm_count, 1)*8u), 256, 0, s>>>(v, m, list, count, zkeys);
}

void launch_setup(const ViewParams &v, const MeshParams &m, const SetupOutputs &out, cudaStream_t s)
{
    if(m.ntri == 0) return;
    const bool tex = m.uv != nullptr, listed = m.tri_list != nullptr, split = m.tris_per_cta > 0;
    const unsigned tpc = split ? (unsigned)m.tris_per_cta : (unsigned)kSetupThreads;
    const unsigned blocks = (m.ntri + tpc - 1)/tpc;
#define B200R_LAUNCH_SETUP(P, T, L) do { if(split) setup_kernel<P, T, L, true><<<blocks, kSetupThreads, 0, s>>>(v, m, out); \
                                         else setup_kernel<P, T, L, false><<<blocks, kSetupThreads, 0, s>>>(v, m, out); } while(0)
    if(listed)
    {
        if(m.phong) { if(tex) B200R_LAUNCH_SETUP(true, true, true); else B200R_LAUNCH_SETUP(true, false, true); }
        else        { if(tex) B200R_LAUNCH_SETUP(false, true, true); else B200R_LAUNCH_SETUP(false, false, true); }
    }
    else
    {
        if(m.phong) { if(tex) B200R_LAUNCH_SETUP(true, true, false); else B200R_LAUNCH_SETUP(true, false, false); }
        else        { if(tex) B200R_LAUNCH_SETUP(false, true, false); else B200R_LAUNCH_SETUP(false, false, false); }
    }
#undef B200R_LAUNCH_SETUP
}

// ---------------------------------------------------------------- clear
__global__ void clear_kernel(uint32_t *color, int cpw, float *depth, int ds, int width, int rows,
                             uint32_t cval, float dval)
{
    int x = blockIdx.x*blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if(x < width && y < rows)
    {
        color[(size_t)y*cpw + x] = cval;
        depth[(size_t)y*ds + x] = dval;
    }
}

void launch_clear(uint32_t *color, int color_pitch_words, float *depth, int depth_stride,
                  int width, int rows, uint32_t cval, float dval, cudaStream_t s)
{
    dim3 grid((width + 255)/256, rows);
    clear_kernel<<<grid, 256, 0, s>>>(color, color_pitch_words, depth, depth_stride, width, rows, cval, dval);
}

} // namespace b200r
