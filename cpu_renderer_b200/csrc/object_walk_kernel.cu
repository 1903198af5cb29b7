// Whole-object mode: DrawModel's intrusive active-edge
// list over ALL edges of an object (projekt.cpp:198-303, 542-597), replayed link by link.
//
// The reference pairs consecutive list entries whatever triangle they belong to (:300-303,
// :584-592) and its exchanges (:562-583) relink nodes without updating ListHead / ListTail; the
// images it produces for multi-triangle objects depend on both (0.1-1.2 % of the demo sphere's
// pixels differ from per-triangle semantics).  Reproducing them needs the list itself, and the list
// is order dependent from the first row to the last: ONE thread walks one object, on the object's
// sorted edge_info array (projekt.h:17-37); the fields DrawModel mutates in place (running values
// and Next, an index here with -1 = null) are kept in a structure of arrays beside it.  For every pair it performs the
// span set-up (:306-412) and emits a span record plus a one-row segment whose owner is the span's
// DRAW ORDER (the depth rule's tie-break: within an object, first drawn wins).  Pixels are filled by
// the same binning and raster kernels as the per-triangle path.
//
// Where the reference dereferences a null pointer (the list ran empty :262, :300; a stale ListTail
// :222, :275) or would follow a cycle, the object STOPS drawing (the stop
// points are pinned against the verbatim build's crash points by the CPU checker).
//
// Two implementations of the same replay:
//   * three phases (default).  What is order dependent in DrawModel is the LIST; the values are not:
//     an edge's value after k steps is the k-fold sequential add (:542-560) from its start value,
//     whichever rows those steps happened in.  So (A) chain_kernel, a thread per edge, computes every
//     edge's chain of values; (B) order_kernel, one lane per object, replays only the list --
//     comparisons and relinking -- with each edge's step COUNT as index into its chain, and emits
//     pairs (chain index left, chain index right, row); (C) emit_kernel, a thread per pair, does the
//     span set-up.  A skipped or repeated step (the reference's odd trailing entry, its exchanges)
//     is just a count that lags or leads; a count that would run past the chain flags the object for
//   * the serial walk (object_walk_kernel): one lane does everything.  10x slower; the fallback.
// This is a compatibility mode either way: objects run in parallel, the list of one object does not.
#include "raster_device.cuh"
#include "edge_walk.cuh"

namespace b200r {

// byte-for-byte edge_info (projekt.h:17-37, include/b200_raster.h); Next is an index here
struct DevEdge
{
    int YMax; float XMin, ZMin, OneOverZMin, Gradient, ZGradient, OneOverZGradient;
    int YMin; float UMin, VMin, UGradient, VGradient;
    int Left; float MinColor[4], ColorGradient[4], MinNormal[3], NormalGradient[3];
    long long Next;
};
static_assert(sizeof(DevEdge) == 120, "edge_info is 120 bytes");

// slot of the g-th span of the frame: striped over the sub-allocator regions so that they fill evenly
__device__ __forceinline__ unsigned striped_slot(unsigned g, unsigned region_size)
{
    return (g % kSubAllocators)*region_size + g/kSubAllocators;
}

// values of one edge at one row
struct EdgeVals { float x, z, c0, c1, c2, c3, n0, n1, n2; };

// Span set-up of the pair (L, R) at row y, projekt.cpp:306-412, and its records.  `index` is the
// pair's position in the object's draw order: its promised slot and, plus prim_base, its owner.
// Returns the number of (span, tile) pairs it produced.
__device__ unsigned emit_span(const ViewParams &v, const ObjectWalkParams &p, const ObjectDesc &o,
                              unsigned index, int y, const EdgeVals &L, const EdgeVals &R)
{
    const bool tex = o.tex >= 0, phong = o.phong != 0;
    const float wf = (float)v.width, wf_m1 = fsub(wf, 1.0f);
    unsigned pairs = 0;
    const float xdiff = roundf(fsub(R.x, L.x));                           // :311-312
    float zi = 0.0f, i0 = 0.0f, i1 = 0.0f, i2 = 0.0f, i3 = 0.0f, ni0 = 0.0f, ni1 = 0.0f, ni2 = 0.0f;
    if(xdiff != 0.0f)                                                     // :333-363
    {
        i0 = fdiv_zn(fsub(R.c0, L.c0), xdiff); i1 = fdiv_zn(fsub(R.c1, L.c1), xdiff);
        i2 = fdiv_zn(fsub(R.c2, L.c2), xdiff); i3 = fdiv_zn(fsub(R.c3, L.c3), xdiff);
        zi = fdiv_zn(fsub(R.z, L.z), xdiff);
        if(phong)
        {
            ni0 = fdiv_zn(fsub(R.n0, L.n0), xdiff); ni1 = fdiv_zn(fsub(R.n1, L.n1), xdiff);
            ni2 = fdiv_zn(fsub(R.n2, L.n2), xdiff);
        }
    }
    float xoff = 0.0f, leftx = L.x;                                       // :381-390
    if(leftx < 0.0f) { xoff = -leftx; leftx = 0.0f; }
    else if(leftx >= wf) { leftx = wf_m1; }
    float rightx = R.x;                                                   // :392-400
    if(rightx < 0.0f) { rightx = 0.0f; }
    else if(rightx >= wf) { rightx = wf_m1; }
    const int minx = round_s32(leftx);                                    // :402-406
    int maxx = round_s32(rightx);
    const float z = fadd(L.z, fmul(xoff, zi));                            // :375, :408
    const float c0 = fadd(L.c0, fmul(xoff, i0)), c1 = fadd(L.c1, fmul(xoff, i1));
    const float c2 = fadd(L.c2, fmul(xoff, i2)), c3 = fadd(L.c3, fmul(xoff, i3));
    float sn0 = 0.0f, sn1 = 0.0f, sn2 = 0.0f;
    if(phong) { sn0 = fadd(L.n0, fmul(xoff, ni0)); sn1 = fadd(L.n1, fmul(xoff, ni1)); sn2 = fadd(L.n2, fmul(xoff, ni2)); }
    // colours of a whole object are not range-checked by a set-up pass: always the guarded pack
    const uint32_t flags = kSpanNonFinite | (phong ? kSpanPhong : 0u) | (tex ? (kSpanTex | ((uint32_t)o.tex << 8)) : 0u);
    const unsigned prim = o.prim_base + index;
    const int sw = p.span_words;
    const bool in_band = y >= v.band_y0 && y < v.band_y1;
    if(maxx >= v.width && minx <= maxx)
    {
        // column == Width: the reference's write lands in column 0 of the next row of a contiguous
        // target (setup_kernel.cu has the full story); reproduced as a one-pixel span of its own
        const int ay = y + 1;
        if(v.alias_rows && ay < v.height && ay >= v.band_y0 && ay < v.band_y1)
        {
            float az = z, a0 = c0, a1 = c1, a2 = c2, a3 = c3, an0 = sn0, an1 = sn1, an2 = sn2;
            for(int sx = minx; sx < v.width; ++sx)
            {
                if(phong) { an0 = fadd(an0, ni0); an1 = fadd(an1, ni1); an2 = fadd(an2, ni2); normalize3f(an0, an1, an2); }
                a0 = fadd(a0, i0); a1 = fadd(a1, i1); a2 = fadd(a2, i2); a3 = fadd(a3, i3);
                az = fadd(az, zi);
            }
            const unsigned ex = atomicAdd(p.extra_total, 1u);
            if(ex < p.span_capacity && ex < p.seg_capacity)
            {
                const unsigned asp = p.span_capacity - 1u - ex, asg = p.seg_capacity - 1u - ex;
                float4 *Q = reinterpret_cast<float4 *>(p.spans + (size_t)asp*sw);
                Q[0] = make_float4(__uint_as_float(prim), __int_as_float(ay), __int_as_float(0), __int_as_float(0));
                Q[1] = make_float4(az, a0, a1, a2);
                Q[2] = make_float4(a3, 0.0f, 0.0f, 0.0f);
                Q[3] = make_float4(0.0f, 0.0f, __uint_as_float(flags | (phong ? kSpanAlias : 0u)), az);
                if(sw > kSpanWords) { Q[4] = make_float4(an0, an1, an2, (float)v.width); Q[5] = make_float4((float)y, 0.0f, 0.0f, 0.0f); }
                SegInfo si;
                const unsigned trow = (unsigned)((ay - v.band_y0) >> v.tile_h_shift);
                si.tile_row = trow; si.tx = 0u; si.span_base = asp; si.nrows = 1u;
                p.segs[asg] = si;
                atomicAdd(&p.tile_count[(trow*v.tiles_x)*kDepthBuckets], 1u);
                pairs += 1u;
            }
        }
        maxx = v.width - 1;
    }
    // every pair consumes one promised slot, drawn or not, so that owners stay in draw order
    const unsigned slot = striped_slot(o.span_base + index, p.region_size);
    SegInfo si; si.tile_row = 0; si.tx = 1u; si.span_base = slot; si.nrows = 0;     // tx0 > tx1: touches no tile
    if(in_band)
    {
        float4 *Q = reinterpret_cast<float4 *>(p.spans + (size_t)slot*sw);
        Q[0] = make_float4(__uint_as_float(prim), __int_as_float(y), __int_as_float(minx), __int_as_float(maxx));
        Q[1] = make_float4(z, c0, c1, c2);
        Q[2] = make_float4(c3, zi, i0, i1);
        Q[3] = make_float4(i2, i3, __uint_as_float(flags), span_depth_bound(z, zi, maxx - minx));
        if(sw > kSpanWords) { Q[4] = make_float4(sn0, sn1, sn2, ni0); Q[5] = make_float4(ni1, ni2, 0.0f, 0.0f); }
        if(minx <= maxx)
        {
            const int tx0 = minx >> v.tile_w_shift, tx1 = maxx >> v.tile_w_shift;
            const unsigned trow = (unsigned)((y - v.band_y0) >> v.tile_h_shift);
            si.tile_row = trow; si.tx = (unsigned)tx0 | ((unsigned)tx1 << 16); si.nrows = 1u;
            for(int tx = tx0; tx <= tx1; ++tx) atomicAdd(&p.tile_count[(trow*v.tiles_x + tx)*kDepthBuckets], 1u);
            pairs += (unsigned)(tx1 - tx0 + 1);
        }
    }
    p.segs[slot] = si;
    return pairs;
}

// One edge step, projekt.cpp:542-560, on explicit values (the colour words of a textured object
// carry u/z, v/z, 1/z; its vertex colours never reach the image).
__device__ __forceinline__ void step_vals(EdgeVals &e, const DevEdge &d, bool phong, bool tex)
{
    e.x = fadd(e.x, __ldg(&d.Gradient));
    e.z = fadd(e.z, __ldg(&d.ZGradient));
    if(tex)
    {
        e.c0 = fadd(e.c0, __ldg(&d.UGradient)); e.c1 = fadd(e.c1, __ldg(&d.VGradient));
        e.c2 = fadd(e.c2, __ldg(&d.OneOverZGradient));
    }
    else
    {
        e.c0 = fadd(e.c0, __ldg(&d.ColorGradient[0])); e.c1 = fadd(e.c1, __ldg(&d.ColorGradient[1]));
        e.c2 = fadd(e.c2, __ldg(&d.ColorGradient[2])); e.c3 = fadd(e.c3, __ldg(&d.ColorGradient[3]));
    }
    if(phong)
    {
        e.n0 = fadd(e.n0, __ldg(&d.NormalGradient[0])); e.n1 = fadd(e.n1, __ldg(&d.NormalGradient[1]));
        e.n2 = fadd(e.n2, __ldg(&d.NormalGradient[2]));
        normalize3f(e.n0, e.n1, e.n2);
    }
}

__device__ __forceinline__ EdgeVals start_vals(const DevEdge &d, bool phong, bool tex)
{
    EdgeVals e;
    e.x = d.XMin; e.z = d.ZMin;
    e.c0 = tex ? d.UMin : d.MinColor[0]; e.c1 = tex ? d.VMin : d.MinColor[1];
    e.c2 = tex ? d.OneOverZMin : d.MinColor[2]; e.c3 = tex ? 0.0f : d.MinColor[3];
    e.n0 = phong ? d.MinNormal[0] : 0.0f; e.n1 = phong ? d.MinNormal[1] : 0.0f; e.n2 = phong ? d.MinNormal[2] : 0.0f;
    return e;
}

// Mutable per-edge state of the walk, structure of arrays: what DrawModel keeps updating inside the
// edge records (XMin, ZMin, MinColor or -- textured -- u/z v/z 1/z, MinNormal, Next).  It lives in
// shared memory when the object fits (28 / 40 bytes per edge), else in a global scratch area with
// the same layout; everything DrawModel only reads (YMin, YMax, Left, the gradients) stays in the
// edge_info array behind the read-only cache.  (First version: all of it in global memory, every
// value re-read from L2 one row after it was written -- 3x slower.)
struct WalkState
{
    float *x, *z, *c;           // c: 4 per edge
    float *n;                   // 3 per edge, Phong only
    int *next;
};

struct Walker
{
    const ViewParams &v;
    const ObjectWalkParams &p;
    const ObjectDesc &o;
    const DevEdge *E;           // read-only fields
    WalkState s;
    unsigned produced;          // spans emitted so far (draw order)
    unsigned pairs;             // tile pairs
    float wf, wf_m1;

    __device__ bool before(int a, int b) const      // projekt.cpp:212-216 / 229-233
    {
        const float ax = s.x[a], bx = s.x[b];
        if(ax < bx) return true;
        if(ax != bx) return false;
        const float ag = __ldg(&E[a].Gradient), bg = __ldg(&E[b].Gradient);
        return ag < bg || (ag == bg && __ldg(&E[a].Left) < __ldg(&E[b].Left));
    }

    __device__ EdgeVals vals(int i) const
    {
        EdgeVals e;
        e.x = s.x[i]; e.z = s.z[i];
        e.c0 = s.c[4*i]; e.c1 = s.c[4*i + 1]; e.c2 = s.c[4*i + 2]; e.c3 = s.c[4*i + 3];
        e.n0 = e.n1 = e.n2 = 0.0f;
        if(o.phong) { e.n0 = s.n[3*i]; e.n1 = s.n[3*i + 1]; e.n2 = s.n[3*i + 2]; }
        return e;
    }

    __device__ void emit(int l, int r, int y)
    {
        pairs += emit_span(v, p, o, produced, y, vals(l), vals(r));
        ++produced;
    }

    __device__ void step(int i)                                              // projekt.cpp:542-560
    {
        const DevEdge &e = E[i];
        s.x[i] = fadd(s.x[i], __ldg(&e.Gradient));
        s.z[i] = fadd(s.z[i], __ldg(&e.ZGradient));
        if(o.tex >= 0)                                                       // :554-560 (the colours of a textured
        {                                                                    //  object never reach the image)
            s.c[4*i + 0] = fadd(s.c[4*i + 0], __ldg(&e.UGradient));
            s.c[4*i + 1] = fadd(s.c[4*i + 1], __ldg(&e.VGradient));
            s.c[4*i + 2] = fadd(s.c[4*i + 2], __ldg(&e.OneOverZGradient));
        }
        else
        {
#pragma unroll
            for(int k = 0; k < 4; ++k) s.c[4*i + k] = fadd(s.c[4*i + k], __ldg(&e.ColorGradient[k]));   // :548-549
        }
        if(o.phong)                                                          // :551-552
        {
            float n0 = fadd(s.n[3*i], __ldg(&e.NormalGradient[0])), n1 = fadd(s.n[3*i + 1], __ldg(&e.NormalGradient[1]));
            float n2 = fadd(s.n[3*i + 2], __ldg(&e.NormalGradient[2]));
            normalize3f(n0, n1, n2);
            s.n[3*i] = n0; s.n[3*i + 1] = n1; s.n[3*i + 2] = n2;
        }
    }

    // returns false where the reference dereferences null / would cycle
    __device__ bool walk()
    {
        const int n = (int)o.edge_count;
        if(n == 0) return true;
        int *next = s.next;
        const int first_row = __ldg(&E[0].YMin);                             // :173
        int max_row = __ldg(&E[0].YMax);                                     // :176-185
        for(int e = 1; e < n; ++e) max_row = max(max_row, __ldg(&E[e].YMax));
        const int max_y = min(max_row, v.height);                            // :187-196
        int head = -1, tail = -1;
        const long long fuse = 4ll*n + 64;
        // The reference scans every edge on every row for YMin == RowIndex (:202-208).  The array is
        // sorted by YMin (MergeSort, :4117) and rows advance by one from Edges[0].YMin, so those edges
        // are the contiguous run at a cursor, visited in the same (array) order.
        int cursor = 0;
        for(int row = first_row; row < max_y; ++row)                         // :198
        {
            for(; cursor < n && __ldg(&E[cursor].YMin) <= row; ++cursor)     // :202-260
            {
                const int cur = cursor;
                if(__ldg(&E[cur].YMin) != row) continue;
                if(head >= 0)
                {
                    if(before(cur, head)) { next[cur] = head; head = cur; }
                    else
                    {
                        int compared = head, previous = head;
                        long long steps = 0;
                        while(compared != tail)
                        {
                            compared = next[compared];
                            if(compared < 0 || ++steps > fuse) return false;
                            if(before(cur, compared)) { next[cur] = compared; next[previous] = cur; compared = tail; }
                            else previous = compared;
                        }
                        if(previous == compared) { next[tail] = cur; tail = cur; }
                    }
                }
                else { head = cur; tail = head; }
            }
            for(long long steps = 0;; ++steps)                                // :262-267
            {
                if(head < 0 || steps > fuse) return false;
                if(!(__ldg(&E[head].YMax) <= row)) break;
                const int removed = head; head = next[head]; next[removed] = -1;
            }
            {
                int previous = head, checked = head;                         // :269-296
                long long steps = 0;
                while(checked != tail)
                {
                    checked = next[checked];
                    if(checked < 0 || ++steps > fuse) return false;
                    if(__ldg(&E[checked].YMax) <= row)
                    {
                        if(checked == tail) { tail = previous; next[tail] = -1; checked = tail; }
                        else { next[previous] = next[checked]; checked = previous; }
                    }
                    previous = checked;
                }
            }
            int prev_cur = -1, prev_next = -1;                                // :298-303
            int cur = head, nxt = next[cur];
            long long npairs = 0;
            while(nxt >= 0)
            {
                if(++npairs > fuse) return false;
                if(produced >= o.span_bound) return false;                   // cannot happen: the bound counts edge rows
                emit(cur, nxt, row);                                         // :306-540
                step(cur); step(nxt);                                        // :542-560
                if(s.x[cur] > s.x[nxt])                                      // :562-572
                {
                    next[cur] = next[nxt];
                    next[nxt] = cur;
                    if(prev_next >= 0) next[prev_next] = nxt;
                    cur = nxt;
                    nxt = next[cur];
                }
                if(prev_next >= 0 && s.x[prev_next] > s.x[cur])              // :574-584
                {
                    next[prev_next] = next[cur];
                    next[cur] = prev_next;
                    next[prev_cur] = cur;
                    prev_next = cur;
                    cur = next[prev_next];
                    if(cur < 0) return false;
                }
                prev_cur = cur; prev_next = nxt;                              // :586-587
                if(nxt < 0) return false;
                if(next[nxt] >= 0) { cur = next[nxt]; nxt = next[cur]; }     // :589-597
                else nxt = -1;
            }
        }
        return true;
    }
};

constexpr int kObjectThreads = 32;

// One CTA per object: all lanes load the mutable state, lane 0 walks.
__global__ void __launch_bounds__(kObjectThreads)
object_walk_kernel(ViewParams v, ObjectWalkParams p)
{
    extern __shared__ __align__(16) float s_state[];
    const unsigned oi = blockIdx.x;
    if(oi >= p.nobjects) return;
    if(p.chain_base != nullptr && p.fallback[oi] == 0u) return;     // the three-phase path handled it
    const ObjectDesc o = p.objects[oi];
    const DevEdge *E = reinterpret_cast<const DevEdge *>(p.edges) + o.first_edge;
    const unsigned n = o.edge_count;
    const unsigned words = o.phong ? 10u : 7u;              // x, z, 4 interpolants, next (+ 3 normal)
    // the object's state: shared memory if it fits, else its slice of the global scratch area
    float *base = ((size_t)n*words*sizeof(float) <= p.state_smem_bytes) ? s_state
                                                                         : p.state_scratch + (size_t)o.first_edge*10u;
    WalkState st;
    st.x = base; st.z = base + n; st.c = base + 2*n; st.next = reinterpret_cast<int *>(base + 6*n); st.n = base + 7*n;
    const bool tex = o.tex >= 0;
    for(unsigned e = threadIdx.x; e < n; e += kObjectThreads)
    {
        const DevEdge &d = E[e];
        st.x[e] = d.XMin; st.z[e] = d.ZMin;
        st.c[4*e + 0] = tex ? d.UMin : d.MinColor[0];
        st.c[4*e + 1] = tex ? d.VMin : d.MinColor[1];
        st.c[4*e + 2] = tex ? d.OneOverZMin : d.MinColor[2];
        st.c[4*e + 3] = tex ? 0.0f : d.MinColor[3];
        st.next[e] = -1;                                    // FillEdgeTable: Next = 0 (:4094)
        if(o.phong) { st.n[3*e] = d.MinNormal[0]; st.n[3*e + 1] = d.MinNormal[1]; st.n[3*e + 2] = d.MinNormal[2]; }
    }
    __syncthreads();
    if(threadIdx.x != 0) return;
    Walker w = { v, p, o, E, st, 0u, 0u, (float)v.width, fsub((float)v.width, 1.0f) };
    const bool ok = w.walk();
    if(!ok) atomicAdd(p.stopped, 1u);
    // promised slots the object did not use
    for(unsigned g = w.produced; g < o.span_bound; ++g)
    {
        const unsigned slot = striped_slot(o.span_base + g, p.region_size);
        SegInfo si; si.tile_row = 0; si.tx = 1u; si.span_base = slot; si.nrows = 0;
        p.segs[slot] = si;
    }
    if(w.pairs)
    {
        atomicAdd(&p.counters[0], 1ull);
        atomicAdd(&p.counters[1], (unsigned long long)w.pairs);
    }
}

// ---------------------------------------------------------------- three-phase path
// chain entry k of an edge = its values after k steps; rows + 1 entries (the value after the last
// step takes part in that row's exchange test, :562)
__device__ __forceinline__ unsigned chain_entries(const DevEdge &d, int max_y)
{
    const int rows = min(__ldg(&d.YMax), max_y) - __ldg(&d.YMin);
    return (unsigned)max(rows, 0) + 1u;
}

__device__ __forceinline__ void store_vals(const ObjectWalkParams &p, unsigned at, const EdgeVals &e, bool phong)
{
    const size_t T = p.chain_T;
    p.chains[at] = e.x; p.chains[T + at] = e.z;
    p.chains[2*T + 4*(size_t)at + 0] = e.c0; p.chains[2*T + 4*(size_t)at + 1] = e.c1;
    p.chains[2*T + 4*(size_t)at + 2] = e.c2; p.chains[2*T + 4*(size_t)at + 3] = e.c3;
    if(phong) { p.chains[6*T + 3*(size_t)at + 0] = e.n0; p.chains[6*T + 3*(size_t)at + 1] = e.n1; p.chains[6*T + 3*(size_t)at + 2] = e.n2; }
}

__device__ __forceinline__ EdgeVals load_vals(const ObjectWalkParams &p, unsigned at, bool phong)
{
    const size_t T = p.chain_T;
    EdgeVals e;
    e.x = p.chains[at]; e.z = p.chains[T + at];
    const float4 c = *reinterpret_cast<const float4 *>(p.chains + 2*T + 4*(size_t)at);
    e.c0 = c.x; e.c1 = c.y; e.c2 = c.z; e.c3 = c.w;
    e.n0 = e.n1 = e.n2 = 0.0f;
    if(phong) { e.n0 = p.chains[6*T + 3*(size_t)at + 0]; e.n1 = p.chains[6*T + 3*(size_t)at + 1]; e.n2 = p.chains[6*T + 3*(size_t)at + 2]; }
    return e;
}

// (A) a thread per edge: its chain of values.  grid.y = object.
__global__ void __launch_bounds__(128)
chain_kernel(ObjectWalkParams p)
{
    const ObjectDesc o = p.objects[blockIdx.y];
    const unsigned e = blockIdx.x*blockDim.x + threadIdx.x;
    if(e >= o.edge_count) return;
    const DevEdge &d = reinterpret_cast<const DevEdge *>(p.edges)[o.first_edge + e];
    const bool phong = o.phong != 0, tex = o.tex >= 0;
    const unsigned n = chain_entries(d, o.max_y);
    unsigned at = p.chain_base[o.first_edge + e];
    EdgeVals v = start_vals(d, phong, tex);
    store_vals(p, at, v, phong);
    for(unsigned k = 1; k < n; ++k)
    {
        step_vals(v, d, phong, tex);
        store_vals(p, at + k, v, phong);
    }
}

// (B) one lane per object: DrawModel's list with chain positions instead of values.  Per edge, in
// shared memory: next, the current chain position (relative to the object's first chain entry), the
// last valid position, YMax.  (First version: position recomputed from a global base + a count and
// the chain length from two more global loads, 64-bit loop guards: 480 instructions per pair.)
struct OrderWalker
{
    const ObjectWalkParams &p;
    const ObjectDesc &o;
    const DevEdge *E;
    const float *xs;            // x chains, indexed by position: shared-memory copy or the global array + chain_first
    int *next;
    unsigned *pos;              // current chain entry of each edge
    const unsigned *end;        // its last entry
    const int *ymax;
    unsigned produced;

    __device__ float x(int i) const { return xs[pos[i]]; }
    __device__ bool before(int a, int b) const      // projekt.cpp:212-216 / 229-233
    {
        const float ax = x(a), bx = x(b);
        if(ax < bx) return true;
        if(ax != bx) return false;
        const float ag = __ldg(&E[a].Gradient), bg = __ldg(&E[b].Gradient);
        return ag < bg || (ag == bg && __ldg(&E[a].Left) < __ldg(&E[b].Left));
    }
    // 0 = done, 1 = stopped where the reference dereferences null, 2 = a step ran past a chain
    __device__ int walk()
    {
        const int n = (int)o.edge_count;
        if(n == 0) return 0;
        const int first_row = __ldg(&E[0].YMin);                             // :173
        const int max_y = o.max_y;                                           // :176-196
        int head = -1, tail = -1;
        const unsigned fuse = 4u*(unsigned)n + 64u;
        uint4 *out = p.pair_list + o.span_base;
        const unsigned chain_first = o.chain_first;
        int cursor = 0;                                                      // see Walker::walk
        int next_ymin = first_row;
        for(int row = first_row; row < max_y; ++row)                         // :198
        {
            while(next_ymin <= row && cursor < n)                            // :202-260
            {
                const int cur = cursor++;
                const int ymin = next_ymin;
                next_ymin = (cursor < n) ? __ldg(&E[cursor].YMin) : 0x7fffffff;
                if(ymin != row) continue;
                if(head >= 0)
                {
                    if(before(cur, head)) { next[cur] = head; head = cur; }
                    else
                    {
                        // The scan is 58 % of this kernel's instructions (48 hops per inserted edge on the
                        // demo sphere).  A hop's comparison (position -> x -> compare) does not feed the
                        // next hop's link load, so that load is issued first and the two chains overlap.
                        int compared = head, previous = head;
                        unsigned steps = 0;
                        const float cx = x(cur);
                        int ahead = next[head];
                        while(compared != tail)
                        {
                            compared = ahead;
                            if(compared < 0 || ++steps > fuse) return 1;
                            ahead = next[compared];
                            const float ox = xs[pos[compared]];
                            bool first = cx < ox;
                            if(cx == ox)
                            {
                                const float ag = __ldg(&E[cur].Gradient), bg = __ldg(&E[compared].Gradient);
                                first = ag < bg || (ag == bg && __ldg(&E[cur].Left) < __ldg(&E[compared].Left));
                            }
                            if(first) { next[cur] = compared; next[previous] = cur; compared = tail; }
                            else previous = compared;
                        }
                        if(previous == compared) { next[tail] = cur; tail = cur; }
                    }
                }
                else { head = cur; tail = head; }
            }
            for(unsigned steps = 0;; ++steps)                                 // :262-267
            {
                if(head < 0 || steps > fuse) return 1;
                if(!(ymax[head] <= row)) break;
                const int removed = head; head = next[head]; next[removed] = -1;
            }
            {
                int previous = head, checked = head;                         // :269-296
                unsigned steps = 0;
                int ahead = next[head];
                while(checked != tail)
                {
                    checked = ahead;
                    if(checked < 0 || ++steps > fuse) return 1;
                    ahead = next[checked];                                    // the successor, before any relinking below
                    if(ymax[checked] <= row)
                    {
                        if(checked == tail) { tail = previous; next[tail] = -1; checked = tail; }
                        else { next[previous] = ahead; checked = previous; }
                    }
                    previous = checked;
                }
            }
            int prev_cur = -1, prev_next = -1;                                // :298-303
            int cur = head, nxt = next[cur];
            unsigned npairs = 0;
            while(nxt >= 0)
            {
                if(++npairs > fuse) return 1;
                if(produced >= o.span_bound) return 2;
                // :306-540: the pair, as the chain entries of its two edges at this moment
                const unsigned pc = pos[cur], pn = pos[nxt];
                out[produced] = make_uint4(chain_first + pc, chain_first + pn, (unsigned)row, 0u);
                ++produced;
                // :542-560: one step each = the next chain entry
                if(pc >= end[cur] || pn >= end[nxt]) return 2;
                pos[cur] = pc + 1u; pos[nxt] = pn + 1u;
                float xc = xs[pc + 1u], xn = xs[pn + 1u];
                if(xc > xn)                                                  // :562-572
                {
                    next[cur] = next[nxt];
                    next[nxt] = cur;
                    if(prev_next >= 0) next[prev_next] = nxt;
                    cur = nxt;
                    nxt = next[cur];
                    xc = xn;
                }
                if(prev_next >= 0 && x(prev_next) > xc)                      // :574-584
                {
                    next[prev_next] = next[cur];
                    next[cur] = prev_next;
                    next[prev_cur] = cur;
                    prev_next = cur;
                    cur = next[prev_next];
                    if(cur < 0) return 1;
                }
                prev_cur = cur; prev_next = nxt;                              // :586-587
                if(nxt < 0) return 1;
                const int after = next[nxt];
                if(after >= 0) { cur = after; nxt = next[cur]; }             // :589-597
                else nxt = -1;
            }
        }
        return 0;
    }
};

__global__ void __launch_bounds__(kObjectThreads)
order_kernel(ObjectWalkParams p)
{
    extern __shared__ __align__(16) float s_state[];
    const unsigned oi = blockIdx.x;
    if(oi >= p.nobjects) return;
    const ObjectDesc o = p.objects[oi];
    const unsigned n = o.edge_count;
    const DevEdge *E = reinterpret_cast<const DevEdge *>(p.edges) + o.first_edge;
    // four words per edge always in shared memory; the x chains too when there is room
    const size_t need_links = (size_t)n*4*sizeof(int);
    const size_t need_all = need_links + (size_t)o.chain_total*sizeof(float);
    if(need_links > p.order_smem_bytes)                     // absurdly large object: serial walk
    {
        if(threadIdx.x == 0) { p.fallback[oi] = 1u; p.produced[oi] = 0u; }
        return;
    }
    int *next = reinterpret_cast<int *>(s_state);
    unsigned *pos = reinterpret_cast<unsigned *>(s_state) + n;
    unsigned *end = reinterpret_cast<unsigned *>(s_state) + 2*n;
    int *ymax = reinterpret_cast<int *>(s_state) + 3*n;
    float *xs_s = s_state + 4*n;
    const bool xs_shared = need_all <= p.order_smem_bytes;
    const unsigned *base = p.chain_base + o.first_edge;
    for(unsigned e = threadIdx.x; e < n; e += kObjectThreads)
    {
        next[e] = -1;
        const unsigned b = base[e] - o.chain_first;
        pos[e] = b; end[e] = b + chain_entries(E[e], o.max_y) - 1u;
        ymax[e] = E[e].YMax;
    }
    if(xs_shared)
        for(unsigned k = threadIdx.x; k < o.chain_total; k += kObjectThreads) xs_s[k] = p.chains[o.chain_first + k];
    __syncthreads();
    if(threadIdx.x != 0) return;
    OrderWalker w = { p, o, E, xs_shared ? xs_s : p.chains + o.chain_first, next, pos, end, ymax, 0u };
    const int rc = w.walk();
    if(rc == 2) { p.fallback[oi] = 1u; p.produced[oi] = 0u; return; }
    p.fallback[oi] = 0u;
    p.produced[oi] = w.produced;
    if(rc == 1) atomicAdd(p.stopped, 1u);
    if(w.produced) atomicAdd(&p.counters[0], 1ull);
}

// (C) a thread per promised slot: span set-up of a pair, or the blank record of an unused slot.
__global__ void __launch_bounds__(128)
emit_kernel(ViewParams v, ObjectWalkParams p)
{
    const unsigned oi = blockIdx.y;
    if(p.fallback[oi] != 0u) return;
    const ObjectDesc o = p.objects[oi];
    const unsigned i = blockIdx.x*blockDim.x + threadIdx.x;
    if(i >= o.span_bound) return;
    unsigned pairs = 0;
    if(i < p.produced[oi])
    {
        const uint4 pr = p.pair_list[o.span_base + i];
        const bool phong = o.phong != 0;
        pairs = emit_span(v, p, o, i, (int)pr.z, load_vals(p, pr.x, phong), load_vals(p, pr.y, phong));
    }
    else
    {
        const unsigned slot = striped_slot(o.span_base + i, p.region_size);
        SegInfo si; si.tile_row = 0; si.tx = 1u; si.span_base = slot; si.nrows = 0;
        p.segs[slot] = si;
    }
    // tile pairs of the warp in one atomic
    for(int d = 16; d > 0; d >>= 1) pairs += __shfl_down_sync(0xffffffffu, pairs, d);
    if((threadIdx.x & 31) == 0 && pairs) atomicAdd(&p.counters[1], (unsigned long long)pairs);
}

cudaError_t launch_object_walk(const ViewParams &v, const ObjectWalkParams &p, cudaStream_t s)
{
    if(p.nobjects == 0) return cudaSuccess;
    if(p.nobjects > 65535u) return cudaErrorInvalidConfiguration;      // objects sit on gridDim.y (api.cu rejects earlier)
    cudaError_t e;
    if(p.chain_base != nullptr)
    {
        // grid.x covers the largest object; smaller ones leave CTAs idle (objects are few)
        chain_kernel<<<dim3((p.max_edges + 127)/128, p.nobjects), 128, 0, s>>>(p);
        e = cudaFuncSetAttribute(order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.order_smem_bytes);
        if(e != cudaSuccess) return e;
        order_kernel<<<p.nobjects, kObjectThreads, p.order_smem_bytes, s>>>(p);
        emit_kernel<<<dim3((p.max_bound + 127)/128, p.nobjects), 128, 0, s>>>(v, p);
    }
    // the serial walk: every object without the three-phase path, only the flagged ones with it
    e = cudaFuncSetAttribute(object_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.state_smem_bytes);
    if(e != cudaSuccess) return e;
    object_walk_kernel<<<p.nobjects, kObjectThreads, p.state_smem_bytes, s>>>(v, p);
    return cudaGetLastError();
}

} // namespace b200r
