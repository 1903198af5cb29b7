// Whole-object mode: DrawModel's intrusive active-edge
// list over ALL edges of an object (projekt.cpp:198-303, 542-597), replayed link by link.
//
// The reference pairs consecutive list entries whatever triangle they belong to (:300-303,
// :584-592) and its exchanges (:562-583) relink nodes without updating ListHead / ListTail; the
// images it produces for multi-triangle objects depend on both (0.1-1.2 % of the demo sphere's
// pixels differ from per-triangle semantics).  Reproducing them needs the list itself, and the list
// is order dependent from the first row to the last: ONE thread walks one object, on a device copy
// of the object's sorted edge_info array (projekt.h:17-37) that it mutates in place exactly as
// DrawModel does, with edge_info::Next holding an index (-1 = null).  For every pair it performs the
// span set-up (:306-412) and emits a span record plus a one-row segment whose owner is the span's
// DRAW ORDER (the depth rule's tie-break: within an object, first drawn wins).  Pixels are filled by
// the same binning and raster kernels as the per-triangle path.
//
// Where the reference dereferences a null pointer (the list ran empty :262, :300; a stale ListTail
// :222, :275) or would follow a cycle, the object STOPS drawing (the stop
// points are pinned against the verbatim build's crash points by the CPU checker).
//
// This is a compatibility mode: objects run in parallel, rows of one object do not.
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

struct Walker
{
    const ViewParams &v;
    const ObjectWalkParams &p;
    const ObjectDesc &o;
    DevEdge *E;
    unsigned produced;          // spans emitted so far (draw order)
    unsigned pairs;             // tile pairs
    float wf, wf_m1;

    __device__ bool before(int a, int b) const      // projekt.cpp:212-216 / 229-233
    {
        const DevEdge &A = E[a], &B = E[b];
        return A.XMin < B.XMin || (A.XMin == B.XMin && (A.Gradient < B.Gradient || (A.Gradient == B.Gradient && A.Left < B.Left)));
    }

    // span set-up of the pair (L, R) at row y, projekt.cpp:306-412, and its records
    __device__ void emit(const DevEdge &L, const DevEdge &R, int y)
    {
        const bool tex = o.tex >= 0, phong = o.phong != 0;
        // textured objects interpolate u/z, v/z, 1/z in the colour words (MeshParams::uv)
        const float Lc0 = tex ? L.UMin : L.MinColor[0], Lc1 = tex ? L.VMin : L.MinColor[1];
        const float Lc2 = tex ? L.OneOverZMin : L.MinColor[2], Lc3 = tex ? 0.0f : L.MinColor[3];
        const float Rc0 = tex ? R.UMin : R.MinColor[0], Rc1 = tex ? R.VMin : R.MinColor[1];
        const float Rc2 = tex ? R.OneOverZMin : R.MinColor[2], Rc3 = tex ? 0.0f : R.MinColor[3];
        const float xdiff = roundf(fsub(R.XMin, L.XMin));                     // :311-312
        float zi = 0.0f, i0 = 0.0f, i1 = 0.0f, i2 = 0.0f, i3 = 0.0f, ni0 = 0.0f, ni1 = 0.0f, ni2 = 0.0f;
        if(xdiff != 0.0f)                                                     // :333-363
        {
            i0 = fdiv_zn(fsub(Rc0, Lc0), xdiff); i1 = fdiv_zn(fsub(Rc1, Lc1), xdiff);
            i2 = fdiv_zn(fsub(Rc2, Lc2), xdiff); i3 = fdiv_zn(fsub(Rc3, Lc3), xdiff);
            zi = fdiv_zn(fsub(R.ZMin, L.ZMin), xdiff);
            if(phong)
            {
                ni0 = fdiv_zn(fsub(R.MinNormal[0], L.MinNormal[0]), xdiff);
                ni1 = fdiv_zn(fsub(R.MinNormal[1], L.MinNormal[1]), xdiff);
                ni2 = fdiv_zn(fsub(R.MinNormal[2], L.MinNormal[2]), xdiff);
            }
        }
        float xoff = 0.0f, leftx = L.XMin;                                    // :381-390
        if(leftx < 0.0f) { xoff = -leftx; leftx = 0.0f; }
        else if(leftx >= wf) { leftx = wf_m1; }
        float rightx = R.XMin;                                                // :392-400
        if(rightx < 0.0f) { rightx = 0.0f; }
        else if(rightx >= wf) { rightx = wf_m1; }
        const int minx = round_s32(leftx);                                    // :402-406
        int maxx = round_s32(rightx);
        const float z = fadd(L.ZMin, fmul(xoff, zi));                         // :375, :408
        const float c0 = fadd(Lc0, fmul(xoff, i0)), c1 = fadd(Lc1, fmul(xoff, i1));
        const float c2 = fadd(Lc2, fmul(xoff, i2)), c3 = fadd(Lc3, fmul(xoff, i3));
        float sn0 = 0.0f, sn1 = 0.0f, sn2 = 0.0f;
        if(phong)
        {
            sn0 = fadd(L.MinNormal[0], fmul(xoff, ni0)); sn1 = fadd(L.MinNormal[1], fmul(xoff, ni1));
            sn2 = fadd(L.MinNormal[2], fmul(xoff, ni2));
        }
        // colours of a whole object are not range-checked by a set-up pass: always the guarded pack
        const uint32_t flags = kSpanNonFinite | (phong ? kSpanPhong : 0u) | (tex ? (kSpanTex | ((uint32_t)o.tex << 8)) : 0u);
        const unsigned prim = o.prim_base + produced;
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
        const unsigned slot = striped_slot(o.span_base + produced, p.region_size);
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
        ++produced;
    }

    __device__ void step(DevEdge &e) const                                   // projekt.cpp:542-560
    {
        e.XMin = fadd(e.XMin, e.Gradient);
        e.ZMin = fadd(e.ZMin, e.ZGradient);
#pragma unroll
        for(int i = 0; i < 4; ++i) e.MinColor[i] = fadd(e.MinColor[i], e.ColorGradient[i]);
        if(o.phong)
        {
            float n0 = fadd(e.MinNormal[0], e.NormalGradient[0]), n1 = fadd(e.MinNormal[1], e.NormalGradient[1]);
            float n2 = fadd(e.MinNormal[2], e.NormalGradient[2]);
            normalize3f(n0, n1, n2);
            e.MinNormal[0] = n0; e.MinNormal[1] = n1; e.MinNormal[2] = n2;
        }
        if(o.tex >= 0)
        {
            e.UMin = fadd(e.UMin, e.UGradient); e.VMin = fadd(e.VMin, e.VGradient);
            e.OneOverZMin = fadd(e.OneOverZMin, e.OneOverZGradient);
        }
    }

    // returns false where the reference dereferences null / would cycle
    __device__ bool walk()
    {
        const int n = (int)o.edge_count;
        if(n == 0) return true;
        const int first_row = E[0].YMin;                                     // :173
        int max_row = E[0].YMax;                                             // :176-185
        for(int e = 1; e < n; ++e) max_row = max(max_row, E[e].YMax);
        const int max_y = min(max_row, v.height);                            // :187-196
        int head = -1, tail = -1;
        const long long fuse = 4ll*n + 64;
        for(int row = first_row; row < max_y; ++row)                         // :198
        {
            for(int cur = 0; cur < n; ++cur)                                 // :202-260
            {
                if(E[cur].YMin != row) continue;
                if(head >= 0)
                {
                    if(before(cur, head)) { E[cur].Next = head; head = cur; }
                    else
                    {
                        int compared = head, previous = head;
                        long long steps = 0;
                        while(compared != tail)
                        {
                            compared = (int)E[compared].Next;
                            if(compared < 0 || ++steps > fuse) return false;
                            if(before(cur, compared)) { E[cur].Next = compared; E[previous].Next = cur; compared = tail; }
                            else previous = compared;
                        }
                        if(previous == compared) { E[tail].Next = cur; tail = cur; }
                    }
                }
                else { head = cur; tail = head; }
            }
            for(long long steps = 0;; ++steps)                                // :262-267
            {
                if(head < 0 || steps > fuse) return false;
                if(!(E[head].YMax <= row)) break;
                const int removed = head; head = (int)E[head].Next; E[removed].Next = -1;
            }
            {
                int previous = head, checked = head;                         // :269-296
                long long steps = 0;
                while(checked != tail)
                {
                    checked = (int)E[checked].Next;
                    if(checked < 0 || ++steps > fuse) return false;
                    if(E[checked].YMax <= row)
                    {
                        if(checked == tail) { tail = previous; E[tail].Next = -1; checked = tail; }
                        else { E[previous].Next = E[checked].Next; checked = previous; }
                    }
                    previous = checked;
                }
            }
            int prev_cur = -1, prev_next = -1;                                // :298-303
            int cur = head, nxt = (int)E[cur].Next;
            long long npairs = 0;
            while(nxt >= 0)
            {
                if(++npairs > fuse) return false;
                if(produced >= o.span_bound) return false;                   // cannot happen: the bound counts edge rows
                emit(E[cur], E[nxt], row);                                   // :306-540
                step(E[cur]); step(E[nxt]);                                  // :542-560
                if(E[cur].XMin > E[nxt].XMin)                                // :562-572
                {
                    E[cur].Next = E[nxt].Next;
                    E[nxt].Next = cur;
                    if(prev_next >= 0) E[prev_next].Next = nxt;
                    cur = nxt;
                    nxt = (int)E[cur].Next;
                }
                if(prev_next >= 0 && E[prev_next].XMin > E[cur].XMin)        // :574-584
                {
                    E[prev_next].Next = E[cur].Next;
                    E[cur].Next = prev_next;
                    E[prev_cur].Next = cur;
                    prev_next = cur;
                    cur = (int)E[prev_next].Next;
                    if(cur < 0) return false;
                }
                prev_cur = cur; prev_next = nxt;                              // :586-587
                if(nxt < 0) return false;
                if(E[nxt].Next >= 0) { cur = (int)E[nxt].Next; nxt = (int)E[cur].Next; }   // :589-597
                else nxt = -1;
            }
        }
        return true;
    }
};

__global__ void __launch_bounds__(32)
object_walk_kernel(ViewParams v, ObjectWalkParams p)
{
    const unsigned oi = blockIdx.x*blockDim.x + threadIdx.x;
    if(oi >= p.nobjects) return;
    const ObjectDesc o = p.objects[oi];
    Walker w = { v, p, o, reinterpret_cast<DevEdge *>(p.edges) + o.first_edge, 0u, 0u, (float)v.width, fsub((float)v.width, 1.0f) };
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

void launch_object_walk(const ViewParams &v, const ObjectWalkParams &p, cudaStream_t s)
{
    if(p.nobjects == 0) return;
    object_walk_kernel<<<(p.nobjects + 31)/32, 32, 0, s>>>(v, p);
}

} // namespace b200r
