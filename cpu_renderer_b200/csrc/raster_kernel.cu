// Tile raster kernel (sm_100a): coverage, interpolation, depth test and Gouraud shading.
//
// Restates DrawModel's scalar Gouraud path for per-triangle objects:
//   active-edge insert / expire           projekt.cpp:202-296
//   span set-up                           projekt.cpp:306-412
//   pixel loop, depth test, ARGB pack     projekt.cpp:423-425, 510-538
//   edge step and crossing exchange       projekt.cpp:542-572
//
// "The reference's own arithmetic" is a chain of rounded binary32 additions: the value at pixel
// k of a span is k sequential adds from the span's left end, and an edge's value at row r is r
// sequential adds from its top.  There is no closed form, so the unit of
// parallel work is the triangle, not the pixel:
//
//   * a WARP owns one screen tile at a time, staged in its private slice of shared memory:
//     depth+owner as one 64-bit word per pixel, colour as one 32-bit word per pixel.  Tile rows
//     enter and leave with TMA bulk copies (cp.async.bulk, mbarrier completion);
//   * each LANE takes one triangle of the tile's bin and replays the reference's row walk and
//     span walk for it: rows before the tile and pixels left of the tile are replayed in
//     registers only (adds, no memory), pixels inside the tile are depth-tested in shared memory;
//   * lanes of a warp hold different triangles that may hit the same pixel, so the depth word is
//     updated with a 64-bit compare-and-swap carrying (z, submission index) and the rule
//         z > zold || (z == zold && prim < primold)
//     which is the reference's strict '>' with first-submitted-wins (projekt.cpp:525) made
//     order independent.  After a __syncwarp the lane that still owns the pixel writes colour.
//
// Warps never share a tile, so no block-level barrier is used after start-up.
#include "raster_device.cuh"

namespace b200r {

struct ActiveEdge
{
    float x, z, c0, c1, c2, c3;         // running XMin, ZMin, MinColor
    float dx, dz, d0, d1, d2, d3;       // Gradient, ZGradient, ColorGradient
    int ymax;
    int id;                             // slot in the triangle record
};

__device__ __forceinline__ void load_edge(ActiveEdge &a, const uint32_t *__restrict__ rec, int id)
{
    const uint32_t *E = rec + R_EDGE0 + id*kEdgeWords;
    a.ymax = (int)__ldg(E + E_YMAX);
    a.x = __uint_as_float(__ldg(E + E_X));    a.dx = __uint_as_float(__ldg(E + E_DX));
    a.z = __uint_as_float(__ldg(E + E_Z));    a.dz = __uint_as_float(__ldg(E + E_DZ));
    a.c0 = __uint_as_float(__ldg(E + E_C + 0)); a.c1 = __uint_as_float(__ldg(E + E_C + 1));
    a.c2 = __uint_as_float(__ldg(E + E_C + 2)); a.c3 = __uint_as_float(__ldg(E + E_C + 3));
    a.d0 = __uint_as_float(__ldg(E + E_DC + 0)); a.d1 = __uint_as_float(__ldg(E + E_DC + 1));
    a.d2 = __uint_as_float(__ldg(E + E_DC + 2)); a.d3 = __uint_as_float(__ldg(E + E_DC + 3));
    a.id = id;
}

// projekt.cpp:542-549: one row down each edge of the pair.
__device__ __forceinline__ void step_edge(ActiveEdge &a)
{
    a.x = fadd(a.x, a.dx);   a.z = fadd(a.z, a.dz);
    a.c0 = fadd(a.c0, a.d0); a.c1 = fadd(a.c1, a.d1);
    a.c2 = fadd(a.c2, a.d2); a.c3 = fadd(a.c3, a.d3);
}

// The active list of one triangle at row y (projekt.cpp:202-296 restated for <= 3 edges with
// array storage): insert, in record order, every edge whose YMin == y before the first entry
// it sorts strictly before (XMin, then Gradient, then Left; :212-216), then drop entries with
// YMax <= y.  L/R receive the first two survivors, keeping the running values of edges that
// were already active.  For finite vertices at most two edges survive a row (the upper and
// lower short edge of a triangle never overlap in rows); a third survivor is ignored.
__device__ __noinline__ void active_list_event(int y, const uint32_t *__restrict__ rec, int nedges,
                                               ActiveEdge &L, ActiveEdge &R, int &nact, int &next_ev)
{
    int ids[3] = {0, 0, 0};
    float xs[3] = {0.0f, 0.0f, 0.0f};
    int n = 0;
    if(nact >= 1) { ids[0] = L.id; xs[0] = L.x; n = 1; }
    if(nact >= 2) { ids[1] = R.id; xs[1] = R.x; n = 2; }
    for(int e = 0; e < nedges; ++e)
    {
        const uint32_t *E = rec + R_EDGE0 + e*kEdgeWords;
        if((int)__ldg(E + E_YMIN) != y) continue;
        float nx = __uint_as_float(__ldg(E + E_X)), ng = __uint_as_float(__ldg(E + E_DX));
        int nl = (int)__ldg(E + E_LEFT);
        int at = n;
        for(int k = n - 1; k >= 0; --k)
        {
            const uint32_t *O = rec + R_EDGE0 + ids[k]*kEdgeWords;
            float ox = xs[k], og = __uint_as_float(__ldg(O + E_DX));
            int ol = (int)__ldg(O + E_LEFT);
            if(nx < ox || (nx == ox && (ng < og || (ng == og && nl < ol)))) at = k;   // first such k
        }
        if(n < 3)
        {
            for(int k = n; k > at; --k) { ids[k] = ids[k - 1]; xs[k] = xs[k - 1]; }
            ids[at] = e; xs[at] = nx; ++n;
        }
    }
    int kept = 0, kid[3] = {0, 0, 0};
    for(int k = 0; k < n; ++k)
    {
        int ym = (int)__ldg(rec + R_EDGE0 + ids[k]*kEdgeWords + E_YMAX);
        if(ym <= y) continue;
        kid[kept++] = ids[k];
    }
    ActiveEdge oldL = L, oldR = R;
    if(kept >= 1)
    {
        if(nact >= 1 && kid[0] == oldL.id) L = oldL;
        else if(nact >= 2 && kid[0] == oldR.id) L = oldR;
        else load_edge(L, rec, kid[0]);
    }
    if(kept >= 2)
    {
        if(nact >= 1 && kid[1] == oldL.id) R = oldL;
        else if(nact >= 2 && kid[1] == oldR.id) R = oldR;
        else load_edge(R, rec, kid[1]);
    }
    nact = (kept > 2) ? 2 : kept;
    int ev = 0x7fffffff;
    for(int e = 0; e < nedges; ++e)
    {
        int ym = (int)__ldg(rec + R_EDGE0 + e*kEdgeWords + E_YMIN);
        if(ym > y && ym < ev) ev = ym;
    }
    if(nact >= 1 && L.ymax < ev) ev = L.ymax;
    if(nact >= 2 && R.ymax < ev) ev = R.ymax;
    next_ev = ev;
}

// RoundR32ToU32(c*255) per channel, A R G B from a r g b (projekt.cpp:520-523); no clamp.
__device__ __forceinline__ uint32_t pack_argb(float r, float g, float b, float a)
{
    return ((uint32_t)round_s32(fmul(a, 255.0f)) << 24) | ((uint32_t)round_s32(fmul(r, 255.0f)) << 16) |
           ((uint32_t)round_s32(fmul(g, 255.0f)) << 8) | ((uint32_t)round_s32(fmul(b, 255.0f)) << 0);
}

template<int TW, int TH>
struct TileLayout
{
    static constexpr int kPix = TW*TH;
    static constexpr int kBytesPerWarp = kPix*8 + kPix*4 + 16;     // depth+owner, colour, mbarrier
};

template<int TW, int TH, int WARPS>
__global__ void __launch_bounds__(WARPS*32)
raster_kernel(const RasterParams p)
{
    using Lay = TileLayout<TW, TH>;
    constexpr int NPIX = Lay::kPix;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    if(*p.pair_total > p.pair_capacity) return;            // host grows the list and re-issues

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *mine_smem = smem_raw + (size_t)warp*Lay::kBytesPerWarp;
    unsigned long long *zp = reinterpret_cast<unsigned long long *>(mine_smem);   // (prim << 32) | zbits
    uint32_t *col = reinterpret_cast<uint32_t *>(mine_smem + NPIX*8);
    uint64_t *bar = reinterpret_cast<uint64_t *>(mine_smem + NPIX*12);
    float *zstage = reinterpret_cast<float *>(zp + NPIX/2);        // upper half of the zp area

    if(lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    uint32_t phase = 0;

    const float wf = (float)p.v.width;
    const float wf_m1 = fsub(wf, 1.0f);
    const int band_rows = p.v.band_y1 - p.v.band_y0;

    while(true)
    {
        unsigned tile = 0;
        if(lane == 0) tile = atomicAdd(p.work_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if(tile >= p.ntiles) break;
        const unsigned cnt = p.tile_count[tile];
        if(cnt == 0) continue;
        const unsigned off = p.tile_offset[tile];
        const int tx = (int)(tile % (unsigned)p.v.tiles_x), ty = (int)(tile / (unsigned)p.v.tiles_x);
        const int x0 = tx*TW;
        const int yb = ty*TH;                              // band-relative first row
        const int ys0 = p.v.band_y0 + yb;                  // screen row of the tile's first row
        const int cols = min(TW, p.v.width - x0);
        const int rows = min(TH, band_rows - yb);

        // ---------------- stage the tile: depth -> zstage, colour -> col -----------------
        bulk_wait_read();                                  // this lane's earlier stores have read smem
        fence_proxy_async();
        __syncwarp();
        if(p.bulk_ok)
        {
            if(lane == 0) mbar_expect_tx(bar, (uint32_t)(rows*cols*8));
            __syncwarp();
            for(int r = lane; r < rows; r += 32)
            {
                bulk_g2s(zstage + r*TW, p.depth + (size_t)(yb + r)*p.depth_stride + x0, (uint32_t)(cols*4), bar);
                bulk_g2s(col + r*TW, p.color + (size_t)(yb + r)*p.color_pitch_words + x0, (uint32_t)(cols*4), bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        else
        {
            for(int r = 0; r < rows; ++r)
                for(int c = lane; c < cols; c += 32)
                {
                    zstage[r*TW + c] = p.depth[(size_t)(yb + r)*p.depth_stride + x0 + c];
                    col[r*TW + c] = p.color[(size_t)(yb + r)*p.color_pitch_words + x0 + c];
                }
            __syncwarp();
        }
        // expand depth to (z, owner = -1): owner -1 is "already in the target", wins every tie
        for(int c = 0; c < NPIX; c += 64)
        {
            float2 zz = *reinterpret_cast<const float2 *>(zstage + c + 2*lane);
            __syncwarp();
            ulonglong2 w;
            w.x = 0xffffffff00000000ull | (unsigned long long)__float_as_uint(zz.x);
            w.y = 0xffffffff00000000ull | (unsigned long long)__float_as_uint(zz.y);
            *reinterpret_cast<ulonglong2 *>(zp + c + 2*lane) = w;
        }
        __syncwarp();

        // ---------------- rasterise the bin, 32 triangles at a time ------------------------
        for(unsigned b = 0; b < cnt; b += 32)
        {
            const bool have = (b + lane) < cnt;
            const unsigned tri = have ? __ldg(p.pair_list + off + b + lane) : 0u;
            const uint32_t *rec = p.recs + (size_t)tri*kRecWords;
            int nedges = 0, first_row = 0, max_y = 0;
            if(have)
            {
                uint4 h = __ldg(reinterpret_cast<const uint4 *>(rec));
                nedges = (int)h.x; first_row = (int)h.y; max_y = (int)h.z;
            }
            const int prim = (int)tri;
            const int r0 = max(first_row, ys0), r1 = min(max_y, ys0 + rows);
            const bool valid = have && nedges >= 2 && r0 < r1;

            ActiveEdge L, R;
            L.x = L.z = L.c0 = L.c1 = L.c2 = L.c3 = L.dx = L.dz = L.d0 = L.d1 = L.d2 = L.d3 = 0.0f;
            L.ymax = 0; L.id = -1; R = L;
            int nact = 0, next_ev = first_row;

            // rows above the tile: replay the walk without touching memory
            if(valid)
            {
                for(int y = first_row; y < r0; ++y)
                {
                    if(y == next_ev) active_list_event(y, rec, nedges, L, R, nact, next_ev);
                    if(nact == 2)
                    {
                        step_edge(L); step_edge(R);
                        if(L.x > R.x) { ActiveEdge t = L; L = R; R = t; }      // :562-572
                    }
                }
            }

            const int nrows = valid ? (r1 - r0) : 0;
            const int nrows_max = __reduce_max_sync(0xffffffffu, nrows);
            for(int k = 0; k < nrows_max; ++k)
            {
                const bool rowact = k < nrows;
                const int y = r0 + k;
                int n = 0, xs = 0;
                float z = 0.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
                float zi = 0.0f, i0 = 0.0f, i1 = 0.0f, i2 = 0.0f, i3 = 0.0f;
                bool span = false;
                if(rowact)
                {
                    if(y == next_ev) active_list_event(y, rec, nedges, L, R, nact, next_ev);
                    span = (nact == 2);
                }
                if(span)
                {
                    // ---- span set-up, projekt.cpp:306-412 ----
                    float xdiff = roundf(fsub(R.x, L.x));                     // :311-312
                    if(xdiff != 0.0f)                                         // :333-363
                    {
                        i0 = fdiv(fsub(R.c0, L.c0), xdiff); i1 = fdiv(fsub(R.c1, L.c1), xdiff);
                        i2 = fdiv(fsub(R.c2, L.c2), xdiff); i3 = fdiv(fsub(R.c3, L.c3), xdiff);
                        zi = fdiv(fsub(R.z, L.z), xdiff);
                    }
                    z = L.z; c0 = L.c0; c1 = L.c1; c2 = L.c2; c3 = L.c3;     // :375-379
                    float xoff = 0.0f, leftx = L.x;                           // :381-390
                    if(leftx < 0.0f) { xoff = -leftx; leftx = 0.0f; }
                    else if(leftx >= wf) { leftx = wf_m1; }
                    float rightx = R.x;                                       // :392-400
                    if(rightx < 0.0f) { rightx = 0.0f; }
                    else if(rightx >= wf) { rightx = wf_m1; }
                    const int minx = round_s32(leftx), maxx = round_s32(rightx);   // :402-406
                    z = fadd(z, fmul(xoff, zi));                              // :408
                    c0 = fadd(c0, fmul(xoff, i0)); c1 = fadd(c1, fmul(xoff, i1));  // :412
                    c2 = fadd(c2, fmul(xoff, i2)); c3 = fadd(c3, fmul(xoff, i3));
                    // pixels left of the tile: the reference's per-pixel adds (:534-535), registers only
                    int skip = x0 - minx;
                    if(skip > maxx - minx + 1) skip = maxx - minx + 1;
                    for(int s = 0; s < skip; ++s)
                    {
                        c0 = fadd(c0, i0); c1 = fadd(c1, i1); c2 = fadd(c2, i2); c3 = fadd(c3, i3);
                        z = fadd(z, zi);
                    }
                    xs = max(minx, x0);
                    const int xe = min(maxx, x0 + cols - 1);
                    n = max(xe - xs + 1, 0);
                }

                // ---- pixel loop, projekt.cpp:423-425, 510-538 ----
                const int nmax = __reduce_max_sync(0xffffffffu, n);
                unsigned long long *zrow = zp + (y - ys0)*TW - x0;
                uint32_t *crow = col + (y - ys0)*TW - x0;
                for(int i = 0; i < nmax; ++i)
                {
                    bool won = false;
                    unsigned long long mine = 0;
                    const int x = xs + i;
                    if(i < n)
                    {
                        const float zo = reinterpret_cast<volatile float *>(zrow + x)[0];
                        if(z >= zo)
                        {
                            mine = ((unsigned long long)(unsigned)prim << 32) | (unsigned long long)__float_as_uint(z);
                            unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(zrow + x);
                            while(true)
                            {
                                const float oz = __uint_as_float((unsigned)old);
                                const int op = (int)(old >> 32);
                                if(!(z > oz || (z == oz && prim < op))) break;        // :525 + tie rule
                                const unsigned long long prev = atomicCAS(zrow + x, old, mine);
                                if(prev == old) { won = true; break; }
                                old = prev;
                            }
                        }
                    }
                    __syncwarp();
                    if(won)
                    {
                        if(*reinterpret_cast<volatile unsigned long long *>(zrow + x) == mine)
                            *reinterpret_cast<volatile uint32_t *>(crow + x) = pack_argb(c0, c1, c2, c3);
                    }
                    if(i < n)
                    {
                        c0 = fadd(c0, i0); c1 = fadd(c1, i1); c2 = fadd(c2, i2); c3 = fadd(c3, i3);   // :534
                        z = fadd(z, zi);                                                              // :535
                    }
                }

                if(span)
                {
                    step_edge(L); step_edge(R);                               // :542-549
                    if(L.x > R.x) { ActiveEdge t = L; L = R; R = t; }          // :562-572
                }
            }
            __syncwarp();
        }

        // ---------------- write the tile back ------------------------------------------------
        __syncwarp();
        float *zout = reinterpret_cast<float *>(zp);       // compact depth in place, front to back
        for(int c = 0; c < NPIX; c += 64)
        {
            ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(zp + c + 2*lane);
            __syncwarp();
            float2 zz = make_float2(__uint_as_float((unsigned)w.x), __uint_as_float((unsigned)w.y));
            *reinterpret_cast<float2 *>(zout + c + 2*lane) = zz;
        }
        __syncwarp();
        if(p.bulk_ok)
        {
            fence_proxy_async();
            __syncwarp();
            for(int r = lane; r < rows; r += 32)
            {
                bulk_s2g(p.depth + (size_t)(yb + r)*p.depth_stride + x0, zout + r*TW, (uint32_t)(cols*4));
                bulk_s2g(p.color + (size_t)(yb + r)*p.color_pitch_words + x0, col + r*TW, (uint32_t)(cols*4));
            }
            bulk_commit();
        }
        else
        {
            for(int r = 0; r < rows; ++r)
                for(int c = lane; c < cols; c += 32)
                {
                    p.depth[(size_t)(yb + r)*p.depth_stride + x0 + c] = zout[r*TW + c];
                    p.color[(size_t)(yb + r)*p.color_pitch_words + x0 + c] = col[r*TW + c];
                }
            __syncwarp();
        }
    }
    bulk_wait_read();
}

template<int TW, int TH, int WARPS>
static cudaError_t launch_one(const RasterParams &p, int sm_count, cudaStream_t s)
{
    const int smem = WARPS*TileLayout<TW, TH>::kBytesPerWarp;
    auto kern = raster_kernel<TW, TH, WARPS>;
    static bool configured = false;
    if(!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if(e != cudaSuccess) return e;
        configured = true;
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS*32, smem);
    if(per_sm < 1) per_sm = 1;
    unsigned need = (p.ntiles + WARPS - 1)/WARPS;
    unsigned grid = (unsigned)(sm_count*per_sm);           // persistent: a multiple of the SM count
    if(grid > need) grid = need;
    if(grid < 1) grid = 1;
    kern<<<grid, WARPS*32, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_raster(const RasterParams &p, int sm_count, cudaStream_t s)
{
    const int tw = p.v.tile_w, th = p.v.tile_h;
    if(tw == 64 && th == 32) return launch_one<64, 32, 4>(p, sm_count, s);     // 24.6 KB / warp
    if(tw == 32 && th == 32) return launch_one<32, 32, 6>(p, sm_count, s);     // 12.3 KB / warp
    if(tw == 128 && th == 16) return launch_one<128, 16, 4>(p, sm_count, s);
    if(tw == 64 && th == 16) return launch_one<64, 16, 6>(p, sm_count, s);
    return cudaErrorInvalidValue;
}

} // namespace b200r
