// Tile raster kernel (sm_100a): coverage, interpolation, depth test and Gouraud shading.
//
// Restates the per-row body of DrawModel's scalar Gouraud path:
//   span set-up                           projekt.cpp:306-412
//   pixel loop, depth test, ARGB pack     projekt.cpp:423-425, 510-538
//   edge step and crossing exchange       projekt.cpp:542-572
// (the active-edge insert/expire of :202-296 has already been resolved by the set-up kernel,
// which hands this kernel trapezoid segments with both edges' running values).
//
// "The reference's own arithmetic" is a chain of rounded binary32 additions: the value at pixel
// k of a span is k sequential adds from the span's left end.  There is no
// closed form, so the unit of parallel work is the segment, not the pixel:
//
//   * a CTA owns one screen tile at a time, staged in shared memory as ONE 128-bit word per
//     pixel: { depth bits, owner (submission index), ARGB colour, 0 }.  Tile rows enter and
//     leave with TMA bulk copies (cp.async.bulk + mbarrier) and 128-bit shared accesses;
//   * the tile's bin is consumed 32 segments at a time by whichever warp is free (shared-memory
//     ticket counter); each LANE walks one segment: per row the reference's span set-up, then the
//     span pixel by pixel.  Pixels left of the tile are replayed in registers only (adds, no
//     memory); pixels inside the tile are depth-tested against shared memory;
//   * different lanes and warps hold different triangles that may hit the same pixel, so a
//     pixel is updated with one 128-bit compare-and-swap (ATOMS.CAS.128) under the rule
//         z > zold || (z == zold && prim < primold)
//     which is the reference's strict '>' with first-submitted-wins (projekt.cpp:525) made
//     order independent; depth, owner and colour change together, so no ordering between lanes
//     or warps is needed and the pixel loop contains no barrier.
#include "raster_device.cuh"

namespace b200r {

struct __align__(16) Pixel { unsigned z, prim, color, pad; };

__device__ __forceinline__ Pixel lds_pixel(const Pixel *p)
{
    Pixel r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.z), "=r"(r.prim), "=r"(r.color), "=r"(r.pad) : "r"(smem_addr(p)) : "memory");
    return r;
}
__device__ __forceinline__ float lds_depth(const Pixel *p)
{
    float z;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(z) : "r"(smem_addr(p)) : "memory");
    return z;
}

// RoundR32ToU32(c*255) per channel, A R G B from a r g b (projekt.cpp:520-523); no clamp.
// guarded: exact cvtss2si behaviour for NaN / |v| >= 2^31 (only segments flagged by set-up).
__device__ __forceinline__ uint32_t pack_argb(float r, float g, float b, float a, bool guarded)
{
    const float fa = fmul(a, 255.0f), fr = fmul(r, 255.0f), fg = fmul(g, 255.0f), fb = fmul(b, 255.0f);
    if(guarded)
        return ((uint32_t)round_s32(fa) << 24) | ((uint32_t)round_s32(fr) << 16) |
               ((uint32_t)round_s32(fg) << 8) | ((uint32_t)round_s32(fb) << 0);
    return ((uint32_t)__float2int_rn(fa) << 24) | ((uint32_t)__float2int_rn(fr) << 16) |
           ((uint32_t)__float2int_rn(fg) << 8) | ((uint32_t)__float2int_rn(fb) << 0);
}

template<int TW, int TH>
struct TileLayout
{
    static constexpr int kPix = TW*TH;
    static constexpr int kBytes = kPix*16 + 16;     // pixels + mbarrier
};

template<int TW, int TH, int WARPS>
__global__ void __launch_bounds__(WARPS*32)
raster_kernel(const RasterParams p)
{
    constexpr int NPIX = TW*TH;
    constexpr int NT = WARPS*32;
    constexpr int PPT = NPIX/NT;                           // pixels per thread when (un)packing
    static_assert(NPIX % NT == 0, "tile must divide evenly over the CTA");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned s_tile, s_ticket;

    if(*p.seg_total > p.seg_capacity || *p.pair_total > p.pair_capacity) return;   // host re-issues

    const int tid = threadIdx.x, lane = tid & 31;
    Pixel *tile = reinterpret_cast<Pixel *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + NPIX*16);
    // staging inside the pixel area: packed depth rows at byte 12*NPIX, packed colour rows at 8*NPIX
    float *zstage = reinterpret_cast<float *>(smem_raw + NPIX*12);
    uint32_t *cstage = reinterpret_cast<uint32_t *>(smem_raw + NPIX*8);

    if(tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t phase = 0;

    const float wf = (float)p.v.width;
    const float wf_m1 = fsub(wf, 1.0f);
    const int band_rows = p.v.band_y1 - p.v.band_y0;

    while(true)
    {
        if(tid == 0) { s_tile = atomicAdd(p.work_counter, 1u); s_ticket = 0; }
        __syncthreads();
        const unsigned tile_id = s_tile;
        if(tile_id >= p.ntiles) break;
        const unsigned cnt = p.tile_count[tile_id];
        const unsigned off = p.tile_offset[tile_id];
        const int tx = (int)(tile_id % (unsigned)p.v.tiles_x), ty = (int)(tile_id / (unsigned)p.v.tiles_x);
        const int x0 = tx*TW;
        const int yb = ty*TH;                              // band-relative first row
        const int ys0 = p.v.band_y0 + yb;                  // screen row of the tile's first row
        const int cols = min(TW, p.v.width - x0);
        const int rows = min(TH, band_rows - yb);
        if(cnt == 0) { __syncthreads(); continue; }

        // ---------------- stage the tile: depth -> zstage, colour -> cstage ------------------
        fence_proxy_async();                               // generic writes above -> async-proxy writes below
        __syncthreads();                                   // (also: everyone has read s_tile)
        if(p.bulk_ok)
        {
            if(tid == 0) mbar_expect_tx(bar, (uint32_t)(rows*cols*8));
            __syncthreads();
            for(int r = tid; r < rows; r += NT)
            {
                bulk_g2s(zstage + r*TW, p.depth + (size_t)(yb + r)*p.depth_stride + x0, (uint32_t)(cols*4), bar);
                bulk_g2s(cstage + r*TW, p.color + (size_t)(yb + r)*p.color_pitch_words + x0, (uint32_t)(cols*4), bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        else
        {
            for(int i = tid; i < rows*cols; i += NT)
            {
                const int r = i / cols, c = i % cols;
                zstage[r*TW + c] = p.depth[(size_t)(yb + r)*p.depth_stride + x0 + c];
                cstage[r*TW + c] = p.color[(size_t)(yb + r)*p.color_pitch_words + x0 + c];
            }
            __syncthreads();
        }
        // expand to { z, owner = -1, colour, 0 }: owner -1 is "already in the target", wins ties
        {
            float zr[PPT]; uint32_t cr[PPT];
#pragma unroll
            for(int k = 0; k < PPT; ++k) { zr[k] = zstage[tid + k*NT]; cr[k] = cstage[tid + k*NT]; }
            bulk_wait_read();                              // the previous tile's stores have left [0, 8N)
            __syncthreads();
#pragma unroll
            for(int k = 0; k < PPT; ++k)
            {
                Pixel px; px.z = __float_as_uint(zr[k]); px.prim = 0xffffffffu; px.color = cr[k]; px.pad = 0;
                tile[tid + k*NT] = px;
            }
        }
        __syncthreads();

        // ---------------- rasterise the bin, 32 segments per ticket ----------------------------
        while(true)
        {
            unsigned b = 0;
            if(lane == 0) b = atomicAdd(&s_ticket, 32u);
            b = __shfl_sync(0xffffffffu, b, 0);
            if(b >= cnt) break;
            const bool have = (b + lane) < cnt;
            const unsigned seg = have ? __ldg(p.pair_list + off + b + lane) : 0u;
            const float4 *S = reinterpret_cast<const float4 *>(p.segs + (size_t)seg*kSegWords);
            const uint4 h = __ldg(reinterpret_cast<const uint4 *>(S));
            const float4 q1 = __ldg(S + 1), q2 = __ldg(S + 2), q3 = __ldg(S + 3);
            const float4 q4 = __ldg(S + 4), q5 = __ldg(S + 5), q6 = __ldg(S + 6);
            const int prim = (int)h.x;
            const int y0 = (int)h.y;
            const int nrows = have ? (int)(h.z & 0xffffu) : 0;
            const bool guarded = (h.z & kSegNonFinite) != 0;
            // L / R: running XMin, ZMin, MinColor and their per-row gradients
            float lx = q1.x, lz = q1.y, l0 = q1.z, l1 = q1.w, l2 = q2.x, l3 = q2.y;
            float ldx = q2.z, ldz = q2.w, ld0 = q3.x, ld1 = q3.y, ld2 = q3.z, ld3 = q3.w;
            float rx = q4.x, rz = q4.y, r0 = q4.z, r1 = q4.w, r2 = q5.x, r3 = q5.y;
            float rdx = q5.z, rdz = q5.w, rd0 = q6.x, rd1 = q6.y, rd2 = q6.z, rd3 = q6.w;

            for(int k = 0; k < nrows; ++k)
            {
                const int y = y0 + k;
                // ---- span set-up, projekt.cpp:306-412 ----
                const float xdiff = roundf(fsub(rx, lx));                     // :311-312
                float zi = 0.0f, i0 = 0.0f, i1 = 0.0f, i2 = 0.0f, i3 = 0.0f;
                if(xdiff != 0.0f)                                             // :333-363
                {
                    i0 = fdiv(fsub(r0, l0), xdiff); i1 = fdiv(fsub(r1, l1), xdiff);
                    i2 = fdiv(fsub(r2, l2), xdiff); i3 = fdiv(fsub(r3, l3), xdiff);
                    zi = fdiv(fsub(rz, lz), xdiff);
                }
                float z = lz, c0 = l0, c1 = l1, c2 = l2, c3 = l3;             // :375-379
                float xoff = 0.0f, leftx = lx;                                // :381-390
                if(leftx < 0.0f) { xoff = -leftx; leftx = 0.0f; }
                else if(leftx >= wf) { leftx = wf_m1; }
                float rightx = rx;                                            // :392-400
                if(rightx < 0.0f) { rightx = 0.0f; }
                else if(rightx >= wf) { rightx = wf_m1; }
                const int minx = round_s32(leftx), maxx = round_s32(rightx);  // :402-406
                z = fadd(z, fmul(xoff, zi));                                  // :408
                c0 = fadd(c0, fmul(xoff, i0)); c1 = fadd(c1, fmul(xoff, i1)); // :412
                c2 = fadd(c2, fmul(xoff, i2)); c3 = fadd(c3, fmul(xoff, i3));
                const int xe = min(maxx, x0 + cols - 1);
                if(minx <= xe && maxx >= x0)
                {
                    // pixels left of the tile: the reference's per-pixel adds (:534-535), registers only
                    for(int s = minx; s < x0; ++s)
                    {
                        c0 = fadd(c0, i0); c1 = fadd(c1, i1); c2 = fadd(c2, i2); c3 = fadd(c3, i3);
                        z = fadd(z, zi);
                    }
                    // ---- pixel loop, projekt.cpp:423-425, 510-538 ----
                    Pixel *row = tile + (y - ys0)*TW - x0;
                    for(int x = max(minx, x0); x <= xe; ++x)
                    {
                        const float zo = lds_depth(row + x);
                        if(z >= zo)
                        {
                            Pixel mine;
                            mine.z = __float_as_uint(z); mine.prim = (unsigned)prim;
                            mine.color = pack_argb(c0, c1, c2, c3, guarded); mine.pad = 0;
                            Pixel old = lds_pixel(row + x);
                            while(true)
                            {
                                const float oz = __uint_as_float(old.z);
                                const int op = (int)old.prim;
                                if(!(z > oz || (z == oz && prim < op))) break;        // :525 + tie rule
                                const Pixel prev = atomicCAS(row + x, old, mine);
                                if(prev.z == old.z && prev.prim == old.prim && prev.color == old.color) break;
                                old = prev;
                            }
                        }
                        c0 = fadd(c0, i0); c1 = fadd(c1, i1); c2 = fadd(c2, i2); c3 = fadd(c3, i3);   // :534
                        z = fadd(z, zi);                                                              // :535
                    }
                }
                // ---- one row down both edges, projekt.cpp:542-549; exchange if crossed, :562-572 ----
                lx = fadd(lx, ldx); lz = fadd(lz, ldz);
                l0 = fadd(l0, ld0); l1 = fadd(l1, ld1); l2 = fadd(l2, ld2); l3 = fadd(l3, ld3);
                rx = fadd(rx, rdx); rz = fadd(rz, rdz);
                r0 = fadd(r0, rd0); r1 = fadd(r1, rd1); r2 = fadd(r2, rd2); r3 = fadd(r3, rd3);
                if(lx > rx)
                {
                    float t;
                    t = lx; lx = rx; rx = t;       t = lz; lz = rz; rz = t;
                    t = l0; l0 = r0; r0 = t;       t = l1; l1 = r1; r1 = t;
                    t = l2; l2 = r2; r2 = t;       t = l3; l3 = r3; r3 = t;
                    t = ldx; ldx = rdx; rdx = t;   t = ldz; ldz = rdz; rdz = t;
                    t = ld0; ld0 = rd0; rd0 = t;   t = ld1; ld1 = rd1; rd1 = t;
                    t = ld2; ld2 = rd2; rd2 = t;   t = ld3; ld3 = rd3; rd3 = t;
                }
            }
        }
        __syncthreads();

        // ---------------- write the tile back: pack depth to [0,4N), colour to [4N,8N) -----------
        {
            Pixel px[PPT];
#pragma unroll
            for(int k = 0; k < PPT; ++k) px[k] = tile[tid + k*NT];
            __syncthreads();
            float *zout = reinterpret_cast<float *>(smem_raw);
            uint32_t *cout_ = reinterpret_cast<uint32_t *>(smem_raw + NPIX*4);
#pragma unroll
            for(int k = 0; k < PPT; ++k) { zout[tid + k*NT] = __uint_as_float(px[k].z); cout_[tid + k*NT] = px[k].color; }
            __syncthreads();
            if(p.bulk_ok)
            {
                fence_proxy_async();
                __syncthreads();
                for(int r = tid; r < rows; r += NT)
                {
                    bulk_s2g(p.depth + (size_t)(yb + r)*p.depth_stride + x0, zout + r*TW, (uint32_t)(cols*4));
                    bulk_s2g(p.color + (size_t)(yb + r)*p.color_pitch_words + x0, cout_ + r*TW, (uint32_t)(cols*4));
                }
                bulk_commit();
            }
            else
            {
                for(int i = tid; i < rows*cols; i += NT)
                {
                    const int r = i / cols, c = i % cols;
                    p.depth[(size_t)(yb + r)*p.depth_stride + x0 + c] = zout[r*TW + c];
                    p.color[(size_t)(yb + r)*p.color_pitch_words + x0 + c] = cout_[r*TW + c];
                }
            }
        }
    }
    bulk_wait_read();
}

template<int TW, int TH, int WARPS>
static cudaError_t launch_one(const RasterParams &p, int sm_count, cudaStream_t s)
{
    const int smem = TileLayout<TW, TH>::kBytes;
    auto kern = raster_kernel<TW, TH, WARPS>;
    static bool configured = false;
    static int per_sm = 1;
    if(!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if(e != cudaSuccess) return e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS*32, smem);
        if(per_sm < 1) per_sm = 1;
        configured = true;
    }
    unsigned grid = (unsigned)(sm_count*per_sm);           // persistent: a multiple of the SM count
    if(grid > p.ntiles) grid = p.ntiles;
    if(grid < 1) grid = 1;
    kern<<<grid, WARPS*32, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_raster(const RasterParams &p, int sm_count, cudaStream_t s)
{
    const int tw = p.v.tile_w, th = p.v.tile_h;
    if(tw == 64 && th == 32) return launch_one<64, 32, 8>(p, sm_count, s);     // 32 KB tile
    if(tw == 32 && th == 32) return launch_one<32, 32, 8>(p, sm_count, s);     // 16 KB
    if(tw == 128 && th == 16) return launch_one<128, 16, 8>(p, sm_count, s);   // 32 KB
    if(tw == 64 && th == 16) return launch_one<64, 16, 8>(p, sm_count, s);     // 16 KB
    if(tw == 128 && th == 32) return launch_one<128, 32, 8>(p, sm_count, s);   // 64 KB
    return cudaErrorInvalidValue;
}

} // namespace b200r
