// Tile raster kernel (sm_100a): per-pixel coverage, interpolation, depth test and colour write.
//
// Restates DrawModel's pixel loop -- Gouraud (projekt.cpp:423-425, 510-538), per-pixel Phong
// (:450-509) and texturing (:427-446), one kernel variant per kind of frame (see MODE below) -- on
// the span records the set-up kernel produced (span set-up, :306-412, is done there, once per row).
//
// "The reference's own arithmetic" is a chain of rounded binary32 additions: the value at pixel
// k of a span is k sequential adds from the span's left end.  There is no
// closed form, so the unit of parallel work is the SPAN, not the pixel:
//
//   * a CTA owns one screen tile at a time, staged in shared memory as TWO arrays: a plane of
//     4-byte depths that the early test reads (32 consecutive pixels = 32 banks), and one 128-bit
//     word per pixel { depth bits, owner (submission index), ARGB colour, 0 } that only the update
//     path touches.  Tile rows enter and leave with TMA bulk copies (cp.async.bulk + mbarrier);
//   * the tile's queue is a flat list of spans.  Lanes are persistent: a lane that runs out of
//     pixels waits until kRefill lanes of its warp are idle, then the idle lanes take new spans
//     with ONE warp-aggregated ticket (ballot + popc + shuffle) on a shared-memory counter;
//   * lanes with pixels advance in lock step, kRound pixels between two warp votes: the round's
//     depths are loaded up front (independent loads), then the adds of :534-535 and the early
//     test run as one predicated chain.  Pixels left of the tile are the same steps with the
//     load predicated off (adds only, in registers);
//   * a pixel whose depth test may pass parks its lane; as soon as kPend lanes are parked they run
//     the update together: pack ARGB (:520-523) and one 128-bit compare-and-swap (ATOMS.CAS.128)
//     under the rule   z > zold || (z == zold && prim < primold)
//     which is the reference's strict '>' with first-submitted-wins (:525) made order
//     independent.  Depth, owner and colour change together, so no ordering between lanes or
//     warps is needed and the pixel loop contains no barrier.  The depth plane is a conservative
//     copy (never above the word's depth: every value written to it was in the word at some time
//     and depths only rise), so the early test can only err towards a needless exact test.
#include "raster_device.cuh"

#include <mutex>

namespace b200r {

struct __align__(16) Pixel { unsigned z, prim, color, pad; };

// -DB200R_STATS: developer build that counts what the lanes do (tools/raster_stats.py); never shipped
#if defined(B200R_STATS)
__device__ unsigned long long g_raster_stats[32];
#define STAT_CLK_ADD(i, v) atomicAdd(&g_raster_stats[i], (unsigned long long)(v))
#if B200R_STATS == 2                       // clocks only: the counters' atomics distort them
#define STAT_ADD(i, v) ((void)0)
#else
#define STAT_ADD(i, v) atomicAdd(&g_raster_stats[i], (unsigned long long)(v))
#endif
#define STAT_WARP(i, v) do { if((threadIdx.x & 31) == 0) STAT_ADD(i, v); } while(0)
#define STAT_CLOCK(var) const long long var = clock64()
#else
#define STAT_ADD(i, v) ((void)0)
#define STAT_WARP(i, v) ((void)0)
#define STAT_CLOCK(var) ((void)0)
#endif

// shared-state-space (32-bit) addressing: no generic-address conversion in the pixel loop
__device__ __forceinline__ Pixel lds_pixel(uint32_t addr)
{
    Pixel r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.z), "=r"(r.prim), "=r"(r.color), "=r"(r.pad) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ float lds_depth(uint32_t addr)
{
    float z;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(z) : "r"(addr) : "memory");
    return z;
}
// the early test's load: predicated off (and NaN, which no depth compares >= to) outside the tile / span
__device__ __forceinline__ float lds_depth_if(uint32_t addr, bool on)
{
    float z;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.u32 p, %2, 0;\n\t"
                 "mov.b32 %0, 0x7fc00000;\n\t"
                 "@p ld.shared.f32 %0, [%1];\n\t}"
                 : "=f"(z) : "r"(addr), "r"((unsigned)on) : "memory");
    return z;
}
__device__ __forceinline__ void sts_depth(uint32_t addr, float z)
{
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(z) : "memory");
}
// 128-bit compare-and-swap on a shared-memory pixel (SASS: ATOMS.CAS.128)
__device__ __forceinline__ Pixel cas_pixel(uint32_t addr, const Pixel &cmp, const Pixel &val)
{
    unsigned long long clo = ((unsigned long long)cmp.prim << 32) | cmp.z, chi = ((unsigned long long)cmp.pad << 32) | cmp.color;
    unsigned long long vlo = ((unsigned long long)val.prim << 32) | val.z, vhi = ((unsigned long long)val.pad << 32) | val.color;
    unsigned long long rlo, rhi;
    asm volatile("{\n\t.reg .b128 c, s, r;\n\t"
                 "mov.b128 c, {%3, %4};\n\t"
                 "mov.b128 s, {%5, %6};\n\t"
                 "atom.shared.cas.b128 r, [%2], c, s;\n\t"
                 "mov.b128 {%0, %1}, r;\n\t}"
                 : "=l"(rlo), "=l"(rhi) : "r"(addr), "l"(clo), "l"(chi), "l"(vlo), "l"(vhi) : "memory");
    Pixel r;
    r.z = (unsigned)rlo; r.prim = (unsigned)(rlo >> 32); r.color = (unsigned)rhi; r.pad = (unsigned)(rhi >> 32);
    return r;
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

// ---- per-pixel Phong shading, projekt.cpp:450-483 (UnprojectVertex :147-160) -----------------
__device__ __forceinline__ void normalize3r(float &x, float &y, float &z)
{
    const float s = fdiv(1.0f, __fsqrt_rn(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z))));
    x = fmul(s, x); y = fmul(s, y); z = fmul(s, z);
}

// c: interpolated (unlit) colour, n: interpolated normal, (X, Row, Z): the pixel's screen position
// and depth.  pow(x, 16) is a double-precision pow narrowed to r32 in the reference (:478); x^16
// by four double squarings differs from it by at most 3 double roundings, far inside the +-1 LSB
// colour tolerance of this path.  The pack is always the guarded one: a degenerate normal or half
// vector yields NaN, which cvtss2si turns into 0x80000000 (:490-493).
__device__ __noinline__ uint32_t phong_pixel(const ViewParams &v, float c0, float c1, float c2, float c3,
                                             float n0, float n1, float n2, float X, float Row, float Z, bool guarded)
{
    const float dist = fsub(v.dist, Z);                                     // :152
    const float inv = fdiv(1.0f, v.m2p);
    const float ax = fmul(inv, fsub(X, v.cx)), ay = fmul(inv, fsub(Row, v.cy));   // :154
    const float sc = fdiv(dist, v.focal);                                   // :155
    const float px = fmul(sc, ax), py = fmul(sc, ay), pz = Z;
    const float c[4] = { c0, c1, c2, c3 };
    float f[4] = { 0.0f, 0.0f, 0.0f, 0.0f };                                // :448
    for(int l = 0; l < v.nlights; ++l)
    {
        const DevLight &L = v.lights[l];
        if(l == 0)                                                          // :464-467
        {
#pragma unroll
            for(int i = 0; i < 4; ++i) f[i] = fmul(c[i], v.amb[i]);
        }
        float lx = fsub(L.px, px), ly = fsub(L.py, py), lz = fsub(L.pz, pz);
        normalize3r(lx, ly, lz);                                            // :471
        const float cosi = clamp01(fadd(fadd(fmul(n0, lx), fmul(n1, ly)), fmul(n2, lz)));   // :474
        float vx = -px, vy = -py, vz = -pz;
        normalize3r(vx, vy, vz);                                            // :475
        float hx = fadd(lx, vx), hy = fadd(ly, vy), hz = fadd(lz, vz);
        normalize3r(hx, hy, hz);                                            // :476
        float term = clamp01(fadd(fadd(fmul(n0, hx), fmul(n1, hy)), fmul(n2, hz)));         // :477
        double d = (double)term;
        d = __dmul_rn(d, d); d = __dmul_rn(d, d); d = __dmul_rn(d, d); d = __dmul_rn(d, d);
        term = __double2float_rn(d);                                        // :478
        const float I[4] = { L.ir, L.ig, L.ib, L.ia };
#pragma unroll
        for(int i = 0; i < 4; ++i)                                          // :480
            f[i] = fadd(f[i], fadd(fmul(cosi, fmul(c[i], I[i])), fmul(term, fmul(1.0f, I[i]))));
    }
    return pack_argb(clamp01(f[0]), clamp01(f[1]), clamp01(f[2]), clamp01(f[3]), guarded);   // :483-493
}

// ---- textured pixel, projekt.cpp:427-446 ----------------------------------------------------
// (uz, vz, oz): the span's interpolated u/z, v/z, 1/z at this pixel.  Nearest texel at
// Round(uv*(dim-1)); coordinates that leave the bitmap are clamped (the reference reads outside it;
// cvtss2si's INT_MIN for NaN / overflow clamps to 0).  Unlit (Gouraud) pixels are the texel word
// itself: Round((b/255)*255) == b for every byte (tests/test_oracle_tex.py), and the channel order
// of the pack (:520-523) is the texel's.  Phong pixels take the texel as base colour (:444, :466).
__device__ __noinline__ uint32_t tex_pixel(const ViewParams &v, const TexDesc *textures, int tex_id,
                                           float uz, float vz, float oz, bool phong,
                                           float n0, float n1, float n2, float X, float Row, float Z)
{
    const TexDesc td = textures[tex_id];
    const float inv = fdiv(1.0f, oz);                                       // :429
    const float fu = fmul(inv, uz), fv = fmul(inv, vz);
    const float tx = fmul(fu, (float)(td.w - 1)), ty = fmul(fv, (float)(td.h - 1));   // :430-432
    int ix = round_s32(tx), iy = round_s32(ty);                             // :433-434
    ix = min(max(ix, 0), td.w - 1);
    iy = min(max(iy, 0), td.h - 1);
    const uint32_t texel = __ldg(reinterpret_cast<const uint32_t *>(
        reinterpret_cast<const unsigned char *>(td.mem) + (size_t)iy*(size_t)td.pitch) + ix);   // :436-438
    if(!phong) return texel;
    const float sa = fdiv((float)((texel >> 24) & 0xFFu), 255.0f);          // :440-443
    const float sr = fdiv((float)((texel >> 16) & 0xFFu), 255.0f);
    const float sg = fdiv((float)((texel >> 8) & 0xFFu), 255.0f);
    const float sb = fdiv((float)(texel & 0xFFu), 255.0f);
    return phong_pixel(v, sr, sg, sb, sa, n0, n1, n2, X, Row, Z, true);
}

// Pixels a lane walks between two warp votes.  The round's depths are loaded before the chain
// of adds starts, so a lane's per-pixel latency is one FADD, not one shared-memory round trip.
// A lane that parks idles for the rest of its round: on C3 (11 % of the fragments pass at the time
// they are tested) rounds of 8 / 6 / 4 / 2 pixels give 0.75 / 0.70 / 0.67 / 0.66 ms.
#ifndef B200R_ROUND
#define B200R_ROUND 4
#endif
constexpr int kRound = B200R_ROUND;
// Pixels a parked lane may update in one visit of the update path while its run of passing pixels lasts.
#ifndef B200R_HOT
#define B200R_HOT 4
#endif
constexpr int kHot = B200R_HOT;
// Spans a lane may take (and cull) in one visit of the refill path.
#ifndef B200R_CULL_TRIES
#define B200R_CULL_TRIES 3
#endif
constexpr int kCullTries = B200R_CULL_TRIES;
// queue entries between two recomputations of a tile's depth floors
#ifndef B200R_FLOOR_EVERY
#define B200R_FLOOR_EVERY 256
#endif
// finished tiles are stored from registers (1) or packed into shared memory and bulk-copied row by row (0)
#ifndef B200R_DIRECT_WB
#define B200R_DIRECT_WB 1
#endif
// resident CTAs per SM the register allocation is held to (a 40 KB tile allows 5)
#ifndef B200R_MINB
#define B200R_MINB 5
#endif

template<int TW, int TH>
struct TileLayout
{
    static constexpr int kPix = TW*TH;
    static constexpr int kPlane = kPix*16;                      // byte offset of the depth plane
    static constexpr int kBar = kPix*20 + 64;                   // mbarrier, behind the plane's over-read pad
    static constexpr int kBytes = kBar + 16;
};

// MODE (RasterParams::mode), one kernel per kind of frame so that the common one stays lean:
//   kRasterPlain     Gouraud only: 16-word span records
//   kRasterGeneral   the frame contains Phong meshes: 24-word records, per-span flags select per-pixel
//                    Phong shading and / or texturing
//   kRasterTextured  textured meshes but no Phong mesh: 16-word records whose colour words carry
//                    u/z, v/z, 1/z for textured spans; a candidate pixel fetches its texel inline
// __grid_constant__: the shaders below take p.v by reference; without it the parameter struct is
// copied to local memory first (440 bytes of stack in the general kernel).
template<int TW, int TH, int WARPS, int MODE>
__global__ void __launch_bounds__(WARPS*32, (MODE == kRasterGeneral ? 3 : B200R_MINB)*8/WARPS)
raster_kernel(const __grid_constant__ RasterParams p)
{
    constexpr bool PHONG = MODE == kRasterGeneral;         // normals, lighting, 24-word records
    constexpr bool TEX = MODE != kRasterPlain;             // spans may be textured
    constexpr int kMaxSharedTex = 16;
    __shared__ TexDesc s_tex[TEX ? kMaxSharedTex : 1];     // the first textures of the table, one load per CTA
    constexpr int NPIX = TW*TH;
    constexpr int NT = WARPS*32;
    constexpr int PPT = NPIX/NT;                           // pixels per thread when (un)packing
    static_assert(NPIX % NT == 0, "tile must divide evenly over the CTA");
    static_assert(TW <= 256 && TH <= 32, "stash entries hold 8 bits of tile column and 5 bits of tile row");
    using Layout = TileLayout<TW, TH>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned s_tile, s_ticket;
    constexpr int kBlocks = TW/32;                         // depth floors are kept per tile row and 32-pixel block
    __shared__ float s_floor[TH][kBlocks];                 // lower bound of the depths of a block
    // per warp: the spans of its last queue chunk that survived the cull
    // { span, tile row | (last column - tile's first column) << 5 | (first column - tile's first column) << 13 }
    __shared__ uint2 s_stash[WARPS][32];

    if(*p.overflow) return;                                // host grows the lists and re-issues the frame

    const int tid = threadIdx.x, lane = tid & 31;
    Pixel *tile = reinterpret_cast<Pixel *>(smem_raw);
    float *zplane = reinterpret_cast<float *>(smem_raw + Layout::kPlane);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + Layout::kBar);
    // colour rows are staged inside the pixel area (bytes [12N, 16N)) until they are expanded;
    // on the way out the colours are packed to bytes [0, 4N) and the depths into the plane
    uint32_t *cstage = reinterpret_cast<uint32_t *>(smem_raw + NPIX*12);

    if(tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if(TEX && tid < kMaxSharedTex && tid < p.texture_count) s_tex[tid] = p.textures[tid];
    __syncthreads();
    uint32_t phase = 0;

    const int band_rows = p.v.band_y1 - p.v.band_y0;

    while(true)
    {
        if(tid == 0) { s_tile = atomicAdd(p.work_counter, 1u); s_ticket = 0; }
        __syncthreads();
        const unsigned tile_id = p.tile_begin + s_tile;
        if(tile_id >= p.tile_end) break;
        // the tile's queue = its kDepthBuckets sub-queues, contiguous and nearest bucket first
        const unsigned off = p.tile_offset[tile_id*kDepthBuckets];
        const unsigned cnt = p.tile_offset[(tile_id + 1)*kDepthBuckets] - off;
        const int tx = (int)(tile_id % (unsigned)p.v.tiles_x), ty = (int)(tile_id / (unsigned)p.v.tiles_x);
        const int x0 = tx*TW;
        const int yb = ty*TH;                              // band-relative first row
        const int ys0 = p.v.band_y0 + yb;                  // screen row of the tile's first row
        const int cols = min(TW, p.v.width - x0);
        const int rows = min(TH, band_rows - yb);
        // a tile nothing touches is skipped -- unless the band is mirrored into a gather target, which
        // then receives the tile as it is (the gathered image must equal the band, whatever it held before)
        const bool copy_through = cnt == 0;
        if(copy_through && p.gather_color == nullptr) { if(tid == 0) STAT_ADD(18, 1); __syncthreads(); continue; }
        const bool gather_bulk = p.bulk_ok && p.gather_bulk_ok;
        STAT_CLOCK(t0);

        // ---------------- stage the tile: depth -> plane, colour -> cstage ------------------
        bulk_wait_read();                                  // the previous tile's stores have read the plane and [0, 4N)
        fence_proxy_async();                               // generic writes above -> async-proxy writes below
        __syncthreads();                                   // (also: everyone has read s_tile)
        if(p.bulk_ok)
        {
            if(tid == 0) mbar_expect_tx(bar, (uint32_t)(rows*cols*8));
            __syncthreads();
            for(int r = tid; r < rows; r += NT)
            {
                bulk_g2s(zplane + r*TW, p.depth + (size_t)(yb + r)*p.depth_stride + x0, (uint32_t)(cols*4), bar);
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
                zplane[r*TW + c] = p.depth[(size_t)(yb + r)*p.depth_stride + x0 + c];
                cstage[r*TW + c] = p.color[(size_t)(yb + r)*p.color_pitch_words + x0 + c];
            }
            __syncthreads();
        }
        STAT_CLOCK(t1);
        if(copy_through)
        {
            if(gather_bulk)
            {
                for(int r = tid; r < rows; r += NT)
                {
                    if(p.gather_depth) bulk_s2g(p.gather_depth + (size_t)(ys0 + r)*p.gather_depth_stride + x0, zplane + r*TW, (uint32_t)(cols*4));
                    bulk_s2g(p.gather_color + (size_t)(ys0 + r)*p.gather_pitch_words + x0, cstage + r*TW, (uint32_t)(cols*4));
                }
                bulk_commit();
            }
            else
            {
                for(int i = tid; i < rows*cols; i += NT)
                {
                    const int r = i / cols, c = i % cols;
                    if(p.gather_depth) p.gather_depth[(size_t)(ys0 + r)*p.gather_depth_stride + x0 + c] = zplane[r*TW + c];
                    p.gather_color[(size_t)(ys0 + r)*p.gather_pitch_words + x0 + c] = cstage[r*TW + c];
                }
            }
            __syncthreads();
            continue;
        }
        // expand to { z, owner = -1, colour, 0 }: owner -1 is "already in the target", wins ties
        {
            float zr[PPT]; uint32_t cr[PPT];
#pragma unroll
            for(int k = 0; k < PPT; ++k) { zr[k] = zplane[tid + k*NT]; cr[k] = cstage[tid + k*NT]; }
            __syncthreads();
#pragma unroll
            for(int k = 0; k < PPT; ++k)
            {
                Pixel px; px.z = __float_as_uint(zr[k]); px.prim = 0xffffffffu; px.color = cr[k]; px.pad = 0;
                tile[tid + k*NT] = px;
            }
        }
        __syncthreads();

        // Conservative depth floors for span culling, one per tile row and 32-pixel block.  Depth only
        // ever rises, so a value read while other warps update pixels is at worst too low; recomputed
        // now and every kFloorEvery queue entries.  A span whose depth upper bound (set-up kernel) is
        // strictly below the floor of every block it touches cannot win or tie any pixel and is
        // skipped without walking it.
        constexpr unsigned kFloorEvery = B200R_FLOOR_EVERY;
        const int warp_id = tid >> 5;
        auto refresh_row_floor = [&](int first_row, int row_step)
        {
            const float inf = __int_as_float(0x7f800000);
            for(int r = first_row; r < TH; r += row_step)
            {
#pragma unroll
                for(int c4 = lane; c4 < (TW/4 + 31)/32*32; c4 += 32)
                {
                    float m = inf;                                   // a NaN depth never lets anything pass: no constraint
                    if(c4 < TW/4 && r < rows)
                    {
                        float4 q;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                                     : "r"(smem_addr(zplane + r*TW + 4*c4)) : "memory");
                        const int c = 4*c4;
                        if(c + 0 < cols) m = fminf(m, q.x);
                        if(c + 1 < cols) m = fminf(m, q.y);
                        if(c + 2 < cols) m = fminf(m, q.z);
                        if(c + 3 < cols) m = fminf(m, q.w);
                    }
                    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                    if((lane & 7) == 0 && c4 < TW/4) s_floor[r][c4 >> 3] = m;
                }
            }
        };
        refresh_row_floor(warp_id, WARPS);
        __syncthreads();
        STAT_CLOCK(t2);

        // ---------------- rasterise the tile's span queue: persistent lanes --------------------
        {
            const int kRefill = p.refill_lanes;            // idle lanes that trigger a refill
            // queue entries a warp takes per ticket: 32, or an equal share when the tile's queue is too short to
            // give every warp a full chunk (a tile touched by a handful of triangles is walked by all warps)
            const unsigned chunk = (cnt >= 32u*WARPS) ? 32u : max((cnt + WARPS - 1u)/WARPS, 4u);
            const int kPend = p.pend_lanes;                // parked lanes that trigger the update path
            constexpr unsigned FULL = 0xffffffffu;
            const uint32_t tile_addr = smem_addr(tile), plane_addr = smem_addr(zplane);
            const int xlast = x0 + cols - 1;
            float z = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, zi = 0, i0 = 0, i1 = 0, i2 = 0, i3 = 0;
            float n0 = 0, n1 = 0, n2 = 0, ni0 = 0, ni1 = 0, ni2 = 0;       // Phong: normal and its per-pixel increment
            float shade_dx = 0, shade_row = 0;                             // Phong: X = x + shade_dx, Row for UnprojectVertex
            bool phong_span = false;
            int tex_id = -1;                                               // textured span: c0..c2 carry u/z, v/z, 1/z
            int prim = 0, n_left = 0;                      // n_left: pixels still to visit, the parked one included
            uint32_t zrow = plane_addr;                    // shared address of the depth of the row's first pixel IN the tile
            uint32_t zaddr = plane_addr;                   // ... of the lane's current pixel; below zrow: left of the tile
            bool guarded = false, exhausted = false, pending = false;
            // the warp's stash (warp-uniform): entries [stash_pos, stash_n) are still to hand out
            unsigned stash_n = 0, stash_pos = 0;
            bool queue_done = false;
            unsigned retired_mask = 0;
            while(true)
            {
                // every lane is parked, walking, idle or retired; retired_mask only changes in the refill path
                const unsigned pend_mask = __ballot_sync(FULL, pending);
                const unsigned busy_mask = __ballot_sync(FULL, n_left > 0 && !pending);
                const unsigned need_mask = ~(pend_mask | busy_mask | retired_mask);
                if((need_mask | busy_mask | pend_mask) == 0) break;
                STAT_WARP(22, 1);

                if(pend_mask && (busy_mask == 0 || __popc(pend_mask) >= kPend))
                {
                    // ---- depth-test passes, projekt.cpp:520-529: pack + one 128-bit compare-and-swap ----
                    // Visible pixels come in runs: a lane whose NEXT pixel passes the early test as well
                    // stays here ("hot") for up to kHot pixels instead of going back through a pixel round
                    // to be parked again one pixel later.
                    STAT_WARP(6, 1); STAT_WARP(7, __popc(pend_mask));
                    for(int h = 0; ; )
                    {
                        if(pending)
                        {
                            const uint32_t pa = tile_addr + (zaddr - plane_addr)*4u;
                            const int x = x0 + ((int)(zaddr - zrow) >> 2);
                            B200R_ASSERT(zaddr >= zrow && zaddr < zrow + (uint32_t)cols*4u && zrow >= plane_addr &&
                                         zrow < plane_addr + (uint32_t)(rows*TW*4) && n_left > 0);
                            Pixel old = lds_pixel(pa);
                            Pixel mine;
                            mine.z = __float_as_uint(z); mine.prim = (unsigned)prim;
                            if(PHONG && tex_id >= 0)
                                mine.color = tex_pixel(p.v, p.textures, tex_id, c0, c1, c2, phong_span, n0, n1, n2,
                                                       fadd((float)x, shade_dx), shade_row, z);
                            else if(MODE == kRasterTextured && tex_id >= 0)
                            {
                                // projekt.cpp:427-446, unlit: the texel word itself (see tex_pixel)
                                const TexDesc td = (tex_id < kMaxSharedTex) ? s_tex[tex_id] : p.textures[tex_id];
                                const float inv = fdiv(1.0f, c2);
                                const float tx = fmul(fmul(inv, c0), (float)(td.w - 1)), ty = fmul(fmul(inv, c1), (float)(td.h - 1));
                                const int ix = min(max(round_s32(tx), 0), td.w - 1), iy = min(max(round_s32(ty), 0), td.h - 1);
                                mine.color = __ldg(reinterpret_cast<const uint32_t *>(
                                    reinterpret_cast<const unsigned char *>(td.mem) + (size_t)iy*(size_t)td.pitch) + ix);
                            }
                            else
                                mine.color = (PHONG && phong_span)
                                             ? phong_pixel(p.v, c0, c1, c2, c3, n0, n1, n2, fadd((float)x, shade_dx), shade_row, z, true)
                                             : pack_argb(c0, c1, c2, c3, guarded);
                            mine.pad = 0;
                            float now = z;                                              // the word's depth when we leave
                            while(true)
                            {
                                const float oz = __uint_as_float(old.z);
                                const int op = (int)old.prim;
                                // :525 + tie rule (first submitted wins; with B200R_AVX_DEPTH_GE, :3205, the last one)
                                if(!(z > oz || (z == oz && (p.v.depth_ge ? prim > op : prim < op)))) { now = oz; STAT_ADD(10, 1); break; }
                                const Pixel prev = cas_pixel(pa, old, mine);
                                STAT_ADD(8, 1);
                                // (depth, owner) identify a pixel's contents: no need to compare the colour
                                if(prev.z == old.z && prev.prim == old.prim) break;
                                STAT_ADD(9, 1);
                                old = prev;
                            }
                            sts_depth(zaddr, now);                                      // never above the word's depth
                            if(PHONG && phong_span) { n0 = fadd(n0, ni0); n1 = fadd(n1, ni1); n2 = fadd(n2, ni2); normalize3r(n0, n1, n2); }   // :504
                            c0 = fadd(c0, i0); c1 = fadd(c1, i1); c2 = fadd(c2, i2); c3 = fadd(c3, i3);   // :534
                            z = fadd(z, zi);                                                              // :535
                            zaddr += 4u; --n_left;
                            // the next pixel lies in the tile (n_left counts pixels up to the tile's last column)
                            pending = kHot > 1 && n_left > 0 && z >= lds_depth(zaddr);
                        }
                        if(++h >= kHot) break;
                        const unsigned hot_mask = __ballot_sync(FULL, pending);
                        if(hot_mask == 0 || (busy_mask != 0 && __popc(hot_mask) < kPend)) break;
                        STAT_WARP(26, 1); STAT_WARP(27, __popc(hot_mask));
                    }
                }
                else if(need_mask && (busy_mask == 0 || __popc(need_mask) >= kRefill))
                {
                    // ---- idle lanes take the next spans of the queue ----
                    // The WARP takes a chunk of 32 queue entries with one ticket on the tile's counter: all 32
                    // lanes load one entry's column range and depth bound each and cull it against the depth
                    // floors of the blocks it touches (more than half of C3's entries die here; doing it at
                    // full lane width costs a third of per-lane culling).  The survivors are compacted into the
                    // warp's stash in shared memory; idle lanes pop them, a few at a time, without a global
                    // ticket.  A chunk whose entries are all culled is replaced at once.
                    STAT_WARP(11, 1); STAT_WARP(12, __popc(need_mask));
                    for(int tries = 0; stash_pos >= stash_n && !queue_done && tries < kCullTries; ++tries)
                    {
                        unsigned b = 0;
                        if(lane == 0) b = atomicAdd(&s_ticket, chunk);
                        b = __shfl_sync(FULL, b, 0);
                        if(b >= cnt) { queue_done = true; break; }
                        const unsigned chunk_n = min(chunk, cnt - b);
                        B200R_ASSERT(off + b + chunk_n <= p.pair_capacity);
                        if(b/kFloorEvery != (b + chunk)/kFloorEvery) refresh_row_floor(0, 1);
                        bool live = false;
                        uint2 e = make_uint2(0u, 0u);
                        if((unsigned)lane < chunk_n)
                        {
                            const unsigned sp = __ldg(p.pair_list + off + b + lane);
                            B200R_ASSERT(sp < p.span_capacity);
                            const float4 *S = reinterpret_cast<const float4 *>(p.spans + (size_t)sp*(PHONG ? kSpanWordsPhong : kSpanWords));
                            const float4 q0 = __ldg(S), q3 = __ldg(S + 3);
                            const int y = __float_as_int(q0.y);
                            const int minx = __float_as_int(q0.z), maxx = __float_as_int(q0.w);
                            const int xe = min(maxx, xlast), xs = max(minx, x0);
                            B200R_ASSERT(y >= ys0 && y < ys0 + rows && minx >= 0);
                            const int n = (minx <= xe && maxx >= x0) ? (xe - minx + 1) : 0;
                            STAT_ADD(3, 1); if(n == 0) STAT_ADD(5, 1);
                            if(n > 0)
                            {
                                const int b0 = (xs - x0) >> 5, b1 = (xe - x0) >> 5;
                                float fl = __int_as_float(0x7f800000);
#pragma unroll
                                for(int k = 0; k < kBlocks; ++k)
                                    if(k >= b0 && k <= b1) fl = fminf(fl, s_floor[y - ys0][k]);
                                live = !(q3.w < fl);                          // below every floor: cannot win or tie anywhere
                                if(!live) STAT_ADD(4, 1);
                                e = make_uint2(sp, (unsigned)(y - ys0) | ((unsigned)(xe - x0) << 5) | ((unsigned)(minx - x0) << 13));
                                if(live) { STAT_ADD(2, max(x0 - minx, 0)); STAT_ADD(1, n - max(x0 - minx, 0)); }
                            }
                        }
                        const unsigned lm = __ballot_sync(FULL, live);
                        if(live) s_stash[warp_id][__popc(lm & ((1u << lane) - 1u))] = e;
                        stash_n = (unsigned)__popc(lm); stash_pos = 0;
                        __syncwarp();
                    }
                    const bool need = n_left == 0 && !exhausted;
                    const unsigned avail = stash_n - stash_pos;
                    if(avail == 0)
                    {
                        // nothing to hand out: the queue is finished (lanes retire), or kCullTries chunks in a
                        // row were culled completely (the lanes ask again)
                        if(need && queue_done) exhausted = true;
                        retired_mask = __ballot_sync(FULL, exhausted);
                    }
                    else
                    {
                        const unsigned rank = (unsigned)__popc(need_mask & ((1u << lane) - 1u));
                        if(need && rank < avail)
                        {
                            const uint2 e = s_stash[warp_id][stash_pos + rank];
                            const float4 *S = reinterpret_cast<const float4 *>(p.spans + (size_t)e.x*(PHONG ? kSpanWordsPhong : kSpanWords));
                            const float4 q0 = __ldg(S), q1 = __ldg(S + 1), q2 = __ldg(S + 2), q3 = __ldg(S + 3);
                            const int first = (int)e.y >> 13;                        // first column - x0: negative left of the tile
                            n_left = (int)((e.y >> 5) & 255u) - first + 1;
                            zrow = plane_addr + (e.y & 31u)*(uint32_t)(TW*4);
                            zaddr = zrow + (uint32_t)(first*4);                       // wraps below zrow left of the tile
                            prim = __float_as_int(q0.x);
                            z = q1.x; c0 = q1.y; c1 = q1.z; c2 = q1.w;
                            c3 = q2.x; zi = q2.y; i0 = q2.z; i1 = q2.w;
                            i2 = q3.x; i3 = q3.y;
                            guarded = (__float_as_uint(q3.z) & kSpanNonFinite) != 0;
                            if(TEX)
                            {
                                const unsigned fl = __float_as_uint(q3.z);
                                tex_id = (fl & kSpanTex) ? (int)((fl >> 8) & 0xffffu) : -1;
                            }
                            if(PHONG)
                            {
                                const unsigned fl = __float_as_uint(q3.z);
                                phong_span = (fl & kSpanPhong) != 0;
                                shade_dx = 0.0f; shade_row = (float)__float_as_int(q0.y);
                                if(phong_span)
                                {
                                    const float4 q4 = __ldg(S + 4), q5 = __ldg(S + 5);
                                    n0 = q4.x; n1 = q4.y; n2 = q4.z; ni0 = q4.w; ni1 = q5.x; ni2 = q5.y;
                                    if(fl & kSpanAlias) { shade_dx = q4.w; shade_row = q5.x; ni0 = ni1 = ni2 = 0.0f; }
                                }
                            }
                        }
                        stash_pos += min((unsigned)__popc(need_mask), avail);
                        __syncwarp();
                    }
                    continue;
                }

                // ---- kRound pixel steps, projekt.cpp:423-425, 525, 534-535 ----
                // A lane whose depth test may pass parks at that pixel (pending); the votes at the top of
                // the loop run the update path once kPend lanes are parked (low-overdraw scenes pass on
                // nearly every pixel, high-overdraw scenes rarely: both stay converged).
                {
                    const int m = pending ? 0 : n_left;                       // pixels this lane may visit in this round
                    STAT_WARP(0, 1); STAT_WARP(19, __popc(busy_mask)); STAT_WARP(20, __popc(pend_mask)); STAT_WARP(21, __popc(need_mask));
                    STAT_ADD(23, min(m, kRound));
                    const int lo = max((int)(zrow - zaddr), 0);               // bytes until the lane's first pixel IN the tile
                    float zo[kRound];
#pragma unroll
                    for(int k = 0; k < kRound; ++k)
                    {
                        B200R_ASSERT(!(k < m && 4*k >= lo) || (zaddr + 4u*k >= zrow && zaddr + 4u*k < zrow + (uint32_t)cols*4u));
                        zo[k] = lds_depth_if(zaddr + 4u*k, k < m && 4*k >= lo);   // NaN when not loaded: never >= anything
                    }
                    const int before = n_left;
#pragma unroll
                    for(int k = 0; k < kRound; ++k)
                    {
                        pending = pending || (z >= zo[k]);                    // sticky: a parked lane's z no longer moves
                        if(k < m && !pending)
                        {
                            if(PHONG && phong_span) { n0 = fadd(n0, ni0); n1 = fadd(n1, ni1); n2 = fadd(n2, ni2); normalize3r(n0, n1, n2); }   // :504
                            c0 = fadd(c0, i0); c1 = fadd(c1, i1); c2 = fadd(c2, i2); c3 = fadd(c3, i3);   // :534
                            z = fadd(z, zi);                                                              // :535
                            --n_left;
                        }
                    }
                    zaddr += 4u*(uint32_t)(before - n_left);
                }
            }
        }
        STAT_CLOCK(t3);
        __syncthreads();
        STAT_CLOCK(t4);

        // ---------------- write the tile back ------------------------------------------------------
        // Straight from registers: thread t holds pixels t, t + NT, ... of the row-major tile, so consecutive
        // threads write consecutive columns of a target row (128-byte stores per warp).  The first version
        // packed depth and colour back into shared memory and issued two bulk copies per tile row; on frames
        // whose tiles carry little work that write-back was 37 % of a tile's time (19 k of 52 k cycles per
        // tile on a 500-triangle 4K frame, tools/raster_stats.py), the issue of 2 x rows small bulk copies
        // behind three barriers and a proxy fence.  The shared-memory route remains for the fused gather,
        // whose second copy of the tile goes to peer memory with bulk stores.
        if(B200R_DIRECT_WB && p.gather_color == nullptr)
        {
            Pixel px[PPT];
#pragma unroll
            for(int k = 0; k < PPT; ++k) px[k] = tile[tid + k*NT];
#pragma unroll
            for(int k = 0; k < PPT; ++k)
            {
                const int idx = tid + k*NT, r = idx / TW, c = idx % TW;
                if(r < rows && c < cols)
                {
                    p.depth[(size_t)(yb + r)*p.depth_stride + x0 + c] = __uint_as_float(px[k].z);
                    p.color[(size_t)(yb + r)*p.color_pitch_words + x0 + c] = px[k].color;
                }
            }
        }
        else
        {
            Pixel px[PPT];
#pragma unroll
            for(int k = 0; k < PPT; ++k) px[k] = tile[tid + k*NT];
            __syncthreads();
            uint32_t *cout_ = reinterpret_cast<uint32_t *>(smem_raw);
#pragma unroll
            for(int k = 0; k < PPT; ++k) { zplane[tid + k*NT] = __uint_as_float(px[k].z); cout_[tid + k*NT] = px[k].color; }
            __syncthreads();
            if(p.bulk_ok)
            {
                fence_proxy_async();
                __syncthreads();
                for(int r = tid; r < rows; r += NT)
                {
                    bulk_s2g(p.depth + (size_t)(yb + r)*p.depth_stride + x0, zplane + r*TW, (uint32_t)(cols*4));
                    bulk_s2g(p.color + (size_t)(yb + r)*p.color_pitch_words + x0, cout_ + r*TW, (uint32_t)(cols*4));
                    // fused gather: the same rows go straight into the (peer-mapped) gather target
                    if(p.gather_color && gather_bulk)
                    {
                        if(p.gather_depth) bulk_s2g(p.gather_depth + (size_t)(ys0 + r)*p.gather_depth_stride + x0, zplane + r*TW, (uint32_t)(cols*4));
                        bulk_s2g(p.gather_color + (size_t)(ys0 + r)*p.gather_pitch_words + x0, cout_ + r*TW, (uint32_t)(cols*4));
                    }
                }
                bulk_commit();
            }
            else
            {
                for(int i = tid; i < rows*cols; i += NT)
                {
                    const int r = i / cols, c = i % cols;
                    p.depth[(size_t)(yb + r)*p.depth_stride + x0 + c] = zplane[r*TW + c];
                    p.color[(size_t)(yb + r)*p.color_pitch_words + x0 + c] = cout_[r*TW + c];
                }
            }
            if(p.gather_color && !gather_bulk)
            {
                for(int i = tid; i < rows*cols; i += NT)
                {
                    const int r = i / cols, c = i % cols;
                    if(p.gather_depth) p.gather_depth[(size_t)(ys0 + r)*p.gather_depth_stride + x0 + c] = zplane[r*TW + c];
                    p.gather_color[(size_t)(ys0 + r)*p.gather_pitch_words + x0 + c] = cout_[r*TW + c];
                }
            }
        }
#if defined(B200R_STATS)
        {
            const long long t5 = clock64();
            if(lane == 0)
            {
                STAT_CLK_ADD(13, t1 - t0); STAT_CLK_ADD(14, t2 - t1); STAT_CLK_ADD(15, t3 - t2); STAT_CLK_ADD(16, t4 - t3); STAT_CLK_ADD(24, t5 - t4);
                STAT_CLK_ADD(25, 1);
            }
            if(tid == 0) STAT_ADD(17, 1);
        }
#endif
    }
    bulk_wait_read();
}

#if defined(B200R_STATS)
extern "C" int b200r_debug_raster_stats(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out, g_raster_stats, sizeof(g_raster_stats));
    if(e != cudaSuccess) return -2;
    if(reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g_raster_stats, z, sizeof(z)); }
    return 0;
}
#endif

// cudaFuncSetAttribute and the occupancy answer are per DEVICE: keep them per device ordinal
// (a process may hold contexts on several GPUs).
template<int TW, int TH, int WARPS, int MODE>
static cudaError_t launch_one(const RasterParams &p, int sm_count, cudaStream_t s)
{
    constexpr int kMaxDevices = 64;
    const int smem = TileLayout<TW, TH>::kBytes;
    auto kern = raster_kernel<TW, TH, WARPS, MODE>;
    static std::mutex mu;
    static int per_sm_of[kMaxDevices] = {};                // 0: not configured on that device yet
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if(e != cudaSuccess) return e;
    if(dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    int per_sm;
    {
        std::lock_guard<std::mutex> lock(mu);
        if(per_sm_of[dev] == 0)
        {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if(e != cudaSuccess) return e;
            int n = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, WARPS*32, smem);
            if(e != cudaSuccess) return e;
            per_sm_of[dev] = n < 1 ? 1 : n;
        }
        per_sm = per_sm_of[dev];
    }
    unsigned grid = (unsigned)(sm_count*per_sm);           // persistent: a multiple of the SM count
    if(grid > p.tile_end - p.tile_begin) grid = p.tile_end - p.tile_begin;
    if(grid < 1) grid = 1;
    kern<<<grid, WARPS*32, smem, s>>>(p);
    return cudaGetLastError();
}

template<int MODE>
static cudaError_t launch_mode(const RasterParams &p, int sm_count, cudaStream_t s)
{
    const int tw = p.v.tile_w, th = p.v.tile_h;
    if(tw == 64 && th == 32) return launch_one<64, 32, 8, MODE>(p, sm_count, s);     // 40 KB tile
    if(tw == 32 && th == 32) return launch_one<32, 32, 8, MODE>(p, sm_count, s);     // 20 KB
    if(tw == 128 && th == 16) return launch_one<128, 16, 8, MODE>(p, sm_count, s);   // 40 KB
    if(tw == 64 && th == 16) return launch_one<64, 16, 8, MODE>(p, sm_count, s);     // 20 KB
    if(tw == 128 && th == 32) return launch_one<128, 32, 8, MODE>(p, sm_count, s);   // 80 KB
    if(tw == 256 && th == 8) return launch_one<256, 8, 8, MODE>(p, sm_count, s);     // 40 KB
    if(tw == 128 && th == 8) return launch_one<128, 8, 4, MODE>(p, sm_count, s);     // 20 KB, 4 warps
    if(tw == 256 && th == 4) return launch_one<256, 4, 4, MODE>(p, sm_count, s);     // 20 KB, 4 warps
    return cudaErrorInvalidValue;
}

cudaError_t launch_raster(const RasterParams &p, int sm_count, cudaStream_t s)
{
    if(p.mode == kRasterGeneral) return launch_mode<kRasterGeneral>(p, sm_count, s);
    if(p.mode == kRasterTextured) return launch_mode<kRasterTextured>(p, sm_count, s);
    return launch_mode<kRasterPlain>(p, sm_count, s);
}

} // namespace b200r
