// Shared device-side definitions for the sm_100a rasterization kernels.
//
// Arithmetic contract (SURVEY.md section 7 "hard parts"): every value that reaches a
// coverage, depth or colour decision is produced by the same sequence of IEEE-754 binary32
// operations as the reference's scalar x86 code (projekt.cpp:74-93, 162-601, 3882-4121),
// one rounding per operation.  All such arithmetic goes through the __f*_rn intrinsics below
// (never contracted to FMA, never reordered, independent of compiler flags); the translation
// units are additionally built with -fmad=false and the default -prec-div/-prec-sqrt/-ftz.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// -DB200R_CHECKED builds the kernels with bounds checks on every list index and shared-memory address they
// compute (device assert: the kernel traps, the next API call returns B200R_E_CUDA).  compute-sanitizer is
// not available on every pool; tools/build_variant.sh checked -DB200R_CHECKED + the GPU test suite under
// B200R_LIB is the substitute (profiles/r02_checked_build.txt).
#if defined(B200R_CHECKED)
#include <assert.h>
#define B200R_ASSERT(x) assert(x)
#else
#define B200R_ASSERT(x) ((void)0)
#endif

namespace b200r {

constexpr int kMaxLights = 8;
// Every tile's span queue is split into kDepthBuckets sub-queues by the camera-space depth of the
// owning triangle, nearest first.  Processing order never changes the image (the depth rule is
// order independent); near-to-far order only makes the cheap early depth test fail more often.
#ifndef B200R_BUCKETS
#define B200R_BUCKETS 8
#endif
constexpr int kDepthBuckets = B200R_BUCKETS;
// The span and segment arrays are carved into kSubAllocators equal regions, each with its own
// fill counter; CTA b of the set-up kernel allocates from region b % kSubAllocators.  One global
// counter pair was the set-up kernel's bottleneck: every CTA does one returning atomicAdd on it
// and waits for the answer at a barrier (C4: 156 k CTAs on two addresses).
constexpr int kSubAllocators = 64;

struct DevLight { float px, py, pz; float ir, ig, ib, ia; };

// Per-call view state: projective_transform + light_data (projekt.cpp:79-89, 3885-3892) and
// the screen / band / tile geometry.
struct ViewParams
{
    float m2p, cx, cy, focal, dist;
    float amb[4];
    int nlights;
    DevLight lights[kMaxLights];
    int width, height;          // logical screen (loaded_bitmap Width/Height)
    int band_y0, band_y1;       // screen rows owned by this target
    int tile_w, tile_h;         // powers of two
    int tile_w_shift, tile_h_shift;
    int tiles_x, tiles_y;
    int right_end_exclusive;    // B200R_AVX_RIGHT_END_EXCLUSIVE: spans cover [MinX, MaxX)
    int depth_ge;               // B200R_AVX_DEPTH_GE: equal depth goes to the LAST submitted fragment
    int alias_rows;             // rows are contiguous (Pitch == Width*4, depth stride == Width): a pixel the
                                // reference writes at column == Width lands in column 0 of the next row
};

struct MeshParams
{
    const float *pos;           // v3 x 3 per triangle
    const float *col;           // v4 x 3 per triangle
    const float *nrm;           // v3 x 3 per triangle
    unsigned ntri;
    float px, py, pz;           // render_entry_3d_object::P
    unsigned prim_base;         // submission index of this mesh's first triangle
    int phong;                  // render_entry_3d_object::PhongShading
    // Textured mesh (render_entry_3d_object::Bitmap != 0).  Every pixel of such a mesh takes its
    // colour from the texel (projekt.cpp:427-446), so the interpolated vertex colours never reach the
    // image, and u/z, v/z, 1/z are interpolated by exactly the operations the colour channels use
    // (edge step += gradient :554-560, span increment (R-L)/XDifference :336-342, start += XOffset*inc
    // :409-410): in tex mode the four colour interpolants CARRY (u/z, v/z, 1/z, 0).
    // Row-band pre-selection (multi-GPU bands): when set, the set-up kernel processes only the
    // triangles listed here (their ORIGINAL indices, in any order) -- the ones select_kernel found
    // able to reach this GPU's band.  Owners stay prim_base + original index.
    const unsigned *tri_list;   // device, or null: all ntri triangles in order
    const unsigned *tri_count;  // device: number of listed triangles
    // Row-parallel walks (setup_kernel<..., SPLIT>): a CTA stages tris_per_cta (8..64) triangles and cuts
    // tall ones into walkers of part_rows screen rows (a multiple of the tile height); 0: off.
    int tris_per_cta, part_rows;
    // Frames of tall triangles (the host decides from target pixels per triangle): walkers are sorted by
    // height >> sort_shift, and the lock-step row loop runs at most row_chunk rows between two looks at
    // the lanes' list events (INT_MAX: up to the next event, right for triangles of a few rows).
    int sort_shift, row_chunk;
    const float *uv;            // v2 x 3 per triangle, or null
    int tex;                    // index into the frame's texture table, -1: untextured
    int white;                  // b200r_fill_edge_table of a textured Gouraud object: light white vertices (:4034-4060)
};

struct TexDesc { const uint32_t *mem; int w, h, pitch; };   // loaded_bitmap of a texture; pitch in bytes

// Compact per-triangle record written by the setup kernel and read by the raster kernel:
// the fields of edge_info (projekt.h:17-37) that the Gouraud path defines, for the <= 3 edges
// of one triangle in the reference's MergeSort order (projekt.cpp:2-72).
constexpr int kEdgeWords = 15;      // ymin ymax x dx z dz c[4] dc[4] left
constexpr int kRecWords = 52;       // 4 header + 3*15 + 3 pad  = 208 bytes = 13 float4
constexpr int kRecVec4 = kRecWords/4;
enum { E_YMIN = 0, E_YMAX = 1, E_X = 2, E_DX = 3, E_Z = 4, E_DZ = 5, E_C = 6, E_DC = 10, E_LEFT = 14 };
// E_LEFT word: bit 0 = edge_info::Left; bits 8-9 / 16-17 = index (0..2) of the edge's upper / lower
// vertex in the triangle (Phong: normals are re-read from the staged vertex normals)
enum { R_NEDGES = 0, R_FIRSTROW = 1, R_MAXY = 2, R_PRIM = 3, R_EDGE0 = 4 };

// Span record: one row of one triangle, fully set up (projekt.cpp:306-412 done once, in the
// set-up kernel): the inclusive column range, the values at the first column and the per-pixel
// increments.  The raster kernel only replays the per-pixel adds and depth-tests.
constexpr int kSpanWords = 16;      // Gouraud frame: 64 bytes = 4 float4
constexpr int kSpanWordsPhong = 24; // frame with a Phong mesh: + normal at the first column, per-pixel normal increment
constexpr int kSpanVec4 = kSpanWords/4;
enum { P_PRIM = 0, P_Y = 1, P_MINX = 2, P_MAXX = 3, P_Z = 4, P_C = 5, P_ZI = 9, P_CI = 10, P_FLAGS = 14, P_ZUB = 15 };
// P_ZUB: an upper bound of every depth value the span can produce (see span_depth_bound)
constexpr unsigned kSpanNonFinite = 1u;   // colours may be NaN/Inf/huge -> guarded pack
constexpr unsigned kSpanAlias = 4u;       // Phong alias pixel: shade at X = word 19, Row = word 20 (not at its own column/row)
constexpr unsigned kSpanTex = 8u;         // textured: colour words hold u/z, v/z, 1/z; texture index in bits 8..23 of P_FLAGS
constexpr unsigned kSpanPhong = 2u;       // per-pixel Phong shading (projekt.cpp:450-509): words 16..21 hold the normals

// Segment: the consecutive spans of one triangle that share one pair of active edges and lie in
// one tile-row band; the unit the binner scatters (its spans are contiguous in the span array).
struct SegInfo
{
    unsigned tile_row;          // band-relative tile row | depth bucket << 24
    unsigned tx;                // tx0 | tx1 << 16 (tx0 > tx1: touches no tile)
    unsigned span_base;         // index of its first span record
    unsigned nrows;             // number of span records
};

enum { kRasterPlain = 0, kRasterGeneral = 1, kRasterTextured = 2 };

struct RasterParams
{
    ViewParams v;
    const uint32_t *spans;      // span_words per span
    int span_words;             // kSpanWords, or kSpanWordsPhong when any mesh of the frame is Phong
    const unsigned *overflow;   // device word set by finalize_kernel: some list did not fit, skip the frame
    unsigned seg_capacity, span_capacity;
    const unsigned *tile_offset;    // [tile*kDepthBuckets + bucket], plus one end entry
    const unsigned *pair_list;
    const unsigned *pair_total; // device word: total (triangle,tile) pairs this frame
    unsigned pair_capacity;
    unsigned *work_counter;
    unsigned ntiles;
    unsigned tile_begin, tile_end;  // this launch renders tiles [tile_begin, tile_end) (row-major: a band of tile rows)
    uint32_t *color;            // band rows
    float *depth;
    int color_pitch_words;      // u32 per row
    int depth_stride;           // floats per row
    int bulk_ok;                // rows may be moved with cp.async.bulk (16-byte aligned)
    // Fused gather (b200r_set_gather_target): every tile of the band is ALSO stored into this image of the
    // whole screen, typically peer memory of the GPU that assembles the frame (NVLink); null: off.
    uint32_t *gather_color;     // row 0 of the whole screen
    float *gather_depth;        // optional
    int gather_pitch_words, gather_depth_stride;
    int gather_bulk_ok;
    const TexDesc *textures;    // device: the frame's texture table (null without textured meshes)
    unsigned texture_count;
    int mode;                   // kRasterPlain / kRasterGeneral / kRasterTextured (raster_kernel.cu)
    int refill_lanes;           // idle lanes of a warp that trigger a refill from the span queue
    int pend_lanes;             // parked lanes of a warp that trigger the depth-pass path
};

// ---------------------------------------------------------------- exact binary32 helpers
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// The same quotient, bit for bit, with the two operand classes that are ROUTINE on this path
// answered by selects instead of by div.rn's slow-path subroutine (its range check rejects zero
// operands, and the subroutine serialises the warp; measured on C2: 12 % of the set-up kernel's
// instructions at 10 active lanes):
//   0/b  a constant colour channel or a flat normal gives zero gradients (projekt.cpp:333-349, :4096)
//   a/0  an edge that lies inside one pixel row has YDifference == 0        (projekt.cpp:4070-4096)
// IEEE 754: 0/b = +-0 and a/0 = +-inf, sign = sign(a) xor sign(b).  0/0 and NaN operands are left
// to the real divide.
__device__ __forceinline__ float fdiv_zq(float a, float b)
{
    const bool az = (a == 0.0f), bz = (b == 0.0f);
    const bool special = (az != bz) && (a == a) && (b == b);
    const float q = __fdiv_rn(special ? 1.0f : a, special ? 1.0f : b);
    const uint32_t sign = (__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u;
    return special ? __uint_as_float(sign | (bz ? 0x7f800000u : 0u)) : q;
}
// b is known to be non-zero (span increments are only computed for XDifference != 0, :333)
__device__ __forceinline__ float fdiv_zn(float a, float b)
{
    const bool special = (a == 0.0f) && (b == b);
    const float q = __fdiv_rn(special ? b : a, b);
    const uint32_t sign = (__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u;
    return special ? __uint_as_float(sign) : q;
}

// RoundR32ToS32 = cvtss2si (projekt.cpp:402, 3988): nearest-even; NaN / out of range -> INT_MIN.
__device__ __forceinline__ int round_s32(float v)
{
    return (fabsf(v) < 2147483648.0f) ? __float2int_rn(v) : (int)0x80000000;
}
// Clamp01 (projekt.cpp:474, 4047): comparisons, so NaN passes through as on the host.
__device__ __forceinline__ float clamp01(float v)
{
    if(v < 0.0f) v = 0.0f; else if(v > 1.0f) v = 1.0f;
    return v;
}

// ---------------------------------------------------------------- TMA bulk copy + mbarrier
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while(!mbar_try_wait(bar, parity)) { }
}
// global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// shared -> global, tracked by the bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------- launchers (one per kernel)
struct SetupOutputs
{
    uint32_t *recs;             // optional (b200r_fill_edge_table): kRecWords per triangle
    uint32_t *spans;            // span_words per span (null: no row walk, records only)
    int span_words;
    SegInfo *segs;
    unsigned *seg_fill;         // [kSubAllocators] device counters, one per region of the arrays
    unsigned *span_fill;        // [kSubAllocators]
    unsigned *extra_total;      // alias pixels: one-pixel span + segment each, allocated downwards
                                // from the END of the span / segment arrays (i.e. of the last region)
    unsigned seg_capacity, span_capacity;   // whole arrays; a region holds capacity / kSubAllocators
    unsigned *tile_count;       // [tile*kDepthBuckets + bucket]
    const float *zrange;        // device: {largest camera z, 1/(largest - smallest)} of this frame's vertices
    unsigned long long *counters;   // [0] binned triangles, [1] tile pairs
};

void launch_setup(const ViewParams &v, const MeshParams &m, const SetupOutputs &out, cudaStream_t s);
// positions only: indices of the triangles whose projected rows can reach the band -> list, *count; the same
// pass folds the mesh's camera z into zkeys (replaces launch_zrange for that mesh)
void launch_select(const ViewParams &v, const MeshParams &m, unsigned *list, unsigned *count, unsigned *zkeys, int sm_count, cudaStream_t s);

// ---- FillEdgeTable's tail on the device (edge_table_kernels.cu): emission order, MergeSort order, edge_info ----
void launch_edge_counts(const uint32_t *recs, unsigned ntri, unsigned *counts, cudaStream_t s);
void launch_edge_keys(const uint32_t *recs, unsigned ntri, const unsigned *offsets, const unsigned *total,
                      unsigned long long *keys, unsigned *vals, cudaStream_t s);
int launch_edge_sort(unsigned long long *keys[2], unsigned *vals[2], unsigned n, const unsigned *total, cudaStream_t s);
void launch_edge_assemble(const uint32_t *recs, const uint32_t *uvrecs, const float *normals, const unsigned *vals,
                          unsigned n, const unsigned *total, void *out, cudaStream_t s);

// ---- whole-object mode (object_walk_kernel.cu) ----
struct ObjectDesc
{
    unsigned first_edge, edge_count;    // the object's sorted edge_info records
    unsigned span_base, span_bound;     // its promised span / segment slots (frame-wide draw-order index of the first one)
    unsigned prim_base;                 // owner of its first span
    int phong;                          // render_entry_3d_object::PhongShading
    int tex;                            // texture table index, -1: untextured
    int max_y;                          // min(largest YMax, Height): one past the object's last row (projekt.cpp:187-196)
    unsigned chain_first, chain_total;  // its slice of the value chains (three-phase path)
};
struct ObjectWalkParams
{
    const void *edges;                  // the objects' sorted edge_info arrays (read only)
    float *state_scratch;               // 10 words per edge: mutable walk state of objects too large for shared memory
    unsigned state_smem_bytes;          // dynamic shared memory per CTA (the largest object that fits, at most 200 KB)
    const ObjectDesc *objects;
    unsigned nobjects;
    uint32_t *spans; int span_words;
    SegInfo *segs;
    unsigned *extra_total;
    unsigned seg_capacity, span_capacity;   // equal in this mode: one slot index addresses both arrays
    unsigned region_size;               // capacity / kSubAllocators
    unsigned *tile_count;
    unsigned long long *counters;
    unsigned *stopped;                  // objects that stopped where the reference dereferences null
    // ---- three-phase path: value chains per edge, list order per object, span set-up per pair ----
    const unsigned *chain_base;         // per edge: frame-wide index of its chain (rows + 1 entries); null: serial walk only
    float *chains;                      // structure of arrays over chain_T entries: x | z | c[4] | n[3]
    unsigned chain_T;
    uint4 *pair_list;                   // per promised slot: chain index left, chain index right, row
    unsigned *produced;                 // per object: pairs the order phase emitted
    unsigned *fallback;                 // per object: 1 = a step ran past its chain, the serial walk redoes the object
    unsigned order_smem_bytes;
    unsigned max_bound;                 // largest span_bound (grid of the emit phase)
    unsigned max_edges;                 // largest edge_count (grid of the chain phase)
};
cudaError_t launch_object_walk(const ViewParams &v, const ObjectWalkParams &p, cudaStream_t s);
// zrange[0..1] start as {-inf as ordered key, +inf as ordered key}; zrange_finish turns them into
// {zmax, 1/(zmax - zmin)}
void launch_zrange(const MeshParams &m, unsigned *zkeys, cudaStream_t s);
void launch_zrange_finish(unsigned *zkeys, cudaStream_t s);
constexpr unsigned kScanMaxChunks = 2048;   // look-back scan: 8192 bins per chunk -> 16 M bins per frame
void launch_tile_scan(const unsigned *tile_count, unsigned *tile_offset, unsigned ntiles,
                      unsigned *pair_total, unsigned long long *state, unsigned *ticket, cudaStream_t s);
struct ScatterParams
{
    const SegInfo *segs;
    const unsigned *seg_fill;   // [kSubAllocators]
    const unsigned *extra_total;
    const unsigned *overflow;
    unsigned seg_capacity;
    int tiles_x;
    const unsigned *tile_offset;
    unsigned *tile_fill;
    unsigned *pair_list;        // per tile: span indices
    unsigned pair_capacity;
};
void launch_scatter(const ScatterParams &p, cudaStream_t s);
// After the scan: one word that tells scatter and raster whether every list fitted.
struct FinalizeParams
{
    const unsigned *seg_fill, *span_fill, *extra_total, *pair_total;
    unsigned seg_capacity, span_capacity, pair_capacity;
    unsigned *overflow;
    unsigned *seg_max, *span_max;   // largest region fill, for the host's growth decision
};
void launch_finalize(const FinalizeParams &p, cudaStream_t s);
cudaError_t launch_raster(const RasterParams &p, int sm_count, cudaStream_t s);
void launch_clear(uint32_t *color, int color_pitch_words, float *depth, int depth_stride,
                  int width, int rows, uint32_t cval, float dval, cudaStream_t s);

} // namespace b200r
