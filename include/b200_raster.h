/* b200_raster.h -- C ABI of the B200 rasterization back end (libb200raster.so).
 *
 * Drop-in boundary for the reference's per-object call pair:
 *     u32  FillEdgeTable(render_entry_3d_object*, game_render_commands*, b32)   projekt.cpp:3882
 *     void DrawModel(loaded_bitmap*, edge_info*, u32, game_render_commands*,
 *                    loaded_bitmap *Bitmap, b32 Phong)                          projekt.cpp:162
 * (and its thread-pool variants projekt.cpp:2350, 3362, 3615, which only redistribute the same
 * work).  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * The structs of projekt.h:2-37 are reproduced byte-for-byte.  The structs projekt.h *uses* but
 * does not define (loaded_bitmap, game_render_commands, light_data, light_info,
 * projective_transform, v2/v3/v4) are absent from the reference snapshot; their layouts are
 * pinned here and become part of this published header.  When this file
 * is included from inside the reference's own unity build, define B200R_NO_REFERENCE_TYPES
 * first so the renderer's own definitions are used.
 *
 * Semantics: Gouraud (PhongShading == 0) or per-pixel Phong (PhongShading != 0,
 * projekt.cpp:450-509) objects, untextured (Bitmap == 0) or textured with nearest-texel,
 * perspective-correct sampling (Bitmap != 0, projekt.cpp:427-446; texel coordinates that leave the
 * bitmap are clamped where the reference reads outside it).  One triangle = one object ("level 1",
 * SURVEY.md section 0) unless B200R_WHOLE_OBJECT_AEL asks for the reference's whole-object
 * active-edge list.  Coverage and depth are bit-exact with the reference's scalar arithmetic,
 * Gouraud and unlit textured colour bit-exact, Phong colour within +-1 LSB per channel (pow(x,16)
 * in double); equal depth is resolved as in the reference (first submitted wins, projekt.cpp:525).
 *
 * Internal lists (segments, spans, per-tile queues) are sized from earlier frames and grow on
 * demand: a frame whose lists did not fit is not drawn at all (one device-side verdict word makes
 * its binning and raster kernels return at once), the lists are grown and the frame is issued again
 * (b200r_frame_stats::Reruns).  Every render call resolves that verdict before it returns, so
 * work enqueued on the same stream afterwards sees the finished frame; B200R_DEFER_VERDICT opts out.
 *
 * Every entry point makes the context's device current for the duration of the call and restores
 * the calling thread's current device before it returns.
 *
 * There is no CPU fallback: every entry point fails with B200R_E_NO_DEVICE when no sm_100 device
 * is usable.
 */
#ifndef B200_RASTER_H
#define B200_RASTER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef B200R_NO_REFERENCE_TYPES
typedef uint8_t u8;
typedef uint32_t u32;
typedef int32_t s32;
typedef float r32;
typedef int32_t b32;

typedef struct v2 { r32 x, y; } v2;
typedef struct v3 { r32 x, y, z; } v3;
typedef struct v4 { r32 x, y, z, w; } v4;          /* colours: x=r y=g z=b w=a */

typedef struct loaded_bitmap               /* projekt.cpp:387, 414-416 */
{
    s32 Width;
    s32 Height;
    s32 Pitch;                             /* bytes per row */
    void *Memory;                          /* u32 ARGB8 */
} loaded_bitmap;

typedef struct projective_transform        /* projekt.cpp:79-89, 152-155 */
{
    r32 MetersToPixels;
    v2 ScreenCenter;
    r32 FocalLength;
    r32 DistanceAboveTarget;
} projective_transform;

typedef struct light_info { v3 P; v4 Intensity; } light_info;      /* projekt.cpp:4026-4027 */

typedef struct light_data                  /* projekt.cpp:3885, 3892, 4010, 4023 */
{
    v4 AmbientIntensity;
    u32 LightCount;
    light_info *Lights;
} light_data;

typedef struct game_render_commands        /* projekt.cpp:170-171, 452, 1017, 2325-2331, 4117 */
{
    u32 Width;                             /* depth row stride, in floats */
    r32 *ZBuffer;
    u8 *ZMask;                             /* CPU spin-lock bytes; unused on the GPU path */
    light_data LightData;
    projective_transform Transform;
    void *ThreadMemory;                    /* CPU work arena; unused on the GPU path */
    u32 ThreadMemorySize;
    u32 ThreadMemorySizeUsed;
    void *SortMemory;                      /* MergeSort scratch; unused on the GPU path */
} game_render_commands;

typedef struct render_entry_3d_object      /* projekt.h:2-15 */
{
    v3 P;
    u32 VertexCount;
    b32 Optimized;
    b32 PhongShading;
    void *VertexData;                      /* v3[VertexCount], 3 per triangle */
    void *ColorData;                       /* v4[VertexCount] */
    void *NormalData;                      /* v3[VertexCount] */
    void *UVData;                          /* v2[VertexCount] */
    void *EdgeMemory;                      /* edge_info[VertexCount] */
    loaded_bitmap *Bitmap;
} render_entry_3d_object;

typedef struct edge_info                   /* projekt.h:17-37, 120 bytes */
{
    s32 YMax;
    r32 XMin;
    r32 ZMin;
    r32 OneOverZMin;
    r32 Gradient;
    r32 ZGradient;
    r32 OneOverZGradient;
    s32 YMin;
    r32 UMin;
    r32 VMin;
    r32 UGradient;
    r32 VGradient;
    b32 Left;
    v4 MinColor;
    v4 ColorGradient;
    v3 MinNormal;
    v3 NormalGradient;
    struct edge_info *Next;
} edge_info;
#endif /* B200R_NO_REFERENCE_TYPES */

/* ---- status codes: the reference has only Assert (projekt.cpp:25, 2327); we never abort --- */
#define B200R_OK             0
#define B200R_E_INVALID     (-1)   /* null / inconsistent arguments                           */
#define B200R_E_CUDA        (-2)   /* a CUDA call failed; see b200r_last_error                 */
#define B200R_E_UNSUPPORTED (-3)   /* > 8 lights, LightCount == 0 on the Gouraud path, > 65535 tiles per axis,
                                      > 65536 textures or (whole-object mode) > 65535 objects per call    */
#define B200R_E_NOMEM       (-4)
#define B200R_E_NO_DEVICE   (-5)   /* no CUDA device of compute capability 10.x                */

#define B200R_MAX_LIGHTS 8

/* flags of the render calls */
#define B200R_WHOLE_OBJECT_AEL 1u  /* b200r_render_objects: reproduce DrawModel's whole-object active-edge list
                                      (projekt.cpp:198-303, 542-597) link by link: consecutive entries are
                                      paired whatever triangle they belong to, as in the reference's own
                                      images of multi-triangle objects.  A compatibility mode: objects run in
                                      parallel, the rows of one object do not.  Where the reference dereferences
                                      a null list pointer the object stops drawing (b200r_frame_stats::
                                      StoppedObjects).  b200r_render_device: B200R_E_UNSUPPORTED.          */

#define B200R_DEFER_VERDICT 2u     /* b200r_render_device: return without waiting for the binning verdict of this
                                      frame.  The frame is then final only after the NEXT call on the context or
                                      b200r_sync: if a list overflowed, the frame's raster kernel did nothing and
                                      the frame is re-issued at that later point, reading the mesh and target
                                      pointers again.  For pipelines that re-submit frames of known size.       */

/* Compatibility switches, for hosts that compare against images of the reference's AVX
 * fillers.  Those differ from the scalar path in more than can be switched (lane-wise start + k*inc
 * arithmetic, truncated texel coordinates -- SURVEY.md section 0); these are the two rules that can be stated
 * on top of the scalar arithmetic.  Per-triangle mode only (with B200R_WHOLE_OBJECT_AEL: B200R_E_UNSUPPORTED). */
#define B200R_AVX_RIGHT_END_EXCLUSIVE 4u   /* a span covers [MinX, MaxX) instead of [MinX, MaxX]: the end-clip masks of
                                              FillLinesOptimized (projekt.cpp:782-794); no pixel is ever written at
                                              column == Width                                                          */
#define B200R_AVX_DEPTH_GE 8u              /* the depth test is CurrentZ >= *Z (DrawModelOptimized, projekt.cpp:3205)
                                              instead of > (:525): equal depth goes to the LAST submitted fragment, and
                                              pre-existing target contents lose ties                                    */

typedef struct b200r_context b200r_context;

/* One context per GPU.  Device < 0 keeps the calling thread's current device. */
int b200r_create(b200r_context **Context, int Device);
void b200r_destroy(b200r_context *Context);
const char *b200r_last_error(const b200r_context *Context);

/* Launch on an existing CUDA stream (a cudaStream_t cast to void*); 0 = the context's own. */
int b200r_set_stream(b200r_context *Context, void *CudaStream);
int b200r_sync(b200r_context *Context);

/* Screen tile staged in shared memory by the raster kernel: 64x32, 32x32, 128x16, 64x16, 128x32,
 * 256x8, 128x8 or 256x4 pixels.  0x0 (the default) lets every render call choose: 64x16 for frames of
 * small triangles (fewer than 4 target pixels per submitted triangle), 128x8 otherwise.  Targets wider
 * than 262143 pixels are not supported. */
int b200r_set_tile(b200r_context *Context, int TileWidth, int TileHeight);

/* ------------------------------------------------------------------------------------------
 * Host-pointer drop-in: replaces, for a batch of objects, the pair
 *   FillEdgeTable(Object, Commands, 0); DrawModel(OutputTarget, Object->EdgeMemory, n, Commands)
 * (projekt.cpp:3882 + 162).  Colour (OutputTarget->Memory) and depth (Commands->ZBuffer) are
 * read from and written back to the caller's host buffers, so pre-existing contents take part
 * in the depth test exactly as in the reference (projekt.cpp:525); nothing is cleared.
 * EdgeMemory / SortMemory / ThreadMemory / ZMask may be null.  Blocking.
 * ------------------------------------------------------------------------------------------ */
int b200r_render_objects(b200r_context *Context, const render_entry_3d_object *Objects,
                         u32 ObjectCount, const game_render_commands *Commands,
                         const loaded_bitmap *OutputTarget, u32 Flags);

/* Replaces FillEdgeTable alone (projekt.cpp:3882): sorted edge_info records are written to
 * Object->EdgeMemory (host, room for VertexCount records) in the reference's MergeSort order
 * (projekt.cpp:2-72, ties included; the order is produced on the device).  Every record is
 * written whole: the fields the selected path defines (YMin YMax XMin Gradient ZMin ZGradient
 * MinColor ColorGradient Left; with PhongShading also MinNormal and NormalGradient; with
 * Object->Bitmap also UMin VMin OneOverZMin and their gradients, and Gouraud colours are lit white
 * as in projekt.cpp:4034-4060), zero in the fields the reference leaves untouched for that kind
 * of object, Next = 0.  At most 16 M triangles per object.
 * Returns the edge count (>= 0) or a negative status. */
int b200r_fill_edge_table(b200r_context *Context, const render_entry_3d_object *Object,
                          const game_render_commands *Commands, b32 PhongShading);

/* ------------------------------------------------------------------------------------------
 * Device-resident path (what the host-pointer call is built from).  All pointers below are
 * device pointers; the kernels run asynchronously on the context's stream.  b200r_render_device
 * returns once the frame's binning verdict is known (its set-up and scan kernels have finished;
 * the raster kernel is still running) -- see "Internal lists" above and B200R_DEFER_VERDICT.
 * ------------------------------------------------------------------------------------------ */
/* render_entry_3d_object::Bitmap (projekt.h:13) resident on the device: nearest-texel, perspective
 * correct sampling at Round(uv * (dim - 1)) (projekt.cpp:427-446).  The reference does not
 * range-check texel coordinates and reads outside the bitmap when they leave it; here they are
 * clamped to the bitmap (NaN samples texel 0,0). */
typedef struct b200r_device_texture
{
    const u32 *Memory;         /* device pointer, ARGB8 (loaded_bitmap::Memory) */
    s32 Width, Height;
    s32 Pitch;                 /* bytes per row, a multiple of 4                */
} b200r_device_texture;

typedef struct b200r_device_mesh
{
    const r32 *Positions;      /* v3 per vertex (VertexData)  */
    const r32 *Colors;         /* v4 per vertex (ColorData)   */
    const r32 *Normals;        /* v3 per vertex (NormalData)  */
    u32 TriangleCount;         /* VertexCount / 3             */
    v3 P;                      /* render_entry_3d_object::P   */
    u32 Flags;                 /* B200R_MESH_PHONG = render_entry_3d_object::PhongShading        */
    const r32 *UVs;            /* v2 per vertex (UVData); read only when Texture != 0            */
    const b200r_device_texture *Texture;   /* host struct describing device memory; 0: untextured */
} b200r_device_mesh;
#define B200R_MESH_PHONG 1u

typedef struct b200r_device_target
{
    u32 *Color;                /* first row of the band, ARGB8                                  */
    r32 *Depth;
    s32 Width, Height;         /* the logical screen (loaded_bitmap Width/Height)              */
    s32 ColorPitch;            /* bytes per row                                                 */
    s32 DepthStride;           /* floats per row (game_render_commands::Width)                  */
    s32 BandFirstRow;          /* this target holds screen rows [BandFirstRow, +BandRows)       */
    s32 BandRows;              /* = Height for a whole frame                                    */
} b200r_device_target;

/* Lights and transform are taken from Commands (host struct; ZBuffer etc. ignored). */
int b200r_render_device(b200r_context *Context, const b200r_device_mesh *Meshes, u32 MeshCount,
                        const game_render_commands *Commands, const b200r_device_target *Target,
                        u32 Flags);

/* Fill a device target band with a clear colour / depth (the reference never clears,
 * SURVEY.md 8b "Persistence"; callers do). */
int b200r_clear_device(b200r_context *Context, const b200r_device_target *Target, u32 Color,
                       r32 Depth);

/* ------------------------------------------------------------------------------------------
 * Fused gather for multi-GPU frames (the path's only exchange step is the gather
 * of the finished band / frame images).  With a gather target set, the raster kernel stores every
 * tile of the band it renders a second time: into GatherTarget, an image of the WHOLE screen
 * (Width x Height as in the render target; BandFirstRow / BandRows ignored; Depth may be 0) --
 * typically memory of the GPU that assembles the frame, mapped into this process with
 * b200r_peer_open, so that the band travels over NVLink tile by tile while the kernel is still
 * rasterising and no separate gather pass follows.  Tiles nothing was drawn into are copied
 * through, so after the frame the band's rows of GatherTarget equal the band exactly.  The writes
 * are complete when the frame's kernels are (stream order on the rendering GPU); the assembling
 * rank must order its reads after that (e.g. a stream-ordered barrier across the ranks).
 * Applies to b200r_render_device only.  0 switches it off.
 * ------------------------------------------------------------------------------------------ */
int b200r_set_gather_target(b200r_context *Context, const b200r_device_target *GatherTarget);

/* Device memory other processes on the same machine can map (CUDA IPC).  b200r_peer_alloc
 * allocates Bytes on the context's device and returns the handle to send to the peers (any byte
 * transport); b200r_peer_open maps a peer's allocation into this process (peer access over
 * NVLink / PCIe is enabled on first use); b200r_peer_release unmaps / frees either kind.
 * Whatever is left is released by b200r_destroy. */
typedef struct b200r_peer_handle { unsigned char Bytes[64]; } b200r_peer_handle;
int b200r_peer_alloc(b200r_context *Context, uint64_t Bytes, void **DevicePointer, b200r_peer_handle *Handle);
int b200r_peer_open(b200r_context *Context, const b200r_peer_handle *Handle, void **DevicePointer);
int b200r_peer_release(b200r_context *Context, void *DevicePointer);

typedef struct b200r_frame_stats
{
    uint64_t Triangles;        /* submitted in the last render call                             */
    uint64_t Binned;           /* triangles that produced at least one (segment, tile) pair     */
    uint64_t Segments;         /* segments (edge pair x tile-row band) emitted by the set-up kernel */
    uint64_t Spans;            /* span records (one per covered row) emitted by the set-up kernel */
    uint64_t AliasPixels;      /* pixels the reference writes at column == Width, i.e. into column 0
                                  of the next row of a contiguous target (projekt.cpp:402-419)    */
    uint64_t TilePairs;        /* (span, tile) queue entries produced by the binner             */
    uint64_t Tiles;            /* screen tiles of the band                                      */
    uint64_t KernelLaunches;   /* kernels launched by this context since creation               */
    uint64_t Reruns;           /* frames re-issued because a segment, span or queue list had to grow */
    uint64_t StoppedObjects;   /* B200R_WHOLE_OBJECT_AEL: objects that stopped drawing where the
                                  reference dereferences a null list pointer (it crashes there)  */
} b200r_frame_stats;

/* Valid after b200r_sync (or any blocking call). */
int b200r_get_stats(b200r_context *Context, b200r_frame_stats *Stats);

/* Per-kernel timing of the device path with CUDA events recorded on the launching stream
 * (the reference has no timers at all).  While enabled, every frame
 * records an event between stages; b200r_get_stage_ms syncs and returns the durations of the
 * last frame: [0] setup_kernel (with zrange_kernel and, on partial bands, select_kernel; in
 * whole-object mode the chain / order / emit kernels), [1] tile_scan_kernel + finalize_kernel,
 * [2] scatter_kernel, [3] raster_kernel. */
#define B200R_STAGES 4
int b200r_set_profiling(b200r_context *Context, int Enable);
int b200r_get_stage_ms(b200r_context *Context, float StageMs[B200R_STAGES]);

#ifdef __cplusplus
}
#endif
#endif /* B200_RASTER_H */
