/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's scalar
 * rasterization path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (cpu_renderer_b200/) never does.
 *
 * What it restates (file:line into /root/reference):
 *   orc_project_vertex      projekt.cpp:74-93     ProjectVertex
 *   orc_fill_edge_table     projekt.cpp:3882-4121 FillEdgeTable, Gouraud branch (4020-4064)
 *   orc_merge_sort          projekt.cpp:2-72      MergeSort (tie order: right half first)
 *   orc_draw_triangle       projekt.cpp:198-303, 542-597 active-edge walk, "level 1":
 *                           one triangle = one object, defined behaviour where the reference
 *                           dereferences a null list pointer
 *   orc_fill_span           projekt.cpp:306-425, 510-538 span set-up + Gouraud pixel loop;
 *                           450-509 per-pixel Phong loop (UnprojectVertex 147-160)
 *
 * Pinning: the missing math layer is pinned by oracle/ref_shim.h, so
 * with respect to upstream this path is PARITY UNPINNED (the reference ships no tests, golden
 * vectors or math library).  With respect to the reference text that *is* in the snapshot the
 * oracle is pinned: tests/test_oracle_vs_ref.py checks it bit-for-bit against the verbatim
 * reference functions compiled into oracle/_ref/libprojekt_ref.so, and tests/golden/ holds
 * vectors generated from that verbatim build (tests/golden/make_golden.py).
 *
 * Build flags are part of the definition: -O2 -ffp-contract=off, no -mfma, no -ffast-math. */
#ifndef B200R_RASTER_ORACLE_H
#define B200R_RASTER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_transform {          /* projective_transform, projekt.cpp:79-89 */
    float MetersToPixels;
    float ScreenCenterX, ScreenCenterY;
    float FocalLength;
    float DistanceAboveTarget;
} orc_transform;

typedef struct orc_light { float P[3]; float Intensity[4]; } orc_light;   /* light_info */

typedef struct orc_scene {
    float Ambient[4];                   /* light_data.AmbientIntensity, projekt.cpp:3892 */
    uint32_t LightCount;                /* must be >= 1: with 0 lights the reference leaves
                                           MinColor uninitialised (projekt.cpp:4022-4045)   */
    const orc_light *Lights;
    orc_transform Transform;
} orc_scene;

/* The fields of edge_info (projekt.h:17-37) that the Gouraud path defines and reads. */
typedef struct orc_edge {
    int32_t YMin, YMax;
    float XMin, Gradient;
    float ZMin, ZGradient;
    float MinColor[4];                  /* r g b a */
    float ColorGradient[4];
    int32_t Left;
    int32_t Triangle;                   /* provenance only (not in the reference) */
    float MinNormal[3];                 /* Phong only (projekt.cpp:4017, 4104-4109) */
    float NormalGradient[3];
    float UMin, VMin, OneOverZMin;      /* textured only (projekt.cpp:4002-4004, 4078-4089): u/z, v/z, 1/z */
    float UGradient, VGradient, OneOverZGradient;
} orc_edge;

/* loaded_bitmap as a texture (projekt.cpp:427-446): ARGB8 texels, Pitch in bytes. */
typedef struct orc_texture {
    int32_t Width, Height, Pitch;
    const uint32_t *Memory;
} orc_texture;

typedef struct orc_target {
    int32_t Width, Height;              /* loaded_bitmap.Width/Height, projekt.cpp:193, 387 */
    int32_t Pitch;                      /* bytes, projekt.cpp:416 */
    uint32_t *Color;                    /* loaded_bitmap.Memory, ARGB8 */
    float *Z;                           /* game_render_commands.ZBuffer, projekt.cpp:170 */
    uint32_t ZStride;                   /* game_render_commands.Width, projekt.cpp:171 */
    int32_t *Prim;                      /* optional: index of the triangle owning each pixel
                                           (same stride as Z); not in the reference */
} orc_target;

typedef struct orc_stats {
    uint64_t Triangles;                 /* submitted */
    uint64_t Visible;                   /* with >= 1 edge record */
    uint64_t SpanRows;                  /* (triangle,row) pairs that produced a span */
    uint64_t Fragments;                 /* depth-tested pixels */
    uint64_t DepthPasses;               /* fragments that won the test when drawn */
    uint64_t RefWouldCrash;             /* triangles on which the verbatim reference
                                           null-dereferences (edges cross before the last row) */
    uint64_t TexelClamps;               /* textured pixels whose texel coordinates left the bitmap */
} orc_stats;

/* Compatibility switches of the span fill (mirror B200R_AVX_RIGHT_END_EXCLUSIVE / B200R_AVX_DEPTH_GE). */
#define ORC_COMPAT_RIGHT_END_EXCLUSIVE 4
#define ORC_COMPAT_DEPTH_GE 8
void orc_set_compat(int32_t Flags);

void orc_project_vertex(const float Cam[3], const orc_transform *T, float Out[3]);

/* Whole object (VertexCount/3 triangles) -> sorted edge records.  Edges/Temp need room for
 * VertexCount records.  Returns the number of records, or -1 for LightCount == 0. */
int32_t orc_fill_edge_table(const float *Pos, const float *Col, const float *Nrm,
                            uint32_t VertexCount, const float P[3], const orc_scene *Scene,
                            orc_edge *Edges, orc_edge *Temp);
/* Same with PhongShading != 0 (projekt.cpp:4012-4019, 4104-4109): unlit colours, vertex normals. */
int32_t orc_fill_edge_table_ex(const float *Pos, const float *Col, const float *Nrm,
                               uint32_t VertexCount, const float P[3], const orc_scene *Scene,
                               int32_t Phong, orc_edge *Edges, orc_edge *Temp);
/* Same for an object with a Bitmap (UV != 0; projekt.cpp:3919-3923, 4002-4008, 4034-4060,
 * 4078-4089): perspective-correct u/z, v/z, 1/z per edge; Gouraud colours are lit WHITE. */
int32_t orc_fill_edge_table_tex(const float *Pos, const float *Col, const float *Nrm, const float *UV,
                                uint32_t VertexCount, const float P[3], const orc_scene *Scene,
                                int32_t Phong, orc_edge *Edges, orc_edge *Temp);

void orc_merge_sort(uint32_t Count, orc_edge *First, orc_edge *Temp);

/* Level-1 walk of one triangle's (<= 3, sorted) edge records into the target.
 * Returns a bit mask: 1 = drew at least one span, 2 = the verbatim reference would crash. */
int32_t orc_draw_triangle(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimIndex,
                          orc_target *Target, orc_stats *Stats);
/* Same with per-pixel Phong shading (projekt.cpp:450-509, 551-552); Scene gives lights + transform. */
int32_t orc_draw_triangle_ex(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimIndex,
                             orc_target *Target, orc_stats *Stats, const orc_scene *Scene, int32_t Phong);
/* Same with a texture (projekt.cpp:427-446): every pixel's colour is the texel at
 * Round(uv * (dim - 1)), nearest.  The reference does not range-check the texel coordinates and
 * reads outside the bitmap when they leave it; DEFINED here: they are clamped to the bitmap
 * (cvtss2si's INT_MIN for NaN / overflow clamps to 0).  Stats->TexelClamps counts such pixels, so a
 * caller can tell whether a comparison with the verbatim reference is meaningful. */
int32_t orc_draw_triangle_tex(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimIndex,
                              orc_target *Target, orc_stats *Stats, const orc_scene *Scene, int32_t Phong,
                              const orc_texture *Texture);

/* LEVEL 0, whole-object semantics: DrawModel's intrusive active-edge list over
 * ALL edges of an object (projekt.cpp:198-303, 542-597), replayed link by link with indices instead
 * of pointers -- including what makes its images differ from level 1: consecutive list entries are
 * paired whatever triangle they belong to (:300-303, :584-592), and the exchanges (:562-583) relink
 * nodes without updating ListHead / ListTail.  Where the reference dereferences a null pointer
 * (empty list :262, :300; stale tail :222, :275) or would follow a cycle, the object STOPS drawing:
 * bit 2 of the result (the verbatim build crashes there).  Spans get PrimBase + their draw order as
 * owner.  Returns 1 if anything was drawn, | 2 if the reference would crash. */
int32_t orc_draw_object(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimBase,
                        orc_target *Target, orc_stats *Stats, const orc_scene *Scene, int32_t Phong,
                        const orc_texture *Texture, int32_t *SpansDrawn);
/* FillEdgeTable over the whole vertex array as ONE object, then orc_draw_object. */
int32_t orc_render_object(const float *Pos, const float *Col, const float *Nrm, const float *UV,
                          uint32_t VertexCount, const float P[3], const orc_scene *Scene, int32_t Phong,
                          const orc_texture *Texture, orc_target *Target, int32_t PrimBase, orc_stats *Stats);

/* Per-triangle semantics over a soup, in submission order.  PrimBase is added to the
 * triangle index stored in Target->Prim.  WouldCrash (optional, one byte per triangle). */
int32_t orc_render_triangles(const float *Pos, const float *Col, const float *Nrm,
                             uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                             orc_target *Target, int32_t PrimBase, uint8_t *WouldCrash,
                             orc_stats *Stats);
int32_t orc_render_triangles_ex(const float *Pos, const float *Col, const float *Nrm,
                                uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                                int32_t Phong, orc_target *Target, int32_t PrimBase, uint8_t *WouldCrash,
                                orc_stats *Stats);
/* UV and Texture both non-null: the soup is textured. */
int32_t orc_render_triangles_tex(const float *Pos, const float *Col, const float *Nrm, const float *UV,
                                 uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                                 int32_t Phong, const orc_texture *Texture, orc_target *Target,
                                 int32_t PrimBase, uint8_t *WouldCrash, orc_stats *Stats);

/* Same result, Threads workers with private targets folded in submission order. */
int32_t orc_render_triangles_mt(const float *Pos, const float *Col, const float *Nrm,
                                uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                                orc_target *Target, uint32_t Threads, orc_stats *Stats);

/* Callback with the signature oracle/ref_exports.inc expects for triangles the verbatim
 * reference cannot survive; User points at an orc_fallback_ctx. */
typedef struct orc_fallback_ctx {
    const float *Pos, *Col, *Nrm;
    float P[3];
    const orc_scene *Scene;
    int32_t Phong;
    const float *UV;                    /* textured soups: both non-null */
    const orc_texture *Texture;
} orc_fallback_ctx;
void orc_ref_fallback(void *User, uint32_t TriangleIndex, void *RefLoadedBitmap,
                      void *RefGameRenderCommands);

/* ConstructSphere (projekt.cpp:4123-4289) with StepCount as a parameter; arrays sized for
 * 3*(4*S*S - 4*S) vertices.  Returns the vertex count. */
uint32_t orc_construct_sphere(uint32_t StepCount, float *Pos, float *Col, float *Nrm, float *UV);

#ifdef __cplusplus
}
#endif
#endif
