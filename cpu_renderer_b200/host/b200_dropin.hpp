// Host-side mirror of the reference's call pair for C++ callers (the reference is a C++ unity
// build, projekt.cpp).  Same names, argument meaning and ownership as
//     u32  FillEdgeTable(render_entry_3d_object*, game_render_commands*, b32)     projekt.cpp:3882
//     void DrawModel(loaded_bitmap*, edge_info*, u32, game_render_commands*,
//                    loaded_bitmap *Bitmap = 0, b32 PhongShading = 0)             projekt.cpp:162
// but executed by libb200raster.so.  Inside the reference's own build define
// B200R_NO_REFERENCE_TYPES before including this header (the renderer's structs are used as is).
//
// Because the GPU path fuses the pair, DrawModel here takes the *object* whose EdgeMemory was
// filled: B200DrawModel(Target, Object, Commands).  Errors come back as negative codes instead
// of the reference's Assert.
#pragma once
#include "../../include/b200_raster.h"

namespace b200 {

class Context
{
public:
    explicit Context(int Device = -1) { Status = b200r_create(&Handle, Device); }
    ~Context() { b200r_destroy(Handle); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    bool Ok() const { return Status == B200R_OK; }
    const char *Error() const { return b200r_last_error(Handle); }
    b200r_context *Handle = nullptr;
    int Status = B200R_E_NO_DEVICE;
};

// FillEdgeTable: sorted edge_info records into Object->EdgeMemory; returns the edge count.
inline int FillEdgeTable(Context &C, render_entry_3d_object *Object, game_render_commands *Commands,
                         b32 PhongShading = 0)
{
    return b200r_fill_edge_table(C.Handle, Object, Commands, PhongShading);
}

// The render-group walker's per-object body (FillEdgeTable + DrawModel) for a batch of objects.
// Object->PhongShading and Object->Bitmap (+ UVData) select the shading exactly as in the reference.
// Flags: 0 = every triangle is its own object (fast path); B200R_WHOLE_OBJECT_AEL = the reference's
// whole-object active-edge list, pixel-identical to its images of multi-triangle objects (slow path).
inline int DrawModels(Context &C, loaded_bitmap *Buffer, const render_entry_3d_object *Objects, u32 Count,
                      game_render_commands *Commands, u32 Flags = 0)
{
    return b200r_render_objects(C.Handle, Objects, Count, Commands, Buffer, Flags);
}

inline int DrawModel(Context &C, loaded_bitmap *Buffer, const render_entry_3d_object *Object,
                     game_render_commands *Commands, u32 Flags = 0)
{
    return DrawModels(C, Buffer, Object, 1, Commands, Flags);
}

} // namespace b200
