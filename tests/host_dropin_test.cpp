// C++ host program written the way the reference's (absent) render-group walker would call the
// path: build render_entry_3d_object / game_render_commands / loaded_bitmap, then the call pair.
// Prints FNV-1a-64 hashes of colour and depth; tests/test_gpu_host_cpp.py compares them with the
// oracle's.  Usage: host_dropin_test <seed> <triangles> <width> <height>
#include "../cpu_renderer_b200/host/b200_dropin.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static uint64_t Fnv(const void *Data, size_t Words)
{
    const uint32_t *P = (const uint32_t *)Data;
    uint64_t H = 0xcbf29ce484222325ull;
    for(size_t i = 0; i < Words; ++i) H = (H ^ P[i])*0x100000001b3ull;
    return H;
}

int main(int argc, char **argv)
{
    if(argc < 5) return 2;
    // vertex streams are read from stdin as raw float32: positions, colours, normals (the Python
    // test generates them with the shared scene generator so both sides see identical inputs)
    const u32 Tris = (u32)atoi(argv[2]);
    const int W = atoi(argv[3]), H = atoi(argv[4]);
    std::vector<float> Pos(9u*Tris), Col(12u*Tris), Nrm(9u*Tris), UV(6u*Tris, 0.0f);
    if(fread(Pos.data(), 4, Pos.size(), stdin) != Pos.size()) return 3;
    if(fread(Col.data(), 4, Col.size(), stdin) != Col.size()) return 3;
    if(fread(Nrm.data(), 4, Nrm.size(), stdin) != Nrm.size()) return 3;

    std::vector<u32> Color((size_t)W*H, 0u);
    std::vector<r32> Depth((size_t)W*H, -1e30f);
    std::vector<edge_info> Edges(3u*Tris);

    light_info Light = { {5.0f, 5.0f, 8.0f}, {0.8f, 0.8f, 0.8f, 0.0f} };
    game_render_commands Commands; memset(&Commands, 0, sizeof(Commands));
    Commands.Width = (u32)W;
    Commands.ZBuffer = Depth.data();
    Commands.LightData.AmbientIntensity = {0.2f, 0.2f, 0.2f, 1.0f};
    Commands.LightData.LightCount = 1;
    Commands.LightData.Lights = &Light;
    Commands.Transform.MetersToPixels = H/2.0f;
    Commands.Transform.ScreenCenter = {W/2.0f, H/2.0f};
    Commands.Transform.FocalLength = 1.0f;
    Commands.Transform.DistanceAboveTarget = 10.0f;

    loaded_bitmap Target = { W, H, W*4, Color.data() };
    render_entry_3d_object Object; memset(&Object, 0, sizeof(Object));
    Object.VertexCount = 3u*Tris;
    Object.VertexData = Pos.data(); Object.ColorData = Col.data();
    Object.NormalData = Nrm.data(); Object.UVData = UV.data();
    Object.EdgeMemory = Edges.data();

    b200::Context Ctx(0);
    if(!Ctx.Ok()) { fprintf(stderr, "no device: %d\n", Ctx.Status); return 4; }
    int EdgeCount = b200::FillEdgeTable(Ctx, &Object, &Commands);            // projekt.cpp:3882
    if(EdgeCount < 0) { fprintf(stderr, "FillEdgeTable: %s\n", Ctx.Error()); return 5; }
    int Status = b200::DrawModel(Ctx, &Target, &Object, &Commands);           // projekt.cpp:162
    if(Status != B200R_OK) { fprintf(stderr, "DrawModel: %s\n", Ctx.Error()); return 6; }
    printf("%d %016llx %016llx\n", EdgeCount, (unsigned long long)Fnv(Color.data(), Color.size()),
           (unsigned long long)Fnv(Depth.data(), Depth.size()));
    return 0;
}
