// C ABI of libb200raster.so (include/b200_raster.h) and the host-side frame orchestration:
//   zrange_kernel -> [select_kernel] -> setup_kernel -> tile_scan_kernel -> finalize_kernel
//   -> scatter_kernel -> raster_kernel
// (whole-object mode: chain / order / emit kernels instead of the first three) on one CUDA stream,
// no host synchronisation inside a frame.  The only host/device handshake is the fill of the span,
// segment and queue lists: it is copied back right after the scan and looked at when the *next*
// call arrives (or at b200r_sync); if a list overflowed, the scatter and raster kernels of that
// frame returned without touching the targets, the lists are grown and the frame is issued again.
// The host-pointer call adds two copy streams: chunked uploads in front of the set-up launches,
// and targets that go up, are rastered and come back in bands of tile rows.
//
// There is no CPU fallback anywhere in this file: without a compute-capability-10 device every
// entry point returns B200R_E_NO_DEVICE.
#include "../../include/b200_raster.h"
#include "raster_device.cuh"

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace b200r;

namespace {

struct DeviceBuffer
{
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need)
    {
        if(need <= bytes) return cudaSuccess;
        if(ptr) { cudaFree(ptr); ptr = nullptr; bytes = 0; }
        size_t want = need + need/4 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if(e == cudaSuccess) bytes = want;
        return e;
    }
    void release() { if(ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
};

// The host-pointer call splits the frame into bands of tile rows: a band's targets are uploaded, the
// band rastered and read back while the next band's targets are still on the bus (PCIe is full duplex).
constexpr int kHostBands = 4;
// meshes smaller than this are not worth a pre-selection pass in front of a partial band
constexpr unsigned kSelectMinTriangles = 1u << 16;

// control words that are zeroed once per frame with a single memset
struct FrameWords
{
    unsigned pair_total;
    unsigned work_counter;
    unsigned band_counters[kHostBands]; // host-pointer path: one tile counter per raster band
    unsigned extra_total;
    unsigned overflow;                  // finalize_kernel: some list did not fit
    unsigned seg_max, span_max;         // finalize_kernel: largest region fill (incl. alias entries)
    unsigned stopped;                   // whole-object mode: objects that stopped where the reference dereferences null
    unsigned zkeys[2];                  // z range of the frame: ordered keys, then {zmax, 1/range} as floats
    unsigned long long counters[4];     // binned triangles, queue entries, segments, spans
    unsigned seg_fill[kSubAllocators];
    unsigned span_fill[kSubAllocators];
    unsigned scan_ticket;               // look-back scan of the bin counts: chunk numbers
    unsigned pad_;
    unsigned long long scan_state[kScanMaxChunks];   // ... and the published chunk sums / prefixes
};

} // namespace

struct b200r_context
{
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t total_ready = nullptr;
    std::string error;
    int tile_w = 64, tile_h = 32;
    bool tile_auto = true;              // b200r_set_tile(0, 0): the render call picks the tile from the triangle density
    int span_words = kSpanWords;        // of the last issued frame (kSpanWordsPhong if it had a Phong mesh)
    int refill_lanes = 12, pend_lanes = 4;   // raster_kernel thresholds (env B200R_REFILL / B200R_PEND)
    int tall_mode = 1, tall_shift = 2, tall_chunk = 4;   // tall-triangle frames: height bins, row chunk (env B200R_TALL=0 off, _SHIFT, _CHUNK)
    int split_mode = 1, split_tpc = 0, split_rows = 0;   // row-parallel set-up (env B200R_SPLIT=0 off, B200R_SPLIT_TPC / _ROWS force)

    DeviceBuffer recs, segs, spans, tiles, pairs, words;
    // b200r_fill_edge_table: second set-up pass of a textured object, edge counts + offsets, sort keys / payloads
    // (two buffers each), the assembled edge_info table
    DeviceBuffer recs_uv, et_counts, et_keys, et_vals, et_out;
    FrameWords *h_words = nullptr;      // pinned
    // row-band pre-selection (select_kernel): one index list for the frame's meshes, one count per mesh
    DeviceBuffer sel_list, sel_counts;
    // whole-object mode (B200R_WHOLE_OBJECT_AEL): the last issued frame's objects
    bool object_mode = false;
    std::vector<ObjectDesc> obj_host;
    DeviceBuffer obj_dev, edges_pristine, edges_work;
    size_t edges_bytes = 0;
    unsigned obj_total_slots = 0;       // sum of the objects' span bounds
    unsigned obj_smem_bytes = 0;        // walk state of the largest object that fits into shared memory
    // three-phase path: chain offsets per edge, the value chains, the pair list, per-object counters
    DeviceBuffer obj_chain_base, obj_chains, obj_pairs, obj_flags;
    unsigned obj_chain_T = 0, obj_order_smem = 0, obj_max_bound = 0, obj_max_edges = 0;
    bool obj_three_phase = false;
    bool obj_force_serial = false;      // env B200R_OBJECT_SERIAL=1 (tests: keeps the fallback kernel covered)
    // textures of the last issued frame: distinct b200r_device_texture descriptors, uploaded as a table
    std::vector<TexDesc> tex_host;
    DeviceBuffer tex_dev;

    // the last issued frame, kept so it can be issued again after the pair list grew
    bool pending = false;
    ViewParams view;
    std::vector<MeshParams> meshes;
    b200r_device_target target;
    unsigned total_tris = 0;
    unsigned ntiles = 0;

    // fused gather (b200r_set_gather_target): an image of the whole screen every band is mirrored into
    bool has_gather = false;
    b200r_device_target gather = {};
    std::vector<void *> peer_owned, peer_opened;    // b200r_peer_alloc / b200r_peer_open

    b200r_frame_stats stats = {};
    bool profiling = false;
    cudaEvent_t stage_ev[B200R_STAGES + 1] = {};

    // host-pointer path mirrors
    DeviceBuffer d_pos, d_col, d_nrm, d_uv, d_color, d_depth;
    struct HostTexture { const loaded_bitmap *host; DeviceBuffer pixels; b200r_device_texture desc; };
    std::vector<HostTexture> host_textures;     // device copies of the objects' Bitmaps, one per distinct pointer
    // host-pointer path: uploads run on their own stream and the frame's kernels wait only for
    // what they read -- the z-range pass for all positions, each chunk's set-up for that chunk's
    // colours and normals, the raster kernel for the targets -- so set-up and binning overlap the
    // rest of the upload instead of following it
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t pos_ready = nullptr, target_ready = nullptr;
    cudaEvent_t band_ready[kHostBands] = {}, band_done[kHostBands] = {};
    int host_bands = 1;                     // 1: whole frame at once (small targets)
    int host_alias_rows = -1;               // host-pointer calls: rows of the caller's targets are contiguous (0/1); -1: device call
    struct { void *color; size_t color_pitch; void *depth; size_t depth_pitch; int W, H, wpad; } host_out = {};
    std::vector<cudaEvent_t> chunk_ready;   // pool, grown on demand
    bool host_path = false;                 // issue_frame: honour the events above
};

static int fail(b200r_context *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if(c)
    {
        c->error = what;
        if(e != cudaSuccess) { c->error += ": "; c->error += cudaGetErrorString(e); }
    }
    return code;
}

// Makes the context's device current and restores the caller's on scope exit (a host process may
// drive several GPUs from one thread, with other CUDA libraries in between).
struct DeviceGuard
{
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if(err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if(err == cudaSuccess) prev = -1;             // nothing to restore
    }
    ~DeviceGuard() { if(prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define ENTER(c) DeviceGuard guard_((c)->device); \
                 if(guard_.err != cudaSuccess) return fail((c), B200R_E_CUDA, "cudaSetDevice", guard_.err)

#define CU(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) return fail(c, B200R_E_CUDA, #call, e_); } while(0)

static int fill_view(b200r_context *c, const game_render_commands *cmd, const b200r_device_target *t, ViewParams &v,
                     bool all_phong = false)
{
    if(!cmd || !t) return fail(c, B200R_E_INVALID, "null Commands / Target");
    const light_data &ld = cmd->LightData;
    // Gouraud: 0 lights leaves MinColor uninitialised in the reference (projekt.cpp:4022-4045).
    // Phong: 0 lights is defined (FinalColor stays 0, :448).
    if((ld.LightCount == 0 && !all_phong) || ld.LightCount > (u32)kMaxLights || (ld.LightCount && !ld.Lights))
        return fail(c, B200R_E_UNSUPPORTED, "LightCount must be 1..8 (0 only when every object is Phong shaded)");
    if(t->Width <= 0 || t->Height <= 0 || t->BandRows <= 0 || t->BandFirstRow < 0 ||
       t->BandFirstRow + t->BandRows > t->Height || !t->Color || !t->Depth)
        return fail(c, B200R_E_INVALID, "bad target geometry");
    if(t->ColorPitch < t->Width*4 || (t->ColorPitch & 3) || t->DepthStride < t->Width)
        return fail(c, B200R_E_INVALID, "bad target pitch");
    v.m2p = cmd->Transform.MetersToPixels;
    v.cx = cmd->Transform.ScreenCenter.x; v.cy = cmd->Transform.ScreenCenter.y;
    v.focal = cmd->Transform.FocalLength; v.dist = cmd->Transform.DistanceAboveTarget;
    v.amb[0] = ld.AmbientIntensity.x; v.amb[1] = ld.AmbientIntensity.y;
    v.amb[2] = ld.AmbientIntensity.z; v.amb[3] = ld.AmbientIntensity.w;
    v.nlights = (int)ld.LightCount;
    for(int i = 0; i < kMaxLights; ++i)
    {
        DevLight L = {0, 0, 0, 0, 0, 0, 0};
        if(i < v.nlights)
        {
            const light_info &s = ld.Lights[i];
            L.px = s.P.x; L.py = s.P.y; L.pz = s.P.z;
            L.ir = s.Intensity.x; L.ig = s.Intensity.y; L.ib = s.Intensity.z; L.ia = s.Intensity.w;
        }
        v.lights[i] = L;
    }
    v.width = t->Width; v.height = t->Height;
    v.band_y0 = t->BandFirstRow; v.band_y1 = t->BandFirstRow + t->BandRows;
    v.tile_w = c->tile_w; v.tile_h = c->tile_h;
    v.tile_w_shift = 0; while((1 << v.tile_w_shift) < v.tile_w) ++v.tile_w_shift;
    v.tile_h_shift = 0; while((1 << v.tile_h_shift) < v.tile_h) ++v.tile_h_shift;
    v.tiles_x = (t->Width + c->tile_w - 1)/c->tile_w;
    v.tiles_y = (t->BandRows + c->tile_h - 1)/c->tile_h;
    v.right_end_exclusive = 0; v.depth_ge = 0;             // set from the call's flags by b200r_render_device
    v.alias_rows = (t->ColorPitch == t->Width*4 && t->DepthStride == t->Width) ? 1 : 0;
    // host-pointer calls render into a device mirror whose rows are padded to 64 pixels; whether a
    // column == Width write lands in the next row (projekt.cpp:414-419) is a property of the CALLER's
    // layout, not of the mirror's
    if(c->host_alias_rows >= 0) v.alias_rows = c->host_alias_rows;
    if(v.tiles_x > 65535 || v.tiles_y > 65535) return fail(c, B200R_E_UNSUPPORTED, "more than 65535 tiles per axis");
    // the raster kernel's stash packs (column - tile column) into 19 signed bits
    if(t->Width > 262143) return fail(c, B200R_E_UNSUPPORTED, "target wider than 262143 pixels");
    return B200R_OK;
}

// Automatic tile shape (b200r_set_tile(0, 0), the default).  Small triangles (a few pixels of target per
// triangle) bin and rasterise best in 64x16 tiles: short spans, little replay, many tiles in flight.
// Larger ones take 128x8: spans cross fewer tile columns.  Both are 20 KB tiles, run by 8 and 4 warps
// (measured on C2 / C3 / C4, profiles/README.md).
static void choose_tile(b200r_context *c, uint64_t triangles, int width, int rows)
{
    if(!c->tile_auto) return;
    const double px_per_tri = (double)width*(double)(rows > 0 ? rows : 0)/(double)(triangles ? triangles : 1);
    if(px_per_tri < 4.0) { c->tile_w = 64; c->tile_h = 16; } else { c->tile_w = 128; c->tile_h = 8; }
}

// Enqueue every kernel of the frame described by c->view / c->meshes / c->target.
static int issue_frame(b200r_context *c)
{
    const ViewParams &v = c->view;
    const unsigned ntiles = c->ntiles;
    const unsigned nbins = ntiles*kDepthBuckets;        // one sub-queue per tile and depth bucket
    if((unsigned long long)ntiles*kDepthBuckets > (unsigned long long)kScanMaxChunks*8192ull)
        return fail(c, B200R_E_UNSUPPORTED, "more than 16 M (tile, depth bucket) bins in one band");
    unsigned *tile_count = (unsigned *)c->tiles.ptr;
    unsigned *tile_fill = tile_count + nbins;
    unsigned *tile_offset = tile_fill + nbins;          // nbins + 1 entries
    FrameWords *words = (FrameWords *)c->words.ptr;

    CU(cudaMemsetAsync(tile_count, 0, (size_t)nbins*2*sizeof(unsigned), c->stream));
    CU(cudaMemsetAsync(words, 0, sizeof(FrameWords), c->stream));

    unsigned seg_cap = (unsigned)std::min<size_t>(c->segs.bytes/sizeof(SegInfo), 0xffffffffu);
    unsigned span_cap = (unsigned)std::min<size_t>(c->spans.bytes/(c->span_words*sizeof(uint32_t)), 0xffffffffu);
    if(c->object_mode)
    {
        // one slot index addresses a span AND its one-row segment: both arrays get the same geometry
        seg_cap = span_cap = std::min(seg_cap, span_cap)/kSubAllocators*kSubAllocators;
    }
    SetupOutputs so;
    so.recs = nullptr;
    so.spans = (uint32_t *)c->spans.ptr;
    so.span_words = c->span_words;
    so.segs = (SegInfo *)c->segs.ptr;
    so.seg_fill = words->seg_fill;
    so.span_fill = words->span_fill;
    so.extra_total = &words->extra_total;
    so.seg_capacity = seg_cap;
    so.span_capacity = span_cap;
    so.tile_count = tile_count;
    so.zrange = reinterpret_cast<const float *>(words->zkeys);
    so.counters = words->counters;
    if(c->profiling) CU(cudaEventRecord(c->stage_ev[0], c->stream));
    if(c->object_mode)
    {
        // slots are striped over the regions (object_walk_kernel.cu), so the fills are known up front
        unsigned fills[kSubAllocators];
        for(int r = 0; r < kSubAllocators; ++r)
            fills[r] = c->obj_total_slots/kSubAllocators + ((unsigned)r < c->obj_total_slots%kSubAllocators ? 1u : 0u);
        CU(cudaMemcpyAsync(words->seg_fill, fills, sizeof(fills), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(words->span_fill, fills, sizeof(fills), cudaMemcpyHostToDevice, c->stream));
        // the walk keeps what DrawModel mutates (running values, Next) beside the edge records, in shared
        // memory when the object fits and in edges_work otherwise: a re-issue starts from the same records
        ObjectWalkParams op;
        op.edges = c->edges_pristine.ptr;
        op.state_scratch = (float *)c->edges_work.ptr;
        op.state_smem_bytes = c->obj_smem_bytes;
        op.objects = (const ObjectDesc *)c->obj_dev.ptr; op.nobjects = (unsigned)c->obj_host.size();
        op.spans = so.spans; op.span_words = c->span_words; op.segs = so.segs;
        op.extra_total = &words->extra_total;
        op.seg_capacity = seg_cap; op.span_capacity = span_cap; op.region_size = seg_cap/kSubAllocators;
        op.tile_count = tile_count; op.counters = words->counters; op.stopped = &words->stopped;
        op.chain_base = nullptr; op.chains = nullptr; op.chain_T = 0; op.pair_list = nullptr;
        op.produced = nullptr; op.fallback = nullptr; op.order_smem_bytes = 0;
        op.max_bound = c->obj_max_bound; op.max_edges = c->obj_max_edges;
        if(c->obj_three_phase)
        {
            op.chain_base = (const unsigned *)c->obj_chain_base.ptr;
            op.chains = (float *)c->obj_chains.ptr; op.chain_T = c->obj_chain_T;
            op.pair_list = (uint4 *)c->obj_pairs.ptr;
            op.produced = (unsigned *)c->obj_flags.ptr; op.fallback = op.produced + op.nobjects;
            op.order_smem_bytes = c->obj_order_smem;
            CU(cudaMemsetAsync(c->obj_flags.ptr, 0, (size_t)op.nobjects*2*sizeof(unsigned), c->stream));
            c->stats.KernelLaunches += 3;
        }
        CU(launch_object_walk(v, op, c->stream));
        c->stats.KernelLaunches += 1;
    }
    else
    {
    if(c->host_path) CU(cudaStreamWaitEvent(c->stream, c->pos_ready, 0));
    // A partial band of a big mesh (multi-GPU row bands): pre-select the triangles that can reach it
    // from their positions alone, so that the set-up kernel's front end runs on those only.  The
    // selection pass reads every position anyway and folds the z range in; the other meshes take
    // zrange_kernel.  Every set-up launch needs the z range of ALL meshes, so these passes come first.
    const bool partial_band = v.band_y0 > 0 || v.band_y1 < v.height;
    bool selecting = false;
    if(partial_band && !c->host_path)
        for(const MeshParams &m : c->meshes) selecting |= m.ntri >= kSelectMinTriangles;
    if(selecting)
    {
        CU(c->sel_list.reserve((size_t)std::max<unsigned>(c->total_tris, 1)*sizeof(unsigned)));
        CU(c->sel_counts.reserve(c->meshes.size()*sizeof(unsigned)));
        CU(cudaMemsetAsync(c->sel_counts.ptr, 0, c->meshes.size()*sizeof(unsigned), c->stream));
    }
    for(size_t i = 0; i < c->meshes.size(); ++i)
    {
        MeshParams &m = c->meshes[i];
        m.tri_list = nullptr; m.tri_count = nullptr;
        if(selecting && m.ntri >= kSelectMinTriangles)
        {
            unsigned *list = (unsigned *)c->sel_list.ptr + m.prim_base, *count = (unsigned *)c->sel_counts.ptr + i;
            launch_select(v, m, list, count, words->zkeys, c->sm_count, c->stream);
            m.tri_list = list; m.tri_count = count;
        }
        else launch_zrange(m, words->zkeys, c->stream);
        if(m.ntri) c->stats.KernelLaunches += 1;
    }
    launch_zrange_finish(words->zkeys, c->stream);
    c->stats.KernelLaunches += 1;
    for(size_t i = 0; i < c->meshes.size(); ++i)
    {
        const MeshParams &m = c->meshes[i];
        if(c->host_path && i < c->chunk_ready.size()) CU(cudaStreamWaitEvent(c->stream, c->chunk_ready[i], 0));
        launch_setup(v, m, so, c->stream);
        if(m.ntri) c->stats.KernelLaunches += 1;
    }
    }
    if(c->profiling) CU(cudaEventRecord(c->stage_ev[1], c->stream));
    launch_tile_scan(tile_count, tile_offset, nbins, &words->pair_total, words->scan_state, &words->scan_ticket, c->stream);
    c->stats.KernelLaunches += 1;
    const unsigned pair_cap_f = (unsigned)(c->pairs.bytes/sizeof(unsigned));
    FinalizeParams fp;
    fp.seg_fill = words->seg_fill; fp.span_fill = words->span_fill;
    fp.extra_total = &words->extra_total; fp.pair_total = &words->pair_total;
    fp.seg_capacity = seg_cap; fp.span_capacity = span_cap; fp.pair_capacity = pair_cap_f;
    fp.overflow = &words->overflow; fp.seg_max = &words->seg_max; fp.span_max = &words->span_max;
    launch_finalize(fp, c->stream);
    c->stats.KernelLaunches += 1;
    if(c->profiling) CU(cudaEventRecord(c->stage_ev[2], c->stream));
    CU(cudaMemcpyAsync(c->h_words, words, offsetof(FrameWords, scan_ticket), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->total_ready, c->stream));

    const unsigned pair_cap = (unsigned)(c->pairs.bytes/sizeof(unsigned));
    ScatterParams sp;
    sp.segs = so.segs;
    sp.seg_fill = words->seg_fill; sp.extra_total = &words->extra_total; sp.overflow = &words->overflow;
    sp.seg_capacity = seg_cap;
    sp.tiles_x = v.tiles_x;
    sp.tile_offset = tile_offset; sp.tile_fill = tile_fill; sp.pair_list = (unsigned *)c->pairs.ptr;
    sp.pair_capacity = pair_cap;
    launch_scatter(sp, c->stream);
    if(c->total_tris || c->object_mode) c->stats.KernelLaunches += 1;
    if(c->profiling) CU(cudaEventRecord(c->stage_ev[3], c->stream));

    RasterParams rp;
    rp.v = v;
    rp.spans = so.spans;
    rp.span_words = c->span_words;
    rp.overflow = &words->overflow;
    rp.seg_capacity = seg_cap;
    rp.span_capacity = span_cap;
    rp.tile_offset = tile_offset;
    rp.pair_list = (const unsigned *)c->pairs.ptr;
    rp.pair_total = &words->pair_total;
    rp.pair_capacity = pair_cap;
    rp.work_counter = &words->work_counter;
    rp.tile_begin = 0; rp.tile_end = ntiles;
    rp.ntiles = ntiles;
    rp.color = c->target.Color;
    rp.depth = c->target.Depth;
    rp.color_pitch_words = c->target.ColorPitch/4;
    rp.depth_stride = c->target.DepthStride;
    rp.bulk_ok = ((((uintptr_t)c->target.Color) & 15) == 0 && (((uintptr_t)c->target.Depth) & 15) == 0 &&
                  (c->target.ColorPitch & 15) == 0 && ((c->target.DepthStride*4) & 15) == 0 &&
                  (c->target.Width & 3) == 0) ? 1 : 0;
    rp.gather_color = nullptr; rp.gather_depth = nullptr;
    rp.gather_pitch_words = 0; rp.gather_depth_stride = 0; rp.gather_bulk_ok = 0;
    if(c->has_gather && !c->host_path)
    {
        const b200r_device_target &g = c->gather;
        rp.gather_color = g.Color; rp.gather_depth = g.Depth;
        rp.gather_pitch_words = g.ColorPitch/4; rp.gather_depth_stride = g.DepthStride;
        rp.gather_bulk_ok = ((((uintptr_t)g.Color) & 15) == 0 && (g.ColorPitch & 15) == 0 &&
                             (!g.Depth || ((((uintptr_t)g.Depth) & 15) == 0 && ((g.DepthStride*4) & 15) == 0))) ? 1 : 0;
    }
    rp.textures = nullptr;
    rp.texture_count = (unsigned)c->tex_host.size();
    rp.mode = (c->span_words == kSpanWordsPhong) ? kRasterGeneral : (c->tex_host.empty() ? kRasterPlain : kRasterTextured);
    if(!c->tex_host.empty())
    {
        CU(c->tex_dev.reserve(c->tex_host.size()*sizeof(TexDesc)));
        CU(cudaMemcpyAsync(c->tex_dev.ptr, c->tex_host.data(), c->tex_host.size()*sizeof(TexDesc),
                           cudaMemcpyHostToDevice, c->stream));
        rp.textures = (const TexDesc *)c->tex_dev.ptr;
    }
    rp.refill_lanes = c->refill_lanes;
    rp.pend_lanes = c->pend_lanes;
    const int nbands = (c->host_path && c->host_bands > 1) ? c->host_bands : 1;
    for(int b = 0; b < nbands; ++b)
    {
        const int tr0 = v.tiles_y*b/nbands, tr1 = v.tiles_y*(b + 1)/nbands;
        rp.tile_begin = (unsigned)(tr0*v.tiles_x); rp.tile_end = (unsigned)(tr1*v.tiles_x);
        if(rp.tile_end <= rp.tile_begin) continue;
        rp.work_counter = (nbands == 1) ? &words->work_counter : &words->band_counters[b];
        if(c->host_path) CU(cudaStreamWaitEvent(c->stream, nbands == 1 ? c->target_ready : c->band_ready[b], 0));
        cudaError_t e = launch_raster(rp, c->sm_count, c->stream);
        if(e != cudaSuccess) return fail(c, B200R_E_CUDA, "raster_kernel launch", e);
        c->stats.KernelLaunches += 1;
        if(nbands > 1)
        {
            // read the band back on the second copy stream while the next band is rastered.  (After an
            // overflow verdict the kernel did nothing and this returns the caller's own pixels; the
            // re-issued frame copies again.)
            const int y0 = tr0*v.tile_h, y1 = std::min(tr1*v.tile_h, c->host_out.H);
            const size_t dp = (size_t)c->host_out.wpad*4, rowbytes = (size_t)c->host_out.W*4;
            CU(cudaEventRecord(c->band_done[b], c->stream));
            CU(cudaStreamWaitEvent(c->d2h_stream, c->band_done[b], 0));
            CU(cudaMemcpy2DAsync((char *)c->host_out.color + (size_t)y0*c->host_out.color_pitch, c->host_out.color_pitch,
                                 (char *)c->d_color.ptr + (size_t)y0*dp, dp, rowbytes, (size_t)(y1 - y0),
                                 cudaMemcpyDeviceToHost, c->d2h_stream));
            CU(cudaMemcpy2DAsync((char *)c->host_out.depth + (size_t)y0*c->host_out.depth_pitch, c->host_out.depth_pitch,
                                 (char *)c->d_depth.ptr + (size_t)y0*dp, dp, rowbytes, (size_t)(y1 - y0),
                                 cudaMemcpyDeviceToHost, c->d2h_stream));
        }
    }
    if(c->profiling) CU(cudaEventRecord(c->stage_ev[4], c->stream));
    CU(cudaGetLastError());
    c->pending = true;
    return B200R_OK;
}

// Look at the pair total of the last issued frame; grow the list and re-issue if it overflowed.
static int settle_pending(b200r_context *c)
{
    while(c->pending)
    {
        CU(cudaEventSynchronize(c->total_ready));
        const FrameWords &hw = *c->h_words;
        const unsigned total = hw.pair_total;
        uint64_t nseg = hw.extra_total, nspan = hw.extra_total;
        for(int r = 0; r < kSubAllocators; ++r) { nseg += hw.seg_fill[r]; nspan += hw.span_fill[r]; }
        c->stats.Binned = hw.counters[0];
        c->stats.TilePairs = hw.counters[1];
        c->stats.Segments = nseg;
        c->stats.Spans = nspan;
        c->stats.AliasPixels = hw.extra_total;
        c->stats.StoppedObjects = hw.stopped;
        c->pending = false;
        if(hw.overflow)
        {
            // the scatter and raster kernels of that frame saw the same verdict and did nothing.
            // Every region must hold the fullest region's fill (plus slack for run-to-run jitter
            // in which CTA lands in which region).
            CU(cudaStreamSynchronize(c->stream));
            const unsigned seg_region = (unsigned)(c->segs.bytes/sizeof(SegInfo))/kSubAllocators;
            const unsigned span_region = (unsigned)(c->spans.bytes/(c->span_words*sizeof(uint32_t)))/kSubAllocators;
            const bool truncated = hw.seg_max > seg_region || hw.span_max > span_region;
            if(hw.seg_max > seg_region)
                CU(c->segs.reserve(((size_t)hw.seg_max + hw.seg_max/8 + 64)*kSubAllocators*sizeof(SegInfo)));
            if(hw.span_max > span_region)
                CU(c->spans.reserve(((size_t)hw.span_max + hw.span_max/8 + 64)*kSubAllocators*c->span_words*sizeof(uint32_t)));
            // with truncated lists the queue total was an under-count: leave headroom
            CU(c->pairs.reserve((size_t)std::max<uint64_t>(total, truncated ? nspan*2 : 0)*sizeof(unsigned)));
            c->stats.Reruns += 1;
            int rc = issue_frame(c);
            if(rc != B200R_OK) return rc;
        }
    }
    return B200R_OK;
}

extern "C" {

int b200r_create(b200r_context **out, int device)
{
    if(!out) return B200R_E_INVALID;
    *out = nullptr;
    int count = 0;
    if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return B200R_E_NO_DEVICE;
    if(device < 0) { if(cudaGetDevice(&device) != cudaSuccess) return B200R_E_NO_DEVICE; }
    if(device >= count) return B200R_E_INVALID;
    cudaDeviceProp prop;
    if(cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200R_E_NO_DEVICE;
    if(prop.major != 10) return B200R_E_NO_DEVICE;          // kernels are built for sm_100a only
    b200r_context *c = new (std::nothrow) b200r_context;
    if(!c) return B200R_E_NOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    DeviceGuard guard(device);
    if(guard.err != cudaSuccess ||
       cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
       cudaEventCreateWithFlags(&c->total_ready, cudaEventDisableTiming) != cudaSuccess ||
       cudaMallocHost((void **)&c->h_words, sizeof(FrameWords)) != cudaSuccess ||
       c->words.reserve(sizeof(FrameWords)) != cudaSuccess)
    {
        b200r_destroy(c);
        return B200R_E_CUDA;
    }
    c->stream = c->own_stream;
    memset(c->h_words, 0, sizeof(FrameWords));
    if(const char *e = getenv("B200R_REFILL")) c->refill_lanes = std::max(1, std::min(32, atoi(e)));
    if(const char *e = getenv("B200R_PEND")) c->pend_lanes = std::max(1, std::min(32, atoi(e)));
    if(const char *e = getenv("B200R_SPLIT")) c->split_mode = atoi(e);
    if(const char *e = getenv("B200R_TALL")) c->tall_mode = atoi(e);
    if(const char *e = getenv("B200R_TALL_SHIFT")) c->tall_shift = std::max(0, std::min(6, atoi(e)));
    if(const char *e = getenv("B200R_TALL_CHUNK")) c->tall_chunk = std::max(1, atoi(e));
    if(const char *e = getenv("B200R_SPLIT_TPC"))
    {
        const int q = atoi(e);
        if(q == 8 || q == 16 || q == 32 || q == 64) c->split_tpc = q;
    }
    if(const char *e = getenv("B200R_SPLIT_ROWS"))
    {
        const int q = atoi(e);
        if(q == 8 || q == 16 || q == 32 || q == 64 || q == 128) c->split_rows = q;
    }
    if(const char *e = getenv("B200R_OBJECT_SERIAL")) c->obj_force_serial = atoi(e) != 0;   // whole-object mode: serial walk only
    *out = c;
    return B200R_OK;
}

void b200r_destroy(b200r_context *c)
{
    if(!c) return;
    DeviceGuard guard(c->device);
    if(c->stream) cudaStreamSynchronize(c->stream);
    c->recs.release(); c->segs.release(); c->spans.release(); c->tiles.release(); c->pairs.release(); c->words.release();
    c->d_pos.release(); c->d_col.release(); c->d_nrm.release(); c->d_uv.release(); c->d_color.release(); c->d_depth.release();
    c->tex_dev.release();
    c->recs_uv.release(); c->et_counts.release(); c->et_keys.release(); c->et_vals.release(); c->et_out.release();
    c->obj_dev.release(); c->edges_pristine.release(); c->edges_work.release();
    c->sel_list.release(); c->sel_counts.release();
    c->obj_chain_base.release(); c->obj_chains.release(); c->obj_pairs.release(); c->obj_flags.release();
    for(b200r_context::HostTexture &ht : c->host_textures) ht.pixels.release();
    for(void *q : c->peer_opened) cudaIpcCloseMemHandle(q);
    for(void *q : c->peer_owned) cudaFree(q);
    for(cudaEvent_t e : c->stage_ev) if(e) cudaEventDestroy(e);
    if(c->h_words) cudaFreeHost(c->h_words);
    if(c->total_ready) cudaEventDestroy(c->total_ready);
    if(c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if(c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for(cudaEvent_t e : c->band_ready) if(e) cudaEventDestroy(e);
    for(cudaEvent_t e : c->band_done) if(e) cudaEventDestroy(e);
    if(c->pos_ready) cudaEventDestroy(c->pos_ready);
    if(c->target_ready) cudaEventDestroy(c->target_ready);
    for(cudaEvent_t e : c->chunk_ready) if(e) cudaEventDestroy(e);
    if(c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char *b200r_last_error(const b200r_context *c) { return c ? c->error.c_str() : "null context"; }

int b200r_set_stream(b200r_context *c, void *s)
{
    if(!c) return B200R_E_INVALID;
    ENTER(c);
    int rc = settle_pending(c);
    if(rc != B200R_OK) return rc;
    CU(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return B200R_OK;
}

int b200r_sync(b200r_context *c)
{
    if(!c) return B200R_E_INVALID;
    ENTER(c);
    int rc = settle_pending(c);
    if(rc != B200R_OK) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return B200R_OK;
}

int b200r_set_tile(b200r_context *c, int w, int h)
{
    if(!c) return B200R_E_INVALID;
    if(w == 0 && h == 0) { c->tile_auto = true; return B200R_OK; }
    if(!((w == 64 && h == 32) || (w == 32 && h == 32) || (w == 128 && h == 16) || (w == 64 && h == 16) ||
         (w == 128 && h == 32) || (w == 256 && h == 8) || (w == 128 && h == 8) || (w == 256 && h == 4)))
        return fail(c, B200R_E_INVALID, "tile must be 64x32, 32x32, 128x16, 64x16, 128x32, 256x8, 128x8 or 256x4");
    int rc = b200r_sync(c);
    if(rc != B200R_OK) return rc;
    c->tile_w = w; c->tile_h = h; c->tile_auto = false;
    return B200R_OK;
}

int b200r_render_device(b200r_context *c, const b200r_device_mesh *meshes, u32 mesh_count,
                        const game_render_commands *cmd, const b200r_device_target *target, u32 flags)
{
    if(!c) return B200R_E_INVALID;
    if(flags & B200R_WHOLE_OBJECT_AEL)
        return fail(c, B200R_E_UNSUPPORTED, "whole-object mode needs the host-pointer call (b200r_render_objects)");
    if(mesh_count && !meshes) return fail(c, B200R_E_INVALID, "null Meshes");
    ENTER(c);
    int rc = settle_pending(c);                 // the previous frame must be complete in the stream
    if(rc != B200R_OK) return rc;

    bool any_phong = false, all_phong = mesh_count > 0, any_tex = false;
    for(u32 i = 0; i < mesh_count; ++i)
    {
        const bool ph = (meshes[i].Flags & B200R_MESH_PHONG) != 0;
        any_phong |= ph; all_phong &= ph;
        any_tex |= meshes[i].Texture != nullptr;
    }
    if(target)
    {
        uint64_t tris = 0;
        for(u32 i = 0; i < mesh_count; ++i) tris += meshes[i].TriangleCount;
        choose_tile(c, tris, target->Width, target->Height);    // triangle density of the SCREEN: a band sees the same triangles
    }
    ViewParams v;
    rc = fill_view(c, cmd, target, v, all_phong);
    if(rc != B200R_OK) return rc;

    v.right_end_exclusive = (flags & B200R_AVX_RIGHT_END_EXCLUSIVE) ? 1 : 0;
    v.depth_ge = (flags & B200R_AVX_DEPTH_GE) ? 1 : 0;
    if(c->has_gather && (c->gather.Width != target->Width || c->gather.Height != target->Height))
        return fail(c, B200R_E_INVALID, "gather target and render target describe different screens");

    uint64_t total = 0;
    std::vector<MeshParams> ms;
    std::vector<TexDesc> texs;
    ms.reserve(mesh_count);
    for(u32 i = 0; i < mesh_count; ++i)
    {
        const b200r_device_mesh &m = meshes[i];
        if(m.TriangleCount && (!m.Positions || !m.Colors || !m.Normals))
            return fail(c, B200R_E_INVALID, "mesh with null attribute pointer");
        MeshParams mp;
        mp.pos = m.Positions; mp.col = m.Colors; mp.nrm = m.Normals;
        mp.ntri = m.TriangleCount;
        mp.px = m.P.x; mp.py = m.P.y; mp.pz = m.P.z;
        mp.prim_base = (unsigned)total;
        mp.phong = (m.Flags & B200R_MESH_PHONG) ? 1 : 0;
        mp.uv = nullptr; mp.tex = -1; mp.white = 0;
        mp.tri_list = nullptr; mp.tri_count = nullptr;
        mp.tris_per_cta = 0; mp.part_rows = 0; mp.sort_shift = 0; mp.row_chunk = 0x7fffffff;
        if(m.Texture)
        {
            const b200r_device_texture &t = *m.Texture;
            if(m.TriangleCount && !m.UVs) return fail(c, B200R_E_INVALID, "textured mesh without UVs");
            if(!t.Memory || t.Width <= 0 || t.Height <= 0 || t.Pitch < t.Width*4 || (t.Pitch & 3))
                return fail(c, B200R_E_INVALID, "bad texture geometry");
            size_t k = 0;
            for(; k < texs.size(); ++k)
                if(texs[k].mem == t.Memory && texs[k].w == t.Width && texs[k].h == t.Height && texs[k].pitch == t.Pitch) break;
            if(k == texs.size())
            {
                if(k >= 65536) return fail(c, B200R_E_UNSUPPORTED, "more than 65536 distinct textures per call");
                TexDesc d; d.mem = t.Memory; d.w = t.Width; d.h = t.Height; d.pitch = t.Pitch;
                texs.push_back(d);
            }
            mp.uv = m.UVs; mp.tex = (int)k;
        }
        total += m.TriangleCount;
        ms.push_back(mp);
    }
    if(total > 0x7fffffffull) return fail(c, B200R_E_UNSUPPORTED, "more than 2^31-1 triangles per call");
    // Frames of few triangles: too few threads for a thread per triangle, and if the triangles are tall each
    // thread is a long latency chain.  Stage fewer triangles per CTA (aim: 8 CTAs per SM) and cut tall
    // triangles into row slabs walked by separate threads (setup_kernel<..., SPLIT>).
    // (only where a triangle has at least 64 target pixels to itself: small triangles are short, and half-empty
    // CTAs cost a frame of 100 000 ten-pixel triangles a quarter of its set-up time)
    // Measured (C3 scaled, 4K): 500 / 2 500 / 10 000 triangles 0.19 / 0.20 / 0.20 -> 0.08 / 0.08 / 0.12 ms; from
    // 25 000 triangles on a thread per triangle is faster again (0.23 against 0.26 ms).
    const bool tall = total > 0 && (uint64_t)target->Width*(uint64_t)target->Height >= total*64;
    if(tall && c->tall_mode != 0)
        for(MeshParams &mp : ms) { mp.sort_shift = c->tall_shift; mp.row_chunk = c->tall_chunk; }
    if(total > 0 && c->split_mode != 0 && (c->split_mode == 2 || (tall && total <= 16384)))
    {
        int tpc = total >= 6000 ? 16 : 8;
        int rows = total >= 6000 ? 32 : 16;
        if(c->split_tpc > 0) tpc = c->split_tpc;
        if(c->split_rows > 0) rows = c->split_rows;
        rows = std::max(rows, c->tile_h);
        for(MeshParams &mp : ms) { mp.tris_per_cta = tpc; mp.part_rows = rows; }
    }

    const unsigned ntiles = (unsigned)(v.tiles_x*v.tiles_y);
    // first guesses (2.5 segments, 6 spans, 8 queue entries per triangle); all lists grow on demand
    if(c->segs.bytes == 0) CU(c->segs.reserve((size_t)std::max<uint64_t>(total*5/2, 1u << 16)*sizeof(SegInfo)));
    // a frame with a Phong mesh uses the wider span record (normals) and the raster kernel's general
    // variant for all its spans; textured frames without Phong keep the 16-word record (the colour
    // words carry u/z, v/z, 1/z) and take the textured variant
    c->span_words = any_phong ? kSpanWordsPhong : kSpanWords;
    (void)any_tex;
    if(c->spans.bytes == 0) CU(c->spans.reserve((size_t)std::max<uint64_t>(total*6, 1u << 16)*c->span_words*sizeof(uint32_t)));
    // counts, cursors, offsets (+1 end entry), and the scan's chunk scratch
    CU(c->tiles.reserve(((size_t)ntiles*kDepthBuckets*3 + 1 + 2*((size_t)ntiles*kDepthBuckets/8192 + 2))*sizeof(unsigned)));
    if(c->pairs.bytes == 0)
        CU(c->pairs.reserve((size_t)std::max<uint64_t>(total*8, 1u << 16)*sizeof(unsigned)));

    c->view = v;
    c->meshes.swap(ms);
    c->tex_host.swap(texs);
    c->target = *target;
    c->total_tris = (unsigned)total;
    c->ntiles = ntiles;
    c->stats.Triangles = total;
    c->stats.Tiles = ntiles;
    rc = issue_frame(c);
    // The frame's overflow verdict is known once set-up, scan and finalize have run (the raster kernel
    // is still in flight): resolve it now, so that whatever the caller enqueues next on this stream
    // reads a finished frame.  B200R_DEFER_VERDICT leaves it to the next call / b200r_sync.
    if(rc == B200R_OK && !(flags & B200R_DEFER_VERDICT) && !c->host_path) rc = settle_pending(c);
    return rc;
}

int b200r_clear_device(b200r_context *c, const b200r_device_target *t, u32 color, r32 depth)
{
    if(!c || !t || !t->Color || !t->Depth || t->Width <= 0 || t->BandRows <= 0) return fail(c, B200R_E_INVALID, "bad clear target");
    ENTER(c);
    int rc = settle_pending(c);
    if(rc != B200R_OK) return rc;
    launch_clear(t->Color, t->ColorPitch/4, t->Depth, t->DepthStride, t->Width, t->BandRows, color, depth, c->stream);
    c->stats.KernelLaunches += 1;
    CU(cudaGetLastError());
    return B200R_OK;
}

int b200r_set_gather_target(b200r_context *c, const b200r_device_target *g)
{
    if(!c) return B200R_E_INVALID;
    ENTER(c);
    int rc = b200r_sync(c);                      // the last frame may still be writing the old one
    if(rc != B200R_OK) return rc;
    if(!g) { c->has_gather = false; return B200R_OK; }
    if(!g->Color || g->Width <= 0 || g->Height <= 0 || g->ColorPitch < g->Width*4 || (g->ColorPitch & 3) ||
       (g->Depth && g->DepthStride < g->Width))
        return fail(c, B200R_E_INVALID, "bad gather target geometry");
    c->gather = *g;
    c->has_gather = true;
    return B200R_OK;
}

static_assert(sizeof(b200r_peer_handle) == sizeof(cudaIpcMemHandle_t), "b200r_peer_handle carries a cudaIpcMemHandle_t");

int b200r_peer_alloc(b200r_context *c, uint64_t bytes, void **ptr, b200r_peer_handle *handle)
{
    if(!c || !ptr || !handle || bytes == 0) return fail(c, B200R_E_INVALID, "b200r_peer_alloc: null argument");
    ENTER(c);
    void *q = nullptr;
    CU(cudaMalloc(&q, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, q);
    if(e != cudaSuccess) { cudaFree(q); return fail(c, B200R_E_CUDA, "cudaIpcGetMemHandle", e); }
    memcpy(handle->Bytes, &h, sizeof(h));
    c->peer_owned.push_back(q);
    *ptr = q;
    return B200R_OK;
}

int b200r_peer_open(b200r_context *c, const b200r_peer_handle *handle, void **ptr)
{
    if(!c || !ptr || !handle) return fail(c, B200R_E_INVALID, "b200r_peer_open: null argument");
    ENTER(c);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->Bytes, sizeof(h));
    void *q = nullptr;
    CU(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_opened.push_back(q);
    *ptr = q;
    return B200R_OK;
}

int b200r_peer_release(b200r_context *c, void *ptr)
{
    if(!c || !ptr) return B200R_E_INVALID;
    ENTER(c);
    int rc = b200r_sync(c);
    if(rc != B200R_OK) return rc;
    if(c->has_gather && (c->gather.Color == ptr || (void *)c->gather.Depth == ptr)) c->has_gather = false;
    for(size_t i = 0; i < c->peer_opened.size(); ++i)
        if(c->peer_opened[i] == ptr)
        {
            c->peer_opened.erase(c->peer_opened.begin() + (long)i);
            CU(cudaIpcCloseMemHandle(ptr));
            return B200R_OK;
        }
    for(size_t i = 0; i < c->peer_owned.size(); ++i)
        if(c->peer_owned[i] == ptr)
        {
            c->peer_owned.erase(c->peer_owned.begin() + (long)i);
            CU(cudaFree(ptr));
            return B200R_OK;
        }
    return fail(c, B200R_E_INVALID, "b200r_peer_release: not a pointer of this context");
}

int b200r_set_profiling(b200r_context *c, int enable)
{
    if(!c) return B200R_E_INVALID;
    ENTER(c);
    int rc = b200r_sync(c);
    if(rc != B200R_OK) return rc;
    if(enable && !c->stage_ev[0])
        for(cudaEvent_t &e : c->stage_ev) CU(cudaEventCreate(&e));
    c->profiling = enable != 0;
    return B200R_OK;
}

int b200r_get_stage_ms(b200r_context *c, float ms[B200R_STAGES])
{
    if(!c || !ms) return B200R_E_INVALID;
    if(!c->profiling) return fail(c, B200R_E_INVALID, "profiling is not enabled");
    int rc = b200r_sync(c);
    if(rc != B200R_OK) return rc;
    for(int i = 0; i < B200R_STAGES; ++i) CU(cudaEventElapsedTime(&ms[i], c->stage_ev[i], c->stage_ev[i + 1]));
    return B200R_OK;
}

int b200r_get_stats(b200r_context *c, b200r_frame_stats *s)
{
    if(!c || !s) return B200R_E_INVALID;
    *s = c->stats;
    return B200R_OK;
}

// ---------------------------------------------------------------------------- host-pointer path
// Objects become device meshes of at most kUploadChunk triangles each (same P and flags), so that
// the set-up of one chunk can start while the next chunk's colours and normals are still on the bus.
constexpr u32 kUploadChunk = 1u << 17;

static int upload_objects(b200r_context *c, const render_entry_3d_object *objs, u32 n,
                          std::vector<b200r_device_mesh> &meshes)
{
    uint64_t verts = 0;
    size_t chunks = 0;
    bool any_uv = false;
    for(u32 i = 0; i < n; ++i)
    {
        const render_entry_3d_object &o = objs[i];
        u32 tris = o.VertexCount/3;                          // projekt.cpp:3886
        if(tris && (!o.VertexData || !o.ColorData || !o.NormalData)) return fail(c, B200R_E_INVALID, "object with null vertex stream");
        if(o.Bitmap)
        {
            const loaded_bitmap &b = *o.Bitmap;
            if(tris && !o.UVData) return fail(c, B200R_E_INVALID, "textured object without UVData");
            if(!b.Memory || b.Width <= 0 || b.Height <= 0 || b.Pitch < b.Width*4 || (b.Pitch & 3))
                return fail(c, B200R_E_INVALID, "bad Bitmap geometry");
            any_uv = true;
        }
        verts += (uint64_t)tris*3;
        chunks += std::max<size_t>(1, ((size_t)tris + kUploadChunk - 1)/kUploadChunk);
    }
    CU(c->d_pos.reserve((size_t)std::max<uint64_t>(verts, 1)*12));
    CU(c->d_col.reserve((size_t)std::max<uint64_t>(verts, 1)*16));
    CU(c->d_nrm.reserve((size_t)std::max<uint64_t>(verts, 1)*12));
    if(any_uv) CU(c->d_uv.reserve((size_t)std::max<uint64_t>(verts, 1)*8));
    if(!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if(!c->pos_ready) CU(cudaEventCreateWithFlags(&c->pos_ready, cudaEventDisableTiming));
    if(!c->target_ready) CU(cudaEventCreateWithFlags(&c->target_ready, cudaEventDisableTiming));
    while(c->chunk_ready.size() < chunks)
    {
        cudaEvent_t e = nullptr;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_ready.push_back(e);
    }

    // 0. the objects' Bitmaps, one device copy per distinct pointer (tightly packed rows); the raster
    //    kernel is the only reader and waits for the targets, which are enqueued after these
    size_t ntex = 0;
    std::vector<int> tex_of(n, -1);
    for(u32 i = 0; i < n; ++i)
    {
        const loaded_bitmap *b = objs[i].Bitmap;
        if(!b) continue;
        size_t k = 0;
        for(; k < ntex; ++k) if(c->host_textures[k].host == b) break;
        if(k == ntex)
        {
            if(c->host_textures.size() <= ntex) c->host_textures.emplace_back();
            b200r_context::HostTexture &ht = c->host_textures[ntex];
            CU(ht.pixels.reserve((size_t)b->Width*b->Height*4));
            ht.host = b;
            ht.desc.Memory = (const u32 *)ht.pixels.ptr; ht.desc.Width = b->Width; ht.desc.Height = b->Height;
            ht.desc.Pitch = b->Width*4;
            CU(cudaMemcpy2DAsync(ht.pixels.ptr, (size_t)b->Width*4, b->Memory, (size_t)b->Pitch, (size_t)b->Width*4,
                                 b->Height, cudaMemcpyHostToDevice, c->copy_stream));
            ++ntex;
        }
        tex_of[i] = (int)k;
    }

    // 1. every position (the z-range pass reads them all before any set-up can start)
    uint64_t at = 0;
    for(u32 i = 0; i < n; ++i)
    {
        const render_entry_3d_object &o = objs[i];
        const size_t nv = (size_t)(o.VertexCount/3)*3;
        if(nv) CU(cudaMemcpyAsync((r32 *)c->d_pos.ptr + at*3, o.VertexData, nv*12, cudaMemcpyHostToDevice, c->copy_stream));
        at += nv;
    }
    CU(cudaEventRecord(c->pos_ready, c->copy_stream));

    // 2. colours and normals chunk by chunk, one event per chunk
    at = 0;
    for(u32 i = 0; i < n; ++i)
    {
        const render_entry_3d_object &o = objs[i];
        const u32 tris = o.VertexCount/3;
        u32 done = 0;
        do
        {
            const u32 ct = std::min(tris - done, kUploadChunk);
            const size_t nv = (size_t)ct*3, v0 = (size_t)done*3;
            b200r_device_mesh m;
            m.Positions = (const r32 *)c->d_pos.ptr + at*3;
            m.Colors = (const r32 *)c->d_col.ptr + at*4;
            m.Normals = (const r32 *)c->d_nrm.ptr + at*3;
            m.TriangleCount = ct;
            m.P = o.P;
            m.Flags = o.PhongShading ? B200R_MESH_PHONG : 0u;
            m.UVs = nullptr; m.Texture = nullptr;
            if(tex_of[i] >= 0)
            {
                m.UVs = (const r32 *)c->d_uv.ptr + at*2;
                m.Texture = &c->host_textures[tex_of[i]].desc;
            }
            if(nv)
            {
                // a textured object's vertex colours never reach the image (MeshParams::uv): not uploaded
                if(tex_of[i] < 0)
                    CU(cudaMemcpyAsync((void *)m.Colors, (const r32 *)o.ColorData + v0*4, nv*16, cudaMemcpyHostToDevice, c->copy_stream));
                else
                    CU(cudaMemcpyAsync((void *)m.UVs, (const r32 *)o.UVData + v0*2, nv*8, cudaMemcpyHostToDevice, c->copy_stream));
                CU(cudaMemcpyAsync((void *)m.Normals, (const r32 *)o.NormalData + v0*3, nv*12, cudaMemcpyHostToDevice, c->copy_stream));
            }
            CU(cudaEventRecord(c->chunk_ready[meshes.size()], c->copy_stream));
            at += nv;
            done += ct;
            meshes.push_back(m);
        } while(done < tris);
    }
    return B200R_OK;
}

static int render_objects_whole(b200r_context *c, const render_entry_3d_object *objs, u32 n,
                                const game_render_commands *cmd, const loaded_bitmap *out);

int b200r_render_objects(b200r_context *c, const render_entry_3d_object *objs, u32 n,
                         const game_render_commands *cmd, const loaded_bitmap *out, u32 flags)
{
    if(!c) return B200R_E_INVALID;
    if(!cmd || !out || !out->Memory || !cmd->ZBuffer || (n && !objs)) return fail(c, B200R_E_INVALID, "null argument");
    if(out->Width <= 0 || out->Height <= 0 || out->Pitch < out->Width*4 || cmd->Width < (u32)out->Width)
        return fail(c, B200R_E_INVALID, "bad OutputTarget / Commands->Width");
    ENTER(c);
    int rc = settle_pending(c);
    if(rc != B200R_OK) return rc;
    if((flags & B200R_WHOLE_OBJECT_AEL) && (flags & (B200R_AVX_RIGHT_END_EXCLUSIVE | B200R_AVX_DEPTH_GE)))
        return fail(c, B200R_E_UNSUPPORTED, "the AVX compatibility switches apply to the per-triangle mode only");
    if(flags & B200R_WHOLE_OBJECT_AEL) return render_objects_whole(c, objs, n, cmd, out);

    std::vector<b200r_device_mesh> meshes;
    rc = upload_objects(c, objs, n, meshes);
    if(rc != B200R_OK)
    {
        if(c->copy_stream) cudaStreamSynchronize(c->copy_stream);   // copies already enqueued read the caller's memory
        return rc;
    }

    // device mirrors of the targets: rows padded to 64 pixels so every tile row is a 16-byte
    // aligned bulk copy
    const int W = out->Width, H = out->Height;
    const int wpad = (W + 63) & ~63;
    CU(c->d_color.reserve((size_t)wpad*H*4));
    CU(c->d_depth.reserve((size_t)wpad*H*4));
    // 3. the targets, last: only the raster kernel reads them -- band by band of tile rows, so that
    //    the first band can be rastered and read back while the others are still being uploaded
    {
        uint64_t tris = 0;
        for(const b200r_device_mesh &m : meshes) tris += m.TriangleCount;
        choose_tile(c, tris, W, H);             // the same answer b200r_render_device gets below
    }
    const int tiles_y = (H + c->tile_h - 1)/c->tile_h;
    c->host_bands = (tiles_y >= 2*kHostBands) ? kHostBands : 1;
    c->host_out.color = out->Memory; c->host_out.color_pitch = (size_t)out->Pitch;
    c->host_out.depth = cmd->ZBuffer; c->host_out.depth_pitch = (size_t)cmd->Width*4;
    c->host_out.W = W; c->host_out.H = H; c->host_out.wpad = wpad;
    if(!c->d2h_stream) CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    for(int b = 0; b < c->host_bands; ++b)
    {
        if(!c->band_ready[b]) CU(cudaEventCreateWithFlags(&c->band_ready[b], cudaEventDisableTiming));
        if(!c->band_done[b]) CU(cudaEventCreateWithFlags(&c->band_done[b], cudaEventDisableTiming));
        const int y0 = std::min(tiles_y*b/c->host_bands*c->tile_h, H), y1 = std::min(tiles_y*(b + 1)/c->host_bands*c->tile_h, H);
        if(y1 > y0)
        {
            CU(cudaMemcpy2DAsync((char *)c->d_color.ptr + (size_t)y0*wpad*4, (size_t)wpad*4,
                                 (const char *)out->Memory + (size_t)y0*out->Pitch, (size_t)out->Pitch, (size_t)W*4, (size_t)(y1 - y0),
                                 cudaMemcpyHostToDevice, c->copy_stream));
            CU(cudaMemcpy2DAsync((char *)c->d_depth.ptr + (size_t)y0*wpad*4, (size_t)wpad*4,
                                 (const char *)cmd->ZBuffer + (size_t)y0*cmd->Width*4, (size_t)cmd->Width*4, (size_t)W*4, (size_t)(y1 - y0),
                                 cudaMemcpyHostToDevice, c->copy_stream));
        }
        CU(cudaEventRecord(c->band_ready[b], c->copy_stream));
    }
    CU(cudaEventRecord(c->target_ready, c->copy_stream));
    b200r_device_target t;
    t.Color = (u32 *)c->d_color.ptr; t.Depth = (r32 *)c->d_depth.ptr;
    t.Width = W; t.Height = H; t.ColorPitch = wpad*4; t.DepthStride = wpad;
    t.BandFirstRow = 0; t.BandRows = H;
    c->host_path = true;
    c->host_alias_rows = (out->Pitch == W*4 && cmd->Width == (u32)W) ? 1 : 0;
    rc = b200r_render_device(c, meshes.data(), (u32)meshes.size(), cmd, &t, flags);
    if(rc == B200R_OK) rc = settle_pending(c);  // re-issue before the read-back if the list grew
    c->host_path = false;
    c->host_alias_rows = -1;
    if(rc != B200R_OK)
    {
        cudaStreamSynchronize(c->copy_stream);  // the caller may free its buffers once we return
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->d2h_stream);
        return rc;
    }
    if(c->host_bands > 1)
    {
        CU(cudaStreamSynchronize(c->d2h_stream));       // every band was read back as it finished
        CU(cudaStreamSynchronize(c->stream));
        return B200R_OK;
    }
    CU(cudaMemcpy2DAsync(out->Memory, (size_t)out->Pitch, c->d_color.ptr, (size_t)wpad*4, (size_t)W*4, H,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpy2DAsync(cmd->ZBuffer, (size_t)cmd->Width*4, c->d_depth.ptr, (size_t)wpad*4, (size_t)W*4, H,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200R_OK;
}

// FillEdgeTable of one object (projekt.cpp:3882) into outp (room for VertexCount records), in the
// reference's MergeSort order (projekt.cpp:2-72; reproduced on the device, edge_table_kernels.cu).  Every
// field of a record is written: the ones the reference leaves untouched for this kind of object (texture
// fields of an untextured object, normals of a Gouraud object, Next) are zero.  Returns the edge count or a
// negative status.
static int build_edge_table(b200r_context *c, const render_entry_3d_object *obj,
                            const game_render_commands *cmd, b32 phong, edge_info *outp)
{
    const bool textured = obj->Bitmap != nullptr;
    if(textured && obj->VertexCount >= 3 && !obj->UVData) return fail(c, B200R_E_INVALID, "textured object without UVData");
    ENTER(c);
    int rc = settle_pending(c);
    if(rc != B200R_OK) return rc;
    // A textured object takes two passes of the set-up kernel: colours (lit white on the Gouraud
    // path, projekt.cpp:4034-4060; the vertex colours on the Phong path, :4014) and then u/z, v/z,
    // 1/z, which travel in the colour words (MeshParams::uv).  The first pass is an untextured upload.
    render_entry_3d_object plain = *obj;
    plain.Bitmap = nullptr;
    std::vector<b200r_device_mesh> meshes;
    rc = upload_objects(c, &plain, 1, meshes);
    if(rc != B200R_OK) return rc;
    CU(cudaStreamWaitEvent(c->stream, c->pos_ready, 0));
    CU(cudaStreamWaitEvent(c->stream, c->chunk_ready[meshes.size() - 1], 0));
    // the upload is chunked (one mesh per 131 072 triangles) but contiguous: one set-up pass over the whole object
    const u32 tris = obj->VertexCount/3;
    if(tris == 0) return 0;
    if(tris > kScanMaxChunks*8192u) return fail(c, B200R_E_UNSUPPORTED, "b200r_fill_edge_table: more than 16 M triangles in one object");

    // Only the set-up kernel runs.  Height only clamps MaxY in the record header (unused here);
    // a 1x1 dummy band keeps the tile bookkeeping trivial.
    b200r_device_target t;
    memset(&t, 0, sizeof(t));
    t.Width = 1 << 17; t.Height = 1 << 20; t.BandFirstRow = 0; t.BandRows = 1;
    t.ColorPitch = t.Width*4; t.DepthStride = t.Width;
    t.Color = (u32 *)16; t.Depth = (r32 *)16;    // never dereferenced: raster is not launched
    ViewParams v;
    rc = fill_view(c, cmd, &t, v, phong != 0);
    if(rc != B200R_OK) return rc;
    v.tiles_x = 1; v.tiles_y = 1; v.tile_w = 1 << 21; v.tile_h = 1 << 21; v.tile_w_shift = v.tile_h_shift = 21;
    CU(c->recs.reserve((size_t)tris*kRecWords*sizeof(uint32_t)));
    CU(c->tiles.reserve(((size_t)kDepthBuckets*3 + 1)*sizeof(unsigned)));
    unsigned *tile_count = (unsigned *)c->tiles.ptr;
    FrameWords *words = (FrameWords *)c->words.ptr;
    CU(cudaMemsetAsync(tile_count, 0, ((size_t)kDepthBuckets*3 + 1)*sizeof(unsigned), c->stream));
    CU(cudaMemsetAsync(words, 0, sizeof(FrameWords), c->stream));
    SetupOutputs so;
    so.recs = (uint32_t *)c->recs.ptr; so.spans = nullptr; so.segs = nullptr;
    so.seg_fill = words->seg_fill; so.span_fill = words->span_fill; so.extra_total = &words->extra_total;
    so.seg_capacity = 0; so.span_capacity = 0; so.span_words = kSpanWords;
    so.tile_count = tile_count; so.counters = words->counters;
    so.zrange = reinterpret_cast<const float *>(words->zkeys);
    MeshParams mp;
    mp.pos = meshes[0].Positions; mp.col = meshes[0].Colors; mp.nrm = meshes[0].Normals;
    mp.ntri = tris; mp.px = obj->P.x; mp.py = obj->P.y; mp.pz = obj->P.z; mp.prim_base = 0;
    mp.phong = phong ? 1 : 0;
    mp.uv = nullptr; mp.tex = -1; mp.white = (textured && !phong) ? 1 : 0;
    mp.tri_list = nullptr; mp.tri_count = nullptr;
    mp.tris_per_cta = 0; mp.part_rows = 0; mp.sort_shift = 0; mp.row_chunk = 0x7fffffff;
    launch_setup(v, mp, so, c->stream);
    c->stats.KernelLaunches += 1;
    CU(cudaGetLastError());
    // FillEdgeTable's tail on the device (edge_table_kernels.cu): the emission index of every edge (a scan
    // of the per-triangle edge counts), MergeSort's order (projekt.cpp:2-72) as a sort on unique keys, and
    // the edge_info records themselves; only the finished table crosses the bus.
    const uint32_t *d_uvrecs = nullptr;
    if(textured)
    {
        CU(c->recs_uv.reserve((size_t)tris*kRecWords*sizeof(uint32_t)));
        CU(c->d_uv.reserve((size_t)tris*3*8));
        CU(cudaMemcpyAsync(c->d_uv.ptr, obj->UVData, (size_t)tris*3*8, cudaMemcpyHostToDevice, c->stream));
        mp.uv = (const float *)c->d_uv.ptr; mp.tex = 0; mp.white = 0;
        so.recs = (uint32_t *)c->recs_uv.ptr;
        launch_setup(v, mp, so, c->stream);
        c->stats.KernelLaunches += 1;
        CU(cudaGetLastError());
        d_uvrecs = (const uint32_t *)c->recs_uv.ptr;
    }
    CU(c->et_counts.reserve(((size_t)tris*2 + 2)*sizeof(unsigned)));
    unsigned *counts = (unsigned *)c->et_counts.ptr, *offsets = counts + tris;      // tris + 1 offsets
    launch_edge_counts((const uint32_t *)c->recs.ptr, tris, counts, c->stream);
    launch_tile_scan(counts, offsets, tris, &words->pair_total, words->scan_state, &words->scan_ticket, c->stream);
    c->stats.KernelLaunches += 2;
    unsigned n = 0;
    CU(cudaMemcpyAsync(&n, &words->pair_total, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if(n > obj->VertexCount) return fail(c, B200R_E_CUDA, "edge count exceeds VertexCount");    // cannot happen: <= 3 per triangle
    if(n == 0) return 0;
    CU(c->et_keys.reserve((size_t)n*2*sizeof(unsigned long long)));
    CU(c->et_vals.reserve((size_t)n*2*sizeof(unsigned)));
    CU(c->et_out.reserve((size_t)n*sizeof(edge_info)));
    unsigned long long *keys[2] = { (unsigned long long *)c->et_keys.ptr, (unsigned long long *)c->et_keys.ptr + n };
    unsigned *vals[2] = { (unsigned *)c->et_vals.ptr, (unsigned *)c->et_vals.ptr + n };
    launch_edge_keys((const uint32_t *)c->recs.ptr, tris, offsets, &words->pair_total, keys[0], vals[0], c->stream);
    const int at = launch_edge_sort(keys, vals, n, &words->pair_total, c->stream);
    launch_edge_assemble((const uint32_t *)c->recs.ptr, d_uvrecs, phong ? (const float *)mp.nrm : nullptr, vals[at], n,
                         &words->pair_total, c->et_out.ptr, c->stream);
    c->stats.KernelLaunches += 3;
    for(unsigned long long run = 2048; run < n; run <<= 1) c->stats.KernelLaunches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(outp, c->et_out.ptr, (size_t)n*sizeof(edge_info), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return (int)n;
}

int b200r_fill_edge_table(b200r_context *c, const render_entry_3d_object *obj,
                          const game_render_commands *cmd, b32 phong)
{
    if(!c) return B200R_E_INVALID;
    if(!obj || !cmd) return fail(c, B200R_E_INVALID, "null argument");
    if(!obj->EdgeMemory) return fail(c, B200R_E_INVALID, "null EdgeMemory");
    return build_edge_table(c, obj, cmd, phong, (edge_info *)obj->EdgeMemory);
}

// ---------------------------------------------------------------------------- whole-object mode
// b200r_render_objects with B200R_WHOLE_OBJECT_AEL: per object, the sorted edge_info array (the
// same records b200r_fill_edge_table returns) goes to the device and ONE thread replays DrawModel's
// active-edge list on it (object_walk_kernel.cu); binning and raster are the usual kernels.
static int render_objects_whole(b200r_context *c, const render_entry_3d_object *objs, u32 n,
                                const game_render_commands *cmd, const loaded_bitmap *out)
{
    const int W = out->Width, H = out->Height;
    std::vector<edge_info> all;
    std::vector<ObjectDesc> descs;
    std::vector<TexDesc> texs;
    std::vector<edge_info> scratch;
    uint64_t slots = 0;
    bool general = false, any_phong_obj = false;
    size_t ntex = 0;
    std::vector<unsigned> chain_base;
    uint64_t chain_T = 0;
    unsigned max_edges = 0, max_bound = 0;
    for(u32 i = 0; i < n; ++i)
    {
        const render_entry_3d_object &o = objs[i];
        if(o.VertexCount < 3) continue;
        scratch.assign(o.VertexCount, edge_info());
        const int ne = build_edge_table(c, &o, cmd, o.PhongShading, scratch.data());
        if(ne < 0) return ne;
        ObjectDesc d;
        d.first_edge = (unsigned)all.size(); d.edge_count = (unsigned)ne;
        d.phong = o.PhongShading ? 1 : 0; d.tex = -1;
        // projekt.cpp:176-196: one past the object's last row
        int max_row = ne ? scratch[0].YMax : 0;
        for(int e = 1; e < ne; ++e) max_row = std::max<int>(max_row, scratch[e].YMax);
        d.max_y = std::min(max_row, H);
        // value chains: rows + 1 entries per edge (object_walk_kernel.cu, chain_entries)
        d.chain_first = (unsigned)chain_T;
        for(int e = 0; e < ne; ++e)
        {
            const int64_t rows = (int64_t)std::min<int>(scratch[e].YMax, d.max_y) - scratch[e].YMin;
            chain_base.push_back((unsigned)chain_T);
            chain_T += (uint64_t)std::max<int64_t>(rows, 0) + 1;
        }
        d.chain_total = (unsigned)(chain_T - d.chain_first);
        max_edges = std::max<unsigned>(max_edges, (unsigned)ne);
        // every pair consumes one row of two active edges: half the edge rows bound the spans
        uint64_t edge_rows = 0;
        for(int e = 0; e < ne; ++e)
        {
            scratch[e].Next = (edge_info *)(intptr_t)-1;           // an index on the device, -1 = null
            const int64_t y0 = std::max<int64_t>(scratch[e].YMin, 0), y1 = std::min<int64_t>(scratch[e].YMax, H);
            if(y1 > y0) edge_rows += (uint64_t)(y1 - y0);
        }
        const uint64_t bound = edge_rows/2 + 1;
        if(slots + bound > 0x7fffffffull) return fail(c, B200R_E_UNSUPPORTED, "more than 2^31-1 spans per call");
        d.span_base = d.prim_base = (unsigned)slots; d.span_bound = (unsigned)bound;
        slots += bound;
        max_bound = std::max<unsigned>(max_bound, (unsigned)bound);
        if(o.Bitmap)
        {
            const loaded_bitmap *b = o.Bitmap;
            size_t k = 0;
            for(; k < ntex; ++k) if(c->host_textures[k].host == b) break;
            if(k == ntex)
            {
                if(k >= 65536) return fail(c, B200R_E_UNSUPPORTED, "more than 65536 distinct textures per call");
                if(c->host_textures.size() <= ntex) c->host_textures.emplace_back();
                b200r_context::HostTexture &ht = c->host_textures[ntex];
                CU(ht.pixels.reserve((size_t)b->Width*b->Height*4));
                ht.host = b;
                CU(cudaMemcpy2DAsync(ht.pixels.ptr, (size_t)b->Width*4, b->Memory, (size_t)b->Pitch, (size_t)b->Width*4,
                                     b->Height, cudaMemcpyHostToDevice, c->stream));
                TexDesc td; td.mem = (const uint32_t *)ht.pixels.ptr; td.w = b->Width; td.h = b->Height; td.pitch = b->Width*4;
                texs.push_back(td);
                ++ntex;
            }
            d.tex = (int)k;
        }
        general |= d.phong != 0 || d.tex >= 0;
        any_phong_obj |= d.phong != 0;
        all.insert(all.end(), scratch.begin(), scratch.begin() + ne);
        descs.push_back(d);
        // chain and emit kernels put the objects on gridDim.y
        if(descs.size() > 65535) return fail(c, B200R_E_UNSUPPORTED, "whole-object mode: more than 65535 objects per call");
    }

    // targets (as in the per-triangle call): device mirrors with rows padded to 64 pixels
    const int wpad = (W + 63) & ~63;
    CU(c->d_color.reserve((size_t)wpad*H*4));
    CU(c->d_depth.reserve((size_t)wpad*H*4));
    CU(cudaMemcpy2DAsync(c->d_color.ptr, (size_t)wpad*4, out->Memory, (size_t)out->Pitch, (size_t)W*4, H,
                         cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpy2DAsync(c->d_depth.ptr, (size_t)wpad*4, cmd->ZBuffer, (size_t)cmd->Width*4, (size_t)W*4, H,
                         cudaMemcpyHostToDevice, c->stream));
    b200r_device_target t;
    t.Color = (u32 *)c->d_color.ptr; t.Depth = (r32 *)c->d_depth.ptr;
    t.Width = W; t.Height = H; t.ColorPitch = wpad*4; t.DepthStride = wpad;
    t.BandFirstRow = 0; t.BandRows = H;
    ViewParams v;
    c->host_alias_rows = (out->Pitch == W*4 && cmd->Width == (u32)W) ? 1 : 0;
    int rc = fill_view(c, cmd, &t, v, true);            // light-count rules were applied per object above
    c->host_alias_rows = -1;
    if(rc != B200R_OK) return rc;

    if(!descs.empty())
    {
        c->edges_bytes = all.size()*sizeof(edge_info);
        CU(c->edges_pristine.reserve(std::max<size_t>(c->edges_bytes, 16)));
        CU(c->edges_work.reserve(std::max<size_t>(all.size()*10*sizeof(float), 16)));     // WalkState scratch
        CU(c->obj_dev.reserve(descs.size()*sizeof(ObjectDesc)));
        if(c->edges_bytes) CU(cudaMemcpyAsync(c->edges_pristine.ptr, all.data(), c->edges_bytes, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->obj_dev.ptr, descs.data(), descs.size()*sizeof(ObjectDesc), cudaMemcpyHostToDevice, c->stream));
        // the chain arrays are x | z | c[4] | n[3] over chain_T entries and c is read with 128-bit loads:
        // every section must start 16-byte aligned
        chain_T = (chain_T + 3) & ~(uint64_t)3;
        // three-phase path unless the value chains would be unreasonably large (9 words per edge row)
        c->obj_three_phase = !c->obj_force_serial && chain_T > 0 && chain_T*9*sizeof(float) <= (1ull << 30) &&
                             chain_T < 0xffffffffull;
        if(c->obj_three_phase)
        {
            CU(c->obj_chain_base.reserve(chain_base.size()*sizeof(unsigned)));
            CU(c->obj_chains.reserve((size_t)chain_T*9*sizeof(float) + 64));
            CU(c->obj_pairs.reserve((size_t)slots*sizeof(uint4)));
            CU(c->obj_flags.reserve(descs.size()*2*sizeof(unsigned)));
            CU(cudaMemcpyAsync(c->obj_chain_base.ptr, chain_base.data(), chain_base.size()*sizeof(unsigned),
                               cudaMemcpyHostToDevice, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));           // `all`, `descs`, `chain_base` are stack vectors
    }
    const unsigned ntiles = (unsigned)(v.tiles_x*v.tiles_y);
    c->span_words = any_phong_obj ? kSpanWordsPhong : kSpanWords;
    (void)general;
    // exact needs are known: the promised slots, striped over the regions, plus room for alias pixels
    const size_t region = (size_t)(slots/kSubAllocators) + 2 + 1024;
    CU(c->segs.reserve(region*kSubAllocators*sizeof(SegInfo)));
    CU(c->spans.reserve(region*kSubAllocators*c->span_words*sizeof(uint32_t)));
    CU(c->tiles.reserve(((size_t)ntiles*kDepthBuckets*3 + 1 + 2*((size_t)ntiles*kDepthBuckets/8192 + 2))*sizeof(unsigned)));
    if(c->pairs.bytes == 0) CU(c->pairs.reserve((size_t)std::max<uint64_t>(slots*2, 1u << 16)*sizeof(unsigned)));

    c->view = v;
    c->meshes.clear();
    c->tex_host.swap(texs);
    c->target = t;
    c->total_tris = 0;
    c->ntiles = ntiles;
    c->obj_host.swap(descs);
    c->obj_total_slots = (unsigned)slots;
    c->obj_smem_bytes = 0;
    c->obj_order_smem = 0;
    c->obj_chain_T = (unsigned)chain_T; c->obj_max_bound = max_bound; c->obj_max_edges = max_edges;
    for(const ObjectDesc &d : c->obj_host)
    {
        const size_t need = (size_t)d.edge_count*(d.phong ? 10 : 7)*sizeof(float);
        if(need <= 200u*1024u) c->obj_smem_bytes = std::max<unsigned>(c->obj_smem_bytes, (unsigned)need);
        // order phase: links + step counts, and the x chains when they fit too
        const size_t links = (size_t)d.edge_count*4*sizeof(int), all_ = links + (size_t)d.chain_total*sizeof(float);
        if(all_ <= 200u*1024u) c->obj_order_smem = std::max<unsigned>(c->obj_order_smem, (unsigned)all_);
        else if(links <= 200u*1024u) c->obj_order_smem = std::max<unsigned>(c->obj_order_smem, (unsigned)links);
    }
    c->stats.Triangles = 0;
    for(u32 i = 0; i < n; ++i) c->stats.Triangles += objs[i].VertexCount/3;
    c->stats.Tiles = ntiles;
    c->object_mode = true;
    rc = issue_frame(c);
    if(rc == B200R_OK) rc = settle_pending(c);
    c->object_mode = false;
    if(rc != B200R_OK) { cudaStreamSynchronize(c->stream); return rc; }
    CU(cudaMemcpy2DAsync(out->Memory, (size_t)out->Pitch, c->d_color.ptr, (size_t)wpad*4, (size_t)W*4, H,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpy2DAsync(cmd->ZBuffer, (size_t)cmd->Width*4, c->d_depth.ptr, (size_t)wpad*4, (size_t)W*4, H,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200R_OK;
}

} // extern "C"
