// Sort-middle binner (sm_100a): per-tile, per-depth-bucket queues of spans, built from the segments
// (runs of a triangle's spans inside one tile-row band, with their exact tile columns) the set-up
// kernel -- or, in whole-object mode, the emit kernel -- wrote.  The reference has no binning (its nearest analogue is MergeSort by YMin plus
// one work item per scan line, projekt.cpp:2-72, 3509-3609); this stage exists because the
// raster kernel keeps a screen tile on chip.
//
//   count    (inside setup_kernel)  tile_count[tile][bucket] += rows, per tile column a segment touches
//   scan     tile_scan_kernel       exclusive prefix sum over the bins -> tile_offset, pair_total
//   finalize finalize_kernel        one overflow verdict for the frame (lists too small: skip and re-issue)
//   scatter  scatter_kernel         queue[tile_offset[bin] + slot .. + rows) = the segment's span indices
//
// List order inside a bin is NOT submission order (slots are handed out by atomics): the
// raster kernel resolves depth with the order-independent rule
//   z > zold || (z == zold && prim < primold)
// which equals the reference's "strict >, first submitted wins" (projekt.cpp:525), so the
// image does not depend on bin order.
#include "raster_device.cuh"

namespace b200r {

constexpr int kScanThreads = 1024;

// Single-CTA scan: each thread owns a contiguous chunk; chunk sums are scanned with warp
// shuffles (__shfl_up_sync) and one cross-warp step in shared memory.
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(const unsigned *__restrict__ count, unsigned *__restrict__ offset, unsigned n,
                 unsigned *__restrict__ total)
{
    __shared__ unsigned warp_sums[kScanThreads/32];
    const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned chunk = (n + kScanThreads - 1)/kScanThreads;
    const unsigned lo = min(t*chunk, n), hi = min(lo + chunk, n);
    unsigned sum = 0;
    for(unsigned i = lo; i < hi; ++i) sum += count[i];
    unsigned incl = sum;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1)
    {
        unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
        if(lane >= (unsigned)d) incl += up;
    }
    if(lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if(warp == 0)
    {
        unsigned w = warp_sums[lane];
        unsigned wi = w;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1)
        {
            unsigned up = __shfl_up_sync(0xffffffffu, wi, d);
            if(lane >= (unsigned)d) wi += up;
        }
        warp_sums[lane] = wi - w;                       // exclusive
        if(lane == 31) *total = wi;
    }
    __syncthreads();
    unsigned run = warp_sums[warp] + (incl - sum);
    for(unsigned i = lo; i < hi; ++i) { offset[i] = run; run += count[i]; }
    if(hi == n && lo <= n) offset[n] = run;             // end entry (several threads may write the same total)
}

// One lane per segment; per tile column the segment touches, ONE atomic reserves nrows consecutive slots of
// that tile's (per-bucket) list, which are filled with the segment's span indices, so the raster kernel's
// queue is a flat list of spans while the number of atomics stays per segment.
//
// The fill is done by the warp together: lanes push their runs {first slot, first span, rows} into a small
// per-warp queue in shared memory and groups of 8 lanes write one run each -- consecutive entries, whole
// sectors.  The first version had every lane write its own runs, one 4-byte store per entry: a warp-wide
// store instruction touched 32 different sectors, and the kernel (13 % issue utilisation) was bound by those
// store transactions, not by the atomics.  Measured: C2 0.063 -> 0.050 ms, C3 0.075 -> 0.050 ms, C4 (20 M
// triangles, 16384^2) 3.96 -> 2.09 ms; frames of a few thousand segments lose 2 us to the queue.
#ifndef B200R_SCATTER_COOP
#define B200R_SCATTER_COOP 1
#endif
constexpr int kScatterThreads = 256;
constexpr int kRunQueue = 128;           // runs a warp collects before its lanes write them out

__global__ void __launch_bounds__(kScatterThreads)
scatter_kernel(const ScatterParams p)
{
    if(*p.overflow) return;                              // host grows the lists and re-issues the frame
    // blockIdx.y = region of the segment array; the alias-pixel segments sit at the very end
    const unsigned region = blockIdx.y;
    const unsigned region_size = p.seg_capacity/kSubAllocators;
    const unsigned nseg = p.seg_fill[region];
    const unsigned nextra = (region == kSubAllocators - 1) ? *p.extra_total : 0u;
#if B200R_SCATTER_COOP
    __shared__ unsigned s_slot[kScatterThreads/32][kRunQueue], s_base[kScatterThreads/32][kRunQueue];
    __shared__ unsigned char s_rows[kScatterThreads/32][kRunQueue];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned qn = 0;                                     // runs in this warp's queue (warp-uniform)
    auto drain = [&]()
    {
        __syncwarp();
        for(unsigned q = lane >> 3; q < qn; q += 4)
        {
            const unsigned slot = s_slot[warp][q], base = s_base[warp][q], rows = s_rows[warp][q];
            for(unsigned r = lane & 7u; r < rows; r += 8) p.pair_list[slot + r] = base + r;
        }
        __syncwarp();
        qn = 0;
    };
    // every lane of a warp runs the same number of iterations (whole warps of consecutive segments)
    for(unsigned i0 = blockIdx.x*blockDim.x + warp*32u; i0 < nseg + nextra; i0 += gridDim.x*blockDim.x)
    {
        const unsigned i = i0 + lane;
        SegInfo si; si.tile_row = 0; si.tx = 1u; si.span_base = 0; si.nrows = 0;       // tx0 > tx1: touches no tile
        if(i < nseg + nextra) si = p.segs[(i < nseg) ? region*region_size + i : p.seg_capacity - 1u - (i - nseg)];
        const int tx0 = si.tx & 0xffff, tx1 = si.tx >> 16;
        const unsigned trow = si.tile_row & 0xffffffu, bucket = si.tile_row >> 24;
        const int ncols = (tx1 >= tx0 && si.nrows) ? tx1 - tx0 + 1 : 0;
        const int maxcols = __reduce_max_sync(0xffffffffu, ncols);
        for(int c = 0; c < maxcols; ++c)
        {
            const bool has = c < ncols;
            unsigned slot = 0;
            if(has)
            {
                const unsigned tile = (trow*(unsigned)p.tiles_x + (unsigned)(tx0 + c))*kDepthBuckets + bucket;
                slot = p.tile_offset[tile] + atomicAdd(&p.tile_fill[tile], si.nrows);
                B200R_ASSERT(slot + si.nrows <= p.tile_offset[tile + 1] && slot + si.nrows <= p.pair_capacity);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if(has)
            {
                const unsigned q = qn + (unsigned)__popc(bal & ((1u << lane) - 1u));
                B200R_ASSERT(q < (unsigned)kRunQueue && si.nrows <= 255u);
                s_slot[warp][q] = slot; s_base[warp][q] = si.span_base; s_rows[warp][q] = (unsigned char)si.nrows;
            }
            qn += (unsigned)__popc(bal);
            if(qn > (unsigned)kRunQueue - 32u) drain();
        }
    }
    drain();
#else
    for(unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < nseg + nextra; i += gridDim.x*blockDim.x)
    {
        const unsigned seg = (i < nseg) ? region*region_size + i : p.seg_capacity - 1u - (i - nseg);
        const SegInfo si = p.segs[seg];
        const int tx0 = si.tx & 0xffff, tx1 = si.tx >> 16;
        const unsigned trow = si.tile_row & 0xffffffu, bucket = si.tile_row >> 24;
        for(int tx = tx0; tx <= tx1; ++tx)
        {
            const unsigned tile = (trow*(unsigned)p.tiles_x + (unsigned)tx)*kDepthBuckets + bucket;
            const unsigned slot = p.tile_offset[tile] + atomicAdd(&p.tile_fill[tile], si.nrows);
            B200R_ASSERT(slot + si.nrows <= p.tile_offset[tile + 1] && slot + si.nrows <= p.pair_capacity);
            for(unsigned r = 0; r < si.nrows; ++r) p.pair_list[slot + r] = si.span_base + r;
        }
    }
#endif
}

__global__ void finalize_kernel(const FinalizeParams p)
{
    // one warp: largest region fills, then the overflow verdict every later kernel obeys
    const unsigned lane = threadIdx.x;
    unsigned g = 0, s = 0;
    for(unsigned r = lane; r < (unsigned)kSubAllocators; r += 32) { g = max(g, p.seg_fill[r]); s = max(s, p.span_fill[r]); }
    g = __reduce_max_sync(0xffffffffu, g);
    s = __reduce_max_sync(0xffffffffu, s);
    if(lane == 0)
    {
        const unsigned nextra = *p.extra_total;
        const unsigned seg_region = p.seg_capacity/kSubAllocators, span_region = p.span_capacity/kSubAllocators;
        const unsigned long long last_g = (unsigned long long)p.seg_fill[kSubAllocators - 1] + nextra;
        const unsigned long long last_s = (unsigned long long)p.span_fill[kSubAllocators - 1] + nextra;
        *p.seg_max = (unsigned)min(max((unsigned long long)g, last_g), 0xffffffffull);
        *p.span_max = (unsigned)min(max((unsigned long long)s, last_s), 0xffffffffull);
        *p.overflow = (g > seg_region || s > span_region || last_g > seg_region || last_s > span_region ||
                       *p.pair_total > p.pair_capacity) ? 1u : 0u;
    }
}

// More than one chunk of bins: ONE launch, a chained scan with decoupled look-back.  CTAs take chunk
// numbers from a ticket (so every chunk's predecessors are running or done), scan their 8192 bins
// locally, publish their sum, and add up the published sums / prefixes of the chunks before them.
// state[c] = flag << 32 | value, flag 1: the chunk's own sum, flag 2: the inclusive prefix up to it.
constexpr unsigned kChunk = kScanThreads*8;

__global__ void __launch_bounds__(kScanThreads)
lookback_scan_kernel(const unsigned *__restrict__ count, unsigned *__restrict__ offset, unsigned n,
                     unsigned *__restrict__ total, unsigned long long *state, unsigned *ticket)
{
    __shared__ unsigned warp_sums[kScanThreads/32];
    __shared__ unsigned s_chunk, s_prefix;
    const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if(t == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned chunk = s_chunk;
    const unsigned lo = min(chunk*kChunk + t*8, n), hi = min(lo + 8, n);
    unsigned v[8], sum = 0;
    if(hi - lo == 8 && (lo & 3u) == 0)
    {
        const uint4 a = *reinterpret_cast<const uint4 *>(count + lo), b = *reinterpret_cast<const uint4 *>(count + lo + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    else
    {
#pragma unroll
        for(int i = 0; i < 8; ++i) v[i] = (lo + i < hi) ? count[lo + i] : 0u;
    }
#pragma unroll
    for(int i = 0; i < 8; ++i) sum += v[i];
    unsigned incl = sum;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1)
    {
        unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
        if(lane >= (unsigned)d) incl += up;
    }
    if(lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if(warp == 0)
    {
        unsigned w = warp_sums[lane], wi = w;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1)
        {
            unsigned up = __shfl_up_sync(0xffffffffu, wi, d);
            if(lane >= (unsigned)d) wi += up;
        }
        warp_sums[lane] = wi - w;
        if(lane == 31)
        {
            // publish, then look back
            const unsigned aggregate = wi;
            volatile unsigned long long *st = state;
            unsigned prefix = 0;
            if(chunk > 0)
            {
                st[chunk] = (1ull << 32) | aggregate;
                __threadfence();
                for(int j = (int)chunk - 1; j >= 0; )
                {
                    const unsigned long long sv = st[j];
                    const unsigned flag = (unsigned)(sv >> 32);
                    if(flag == 0u) continue;                     // not published yet
                    prefix += (unsigned)sv;
                    if(flag == 2u) break;
                    --j;
                }
            }
            __threadfence();
            st[chunk] = (2ull << 32) | (prefix + aggregate);
            s_prefix = prefix;
            if(chunk == gridDim.x - 1) { *total = prefix + aggregate; offset[n] = prefix + aggregate; }
        }
    }
    __syncthreads();
    unsigned run = s_prefix + warp_sums[warp] + (incl - sum);
#pragma unroll
    for(int i = 0; i < 8; ++i) { if(lo + i < hi) offset[lo + i] = run; run += v[i]; }
}

// state: ceil(n/8192) zeroed 64-bit words, ticket: one zeroed word (both part of the frame's control words)
void launch_tile_scan(const unsigned *tile_count, unsigned *tile_offset, unsigned ntiles,
                      unsigned *pair_total, unsigned long long *state, unsigned *ticket, cudaStream_t s)
{
    if(ntiles <= kChunk)
    {
        tile_scan_kernel<<<1, kScanThreads, 0, s>>>(tile_count, tile_offset, ntiles, pair_total);
        return;
    }
    const unsigned chunks = (ntiles + kChunk - 1)/kChunk;
    lookback_scan_kernel<<<chunks, kScanThreads, 0, s>>>(tile_count, tile_offset, ntiles, pair_total, state, ticket);
}

void launch_scatter(const ScatterParams &p, cudaStream_t s)
{
    if(p.seg_capacity == 0) return;
    // the fill counts live on the device: size the grid for a region's capacity, grid-stride inside
    unsigned blocks = (p.seg_capacity/kSubAllocators + 255)/256;
    if(blocks > 64) blocks = 64;
    if(blocks < 1) blocks = 1;
    scatter_kernel<<<dim3(blocks, kSubAllocators), kScatterThreads, 0, s>>>(p);
}

void launch_finalize(const FinalizeParams &p, cudaStream_t s)
{
    finalize_kernel<<<1, 32, 0, s>>>(p);
}

} // namespace b200r
