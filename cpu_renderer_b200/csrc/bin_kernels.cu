// Sort-middle binner (sm_100a): per-tile lists of trapezoid segments from the set-up kernel's
// exact per-segment tile columns.  The reference has no binning (its nearest analogue is MergeSort by YMin plus
// one work item per scan line, projekt.cpp:2-72, 3509-3609); this stage exists because the
// raster kernel keeps a screen tile on chip.
//
//   count   (inside setup_kernel)   tile_count[t] += 1 per tile a segment touches
//   scan    tile_scan_kernel        exclusive prefix sum over tiles -> tile_offset, pair_total
//   scatter scatter_kernel          list[tile_offset[t] + slot] = segment
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
}

// One thread per segment; lanes whose segments fall into the same tile contend on that tile's
// cursor, which the L2 atomic unit serialises.
__global__ void __launch_bounds__(256)
scatter_kernel(const uint2 *__restrict__ seg_tiles, const unsigned *__restrict__ seg_total,
               unsigned seg_capacity, int tiles_x,
               const unsigned *__restrict__ tile_offset, unsigned *__restrict__ tile_fill,
               unsigned *__restrict__ pair_list, const unsigned *__restrict__ pair_total,
               unsigned pair_capacity)
{
    const unsigned nseg = *seg_total;
    if(nseg > seg_capacity || *pair_total > pair_capacity) return;   // host grows and re-issues
    for(unsigned seg = blockIdx.x*blockDim.x + threadIdx.x; seg < nseg; seg += gridDim.x*blockDim.x)
    {
        const uint2 r = seg_tiles[seg];
        const int tx0 = r.y & 0xffff, tx1 = r.y >> 16;
        for(int tx = tx0; tx <= tx1; ++tx)
        {
            const unsigned tile = r.x*(unsigned)tiles_x + (unsigned)tx;
            const unsigned slot = atomicAdd(&tile_fill[tile], 1u);
            pair_list[tile_offset[tile] + slot] = seg;
        }
    }
}

void launch_tile_scan(const unsigned *tile_count, unsigned *tile_offset, unsigned ntiles,
                      unsigned *pair_total, cudaStream_t s)
{
    tile_scan_kernel<<<1, kScanThreads, 0, s>>>(tile_count, tile_offset, ntiles, pair_total);
}

void launch_scatter(const uint2 *seg_tiles, const unsigned *seg_total, unsigned seg_capacity,
                    unsigned max_segments, int tiles_x, const unsigned *tile_offset,
                    unsigned *tile_fill, unsigned *pair_list, const unsigned *pair_total,
                    unsigned pair_capacity, cudaStream_t s)
{
    if(max_segments == 0) return;
    // the segment count lives on the device: size the grid for the capacity, grid-stride inside
    unsigned blocks = (max_segments + 255)/256;
    if(blocks > 148*16) blocks = 148*16;
    scatter_kernel<<<blocks, 256, 0, s>>>(seg_tiles, seg_total, seg_capacity, tiles_x, tile_offset, tile_fill,
                                          pair_list, pair_total, pair_capacity);
}

} // namespace b200r
