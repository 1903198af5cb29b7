// Tail of FillEdgeTable on the device (sm_100a): append every triangle's edges in emission order
// (projekt.cpp:3947, 4113-4115) and put them into the order the reference's MergeSort leaves them in
// (projekt.cpp:2-72, called at :4117), then assemble edge_info records (projekt.h:17-37).
//
// MergeSort is NOT stable, but its order is a pure function of (YMin, position before the sort)
// (merge_order.h): under the key
//
//   key(i) = (YMin(i) as ordered 32 bits) << 32 | b200r_merge_tie_path(i, n)
//
// ANY correct sort reproduces the reference's permutation -- keys are unique, so stability and the shape
// of the device sort do not matter.
//
// Kernels: edge counts per triangle -> chained scan (bin_kernels.cu) -> keys and payloads ->
// 2048-element bitonic tiles in shared memory -> log2(n / 2048) rank-merge passes (every element finds
// its rank in the sibling run by binary search; unique keys make lower_bound exact for both sides) ->
// assemble.  The first version did this on the host (D2H of the per-triangle records, std::sort, H2D
// for the whole-object mode).
#include "raster_device.cuh"
#include "merge_order.h"

namespace b200r {

namespace {

constexpr int kSortTile = 2048;         // elements per bitonic tile
constexpr int kSortThreads = 1024;

// byte-for-byte edge_info (projekt.h:17-37, include/b200_raster.h)
struct EdgeOut
{
    int YMax; float XMin, ZMin, OneOverZMin, Gradient, ZGradient, OneOverZGradient;
    int YMin; float UMin, VMin, UGradient, VGradient;
    int Left; float MinColor[4], ColorGradient[4], MinNormal[3], NormalGradient[3];
    long long Next;
};
static_assert(sizeof(EdgeOut) == 120, "edge_info is 120 bytes");

__global__ void __launch_bounds__(256)
edge_count_kernel(const uint32_t *__restrict__ recs, unsigned ntri, unsigned *__restrict__ counts)
{
    const unsigned tri = blockIdx.x*blockDim.x + threadIdx.x;
    if(tri < ntri) counts[tri] = recs[(size_t)tri*kRecWords + R_NEDGES];
}

// thread per triangle: its edges' keys and payloads (triangle << 2 | slot of the record) at their emission index
__global__ void __launch_bounds__(256)
edge_key_kernel(const uint32_t *__restrict__ recs, unsigned ntri, const unsigned *__restrict__ offsets,
                const unsigned *__restrict__ total, unsigned long long *__restrict__ keys, unsigned *__restrict__ vals)
{
    const unsigned tri = blockIdx.x*blockDim.x + threadIdx.x;
    if(tri >= ntri) return;
    const uint32_t *rec = recs + (size_t)tri*kRecWords;
    const unsigned ne = rec[R_NEDGES], n = *total, at = offsets[tri];
    const uint32_t emit = rec[R_EDGE0 + 3*kEdgeWords];      // slot of the k-th emitted edge, 2 bits each
    for(unsigned k = 0; k < ne && k < 3u; ++k)
    {
        const unsigned slot = (emit >> (2*k)) & 3u;
        const uint32_t ymin = rec[R_EDGE0 + slot*kEdgeWords + E_YMIN];
        const unsigned i = at + k;
        B200R_ASSERT(i < n);
        keys[i] = ((unsigned long long)(ymin ^ 0x80000000u) << 32) | b200r_merge_tie_path(i, n);
        vals[i] = (tri << 2) | slot;
    }
}

// one CTA sorts one tile of kSortTile elements in shared memory (bitonic network).  The tail tile is padded
// with the largest key and the largest payload; the payload breaks a tie, so padding ends up behind every
// real element and is never written back.
__global__ void __launch_bounds__(kSortThreads)
edge_tile_sort_kernel(unsigned long long *__restrict__ keys, unsigned *__restrict__ vals, const unsigned *__restrict__ total)
{
    __shared__ unsigned long long s_key[kSortTile];
    __shared__ unsigned s_val[kSortTile];
    const unsigned n = *total, base = blockIdx.x*kSortTile;
    if(base >= n) return;
    for(unsigned j = threadIdx.x; j < (unsigned)kSortTile; j += kSortThreads)
    {
        const bool in = base + j < n;
        s_key[j] = in ? keys[base + j] : ~0ull;
        s_val[j] = in ? vals[base + j] : 0xffffffffu;
    }
    __syncthreads();
    for(unsigned k = 2; k <= (unsigned)kSortTile; k <<= 1)
        for(unsigned j = k >> 1; j > 0; j >>= 1)
        {
            // every thread owns one compare-exchange of this step
            const unsigned t = threadIdx.x;
            const unsigned a = ((t & ~(j - 1u)) << 1) | (t & (j - 1u)), b = a | j;
            const bool up = (a & k) == 0;
            const unsigned long long ka = s_key[a], kb = s_key[b];
            // padding (val 0xffffffff) must end up behind every real element even on an equal key
            const bool gt = ka > kb || (ka == kb && s_val[a] > s_val[b]);
            if(gt == up)
            {
                s_key[a] = kb; s_key[b] = ka;
                const unsigned va = s_val[a]; s_val[a] = s_val[b]; s_val[b] = va;
            }
            __syncthreads();
        }
    for(unsigned j = threadIdx.x; j < (unsigned)kSortTile; j += kSortThreads)
        if(base + j < n) { keys[base + j] = s_key[j]; vals[base + j] = s_val[j]; }
}

// one merge level: runs of `run` sorted elements are merged pairwise; every element computes where it
// lands -- its position in its own run plus its rank in the sibling run
__global__ void __launch_bounds__(256)
edge_merge_kernel(const unsigned long long *__restrict__ kin, const unsigned *__restrict__ vin,
                  unsigned long long *__restrict__ kout, unsigned *__restrict__ vout,
                  const unsigned *__restrict__ total, unsigned run)
{
    const unsigned n = *total;
    const unsigned i = blockIdx.x*blockDim.x + threadIdx.x;
    if(i >= n) return;
    const unsigned r = i/run, mine = r*run, pair0 = (r & ~1u)*run;
    const unsigned sib = (r ^ 1u)*run;
    const unsigned long long key = kin[i];
    unsigned lo = 0, hi = (sib < n) ? min(run, n - sib) : 0u;       // sibling run [sib, sib + hi)
    while(lo < hi)
    {
        const unsigned mid = (lo + hi) >> 1;
        if(kin[sib + mid] < key) lo = mid + 1; else hi = mid;
    }
    const unsigned dst = pair0 + (i - mine) + lo;
    B200R_ASSERT(dst < n);
    kout[dst] = key; vout[dst] = vin[i];
}

// SSE's invalid-operation result is the default NaN 0xFFC00000 (and it propagates), the GPU's is
// 0x7FFFFFFF.  NaN sign / payload is not part of the arithmetic contract; such values only occur in edges
// that are never drawn (YMax == YMin: gradients 0/0) and are exported in the x86 encoding so that the
// table is byte-identical to the reference's.
__device__ __forceinline__ float x86_nan(uint32_t u)
{
    if((u & 0x7fffffffu) > 0x7f800000u) u = 0xffc00000u;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(128)
edge_assemble_kernel(const uint32_t *__restrict__ recs, const uint32_t *__restrict__ uvrecs,
                     const float *__restrict__ normals, const unsigned *__restrict__ vals,
                     const unsigned *__restrict__ total, EdgeOut *__restrict__ out)
{
    const unsigned i = blockIdx.x*blockDim.x + threadIdx.x;
    if(i >= *total) return;
    const unsigned v = vals[i], tri = v >> 2, slot = v & 3u;
    const size_t at = (size_t)tri*kRecWords + R_EDGE0 + slot*kEdgeWords;
    const uint32_t *E = recs + at;
    EdgeOut o;
    o.YMin = (int)E[E_YMIN]; o.YMax = (int)E[E_YMAX];
    o.XMin = x86_nan(E[E_X]); o.Gradient = x86_nan(E[E_DX]); o.ZMin = x86_nan(E[E_Z]); o.ZGradient = x86_nan(E[E_DZ]);
#pragma unroll
    for(int k = 0; k < 4; ++k) { o.MinColor[k] = x86_nan(E[E_C + k]); o.ColorGradient[k] = x86_nan(E[E_DC + k]); }
    o.Left = (int)(E[E_LEFT] & 1u);
    o.Next = 0;
    o.UMin = o.VMin = o.OneOverZMin = o.UGradient = o.VGradient = o.OneOverZGradient = 0.0f;
    o.MinNormal[0] = o.MinNormal[1] = o.MinNormal[2] = 0.0f;
    o.NormalGradient[0] = o.NormalGradient[1] = o.NormalGradient[2] = 0.0f;
    if(uvrecs)
    {
        // same triangle, same slot of the second set-up pass: u/z, v/z, 1/z travel in the colour words
        const uint32_t *U = uvrecs + at;
        o.UMin = x86_nan(U[E_C + 0]); o.VMin = x86_nan(U[E_C + 1]); o.OneOverZMin = x86_nan(U[E_C + 2]);
        o.UGradient = x86_nan(U[E_DC + 0]); o.VGradient = x86_nan(U[E_DC + 1]); o.OneOverZGradient = x86_nan(U[E_DC + 2]);
    }
    if(normals)
    {
        // projekt.cpp:4017-4018, 4104-4109: MinNormal = the upper vertex's normal (not advanced by the top
        // clip), NormalGradient = (MaxNormal - MinNormal)/YDifference: one subtraction and one division each
        const float *N = normals + (size_t)tri*9;
        const unsigned mn = (E[E_LEFT] >> 8) & 3u, mx = (E[E_LEFT] >> 16) & 3u;
        const float ydiff = fsub(__int2float_rn(o.YMax), __int2float_rn(o.YMin));
#pragma unroll
        for(int k = 0; k < 3; ++k)
        {
            const float a = N[3*mn + k], b = N[3*mx + k];
            o.MinNormal[k] = a;
            o.NormalGradient[k] = x86_nan(__float_as_uint(fdiv(fsub(b, a), ydiff)));
        }
    }
    out[i] = o;
}

} // namespace

void launch_edge_counts(const uint32_t *recs, unsigned ntri, unsigned *counts, cudaStream_t s)
{
    if(ntri) edge_count_kernel<<<(ntri + 255)/256, 256, 0, s>>>(recs, ntri, counts);
}

void launch_edge_keys(const uint32_t *recs, unsigned ntri, const unsigned *offsets, const unsigned *total,
                      unsigned long long *keys, unsigned *vals, cudaStream_t s)
{
    if(ntri) edge_key_kernel<<<(ntri + 255)/256, 256, 0, s>>>(recs, ntri, offsets, total, keys, vals);
}

// sorts n (key, payload) pairs; n is also on the device (*total).  Returns which of the two buffer pairs
// holds the result (0 or 1).
int launch_edge_sort(unsigned long long *keys[2], unsigned *vals[2], unsigned n, const unsigned *total, cudaStream_t s)
{
    if(n == 0) return 0;
    edge_tile_sort_kernel<<<(n + kSortTile - 1)/kSortTile, kSortThreads, 0, s>>>(keys[0], vals[0], total);
    int cur = 0;
    for(unsigned long long run = kSortTile; run < n; run <<= 1)
    {
        edge_merge_kernel<<<(n + 255)/256, 256, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], total, (unsigned)run);
        cur ^= 1;
    }
    return cur;
}

void launch_edge_assemble(const uint32_t *recs, const uint32_t *uvrecs, const float *normals, const unsigned *vals,
                          unsigned n, const unsigned *total, void *out, cudaStream_t s)
{
    if(n) edge_assemble_kernel<<<(n + 127)/128, 128, 0, s>>>(recs, uvrecs, normals, vals, total, reinterpret_cast<EdgeOut *>(out));
}

} // namespace b200r
