// MergeSort's tie order as a sort key (projekt.cpp:2-72), shared by the device code (edge_table_kernels.cu)
// and the CPU test that pins it against the verbatim reference's permutations (tests/test_merge_order.py).
//
// The reference's MergeSort is NOT stable: on equal YMin the merge takes the RIGHT half first (:51-58) while
// the two-element base case keeps the left element first (:13).  The resulting order is nevertheless a pure
// function of (YMin, position before the sort): follow the recursion from the root (Half0 = Count/2, :21-22)
// down to an element and note, per level, whether it sits on the side that loses ties.  Those bits, most
// significant first, are a key under which ANY correct sort reproduces the reference's permutation:
//   key(i) = (YMin(i) as ordered 32 bits) << 32 | b200r_merge_tie_path(i, n)      -- unique per element
#pragma once

#if defined(__CUDACC__)
#define B200R_HOST_DEVICE __host__ __device__
#else
#define B200R_HOST_DEVICE
#endif

// position of element i of n in MergeSort's tie order, as left-aligned path bits (0 = wins ties)
static inline B200R_HOST_DEVICE unsigned b200r_merge_tie_path(unsigned i, unsigned n)
{
    unsigned key = 0, lo = 0, cnt = n;
    int depth = 0;
    while(cnt > 2)
    {
        const unsigned half0 = cnt/2;                       // projekt.cpp:21
        unsigned bit;
        if(i - lo < half0) { bit = 1; cnt = half0; }        // left half: loses ties (:51-58)
        else { bit = 0; lo += half0; cnt -= half0; }
        key = (key << 1) | bit; ++depth;
    }
    if(cnt == 2) { key = (key << 1) | (i - lo); ++depth; }  // base case: left first (:13)
    return depth ? key << (32 - depth) : 0u;
}
