// xsb_fold.cuh -- device helpers shared by the per-column kernels: the insertion-order fold
// state machines and an in-register warp bitonic sort.
#pragma once
#include "xsb_common.cuh"

namespace xsb {

// Insertion-order fold of one run of equal (col,row): the semantics of the three insert
// flavours (src/matrix/sparsematrixlnk.jl:178-253) behind the CSC-hit router
// (src/matrix/extendable.jl:159-218), per partition buffer, then partitions summed in tid
// order with the first one copied (sparse! combine, src/matrix/sparsematrixdilnkc.jl:416-432).
// Same state machine as RunFold in xsb_flush.cu (kept per translation unit on purpose).
struct ColFold
{
    bool seeded = false, has_old = false, exists = false, pexists = false;
    double old = 0.0, acc = 0.0, pacc = 0.0;
    u32 ptid = 0xffffffffu;
    __device__ __forceinline__ void commit()
    {
        if (pexists)
        {
            if (exists)
                acc = acc + pacc;
            else
            {
                acc = pacc;
                exists = true;
            }
            pexists = false;
        }
    }
    __device__ __forceinline__ void apply(u32 fl, u32 tid, double v, int combine)
    {
        if (fl == FL_OLD)
        {
            if (combine == 0)
            {
                seeded = true;
                exists = true;
                acc = v;
            }
            else
            {
                has_old = true;
                old = v;
            }
            return;
        }
        if (seeded)
        {
            acc = (fl == FL_ASSIGN) ? v : acc + v;
            return;
        }
        if (tid != ptid)
        {
            commit();
            ptid = tid;
        }
        if (fl == FL_RAW)
        {
            pacc = pexists ? pacc + v : 0.0 + v;
            pexists = true;
        }
        else if (fl == FL_UPDATE)
        {
            if (pexists)
                pacc = pacc + v;
            else if (v != 0.0)
            {
                pexists = true;
                pacc = 0.0 + v;
            }
        }
        else
        {
            if (pexists)
                pacc = v;
            else if (v != 0.0)
            {
                pexists = true;
                pacc = v;
            }
        }
    }
    __device__ __forceinline__ void finish()
    {
        commit();
        if (has_old)
        {
            acc = exists ? old + acc : old;
            exists = true;
        }
    }
};

// The fold when a flush holds one partition, no assign flavour and old entries seed the sum: a
// running sum from +0.0 (sparsematrixlnk.jl:225), an old CSC value (always the first record of its
// run) replaces the seed (extendable.jl:165-166); the entry exists once any record would have
// created it (raw always, update only with v != 0: sparsematrixlnk.jl:212,223,239).
struct SimpleFold
{
    double acc = 0.0;
    bool exists = false;
    __device__ __forceinline__ void apply(u32 fl, u32, double v, int)
    {
        acc = (fl == FL_OLD) ? v : acc + v;
        exists |= (fl != FL_UPDATE) | (v != 0.0);
    }
    __device__ __forceinline__ void finish() {}
};

// bitonic sort of 32*E keys held E per lane (element index = lane*E + e), ascending
template <int E> __device__ __forceinline__ void warp_bitonic(u32 (&k)[E], int lane)
{
    constexpr int N = 32 * E;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1)
    {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
        {
            if (stride >= E)
            { // partner in another lane, same register slot
                const int lstride = stride / E;
#pragma unroll
                for (int e = 0; e < E; ++e)
                {
                    const int g = lane * E + e;
                    const u32 other = __shfl_xor_sync(0xffffffffu, k[e], lstride);
                    const bool up = (g & size) == 0;      // ascending block
                    const bool lower = (g & stride) == 0; // this element is the lower index of the pair
                    const u32 mn = min(k[e], other), mx = max(k[e], other);
                    k[e] = (up == lower) ? mn : mx;
                }
            }
            else
            { // both elements in this lane
#pragma unroll
                for (int e = 0; e < E; ++e)
                {
                    if ((e & stride) == 0)
                    {
                        const int g = lane * E + e;
                        const bool up = (g & size) == 0;
                        const u32 a = k[e], b = k[e + stride];
                        const u32 mn = min(a, b), mx = max(a, b);
                        k[e] = up ? mn : mx;
                        k[e + stride] = up ? mx : mn;
                    }
                }
            }
        }
    }
}

} // namespace xsb
