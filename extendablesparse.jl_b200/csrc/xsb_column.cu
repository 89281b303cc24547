// xsb_column.cu -- flush! after a COLUMN-ONLY radix sort (half the passes of a (col,row) sort):
//
//   column_reduce_kernel : a tile owns the columns that start inside it.  It stages the tile
//     (+ look-ahead) in shared memory, finds the column boundaries, and one warp per column
//       * orders the column's records by (row, position in column) with a register bitonic
//         sort on a packed 32-bit key -- position is the insertion order, so the order inside a
//         run of equal rows is the stream order (what the stable global sort would have given),
//       * folds every run of equal rows sequentially (bit-exact left fold, same RunFold state
//         machine as the (col,row) path),
//       * leaves the column's entries (sorted rows, values) and their count in shared memory.
//     Tile output offsets come from a decoupled look-back over the tiles' entry counts; the
//     entries are written to rowval/nzval and the per-column counts to colcount.
//   colcount -> colptr   : exclusive sum scan.
//
// This is the reference's own per-column structure (gather column, sort by row, merge:
// src/matrix/sparsematrixlnk.jl:328-377) with the column gather done by the radix sort.
// Columns longer than the in-warp limit raise an overflow flag; the caller then finishes with
// the general (col,row) sort path.
#include "xsb_internal.h"
#include "xsb_fold.cuh"

namespace xsb {

constexpr int CK_THREADS = 256;
constexpr int CK_WARPS = CK_THREADS / 32;
constexpr int CK_T = 2048;                 // nominal tile: columns starting in [tile*CK_T, (tile+1)*CK_T) are owned
constexpr int CK_CAP = 3072;               // records staged per tile (tile + look-ahead for the last owned column)
constexpr int CK_IPT = CK_CAP / CK_THREADS; // 12
constexpr int CK_MAXE = 8;                 // up to 32*8 = 256 records per column in the warp sort

constexpr u64 CS_LOCAL = 1ull << 62;
constexpr u64 CS_INCL = 2ull << 62;
constexpr u64 CS_VALUE = (1ull << 62) - 1ull;

// exclusive rank of every element among the set flags, flags laid out as (round i, thread)
template <int IPT, int WARPS>
__device__ __forceinline__ u32 block_rank(const bool (&flag)[IPT], u32 (&excl)[IPT], u32 *s_cnt, u32 *s_total,
                                          int lane, int warp, u32 lt)
{
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
        const u32 bal = __ballot_sync(0xffffffffu, flag[i]);
        if (lane == 0)
            s_cnt[i * WARPS + warp] = __popc(bal);
        excl[i] = __popc(bal & lt);
    }
    __syncthreads();
    if (warp == 0)
    {
        constexpr int N = IPT * WARPS;
        constexpr int PER = (N + 31) / 32;
        u32 c[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k)
        {
            const int idx = lane * PER + k;
            c[k] = idx < N ? s_cnt[idx] : 0u;
            sum += c[k];
        }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        u32 run = incl - sum;
#pragma unroll
        for (int k = 0; k < PER; ++k)
        {
            const int idx = lane * PER + k;
            if (idx < N)
                s_cnt[idx] = run;
            run += c[k];
        }
        if (lane == 31)
            *s_total = incl;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        excl[i] += s_cnt[i * WARPS + warp];
    const u32 total = *s_total;
    __syncthreads(); // s_cnt / s_total may be reused right away
    return total;
}

// Sorts one column (len <= 32*E records at s_col[0..len)) by (row, position) and permutes the
// records in place into that order: keys are sorted in registers, every lane then gathers the
// records of its E ranks, and only after the whole warp has read are they written back.
template <int E>
__device__ __forceinline__ void sort_column(const u32 *s_keycol, Rec *s_col, u32 len, u32 posmask, int lane)
{
    u32 k[E];
#pragma unroll
    for (int e = 0; e < E; ++e)
    {
        const u32 i = e * 32 + lane; // any initial arrangement will do: the position is part of the key
        k[e] = i < len ? s_keycol[i] : 0xffffffffu;
    }
    warp_bitonic<E>(k, lane);
    Rec t[E];
#pragma unroll
    for (int e = 0; e < E; ++e)
    {
        const u32 r = lane * E + e;
        if (r < len)
            t[e] = s_col[k[e] & posmask];
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; ++e)
    {
        const u32 r = lane * E + e;
        if (r < len)
            s_col[r] = t[e];
    }
}

#include "xsb_column_kernel.cuh"


// ------------------------------------------------------------------------
// colptr = base + exclusive sum of colcount
// ------------------------------------------------------------------------
constexpr int CS_THREADS = 256;
constexpr int CS_IPT = 8;
constexpr int CS_TILE = CS_THREADS * CS_IPT;

__global__ void __launch_bounds__(CS_THREADS)
colcount_tilesum_kernel(const u32 *__restrict__ colcount, i64 n, u64 *__restrict__ tsum)
{
    __shared__ u64 s_w[CS_THREADS / 32];
    const i64 b0 = (i64)blockIdx.x * CS_TILE;
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < CS_IPT; ++i)
    {
        const i64 j = b0 + i * CS_THREADS + threadIdx.x;
        if (j < n)
            s += colcount[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
        s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < CS_THREADS / 32; ++w)
            s += s_w[w];
        tsum[blockIdx.x] = s;
    }
}

// single block: in-place exclusive scan of the tile sums
__global__ void __launch_bounds__(1024) tilesum_scan_kernel(u64 *__restrict__ tsum, i64 nt)
{
    __shared__ u64 s_w[32];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0)
        s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (i64 b0 = 0; b0 < nt; b0 += 1024)
    {
        const i64 j = b0 + threadIdx.x;
        const u64 x = j < nt ? tsum[j] : 0;
        u64 v = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const u64 t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o)
                v += t;
        }
        if (lane == 31)
            s_w[warp] = v;
        __syncthreads();
        u64 pre = s_carry;
        for (int w = 0; w < warp; ++w)
            pre += s_w[w];
        if (j < nt)
            tsum[j] = pre + v - x;
        __syncthreads();
        if (threadIdx.x == 1023)
            s_carry = pre + v;
        __syncthreads();
    }
}

template <typename Ti>
__global__ void __launch_bounds__(CS_THREADS)
colptr_from_counts_kernel(const u32 *__restrict__ colcount, const u64 *__restrict__ tsum, i64 n, Ti base,
                          Ti *__restrict__ colptr)
{
    __shared__ u64 s_w[CS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 b0 = (i64)blockIdx.x * CS_TILE + (i64)threadIdx.x * CS_IPT; // blocked: thread owns CS_IPT columns
    u64 c[CS_IPT], sum = 0;
#pragma unroll
    for (int i = 0; i < CS_IPT; ++i)
    {
        const i64 j = b0 + i;
        c[i] = j < n ? (u64)colcount[j] : 0ull;
        sum += c[i];
    }
    u64 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const u64 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s_w[warp] = incl;
    __syncthreads();
    u64 run = tsum[blockIdx.x];
    for (int w = 0; w < warp; ++w)
        run += s_w[w];
    run += incl - sum;
#pragma unroll
    for (int i = 0; i < CS_IPT; ++i)
    {
        const i64 j = b0 + i;
        if (j < n)
            colptr[j] = (Ti)run + base; // entries before column j
        run += c[i];
        if (j == n - 1)
            colptr[n] = (Ti)run + base;
    }
}

size_t column_workspace_bytes(u64 nrec, i64 ncols)
{
    const u64 ntiles = (nrec + CK_T - 1) / CK_T;
    const u64 ctiles = ((u64)ncols + CS_TILE - 1) / CS_TILE;
    return 256 + sizeof(u32) * (((size_t)ncols + 1) & ~(size_t)1) + sizeof(u64) * (ntiles + 1) +
           sizeof(u64) * (ctiles + 1) + 64;
}

size_t column_kernel_smem()
{
    return sizeof(Rec) * CK_CAP + sizeof(u32) * CK_CAP + sizeof(unsigned short) * (CK_CAP + 8 + CK_T) + 16;
}

bool column_path_supported(const KeyLayout &L) { return L.rowbits <= 27 && L.colbits <= 32; }

// records sorted by column (stable) -> CSC.  *h_overflow is set when a column was too long for the
// in-warp path (outputs are then undefined and the caller falls back to the (col,row) sort).
void column_reduce_emit_csc(cudaStream_t stream, const Rec *sorted, u64 nrec, KeyLayout L, int combine,
                            bool plain_adds, i64 ncols, int idx64, int base, void *rowval_out, double *nzval_out,
                            void *colptr_out, void *workspace, u64 *d_nnz, u32 *d_overflow, LaunchCounter &lc,
                            StageTimer *timer)
{
    const u64 ntiles = (nrec + CK_T - 1) / CK_T;
    const u64 ctiles = ((u64)ncols + CS_TILE - 1) / CS_TILE;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    u32 *counter = reinterpret_cast<u32 *>(ws);
    u32 *colcount = reinterpret_cast<u32 *>(ws + 256);
    u64 *status = reinterpret_cast<u64 *>(colcount + (((size_t)ncols + 1) & ~(size_t)1));
    u64 *tsum = status + ntiles + 1;
    const int posbits = std::min(8, 32 - L.rowbits);
    const size_t smem = column_kernel_smem();

    if (timer)
        timer->begin(stream);
    XSB_CUDA(cudaMemsetAsync(ws, 0, 256 + sizeof(u32) * (((size_t)ncols + 1) & ~(size_t)1) + sizeof(u64) * (ntiles + 1),
                             stream));
    XSB_CUDA(cudaMemsetAsync(d_nnz, 0, sizeof(u64), stream));
    XSB_CUDA(cudaMemsetAsync(d_overflow, 0, sizeof(u32), stream));
    const bool simple = plain_adds && L.tidbits == 0 && combine == 0;
    static FuncAttrOnce once[4];
    once[0].set(column_reduce_kernel<int64_t, true>, (int)smem);
    once[1].set(column_reduce_kernel<int64_t, false>, (int)smem);
    once[2].set(column_reduce_kernel<int32_t, true>, (int)smem);
    once[3].set(column_reduce_kernel<int32_t, false>, (int)smem);
#define XSB_LAUNCH_COL(TI, SIMPLE)                                                                                 \
    column_reduce_kernel<TI, SIMPLE><<<(unsigned)ntiles, CK_THREADS, smem, stream>>>(                              \
        sorted, nrec, L, posbits, combine, (TI)base, (TI *)rowval_out, nzval_out, colcount, status, counter, d_nnz, \
        d_overflow, (u32)ntiles)
    if (nrec > 0)
    {
        if (idx64)
        {
            if (simple)
                XSB_LAUNCH_COL(int64_t, true);
            else
                XSB_LAUNCH_COL(int64_t, false);
        }
        else
        {
            if (simple)
                XSB_LAUNCH_COL(int32_t, true);
            else
                XSB_LAUNCH_COL(int32_t, false);
        }
        lc.add();
        XSB_CUDA(cudaGetLastError());
    }
#undef XSB_LAUNCH_COL
    if (timer)
        timer->end(stream, &StageTimes::reduce);

    if (timer)
        timer->begin(stream);
    colcount_tilesum_kernel<<<(unsigned)ctiles, CS_THREADS, 0, stream>>>(colcount, ncols, tsum);
    tilesum_scan_kernel<<<1, 1024, 0, stream>>>(tsum, (i64)ctiles);
    if (idx64)
        colptr_from_counts_kernel<int64_t><<<(unsigned)ctiles, CS_THREADS, 0, stream>>>(colcount, tsum, ncols, (int64_t)base,
                                                                                        (int64_t *)colptr_out);
    else
        colptr_from_counts_kernel<int32_t><<<(unsigned)ctiles, CS_THREADS, 0, stream>>>(colcount, tsum, ncols, (int32_t)base,
                                                                                        (int32_t *)colptr_out);
    lc.add(3);
    XSB_CUDA(cudaGetLastError());
    if (timer)
        timer->end(stream, &StageTimes::colptr);
}

} // namespace xsb
