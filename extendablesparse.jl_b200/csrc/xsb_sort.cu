// xsb_sort.cu -- stable LSD radix sort of 16-byte records on a bit range of the
// packed (col,row) key, "onesweep" style: one upfront digit histogram for all
// passes, then per pass a single kernel that ranks a tile with warp ballots,
// resolves its global offsets by decoupled look-back over per-tile digit counts
// and scatters through shared memory so that each digit's run leaves the SM as
// coalesced 16-byte stores.
//
// Reference counterpart (CPU): the per-column QuickSort of
// src/matrix/sparsematrixlnk.jl:339 and the counting sort inside stdlib
// sparse! reached from src/matrix/sparsematrixdilnkc.jl:428-432.
#include "xsb_internal.h"

namespace xsb {

// ------------------------------------------------------------------------
// digit histogram for all passes
// ------------------------------------------------------------------------
constexpr int HIST_THREADS = 512;

__global__ void __launch_bounds__(HIST_THREADS)
histogram_kernel(const Rec *__restrict__ in, u64 n, SortPlan plan, u64 *__restrict__ ghist)
{
    __shared__ u32 s_hist[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += HIST_THREADS)
        (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * HIST_THREADS;
    // 32-bit shared counters: one block sees n/gridDim.x < 2^32 keys for every n that fits in HBM
    for (u64 k = (u64)blockIdx.x * HIST_THREADS + threadIdx.x; k < n; k += stride)
    {
        const u64 key = in[k].key;
#pragma unroll
        for (int p = 0; p < kMaxPasses; ++p)
            if (p < plan.npasses)
                atomicAdd(&s_hist[p][(key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npasses * kRadix; i += HIST_THREADS)
    {
        u32 c = (&s_hist[0][0])[i];
        if (c)
            atomicAdd(&ghist[i], (u64)c);
    }
}

// exclusive scan of each pass's histogram -> first output index of every digit
__global__ void __launch_bounds__(kRadix) histogram_scan_kernel(u64 *__restrict__ ghist, int npasses)
{
    __shared__ u64 s_warp[kRadix / 32];
    const int p = blockIdx.x;
    if (p >= npasses)
        return;
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    const u64 c = ghist[p * kRadix + b];
    u64 incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        u64 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    u64 pre = 0;
    for (int w = 0; w < warp; ++w)
        pre += s_warp[w];
    ghist[p * kRadix + b] = pre + incl - c;
}

// ------------------------------------------------------------------------
// one onesweep pass
// ------------------------------------------------------------------------
constexpr int OS_THREADS = 256;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_IPT = 16;
constexpr int OS_TILE = OS_THREADS * OS_IPT; // 4096 records = 64 KB of shared memory
static_assert(OS_THREADS == kRadix, "one thread per digit in the look-back");

// look-back words carry 30-bit counts: longer inputs are sorted in portions
constexpr u64 kSortPortion = ((1ull << 30) - 1ull) / OS_TILE * OS_TILE;
constexpr u32 ST_LOCAL = 1u << 30;
constexpr u32 ST_INCL = 2u << 30;
constexpr u32 ST_VALUE = (1u << 30) - 1u;

__global__ void __launch_bounds__(OS_THREADS, 2)
onesweep_kernel(const Rec *__restrict__ in, Rec *__restrict__ out, u32 n, u32 ntiles, int shift,
                int bits, const u64 *__restrict__ gbase, u64 *__restrict__ gbase_next,
                u32 *__restrict__ status, u32 *__restrict__ tile_counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec *s_rec = reinterpret_cast<Rec *>(smem_raw);
    __shared__ u32 s_whist[OS_WARPS][kRadix];
    __shared__ u32 s_binstart[kRadix];
    __shared__ i64 s_gofs[kRadix];
    __shared__ u32 s_warpsum[OS_WARPS];
    __shared__ u32 s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_tile = atomicAdd(tile_counter, 1u); // tiles are claimed in launch order: predecessors always run
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w)
        s_whist[w][tid] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 base = tile * OS_TILE;
    const u32 valid = min((u32)OS_TILE, n - base);
    const u32 mask = (1u << bits) - 1u;
    const u32 nbins = 1u << bits;

    // ---- load: warp-striped, one 16-byte vector per lane per step
    u64 key[OS_IPT];
    double val[OS_IPT];
    const u32 wbase = warp * (32 * OS_IPT);
#pragma unroll
    for (int k = 0; k < OS_IPT; ++k)
    {
        const u32 idx = wbase + k * 32 + lane;
        if (idx < valid)
        {
            Rec r = ld_rec_stream(in + base + idx);
            key[k] = r.key;
            val[k] = r.val;
        }
        else
        {
            key[k] = ~0ull; // padding sorts behind every real record of the tile
            val[k] = 0.0;
        }
    }

    // ---- rank inside the warp with ballots (stable: lower lane, lower step first)
    const u32 lt = lanemask_lt();
    u32 rank[OS_IPT];
#pragma unroll
    for (int k = 0; k < OS_IPT; ++k)
    {
        const u32 d = (u32)(key[k] >> shift) & mask;
        u32 peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 8; ++b)
        {
            if (b < bits)
            {
                const bool bit = (d >> b) & 1u;
                const u32 bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
        }
        const int leader = __ffs(peers) - 1;
        u32 pre = 0;
        if (lane == leader)
        {
            pre = s_whist[warp][d];
            s_whist[warp][d] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[k] = pre + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive prefix over the warps, tile count
    u32 total = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w)
    {
        const u32 c = s_whist[w][tid];
        s_whist[w][tid] = total;
        total += c;
    }
    // exclusive scan of the tile counts over the digits
    u32 incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s_warpsum[warp] = incl;
    __syncthreads();
    u32 wpre = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w)
        if (w < warp)
            wpre += s_warpsum[w];
    const u32 excl = wpre + incl - total;
    s_binstart[tid] = excl;

    // ---- decoupled look-back: one thread per digit
    if ((u32)tid < nbins)
    {
        u32 *mine = status + (size_t)tile * nbins + tid;
        u32 prefix = 0;
        if (tile == 0)
        {
            st_relaxed_u32(mine, ST_INCL | total);
        }
        else
        {
            st_relaxed_u32(mine, ST_LOCAL | total);
            for (i64 t = (i64)tile - 1;; --t)
            {
                const u32 *p = status + (size_t)t * nbins + tid;
                u32 v;
                do
                {
                    v = ld_relaxed_u32(p);
                } while ((v >> 30) == 0u);
                prefix += v & ST_VALUE;
                if ((v >> 30) == 2u)
                    break;
            }
            st_relaxed_u32(mine, ST_INCL | (prefix + total));
        }
        const u64 g = gbase[tid];
        s_gofs[tid] = (i64)(g + prefix) - (i64)excl;
        if (gbase_next != nullptr && tile == ntiles - 1)
            gbase_next[tid] = g + prefix + total;
    }
    __syncthreads();

    // ---- reorder through shared memory
#pragma unroll
    for (int k = 0; k < OS_IPT; ++k)
    {
        const u32 d = (u32)(key[k] >> shift) & mask;
        const u32 pos = s_binstart[d] + s_whist[warp][d] + rank[k];
        Rec r;
        r.key = key[k];
        r.val = val[k];
        s_rec[pos] = r;
    }
    __syncthreads();

    // ---- scatter: consecutive threads hold consecutive ranks, so each digit's run is coalesced
#pragma unroll
    for (int i = 0; i < OS_IPT; ++i)
    {
        const u32 idx = i * OS_THREADS + tid;
        if (idx < valid)
        {
            const Rec r = s_rec[idx];
            const u32 d = (u32)(r.key >> shift) & mask;
            st_rec(out + (s_gofs[d] + (i64)idx), r);
        }
    }
}

// ------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------
SortPlan make_sort_plan(int begin_bit, int nbits)
{
    SortPlan p{};
    if (nbits <= 0)
    {
        p.npasses = 0;
        return p;
    }
    p.npasses = (nbits + 7) / 8;
    if (p.npasses > kMaxPasses)
        throw std::runtime_error("key too wide for the radix sort");
    int shift = begin_bit;
    for (int i = 0; i < p.npasses; ++i)
    {
        int b = nbits / p.npasses + (i < nbits % p.npasses ? 1 : 0);
        p.bits[i] = b;
        p.shift[i] = shift;
        shift += b;
    }
    return p;
}

size_t sort_workspace_bytes(u64 n)
{
    const u64 portion = std::min<u64>(n, kSortPortion);
    const u64 ntiles = (portion + OS_TILE - 1) / OS_TILE;
    // ghist [kMaxPasses][256] u64, gbase_next [256] u64 x2, counter u32 (padded), status [ntiles][256] u32
    return sizeof(u64) * kMaxPasses * kRadix + 2 * sizeof(u64) * kRadix + 256 +
           sizeof(u32) * ntiles * kRadix;
}

// Sorts n records by key bits [begin_bit, begin_bit+nbits).  `a` holds the input;
// `b` is scratch of the same size.  Returns the buffer that holds the result.
Rec *radix_sort_records(cudaStream_t stream, Rec *a, Rec *b, u64 n, const SortPlan &plan, void *workspace,
                        LaunchCounter &lc, StageTimer *timer)
{
    if (n <= 1 || plan.npasses == 0)
        return a;
    static bool attr_set = false;
    if (!attr_set)
    {
        XSB_CUDA(cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(OS_TILE * sizeof(Rec))));
        attr_set = true;
    }
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    u64 *ghist = reinterpret_cast<u64 *>(ws);
    u64 *gnext0 = ghist + kMaxPasses * kRadix;
    u64 *gnext1 = gnext0 + kRadix;
    u32 *counter = reinterpret_cast<u32 *>(gnext1 + kRadix);
    u32 *status = counter + 64;

    if (timer)
        timer->begin(stream);
    XSB_CUDA(cudaMemsetAsync(ghist, 0, sizeof(u64) * kMaxPasses * kRadix, stream));
    {
        const u64 want = (n + HIST_THREADS * 8 - 1) / (HIST_THREADS * 8);
        const int blocks = (int)std::min<u64>(std::max<u64>(want, 1), (u64)kNumSM * 4);
        histogram_kernel<<<blocks, HIST_THREADS, 0, stream>>>(a, n, plan, ghist);
        lc.add();
        histogram_scan_kernel<<<plan.npasses, kRadix, 0, stream>>>(ghist, plan.npasses);
        lc.add();
        XSB_CUDA(cudaGetLastError());
    }
    if (timer)
        timer->end(stream, &StageTimes::histogram);

    if (timer)
        timer->begin(stream);
    Rec *src = a, *dst = b;
    for (int p = 0; p < plan.npasses; ++p)
    {
        const u32 nbins = 1u << plan.bits[p];
        const u64 *gbase = ghist + p * kRadix;
        u64 *gn[2] = {gnext0, gnext1};
        int flip = 0;
        for (u64 off = 0; off < n; off += kSortPortion)
        {
            const u32 cnt = (u32)std::min<u64>(kSortPortion, n - off);
            const u32 ntiles = (cnt + OS_TILE - 1) / OS_TILE;
            const bool more = off + kSortPortion < n;
            XSB_CUDA(cudaMemsetAsync(counter, 0, 256 + sizeof(u32) * (size_t)ntiles * nbins, stream));
            onesweep_kernel<<<ntiles, OS_THREADS, OS_TILE * sizeof(Rec), stream>>>(
                src + off, dst, cnt, ntiles, plan.shift[p], plan.bits[p], gbase, more ? gn[flip] : nullptr,
                status, counter);
            lc.add();
            if (more)
            {
                gbase = gn[flip];
                flip ^= 1;
            }
        }
        XSB_CUDA(cudaGetLastError());
        std::swap(src, dst);
    }
    if (timer)
        timer->end(stream, &StageTimes::sort);
    return src;
}

} // namespace xsb
