// xsb_route.cu -- multi-GPU routing of staged records by column ownership without moving the
// records a rank keeps.
//
// A staged key carries [owner | column relative to the owner's slab | row | tid | flavour], so a
// record is already in its owner's flush layout.  Routing therefore only has to COPY OUT the
// records owned by other ranks (in stream order, destination after destination); they stay
// behind in the staging buffer as records the flush skips by their owner bits.  For a
// partitioned assembly the foreign part is the interface between slabs (P1-FEM 128^2 x 127
// layers: 0.4 % of the stream), so routing costs one read of the keys instead of two passes over
// all records.
//
//   route_count_kernel   : records per destination rank in every tile of RT_TILE records
//   route_scan_kernel    : per destination, exclusive scan over the tiles (one block each)
//   route_extract_kernel : tiles that hold foreign records copy them to the send buffer, stable
//   route_check_kernel   : received records must be owned by this rank; notes A[i,j]=v records
//
// Reference analogue: the per-partition buffers of GenericMTExtendableSparseMatrixCSC, summed in
// partition order (src/matrix/genericmtextendablesparsematrixcsc.jl:45-51,
// src/matrix/sparsematrixdilnkc.jl:416-426); the reference itself is single-process.
#include "xsb_internal.h"

namespace xsb {

constexpr int RT_THREADS = 256;
constexpr int RT_IPT = 8;
constexpr int RT_TILE = RT_THREADS * RT_IPT;
static_assert(RT_TILE == (1 << kRouteTileShift), "tile size of the producers' flags");

// A block looks after RT_GROUP consecutive tiles and only reads the ones the producers marked.
constexpr int RT_GROUP = RT_THREADS;

// Blocks of the tile kernels: thread x of block b looks at tile x * gridDim + b, so a block needs at most RT_GROUP
// tiles -- and a SHORT stream (an interface layer routed ahead of the interior: every tile marked) is spread over
// enough blocks to fill the GPU instead of RT_GROUP marked tiles per block.
static unsigned rt_blocks(u64 ntiles)
{
    return (unsigned)std::max<u64>((ntiles + RT_GROUP - 1) / RT_GROUP, std::min<u64>(ntiles, (u64)kNumSM * 8));
}

__global__ void __launch_bounds__(RT_THREADS)
route_count_kernel(const Rec *__restrict__ in, u64 n, u64 ntiles, int ownershift, u32 me, int nranks,
                   u32 *__restrict__ tilecnt, const unsigned char *__restrict__ tileflags)
{
    __shared__ u32 s_cnt[kMaxRanks];
    __shared__ u32 s_list[RT_GROUP];
    __shared__ u32 s_nlist;
    if (threadIdx.x == 0)
        s_nlist = 0;
    __syncthreads();
    { // tiles are dealt out round-robin: the marked ones usually sit next to each other in the stream
        const u64 t = (u64)threadIdx.x * gridDim.x + blockIdx.x;
        if (t < ntiles && (tileflags == nullptr || tileflags[t] != 0))
            s_list[atomicAdd(&s_nlist, 1u)] = (u32)threadIdx.x;
    }
    __syncthreads();
    const u32 nlist = s_nlist;
    for (u32 q = 0; q < nlist; ++q)
    {
        const u64 tile = (u64)s_list[q] * gridDim.x + blockIdx.x;
        if (threadIdx.x < kMaxRanks)
            s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const u64 b0 = tile * RT_TILE;
        bool any = false;
#pragma unroll
        for (int i = 0; i < RT_IPT; ++i)
        {
            const u64 k = b0 + (u64)i * RT_THREADS + threadIdx.x;
            if (k < n)
            {
                const u32 o = (u32)(in[k].key >> ownershift);
                if (o != me)
                {
                    atomicAdd(&s_cnt[o], 1u);
                    any = true;
                }
            }
        }
        if (__syncthreads_or(any))
        {
            if ((int)threadIdx.x < nranks)
                tilecnt[tile * nranks + threadIdx.x] = s_cnt[threadIdx.x];
        }
        __syncthreads();
    }
}

// block d: tileoff[t][d] = records for rank d in the tiles before t; total[d].
// tilecnt holds the raw counts (cnt) and tileoff the scanned offsets: the extract kernel needs both
__global__ void __launch_bounds__(1024)
route_scan_kernel(const u32 *__restrict__ tilecnt, u32 *__restrict__ tileoff, u64 ntiles, int nranks,
                  u64 *__restrict__ total)
{
    __shared__ u64 s_w[32];
    const int d = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // every thread owns a run of consecutive tiles: one block-wide scan for the whole column
    const u64 per = (ntiles + 1023) / 1024;
    const u64 t0 = (u64)threadIdx.x * per, t1 = min(t0 + per, ntiles);
    u64 sum = 0;
    for (u64 t = t0; t < t1; ++t)
        sum += tilecnt[t * nranks + d];
    u64 v = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const u64 y = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o)
            v += y;
    }
    if (lane == 31)
        s_w[warp] = v;
    __syncthreads();
    u64 run = v - sum;
    for (int w = 0; w < warp; ++w)
        run += s_w[w];
    for (u64 t = t0; t < t1; ++t)
    { // offsets of one destination stay below 2^32 (a staging buffer holds < 2^32 records)
        tileoff[t * nranks + d] = (u32)run;
        run += tilecnt[t * nranks + d];
    }
    if (threadIdx.x == 1023)
        total[d] = run;
}

// The same scan for long staging buffers, single pass with decoupled look-back: blockIdx.y = destination, a block
// (in ticket order per destination) scans RS_TILE consecutive tiles.  status / ticket: zeroed, (nblocks + 2) u64 and
// one u32 per destination.
constexpr int RS_THREADS = 256;
constexpr int RS_IPT = 8;
constexpr int RS_TILE = RS_THREADS * RS_IPT;

__global__ void __launch_bounds__(RS_THREADS)
route_scan_lookback_kernel(const u32 *__restrict__ tilecnt, u32 *__restrict__ tileoff, u64 ntiles, int nranks,
                           u64 *__restrict__ total, u64 *__restrict__ status, u32 *__restrict__ ticket, u32 nblocks)
{
    __shared__ u32 s_w[RS_THREADS / 32];
    __shared__ u64 s_prefix;
    __shared__ u32 s_blk;
    const int d = blockIdx.y;
    if (threadIdx.x == 0)
        s_blk = atomicAdd(ticket + d, 1u);
    __syncthreads();
    const u32 blk = s_blk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 t0 = (u64)blk * RS_TILE + (u64)threadIdx.x * RS_IPT;
    u32 v[RS_IPT];
    u32 sum = 0;
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i)
    {
        v[i] = t0 + i < ntiles ? tilecnt[(t0 + i) * nranks + d] : 0u;
        sum += v[i];
    }
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += y;
    }
    if (lane == 31)
        s_w[warp] = incl;
    __syncthreads();
    u32 wpre = 0, btotal = 0;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w)
    {
        const u32 c = s_w[w];
        if (w < warp)
            wpre += c;
        btotal += c;
    }
    if (warp == 0)
    {
        const u64 prefix = warp_lookback(status + (size_t)d * (nblocks + 2), blk, (u64)btotal, lane);
        if (lane == 0)
        {
            s_prefix = prefix;
            if (blk == nblocks - 1)
                total[d] = prefix + btotal;
        }
    }
    __syncthreads();
    u32 run = (u32)s_prefix + wpre + incl - sum;
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i)
    {
        if (t0 + i < ntiles)
            tileoff[(t0 + i) * nranks + d] = run;
        run += v[i];
    }
}

// bucket_addr[d] = address of the first record slot of rank d's bucket: in the send buffer, or in the mailbox of
// rank d itself (peer memory over NVLink: the copy-out IS the transfer)
__global__ void __launch_bounds__(RT_THREADS)
route_extract_kernel(const Rec *__restrict__ in, u64 n, u64 ntiles, int ownershift, u32 me, int nranks,
                     const u32 *__restrict__ tilecnt, const u32 *__restrict__ tileoff,
                     const u64 *__restrict__ bucket_addr, const u64 *__restrict__ bucket_cap)
{
    __shared__ u32 s_w[RT_THREADS / 32];
    __shared__ u32 s_run;
    __shared__ u32 s_list[RT_GROUP];
    __shared__ u32 s_nlist;
    if (threadIdx.x == 0)
        s_nlist = 0;
    __syncthreads();
    {
        const u64 t = (u64)threadIdx.x * gridDim.x + blockIdx.x;
        if (t < ntiles)
        {
            u32 has = 0;
            for (int d = 0; d < nranks; ++d)
                has |= tilecnt[t * nranks + d];
            if (has)
                s_list[atomicAdd(&s_nlist, 1u)] = (u32)threadIdx.x;
        }
    }
    __syncthreads();
    const u32 nlist = s_nlist;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 q = 0; q < nlist; ++q)
    {
        const u64 tile = (u64)s_list[q] * gridDim.x + blockIdx.x;
        const u32 *cnt = tilecnt + tile * nranks;
        const u64 b0 = tile * RT_TILE;
        Rec r[RT_IPT];
        u32 own[RT_IPT];
#pragma unroll
        for (int i = 0; i < RT_IPT; ++i)
        {
            const u64 k = b0 + (u64)i * RT_THREADS + threadIdx.x;
            own[i] = me;
            if (k < n)
            {
                r[i] = in[k];
                own[i] = (u32)(r[i].key >> ownershift);
            }
        }
        for (int d = 0; d < nranks; ++d)
        {
            if (cnt[d] == 0u) // uniform over the block
                continue;
            __syncthreads();
            if (threadIdx.x == 0)
                s_run = 0;
            __syncthreads();
            const u32 toff = tileoff[tile * nranks + d];
            Rec *dst = reinterpret_cast<Rec *>(bucket_addr[d]) + toff;
            // fixed-capacity exchange: a bucket that is fuller than its block is cut off (its header says so)
            const u64 room = bucket_cap != nullptr ? bucket_cap[d] : ~0ull;
#pragma unroll
            for (int i = 0; i < RT_IPT; ++i)
            { // round i holds RT_THREADS consecutive records: rank them in thread order
                const bool mine = own[i] == (u32)d && (u32)d != me;
                const u32 bal = __ballot_sync(0xffffffffu, mine);
                if (lane == 0)
                    s_w[warp] = __popc(bal);
                __syncthreads();
                u32 pre = s_run;
                for (int w = 0; w < warp; ++w)
                    pre += s_w[w];
                if (mine && (u64)toff + pre + __popc(bal & lanemask_lt()) < room)
                    st_rec(dst + pre + __popc(bal & lanemask_lt()), r[i]);
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    u32 t = 0;
                    for (int w = 0; w < RT_THREADS / 32; ++w)
                        t += s_w[w];
                    s_run += t;
                }
                __syncthreads();
            }
        }
    }
}

__global__ void __launch_bounds__(256)
route_check_kernel(const Rec *__restrict__ in, i64 count, int ownershift, u32 me, u64 colmask, int colshift, u64 ncols,
                   u64 *__restrict__ d_err, u64 *__restrict__ d_has_assign)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
    {
        const u64 key = in[k].key;
        if ((key & 3ull) == FL_ASSIGN && *d_has_assign == 0ull)
            *d_has_assign = 1ull; // benign race: every writer stores the same value
        if ((u32)(key >> ownershift) != me || ((key >> colshift) & colmask) >= ncols)
            atomicMin(d_err, (u64)k); // a record routed to the wrong owner
    }
}

// padding records between the regions of a slab handle's buffer: owned by "another rank", so the
// flush skips them like the records that were sent away
__global__ void __launch_bounds__(256) route_fill_kernel(Rec *__restrict__ out, i64 count, u64 key)
{
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count)
    {
        Rec r;
        r.key = key;
        r.val = 0.0;
        st_rec(out + k, r);
    }
}

void route_fill_skipped(cudaStream_t stream, Rec *out, i64 count, const KeyLayout &L, LaunchCounter &lc)
{
    if (count <= 0)
        return;
    const u64 key = (u64)((u32)L.self ^ 1u) << L.ownershift();
    route_fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(out, count, key);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

size_t route_workspace_bytes(u64 n, int nranks)
{
    const u64 ntiles = (n + RT_TILE - 1) / RT_TILE;
    const u64 sblocks = (ntiles + RS_TILE - 1) / RS_TILE;
    return 2 * sizeof(u32) * (size_t)ntiles * nranks + 3 * sizeof(u64) * kMaxRanks + 256 +
           sizeof(u64) * (size_t)(sblocks + 2) * kMaxRanks + sizeof(u32) * kMaxRanks + 64;
}

// counts_host[d] = staged records owned by rank d (d == me: the ones that stay)
void route_count(cudaStream_t stream, const Rec *in, u64 n, const KeyLayout &L, void *workspace, u64 *counts_host,
                 LaunchCounter &lc, const unsigned char *tileflags)
{
    const int nr = L.nranks;
    for (int d = 0; d < nr; ++d)
        counts_host[d] = 0;
    if (n == 0)
        return;
    const u64 ntiles = (n + RT_TILE - 1) / RT_TILE;
    u32 *tilecnt = static_cast<u32 *>(workspace);
    u32 *tileoff = tilecnt + (size_t)ntiles * nr;
    u64 *total = reinterpret_cast<u64 *>(tileoff + (size_t)ntiles * nr);
    XSB_CUDA(cudaMemsetAsync(tilecnt, 0, sizeof(u32) * (size_t)ntiles * nr, stream));
    route_count_kernel<<<rt_blocks(ntiles), RT_THREADS, 0, stream>>>(
        in, n, ntiles, L.ownershift(), (u32)L.self, nr, tilecnt, tileflags);
    route_scan_kernel<<<nr, 1024, 0, stream>>>(tilecnt, tileoff, ntiles, nr, total);
    lc.add(2);
    XSB_CUDA(cudaGetLastError());
    u64 h[kMaxRanks];
    XSB_CUDA(cudaMemcpyAsync(h, total, sizeof(u64) * nr, cudaMemcpyDeviceToHost, stream));
    XSB_CUDA(cudaStreamSynchronize(stream));
    u64 foreign = 0;
    for (int d = 0; d < nr; ++d)
    {
        counts_host[d] = h[d];
        foreign += h[d];
    }
    counts_host[L.self] = n - foreign;
}

// after route_count on the same records and workspace: foreign records -> send (destination after
// destination, each bucket in stream order)
void route_extract(cudaStream_t stream, const Rec *in, u64 n, const KeyLayout &L, void *workspace,
                   const u64 *counts_host, Rec *send, LaunchCounter &lc)
{
    if (n == 0)
        return;
    const int nr = L.nranks;
    const u64 ntiles = (n + RT_TILE - 1) / RT_TILE;
    u32 *tilecnt = static_cast<u32 *>(workspace);
    u32 *tileoff = tilecnt + (size_t)ntiles * nr;
    u64 *total = reinterpret_cast<u64 *>(tileoff + (size_t)ntiles * nr);
    u64 *bucket_base = total + kMaxRanks;
    u64 hb[kMaxRanks];
    u64 run = 0;
    for (int d = 0; d < nr; ++d)
    {
        hb[d] = reinterpret_cast<u64>(send + run);
        if (d != L.self)
            run += counts_host[d];
    }
    if (run == 0)
        return;
    XSB_CUDA(cudaMemcpyAsync(bucket_base, hb, sizeof(u64) * nr, cudaMemcpyHostToDevice, stream));
    route_extract_kernel<<<rt_blocks(ntiles), RT_THREADS, 0, stream>>>(
        in, n, ntiles, L.ownershift(), (u32)L.self, nr, tilecnt, tileoff, bucket_base, nullptr);
    lc.add();
    XSB_CUDA(cudaGetLastError());
    XSB_CUDA(cudaStreamSynchronize(stream)); // hb lives on this stack frame
}

// ------------------------------------------------------------------------
// Fixed-capacity exchange: no count ever visits the host.  The send buffer holds one BLOCK per destination
// (own rank: none): a header record {records of the bucket, magic} followed by cap[d] record slots.  Sender and
// receiver agree on the capacities beforehand (e.g. from the previous step of an assembly loop), so the
// all-to-all runs with host-known split sizes and the whole step is stream-ordered.
// ------------------------------------------------------------------------
constexpr u64 kRouteMagic = 0x5853425f524f5554ull; // "XSB_ROUT"

// flags of a mailbox are read and written by two GPUs
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p)
{
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 global_ns()
{
    u64 t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Headers of the blocks (bucket_addr[d] != 0: this rank sends to d), written behind the extraction in stream
// order.  Peer exchange: the block lives in rank d's mailbox; once the header is out, `seq` is stored with
// release semantics at system scope into d's ready flag: everything this stream wrote to d's memory before
// -- the records of the extraction kernel, the header -- is visible to a reader on d that saw the flag.
__global__ void route_header_kernel(const u64 *__restrict__ total, const u64 *__restrict__ bucket_addr,
                                    const u64 *__restrict__ bucket_cap, int nranks, u32 me, PeerFlags sig,
                                    u64 *__restrict__ d_flags)
{
    const int d = threadIdx.x;
    if (d < nranks && (u32)d != me && total[d] > bucket_cap[d]) // also for ranks without a block (capacity 0)
        atomicOr(reinterpret_cast<unsigned long long *>(d_flags), 2ull);
    if (d < nranks && (u32)d != me && bucket_addr[d] != 0ull)
    {
        Rec r;
        r.key = total[d];
        r.val = __longlong_as_double((long long)kRouteMagic);
        st_rec(reinterpret_cast<Rec *>(bucket_addr[d]) - 1, r);
        if (sig.addr[d] != 0ull)
        {
            __threadfence_system();
            st_release_sys(reinterpret_cast<u64 *>(sig.addr[d]), sig.value[d]);
        }
    }
}

// One thread per flag: waits until the flag (in THIS GPU's memory, written by a peer) reaches value[k].  Polite
// polling; gives up after timeout_ns and raises bit 3 of *d_flags (the flush then fails instead of hanging).
__global__ void peer_wait_kernel(PeerFlags w, u64 timeout_ns, u64 *__restrict__ d_flags)
{
    const int k = threadIdx.x;
    if (k >= kMaxRanks || w.addr[k] == 0ull)
        return;
    const u64 *p = reinterpret_cast<const u64 *>(w.addr[k]);
    const u64 t0 = global_ns();
    unsigned ns = 32;
    while (ld_acquire_sys(p) < w.value[k])
    {
        __nanosleep(ns);
        ns = min(ns * 2u, 2048u);
        if (global_ns() - t0 > timeout_ns)
        {
            atomicOr(reinterpret_cast<unsigned long long *>(d_flags), 8ull);
            break;
        }
    }
}

// value[k] -> flag k (in a peer's memory), after everything this stream did before
__global__ void peer_signal_kernel(PeerFlags sgn)
{
    const int k = threadIdx.x;
    if (k < kMaxRanks && sgn.addr[k] != 0ull)
    {
        __threadfence_system();
        st_release_sys(reinterpret_cast<u64 *>(sgn.addr[k]), sgn.value[k]);
    }
}

// Records [from, n) of the staged stream, staged after an early xsb_route_pack_peer: every one must be owned by this
// rank.  A warp per tile the producers flagged (normally only the tile `from` lies in); bit 5 of *d_flags otherwise.
__global__ void __launch_bounds__(256)
route_tailcheck_kernel(const Rec *__restrict__ in, u64 from, u64 n, int ownershift, u32 me,
                       const unsigned char *__restrict__ tileflags, u64 *__restrict__ d_flags)
{
    const u64 tile = (from >> kRouteTileShift) + (u64)blockIdx.x * 8 + (threadIdx.x >> 5);
    const u64 b0 = tile << kRouteTileShift;
    if (b0 >= n || (tileflags != nullptr && tileflags[tile] == 0))
        return;
    const u64 e = min(n, b0 + RT_TILE);
    bool bad = false;
    for (u64 k = max(b0, from) + (threadIdx.x & 31); k < e; k += 32)
        bad |= (u32)(in[k].key >> ownershift) != me;
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0)
        atomicOr(reinterpret_cast<unsigned long long *>(d_flags), 32ull);
}

void route_tailcheck(cudaStream_t stream, const Rec *in, u64 from, u64 n, const KeyLayout &L, const unsigned char *tileflags,
                     u64 *d_flags, LaunchCounter &lc)
{
    if (from >= n)
        return;
    const u64 tiles = ((n + RT_TILE - 1) >> kRouteTileShift) - (from >> kRouteTileShift);
    route_tailcheck_kernel<<<(unsigned)((tiles + 7) / 8), 256, 0, stream>>>(in, from, n, L.ownershift(), (u32)L.self, tileflags,
                                                                          d_flags);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

void peer_wait(cudaStream_t stream, const PeerFlags &w, u64 timeout_ns, u64 *d_flags, LaunchCounter &lc)
{
    peer_wait_kernel<<<1, kMaxRanks, 0, stream>>>(w, timeout_ns, d_flags);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

void peer_signal(cudaStream_t stream, const PeerFlags &sgn, LaunchCounter &lc)
{
    peer_signal_kernel<<<1, kMaxRanks, 0, stream>>>(sgn);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

// pinned_addr / pinned_caps: host staging (pinned, kMaxRanks entries each) that outlives the call.
// pinned_addr[d] (filled by the caller) = address of the first record slot of the block for rank d (its header sits
// one record below), 0 = nothing is sent to d; caps[d] = slots of that block.  sig: flags to raise once a block is
// complete (peer exchange), all-zero otherwise.  A bucket that outgrew its block raises bit 1 of *d_flags on the
// SENDER too (the receiver sees it in the header).
void route_pack(cudaStream_t stream, const Rec *in, u64 n, const KeyLayout &L, void *workspace, const i64 *caps,
                u64 *pinned_addr, u64 *pinned_caps, const PeerFlags &sig, u64 *d_flags, LaunchCounter &lc,
                const unsigned char *tileflags)
{
    const int nr = L.nranks;
    const u64 ntiles = (n + RT_TILE - 1) / RT_TILE;
    u32 *tilecnt = static_cast<u32 *>(workspace);
    u32 *tileoff = tilecnt + (size_t)ntiles * nr;
    u64 *total = reinterpret_cast<u64 *>(tileoff + (size_t)ntiles * nr);
    u64 *bucket_addr = total + kMaxRanks;
    u64 *bucket_cap = bucket_addr + kMaxRanks;
    for (int d = 0; d < nr; ++d)
        pinned_caps[d] = (d == L.self || pinned_addr[d] == 0ull) ? 0ull : (u64)caps[d];
    XSB_CUDA(cudaMemcpyAsync(bucket_addr, pinned_addr, sizeof(u64) * nr, cudaMemcpyHostToDevice, stream));
    XSB_CUDA(cudaMemcpyAsync(bucket_cap, pinned_caps, sizeof(u64) * nr, cudaMemcpyHostToDevice, stream));
    XSB_CUDA(cudaMemsetAsync(total, 0, sizeof(u64) * kMaxRanks, stream));
    if (n > 0)
    {
        XSB_CUDA(cudaMemsetAsync(tilecnt, 0, sizeof(u32) * (size_t)ntiles * nr, stream));
        route_count_kernel<<<rt_blocks(ntiles), RT_THREADS, 0, stream>>>(
            in, n, ntiles, L.ownershift(), (u32)L.self, nr, tilecnt, tileflags);
        { // offsets of every tile's records in their bucket: single-pass scan per destination
            const u32 sblocks = (u32)((ntiles + RS_TILE - 1) / RS_TILE);
            u64 *status = bucket_cap + kMaxRanks + 4;
            u32 *ticket = reinterpret_cast<u32 *>(status + (size_t)(sblocks + 2) * kMaxRanks);
            XSB_CUDA(cudaMemsetAsync(status, 0, sizeof(u64) * (size_t)(sblocks + 2) * kMaxRanks + sizeof(u32) * kMaxRanks, stream));
            route_scan_lookback_kernel<<<dim3(sblocks, (unsigned)nr), RS_THREADS, 0, stream>>>(tilecnt, tileoff, ntiles, nr, total,
                                                                                             status, ticket, sblocks);
        }
        route_extract_kernel<<<rt_blocks(ntiles), RT_THREADS, 0, stream>>>(
            in, n, ntiles, L.ownershift(), (u32)L.self, nr, tilecnt, tileoff, bucket_addr, bucket_cap);
        lc.add(3);
    }
    route_header_kernel<<<1, kMaxRanks, 0, stream>>>(total, bucket_addr, bucket_cap, nr, (u32)L.self, sig, d_flags);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

// One received block -> `cap` records behind the staged ones: the bucket's records (checked: owned by this rank),
// then padding records the flush skips.  flags: bit 0 a record of another owner, bit 1 the bucket did not fit its
// block (records were cut off), bit 2 not a block (bad magic), bit 4 an A[i,j] = v record was received.  counts[which] += records taken.
__global__ void __launch_bounds__(256)
route_unpack_kernel(const Rec *__restrict__ block, u64 cap, Rec *__restrict__ out, int ownershift, u32 me, u64 colmask,
                    int colshift, u64 ncols, u64 padkey, u64 *__restrict__ flags, u64 *__restrict__ counts, int which)
{
    const Rec hdr = block[0];
    const u64 count = hdr.key;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        u64 f = 0;
        if ((u64)__double_as_longlong(hdr.val) != kRouteMagic)
            f |= 4ull;
        else if (count > cap)
            f |= 2ull;
        if (f)
            atomicOr(reinterpret_cast<unsigned long long *>(flags), (unsigned long long)f);
        atomicAdd(reinterpret_cast<unsigned long long *>(counts + which), (unsigned long long)min(count, cap));
    }
    const u64 take = (u64)__double_as_longlong(hdr.val) == kRouteMagic ? min(count, cap) : 0ull;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < cap; k += stride)
    {
        Rec r;
        if (k < take)
        {
            r = block[1 + k];
            if ((u32)(r.key >> ownershift) != me || ((r.key >> colshift) & colmask) >= ncols)
                atomicOr(reinterpret_cast<unsigned long long *>(flags), 1ull);
            if ((r.key & 3ull) == FL_ASSIGN && (*reinterpret_cast<volatile u64 *>(flags) & 16ull) == 0ull)
                atomicOr(reinterpret_cast<unsigned long long *>(flags), 16ull);
        }
        else
        {
            r.key = padkey;
            r.val = 0.0;
        }
        st_rec(out + k, r);
    }
}

void route_unpack(cudaStream_t stream, const Rec *block, i64 cap, Rec *out, const KeyLayout &L, i64 ncols, u64 *d_flags,
                  u64 *d_counts, int which, LaunchCounter &lc)
{
    const u64 padkey = (u64)((u32)L.self ^ 1u) << L.ownershift();
    const int blocks = (int)std::max<i64>(1, std::min<i64>((cap + 255) / 256, (i64)kNumSM * 8));
    route_unpack_kernel<<<blocks, 256, 0, stream>>>(block, (u64)cap, out, L.ownershift(), (u32)L.self,
                                                    (1ull << L.colbits) - 1ull, L.low + L.rowbits, (u64)ncols, padkey, d_flags,
                                                    d_counts, which);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

void route_check(cudaStream_t stream, const Rec *in, i64 count, const KeyLayout &L, i64 ncols, u64 *d_err,
                 u64 *d_has_assign, LaunchCounter &lc)
{
    if (count <= 0)
        return;
    const int blocks = (int)std::min<i64>((count + 255) / 256, (i64)kNumSM * 16);
    route_check_kernel<<<blocks, 256, 0, stream>>>(in, count, L.ownershift(), (u32)L.self, (1ull << L.colbits) - 1ull,
                                                   L.low + L.rowbits, (u64)ncols, d_err, d_has_assign);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

} // namespace xsb
