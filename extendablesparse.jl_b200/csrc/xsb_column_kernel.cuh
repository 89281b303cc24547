// xsb_column_kernel.cuh -- column_reduce_kernel (included by xsb_column.cu only).
//
// Phases of one tile (CK_T nominal records, CK_CAP staged):
//   1. stage the records in shared memory, flag column starts, build the packed
//      (row, position-in-column) sort keys;
//   2. one warp per owned column: register bitonic sort of the keys, records permuted in place
//      into (row, stream) order;
//   3. block-wide: flag the heads of the runs of equal rows, compact them, ONE RUN PER THREAD is
//      folded sequentially (bit-exact order), created entries are compacted;
//   4. decoupled look-back over the tiles' entry counts, coalesced write of rowval/nzval,
//      per-column entry counts to colcount.
#pragma once

template <typename Ti, bool SIMPLE>
__global__ void __launch_bounds__(CK_THREADS, 3)
column_reduce_kernel(const Rec *__restrict__ sorted, u64 nrec, KeyLayout L, int posbits, int combine, Ti base,
                     Ti *__restrict__ rowval, double *__restrict__ nzval, u32 *__restrict__ colcount,
                     u64 *__restrict__ status, u32 *__restrict__ tile_counter, u64 *__restrict__ d_nnz,
                     u32 *__restrict__ d_overflow, u32 ntiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec *s_rec = reinterpret_cast<Rec *>(smem_raw);                            // CK_CAP records
    u32 *s_key = reinterpret_cast<u32 *>(s_rec + CK_CAP);                      // sort keys; then the run heads
    unsigned short *s_cs = reinterpret_cast<unsigned short *>(s_key + CK_CAP); // column starts (CK_CAP + 8)
    u32 *s_cc = reinterpret_cast<u32 *>(s_cs + CK_CAP + 8);                    // entries per owned column, 2 x u16 per word
    __shared__ u32 s_cnt[CK_IPT * CK_WARPS];
    __shared__ u32 s_total, s_tile, s_next;
    __shared__ u64 s_prev, s_tileoff;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
    {
        s_tile = atomicAdd(tile_counter, 1u);
        s_next = 0;
    }
    __syncthreads();
    const u32 tile = s_tile;
    const u64 g0 = (u64)tile * CK_T;
    const u32 avail = (u32)min((u64)CK_CAP, nrec - g0);
    const bool at_end = g0 + avail == nrec;

#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
    {
        const u32 e = i * CK_THREADS + tid;
        if (e < avail)
            s_rec[e] = ld_rec_stream(sorted + g0 + e);
    }
    if (tid == 0)
        s_prev = g0 > 0 ? sorted[g0 - 1].key : 0ull;
    if (tid < CK_T / 8) // zero the per-column counts (CK_T u16 = CK_T/8 uint4)
        reinterpret_cast<uint4 *>(s_cc)[tid] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // ---- 1. column starts
    const u32 lt = lanemask_lt();
    bool flag[CK_IPT];
    u32 excl[CK_IPT];
#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
    {
        const u32 e = i * CK_THREADS + tid;
        bool st = false;
        if (e < avail)
        {
            const u64 c = L.col(s_rec[e].key);
            st = e > 0 ? c != L.col(s_rec[e - 1].key) : (g0 == 0 || c != L.col(s_prev));
        }
        flag[i] = st;
    }
    const u32 nstarts = block_rank<CK_IPT, CK_WARPS>(flag, excl, s_cnt, &s_total, lane, warp, lt);
#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
        if (flag[i])
            s_cs[excl[i]] = (unsigned short)(i * CK_THREADS + tid);
    __syncthreads();
    // columns that start inside the nominal tile are owned: the first `nown` of the sorted list
    u32 nown = 0;
    {
        u32 lo = 0, hi = nstarts;
        while (lo < hi)
        {
            const u32 mid = (lo + hi) >> 1;
            if (s_cs[mid] < CK_T)
                lo = mid + 1;
            else
                hi = mid;
        }
        nown = lo;
    }
    u32 own_begin = 0, own_end = 0; // owned records: [own_begin, own_end)
    bool overflow = false;
    if (nown > 0)
    {
        own_begin = s_cs[0];
        if (nown < nstarts)
            own_end = s_cs[nown];
        else if (at_end)
            own_end = avail;
        else
            overflow = true; // the last owned column runs past the look-ahead
    }
    const u32 posmask = (1u << posbits) - 1u;
    int ci[CK_IPT]; // column (index into s_cs) of each of this thread's records, -1 / >= nown: not owned
#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
    {
        const u32 e = i * CK_THREADS + tid;
        ci[i] = (int)excl[i] + (flag[i] ? 1 : 0) - 1;
        if (e < avail && ci[i] >= 0 && (u32)ci[i] < nown)
        {
            const u32 pos = e - s_cs[ci[i]];
            s_key[e] = ((u32)L.row(s_rec[e].key) << posbits) | (pos & posmask);
        }
    }
    __syncthreads();

    // ---- 2. one warp per owned column: sort by (row, position), permute the records in place
    const u32 maxlen = min(32u * CK_MAXE, 1u << posbits);
    if (!overflow)
    {
        for (;;)
        {
            u32 k = 0;
            if (lane == 0)
                k = atomicAdd(&s_next, 1u);
            k = __shfl_sync(0xffffffffu, k, 0);
            if (k >= nown)
                break;
            const u32 c0 = s_cs[k];
            const u32 c1 = (k + 1 < nown) ? s_cs[k + 1] : own_end;
            const u32 len = c1 - c0;
            if (len > maxlen)
            {
                overflow = true; // warp-uniform
                break;
            }
            if (len <= 1)
                continue;
            if (len <= 32)
                sort_column<1>(s_key + c0, s_rec + c0, len, posmask, lane);
            else if (len <= 64)
                sort_column<2>(s_key + c0, s_rec + c0, len, posmask, lane);
            else if (len <= 128)
                sort_column<4>(s_key + c0, s_rec + c0, len, posmask, lane);
            else
                sort_column<8>(s_key + c0, s_rec + c0, len, posmask, lane);
        }
    }
    if (overflow && lane == 0)
        atomicExch(d_overflow, 1u);
    __syncthreads();
    // an overflowing tile still walks through the rest (its output is discarded by the caller)

    // ---- 3. heads of the runs of equal rows inside the owned columns
#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
    {
        const u32 e = i * CK_THREADS + tid;
        bool head = false;
        if (e >= own_begin && e < own_end)
            head = flag[i] || L.row(s_rec[e].key) != L.row(s_rec[e - 1].key);
        flag[i] = head;
    }
    const u32 nruns = block_rank<CK_IPT, CK_WARPS>(flag, excl, s_cnt, &s_total, lane, warp, lt);
#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
        if (flag[i])
            s_key[excl[i]] = (u32)(i * CK_THREADS + tid) | ((u32)ci[i] << 16); // position | column << 16
    __syncthreads();

    // one run per thread and round
    double myval[CK_IPT];
    u32 myrow[CK_IPT];
#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
    {
        const u32 u = i * CK_THREADS + tid;
        bool created = false;
        double out = 0.0;
        u32 row = 0;
        if (i * CK_THREADS < nruns) // uniform: rounds beyond the last run are skipped by the whole block
        {
            if (u < nruns)
            {
                const u32 hk = s_key[u];
                const u32 q0 = hk & 0xffffu;
                const u32 q1 = (u + 1 < nruns) ? (s_key[u + 1] & 0xffffu) : own_end;
                const Rec first = s_rec[q0];
                row = (u32)L.row(first.key);
                if (SIMPLE)
                { // an old CSC entry can only be the first record of its run (it precedes the stream)
                    const u32 fl0 = L.flavour(first.key);
                    double acc = (fl0 == FL_OLD) ? first.val : 0.0 + first.val;
                    created = (fl0 != FL_UPDATE) | (first.val != 0.0);
                    for (u32 q = q0 + 1; q < q1; ++q)
                    {
                        const Rec r = s_rec[q];
                        acc = acc + r.val;
                        created |= (L.flavour(r.key) != FL_UPDATE) | (r.val != 0.0);
                    }
                    out = acc;
                }
                else
                {
                    ColFold f;
                    for (u32 q = q0; q < q1; ++q)
                    {
                        const Rec r = s_rec[q];
                        f.apply(L.flavour(r.key), L.tid(r.key), r.val, combine);
                    }
                    f.finish();
                    created = f.exists;
                    out = f.acc;
                }
                if (created)
                { // count the entry for its column (two 16-bit counters per word)
                    const u32 c = hk >> 16;
                    atomicAdd(&s_cc[c >> 1], (c & 1u) ? 0x10000u : 1u);
                }
            }
        }
        flag[i] = created;
        myval[i] = out;
        myrow[i] = row;
    }
    const u32 total = block_rank<CK_IPT, CK_WARPS>(flag, excl, s_cnt, &s_total, lane, warp, lt);

    // ---- 4. decoupled look-back over the tiles' entry counts
    if (warp == 0)
    {
        u64 prefix = 0;
        if (tile == 0)
        {
            if (lane == 0)
                st_relaxed_u64(status, CS_INCL | (u64)total);
        }
        else
        {
            if (lane == 0)
                st_relaxed_u64(status + tile, CS_LOCAL | (u64)total);
            i64 t = (i64)tile - 1;
            for (;;)
            {
                const i64 idx = t - lane;
                u64 v = CS_INCL;
                if (idx >= 0)
                {
                    do
                    {
                        v = ld_relaxed_u64(status + idx);
                    } while ((v >> 62) == 0ull);
                }
                const u32 incl_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                u64 contrib = v & CS_VALUE;
                if (incl_mask)
                {
                    const int first = __ffs(incl_mask) - 1;
                    if (lane > first)
                        contrib = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                prefix += contrib;
                if (incl_mask)
                    break;
                t -= 32;
            }
            if (lane == 0)
                st_relaxed_u64(status + tile, CS_INCL | (prefix + total));
        }
        if (lane == 0)
        {
            s_tileoff = prefix;
            if (tile == ntiles - 1)
                *d_nnz = prefix + total;
        }
    }
    __syncthreads();
    const u64 tileoff = s_tileoff;
#pragma unroll
    for (int i = 0; i < CK_IPT; ++i)
    {
        if (flag[i])
        {
            const u64 p = tileoff + excl[i];
            rowval[p] = (Ti)myrow[i] + base;
            nzval[p] = myval[i];
        }
    }
    // per-column entry counts (every column is owned by exactly one tile: plain stores)
    for (u32 k = tid; k < nown; k += CK_THREADS)
    {
        const u32 w = s_cc[k >> 1];
        colcount[L.col(s_rec[s_cs[k]].key)] = (k & 1u) ? (w >> 16) : (w & 0xffffu);
    }
}
