// xsb_mul.cu -- y = A*x on the resident CSC (the step right after assembly
// in every solver loop).
//
// Reference: mul!(r, A, x) flushes and delegates to SparseArrays
// (src/matrix/abstractextendablesparsematrixcsc.jl:170-181); the stdlib kernel walks the columns in
// ascending order and does y[rowval[k]] += nzval[k] * x[j] -- so y[i] is the left fold, over the
// entries of row i in COLUMN order, of separately rounded products.  (The MT wrapper's threaded
// variant, src/matrix/genericmtextendablesparsematrixcsc.jl:124-143, sums the same terms partition
// by partition; its own test only asks for 1e-8, test/test_parallel.jl:94-118.)
//
// To reproduce that bit for bit a row's terms must be added in column order by ONE thread, which
// needs the row-major view of the pattern.  It is built once per pattern (and dropped when a flush
// changes it): the entries tagged (row, CSC position, column) go through the library's stable radix
// sort on the row bits; the sorted tags give rowptr, and per entry its column and where its value
// lives in nzval.  The product itself is then one thread per row, no atomics.
// Compiled with -fmad=false: product and sum round separately, like the reference.
#include "xsb_internal.h"

namespace xsb {

template <typename Ti>
__global__ void __launch_bounds__(256)
csr_tag_kernel(const Ti *__restrict__ colptr, const Ti *__restrict__ rowval, i64 n, Ti base, Rec *__restrict__ out)
{
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = warp; j < n; j += nwarps)
    {
        const i64 s = (i64)colptr[j] - base, e = (i64)colptr[j + 1] - base;
        for (i64 k = s + lane; k < e; k += 32)
        {
            Rec r;
            r.key = (u64)((i64)rowval[k] - base);
            r.val = __longlong_as_double((long long)(((u64)j << 32) | (u64)k)); // column | CSC position
            st_rec(out + k, r);
        }
    }
}

// sorted tags -> column and nzval position of every entry in row-major order, rowptr
__global__ void __launch_bounds__(256)
csr_emit_kernel(const Rec *__restrict__ tags, i64 nnz, i64 m, u32 *__restrict__ rowptr, u32 *__restrict__ csr_col,
                u32 *__restrict__ csr_src)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride)
    {
        const Rec r = tags[e];
        const u64 pl = (u64)__double_as_longlong(r.val);
        csr_col[e] = (u32)(pl >> 32);
        csr_src[e] = (u32)pl;
        const i64 row = (i64)r.key;
        const i64 prev = e ? (i64)tags[e - 1].key : -1;
        for (i64 q = prev + 1; q <= row; ++q) // this row and the empty rows right before it start here
            rowptr[q] = (u32)e;
        if (e == nnz - 1)
            for (i64 q = row + 1; q <= m; ++q)
                rowptr[q] = (u32)nnz;
    }
}

__global__ void __launch_bounds__(256) csr_empty_kernel(i64 m, u32 *__restrict__ rowptr)
{
    const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= m)
        rowptr[q] = 0u;
}

size_t csr_map_bytes(i64 m, i64 nnz)
{
    return sizeof(u32) * ((size_t)m + 1 + 2 * (size_t)std::max<i64>(nnz, 1)) + 64;
}

// map = [rowptr (m+1) | csr_col (nnz) | csr_src (nnz)]; tags_a / tags_b: nnz records of scratch each
void build_csr_map(cudaStream_t stream, const CscView &csc, i64 m, i64 n, int idx64, int base, Rec *tags_a, Rec *tags_b,
                   void *sort_ws, u32 *map, LaunchCounter &lc)
{
    u32 *rowptr = map;
    u32 *csr_col = map + (m + 1);
    u32 *csr_src = csr_col + std::max<i64>(csc.nnz, 1);
    if (csc.nnz == 0)
    {
        csr_empty_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, stream>>>(m, rowptr);
        lc.add();
        return;
    }
    const int threads = 256;
    const i64 want = (n * 32 + threads - 1) / threads;
    const int blocks = (int)std::min<i64>(std::max<i64>(want, 1), (i64)kNumSM * 16);
    if (idx64)
        csr_tag_kernel<int64_t><<<blocks, threads, 0, stream>>>((const int64_t *)csc.colptr, (const int64_t *)csc.rowval,
                                                                n, (int64_t)base, tags_a);
    else
        csr_tag_kernel<int32_t><<<blocks, threads, 0, stream>>>((const int32_t *)csc.colptr, (const int32_t *)csc.rowval,
                                                                n, (int32_t)base, tags_a);
    lc.add();
    int rowbits = 1;
    while ((1ll << rowbits) < m)
        ++rowbits;
    const SortPlan plan = make_sort_plan(0, rowbits);
    const Rec *sorted = radix_sort_records(stream, tags_a, tags_b, (u64)csc.nnz, plan, sort_ws, lc, nullptr);
    const int eblocks = (int)std::min<i64>((csc.nnz + 255) / 256, (i64)kNumSM * 16);
    csr_emit_kernel<<<eblocks, 256, 0, stream>>>(sorted, csc.nnz, m, rowptr, csr_col, csr_src);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

// one thread per row: y[i] = (((0 + a_ij1 x_j1) + a_ij2 x_j2) + ...), j ascending
__global__ void __launch_bounds__(128)
csr_mul_kernel(const u32 *__restrict__ rowptr, const u32 *__restrict__ csr_col, const u32 *__restrict__ csr_src,
               const double *__restrict__ nzval, const double *__restrict__ x, i64 m, double *__restrict__ y)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m)
        return;
    const u32 b = rowptr[i], e = rowptr[i + 1];
    double acc = 0.0;
    u32 k = b;
    for (; k + 4 <= e; k += 4)
    { // the gathers of four terms travel together; the adds stay in order
        const double p0 = nzval[csr_src[k]] * x[csr_col[k]];
        const double p1 = nzval[csr_src[k + 1]] * x[csr_col[k + 1]];
        const double p2 = nzval[csr_src[k + 2]] * x[csr_col[k + 2]];
        const double p3 = nzval[csr_src[k + 3]] * x[csr_col[k + 3]];
        acc = (((acc + p0) + p1) + p2) + p3;
    }
    for (; k < e; ++k)
        acc = acc + nzval[csr_src[k]] * x[csr_col[k]];
    y[i] = acc;
}

void csr_mul(cudaStream_t stream, const u32 *map, i64 m, i64 nnz, const double *nzval, const double *x, double *y,
             LaunchCounter &lc)
{
    const u32 *rowptr = map;
    const u32 *csr_col = map + (m + 1);
    const u32 *csr_src = csr_col + std::max<i64>(nnz, 1);
    csr_mul_kernel<<<(unsigned)((m + 127) / 128), 128, 0, stream>>>(rowptr, csr_col, csr_src, nzval, x, m, y);
    lc.add();
    XSB_CUDA(cudaGetLastError());
}

} // namespace xsb
