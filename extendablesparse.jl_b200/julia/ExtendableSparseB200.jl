# ExtendableSparseB200.jl -- Julia-side binding of libxsparse_b200.so.
#
# Plugs the CUDA library into ExtendableSparse.jl through the package's own extension
# interface (src/matrix/abstractsparsematrixextension.jl:1-19): `SparseMatrixB200` is an
# `AbstractSparseMatrixExtension`, and the user-facing matrix types are obtained exactly like
# the package's own aliases (src/ExtendableSparse.jl:35-39):
#
#     const B200ExtendableSparseMatrixCSC{Tv,Ti} =
#         GenericMTExtendableSparseMatrixCSC{SparseMatrixB200{Tv,Ti},Tv,Ti}
#
# NOTE: no Julia runtime exists in the build container of this repository, so this file has
# been written against the reference sources but NOT executed.  The C ABI it binds is
# exercised by the Python tests (tests/test_gpu_parity.py) through ctypes.
#
# Per-entry calls are appended to a per-partition host buffer of 16-byte triplets
# (`XsbTriplet` = `xsb_triplet` of include/xsparse_b200.h: one 16-byte store per call, no ccall per
# entry) and shipped with one `xsb_insert_triplets` when the buffer is full, the flavour changes, or
# at `flush!`: 16 bytes per insertion cross PCIe instead of the 24 of three Int64/Int64/Float64 arrays.

module ExtendableSparseB200

using SparseArrays
using ExtendableSparse
import ExtendableSparse: AbstractSparseMatrixExtension, rawupdateindex!, updateindex!, flush!, reset!

const libxsb = get(ENV, "XSPARSE_B200_LIB", "libxsparse_b200.so")

# status codes / enums of include/xsparse_b200.h
const XSB_OK, XSB_EBOUNDS, XSB_ESIZE, XSB_EILLEGAL = Int32(0), Int32(1), Int32(2), Int32(3)
const XSB_F64, XSB_I32, XSB_I64 = Int32(0), Int32(0), Int32(1)
const XSB_UPDATE, XSB_RAW, XSB_ASSIGN = Int32(0), Int32(1), Int32(2)
const XSB_DETERMINISTIC, XSB_FAST = Int32(0), Int32(1)
const XSB_COMBINE_SEED, XSB_COMBINE_ADD = Int32(0), Int32(1)

const CHUNK = 1 << 16

function check(h::Ptr{Cvoid}, rc::Int32)
    rc == XSB_OK && return nothing
    msg = unsafe_string(ccall((:xsb_last_error, libxsb), Cstring, (Ptr{Cvoid},), h))
    rc == XSB_EBOUNDS && throw(BoundsError())                    # sparsematrixcsc.jl:8-10
    rc == XSB_ESIZE && throw(AssertionError(msg))                # sparsematrixlnk.jl:296-297
    error("libxsparse_b200: [$rc] $msg")                        # genericmt...:67,80
end

"`xsb_triplet`: what one updateindex!/rawupdateindex!/setindex! call appends (1-based row/col)."
struct XsbTriplet
    row::UInt32
    col::UInt32
    val::Float64
end

idxcode(::Type{Int64}) = XSB_I64
idxcode(::Type{Int32}) = XSB_I32

"""
    SparseMatrixB200{Tv,Ti}(m, n)

Insert buffer living in B200 HBM.  Constructor signature `T_ext(m,n)` as required by
abstractsparsematrixextension.jl:10.
"""
mutable struct SparseMatrixB200{Tv, Ti <: Integer} <: AbstractSparseMatrixExtension{Tv, Ti}
    m::Ti
    n::Ti
    handle::Ptr{Cvoid}
    T::Vector{XsbTriplet}
    fill::Int
    flavour::Int32
    shipped::Int            # insertions already on the device

    function SparseMatrixB200{Tv, Ti}(m, n) where {Tv, Ti <: Integer}
        Tv === Float64 || error("libxsparse_b200 implements Float64 values")
        (m < 2^32 && n < 2^32) || error("triplet buffers carry 32-bit indices")
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:xsb_create, libxsb), Int32,
                   (Int64, Int64, Int32, Int32, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
                   m, n, XSB_F64, idxcode(Ti), 1, 1, 0, h)
        check(Ptr{Cvoid}(C_NULL), rc)
        x = new{Tv, Ti}(m, n, h[], Vector{XsbTriplet}(undef, CHUNK), 0, XSB_RAW, 0)
        finalizer(x) do y
            y.handle == C_NULL || ccall((:xsb_destroy, libxsb), Int32, (Ptr{Cvoid},), y.handle)
            y.handle = C_NULL
        end
        x
    end
end

Base.size(x::SparseMatrixB200) = (x.m, x.n)
# upper bound of the distinct new entries, enough for `nnz(ext)>0` in flush!
# (genericextendablesparsematrixcsc.jl:31-37)
SparseArrays.nnz(x::SparseMatrixB200) = x.shipped + x.fill

function ship!(x::SparseMatrixB200)
    x.fill == 0 && return
    n = x.fill
    x.fill = 0
    rc = ccall((:xsb_insert_triplets, libxsb), Int32,
               (Ptr{Cvoid}, Int32, Ptr{XsbTriplet}, Int64, Int32),
               x.handle, 0, x.T, n, x.flavour)
    check(x.handle, rc)
    x.shipped += n
end

@inline function push_entry!(x::SparseMatrixB200{Tv, Ti}, flavour::Int32, v, i, j) where {Tv, Ti}
    (1 <= i <= x.m && 1 <= j <= x.n) || throw(BoundsError(x, (i, j)))
    if x.fill == CHUNK || (x.fill > 0 && x.flavour != flavour)
        ship!(x)
    end
    x.flavour = flavour
    k = (x.fill += 1)
    @inbounds x.T[k] = XsbTriplet(i % UInt32, j % UInt32, v)
    x
end

signed(op, v) = op === (+) ? v : op === (-) ? -v :
    error("libxsparse_b200 implements op in {+,-}; use the CPU buffers for other ops")

# the calls the generic wrappers forward to (genericmt...:87-114, genericextendable...:44-92)
rawupdateindex!(x::SparseMatrixB200, op, v, i, j) = push_entry!(x, XSB_RAW, signed(op, v), i, j)
rawupdateindex!(x::SparseMatrixB200, op, v, i, j, tid) = rawupdateindex!(x, op, v, i, j)
updateindex!(x::SparseMatrixB200, op, v, i, j) = push_entry!(x, XSB_UPDATE, signed(op, v), i, j)
Base.setindex!(x::SparseMatrixB200, v, i::Integer, j::Integer) = push_entry!(x, XSB_ASSIGN, v, i, j)
function Base.getindex(x::SparseMatrixB200{Tv}, i::Integer, j::Integer) where {Tv}
    # entries still in the buffer are not visible before flush! (same contract as the MT wrapper,
    # genericmt...:71-82); the single-buffer wrapper reaches here only on a CSC miss.
    nnz(x) == 0 ? zero(Tv) : error("flush! before reading unflushed entries of a B200 matrix")
end

"""
    Base.sum(exts::Vector{SparseMatrixB200}, csc) -> SparseMatrixCSC

The flush of the plug-in contract (abstractsparsematrixextension.jl:13): partitions are summed
in vector order like sparsematrixdilnkc.jl:416-426.  The old CSC seeds the device matrix, all
buffers are replayed onto it, and the merged CSC is fetched into Julia-owned arrays.
"""
function Base.sum(exts::Vector{SparseMatrixB200{Tv, Ti}}, csc::SparseMatrixCSC{Tv, Ti};
                  mode = XSB_DETERMINISTIC) where {Tv, Ti}
    sum(nnz, exts) == 0 && return csc
    acc = exts[1]
    foreach(ship!, exts)
    check(acc.handle, ccall((:xsb_synchronize, libxsb), Int32, (Ptr{Cvoid},), acc.handle))
    if length(exts) > 1
        # one device matrix with one staging buffer per partition keeps the partition order
        big = B200Assembler{Tv, Ti}(csc.m, csc.n, length(exts))
        set_csc!(big, csc)
        for (t, e) in enumerate(exts)
            replay!(big, e, t - 1)
        end
        return fetch!(big, mode)
    end
    h = acc.handle
    # pending records were staged against an empty device CSC; merging with `csc` needs the old
    # entries on the device first.  xsb_set_csc would drop the staged records, so the old matrix
    # is inserted as the first partition of a fresh assembler instead.
    big = B200Assembler{Tv, Ti}(csc.m, csc.n, 1)
    set_csc!(big, csc)
    replay!(big, acc, 0)
    fetch!(big, mode)
end

Base.:+(x::SparseMatrixB200, csc::SparseMatrixCSC) = sum([x], csc)

# ---------------------------------------------------------------------------------------
# B200Assembler: the device matrix proper (resident CSC + per-partition staging buffers).
# Users who can afford to keep the matrix on the GPU between flushes should use this type
# directly; the AbstractSparseMatrixExtension path above round-trips the CSC every flush!.
# ---------------------------------------------------------------------------------------
mutable struct B200Assembler{Tv, Ti <: Integer}
    m::Int
    n::Int
    handle::Ptr{Cvoid}
    function B200Assembler{Tv, Ti}(m, n, nparts = 1; device = 0) where {Tv, Ti}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:xsb_create, libxsb), Int32,
                   (Int64, Int64, Int32, Int32, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
                   m, n, XSB_F64, idxcode(Ti), 1, nparts, device, h)
        check(Ptr{Cvoid}(C_NULL), rc)
        x = new{Tv, Ti}(m, n, h[])
        finalizer(y -> (y.handle == C_NULL || ccall((:xsb_destroy, libxsb), Int32, (Ptr{Cvoid},), y.handle);
                        y.handle = C_NULL), x)
        x
    end
end

function set_csc!(a::B200Assembler{Tv, Ti}, csc::SparseMatrixCSC{Tv, Ti}) where {Tv, Ti}
    check(a.handle, ccall((:xsb_set_csc, libxsb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                          a.handle, csc.colptr, csc.rowval, csc.nzval))
end

"bulk insertion: k-th element == k-th updateindex!/rawupdateindex!/setindex! call"
function insert!(a::B200Assembler{Tv, Ti}, I::Vector{Ti}, J::Vector{Ti}, V::Vector{Tv};
                 flavour = XSB_UPDATE, tid = 0) where {Tv, Ti}
    check(a.handle, ccall((:xsb_insert_batch, libxsb), Int32,
                          (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32),
                          a.handle, tid, I, J, V, length(V), flavour))
end

function replay!(a::B200Assembler{Tv, Ti}, e::SparseMatrixB200{Tv, Ti}, tid) where {Tv, Ti}
    cnt = Ref{Int64}(0)
    check(e.handle, ccall((:xsb_debug_fetch_staged, libxsb), Int32,
                          (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int64}),
                          e.handle, 0, C_NULL, C_NULL, C_NULL, C_NULL, 0, cnt))
    n = cnt[]
    I = Vector{Ti}(undef, n); J = Vector{Ti}(undef, n); V = Vector{Tv}(undef, n); F = Vector{Int32}(undef, n)
    check(e.handle, ccall((:xsb_debug_fetch_staged, libxsb), Int32,
                          (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int64}),
                          e.handle, 0, I, J, V, F, n, cnt))
    # runs of equal flavour keep the call order
    s = 1
    while s <= n
        t = s
        while t < n && F[t + 1] == F[s]
            t += 1
        end
        insert!(a, I[s:t], J[s:t], V[s:t]; flavour = F[s], tid = tid)
        s = t + 1
    end
end

"flush! + sparse(A): two-phase -- query nnz, allocate Julia arrays, fetch (ownership stays with Julia)"
function fetch!(a::B200Assembler{Tv, Ti}, mode = XSB_DETERMINISTIC) where {Tv, Ti}
    nnz = Ref{Int64}(0); changed = Ref{Int32}(0)
    check(a.handle, ccall((:xsb_flush, libxsb), Int32, (Ptr{Cvoid}, Int32, Ref{Int64}, Ref{Int32}),
                          a.handle, mode, nnz, changed))
    colptr = Vector{Ti}(undef, a.n + 1)
    rowval = Vector{Ti}(undef, nnz[])
    nzval = Vector{Tv}(undef, nnz[])
    check(a.handle, ccall((:xsb_fetch_csc, libxsb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                          a.handle, colptr, rowval, nzval))
    SparseMatrixCSC{Tv, Ti}(a.m, a.n, colptr, rowval, nzval)
end

"values-only Newton/transient loop: freeze the stream's positions once, then push values"
function freeze!(a::B200Assembler{Tv, Ti}, I::Vector{Ti}, J::Vector{Ti}) where {Tv, Ti}
    check(a.handle, ccall((:xsb_freeze_pattern, libxsb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                          a.handle, I, J, length(I)))
end
function reassemble!(a::B200Assembler{Tv}, V::Vector{Tv}; mode = XSB_DETERMINISTIC, zero = true) where {Tv}
    zero && check(a.handle, ccall((:xsb_zero_values, libxsb), Int32, (Ptr{Cvoid},), a.handle))
    check(a.handle, ccall((:xsb_reassemble_values, libxsb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32),
                          a.handle, V, length(V), mode))
end

"""
    pointblock(a::B200Assembler, blocksize) -> (colptr, rowval, blocks)

Device version of `pointblock` (src/matrix/extendable.jl:292-318): the block pattern is assembled on
the GPU from the resident CSC; `blocks[:, :, k]` is the k-th stored `blocksize x blocksize` block
(column-major, the memory layout of `SMatrix{bs,bs}`), so
`reinterpret(SMatrix{bs,bs,Tv,bs*bs}, vec(blocks))` is the `nzval` of the reference's result.
"""
function pointblock(a::B200Assembler{Tv, Ti}, blocksize::Integer) where {Tv, Ti}
    hb = Ref{Ptr{Cvoid}}(C_NULL)
    check(a.handle, ccall((:xsb_pointblock, libxsb), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}),
                          a.handle, blocksize, hb))
    try
        nnzb = Ref{Int64}(0)
        check(hb[], ccall((:xsb_nnz, libxsb), Int32, (Ptr{Cvoid}, Ref{Int64}), hb[], nnzb))
        nb = a.n ÷ blocksize
        colptr = Vector{Ti}(undef, nb + 1)
        rowval = Vector{Ti}(undef, nnzb[])
        blocks = Array{Tv, 3}(undef, blocksize, blocksize, nnzb[])
        check(hb[], ccall((:xsb_fetch_csc, libxsb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                          hb[], colptr, rowval, C_NULL))
        check(hb[], ccall((:xsb_fetch_blocks, libxsb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), hb[], blocks))
        return colptr, rowval, blocks
    finally
        ccall((:xsb_destroy, libxsb), Int32, (Ptr{Cvoid},), hb[])
    end
end

# the drop-in aliases, mirroring src/ExtendableSparse.jl:35-39
const B200ExtendableSparseMatrixCSC{Tv, Ti} =
    ExtendableSparse.GenericMTExtendableSparseMatrixCSC{SparseMatrixB200{Tv, Ti}, Tv, Ti}
const STB200ExtendableSparseMatrixCSC{Tv, Ti} =
    ExtendableSparse.GenericExtendableSparseMatrixCSC{SparseMatrixB200{Tv, Ti}, Tv, Ti}

export SparseMatrixB200, B200Assembler, B200ExtendableSparseMatrixCSC, STB200ExtendableSparseMatrixCSC,
       set_csc!, insert!, fetch!, freeze!, reassemble!, pointblock, XsbTriplet

end # module
