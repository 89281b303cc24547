# tools/julia_ref.jl -- runs the REAL ExtendableSparse.jl on the streams the oracle is checked with.
# Written against the reference sources; no Julia runtime exists in the build container, so it has not been executed.
#
#   julia tools/julia_ref.jl dump [outdir]      CSC of every golden stream (tests/golden/<case>.csv holds I,J,V,flavour)
#   julia tools/julia_ref.jl time <mesh> <reps> seconds per assembly of the Kuhn-mesh FEM stream (rawupdateindex! + flush!)
using ExtendableSparse, SparseArrays, Printf

"The Kuhn 6-tet tensor mesh stream of this repository (oracle/xsb_streams.c: ora_fem_stream), test/femtools.jl:45-72 per cell."
function fem_assemble!(A, n)
    kuhn = ((1, 2, 3), (1, 3, 2), (2, 1, 3), (2, 3, 1), (3, 1, 2), (3, 2, 1))
    G = zeros(4, 4); C = zeros(4, 4)
    for cz in 0:n-2, cy in 0:n-2, cx in 0:n-2, perm in kuhn
        idx = zeros(Int, 4, 3); idx[1, :] .= (cx, cy, cz)
        for v in 2:4
            idx[v, :] .= idx[v-1, :]; idx[v, perm[v-1]] += 1
        end
        nodes = [idx[v, 1] + n * idx[v, 2] + n * n * idx[v, 3] + 1 for v in 1:4]
        p = [idx[v, d] / (n - 1) for v in 1:4, d in 1:3]
        a = [p[c+1, r] - p[1, r] for r in 1:3, c in 1:3]
        det3 = (a[1, 1] * (a[2, 2] * a[3, 3] - a[2, 3] * a[3, 2]) + a[1, 2] * (a[2, 3] * a[3, 1] - a[2, 1] * a[3, 3])) +
               a[1, 3] * (a[2, 1] * a[3, 2] - a[2, 2] * a[3, 1])
        inva = inv(a)   # NB: the oracle divides cofactors by det in a fixed order; compare with isapprox unless restated
        gr = zeros(4, 3); gr[2:4, :] .= inva
        gr[1, :] .= -(gr[2, :] .+ gr[3, :] .+ gr[4, :])
        vol = abs(det3) / 6
        for il in 1:4
            i = nodes[il]
            rawupdateindex!(A, +, 0.1 * vol / 4, i, i)
            for jl in 1:4
                rawupdateindex!(A, +, vol * sum(gr[jl, d] * gr[il, d] for d in 1:3), i, nodes[jl])
            end
        end
    end
    flush!(A)
end

function dump(outdir)
    for f in filter(x -> endswith(x, ".csv"), readdir(outdir; join = true))
        rows = [split(l, ',') for l in eachline(f)][2:end]
        m, n = parse.(Int, split(readline(replace(f, ".csv" => ".dims")), ','))
        A = ExtendableSparseMatrix(Float64, Int64, m, n)
        for r in rows
            i, j, v, fl = parse(Int, r[1]), parse(Int, r[2]), parse(Float64, r[3]), parse(Int, r[4])
            fl == 0 ? updateindex!(A, +, v, i, j) : fl == 1 ? rawupdateindex!(A, +, v, i, j) : (A[i, j] = v)
        end
        flush!(A)
        S = sparse(A)
        open(replace(f, ".csv" => ".julia_csc"), "w") do io
            println(io, join(S.colptr, ','))
            println(io, join(S.rowval, ','))
            println(io, join((@sprintf("%a", x) for x in S.nzval), ','))   # hex floats: bit-exact
        end
    end
end

if ARGS[1] == "dump"
    dump(length(ARGS) > 1 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden"))
elseif ARGS[1] == "time"
    n, reps = parse(Int, ARGS[2]), parse(Int, ARGS[3])
    best = Inf
    for _ in 1:reps+1
        A = ExtendableSparseMatrix(Float64, Int64, n^3, n^3)
        best = min(best, @elapsed fem_assemble!(A, n))
    end
    println(best)
end
