/*
 * xsb_streams.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Sequential generators of the insertion streams the benchmark configurations
 * are defined on.  Each emits the exact (i,j,v) call
 * sequence a user loop would issue, 1-based indices:
 *
 *   fdrand!   src/matrix/sprand.jl:58-126  (update_pair order :87-92)
 *   P1 FEM    test/femtools.jl:45-72       (20 rawupdateindex! calls per tetrahedron)
 *   block RD  SURVEY.md 8(d) cfg 4         (this build's own definition; the
 *                                           reference has no such test)
 *
 * Random numbers: the reference calls an unseeded rand() (sprand.jl:63); the
 * build fixes a counter-based Philox4x32-10 so that the CPU oracle and the GPU
 * emitters produce identical bits: the c-th rand() call of a stream (c = 0,1,..)
 * is  u_c = (philox(key=seed, ctr=(c_lo,c_hi,0,0)).w[1:0] >> 11) * 2^-53.
 *
 * Compile with -ffp-contract=off: the GPU emitters use non-fused IEEE double
 * operations in the same order, which makes the values bit-identical.
 */
#include <stdint.h>
#include <stdlib.h>

typedef int64_t i64;

/* ---------------- Philox4x32-10 (Salmon et al., SC'11) ---------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
}

void ora_philox4x32_10(uint64_t seed, uint64_t counter, uint32_t out[4])
{
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; r++)
    {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0];
    out[1] = c[1];
    out[2] = c[2];
    out[3] = c[3];
}

double ora_uniform(uint64_t seed, uint64_t counter)
{
    uint32_t w[4];
    ora_philox4x32_10(seed, counter, w);
    uint64_t bits = ((uint64_t)w[1] << 32) | (uint64_t)w[0];
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

/* ---------------- fdrand! stream (sprand.jl:58-126) ---------------- */
typedef struct
{
    i64 *I, *J;
    double *V;
    i64 pos;
    uint64_t seed, calls;
    int ones;
} emitter;

static inline double next_rand(emitter *e)
{
    double u = e->ones ? 1.0 : ora_uniform(e->seed, e->calls);
    e->calls++;
    return u;
}

static inline void emit(emitter *e, double v, i64 i, i64 j)
{
    if (e->I)
    {
        e->I[e->pos] = i;
        e->J[e->pos] = j;
        e->V[e->pos] = v;
    }
    e->pos++;
}

static inline void update_pair(emitter *e, double v, i64 i, i64 j)
{ /* sprand.jl:87-92 */
    emit(e, -v, i, j);
    emit(e, -v, j, i);
    emit(e, v, i, i);
    emit(e, v, j, j);
}

static i64 fdrand_run(emitter *e, i64 nx, i64 ny, i64 nz)
{
    const double hx = 1.0 / (double)nx, hy = 1.0 / (double)ny, hz = 1.0 / (double)nz;
    const i64 nxy = nx * ny;
    i64 l = 1;
    for (i64 k = 1; k <= nz; k++)
        for (i64 j = 1; j <= ny; j++)
            for (i64 i = 1; i <= nx; i++)
            {
                if (i < nx)
                    update_pair(e, next_rand(e) * hy * hz / hx, l, l + 1);
                if (i == 1 || i == nx)
                    emit(e, next_rand(e) * hy * hz, l, l);
                if (j < ny)
                    update_pair(e, next_rand(e) * hx * hz / hy, l, l + nx);
                if (ny > 2 && (j == 1 || j == ny))
                    emit(e, next_rand(e) * hx * hz, l, l);
                if (k < nz)
                    update_pair(e, next_rand(e) * hx * hy / hz, l, l + nxy);
                if (nz > 2 && (k == 1 || k == nz))
                    emit(e, next_rand(e) * hx * hy, l, l);
                l++;
            }
    return e->pos;
}

i64 ora_fdrand_count(i64 nx, i64 ny, i64 nz)
{
    emitter e = {0};
    e.ones = 1;
    return fdrand_run(&e, nx, ny, nz);
}

void ora_fdrand_stream(i64 nx, i64 ny, i64 nz, uint64_t seed, int ones, i64 *I, i64 *J, double *V)
{
    emitter e = {0};
    e.I = I;
    e.J = J;
    e.V = V;
    e.seed = seed;
    e.ones = ones;
    fdrand_run(&e, nx, ny, nz);
}

/* ---------------- P1 FEM stream (femtools.jl:45-72) ----------------
 * Mesh (this build's definition, SURVEY.md 8d cfg 2): tensor grid of
 * nxn*nyn*nzn nodes on [0,1]^3, node id = 1 + ix + nxn*iy + nxn*nyn*iz,
 * coordinate = index/(count-1).  Each cube is cut into the 6 Kuhn tetrahedra:
 * for the axis permutation (a,b,c) the vertices are base, base+e_a,
 * base+e_a+e_b, base+e_a+e_b+e_c.  Cells are numbered cube-major
 * (x fastest), permutation index minor, in the order of KUHN below.
 *
 * Element math: P1 gradients from the inverse of the edge matrix (own
 * restatement of coordmatrix!/gradient!/stiffness!, femtools.jl:9-43; the
 * reference goes through a pivoted LU of the 4x4 coordinate matrix, which
 * agrees to rounding).  vol = |det|/6.
 */
static const int KUHN[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};

i64 ora_fem_count(i64 nxn, i64 nyn, i64 nzn) { return 20 * 6 * (nxn - 1) * (nyn - 1) * (nzn - 1); }

static void tet_matrix(const double p[4][3], double *vol_out, double S[4][4])
{
    double a[3][3]; /* a[r][c] = p[c+1][r] - p[0][r] : columns are edge vectors */
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            a[r][c] = p[c + 1][r] - p[0][r];
    /* cofactors */
    const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    const double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    const double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    const double c10 = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    const double c11 = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    const double c12 = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    const double c20 = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    const double c21 = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    const double c22 = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    const double det = (a[0][0] * c00 + a[0][1] * c01) + a[0][2] * c02;
    /* inverse: inv[r][c] = cof[c][r]/det ; gradient of lambda_{r+1} = row r of inv */
    double g[4][3];
    g[1][0] = c00 / det;
    g[1][1] = c10 / det;
    g[1][2] = c20 / det;
    g[2][0] = c01 / det;
    g[2][1] = c11 / det;
    g[2][2] = c21 / det;
    g[3][0] = c02 / det;
    g[3][1] = c12 / det;
    g[3][2] = c22 / det;
    for (int d = 0; d < 3; d++)
        g[0][d] = -((g[1][d] + g[2][d]) + g[3][d]);
    const double adet = det < 0 ? -det : det;
    *vol_out = adet / 6.0;
    for (int il = 0; il < 4; il++)
        for (int jl = il; jl < 4; jl++)
        {
            double s = 0.0;
            for (int d = 0; d < 3; d++)
                s += g[jl][d] * g[il][d];
            S[il][jl] = s;
            S[jl][il] = s;
        }
}

void ora_fem_stream(i64 nxn, i64 nyn, i64 nzn, i64 *I, i64 *J, double *V)
{
    i64 pos = 0;
    const double dx = (double)(nxn - 1), dy = (double)(nyn - 1), dz = (double)(nzn - 1);
    for (i64 cz = 0; cz < nzn - 1; cz++)
        for (i64 cy = 0; cy < nyn - 1; cy++)
            for (i64 cx = 0; cx < nxn - 1; cx++)
                for (int t = 0; t < 6; t++)
                {
                    i64 idx[4][3];
                    idx[0][0] = cx;
                    idx[0][1] = cy;
                    idx[0][2] = cz;
                    for (int v = 1; v < 4; v++)
                    {
                        for (int d = 0; d < 3; d++)
                            idx[v][d] = idx[v - 1][d];
                        idx[v][KUHN[t][v - 1]] += 1;
                    }
                    double p[4][3];
                    i64 node[4];
                    for (int v = 0; v < 4; v++)
                    {
                        p[v][0] = (double)idx[v][0] / dx;
                        p[v][1] = (double)idx[v][1] / dy;
                        p[v][2] = (double)idx[v][2] / dz;
                        node[v] = 1 + idx[v][0] + nxn * idx[v][1] + nxn * nyn * idx[v][2];
                    }
                    double vol, S[4][4];
                    tet_matrix(p, &vol, S);
                    for (int il = 0; il < 4; il++)
                    { /* femtools.jl:62-69 */
                        I[pos] = node[il];
                        J[pos] = node[il];
                        V[pos] = 0.1 * vol / 4.0;
                        pos++;
                        for (int jl = 0; jl < 4; jl++)
                        {
                            I[pos] = node[il];
                            J[pos] = node[jl];
                            V[pos] = vol * S[il][jl];
                            pos++;
                        }
                    }
                }
}

/* ---------------- block reaction-diffusion stream (cfg 4) ----------------
 * nx*ny*nz grid nodes, ns species, unknown id = ns*(node-1)+s (s = 1..ns).
 * Per node l in lexicographic order (x fastest): for each of the x-, y-, z-edges
 * leaving l a random dense ns*ns block B (ns*ns rand() calls, row-major) applied
 * as update_pair per block entry (-B at (i_a,j_b) and (j_a,i_b), +B at (i_a,i_b)
 * and (j_a,j_b)); then an ns*ns reaction block R at (l_a,l_b). */
i64 ora_blockrd_count(i64 nx, i64 ny, i64 nz, i64 ns)
{
    i64 edges = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1);
    return edges * 4 * ns * ns + nx * ny * nz * ns * ns;
}

void ora_blockrd_stream(i64 nx, i64 ny, i64 nz, i64 ns, uint64_t seed, i64 *I, i64 *J, double *V)
{
    emitter e = {0};
    e.I = I;
    e.J = J;
    e.V = V;
    e.seed = seed;
    const i64 nxy = nx * ny;
    const i64 step[3] = {1, nx, nxy};
    i64 l = 1;
    for (i64 k = 1; k <= nz; k++)
        for (i64 j = 1; j <= ny; j++)
            for (i64 i = 1; i <= nx; i++)
            {
                const int has[3] = {i < nx, j < ny, k < nz};
                for (int d = 0; d < 3; d++)
                {
                    if (!has[d])
                        continue;
                    const i64 l2 = l + step[d];
                    for (i64 a = 1; a <= ns; a++)
                        for (i64 b = 1; b <= ns; b++)
                        {
                            const double v = next_rand(&e);
                            const i64 ia = ns * (l - 1) + a, ib = ns * (l - 1) + b;
                            const i64 ja = ns * (l2 - 1) + a, jb = ns * (l2 - 1) + b;
                            emit(&e, -v, ia, jb);
                            emit(&e, -v, ja, ib);
                            emit(&e, v, ia, ib);
                            emit(&e, v, ja, jb);
                        }
                }
                for (i64 a = 1; a <= ns; a++)
                    for (i64 b = 1; b <= ns; b++)
                        emit(&e, next_rand(&e), ns * (l - 1) + a, ns * (l - 1) + b);
                l++;
            }
}
