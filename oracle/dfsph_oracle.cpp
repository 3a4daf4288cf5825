// TEST INFRASTRUCTURE ONLY (oracle/).  Plain C++ CPU restatement of the reference's hot path: neighbourhood search +
// one DFSPH time step, in the reference's two variants (float = AVX-variant semantics, double = scalar-variant
// semantics).  No reference code is included or linked; every function cites the reference lines it
// restates (paths relative to the reference root).  Exposes the SAME C harness ABI as oracle/ref_driver.cpp (names
// ref_*), so oracle/refsim.RefSim drives either library.
//
// PINNING: the reference ships no golden vectors for this path.  This restatement is pinned against
// the reference ITSELF: oracle/_ref (the reference's unmodified DFSPH sources compiled here) -- see
// tests/test_oracle.py (field agreement per step, identical neighbour sets, identical iteration counts) and the
// committed fixtures under tests/golden/ generated from oracle/_ref by tests/golden/make_golden.py.
// Neighbour-set parity against upstream CompactNSearch @ a9ab7c71 (absent from the reference tree) is UNPINNED; the
// predicate is restated from its published algorithm (see oracle/standin/CompactNSearch.h).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <omp.h>

#ifdef ORACLE_DOUBLE
typedef double Real;
#define ORACLE_AVX_VARIANT 0
#else
typedef float Real;
#define ORACLE_AVX_VARIANT 1
#endif

namespace {

struct V3 { Real x, y, z; };
static inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static inline Real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

const unsigned LUT_RES = 10000u;
const Real EPS = static_cast<Real>(1.0e-5);   // TimeStepDFSPH.h:28

struct Sim {
    // Simulation / TimeManager parameters (Simulation.cpp:67-88, TimeManager.cpp:12)
    Real radius = 0.025f, support = 0.1f;
    Real h = static_cast<Real>(0.001);
    double time = 0.0;
    Real gravity[3] = { 0, static_cast<Real>(-9.81), 0 };
    int cflMethod = 1;
    Real cflFactor = 0.5f, cflMin = static_cast<Real>(0.0001), cflMax = static_cast<Real>(0.005);
    int kernel = 4, gradKernel = 4;   // Simulation "kernel" / "gradKernel" (Simulation.cpp:306-393)
    // TimeStepDFSPH parameters (TimeStepDFSPH.cpp:28-41)
    unsigned minIter = 2, maxIter = 100, maxIterV = 100;
    Real maxError = static_cast<Real>(0.01), maxErrorV = static_cast<Real>(0.1);
    bool enableDiv = true;
    int viscosityMethod = 0;       // 0 none, 1 Viscosity_Standard
    Real viscosity = static_cast<Real>(0.01), viscosityBoundary = 0;
    std::vector<V3> acc;           // non-pressure accelerations (FluidModel::m_a)
    unsigned iterations = 0, iterationsV = 0;
    // kernel constants
    Real k = 0, l = 0, W_zero = 0, simW_zero = 0, invR = 0, invR2 = 0, lutInvStep = 0;
    Real wend_k = 0, wend_l = 0, poly_k = 0, poly_l = 0, spiky_k = 0, spiky_l = 0;
    std::vector<Real> lutW, lutG;
    // fluid
    Real density0 = 1000, V = 0;
    std::vector<V3> x, v, pa;
    std::vector<Real> density, factor, density_adv, kappa, kappa_v;
    std::vector<unsigned> id;
    // boundary
    std::vector<V3> bx;
    std::vector<Real> bV;
    // neighbour lists (CSR)
    std::vector<unsigned long long> off_f, off_b;
    std::vector<unsigned> nbr_f, nbr_b;
    bool created = false;
    double step_seconds = 0.0;
};
Sim* g = nullptr;

// ---- kernels ---------------------------------------------------------------------------------------------------------
// CubicKernel::W (SPHKernels.h:37-56)
Real cubicW(Real r)
{
    Real res = 0.0;
    const Real q = r / g->support;
    if (q <= 1.0) {
        if (q <= 0.5) { const Real q2 = q * q, q3 = q2 * q; res = g->k * (static_cast<Real>(6.0) * q3 - static_cast<Real>(6.0) * q2 + static_cast<Real>(1.0)); }
        else res = g->k * (static_cast<Real>(2.0) * std::pow(static_cast<Real>(1.0) - q, static_cast<Real>(3.0)));
    }
    return res;
}
// CubicKernel::gradW (SPHKernels.h:63-85)
V3 cubicGradW(V3 r)
{
    V3 res = { 0, 0, 0 };
    const Real rl = std::sqrt(dot(r, r));
    const Real q = rl / g->support;
    if ((rl > 1.0e-9) && (q <= 1.0)) {
        V3 gradq = { r.x / rl, r.y / rl, r.z / rl };
        gradq = { gradq.x / g->support, gradq.y / g->support, gradq.z / g->support };
        Real s;
        if (q <= 0.5) s = g->l * q * ((Real)3.0 * q - static_cast<Real>(2.0));
        else { const Real f = static_cast<Real>(1.0) - q; s = g->l * (-f * f); }
        res = { s * gradq.x, s * gradq.y, s * gradq.z };
    }
    return res;
}
// PrecomputedKernel (SPHKernels.h:649-660, 673-687)
Real lutWf(V3 r)
{
    Real res = 0.0;
    const Real r2 = dot(r, r);
    if (r2 <= g->support * g->support) {
        const Real rl = std::sqrt(r2);
        const unsigned pos = std::min<unsigned>((unsigned)(rl * g->lutInvStep), LUT_RES - 2u);
        res = static_cast<Real>(0.5) * (g->lutW[pos] + g->lutW[pos + 1]);
    }
    return res;
}
V3 lutGradW(V3 r)
{
    V3 res = { 0, 0, 0 };
    const Real rl = std::sqrt(dot(r, r));
    if (rl <= g->support) {
        const unsigned pos = std::min<unsigned>(static_cast<unsigned>(rl * g->lutInvStep), LUT_RES - 2u);
        const Real s = static_cast<Real>(0.5) * (g->lutG[pos] + g->lutG[pos + 1]);
        res = { s * r.x, s * r.y, s * r.z };
    }
    return res;
}
// CubicKernel_AVX (SPHKernels.h:743-786), one lane
Real avxW(V3 r)
{
    const Real rl = std::sqrt(dot(r, r));
    const Real q = rl * g->invR, v = static_cast<Real>(1.0) - q;
    const Real res1 = g->k * (static_cast<Real>(-6.0) * q * q * v + static_cast<Real>(1.0));
    const Real res2 = g->k * static_cast<Real>(2.0) * (v * v * v);
    Real res = q <= 1.0 ? res2 : static_cast<Real>(0.0);
    return q <= 0.5 ? res1 : res;
}
V3 avxGradW(V3 r)
{
    const Real rl = std::sqrt(dot(r, r));
    const Real q = rl * g->invR;
    const Real res1 = g->l * g->invR2 * (static_cast<Real>(3.0) * q - static_cast<Real>(2.0));
    const Real v = static_cast<Real>(1.0) - q;
    const Real gradq = g->invR / rl;
    const Real res2 = gradq * (-g->l * (v * v));
    Real res = q <= 1.0 ? res2 : static_cast<Real>(0.0);
    res = q <= 0.5 ? res1 : res;
    res = rl > static_cast<Real>(1.0e-9) ? res : static_cast<Real>(0.0);
    return { r.x * res, r.y * res, r.z * res };
}
// WendlandQuinticC2Kernel (SPHKernels.h:291-320)
Real wendlandW(Real r)
{
    Real res = 0.0;
    const Real q = r / g->support;
    if (q <= 1.0) res = g->wend_k * std::pow(static_cast<Real>(1.0) - q, static_cast<Real>(4.0)) * (static_cast<Real>(4.0) * q + static_cast<Real>(1.0));
    return res;
}
V3 wendlandGradW(V3 r)
{
    V3 res = { 0, 0, 0 };
    const Real rl = std::sqrt(dot(r, r));
    const Real q = rl / g->support;
    if (q <= 1.0) {
        const Real f = static_cast<Real>(1.0) / (rl * g->support);
        const Real s = g->wend_l * q * std::pow(static_cast<Real>(1.0) - q, static_cast<Real>(3.0));
        res = { s * (r.x * f), s * (r.y * f), s * (r.z * f) };
    }
    return res;
}
// Poly6Kernel (SPHKernels.h:138-167)
Real poly6W(V3 r)
{
    Real res = 0.0;
    const Real r2 = dot(r, r), radius2 = g->support * g->support;
    if (r2 <= radius2) res = std::pow(radius2 - r2, static_cast<Real>(3.0)) * g->poly_k;
    return res;
}
V3 poly6GradW(V3 r)
{
    V3 res = { 0, 0, 0 };
    const Real r2 = dot(r, r), radius2 = g->support * g->support;
    if (r2 <= radius2) { const Real tmp = radius2 - r2; const Real s = g->poly_l * tmp * tmp; res = { s * r.x, s * r.y, s * r.z }; }
    return res;
}
// SpikyKernel (SPHKernels.h:225-257)
Real spikyW(V3 r)
{
    Real res = 0.0;
    const Real r2 = dot(r, r), radius2 = g->support * g->support;
    if (r2 <= radius2) res = g->spiky_k * std::pow(g->support - std::sqrt(r2), static_cast<Real>(3.0));
    return res;
}
V3 spikyGradW(V3 r)
{
    V3 res = { 0, 0, 0 };
    const Real r2 = dot(r, r), radius2 = g->support * g->support;
    if (r2 <= radius2) {
        const Real r_l = std::sqrt(r2), hr = g->support - r_l;
        const Real s = g->spiky_l * (hr * hr), f = static_cast<Real>(1.0) / r_l;
        res = { s * r.x * f, s * r.y * f, s * r.z * f };
    }
    return res;
}
// sim->W / sim->gradW: the configured scalar kernels (Simulation.h:381-382; setKernel / setGradKernel, Simulation.cpp:306-393)
Real simW(V3 r)
{
    switch (g->kernel) {
        case 1: return wendlandW(std::sqrt(dot(r, r)));
        case 2: return poly6W(r);
        case 3: return spikyW(r);
        case 4: return lutWf(r);
        default: return cubicW(std::sqrt(dot(r, r)));
    }
}
V3 simGradW(V3 r)
{
    switch (g->gradKernel) {
        case 1: return wendlandGradW(r);
        case 2: return poly6GradW(r);
        case 3: return spikyGradW(r);
        case 4: return lutGradW(r);
        default: return cubicGradW(r);
    }
}
// kernel used inside the solver sums
Real solverW(V3 r) { return ORACLE_AVX_VARIANT ? avxW(r) : simW(r); }
V3 solverGradW(V3 r) { return ORACLE_AVX_VARIANT ? avxGradW(r) : simGradW(r); }

// Simulation::setParticleRadius + initKernels (Simulation.cpp:280-330), setRadius of the kernels
void initKernels()
{
    g->support = static_cast<Real>(4.0) * g->radius;
    const Real pi = static_cast<Real>(M_PI);
    const Real h3 = g->support * g->support * g->support;
#if ORACLE_AVX_VARIANT
    g->invR = 1.0f / g->support;
    g->k = 8.0f / static_cast<float>(pi * h3);
    g->l = 48.0f / static_cast<float>(pi * h3);
#else
    g->invR = 1.0 / g->support;
    g->k = static_cast<Real>(8.0) / (pi * h3);
    g->l = static_cast<Real>(48.0) / (pi * h3);
#endif
    g->invR2 = g->invR * g->invR;
    {   // setRadius of the other kernels (SPHKernels.h:107-118, 196-204, 268-277)
        const Real h6 = std::pow(g->support, static_cast<Real>(6.0)), h9 = std::pow(g->support, static_cast<Real>(9.0));
        g->wend_k = static_cast<Real>(21.0) / (static_cast<Real>(2.0) * pi * h3);
        g->wend_l = -static_cast<Real>(210.0) / (pi * h3);
        g->poly_k = static_cast<Real>(315.0) / (static_cast<Real>(64.0) * pi * h9);
        g->poly_l = -static_cast<Real>(945.0) / (static_cast<Real>(32.0) * pi * h9);
        g->spiky_k = static_cast<Real>(15.0) / (pi * h6);
        g->spiky_l = -static_cast<Real>(45.0) / (pi * h6);
    }
    // sim->W_zero() of the configured kernel; the AVX solver sums use CubicKernel_AVX::W_zero() regardless (TimeStep.cpp:70)
    g->simW_zero = g->kernel == 1 ? wendlandW(0) : g->kernel == 2 ? poly6W({ 0, 0, 0 }) : g->kernel == 3 ? spikyW({ 0, 0, 0 }) : cubicW(0);
    g->W_zero = ORACLE_AVX_VARIANT ? cubicW(0) : g->simW_zero;
    g->lutW.resize(LUT_RES);
    g->lutG.resize(LUT_RES + 1);
    const Real stepSize = g->support / (Real)(LUT_RES - 1);
    g->lutInvStep = static_cast<Real>(1.0) / stepSize;
    for (unsigned i = 0; i < LUT_RES; i++) {
        const Real posX = stepSize * (Real)i;
        g->lutW[i] = cubicW(posX);
        if (posX > 1.0e-9) g->lutG[i] = cubicGradW({ posX, 0, 0 }).x / posX;
        else g->lutG[i] = 0.0;
    }
    g->lutG[LUT_RES] = 0.0;
}

// ---- neighbourhood search (CompactNSearch contract, see oracle/standin/CompactNSearch.h) ------------------------------
static inline int cellOf(Real x, Real inv) { return x >= 0.0 ? static_cast<int>(inv * x) : static_cast<int>(inv * x) - 1; }

struct Grid {
    int lo[3], n[3];
    std::vector<unsigned> start, sorted;
    void build(const std::vector<V3>& p, const int glo[3], const int gn[3], Real inv)
    {
        for (int k = 0; k < 3; ++k) { lo[k] = glo[k]; n[k] = gn[k]; }
        const size_t nc = (size_t)n[0] * n[1] * n[2];
        start.assign(nc + 1, 0u);
        std::vector<unsigned> cid(p.size());
        for (size_t i = 0; i < p.size(); ++i) {
            const int c[3] = { cellOf(p[i].x, inv) - lo[0], cellOf(p[i].y, inv) - lo[1], cellOf(p[i].z, inv) - lo[2] };
            cid[i] = (unsigned)(((size_t)c[0] * n[1] + c[1]) * n[2] + c[2]);
            start[cid[i] + 1]++;
        }
        for (size_t c = 0; c < nc; ++c) start[c + 1] += start[c];
        sorted.resize(p.size());
        std::vector<unsigned> cur(start.begin(), start.end() - 1);
        for (size_t i = 0; i < p.size(); ++i) sorted[cur[cid[i]]++] = (unsigned)i;
    }
};

// neighbours of every point of `a` inside set `b` (self: a and b are the same set)
void query(const std::vector<V3>& a, const std::vector<V3>& b, const Grid& gb, bool self, Real inv, Real r2,
           std::vector<unsigned long long>& off, std::vector<unsigned>& out)
{
    const long n = (long)a.size();
    std::vector<std::vector<unsigned>> lists(n);
    #pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        const V3 xa = a[i];
        const int c[3] = { cellOf(xa.x, inv) - gb.lo[0], cellOf(xa.y, inv) - gb.lo[1], cellOf(xa.z, inv) - gb.lo[2] };
        for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
            const int cx = c[0] + dx, cy = c[1] + dy, cz = c[2] + dz;
            if (cx < 0 || cy < 0 || cz < 0 || cx >= gb.n[0] || cy >= gb.n[1] || cz >= gb.n[2]) continue;
            const size_t cell = ((size_t)cx * gb.n[1] + cy) * gb.n[2] + cz;
            for (unsigned k = gb.start[cell]; k < gb.start[cell + 1]; ++k) {
                const unsigned j = gb.sorted[k];
                if (self && j == (unsigned)i) continue;
                const V3 xb = b[j];
                Real tmp = xa.x - xb.x;
                Real l2 = tmp * tmp;
                tmp = xa.y - xb.y;
                l2 += tmp * tmp;
                tmp = xa.z - xb.z;
                l2 += tmp * tmp;
                if (l2 < r2) lists[i].push_back(j);
            }
        }
    }
    off.assign(n + 1, 0ull);
    for (long i = 0; i < n; ++i) off[i + 1] = off[i] + lists[i].size();
    out.resize(off[n]);
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) std::copy(lists[i].begin(), lists[i].end(), out.begin() + off[i]);
}

void gridExtent(const std::vector<V3>& a, const std::vector<V3>& b, Real inv, int lo[3], int n[3])
{
    int hi[3] = { -2147483647, -2147483647, -2147483647 };
    lo[0] = lo[1] = lo[2] = 2147483647;
    for (const std::vector<V3>* s : { &a, &b })
        for (const V3& p : *s) {
            const int c[3] = { cellOf(p.x, inv), cellOf(p.y, inv), cellOf(p.z, inv) };
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], c[k]); hi[k] = std::max(hi[k], c[k]); }
        }
    for (int k = 0; k < 3; ++k) n[k] = hi[k] >= lo[k] ? hi[k] - lo[k] + 1 : 1;
}

// Simulation::performNeighborhoodSearch -> find_neighbors (Simulation.cpp:606-619); fluid searches fluid + boundary
void findNeighbors()
{
    const Real inv = static_cast<Real>(1.0 / g->support);
    const Real r2 = g->support * g->support;
    int lo[3], n[3];
    gridExtent(g->x, g->bx, inv, lo, n);
    Grid gf, gb;
    gf.build(g->x, lo, n, inv);
    gb.build(g->bx, lo, n, inv);
    query(g->x, g->x, gf, true, inv, r2, g->off_f, g->nbr_f);
    query(g->x, g->bx, gb, false, inv, r2, g->off_b, g->nbr_b);
}

// Simulation::updateBoundaryVolume + BoundaryModel_Akinci2012::computeBoundaryVolume
// (Simulation.cpp:696-756, BoundaryModel_Akinci2012.cpp:48-75)
void computeBoundaryVolume()
{
    const Real inv = static_cast<Real>(1.0 / g->support);
    const Real r2 = g->support * g->support;
    int lo[3], n[3];
    std::vector<V3> none;
    gridExtent(g->bx, none, inv, lo, n);
    Grid gb;
    gb.build(g->bx, lo, n, inv);
    std::vector<unsigned long long> off;
    std::vector<unsigned> nb;
    query(g->bx, g->bx, gb, true, inv, r2, off, nb);
    g->bV.resize(g->bx.size());
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)g->bx.size(); ++i) {
        Real delta = g->simW_zero;
        for (unsigned long long k = off[i]; k < off[i + 1]; ++k) delta += simW(g->bx[i] - g->bx[nb[k]]);
        g->bV[i] = static_cast<Real>(1.0) / delta;
    }
}

// ---- solver pieces -----------------------------------------------------------------------------------------------------
#define FOR_FLUID(i, j) for (unsigned long long _k = g->off_f[i]; _k < g->off_f[i + 1]; ++_k) { const unsigned j = g->nbr_f[_k];
#define FOR_BOUNDARY(i, j) for (unsigned long long _k = g->off_b[i]; _k < g->off_b[i + 1]; ++_k) { const unsigned j = g->nbr_b[_k];
#define END_FOR }

// TimeStep::computeDensities (TimeStep.cpp:54-112 AVX / 116-169 scalar)
void computeDensities()
{
    const long n = (long)g->x.size();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        Real density = g->V * g->W_zero;
        Real sum = 0;
        FOR_FLUID(i, j) sum += g->V * solverW(g->x[i] - g->x[j]); END_FOR
        FOR_BOUNDARY(i, j) sum += g->bV[j] * solverW(g->x[i] - g->bx[j]); END_FOR
        density += sum;
        g->density[i] = density * g->density0;
    }
}

// TimeStepDFSPH::computeDFSPHFactor (TimeStepDFSPH.cpp:735-825 / 1106-1186)
void computeFactor()
{
    const long n = (long)g->x.size();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        Real sum = 0;
        V3 gi = { 0, 0, 0 };
        FOR_FLUID(i, j)
            const V3 w = solverGradW(g->x[i] - g->x[j]);
            const V3 p = { g->V * w.x, g->V * w.y, g->V * w.z };
            sum += dot(p, p);
            gi = { gi.x + p.x, gi.y + p.y, gi.z + p.z };
        END_FOR
        FOR_BOUNDARY(i, j)
            const V3 w = solverGradW(g->x[i] - g->bx[j]);
            gi = { gi.x + g->bV[j] * w.x, gi.y + g->bV[j] * w.y, gi.z + g->bV[j] * w.z };
        END_FOR
        sum += dot(gi, gi);
        g->factor[i] = sum > EPS ? static_cast<Real>(1.0) / sum : static_cast<Real>(0.0);
    }
}

// shared velocity-divergence sum of computeDensityChange (:894-950 / 1247-1295) and computeDensityAdv (:830-889 / 1191-1242)
Real velocityDivergence(long i)
{
    Real d = 0;
    const V3 vi = g->v[i];
    FOR_FLUID(i, j)
#if ORACLE_AVX_VARIANT
        const V3 w = solverGradW(g->x[i] - g->x[j]);
        d += dot(vi - g->v[j], { w.x * g->V, w.y * g->V, w.z * g->V });
#else
        d += dot(vi - g->v[j], solverGradW(g->x[i] - g->x[j]));
#endif
    END_FOR
#if !ORACLE_AVX_VARIANT
    d *= g->V;
#endif
    FOR_BOUNDARY(i, j) d += g->bV[j] * dot(vi, solverGradW(g->x[i] - g->bx[j])); END_FOR   // static boundary: v_b = 0
    return d;
}

// TimeStepDFSPH::computePressureAccel (:954-1039 / 1299-1367); all particles are Active here
void computePressureAccel(const std::vector<Real>& p)
{
    const long n = (long)g->x.size();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        V3 a = { 0, 0, 0 };
        const Real pi = p[i];
        FOR_FLUID(i, j)
            const Real pSum = pi + p[j];
#if ORACLE_AVX_VARIANT
            const V3 w = solverGradW(g->x[i] - g->x[j]);
            a = { a.x - w.x * g->V * pSum, a.y - w.y * g->V * pSum, a.z - w.z * g->V * pSum };
#else
            if (std::fabs(pSum) > EPS) {
                const V3 w = solverGradW(g->x[i] - g->x[j]);
                a = { a.x + pSum * (-g->V * w.x), a.y + pSum * (-g->V * w.y), a.z + pSum * (-g->V * w.z) };
            }
#endif
        END_FOR
        if (std::fabs(pi) > EPS) {
            FOR_BOUNDARY(i, j)
                const V3 w = solverGradW(g->x[i] - g->bx[j]);
                a = { a.x + pi * (-g->bV[j] * w.x), a.y + pi * (-g->bV[j] * w.y), a.z + pi * (-g->bV[j] * w.z) };
            END_FOR
        }
        g->pa[i] = a;
    }
}

// TimeStepDFSPH::compute_aij_pj (:1042-1100 / 1370-1420)
Real aij_pj(long i)
{
    Real s = 0;
    const V3 ai = g->pa[i];
    FOR_FLUID(i, j)
#if ORACLE_AVX_VARIANT
        const V3 w = solverGradW(g->x[i] - g->x[j]);
        s += dot(ai - g->pa[j], { w.x * g->V, w.y * g->V, w.z * g->V });
#else
        s += dot(ai - g->pa[j], solverGradW(g->x[i] - g->x[j]));
#endif
    END_FOR
#if !ORACLE_AVX_VARIANT
    s *= g->V;
#endif
    FOR_BOUNDARY(i, j) s += g->bV[j] * dot(ai, solverGradW(g->x[i] - g->bx[j])); END_FOR
    return s;
}

unsigned numNeighbors(long i) { return (unsigned)(g->off_f[i + 1] - g->off_f[i] + g->off_b[i + 1] - g->off_b[i]); }

// TimeStepDFSPH::divergenceSolve (:386-541) + divergenceSolveIteration (:621-706)
void divergenceSolve()
{
    const Real h = g->h, invH = static_cast<Real>(1.0) / h;
    const long n = (long)g->x.size();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        g->density_adv[i] = velocityDivergence(i);
        Real d = std::max(g->density_adv[i], static_cast<Real>(0.0));
        if (numNeighbors(i) < 20) d = 0.0;
        g->factor[i] *= invH;
        if (d > 0.0) g->kappa_v[i] = static_cast<Real>(0.5) * std::min(g->kappa_v[i], static_cast<Real>(0.5)) * invH;
        else g->kappa_v[i] = 0.0;
    }
    g->iterationsV = 0;
    Real avg = 0.0;
    bool chk = false;
    while ((!chk || (g->iterationsV < 1)) && (g->iterationsV < g->maxIterV)) {
        chk = true;
        avg = 0.0;
        if (n > 0) {
            computePressureAccel(g->kappa_v);
            Real err = 0.0;
            #pragma omp parallel for reduction(+:err) schedule(static)
            for (long i = 0; i < n; ++i) {
                const Real ap = aij_pj(i) * h;
                const Real s_i = -g->density_adv[i];
                Real residuum = std::min(s_i - ap, static_cast<Real>(0.0));
                if (numNeighbors(i) < 20) residuum = 0.0;
                g->kappa_v[i] = std::max(g->kappa_v[i] - static_cast<Real>(0.5) * (s_i - ap) * g->factor[i], static_cast<Real>(0.0));
                err -= g->density0 * residuum;
            }
            avg = err / n;
        }
        const Real eta = (static_cast<Real>(1.0) / h) * g->maxErrorV * static_cast<Real>(0.01) * g->density0;
        chk = chk && (avg <= eta);
        g->iterationsV++;
    }
    computePressureAccel(g->kappa_v);
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        g->v[i] = { g->v[i].x + h * g->pa[i].x, g->v[i].y + h * g->pa[i].y, g->v[i].z + h * g->pa[i].z };
        g->factor[i] *= h;
        g->kappa_v[i] *= h;
    }
}

// TimeStepDFSPH::pressureSolve (:252-384) + pressureSolveIteration (:544-619)
void pressureSolve()
{
    const Real h = g->h, h2 = h * h, invH2 = static_cast<Real>(1.0) / h2;
    const long n = (long)g->x.size();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        g->density_adv[i] = g->density[i] / g->density0 + h * velocityDivergence(i);
        g->factor[i] *= invH2;
        if (g->density_adv[i] > 1.0) g->kappa[i] = static_cast<Real>(0.5) * std::min(g->kappa[i], static_cast<Real>(0.00025)) * invH2;
        else g->kappa[i] = 0.0;
    }
    g->iterations = 0;
    Real avg = 0.0;
    bool chk = false;
    while ((!chk || (g->iterations < g->minIter)) && (g->iterations < g->maxIter)) {
        chk = true;
        avg = 0.0;
        if (n > 0) {
            computePressureAccel(g->kappa);
            Real err = 0.0;
            #pragma omp parallel for reduction(+:err) schedule(static)
            for (long i = 0; i < n; ++i) {
                const Real ap = aij_pj(i) * h * h;
                const Real s_i = static_cast<Real>(1.0) - g->density_adv[i];
                const Real residuum = std::min(s_i - ap, static_cast<Real>(0.0));
                g->kappa[i] = std::max(g->kappa[i] - static_cast<Real>(0.5) * (s_i - ap) * g->factor[i], static_cast<Real>(0.0));
                err -= g->density0 * residuum;
            }
            avg = err / n;
        }
        const Real eta = g->maxError * static_cast<Real>(0.01) * g->density0;
        chk = chk && (avg <= eta);
        g->iterations++;
    }
    computePressureAccel(g->kappa);
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        g->v[i] = { g->v[i].x + h * g->pa[i].x, g->v[i].y + h * g->pa[i].y, g->v[i].z + h * g->pa[i].z };
        g->kappa[i] *= h2;
    }
}

// TimeStep::clearAccelerations (TimeStep.cpp:35-50) + Viscosity_Standard::step (Viscosity/Viscosity_Standard.cpp:48-238 /
// 242-400): a_i = g + 10 mu sum_j (m_j/rho_j) (v_ij.x_ij)/(|x_ij|^2 + 0.01 h^2) gradW_ij (+ boundary term)
void computeNonPressureForces()
{
    const long n = (long)g->x.size();
    g->acc.resize(n);
    const Real h2 = g->support * g->support;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        V3 a = { g->gravity[0], g->gravity[1], g->gravity[2] };
        if (g->viscosityMethod == 1) {
            const V3 xi = g->x[i], vi = g->v[i];
            FOR_FLUID(i, j)
                const V3 xixj = xi - g->x[j];
                const V3 w = solverGradW(xixj);
                const Real s = static_cast<Real>(10.0) * g->viscosity * (g->V * g->density0 / g->density[j]) * dot(vi - g->v[j], xixj) / (dot(xixj, xixj) + static_cast<Real>(0.01) * h2);
                a = { a.x + s * w.x, a.y + s * w.y, a.z + s * w.z };
            END_FOR
            if (g->viscosityBoundary != 0.0) {
                FOR_BOUNDARY(i, j)
                    const V3 xixj = xi - g->bx[j];
                    const V3 w = solverGradW(xixj);
                    const Real s = static_cast<Real>(10.0) * g->viscosityBoundary * (g->density0 * g->bV[j] / g->density[i]) * dot(vi, xixj) / (dot(xixj, xixj) + static_cast<Real>(0.01) * h2);
                    a = { a.x + s * w.x, a.y + s * w.y, a.z + s * w.z };
                END_FOR
            }
        }
        g->acc[i] = a;
    }
}

// Simulation::updateTimeStepSize (Simulation.cpp:395-493)
void updateTimeStepSize()
{
    if (g->cflMethod != 1 && g->cflMethod != 2) return;
    const Real hOld = g->h;
    Real maxVel = 0.0;
    for (size_t i = 0; i < g->x.size(); ++i) {
        const V3 t = { g->v[i].x + g->acc[i].x * hOld, g->v[i].y + g->acc[i].y * hOld, g->v[i].z + g->acc[i].z * hOld };
        const Real m = dot(t, t);
        if (m > maxVel) maxVel = m;
    }
    if (maxVel < static_cast<Real>(1.0e-9)) maxVel = static_cast<Real>(1.0e-9);
    const Real diameter = static_cast<Real>(2.0) * g->radius;
    Real h = g->cflFactor * static_cast<Real>(0.4) * (diameter / (std::sqrt(maxVel)));
    h = std::min(h, g->cflMax);
    h = std::max(h, g->cflMin);
    if (g->cflMethod == 2 && g->iterations != 0) {
        Real h2 = hOld;
        if (g->iterations > 10) h2 *= static_cast<Real>(0.9);
        else if (g->iterations < 5) h2 *= static_cast<Real>(1.1);
        h = std::min(h2, h);
    }
    g->h = h;
}

// TimeStepDFSPH::step (TimeStepDFSPH.cpp:117-249)
void step()
{
    const Real h = g->h;
    const long n = (long)g->x.size();
    findNeighbors();
    computeDensities();
    computeFactor();
    if (g->enableDiv) divergenceSolve(); else g->iterationsV = 0;
    computeNonPressureForces();
    updateTimeStepSize();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i)
        g->v[i] = { g->v[i].x + h * g->acc[i].x, g->v[i].y + h * g->acc[i].y, g->v[i].z + h * g->acc[i].z };
    pressureSolve();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i)
        g->x[i] = { g->x[i].x + h * g->v[i].x, g->x[i].y + h * g->v[i].y, g->x[i].z + h * g->v[i].z };
    g->time += h;
}

}  // namespace

// ---- harness ABI (same as oracle/ref_driver.cpp) -------------------------------------------------------------------------
extern "C" {

int ref_sizeof_real() { return (int)sizeof(Real); }
int ref_uses_avx() { return 0; }
int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

int ref_create(double particleRadius)
{
    if (g) return -1;
    g = new Sim();
    g->radius = static_cast<Real>(particleRadius);
    initKernels();
    g->created = true;
    return 0;
}

int ref_add_fluid(const Real* x, const Real* v, unsigned n, double density0)
{
    if (!g->x.empty()) return -1;   // single fluid model
    g->density0 = static_cast<Real>(density0);
    const Real diam = static_cast<Real>(2.0) * g->radius;
    g->V = static_cast<Real>(0.8) * diam * diam * diam;   // FluidModel.cpp:242
    g->x.resize(n); g->v.resize(n); g->pa.assign(n, { 0, 0, 0 }); g->id.resize(n);
    for (unsigned i = 0; i < n; ++i) {
        g->x[i] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
        g->v[i] = v ? V3{ v[3 * i], v[3 * i + 1], v[3 * i + 2] } : V3{ 0, 0, 0 };
        g->id[i] = i;
    }
    g->density.assign(n, 0); g->factor.assign(n, 0); g->density_adv.assign(n, 0); g->kappa.assign(n, 0); g->kappa_v.assign(n, 0);
    return 0;
}

int ref_configure(int kernel, int gradKernel)
{
    g->kernel = (kernel < 0 || kernel > 4) ? 0 : kernel;                 // Simulation.cpp:346-347
    g->gradKernel = (gradKernel < 0 || gradKernel > 4) ? 0 : gradKernel; // Simulation.cpp:312-313
    initKernels();   // W_zero follows the configured kernel
    return 0;
}

int ref_add_boundary(const Real* x, unsigned n)
{
    for (unsigned i = 0; i < n; ++i) g->bx.push_back({ x[3 * i], x[3 * i + 1], x[3 * i + 2] });
    return 0;
}

int ref_finalize() { computeBoundaryVolume(); return 0; }

int ref_set_real(const char* name, double v)
{
    const std::string s(name);
    if (s == "timeStepSize") g->h = static_cast<Real>(v);
    else if (s == "maxError") g->maxError = std::max(static_cast<Real>(v), static_cast<Real>(1e-6));
    else if (s == "maxErrorV") g->maxErrorV = std::max(static_cast<Real>(v), static_cast<Real>(1e-6));
    else if (s == "cflFactor") g->cflFactor = static_cast<Real>(v);
    else if (s == "cflMinTimeStepSize") g->cflMin = static_cast<Real>(v);
    else if (s == "cflMaxTimeStepSize") g->cflMax = static_cast<Real>(v);
    else return -1;
    return 0;
}

int ref_set_int(const char* name, int v)
{
    const std::string s(name);
    if (s == "minIterations") g->minIter = (unsigned)v;
    else if (s == "maxIterations") g->maxIter = (unsigned)std::max(v, 1);
    else if (s == "maxIterationsV") g->maxIterV = (unsigned)std::max(v, 1);
    else if (s == "enableDivergenceSolver") g->enableDiv = v != 0;
    else if (s == "cflMethod") g->cflMethod = v;
    else if (s == "enableZSort" || s == "stepsPerZSort") {}
    else return -1;
    return 0;
}

int ref_set_viscosity(int, int method, double viscosity, double viscosityBoundary)
{
    if (method != 0 && method != 1) return -1;
    g->viscosityMethod = method;
    g->viscosity = static_cast<Real>(viscosity);
    g->viscosityBoundary = static_cast<Real>(viscosityBoundary);
    return 0;
}

int ref_set_gravity(double gx, double gy, double gz)
{
    g->gravity[0] = static_cast<Real>(gx); g->gravity[1] = static_cast<Real>(gy); g->gravity[2] = static_cast<Real>(gz);
    return 0;
}

int ref_step(int n)
{
    for (int k = 0; k < n; ++k) {
        auto t0 = std::chrono::high_resolution_clock::now();
        step();
        g->step_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    }
    return 0;
}
double ref_step_seconds() { return g->step_seconds; }
void ref_reset_step_seconds() { g->step_seconds = 0.0; }
double ref_avg_timer_ms(const char*) { return -1.0; }
int ref_search_and_density() { findNeighbors(); computeDensities(); return 0; }

unsigned ref_num_particles(int) { return (unsigned)g->x.size(); }
unsigned ref_num_boundary_particles(int) { return (unsigned)g->bx.size(); }
double ref_time() { return g->time; }
double ref_time_step_size() { return g->h; }
int ref_iterations() { return (int)g->iterations; }
int ref_iterations_v() { return (int)g->iterationsV; }
double ref_w_zero() { return g->simW_zero; }
double ref_fluid_volume(int) { return g->V; }
int ref_kernel() { return g->kernel; }

int ref_eval_kernel(int kind, unsigned n, const Real* r, Real* W, Real* gradW)
{
    for (unsigned i = 0; i < n; ++i) {
        const V3 x = { r[3 * i], r[3 * i + 1], r[3 * i + 2] };
        Real w; V3 gr;
        switch (kind) {
            case 0: w = cubicW(std::sqrt(dot(x, x))); gr = cubicGradW(x); break;
            case 1: w = wendlandW(std::sqrt(dot(x, x))); gr = wendlandGradW(x); break;
            case 2: w = poly6W(x); gr = poly6GradW(x); break;
            case 3: w = spikyW(x); gr = spikyGradW(x); break;
            case 4: w = lutWf(x); gr = lutGradW(x); break;
            default: return -1;
        }
        W[i] = w; gradW[3 * i] = gr.x; gradW[3 * i + 1] = gr.y; gradW[3 * i + 2] = gr.z;
    }
    return 0;
}

int ref_get_field(int, const char* name, Real* out, int dim)
{
    const std::string s(name);
    const size_t n = g->x.size();
    const std::vector<V3>* v3 = nullptr;
    const std::vector<Real>* v1 = nullptr;
    if (s == "position") v3 = &g->x; else if (s == "velocity") v3 = &g->v; else if (s == "pressure acceleration") v3 = &g->pa;
    else if (s == "density") v1 = &g->density; else if (s == "factor") v1 = &g->factor; else if (s == "advected density") v1 = &g->density_adv;
    else if (s == "p / rho^2") v1 = &g->kappa; else if (s == "p_v / rho^2") v1 = &g->kappa_v;
    else if (s == "acceleration") { for (size_t i = 0; i < n; ++i) { const V3 a = i < g->acc.size() ? g->acc[i] : V3{ g->gravity[0], g->gravity[1], g->gravity[2] }; out[3 * i] = a.x; out[3 * i + 1] = a.y; out[3 * i + 2] = a.z; } return 0; }
    else return -1;
    if (v3) for (size_t i = 0; i < n; ++i) { out[3 * i] = (*v3)[i].x; out[3 * i + 1] = (*v3)[i].y; out[3 * i + 2] = (*v3)[i].z; }
    if (v1) for (size_t i = 0; i < n; ++i) out[i] = (*v1)[i];
    (void)dim;
    return 0;
}

int ref_get_ids(int, unsigned* out) { std::copy(g->id.begin(), g->id.end(), out); return 0; }

int ref_set_state(int, const Real* x, const Real* v)
{
    for (size_t i = 0; i < g->x.size(); ++i) {
        if (x) g->x[i] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
        if (v) g->v[i] = { v[3 * i], v[3 * i + 1], v[3 * i + 2] };
    }
    return 0;
}

int ref_set_field(int, const char* name, const Real* in, int dim)
{
    const std::string s(name);
    const size_t n = g->x.size();
    std::vector<V3>* v3 = nullptr;
    std::vector<Real>* v1 = nullptr;
    if (s == "position") v3 = &g->x; else if (s == "velocity") v3 = &g->v;
    else if (s == "p / rho^2") v1 = &g->kappa; else if (s == "p_v / rho^2") v1 = &g->kappa_v;
    else return -1;
    if (v3) for (size_t i = 0; i < n; ++i) (*v3)[i] = { in[3 * i], in[3 * i + 1], in[3 * i + 2] };
    if (v1) for (size_t i = 0; i < n; ++i) (*v1)[i] = in[i];
    (void)dim;
    return 0;
}

int ref_get_boundary(int, Real* x, Real* V)
{
    for (size_t i = 0; i < g->bx.size(); ++i) {
        if (x) { x[3 * i] = g->bx[i].x; x[3 * i + 1] = g->bx[i].y; x[3 * i + 2] = g->bx[i].z; }
        if (V) V[i] = g->bV[i];
    }
    return 0;
}

int ref_neighbor_counts(int, int pid, unsigned* counts)
{
    const auto& off = pid == 0 ? g->off_f : g->off_b;
    for (size_t i = 0; i + 1 < off.size(); ++i) counts[i] = (unsigned)(off[i + 1] - off[i]);
    return 0;
}

int ref_neighbor_lists(int, int pid, const unsigned long long*, unsigned* idx)
{
    const auto& nb = pid == 0 ? g->nbr_f : g->nbr_b;
    std::copy(nb.begin(), nb.end(), idx);
    return 0;
}

int ref_destroy() { delete g; g = nullptr; return 0; }

}
