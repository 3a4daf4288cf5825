// CPU model of the L1 data-pipe cost of the neighbour gathers (no GPU needed): for a jittered lattice block, build the
// particle order and the neighbour table exactly as the device code does and count, per warp-level gather of 16-byte
// records, the distinct 128-byte lines each quarter-warp (8 lanes) touches -- the quantity that the LDG.128 gathers of
// the sweeps pay for (profiles/r1_gather_microbench.md).  Used to compare orderings / list layouts before building them.
//   g++ -O2 -o wavefront_model wavefront_model.cpp && ./wavefront_model [n_per_axis] [jitter] [mode]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <map>
#include <vector>
struct P { float x, y, z; };
int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 64;
    const float jitter = argc > 2 ? atof(argv[2]) : 0.2f;   // fraction of the spacing
    const int order = argc > 3 ? atoi(argv[3]) : 0;         // 0: 8^3 blocks, x fastest in block; 1: Morton in block
    const int cluster = argc > 4 ? atoi(argv[4]) : 1;       // particles per lane (1 or 2 or 4)
    const int sortlist = argc > 5 ? atoi(argv[5]) : 0;      // 1: lists sorted by address
    const int split = argc > 6 ? atoi(argv[6]) : 1;         // lanes per particle (1, 2, 4, 8): lane s of particle p reads entries k*split + s
    const float d = 0.05f, R = 0.1f, S = R * 1.00001f;
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> U(-jitter * d, jitter * d);
    std::vector<P> p;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) for (int k = 0; k < n; ++k) p.push_back({(i + 0.5f) * d + U(rng), (j + 0.5f) * d + U(rng), (k + 0.5f) * d + U(rng)});
    const int N = (int)p.size();
    const int nc = (int)std::ceil(n * d / S) + 1, nb = (nc + 7) / 8;
    auto spread3 = [](unsigned v) { return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4); };
    auto key_of = [&](int cx, int cy, int cz) {
        const unsigned b = ((unsigned)(cx >> 3) * nb + (cy >> 3)) * nb + (cz >> 3);
        // block rank: Morton over block coords
        unsigned br = 0; { unsigned bx = cx >> 3, by = cy >> 3, bz = cz >> 3; for (int t = 0; t < 8; ++t) br |= ((bx >> t) & 1u) << (3 * t) | ((by >> t) & 1u) << (3 * t + 1) | ((bz >> t) & 1u) << (3 * t + 2); }
        (void)b;
        const unsigned l = order != 1 ? ((cx & 7) | ((cy & 7) << 3) | ((cz & 7) << 6)) : (spread3(cx & 7) | (spread3(cy & 7) << 1) | (spread3(cz & 7) << 2));
        return (unsigned long long)br * 512ull + l;
    };
    struct Rec { unsigned long long key; unsigned fine; int idx; float x; };
    std::vector<Rec> recs(N);
    std::vector<int> ccx(N), ccy(N), ccz(N);
    for (int i = 0; i < N; ++i) {
        const float tx = p[i].x / S, ty = p[i].y / S, tz = p[i].z / S;
        const int cx = std::max(0, (int)std::floor(tx)), cy = std::max(0, (int)std::floor(ty)), cz = std::max(0, (int)std::floor(tz));
        const unsigned sx = std::min(7, std::max(0, (int)((tx - cx) * 8))), sy = std::min(7, std::max(0, (int)((ty - cy) * 8))), sz = std::min(7, std::max(0, (int)((tz - cz) * 8)));
        recs[i] = {key_of(cx, cy, cz), spread3(sx) | (spread3(sy) << 1) | (spread3(sz) << 2), i, p[i].x};
        if (order >= 6 && order <= 8) {   // pencil order inside B^3-cell blocks (B = 4, 2, 8 for order 6, 7, 8): fine rows of half a cell
            const int B = order == 6 ? 4 : (order == 7 ? 2 : 8), F = 2;
            const int bx = cx / B, by = cy / B, bz = cz / B;
            unsigned long long br = 0; for (int t = 0; t < 10; ++t) br |= (unsigned long long)((bx >> t) & 1) << (3 * t) | (unsigned long long)((by >> t) & 1) << (3 * t + 1) | (unsigned long long)((bz >> t) & 1) << (3 * t + 2);
            const unsigned fy = std::min(B * F - 1, (int)((ty - by * B) * F)), fz = std::min(B * F - 1, (int)((tz - bz * B) * F));
            recs[i].key = br * 4096ull + fz * 64ull + fy;
            recs[i].fine = 0;
        } else if (order == 5) {   // x fastest in block, cells split into 4 x-slices, sub-position Morton inside a slice
            recs[i].key = key_of(cx, cy, cz) * 4ull + (sx >> 1);
        } else if (order >= 2) {
            // pencil order: block (bx x 8 cells), then fine row (z, y at 1/F cell), then x
            const int F = order == 2 ? 2 : (order == 3 ? 4 : 1);
            const unsigned fy = std::min(8 * F - 1, (int)((ty - (cy & ~7)) * F)), fz = std::min(8 * F - 1, (int)((tz - (cz & ~7)) * F));
            recs[i].key = (key_of(cx & ~7, cy & ~7, cz & ~7) >> 9) * 4096ull + fz * 64ull + fy;
            recs[i].fine = 0;
        }
    }
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.key != b.key ? a.key < b.key : (a.fine != b.fine ? a.fine < b.fine : (a.x != b.x ? a.x < b.x : a.idx < b.idx)); });
    std::vector<P> q(N);
    for (int i = 0; i < N; ++i) q[i] = p[recs[i].idx];
    // cell table over sorted particles
    std::vector<std::vector<int>> cells((size_t)nc * nc * nc);
    auto cid = [&](int cx, int cy, int cz) { return ((size_t)cx * nc + cy) * nc + cz; };
    for (int i = 0; i < N; ++i) {
        ccx[i] = std::max(0, (int)std::floor(q[i].x / S)); ccy[i] = std::max(0, (int)std::floor(q[i].y / S)); ccz[i] = std::max(0, (int)std::floor(q[i].z / S));
        cells[cid(ccx[i], ccy[i], ccz[i])].push_back(i);
    }
    // neighbour lists in walk order (z outer, y, x inner)
    std::vector<std::vector<int>> nl(N);
    for (int i = 0; i < N; ++i)
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            const int x = ccx[i] + dx, y = ccy[i] + dy, z = ccz[i] + dz;
            if (x < 0 || y < 0 || z < 0 || x >= nc || y >= nc || z >= nc) continue;
            std::vector<int> cand = cells[cid(x, y, z)];
            for (int j : cand) {
                if (j == i) continue;
                const float ax = q[i].x - q[j].x, ay = q[i].y - q[j].y, az = q[i].z - q[j].z;
                if (ax * ax + ay * ay + az * az < R * R) nl[i].push_back(j);
            }
        }
    // lanes: `cluster` consecutive particles; lane list = union in walk order of first particle, then extras
    const int L = (N + cluster - 1) / cluster;
    std::vector<std::vector<int>> ll(L);
    double real_pairs = 0;
    for (int l = 0; l < L; ++l) {
        std::vector<int> u;
        for (int c = 0; c < cluster; ++c) {
            const int i = l * cluster + c;
            if (i >= N) break;
            real_pairs += nl[i].size();
            for (int j : nl[i]) if (std::find(u.begin(), u.end(), j) == u.end()) u.push_back(j);
        }
        if (sortlist || cluster > 1) std::sort(u.begin(), u.end());
        ll[l] = u;
    }
    if (split > 1) {
        // `split` lanes per particle: a warp holds 32/split particles, step k of lane (p, s) reads entry k*split + s
        const int PW = 32 / split;
        double wsteps = 0, lines = 0, entries = 0;
        for (int t = 0; t < (N + PW - 1) / PW; ++t) {
            size_t mx = 0;
            for (int i = t * PW; i < std::min(N, t * PW + PW); ++i) mx = std::max(mx, (ll[i].size() + split - 1) / split);
            wsteps += mx;
            for (size_t k = 0; k < mx; ++k) {
                std::map<int, int> wl;
                for (int i = t * PW; i < std::min(N, t * PW + PW); ++i)
                    for (int s2 = 0; s2 < split; ++s2) { const size_t e = k * split + s2; const int j = e < ll[i].size() ? ll[i][e] : N; wl[j / 8]++; if (e < ll[i].size()) entries += 1; }
                for (auto& kv : wl) lines += (kv.second + 7) / 8;
            }
        }
        printf("N=%d order=%d split=%d sorted=%d jitter=%.2f: neighbours/particle %.2f, warp-steps/particle %.4f (x32 = %.1f slots/particle), lines per warp-gather %.2f, lines/particle %.3f\n",
               N, order, split, sortlist, jitter, entries / N, wsteps / N, 32.0 * wsteps / N, lines / wsteps, lines / N);
        return 0;
    }
    // per warp tile: padded length = max rounded to 4; count lines per quarter-warp
    double steps = 0, wf = 0, slots = 0, wf_warp = 0;
    for (int t = 0; t < (L + 31) / 32; ++t) {
        size_t mx = 0;
        for (int l = t * 32; l < std::min(L, t * 32 + 32); ++l) mx = std::max(mx, ll[l].size());
        mx = (mx + 3) & ~(size_t)3;
        steps += mx;
        for (size_t k = 0; k < mx; ++k) {
            std::map<int, int> wl;
            for (int l = t * 32; l < std::min(L, t * 32 + 32); ++l) { const int j = k < ll[l].size() ? ll[l][k] : N; wl[j / 8]++; }
            for (auto& kv : wl) { wf_warp += (kv.second + 7) / 8; }
            for (int qd = 0; qd < 4; ++qd) {
                std::set<int> lines;
                for (int l = t * 32 + qd * 8; l < std::min(L, t * 32 + qd * 8 + 8); ++l) {
                    const int j = k < ll[l].size() ? ll[l][k] : N;   // sentinel
                    lines.insert(j / 8);
                    if (k < ll[l].size()) slots += 1;
                }
                wf += lines.size();
            }
        }
    }
    printf("N=%d order=%d cluster=%d sorted=%d jitter=%.2f: neighbours/particle %.2f, list entries/particle %.2f, warp-steps/particle %.4f, lines per warp-gather %.2f, wavefronts/particle %.3f | per-warp line model: %.2f per gather, %.3f per particle\n",
           N, order, cluster, sortlist, jitter, real_pairs / N, slots / N, steps / N, wf / steps, wf / N, wf_warp / steps, wf_warp / N);
    return 0;
}
