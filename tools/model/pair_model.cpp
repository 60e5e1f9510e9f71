// Development tool: a sequential CPU model of the child-pair traversal (csrc/traversal_pairs.cuh) next to the reference's
// one-node-per-visit traversal, run over a dump of rays.  Checks that hits and visit counts agree and prints the statistics
// that size the kernel (records loaded per ray, stack depths, pops that need no memory access).
//
//   g++ -O2 -std=c++20 -ffp-contract=off -I rayfinder_b200/csrc tools/model/pair_model.cpp -o /tmp/pair_model
//   /tmp/pair_model nodes.bin tris9.bin rays.bin kinds.bin        (written by tools/model/dump_rays.py)
#include "pair_records.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

using namespace rfb200;

rf_status rfb200::setError(rf_status code, const char*, ...) { return code; }

template<class T>
static std::vector<T> readAll(const char* path)
{
    std::ifstream     f(path, std::ios::binary | std::ios::ate);
    const std::size_t n = static_cast<std::size_t>(f.tellg());
    std::vector<T>    v(n / sizeof(T));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
    return v;
}

struct Slab
{
    bool  ok;   // tmax-independent part
    float tmin; // compared with tmax
};
// ray_intersection.cpp:101-136, split into its tmax-independent part and tmin
static Slab slab(const float* lo, const float* hi, const float* o, const float* inv, const int* neg)
{
    const float* b[2] = {lo, hi};
    float        tmin = (b[neg[0]][0] - o[0]) * inv[0];
    float        tmax = (b[1 - neg[0]][0] - o[0]) * inv[0];
    const float  tymin = (b[neg[1]][1] - o[1]) * inv[1];
    const float  tymax = (b[1 - neg[1]][1] - o[1]) * inv[1];
    bool         hit = !((tmin > tymax) || (tymin > tmax));
    tmin = std::max(tymin, tmin);
    tmax = std::min(tymax, tmax);
    const float tzmin = (b[neg[2]][2] - o[2]) * inv[2];
    const float tzmax = (b[1 - neg[2]][2] - o[2]) * inv[2];
    hit = hit && !((tmin > tzmax) || (tzmin > tmax));
    tmin = std::max(tzmin, tmin);
    tmax = std::min(tzmax, tmax);
    return {hit && (tmax > 0.0f), tmin};
}

static bool triangle(const float* t9, const float* o, const float* d, float tmax, float& tOut)
{
    const V3    v0 = v3(t9), e1 = v3(t9 + 3) - v0, e2 = v3(t9 + 6) - v0, dd = v3(d), oo = v3(o);
    const V3    h = cross(dd, e2);
    const float det = dot(e1, h);
    if (det > -0.00001f && det < 0.00001f) return false;
    const float invDet = 1.0f / det;
    const V3    s = oo - v0;
    const float u = invDet * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    const V3    q = cross(s, e1);
    const float v = invDet * dot(dd, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    const float t = invDet * dot(e2, q);
    if (t > 0.00001f && t < tmax)
    {
        tOut = t;
        return true;
    }
    return false;
}

struct Result
{
    std::uint32_t tri = 0xFFFFFFFFu, nodes = 0, tris = 0;
    float         t = 0.f;
};

int main(int argc, char** argv)
{
    if (argc < 5) return 1;
    const auto nodes = readAll<rf_bvh_node>(argv[1]);
    const auto tris = readAll<float>(argv[2]);
    const auto rays = readAll<float>(argv[3]);
    const auto kinds = readAll<std::uint8_t>(argv[4]);
    const PairScene ps = buildPairRecords(nodes.data(), nodes.size());
    std::printf("nodes %zu records %zu usable %d rays %zu\n", nodes.size(), ps.records.size(), ps.usable, kinds.size());

    std::uint64_t visits = 0, loadsRef = 0, expands = 0, freePops = 0, hitPops = 0, pushes = 0, pushesDefiniteMiss = 0, pushesAlways = 0, mismatches = 0;
    std::uint64_t depthHist[40] = {}, depthHistSkip[40] = {}, triTests = 0, leafVisits = 0;
    // breadth-first rank of every record (what a top-of-tree treelet in shared memory would hold first)
    std::vector<std::uint32_t> bfsRank(ps.records.size(), 0xFFFFFFFFu);
    {
        std::vector<std::uint32_t> queue;
        if (!pairLinkIsLeaf(ps.rootLink)) queue.push_back(ps.rootLink);
        for (std::size_t head = 0; head < queue.size(); ++head)
        {
            const std::uint32_t r = queue[head];
            bfsRank[r] = static_cast<std::uint32_t>(head);
            if (!pairLinkIsLeaf(ps.records[r].link0)) queue.push_back(ps.records[r].link0);
            if (!pairLinkIsLeaf(ps.records[r].link1)) queue.push_back(ps.records[r].link1);
        }
    }
    const std::uint32_t treeletSizes[8] = {64, 128, 256, 512, 1024, 2048, 4096, 8192};
    std::uint64_t       treeletHits[8] = {};
    for (std::size_t r = 0; r < kinds.size(); ++r)
    {
        const float* o = &rays[6 * r];
        const float* d = o + 3;
        const bool   anyHit = kinds[r] != 0;
        const float  inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        const int    neg[3] = {inv[0] < 0.f, inv[1] < 0.f, inv[2] < 0.f};
        // ---- reference
        Result a;
        {
            float         tmax = 10000.0f;
            std::uint32_t stack[64], sp = 0, cur = 0;
            for (;;)
            {
                ++a.nodes;
                const rf_bvh_node& n = nodes[cur];
                const Slab         s = slab(n.aabb_min, n.aabb_max, o, inv, neg);
                bool               done = false;
                if (s.ok && s.tmin < tmax)
                {
                    if (n.triangle_count > 0)
                    {
                        for (std::uint32_t k = 0; k < n.triangle_count && !done; ++k)
                        {
                            ++a.tris;
                            float t;
                            if (triangle(&tris[9 * (n.triangles_offset + k)], o, d, tmax, t))
                            {
                                a.tri = n.triangles_offset + k, a.t = t;
                                if (anyHit) done = true; else tmax = t;
                            }
                        }
                        if (done || sp == 0) break;
                        cur = stack[--sp];
                    }
                    else if (neg[n.split_axis])
                    {
                        stack[sp++] = cur + 1;
                        cur = n.second_child_offset;
                    }
                    else
                    {
                        stack[sp++] = n.second_child_offset;
                        cur = cur + 1;
                    }
                }
                else
                {
                    if (sp == 0) break;
                    cur = stack[--sp];
                }
            }
        }
        // ---- child pairs
        Result b;
        {
            float tmax = 10000.0f;
            struct Entry { std::uint32_t link; float t; bool definite; };
            Entry         stack[64];
            std::uint32_t sp = 0, maxDepth = 0, spSkip = 0, maxDepthSkip = 0;
            std::uint32_t skipDepthAt[64];
            ++b.nodes; // the root's visit
            const Slab    root = slab(ps.rootBox, ps.rootBox + 3, o, inv, neg);
            std::uint32_t link = ps.rootLink;
            bool          enter = root.ok && root.tmin < tmax, done = false;
            for (;;)
            {
                if (enter && !pairLinkIsLeaf(link))
                {
                    // EXPAND: one record, two slab tests
                    ++expands;
                    for (int k = 0; k < 8; ++k) treeletHits[k] += bfsRank[link] < treeletSizes[k];
                    const PairRecord& rec = ps.records[link];
                    const Slab        s0 = slab(rec.box0, rec.box0 + 3, o, inv, neg), s1 = slab(rec.box1, rec.box1 + 3, o, inv, neg);
                    const bool        nearIsSecond = neg[rec.meta];
                    const Slab        sn = nearIsSecond ? s1 : s0, sf = nearIsSecond ? s0 : s1;
                    const std::uint32_t ln = nearIsSecond ? rec.link1 : rec.link0, lf = nearIsSecond ? rec.link0 : rec.link1;
                    const float       tf = sf.ok ? sf.tmin : __builtin_inff();
                    skipDepthAt[sp] = spSkip;
                    stack[sp++] = Entry{lf, tf, !sf.ok};
                    ++pushes;
                    if (!sf.ok) ++pushesDefiniteMiss; else { ++spSkip; if (sf.tmin <= 0.f) ++pushesAlways; }
                    maxDepth = std::max(maxDepth, sp), maxDepthSkip = std::max(maxDepthSkip, spSkip);
                    ++b.nodes; // the near child's visit
                    link = ln;
                    enter = sn.ok && sn.tmin < tmax;
                    continue;
                }
                if (enter)
                {
                    ++leafVisits;
                    const std::uint32_t first = link & 0xFFFFFFu, count = ((link >> 24) & 127u) + 1u;
                    for (std::uint32_t k = 0; k < count && !done; ++k)
                    {
                        ++b.tris;
                        float t;
                        if (triangle(&tris[9 * (first + k)], o, d, tmax, t))
                        {
                            b.tri = first + k, b.t = t;
                            if (anyHit) done = true; else tmax = t;
                        }
                    }
                    if (done) break;
                }
                // POP
                if (sp == 0) break;
                const Entry e = stack[--sp];
                spSkip = skipDepthAt[sp];
                ++b.nodes;
                link = e.link;
                enter = e.t < tmax;
                if (enter) ++hitPops; else ++freePops;
            }
            ++depthHist[std::min(maxDepth, 39u)];
            ++depthHistSkip[std::min(maxDepthSkip, 39u)];
        }
        visits += a.nodes, loadsRef += a.nodes, triTests += a.tris;
        if (a.tri != b.tri || a.nodes != b.nodes || a.tris != b.tris || std::memcmp(&a.t, &b.t, 4) != 0)
        {
            if (mismatches++ < 5) std::printf("MISMATCH ray %zu: ref tri %u nodes %u tris %u | pairs tri %u nodes %u tris %u\n", r, a.tri, a.nodes, a.tris, b.tri, b.nodes, b.tris);
        }
    }
    const double n = static_cast<double>(kinds.size());
    std::printf("mismatches %llu\n", (unsigned long long)mismatches);
    std::printf("per ray: visits %.2f (= node loads of the per-node layout)  records loaded %.2f  triangle tests %.2f  leaf visits %.2f\n", visits / n, expands / n, triTests / n, leafVisits / n);
    std::printf("pops: %.2f hit, %.2f free misses per ray;  pushes %.2f per ray: %.1f%% definite misses, %.1f%% always-hit (tmin <= 0)\n", hitPops / n, freePops / n, pushes / n,
                100.0 * pushesDefiniteMiss / pushes, 100.0 * pushesAlways / pushes);
    for (int k = 0; k < 8; ++k) std::printf("records loaded from the first %5u records in breadth-first order: %.1f%%\n", treeletSizes[k], 100.0 * treeletHits[k] / expands);
    std::printf("max stack depth per ray (all far children pushed / definite misses not pushed):\n");
    std::uint64_t cum = 0, cumSkip = 0;
    for (int k = 0; k < 40; ++k)
    {
        cum += depthHist[k], cumSkip += depthHistSkip[k];
        if (depthHist[k] || depthHistSkip[k]) std::printf("  depth %2d: %9llu (cum %.4f)   %9llu (cum %.4f)\n", k, (unsigned long long)depthHist[k], cum / n, (unsigned long long)depthHistSkip[k], cumSkip / n);
    }
    return mismatches != 0;
}
