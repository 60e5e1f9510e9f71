// Dependent-load latency of 32-byte node records on B200: what bounds a lone ray's BVH walk.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chase chase.cu && ./chase
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

struct __align__(32) Node { float f[6]; uint32_t a, b; };

template<int MODE> // 0: nc v8 load, 1: plain v8 load, 2: nc v8 + slab-test-like ALU chain, 3: two 16 B loads
__global__ void chase(const Node* nodes, uint32_t start, int steps, long long* cycles, uint32_t* sink)
{
    uint32_t cur = start;
    float acc = 0.f;
    const float ox = 0.1f, oy = 0.2f, oz = 0.3f, ix = 1.5f, iy = -2.5f, iz = 0.7f;
    const long long t0 = clock64();
    for (int k = 0; k < steps; ++k)
    {
        Node n;
        if (MODE == 1)
            asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(n.f[0]), "=f"(n.f[1]), "=f"(n.f[2]), "=f"(n.f[3]), "=f"(n.f[4]), "=f"(n.f[5]), "=r"(n.a), "=r"(n.b) : "l"(nodes + cur));
        else if (MODE == 3)
        {
            const float4 lo = __ldg(reinterpret_cast<const float4*>(nodes + cur));
            const uint4  hi = __ldg(reinterpret_cast<const uint4*>(nodes + cur) + 1);
            n.f[0] = lo.x, n.f[1] = lo.y, n.f[2] = lo.z, n.f[3] = lo.w, n.f[4] = __uint_as_float(hi.x), n.f[5] = __uint_as_float(hi.y), n.a = hi.z, n.b = hi.w;
        }
        else
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(n.f[0]), "=f"(n.f[1]), "=f"(n.f[2]), "=f"(n.f[3]), "=f"(n.f[4]), "=f"(n.f[5]), "=r"(n.a), "=r"(n.b) : "l"(nodes + cur));
        if (MODE == 2)
        {
            const float x0 = (n.f[0] - ox) * ix, x1 = (n.f[3] - ox) * ix;
            const float y0 = (n.f[1] - oy) * iy, y1 = (n.f[4] - oy) * iy;
            const float z0 = (n.f[2] - oz) * iz, z1 = (n.f[5] - oz) * iz;
            const float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
            const float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
            const bool hit = tmin <= tmax && tmax > -1e30f; // always true for finite data: keeps the chain, not the branch
            acc += tmin;
            cur = hit ? n.a : n.b;
        }
        else
            cur = n.a;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = cur + (acc == 1.2345f);
}

int main()
{
    long long* dCycles; uint32_t* dSink;
    cudaMalloc(&dCycles, 8); cudaMalloc(&dSink, 4 << 20);
    for (size_t bytes : {size_t(16) << 10, size_t(16) << 20, size_t(2) << 30})
    {
        const size_t count = bytes / sizeof(Node);
        std::vector<uint32_t> perm(count);
        std::iota(perm.begin(), perm.end(), 0u);
        std::mt19937 rng(1);
        std::shuffle(perm.begin() + 1, perm.end(), rng);
        std::vector<Node> h(count);
        for (size_t i = 0; i < count; ++i)
        {
            Node& n = h[perm[i]];
            for (int k = 0; k < 6; ++k) n.f[k] = float((i * 7 + k) % 13) * 0.25f;
            n.a = perm[(i + 1) % count], n.b = n.a;
        }
        Node* d; cudaMalloc(&d, bytes);
        cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
        const int steps = 20000;
        auto run = [&](auto kernel, const char* name, int threads, int blocks) {
            long long c = 0;
            for (int rep = 0; rep < 3; ++rep)
            {
                kernel<<<blocks, threads>>>(d, 0u, steps, dCycles, dSink);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c, dCycles, 8, cudaMemcpyDeviceToHost);
            printf("  %-44s %3d thr x %3d blk: %7.1f cycles/step\n", name, threads, blocks, double(c) / steps);
        };
        printf("working set %zu KB (%s)\n", bytes >> 10, bytes <= (64 << 10) ? "L1" : bytes <= (64 << 20) ? "L2" : "DRAM");
        run(chase<0>, "ld.global.nc.v8.b32", 1, 1);
        run(chase<1>, "ld.global.v8.b32", 1, 1);
        run(chase<3>, "2 x ld.global.nc.v4", 1, 1);
        run(chase<2>, "nc.v8 + slab-test ALU chain", 1, 1);
        run(chase<2>, "nc.v8 + slab-test ALU chain, full warp", 32, 1);
        cudaFree(d);
    }
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
