// dev probe: how expensive is first-touch growth of the stream-ordered memory pool?
#include <chrono>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaSetDevice(0);
    cudaFree(0);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    const size_t GB = 1ull << 30;
    {   // (a) 40 x 100 MB
        std::vector<void*> p(40); double t = now();
        for (auto &q : p) cudaMallocAsync(&q, 100ull << 20, st);
        cudaStreamSynchronize(st); double t1 = now();
        for (auto &q : p) cudaFreeAsync(q, st);
        cudaStreamSynchronize(st);
        printf("40 x 100 MB first time: %.1f ms\n", t1 - t);
        t = now();
        for (auto &q : p) cudaMallocAsync(&q, 100ull << 20, st);
        cudaStreamSynchronize(st); t1 = now();
        for (auto &q : p) cudaFreeAsync(q, st);
        cudaStreamSynchronize(st);
        printf("40 x 100 MB again: %.1f ms\n", t1 - t);
    }
    {   // (b) one 8 GB block (pool must grow by ~4 GB)
        void *q; double t = now();
        cudaMallocAsync(&q, 8 * GB, st); cudaStreamSynchronize(st); double t1 = now();
        cudaFreeAsync(q, st); cudaStreamSynchronize(st);
        printf("1 x 8 GB first time: %.1f ms\n", t1 - t);
        std::vector<void*> p(60); t = now();
        for (auto &r : p) cudaMallocAsync(&r, 128ull << 20, st);
        cudaStreamSynchronize(st); t1 = now();
        printf("60 x 128 MB carved from it: %.1f ms\n", t1 - t);
        for (auto &r : p) cudaFreeAsync(r, st);
        cudaStreamSynchronize(st);
    }
    {   // (c) plain cudaMalloc
        void *q; double t = now(); cudaMalloc(&q, 4 * GB); double t1 = now(); cudaFree(q);
        printf("cudaMalloc 4 GB: %.1f ms, free %.1f ms\n", t1 - t, now() - t1);
    }
    return 0;
}
