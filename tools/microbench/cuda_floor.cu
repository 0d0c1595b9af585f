// The floor of any CUDA process on this box: driver initialisation + one primary context + one empty kernel.
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void empty() {}
int main() {
    auto t0 = std::chrono::steady_clock::now();
    int n = 0;
    cudaGetDeviceCount(&n);
    auto t1 = std::chrono::steady_clock::now();
    cudaSetDevice(0);
    cudaFree(nullptr);
    auto t2 = std::chrono::steady_clock::now();
    empty<<<1, 1>>>();
    cudaDeviceSynchronize();
    auto t3 = std::chrono::steady_clock::now();
    auto s = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    printf("{\"visible_gpus\": %d, \"driver_init_s\": %.4f, \"context_s\": %.4f, \"first_kernel_s\": %.4f, \"floor_s\": %.4f}\n", n, s(t0, t1),
           s(t1, t2), s(t2, t3), s(t0, t3));
    fflush(stdout);
    _Exit(0);
}
