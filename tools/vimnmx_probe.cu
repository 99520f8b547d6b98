// Probe: does a fully unrolled int min/max tree (ptxas fuses it into 3-input VIMNMX3 on sm_100a)
// compute the same FAST corner score as a compare/bit-logic bisection?  Prints mismatch counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o vimnmx_probe tools/vimnmx_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ int score_tree(const int* d)
{
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
    int best = -1000;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int lo9 = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int hi9 = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = max(best, max(lo9, -hi9));
    }
    return best - 1;
}
__device__ int score_ref(const int* d)
{
    int best = -1000;
    for (int k = 0; k < 16; ++k) {
        int mn = 1000, mx = -1000;
        for (int j = 0; j < 9; ++j) { int v = d[(k + j) & 15]; if (v < mn) mn = v; if (v > mx) mx = v; }
        int a = mn > -mx ? mn : -mx;
        if (a > best) best = a;
    }
    return best - 1;
}
__global__ void probe(const int* in, int n, int* bad, int* first)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = in[i * 16 + k];
    int a = score_tree(d), b = score_ref(d);
    if (a != b) { if (atomicAdd(bad, 1) == 0) { first[0] = i; first[1] = a; first[2] = b; } }
}

// same tree, but the differences come from BYTES in shared memory (centre minus ring pixel), as in k_fast_cells:
// the compiler now knows every d[k] fits in 9 bits and is free to pick narrower / packed min-max forms
__global__ void probe_u8(const unsigned char* in, int n, int* bad, int* first)
{
    __shared__ unsigned char s[256 * 17];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < 17; ++k) s[threadIdx.x * 17 + k] = i < n ? in[(size_t)i * 17 + k] : 0;
    __syncthreads();
    if (i >= n) return;
    const unsigned char* p = &s[threadIdx.x * 17];
    const int v = p[0];
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - p[1 + k];
    int a = score_tree(d), b = score_ref(d);
    if (a != b) { if (atomicAdd(bad, 1) == 0) { first[0] = i; first[1] = a; first[2] = b; } }
}
int main()
{
    const int n = 1 << 20;
    std::vector<int> h(n * 16);
    srand(7);
    for (auto& v : h) v = rand() % 201 - 100;
    int *d_in, *d_bad, *d_first;
    cudaMalloc(&d_in, h.size() * 4); cudaMalloc(&d_bad, 4); cudaMalloc(&d_first, 12);
    cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_bad, 0, 4);
    probe<<<n / 256, 256>>>(d_in, n, d_bad, d_first);
    int bad = -1, first[3] = {0, 0, 0};
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(first, d_first, 12, cudaMemcpyDeviceToHost);
    printf("vimnmx_probe: %d of %d mismatches (first idx %d tree %d ref %d) err=%s\n", bad, n, first[0], first[1], first[2], cudaGetErrorString(cudaGetLastError()));
    std::vector<unsigned char> hb((size_t)n * 17);
    for (auto& v : hb) v = (unsigned char)(rand() & 255);
    unsigned char* d_b; cudaMalloc(&d_b, hb.size());
    cudaMemcpy(d_b, hb.data(), hb.size(), cudaMemcpyHostToDevice);
    cudaMemset(d_bad, 0, 4);
    probe_u8<<<n / 256, 256>>>(d_b, n, d_bad, d_first);
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(first, d_first, 12, cudaMemcpyDeviceToHost);
    printf("vimnmx_probe_u8: %d of %d mismatches (first idx %d tree %d ref %d) err=%s\n", bad, n, first[0], first[1], first[2], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
