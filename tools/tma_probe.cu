// Development probe: which form of the 3-D u8 tensor-map load (cp.async.bulk.tensor, SASS UTMALDG) runs on the B200 -- descriptor in the kernel's
// parameter space (__grid_constant__) or in global memory -- and does the 80x80 box arrive as the plain loads see it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe.bin tools/tma_probe.cu && tools/tma_probe.bin
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

struct Maps { CUtensorMap level[16]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k_probe(const __grid_constant__ Maps maps, const CUtensorMap* gmaps, int level, const uint8_t* src, int pitch, long long fstride,
                                               int x0, int y0, int z, int box, int* mismatches)
{
    __shared__ __align__(128) uint8_t s_img[80 * 80];
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x;
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box * box) : "memory");
        const CUtensorMap* m = MODE == 0 ? &maps.level[level] : &gmaps[level];
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(s_img)), "l"(reinterpret_cast<uint64_t>(m)), "r"(x0), "r"(y0), "r"(z), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(0) : "memory");
    int bad = 0;
    for (int i = tid; i < box * box; i += 256) {
        const int r = i / box, c = i - r * box;
        if (s_img[r * box + c] != src[(long long)z * fstride + (long long)(y0 + r) * pitch + x0 + c]) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

int main(int argc, char** argv)
{
    const int w = 640, h = 480, pitch = 640, B = 3, box = 80;
    const long long fstride = (long long)pitch * h + 256;
    std::vector<uint8_t> img((size_t)fstride * B);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t* d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    Maps maps{};
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    const cuuint32_t bx[3] = {(cuuint32_t)box, (cuuint32_t)box, 1u}, es[3] = {1u, 1u, 1u};
    for (int l = 0; l < 16; ++l) {
        CUresult r = ((EncodeFn)fn)(&maps.level[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (l == 0) printf("encode: %d\n", (int)r);
    }
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(maps)); cudaMemcpy(gm, &maps, sizeof(maps), cudaMemcpyHostToDevice);
    int* dm; cudaMalloc(&dm, 4);
    // one configuration per process (a faulting launch takes the context with it): tma_probe.bin mode x0 y0 z level
    const int mode = argc > 1 ? atoi(argv[1]) : 1, x0 = argc > 2 ? atoi(argv[2]) : 16, y0 = argc > 3 ? atoi(argv[3]) : 19, z = argc > 4 ? atoi(argv[4]) : 0, level = argc > 5 ? atoi(argv[5]) : 0;
    cudaMemset(dm, 0, 4);
    if (mode == 0) k_probe<0><<<1, 256>>>(maps, gm, level, d, pitch, fstride, x0, y0, z, box, dm);
    else k_probe<1><<<1, 256>>>(maps, gm, level, d, pitch, fstride, x0, y0, z, box, dm);
    cudaError_t st = cudaDeviceSynchronize();
    int bad = -1; cudaMemcpy(&bad, dm, 4, cudaMemcpyDeviceToHost);
    printf("mode %d (%s) level %d at (%d,%d,%d): %s, mismatches %d\n", mode, mode ? "global descriptor" : "param descriptor", level, x0, y0, z, cudaGetErrorString(st), bad);
    return 0;
}
