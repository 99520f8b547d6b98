// Library-wide helpers: error string, version, device count, pinned host memory.
#include <cstdarg>
#include <fstream>
#include "common.cuh"

namespace plvs {
int g_profiling = 0;
std::atomic<long long> g_io_bytes[2];
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace plvs

extern "C" {

const char* plvs_version(void) { return "plvs_b200 0.1 (sm_100a)"; }
const char* plvs_last_error(void) { return plvs::g_err; }

int plvs_set_profiling(int mask) { plvs::g_profiling = mask; return PLVS_OK; }

int plvs_io_bytes(long long* h2d, long long* d2h, int reset)
{
    if (h2d) *h2d = plvs::g_io_bytes[0].load();
    if (d2h) *d2h = plvs::g_io_bytes[1].load();
    if (reset) { plvs::g_io_bytes[0] = 0; plvs::g_io_bytes[1] = 0; }
    return PLVS_OK;
}

int plvs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int plvs_enable_peer_access(int device, int peer)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || peer < 0 || device >= n || peer >= n || device == peer) { plvs::set_error("bad device pair %d -> %d", device, peer); return PLVS_EINVAL; }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device, peer) != cudaSuccess || !can) { plvs::set_error("device %d cannot map the memory of device %d", device, peer); return PLVS_ENODEV; }
    int prev = 0; cudaGetDevice(&prev);
    cudaSetDevice(device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    cudaSetDevice(prev);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { plvs::set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", device, peer, cudaGetErrorString(e)); return PLVS_ENODEV; }
    cudaGetLastError();
    return PLVS_OK;
}

int plvs_host_alloc(void** p, size_t bytes)
{
    if (!p) return PLVS_EINVAL;
    if (cudaHostAlloc(p, bytes, cudaHostAllocDefault) != cudaSuccess) { plvs::set_error("cudaHostAlloc(%zu) failed", bytes); return PLVS_ENOMEM; }
    return PLVS_OK;
}

int plvs_host_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) return PLVS_ENODEV;
    return PLVS_OK;
}

// Chisel::SaveAllMeshesToPLY + SaveMeshPLYASCII (Thirdparty/open_chisel/src/Chisel.cpp:79-118, src/io/PLY.cpp:29-86): one ASCII PLY of all mesh
// vertices (default ostream float formatting, colours as static_cast<int>(c * 255.0f)), faces = consecutive vertex triples.  Host I/O only.
int plvs_mesh_save_ply(const char* path, const float* verts, const float* colors, long long n_verts)
{
    if (!path || n_verts < 0 || (n_verts && !verts)) { plvs::set_error("bad argument"); return PLVS_EINVAL; }
    std::ofstream stream(path);
    if (!stream) { plvs::set_error("cannot open %s", path); return PLVS_EINVAL; }
    stream << "ply" << std::endl;
    stream << "format ascii 1.0" << std::endl;
    stream << "element vertex " << (size_t)n_verts << std::endl;
    stream << "property float x" << std::endl;
    stream << "property float y" << std::endl;
    stream << "property float z" << std::endl;
    if (colors) {
        stream << "property uchar red" << std::endl;
        stream << "property uchar green" << std::endl;
        stream << "property uchar blue" << std::endl;
    }
    stream << "element face " << (size_t)n_verts / 3 << std::endl;
    stream << "property list uchar int vertex_index" << std::endl;
    stream << "end_header" << std::endl;
    for (long long i = 0; i < n_verts; ++i) {
        stream << verts[3 * i] << " " << verts[3 * i + 1] << " " << verts[3 * i + 2];
        if (colors) stream << " " << static_cast<int>(colors[3 * i] * 255.0f) << " " << static_cast<int>(colors[3 * i + 1] * 255.0f) << " " << static_cast<int>(colors[3 * i + 2] * 255.0f);
        stream << std::endl;
    }
    for (long long i = 0; i + 2 < n_verts; i += 3) stream << "3 " << (size_t)i << " " << (size_t)(i + 1) << " " << (size_t)(i + 2) << " " << std::endl;
    return stream ? PLVS_OK : PLVS_EINVAL;
}

}  // extern "C"
