// Library-wide helpers: error string, version, device count, pinned host memory.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
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

// PointCloudMap<PointT>::WritePLY (src/PointCloudMap.cc:325-437; USE_NORMALS = USE_POINTSURFELSEGMENT = 1, include/PointDefinitions.h:28-29): the volumetric
// map as PLVS writes it -- vertices x y z, red green blue, normal_x normal_y normal_z, label, kfid; with is_mesh a face per three consecutive vertices.  The
// binary form writes the first three bytes of PCL's rgba word under the names red, green, blue: that is b, g, r (PCL's union is {b, g, r, a}); the ASCII form
// writes r, g, b, and both put a newline where the reference puts one (`"\n" <<` in front of every record, normals and custom data on lines of their own).
// bgra: 4 bytes per point in PCL's memory order (b, g, r, a).  Host I/O only.
int plvs_map_save_ply(const char* path, const float* xyz, const uint8_t* bgra, const float* normals, const uint32_t* label, const uint32_t* kfid, long long n,
                      int is_mesh, int binary)
{
    if (!path || n < 0 || (n && (!xyz || !bgra || !normals || !label || !kfid))) { plvs::set_error("bad argument"); return PLVS_EINVAL; }
    std::fstream file;
    file.open(path, std::ios::out | std::ios::binary);
    if (!file.is_open()) { plvs::set_error("cannot open %s for writing", path); return PLVS_EINVAL; }
    const int verticesPerFace = 3;
    const size_t nfaces_idx = is_mesh ? (size_t)n : 0;
    file << "ply";
    if (binary) file << "\nformat binary_little_endian 1.0";
    else file << "\nformat ascii 1.0";
    file << "\nelement vertex " << (size_t)n;
    file << "\nproperty float32 x\nproperty float32 y\nproperty float32 z";
    file << "\nproperty uchar red\nproperty uchar green\nproperty uchar blue";
    file << "\nproperty float32 normal_x\nproperty float32 normal_y\nproperty float32 normal_z";
    file << "\nproperty uint32 label";
    file << "\nproperty uint32 kfid";
    if (nfaces_idx) {
        file << "\nelement face " << nfaces_idx / verticesPerFace;
        file << "\nproperty list uint8 int32 vertex_indices";
    }
    file << "\nend_header";
    if (binary) file << "\n";
    for (long long i = 0; i < n; ++i) {
        const float* p = xyz + 3 * i; const uint8_t* c = bgra + 4 * i; const float* nm = normals + 3 * i;
        if (binary) {
            file.write((const char*)p, 3 * sizeof(float));
            file.write((const char*)c, 3 * sizeof(uint8_t));
            file.write((const char*)nm, 3 * sizeof(float));
            file.write((const char*)&label[i], sizeof(uint32_t));
            file.write((const char*)&kfid[i], sizeof(uint32_t));
        } else {
            file << "\n" << p[0] << " " << p[1] << " " << p[2];
            file << " " << (int)c[2] << " " << (int)c[1] << " " << (int)c[0];
            file << "\n" << nm[0] << " " << nm[1] << " " << nm[2];
            file << "\n" << label[i] << " " << kfid[i];
        }
    }
    for (size_t i = 0; i + 0 < nfaces_idx; i += verticesPerFace) {
        if (binary) file.write((const char*)&verticesPerFace, sizeof(uint8_t));
        else file << "\n" << (int)verticesPerFace;
        for (int j = 0; j < verticesPerFace; ++j) {
            const unsigned int idx = (unsigned int)(i + j);
            if (binary) file.write((const char*)&idx, sizeof(unsigned int));
            else file << " " << idx;
        }
    }
    file.close();
    return PLVS_OK;
}

namespace {
struct PlyProp { std::string name, type; };
int ply_type_size(const std::string& t)
{
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "int32" || t == "uint32" || t == "float" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
double ply_read_binary(const unsigned char* p, const std::string& t)
{
    if (t == "char" || t == "int8") return (double)*(const signed char*)p;
    if (t == "uchar" || t == "uint8") return (double)*p;
    if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, p, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, p, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, p, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; std::memcpy(&v, p, 4); return v; }
    if (t == "float" || t == "float32") { float v; std::memcpy(&v, p, 4); return v; }
    double v; std::memcpy(&v, p, 8); return v;
}
}  // namespace

// The reading half of PointCloudMap<PointT>::LoadMap (src/PointCloudMap.cc:466-503: pcl::io::loadPLYFile + fromPCLPointCloud2) for the files PLVS writes
// (plvs_map_save_ply above) and for any PLY whose vertex element has scalar properties: ASCII or binary_little_endian, properties matched BY NAME --
// x y z, red green blue (-> rgb[3i..], as named in the file), normal_x normal_y normal_z, label, kfid; other properties and the face element are skipped.
// Missing properties leave their arrays untouched and clear their bit in *fields (1 xyz, 2 rgb, 4 normals, 8 label, 16 kfid).  cap = points the arrays
// hold; *n_out = points in the file (PLVS_ECAP if it does not fit, call again).  What LoadMap does next -- InvertColors (swap r and b, :505-514), then
// IntegrateWorldPointCloud -- is the caller's: plvs_tsdf_integrate_world_cloud.  Host I/O only.
int plvs_map_load_ply(const char* path, float* xyz, uint8_t* rgb, float* normals, uint32_t* label, uint32_t* kfid, long long cap, long long* n_out, int* fields)
{
    if (!path || !n_out) { plvs::set_error("bad argument"); return PLVS_EINVAL; }
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f) { plvs::set_error("cannot open dense map file: %s", path); return PLVS_EINVAL; }
    std::string line, fmt;
    long long nv = -1;
    std::vector<PlyProp> props;
    bool in_vertex = false, header_ok = false;
    if (!std::getline(f, line) || line.substr(0, 3) != "ply") { plvs::set_error("%s is not a PLY file", path); return PLVS_EINVAL; }
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string w; ls >> w;
        if (w == "format") ls >> fmt;
        else if (w == "element") { std::string e; long long cnt = 0; ls >> e >> cnt; in_vertex = e == "vertex"; if (in_vertex) nv = cnt; }
        else if (w == "property" && in_vertex) {
            std::string t, nm; ls >> t;
            if (t == "list") { plvs::set_error("list property in the vertex element is not supported"); return PLVS_EINVAL; }
            ls >> nm;
            if (!ply_type_size(t)) { plvs::set_error("unknown PLY type %s", t.c_str()); return PLVS_EINVAL; }
            props.push_back(PlyProp{nm, t});
        } else if (w == "end_header") { header_ok = true; break; }
    }
    if (!header_ok || nv < 0 || (fmt != "ascii" && fmt != "binary_little_endian")) { plvs::set_error("unsupported PLY header (format '%s')", fmt.c_str()); return PLVS_EINVAL; }
    *n_out = nv;
    int have = 0;
    auto slot = [&](const std::string& nm) -> int {      // 0-2 xyz, 3-5 rgb, 6-8 normal, 9 label, 10 kfid, -1 skip
        static const char* names[11] = {"x", "y", "z", "red", "green", "blue", "normal_x", "normal_y", "normal_z", "label", "kfid"};
        for (int i = 0; i < 11; ++i) if (nm == names[i]) return i;
        return -1;
    };
    std::vector<int> slots;
    for (const PlyProp& p : props) { const int s = slot(p.name); slots.push_back(s); if (s >= 0) have |= s < 3 ? 1 : s < 6 ? 2 : s < 9 ? 4 : s == 9 ? 8 : 16; }
    if (fields) *fields = have;
    if (nv > cap) { plvs::set_error("PLY holds %lld points, capacity %lld", nv, cap); return PLVS_ECAP; }
    auto store = [&](long long i, int s, double v) {
        if (s < 0) return;
        if (s < 3) { if (xyz) xyz[3 * i + s] = (float)v; }
        else if (s < 6) { if (rgb) rgb[3 * i + (s - 3)] = (uint8_t)v; }
        else if (s < 9) { if (normals) normals[3 * i + (s - 6)] = (float)v; }
        else if (s == 9) { if (label) label[i] = (uint32_t)v; }
        else if (kfid) kfid[i] = (uint32_t)v;
    };
    if (fmt == "ascii") {
        for (long long i = 0; i < nv; ++i)
            for (size_t k = 0; k < props.size(); ++k) {
                const std::string& t = props[k].type;
                double v;
                if (t == "float" || t == "float32" || t == "double" || t == "float64") { std::string tok; f >> tok; v = std::strtod(tok.c_str(), nullptr); if (t != "double" && t != "float64") v = (double)std::strtof(tok.c_str(), nullptr); }
                else { long long iv; f >> iv; v = (double)iv; }
                if (!f) { plvs::set_error("truncated PLY (vertex %lld)", i); return PLVS_EINVAL; }
                store(i, slots[k], v);
            }
    } else {
        size_t rec = 0;
        for (const PlyProp& p : props) rec += (size_t)ply_type_size(p.type);
        std::vector<unsigned char> buf(rec);
        for (long long i = 0; i < nv; ++i) {
            f.read((char*)buf.data(), (std::streamsize)rec);
            if (!f) { plvs::set_error("truncated PLY (vertex %lld)", i); return PLVS_EINVAL; }
            size_t off = 0;
            for (size_t k = 0; k < props.size(); ++k) { store(i, slots[k], ply_read_binary(buf.data() + off, props[k].type)); off += (size_t)ply_type_size(props[k].type); }
        }
    }
    return PLVS_OK;
}

}  // extern "C"
