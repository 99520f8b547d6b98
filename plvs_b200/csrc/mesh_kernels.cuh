// Device code only (needs tsdf_hash.cuh first): mesh read-out kernels.  tsdf.cu includes it inside its anonymous namespace;
// tests/native/emu_kernels.cpp compiles the same text for the CPU (tests/native/cuda_emu.hpp).
#pragma once

// ---------------------------------------------------------------------------------------------
// Read-out: ChunkManager::RecomputeMesh for every live block -- GenerateMesh (Thirdparty/open_chisel/src/
// ChunkManager.cpp:577-664), MarchingCubes::MeshCube (include/open_chisel/marching_cubes/MarchingCubes.h:76-271), ColorizeMesh /
// InterpolateColor (:715-806), ComputeNormalsFromGradients (:838-856).  The reference appends triangles voxel by voxel in a fixed order
// (inside voxels z,y,x; then the max-X, max-Y, max-Z planes, whose cubes reach into the +x/+y/+z neighbour blocks): `rank` below is
// a voxel's position in that order, so counting per voxel + a prefix sum reproduces the vertex order exactly.
//   k_mesh_count : CTA per block (17^3 corner lattice staged in shared memory), thread = 16 consecutive ranks -> triangles per block
//   k_mesh_emit  : same walk, in-block prefix sum, writes positions + face normals at the block's base offset
//   k_mesh_shade : thread per vertex -> colour (InterpolateColor as written, including its look-ups by voxel index) and the
//                  normal from the SDF gradient (kept in double where the reference is)
// HBM-bound: 8 B per voxel read (coalesced rows into shared memory; the 8-corner reuse is on chip) + 36 B per vertex written.
// ---------------------------------------------------------------------------------------------
__device__ const uint64_t d_tri_table[256] = {
#include "mc_tables.inc"
};
__device__ const int8_t d_edge_pairs[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

struct MeshParams { float res, inv, half, rounding; int use_color; };

__device__ __forceinline__ void mesh_rank_to_voxel(int r, int& x, int& y, int& z)
{
    if (r < 3375) { x = r % 15; y = (r / 15) % 15; z = r / 225; }
    else if (r < 3615) { r -= 3375; x = 15; y = r % 16; z = r / 16; }
    else if (r < 3840) { r -= 3615; y = 15; x = r % 15; z = r / 15; }
    else { r -= 3840; z = 15; x = r % 16; y = r / 16; }
}

// The 17^3 corner lattice of a block -- its own 16^3 voxels plus the first layer of the +x/+y/+z neighbours -- staged in shared memory: the block's
// own voxels arrive as coalesced rows, every cube then reads its 8 corners on chip.  s_ok = 0 marks a corner that is unobserved (weight <= 1e-15,
// ChunkManager.cpp:336,366) or whose block does not exist; nb[k] = pool index of the neighbour at offset (k&1, k>>1&1, k>>2&1), -1 if absent.
constexpr int kLat = 17, kLatN = kLat * kLat * kLat;

__device__ __forceinline__ void mesh_stage_lattice(const float* __restrict__ sdf_pool, const float* __restrict__ w_pool, const int* nb, float* s_sdf, uint8_t* s_ok)
{
    for (int i = threadIdx.x; i < kLatN; i += blockDim.x) {
        int x = i % kLat, y = (i / kLat) % kLat, z = i / (kLat * kLat), k = 0;
        if (x == 16) { k |= 1; x = 0; }
        if (y == 16) { k |= 2; y = 0; }
        if (z == 16) { k |= 4; z = 0; }
        const int blk = nb[k];
        float s = 0.f; uint8_t ok = 0;
        if (blk >= 0) {
            const size_t id = (size_t)blk * kBlockVox + ((z * 16 + y) * 16 + x);
            ok = !((double)w_pool[id] <= 1e-15);
            s = sdf_pool[id];
        }
        s_sdf[i] = s; s_ok[i] = ok;
    }
}

// corner SDFs of the cube at voxel (ix,iy,iz); returns the configuration, 0 when a corner is unobserved / its block is missing
__device__ __forceinline__ int mesh_cube(const float* s_sdf, const uint8_t* s_ok, int ix, int iy, int iz, float* sdf)
{
    int cfg = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ox = (i == 1 || i == 2 || i == 5 || i == 6), oy = (i == 2 || i == 3 || i == 6 || i == 7), oz = i >> 2;     // cubeIndexOffsets (:84-86)
        const int id = ((iz + oz) * kLat + (iy + oy)) * kLat + (ix + ox);
        if (!s_ok[id]) return 0;
        const float s = s_sdf[id];
        sdf[i] = s;
        if (s < 0) cfg |= 1 << i;
    }
    return cfg;
}

__global__ void __launch_bounds__(256)
k_mesh_count(const int* __restrict__ list, const int* __restrict__ block_key, const HashEntry* __restrict__ tab, uint32_t mask,
             const float* __restrict__ sdf_pool, const float* __restrict__ w_pool, int* __restrict__ tri_count)
{
    __shared__ int s_nb[8];
    __shared__ int s_sum[8];
    __shared__ float s_sdf[kLatN];
    __shared__ uint8_t s_ok[kLatN];
    const int b = list[blockIdx.x];
    if (threadIdx.x < 8) {
        const int k = threadIdx.x;
        s_nb[k] = k == 0 ? b : hash_find(tab, mask, block_key[3 * b] + (k & 1), block_key[3 * b + 1] + ((k >> 1) & 1), block_key[3 * b + 2] + (k >> 2));
    }
    __syncthreads();
    mesh_stage_lattice(sdf_pool, w_pool, s_nb, s_sdf, s_ok);
    __syncthreads();
    int n = 0;
    for (int j = 0; j < 16; ++j) {
        int x, y, z; float sdf[8];
        mesh_rank_to_voxel(threadIdx.x * 16 + j, x, y, z);
        const int cfg = mesh_cube(s_sdf, s_ok, x, y, z, sdf);
        n += (int)(d_tri_table[cfg] >> 60);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 8; ++i) t += s_sum[i]; tri_count[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(256)
k_mesh_emit(const int* __restrict__ list, const int* __restrict__ block_key, const HashEntry* __restrict__ tab, uint32_t mask,
            const float* __restrict__ sdf_pool, const float* __restrict__ w_pool, const long long* __restrict__ vert_base, MeshParams M,
            float* __restrict__ verts, float* __restrict__ normals, const uint32_t* __restrict__ kfid_pool, uint32_t* __restrict__ vert_kfid)
{
    __shared__ int s_nb[8];
    __shared__ int s_warp[8];
    __shared__ float s_sdf[kLatN];
    __shared__ uint8_t s_ok[kLatN];
    const int b = list[blockIdx.x];
    if (threadIdx.x < 8) {
        const int k = threadIdx.x;
        s_nb[k] = k == 0 ? b : hash_find(tab, mask, block_key[3 * b] + (k & 1), block_key[3 * b + 1] + ((k >> 1) & 1), block_key[3 * b + 2] + (k >> 2));
    }
    __syncthreads();
    mesh_stage_lattice(sdf_pool, w_pool, s_nb, s_sdf, s_ok);
    __syncthreads();
    // pass 1: triangles of this thread's 16 voxels, exclusive prefix over the CTA (thread order == rank order)
    int n = 0;
    for (int j = 0; j < 16; ++j) {
        int x, y, z; float sdf[8];
        mesh_rank_to_voxel(threadIdx.x * 16 + j, x, y, z);
        n += (int)(d_tri_table[mesh_cube(s_sdf, s_ok, x, y, z, sdf)] >> 60);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = incl - n;
    for (int i = 0; i < wid; ++i) before += s_warp[i];
    if (n == 0) return;
    // pass 2: emit
    const float ox = (float)(16 * block_key[3 * b]) * M.res, oy = (float)(16 * block_key[3 * b + 1]) * M.res, oz = (float)(16 * block_key[3 * b + 2]) * M.res;
    long long v = vert_base[blockIdx.x] + 3ll * before;
    for (int j = 0; j < 16; ++j) {
        int x, y, z; float sdf[8];
        mesh_rank_to_voxel(threadIdx.x * 16 + j, x, y, z);
        const int cfg = mesh_cube(s_sdf, s_ok, x, y, z, sdf);
        const uint64_t row = d_tri_table[cfg];
        const int ntri = (int)(row >> 60);
        if (ntri == 0) continue;
        // centroids[i] + chunk->GetOrigin(), then + cubeCoordOffsets (:65-87, :462)
        const float cx = ((float)x * M.res + M.half) + ox, cy = ((float)y * M.res + M.half) + oy, cz = ((float)z * M.res + M.half) + oz;
        float ex[12], ey[12], ez[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            const int a = d_edge_pairs[e][0], c = d_edge_pairs[e][1];
            const float sa = sdf[a], sc = sdf[c];
            if ((sa < 0 && sc >= 0) || (sa >= 0 && sc < 0)) {
                const float ax = cx + (float)(a == 1 || a == 2 || a == 5 || a == 6) * M.res, ay = cy + (float)(a == 2 || a == 3 || a == 6 || a == 7) * M.res, az = cz + (float)(a >> 2) * M.res;
                const float bx = cx + (float)(c == 1 || c == 2 || c == 5 || c == 6) * M.res, by = cy + (float)(c == 2 || c == 3 || c == 6 || c == 7) * M.res, bz = cz + (float)(c >> 2) * M.res;
                const float diff = sa - sc;
                if (fabsf(diff) < 1e-6f) { ex[e] = ax + bx * 0.5f; ey[e] = ay + by * 0.5f; ez[e] = az + bz * 0.5f; }       // InterpolateVertex as written (:249-252)
                else { const float t = sa / diff; ex[e] = ax + (bx - ax) * t; ey[e] = ay + (by - ay) * t; ez[e] = az + (bz - az) * t; }
            }
        }
        // Mesh::kfids: every vertex of the cube carries the keyframe id of the cube's first corner, i.e. of this voxel (ChunkManager.cpp:464-468)
        const uint32_t cube_kfid = (vert_kfid && kfid_pool) ? kfid_pool[(size_t)b * kBlockVox + ((z * 16 + y) * 16 + x)] : 0u;
        for (int t = 0; t < ntri; ++t) {
            const int e0 = (int)(row >> (4 * (3 * t + 2))) & 15, e1 = (int)(row >> (4 * (3 * t + 1))) & 15, e2 = (int)(row >> (4 * (3 * t))) & 15;
            const float p0x = ex[e0], p0y = ey[e0], p0z = ez[e0], p1x = ex[e1], p1y = ey[e1], p1z = ez[e1], p2x = ex[e2], p2y = ey[e2], p2z = ez[e2];
            const float ux = p1x - p0x, uy = p1y - p0y, uz = p1z - p0z, wx = p2x - p0x, wy = p2y - p0y, wz = p2z - p0z;
            float nx = uy * wz - uz * wy, ny = uz * wx - ux * wz, nz = ux * wy - uy * wx;
            const float sq = nx * nx + (ny * ny + nz * nz);
            if (sq > 0.f) { const float s = sqrtf(sq); nx = nx / s; ny = ny / s; nz = nz / s; }
            float* V = verts + 3 * v; float* N = normals + 3 * v;
            V[0] = p0x; V[1] = p0y; V[2] = p0z; V[3] = p1x; V[4] = p1y; V[5] = p1z; V[6] = p2x; V[7] = p2y; V[8] = p2z;
            N[0] = nx; N[1] = ny; N[2] = nz; N[3] = nx; N[4] = ny; N[5] = nz; N[6] = nx; N[7] = ny; N[8] = nz;
            if (vert_kfid) { vert_kfid[v] = cube_kfid; vert_kfid[v + 1] = cube_kfid; vert_kfid[v + 2] = cube_kfid; }
            v += 3;
        }
    }
}

// ChunkManager::GetIDAt + FindChunk + chunk->GetVoxelID(pos - origin): the voxel a metric position falls into, -1 when its block is
// absent or the id leaves [0, 4096) (the coordinates themselves are not range-checked, as in the reference)
__device__ __forceinline__ long long mesh_voxel_at(const HashEntry* __restrict__ tab, uint32_t mask, const MeshParams& M, float px, float py, float pz)
{
    const int kx = (int)floorf(px * M.rounding), ky = (int)floorf(py * M.rounding), kz = (int)floorf(pz * M.rounding);
    const int blk = hash_find(tab, mask, kx, ky, kz);
    if (blk < 0) return -1;
    const float rx = px - (float)(16 * kx) * M.res, ry = py - (float)(16 * ky) * M.res, rz = pz - (float)(16 * kz) * M.res;
    const long long x = (int)floorf(rx * M.inv), y = (int)floorf(ry * M.inv), z = (int)floorf(rz * M.inv);
    const long long id = (z * 16 + y) * 16 + x;
    return (id >= 0 && id < kBlockVox) ? (long long)blk * kBlockVox + id : -1;
}

__device__ __forceinline__ bool mesh_sdf_at(const HashEntry* __restrict__ tab, uint32_t mask, const MeshParams& M, const float* __restrict__ sdf_pool,
                                            const float* __restrict__ w_pool, float px, float py, float pz, double* dist)
{
    const long long id = mesh_voxel_at(tab, mask, M, px, py, pz);
    if (id < 0 || !((double)w_pool[id] > 1e-12)) return false;
    *dist = (double)sdf_pool[id];
    return true;
}

__global__ void __launch_bounds__(256)
k_mesh_shade(long long nv, const HashEntry* __restrict__ tab, uint32_t mask, const int* __restrict__ block_key, const float* __restrict__ sdf_pool,
             const float* __restrict__ w_pool, const uint32_t* __restrict__ rgba_pool, MeshParams M, const float* __restrict__ verts, float* __restrict__ normals,
             float* __restrict__ colors)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= nv) return;
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    float cr = 0.f, cg = 0.f, cb = 0.f;
    if (M.use_color) {
        const int x_0 = (int)floorf(x * M.inv), y_0 = (int)floorf(y * M.inv), z_0 = (int)floorf(z * M.inv);
        const int x_1 = x_0 + 1, y_1 = y_0 + 1, z_1 = z_0 + 1;
        // GetColorVoxel(Vec3(x_0, y_0, z_0)): voxel INDICES handed over as metric positions (:734-741), restated as written
        const long long v000 = mesh_voxel_at(tab, mask, M, (float)x_0, (float)y_0, (float)z_0), v001 = mesh_voxel_at(tab, mask, M, (float)x_0, (float)y_0, (float)z_1);
        const long long v011 = mesh_voxel_at(tab, mask, M, (float)x_0, (float)y_1, (float)z_1), v111 = mesh_voxel_at(tab, mask, M, (float)x_1, (float)y_1, (float)z_1);
        const long long v110 = mesh_voxel_at(tab, mask, M, (float)x_1, (float)y_1, (float)z_0), v100 = mesh_voxel_at(tab, mask, M, (float)x_1, (float)y_0, (float)z_0);
        const long long v010 = mesh_voxel_at(tab, mask, M, (float)x_0, (float)y_1, (float)z_0), v101 = mesh_voxel_at(tab, mask, M, (float)x_1, (float)y_0, (float)z_1);
        if (v000 < 0 || v001 < 0 || v011 < 0 || v111 < 0 || v110 < 0 || v100 < 0 || v010 < 0 || v101 < 0) {
            const int kx = (int)floorf(x * M.rounding), ky = (int)floorf(y * M.rounding), kz = (int)floorf(z * M.rounding);
            const int blk = hash_find(tab, mask, kx, ky, kz);
            if (blk >= 0) {         // Chunk::GetColorAt (src/Chunk.cpp:136-155)
                const float ox = (float)(16 * kx) * M.res, oy = (float)(16 * ky) * M.res, oz = (float)(16 * kz) * M.res;
                const float size = (float)16 * M.res;
                if (x >= ox && y >= oy && z >= oz && x <= ox + size && y <= oy + size && z <= oz + size) {
                    const int vx = (int)((x - ox) * M.inv), vy = (int)((y - oy) * M.inv), vz = (int)((z - oz) * M.inv);
                    if (vx >= 0 && vx < 16 && vy >= 0 && vy < 16 && vz >= 0 && vz < 16) {
                        const uint32_t c = rgba_pool[(size_t)blk * kBlockVox + ((vz * 16 + vy) * 16 + vx)];
                        const float invMax = 1.f / 255.f;
                        cr = (float)(c & 0xffu) * invMax; cg = (float)((c >> 8) & 0xffu) * invMax; cb = (float)((c >> 16) & 0xffu) * invMax;
                    }
                }
            }
        } else {
            const float xd = (x - (float)x_0) / (float)(x_1 - x_0), yd = (y - (float)y_0) / (float)(y_1 - y_0), zd = (z - (float)z_0) / (float)(z_1 - z_0);
            const uint32_t c000 = rgba_pool[v000], c001 = rgba_pool[v001], c011 = rgba_pool[v011], c111 = rgba_pool[v111];
            const uint32_t c110 = rgba_pool[v110], c100 = rgba_pool[v100], c010 = rgba_pool[v010], c101 = rgba_pool[v101];
            float out[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int sh = 8 * ch;
                const float c_00 = (float)((c000 >> sh) & 0xffu) * (1.f - xd) + (float)((c100 >> sh) & 0xffu) * xd;
                const float c_10 = (float)((c010 >> sh) & 0xffu) * (1.f - xd) + (float)((c110 >> sh) & 0xffu) * xd;
                const float c_01 = (float)((c001 >> sh) & 0xffu) * (1.f - xd) + (float)((c101 >> sh) & 0xffu) * xd;
                const float c_11 = (float)((c011 >> sh) & 0xffu) * (1.f - xd) + (float)((c111 >> sh) & 0xffu) * xd;
                const float c_0 = c_00 * (1.f - yd) + c_10 * yd;
                const float c_1 = c_01 * (1.f - yd) + c_11 * yd;
                out[ch] = (c_0 * (1.f - zd) + c_1 * zd) / 255.0f;
            }
            cr = out[0]; cg = out[1]; cb = out[2];
        }
    }
    colors[3 * i] = cr; colors[3 * i + 1] = cg; colors[3 * i + 2] = cb;
    // GetSDFAndGradient (:666-690) + ComputeNormalsFromGradients (:838-856): the face normal stays when a neighbour is unobserved
    const float fx = floorf(x * M.inv) * M.res + M.half, fy = floorf(y * M.inv) * M.res + M.half, fz = floorf(z * M.inv) * M.res + M.half;
    double d0, xp, yp, zp, xm, ym, zm;
    if (!mesh_sdf_at(tab, mask, M, sdf_pool, w_pool, fx, fy, fz, &d0)) return;
    if (!mesh_sdf_at(tab, mask, M, sdf_pool, w_pool, fx + M.res, fy + 0.f, fz + 0.f, &xp)) return;
    if (!mesh_sdf_at(tab, mask, M, sdf_pool, w_pool, fx + 0.f, fy + M.res, fz + 0.f, &yp)) return;
    if (!mesh_sdf_at(tab, mask, M, sdf_pool, w_pool, fx + 0.f, fy + 0.f, fz + M.res, &zp)) return;
    if (!mesh_sdf_at(tab, mask, M, sdf_pool, w_pool, fx - M.res, fy - 0.f, fz - 0.f, &xm)) return;
    if (!mesh_sdf_at(tab, mask, M, sdf_pool, w_pool, fx - 0.f, fy - M.res, fz - 0.f, &ym)) return;
    if (!mesh_sdf_at(tab, mask, M, sdf_pool, w_pool, fx - 0.f, fy - 0.f, fz - M.res, &zm)) return;
    float gx = (float)(xp - xm), gy = (float)(yp - ym), gz = (float)(zp - zm);
    const float sq = gx * gx + (gy * gy + gz * gz);
    if (sq > 0.f) { const float s = sqrtf(sq); gx = gx / s; gy = gy / s; gz = gz / s; }
    const float mag = sqrtf(gx * gx + (gy * gy + gz * gz));
    if ((double)mag > 1e-12) { const float r = 1.0f / mag; normals[3 * i] = gx * r; normals[3 * i + 1] = gy * r; normals[3 * i + 2] = gz * r; }
}
