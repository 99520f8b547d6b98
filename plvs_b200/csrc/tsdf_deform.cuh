// SURVEY.md §8f rank 3, the rest of the TSDF read-out / maintenance surface:
//   * ChunkManager::Deform (Thirdparty/open_chisel/src/ChunkManager.cpp:920-1062; ChiselServer::Deform, src/PointCloudMapChisel.cc:406-489): after a
//     loop closure every known voxel moves with the correction R|t of the keyframe that last wrote it; voxels that land in the same cell are folded
//     in the order the reference meets them (its chunk map's iteration order -- an input here --, then voxel id);
//   * Chisel::IntegrateWorldPointCloudWithNormals (Thirdparty/open_chisel/src/Chisel.cpp:238-379; PointCloudMapChisel::LoadMap): a saved world-frame
//     cloud is turned back into a TSDF, every point updating the voxels within 4 voxel sizes along its normal.
// Both reuse the machinery of the point-cloud route (tsdf.cu): a first kernel only RECORDS (destination voxel <- source) pairs in per-voxel linked
// lists (chunks found or created in the hash on the fly), a second one replays every voxel's list in the reference's order with the reference's
// arithmetic.  Device code only; included by tsdf.cu inside its anonymous namespace (and, through it, by the CPU execution model of the tests).
#pragma once

// ---------------------------------------------------------------------------------------------
// world cloud with normals
// ---------------------------------------------------------------------------------------------
struct WorldParams {
    float r00, r01, r02, r10, r11, r12, r20, r21, r22, tx, ty, tz;      // cameraPose (identity when a saved map is loaded)
    float res, half, rf, round_to_voxel, trunc, weight;                 // trunc = 4 * res; weight = ConstantWeighter: w / (2 * trunc)
};

struct WorldPoint { float wx, wy, wz, dx, dy, dz; };

__device__ __forceinline__ WorldPoint world_point(const WorldParams& C, const float* __restrict__ xyz, const float* __restrict__ normals, int i)
{
    const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    WorldPoint p;
    p.wx = C.r00 * px + (C.r01 * py + C.r02 * pz) + C.tx; p.wy = C.r10 * px + (C.r11 * py + C.r12 * pz) + C.ty; p.wz = C.r20 * px + (C.r21 * py + C.r22 * pz) + C.tz;
    p.dx = normals[3 * i]; p.dy = normals[3 * i + 1]; p.dz = normals[3 * i + 2];
    const float nn = p.dx * p.dx + (p.dy * p.dy + p.dz * p.dz);          // Eigen normalized(): v / sqrt(squaredNorm) when > 0, else v
    if (nn > 0.f) { const float sq = sqrtf(nn); p.dx = p.dx / sq; p.dy = p.dy / sq; p.dz = p.dz / sq; }
    return p;
}

__global__ void __launch_bounds__(256)
k_world_raycast(WorldParams C, const float* __restrict__ xyz, const float* __restrict__ normals, int n, HashEntry* tab, uint32_t mask, int* free_stack, int* free_top,
                int* block_key, uint8_t* live, int* fresh_list, int* n_fresh, int* heads, int* touched_flag, int* touched_list, int* n_touched,
                HitNode* nodes, int node_cap, int* n_nodes, int* error)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const WorldPoint p = world_point(C, xyz, normals, i);
    const float sx = p.wx * C.round_to_voxel, sy = p.wy * C.round_to_voxel, sz = p.wz * C.round_to_voxel;
    const float tx_ = (p.dx * C.trunc) * C.round_to_voxel, ty_ = (p.dy * C.trunc) * C.round_to_voxel, tz_ = (p.dz * C.trunc) * C.round_to_voxel;
    const float stx = sx - tx_, sty = sy - ty_, stz = sz - tz_, enx = sx + tx_, eny = sy + ty_, enz = sz + tz_;
    int x = (int)floorf(stx), y = (int)floorf(sty), z = (int)floorf(stz);
    const int endX = (int)floorf(enx), endY = (int)floorf(eny), endZ = (int)floorf(enz);
    const float ddx = enx - stx, ddy = eny - sty, ddz = enz - stz;
    const float maxDist = ddx * ddx + (ddy * ddy + ddz * ddz);
    const float fdx = (float)(endX - x), fdy = (float)(endY - y), fdz = (float)(endZ - z);
    const int stepX = (fdx > 0) - (fdx < 0), stepY = (fdy > 0) - (fdy < 0), stepZ = (fdz > 0) - (fdz < 0);
    if (stepX == 0 && stepY == 0 && stepZ == 0) return;
    float tMaxX = intbound_dev(stx, (int)fdx), tMaxY = intbound_dev(sty, (int)fdy), tMaxZ = intbound_dev(stz, (int)fdz);
    const float tDeltaX = ((float)stepX) / fdx, tDeltaY = ((float)stepY) / fdy, tDeltaZ = ((float)stepZ) / fdz;
    int last_block = -1, lcx = 0, lcy = 0, lcz = 0;
    for (int guard = 0; guard < 100000; ++guard) {
        {
            const float cx = (float)x * C.res + C.half, cy = (float)y * C.res + C.half, cz = (float)z * C.res + C.half;
            const int kx = (int)floorf(cx * C.rf), ky = (int)floorf(cy * C.rf), kz = (int)floorf(cz * C.rf);
            const int lx = x - kx * 16, ly = y - ky * 16, lz = z - kz * 16;
            const int id = (lz * 16 + ly) * 16 + lx;
            if (id >= 0 && id < kBlockVox) {
                const float ex = cx - p.wx, ey = cy - p.wy, ez = cz - p.wz;
                const float u = ex * p.dx + (ey * p.dy + ez * p.dz);
                if (fabsf(u) < C.trunc) {
                    int block = last_block;
                    if (block < 0 || kx != lcx || ky != lcy || kz != lcz) {
                        block = hash_find_or_create(tab, mask, kx, ky, kz, free_stack, free_top, block_key, live, fresh_list, n_fresh, error);
                        last_block = block; lcx = kx; lcy = ky; lcz = kz;
                    }
                    if (block >= 0) {
                        const int node = atomicAdd(n_nodes, 1);
                        if (node < node_cap) {
                            nodes[node].point = i;
                            nodes[node].next = atomicExch(&heads[(size_t)block * kBlockVox + id], node);
                            if (atomicExch(&touched_flag[block], 1) == 0) touched_list[atomicAdd(n_touched, 1)] = block;
                        } else atomicExch(error, 2);
                    }
                }
            }
        }
        const float ex = (float)x - stx, ey = (float)y - sty, ez = (float)z - stz;
        if (ex * ex + (ey * ey + ez * ez) > maxDist) break;
        if (x == endX && y == endY && z == endZ) break;
        if (tMaxX < tMaxY) { if (tMaxX < tMaxZ) { x += stepX; tMaxX += tDeltaX; } else { z += stepZ; tMaxZ += tDeltaZ; } }
        else { if (tMaxY < tMaxZ) { y += stepY; tMaxY += tDeltaY; } else { z += stepZ; tMaxZ += tDeltaZ; } }
    }
}

// ColorVoxel::Integrate(r, g, b, 1) (ColorVoxel.h:68-89): division form, saturated, truncated to a byte
__device__ __forceinline__ uint32_t color_integrate_div(uint32_t col, uint32_t nr, uint32_t ng, uint32_t nb)
{
    const uint32_t cw = col >> 24;
    if (cw >= 254u) return col;
    const float wsum = (float)(1u + cw);
    const uint32_t in[3] = {nr, ng, nb};
    uint32_t out = (cw + 1u) << 24;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float upd = fminf(fmaxf((float)((float)cw * (float)((col >> (8 * ch)) & 0xffu) + (float)in[ch]) / wsum, 0.0f), 255.0f);
        out |= ((uint32_t)upd & 0xffu) << (8 * ch);
    }
    return out;
}

// collect the sequence numbers of one voxel's hit list in ascending order, 32 at a time above `last` (lists are short; long ones are re-walked)
__device__ __forceinline__ int next_hits(const HitNode* __restrict__ nodes, int head, int last, int* pts)
{
    int cnt = 0;
    for (int nd = head; nd >= 0; nd = nodes[nd].next) {
        const int p = nodes[nd].point;
        if (p <= last) continue;
        if (cnt < 32) pts[cnt++] = p;
        else { int mx = 0; for (int k = 1; k < 32; ++k) if (pts[k] > pts[mx]) mx = k; if (p < pts[mx]) pts[mx] = p; }
    }
    for (int a = 1; a < cnt; ++a) { const int v = pts[a]; int j = a - 1; while (j >= 0 && pts[j] > v) { pts[j + 1] = pts[j]; --j; } pts[j + 1] = v; }
    return cnt;
}

__global__ void __launch_bounds__(256)
k_world_apply(WorldParams C, const float* __restrict__ xyz, const float* __restrict__ rgb, const float* __restrict__ normals, int use_color,
              const int* __restrict__ touched_list, const int* __restrict__ n_touched, const int* __restrict__ block_key, int* heads, int* touched_flag,
              const HitNode* __restrict__ nodes, const uint32_t* __restrict__ kfids, uint32_t kfid_all,
              float* sdf_pool, float* w_pool, uint32_t* rgba_pool, uint32_t* kfid_pool, Totals* tot, int* neg_mask)
{
    if ((int)blockIdx.x >= *n_touched) return;
    const int b = touched_list[blockIdx.x];
    const int kx = block_key[3 * b], ky = block_key[3 * b + 1], kz = block_key[3 * b + 2];
    for (int id = threadIdx.x; id < kBlockVox; id += 256) {
        int* hp = &heads[(size_t)b * kBlockVox + id];
        const int head = *hp;
        if (head < 0) continue;
        *hp = -1;
        const int vx = kx * 16 + (id & 15), vy = ky * 16 + ((id >> 4) & 15), vz = kz * 16 + (id >> 8);
        const float cx = (float)vx * C.res + C.half, cy = (float)vy * C.res + C.half, cz = (float)vz * C.res + C.half;
        const size_t o = (size_t)b * kBlockVox + id;
        float sdf = sdf_pool[o], w = w_pool[o];
        uint32_t col = rgba_pool[o], kf = kfid_pool[o];
        int pts[32], last = -1;
        for (;;) {
            const int cnt = next_hits(nodes, head, last, pts);
            if (cnt == 0) break;
            for (int a = 0; a < cnt; ++a) {
                const int i = pts[a];
                const WorldPoint p = world_point(C, xyz, normals, i);
                const float ex = cx - p.wx, ey = cy - p.wy, ez = cz - p.wz;
                const float u = ex * p.dx + (ey * p.dy + ez * p.dz);
                sdf = (w * sdf + C.weight * u) / (C.weight + w);        // DistVoxel::Integrate(u, weight)
                w = w + C.weight;
                kf = kfids ? kfids[i] : kfid_all;
                if (use_color)
                    col = color_integrate_div(col, (uint32_t)(uint8_t)((rgb ? rgb[3 * i] : 0.f) * 255.0f), (uint32_t)(uint8_t)((rgb ? rgb[3 * i + 1] : 0.f) * 255.0f),
                                              (uint32_t)(uint8_t)((rgb ? rgb[3 * i + 2] : 0.f) * 255.0f));
            }
            last = pts[cnt - 1];
            if (cnt < 32) break;
        }
        sdf_pool[o] = sdf; w_pool[o] = w; rgba_pool[o] = col; kfid_pool[o] = kf;
    }
    __syncthreads();
    if (threadIdx.x == 0) { touched_flag[b] = 0; neg_mask[b] = 0xff; atomicAdd((unsigned long long*)&tot->updated, 1ull); }   // carvable mask: conservative
}

// ---------------------------------------------------------------------------------------------
// Deform
// ---------------------------------------------------------------------------------------------
struct DeformEntry { uint32_t kfid; float T[12]; };          // sorted by kfid (unique) on the host

__device__ __forceinline__ int deform_lookup(const DeformEntry* __restrict__ e, int n, uint32_t kfid)
{
    int lo = 0, hi = n - 1;
    while (lo <= hi) { const int mid = (lo + hi) >> 1; const uint32_t k = e[mid].kfid; if (k == kfid) return mid; if (k < kfid) lo = mid + 1; else hi = mid - 1; }
    return -1;
}

// CTA = old chunk number `rank` of the visiting order, thread = 16 of its voxels.  Records (new voxel <- sequence number rank * 4096 + voxel id) in the
// NEW hash table `tab`, whose chunks are taken from the same pool (the old chunks stay readable until k_deform_release).
__global__ void __launch_bounds__(256)
k_deform_scatter(const int* __restrict__ old_list, int n_old, const int* __restrict__ old_key, const float* __restrict__ w_pool, const uint32_t* __restrict__ kfid_pool,
                 const DeformEntry* __restrict__ entries, int n_entries, float res, float half, float inv_res, float rf,
                 HashEntry* tab, uint32_t mask, int* free_stack, int* free_top, int* block_key, uint8_t* live, int* fresh_list, int* n_fresh,
                 int* heads, int* touched_flag, int* touched_list, int* n_touched, HitNode* nodes, int node_cap, int* n_nodes, int* error, int* n_dropped)
{
    const int rank = blockIdx.x;
    if (rank >= n_old) return;
    const int b = old_list[rank];
    const int kx = old_key[3 * rank], ky = old_key[3 * rank + 1], kz = old_key[3 * rank + 2];      // the key table of the pool is overwritten by new chunks: own copy
    const float ox = (float)(16 * kx) * res, oy = (float)(16 * ky) * res, oz = (float)(16 * kz) * res;
    for (int id = threadIdx.x; id < kBlockVox; id += 256) {
        const size_t o = (size_t)b * kBlockVox + id;
        if ((double)w_pool[o] <= 1e-15) continue;
        const int e = deform_lookup(entries, n_entries, kfid_pool[o]);
        if (e < 0) { atomicAdd(n_dropped, 1); continue; }
        const float* T = entries[e].T;
        const float px = ((float)(id & 15) * res + half) + ox, py = ((float)((id >> 4) & 15) * res + half) + oy, pz = ((float)(id >> 8) * res + half) + oz;
        const float nx = T[0] * px + (T[1] * py + T[2] * pz) + T[3], ny = T[4] * px + (T[5] * py + T[6] * pz) + T[7], nz = T[8] * px + (T[9] * py + T[10] * pz) + T[11];
        const int cx = (int)floorf(nx * rf), cy = (int)floorf(ny * rf), cz = (int)floorf(nz * rf);
        const int lx = (int)floorf(nx * inv_res) - 16 * cx, ly = (int)floorf(ny * inv_res) - 16 * cy, lz = (int)floorf(nz * inv_res) - 16 * cz;
        if (lx < 0 || lx > 15 || ly < 0 || ly > 15 || lz < 0 || lz > 15) { atomicAdd(n_dropped, 1); continue; }     // the reference indexes outside the chunk here
        const int block = hash_find_or_create(tab, mask, cx, cy, cz, free_stack, free_top, block_key, live, fresh_list, n_fresh, error);
        if (block < 0) continue;
        const int node = atomicAdd(n_nodes, 1);
        if (node >= node_cap) { atomicExch(error, 2); continue; }
        nodes[node].point = rank * kBlockVox + id;
        nodes[node].next = atomicExch(&heads[(size_t)block * kBlockVox + ((lz * 16 + ly) * 16 + lx)], node);
        if (atomicExch(&touched_flag[block], 1) == 0) touched_list[atomicAdd(n_touched, 1)] = block;
    }
}

__global__ void __launch_bounds__(256)
k_deform_apply(const int* __restrict__ old_list, int use_color, const int* __restrict__ touched_list, const int* __restrict__ n_touched, int* heads, int* touched_flag,
               const HitNode* __restrict__ nodes, float* sdf_pool, float* w_pool, uint32_t* rgba_pool, uint32_t* kfid_pool, int* neg_mask)
{
    if ((int)blockIdx.x >= *n_touched) return;
    const int b = touched_list[blockIdx.x];
    for (int id = threadIdx.x; id < kBlockVox; id += 256) {
        int* hp = &heads[(size_t)b * kBlockVox + id];
        const int head = *hp;
        if (head < 0) continue;
        *hp = -1;
        const size_t o = (size_t)b * kBlockVox + id;
        float sdf = sdf_pool[o], w = w_pool[o];            // a fresh chunk: 99999 / 0
        uint32_t col = rgba_pool[o], kf = kfid_pool[o];
        int pts[32], last = -1;
        for (;;) {
            const int cnt = next_hits(nodes, head, last, pts);
            if (cnt == 0) break;
            for (int a = 0; a < cnt; ++a) {
                const size_t so = (size_t)old_list[pts[a] >> 12] * kBlockVox + (pts[a] & (kBlockVox - 1));
                const float ssdf = sdf_pool[so], sw = w_pool[so];
                const uint32_t scol = rgba_pool[so];
                if ((double)w <= 1e-15) { sdf = ssdf; w = sw; col = scol; }                        // newDistVoxel = distVoxel; newColorVoxel = colorVoxel
                else {
                    sdf = (w * sdf + sw * ssdf) / (sw + w);                                        // DistVoxel::Integrate(sdf, weight)
                    w = w + sw;
                    if (use_color) col = color_integrate_div(col, scol & 0xffu, (scol >> 8) & 0xffu, (scol >> 16) & 0xffu);
                }
                kf = kfid_pool[so];
            }
            last = pts[cnt - 1];
            if (cnt < 32) break;
        }
        sdf_pool[o] = sdf; w_pool[o] = w; rgba_pool[o] = col; kfid_pool[o] = kf;
    }
    __syncthreads();
    if (threadIdx.x == 0) { touched_flag[b] = 0; neg_mask[b] = 0xff; }
}

// give the chunks in `list` back to the pool (the old map after a deformation, or the new one when it has to be rolled back)
__global__ void __launch_bounds__(256)
k_release_blocks(const int* __restrict__ list, const int* __restrict__ n_list_dev, int n_list_host, int* free_stack, int* free_top, uint8_t* live, int* neg_mask)
{
    const int n = n_list_dev ? *n_list_dev : n_list_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = list[i];
    live[b] = 0; neg_mask[b] = 0;
    free_stack[atomicAdd(free_top, 1)] = b;
}

// fresh chunks of a deformation: voxel state and keyframe ids
__global__ void __launch_bounds__(256)
k_init_fresh_kfid(const int* __restrict__ fresh_list, const int* __restrict__ n_fresh, uint32_t* kfid_pool)
{
    if ((int)blockIdx.x >= *n_fresh) return;
    const int b = fresh_list[blockIdx.x];
    for (int i = threadIdx.x; i < kBlockVox; i += 256) kfid_pool[(size_t)b * kBlockVox + i] = 0u;
}

// the meshes of the last UpdateMesh move with their vertices' keyframes (ChunkManager.cpp:1022-1053): vertex = R * vertex + t, normal = R * normal
__global__ void __launch_bounds__(256)
k_deform_mesh(float* verts, float* normals, const uint32_t* __restrict__ vkfid, long long nverts, const DeformEntry* __restrict__ entries, int n_entries)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nverts) return;
    const int e = deform_lookup(entries, n_entries, vkfid[i]);
    if (e < 0) return;
    const float* T = entries[e].T;
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    verts[3 * i] = T[0] * x + (T[1] * y + T[2] * z) + T[3]; verts[3 * i + 1] = T[4] * x + (T[5] * y + T[6] * z) + T[7]; verts[3 * i + 2] = T[8] * x + (T[9] * y + T[10] * z) + T[11];
    const float a = normals[3 * i], b = normals[3 * i + 1], c = normals[3 * i + 2];
    normals[3 * i] = T[0] * a + (T[1] * b + T[2] * c); normals[3 * i + 1] = T[4] * a + (T[5] * b + T[6] * c); normals[3 * i + 2] = T[8] * a + (T[9] * b + T[10] * c);
}
