// Device code only: the voxel-block hash (ChunkHasher / ChunkMap, Thirdparty/open_chisel/include/open_chisel/ChunkManager.h:42-56) as an
// open-addressing table, look-up side.  tsdf.cu includes it inside its anonymous namespace; tests/native/emu_kernels.cpp compiles the same text
// for the CPU (tests/native/cuda_emu.hpp).
#pragma once

constexpr int kBlockVox = 4096;
constexpr int HASH_EMPTY = -1, HASH_LOCKED = -2;

struct HashEntry { int x, y, z, idx; };

__device__ __forceinline__ uint32_t hash_key(int x, int y, int z, uint32_t mask)
{
    return (((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)z * 83492791u)) & mask;   // ChunkHasher
}

__device__ int hash_find(const HashEntry* __restrict__ tab, uint32_t mask, int x, int y, int z)
{
    uint32_t s = hash_key(x, y, z, mask);
    for (uint32_t probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
        const int idx = tab[s].idx;
        if (idx == HASH_EMPTY) return -1;
        if (idx >= 0 && tab[s].x == x && tab[s].y == y && tab[s].z == z) return idx;
    }
    return -1;
}
