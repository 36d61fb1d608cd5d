// roi_align_tiled.cuh -- shared definitions of the feature-map-stationary ("tiled") RoIAlign kernels:
// tile geometry, the per-RoI sample tables and the per-tile work lists that the prepass kernel
// builds once per call (channel independent) and the main kernels consume.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kTX = 32;                 // tile width in cells (one warp-wide row access)
constexpr int kCG = 32;                 // channels per work item
constexpr int kCellWords = kCG + 4;     // 36 words = 144 B per cell: the pad makes the transposing 128-bit shared accesses conflict-free
#ifndef B200_FWD_WARPS
#define B200_FWD_WARPS 16               // warps per forward CTA: 16 (two CTAs per SM) or 8 (four CTAs per SM, smaller tiles)
#endif
constexpr int kWarps = B200_FWD_WARPS;
constexpr int kTiledThreads = 32 * kWarps;      // 32 resident warps per SM either way (<= 64 registers / thread)
constexpr int kTiledCtasPerSM = 32 / kWarps;
static_assert(kWarps == 8 || kWarps == 16, "the staging loop deals the channel quads to 8 warps per row parity");
constexpr int kAxisMax = 32;            // P * sampling_ratio per axis supported by the tiled paths
constexpr int kStageBins = 8;           // bins staged per warp between compute and global memory
constexpr int kStageWords = kStageBins + 1;

struct __align__(16) AxisEntry {        // one bilinear sample along one axis (channel independent)
    int   low;                          // low cell (clamped into the map even when invalid)
    int   valid;                        // 0 <=> the reference's "outside the map" early-out
    float l, h;                         // weights of the high / low cell
};

struct __align__(16) RoiHeader {
    int batch;                          // -1 if the batch index is out of range
    int y_min, y_max, x_min, x_max;     // range of `low` over the samples of each axis
    int pad0, pad1, pad2;
};

struct TiledPlan {
    int ny, nx;                         // samples per axis = P * sr
    int core_h, core_w;                 // tile core (= tile minus the 1-cell halo / ring)
    int tile_h;                         // rows held per tile
    int tiles_y, tiles_x, tiles_total;  // per image / times N
    size_t smem_bytes;
    // workspace sections (byte offsets)
    size_t hdr_off, ytab_off, xtab_off;
    size_t zero_off, zero_bytes;        // block that must be zero at kernel start: [work counter][3 spare][tile counts]
    int groups_max;                     // 8-bin groups one RoI can contribute to one tile = ceil(PH*PW / 8)
    size_t tile_list_off;               // uint32 [tiles_total][R * groups_max]: RoI index | (8-bin group << 16)
    size_t ws_bytes;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// shared memory of a work item: [tile][per-warp axis tables (ny + nx entries)][per-warp staging][misc]
__host__ __device__ inline size_t tiled_smem_bytes(int tile_h, int ny, int nx) {
    return (size_t)kTX * tile_h * kCellWords * 4 + (size_t)kWarps * (ny + nx) * 16 +
           (size_t)kWarps * kCG * kStageWords * 4 + 64;
}

// Tile geometry + workspace layout.
static inline bool roi_align_tiled_plan(int N, int R, int H, int W, int C, int PH, int PW, int sr, TiledPlan* p) {
    if (sr < 1 || sr > 4 || PH * sr > kAxisMax || PW * sr > kAxisMax) return false;
    if (R <= 0 || R > 65535 || C <= 0 || N <= 0 || H <= 0 || W <= 0) return false;
    if ((long long)R * C * PH * PW >= (1LL << 31) || (long long)N * C * H * W >= (1LL << 31)) return false;
    p->ny = PH * sr; p->nx = PW * sr;
    const size_t budget = (228 * 1024 - kTiledCtasPerSM * 1024) / kTiledCtasPerSM;      // resident work items share the SM's 228 KB
    int th = 64;
    while (th > 4 && tiled_smem_bytes(th, p->ny, p->nx) > budget) --th;
    if (th <= 4) return false;
    int core_h = th - 1;
    int tiles_y = (H + core_h - 1) / core_h;
    core_h = (H + tiles_y - 1) / tiles_y;            // balance the rows over the tiles
    p->core_h = core_h; p->tile_h = core_h + 1; p->tiles_y = tiles_y;
    p->core_w = kTX - 1; p->tiles_x = (W + p->core_w - 1) / p->core_w;
    if ((long long)N * tiles_y * p->tiles_x > (1 << 20)) return false;
    p->tiles_total = N * tiles_y * p->tiles_x;
    p->smem_bytes = tiled_smem_bytes(p->tile_h, p->ny, p->nx);
    size_t off = 0;
    p->hdr_off = off;  off = align_up(off + (size_t)R * sizeof(RoiHeader), 256);
    p->ytab_off = off; off = align_up(off + (size_t)R * p->ny * sizeof(AxisEntry), 256);
    p->xtab_off = off; off = align_up(off + (size_t)R * p->nx * sizeof(AxisEntry), 256);
    p->zero_off = off;
    p->zero_bytes = align_up(sizeof(int) * (size_t)(4 + p->tiles_total), 256);
    off += p->zero_bytes;
    p->groups_max = (PH * PW + kStageBins - 1) / kStageBins;
    p->tile_list_off = off; off = align_up(off + (size_t)p->tiles_total * R * p->groups_max * sizeof(unsigned), 256);
    p->ws_bytes = off;
    return true;
}

// One axis sample of RoI geometry `g` (prepass helper).
__device__ __forceinline__ AxisEntry tiled_axis_entry(const XfromRoi& g, bool isy, int s, int sr, int H, int W) {
    const AxisTap a = isy ? xfrom_axis(xfrom_coord(g.start_h, g.bin_h, s / sr, s % sr, sr), H)
                          : xfrom_axis(xfrom_coord(g.start_w, g.bin_w, s / sr, s % sr, sr), W);
    AxisEntry e; e.low = a.low; e.valid = a.valid ? 1 : 0; e.l = a.l; e.h = a.h;
    return e;
}

// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2).  Each half is an independent IEEE-754 RN
// operation, so results are bit-identical to the scalar _rn intrinsics; it halves the issue slots of
// the interpolation, which is issue-bound, not FLOP-bound.
typedef unsigned long long u64x;
__device__ __forceinline__ u64x pack2(float lo, float hi) { u64x r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64x v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64x fma2(u64x a, u64x b, u64x c) { u64x d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64x mul2(u64x a, u64x b) { u64x d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64x add2(u64x a, u64x b) { u64x d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ ulonglong2 lds128(const float* p) { return *reinterpret_cast<const ulonglong2*>(p); }

}  // namespace b200
