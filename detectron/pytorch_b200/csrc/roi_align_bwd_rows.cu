// roi_align_bwd_rows.cu -- Caffe2-exact RoIAlign BACKWARD, row-stationary gather path (no global atomics).
//
// The scatter formulations (reference kernel, our generic and NHWC paths) are bound by the L2 atomic unit: one
// reduction per tap (per 4 channels at best), plus a zero-fill of dX (and a transpose for the NHWC variant).  Here
// the roles are swapped: a work item is one image ROW of one 32-cell x-tile for one block of 64 / 128 channels; ONE
// WARP owns the item's accumulator ([cell][channel] in shared memory, lane = 2 / 4 adjacent channels), visits every
// bilinear tap that lands in it and read-modify-writes the accumulator with plain 64/128-bit shared accesses.  Each
// dX element is produced by exactly one warp and written exactly once with coalesced 128-byte stores, so there is
// no memset, no atomic and no scratch image; accumulation order inside an item is sequential.
//
//   prep   (1 launch, CTA 0)     per-RoI x-sample tables; every valid y-sample of every RoI contributes two
//                                "units" (row of its low cell with weight hy, row of its high cell with weight ly),
//                                bucketed by image row with a shared-memory counting sort (CSR row_off + units);
//                                rows ranked by unit count so that heavy items are handed out first.
//          (same launch, rest)   dY (R, C, PH, PW) -> dYt (R, PH, PW, C) / count: channel-innermost, so that a
//                                warp reads the gradient of a bin for its channels with one coalesced access.
//   main   (persistent warps)    item <- atomic counter; zero accumulator; for each unit of the row whose RoI
//                                overlaps the x-tile: lanes j < PW*sr build the tap records of x-sample j (cell
//                                offset, wy * wx), staged in shared memory; the gradients of the needed bins are
//                                loaded to registers (software pipelined one unit ahead); taps are applied with
//                                LDS / FFMA2 / STS.  Finally the row segment is transposed to NCHW on the way out.
//
// Numerics: per tap the term is (dY / count) * (wy * wx) with count = sr^2 in {1, 4}: the division is an exact
// scaling, so the term equals the reference's FMUL(dY, w) / count bit for bit (barring underflow); only the order
// of the additions differs, as it does between two runs of the reference's atomicAdd.
//
// Semantics: lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu (reference) :150-193, :195-270.
#include "roi_align_tiled.cuh"
#include <stdlib.h>

namespace b200 {

constexpr int kRowCells = 32;             // cells per x-tile = lanes of the write-out
constexpr int kPrepThreads = 1024;
constexpr int kTransposeChannels = 64;    // channels per transpose CTA
constexpr int kMaxRows = 8192;            // N * H supported by the shared-memory counting sort

struct __align__(16) BwdRoi {
    int batch;                            // -1: batch index out of range (contributes nothing)
    int x_lo, x_hi;                       // range of cells any x-sample of the RoI can touch
    int pad;
};

struct RowsPlan {
    int ny, nx, tiles_x, rows;
    size_t roi_off, xtab_off, ytab_off, row_off_off, row_rank_off, units_off, counter_off, dyt_off, ws_bytes;
};

static bool rows_plan(int N, int R, int C, int H, int W, int PH, int PW, int sr, RowsPlan* p) {
    if (sr < 1 || sr > 2 || (PW != 7 && PW != 14) || PH * sr > kAxisMax || PW * sr > kAxisMax) return false;
    if (R <= 0 || R > 65535 || N <= 0 || C <= 0 || (C % 64) != 0 || H <= 0 || W <= 0) return false;
    if ((long long)N * H > kMaxRows) return false;
    if ((long long)R * C * PH * PW >= (1LL << 31) || (long long)N * C * H * W >= (1LL << 31)) return false;
    p->ny = PH * sr; p->nx = PW * sr; p->tiles_x = (W + kRowCells - 1) / kRowCells; p->rows = N * H;
    size_t off = 0;
    p->roi_off = off;      off = align_up(off + (size_t)R * sizeof(BwdRoi), 256);
    p->xtab_off = off;     off = align_up(off + (size_t)R * p->nx * sizeof(AxisEntry), 256);
    p->ytab_off = off;     off = align_up(off + (size_t)R * p->ny * sizeof(AxisEntry), 256);
    p->row_off_off = off;  off = align_up(off + ((size_t)p->rows + 1) * sizeof(int), 256);
    p->row_rank_off = off; off = align_up(off + (size_t)p->rows * sizeof(int), 256);
    p->units_off = off;    off = align_up(off + (size_t)R * p->ny * 2 * sizeof(uint2), 256);
    p->counter_off = off;  off = align_up(off + 64, 256);
    p->dyt_off = off;      off = align_up(off + (size_t)R * PH * PW * C * sizeof(float), 256);
    p->ws_bytes = off;
    return true;
}

// ------------------------------------------------------------------------------------------------
// prep + transpose
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPrepThreads)
roi_align_bwd_rows_prep(const float* __restrict__ rois, const float* __restrict__ dy, float scale, int N, int R, int C, int H,
                        int W, int PH, int PW, int sr, BwdRoi* __restrict__ roi_out, AxisEntry* __restrict__ xtab,
                        AxisEntry* __restrict__ ytab, int* __restrict__ row_off, int* __restrict__ row_rank, uint2* __restrict__ units,
                        int* __restrict__ counter, float* __restrict__ dyt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int ny = PH * sr, nx = PW * sr;
    if (blockIdx.x > 0) {
        // ---- dY (r, c, bin) -> dYt (r, bin, c) / count for one RoI and one block of channels
        const int bins = PH * PW;
        const int cblocks = (C + kTransposeChannels - 1) / kTransposeChannels;
        const int r = (blockIdx.x - 1) / cblocks, c0 = ((blockIdx.x - 1) % cblocks) * kTransposeChannels;
        const int cc = min(kTransposeChannels, C - c0);
        const int stride = bins | 1;                        // odd: the transposed read below is bank-conflict free
        float* s = reinterpret_cast<float*>(smem_raw);
        const float* src = dy + ((size_t)r * C + c0) * bins;
        const float cnt = (float)(sr * sr);
        for (int k = tid; k < cc * bins; k += kPrepThreads) {
            const int c = k / bins, b = k - c * bins;
            s[c * stride + b] = __fdiv_rn(src[k], cnt);
        }
        __syncthreads();
        float* dst = dyt + (size_t)r * bins * C + c0;
        for (int k = tid; k < cc * bins; k += kPrepThreads) {
            const int b = k / cc, c = k - b * cc;
            dst[(size_t)b * C + c] = s[c * stride + b];
        }
        return;
    }
    // ---- CTA 0: tables and the row-bucketed unit lists
    const int rows = N * H;
    int* cnt = reinterpret_cast<int*>(smem_raw);            // [rows]  units per row, then the write cursor
    int* offs = cnt + rows;                                 // [rows]  exclusive prefix
    __shared__ int s_warp_sum[32];
    for (int k = tid; k < rows; k += kPrepThreads) cnt[k] = 0;
    if (tid == 0) *counter = 0;
    __syncthreads();
    const int per = ny + nx;
    for (int k = tid; k < R * per; k += kPrepThreads) {
        const int r = k / per, s = k - r * per;
        const XfromRoi g = xfrom_roi(rois + 5 * (size_t)r, scale, PH, PW, sr);
        const bool batch_ok = g.batch >= 0 && g.batch < N;
        const bool isy = s < ny;
        const AxisEntry e = tiled_axis_entry(g, isy, isy ? s : s - ny, sr, H, W);
        if (isy) {
            ytab[(size_t)r * ny + s] = e;
            if (s == 0) roi_out[r].batch = batch_ok ? g.batch : -1;
            if (batch_ok && e.valid) {
                atomicAdd(&cnt[g.batch * H + e.low], 1);
                atomicAdd(&cnt[g.batch * H + min(e.low + 1, H - 1)], 1);
            }
        } else {
            xtab[(size_t)r * nx + (s - ny)] = e;
            if (s == ny) roi_out[r].x_lo = e.low;                               // lows are monotone along the axis
            if (s == per - 1) { roi_out[r].x_hi = min(e.low + 1, W - 1); roi_out[r].pad = 0; }
        }
    }
    __syncthreads();
    // exclusive scan of cnt[0 .. rows) (rows <= 8192: 8 per thread)
    {
        const int chunk = (rows + kPrepThreads - 1) / kPrepThreads;
        const int b0 = tid * chunk;
        int local = 0;
        for (int k = 0; k < chunk; ++k) if (b0 + k < rows) local += cnt[b0 + k];
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= d) incl += v; }
        if ((tid & 31) == 31) s_warp_sum[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            int w = s_warp_sum[tid];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, d); if (tid >= d) w += v; }
            s_warp_sum[tid] = w;                            // inclusive over warps
        }
        __syncthreads();
        int run = incl - local + ((tid >> 5) > 0 ? s_warp_sum[(tid >> 5) - 1] : 0);
        for (int k = 0; k < chunk; ++k) if (b0 + k < rows) { offs[b0 + k] = run; run += cnt[b0 + k]; }
        __syncthreads();
        for (int k = tid; k < rows; k += kPrepThreads) row_off[k] = offs[k];
        if (tid == 0) row_off[rows] = s_warp_sum[31];
    }
    // rank of every row by unit count (heaviest first; ties by index): the main kernel hands items out in this order
    for (int k = tid; k < rows; k += kPrepThreads) {
        int rank = k;
        if (rows <= 1024) {                                 // O(rows^2 / threads): skipped (identity order) for very tall batches
            const int mine = cnt[k];
            rank = 0;
            for (int q = 0; q < rows; ++q) { const int o = cnt[q]; rank += (o > mine) || (o == mine && q < k); }
        }
        row_rank[rank] = k;
    }
    __syncthreads();
    for (int k = tid; k < rows; k += kPrepThreads) cnt[k] = offs[k];             // cursors
    __syncthreads();
    for (int k = tid; k < R * ny; k += kPrepThreads) {
        const int r = k / ny, i = k - r * ny;
        const int batch = (int)rois[5 * (size_t)r];
        if (batch < 0 || batch >= N) continue;
        const AxisEntry e = ytab[k];                        // written above by this CTA
        if (!e.valid) continue;
        const unsigned key = (unsigned)r | ((unsigned)i << 16);
        const int p0 = atomicAdd(&cnt[batch * H + e.low], 1);
        units[p0] = make_uint2(key, __float_as_uint(e.h));                      // row of the low cell: weight hy
        const int p1 = atomicAdd(&cnt[batch * H + min(e.low + 1, H - 1)], 1);
        units[p1] = make_uint2(key | 0x80000000u, __float_as_uint(e.l));        // row of the high cell: weight ly
    }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
template <int CPL> struct LaneVec;
template <> struct LaneVec<2> {
    u64x v[1];
    __device__ __forceinline__ void load_shared(const float* p) { v[0] = *reinterpret_cast<const u64x*>(p); }
    __device__ __forceinline__ void store_shared(float* p) const { *reinterpret_cast<u64x*>(p) = v[0]; }
    __device__ __forceinline__ void load_global(const float* p) { v[0] = __ldg(reinterpret_cast<const u64x*>(p)); }
};
template <> struct LaneVec<4> {
    u64x v[2];
    __device__ __forceinline__ void load_shared(const float* p) { const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void store_shared(float* p) const { ulonglong2 t; t.x = v[0]; t.y = v[1]; *reinterpret_cast<ulonglong2*>(p) = t; }
    __device__ __forceinline__ void load_global(const float* p) { const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(p)); v[0] = t.x; v[1] = t.y; }
};

template <int CPL> __host__ __device__ constexpr int rows_acc_stride() { return 33 * CPL; }          // words per cell (odd multiple of CPL)
template <int NX> __host__ __device__ constexpr int rows_stage_bytes() { return ((NX * 16 + 127) / 128) * 128; }
template <int CPL, int NX> __host__ __device__ constexpr int rows_warp_smem() {
    return kRowCells * rows_acc_stride<CPL>() * 4 + 2 * rows_stage_bytes<NX>();
}

__device__ __forceinline__ AxisEntry make_axis_zero() { AxisEntry e; e.low = 0; e.valid = 0; e.l = 0.f; e.h = 0.f; return e; }

// One x-sample of one unit, as the tap loop wants it: byte offsets of the two cells inside the warp's accumulator
// (negative: the cell is outside this x-tile or the sample is invalid) and the two products wy * wx.
struct __align__(16) TapRec {
    int off_lo, off_hi;
    float w_lo, w_hi;
};

// acc[cell][lane's channels] += g * w for one tap (warp-uniform skip when the cell is outside the tile)
template <int CPL>
__device__ __forceinline__ void apply_tap(float* acc, int off, float w, const LaneVec<CPL>& g, int lane) {
    if (off < 0) return;
    float* p = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(acc) + off) + lane * CPL;
    LaneVec<CPL> a;
    a.load_shared(p);
    const u64x ww = pack2(w, w);
#pragma unroll
    for (int q = 0; q < CPL / 2; ++q) a.v[q] = fma2(g.v[q], ww, a.v[q]);
    a.store_shared(p);
}

template <int PW, int SR, int CPL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
roi_align_bwd_rows(const BwdRoi* __restrict__ roi_in, const AxisEntry* __restrict__ xtab, const int* __restrict__ row_off,
                   const int* __restrict__ row_rank, const uint2* __restrict__ units, int* __restrict__ counter,
                   const float* __restrict__ dyt, float* __restrict__ dx, int N, int C, int H, int W, int PH, int tiles_x) {
    constexpr int NX = PW * SR;
    constexpr int S = rows_acc_stride<CPL>();
    constexpr int CHB = 32 * CPL;                       // channels per item
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my = smem_raw + (size_t)warp * rows_warp_smem<CPL, NX>();
    float* acc = reinterpret_cast<float*>(my);
    TapRec* stage = reinterpret_cast<TapRec*>(my + kRowCells * S * 4);          // two buffers of NX records
    constexpr int kStageRecs = rows_stage_bytes<NX>() / 16;
    const int cblocks = C / CHB;
    const int items_per_row = tiles_x * cblocks;
    const int total = N * H * items_per_row;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const int row = row_rank[item / items_per_row];             // heaviest rows first
        const int rem = item - (item / items_per_row) * items_per_row;
        const int cb = rem / tiles_x, tx = rem - cb * tiles_x;
        const int n = row / H, y = row - n * H;
        const int x0 = tx * kRowCells;
        const int c_base = cb * CHB + lane * CPL;

        // zero the accumulator
        {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            float4* a4 = reinterpret_cast<float4*>(acc);
#pragma unroll 4
            for (int k = lane; k < kRowCells * S / 4; k += 32) a4[k] = z;
        }
        __syncwarp();

        const int u_begin = row_off[row], u_end = row_off[row + 1];
        for (int base = u_begin; base < u_end; base += 32) {
            uint2 u = make_uint2(0u, 0u);
            bool ov = false;
            if (base + lane < u_end) {
                u = units[base + lane];
                const BwdRoi h = roi_in[u.x & 0xffffu];
                ov = h.x_lo <= x0 + kRowCells - 1 && h.x_hi >= x0;
            }
            unsigned m = __ballot_sync(0xffffffffu, ov);
            if (m == 0u) continue;

            // software pipeline over the overlapping units of this batch:
            //   X(k+1): x-sample table entry of the next unit in flight | A(k): records + gradient loads | B(k-1): taps
            LaneVec<CPL> g_cur[PW], g_new[PW];
            unsigned bm_cur = 0u, bm_new = 0u;
            int slot = 0;
            bool have_prev = false;
            int l0 = __ffs(m) - 1; m &= m - 1u;
            unsigned key0 = __shfl_sync(0xffffffffu, u.x, l0);
            float wy0 = __uint_as_float(__shfl_sync(0xffffffffu, u.y, l0));
            AxisEntry e0 = make_axis_zero();
            if (lane < NX) e0 = xtab[(size_t)(key0 & 0xffffu) * NX + lane];
            while (l0 >= 0) {
                // X(k+1)
                const int l1 = m ? (__ffs(m) - 1) : -1;
                if (l1 >= 0) m &= m - 1u;
                unsigned key1 = 0u; float wy1 = 0.f;
                AxisEntry e1 = make_axis_zero();
                if (l1 >= 0) {
                    key1 = __shfl_sync(0xffffffffu, u.x, l1);
                    wy1 = __uint_as_float(__shfl_sync(0xffffffffu, u.y, l1));
                    if (lane < NX) e1 = xtab[(size_t)(key1 & 0xffffu) * NX + lane];
                }
                // A(k): tap records of x-sample `lane`
                {
                    TapRec rec; rec.off_lo = -1; rec.off_hi = -1; rec.w_lo = 0.f; rec.w_hi = 0.f;
                    bool any = false;
                    if (lane < NX && e0.valid) {
                        const int lo = e0.low - x0, hi = min(e0.low + 1, W - 1) - x0;
                        if ((unsigned)lo < (unsigned)kRowCells) { rec.off_lo = lo * (S * 4); any = true; }
                        if ((unsigned)hi < (unsigned)kRowCells) { rec.off_hi = hi * (S * 4); any = true; }
                        rec.w_lo = __fmul_rn(wy0, e0.h);
                        rec.w_hi = __fmul_rn(wy0, e0.l);
                    }
                    if (lane < NX) stage[slot * kStageRecs + lane] = rec;
                    bm_new = __ballot_sync(0xffffffffu, any);       // also orders the record stores before the reads below
                    const int r = (int)(key0 & 0xffffu), i = (int)((key0 >> 16) & 0x7fffu);
                    const float* gsrc = dyt + ((size_t)(r * PH + i / SR) * PW) * C + c_base;
#pragma unroll
                    for (int pw = 0; pw < PW; ++pw)
                        if ((bm_new >> (pw * SR)) & ((1u << SR) - 1u)) g_new[pw].load_global(gsrc + (size_t)pw * C);
                }
                // B(k-1)
                if (have_prev) {
                    const TapRec* st = stage + (slot ^ 1) * kStageRecs;
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        if (!((bm_cur >> j) & 1u)) continue;
                        const TapRec rec = st[j];
                        apply_tap<CPL>(acc, rec.off_lo, rec.w_lo, g_cur[j / SR], lane);
                        apply_tap<CPL>(acc, rec.off_hi, rec.w_hi, g_cur[j / SR], lane);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int pw = 0; pw < PW; ++pw) g_cur[pw] = g_new[pw];
                bm_cur = bm_new; have_prev = true; slot ^= 1;
                l0 = l1; key0 = key1; wy0 = wy1; e0 = e1;
            }
            {   // drain: B of the last unit of the batch
                const TapRec* st = stage + (slot ^ 1) * kStageRecs;
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    if (!((bm_cur >> j) & 1u)) continue;
                    const TapRec rec = st[j];
                    apply_tap<CPL>(acc, rec.off_lo, rec.w_lo, g_cur[j / SR], lane);
                    apply_tap<CPL>(acc, rec.off_hi, rec.w_hi, g_cur[j / SR], lane);
                }
                __syncwarp();
            }
        }

        // write-out: lane = cell; CPL channels per shared load, one coalesced 128-byte row segment per channel
        {
            const int x = x0 + lane;
            const bool ok = x < W;
            float* out = dx + (((size_t)n * C + (size_t)cb * CHB) * H + y) * W + x;
            const size_t cstride = (size_t)H * W;
#pragma unroll 4
            for (int c4 = 0; c4 < 32; ++c4) {
                LaneVec<CPL> v;
                v.load_shared(acc + lane * S + c4 * CPL);
                if (ok) {
#pragma unroll
                    for (int q = 0; q < CPL / 2; ++q) {
                        float a, b; unpack2(v.v[q], a, b);
                        out[(size_t)(c4 * CPL + 2 * q) * cstride] = a;
                        out[(size_t)(c4 * CPL + 2 * q + 1) * cstride] = b;
                    }
                }
            }
        }
        __syncwarp();
    }
}

size_t roi_align_bwd_rows_workspace_bytes(int N, int R, int C, int H, int W, int PH, int PW, int sr) {
    RowsPlan p;
    return rows_plan(N, R, C, H, W, PH, PW, sr, &p) ? p.ws_bytes : 0;
}

template <int PW, int SR, int CPL, int WARPS>
static int launch_rows(const RowsPlan& p, unsigned char* ws, float* dx, int N, int C, int H, int W, int PH, cudaStream_t stream) {
    constexpr int NX = PW * SR;
    const size_t smem = (size_t)WARPS * rows_warp_smem<CPL, NX>();
    auto kern = roi_align_bwd_rows<PW, SR, CPL, WARPS>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
    if (err != cudaSuccess) return (int)err;
    if (per_sm < 1) return 1000;
    const int items = N * H * p.tiles_x * (C / (32 * CPL));
    const int warps_needed = items;
    int grid = kNumSMs * per_sm;
    if (grid * WARPS > warps_needed) grid = (warps_needed + WARPS - 1) / WARPS;
    kern<<<grid, WARPS * 32, smem, stream>>>(
        reinterpret_cast<const BwdRoi*>(ws + p.roi_off), reinterpret_cast<const AxisEntry*>(ws + p.xtab_off),
        reinterpret_cast<const int*>(ws + p.row_off_off), reinterpret_cast<const int*>(ws + p.row_rank_off),
        reinterpret_cast<const uint2*>(ws + p.units_off), reinterpret_cast<int*>(ws + p.counter_off),
        reinterpret_cast<const float*>(ws + p.dyt_off), dx, N, C, H, W, PH, p.tiles_x);
    return B200_ROI_OK;
}

// returns 1000 when the path does not apply (caller falls back to another backward path)
int roi_align_backward_rows(const float* top_diff, float scale, int N, int R, int H, int W, int C, int PH, int PW, int sr,
                            const float* rois, float* bottom_diff, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    RowsPlan p;
    if (!rows_plan(N, R, C, H, W, PH, PW, sr, &p)) return 1000;
    if (workspace == nullptr || workspace_bytes < p.ws_bytes) return 1000;
    unsigned char* ws = (unsigned char*)workspace;
    const int bins = PH * PW;
    const size_t smem_prep = (size_t)2 * p.rows * sizeof(int);
    const size_t smem_tr = (size_t)kTransposeChannels * (bins | 1) * sizeof(float);
    const size_t smem = smem_prep > smem_tr ? smem_prep : smem_tr;
    if (smem > 200 * 1024) return 1000;
    cudaError_t err = cudaFuncSetAttribute(roi_align_bwd_rows_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const int cblocks = (C + kTransposeChannels - 1) / kTransposeChannels;
    roi_align_bwd_rows_prep<<<1 + R * cblocks, kPrepThreads, smem, stream>>>(
        rois, top_diff, scale, N, R, C, H, W, PH, PW, sr, reinterpret_cast<BwdRoi*>(ws + p.roi_off),
        reinterpret_cast<AxisEntry*>(ws + p.xtab_off), reinterpret_cast<AxisEntry*>(ws + p.ytab_off),
        reinterpret_cast<int*>(ws + p.row_off_off), reinterpret_cast<int*>(ws + p.row_rank_off),
        reinterpret_cast<uint2*>(ws + p.units_off), reinterpret_cast<int*>(ws + p.counter_off),
        reinterpret_cast<float*>(ws + p.dyt_off));
    const char* e_cpl = getenv("B200_ROI_ALIGN_BWD_CPL");       // channels per lane of the main kernel: 2 | 4 (A/B tests)
    const bool want4 = !(e_cpl && e_cpl[0] == '2');
    int rc;
    if (PW == 7) {
        if (want4 && (C % 128) == 0) rc = (sr == 1) ? launch_rows<7, 1, 4, 13>(p, ws, bottom_diff, N, C, H, W, PH, stream)
                                                     : launch_rows<7, 2, 4, 13>(p, ws, bottom_diff, N, C, H, W, PH, stream);
        else rc = (sr == 1) ? launch_rows<7, 1, 2, 12>(p, ws, bottom_diff, N, C, H, W, PH, stream)
                            : launch_rows<7, 2, 2, 12>(p, ws, bottom_diff, N, C, H, W, PH, stream);
    } else {
        rc = (sr == 1) ? launch_rows<14, 1, 2, 12>(p, ws, bottom_diff, N, C, H, W, PH, stream)
                       : launch_rows<14, 2, 2, 12>(p, ws, bottom_diff, N, C, H, W, PH, stream);
    }
    if (rc != B200_ROI_OK) return rc;
    return finish_launch(2);
}

}  // namespace b200
