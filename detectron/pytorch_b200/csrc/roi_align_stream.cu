// roi_align_stream.cu -- Caffe2-exact RoIAlign FORWARD, "streaming strip" fast path (cp.async + mbarrier ring).
//
// The feature map is cut into vertical strips of WX columns (+ a halo); one persistent CTA per SM owns a
// contiguous piece of the linear space (image, strip, 32-channel group, row) and streams the rows of its
// strips top to bottom through a ring of K row slots in shared memory.  Four producer warps fetch rows with
// 4-byte cp.async (LDGSTS: global -> shared without a register round trip, 128-byte coalesced per warp) and
// signal per-slot mbarriers (cp.async.mbarrier.arrive); they run ahead of 16 consumer warps, which release
// slots on a second set of mbarriers, so loading overlaps the arithmetic completely.
//
// Shared-memory layout of one row slot: [32 channels][SX columns], SX odd.  The compute mapping is
// lane = channel: one warp evaluates one bilinear tap of one sample for 32 channels with ONE conflict-free
// LDS.32 (bank = c * SX + x), and all addressing / weights are warp-uniform.
// Why not TMA (cp.async.bulk.tensor): measured on this B200 (tools/tma_probe.cu) a tile load whose innermost
// start coordinate is not 16-byte aligned faults, and with 16-byte placement granularity every channel row would
// start at a bank that is a multiple of 4 -- a lane = channel access would hit at most 8 banks (4-way conflict).
// The 4-byte cp.async can place a row at any word, which is what the odd channel stride needs.
//
// Work = "fragments": up to 8 consecutive bins of one bin row of one RoI, keyed by the first feature row
// they read.  A fragment is processed by ONE warp when the rows [key, end) it needs are resident, so every
// output element is computed by one lane in the reference's own operation order and written exactly once
// with a plain store: bit-identical to the reference kernel, deterministic, no zero-fill of the output.
// Only bins whose samples cannot be resident together (x-span beyond the strip halo, y-span beyond the
// ring) are cut into per-sample fragments that accumulate with red.global.add onto elements the fill
// kernel zeroed (at most two partial sums per element -> still deterministic).
//
// Prepass (channel independent, two small kernels): `count` builds the per-RoI axis tables (all IEEE
// divisions; ring-slot byte offsets precomputed) and a histogram of fragments per (image, strip, row); its
// last CTA scans the histogram (CSR row pointers) and cuts the linear space into one piece per SM of equal
// estimated cost; `fill` writes the 8-byte fragment records into the CSR.
//
// Semantics: lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu (reference) :16-63, :65-121.
#include "common.cuh"
#include <mutex>
#include <stdlib.h>

namespace b200 {

namespace {

#ifndef B200_STREAM_CWARPS
#define B200_STREAM_CWARPS 16
#endif
constexpr int kConsumerWarps = B200_STREAM_CWARPS;
constexpr int kProducerWarps = 4;
constexpr int kStreamThreads = 32 * (kConsumerWarps + kProducerWarps);
constexpr int kFragBins = 8;                                  // bins per fragment (one flush)
constexpr int kStageStride = 36;                              // words per staged bin: conflict-free both ways
constexpr int kTabWords = 18 * 4;                              // per-warp axis-table buffer: 16 x entries + 2 y entries
constexpr int kStageWordsPerWarp = kFragBins * kStageStride + kTabWords;
constexpr int kMaxSlots = 32;                                 // ring depth K <= 32 (phase bits live in one register)
constexpr int kAxisMaxS = 32;                                 // P * sampling_ratio per axis
constexpr int kPrepThreads = 128;
constexpr unsigned kRowCost = 8u;                             // cost of streaming one row (16 acquires + staging), in bins
constexpr int kScanCap = 6144;                                // keys whose cost prefix is kept in shared memory by the scan
constexpr unsigned kSmemBudget = 227u * 1024u;

typedef unsigned long long u64;

// Debug build (-DB200_STREAM_DEBUG, tools/stream_debug.py): every warp leaves progress markers in a HOST-pinned buffer
// registered with b200_roi_ops_debug_timing_buffer(), so that they survive a trap.  Compiled out by default.
#ifdef B200_STREAM_DEBUG
__device__ u64* g_stream_dbg = nullptr;
#define SDBG(slot, val)                                                                                      \
    do {                                                                                                     \
        if (g_stream_dbg != nullptr && (threadIdx.x & 31) == 0) {                                            \
            volatile u64* dbg_p = g_stream_dbg + ((size_t)blockIdx.x * 17 + (threadIdx.x >> 5)) * 8;        \
            dbg_p[slot] = (u64)(val);                                                                        \
            __threadfence_system();                                                                          \
        }                                                                                                    \
    } while (0)
#else
#define SDBG(slot, val) do { } while (0)
#endif

// Timing build (-DB200_STREAM_TIMING, tools/stream_timing.py): every warp leaves SM clock stamps and wait-cycle totals in a
// DEVICE buffer registered the same way ([CTA][warp][8] u64).  Compiled out by default.
#ifdef B200_STREAM_TIMING
__device__ u64* g_stream_tim = nullptr;
#define STIM_DECL u64 stim_wait = 0, stim_n0 = 0, stim_n1 = 0, stim_first = 0, stim_comp = 0, stim_c0 = 0
#define STIM_CLOCK() ((u64)clock64())
#define STIM(slot, val)                                                                                      \
    do {                                                                                                     \
        if (g_stream_tim != nullptr && (threadIdx.x & 31) == 0)                                              \
            g_stream_tim[((size_t)blockIdx.x * (kConsumerWarps + kProducerWarps) + (threadIdx.x >> 5)) * 8 + (slot)] = (u64)(val); \
    } while (0)
#define STIM_DO(stmt) do { stmt; } while (0)
#else
#define STIM_DECL
#define STIM_CLOCK() 0ull
#define STIM(slot, val) do { } while (0)
#define STIM_DO(stmt) do { } while (0)
#endif

constexpr int kMaxLevels = 6;                                 // feature maps served by one launch (FPN: P2..P6)
constexpr int kMaxCols = 96;                                  // strip columns (level, image, strip) of one launch

struct StreamLevel {                                          // one feature map of the pyramid (a plain call has one)
    const float* bottom;                                      // (N, C, H, W)
    float scale;
    int H, W, S;                                              // S: strips per image
    int q_base;                                               // its first strip column; column = q_base + image * S + strip
    int roi_begin;                                            // its first RoI (RoIs are stored level-major)
    int pad;
};

struct StreamGeom {
    int N, R, C, PH, PW, sr;
    int ny, nx;
    int SX, WX, K;              // columns held per row (odd, 32 m + 1), strip core width, ring depth
    int row_bytes;              // bytes of one ring slot = 32 * SX * 4
    int L;                      // levels
    int Q, keys;                // strip columns, keys = sum over columns of their rows
    int G;                      // 32-channel groups
    int pieces;                 // persistent CTAs
    int max_entries;
    StreamLevel lv[kMaxLevels];
    int colstart[kMaxCols + 1]; // first key of every column (its rows are consecutive keys); colstart[Q] = keys
};

// level of RoI r / of strip column q (at most kMaxLevels entries: linear scan)
__host__ __device__ __forceinline__ int level_of_roi(const StreamGeom& g, int r) {
    int l = 0;
    while (l + 1 < g.L && r >= g.lv[l + 1].roi_begin) ++l;
    return l;
}

struct StreamWs {
    uint4* ytab;                // [R][ny] {hy, ly, byte offset of the ring slot of row y_low, 0}; the lower tap row is the NEXT slot
    uint4* xtab;                // [R][nx] {hx, lx, x_low * 4, 0}
    int* hist;                  // [keys] fragments per key                           (zero block)
    int* cost;                  // [keys] estimated cost per key                      (zero block)
    int* cursor;                // [keys] fill cursors                                (zero block)
    int* maxend;                // [keys] max over the key's fragments of `end`       (zero block)
    int* ticket;                // [4]                                                (zero block)
    int* rowptr;                // [keys + 1] CSR offsets into entries
    unsigned* cpre;             // [keys + 1] exclusive prefix of the cost
    int* piece_start;           // [pieces + 1] linear start index of every piece
    uint2* entries;             // fragment records
};

// fragment record: x = r | ph << 16 | pw0 << 21 | (npw - 1) << 26 | red << 29 ; y = key | (end - key) << 16 | smask << 24
__device__ __forceinline__ uint2 pack_entry(int r, int ph, int pw0, int npw, int red, int key, int end, unsigned smask) {
    uint2 e;
    e.x = (unsigned)r | ((unsigned)ph << 16) | ((unsigned)pw0 << 21) | ((unsigned)(npw - 1) << 26) | ((unsigned)red << 29);
    e.y = (unsigned)key | ((unsigned)(end - key) << 16) | (smask << 24);
    return e;
}

// ------------------------------------------------------------------------------------------------
// axis samples, adjusted so that both taps are always inside the map
// ------------------------------------------------------------------------------------------------
struct AdjTap {
    int   low;                  // low cell in [0, size - 2]
    float l, h;                 // weights of cell low + 1 / cell low (0, 0 for a sample the reference skips)
};

// The reference clamps a sample on the last cell to low = high = size - 1 with weights (h, l) = (1, 0).  Reading
// cells (size - 2, size - 1) with weights (0, 1) instead is the same sum term for term (x*1 and y*0 swap places in
// the FMA chain; the zero products do not change any finite partial sum), and keeps every tap in the map.
__device__ __forceinline__ AdjTap adj_axis(float v, int size) {
    const AxisTap t = xfrom_axis(v, size);
    AdjTap a;
    a.low = t.low; a.l = t.l; a.h = t.h;
    if (t.low >= size - 1) { a.low = size - 2; a.l = 1.f; a.h = 0.f; }
    if (!t.valid) { a.l = 0.f; a.h = 0.f; }
    return a;
}

// Enumerate the fragments of bin row `ph` of one RoI.  yl / yh: the reference's low / high tap row per y sample
// (equal on the clamped last row); xl: adjusted low cell per x sample.
// emit(strip, key, end, pw0, npw, smask, red, zero_owner)
template <int SR, class Emit>
__device__ __forceinline__ void enum_row(const int* yl, const int* yh, const int* xl, int ph, const StreamGeom& g, int S, Emit&& emit) {
    const int i0 = ph * SR, i1 = i0 + SR - 1;
    int ngroups = 1;
    int key[2], end[2];
    unsigned ym[2];
    key[0] = yl[i0]; end[0] = yh[i1] + 1; ym[0] = (1u << SR) - 1u;
    key[1] = 0; end[1] = 0; ym[1] = 0;
    if (SR == 2 && end[0] - key[0] > g.K) {          // rows cannot be resident together: one fragment per y sample
        ngroups = 2;
        end[0] = yh[i0] + 1; ym[0] = 1u;
        key[1] = yl[i1]; end[1] = yh[i1] + 1; ym[1] = 2u;
    }
    constexpr unsigned kXFull = (1u << SR) - 1u;
    for (int gi = 0; gi < ngroups; ++gi) {
        int run_s = -1, run_pw0 = 0, run_n = 0;
        unsigned run_xm = 0;
        auto flush = [&]() {
            if (run_n == 0) return;
            unsigned smask;
            if (SR == 1) smask = 1u;
            else smask = ((ym[gi] & 1u) ? run_xm : 0u) | ((ym[gi] & 2u) ? (run_xm << 2) : 0u);
            const bool red = (ym[gi] != kXFull) || (run_xm != kXFull);
            const bool owner = red && (ym[gi] & 1u) && (run_xm & 1u);
            emit(run_s, key[gi], end[gi], run_pw0, run_n, smask, red ? 1 : 0, owner);
            run_n = 0;
        };
        auto push = [&](int s, unsigned xm, int pw) {
            if (run_n > 0 && s == run_s && xm == run_xm && run_n < kFragBins && pw == run_pw0 + run_n) { ++run_n; return; }
            flush();
            run_s = s; run_xm = xm; run_pw0 = pw; run_n = 1;
        };
        for (int pw = 0; pw < g.PW; ++pw) {
            const int j0 = pw * SR, j1 = j0 + SR - 1;
            const int s0 = min(xl[j0] / g.WX, S - 1);
            if (SR == 1 || xl[j1] + 1 <= s0 * g.WX + g.SX - 1) {
                push(s0, kXFull, pw);
            } else {                                   // x samples in different strips: one fragment per x sample
                push(s0, 1u, pw);
                push(min(xl[j1] / g.WX, S - 1), 2u, pw);
            }
        }
        flush();
    }
}

// ------------------------------------------------------------------------------------------------
// prepass 1: tables + histogram; last CTA: scan + partition
// ------------------------------------------------------------------------------------------------
template <int SR>
__global__ void __launch_bounds__(kPrepThreads)
stream_count(const float* __restrict__ rois, StreamGeom g, StreamWs ws) {
    __shared__ int s_yl[kAxisMaxS], s_yh[kAxisMaxS], s_xl[kAxisMaxS];
    __shared__ int s_last;
    __shared__ unsigned s_h[kPrepThreads], s_c[kPrepThreads];
    const int r = blockIdx.x, t = threadIdx.x;
    const StreamLevel& lv = g.lv[level_of_roi(g, r)];
    const XfromRoi geo = xfrom_roi(rois + 5 * (size_t)r, lv.scale, g.PH, g.PW, g.sr);
    if (t < g.ny + g.nx) {
        const bool isy = t < g.ny;
        const int s = isy ? t : t - g.ny;
        if (isy) {
            // rows exactly as the reference takes them (low == high == H - 1 on the clamped last row, weights (1, 0)): the
            // lower tap then reads whatever the next ring slot holds with weight 0 -- always finite, the ring starts zeroed
            const AxisTap a = xfrom_axis(xfrom_coord(geo.start_h, geo.bin_h, s / SR, s % SR, SR), lv.H);
            s_yl[s] = a.low; s_yh[s] = a.high;
            uint4 e;
            e.x = __float_as_uint(a.valid ? a.h : 0.f); e.y = __float_as_uint(a.valid ? a.l : 0.f);
            e.z = (unsigned)((a.low % g.K) * g.row_bytes);
            e.w = 0u;
            ws.ytab[(size_t)r * g.ny + s] = e;
        } else {
            const AdjTap a = adj_axis(xfrom_coord(geo.start_w, geo.bin_w, s / SR, s % SR, SR), lv.W);
            s_xl[s] = a.low;
            uint4 e;
            e.x = __float_as_uint(a.h); e.y = __float_as_uint(a.l);
            e.z = (unsigned)(a.low * 4); e.w = 0u;
            ws.xtab[(size_t)r * g.nx + s] = e;
        }
    }
    __syncthreads();
    const bool batch_ok = geo.batch >= 0 && geo.batch < g.N;
    if (batch_ok && t < g.PH) {
        const int cbase = lv.q_base + geo.batch * lv.S;
        enum_row<SR>(s_yl, s_yh, s_xl, t, g, lv.S, [&](int s, int key, int end, int pw0, int npw, unsigned smask, int red, bool owner) {
            const int k = g.colstart[cbase + s] + key;
            atomicAdd(&ws.hist[k], 1);
            atomicAdd(&ws.cost[k], 2 + npw);
            (void)end; (void)pw0; (void)smask; (void)red; (void)owner;
        });
    }
    // ---- last CTA: CSR row pointers, cost prefix, piece boundaries
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(&ws.ticket[0], 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // Everything below is latency bound (one CTA): the histogram is pulled into shared memory with all loads of a
    // chunk in flight, scanned there, and the piece boundaries are binary-searched in the shared copy of the prefix.
    __shared__ unsigned s_pre[kScanCap + 1];                  // exclusive cost prefix (keys <= kScanCap)
    __shared__ unsigned s_carry[2];
    const int keys = g.keys;
    const bool in_smem = keys <= kScanCap;
    if (t == 0) { s_carry[0] = 0u; s_carry[1] = 0u; }
    __syncthreads();
    constexpr int kPer = 8, kChunk = kPrepThreads * kPer;
    for (int base = 0; base < keys; base += kChunk) {
        unsigned h[kPer], c[kPer];
#pragma unroll
        for (int j = 0; j < kPer; ++j) {                      // thread t owns keys base + t * kPer + j; loads are independent
            const int k = base + t * kPer + j;
            h[j] = k < keys ? (unsigned)__ldcg(&ws.hist[k]) : 0u;
            c[j] = k < keys ? (unsigned)__ldcg(&ws.cost[k]) + kRowCost : 0u;
        }
        unsigned sh = 0, sc = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { sh += h[j]; sc += c[j]; }
        s_h[t] = sh; s_c[t] = sc;
        __syncthreads();
        if (t < 32) {                                         // exclusive scan of the 128 thread sums by one warp (4 per lane)
            unsigned vh[4], vc[4], th = 0, tc = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { vh[j] = s_h[t * 4 + j]; vc[j] = s_c[t * 4 + j]; th += vh[j]; tc += vc[j]; }
            unsigned ih = th, ic = tc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned uh = __shfl_up_sync(0xffffffffu, ih, d), uc = __shfl_up_sync(0xffffffffu, ic, d);
                if (t >= d) { ih += uh; ic += uc; }
            }
            unsigned rh = ih - th + s_carry[0], rc = ic - tc + s_carry[1];
#pragma unroll
            for (int j = 0; j < 4; ++j) { s_h[t * 4 + j] = rh; s_c[t * 4 + j] = rc; rh += vh[j]; rc += vc[j]; }
            __syncwarp();
            if (t == 31) { s_carry[0] = rh; s_carry[1] = rc; }
        }
        __syncthreads();
        unsigned rh = s_h[t], rc = s_c[t];
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int k = base + t * kPer + j;
            if (k < keys) {
                ws.rowptr[k] = (int)rh; ws.cpre[k] = rc;
                if (in_smem) s_pre[k] = rc;
            }
            rh += h[j]; rc += c[j];
        }
        __syncthreads();
    }
    if (t == 0) {
        ws.rowptr[keys] = (int)s_carry[0]; ws.cpre[keys] = s_carry[1];
        if (in_smem) s_pre[keys] = s_carry[1];
    }
    __threadfence();
    __syncthreads();
    // piece p starts where the cumulative cost over the linear order (column q, channel group, row) reaches p/pieces
    auto pre = [&](size_t k) -> unsigned { return in_smem ? s_pre[k] : __ldcg(&ws.cpre[k]); };
    const unsigned total = s_carry[1];
    const u64 grand = (u64)total * (u64)g.G;
    for (int p = t; p <= g.pieces; p += kPrepThreads) {
        int L;
        if (p == 0) L = 0;
        else if (p == g.pieces) L = g.G * g.keys;
        else {
            const u64 target = grand / (u64)g.pieces * (u64)p + (grand % (u64)g.pieces) * (u64)p / (u64)g.pieces;
            int lo = 0, hi = g.Q - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((u64)g.G * (u64)pre((size_t)g.colstart[mid]) <= target) lo = mid; else hi = mid - 1;
            }
            const int q = lo;
            const int k0 = g.colstart[q], Hq = g.colstart[q + 1] - k0;
            const unsigned c0 = pre((size_t)k0);
            const unsigned colsum = pre((size_t)k0 + Hq) - c0;              // >= kRowCost * Hq > 0
            u64 rem = target - (u64)g.G * (u64)c0;
            int gg = (int)(rem / colsum);
            if (gg > g.G - 1) gg = g.G - 1;
            rem -= (u64)gg * colsum;
            int ylo = 0, yhi = Hq - 1;
            while (ylo < yhi) {
                const int mid = (ylo + yhi + 1) >> 1;
                if ((u64)(pre((size_t)k0 + mid) - c0) <= rem) ylo = mid; else yhi = mid - 1;
            }
            L = g.G * k0 + gg * Hq + ylo;                                   // linear order: (column, channel group, row)
        }
        ws.piece_start[p] = L;
    }
}

// ------------------------------------------------------------------------------------------------
// prepass 2: fragment records into the CSR; zero-fill of the elements that are accumulated with red.add
// ------------------------------------------------------------------------------------------------
template <int SR>
__global__ void __launch_bounds__(kPrepThreads)
stream_fill(const float* __restrict__ rois, StreamGeom g, StreamWs ws, float* __restrict__ out, const int* __restrict__ row_map) {
    __shared__ int s_yl[kAxisMaxS], s_yh[kAxisMaxS], s_xl[kAxisMaxS];
    __shared__ unsigned short s_zero[kAxisMaxS * kAxisMaxS];
    __shared__ int s_nzero;
    const int r = blockIdx.x, t = threadIdx.x;
    const StreamLevel& lv = g.lv[level_of_roi(g, r)];
    const XfromRoi geo = xfrom_roi(rois + 5 * (size_t)r, lv.scale, g.PH, g.PW, g.sr);
    if (t == 0) s_nzero = 0;
    if (t < g.ny + g.nx) {
        const bool isy = t < g.ny;
        const int s = isy ? t : t - g.ny;
        if (isy) {
            const AxisTap a = xfrom_axis(xfrom_coord(geo.start_h, geo.bin_h, s / SR, s % SR, SR), lv.H);
            s_yl[s] = a.low; s_yh[s] = a.high;
        } else {
            s_xl[s] = adj_axis(xfrom_coord(geo.start_w, geo.bin_w, s / SR, s % SR, SR), lv.W).low;
        }
    }
    __syncthreads();
    const bool batch_ok = geo.batch >= 0 && geo.batch < g.N;
    const int bins = g.PH * g.PW;
    if (batch_ok) {
        if (t < g.PH) {
            const int cbase = lv.q_base + geo.batch * lv.S;
            enum_row<SR>(s_yl, s_yh, s_xl, t, g, lv.S, [&](int s, int key, int end, int pw0, int npw, unsigned smask, int red, bool owner) {
                const int k = g.colstart[cbase + s] + key;
                const int pos = ws.rowptr[k] + atomicAdd(&ws.cursor[k], 1);
                if (pos < g.max_entries) ws.entries[pos] = pack_entry(r, t, pw0, npw, red, key, end, smask);
                atomicMax(&ws.maxend[k], end);
                if (owner) {
                    const int z = atomicAdd(&s_nzero, npw);
                    for (int i = 0; i < npw; ++i) s_zero[z + i] = (unsigned short)(t * g.PW + pw0 + i);
                }
            });
        }
    } else {
        for (int i = t; i < bins; i += kPrepThreads) s_zero[i] = (unsigned short)i;       // the reference would read out of bounds
        if (t == 0) s_nzero = bins;
    }
    __syncthreads();
    const int nz = s_nzero;
    if (nz == 0) return;
    float* out_r = out + (size_t)(row_map ? row_map[r] : r) * g.C * bins;
    for (int idx = t; idx < g.C * nz; idx += kPrepThreads) {
        const int c = idx / nz, k = idx - c * nz;
        out_r[(size_t)c * bins + s_zero[k]] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + cp.async
// ------------------------------------------------------------------------------------------------
// Watchdog evidence: a wait that is half way to the trap leaves a record here ([CTA][32] u64; host-pinned, optional).
__device__ u64* g_stream_dead = nullptr;

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"     // %3: the hardware may park the thread this long (ns)
        "selp.u32 %0, 1, 0, p;\n"
        "}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    return ok != 0;
}
// try_wait suspends the thread in hardware for a bounded time per attempt.  A wait that is still pending after 2^20
// attempts (seconds) is a protocol bug: trap (surfaces as a launch failure) instead of hanging the GPU.
#ifdef B200_STREAM_WATCH
#define WCTX(expr) (expr)
#else
#define WCTX(expr) 0ull
#endif
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity, unsigned tag = 0u, u64 ctx0 = 0ull, u64 ctx1 = 0ull) {
    if (mbar_try_wait(bar, parity)) return;
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        ++spins;
        if (spins == (1 << 19) && g_stream_dead != nullptr && (threadIdx.x & 31) == 0) {
            // half way to the trap: say who is stuck on what (host-pinned buffer registered by the caller, else nothing)
            g_stream_dead[blockIdx.x * 32 + (threadIdx.x >> 5)] =
                0xDEAD000000000000ull | ((u64)tag << 16) | ((u64)parity << 8) | (u64)((bar >> 3) & 0xff);
#ifdef B200_STREAM_WATCH
            g_stream_dead[148 * 32 + (blockIdx.x * 32 + (threadIdx.x >> 5)) * 2] = ctx0;
            g_stream_dead[148 * 32 + (blockIdx.x * 32 + (threadIdx.x >> 5)) * 2 + 1] = ctx1;
#endif
            __threadfence_system();
        }
        if (spins > (1 << 20)) {
            SDBG(6, 0xDEAD000000000000ull | ((u64)tag << 16) | ((u64)parity << 8) | (u64)((bar >> 3) & 0xff));
            __trap();
        }
    }
}
// 4-byte cp.async (LDGSTS): global -> shared without registers.
__device__ __forceinline__ void cp_async4(unsigned dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// ok == false copies nothing and zero-fills the word
__device__ __forceinline__ void cp_async4_if(unsigned dst, const float* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0) : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(unsigned bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

struct StreamArgs {
    const uint4* ytab;
    const uint4* xtab;
    const uint2* entries;
    const int* rowptr;
    const int* maxend;
    const int* piece_start;
    float* out;
    const int* row_map;
    int C, G, K, WX, PH, PW, ny, nx, L, Q;
    StreamLevel lv[kMaxLevels];
    int colstart[kMaxCols + 1];
};

// one item = the part of one (strip column q, channel group) that lies in this CTA's piece
struct Item {
    int lvl, n, s, g, ya, yb, e0, e1, yhi, kbase;
};

// Decode the item that starts at linear index L (all lanes of the calling warp; yhi by warp reduction).
__device__ __forceinline__ Item decode_item(int L, int L1, const StreamArgs& a, int lane) {
    Item it;
    int lo = 0, hi = a.Q - 1;                                  // column: the last one whose first linear index is <= L
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.G * a.colstart[mid] <= L) lo = mid; else hi = mid - 1;
    }
    const int q = lo;
    it.kbase = a.colstart[q];
    const int Hq = a.colstart[q + 1] - it.kbase;
    const int rem = L - a.G * it.kbase;
    it.g = rem / Hq;
    it.ya = rem - it.g * Hq;
    it.yb = min(Hq, it.ya + (L1 - L));
    int l = 0;
    while (l + 1 < a.L && q >= a.lv[l + 1].q_base) ++l;
    it.lvl = l;
    const int qq = q - a.lv[l].q_base;
    it.n = qq / a.lv[l].S;
    it.s = qq - it.n * a.lv[l].S;
    it.e0 = __ldg(&a.rowptr[it.kbase + it.ya]);
    it.e1 = __ldg(&a.rowptr[it.kbase + it.yb]);
    int m = 0;
    if (it.e1 > it.e0) {
        for (int y = it.ya + lane; y < it.yb; y += 32) m = max(m, __ldg(&a.maxend[it.kbase + y]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    it.yhi = m;
    return it;
}

// Packed fp32x2 arithmetic (sm_100 FMUL2): both halves are independent IEEE-754 RN multiplications, i.e. bit-identical
// to two __fmul_rn; used for the warp-uniform weight products only (it halves their issue slots).
__device__ __forceinline__ u64 pack2f(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2f(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2f(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct Taps { float v1, v2, v3, v4; };

// The four taps of one sample for this lane's channel: rows (slot, slot + 1) x columns (x_low, x_low + 1).  The lower row
// always sits in the NEXT physical slot (the ring has a mirror of slot 0 behind slot K - 1), so one address serves all four.
template <int RB>
__device__ __forceinline__ Taps load_taps(unsigned at) {
    Taps t;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t.v1) : "r"(at));
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(t.v2) : "r"(at));
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(t.v3) : "r"(at), "n"(RB));
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(t.v4) : "r"(at), "n"(RB + 4));
    return t;
}

// One bilinear sample, the reference's rounding recipe:
// val = FFMA(v4, w4, FFMA(v3, w3, FFMA(v1, w1, FMUL(v2, w2)))) with w1 = hy*hx, w2 = hy*lx, w3 = ly*hx, w4 = ly*lx.
// hyhy = (hy, hy), lyly = (ly, ly), hxlx = (hx, lx) packed.
__device__ __forceinline__ float sample_val(const Taps& t, u64 hyhy, u64 lyly, u64 hxlx) {
    float w1, w2, w3, w4;
    unpack2f(mul2f(hyhy, hxlx), w1, w2);
    unpack2f(mul2f(lyly, hxlx), w3, w4);
    return __fmaf_rn(t.v4, w4, __fmaf_rn(t.v3, w3, __fmaf_rn(t.v1, w1, __fmul_rn(t.v2, w2))));
}

// Stage row y of one (image, strip, channel group) into a ring slot: M full 32-column chunks per channel
// (lane = column: 128-byte coalesced) plus the last column x = 32 M for all channels at once (lane = channel).
// ASYNC: 4-byte cp.async; otherwise LDG -> registers -> STS in batches of 8 channels.
template <int M, bool ASYNC>
__device__ __forceinline__ void stage_row(const float* __restrict__ src_row, size_t plane, unsigned dst, int xw, int cvalid, int lane) {
    // src_row: (channel 0, row y, column x0); xw = number of valid columns from x0; cvalid = number of real channels
    constexpr int SX = 32 * M + 1;
    if (ASYNC) {
        if (cvalid == 32 && xw >= SX) {                       // interior: no predicates, constant offsets
            const float* sc = src_row + lane;
            const unsigned d = dst + (unsigned)lane * 4u;
#pragma unroll 8
            for (int c = 0; c < 32; ++c) {
#pragma unroll
                for (int m = 0; m < M; ++m) cp_async4(d + (unsigned)(c * SX + 32 * m) * 4u, sc + 32 * m);
                sc += plane;
            }
            cp_async4(dst + (unsigned)(lane * SX + 32 * M) * 4u, src_row + (size_t)lane * plane + 32 * M);
        } else {
#pragma unroll 4
            for (int c = 0; c < 32; ++c) {
                const bool cok = c < cvalid;
                const float* sc = src_row + (size_t)(cok ? c : 0) * plane;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int x = lane + 32 * m;
                    const bool ok = cok && x < xw;
                    cp_async4_if(dst + (unsigned)(c * SX + x) * 4u, sc + (ok ? x : 0), ok);
                }
            }
            const bool ok = lane < cvalid && 32 * M < xw;
            cp_async4_if(dst + (unsigned)(lane * SX + 32 * M) * 4u, src_row + (ok ? (size_t)lane * plane + 32 * M : 0), ok);
        }
    } else {
        for (int c8 = 0; c8 < 32; c8 += 8) {
            float v[8][M];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool cok = c8 + j < cvalid;
                const float* sc = src_row + (size_t)(cok ? c8 + j : 0) * plane;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int x = lane + 32 * m;
                    v[j][m] = (cok && x < xw) ? __ldcs(sc + x) : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int m = 0; m < M; ++m)
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + (unsigned)((c8 + j) * SX + lane + 32 * m) * 4u), "f"(v[j][m]) : "memory");
        }
        const bool ok = lane < cvalid && 32 * M < xw;
        const float t = ok ? __ldcs(src_row + (size_t)lane * plane + 32 * M) : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + (unsigned)(lane * SX + 32 * M) * 4u), "f"(t) : "memory");
    }
}

template <int SR, int M, bool ASYNC>
__global__ void __launch_bounds__(kStreamThreads, 1)
roi_align_stream_fwd(const StreamArgs a) {
    constexpr int SX = 32 * M + 1;
    constexpr int RB = 128 * SX;                                  // bytes of one ring slot
    extern __shared__ unsigned char smem_raw[];
    const unsigned ring = (smem_addr(smem_raw) + 127u) & ~127u;
    unsigned char* ring_ptr = smem_raw + (ring - smem_addr(smem_raw));
    // ring: K logical slots + one mirror of slot 0 behind slot K - 1, so that row y + 1 is always the next physical slot
    float* stage_all = reinterpret_cast<float*>(ring_ptr + (size_t)(a.K + 1) * RB);
    const unsigned bars = ring + (unsigned)(a.K + 1) * (unsigned)RB + kConsumerWarps * kStageWordsPerWarp * 4;
    // full[k] at bars + 8k (32 arrivals: the lanes of the producer warp that staged the row),
    // empty[k] at bars + 8 * (kMaxSlots + k) (one arrival per consumer warp)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    {   // the ring starts zeroed: a tap with weight 0 may read a slot that was never written (clamped last row)
        const int n16 = (a.K + 1) * (RB / 16);
        float4* r4 = reinterpret_cast<float4*>(ring_ptr);
        for (int k = tid; k < n16; k += kStreamThreads) r4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        for (int k = 0; k < a.K; ++k) { mbar_init(bars + 8u * k, 32u); mbar_init(bars + 8u * (kMaxSlots + k), kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int L0 = __ldg(&a.piece_start[blockIdx.x]), L1 = __ldg(&a.piece_start[blockIdx.x + 1]);
    SDBG(0, 1); SDBG(1, ((u64)(unsigned)L0 << 32) | (unsigned)L1);
    STIM_DECL;
    STIM(0, STIM_CLOCK()); STIM(7, ((u64)(unsigned)L0 << 32) | (unsigned)L1);

    if (warp >= kConsumerWarps) {
        // =============================== producers ===============================
        // Row i of the CTA's row stream (all items, in order) is staged by producer warp i % kProducerWarps.  Every
        // producer tracks the phase of every slot (one flip per row of the stream) but waits only for its own rows.
        const int pw = warp - kConsumerWarps;
        unsigned ephase = 0xffffffffu;                // waiting for parity 1 on a fresh barrier passes immediately
        int turn = 0;                                 // stream row index modulo kProducerWarps
        for (int L = L0; L < L1;) {
            const Item it = decode_item(L, L1, a, lane);
            L += it.yb - it.ya;
            SDBG(2, ((u64)it.ya << 48) | ((u64)it.yb << 32) | ((u64)it.yhi << 16) | (u64)(it.e1 - it.e0));
            if (it.e1 <= it.e0) continue;
            int slot = it.ya % a.K;
            const StreamLevel& lv = a.lv[it.lvl];
            const size_t plane = (size_t)lv.H * lv.W;
            const int x0 = it.s * a.WX;
            const int c0 = it.g * 32;
            const int cvalid = min(32, a.C - c0);
            const int xw = lv.W - x0;
            const int rowlen = lv.W;
            const float* src = lv.bottom + ((size_t)it.n * a.C + c0) * plane + (size_t)it.ya * lv.W + x0;
            for (int y = it.ya; y < it.yhi; ++y) {
                // EVERY producer waits for EVERY row's slot, not only for the rows it stages: successive uses of a slot can
                // belong to different producer warps (K % kProducerWarps != 0, or an item boundary), and a parity wait that
                // runs two phases ahead of its barrier passes at once -- a producer that skipped the wait of use k - 1 could
                // overtake the warp that stages it and overwrite use k - 2.  Having seen use k - 2 released here, the test
                // for use k can only be one phase ahead (measured: K = 15, four producers, rows 2 and 17 of one item).
                const unsigned full = bars + 8u * slot, empty = bars + 8u * (kMaxSlots + slot);
                STIM_DO(stim_n1 = STIM_CLOCK());
                mbar_wait(empty, (ephase >> slot) & 1u, 0x1000u + (unsigned)y,
                          WCTX(((u64)it.ya << 48) | ((u64)it.yb << 32) | ((u64)it.yhi << 16) | (u64)(it.lvl << 12) | (u64)(it.g << 8) | (u64)it.s),
                          WCTX(((u64)ephase << 32) | (u64)(unsigned)(L - (it.yb - it.ya))));
                STIM_DO(stim_wait += STIM_CLOCK() - stim_n1);
                if (turn == pw) {
                    STIM_DO(++stim_n0; if (!stim_first) stim_first = STIM_CLOCK());
                    SDBG(4, ((u64)y << 32) | (u64)slot);
                    stage_row<M, ASYNC>(src, plane, ring + (unsigned)slot * (unsigned)RB, xw, cvalid, lane);
                    if (slot == 0) stage_row<M, ASYNC>(src, plane, ring + (unsigned)a.K * (unsigned)RB, xw, cvalid, lane);   // mirror
                    if (ASYNC) cp_async_arrive(full); else mbar_arrive(full);
                }
                ephase ^= 1u << slot;
                if (++slot == a.K) slot = 0;
                if (++turn == kProducerWarps) turn = 0;
                src += rowlen;
            }
        }
        SDBG(0, 9);
        STIM(1, stim_first); STIM(2, STIM_CLOCK()); STIM(3, stim_wait); STIM(4, stim_n0);
        return;
    }

    // =============================== consumers ===============================
    float* stage = stage_all + warp * kStageWordsPerWarp;
    const unsigned tbuf = smem_addr(stage + kFragBins * kStageStride);      // 16-byte aligned: 288 words precede it
    const unsigned lane_base = ring + (unsigned)(lane * SX) * 4u;
    const int bins = a.PH * a.PW;
    unsigned fphase = 0u;
    const int chsub = lane >> 3, fb = lane & 7;
    constexpr float kInvCount = 1.f / (float)(SR * SR);

    for (int L = L0; L < L1;) {
        const Item it = decode_item(L, L1, a, lane);
        L += it.yb - it.ya;
        SDBG(2, ((u64)it.ya << 48) | ((u64)it.yb << 32) | ((u64)it.yhi << 16) | (u64)(it.e1 - it.e0));
        if (it.e1 <= it.e0) continue;
        SDBG(0, 2);
        const unsigned lc = lane_base - (unsigned)(it.s * a.WX) * 4u;
        int acq = it.ya, rel = it.ya;
        int acq_slot = it.ya % a.K, rel_slot = acq_slot;
        const int c0 = it.g * 32;
        // every warp waits for every row of the item before it releases it, in stream order (keeps the per-slot
        // phase bits in step with the producers)
        auto acquire_to = [&](int row_end, unsigned tag) {
            while (acq < row_end) {
                STIM_DO(stim_n1 = STIM_CLOCK());
                mbar_wait(bars + 8u * acq_slot, (fphase >> acq_slot) & 1u, tag + (unsigned)acq,
                          WCTX(((u64)it.ya << 48) | ((u64)it.yb << 32) | ((u64)it.yhi << 16) | (u64)(it.lvl << 12) | (u64)(it.g << 8) | (u64)it.s),
                          WCTX(((u64)fphase << 32) | ((u64)(unsigned)rel << 16) | (u64)(unsigned)row_end));
                STIM_DO(stim_wait += STIM_CLOCK() - stim_n1);
                fphase ^= 1u << acq_slot;
                ++acq;
                if (++acq_slot == a.K) acq_slot = 0;
            }
        };
        auto release_to = [&](int row_end) {                  // rows [rel, row_end) have been acquired
            if (rel < row_end) {
                if (lane == 0) {
                    int rs = rel_slot;
                    for (int y = rel; y < row_end; ++y) { mbar_arrive(bars + 8u * (kMaxSlots + rs)); if (++rs == a.K) rs = 0; }
                }
                rel_slot += row_end - rel;                     // one wrap at most in the common case: no division
                while (rel_slot >= a.K) rel_slot -= a.K;
                rel = row_end;
            }
        };

        // Fragment pipeline (global-load latency is the enemy: entries and tables are needed by every lane at once):
        //   entry records are fetched two fragments ahead; the axis tables of the NEXT fragment are fetched one fragment
        //   ahead, lane-distributed (lanes 0..15: its x entries, lanes 16..17: its y entries -- one 16-byte load per lane),
        //   parked in registers, and dropped into this warp's table buffer in shared memory when their fragment starts;
        //   the bin loop then reads them with broadcast LDS.128.
        auto load_tab = [&](const uint2& en) -> uint4 {
            const int rn = en.x & 0xffffu, phn = (en.x >> 16) & 31u, pwn = (en.x >> 21) & 31u, nn = (int)((en.x >> 26) & 7u) + 1;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            const bool isx = lane < nn * SR;                   // one predicated load: the two tables differ in base and index only
            const uint4* base = isx ? a.xtab : a.ytab;
            const int idx = isx ? rn * a.nx + pwn * SR + lane : rn * a.ny + phn * SR + (lane - 16);
            if (isx || (lane >= 16 && lane < 16 + SR)) v = __ldg(base + idx);
            return v;
        };
        int e = it.e0 + warp;
        uint2 ent0 = make_uint2(0u, 0u), ent1 = make_uint2(0u, 0u);
        uint4 tab = make_uint4(0u, 0u, 0u, 0u);
        if (e < it.e1) { ent0 = __ldg(&a.entries[e]); tab = load_tab(ent0); }
        if (e + kConsumerWarps < it.e1) ent1 = __ldg(&a.entries[e + kConsumerWarps]);
        while (e < it.e1) {
            const uint2 cur = ent0;
            const int e_next = e + kConsumerWarps;
            // this fragment's tables -> shared (the previous fragment's readers are past their last LDS: trailing syncwarp)
            if (lane < 18) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tbuf + lane * 16), "r"(tab.x), "r"(tab.y), "r"(tab.z), "r"(tab.w) : "memory");
            __syncwarp();
            ent0 = ent1;
            if (e_next < it.e1) tab = load_tab(ent0);                               // in flight while this fragment computes
            if (e_next + kConsumerWarps < it.e1) ent1 = __ldg(&a.entries[e_next + kConsumerWarps]);
            const int r = cur.x & 0xffffu, ph = (cur.x >> 16) & 31u, pw0 = (cur.x >> 21) & 31u;
            const int npw = (int)((cur.x >> 26) & 7u) + 1;
            const bool red = (cur.x >> 29) & 1u;
            const int key = cur.y & 0xffffu, end = key + (int)((cur.y >> 16) & 0xffu);
            const unsigned smask = (cur.y >> 24) & 0xfu;
            // entries are sorted by key: rows below it are done with.  Rows this warp never needed are passed through one
            // at a time (waited for, then released at once): holding them while waiting for later rows would starve the
            // producers of slots whenever this warp's consecutive fragments lie more than K rows apart.
            release_to(min(acq, key));
            while (acq < key) {
                acquire_to(acq + 1, 0x2000u);
                release_to(acq);
            }
            acquire_to(end, 0x2800u);                           // rows [key, end) must be resident
            SDBG(3, ((u64)cur.x << 32) | (u64)cur.y);
            STIM_DO(++stim_n0; stim_c0 = STIM_CLOCK(); if (!stim_first) stim_first = stim_c0);
            const uint4 yA = lds128(tbuf + 16 * 16);
            uint4 yB = yA;
            if (SR == 2) yB = lds128(tbuf + 17 * 16);
            // row bases for this lane and the packed axis weights (uniform; hoisted out of the bin loop)
            const unsigned rt0 = lc + yA.z, rt1 = lc + yB.z;
            const u64 hh0 = pack2f(__uint_as_float(yA.x), __uint_as_float(yA.x)), ll0 = pack2f(__uint_as_float(yA.y), __uint_as_float(yA.y));
            const u64 hh1 = pack2f(__uint_as_float(yB.x), __uint_as_float(yB.x)), ll1 = pack2f(__uint_as_float(yB.y), __uint_as_float(yB.y));
            unsigned tb = tbuf;                                   // this bin's x entries (broadcast reads)
            unsigned sp = smem_addr(stage) + (unsigned)lane * 4u; // where this lane parks the bin's result
            const unsigned tb_end = tbuf + (unsigned)npw * (SR * 16u);
            if (SR == 2 && smask == 0xfu) {
                // whole bins (the common case): all 16 taps are requested before the first one is used
                do {
                    const uint4 cxA = lds128(tb), cxB = lds128(tb + 16);
                    const u64 wxA = pack2f(__uint_as_float(cxA.x), __uint_as_float(cxA.y));
                    const u64 wxB = pack2f(__uint_as_float(cxB.x), __uint_as_float(cxB.y));
                    const Taps t0 = load_taps<RB>(rt0 + cxA.z);
                    const Taps t1 = load_taps<RB>(rt0 + cxB.z);
                    const Taps t2 = load_taps<RB>(rt1 + cxA.z);
                    const Taps t3 = load_taps<RB>(rt1 + cxB.z);
                    float acc = __fadd_rn(0.f, sample_val(t0, hh0, ll0, wxA));   // the reference's order: iy outer, ix inner
                    acc = __fadd_rn(acc, sample_val(t1, hh0, ll0, wxB));
                    acc = __fadd_rn(acc, sample_val(t2, hh1, ll1, wxA));
                    acc = __fadd_rn(acc, sample_val(t3, hh1, ll1, wxB));
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sp), "f"(__fmul_rn(acc, kInvCount)) : "memory");   // count 4: exact
                    tb += 32u; sp += kStageStride * 4u;
                } while (tb != tb_end);
            } else {
                do {
                    const uint4 cxA = lds128(tb);
                    uint4 cxB = cxA;
                    if (SR == 2) cxB = lds128(tb + 16);
                    const u64 wxA = pack2f(__uint_as_float(cxA.x), __uint_as_float(cxA.y));
                    const u64 wxB = pack2f(__uint_as_float(cxB.x), __uint_as_float(cxB.y));
                    float acc = 0.f;
                    if (SR == 1) {
                        acc = __fadd_rn(acc, sample_val(load_taps<RB>(rt0 + cxA.z), hh0, ll0, wxA));
                    } else {
                        if (smask & 1u) acc = __fadd_rn(acc, sample_val(load_taps<RB>(rt0 + cxA.z), hh0, ll0, wxA));
                        if (smask & 2u) acc = __fadd_rn(acc, sample_val(load_taps<RB>(rt0 + cxB.z), hh0, ll0, wxB));
                        if (smask & 4u) acc = __fadd_rn(acc, sample_val(load_taps<RB>(rt1 + cxA.z), hh1, ll1, wxA));
                        if (smask & 8u) acc = __fadd_rn(acc, sample_val(load_taps<RB>(rt1 + cxB.z), hh1, ll1, wxB));
                    }
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sp), "f"(__fmul_rn(acc, kInvCount)) : "memory");   // count 1 / 4: exact
                    tb += SR * 16u; sp += kStageStride * 4u;
                } while (tb != tb_end);
            }
            __syncwarp();
            // flush: lanes = (channel within a group of 4, bin) -> a store's lanes are consecutive bins of one channel
            if (fb < npw) {
                const int orow = a.row_map ? __ldg(&a.row_map[r]) : r;
                float* dst = a.out + ((size_t)orow * a.C + c0) * bins + ph * a.PW + pw0 + fb;
                const float* src = stage + fb * kStageStride + chsub;
                const int cmax = a.C - c0 - chsub;               // channel 4*cb + chsub is real iff 4*cb < cmax
                if (!red && cmax + chsub >= 32) {                        // a full channel group: no predicates, one pointer walk
                    float* d = dst + (size_t)chsub * bins;
                    const size_t step = (size_t)bins * 4;
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb) { *d = src[4 * cb]; d += step; }
                } else if (!red) {
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb)
                        if (4 * cb < cmax) dst[(size_t)(4 * cb + chsub) * bins] = src[4 * cb];
                } else {
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb)
                        if (4 * cb < cmax) atomicAdd(dst + (size_t)(4 * cb + chsub) * bins, src[4 * cb]);
                }
            }
            __syncwarp();
            STIM_DO(stim_comp += STIM_CLOCK() - stim_c0);
            e = e_next;
        }
        // item tail: nothing more to read -- give back what is held, then pass the remaining rows through one by one
        // (a warp that held rows while waiting for later ones could starve the producers of slots)
        SDBG(0, 5);
        release_to(acq);
        while (acq < it.yhi) {
            acquire_to(acq + 1, 0x3000u);
            release_to(acq);
        }
        __syncwarp();
        SDBG(0, 6);
    }
    SDBG(0, 9);
    STIM(1, stim_first); STIM(2, STIM_CLOCK()); STIM(3, stim_wait); STIM(4, stim_n0); STIM(5, stim_comp);
}

// ------------------------------------------------------------------------------------------------
// host: geometry, workspace, tensor map, launches
// ------------------------------------------------------------------------------------------------
size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct StreamLayout {
    size_t ytab_off, xtab_off, zero_off, zero_bytes, rowptr_off, cpre_off, piece_off, entries_off, ws_bytes;
};

// One launch serves `levels` feature maps (FPN: the pyramid levels of one head; a plain call: one map).  All levels share the
// batch size, the channel count, the pooled size, the ring geometry (SX, K) and therefore the kernel instantiation.
// level_roi_begin: [levels + 1] first RoI of every level (RoIs are stored level-major), level_roi_begin[levels] = R.
bool stream_geometry(int levels, const float* const* bottoms, const int* heights, const int* widths, const float* scales,
                     const int* level_roi_begin, int N, int R, int C, int PH, int PW, int sr, int sm_count, StreamGeom* g,
                     StreamLayout* lay, unsigned* smem_bytes) {
    if (levels < 1 || levels > kMaxLevels) return false;
    if (sr < 1 || sr > 2 || PH * sr > kAxisMaxS || PW * sr > kAxisMaxS || PH > 31 || PW > 31) return false;
    if (N <= 0 || R <= 0 || R > 65535) return false;
    if ((long long)R * PH * PW * sr * sr >= (1LL << 28)) return false;
    for (int l = 0; l < levels; ++l)
        if (heights[l] < 2 || widths[l] < 2 || heights[l] > 65535) return false;
    // ring budget: everything but the per-warp staging, the barriers and the alignment slack
    const unsigned fixed = kConsumerWarps * kStageWordsPerWarp * 4 + 2 * kMaxSlots * 8 + 128;
    const unsigned ring_budget = kSmemBudget - fixed;
    int best_sx = 0, best_wx = 0, best_k = 0;
    long best_score = -1;
    for (int m = 1; m <= 3; ++m) {                            // SX = 32 m + 1: odd channel stride, m coalesced chunks + 1 column
        const int sx = 32 * m + 1;
        int k = (int)(ring_budget / (128u * (unsigned)sx)) - 1;       // one physical slot is the mirror of slot 0
        if (k > kMaxSlots) k = kMaxSlots;
        if (k < 12) continue;
        const int hx = m >= 2 ? 9 : 5;                        // halo: bins whose x samples are <= hx + 1 cells apart stay whole
        const int wx = sx - 1 - hx;
        long score = 0;                                       // columns staged per row of every level, weighted by its rows
        for (int l = 0; l < levels; ++l) {
            int s = 1;
            while ((s - 1) * wx + sx < widths[l]) ++s;        // the last strip needs no halo
            score += (long)s * sx * heights[l];
        }
        if (best_score < 0 || score < best_score) { best_score = score; best_sx = sx; best_wx = wx; best_k = k; }
    }
    if (best_score < 0) return false;
    g->N = N; g->R = R; g->C = C; g->PH = PH; g->PW = PW; g->sr = sr;
    g->ny = PH * sr; g->nx = PW * sr;
    g->SX = best_sx; g->WX = best_wx; g->K = best_k;
    g->row_bytes = 128 * best_sx;
    g->L = levels;
    int q = 0;
    long long keys = 0;
    for (int l = 0; l < levels; ++l) {
        StreamLevel& lv = g->lv[l];
        lv.bottom = bottoms ? bottoms[l] : nullptr;
        lv.scale = scales ? scales[l] : 1.f;
        lv.H = heights[l]; lv.W = widths[l];
        int s = 1;
        while ((s - 1) * best_wx + best_sx < lv.W) ++s;
        lv.S = s;
        lv.q_base = q;
        lv.roi_begin = level_roi_begin ? level_roi_begin[l] : 0;
        lv.pad = 0;
        if (q + N * s > kMaxCols) return false;
        for (int c = 0; c < N * s; ++c) { g->colstart[q + c] = (int)keys; keys += lv.H; }
        q += N * s;
        if (keys >= (1LL << 24)) return false;
    }
    for (int l = levels; l < kMaxLevels; ++l) { g->lv[l] = g->lv[levels - 1]; g->lv[l].q_base = q; g->lv[l].roi_begin = R; }
    g->Q = q; g->keys = (int)keys;
    for (int c = q; c <= kMaxCols; ++c) g->colstart[c] = (int)keys;
    g->G = (C + 31) / 32;
    if (g->G < 1) g->G = 1;
    if (keys * g->G >= (1LL << 30)) return false;
    g->pieces = sm_count > 0 ? sm_count : kNumSMs;
    g->max_entries = R * PH * PW * sr * sr;                    // worst case: every sample its own fragment
    size_t off = 0;
    lay->ytab_off = off; off = align_up_sz(off + (size_t)R * g->ny * 16, 256);
    lay->xtab_off = off; off = align_up_sz(off + (size_t)R * g->nx * 16, 256);
    lay->zero_off = off; lay->zero_bytes = align_up_sz(((size_t)4 * g->keys + 4) * 4, 256); off += lay->zero_bytes;
    lay->rowptr_off = off; off = align_up_sz(off + ((size_t)g->keys + 1) * 4, 256);
    lay->cpre_off = off; off = align_up_sz(off + ((size_t)g->keys + 1) * 4, 256);
    lay->piece_off = off; off = align_up_sz(off + ((size_t)g->pieces + 1) * 4, 256);
    lay->entries_off = off; off = align_up_sz(off + (size_t)g->max_entries * 8, 256);
    lay->ws_bytes = off;
    *smem_bytes = (unsigned)(g->K + 1) * (unsigned)g->row_bytes + fixed;
    return true;
}

struct DeviceInfo {
    bool ok = false;
    int sm_count = 0;
};

// per device, set up once under a lock (the reference calls these ops from one host thread per GPU)
bool stream_device_info(int* sm_count) {
    static std::mutex mu;
    static DeviceInfo info[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    std::lock_guard<std::mutex> lock(mu);
    if (!info[dev].ok) {
        const int max_dyn = (int)kSmemBudget;
        bool ok = cudaDeviceGetAttribute(&info[dev].sm_count, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
#define B200_STREAM_ATTR(SRV, MV, AV) \
        ok = ok && cudaFuncSetAttribute(roi_align_stream_fwd<SRV, MV, AV>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) == cudaSuccess
        B200_STREAM_ATTR(1, 1, true); B200_STREAM_ATTR(1, 2, true); B200_STREAM_ATTR(1, 3, true);
        B200_STREAM_ATTR(2, 1, true); B200_STREAM_ATTR(2, 2, true); B200_STREAM_ATTR(2, 3, true);
        B200_STREAM_ATTR(1, 1, false); B200_STREAM_ATTR(1, 2, false); B200_STREAM_ATTR(1, 3, false);
        B200_STREAM_ATTR(2, 1, false); B200_STREAM_ATTR(2, 2, false); B200_STREAM_ATTR(2, 3, false);
#undef B200_STREAM_ATTR
        if (!ok) {
            (void)cudaGetLastError();
            return false;
        }
        info[dev].ok = true;
    }
    *sm_count = info[dev].sm_count;
    return true;
}

}  // namespace

void roi_align_stream_set_debug_buffer(unsigned long long* host_pinned) {
#if defined(B200_STREAM_DEBUG)
    cudaMemcpyToSymbol(g_stream_dbg, &host_pinned, sizeof(host_pinned));
#elif defined(B200_STREAM_TIMING)
    cudaMemcpyToSymbol(g_stream_tim, &host_pinned, sizeof(host_pinned));
#else
    cudaMemcpyToSymbol(g_stream_dead, &host_pinned, sizeof(host_pinned));      // [CTA][32] u64 watchdog records
#endif
}

size_t roi_align_stream_fpn_workspace_bytes(int levels, const int* heights, const int* widths, int N, int R, int PH, int PW, int sr) {
    StreamGeom g;
    StreamLayout lay;
    unsigned smem = 0;
    if (!stream_geometry(levels, nullptr, heights, widths, nullptr, nullptr, N, R, 32, PH, PW, sr, kNumSMs, &g, &lay, &smem)) return 0;
    return lay.ws_bytes;
}

size_t roi_align_stream_workspace_bytes(int N, int R, int H, int W, int PH, int PW, int sr) {
    return roi_align_stream_fpn_workspace_bytes(1, &H, &W, N, R, PH, PW, sr);
}

// returns B200_ROI_OK when the streaming path ran; 1000 when it does not apply (caller falls back)
int roi_align_forward_stream_fpn(int levels, const float* const* bottoms, const int* heights, const int* widths, const float* scales,
                                 const int* level_roi_begin, int N, int R, int C, int PH, int PW, int sr, const float* rois, float* top,
                                 const int* row_map, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (workspace == nullptr || C <= 0) return 1000;
    if ((long long)N * C >= (1LL << 31) || (long long)R * C * PH * PW >= (1LL << 31)) return 1000;
    int sm_count = 0;
    if (!stream_device_info(&sm_count)) return 1000;
    StreamGeom g;
    StreamLayout lay;
    unsigned smem = 0;
    if (!stream_geometry(levels, bottoms, heights, widths, scales, level_roi_begin, N, R, C, PH, PW, sr, sm_count, &g, &lay, &smem)) return 1000;
    if (workspace_bytes < lay.ws_bytes) return 1000;
    unsigned char* wsb = (unsigned char*)workspace;
    StreamWs ws;
    ws.ytab = (uint4*)(wsb + lay.ytab_off);
    ws.xtab = (uint4*)(wsb + lay.xtab_off);
    int* zero = (int*)(wsb + lay.zero_off);
    ws.hist = zero; ws.cost = zero + g.keys; ws.cursor = zero + 2 * (size_t)g.keys; ws.maxend = zero + 3 * (size_t)g.keys;
    ws.ticket = zero + 4 * (size_t)g.keys;
    ws.rowptr = (int*)(wsb + lay.rowptr_off);
    ws.cpre = (unsigned*)(wsb + lay.cpre_off);
    ws.piece_start = (int*)(wsb + lay.piece_off);
    ws.entries = (uint2*)(wsb + lay.entries_off);

    cudaError_t err = cudaMemsetAsync(zero, 0, lay.zero_bytes, stream);
    if (err != cudaSuccess) return (int)err;
    StreamArgs a;
    a.ytab = ws.ytab; a.xtab = ws.xtab; a.entries = ws.entries; a.rowptr = ws.rowptr; a.maxend = ws.maxend;
    a.piece_start = ws.piece_start; a.out = top; a.row_map = row_map;
    a.C = C; a.G = g.G; a.K = g.K; a.WX = g.WX; a.PH = PH; a.PW = PW; a.ny = g.ny; a.nx = g.nx; a.L = g.L; a.Q = g.Q;
    for (int l = 0; l < kMaxLevels; ++l) a.lv[l] = g.lv[l];
    for (int c = 0; c <= kMaxCols; ++c) a.colstart[c] = g.colstart[c];
    const bool async = option_get(kOptStreamStage) != 'r';          // B200_STREAM_STAGE=regs selects LDG -> registers -> STS (A/B)
    const int m = (g.SX - 1) / 32;
    if (sr == 1) {
        stream_count<1><<<R, kPrepThreads, 0, stream>>>(rois, g, ws);
        stream_fill<1><<<R, kPrepThreads, 0, stream>>>(rois, g, ws, top, row_map);
    } else {
        stream_count<2><<<R, kPrepThreads, 0, stream>>>(rois, g, ws);
        stream_fill<2><<<R, kPrepThreads, 0, stream>>>(rois, g, ws, top, row_map);
    }
#define B200_STREAM_LAUNCH(SRV, MV, AV) roi_align_stream_fwd<SRV, MV, AV><<<g.pieces, kStreamThreads, smem, stream>>>(a)
#define B200_STREAM_LAUNCH_M(SRV, AV) (m == 1 ? B200_STREAM_LAUNCH(SRV, 1, AV) : m == 2 ? B200_STREAM_LAUNCH(SRV, 2, AV) : B200_STREAM_LAUNCH(SRV, 3, AV))
    if (option_get(kOptStreamPhases) == 'p') return finish_launch(2);          // timing probe: prepass only (output undefined)
    if (sr == 1) { if (async) B200_STREAM_LAUNCH_M(1, true); else B200_STREAM_LAUNCH_M(1, false); }
    else         { if (async) B200_STREAM_LAUNCH_M(2, true); else B200_STREAM_LAUNCH_M(2, false); }
#undef B200_STREAM_LAUNCH_M
#undef B200_STREAM_LAUNCH
    return finish_launch(3);
}

int roi_align_forward_stream(const float* bottom, float scale, int N, int R, int H, int W, int C, int PH, int PW, int sr,
                             const float* rois, float* top, const int* row_map, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int begin[2] = {0, R};
    return roi_align_forward_stream_fpn(1, &bottom, &H, &W, &scale, begin, N, R, C, PH, PW, sr, rois, top, row_map, workspace,
                                        workspace_bytes, stream);
}

}  // namespace b200
