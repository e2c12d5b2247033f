// rm_gridbuild.cu — AccGrid::build_from_mesh (acc_grid.rs:36-83) with the cell lists built on the GPU.
//
// Same result as the host build (rm_host.cpp build_grid), bit for bit: the resolution and cell size come from the
// same host function (grid_dims); every triangle's cell range is the reference's arithmetic (f64, no FMA:
// (bounds - mesh.min) / cell_size, cast::<usize>(), clamp to res - 1, acc_grid.rs:43-56); the index is the
// reference's x + res.x * (y + z * res.z) (sic, :61); and inside a cell the triangle indices ascend, which is the
// order the reference's sequential push produces.  count -> exclusive scan -> fill (atomic cursors) -> per-cell sort.
// Failures map to the same statuses: a failed cast -> RM_ERR_GRID_CAST, an index beyond cells.len() ->
// RM_ERR_GRID_INDEX_OOB naming the first (lowest) offending triangle, like the reference's panic would.

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rm_internal.hpp"

namespace rm {

namespace {

constexpr int kGB = 256;
constexpr unsigned kScanItems = 8;                       // cells per thread in the scan kernels
constexpr unsigned kScanTile = kGB * kScanItems;

struct BuildParams {
    double bmin[3], cell[3];
    unsigned res[3];
    unsigned n_tris;
    unsigned long long n_cells;
};

struct BuildFlags {
    unsigned cast_failed;        // lowest triangle whose cell range failed cast::<usize>()
    unsigned oob;                // lowest triangle that touches an index >= cells.len()
};

// Triangle::find_bounds (triangle.rs:70-84) + the cell range of acc_grid.rs:43-56.  false = the cast failed.
__device__ __forceinline__ bool cell_range(const BuildParams& p, const double* __restrict__ pos, unsigned lo[3], unsigned hi[3]) {
    const double seed_min[3] = {125125.0, 1251251.0, 12512512.0};
    const double seed_max[3] = {-123125.0, -125123.0, -512123.0};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double mn = seed_min[a], mx = seed_max[a];
#pragma unroll
        for (int k = 0; k < 3; k++) { mn = fmin(mn, pos[3 * k + a]); mx = fmax(mx, pos[3 * k + a]); }
        const double ql = (mn - p.bmin[a]) / p.cell[a], qh = (mx - p.bmin[a]) / p.cell[a];
        // cgmath cast::<usize>(): Some(trunc toward zero) iff -1 < v < 2^64 (NaN fails)
        if (!(ql > -1.0 && ql < 18446744073709551616.0) || !(qh > -1.0 && qh < 18446744073709551616.0)) return false;
        const unsigned long long l = ql <= 0.0 ? 0ull : __double2ull_rz(ql), h = qh <= 0.0 ? 0ull : __double2ull_rz(qh);
        lo[a] = (unsigned)min(l, (unsigned long long)(p.res[a] - 1u));
        hi[a] = (unsigned)min(h, (unsigned long long)(p.res[a] - 1u));
    }
    return true;
}

// One thread per triangle: cell range, then ++count[cell] for every covered cell.  FILL = false counts,
// FILL = true writes the triangle index at start[cell] + cursor[cell]++.
template <bool FILL>
__global__ void __launch_bounds__(kGB) k_grid_scatter(const __grid_constant__ BuildParams p, const double* __restrict__ pos, unsigned* __restrict__ count,
                                                       const unsigned* __restrict__ start, unsigned* __restrict__ refs, BuildFlags* flags) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_tris) return;
    unsigned lo[3], hi[3];
    if (!cell_range(p, pos + (size_t)i * 9, lo, hi)) { if (!FILL) atomicMin(&flags->cast_failed, i); return; }
    const unsigned long long rx = p.res[0], rz = p.res[2];
    for (unsigned z = lo[2]; z <= hi[2]; z++)
        for (unsigned y = lo[1]; y <= hi[1]; y++) {
            const unsigned long long row = rx * ((unsigned long long)y + (unsigned long long)z * rz);     // sic: res.z (acc_grid.rs:61)
            if (row + hi[0] >= p.n_cells) { if (!FILL) atomicMin(&flags->oob, i); return; }
            for (unsigned x = lo[0]; x <= hi[0]; x++) {
                if (FILL) refs[start[row + x] + atomicAdd(&count[row + x], 1u)] = i;
                else atomicAdd(&count[row + x], 1u);
            }
        }
}

// Exclusive scan of `count` (n entries) into `start` (n + 1 entries), three kernels: tile sums, scan of the tile sums
// (one block), tile-local scan + offset.  Totals are carried in 64 bits so that >= 2^32 references is detected.
__global__ void __launch_bounds__(kGB) k_scan_tile_sums(const unsigned* __restrict__ count, unsigned long long n, unsigned long long* __restrict__ tile_sum) {
    __shared__ unsigned long long warp_sum[kGB / 32];
    const unsigned long long base = (unsigned long long)blockIdx.x * kScanTile + (unsigned long long)threadIdx.x * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (unsigned j = 0; j < kScanItems; j++) if (base + j < n) s += count[base + j];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kGB / 32; w++) t += warp_sum[w];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_scan_tiles(unsigned long long* __restrict__ tile_sum, unsigned n_tiles, unsigned long long* __restrict__ total) {
    // one block: sequential over chunks of 1024 tiles with a running carry
    __shared__ unsigned long long buf[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned c = 0; c < n_tiles; c += 1024) {
        const unsigned i = c + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_sum[i] : 0ull;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (unsigned o = 1; o < 1024; o <<= 1) {              // Hillis-Steele inclusive scan
            const unsigned long long add = threadIdx.x >= o ? buf[threadIdx.x - o] : 0ull;
            __syncthreads();
            buf[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < n_tiles) tile_sum[i] = carry + buf[threadIdx.x] - v;      // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kGB) k_scan_apply(const unsigned* __restrict__ count, unsigned long long n, const unsigned long long* __restrict__ tile_off,
                                                     unsigned* __restrict__ start) {
    __shared__ unsigned warp_sum[kGB / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long base = (unsigned long long)blockIdx.x * kScanTile + (unsigned long long)threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned s = 0;
#pragma unroll
    for (unsigned j = 0; j < kScanItems; j++) { v[j] = base + j < n ? count[base + j] : 0u; s += v[j]; }
    unsigned incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (unsigned w = 0; w < warp; w++) woff += warp_sum[w];
    unsigned run = (unsigned)tile_off[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (unsigned j = 0; j < kScanItems; j++) {
        if (base + j < n) {
            start[base + j] = run;
            run += v[j];
            if (base + j + 1 == n) start[n] = run;      // the closing entry: total references
        }
    }
}

// One thread per cell: ascending triangle order inside the cell (the atomic cursors fill in arbitrary order).
__global__ void __launch_bounds__(kGB) k_sort_cells(const unsigned* __restrict__ start, unsigned long long n_cells, unsigned* __restrict__ refs) {
    const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const unsigned b = start[c], e = start[c + 1];
    const unsigned m = e - b;
    if (m < 2) return;
    unsigned* a = refs + b;
    if (m <= 48) {
        for (unsigned i = 1; i < m; i++) {
            const unsigned v = a[i];
            unsigned j = i;
            while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; j--; }
            a[j] = v;
        }
    } else {
        // heap sort: O(m log m) in place for the rare long list
        auto sift = [&](unsigned root, unsigned end) {
            for (;;) {
                unsigned child = 2 * root + 1;
                if (child >= end) break;
                if (child + 1 < end && a[child] < a[child + 1]) child++;
                if (a[root] >= a[child]) break;
                const unsigned t = a[root]; a[root] = a[child]; a[child] = t;
                root = child;
            }
        };
        for (unsigned i = m / 2; i-- > 0;) sift(i, m);
        for (unsigned end = m - 1; end > 0; end--) {
            const unsigned t = a[0]; a[0] = a[end]; a[end] = t;
            sift(0, end);
        }
    }
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { dev_release(p); }
    int alloc(size_t bytes) {
        const int e = dev_alloc(&p, bytes);
        if (e != 0) return fail(RM_ERR_CUDA, std::string("device allocation of ") + std::to_string(bytes) + " bytes: " + cudaGetErrorString((cudaError_t)e));
        return RM_OK;
    }
};

#define RM_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(RM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
    } while (0)

}  // namespace

int build_grid_device(std::vector<rm_triangle>&& tris, const rm_aabb& bounds, int device, std::shared_ptr<Grid>* out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(RM_ERR_CUDA, "no usable CUDA device " + std::to_string(device) + " for the device grid build (use rm_grid_build for the host build)");
    }
    RM_CUDA(cudaSetDevice(device));
    auto g = std::make_shared<Grid>();
    const size_t n = tris.size();
    uint64_t res[3];
    double cell[3];
    if (int st = grid_dims(bounds, n, res, cell)) return st;
    const uint64_t n_cells = res[0] * res[1] * res[2];
    BuildParams p{};
    p.bmin[0] = bounds.min.x; p.bmin[1] = bounds.min.y; p.bmin[2] = bounds.min.z;
    for (int a = 0; a < 3; a++) { p.cell[a] = cell[a]; p.res[a] = (unsigned)res[a]; }
    p.n_tris = (unsigned)n;
    p.n_cells = n_cells;

    // positions -> pinned staging -> device (72 B per triangle)
    const size_t pos_bytes = std::max<size_t>(n * 9 * sizeof(double), 64);
    double* hpos = (double*)pinned_acquire(pos_bytes);
    if (!hpos) return fail(RM_ERR_OUT_OF_MEMORY, "cannot pin the position staging buffer");
    {
        auto pack = [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) {
                const rm_vertex* v[3] = {&tris[i].v0, &tris[i].v1, &tris[i].v2};
                for (int k = 0; k < 3; k++) { hpos[i * 9 + 3 * k] = v[k]->position.x; hpos[i * 9 + 3 * k + 1] = v[k]->position.y; hpos[i * 9 + 3 * k + 2] = v[k]->position.z; }
            }
        };
        const unsigned hw = std::thread::hardware_concurrency();
        const size_t workers = n < ((size_t)1 << 16) ? 1 : std::min<size_t>(hw ? hw : 4, 8);
        std::vector<std::thread> pool;
        for (size_t w = 1; w < workers; w++) pool.emplace_back(pack, n * w / workers, n * (w + 1) / workers);
        pack(0, n / workers);
        for (std::thread& t : pool) t.join();
    }
    struct PinGuard { void* p; ~PinGuard() { pinned_release(p); } } pin_guard{hpos};

    const unsigned n_tiles = (unsigned)((n_cells + kScanTile - 1) / kScanTile);
    DevBuf d_pos, d_count, d_start, d_tiles, d_flags, d_refs;
    if (int st = d_pos.alloc(pos_bytes)) return st;
    if (int st = d_count.alloc(n_cells * sizeof(unsigned))) return st;
    if (int st = d_start.alloc((n_cells + 1) * sizeof(unsigned))) return st;
    if (int st = d_tiles.alloc(((size_t)n_tiles + 1) * sizeof(unsigned long long))) return st;
    if (int st = d_flags.alloc(sizeof(BuildFlags))) return st;
    cudaStream_t s = 0;
    RM_CUDA(cudaMemcpyAsync(d_pos.p, hpos, n * 9 * sizeof(double), cudaMemcpyHostToDevice, s));
    RM_CUDA(cudaMemsetAsync(d_count.p, 0, n_cells * sizeof(unsigned), s));
    RM_CUDA(cudaMemsetAsync(d_flags.p, 0xff, sizeof(BuildFlags), s));
    const unsigned tri_blocks = (unsigned)((n + kGB - 1) / kGB);
    if (n) k_grid_scatter<false><<<tri_blocks, kGB, 0, s>>>(p, (const double*)d_pos.p, (unsigned*)d_count.p, nullptr, nullptr, (BuildFlags*)d_flags.p);
    k_scan_tile_sums<<<n_tiles, kGB, 0, s>>>((const unsigned*)d_count.p, n_cells, (unsigned long long*)d_tiles.p);
    k_scan_tiles<<<1, 1024, 0, s>>>((unsigned long long*)d_tiles.p, n_tiles, (unsigned long long*)d_tiles.p + n_tiles);
    BuildFlags flags{};
    unsigned long long total = 0;
    RM_CUDA(cudaMemcpyAsync(&flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, s));
    RM_CUDA(cudaMemcpyAsync(&total, (unsigned long long*)d_tiles.p + n_tiles, sizeof(total), cudaMemcpyDeviceToHost, s));
    RM_CUDA(cudaStreamSynchronize(s));
    RM_CUDA(cudaGetLastError());
    // the reference fails at the FIRST offending triangle in mesh order, whichever kind of failure that is
    if (flags.cast_failed != 0xffffffffu && flags.cast_failed <= flags.oob)
        return fail(RM_ERR_GRID_CAST, "Failed to cast cell bounds to usize (acc_grid.rs:47,51)");
    if (flags.oob != 0xffffffffu)
        return fail(RM_ERR_GRID_INDEX_OOB, "cell index out of bounds while inserting triangle " + std::to_string(flags.oob) + " (reference panics at acc_grid.rs:61)");
    if (total >= 0xffffffffull) return fail(RM_ERR_UNSUPPORTED, "grid has 2^32 or more triangle references");

    if (int st = d_refs.alloc(std::max<size_t>(total, 1) * sizeof(unsigned))) return st;
    k_scan_apply<<<n_tiles, kGB, 0, s>>>((const unsigned*)d_count.p, n_cells, (const unsigned long long*)d_tiles.p, (unsigned*)d_start.p);
    RM_CUDA(cudaMemsetAsync(d_count.p, 0, n_cells * sizeof(unsigned), s));       // the counts become the fill cursors
    if (n) k_grid_scatter<true><<<tri_blocks, kGB, 0, s>>>(p, (const double*)d_pos.p, (unsigned*)d_count.p, (const unsigned*)d_start.p, (unsigned*)d_refs.p, (BuildFlags*)d_flags.p);
    k_sort_cells<<<(unsigned)((n_cells + kGB - 1) / kGB), kGB, 0, s>>>((const unsigned*)d_start.p, n_cells, (unsigned*)d_refs.p);
    RM_CUDA(cudaGetLastError());

    g->cell_start.resize(n_cells + 1);
    g->references.resize(total);
    RM_CUDA(cudaMemcpyAsync(g->cell_start.data(), d_start.p, (n_cells + 1) * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    if (total) RM_CUDA(cudaMemcpyAsync(g->references.data(), d_refs.p, total * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    RM_CUDA(cudaStreamSynchronize(s));
    g->triangles = std::move(tris);
    g->bounds = bounds;
    for (int a = 0; a < 3; a++) g->resolution[a] = res[a];
    g->cell_size = rm_vec3{cell[0], cell[1], cell[2]};
    *out = g;
    return RM_OK;
}

}  // namespace rm
