// Bucket partition of a point stream (first half of the accumulate for inputs in no particular order).
//
// The accumulate kernels keep one knot interval (1-D) or one cell (2-D) per warp in registers, which is free when
// consecutive points share it (time-series / raster order) and degenerates to one fp64 RED per band entry per point
// when they do not (measured 11.4 ms instead of 0.27 ms for 1e8 shuffled 1-D points, 35 ms instead of 0.65 ms in
// 2-D).  For such inputs the points are first partitioned into <= 256 buckets of consecutive knot intervals:
//
//   part_hist_kernel      one read of the key coordinate, per-CTA shared-memory histogram, one global add per
//                         (CTA, bucket);
//   part_scan_kernel      exclusive scan (one CTA): bucket starts, write cursors, and the unit table of the second
//                         half (a unit = at most kUnitPoints consecutive points of ONE bucket);
//   part_scatter_kernel   tiles of kPartTile points: bucket + rank within the tile by shared-memory integer atomics,
//                         one global cursor add per (tile, bucket), the tile is sorted by bucket in shared memory and
//                         leaves as runs of consecutive records (component-major output, 8 B per component).
//
// The second half (accum_1d_units_kernel / accum_2d_units_kernel) sorts each unit by interval / cell in shared
// memory and accumulates each run in registers, so the REDs drop from ~14 per point to ~14 per (unit, interval).
// HBM traffic: key read + (REC read + REC write) + REC read = (1 + 3 REC) * 8 B per point (56 B for REC = 2).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace asvgp {

constexpr int kPartBuckets = 256;     // upper bound on the number of buckets
constexpr int kPartThreads = 256;
constexpr int kPartTile = 4096;       // points per tile of the scatter pass (16 per thread)
constexpr int kUnitPoints = 4096;     // largest unit of the second half (each kernel passes the unit size it uses)
constexpr int kUnitMaxBins = 4096;    // intervals (cells) of one bucket the second half can sort in shared memory

typedef unsigned long long u64;

// [ count[NB] | start[NB + 1] | cursor[NB] | unit_start[NB + 1] ] as u64, then the records (component-major)
struct PartWork {
    u64* count;
    u64* start;
    u64* cursor;
    u64* unit_start;
    double* rec;
    __host__ __device__ static int64_t head_words() { return 4 * (int64_t)kPartBuckets + 4; }   // + max units per bucket, + pad
    __host__ __device__ static int64_t bytes(int64_t n, int rec_doubles) {
        return (head_words() + n * rec_doubles) * 8;
    }
    __host__ __device__ static PartWork carve(void* work) {
        PartWork w;
        w.count = static_cast<u64*>(work);
        w.start = w.count + kPartBuckets;
        w.cursor = w.start + kPartBuckets + 1;
        w.unit_start = w.cursor + kPartBuckets;
        w.rec = reinterpret_cast<double*>(w.count + head_words());
        return w;
    }
};

// Src: struct with  __device__ void init()  (device-side set-up, e.g. reading the mesh origin),
// __device__ int bucket(int64_t i) const  (reads only the key coordinate),  __device__ void load(int64_t i, double (&v)[REC]) const
// and  __device__ int bucket_of(const double (&v)[REC]) const.  Buckets only have to GROUP nearby points, so they come
// from the uniform-grid guess of the interval (no gather from the knot array); the exact interval (reference
// basis.py:58) is found in the second half.
template <class Src>
__global__ void __launch_bounds__(kPartThreads) part_hist_kernel(Src src, int64_t n, u64* __restrict__ count) {
    __shared__ int s_cnt[kPartBuckets];
    src.init();
    for (int b = threadIdx.x; b < kPartBuckets; b += kPartThreads) s_cnt[b] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kPartThreads;
    for (int64_t i = blockIdx.x * (int64_t)kPartThreads + threadIdx.x; i < n; i += 4 * stride) {
        int bk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) bk[q] = (i + q * stride < n) ? src.bucket(i + q * stride) : -1;      // four loads in flight
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (bk[q] >= 0) atomicAdd(&s_cnt[bk[q]], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kPartBuckets; b += kPartThreads)
        if (s_cnt[b]) atomicAdd(count + b, (u64)s_cnt[b]);
}

static __global__ void __launch_bounds__(kPartBuckets) part_scan_kernel(PartWork w, int unit_points) {
    __shared__ u64 s_a[kPartBuckets], s_u[kPartBuckets];
    const int b = threadIdx.x;
    const u64 c = w.count[b];
    s_a[b] = c;
    s_u[b] = (c + unit_points - 1) / unit_points;
    __syncthreads();
    for (int o = 1; o < kPartBuckets; o <<= 1) {          // Hillis-Steele inclusive scan of both columns
        const u64 a = b >= o ? s_a[b - o] : 0, u = b >= o ? s_u[b - o] : 0;
        __syncthreads();
        s_a[b] += a;
        s_u[b] += u;
        __syncthreads();
    }
    w.start[b + 1] = s_a[b];
    w.unit_start[b + 1] = s_u[b];
    w.cursor[b] = s_a[b] - c;
    if (b == 0) { w.start[0] = 0; w.unit_start[0] = 0; }
    // the largest number of units in one bucket (the second half walks units bucket-interleaved)
    __syncthreads();
    s_u[b] = (c + unit_points - 1) / unit_points;
    __syncthreads();
    for (int o = kPartBuckets / 2; o > 0; o >>= 1) {
        if (b < o && s_u[b + o] > s_u[b]) s_u[b] = s_u[b + o];
        __syncthreads();
    }
    if (b == 0) w.unit_start[kPartBuckets + 1] = s_u[0];
}

template <class Src, int REC>
__global__ void __launch_bounds__(kPartThreads) part_scatter_kernel(Src src, int64_t n, PartWork w) {
    constexpr int PER = kPartTile / kPartThreads;
    extern __shared__ double s_dyn[];
    double* s_val = s_dyn;                                                  // [REC][kPartTile]
    int* s_cnt = reinterpret_cast<int*>(s_val + REC * kPartTile);           // [NB] counts, then offsets
    int* s_off = s_cnt + kPartBuckets;                                      // [NB] exclusive offsets inside the tile
    long long* s_base = reinterpret_cast<long long*>(s_off + kPartBuckets); // [NB] global position of slot 0 of a bucket
    unsigned short* s_bkt = reinterpret_cast<unsigned short*>(s_base + kPartBuckets);   // [kPartTile] bucket of a sorted slot
    __shared__ int s_wsum[kPartThreads / 32];
    src.init();
    const int64_t n_tiles = (n + kPartTile - 1) / kPartTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * kPartTile;
        for (int b = threadIdx.x; b < kPartBuckets; b += kPartThreads) s_cnt[b] = 0;
        __syncthreads();
        double v[PER][REC];
        int bkt[PER], rank[PER];
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int64_t i = i0 + p * kPartThreads + threadIdx.x;
            if (i < n) src.load(i, v[p]);
        }
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int64_t i = i0 + p * kPartThreads + threadIdx.x;
            bkt[p] = -1;
            if (i < n) {
                bkt[p] = src.bucket_of(v[p]);
                rank[p] = atomicAdd(&s_cnt[bkt[p]], 1);
            }
        }
        __syncthreads();
        // exclusive scan of the tile's bucket counts (kPartBuckets == kPartThreads: one entry per thread)
        {
            const int b = threadIdx.x, lane = b & 31, wrp = b >> 5;
            const int c = s_cnt[b];
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[wrp] = incl;
            __syncthreads();
            int before = 0;
#pragma unroll
            for (int q = 0; q < kPartThreads / 32; ++q) before += (q < wrp) ? s_wsum[q] : 0;
            const int excl = before + incl - c;
            s_off[b] = excl;
            s_base[b] = c ? (long long)atomicAdd(w.cursor + b, (u64)c) - excl : 0;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            if (bkt[p] >= 0) {
                const int slot = s_off[bkt[p]] + rank[p];
#pragma unroll
                for (int c = 0; c < REC; ++c) s_val[c * kPartTile + slot] = v[p][c];
                s_bkt[slot] = (unsigned short)bkt[p];
            }
        }
        __syncthreads();
        const int n_here = (int)((n - i0) < kPartTile ? (n - i0) : kPartTile);
        for (int slot = threadIdx.x; slot < n_here; slot += kPartThreads) {
            const long long pos = s_base[s_bkt[slot]] + slot;
#pragma unroll
            for (int c = 0; c < REC; ++c) w.rec[(int64_t)c * n + pos] = s_val[c * kPartTile + slot];
        }
        __syncthreads();
    }
}

template <int REC>
constexpr size_t part_scatter_smem() {
    return (size_t)REC * kPartTile * 8 + 2 * kPartBuckets * 4 + kPartBuckets * 8 + kPartTile * 2;
}

// histogram -> scan -> scatter on `st`; w.count must have been zeroed on the same stream
template <class Src, int REC>
cudaError_t launch_partition(const Src& src, int64_t n, const PartWork& w, int unit_points, int blocks, cudaStream_t st) {
    const size_t smem = part_scatter_smem<REC>();
    cudaError_t e = cudaFuncSetAttribute(part_scatter_kernel<Src, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // the histogram pass keeps only four 8-byte loads in flight per thread: 8 CTAs per SM to cover the HBM latency
    const int hist_blocks = (int)((n + 4 * kPartThreads - 1) / (4 * kPartThreads) < 4 * (int64_t)blocks
                                      ? (n + 4 * kPartThreads - 1) / (4 * kPartThreads) : 4 * (int64_t)blocks);
    part_hist_kernel<Src><<<hist_blocks, kPartThreads, 0, st>>>(src, n, w.count); ASVGP_LAUNCHED();
    part_scan_kernel<<<1, kPartBuckets, 0, st>>>(w, unit_points); ASVGP_LAUNCHED();
    part_scatter_kernel<Src, REC><<<blocks, kPartThreads, smem, st>>>(src, n, w); ASVGP_LAUNCHED();
    return cudaGetLastError();
}

// The units of the second half, walked BUCKET-INTERLEAVED: slot i is chunk i / NB of bucket i % NB, so that CTAs running
// at the same time work on different buckets and their REDs go to different band entries / cells (with units in storage
// order ~100 CTAs add into the same few hundred addresses at once and the L2 serialises them).
struct UnitTable {
    u64 start[kPartBuckets + 1];
    u64 max_chunks;
    __device__ void stage(const PartWork& w) {
        for (int b = threadIdx.x; b <= kPartBuckets; b += blockDim.x) start[b] = w.start[b];
        if (threadIdx.x == 0) max_chunks = w.unit_start[kPartBuckets + 1];
    }
    __device__ int64_t n_slots() const { return (int64_t)max_chunks * kPartBuckets; }
    // false: the slot is past the end of its bucket
    __device__ bool find(int64_t slot, int unit_points, int& bucket, int64_t& first, int& count) const {
        bucket = (int)(slot % kPartBuckets);
        first = (int64_t)start[bucket] + (slot / kPartBuckets) * unit_points;
        const int64_t left = (int64_t)start[bucket + 1] - first;
        count = (int)(left < unit_points ? left : unit_points);
        return left > 0;
    }
};

}  // namespace asvgp
