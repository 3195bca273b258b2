// Alignment ingestion on the device: unique site patterns in sorted order with their counts, sm_100a.
//
// Replaces the host-side compression of phylo.compress_alignment (dodonaphy/phylo.py:16-58):
//     counts = Counter(zip(*sequences)); patterns = sorted(counts)
// i.e. the columns of the S x n character matrix, de-duplicated, in lexicographic order of their raw bytes
// (taxon 0 most significant), each with its multiplicity.  For the long-alignment configuration (500 taxa x 1M
// sites) that is a sort of a million 500-byte keys.
//
//   1. presence histogram of the 256 byte values -> order-preserving dense codes (c distinct bytes, b = ceil(log2 c)
//      bits each: DNA alignments have <= ~20 distinct characters);
//   2. LSD radix sort of the column permutation, floor(10 / b) taxa per pass (<= 1024 buckets), from the last group of
//      taxa to the first.  A pass is a STABLE counting sort: every warp owns a contiguous segment of 512 columns of
//      the current order (2k warps at 1M columns; all 16 keys of a lane are gathered before the first is used: a pass
//      is bound by the latency of those byte gathers); (a) per-segment bucket histogram, (b) exclusive scan of
//      hist[bucket][segment] (bucket-major, so equal keys keep their order across segments; tiled over many CTAs),
//      (c) scatter -- the warp walks its segment 32 elements at a time, __match_any_sync groups equal keys, the rank
//      inside the group keeps the order inside the warp;
//   3. boundary flags (column differs from its predecessor in sorted order, compared taxon by taxon with early exit)
//      and their exclusive scan = index of every column's unique pattern; the number of patterns goes to the host
//      (which has to size the outputs: one synchronisation, as with any unique());
//   4. gather: tip masks uint8 [S][U] through the IUPAC table and weights int64 [U].
#include <stdlib.h>

#include "common.cuh"

namespace dodo {

constexpr int AL_SEG = 512;       // elements per warp segment (2k warps in flight at 1M columns: a pass is gather-latency bound)
constexpr int AL_KPL = AL_SEG / 32;  // keys per lane: all fetched before the first is used
constexpr int AL_WARPS = 8;       // warps per CTA in the sort passes
constexpr int AL_MAX_BUCKETS = 1024;
constexpr int AL_SCAN_TILE = 4096;  // elements per CTA in the tiled scan
constexpr unsigned AL_NOKEY = 0xffffffffu;

__global__ void al_presence_kernel(const uint8_t *__restrict__ raw, size_t total, unsigned *__restrict__ present) {
    __shared__ unsigned loc[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) loc[i] = 0u;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) loc[raw[i]] = 1u;
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (loc[i]) present[i] = 1u;
}

// code[v] = number of present byte values below v (order-preserving); info[0] = number of distinct values
__global__ void al_codes_kernel(const unsigned *__restrict__ present, uint8_t *__restrict__ code, int *__restrict__ info) {
    if (threadIdx.x == 0) {
        int c = 0;
        for (int v = 0; v < 256; ++v) {
            code[v] = (uint8_t)c;
            if (present[v]) ++c;
        }
        info[0] = c;
    }
}

__device__ __forceinline__ unsigned al_key(const uint8_t *__restrict__ raw, const uint8_t *__restrict__ code, size_t n,
                                           int t0, int tpp, int S, int bits, unsigned col) {
    unsigned k = 0u;
    for (int j = 0; j < tpp; ++j) {
        const int t = t0 + j;
        k = (k << bits) | (t < S ? (unsigned)code[raw[(size_t)t * n + col]] : 0u);
    }
    return k;
}

// (a) hist[bucket][segment].  The eight segments of a CTA are neighbours in a bucket's row, so the counters leave
// (and, in the scatter, the offsets enter) shared memory as whole 32-byte sectors.
__global__ void __launch_bounds__(32 * AL_WARPS) al_hist_kernel(const uint8_t *__restrict__ raw, const uint8_t *__restrict__ code,
                                                              const unsigned *__restrict__ perm, size_t n, int S, int t0,
                                                              int tpp, int bits, int nb, int nseg, unsigned *__restrict__ hist) {
    extern __shared__ unsigned al_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.x * AL_WARPS + warp;
    const int ld = nb + 1;  // (odd stride: the eight counters of a bucket sit in different banks)
    unsigned *cnt = al_sm + (size_t)warp * ld;
    for (int i = lane; i < nb; i += 32) cnt[i] = 0u;
    __syncwarp();
    if (seg < nseg) {
        const size_t lo = (size_t)seg * AL_SEG, hi = lo + AL_SEG < n ? lo + AL_SEG : n;
        unsigned keys[AL_KPL];
#pragma unroll
        for (int j = 0; j < AL_KPL; ++j) {
            const size_t i = lo + (size_t)j * 32 + lane;
            keys[j] = i < hi ? al_key(raw, code, n, t0, tpp, S, bits, perm[i]) : AL_NOKEY;
        }
#pragma unroll
        for (int j = 0; j < AL_KPL; ++j)
            if (keys[j] != AL_NOKEY) atomicAdd(&cnt[keys[j]], 1u);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nb * AL_WARPS; idx += 32 * AL_WARPS) {
        const int b = idx / AL_WARPS, w = idx % AL_WARPS, sg = blockIdx.x * AL_WARPS + w;
        if (sg < nseg) hist[(size_t)b * nseg + sg] = al_sm[(size_t)w * ld + b];
    }
}

// (b) exclusive scan of m unsigned values in place, one CTA (the tile sums of the tiled scan below); total -> *total_out
__global__ void __launch_bounds__(1024) al_scan_kernel(unsigned *__restrict__ v, size_t m, unsigned *__restrict__ total_out) {
    __shared__ unsigned part[1024];
    const int tid = threadIdx.x;
    const size_t chunk = (m + 1023) / 1024;
    const size_t lo = (size_t)tid * chunk < m ? (size_t)tid * chunk : m, hi = lo + chunk < m ? lo + chunk : m;
    unsigned s = 0u;
    for (size_t i = lo; i < hi; ++i) s += v[i];
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        const unsigned add = tid >= d ? part[tid - d] : 0u;
        __syncthreads();
        part[tid] += add;
        __syncthreads();
    }
    unsigned run = tid ? part[tid - 1] : 0u;
    for (size_t i = lo; i < hi; ++i) {
        const unsigned x = v[i];
        v[i] = run;
        run += x;
    }
    if (tid == 1023 && total_out) *total_out = part[1023];
}

// Tiled scan of long arrays (the bucket x segment histogram: 2M values per pass at 1M columns; the boundary flags):
// every CTA scans its own tile of 4096 values in place (four per thread, 16-byte accesses) and leaves the tile's sum;
// al_scan_kernel scans the tile sums; al_scan_add_kernel adds them back.
__global__ void __launch_bounds__(1024) al_scan_tiles_kernel(unsigned *__restrict__ v, size_t m, unsigned *__restrict__ tile_sums) {
    __shared__ unsigned wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t base = (size_t)blockIdx.x * AL_SCAN_TILE + (size_t)tid * 4;
    unsigned x[4] = {0u, 0u, 0u, 0u};
    const bool vec = (reinterpret_cast<uintptr_t>(v) & 15u) == 0u && base + 3 < m;  // (the flag array starts at 8 n bytes)
    if (vec) {
        const uint4 q = *reinterpret_cast<const uint4 *>(v + base);
        x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else {
        for (int k = 0; k < 4; ++k)
            if (base + k < m) x[k] = v[base + k];
    }
    const unsigned s = x[0] + x[1] + x[2] + x[3];
    unsigned inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned w = wsum[lane];
        unsigned winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        wsum[lane] = winc - w;
        if (lane == 31) tile_sums[blockIdx.x] = winc;
    }
    __syncthreads();
    unsigned run = wsum[warp] + inc - s;
    if (vec) {
        uint4 q;
        q.x = run; run += x[0];
        q.y = run; run += x[1];
        q.z = run; run += x[2];
        q.w = run;
        *reinterpret_cast<uint4 *>(v + base) = q;
    } else {
        for (int k = 0; k < 4; ++k)
            if (base + k < m) { v[base + k] = run; run += x[k]; }
    }
}

__global__ void __launch_bounds__(1024) al_scan_add_kernel(unsigned *__restrict__ v, size_t m, const unsigned *__restrict__ tile_offs) {
    const unsigned off = tile_offs[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * AL_SCAN_TILE + (size_t)threadIdx.x * 4;
    if ((reinterpret_cast<uintptr_t>(v) & 15u) == 0u && base + 3 < m) {
        uint4 q = *reinterpret_cast<uint4 *>(v + base);
        q.x += off; q.y += off; q.z += off; q.w += off;
        *reinterpret_cast<uint4 *>(v + base) = q;
    } else {
        for (int k = 0; k < 4; ++k)
            if (base + k < m) v[base + k] += off;
    }
}

// (c) stable scatter
__global__ void __launch_bounds__(32 * AL_WARPS) al_scatter_kernel(const uint8_t *__restrict__ raw, const uint8_t *__restrict__ code,
                                                                 const unsigned *__restrict__ perm, size_t n, int S, int t0,
                                                                 int tpp, int bits, int nb, int nseg,
                                                                 const unsigned *__restrict__ hist, unsigned *__restrict__ out) {
    extern __shared__ unsigned al_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.x * AL_WARPS + warp;
    const int ld = nb + 1;
    unsigned *off = al_sm + (size_t)warp * ld;
    // the keys of the whole segment are requested before the offsets are staged: one gather latency per pass
    unsigned cols[AL_KPL], keys[AL_KPL];
    const size_t lo = (size_t)seg * AL_SEG, hi = lo + AL_SEG < n ? lo + AL_SEG : n;
#pragma unroll
    for (int j = 0; j < AL_KPL; ++j) {
        const size_t i = lo + (size_t)j * 32 + lane;
        const bool valid = seg < nseg && i < hi;
        cols[j] = valid ? perm[i] : 0u;
        keys[j] = valid ? al_key(raw, code, n, t0, tpp, S, bits, cols[j]) : AL_NOKEY;
    }
    for (int idx = threadIdx.x; idx < nb * AL_WARPS; idx += 32 * AL_WARPS) {
        const int b = idx / AL_WARPS, w = idx % AL_WARPS, sg = blockIdx.x * AL_WARPS + w;
        if (sg < nseg) al_sm[(size_t)w * ld + b] = hist[(size_t)b * nseg + sg];
    }
    __syncthreads();
    if (seg >= nseg) return;
#pragma unroll
    for (int j = 0; j < AL_KPL; ++j) {
        const unsigned key = keys[j];
        const bool valid = key != AL_NOKEY;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        unsigned start = 0u;
        if (valid && lane == leader) {
            start = off[key];
            off[key] = start + __popc(peers);
        }
        start = __shfl_sync(0xffffffffu, start, leader);
        if (valid) out[start + rank] = cols[j];
        __syncwarp();
    }
}

__global__ void al_iota_kernel(unsigned *__restrict__ perm, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (unsigned)i;
}

// flags[i] = 1 when the i-th column of the sorted order starts a new pattern
__global__ void al_flags_kernel(const uint8_t *__restrict__ raw, const unsigned *__restrict__ perm, size_t n, int S,
                                unsigned *__restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned f = 1u;
    if (i > 0) {
        const unsigned a = perm[i], b = perm[i - 1];
        f = 0u;
        for (int t = 0; t < S; ++t)
            if (raw[(size_t)t * n + a] != raw[(size_t)t * n + b]) { f = 1u; break; }
    }
    flags[i] = f;
}

// uid[i] = (inclusive scan of flags) - 1 = exclusive scan + flag - 1; we hold the exclusive scan `ex` and the raw flag is
// recomputed from ex[i+1] != ex[i].  first[u] = sorted position where pattern u starts.
__global__ void al_first_kernel(const unsigned *__restrict__ ex, size_t n, unsigned n_unique, unsigned *__restrict__ first) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned nxt = i + 1 < n ? ex[i + 1] : n_unique;
    if (nxt != ex[i]) first[ex[i]] = (unsigned)i;  // position i carries a flag: it is the first column of pattern ex[i]
}

__global__ void al_gather_kernel(const uint8_t *__restrict__ raw, const unsigned *__restrict__ perm, const unsigned *__restrict__ first,
                                 size_t n, int S, unsigned n_unique, const uint8_t *__restrict__ lut, uint8_t *__restrict__ masks,
                                 int64_t *__restrict__ weights) {
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    const unsigned p0 = first[u], p1 = u + 1 < n_unique ? first[u + 1] : (unsigned)n;
    const unsigned col = perm[p0];
    if (blockIdx.y == 0) weights[u] = (int64_t)(p1 - p0);
    for (int t = blockIdx.y; t < S; t += gridDim.y) masks[(size_t)t * n_unique + u] = lut[raw[(size_t)t * n + col]];
}

}  // namespace dodo

using namespace dodo;

// workspace: perm A, perm B, flags/scan [n] (unsigned each), hist [AL_MAX_BUCKETS][nseg], present[256], code[256], info[4],
// tile sums of the tiled scan
static size_t al_nseg(size_t n) { return (n + AL_SEG - 1) / AL_SEG; }
static size_t al_ntiles(size_t m) { return (m + AL_SCAN_TILE - 1) / AL_SCAN_TILE; }
static size_t al_max_tiles(size_t n) { return al_ntiles((size_t)AL_MAX_BUCKETS * al_nseg(n) > n ? (size_t)AL_MAX_BUCKETS * al_nseg(n) : n); }

// exclusive scan of v[0..m) in place (three launches; one for short arrays)
static int al_scan(unsigned *v, size_t m, unsigned *tile_sums, unsigned *total_out, cudaStream_t stream) {
    const size_t nt = al_ntiles(m);
    if (m <= 65536) {  // short arrays: one CTA, one launch (small alignments are launch-bound)
        al_scan_kernel<<<1, 1024, 0, stream>>>(v, m, total_out);
        DODO_LAUNCH_CHECK();
        return DODO_OK;
    }
    al_scan_tiles_kernel<<<(unsigned)nt, 1024, 0, stream>>>(v, m, tile_sums);
    DODO_LAUNCH_CHECK();
    al_scan_kernel<<<1, 1024, 0, stream>>>(tile_sums, nt, total_out);
    DODO_LAUNCH_CHECK();
    al_scan_add_kernel<<<(unsigned)nt, 1024, 0, stream>>>(v, m, tile_sums);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" uint64_t dodo_alignment_workspace_bytes(int S, int64_t n) {
    (void)S;
    const uint64_t nn = (uint64_t)n;
    return align_up(3 * nn * sizeof(unsigned), 256) + align_up((uint64_t)AL_MAX_BUCKETS * al_nseg(nn) * sizeof(unsigned), 256) +
           align_up(256 * sizeof(unsigned) + 256 + 16 * sizeof(int), 256) + align_up(al_max_tiles(nn) * sizeof(unsigned), 256) + 256;
}

// Sorts the columns.  On return (after the stream has been synchronised by the caller) h_n_unique[0] holds the number
// of distinct columns; d_ws holds the sorted permutation and the pattern starts for dodo_alignment_gather.
extern "C" int dodo_alignment_sort(const uint8_t *d_raw, int S, int64_t n64, void *d_ws, int32_t *d_n_unique, void *stream_) {
    DODO_CHECK_ARG(d_raw && d_ws && d_n_unique, "dodo_alignment_sort: NULL argument");
    DODO_CHECK_ARG(S >= 1 && n64 >= 1 && n64 < (1ll << 31), "dodo_alignment_sort: need S >= 1 and 1 <= n < 2^31");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = (size_t)n64;
    const size_t nseg = al_nseg(n);
    char *ws = (char *)d_ws;
    unsigned *permA = (unsigned *)ws;
    unsigned *permB = permA + n;
    unsigned *flags = permB + n;
    unsigned *hist = (unsigned *)(ws + align_up(3 * (uint64_t)n * sizeof(unsigned), 256));
    unsigned *present = (unsigned *)((char *)hist + align_up((uint64_t)AL_MAX_BUCKETS * nseg * sizeof(unsigned), 256));
    uint8_t *code = (uint8_t *)(present + 256);
    int *info = (int *)(code + 256);
    unsigned *tile_sums = (unsigned *)((char *)present + align_up(256 * sizeof(unsigned) + 256 + 16 * sizeof(int), 256));
    DODO_CUDA(cudaMemsetAsync(present, 0, 256 * sizeof(unsigned), stream));
    al_presence_kernel<<<sm_count() * 4, 256, 0, stream>>>(d_raw, (size_t)S * n, present);
    DODO_LAUNCH_CHECK();
    al_codes_kernel<<<1, 32, 0, stream>>>(present, code, info);
    DODO_LAUNCH_CHECK();
    int distinct = 0;  // a handful of bytes to the host: the pass structure depends on the alphabet size
    DODO_CUDA(cudaMemcpyAsync(&distinct, info, sizeof(int), cudaMemcpyDeviceToHost, stream));
    DODO_CUDA(cudaStreamSynchronize(stream));
    int bits = 1;
    while ((1 << bits) < distinct) ++bits;
    const int tpp = bits >= 10 ? 1 : 10 / bits;  // taxa per pass
    const int nb = 1 << (bits * tpp);            // <= 1024 (256 when an alignment uses > 32 distinct bytes)
    al_iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(permA, n);
    DODO_LAUNCH_CHECK();
    const unsigned grid = (unsigned)((nseg + AL_WARPS - 1) / AL_WARPS);
    const size_t smem = (size_t)AL_WARPS * (nb + 1) * sizeof(unsigned);
    unsigned *cur = permA, *nxt = permB;
    const int ngroups = (S + tpp - 1) / tpp;
    for (int g = ngroups - 1; g >= 0; --g) {
        const int t0 = g * tpp;
        al_hist_kernel<<<grid, 32 * AL_WARPS, smem, stream>>>(d_raw, code, cur, n, S, t0, tpp, bits, nb, (int)nseg, hist);
        DODO_LAUNCH_CHECK();
        if (int rc = al_scan(hist, (size_t)nb * nseg, tile_sums, nullptr, stream)) return rc;
        al_scatter_kernel<<<grid, 32 * AL_WARPS, smem, stream>>>(d_raw, code, cur, n, S, t0, tpp, bits, nb, (int)nseg, hist, nxt);
        DODO_LAUNCH_CHECK();
        unsigned *t = cur; cur = nxt; nxt = t;
    }
    if (cur != permA) DODO_CUDA(cudaMemcpyAsync(permA, cur, n * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
    al_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_raw, permA, n, S, flags);
    DODO_LAUNCH_CHECK();
    return al_scan(flags, n, tile_sums, (unsigned *)d_n_unique, stream);
}

extern "C" int dodo_alignment_gather(const uint8_t *d_raw, int S, int64_t n64, void *d_ws, int32_t n_unique, const uint8_t *d_lut,
                                     uint8_t *d_masks, int64_t *d_weights, void *stream_) {
    DODO_CHECK_ARG(d_raw && d_ws && d_lut && d_masks && d_weights, "dodo_alignment_gather: NULL argument");
    DODO_CHECK_ARG(S >= 1 && n64 >= 1 && n_unique >= 1 && n_unique <= n64, "dodo_alignment_gather: bad sizes");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = (size_t)n64;
    unsigned *permA = (unsigned *)d_ws;
    unsigned *first = permA + n;       // perm B is free now
    unsigned *ex = permA + 2 * n;      // exclusive scan of the boundary flags
    al_first_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ex, n, (unsigned)n_unique, first);
    DODO_LAUNCH_CHECK();
    dim3 grid((unsigned)((n_unique + 255) / 256), (unsigned)(S < 64 ? S : 64));
    al_gather_kernel<<<grid, 256, 0, stream>>>(d_raw, permA, first, n, S, (unsigned)n_unique, d_lut, d_masks, d_weights);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}
