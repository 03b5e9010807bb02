// sort_kernels.cu — stable LSD radix sort of the fp64 margins with carried row indices.
//
// Replaces: np.argsort(m) + np.sort(m)  (src/optim/algorithms.py:92-93).  The reference's argsort is
// numpy's unstable introsort; on tie-free keys it equals the stable permutation, which is the only
// well-defined target under ties and what this sort produces (bit-exact vs np.argsort(kind="stable")).
//
// fp64 -> order-preserving u64 keys (-0.0 == +0.0, NaN last, prox_core.h), 8 passes of 8 bits.
// Per pass, three short kernels, none of which spins on another CTA:
//   hist    : per-tile digit histogram (2048 keys per tile; warp-aggregated with match.any so a pass
//             whose keys all share one digit does not serialise on shared-memory atomics)
//   scan    : one CTA per digit: exclusive scan of that digit's counts over the tiles + digit total
//   scatter : digit bases (256-wide scan of the totals, in-CTA) + tile offsets + stable in-tile ranks by
//             warp-level match.any multisplit with per-warp digit counters in shared memory
// The first pass converts doubles on load and synthesises the index payload; the last pass writes the
// sorted doubles and the int32 permutation directly.
#include "common.cuh"

namespace {

constexpr int kSortThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kSortThreads * kItems;  // 2048 keys per tile
constexpr int kWarps = kSortThreads / 32;

__device__ __forceinline__ uint64_t load_key(const void* src, int from_double, int64_t i) {
    if (from_double) return rbl_key_from_bits(reinterpret_cast<const uint64_t*>(src)[i]);
    return reinterpret_cast<const uint64_t*>(src)[i];
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const void* __restrict__ keys, int from_double,
                                                                  int64_t n, int shift, int ntiles,
                                                                  uint32_t* __restrict__ tile_hist) {
    __shared__ uint32_t h[kWarps][257];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    for (int q = tid; q < kWarps * 257; q += kSortThreads) (&h[0][0])[q] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)tile * kTile + (int64_t)warp * (kItems * 32);
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        const uint32_t dg = (i < n) ? (uint32_t)((load_key(keys, from_double, i) >> shift) & 0xff) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, dg);
        if (lane == __ffs(peers) - 1) h[warp][dg] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += h[w][tid];
    tile_hist[(size_t)tid * ntiles + tile] = s;
}

// one CTA per digit: in-place exclusive scan of tile_hist[digit][0..ntiles), total -> digit_tot[digit]
__global__ void __launch_bounds__(kSortThreads) radix_scan_kernel(uint32_t* __restrict__ tile_hist, int ntiles,
                                                                  uint32_t* __restrict__ digit_tot) {
    __shared__ uint32_t wsum[kWarps];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* a = tile_hist + (size_t)blockIdx.x * ntiles;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += kSortThreads) {
        const int i = base + tid;
        const uint32_t v = (i < ntiles) ? a[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t t = wsum[w];
            if (w < warp) woff += t;
            tot += t;
        }
        if (i < ntiles) a[i] = carry + woff + (x - v);
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid == 0) digit_tot[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(
    const void* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int from_double, int64_t n, int shift,
    int ntiles, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ digit_tot,
    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int last, double* __restrict__ sorted_out,
    int32_t* __restrict__ perm_out) {
    __shared__ uint32_t cnt[kWarps][257];
    __shared__ uint32_t wtot[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    for (int q = tid; q < kWarps * 257; q += kSortThreads) (&cnt[0][0])[q] = 0;
    __syncthreads();

    // warp w owns the contiguous span [base + w*256, +256): round j covers 32 consecutive keys, so
    // (warp, round, lane) order == original order inside the tile  => stable
    const int64_t wbase = (int64_t)tile * kTile + (int64_t)warp * (kItems * 32);
    uint64_t key[kItems];
    uint32_t val[kItems];
    uint16_t rank[kItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        const bool ok = i < n;
        key[j] = ok ? load_key(keys_in, from_double, i) : 0ull;
        val[j] = ok ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
        const uint32_t dg = ok ? (uint32_t)((key[j] >> shift) & 0xff) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, dg);
        const int leader = __ffs(peers) - 1;
        uint32_t basec = 0;
        if (lane == leader) {
            basec = cnt[warp][dg];
            cnt[warp][dg] = basec + __popc(peers);
        }
        basec = __shfl_sync(0xffffffffu, basec, leader);
        rank[j] = (uint16_t)(basec + __popc(peers & lt_mask));
        __syncwarp();
    }
    // digit base = exclusive scan of the 256 digit totals (thread t <-> digit t)
    uint32_t dtot = digit_tot[tid];
    uint32_t x = dtot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wtot[warp] = x;
    __syncthreads();  // also orders the per-warp counters
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
        if (w < warp) woff += wtot[w];
    {
        const int dg = tid;
        uint32_t run = woff + (x - dtot) + tile_off[(size_t)dg * ntiles + tile];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = cnt[w][dg];
            cnt[w][dg] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        if (i < n) {
            const uint32_t dg = (uint32_t)((key[j] >> shift) & 0xff);
            const uint32_t pos = cnt[warp][dg] + rank[j];
            if (last) {
                if (sorted_out) reinterpret_cast<uint64_t*>(sorted_out)[pos] = rbl_bits_from_key(key[j]);
                if (perm_out) perm_out[pos] = (int32_t)val[j];
            } else {
                keys_out[pos] = key[j];
                vals_out[pos] = val[j];
            }
        }
    }
}

// ---- the whole 8-pass sort as ONE persistent cooperative kernel ---------------------------------------------
// CTA c owns the contiguous key range [c chunk, (c+1) chunk) in every pass (chunk = ceil(n / grid), one CTA
// per SM).  Per pass: (1) rank the range's keys by digit in shared memory (warp multisplit, as above) and
// publish the CTA's 256 digit counts; grid barrier; (2) every CTA sums the published counts (digit totals and
// the counts of the CTAs before it: G coalesced 1 KB rows), scans the totals, and scatters — keys, payload and
// ranks stay in registers across the barrier when the range is a single 8192-key tile (n <= 1.2M on 148 SMs);
// grid barrier.  16 barriers instead of 24 launches: the sort of 1M (key, index) pairs is latency-bound, not
// bandwidth-bound (12 MB per pass lives in L2).
constexpr int kPT = 1024;
constexpr int kPWarps = kPT / 32;
constexpr int kPTile = kPT * kItems;  // 8192 keys
constexpr size_t kPSortSmem =
    (size_t)(kPWarps * 257 + 256 + 256 + 8 + 2 * 16 * 256 + 3 * 256) * sizeof(uint32_t) + (size_t)kPTile * 12;

struct PSortParams {
    const double* m;
    int64_t n;
    uint64_t *kA, *kB;
    uint32_t *vA, *vB;
    uint32_t* counts;    // [grid][256]
    unsigned int* bar;   // [0] barrier counter, [1] exit ticket; both zero on entry and on exit
    double* sorted_out;
    int32_t* perm_out;
    int64_t chunk;
    unsigned long long* dbg;  // optional [8][6] globaltimer stamps of CTA 0 (dev tool)
};

__device__ __forceinline__ void sort_grid_barrier(unsigned int* ctr, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPT, 1) radix_sort_persistent_kernel(const PSortParams p) {
    extern __shared__ __align__(16) uint32_t psm[];
    uint32_t(*cnt)[257] = reinterpret_cast<uint32_t(*)[257]>(psm);                       // [kPWarps][257]
    uint32_t* base = psm + kPWarps * 257;  // [256]; 32 * 257 is a multiple of 4, so uint4 alignment holds
    uint32_t* tot = base + 256;                                                          // [256]
    uint32_t* wtot = tot + 256;                                                          // [8]
    uint32_t(*part_tot)[256] = reinterpret_cast<uint32_t(*)[256]>(wtot + 8);             // [16][256]
    uint32_t(*part_pre)[256] = reinterpret_cast<uint32_t(*)[256]>(wtot + 8 + 16 * 256);  // [16][256]
    uint32_t* gbase = wtot + 8 + 32 * 256;                                               // [256]
    uint32_t* tcnt = gbase + 256;                                                        // [256]
    uint32_t* lstart = tcnt + 256;                                                       // [256]
    uint64_t* skey = reinterpret_cast<uint64_t*>(lstart + 256);                          // [kPTile] (8-byte aligned)
    uint32_t* sval = reinterpret_cast<uint32_t*>(skey + kPTile);                         // [kPTile]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, me = blockIdx.x;
    const int64_t r0 = (int64_t)me * p.chunk;
    const int64_t r1 = (r0 + p.chunk < p.n) ? r0 + p.chunk : p.n;
    const int ntile = r0 < r1 ? (int)((r1 - r0 + kPTile - 1) / kPTile) : 0;
    const bool single = p.chunk <= kPTile;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned int target = 0;

    uint64_t key[kItems];
    uint32_t val[kItems];
    uint16_t rank[kItems];
    uint32_t qoff = 0;

    for (int pass = 0; pass < 8; ++pass) {
        const int shift = pass * 8;
        const bool from_double = pass == 0, last = pass == 7;
        const void* kin = from_double ? (const void*)p.m : (const void*)((pass & 1) ? p.kA : p.kB);
        const uint32_t* vin = from_double ? nullptr : ((pass & 1) ? p.vA : p.vB);
        uint64_t* kout = (pass & 1) ? p.kB : p.kA;
        uint32_t* vout = (pass & 1) ? p.vB : p.vA;

        auto rank_tile = [&](int k) {
            // all loads of the tile first (independent, L1-bypassing: other CTAs rewrote these buffers earlier
            // in this launch), zero the counters while they are in flight
            const int64_t wbase = r0 + (int64_t)k * kPTile + (int64_t)warp * (kItems * 32);
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int64_t i = wbase + j * 32 + lane;
                const bool ok = i < r1;
                key[j] = ok ? (from_double ? rbl_key_from_bits(reinterpret_cast<const uint64_t*>(kin)[i])
                                           : __ldcg(reinterpret_cast<const unsigned long long*>(kin) + i))
                            : 0ull;
                val[j] = ok ? (vin ? __ldcg(vin + i) : (uint32_t)i) : 0u;
            }
            {
                uint4* c4 = reinterpret_cast<uint4*>(&cnt[0][0]);  // 32 * 257 words = 2056 uint4
                for (int q = tid; q < kPWarps * 257 / 4; q += kPT) c4[q] = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int64_t i = wbase + j * 32 + lane;
                const bool ok = i < r1;
                const uint32_t dg = ok ? (uint32_t)((key[j] >> shift) & 0xff) : 256u;
                // peers = lanes holding the same digit, from 9 ballots (MATCH.ANY serialises over the ~30
                // distinct digits of a warp and was the single largest cost of a pass)
                uint32_t peers = 0xffffffffu;
#pragma unroll
                for (int bit = 0; bit < 9; ++bit) {
                    const bool set = (dg >> bit) & 1u;
                    const uint32_t bal = __ballot_sync(0xffffffffu, set);
                    peers &= set ? bal : ~bal;
                }
                const int leader = __ffs(peers) - 1;
                uint32_t basec = 0;
                if (lane == leader) {
                    basec = cnt[warp][dg];
                    cnt[warp][dg] = basec + __popc(peers);
                }
                basec = __shfl_sync(0xffffffffu, basec, leader);
                rank[j] = (uint16_t)(basec + __popc(peers & lt_mask));
                __syncwarp();
            }
            __syncthreads();
        };

        auto stamp = [&](int slot) {
            if (p.dbg && blockIdx.x == 0 && tid == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                p.dbg[pass * 6 + slot] = t;
            }
        };
        stamp(0);
        // ---- phase 1: digit counts of my range
        if (tid < 256) tot[tid] = 0;
        for (int k = 0; k < ntile; ++k) {
            rank_tile(k);
            {   // 4 threads per digit, 8 warps' counters each (bank-conflict free: 257-word rows)
                const int dg = tid >> 2, q = tid & 3;
                uint32_t t = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) t += cnt[8 * q + w][dg];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                if (q == 0) tot[dg] += t;
            }
            if (!single) __syncthreads();
        }
        __syncthreads();
        if (tid < 256) __stcg(&p.counts[(size_t)me * 256 + tid], tot[tid]);
        stamp(1);
        target += G;
        sort_grid_barrier(p.bar, target);
        stamp(2);

        // ---- phase 2: digit totals, counts of the CTAs before me, digit bases
        {
            // 64 threads x uint4 cover one 256-digit row; 16 row groups; <= 10 independent loads per thread
            const int q = tid >> 6, v4 = tid & 63;
            const uint4* cp = reinterpret_cast<const uint4*>(p.counts);
            uint4 t = make_uint4(0, 0, 0, 0), pre = make_uint4(0, 0, 0, 0);
#pragma unroll 10
            for (int c = q; c < G; c += 16) {
                const uint4 v = __ldcg(&cp[(size_t)c * 64 + v4]);
                t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
                if (c < me) { pre.x += v.x; pre.y += v.y; pre.z += v.z; pre.w += v.w; }
            }
            *reinterpret_cast<uint4*>(&part_tot[q][4 * v4]) = t;
            *reinterpret_cast<uint4*>(&part_pre[q][4 * v4]) = pre;
        }
        __syncthreads();
        uint32_t dtot = 0, dpre = 0, x = 0;
        if (tid < 256) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                dtot += part_tot[q][tid];
                dpre += part_pre[q][tid];
            }
            x = dtot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) wtot[warp] = x;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if (w < warp) woff += wtot[w];
            base[tid] = woff + (x - dtot) + dpre;
        }
        __syncthreads();
        stamp(3);
        for (int k = 0; k < ntile; ++k) {
            if (!single) rank_tile(k);
            // tile-local digit starts (exclusive scan of the tile's digit counts), then per-warp local offsets;
            // gbase[dg] = global start of this tile's run of digit dg minus its local start
            {
                const int dg = tid >> 2, q = tid & 3;
                uint32_t sq = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) sq += cnt[8 * q + w][dg];
                // exclusive prefix over the 4 warp groups of this digit, and the digit's tile count
                uint32_t inc = sq;
                uint32_t y = __shfl_up_sync(0xffffffffu, inc, 1, 4);
                if (q >= 1) inc += y;
                y = __shfl_up_sync(0xffffffffu, inc, 2, 4);
                if (q >= 2) inc += y;
                const uint32_t tcount = __shfl_sync(0xffffffffu, inc, 3, 4);
                qoff = inc - sq;
                if (q == 0) tcnt[dg] = tcount;
            }
            __syncthreads();
            if (warp == 0) {  // 256-wide exclusive scan by one warp: 8 digits per lane
                uint32_t v[8], run = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[e] = tcnt[lane * 8 + e];
                    run += v[e];
                }
                uint32_t x = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                uint32_t ex = x - run;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    lstart[lane * 8 + e] = ex;
                    ex += v[e];
                }
            }
            __syncthreads();
            {
                const int dg = tid >> 2, q = tid & 3;
                const uint32_t ls = lstart[dg];
                uint32_t run = ls + qoff;
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    const uint32_t c = cnt[8 * q + w][dg];
                    cnt[8 * q + w][dg] = run;
                    run += c;
                }
                if (q == 0) {
                    gbase[dg] = base[dg] - ls;
                    base[dg] += tcnt[dg];
                }
            }
            __syncthreads();
            // reorder the tile by digit in shared memory so that the global stores below are coalesced runs
            // (8192 / 256 = 32 keys per digit on average) instead of 8-byte writes to random sectors
            const int64_t tbase = r0 + (int64_t)k * kPTile;
            const int tlen = (int)((r1 - tbase < kPTile) ? (r1 - tbase) : kPTile);
            const int64_t wbase = tbase + (int64_t)warp * (kItems * 32);
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int64_t i = wbase + j * 32 + lane;
                if (i < r1) {
                    const uint32_t dg = (uint32_t)((key[j] >> shift) & 0xff);
                    const uint32_t lpos = cnt[warp][dg] + rank[j];
                    skey[lpos] = key[j];
                    sval[lpos] = val[j];
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int L = tid + j * kPT;
                if (L < tlen) {
                    const uint64_t kk = skey[L];
                    const uint32_t vv = sval[L];
                    const uint32_t dg = (uint32_t)((kk >> shift) & 0xff);
                    const uint32_t pos = gbase[dg] + (uint32_t)L;
                    if (last) {
                        if (p.sorted_out) reinterpret_cast<uint64_t*>(p.sorted_out)[pos] = rbl_bits_from_key(kk);
                        if (p.perm_out) p.perm_out[pos] = (int32_t)vv;
                    } else {
                        kout[pos] = kk;
                        vout[pos] = vv;
                    }
                }
            }
            __syncthreads();
        }
        stamp(4);
        if (!last) {
            target += G;
            sort_grid_barrier(p.bar, target);
        }
        stamp(5);
    }
    // leave the barrier words clean for the next launch: the last CTA out resets them (no memset node needed)
    if (tid == 0) {
        const unsigned int tk = atomicAdd(p.bar + 1, 1u);
        if (tk == gridDim.x - 1) {
            p.bar[0] = 0u;
            p.bar[1] = 0u;
            __threadfence();
        }
    }
}

}  // namespace

// 1 if the persistent sort can run on this device (cooperative launch, one CTA per SM)
int rbl_sort_persistent_ok(rbl_ctx* c) {
    if (c->psort_checked) return c->psort_ok;
    c->psort_checked = 1;
    c->psort_ok = 0;
    int coop = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device) != cudaSuccess || !coop) return 0;
    if (cudaFuncSetAttribute(radix_sort_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kPSortSmem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, radix_sort_persistent_kernel, kPT, kPSortSmem) !=
            cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    c->psort_ok = 1;
    return 1;
}

int rbl_k_sort_persistent(rbl_ctx* c, const double* m, int64_t n, double* sorted_out, int32_t* perm_out,
                          cudaStream_t s) {
    PSortParams p;
    int G = c->num_sms;
    const int64_t min_chunk = 2048;  // do not spread tiny sorts over the whole machine
    if ((int64_t)G * min_chunk > n) G = (int)((n + min_chunk - 1) / min_chunk);
    if (G < 1) G = 1;
    p.m = m;
    p.n = n;
    p.kA = c->keysA;
    p.kB = c->keysB;
    p.vA = c->valsA;
    p.vB = c->valsB;
    p.counts = c->sort_counts;
    p.bar = c->gticket + 16;
    p.sorted_out = sorted_out;
    p.perm_out = perm_out;
    p.chunk = (n + G - 1) / G;
    p.dbg = c->sort_dbg;
    void* args[] = {(void*)&p};
    RBL_CUDA(cudaLaunchCooperativeKernel((const void*)radix_sort_persistent_kernel, dim3(G), dim3(kPT), args,
                                         kPSortSmem, s));
    RBL_LAUNCH_CHECK();
    return RBL_OK;
}

int rbl_sort_tiles(int64_t n) { return (int)((n + kTile - 1) / kTile); }

// sorted_out / perm_out may be null when only one of them is wanted (objective: keys only)
int rbl_k_sort(rbl_ctx* c, const double* m, int64_t n, double* sorted_out, int32_t* perm_out, cudaStream_t s) {
    if (n <= 0) return RBL_OK;
    if (!c->sort_legacy && rbl_sort_persistent_ok(c)) return rbl_k_sort_persistent(c, m, n, sorted_out, perm_out, s);
    const int ntiles = rbl_sort_tiles(n);
    if (ntiles > c->sort_tiles) {
        rbl_set_error("sort of %lld keys exceeds the handle's capacity", (long long)n);
        return RBL_ERR_ARG;
    }
    const void* kin = m;
    const uint32_t* vin = nullptr;
    uint64_t* kout = c->keysA;
    uint32_t* vout = c->valsA;
    uint32_t* digit_tot = c->tile_hist + (size_t)256 * c->sort_tiles;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = pass * 8;
        const int from_double = pass == 0;
        const int last = pass == 7;
        radix_hist_kernel<<<ntiles, kSortThreads, 0, s>>>(kin, from_double, n, shift, ntiles, c->tile_hist);
        RBL_LAUNCH_CHECK();
        radix_scan_kernel<<<256, kSortThreads, 0, s>>>(c->tile_hist, ntiles, digit_tot);
        RBL_LAUNCH_CHECK();
        radix_scatter_kernel<<<ntiles, kSortThreads, 0, s>>>(kin, vin, from_double, n, shift, ntiles, c->tile_hist,
                                                            digit_tot, kout, vout, last, sorted_out, perm_out);
        RBL_LAUNCH_CHECK();
        kin = kout;
        vin = vout;
        if (kout == c->keysA) {
            kout = c->keysB;
            vout = c->valsB;
        } else {
            kout = c->keysA;
            vout = c->valsA;
        }
    }
    return RBL_OK;
}
