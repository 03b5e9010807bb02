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
    const int* gate;          // run only if *gate != 0 (the splitter sort overflowed); null: always run
    int* ss_flag;             // splitter-sort state cleaned up on exit when gated: overflow flag,
    uint32_t* ss_count;       //   bucket counts [ss_nb]
    int ss_nb;
    int* ss_stats;            // [0] route of the last hinted call (1 buckets, 2 LSD fallback), [1] largest bucket
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
    if (p.gate && *p.gate == 0) return;  // the splitter sort handled this call (every CTA sees the same flag)
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
    if (p.gate) {  // fallback run: leave the splitter sort's counters and flag clean for the next call
        if (me == 0 && tid == 0) {
            p.ss_stats[0] = 2;
            p.ss_stats[1] = 0;
        }
        __syncthreads();
        for (int b = me * kPT + tid; b < p.ss_nb; b += G * kPT) {
            atomicMax(&p.ss_stats[1], (int)p.ss_count[b]);
            p.ss_count[b] = 0u;
        }
        __threadfence();
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned int tk = atomicAdd(p.bar + 1, 1u);
        if (tk == gridDim.x - 1) {
            p.bar[0] = 0u;
            p.bar[1] = 0u;
            if (p.gate) {
                // overflow this call: do not try the bucket route for the next few calls (the rank order is still
                // moving fast, typically the first ~10 ADMM iterations); otherwise count the pause down
                int* cool = p.ss_flag + 1;
                *cool = (*cool > 0) ? *cool - 1 : 4;
                *p.ss_flag = 0;
            }
            __threadfence();
        }
    }
}

// ---- splitter sort: partition by the PREVIOUS rank order, then sort every bucket in shared memory -----------
// Between two ADMM iterations the rank order of the margins changes little (even when their values shift or scale
// a lot, as they do while rho is small): the rows that sat at B - 1 evenly spaced ranks in the last z-step are read
// at their NEW values, sorted, and used as splitters — B buckets of ~n/B keys each (B = 512 at n = 1M; measured
// along a solve: largest bucket <= 2.1x the mean from iteration 3 on, <= 1.1x from iteration 10).  One pass
// scatters (key, index) into per-bucket slots (arrival order is irrelevant), then one CTA per bucket sorts its
// <= 4096 keys in shared memory by (key, index) — a total order that equals the stable sort order — and writes
// its slice of the output at the prefix of the bucket counts.  No grid barrier and two launches instead of the
// 16 barriers of the LSD sort.  If any bucket would overflow (first call, margins all equal, a jump in the data)
// the partition raises a device flag: the bucket kernel exits and the gated LSD kernel runs instead.
constexpr int kSSCap = 4096;       // slots per bucket up to 2^20 keys
constexpr int kSSCapBig = 8192;    // ... above (where the alternative, the LSD sort, costs 0.3-0.6 ms): the rank order
                                   // of the AoRR / EHRM problems at 2-4 M keys keeps moving by 6-7 bucket loads
static inline int ss_cap_for(int64_t n) { return n > ((int64_t)1 << 20) ? kSSCapBig : kSSCap; }
constexpr int kSSThreads = 1024;
constexpr int kSSItems = 4;        // keys per thread in the partition kernel
#ifndef RBL_SB_THREADS
#define RBL_SB_THREADS 512
#endif
constexpr int kSBThreads = RBL_SB_THREADS;  // bucket kernel: a 1024-key bucket has 512 pairs per stage

struct SSParams {
    const double* m;
    int64_t n;
    const int32_t* prev_perm;   // permutation of an earlier call (splitters: today's keys of the rows that sat at
                                // evenly spaced ranks then)
    int nb;                     // buckets (power of two, <= 4096)
    uint64_t* spl;              // [nb2] sorted splitters of this call
    int cap;                    // slots per bucket (kSSCap or kSSCapBig)
    uint64_t* bkey;             // [nb][cap]
    uint32_t* bval;
    uint32_t* count;            // [nb], zero on entry
    int* flag;                  // [0] overflow flag, zero on entry; [1] calls left to pause after an overflow
    unsigned int* ticket;       // last-CTA ticket of the bucket kernel, zero on entry
    int* stats;                 // [0] route taken, [1] largest bucket (instrumentation)
    double* sorted_out;
    int32_t* perm_out;
};

__global__ void __launch_bounds__(kSSThreads) ss_partition_kernel(const SSParams p) {
    rbl_pdl_wait();
    extern __shared__ __align__(16) unsigned char ssm[];
    int nb2 = 16;
    while (nb2 < p.nb) nb2 <<= 1;
    uint64_t* spl = reinterpret_cast<uint64_t*>(ssm);          // [nb2] sorted splitters (ss_splitters_kernel)
    uint32_t* hist = reinterpret_cast<uint32_t*>(spl + nb2);   // [nb] tile counts, then tile bases
    const int tid = threadIdx.x, nb = p.nb;
    if (p.flag[1] > 0) {  // pausing after a recent overflow: straight to the LSD kernel
        if (blockIdx.x == 0 && tid == 0) p.flag[0] = 1;
        return;
    }
    for (int j = tid; j < nb2; j += kSSThreads) {
        spl[j] = p.spl[j];
        if (j < nb) hist[j] = 0u;
    }
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (kSSThreads * kSSItems);
    uint64_t key[kSSItems];
    int bkt[kSSItems];
    uint32_t rnk[kSSItems];
#pragma unroll
    for (int q = 0; q < kSSItems; ++q) {
        const int64_t i = base + q * kSSThreads + tid;
        bkt[q] = -1;
        if (i < p.n) {
            key[q] = rbl_key_from_bits(reinterpret_cast<const uint64_t*>(p.m)[i]);
            int lo = 0, hi = nb;  // largest j with spl[j] <= key (spl[0] = 0 always qualifies)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (spl[mid] <= key[q]) lo = mid; else hi = mid;
            }
            bkt[q] = lo;
            rnk[q] = atomicAdd(&hist[lo], 1u);
        }
    }
    __syncthreads();
    for (int j = tid; j < nb; j += kSSThreads) {
        const uint32_t c = hist[j];
        hist[j] = c ? atomicAdd(&p.count[j], c) : 0u;  // reserve this tile's range in bucket j
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kSSItems; ++q) {
        if (bkt[q] >= 0) {
            const uint32_t slot = hist[bkt[q]] + rnk[q];
            if (slot < (uint32_t)p.cap) {
                const size_t at = (size_t)bkt[q] * p.cap + slot;
                p.bkey[at] = key[q];
                p.bval[at] = (uint32_t)(base + q * kSSThreads + tid);
            } else {
                *p.flag = 1;  // overflow: the LSD sort takes this call
            }
        }
    }
}

// The partition when the hint is known to be a permutation (the handle's own last output): walk the keys in the
// PREVIOUS rank order.  The row that sat at rank i then is expected in bucket i * nb / n now — two loads of the
// splitter table verify it (a gallop + bisection from there when the row has moved further), so the 10..12-step
// search through a shared-memory table (4.8 M of its 7.1 M wavefronts were bank conflicts) is gone; a warp's 32
// consecutive ranks fall into one or two buckets, so slots are reserved with one global atomic per (warp, bucket)
// (no per-CTA histogram, no nb atomics per CTA) and the (key, row) stores of a warp are contiguous in the bucket.
__global__ void __launch_bounds__(kSSThreads) ss_partition_ranked_kernel(const SSParams p) {
    rbl_pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, nb = p.nb;
    if (p.flag[1] > 0) {  // pausing after a recent overflow: straight to the LSD kernel
        if (blockIdx.x == 0 && tid == 0) p.flag[0] = 1;
        return;
    }
    const uint64_t* __restrict__ spl = p.spl;
    const int64_t base = (int64_t)blockIdx.x * (kSSThreads * kSSItems);
    uint64_t key[kSSItems];
    int64_t row[kSSItems];
    int bkt[kSSItems];
#pragma unroll
    for (int q = 0; q < kSSItems; ++q) {
        const int64_t i = base + q * kSSThreads + tid;
        row[q] = -1;
        if (i < p.n) {
            int64_t r = p.prev_perm[i];
            row[q] = r < 0 ? 0 : (r >= p.n ? p.n - 1 : r);
        }
    }
#pragma unroll
    for (int q = 0; q < kSSItems; ++q)
        if (row[q] >= 0) key[q] = rbl_key_from_bits(reinterpret_cast<const uint64_t*>(p.m)[row[q]]);
#pragma unroll
    for (int q = 0; q < kSSItems; ++q) {
        bkt[q] = -1;
        if (row[q] < 0) continue;
        const int64_t i = base + q * kSSThreads + tid;
        // rank i lay between the ranks floor(j n / nb) the splitters were taken from: j = the largest with that <= i
        int g = (int)(((i + 1) * nb - 1) / p.n);  // i < 2^22, nb <= 2^12
        g = g < 0 ? 0 : (g >= nb ? nb - 1 : g);
        const uint64_t k = key[q];
        int lo, hi;  // largest j with spl[j] <= k lies in [lo, hi)
        if (spl[g] <= k) {
            lo = g;
            hi = g + 1;
            for (int step = 1; hi < nb && spl[hi] <= k; step <<= 1) {
                lo = hi;
                hi = lo + step > nb ? nb : lo + step;
            }
        } else {  // spl[0] = 0 <= every key
            hi = g;
            lo = g - 1;
            for (int step = 1; lo > 0 && spl[lo] > k; step <<= 1) {
                hi = lo;
                lo = hi - step < 0 ? 0 : hi - step;
            }
        }
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (spl[mid] <= k) lo = mid; else hi = mid;
        }
        bkt[q] = lo;
    }
#pragma unroll
    for (int q = 0; q < kSSItems; ++q) {
        const unsigned live = __ballot_sync(0xffffffffu, bkt[q] >= 0);
        if (bkt[q] < 0) continue;
        const unsigned peers = __match_any_sync(live, bkt[q]);
        const int leader = __ffs(peers) - 1;
        uint32_t slot0 = 0;
        if (lane == leader) slot0 = atomicAdd(&p.count[bkt[q]], (uint32_t)__popc(peers));
        slot0 = __shfl_sync(peers, slot0, leader);
        const uint32_t slot = slot0 + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (slot < (uint32_t)p.cap) {
            const size_t at = (size_t)bkt[q] * p.cap + slot;
            p.bkey[at] = key[q];
            p.bval[at] = (uint32_t)row[q];
        } else {
            *p.flag = 1;  // overflow: the LSD sort takes this call
        }
    }
}

// bitonic network of the bucket kernel over sk / sv [N2 = Na + Nb] in shared memory (see ss_bucket_kernel for the
// A / B split).  kFull: compare (key, index) pairs — a total order, identical to the stable order of the keys;
// otherwise the keys alone (enough when no two keys of the bucket are equal).
template <bool kFull>
__device__ __forceinline__ void ss_bucket_network(uint64_t* __restrict__ sk, uint32_t* __restrict__ sv, const int N2,
                                                  const int Na, const int Nb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // bitonic network on (key, index): a total order, identical to the stable order of the keys.  Stages with
    // distance j > 32 run in shared memory across the block.  The stages with j <= 32 of every phase k stay inside
    // blocks of 64 consecutive elements: a warp takes such a block into registers (2 elements per lane: x and
    // x + 32), runs them with shuffles, and writes the block back once — one shared-memory round trip per phase
    // instead of one per stage (the all-shared-memory version moved 72 x 24 B per key and was bound by it).
    auto cmpx = [&](int t, int j, int k) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
        const int l = i | j;
        const bool up = ((i & k) == 0);
        const uint64_t ka = sk[i], kb = sk[l];
        const uint32_t va = sv[i], vb = sv[l];
        const bool gt = kFull ? ((ka > kb) || (ka == kb && va > vb)) : (ka > kb);
        if (gt == up) {
            sk[i] = kb; sk[l] = ka;
            sv[i] = vb; sv[l] = va;
        }
    };
    // register stages of phase k from distance j0 (<= 32) down to 1 on the 64 keys a warp holds
    auto warp_stages = [&](int g0, int g1, int k, int j0, uint64_t& k0, uint64_t& k1, uint32_t& v0, uint32_t& v1) {
        const bool upa = ((g0 & k) == 0), upb = ((g1 & k) == 0);
        int jj = j0;
        if (jj == 32) {  // partner of x is x + 32: both in this lane (k >= 64: one direction for the pair)
            const bool gt = kFull ? ((k0 > k1) || (k0 == k1 && v0 > v1)) : (k0 > k1);
            if (gt == upa) {
                const uint64_t tk = k0; k0 = k1; k1 = tk;
                const uint32_t tv = v0; v0 = v1; v1 = tv;
            }
            jj = 16;
        }
        for (; jj > 0; jj >>= 1) {
            // element g (g0 or g1): direction from bit k of g, lower of its pair iff bit jj of g is clear
            const uint64_t ok0 = __shfl_xor_sync(0xffffffffu, k0, jj), ok1 = __shfl_xor_sync(0xffffffffu, k1, jj);
            const uint32_t ov0 = __shfl_xor_sync(0xffffffffu, v0, jj), ov1 = __shfl_xor_sync(0xffffffffu, v1, jj);
            const bool lower = (lane & jj) == 0;
            {
                // (key, index) pairs are distinct, so "not greater" is "less".  Keys alone: two EQUAL keys make the
                // upper lane take the lower lane's element while that one keeps it — the row index of the other is
                // lost, the keys are not; the caller finds the equal keys side by side afterwards and starts over
                // from the bucket's slots with kFull (a symmetric keys-only rule costs two more compares per
                // element and stage: 49.6 instead of 38.7 us for the kernel at 1 M keys)
                const bool mine_gt = kFull ? ((k0 > ok0) || (k0 == ok0 && v0 > ov0)) : (k0 > ok0);
                const bool keep_min = (lower == upa);
                if (mine_gt == keep_min) { k0 = ok0; v0 = ov0; }
            }
            {
                const bool mine_gt = kFull ? ((k1 > ok1) || (k1 == ok1 && v1 > ov1)) : (k1 > ok1);
                const bool keep_min = (lower == upb);
                if (mine_gt == keep_min) { k1 = ok1; v1 = ov1; }
            }
        }
    };
    // phases k = 2 .. 64 never leave a block of 64 keys: one load, 21 stages in registers, one store
    for (int blk = warp; blk < (N2 >> 6); blk += kSBThreads / 32) {
        const int g0 = (blk << 6) + lane, g1 = g0 + 32;
        uint64_t k0 = sk[g0], k1 = sk[g1];
        uint32_t v0 = sv[g0], v1 = sv[g1];
#pragma unroll
        for (int k = 2; k <= 64; k <<= 1) warp_stages(g0, g1, k, k >> 1, k0, k1, v0, v1);
        sk[g0] = k0; sk[g1] = k1;
        sv[g0] = v0; sv[g1] = v1;
    }
    __syncthreads();
    for (int k = 128; k <= Na; k <<= 1) {
        // B starts at a multiple of 2 Nb, so (i & k) gives it the same directions as a network of its own
        const int lim = (k <= Nb) ? N2 : Na, npair = lim >> 1;
        for (int j = k >> 1; j > 32; j >>= 1) {  // pairs span warps: block-wide stages
            for (int t = tid; t < npair; t += kSBThreads) cmpx(t, j, k);
            __syncthreads();
        }
        for (int blk = warp; blk < (lim >> 6); blk += kSBThreads / 32) {
            const int g0 = (blk << 6) + lane, g1 = g0 + 32;
            uint64_t k0 = sk[g0], k1 = sk[g1];
            uint32_t v0 = sv[g0], v1 = sv[g1];
            warp_stages(g0, g1, k, 32, k0, k1, v0, v1);
            sk[g0] = k0; sk[g1] = k1;
            sv[g0] = v0; sv[g1] = v1;
        }
        __syncthreads();
    }
}

// the splitters, once per call (ONE CTA): today's keys of the rows that sat at nb - 1 evenly spaced ranks in the
// previous order, put in exact order and written to p.spl for the partition CTAs (each of them used to sort its own
// copy: fine for 1024 splitters, not for the 4096 that 4 M keys need).  Sorted by the bucket kernel's network on the
// keys alone (register stages with shuffles: 12.6 us for 1024 splitters with an all-shared-memory network before).
__global__ void __launch_bounds__(kSBThreads) ss_splitters_kernel(const SSParams p) {
    rbl_pdl_wait();
    extern __shared__ __align__(16) unsigned char ssm[];
    int nb2 = 16;  // power-of-two network: pad with maximal keys
    while (nb2 < p.nb) nb2 <<= 1;
    const int N2 = nb2 < 64 ? 64 : nb2;                        // the network's smallest size
    uint64_t* spl = reinterpret_cast<uint64_t*>(ssm);          // [N2] spl[j] = first key of bucket j (spl[0] = 0)
    uint32_t* tag = reinterpret_cast<uint32_t*>(spl + N2);     // [N2] rides along (the network moves pairs)
    const int tid = threadIdx.x, nb = p.nb;
    if (p.flag[1] > 0) return;  // pausing after a recent overflow (the partition kernel raises the flag)
    for (int j = tid; j < N2; j += kSBThreads) {
        uint64_t k = 0ull;
        if (j >= nb) {
            k = 0xffffffffffffffffull;
        } else if (j > 0) {
            const int64_t pos = (int64_t)(((__int128)j * p.n) / nb);
            int64_t row = p.prev_perm[pos];
            row = row < 0 ? 0 : (row >= p.n ? p.n - 1 : row);
            k = rbl_key_from_bits(reinterpret_cast<const uint64_t*>(p.m)[row]);
        }
        spl[j] = k;
        tag[j] = (uint32_t)j;
    }
    __syncthreads();
    ss_bucket_network<false>(spl, tag, N2, N2, 0);
    for (int j = tid; j < nb2; j += kSBThreads) p.spl[j] = j == 0 ? 0ull : spl[j];  // bucket 0 starts at the smallest key
}

__global__ void __launch_bounds__(kSBThreads) ss_bucket_kernel(const SSParams p) {
    rbl_pdl_wait();
    extern __shared__ __align__(16) unsigned char bsm[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(bsm);         // [cap]
    uint32_t* sv = reinterpret_cast<uint32_t*>(sk + p.cap);   // [cap]
    __shared__ uint32_t s_part[32];
    __shared__ uint32_t s_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x, nb = p.nb;
    if (*p.flag) return;  // overflow: nothing here is valid (the gated LSD kernel cleans up)
    const uint32_t cnt = p.count[b];
    {   // output offset = sum of the counts of the buckets before mine
        uint32_t a = 0;
        for (int j = tid; j < b; j += kSBThreads) a += p.count[j];
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s_part[warp] = a;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kSBThreads / 32; ++w) t += s_part[w];
            s_off = t;
        }
    }
    // The network needs power-of-two lengths, and the loads sit just around one (mean n / nb ~ 1000): rounding
    // 1030 keys up to 2048 would cost 2.4x.  The keys are split into a head A of Na keys (the largest power of two
    // <= cnt) and a tail B padded to Nb < Na slots; both are sorted by the same network (B only joins the phases
    // k <= Nb) and every key then finds its output slot by one binary search in the other part.
    int Na = 64;  // at least one full warp of pairs
    while (Na * 2 <= (int)cnt) Na <<= 1;
    int Nb = 0;
    if ((int)cnt > Na) {
        Nb = 64;
        while (Nb < (int)cnt - Na) Nb <<= 1;
        if (Nb >= Na) {  // no saving: one network over 2 Na
            Na <<= 1;
            Nb = 0;
        }
    }
    const int N2 = Na + Nb;
    auto load_bucket = [&]() {
        for (int i = tid; i < N2; i += kSBThreads) {
            const bool ok = i < (int)cnt;
            sk[i] = ok ? p.bkey[(size_t)b * p.cap + i] : 0xffffffffffffffffull;  // sentinels sort last
            sv[i] = ok ? p.bval[(size_t)b * p.cap + i] : 0xffffffffu;
        }
        __syncthreads();
    };
    load_bucket();
    // The network on the keys alone first: exact ties between margins are rare, and a comparator that does not look
    // at the row index is 2 instead of 7 instructions in an instruction-bound kernel.  The keys come out sorted either
    // way (the key multiset survives, see ss_bucket_network); if any two of them turn out equal (or a padding entry
    // sits among the real ones) the row indices cannot be trusted: the bucket is loaded again and sorted with the
    // full (key, index) comparator — the stable order either way.
    ss_bucket_network<false>(sk, sv, N2, Na, Nb);
    {
        int tie = 0;
        // (equal keys ACROSS the two parts are put in order by the (key, index) search below)
        for (int i = tid; i < (int)cnt; i += kSBThreads)
            tie |= (i > 0 && i != Na && sk[i] == sk[i - 1]) || sv[i] == 0xffffffffu;
        if (__syncthreads_or(tie)) {
            load_bucket();
            ss_bucket_network<true>(sk, sv, N2, Na, Nb);
        }
    }
    const uint32_t off = s_off;
    for (int i = tid; i < (int)cnt; i += kSBThreads) {
        const uint64_t ki = sk[i];
        const uint32_t vi = sv[i];
        int pos = i;
        if (Nb) {  // + the keys of the other part that sort before mine ((key, index) pairs are all distinct)
            const bool inA = i < Na;
            int lo = inA ? Na : 0, hi = inA ? (int)cnt : Na;
            const int first = lo;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const uint64_t km = sk[mid];
                if (km < ki || (km == ki && sv[mid] < vi)) lo = mid + 1; else hi = mid;
            }
            pos = (inA ? i : i - Na) + (lo - first);
        }
        if (p.sorted_out) reinterpret_cast<uint64_t*>(p.sorted_out)[off + pos] = rbl_bits_from_key(ki);
        if (p.perm_out) p.perm_out[off + pos] = (int32_t)vi;
    }
    // last CTA out clears the counters for the next call (every CTA has read all the counts it needs by now)
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int tk = atomicAdd(p.ticket, 1u);
        s_off = (tk == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_off) {
        uint32_t mx = 0;
        for (int j = tid; j < nb; j += kSBThreads) {
            mx = max(mx, p.count[j]);
            p.count[j] = 0u;
        }
        for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_part[warp] = mx;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kSBThreads / 32; ++w) mx = max(mx, s_part[w]);
            p.stats[0] = 1;
            p.stats[1] = (int)mx;
            *p.ticket = 0u;
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
                          cudaStream_t s, int gated) {
    PSortParams p;
    p.gate = gated ? c->ss_flag : nullptr;
    p.ss_flag = c->ss_flag;
    p.ss_count = c->ss_count;
    p.ss_nb = c->ss_nb;
    p.ss_stats = c->ss_flag + 4;
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

// buckets of the splitter sort for n keys (0: n too small to bother)
int rbl_ss_buckets(int64_t n) {
    // (a one-CTA shared-memory network over all n <= 8192 keys was tried for the reference's own 6000-row problem:
    // 114 us, shared-memory-throughput bound on a single SM — no better than the LSD kernel; 16 buckets are)
    if (n < 2048) return 0;
    // mean load n / nb in (512, 1024]: >= 4x headroom to the 4096-slot cap for buckets that grow because the rank
    // order moved.  Up to 4096 buckets (n <= 2^22): the splitters are sorted once per call by ss_splitters_kernel.
    // (With at most 1024 buckets, each partition CTA sorting its own splitter copy, 2 M / 4 M keys meant buckets of
    // 2 k / 4 k keys and the route lost to the LSD sort: 340 / 683 us vs 322 / 603 us.)
    int nb = 16;
    while ((int64_t)nb * 1024 < n && nb < 4096) nb <<= 1;
    if ((int64_t)nb * 1024 < n) return 0;
    return nb;
}

size_t rbl_ss_slots(int64_t n) { return (size_t)rbl_ss_buckets(n) * ss_cap_for(n); }

// sorted_out / perm_out may be null when only one of them is wanted (objective: keys only).
// prev_perm (may be null): permutation of an earlier, similar call — enables the splitter sort with the LSD sort
// as the device-gated fallback.
int rbl_k_sort(rbl_ctx* c, const double* m, int64_t n, double* sorted_out, int32_t* perm_out, cudaStream_t s,
               const int32_t* prev_perm) {
    if (n <= 0) return RBL_OK;
    const bool persistent = !c->sort_legacy && rbl_sort_persistent_ok(c);
    if (persistent && prev_perm && c->ss_nb > 0 && n == c->n_global && !c->ss_off) {
        SSParams q;
        q.m = m;
        q.n = n;
        q.prev_perm = prev_perm;
        q.nb = c->ss_nb;
        q.cap = ss_cap_for(n);
        q.spl = c->ss_spl;
        q.bkey = c->ss_bkey;
        q.bval = c->ss_bval;
        q.count = c->ss_count;
        q.flag = c->ss_flag;
        q.ticket = c->gticket + 24;
        q.stats = c->ss_flag + 4;
        q.sorted_out = sorted_out;
        q.perm_out = perm_out;
        const int tiles = (int)((n + kSSThreads * kSSItems - 1) / (kSSThreads * kSSItems));
        int nb2 = 16;
        while (nb2 < q.nb) nb2 <<= 1;
        const size_t smem = (size_t)nb2 * sizeof(uint64_t) + (size_t)q.nb * sizeof(uint32_t);
        RBL_PER_DEVICE(size_t, attr, c);
        if (smem > 48 * 1024 && smem > attr) {
            RBL_CUDA(cudaFuncSetAttribute(ss_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = smem;
        }
        RBL_PER_DEVICE(bool, sattr, c);
        if (!sattr) {  // 4096 splitters: 48 KB of dynamic shared memory next to the kernel's static words
            RBL_CUDA(cudaFuncSetAttribute(ss_splitters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12));
            sattr = true;
        }
        RBL_CUDA(rbl_launch_pdl(ss_splitters_kernel, dim3(1), dim3(kSBThreads), (size_t)(nb2 < 64 ? 64 : nb2) * 12, s, q));
        RBL_LAUNCH_CHECK();
        // a hint that is the buffer this handle's last sort wrote is a permutation: partition in previous-rank order
        const bool ranked = !c->ss_row_order && prev_perm == c->last_perm && n == c->last_perm_n;
        if (ranked)
            RBL_CUDA(rbl_launch_pdl(ss_partition_ranked_kernel, dim3(tiles), dim3(kSSThreads), 0, s, q));
        else
            RBL_CUDA(rbl_launch_pdl(ss_partition_kernel, dim3(tiles), dim3(kSSThreads), smem, s, q));
        RBL_LAUNCH_CHECK();
        RBL_PER_DEVICE(bool, battr, c);
        if (!battr) {
            RBL_CUDA(cudaFuncSetAttribute(ss_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kSSCapBig * 12));
            battr = true;
        }
        RBL_CUDA(rbl_launch_pdl(ss_bucket_kernel, dim3(q.nb), dim3(kSBThreads), (size_t)q.cap * 12, s, q));
        RBL_LAUNCH_CHECK();
        if (perm_out) c->last_perm = perm_out, c->last_perm_n = n;
        return rbl_k_sort_persistent(c, m, n, sorted_out, perm_out, s, 1);  // runs only if the flag was raised
    }
    if (perm_out) c->last_perm = perm_out, c->last_perm_n = n;
    if (persistent) return rbl_k_sort_persistent(c, m, n, sorted_out, perm_out, s, 0);
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
