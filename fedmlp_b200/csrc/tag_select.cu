// tag_select.cu — K3b: sign split + top-fraction selection, K3c: label / mask fill.
//
// Replaces utils/local_training.py:1061-1112 (np.where sign split, int(frac*len) counts,
// utils/utils.py:24-35 max_m_indices / min_n_indices = Python `sorted(enumerate(list))`) and
// DatasetSplit_pseudo.__getitem__ (:1456-1477).  The reference sorts all N similarities in the
// interpreter twice per class; here one CTA per (segment, class) radix-SELECTS the m-th largest
// clean and k-th smallest noise value (both are "largest |sim|" on their side), so only
// histogram passes over an L2-resident [N] vector are needed.
//
// Ordering key: comp = (bits(|sim|) << 32) | (0xFFFFFFFF - local_row).  Larger comp = picked
// first; equal similarities are therefore taken in increasing row order, which is exactly the
// tie behaviour of Python's stable sort in both max_m_indices (reverse=True keeps the original
// order of equal elements) and min_n_indices.  All comps of an item are distinct, so the
// selection "comp >= T" has exactly m (resp. k) members.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace fmlp {

constexpr int kSelBins = 2048;  // 11-bit digits
// Histogram copies per CTA (indexed by lane & 7).  The leading digit of |sim| (exponent + 3 mantissa bits) takes only
// a few dozen values, so with one copy the 32 lanes of a warp serialised on a handful of shared-memory words in the
// first sweep (r02: ~1 cycle per KEY); the kernel owns its SM anyway (1024 threads), so it spends 128 KB on eight
// copies and adds them up when the bins are scanned.
constexpr int kSelCopies = 8;
constexpr size_t kSelSmemBytes = (size_t)kSelCopies * 2 * kSelBins * sizeof(int);

struct SelArgs {
    const float* sim;
    uint8_t* tag;
    int32_t* counts;     // [S][C][4]
    int32_t* remaining;  // [S][C] candidates (tag == 0) left after this selection, or null
    int32_t* sel;        // [S][C][2][cap]
    unsigned long long* cand;  // ws [S][C][2][cap]
    int64_t ld_sim, ld_tag, cap;
    double clean_frac, noise_frac;
    int C;
    SegTable seg;  // mask_a = missing classes
};

// Block-wide exclusive scan of a packed pair of counts (low / high 32 bits = clean / noise side);
// returns the exclusive prefix and writes the grand total (same value in every thread).
template <int THREADS>
__device__ __forceinline__ unsigned long long block_exclusive_scan2(unsigned long long v, unsigned long long* s_warp,
                                                                    unsigned long long* total) {
    constexpr int NW = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // s_warp reuse
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = (lane < NW) ? s_warp[lane] : 0ull;
        unsigned long long winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < NW) s_warp[lane] = winc - w;  // exclusive warp offsets
        if (lane == NW - 1) s_warp[NW] = winc;
    }
    __syncthreads();
    *total = s_warp[NW];
    return s_warp[warp] + inc - v;
}

__device__ __forceinline__ unsigned long long make_comp(float s, uint32_t local_row) {
    const uint32_t mag = __float_as_uint(fabsf(s));  // -0.0 -> 0, NaN never reaches here
    return ((unsigned long long)mag << 32) | (unsigned long long)(0xFFFFFFFFu - local_row);
}
// side: 0 clean (sim >= 0), 1 noise (sim < 0), -1 neither (NaN) — np.where(sim >= 0) / (sim < 0)
__device__ __forceinline__ int side_of(float s) { return s >= 0.f ? 0 : (s < 0.f ? 1 : -1); }

// ---- thread-block cluster helpers (an item may be split over the CTAs of one cluster) -----------
__device__ __forceinline__ uint32_t sel_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t sel_cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void sel_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared-memory object of this CTA) in the CTA of rank `rank` of the cluster
__device__ __forceinline__ uint32_t sel_dsmem(const void* p, uint32_t rank) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ int sel_dsmem_ld(uint32_t addr) { int v; asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ unsigned long long sel_dsmem_ld64(uint32_t addr) { unsigned long long v; asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void sel_dsmem_st64(uint32_t addr, unsigned long long v) { asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ int sel_dsmem_add(uint32_t addr, int v) { int o; asm volatile("atom.shared::cluster.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); return o; }

// An item (segment, class) is handled by ONE CLUSTER of `cs` CTAs (cs = 1: a plain CTA).  CTA `cr` of the cluster
// owns rows [cr*chunk, (cr+1)*chunk) of the segment; the per-pass histograms are built per CTA and added up by every
// CTA through distributed shared memory, the small-bin gather list, the selection counters and the thresholds
// live in the shared memory of CTA 0.  r02: one CTA per item left 13 CTAs on 148 SMs for the single-client
// 85,000-row workload (84 us, issue-bound on those 13 SMs: ~190 instructions per key over its sweeps).
// IPT > 0: every thread keeps its IPT candidate keys in registers for all passes (CTA chunks of up
// to IPT * THREADS rows); IPT == 0: keys are re-read from global memory in every pass.
// key = comp | side << 63.  Both sides share every histogram pass and every scan.
// COPIES: histogram copies (1 or kSelCopies); RESOLVE: leave the digit loop for the small-bin gather.  Items of up to
// 16,384 rows run as before (one CTA, keys in registers, one 16 KB histogram, six cheap digit sweeps): for them the
// extra shared memory and the gather sweep cost more than they save (r02: 12.4 -> 16.4 us at 8 x 6,875 rows).
template <int IPT, int THREADS, int COPIES, bool RESOLVE>
__global__ void __launch_bounds__(THREADS, 1) tag_select_kernel(const __grid_constant__ SelArgs a) {
    constexpr bool CACHED = IPT > 0;
    constexpr int NK = CACHED ? IPT : 1;
    constexpr int BPT = kSelBins / THREADS;  // bins per thread in the scan (2 or 4)
    extern __shared__ __align__(16) int s_hist_all[];   // [COPIES][2][kSelBins]; copy 0 doubles as the merged
    int (*s_hist)[kSelBins] = reinterpret_cast<int (*)[kSelBins]>(s_hist_all);   // histogram / gather list / rank tile
    __shared__ unsigned long long s_warp[THREADS / 32 + 1];
    __shared__ int s_found[2][3];            // per side: bin, remaining-in-bin, bin count
    __shared__ int s_count[3];               // selected per side, candidates seen (cluster: those of CTA 0 count)
    __shared__ int s_gather[2];              // small-bin gather: keys appended per side (CTA 0's)
    __shared__ unsigned long long s_thresh[2];

    // chained to the similarity kernel (and the fill+loss kernel to this one) by programmatic dependent launches:
    // the launch latency of each hides behind its predecessor; the data dependency is the wait below
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t cs = sel_cluster_size(), cr = sel_cluster_rank();
    const int item = blockIdx.x / cs;
    const int s = item / a.C, c = item - s * a.C;
    int32_t* counts = a.counts + (int64_t)item * 4;
    if (!((a.seg.mask_a[s] >> c) & 1u)) {       // uniform over the cluster
        if (cr == 0) {
            if (threadIdx.x < 4) counts[threadIdx.x] = 0;
            if (threadIdx.x == 0 && a.remaining) a.remaining[item] = 0;
        }
        return;
    }
    const int64_t r0 = a.seg.rows[s];
    const uint32_t n_seg = (uint32_t)(a.seg.rows[s + 1] - r0);
    const uint32_t chunk = (n_seg + cs - 1) / cs;
    const uint32_t lo = min(n_seg, cr * chunk);           // this CTA's rows of the segment: [lo, n)
    const uint32_t n = min(n_seg, lo + chunk);
    const float* sim = a.sim + (int64_t)c * a.ld_sim + r0;
    uint8_t* tag = a.tag + (int64_t)c * a.ld_tag + r0;
    constexpr unsigned long long kSideBit = 1ull << 63;
    if (threadIdx.x < 3) s_count[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_gather[threadIdx.x] = 0;

    // ---- candidate keys ------------------------------------------------------------------
    unsigned long long key[NK];
    uint32_t valid = 0;
    int n_cand = 0;  // tag == 0 rows (including NaN similarities, which are never picked)
    if (CACHED) {
        float v[NK];
        uint8_t tg[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const uint32_t i = lo + (uint32_t)k * THREADS + threadIdx.x;
            const bool in = i < n;
            tg[k] = in ? tag[i] : (uint8_t)1;
            v[k] = in ? sim[i] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const uint32_t i = lo + (uint32_t)k * THREADS + threadIdx.x;
            const int sd = side_of(v[k]);
            key[k] = make_comp(v[k], i) | (sd == 1 ? kSideBit : 0ull);
            if (tg[k] == 0) { ++n_cand; if (sd >= 0) valid |= 1u << k; }
        }
    }
    // visit every candidate key of this thread (count_cand: the uncached path also counts its tag == 0 rows,
    // once, in the first histogram pass — a pass of its own was one more sweep over the segment)
    auto for_each = [&](auto&& fn, bool count_cand = false) {
        // fn(key, ok) is called for every key slot of the thread; ok = the slot holds a candidate
        if (CACHED) {
#pragma unroll
            for (int k = 0; k < NK; ++k) fn(key[k], ((valid >> k) & 1u) != 0u);
        } else {
            // batches of 16 independent (tag, sim) loads per thread: with one dependent pair per iteration a pass
            // over an 85,000-row client was ~80 serial L2 round trips per thread (r02: 136 us for 13 classes;
            // 84 us with batches of 8, eight sweeps per item)
            constexpr int U = 16;
            for (uint32_t base0 = lo; base0 < n; base0 += THREADS * U) {     // same trip count for every thread
                const uint32_t base = base0 + threadIdx.x;
                uint8_t tg[U];
                float v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t i = base + (uint32_t)u * THREADS;
                    const bool in = i < n;
                    tg[u] = in ? tag[i] : (uint8_t)1;
                    v[u] = in ? sim[i] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool cand = tg[u] == 0;
                    if (count_cand && cand) ++n_cand;
                    const int sd = side_of(v[u]);
                    fn(make_comp(v[u], base + (uint32_t)u * THREADS) | (sd == 1 ? kSideBit : 0ull), cand && sd >= 0);
                }
            }
        }
    };

    unsigned long long prefix[2] = {0ull, 0ull};
    unsigned long long thresh[2] = {~0ull, ~0ull};  // comp >= thresh selects; ~0 selects nothing
    int rem[2] = {0, 0};
    bool done[2] = {false, false};
    // A side whose target bin holds at most kSmallBin keys leaves the digit loop: its bin is gathered into shared
    // memory in ONE more sweep and the rem-th largest key found by counting.  All comps are distinct down to the
    // row bits, so the digit loop alone always ran its six sweeps; now it is one or two plus the gather.
    constexpr int kSmallBin = 512;
    bool resolve[2] = {false, false};
    int res_shift[2] = {0, 0};
    int n_side[2] = {0, 0}, want[2] = {0, 0};

    // digits of the 63 significant comp bits, most significant first: 5 x 11 bits + 8 bits
    for (int pass = 0; pass < 6; ++pass) {
        if (done[0] && done[1]) break;
        const int shift = pass < 5 ? 52 - 11 * pass : 0;
        const int width = pass < 5 ? 11 : 8;
        for (int i = threadIdx.x; i < COPIES * 2 * kSelBins / 4; i += THREADS)
            reinterpret_cast<int4*>(s_hist_all)[i] = make_int4(0, 0, 0, 0);
        __syncthreads();
        const unsigned long long dmask = (1ull << width) - 1ull;
        const unsigned long long pre0 = prefix[0], pre1 = prefix[1];
        const bool d0 = done[0], d1 = done[1];
        for_each([&](unsigned long long k, bool ok) {
            const int sd = (int)(k >> 63);
            const unsigned long long comp = k & ~kSideBit;
            const bool hit = ok && !(sd ? d1 : d0) && (pass == 0 || (comp >> (shift + width)) == (sd ? pre1 : pre0));
            if (hit) atomicAdd(&s_hist_all[((threadIdx.x & (COPIES - 1)) * 2 + sd) * kSelBins + (int)((comp >> shift) & dmask)], 1);
        }, !CACHED && pass == 0);
        __syncthreads();
        // bins are visited from the top: thread t owns bins 2047-BPT*t .. 2047-BPT*t-(BPT-1); it adds up the copies
        int h[2][BPT];
        unsigned long long ls = 0ull;
#pragma unroll
        for (int j = 0; j < BPT; ++j) {
            const int b = kSelBins - 1 - (BPT * (int)threadIdx.x + j);
            int v0 = 0, v1 = 0;
#pragma unroll
            for (int cp = 0; cp < COPIES; ++cp) {
                v0 += s_hist_all[(cp * 2 + 0) * kSelBins + b];
                v1 += s_hist_all[(cp * 2 + 1) * kSelBins + b];
            }
            h[0][j] = v0; h[1][j] = v1;
            if (cs > 1) { s_hist[0][b] = v0; s_hist[1][b] = v1; }    // copy 0 := this CTA's histogram (own bins only)
        }
        if (cs > 1) {
            sel_cluster_sync();                                   // every CTA's histogram is complete
            // the item's histogram = sum of the cluster's histograms, read through distributed shared memory
            // (all loads independent; every CTA computes the same sums and therefore takes the same decisions)
            for (uint32_t r = 1; r < cs; ++r) {
                const uint32_t rr = (cr + r) % cs;                // start at the neighbour: spreads the remote reads
#pragma unroll
                for (int j = 0; j < BPT; ++j) {
                    const int b = kSelBins - 1 - (BPT * (int)threadIdx.x + j);
                    h[0][j] += sel_dsmem_ld(sel_dsmem(&s_hist[0][b], rr));
                    h[1][j] += sel_dsmem_ld(sel_dsmem(&s_hist[1][b], rr));
                }
            }
            sel_cluster_sync();                                   // all reads done: the histograms may be reused
        }
#pragma unroll
        for (int j = 0; j < BPT; ++j)
            ls += (unsigned long long)(uint32_t)h[0][j] | ((unsigned long long)(uint32_t)h[1][j] << 32);
        unsigned long long total2;
        const unsigned long long ex2 = block_exclusive_scan2<THREADS>(ls, s_warp, &total2);
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            if (done[sd]) continue;  // uniform across the CTA
            const int total = (int)(uint32_t)(total2 >> (32 * sd));
            const int ex = (int)(uint32_t)(ex2 >> (32 * sd));
            const int lsum = (int)(uint32_t)(ls >> (32 * sd));
            if (pass == 0) {
                n_side[sd] = total;
                const double frac = sd == 0 ? a.clean_frac : a.noise_frac;
                long long w = (long long)__dmul_rn(frac, (double)total);  // int(frac * len), :1069-1070
                if (w < 0) w = 0;
                if (w > total) w = total;
                want[sd] = (int)w;
                rem[sd] = (int)w;
                if (w == 0) { done[sd] = true; continue; }
            }
            if (ex < rem[sd] && rem[sd] <= ex + lsum) {
                int cum = ex;
#pragma unroll
                for (int j = 0; j < BPT; ++j) {
                    if (cum < rem[sd] && rem[sd] <= cum + h[sd][j]) {
                        s_found[sd][0] = kSelBins - 1 - (BPT * (int)threadIdx.x + j);
                        s_found[sd][1] = rem[sd] - cum;
                        s_found[sd][2] = h[sd][j];
                    }
                    cum += h[sd][j];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            if (done[sd]) continue;
            const int bin = s_found[sd][0], rem_in = s_found[sd][1], bin_count = s_found[sd][2];
            prefix[sd] = (prefix[sd] << width) | (unsigned long long)bin;
            if (bin_count == rem_in) {  // whole bin is taken: threshold = smallest comp with this prefix
                thresh[sd] = prefix[sd] << shift;
                done[sd] = true;
            } else {
                rem[sd] = rem_in;
                if (RESOLVE && bin_count <= kSmallBin && shift > 0) { resolve[sd] = true; res_shift[sd] = shift; done[sd] = true; }
            }
        }
        __syncthreads();
    }

    // ---- small target bins: gather the keys that share the prefix, take the rem-th largest ----
    if (RESOLVE && (resolve[0] || resolve[1])) {          // uniform across the cluster
        unsigned long long* list = reinterpret_cast<unsigned long long*>(&s_hist[0][0]);   // [2][kSmallBin] in CTA 0, histogram is dead
        const uint32_t list0 = sel_dsmem(list, 0), gather0 = sel_dsmem(&s_gather[0], 0);
        const unsigned long long pre0 = prefix[0], pre1 = prefix[1];
        const int sh0 = res_shift[0], sh1 = res_shift[1];
        const bool rs0 = resolve[0], rs1 = resolve[1];
        for_each([&](unsigned long long k, bool ok) {
            const int sd = (int)(k >> 63);
            const unsigned long long comp = k & ~kSideBit;
            if (!ok || !(sd ? rs1 : rs0)) return;
            if ((comp >> (sd ? sh1 : sh0)) == (sd ? pre1 : pre0)) {
                if (cs > 1) {
                    const int slot = sel_dsmem_add(gather0 + 4u * sd, 1);
                    if (slot < kSmallBin) sel_dsmem_st64(list0 + 8u * (uint32_t)(sd * kSmallBin + slot), comp);
                } else {
                    const int slot = atomicAdd(&s_gather[sd], 1);
                    if (slot < kSmallBin) list[sd * kSmallBin + slot] = comp;
                }
            }
        });
        if (cs > 1) sel_cluster_sync(); else __syncthreads();
        if (cr == 0) {
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                if (!resolve[sd]) continue;
                const int cnt = min(s_gather[sd], kSmallBin);
                if ((int)threadIdx.x < cnt) {
                    const unsigned long long mine = list[sd * kSmallBin + threadIdx.x];
                    int rank = 0;
                    for (int j = 0; j < cnt; ++j) rank += (list[sd * kSmallBin + j] > mine) ? 1 : 0;
                    if (rank == rem[sd] - 1) s_thresh[sd] = mine;
                }
            }
        }
        if (cs > 1) sel_cluster_sync(); else __syncthreads();
#pragma unroll
        for (int sd = 0; sd < 2; ++sd)
            if (resolve[sd]) thresh[sd] = cs > 1 ? sel_dsmem_ld64(sel_dsmem(&s_thresh[sd], 0)) : s_thresh[sd];
    }

    // ---- compaction: mark the tag state and collect the selected comps -------------------
    const uint32_t count0 = sel_dsmem(&s_count[0], 0);
    n_cand = warp_sum_i(n_cand);
    if ((threadIdx.x & 31) == 0 && n_cand) {
        if (cs > 1) sel_dsmem_add(count0 + 8u, n_cand); else atomicAdd(&s_count[2], n_cand);
    }
    {
        const unsigned long long t0 = thresh[0], t1 = thresh[1];
        for_each([&](unsigned long long k, bool ok) {
            const int sd = (int)(k >> 63);
            const unsigned long long comp = k & ~kSideBit;
            if (ok && comp >= (sd ? t1 : t0)) {
                const int slot = cs > 1 ? sel_dsmem_add(count0 + 4u * sd, 1) : atomicAdd(&s_count[sd], 1);
                if (slot < a.cap) a.cand[((int64_t)item * 2 + sd) * a.cap + slot] = comp;
                tag[0xFFFFFFFFu - (uint32_t)(comp & 0xFFFFFFFFull)] = (uint8_t)(1 + sd);
            }
        });
    }
    if (cs > 1) sel_cluster_sync(); else __syncthreads();     // counters final; the cluster's a.cand writes are visible
    int total_sel[2];
    total_sel[0] = cs > 1 ? sel_dsmem_ld(count0) : s_count[0];
    total_sel[1] = cs > 1 ? sel_dsmem_ld(count0 + 4u) : s_count[1];
    if (threadIdx.x == 0 && cr == 0) {
        counts[0] = n_side[0]; counts[1] = n_side[1]; counts[2] = want[0]; counts[3] = want[1];
        if (a.remaining) a.remaining[item] = s_count[2] - s_count[0] - s_count[1];
    }

    // ---- rank order: position = number of selected comps that are larger (CTA cr ranks every cs-th batch) ----
    unsigned long long* tile = reinterpret_cast<unsigned long long*>(&s_hist[0][0]);  // 2048 entries
    constexpr int kTile = 2048;
    for (int sd = 0; sd < 2; ++sd) {
        const int cnt = (int)min((int64_t)total_sel[sd], a.cap);
        const unsigned long long* cand = a.cand + ((int64_t)item * 2 + sd) * a.cap;
        int32_t* out = a.sel + ((int64_t)item * 2 + sd) * a.cap;
        for (int e0 = (int)cr * THREADS; e0 < cnt; e0 += (int)cs * THREADS) {
            const int e = e0 + threadIdx.x;
            const unsigned long long mine = e < cnt ? cand[e] : 0ull;
            int rank = 0;
            for (int t0 = 0; t0 < cnt; t0 += kTile) {
                __syncthreads();
                for (int j = threadIdx.x; j < kTile && t0 + j < cnt; j += THREADS) tile[j] = cand[t0 + j];
                __syncthreads();
                const int lim = min(kTile, cnt - t0);
                for (int j = 0; j < lim; ++j) rank += (tile[j] > mine) ? 1 : 0;
            }
            if (e < cnt)
                out[rank] = (int32_t)(r0 + (int64_t)(0xFFFFFFFFu - (uint32_t)(mine & 0xFFFFFFFFull)));
        }
        __syncthreads();
    }
    if (cs > 1) sel_cluster_sync();      // nobody leaves while its shared memory may still be read
}

struct FillArgs {
    const float* labels_in;
    const uint8_t* tag;
    float* y;
    float* distill;
    float* sup;
    int64_t ld_tag;
    int C;
    SegTable seg;  // mask_a = active, mask_b = missing
};

__global__ void __launch_bounds__(256) mask_fill_kernel(const __grid_constant__ FillArgs a) {
    const int64_t n_el = a.seg.rows[a.seg.S] * a.C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / a.C;
        const int c = (int)(e - row * a.C);
        const int s = find_segment(a.seg.rows, a.seg.S, row);
        float y = 0.f, dis = 0.f;
        if ((a.seg.mask_a[s] >> c) & 1u) {
            y = a.labels_in[e];                       // annotated class keeps its label (:1458-1460)
        } else if ((a.seg.mask_b[s] >> c) & 1u) {
            const uint8_t t = a.tag[(int64_t)c * a.ld_tag + row];
            y = (t == 2) ? 1.f : 0.f;                 // in the noise list -> pseudo-positive (:1464-1466)
            dis = (t == 0) ? 1.f : 0.f;               // in neither list -> distilled (:1467-1468)
        }
        if (a.y) a.y[e] = y;
        if (a.distill) a.distill[e] = dis;
        if (a.sup) a.sup[e] = 1.f - dis;              // sup_cls = ~distill_cls (:1173)
    }
}

}  // namespace fmlp

using namespace fmlp;

extern "C" size_t fmlp_tag_select_ws_bytes(int S, int C, int64_t cap) {
    if (S < 1 || C < 1 || cap < 0) return 0;
    const size_t n = (size_t)S * C * 2 * (size_t)(cap > 0 ? cap : 1);
    return n * sizeof(unsigned long long);
}

extern "C" int fmlp_tag_select(const float* sim, int64_t ld_sim, uint8_t* tag, int64_t ld_tag, int C,
                               int S, const int64_t* seg_rows, const uint32_t* seg_missing,
                               double clean_frac, double noise_frac, int32_t* counts, int32_t* remaining,
                               int32_t* sel, int64_t cap, void* ws, size_t ws_bytes, fmlp_stream_t stream) {
    if (!sim || !tag || !seg_missing || !counts || !sel || !ws || C < 1 || C > FMLP_MAX_CLASSES || cap < 1)
        return FMLP_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(ws) & 7u) return FMLP_ERR_UNSUPPORTED;
    if (ws_bytes < fmlp_tag_select_ws_bytes(S, C, cap)) return FMLP_ERR_WORKSPACE;
    SelArgs a;
    int rc = fill_seg_table(a.seg, S, seg_rows, seg_missing, nullptr);
    if (rc != FMLP_OK) return rc;
    for (int s = 0; s < S; ++s)
        if (seg_rows[s + 1] - seg_rows[s] > 0x7fffffffLL) return FMLP_ERR_UNSUPPORTED;
    if (ld_sim < seg_rows[S] || ld_tag < seg_rows[S]) return FMLP_ERR_BAD_ARG;
    a.sim = sim; a.tag = tag; a.counts = counts; a.remaining = remaining; a.sel = sel; a.cand = (unsigned long long*)ws;
    a.ld_sim = ld_sim; a.ld_tag = ld_tag; a.cap = cap; a.clean_frac = clean_frac; a.noise_frac = noise_frac;
    a.C = C;
    int64_t max_rows = 0;
    for (int s = 0; s < S; ++s) max_rows = std::max<int64_t>(max_rows, seg_rows[s + 1] - seg_rows[s]);
    // CTAs per item (one cluster): enough for the keys of the largest segment to stay in registers (16 x 1024 per CTA),
    // and, when the launch is small, enough to put the items on more SMs (FMLP_SELECT_CLUSTER overrides: 1, 2, 4, 8)
    int cs = 1;
    while (cs < 8 && (max_rows + cs - 1) / cs > 16 * 1024) cs *= 2;
    {
        const int forced = tuning_value(FMLP_TUNE_SELECT_CLUSTER, "FMLP_SELECT_CLUSTER", 0, 8, 0);
        if (forced == 1 || forced == 2 || forced == 4 || forced == 8) cs = forced;
    }
    const int64_t per_cta = (max_rows + cs - 1) / cs;
    const unsigned grid = (unsigned)(S * C * cs);
    cudaStream_t st = (cudaStream_t)stream;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cs; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cs > 1 ? 2 : 1;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.stream = st;
    const bool big = cs > 1 || per_cta > 16 * 1024;       // split or re-reading items: copies + small-bin gather
    cfg.dynamicSmemBytes = big ? kSelSmemBytes : kSelSmemBytes / kSelCopies;
    {
        static bool configured = false;
        if (!configured) {
            cudaError_t e0 = cudaFuncSetAttribute(tag_select_kernel<8, 1024, kSelCopies, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmemBytes);
            if (e0 == cudaSuccess) e0 = cudaFuncSetAttribute(tag_select_kernel<16, 1024, kSelCopies, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmemBytes);
            if (e0 == cudaSuccess) e0 = cudaFuncSetAttribute(tag_select_kernel<0, 1024, kSelCopies, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmemBytes);
            if (e0 != cudaSuccess) return (int)e0;
            configured = true;
        }
    }
    cudaError_t e;
    if (!big) e = per_cta <= 8 * 1024 ? cudaLaunchKernelEx(&cfg, tag_select_kernel<8, 1024, 1, false>, a)
                                      : cudaLaunchKernelEx(&cfg, tag_select_kernel<16, 1024, 1, false>, a);
    else if (per_cta <= 8 * 1024) e = cudaLaunchKernelEx(&cfg, tag_select_kernel<8, 1024, kSelCopies, true>, a);
    else if (per_cta <= 16 * 1024) e = cudaLaunchKernelEx(&cfg, tag_select_kernel<16, 1024, kSelCopies, true>, a);
    else e = cudaLaunchKernelEx(&cfg, tag_select_kernel<0, 1024, kSelCopies, true>, a);
    return e == cudaSuccess ? launch_status() : (int)e;
}

extern "C" int fmlp_mask_fill(const float* labels_in, const uint8_t* tag, int64_t ld_tag, int C, int S,
                              const int64_t* seg_rows, const uint32_t* seg_active,
                              const uint32_t* seg_missing, float* y, float* distill, float* sup,
                              fmlp_stream_t stream) {
    if (!labels_in || !tag || !seg_active || !seg_missing || C < 1 || C > FMLP_MAX_CLASSES)
        return FMLP_ERR_BAD_ARG;
    FillArgs a;
    int rc = fill_seg_table(a.seg, S, seg_rows, seg_active, seg_missing);
    if (rc != FMLP_OK) return rc;
    if (ld_tag < seg_rows[S]) return FMLP_ERR_BAD_ARG;
    const int64_t n_el = seg_rows[S] * C;
    if (n_el == 0) return FMLP_OK;
    a.labels_in = labels_in; a.tag = tag; a.y = y; a.distill = distill; a.sup = sup; a.ld_tag = ld_tag; a.C = C;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    int64_t blocks = (n_el + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    mask_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}
