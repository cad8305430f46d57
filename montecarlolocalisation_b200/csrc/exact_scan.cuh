// exact_scan.cuh — CUDA kernels for the parallel, bit-exact sequential f64 accumulation (see exact_scan_core.cuh).
//
// Tiles of 2048 fp32 weights (256 threads x 8 consecutive items). Per accumulation:
//   k_xs_tilesum   tile sums of the weights (optionally normalising them first: w <- (float)((double)w / total)); the last
//                  block to finish scans them with one warp                         -> P~ at every tile edge
//   k_xs_scan<0>   per tile: P~_i, binade prediction, parity-monoid segmented scan  -> tile composites + SEQ entries
//   k_xs_chain     one thread: carries across tiles, the SEQ elements with the hardware adder, the grand total
//   k_xs_scan<1>   per tile again: same scan, now applying run-start values         -> exact s_i for every i (the CDF)
// HBM traffic: 4 B/particle per pass (dense weights) + 8 B/particle for the CDF write.
// If a prediction cannot be trusted (flag != 0) the single-thread sequential kernels redo the job, so correctness never
// rests on the margin analysis.
#pragma once
#include "exact_scan_core.cuh"
#include "mcl_device.cuh"

namespace mcl {
namespace xs {

constexpr int XS_THREADS = 256;
constexpr int XS_ITEMS = 8;
constexpr int XS_TILE = XS_THREADS * XS_ITEMS;
constexpr int XS_SEQ_CAP = 16;

struct SeqEntry {
    uint32_t idx;
    float w;
    Par pre;              // composite of the PAR elements between the previous SEQ element of this tile (or tile start) and idx
    int first_in_tile;
    int E_prev;           // predicted binade of s_{idx-1}
};
struct TileSummary {
    Par vlast;            // composite after the last SEQ element of the tile (whole tile if none)
    int seq_count;
    int pad;
};
struct Workspace {
    double* tsum;         // [nt]
    double* toff;         // [nt+1]
    TileSummary* tiles;   // [nt]
    SeqEntry* entries;    // [nt * XS_SEQ_CAP]
    Par* carry;           // [nt+1]
    int* seq_base;        // [nt+1]
    double* seq_s;        // [nt * XS_SEQ_CAP]
    int* flag;            // != 0: fall back to the sequential kernel; flag[1] = ticket of k_xs_tilesum (resets itself)
};

__device__ __forceinline__ void load_items(const float* __restrict__ w, int64_t base, int64_t n, float (&x)[XS_ITEMS]) {
    if (base + XS_ITEMS <= n) {
        const float4* p = reinterpret_cast<const float4*>(w + base);
        float4 a = __ldg(p), b = __ldg(p + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < XS_ITEMS; j++) x[j] = (base + j < n) ? w[base + j] : 0.f;
    }
}

// Deterministic block-wide inclusive scan of one double per thread (fixed association).
__device__ __forceinline__ double block_scan_incl(double v, double* smem8) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = dadd(t, v);
    }
    if (lane == 31) smem8[warp] = v;
    __syncthreads();
    double pre = 0.0;
    for (int k = 0; k < warp; k++) pre = dadd(pre, smem8[k]);
    __syncthreads();
    return dadd(pre, v);
}

// Exclusive scan of the tile sums by one warp, contiguous chunks (fixed association) -> P~ at every tile edge.
__device__ __forceinline__ void tile_offsets_warp(const double* __restrict__ tsum, int nt, double* __restrict__ toff, int lane) {
    const int chunk = (nt + 31) / 32;
    const int a = min(nt, lane * chunk), b = min(nt, a + chunk);
    double s = 0.0;
    for (int t = a; t < b; t++) s = dadd(s, __ldcg(tsum + t));
    double incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = dadd(up, incl);
    }
    double run = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) run = 0.0;
    for (int t = a; t < b; t++) { toff[t] = run; run = dadd(run, __ldcg(tsum + t)); }
    if (lane == 31) toff[nt] = run;          // lane 31's chunk always ends at nt
}

// ---- pass 1: tile sums (and normalisation); the last block to finish also scans them (pass 2) ---------------------------
template <bool NORMALISE>
__global__ void __launch_bounds__(XS_THREADS) k_xs_tilesum(const float* __restrict__ w_in, float* __restrict__ w_out,
                                                           float4* __restrict__ part, int64_t n, const double* __restrict__ total,
                                                           double* __restrict__ tsum, double* __restrict__ toff, int* __restrict__ flag) {
    pdl_enter();
    __shared__ double sm[8];
    __shared__ bool last;
    const int64_t base = (int64_t)blockIdx.x * XS_TILE + (int64_t)threadIdx.x * XS_ITEMS;
    float x[XS_ITEMS];
    load_items(w_in, base, n, x);
    if (NORMALISE) {
        const double tot = *total;
#pragma unroll
        for (int j = 0; j < XS_ITEMS; j++)
            if (base + j < n) {
                x[j] = __double2float_rn(ddiv((double)x[j], tot));         // MC:497,503
                w_out[base + j] = x[j];
                // (the reference also stores it into particles(3,i); that buffer is consumed by the resampling that
                // follows in the same call and never visible again, so the 4-byte scatter into 16-byte records is skipped)
            }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < XS_ITEMS; j++) s = dadd(s, (double)x[j]);
    double incl = block_scan_incl(s, sm);
    if (threadIdx.x == XS_THREADS - 1) {
        tsum[blockIdx.x] = incl;
        __threadfence();
        last = atomicAdd(flag + 1, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        if (threadIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
        tile_offsets_warp(tsum, (int)gridDim.x, toff, threadIdx.x);
    }
}

// ---- passes 3 and 5 ---------------------------------------------------------------------------------------------------
struct ScanState {        // segmented parity-monoid scan state of a span of elements
    Par v;                // composite since the last SEQ element in the span (whole span if none)
    int reset;            // span contains a SEQ element
    int cnt;              // number of SEQ elements in the span
};
__device__ __forceinline__ ScanState st_combine(const ScanState& a, const ScanState& b) {
    ScanState r;
    r.v = b.reset ? b.v : par_compose(a.v, b.v);
    r.reset = a.reset | b.reset;
    r.cnt = a.cnt + b.cnt;
    return r;
}
__device__ __forceinline__ ScanState st_shfl_up(const ScanState& s, int o) {
    ScanState r;
    r.v.e = __shfl_up_sync(0xffffffffu, s.v.e, o);
    r.v.o = __shfl_up_sync(0xffffffffu, s.v.o, o);
    r.reset = __shfl_up_sync(0xffffffffu, s.reset, o);
    r.cnt = __shfl_up_sync(0xffffffffu, s.cnt, o);
    return r;
}

template <bool APPLY>
__global__ void __launch_bounds__(XS_THREADS) k_xs_scan(const float* __restrict__ w, int64_t n, int nt, Workspace ws,
                                                        double* __restrict__ out) {
    pdl_enter();
    __shared__ double sm_d[8];
    __shared__ double sm_last[8];
    __shared__ ScanState sm_st[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;
    if (APPLY && *ws.flag) return;            // uniform: the sequential kernel redoes the whole job
    const int64_t base = (int64_t)t * XS_TILE + (int64_t)tid * XS_ITEMS;
    float x[XS_ITEMS];
    load_items(w, base, n, x);
    // P~: thread-serial partial sums, block scan of thread totals, plus the tile offset
    double loc[XS_ITEMS];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < XS_ITEMS; j++) { s = dadd(s, (double)x[j]); loc[j] = s; }
    const double incl = block_scan_incl(s, sm_d);
    double excl_thr = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 31) sm_last[warp] = incl;
    __syncthreads();
    if (lane == 0) excl_thr = warp ? sm_last[warp - 1] : 0.0;
    const double toff = ws.toff[t];
    const int64_t tile_end = min(n, (int64_t)(t + 1) * XS_TILE);       // one past the last valid element of this tile
    double pt[XS_ITEMS];
#pragma unroll
    for (int j = 0; j < XS_ITEMS; j++) {
        pt[j] = dadd(toff, dadd(excl_thr, loc[j]));
        if (base + j == tile_end - 1) pt[j] = ws.toff[t + 1];           // tile edges are shared values
    }
    // P~ of the element before this thread's first one
    double prev_last = __shfl_up_sync(0xffffffffu, pt[XS_ITEMS - 1], 1);
    __syncthreads();
    if (lane == 31) sm_last[warp] = pt[XS_ITEMS - 1];
    __syncthreads();
    if (lane == 0) prev_last = warp ? sm_last[warp - 1] : toff;
    const uint64_t depth = 2ull * (uint64_t)((nt + 31) / 32) + 64;
    // classify + local segmented scan
    Par f[XS_ITEMS];
    bool is_seq[XS_ITEMS];
    int Eof[XS_ITEMS], Eprev[XS_ITEMS];
    ScanState agg;
    agg.v = par_identity(); agg.reset = 0; agg.cnt = 0;
    bool bad = false;
    {
        double pprev = prev_last;
#pragma unroll
        for (int j = 0; j < XS_ITEMS; j++) {
            const int64_t i = base + j;
            f[j] = par_identity();
            is_seq[j] = false;
            Eof[j] = 0; Eprev[j] = 0;
            if (i < n) {
                Pred cur = predict(pt[j], margin_for((uint64_t)i, depth));
                Pred prv = predict(pprev, margin_for(i ? (uint64_t)(i - 1) : 0, depth));
                if (i == 0) { prv.ok = false; prv.zero = true; }
                Eof[j] = cur.E; Eprev[j] = prv.E;
                if (!cur.zero) {
                    bool par = cur.ok && prv.ok && cur.E == prv.E;
                    if (par) { if (!par_of_weight(x[j], cur.E, f[j])) bad = true; }
                    else is_seq[j] = true;
                }
                if (is_seq[j]) { agg.v = par_identity(); agg.reset = 1; agg.cnt++; }
                else agg.v = par_compose(agg.v, f[j]);
            }
            pprev = pt[j];
        }
    }
    // block-wide exclusive scan of the thread aggregates
    ScanState inc = agg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        ScanState up = st_shfl_up(inc, o);
        if (lane >= o) inc = st_combine(up, inc);
    }
    if (lane == 31) sm_st[warp] = inc;
    __syncthreads();
    ScanState pre;
    pre.v = par_identity(); pre.reset = 0; pre.cnt = 0;
    for (int k = 0; k < warp; k++) pre = st_combine(pre, sm_st[k]);
    ScanState lane_excl = st_shfl_up(inc, 1);
    if (lane > 0) pre = st_combine(pre, lane_excl);
    // replay this thread's items from its exclusive prefix
    ScanState run = pre;
    Par carry_in = par_identity();
    int seq_base = 0;
    if (APPLY) { carry_in = ws.carry[t]; seq_base = ws.seq_base[t]; }
    bool okflag = true;
#pragma unroll
    for (int j = 0; j < XS_ITEMS; j++) {
        const int64_t i = base + j;
        if (i >= n) break;
        if (is_seq[j]) {
            if (!APPLY) {
                if (run.cnt < XS_SEQ_CAP) {
                    SeqEntry e;
                    e.idx = (uint32_t)i; e.w = x[j]; e.pre = run.v; e.first_in_tile = run.reset ? 0 : 1; e.E_prev = Eprev[j];
                    ws.entries[(size_t)t * XS_SEQ_CAP + run.cnt] = e;
                } else bad = true;
            } else {
                out[i] = ws.seq_s[seq_base + run.cnt];
            }
            run.v = par_identity(); run.reset = 1; run.cnt++;
        } else {
            run.v = par_compose(run.v, f[j]);
            if (APPLY) {
                const int rank = seq_base + run.cnt;
                const double start = rank ? ws.seq_s[rank - 1] : 0.0;
                const Par comp = run.reset ? run.v : par_compose(carry_in, run.v);
                out[i] = par_apply(start, comp, Eof[j], okflag);
            }
        }
    }
    if (!APPLY && tid == XS_THREADS - 1) {
        TileSummary ts;
        ts.vlast = run.v; ts.seq_count = run.cnt; ts.pad = 0;
        ws.tiles[t] = ts;
    }
    if (bad || !okflag) atomicOr(ws.flag, 1);
}

// ---- pass 4: carries across tiles (parallel), then the SEQ elements one by one (sequential) ---------------------------------
// The composite carried into tile t is a segmented scan of the tile summaries under the parity monoid (a tile with SEQ
// elements restarts the composite), done by one block with warp shuffles. Only the few SEQ elements and the binade runs
// between them need the hardware adder in order; one thread walks those.
constexpr int XS_CHAIN_THREADS = 512;
constexpr int XS_CHAIN_LIST = 2048;
constexpr int XS_CHAIN_PRE = 32;              // listed tiles whose SEQ entries are prefetched into shared memory
constexpr int XS_SEQ_TILE = 1024;

// The always-correct fallback for the total: one thread adds left to right (MC:675) while the rest of the block stages
// the next 1024 weights in shared memory.
__device__ __forceinline__ double seq_total_block(const float* __restrict__ w, int64_t n, float (*tile)[XS_SEQ_TILE]) {
    double acc = 0.0;
    const int64_t n_tiles = (n + XS_SEQ_TILE - 1) / XS_SEQ_TILE;
    for (int i = threadIdx.x; i < XS_SEQ_TILE; i += blockDim.x) { int64_t g = i; tile[0][i] = g < n ? w[g] : 0.f; }
    __syncthreads();
    for (int64_t t = 0; t < n_tiles; t++) {
        const int cur = t & 1;
        if (threadIdx.x == 0) {
            const int64_t cnt = min((int64_t)XS_SEQ_TILE, n - t * XS_SEQ_TILE);
            for (int i = 0; i < cnt; i++) acc = dadd(acc, (double)tile[cur][i]);
        } else if (t + 1 < n_tiles) {
            const int64_t base = (t + 1) * XS_SEQ_TILE;
            for (int i = threadIdx.x - 1; i < XS_SEQ_TILE; i += blockDim.x - 1) { int64_t g = base + i; tile[cur ^ 1][i] = g < n ? w[g] : 0.f; }
        }
        __syncthreads();
    }
    return acc;      // valid in thread 0
}

// w / n: the accumulated terms, for the in-kernel sequential fallback of the total (the CDF has its own fallback kernel
// after the apply pass).
__global__ void __launch_bounds__(XS_CHAIN_THREADS) k_xs_chain(int nt, Workspace ws, double* __restrict__ total_out, const float* __restrict__ w,
                                                               int64_t n) {
    pdl_enter();
    __shared__ ScanState sm_warp[XS_CHAIN_THREADS / 32];
    __shared__ ScanState sm_carry_in;          // running state entering the current chunk of tiles
    __shared__ int sm_fail;
    __shared__ int sm_list[XS_CHAIN_LIST];     // tiles that contain SEQ elements, in order
    __shared__ int sm_list_n;
    __shared__ int sm_wcount[XS_CHAIN_THREADS / 32];
    __shared__ SeqEntry sm_entries[XS_CHAIN_PRE * XS_SEQ_CAP];
    __shared__ Par sm_pcarry[XS_CHAIN_PRE];
    __shared__ int sm_pcnt[XS_CHAIN_PRE], sm_pbase[XS_CHAIN_PRE];
    __shared__ float sm_seq[2][XS_SEQ_TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { sm_carry_in.v = par_identity(); sm_carry_in.reset = 0; sm_carry_in.cnt = 0; sm_fail = *ws.flag ? 1 : 0; sm_list_n = 0; }
    __syncthreads();
    const bool skip = sm_fail != 0;            // uniform: an earlier pass already gave up
    for (int c0 = 0; c0 < nt && !skip; c0 += XS_CHAIN_THREADS) {
        const int t = c0 + tid;
        ScanState mine;
        mine.v = par_identity(); mine.reset = 0; mine.cnt = 0;
        if (t < nt) {
            const TileSummary ts = ws.tiles[t];
            mine.v = ts.vlast; mine.reset = ts.seq_count ? 1 : 0; mine.cnt = ts.seq_count;
            if (ts.seq_count > XS_SEQ_CAP) sm_fail = 1;
        }
        ScanState inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            ScanState up = st_shfl_up(inc, o);
            if (lane >= o) inc = st_combine(up, inc);
        }
        // ordered compaction of the tiles that hold SEQ elements
        const unsigned has = __ballot_sync(0xffffffffu, mine.cnt > 0);
        if (lane == 31) sm_warp[warp] = inc;
        if (lane == 0) sm_wcount[warp] = __popc(has);
        __syncthreads();
        ScanState pre = sm_carry_in;
        for (int k = 0; k < warp; k++) pre = st_combine(pre, sm_warp[k]);
        ScanState lane_excl = st_shfl_up(inc, 1);
        if (lane > 0) pre = st_combine(pre, lane_excl);          // state entering tile t
        {
            int pos = sm_list_n;
            for (int k = 0; k < warp; k++) pos += sm_wcount[k];
            pos += __popc(has & ((1u << lane) - 1u));
            if (mine.cnt > 0) {
                if (pos < XS_CHAIN_LIST) sm_list[pos] = t; else sm_fail = 1;
                if (pos < XS_CHAIN_PRE) {                        // keep what the sequential walk will need on chip
                    sm_pcarry[pos] = pre.v; sm_pcnt[pos] = mine.cnt; sm_pbase[pos] = pre.cnt;
                    for (int k = 0; k < mine.cnt && k < XS_SEQ_CAP; k++) sm_entries[pos * XS_SEQ_CAP + k] = ws.entries[(size_t)t * XS_SEQ_CAP + k];
                }
            }
        }
        if (t < nt) { ws.carry[t] = pre.v; ws.seq_base[t] = pre.cnt; }
        __syncthreads();
        if (tid == XS_CHAIN_THREADS - 1) {
            sm_carry_in = st_combine(pre, mine);      // state leaving the chunk
            int add = 0;
            for (int k = 0; k < XS_CHAIN_THREADS / 32; k++) add += sm_wcount[k];
            sm_list_n += add;
        }
        __syncthreads();
    }
    if (tid == 0 && !skip) {
        const ScanState fin = sm_carry_in;
        ws.carry[nt] = fin.v;
        ws.seq_base[nt] = fin.cnt;
        bool ok = sm_fail == 0;
        double s = 0.0;
        if (ok) {
            const int n_list = min(sm_list_n, XS_CHAIN_LIST);
            for (int li = 0; li < n_list; li++) {
                const bool pre = li < XS_CHAIN_PRE;
                const int t = sm_list[li];
                const int cnt = pre ? sm_pcnt[li] : ws.tiles[t].seq_count;
                const Par carry = pre ? sm_pcarry[li] : ws.carry[t];
                const int base = pre ? sm_pbase[li] : ws.seq_base[t];
                for (int k = 0; k < cnt; k++) {
                    const SeqEntry e = pre ? sm_entries[li * XS_SEQ_CAP + k] : ws.entries[(size_t)t * XS_SEQ_CAP + k];
                    const Par comp = e.first_in_tile ? par_compose(carry, e.pre) : e.pre;
                    s = par_apply(s, comp, e.E_prev, ok);
                    s = dadd(s, (double)e.w);            // the hardware adder: exactly the reference's rounding
                    ws.seq_s[base + k] = s;
                }
            }
            // the run after the last SEQ element: its binade is that of P~_{n-1} = toff[nt]
            if (ok) s = par_apply(s, fin.v, f64_exponent(ws.toff[nt]), ok);
        }
        if (!ok) { atomicOr(ws.flag, 1); sm_fail = 1; }
        else if (total_out) *total_out = s;
    }
    __syncthreads();
    if (sm_fail && total_out) {                  // uniform: redo the total with the single chain
        const double s = seq_total_block(w, n, sm_seq);
        if (tid == 0) *total_out = s;
    }
}

}  // namespace xs
}  // namespace mcl
