// gsc_seed.cuh -- K3: yakmo's k-means++ seeding (init(), as called at enc:824-828), exact and INCREMENTAL.
//
// What the reference does per seed i: draw u (xor128), pick c = lower_bound(r, u * obj) over the
// running FLOAT prefix sum r[j] = r[j-1] + up[j] of the points' current squared distances (i = 0: c = floor(u*N)),
// probe linearly past points already chosen, then lower every point's distance against the new seed and rebuild r.
// 4096 dependent steps; a step that touches all N points three times (distances, prefix, store) is HBM-bound with
// hundreds of frames in flight.  This kernel touches only what a step can change:
//
//  * Distance pass over a NORM-BUCKETED copy of the points (k_seed_prep: counting sort on 12 bits of |p|^2, rows,
//    norms and current distances stored in that order).  d(p, c) >= (|p| - |c|)^2, so a 256-point block whose norm
//    range is farther from |c| than the square root of the block's largest current distance cannot change and is
//    skipped by one test; inside a surviving block the same test per point decides whether the row is read at all.
//    Changed distances are written through to the original-order array up[] (the prefix sum's order) and mark their
//    256-element WINDOW dirty.
//  * Prefix sum without storing r[]: float addition is not associative, but while the running sum stays inside one
//    binade every partial sum is an integer multiple S of the ulp and an element acts on S as a function
//    parity -> increment (round-half-even depends only on parity); these functions compose associatively.  Each
//    window keeps a SUMMARY under the exponent its entry sum had at the last step: (f0, f1) = increment for entry
//    parity 0 / 1, and bounds (neg, pos) on how far partial sums can fall below / rise above the entry.  Only dirty
//    windows are re-summarised (one warp each).  One warp then CHAINS the windows in order: a summary is used iff
//    the actual entry exponent equals the predicted one and S - neg >= 2^23, S + pos < 2^24 (every partial sum
//    stays in the binade, so the composition IS the sequential result); any other window (a binade crossing inside
//    it, ~10 per step) is added up element by element (gsc_warp_chain).
//    The chain yields the exact entry sum of every window and the exact total obj.
//  * lower_bound follows std::lower_bound's probes exactly: a probe in a window without negative elements is
//    decided from the window's entry / exit sums when the target lies outside them (the sums are monotone there);
//    otherwise that window's prefix values are evaluated exactly (gsc_warp_chain) and the probe reads them.
//
// Results are those of the sequential reference bit for bit: seeds, seed cells (sid) and distances (tests:
// test_yakmo_seeding*, golden fixtures, cross-check against the one-warp serial evaluation in k_seed).
#pragma once
#include "gsc_device.cuh"

#define GSC_SW 256            // elements per prefix-sum window = points per norm block
#define GSC_SF_BAD 1          // window flags: an element that no binade summary can hold (non-finite)
#define GSC_SF_NEG 2          //               a negative element (prefix sums not monotone inside the window)

// window / block arrays of a frame start at this offset (non-overlapping for consecutive frames, SURVEY 8a N = cc*C)
__device__ __forceinline__ long long gsc_win_off(const GscFrame &f) { return f.chunk_off / GSC_SW + f.slot; }

struct GscXor128b { unsigned long long x, y, z, w; };
__device__ __forceinline__ float gsc_xor128b(GscXor128b &g) {   // init() RVA 0x18c4-0x18f3
    unsigned long long t = g.x ^ (g.x << 11);
    g.x = g.y; g.y = g.z; g.z = g.w;
    g.w = (g.w ^ (g.w >> 19)) ^ (t ^ (t >> 8));
    return (float)((double)g.w * 5.42101086242752217e-20);  // 2^-64
}

// (f then g)(p) = f(p) + g((p + f(p)) & 1), functions given as (value at parity 0, value at parity 1)
__device__ __forceinline__ void gsc_pf_compose(int &f0, int &f1, int g0, int g1) {
    const int h0 = f0 + ((f0 & 1) ? g1 : g0);
    const int h1 = f1 + (((1 + f1) & 1) ? g1 : g0);
    f0 = h0; f1 = h1;
}

// Element classes for a running sum in the binade with exponent field E0 (24..250): a / ulp = I + fraction;
// cls bit0 = round up whatever the parity, bit1 = exact tie (round to even: depends on the parity of S), 4 = bad.
__device__ __forceinline__ void gsc_classify(float v, float sc, float huge, int &I, int &cls) {
    const float av = fabsf(v);
    if (!(av < huge)) { I = 1 << 26; cls = 4; return; }   // non-finite / two binades up: never inside the binade
    const float aq = av * sc;                              // exact (power-of-two scaling), < 2^25
    const float fl = floorf(aq);
    const float g = aq - fl;                               // exact fraction
    const int ni = (int)fl;
    const int tie = (g == 0.5f) ? 2 : 0;
    if (v >= 0.0f) { I = ni; cls = tie | ((g > 0.5f) ? 1 : 0); }
    else if (g == 0.0f) { I = -ni; cls = 0; }
    else { I = -(ni + 1); cls = tie | ((g < 0.5f) ? 1 : 0); }   // floor(-x) = -(n+1), fraction 1 - g
}
__device__ __forceinline__ int gsc_pf_step(int S, int I, int cls) {
    const int t = S + I;
    return t + ((cls & 1) | ((cls >> 1) & t & 1));
}

// 8 consecutive floats a[j0 .. j0+8), zero beyond n.  (The distance arrays are written by the same kernel: the
// helpers take plain pointers, no __restrict__, so that the loads are ordinary coherent ld.global.)
__device__ __forceinline__ void gsc_load8(const float *a, int j0, int n, float (&v)[8]) {
    if (j0 + 8 <= n && ((reinterpret_cast<unsigned long long>(a + j0) & 15ull) == 0ull)) {
        const float4 t0 = *reinterpret_cast<const float4 *>(a + j0), t1 = *reinterpret_cast<const float4 *>(a + j0 + 4);
        v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (j0 + e < n) ? a[j0 + e] : 0.0f;
    }
}

// ---------------------------------------------------------------------------
// Exact sequential float sum  run := fl(run + a[j]),  j = 0..n-1  (n <= 256) of ONE window, called by a whole warp
// with the same arguments: the warp stages the window in shared memory with one coalesced load, lane 0 adds the
// elements one after the other (a 256-long FADD chain costs about what two certified scan rounds would, with
// none of their special cases) and the sum is handed to every lane.  prefix: buf[j] := the sum after element j.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float gsc_warp_chain(const float *a, int n, float run, float *buf, bool prefix) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float v[8];
    gsc_load8(a, lane * 8, n, v);
    __syncwarp();
    *reinterpret_cast<float4 *>(buf + lane * 8) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4 *>(buf + lane * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    __syncwarp();
    float r = run;
    if (lane == 0) {
        const int n8 = n & ~7;
        for (int j = 0; j < n8; j += 8) {
            const float4 x0 = *reinterpret_cast<const float4 *>(buf + j), x1 = *reinterpret_cast<const float4 *>(buf + j + 4);
            float4 y0, y1;
            r = r + x0.x; y0.x = r; r = r + x0.y; y0.y = r; r = r + x0.z; y0.z = r; r = r + x0.w; y0.w = r;
            r = r + x1.x; y1.x = r; r = r + x1.y; y1.y = r; r = r + x1.z; y1.z = r; r = r + x1.w; y1.w = r;
            if (prefix) { *reinterpret_cast<float4 *>(buf + j) = y0; *reinterpret_cast<float4 *>(buf + j + 4) = y1; }
        }
        for (int j = n8; j < n; ++j) { r = r + buf[j]; if (prefix) buf[j] = r; }
    }
    __syncwarp();
    return __shfl_sync(FULL, r, 0);
}

// ---------------------------------------------------------------------------
// Summary of one window (elements a[0..n), n <= 256) under the exponent field e0 of its entry sum, by one warp:
// (f0, f1, neg, pos) and flags = e0 << 8 | GSC_SF_*; e0 outside 26..250 yields GSC_SF_BAD (no summary).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void gsc_window_summary(const float *a, int n, int e0, int4 &sum, int &flags) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float v[8];
    gsc_load8(a, lane * 8, n, v);
    const bool valid = e0 >= 26 && e0 <= 250;
    // a negative element below a quarter ulp of binade e0 never changes a sum that is >= 2^(e0-127): the prefix sums
    // stay monotone (fl(S ulp - t) = S ulp for t <= ulp/4, also at S = 2^23 where the spacing below halves)
    const float harmless = valid ? __uint_as_float((unsigned)(e0 - 25) << 23) : 0.0f;
    bool bad = false, hasneg = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) { hasneg |= (v[e] < 0.0f) && !(fabsf(v[e]) < harmless); bad |= !(fabsf(v[e]) < 3.0e38f); }
    int f0 = 0, f1 = 0, neg = 0, pos = 0;
    if (valid) {
        const float sc = __uint_as_float((unsigned)(277 - e0) << 23);
        const float huge = __uint_as_float((unsigned)(e0 + 2) << 23);
        int g0 = 0, g1 = 1;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int I, cls;
            gsc_classify(v[e], sc, huge, I, cls);
            bad |= cls >= 4;
            g0 = gsc_pf_step(g0, I, cls); g1 = gsc_pf_step(g1, I, cls);
            if (I < 0) neg += -I; else pos += I + 1;
        }
        g1 -= 1;
        neg = min(neg, 1 << 25); pos = min(pos, 1 << 25);   // (clamped: such a window is rejected anyway; no overflow)
        int s0 = g0, s1 = g1;
#pragma unroll
        for (int of = 1; of < 32; of <<= 1) {
            const int p0 = __shfl_up_sync(FULL, s0, of), p1 = __shfl_up_sync(FULL, s1, of);
            if (lane >= of) { int q0 = p0, q1 = p1; gsc_pf_compose(q0, q1, s0, s1); s0 = q0; s1 = q1; }
        }
        f0 = __shfl_sync(FULL, s0, 31); f1 = __shfl_sync(FULL, s1, 31);
        neg = (int)__reduce_add_sync(FULL, (unsigned)neg);
        pos = (int)__reduce_add_sync(FULL, (unsigned)pos);
    } else {
        bad = true;     // no summary for this exponent
    }
    const bool anybad = __any_sync(FULL, bad), anyneg = __any_sync(FULL, hasneg);
    sum = make_int4(f0, f1, neg, pos);
    flags = ((valid ? e0 : 0) << 8) | (anybad ? GSC_SF_BAD : 0) | (anyneg ? GSC_SF_NEG : 0);
}

// ---------------------------------------------------------------------------
// k_seed_prep: per frame, norms (load(): sum v*v left to right in float) and the norm-bucketed order.
//   pn[j]    original order (yakmo reassignment, gsc_yakmo)
//   perm[k]  original index of the point at bucketed position k; pns[k] its norm; Xs[k] its row
//   blo/bhi  per 256-point block: min / max of sqrtf(norm)
// Counting sort on the top 12 bits of the (non-negative) float norm: 16 buckets per octave, ascending.  The order
// inside a bucket is whatever the atomics give -- nothing downstream depends on it (every point is updated on its
// own, results are indexed by the original j).
// grid = F, block 512, static smem 32 KB.
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(512) k_seed_prep(const GscFrame *__restrict__ frames, const float *__restrict__ X,
                                                   float *__restrict__ pn, int *__restrict__ perm,
                                                   float *__restrict__ pns, float *__restrict__ Xs,
                                                   float *__restrict__ blo, float *__restrict__ bhi) {
    __shared__ int s_cnt[4096];
    __shared__ int s_cur[4096];
    __shared__ int s_wsum[16];
    const GscFrame f = frames[blockIdx.x];
    if (f.K <= 0) return;
    const int N = f.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *Xf = X + f.chunk_off * D;
    float *pnf = pn + f.chunk_off, *pnsf = pns + f.chunk_off, *Xsf = Xs + f.chunk_off * D;
    int *permf = perm + f.chunk_off;
    for (int b = tid; b < 4096; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    for (int j = tid; j < N; j += blockDim.x) {
        float p[D];
        gsc_load_row<D>(Xf, j, p);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) { float m = p[k] * p[k]; s = s + m; }
        pnf[j] = s;
        atomicAdd(&s_cnt[(__float_as_uint(s) >> 19) & 0xfffu], 1);
    }
    __syncthreads();
    // exclusive scan of the 4096 counts: 8 per thread, warp scan, 16 warp totals
    {
        int c[8], t = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) { c[e] = s_cnt[tid * 8 + e]; t += c[e]; }
        int incl = t;
#pragma unroll
        for (int of = 1; of < 32; of <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, of); if (lane >= of) incl += u; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int base = incl - t;
        for (int w = 0; w < warp; ++w) base += s_wsum[w];
#pragma unroll
        for (int e = 0; e < 8; ++e) { s_cur[tid * 8 + e] = base; base += c[e]; }
    }
    __syncthreads();
    for (int j = tid; j < N; j += blockDim.x) {
        const float s = pnf[j];
        const int k = atomicAdd(&s_cur[(__float_as_uint(s) >> 19) & 0xfffu], 1);
        permf[k] = j;
        pnsf[k] = s;
        float p[D];
        gsc_load_row<D>(Xf, j, p);
        if (D % 4 == 0) {
#pragma unroll
            for (int q = 0; q < D / 4; ++q)
                *reinterpret_cast<float4 *>(Xsf + (long long)k * D + 4 * q) = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < D; ++q) Xsf[(long long)k * D + q] = p[q];
        }
    }
    __syncthreads();
    // block ranges of sqrt(norm)
    const long long wo = gsc_win_off(f);
    const int nb = (N + GSC_SW - 1) / GSC_SW;
    for (int b = warp; b < nb; b += blockDim.x / 32) {
        float lo = INFINITY, hi = -INFINITY;
        for (int k = b * GSC_SW + lane; k < min(N, (b + 1) * GSC_SW); k += 32) {
            const float r = sqrtf(pnsf[k]);
            lo = fminf(lo, r); hi = fmaxf(hi, r);
        }
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, of)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, of)); }
        if (lane == 0) { blo[wo + b] = lo; bhi[wo + b] = hi; }
    }
}

// ---------------------------------------------------------------------------
// k_seed2: the seeding loop, one CTA of 256 threads per frame.
// Dynamic shared memory: chosen bitmap [ceil(N/32)] | entry sums [nw+1] | window exponent predictions [nw] |
//                        block maxima [nw] | list [nw] | dirty bitmap [ceil(nw/32)] | window flags [nw] (bytes) |
//                        prefix values of one window [256]
// ---------------------------------------------------------------------------
__host__ __device__ inline size_t gsc_seed2_smem(int N) {
    const size_t nw = (size_t)(N + GSC_SW - 1) / GSC_SW;
    return 4 * ((size_t)(N + 31) / 32) + 4 * (nw + 1) + 7 * 4 * nw + 4 * ((nw + 31) / 32) + 4 * GSC_SW + 64;
}

__device__ __forceinline__ float gsc_mkf(int e, int S) { return __uint_as_float(((unsigned)e << 23) | ((unsigned)S & 0x7fffffu)); }

// T threads per CTA: every phase of a step is a latency chain (global loads, warp scans, a 256-long FADD chain), so
// what keeps an SM busy is the number of independent frames resident on it, not the width of one CTA.
template <int D, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_seed2(const GscFrame *__restrict__ frames,
                                                          const float *__restrict__ X,     // [sumN][D] original order
                                                          const float *__restrict__ Xs,    // [sumN][D] bucketed order
                                                          const float *__restrict__ pns,   // [sumN] norms, bucketed order
                                                          const int *__restrict__ perm,    // [sumN] bucketed -> original
                                                          const float *__restrict__ blo, const float *__restrict__ bhi,
                                                          int init_type,
                                                          float *__restrict__ ups,         // [sumN] distances, bucketed order
                                                          float *__restrict__ up,          // [sumN] distances, original order
                                                          int *__restrict__ sid,           // [sumN] seed cell, original order
                                                          int4 *__restrict__ wsum,         // [windows] summaries
                                                          int *__restrict__ seeds,         // [F][Kmax] or null
                                                          float *__restrict__ cen,         // [F][Kmax][D] seeds out
                                                          float *__restrict__ cnorm,       // [F][Kmax]
                                                          int Kmax, unsigned long long *__restrict__ sdbg) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int W = T / 32;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ float s_c[D];
    __shared__ float s_cn, s_obj;
    // list counters: each is reset at a point that barriers separate from its next use
    __shared__ int s_nlist, s_ndirty;
    const GscFrame f = frames[blockIdx.x];
    const int K = f.K, N = f.N;
    if (K <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = (N + GSC_SW - 1) / GSC_SW;
    unsigned *chosen = reinterpret_cast<unsigned *>(smraw);
    float *entry = reinterpret_cast<float *>(chosen + (N + 31) / 32);      // [nw+1]: entry[w] = sum before window w, entry[nw] = obj
    int *wexp = reinterpret_cast<int *>(entry + nw + 1);                   // [nw] exponent the NEXT summary is to be made for (-1: none)
    int *wfl = wexp + nw;                                                  // [nw] exponent the summary was made for << 8 | GSC_SF_*
    float *bmax = reinterpret_cast<float *>(wfl + nw);                     // [nw] >= every current distance of the block (>= 0)
    float *s_blo = bmax + nw, *s_bhi = s_blo + nw;                         // [nw] norm range of the block
    int *list = reinterpret_cast<int *>(s_bhi + nw);                       // [nw]
    unsigned *dirty = reinterpret_cast<unsigned *>(list + nw);             // [ceil(nw/32)]
    float *wout = reinterpret_cast<float *>((reinterpret_cast<unsigned long long>(dirty + (nw + 31) / 32) + 15ull) & ~15ull);   // [256], 16-byte aligned: one staged window / its prefix values

    const float *Xf = X + f.chunk_off * D, *Xsf = Xs + f.chunk_off * D, *pnsf = pns + f.chunk_off;
    const int *permf = perm + f.chunk_off;
    float *upsf = ups + f.chunk_off, *upf = up + f.chunk_off;
    int *sidf = sid + f.chunk_off;
    const long long wo = gsc_win_off(f);
    int4 *wsumf = wsum + wo;
    float *cenf = cen + (long long)f.slot * Kmax * D;
    float *cnf = cnorm + (long long)f.slot * Kmax;

    for (int w = tid; w < (N + 31) / 32; w += T) chosen[w] = 0u;
    for (int w = tid; w < nw; w += T) {
        wexp[w] = -1; wfl[w] = GSC_SF_BAD; bmax[w] = INFINITY; entry[w] = 0.0f;
        s_blo[w] = blo[wo + w]; s_bhi[w] = bhi[wo + w];
    }
    for (int w = tid; w < (nw + 31) / 32; w += T) dirty[w] = 0xffffffffu;   // every window is summarised after step 0
    if (tid == 0) { entry[nw] = 0.0f; s_obj = 0.0f; s_nlist = 0; s_ndirty = 0; }
    GscXor128b g = {123456789ull, 362436069ull, 521288629ull, 88675123ull};
    int cached_w = -1;          // warp 0: window whose prefix values are in wout
    unsigned long long c_pick = 0, c_dist = 0, c_sum = 0, c_chain = 0, c_t = 0, n_exact = 0, n_blocks = 0, n_dirty = 0;
    __syncthreads();

    for (int i = 0; i < K; ++i) {
        if (tid == 0) c_t = clock64();
        // ================= pick (warp 0) =================
        if (warp == 0) {
            const float u = gsc_xor128b(g);      // every lane runs the generator: same state everywhere
            unsigned c;
            if (init_type == 0 || i == 0) {
                c = (unsigned)(long long)floorf(u * (float)N);
            } else {
                // std::lower_bound(r.begin(), r.end(), u * obj), probe by probe
                const float target = u * s_obj;
                int first = 0, count = N;
                while (count > 0) {
                    const int half = count >> 1, idx = first + half, w = idx / GSC_SW;
                    bool gt;                    // target > r[idx]
                    const float lo = entry[w], hi = entry[w + 1];
                    // monotone window: no element that can lower a sum of the entry's magnitude (flags made for an
                    // exponent <= the entry's), nothing non-finite
                    const int wf = wfl[w];
                    const unsigned lob = __float_as_uint(lo);
                    const bool mono = !(wf & (GSC_SF_NEG | GSC_SF_BAD)) && (lob >> 31) == 0u && (int)((lob >> 23) & 0xffu) >= (wf >> 8);
                    if (mono && target > hi) gt = true;             // r[idx] <= exit sum < target
                    else if (mono && !(target > lo)) gt = false;    // r[idx] >= entry sum >= target
                    else {
                        if (cached_w != w) {
                            gsc_warp_chain(upf + w * GSC_SW, min(GSC_SW, N - w * GSC_SW), lo, wout, true);
                            cached_w = w;
                        }
                        gt = target > wout[idx - w * GSC_SW];
                    }
                    if (gt) { first = idx + 1; count -= half + 1; } else count = half;
                }
                c = (unsigned)first;
            }
            if (lane == 0) {
                while (c < (unsigned)N && ((chosen[c >> 5] >> (c & 31)) & 1u)) c = (c >= (unsigned)(N - 1)) ? 0u : c + 1u;   // RVA 0x1b20-0x1bd8
                if (c >= (unsigned)N) c = (unsigned)(N - 1);
                chosen[c >> 5] |= 1u << (c & 31);
                if (seeds) seeds[(long long)f.slot * Kmax + i] = (int)c;
                float r[D];
                gsc_load_row<D>(Xf, (long long)c, r);
                float sn = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    s_c[k] = r[k]; cenf[(long long)i * D + k] = r[k];
                    const float m = r[k] * r[k]; sn = sn + m;       // the norm load() cached for this point
                }
                s_cn = sn; cnf[i] = sn;
            }
            cached_w = -1;      // the distances change below
        }
        __syncthreads();
        if (tid == 0) { const unsigned long long t1 = clock64(); c_pick += t1 - c_t; c_t = t1; }
        // ================= distance pass =================
        const float cn = s_cn, scn = sqrtf(cn);
        const float mrg = 4e-6f;
        float crow[D];
#pragma unroll
        for (int k = 0; k < D; ++k) crow[k] = s_c[k];
        if (i == 0) {
            for (int b = warp; b < nw; b += W) {
                float mx = 0.0f;
                for (int k = b * GSC_SW + lane; k < min(N, (b + 1) * GSC_SW); k += 32) {
                    float p[D];
                    gsc_load_row<D>(Xsf, k, p);
                    const float d = gsc_yakmo_dist<D>(p, pnsf[k], crow, cn);
                    const int j = permf[k];
                    upsf[k] = d; upf[j] = d; sidf[j] = 0;
                    mx = fmaxf(mx, d);      // (NaN ignored)
                }
                mx = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(fmaxf(mx, 0.0f))));
                if (lane == 0) bmax[b] = mx;
            }
        } else {
            // blocks whose norm range is close enough to |c| for a distance to drop
            for (int b0 = 0; b0 < nw; b0 += T) {
                const int b = b0 + tid;
                bool take = false;
                if (b < nw) {
                    const float lo = s_blo[b], hi = s_bhi[b];
                    const float gp = fmaxf(lo - scn, scn - hi);
                    // every point of the block has (sqrt(pn) - scn)^2 (1 - mrg) - mrg (cn + pn) - 1e-37 >= lbb
                    const float lbb = gp * gp * (1.0f - 2.0f * mrg) - mrg * (cn + hi * hi * 1.000001f) - 1e-37f;
                    take = !(gp > 0.0f && lbb > bmax[b]);
                }
                const unsigned m = __ballot_sync(FULL, take);
                int basei = 0;
                if (lane == 0 && m) basei = atomicAdd(&s_nlist, __popc(m));
                basei = __shfl_sync(FULL, basei, 0);
                if (take) list[basei + __popc(m & ((1u << lane) - 1u))] = b;
            }
            __syncthreads();
            const int nl = s_nlist;
            if (tid == 0) n_blocks += nl;
            for (int li = warp; li < nl; li += W) {
                const int kb = list[li] * GSC_SW;
                const int kend = min(N, kb + GSC_SW);
                // the norms and current distances of the lane's 8 points first (16 loads in flight), then the rows of
                // the points the seed can still reach, two at a time.  A row is only fetched when the new seed can
                // lower the distance: d >= (|p| - |c|)^2, and yakmo's float evaluation of d stays within a few ulps of
                // (cn + pn) of the true value; the margins cover both, so `up > d` is false for every skipped point
                float pj[8], uj[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = kb + q * 32 + lane;
                    pj[q] = (k < kend) ? pnsf[k] : 0.0f;
                    uj[q] = (k < kend) ? upsf[k] : -INFINITY;
                }
                unsigned pass = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float t = sqrtf(pj[q]) - scn;
                    const float lb = t * t * (1.0f - mrg) - mrg * (cn + pj[q]) - 1e-37f;
                    if (kb + q * 32 + lane < kend && !(lb > uj[q])) pass |= 1u << q;
                }
#pragma unroll
                for (int h = 0; h < 8; h += 2) {
                    float p[2][D];
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if ((pass >> (h + q)) & 1u) gsc_load_row<D>(Xsf, kb + (h + q) * 32 + lane, p[q]);
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if ((pass >> (h + q)) & 1u) {
                            const int k = kb + (h + q) * 32 + lane;
                            const float d = gsc_yakmo_dist<D>(p[q], pj[h + q], crow, cn);
                            if (uj[h + q] > d) {
                                const int j = permf[k];
                                upsf[k] = d; upf[j] = d; sidf[j] = i;
                                atomicOr(&dirty[(j / GSC_SW) >> 5], 1u << ((j / GSC_SW) & 31));
                                uj[h + q] = d;
                            }
                        }
                }
                float mx = 0.0f;
#pragma unroll
                for (int q = 0; q < 8; ++q) mx = fmaxf(mx, uj[q]);     // (-inf for the lanes past the end, NaN ignored)
                mx = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(fmaxf(mx, 0.0f))));
                if (lane == 0) bmax[list[li]] = mx;
            }
        }
        __syncthreads();
        if (tid == 0) { s_nlist = 0; const unsigned long long t1 = clock64(); c_dist += t1 - c_t; c_t = t1; }   // (next use: next step)
        if (i < K - 1 && init_type == 1) {
            // ================= summaries of the dirty windows =================
            for (int w0 = 0; w0 < nw; w0 += T) {
                const int w = w0 + tid;
                const bool take = (w < nw) && ((dirty[w >> 5] >> (w & 31)) & 1u);
                const unsigned m = __ballot_sync(FULL, take);
                int basei = 0;
                if (lane == 0 && m) basei = atomicAdd(&s_ndirty, __popc(m));
                basei = __shfl_sync(FULL, basei, 0);
                if (take) list[basei + __popc(m & ((1u << lane) - 1u))] = w;
            }
            __syncthreads();
            {
                const int nl = s_ndirty;
                if (tid == 0) n_dirty += nl;
                for (int w = tid; w < (nw + 31) / 32; w += T) dirty[w] = 0u;
                for (int li = warp; li < nl; li += W) {
                    const int w = list[li];
                    int4 sm;
                    int fl;
                    gsc_window_summary(upf + w * GSC_SW, min(GSC_SW, N - w * GSC_SW), wexp[w], sm, fl);
                    if (lane == 0) { wsumf[w] = sm; wfl[w] = fl; }
                }
            }
            __syncthreads();
            if (tid == 0) { s_ndirty = 0; const unsigned long long t1 = clock64(); c_sum += t1 - c_t; c_t = t1; }   // (next use: next step)
            // ================= chain (warp 0): exact entry sum of every window, exact total =================
            // 32 windows at a time, one per lane: the summaries that can apply in the running sum's binade are composed
            // by a warp scan, every lane checks its own window against the entry sum the scan gives it, and the windows
            // before the first one that fails are final.  That window is evaluated element by element (gsc_warp_chain);
            // the scan resumes behind it.
            if (warp == 0) {
                float run = 0.0f;
                for (int wb = 0; wb < nw; wb += 32) {
                    const int wl = wb + lane, cntw = min(32, nw - wb);
                    int4 sm = make_int4(0, 0, 0, 0);
                    int ep = -2, fl = GSC_SF_BAD, fe = 0;
                    if (wl < nw) { sm = wsumf[wl]; ep = wexp[wl]; const int wf = wfl[wl]; fl = wf & 3; fe = wf >> 8; }
                    float my_entry = 0.0f;
                    int my_newexp = ep;
                    int start = 0;
                    while (start < cntw) {
                        const unsigned rb = __float_as_uint(run);
                        const int ea = (int)((rb >> 23) & 0xffu);
                        const int S0 = (int)((rb & 0x7fffffu) | 0x800000u);
                        const bool normal = (rb >> 31) == 0u && ea >= 26 && ea <= 250;
                        int firstbad = start;
                        if (normal) {
                            const bool in = lane >= start && lane < cntw;
                            const bool cand = in && !(fl & GSC_SF_BAD) && fe == ea;    // a summary made for this binade
                            int s0 = cand ? sm.x : 0, s1 = cand ? sm.y : 0;            // (identity for the others)
#pragma unroll
                            for (int of = 1; of < 32; of <<= 1) {
                                const int p0 = __shfl_up_sync(FULL, s0, of), p1 = __shfl_up_sync(FULL, s1, of);
                                if (lane >= of) { int q0 = p0, q1 = p1; gsc_pf_compose(q0, q1, s0, s1); s0 = q0; s1 = q1; }
                            }
                            int x0 = __shfl_up_sync(FULL, s0, 1), x1 = __shfl_up_sync(FULL, s1, 1);
                            if (lane == 0) { x0 = 0; x1 = 0; }
                            const int Sen = S0 + ((S0 & 1) ? x1 : x0);      // entry of this lane's window if all before it hold
                            const bool ok = !in || (cand && Sen - sm.z >= (1 << 23) && Sen + sm.w < (1 << 24));
                            const unsigned badm = __ballot_sync(FULL, !ok);
                            firstbad = badm ? (__ffs(badm) - 1) : cntw;
                            if (in && lane < firstbad) my_entry = gsc_mkf(ea, Sen);
                            const int e0 = __shfl_sync(FULL, s0, 31), e1 = __shfl_sync(FULL, s1, 31);
                            const int Sfb = __shfl_sync(FULL, Sen, firstbad & 31);
                            run = gsc_mkf(ea, (firstbad < cntw) ? Sfb : S0 + ((S0 & 1) ? e1 : e0));
                        }
                        if (firstbad < cntw) {
                            const int w = wb + firstbad;
                            const unsigned rb2 = __float_as_uint(run);
                            const int ea2 = (int)((rb2 >> 23) & 0xffu);
                            if (lane == firstbad) my_entry = run;
                            run = gsc_warp_chain(upf + w * GSC_SW, min(GSC_SW, N - w * GSC_SW), run, wout, false);
                            if (lane == 0) ++n_exact;
                            // a summary made for another exponent (or none): make it for this one at the next step
                            const int eu = ((rb2 >> 31) == 0u && ea2 >= 26 && ea2 <= 250) ? ea2 : -1;
                            if (lane == firstbad && eu != ep) my_newexp = eu;
                            start = firstbad + 1;
                        } else {
                            start = cntw;
                        }
                    }
                    if (wl < nw) {
                        entry[wl] = my_entry;
                        if (my_newexp != ep) { wexp[wl] = my_newexp; atomicOr(&dirty[wl >> 5], 1u << (wl & 31)); }
                    }
                }
                if (lane == 0) { entry[nw] = run; s_obj = run; }
            }
            __syncthreads();
            if (tid == 0) { const unsigned long long t1 = clock64(); c_chain += t1 - c_t; c_t = t1; }
        }
    }
    if (tid == 0 && sdbg) {
        unsigned long long *o = sdbg + (long long)f.slot * 8;
        o[0] = c_pick; o[1] = c_dist; o[2] = c_sum + c_chain; o[3] = K > 0 ? (unsigned long long)K : 1ull;
        o[4] = n_exact; o[5] = n_blocks; o[6] = n_dirty; o[7] = c_chain;
    }
}
