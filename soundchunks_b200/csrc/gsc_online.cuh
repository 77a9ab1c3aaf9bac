// gsc_online.cuh -- K4/K5: the reference's online k-means (enc:699-765
// TFrame.KNNScanReduce), one point at a time against the LIVE centroids.
//
// The rule is sequential in the points (each update moves one centroid before
// the next query), so the parallelism is (a) across the K centroids inside a
// frame and (b) across the frames of a batch: ONE CTA PER FRAME.
//
// Per point (arithmetic of the result is the reference's, bit for bit):
//   best  := argmin_c  d(x, c),  d = sum_k (x_k - c_k)^2   float, left to
//            right, no FMA (ANN annkSearch, eps = 0; lowest index on ties)
//   rate  := Single(1 / sqrt(cnt_prev[best]))               enc:735
//   c_best += (x - c_best) * rate                            enc:736-740
//   err   += sqrt(d / D)   (Single sqrt, Double accumulate)  enc:743
//
// Data placement (K = 4096, D = 8):
//   shared memory  exact centroid rows K x 8 fp32 = 128 KB (the truth; read
//                  only for exact scoring and updates), per-centroid rate
//                  16 KB, counts 2 x 16 KB, point tile, candidate lists
//   registers      FILTER copy: first DF = 4 dims of the thread's CPT = 16
//                  centroids (64 regs) + h_c = -0.5|c|^2(1-g) (16 regs)
//   HBM            one 32-byte row per point per pass (streamed through smem)
//
// Schedule: points are taken in batches of up to B = 8.
//  Phase 1 (all warps, codebook frozen).  For every point b of the batch each
//    thread evaluates, over the first DF dimensions of its centroids,
//        s_c = x.c + h_c                          (DF FFMAs per centroid)
//    which certifies the LOWER bound
//        lb_c = |x|^2(1-g) - 2 s_c <= sum_{k<DF}(x_k-c_k)^2 <= d(x,c)
//    (dropped squared terms are >= 0 and tiny -- the 1e-5-scaled cepstral
//    features, enc:362; g = 2^-17 covers every rounding of both forms).
//    Centroids with lb_c <= U_b are scored in the exact operation order and
//    their (d bits << 32 | index) keys appended to the point's list.  U_b is
//    the exact distance to the centroid the point chose in the previous pass
//    (its seed cell in pass 0).
//  Phase 2 (one warp, lane b = point b, strictly in point order).  The winner
//    of point t is the minimum key over (a) its list minus entries of
//    centroids moved by points 0..t-1 of this batch and (b) fresh exact
//    distances to those moved centroids; it is ACCEPTED iff d_win <= U_t --
//    then every unmoved centroid with d <= d_win had lb <= U_t and was scored,
//    and every moved one is scored fresh, so the key minimum is the exact
//    argmin with the lowest index on ties.  If the test fails (or a list
//    overflowed) the batch is cut before point t, which then leads the next
//    batch with a fresh bound (always sufficient); a cut at t = 0 switches to
//    an exhaustive scoring of that one point.  After each accepted point the
//    row, counts, label and error term are updated and the later lanes score
//    their points against the new row.
// Two block barriers per batch.  Shared memory is addressed through an opaque
// 32-bit base (inline ld/st.shared) so the shared-window base is not
// rematerialised (S2UR SR_CgaCtaId) inside the loops.
#pragma once
#include "gsc_device.cuh"

#define GSC_ON_TP 256         // points per shared-memory tile
#define GSC_ON_B 8            // points per batch
#define GSC_ON_L 16           // candidate list capacity per point
#define GSC_ON_G 7.62939453125e-06f   // 2^-17
#define GSC_NONE 0xffffffffu
#define GSC_KNONE 0xffffffffffffffffull

// ---- shared-memory access through an opaque 32-bit address -------------------
__device__ __forceinline__ unsigned gsc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned gsc_opaque(unsigned v) { unsigned r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ float gsc_lds_f(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int gsc_lds_i(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ unsigned long long gsc_lds_u64(unsigned a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double gsc_lds_d(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 gsc_lds_f4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void gsc_sts_f(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void gsc_sts_i(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void gsc_sts_u64(unsigned a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void gsc_sts_d(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void gsc_sts_f4(unsigned a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int gsc_atoms_add(unsigned a, int v) { int o; asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }

template <int D>
__device__ __forceinline__ void gsc_lds_row(unsigned a, float (&r)[D]) {
    if (D % 4 == 0) {
#pragma unroll
        for (int k = 0; k < D / 4; ++k) {
            const float4 t = gsc_lds_f4(a + 16u * k);
            r[4 * k] = t.x; r[4 * k + 1] = t.y; r[4 * k + 2] = t.z; r[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) r[k] = gsc_lds_f(a + 4u * k);
    }
}
template <int D>
__device__ __forceinline__ void gsc_sts_row(unsigned a, const float (&r)[D]) {
    if (D % 4 == 0) {
#pragma unroll
        for (int k = 0; k < D / 4; ++k) gsc_sts_f4(a + 16u * k, make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]));
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) gsc_sts_f(a + 4u * k, r[k]);
    }
}

__device__ __forceinline__ unsigned long long gsc_pack(unsigned dbits, unsigned idx) {
    return ((unsigned long long)dbits << 32) | idx;
}
__device__ __forceinline__ unsigned gsc_kd(unsigned long long k) { return (unsigned)(k >> 32); }
__device__ __forceinline__ unsigned gsc_ki(unsigned long long k) { return (unsigned)(k & 0xffffffffu); }
__device__ __forceinline__ float gsc_rate(int cnt) {  // enc:735
    return (float)(1.0 / sqrt((double)cnt));
}

// shared-memory layout (byte offsets from the base)
template <int D, int CPT, int T>
struct GscOnLayout {
    static constexpr int KP = T * CPT;
    static constexpr unsigned X = 0;                                  // float [TP][D]
    static constexpr unsigned HX = X + GSC_ON_TP * D * 4;             // float [TP]
    static constexpr unsigned G = HX + GSC_ON_TP * 4;                 // int   [TP]
    static constexpr unsigned ET = G + GSC_ON_TP * 4;                 // float [TP]
    static constexpr unsigned LIST = ET + GSC_ON_TP * 4;              // u64   [B][L]
    static constexpr unsigned LISTN = LIST + GSC_ON_B * GSC_ON_L * 8; // int   [B]
    static constexpr unsigned MOVED = LISTN + GSC_ON_B * 4;           // int   [B]
    static constexpr unsigned WKEY = MOVED + GSC_ON_B * 4;            // u64   [32]
    static constexpr unsigned NMOVED = WKEY + 32 * 8;                 // int
    static constexpr unsigned POSN = NMOVED + 4;                      // int
    static constexpr unsigned EXH = POSN + 4;                         // int
    static constexpr unsigned STOP = EXH + 4;                         // int
    static constexpr unsigned ERR = STOP + 4;                         // double (8-aligned)
    static constexpr unsigned RATE = ((ERR + 8 + 15) / 16) * 16;      // float [KP]
    static constexpr unsigned CNT = RATE + KP * 4;                    // int   [2][KP]
    static constexpr unsigned C = ((CNT + 2 * KP * 4 + 15) / 16) * 16;  // float [KP][D]
    static constexpr unsigned TOTAL = C + KP * D * 4;
};

template <int D, int CPT, int T>
__global__ void __launch_bounds__(T) k_online(const GscFrame *__restrict__ frames,
                                              const float *__restrict__ X,       // [sumN][D]
                                              float *__restrict__ cen,           // [F][Kmax][D] in/out
                                              int *__restrict__ labels,          // [sumN] in: guesses, out: labels
                                              int *__restrict__ passes_out,      // [F]
                                              double *__restrict__ err_out,      // [F]
                                              double tol, int max_passes, int Kmax, int force_exact,
                                              unsigned long long *__restrict__ dbg) {
    using Ly = GscOnLayout<D, CPT, T>;
    constexpr int DF = (D >= 8) ? D / 2 : D;   // filter dimensions
    constexpr int W = T / 32;
    constexpr int KP = T * CPT;
    constexpr int B = GSC_ON_B, L = GSC_ON_L;
    extern __shared__ __align__(16) unsigned char smraw[];
    const unsigned sb = gsc_opaque(gsc_smem_u32(smraw));

    const GscFrame f = frames[blockIdx.x];
    const int K = f.K, N = f.N;
    if (K <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int first = tid * CPT;
    const float *Xf = X + f.chunk_off * D;
    int *lab = labels + f.chunk_off;
    float *cf = cen + (long long)f.slot * Kmax * D;

    // codebook -> shared rows + register filter copy; dead slots (idx >= K) are NaN and never win
    float fc[CPT][DF], h[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int idx = first + j;
        float r[D];
        float nc = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r[k] = (idx < K) ? cf[(long long)idx * D + k] : __int_as_float(0x7fc00000);
            if (k < DF) { fc[j][k] = r[k]; nc = fmaf(r[k], r[k], nc); }
        }
        gsc_sts_row<D>(sb + Ly::C + (unsigned)idx * D * 4, r);
        h[j] = -0.5f * nc * (1.0f - GSC_ON_G);
    }
    for (int j = tid; j < 2 * KP; j += T) gsc_sts_i(sb + Ly::CNT + 4u * j, 1);  // enc:717-721
    if (tid < B) { gsc_sts_i(sb + Ly::LISTN + 4u * tid, 0); gsc_sts_i(sb + Ly::MOVED + 4u * tid, 0); }
    if (tid == 0) {
        gsc_sts_d(sb + Ly::ERR, 3.40282346638528860e+38);
        gsc_sts_i(sb + Ly::STOP, 0); gsc_sts_i(sb + Ly::NMOVED, 0); gsc_sts_i(sb + Ly::EXH, 0); gsc_sts_i(sb + Ly::POSN, 0);
    }
    __syncthreads();

    unsigned long long c_ph1 = 0, c_ph2 = 0, c_t0 = 0; unsigned long long c_batches = 0, c_exh = 0, c_cut_verify = 0, c_cut_over = 0, c_points = 0, c_cands = 0;
    int iter = 0;
    for (;;) {
        const int odd = iter & 1;
        const unsigned cnt_prev = sb + Ly::CNT + (odd ? 0u : (unsigned)KP * 4u);   // cnts[not Odd(iter)]
        const unsigned cnt_cur = sb + Ly::CNT + (odd ? (unsigned)KP * 4u : 0u);    // cnts[Odd(iter)]
        const double prevErr = gsc_lds_d(sb + Ly::ERR);                            // uniform copy
        // rate of every centroid for this pass (cnt_prev is constant during a pass), enc:735
        for (int j = tid; j < KP; j += T) gsc_sts_f(sb + Ly::RATE + 4u * j, gsc_rate(gsc_lds_i(cnt_prev + 4u * j)));
        __syncthreads();
        if (tid == 0) gsc_sts_d(sb + Ly::ERR, 0.0);

        for (int base = 0; base < N; base += GSC_ON_TP) {
            const int tn = min(GSC_ON_TP, N - base);
            __syncthreads();  // (A) previous tile fully consumed
            for (int t = tid; t < tn * D; t += T) gsc_sts_f(sb + Ly::X + 4u * t, Xf[(long long)base * D + t]);
            for (int t = tid; t < tn; t += T) {
                int gg = lab[base + t];
                gsc_sts_i(sb + Ly::G + 4u * t, (gg < 0 || gg >= K) ? 0 : gg);
            }
            __syncthreads();  // (B)
            for (int t = tid; t < tn; t += T) {
                float nx = 0.0f;
#pragma unroll
                for (int k = 0; k < DF; ++k) { const float v = gsc_lds_f(sb + Ly::X + (unsigned)(t * D + k) * 4u); nx = fmaf(v, v, nx); }
                gsc_sts_f(sb + Ly::HX + 4u * t, 0.5f * nx * (1.0f - GSC_ON_G) - 1e-30f);
            }
            __syncthreads();  // (C)

            int pos = 0;
            while (pos < tn) {
                const int nb = min(B, tn - pos);
                const int exh = force_exact ? 1 : gsc_lds_i(sb + Ly::EXH);
                if (tid == 0) c_t0 = clock64();
                // ============ phase 1: all warps, codebook frozen ============
                {   // refresh the filter copy of centroids moved by the previous batch
                    const int nm = gsc_lds_i(sb + Ly::NMOVED);
                    for (int m = 0; m < nm; ++m) {
                        const int id = gsc_lds_i(sb + Ly::MOVED + 4u * m);
                        if (id >= first && id < first + CPT) {
                            float r[D];
                            gsc_lds_row<D>(sb + Ly::C + (unsigned)id * D * 4, r);
                            float nc = 0.0f;
#pragma unroll
                            for (int k = 0; k < DF; ++k) nc = fmaf(r[k], r[k], nc);
                            const float hn = -0.5f * nc * (1.0f - GSC_ON_G);
                            const unsigned um = 1u << (id - first);   // bit mask, not `id - first == j` (see gsc_device.cuh)
#pragma unroll
                            for (int j = 0; j < CPT; ++j)
                                if (um & (1u << j)) {
#pragma unroll
                                    for (int k = 0; k < DF; ++k) fc[j][k] = r[k];
                                    h[j] = hn;
                                }
                        }
                    }
                }
                float Umine = INFINITY;   // lane b of every warp: bound of point pos+b
                if (exh) {
                    // exhaustive scoring of ONE point (validation mode, or a batch cut at its first point)
                    float x[D];
                    gsc_lds_row<D>(sb + Ly::X + (unsigned)pos * D * 4, x);
                    unsigned bd = GSC_NONE, bi = GSC_NONE;
                    for (int j = 0; j < CPT; ++j) {
                        float r[D];
                        gsc_lds_row<D>(sb + Ly::C + (unsigned)(first + j) * D * 4, r);
                        const float d = gsc_ann_dist<D>(x, r);
                        const unsigned b = __float_as_uint(d);
                        if (d == d && b < bd) { bd = b; bi = (unsigned)(first + j); }   // strict <: lowest index wins
                    }
                    const unsigned m = __reduce_min_sync(0xffffffffu, bd);
                    const unsigned mi = __reduce_min_sync(0xffffffffu, bd == m ? bi : GSC_NONE);
                    if (lane == 0) gsc_sts_u64(sb + Ly::WKEY + 8u * warp, gsc_pack(m, mi));
                } else {
                    if (lane < nb) {
                        const int p = pos + lane;
                        const int g = gsc_lds_i(sb + Ly::G + 4u * p);
                        float x[D], r[D];
                        gsc_lds_row<D>(sb + Ly::X + (unsigned)p * D * 4, x);
                        gsc_lds_row<D>(sb + Ly::C + (unsigned)g * D * 4, r);
                        const float d = gsc_ann_dist<D>(x, r);
                        Umine = (d == d) ? d : INFINITY;
                    }
#pragma unroll
                    for (int b = 0; b < B; ++b) {
                        if (b < nb) {
                            const int p = pos + b;
                            const float U = __shfl_sync(0xffffffffu, Umine, b);
                            const float thr = gsc_lds_f(sb + Ly::HX + 4u * p) - 0.5f * U;   // candidate iff s >= thr (lb <= U)
                            float xq[DF];
                            if (DF == 4) {
                                const float4 t4 = gsc_lds_f4(sb + Ly::X + (unsigned)p * D * 4);
                                xq[0] = t4.x; xq[1] = t4.y; xq[2] = t4.z; xq[3] = t4.w;
                            } else {
#pragma unroll
                                for (int k = 0; k < DF; ++k) xq[k] = gsc_lds_f(sb + Ly::X + (unsigned)(p * D + k) * 4u);
                            }
                            float s[CPT];
#pragma unroll
                            for (int j = 0; j < CPT; ++j) s[j] = h[j];
#pragma unroll
                            for (int k = 0; k < DF; ++k)
#pragma unroll
                                for (int j = 0; j < CPT; ++j) s[j] = fmaf(xq[k], fc[j][k], s[j]);
                            float smax = s[0];
#pragma unroll
                            for (int j = 1; j < CPT; ++j) smax = fmaxf(smax, s[j]);   // NaN-safe: fmaxf ignores NaN
                            if (smax >= thr) {
                                unsigned m = 0;
#pragma unroll
                                for (int j = 0; j < CPT; ++j) m |= (s[j] >= thr) ? (1u << j) : 0u;
                                float x[D];
                                gsc_lds_row<D>(sb + Ly::X + (unsigned)p * D * 4, x);
                                while (m) {
                                    const int j = __ffs(m) - 1;
                                    m &= m - 1;
                                    float r[D];
                                    gsc_lds_row<D>(sb + Ly::C + (unsigned)(first + j) * D * 4, r);
                                    const float d = gsc_ann_dist<D>(x, r);
                                    if (d == d) {
                                        const int slot = gsc_atoms_add(sb + Ly::LISTN + 4u * b, 1);
                                        if (slot < L) gsc_sts_u64(sb + Ly::LIST + (unsigned)(b * L + slot) * 8u, gsc_pack(__float_as_uint(d), (unsigned)(first + j)));
                                    }
                                }
                            }
                        }
                    }
                }
                __syncthreads();   // ---- bar A: lists / warp keys complete ----
                if (tid == 0) { const unsigned long long t1 = clock64(); c_ph1 += t1 - c_t0; c_t0 = t1; }
                // ============ phase 2: warp 0 resolves the batch in point order ============
                if (warp == 0) {
                    int done = 0;
                    int wids[B];
#pragma unroll
                    for (int t = 0; t < B; ++t) wids[t] = -1;
                    if (exh) {
                        unsigned long long k = (lane < W) ? gsc_lds_u64(sb + Ly::WKEY + 8u * lane) : GSC_KNONE;
                        unsigned kd = gsc_kd(k), ki = gsc_ki(k);
                        const unsigned m = __reduce_min_sync(0xffffffffu, kd);
                        const unsigned mi = __reduce_min_sync(0xffffffffu, kd == m ? ki : GSC_NONE);
                        const int w = (m == GSC_NONE) ? 0 : (int)mi;
                        const float dwin = (m == GSC_NONE) ? INFINITY : __uint_as_float(m);
                        float xt[D], r[D];
                        gsc_lds_row<D>(sb + Ly::X + (unsigned)pos * D * 4, xt);
                        gsc_lds_row<D>(sb + Ly::C + (unsigned)w * D * 4, r);
                        const float rate = gsc_lds_f(sb + Ly::RATE + 4u * w);
#pragma unroll
                        for (int k2 = 0; k2 < D; ++k2) { float v = xt[k2] - r[k2]; float mm = v * rate; r[k2] = r[k2] + mm; }   // enc:736-740
                        if (lane == 0) {
                            gsc_sts_row<D>(sb + Ly::C + (unsigned)w * D * 4, r);
                            gsc_sts_i(cnt_cur + 4u * w, gsc_lds_i(cnt_cur + 4u * w) + 1);     // enc:744
                            lab[base + pos] = w;                                              // enc:742
                            gsc_sts_f(sb + Ly::ET + 4u * pos, sqrtf(dwin / (float)D));        // enc:743 (term)
                        }
                        wids[0] = w;
                        done = 1;
                    } else {
                        // lane b < nb: top-2 (distinct ids) of its list, its bound and its point
                        unsigned long long top1 = GSC_KNONE, top2 = GSC_KNONE;
                        int nl = 0, over = 0;
                        float xb[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) xb[k] = 0.0f;
                        if (lane < nb) {
                            nl = gsc_lds_i(sb + Ly::LISTN + 4u * lane);
                            over = nl > L;
                            nl = min(nl, L);
                            for (int e = 0; e < nl; ++e) {
                                const unsigned long long k = gsc_lds_u64(sb + Ly::LIST + (unsigned)(lane * L + e) * 8u);
                                if (k < top1) { top2 = top1; top1 = k; } else if (k < top2) top2 = k;
                            }
                            gsc_lds_row<D>(sb + Ly::X + (unsigned)(pos + lane) * D * 4, xb);
                        }
                        unsigned long long fresh[B];
#pragma unroll
                        for (int t = 0; t < B; ++t) fresh[t] = GSC_KNONE;
#pragma unroll
                        for (int t = 0; t < B; ++t) {
                            if (t < nb && done == t) {   // uniform
                                // ---- lane t: its exact winner under the current state ----
                                unsigned long long key = GSC_KNONE;
                                int ok = 0;
                                if (lane == t) {
                                    unsigned long long bl = top1;
                                    bool m1 = false, m2 = false;
#pragma unroll
                                    for (int u = 0; u < B; ++u) if (u < t) { m1 |= ((int)gsc_ki(top1) == wids[u]); m2 |= ((int)gsc_ki(top2) == wids[u]); }
                                    if (m1) {
                                        bl = top2;
                                        if (m2 || top2 == GSC_KNONE) {
                                            // both leaders stale: rescan the list without moved centroids
                                            bl = GSC_KNONE;
                                            for (int e = 0; e < nl; ++e) {
                                                const unsigned long long k = gsc_lds_u64(sb + Ly::LIST + (unsigned)(lane * L + e) * 8u);
                                                bool mv = false;
#pragma unroll
                                                for (int u = 0; u < B; ++u) if (u < t) mv |= ((int)gsc_ki(k) == wids[u]);
                                                if (!mv && k < bl) bl = k;
                                            }
                                        }
                                    }
                                    key = bl;
#pragma unroll
                                    for (int u = 0; u < B; ++u) if (u < t && fresh[u] < key) key = fresh[u];
                                    ok = (!over) && (key != GSC_KNONE) && (__uint_as_float(gsc_kd(key)) <= Umine);
                                }
                                ok = __shfl_sync(0xffffffffu, ok, t);
                                if (!ok) { if (__shfl_sync(0xffffffffu, over, t)) ++c_cut_over; else ++c_cut_verify; }
                                if (ok) {   // uniform
                                    const int w = (int)__shfl_sync(0xffffffffu, gsc_ki(key), t);
                                    const unsigned dbits = __shfl_sync(0xffffffffu, gsc_kd(key), t);
                                    // ---- update the winner (enc:735-744); every lane holds the new row ----
                                    float xt[D], r[D];
                                    gsc_lds_row<D>(sb + Ly::X + (unsigned)(pos + t) * D * 4, xt);
                                    gsc_lds_row<D>(sb + Ly::C + (unsigned)w * D * 4, r);
                                    const float rate = gsc_lds_f(sb + Ly::RATE + 4u * w);
#pragma unroll
                                    for (int k2 = 0; k2 < D; ++k2) { float v = xt[k2] - r[k2]; float mm = v * rate; r[k2] = r[k2] + mm; }
                                    if (lane == 0) {
                                        gsc_sts_row<D>(sb + Ly::C + (unsigned)w * D * 4, r);
                                        gsc_sts_i(cnt_cur + 4u * w, gsc_lds_i(cnt_cur + 4u * w) + 1);
                                        lab[base + pos + t] = w;
                                        gsc_sts_f(sb + Ly::ET + 4u * (pos + t), sqrtf(__uint_as_float(dbits) / (float)D));
                                    }
                                    __syncwarp();   // the new row is visible to every lane's later reads
                                    // a centroid moved twice: its older fresh keys are stale
#pragma unroll
                                    for (int u = 0; u < B; ++u) if (u < t && wids[u] == w) fresh[u] = GSC_KNONE;
                                    wids[t] = w;
                                    // ---- later points score the moved centroid in its new position ----
                                    if (lane > t && lane < nb) {
                                        const float dn = gsc_ann_dist<D>(xb, r);
                                        if (dn == dn) fresh[t] = gsc_pack(__float_as_uint(dn), (unsigned)w);
                                    }
                                    done = t + 1;
                                }
                            }
                        }
                    }
                    ++c_batches; c_points += done; if (exh) ++c_exh;
                    if (!exh) { int nn = (lane < nb) ? gsc_lds_i(sb + Ly::LISTN + 4u * lane) : 0; for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o); c_cands += nn; }
                    // ---- hand over to the next batch ----
#pragma unroll
                    for (int t = 0; t < B; ++t) if (lane == 0 && t < done) gsc_sts_i(sb + Ly::MOVED + 4u * t, wids[t]);
                    if (lane < B) gsc_sts_i(sb + Ly::LISTN + 4u * lane, 0);
                    if (lane == 0) {
                        gsc_sts_i(sb + Ly::NMOVED, done);
                        gsc_sts_i(sb + Ly::POSN, pos + done);
                        gsc_sts_i(sb + Ly::EXH, done == 0 ? 1 : 0);
                    }
                }
                __syncthreads();   // ---- bar B ----
                if (tid == 0) c_ph2 += clock64() - c_t0;
                pos = gsc_lds_i(sb + Ly::POSN);
            }
            __syncthreads();  // (E) all terms of the tile written
            if (tid == 0) {
                double e = gsc_lds_d(sb + Ly::ERR);                            // enc:743 (Double sum, point order)
                for (int p = 0; p < tn; ++p) e += (double)gsc_lds_f(sb + Ly::ET + 4u * p);
                gsc_sts_d(sb + Ly::ERR, e);
                gsc_sts_i(sb + Ly::POSN, 0);
            }
        }
        // ---- end of pass: enc:754-761 ----
        __syncthreads();
        for (int j = tid; j < KP; j += T) gsc_sts_i(cnt_prev + 4u * j, 1);
        ++iter;
        if (tid == 0) {
            const double e = gsc_lds_d(sb + Ly::ERR);
            const bool same = (e > prevErr) ? ((e - prevErr) <= tol) : ((prevErr - e) <= tol);
            gsc_sts_i(sb + Ly::STOP, (same || iter >= max_passes) ? 1 : 0);
        }
        __syncthreads();
        if (gsc_lds_i(sb + Ly::STOP)) break;
    }
    // the filter copies of the last batch's moved centroids are stale, the shared rows are the truth
    for (int j = 0; j < CPT; ++j) {
        const int idx = first + j;
        if (idx < K) {
            float r[D];
            gsc_lds_row<D>(sb + Ly::C + (unsigned)idx * D * 4, r);
#pragma unroll
            for (int k = 0; k < D; ++k) cf[(long long)idx * D + k] = r[k];
        }
    }
    if (tid == 0) {
        passes_out[f.slot] = iter;
        err_out[f.slot] = gsc_lds_d(sb + Ly::ERR);
        if (dbg) {
            unsigned long long *o = dbg + (long long)f.slot * 8;
            o[0] = c_batches; o[1] = c_points; o[2] = c_exh; o[3] = c_cut_verify; o[4] = c_cut_over; o[5] = c_cands; o[6] = c_ph1; o[7] = c_ph2;
        }
    }
}
