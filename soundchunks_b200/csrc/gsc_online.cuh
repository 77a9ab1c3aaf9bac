// gsc_online.cuh -- K4/K5: the reference's online k-means (enc:699-765
// TFrame.KNNScanReduce), one point at a time against the LIVE centroids.
//
// The rule is sequential in the points (each update moves one centroid before
// the next query), so the parallelism is (a) across the K centroids inside a
// frame and (b) across the frames of a batch: ONE CTA PER FRAME, the whole
// codebook resident in REGISTERS (K*D floats = 128 KB at K = 4096, D = 8:
// thread t of 512 owns centroids t*CPT .. t*CPT+CPT-1), so a query reads the
// codebook at register-file bandwidth and never touches shared memory or HBM
// for it.  HBM traffic is one 32-byte row per point per pass.
//
// Per point (arithmetic of the result is the reference's, bit for bit):
//   best  := argmin_c  d(x, c),  d = sum_k (x_k - c_k)^2   float, left to
//            right, no FMA (ANN annkSearch, eps = 0; lowest index on ties)
//   rate  := Single(1 / sqrt(cnt_prev[best]))               enc:735
//   c_best += (x - c_best) * rate                            enc:736-740
//   err   += sqrt(d / D)   (Single sqrt, Double accumulate)  enc:743
//
// Finding the argmin without scoring all K centroids exactly.  Each thread
// evaluates for its centroids, over the first DF = D/2 dimensions only,
//   s_c = x.c - 0.5*|c|^2*(1-g)              (DF FFMAs, chain starts at h_c)
// which certifies the LOWER bound
//   lb_c = |x|^2(1-g) - 2 s_c  <=  sum_{k<DF}(x_k-c_k)^2  <=  d(x,c)
// (the dropped squared terms are >= 0 -- and tiny: they are the 1e-5-scaled
// cepstral features, enc:362 -- and g = 2^-17 covers every rounding of both
// forms).  Given a bound U, only centroids with lb_c <= U are
// scored in the exact operation order; each warp merges its (d bits, index)
// keys with two 32-bit REDUX and publishes one key; the minimum key IS
// "smallest d, lowest index on ties".
// U is the exact distance to the centroid the point chose in the previous
// pass (its seed cell in pass 0), computed one point ahead by that centroid's
// owner.  If that centroid is the one the previous point just moved, U is
// stale; instead of a barrier the result is VERIFIED: the owner of the moved
// centroid always contributes its exact distance, and the winner is accepted
// iff d_win <= U (then every centroid with d <= d_win had lb <= U and was
// scored).  Otherwise the point is redone exhaustively.
// Synchronisation: ONE split-phase mbarrier per point (arrive after
// publishing, bookkeeping overlaps the wait).
#pragma once
#include "gsc_device.cuh"

#define GSC_ON_T 512          // threads per CTA
#define GSC_ON_W (GSC_ON_T / 32)
#define GSC_ON_TP 256         // points per shared-memory tile
#define GSC_ON_G 7.62939453125e-06f   // 2^-17
#define GSC_ON_RLUT 2048      // rate LUT entries
#define GSC_NONE 0xffffffffu

template <int D>
struct GscOnlineSmem {
    float x[GSC_ON_TP][D];
    float hx[GSC_ON_TP];           // 0.5*|x[0..DF)|^2*(1-g) - tiny
    int g[GSC_ON_TP];              // guess = previous label (sanitised)
    float rate[GSC_ON_RLUT];       // Single(1/sqrt(cnt)), enc:735
    float et[GSC_ON_TP];           // per-point sqrt(d/D) terms, summed in point order by thread 0
    unsigned long long wkey[2][GSC_ON_W];   // per-warp (d bits << 32 | index), double buffered per point
    unsigned long long xkey[GSC_ON_W];      // exhaustive redo
    unsigned long long mbar;       // split-phase barrier, 16 arrivals (lane 0 of each warp)
    float U[2];
    double err;
    int stop;
};

__device__ __forceinline__ unsigned gsc_smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void gsc_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gsc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gsc_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gsc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gsc_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "GSC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GSC_DONE_%=;\n\t"
        "bra GSC_WAIT_%=;\n\t"
        "GSC_DONE_%=:\n\t"
        "}" ::"r"(gsc_smem_u32(bar)), "r"(parity)
        : "memory");
}

__device__ __forceinline__ unsigned long long gsc_pack(unsigned dbits, unsigned idx) {
    return ((unsigned long long)dbits << 32) | idx;
}
__device__ __forceinline__ float gsc_rate(int cnt) {  // enc:735
    return (float)(1.0 / sqrt((double)cnt));
}

// (d bits, idx) minimum across the warp; NONE/NONE if no lane has a key.
__device__ __forceinline__ void gsc_warp_min(unsigned &dbits, unsigned &idx) {
    const unsigned m = __reduce_min_sync(0xffffffffu, dbits);
    const unsigned i = __reduce_min_sync(0xffffffffu, dbits == m ? idx : GSC_NONE);
    dbits = m; idx = i;
}

// exact scoring of every own centroid (validation mode and redo path)
template <int D, int CPT>
__device__ __forceinline__ void gsc_local_exact(const float (&c)[CPT][D], const float (&x)[D], int first,
                                                unsigned &dbits, unsigned &idx) {
    dbits = GSC_NONE; idx = GSC_NONE;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const float d = gsc_ann_dist<D>(x, c[j]);
        const unsigned b = __float_as_uint(d);
        if (d == d && b < dbits) { dbits = b; idx = (unsigned)(first + j); }   // strict <: lowest index wins
    }
}

// Exact distance to the one own centroid selected by `mask` (= 1 << slot).
// A bit mask, not `slot == j`: nvcc turns an equality chain over j into a
// dynamically indexed local-memory copy of the whole codebook.
template <int D, int CPT>
__device__ __forceinline__ float gsc_owner_dist(const float (&c)[CPT][D], const float (&x)[D], unsigned mask) {
    float d = INFINITY;
#pragma unroll
    for (int j = 0; j < CPT; ++j)
        if (mask & (1u << j)) d = gsc_ann_dist<D>(x, c[j]);
    return (d == d) ? d : INFINITY;
}

template <int D, int CPT>
__global__ void __launch_bounds__(GSC_ON_T, 1) k_online(const GscFrame *__restrict__ frames,
                                                        const float *__restrict__ X,       // [sumN][D]
                                                        float *__restrict__ cen,           // [F][Kmax][D] in/out
                                                        int *__restrict__ labels,          // [sumN] in: guesses, out: labels
                                                        int *__restrict__ passes_out,      // [F]
                                                        double *__restrict__ err_out,      // [F]
                                                        double tol, int max_passes, int Kmax, int force_exact) {
    constexpr int DF = (D >= 8) ? D / 2 : D;   // filter dimensions
    extern __shared__ __align__(16) unsigned char smraw[];
    GscOnlineSmem<D> &sm = *reinterpret_cast<GscOnlineSmem<D> *>(smraw);
    int *cnts = reinterpret_cast<int *>(smraw + sizeof(GscOnlineSmem<D>));  // [2][T*CPT]
    constexpr int KP = GSC_ON_T * CPT;

    const GscFrame f = frames[blockIdx.x];
    const int K = f.K, N = f.N;
    if (K <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int first = tid * CPT;
    const float *Xf = X + f.chunk_off * D;
    int *lab = labels + f.chunk_off;
    float *cf = cen + (long long)f.slot * Kmax * D;

    // codebook -> registers; dead slots (idx >= K) are NaN and never win
    float c[CPT][D], h[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int idx = first + j;
        float nc = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            c[j][k] = (idx < K) ? cf[(long long)idx * D + k] : __int_as_float(0x7fc00000);
            if (k < DF) nc = fmaf(c[j][k], c[j][k], nc);
        }
        h[j] = -0.5f * nc * (1.0f - GSC_ON_G);
    }
    for (int j = tid; j < 2 * KP; j += GSC_ON_T) cnts[j] = 1;  // enc:717-721
    for (int j = tid; j < GSC_ON_RLUT; j += GSC_ON_T) sm.rate[j] = gsc_rate(j < 1 ? 1 : j);
    if (tid == 0) {
        sm.err = 3.40282346638528860e+38; sm.stop = 0;
        gsc_mbar_init(&sm.mbar, GSC_ON_W);
    }
    __syncthreads();

    unsigned phase = 0;   // parity of the mbarrier phase to wait for next (uniform)
    int iter = 0;
    double prevErr;
    for (;;) {
        const int odd = iter & 1;
        int *cnt_prev = cnts + (odd ? 0 : KP);   // cnts[not Odd(iter)]
        int *cnt_cur = cnts + (odd ? KP : 0);    // cnts[Odd(iter)]
        prevErr = sm.err;                        // uniform copy
        __syncthreads();
        if (tid == 0) sm.err = 0.0;

        for (int base = 0; base < N; base += GSC_ON_TP) {
            const int tn = min(GSC_ON_TP, N - base);
            __syncthreads();  // (A) previous tile fully consumed
            for (int t = tid; t < tn * D; t += GSC_ON_T) (&sm.x[0][0])[t] = Xf[(long long)base * D + t];
            for (int t = tid; t < tn; t += GSC_ON_T) {
                int gg = lab[base + t];
                sm.g[t] = (gg < 0 || gg >= K) ? 0 : gg;
            }
            __syncthreads();  // (B)
            for (int t = tid; t < tn; t += GSC_ON_T) {
                float nx = 0.0f;
#pragma unroll
                for (int k = 0; k < DF; ++k) nx = fmaf(sm.x[t][k], sm.x[t][k], nx);
                sm.hx[t] = 0.5f * nx * (1.0f - GSC_ON_G) - 1e-30f;
            }
            {   // bound for the first point of the tile
                const int g0 = sm.g[0];
                if (g0 >= first && g0 < first + CPT) {
                    float x0[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) x0[k] = sm.x[0][k];
                    sm.U[0] = gsc_owner_dist<D, CPT>(c, x0, 1u << (g0 - first));
                }
            }
            __syncthreads();  // (C)

            float Uprev = INFINITY;
            // deferred bookkeeping of the point resolved in the previous interval (owner only)
            int bk_b = -1, bk_p = 0;
            float bk_d = 0.0f;
            for (int ii = 0; ii <= tn; ++ii) {
                unsigned fd = GSC_NONE, fi = GSC_NONE;   // this thread's key for point ii
                // ---- resolve point ii-1, apply its update (enc:733-744) ----
                if (ii > 0) {
                    const int p = ii - 1;
                    gsc_mbar_wait(&sm.mbar, phase);
                    phase ^= 1u;
                    unsigned long long wk = sm.wkey[p & 1][lane & (GSC_ON_W - 1)];
                    unsigned kd = (unsigned)(wk >> 32), ki = (unsigned)(wk & 0xffffffffu);
                    gsc_warp_min(kd, ki);
                    float xp[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) xp[k] = sm.x[p][k];
                    if (kd == GSC_NONE || !(__uint_as_float(kd) <= Uprev)) {
                        // stale / missing bound (block-uniform): redo exhaustively
                        unsigned ed, ei;
                        gsc_local_exact<D, CPT>(c, xp, first, ed, ei);
                        gsc_warp_min(ed, ei);
                        if (lane == 0) sm.xkey[warp] = gsc_pack(ed, ei);
                        __syncthreads();
                        wk = sm.xkey[lane & (GSC_ON_W - 1)];
                        kd = (unsigned)(wk >> 32); ki = (unsigned)(wk & 0xffffffffu);
                        gsc_warp_min(kd, ki);
                        __syncthreads();   // xkey may be rewritten by a later redo
                    }
                    const int b = (kd == GSC_NONE) ? 0 : (int)ki;
                    const float dbest = (kd == GSC_NONE) ? INFINITY : __uint_as_float(kd);
                    if (b >= first && b < first + CPT) {
                        // executed by the centroid's owner only
                        const int cp = cnt_prev[b];
                        const float rate = (cp < GSC_ON_RLUT) ? sm.rate[cp] : gsc_rate(cp);
                        const unsigned um = 1u << (b - first);
                        float xn[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) xn[k] = sm.x[ii < tn ? ii : p][k];
#pragma unroll
                        for (int j = 0; j < CPT; ++j)
                            if (um & (1u << j)) {
                                float nc = 0.0f;
#pragma unroll
                                for (int k = 0; k < D; ++k) {
                                    float v = xp[k] - c[j][k];
                                    float m = v * rate;
                                    c[j][k] = c[j][k] + m;
                                    if (k < DF) nc = fmaf(c[j][k], c[j][k], nc);
                                }
                                h[j] = -0.5f * nc * (1.0f - GSC_ON_G);
                                // the moved centroid always contributes its exact distance to the next point
                                const float dn = gsc_ann_dist<D>(xn, c[j]);
                                if (dn == dn) { fd = __float_as_uint(dn); fi = (unsigned)b; }
                            }
                        bk_b = b; bk_p = p; bk_d = dbest;
                    }
                }
                if (ii < tn) {
                    // ---- scoring of point ii ----
                    float x[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = sm.x[ii][k];
                    if (force_exact) {
                        gsc_local_exact<D, CPT>(c, x, first, fd, fi);
                        Uprev = INFINITY;
                    } else {
                        const float U = sm.U[ii & 1];
                        Uprev = U;
                        const float thr = sm.hx[ii] - 0.5f * U;  // candidate iff s >= thr  (lb <= U)
                        float s[CPT];
#pragma unroll
                        for (int j = 0; j < CPT; ++j) s[j] = h[j];
#pragma unroll
                        for (int k = 0; k < DF; ++k)
#pragma unroll
                            for (int j = 0; j < CPT; ++j) s[j] = fmaf(x[k], c[j][k], s[j]);
                        bool any = false;
#pragma unroll
                        for (int j = 0; j < CPT; ++j) any |= (s[j] >= thr);
                        if (any) {
#pragma unroll
                            for (int j = 0; j < CPT; ++j)
                                if (s[j] >= thr) {
                                    const float d = gsc_ann_dist<D>(x, c[j]);
                                    const unsigned b = __float_as_uint(d);
                                    const unsigned id = (unsigned)(first + j);
                                    if (d == d && (b < fd || (b == fd && id < fi))) { fd = b; fi = id; }
                                }
                        }
                    }
                    // ---- publish this warp's key ----
                    if (__any_sync(0xffffffffu, fd != GSC_NONE)) gsc_warp_min(fd, fi);
                    if (lane == 0) sm.wkey[ii & 1][warp] = gsc_pack(fd, fi);
                    // ---- bound for point ii+1, one point ahead ----
                    if (ii + 1 < tn) {
                        const int g1 = sm.g[ii + 1];
                        if (g1 >= first && g1 < first + CPT) {
                            float x1[D];
#pragma unroll
                            for (int k = 0; k < D; ++k) x1[k] = sm.x[ii + 1][k];
                            sm.U[(ii + 1) & 1] = gsc_owner_dist<D, CPT>(c, x1, 1u << (g1 - first));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) gsc_mbar_arrive(&sm.mbar);
                }
                // ---- bookkeeping of the resolved point, overlapped with the wait ----
                if (bk_b >= 0) {
                    lab[base + bk_p] = bk_b;                                  // enc:742
                    sm.et[bk_p] = sqrtf(bk_d / (float)D);                     // enc:743 (term)
                    cnt_cur[bk_b] += 1;                                       // enc:744
                    bk_b = -1;
                }
            }
            __syncthreads();  // (E) all terms of the tile written
            if (tid == 0) {
                double e = sm.err;                                            // enc:743 (Double sum, point order)
                for (int p = 0; p < tn; ++p) e += (double)sm.et[p];
                sm.err = e;
            }
        }
        // ---- end of pass: enc:754-761 ----
        __syncthreads();
        for (int j = tid; j < KP; j += GSC_ON_T) cnt_prev[j] = 1;
        ++iter;
        if (tid == 0) {
            const double e = sm.err;
            const bool same = (e > prevErr) ? ((e - prevErr) <= tol) : ((prevErr - e) <= tol);
            sm.stop = (same || iter >= max_passes) ? 1 : 0;
        }
        __syncthreads();
        if (sm.stop) break;
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int idx = first + j;
        if (idx < K) {
#pragma unroll
            for (int k = 0; k < D; ++k) cf[(long long)idx * D + k] = c[j][k];
        }
    }
    if (tid == 0) {
        passes_out[f.slot] = iter;
        err_out[f.slot] = sm.err;
    }
}
