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
// Finding the argmin without evaluating all K exact distances: each thread
// evaluates for its centroids the FFMA form
//   s_c = x.c - 0.5*|c|^2*(1-g)          (D FFMAs, the chain starts at h_c)
// which gives a certified LOWER bound  lb_c = |x|^2(1-g) - 2 s_c <= d(x,c)
// (g = 2^-17 covers every rounding of both forms).  Given any
// upper bound U >= min_c d(x,c), only centroids with lb_c <= U can be the
// argmin; those few are re-scored in the exact operation order and the
// (d, index) minimum over them is the exact result.  U is the exact distance
// to the centroid this point chose in the previous pass (or its seed cell in
// pass 0), computed by that centroid's owner one point ahead, so the common
// case costs one block barrier per point.
#pragma once
#include "gsc_device.cuh"

#define GSC_ON_T 512          // threads per CTA
#define GSC_ON_TP 256         // points per shared-memory tile
#define GSC_ON_CAP 48         // candidate list capacity per point
#define GSC_ON_G 7.62939453125e-06f   // 2^-17

template <int D>
struct GscOnlineSmem {
    float x[GSC_ON_TP][D];
    float hx[GSC_ON_TP];           // 0.5*|x|^2*(1-g) - tiny
    int g[GSC_ON_TP];              // guess = previous label (sanitised)
    unsigned long long cand[3][GSC_ON_CAP];
    int cand_n[3];
    float U[2];
    unsigned long long wkey[GSC_ON_T / 32];
    double err;
    int stop;
};

__device__ __forceinline__ unsigned long long gsc_pack(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}

// exact local best over ALL own centroids (slow path / validation path)
template <int D, int CPT>
__device__ __forceinline__ unsigned long long gsc_local_exact(const float (&c)[CPT][D], const float (&x)[D], int first) {
    unsigned long long key = ~0ull;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        float d = gsc_ann_dist<D>(x, c[j]);
        if (d == d) {
            unsigned long long k = gsc_pack(d, first + j);
            key = k < key ? k : key;
        }
    }
    return key;
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

// Block-wide exact argmin (one extra barrier). All threads return the same key.
template <int D, int CPT>
__device__ __forceinline__ unsigned long long gsc_block_exact(GscOnlineSmem<D> &sm, const float (&c)[CPT][D],
                                                              const float (&x)[D], int first) {
    unsigned long long key = gsc_local_exact<D, CPT>(c, x, first);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) sm.wkey[threadIdx.x >> 5] = key;
    __syncthreads();
    unsigned long long best = ~0ull;
#pragma unroll
    for (int w = 0; w < GSC_ON_T / 32; ++w) {
        unsigned long long k = sm.wkey[w];
        best = k < best ? k : best;
    }
    return best;
}

template <int D, int CPT>
__global__ void __launch_bounds__(GSC_ON_T, 1) k_online(const GscFrame *__restrict__ frames,
                                                        const float *__restrict__ X,       // [sumN][D]
                                                        float *__restrict__ cen,           // [F][Kmax][D] in/out
                                                        int *__restrict__ labels,          // [sumN] in: guesses, out: labels
                                                        int *__restrict__ passes_out,      // [F]
                                                        double *__restrict__ err_out,      // [F]
                                                        double tol, int max_passes, int Kmax, int force_exact) {
    extern __shared__ __align__(16) unsigned char smraw[];
    GscOnlineSmem<D> &sm = *reinterpret_cast<GscOnlineSmem<D> *>(smraw);
    int *cnts = reinterpret_cast<int *>(smraw + sizeof(GscOnlineSmem<D>));  // [2][T*CPT]
    constexpr int KP = GSC_ON_T * CPT;

    const GscFrame f = frames[blockIdx.x];
    const int K = f.K, N = f.N;
    if (K <= 0) return;
    const int tid = threadIdx.x;
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
            nc = fmaf(c[j][k], c[j][k], nc);
        }
        h[j] = -0.5f * nc * (1.0f - GSC_ON_G);
    }
    for (int j = tid; j < 2 * KP; j += GSC_ON_T) cnts[j] = 1;  // enc:717-721
    if (tid == 0) { sm.err = 3.40282346638528860e+38; sm.stop = 0; }
    __syncthreads();

    int iter = 0;
    double prevErr;
    for (;;) {
        const int odd = iter & 1;
        int *cnt_prev = cnts + (odd ? 0 : KP);   // cnts[not Odd(iter)]
        int *cnt_cur = cnts + (odd ? KP : 0);    // cnts[Odd(iter)]
        prevErr = sm.err;                        // every thread keeps a copy (uniform)
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
            if (tid < 3) sm.cand_n[tid] = 0;
            __syncthreads();  // (B)
            for (int t = tid; t < tn; t += GSC_ON_T) {
                float nx = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) nx = fmaf(sm.x[t][k], sm.x[t][k], nx);
                sm.hx[t] = 0.5f * nx * (1.0f - GSC_ON_G) - 1e-30f;
            }
            // prologue: U for the first point of the tile
            {
                const int g0 = sm.g[0];
                if (g0 >= first && g0 < first + CPT) {
                    float x0[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) x0[k] = sm.x[0][k];
                    sm.U[0] = gsc_owner_dist<D, CPT>(c, x0, 1u << (g0 - first));
                }
            }
            __syncthreads();  // (C)

            for (int ii = 0; ii <= tn; ++ii) {
                // ---- resolve point ii-1, apply its update ----
                int bprev = -1;
                if (ii > 0) {
                    const int p = ii - 1;
                    float xp[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) xp[k] = sm.x[p][k];
                    unsigned long long key;
                    const int n = force_exact ? (GSC_ON_CAP + 1) : sm.cand_n[p % 3];
                    if (n <= GSC_ON_CAP) {
                        key = ~0ull;
                        for (int e = 0; e < n; ++e) {
                            unsigned long long k = sm.cand[p % 3][e];
                            key = k < key ? k : key;
                        }
                    } else {
                        key = gsc_block_exact<D, CPT>(sm, c, xp, first);  // + one barrier (block-uniform)
                    }
                    bprev = (key == ~0ull) ? 0 : (int)(unsigned)(key & 0xffffffffu);
                    const float dbest = (key == ~0ull) ? INFINITY : __uint_as_float((unsigned)(key >> 32));
                    if (bprev >= first && bprev < first + CPT) {
                        // enc:735-744, executed by the centroid's owner
                        const float rate = (float)(1.0 / sqrt((double)cnt_prev[bprev]));
                        const unsigned um = 1u << (bprev - first);
#pragma unroll
                        for (int j = 0; j < CPT; ++j)
                            if (um & (1u << j)) {
                                float nc = 0.0f;
#pragma unroll
                                for (int k = 0; k < D; ++k) {
                                    float v = xp[k] - c[j][k];
                                    float m = v * rate;
                                    c[j][k] = c[j][k] + m;
                                    nc = fmaf(c[j][k], c[j][k], nc);
                                }
                                h[j] = -0.5f * nc * (1.0f - GSC_ON_G);
                            }
                        lab[base + p] = bprev;                                    // enc:742
                        sm.err += (double)sqrtf(dbest / (float)D);                // enc:743
                        cnt_cur[bprev] += 1;                                      // enc:744
                    }
                }
                if (ii == tn) break;
                // ---- conflict: the guess of point ii is the centroid that just moved ----
                const int gi = sm.g[ii];
                if (ii > 0 && gi == bprev) {
                    if (gi >= first && gi < first + CPT) {
                        float xi[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) xi[k] = sm.x[ii][k];
                        sm.U[ii & 1] = gsc_owner_dist<D, CPT>(c, xi, 1u << (gi - first));
                    }
                    __syncthreads();  // block-uniform condition
                }
                // ---- local phase of point ii ----
                float x[D];
#pragma unroll
                for (int k = 0; k < D; ++k) x[k] = sm.x[ii][k];
                if (!force_exact) {
                    const float U = sm.U[ii & 1];
                    const float thr = sm.hx[ii] - 0.5f * U;  // candidate iff s >= thr
                    float s[CPT];
                    bool any = false;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) s[j] = h[j];
#pragma unroll
                    for (int k = 0; k < D; ++k)
#pragma unroll
                        for (int j = 0; j < CPT; ++j) s[j] = fmaf(x[k], c[j][k], s[j]);
#pragma unroll
                    for (int j = 0; j < CPT; ++j) any |= (s[j] >= thr);
                    if (any) {
                        unsigned long long key = ~0ull;
#pragma unroll
                        for (int j = 0; j < CPT; ++j)
                            if (s[j] >= thr) {
                                float d = gsc_ann_dist<D>(x, c[j]);
                                if (d == d) {
                                    unsigned long long k = gsc_pack(d, first + j);
                                    key = k < key ? k : key;
                                }
                            }
                        if (key != ~0ull) {
                            const int slot = atomicAdd(&sm.cand_n[ii % 3], 1);
                            if (slot < GSC_ON_CAP) sm.cand[ii % 3][slot] = key;
                        }
                    }
                }
                if (tid == 0) sm.cand_n[(ii + 1) % 3] = 0;
                // ---- bound for point ii+1 (one point ahead) ----
                if (ii + 1 < tn) {
                    const int g1 = sm.g[ii + 1];
                    if (g1 >= first && g1 < first + CPT) {
                        float x1[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) x1[k] = sm.x[ii + 1][k];
                        sm.U[(ii + 1) & 1] = gsc_owner_dist<D, CPT>(c, x1, 1u << (g1 - first));
                    }
                }
                __syncthreads();  // B_ii
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
